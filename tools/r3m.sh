mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pressure_gpu.py tests/test_stepping_gpu.py -m gpu -x -q > gpurun_out/r3m_tests.log 2>&1; tail -4 gpurun_out/r3m_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-shuffle > gpurun_out/r3m_bench.json 2>gpurun_out/r3m_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r3m_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['config']['krylov_iters_per_step'],d['submetrics']['phase_ms_per_step'],d['e2e']['ms_per_step'],d['submetrics']['plain_cg']['ms_per_step'])"; tail -3 gpurun_out/r3m_bench.err
