mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_remesh_gpu.py tests/test_io_populate.py -m gpu -x -q > gpurun_out/r3e_tests.log 2>&1; tail -4 gpurun_out/r3e_tests.log
LV_CLIP_STATS=1 timeout 300 python tools/prof_one.py 4096 1 2>&1 | grep "clip stats" | head -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong > gpurun_out/r3e_bench.json 2>gpurun_out/r3e_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r3e_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['submetrics']['phase_ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['host_wall_ms_per_call'],d['submetrics']['checks'])"; tail -3 gpurun_out/r3e_bench.err
