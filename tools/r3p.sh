mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pressure_gpu.py tests/test_stepping_gpu.py tests/test_io_populate.py -m gpu -x -q > gpurun_out/r3p_tests.log 2>&1; tail -4 gpurun_out/r3p_tests.log
LV_CLIP_MODE=plain timeout 900 python -m pytest tests/test_pressure_gpu.py -m gpu -x -q -k "operator or solve or walls" > gpurun_out/r3p_tests_plain.log 2>&1; tail -2 gpurun_out/r3p_tests_plain.log
for mode in tile rows; do
LV_ASSEMBLE=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-shuffle --no-e2e > gpurun_out/r3p_bench_$mode.json 2>gpurun_out/r3p_bench_$mode.err; python -c "
import json;d=json.loads(open('gpurun_out/r3p_bench_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['ms_per_step'],d['config']['krylov_iters_per_step'],d['submetrics']['phase_ms_per_step'])"; tail -2 gpurun_out/r3p_bench_$mode.err
done
