set -x
timeout 600 python -m pytest tests/test_remesh_gpu.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; tail -4 gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-strong > gpurun_out/r2i_bench_refill.json 2>gpurun_out/r2i_bench_refill.err; python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench_refill.json').read().strip().splitlines()[-1]);print('refill',d['ms_per_step'],d['submetrics']['phase_ms_per_step'],d['submetrics']['checks'])"; tail -2 gpurun_out/r2i_bench_refill.err
LV_CLIP_MODE=tile timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-strong > gpurun_out/r2i_bench_tile.json 2>gpurun_out/r2i_bench_tile.err; python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench_tile.json').read().strip().splitlines()[-1]);print('tile',d['ms_per_step'],d['submetrics']['phase_ms_per_step'])"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests_all.log 2>&1; tail -4 gpurun_out/r2i_tests_all.log
NCU="ncu --set full --import-source on --clock-control none"
timeout 400 $NCU -k regex:'k_clip_refill|k_clip_emit' -s 2 -c 2 -o gpurun_out/r2i_clip -f python tools/prof_one.py 4096 1 > gpurun_out/r2i_ncu.log 2>&1; tail -2 gpurun_out/r2i_ncu.log
