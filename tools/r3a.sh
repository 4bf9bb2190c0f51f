set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; tail -15 gpurun_out/r3a_tests.log
LV_CLIP_STATS=1 timeout 300 python tools/prof_one.py 4096 1 > gpurun_out/r3a_stats.log 2>&1; grep "clip stats" gpurun_out/r3a_stats.log | head -3
for mode in pipeline all; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --e2e-mode $mode > gpurun_out/r3a_bench_$mode.json 2>gpurun_out/r3a_bench_$mode.err; python -c "
import json;d=json.loads(open('gpurun_out/r3a_bench_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['ms_per_step'],d['submetrics']['phase_ms_per_step'],d['e2e'])"; tail -3 gpurun_out/r3a_bench_$mode.err
done
