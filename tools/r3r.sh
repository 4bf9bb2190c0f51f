mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu --no-strong > gpurun_out/r3r_bench.json 2>gpurun_out/r3r_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r3r_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['mode'],d['submetrics']['shuffled_labels']['ms_per_step'])"; tail -3 gpurun_out/r3r_bench.err
LV_HOST_THREADS=1 timeout 600 python bench.py --no-cpu --no-strong --no-shuffle --e2e-mode all > gpurun_out/r3r_bench_all.json 2>gpurun_out/r3r_bench_all.err; python -c "
import json;d=json.loads(open('gpurun_out/r3r_bench_all.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['mode'])"; tail -3 gpurun_out/r3r_bench_all.err
