set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; tail -5 gpurun_out/r2e_tests.log
timeout 900 python tools/configs_at_scale.py > gpurun_out/r2_configs_at_scale.json 2> gpurun_out/r2_configs_at_scale.err; tail -c 3000 gpurun_out/r2_configs_at_scale.json; tail -5 gpurun_out/r2_configs_at_scale.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1800 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
bash tools/sanitize.sh gpurun_out
