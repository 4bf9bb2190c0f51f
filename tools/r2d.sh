set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; tail -5 gpurun_out/r2d_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-strong > gpurun_out/r2d_bench6.json 2>gpurun_out/r2d_bench6.err; tail -c 1500 gpurun_out/r2d_bench6.json
LV_MV_MINB=5 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-strong > gpurun_out/r2d_bench5.json 2>gpurun_out/r2d_bench5.err; tail -c 1500 gpurun_out/r2d_bench5.json
LV_MV_MINB=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-strong > gpurun_out/r2d_bench1.json 2>gpurun_out/r2d_bench1.err; tail -c 1500 gpurun_out/r2d_bench1.json
NCU="ncu --set full --import-source on --clock-control none"
timeout 400 $NCU -k regex:k_matvec -s 4 -c 1 -o gpurun_out/r2d_mv -f python tools/prof_one.py 4096 2 > gpurun_out/r2d_ncu.log 2>&1
timeout 400 $NCU -k regex:'k_rhs_gp|k_rhs_corr|k_assemble' -c 3 -o gpurun_out/r2d_rhs -f python tools/prof_one.py 4096 2 >> gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
