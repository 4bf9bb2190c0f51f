set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; tail -5 gpurun_out/r2c_tests.log
NCU="ncu --set full --import-source on --clock-control none"
timeout 300 $NCU -k regex:k_clip_fast -s 1 -c 1 -o gpurun_out/r2c_clip -f python tools/prof_one.py 2048 > gpurun_out/r2c_ncu.log 2>&1
timeout 300 $NCU -k regex:k_rhs_gp_mv -c 1 -o gpurun_out/r2c_gpmv -f python tools/prof_one.py 2048 >> gpurun_out/r2c_ncu.log 2>&1
timeout 300 $NCU -k regex:k_rhs_corr -c 1 -o gpurun_out/r2c_corr -f python tools/prof_one.py 2048 >> gpurun_out/r2c_ncu.log 2>&1
timeout 300 $NCU -k regex:k_assemble -c 1 -o gpurun_out/r2c_asm -f python tools/prof_one.py 2048 >> gpurun_out/r2c_ncu.log 2>&1
timeout 300 $NCU -k regex:k_matvec -s 4 -c 1 -o gpurun_out/r2c_mv -f python tools/prof_one.py 2048 >> gpurun_out/r2c_ncu.log 2>&1
timeout 300 $NCU -k regex:k_cg_update -s 4 -c 2 -o gpurun_out/r2c_upd -f python tools/prof_one.py 2048 >> gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 2500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
ls -la gpurun_out/
