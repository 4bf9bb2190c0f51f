set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; tail -4 gpurun_out/r2h_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -c 1200 gpurun_out/r2h_bench_n1.json; tail -3 gpurun_out/r2h_bench_n1.err
NCU="ncu --set full --import-source on --clock-control none"
timeout 400 $NCU -k regex:k_matvec -s 4 -c 1 -o gpurun_out/r2h_mv -f python tools/prof_one.py 4096 2 > gpurun_out/r2h_ncu.log 2>&1
timeout 400 $NCU -k regex:k_clip_fast -s 1 -c 1 -o gpurun_out/r2h_clip -f python tools/prof_one.py 4096 1 >> gpurun_out/r2h_ncu.log 2>&1
timeout 400 $NCU -k regex:k_cg_update -s 4 -c 2 -o gpurun_out/r2h_upd -f python tools/prof_one.py 4096 2 >> gpurun_out/r2h_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-strong --no-cpu --side 2048 > gpurun_out/r2h_launches.out 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-strong --no-cpu --sweep > gpurun_out/r2h_sweep.json 2> gpurun_out/r2h_sweep.err; tail -c 600 gpurun_out/r2h_sweep.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2h_ref_n1.json 2> gpurun_out/r2h_ref_n1.err; cat gpurun_out/r2h_ref_n1.json | cut -c1-600
