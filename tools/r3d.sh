mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1; tail -4 gpurun_out/r3d_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3d_bench_n1.json 2> gpurun_out/r3d_bench_n1.err; tail -c 1500 gpurun_out/r3d_bench_n1.json; tail -3 gpurun_out/r3d_bench_n1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3d_ref_n1.json 2> gpurun_out/r3d_ref_n1.err; cat gpurun_out/r3d_ref_n1.json | cut -c1-700
