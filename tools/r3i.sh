mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3i_bench_n1.json 2> gpurun_out/r3i_bench_n1.err; tail -c 600 gpurun_out/r3i_bench_n1.json; tail -3 gpurun_out/r3i_bench_n1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3i_ref_n1.json 2> gpurun_out/r3i_ref_n1.err; cat gpurun_out/r3i_ref_n1.json | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-strong --no-cpu --sweep > gpurun_out/r3i_sweep.json 2> gpurun_out/r3i_sweep.err; tail -c 300 gpurun_out/r3i_sweep.json
