mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py > gpurun_out/r3g_multigpu_check_2gpu.log 2>&1; tail -5 gpurun_out/r3g_multigpu_check_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3g_n2.json 2> gpurun_out/r3g_n2.err; tail -c 1800 gpurun_out/r3g_n2.json; tail -3 gpurun_out/r3g_n2.err
