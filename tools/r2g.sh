set -x
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py > gpurun_out/r2g_multigpu4.log 2>&1; grep -n "case \|stepping\|rror\|MULTIGPU" gpurun_out/r2g_multigpu4.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench8.json 2> gpurun_out/r2g_bench8.err; tail -c 3000 gpurun_out/r2g_bench8.json; tail -5 gpurun_out/r2g_bench8.err
