set -x
LV_DEBUG_STEP=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py > gpurun_out/r2f_multigpu.log 2>&1; grep -n "case \|stepping\|rror\|non-finite" gpurun_out/r2f_multigpu.log | head -20
timeout 600 python -m pytest tests/test_pressure_gpu.py -m gpu -x -q -s 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err; tail -c 1500 gpurun_out/r2f_bench2.json; tail -5 gpurun_out/r2f_bench2.err
