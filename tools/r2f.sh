set -x
LV_DEBUG_STEP=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multigpu_check.py > gpurun_out/r2f_multigpu.log 2>&1; grep -n "case \|stepping\|rror\|non-finite" gpurun_out/r2f_multigpu.log | head -20
