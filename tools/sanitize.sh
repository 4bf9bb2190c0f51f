#!/bin/bash
# compute-sanitizer over the small end-to-end case (single GPU) and, when two GPUs are present, memcheck over the multi-GPU
# parity check.  Logs go to gpurun_out/ (copy the summaries to profiles/).
set -x
CS=/usr/local/cuda/bin/compute-sanitizer
OUT=${1:-gpurun_out}
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_case.py ${SAN_M:-32} > $OUT/r2_sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> $OUT/r2_sanitize_$tool.log
  tail -4 $OUT/r2_sanitize_$tool.log
done
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 1200 $CS --tool memcheck --target-processes all --error-exitcode 9 --print-limit 20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > $OUT/r2_sanitize_multigpu_memcheck.log 2>&1
  echo "multigpu memcheck exit $?" >> $OUT/r2_sanitize_multigpu_memcheck.log
  tail -6 $OUT/r2_sanitize_multigpu_memcheck.log
fi
