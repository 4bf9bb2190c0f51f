set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_remesh_gpu.py -m gpu -x -q > gpurun_out/r3b_tests.log 2>&1; tail -5 gpurun_out/r3b_tests.log
timeout 300 python tools/e2e_trace.py 4096 > gpurun_out/r3b_trace.log 2>&1; grep -v "^\[pipe\] remesh\|Warning" gpurun_out/r3b_trace.log | tail -40
LV_HOST_THREADS=8 timeout 300 python tools/e2e_trace.py 4096 > gpurun_out/r3b_trace8.log 2>&1; grep "^step\|job" gpurun_out/r3b_trace8.log | tail -6
