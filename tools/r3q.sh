mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pressure_gpu.py -m gpu -x -q > gpurun_out/r3q_tests.log 2>&1; tail -6 gpurun_out/r3q_tests.log
