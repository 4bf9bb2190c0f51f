mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3k_tests.log 2>&1; tail -15 gpurun_out/r3k_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
