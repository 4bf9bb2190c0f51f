mkdir -p gpurun_out
LV_CLIP_MODE=refill LV_CLIP_STATS=1 timeout 300 python tools/prof_one.py 4096 1 2>&1 | grep "clip stats" | head -1
timeout 300 python tools/e2e_trace.py 4096 2>&1 | grep "^step\|job" | tail -4
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_clip_fast -s 1 -c 1 -o gpurun_out/r3c_clip -f python tools/prof_one.py 4096 1 > gpurun_out/r3c_ncu.log 2>&1; tail -2 gpurun_out/r3c_ncu.log; ls -la gpurun_out/
