mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_clip_fast -s 1 -c 1 -o gpurun_out/r3o_clip -f python tools/prof_one.py 4096 1 > gpurun_out/r3o_ncu.log 2>&1; tail -1 gpurun_out/r3o_ncu.log
