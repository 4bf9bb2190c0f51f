mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
timeout 400 $NCU -k regex:k_clip_fast -s 1 -c 1 -o gpurun_out/r3f_clip -f python tools/prof_one.py 4096 1 > gpurun_out/r3f_ncu.log 2>&1
timeout 400 $NCU -k regex:'k_pipe_copy|k_pipe_deg|k_pipe_hdr' -c 3 -o gpurun_out/r3f_pipe -f python tools/e2e_trace.py 4096 >> gpurun_out/r3f_ncu.log 2>&1
tail -2 gpurun_out/r3f_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-strong --no-cpu --side 2048 > gpurun_out/r3f_launches.out 2>&1
bash tools/sanitize.sh gpurun_out > gpurun_out/r3f_sanitize.out 2>&1; tail -12 gpurun_out/r3f_sanitize.out
ls -la gpurun_out
