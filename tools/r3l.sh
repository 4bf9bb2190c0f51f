mkdir -p gpurun_out
timeout 1500 python tools/configs_at_scale.py > gpurun_out/r3l_configs_at_scale.json 2> gpurun_out/r3l_configs.err; tail -c 2500 gpurun_out/r3l_configs_at_scale.json; tail -3 gpurun_out/r3l_configs.err
