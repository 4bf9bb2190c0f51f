mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_tests.log 2>&1; tail -4 gpurun_out/r3h_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-strong > gpurun_out/r3h_n2.json 2> gpurun_out/r3h_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/r3h_n2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['e2e'])"; tail -3 gpurun_out/r3h_n2.err
