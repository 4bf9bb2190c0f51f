mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3n_tests.log 2>&1; tail -3 gpurun_out/r3n_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3n_n2.json 2> gpurun_out/r3n_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/r3n_n2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['e2e']['mode']);s=d['submetrics']['strong_64M'];print('strong',s['ms_per_step'],s['checks']['mesh_witness_matches_1gpu'])"; tail -2 gpurun_out/r3n_n2.err
