mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r3j_n8.json 2> gpurun_out/r3j_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/r3j_n8.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']);s=d['submetrics']['strong_64M'];print('strong',s['ms_per_step'],s['checks']['mesh_witness_matches_1gpu'],s.get('e2e'))"; tail -3 gpurun_out/r3j_n8.err
