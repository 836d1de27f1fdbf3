set -x
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 profiles/scripts/ba_shard_soak.py 60 > gpurun_out/soak_shard2.log 2>&1; echo rc=$?; tail -6 gpurun_out/soak_shard2.log | cut -c1-300
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29752 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().split('\n')[-1])
print('N2 sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9)
print('full', d['full_sweep']['value']/1e9, d['full_sweep']['wall_s'], d['full_sweep']['parity']['ok'])
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['ms_per_attempt_by_part'], b['parity']['ok'])
PY
