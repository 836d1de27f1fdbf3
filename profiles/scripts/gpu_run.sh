set -x
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2975$N bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
print('N$N sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9)
print('full', d['full_sweep']['value']/1e9, d['full_sweep']['wall_s'], d['full_sweep']['gather_s'], d['full_sweep']['parity']['ok'])
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['ms_per_attempt_by_part'], b['parity']['ok'], b['parity']['max_err_over_tolerance'])
PY
