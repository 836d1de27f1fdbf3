set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python -m pytest tests/test_host_dropin_gpu.py -x -q -m gpu -k local_mapping -s 2>&1 | grep -E "assemble|passed|failed"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().split('\n')[-1])
print('sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['roofline']['frac'])
print('full', d['full_sweep']['value']/1e9, d['full_sweep']['wall_s'], d['full_sweep']['gather_s'], d['full_sweep']['parity'])
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['kernel'], b['roofline']['frac'], b['roofline']['ms_per_attempt_by_part'], b['parity'])
PY
