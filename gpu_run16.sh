for N in 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 5 --warmup 3 --no-extras 2>&1 | tail -1 > gpurun_out/bench_sweep_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/bench_sweep_n$N.json')); print('sweep N=$N', d['value']/1e9, 'G pairs/s  e2e', d['e2e']['value']/1e9, d['ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --workload ba_large --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_large_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/bench_ba_large_n$N.json')); print('ba_large N=$N', d['value']/1e6, 'M obs*it/s', d['ms_per_lm_iteration'], 'ms/iter e2e', d['e2e']['value']/1e6)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --workload ba_batched --windows 64 --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_batched_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/bench_ba_batched_n$N.json')); print('ba_batched N=$N', d['value']/1e6, 'M obs*it/s', d['ms_per_local_ba'], 'ms/window')"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 4 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-150
