# bench lines under torchrun on N GPUs of one box: bash profiles/scripts/bench_multi.sh N
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/bench_n${N}_$2.json 2> gpurun_out/bench_n${N}_$2.err; }
run 29611 sweep --steps 5 --warmup 3
run 29612 reference --impl reference --steps 2 --warmup 1
run 29613 ba_batched --workload ba_batched --steps 3 --warmup 3
run 29614 ba_large --workload ba_large --steps 3 --warmup 3
python - <<PY
import json
for f in ("sweep", "reference", "ba_batched", "ba_large"):
    try:
        d = json.loads(open("gpurun_out/bench_n${N}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "%.4g" % d["value"], d["unit"], "e2e %.4g" % d["e2e"]["value"], d.get("scaling"))
    except Exception as e:
        print(f, "FAILED", e)
PY
