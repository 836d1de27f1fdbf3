"""Soak of the point-sharded BA (one NCCL all-reduce of the reduced camera system per LM attempt) over
every solver path and any world size, against the fp64 oracle on rank 0:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/ba_shard_soak.py [rounds]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lorb_slam_b200 import capi, synth  # noqa: E402
from oracle import ref  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = capi.Context(local)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(capi.Context.dist_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
ctx.dist_init(bytes(uid.cpu().tolist()), rank, world)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(os.environ.get("LORB_SOAK_SEED", 5)))  # same stream on every rank
bad = 0
for s in range(n):
    C = int(rng.choice([3, 5, 8, 10, 11, 14, 16, 17, 24, 40, 60]))
    P = int(rng.integers(60, 1500))
    k = tuple(int(v) for v in rng.choice(np.arange(3, min(C, 10) + 1), size=3))
    ff = float(rng.choice([0.0, 0.0, 0.2]))
    it = int(rng.choice([3, 6, 20]))
    pb = synth.make_ba_problem(7000 + s, C=C, P=P, obs_per_point=k, fixed_frac=ff, traj_len=float(max(3.0, C * 0.4)))
    opt = capi.ba_options(max_num_iterations=it)
    prob = ctx.ba_problem_sharded(pb, rank, world)
    sm = prob.solve(opt, sharded=True)
    cams, pts = prob.download()
    ids = prob.point_ids
    prob.close()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(ids)], dtype=torch.int64, device="cuda"))
    mx = max(int(v) for v in sizes)
    buf = torch.zeros((mx, 4), dtype=torch.float64, device="cuda")
    buf[:len(ids), :3] = torch.from_numpy(pts).cuda()
    buf[:len(ids), 3] = torch.from_numpy(ids.astype(np.float64)).cuda()
    allb = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf)
    if rank == 0:
        full = np.zeros((P, 3))
        for r_ in range(world):
            a = allb[r_][:int(sizes[r_])].cpu().numpy()
            full[a[:, 3].astype(np.int64)] = a[:, :3]
        oc, op, so = ref.ba_local(pb, ref.ba_options(max_num_iterations=it))
        ok = (np.allclose(cams, oc, rtol=1e-6, atol=1e-8) and np.allclose(full, op, rtol=1e-6, atol=1e-8)
              and sm["iterations"] == so["iterations"] and sm["termination"] == so["termination"])
        if not ok:
            bad += 1
            print("SHARD MISMATCH", s, C, P, k, ff, it, np.abs(cams - oc).max(), np.abs(full - op).max(),
                  sm["iterations"], so["iterations"], flush=True)
dist.barrier()
ctx.dist_finalize()
ctx.close()
if rank == 0:
    print("%d sharded solves at world size %d, %d differ" % (n, world, bad))
dist.destroy_process_group()
