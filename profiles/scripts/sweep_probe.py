"""Times the keyframe-pair sweep kernels (popc / tensor) on one block of the BASELINE cfg-5 bank.

    python profiles/scripts/sweep_probe.py [n_kf_block] [reps] [impl,impl,...]
Prints ms per step and G descriptor-pairs/s per implementation, and checks the two agree."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lorb_slam_b200 import capi  # noqa: E402

blk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
impls = sys.argv[3].split(",") if len(sys.argv) > 3 else ["popc", "tensor"]
n_desc = int(os.environ.get("N_DESC", "2000"))
rng = np.random.default_rng(0)
bank = rng.integers(0, 256, size=(2 * blk, n_desc, 32), dtype=np.uint8)
a, b = np.triu_indices(blk, k=1)
a, b = a.astype(np.int32), b.astype(np.int32)
res = {}
with capi.Context(0) as ctx:
    st = torch.cuda.ExternalStream(ctx.stream)
    for impl in impls:
        ctx.sweep_set_impl(impl)
        ctx.bank_upload(bank)
        ctx.sweep_plan_upload(a, b)
        for i in range(2):
            ctx.sweep_plan_run((i % 2) * blk)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(reps):
            ctx.sweep_plan_run((i % 2) * blk)
        e1.record(st)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        res[impl] = ctx.sweep_plan_download()
        pairs = len(a) * n_desc * n_desc
        print("%s: %.3f ms/step  %.1f G pairs/s  (%d kf pairs x %d^2)" % (impl, ms, pairs / ms / 1e6, len(a), n_desc), flush=True)
if len(res) == 2:
    r0, r1 = res[impls[0]], res[impls[1]]
    same = all(np.array_equal(x, y) for x, y in zip(r0, r1))
    print("results identical:", same)
    if not same:
        bad = np.nonzero((r0[0] != r1[0]) | (r0[1] != r1[1]) | (r0[2] != r1[2]))[0]
        print("first mismatches", bad[:10], [(r0[0][k], r0[1][k], r0[2][k], r1[0][k], r1[1][k], r1[2][k]) for k in bad[:10]])
        sys.exit(1)
