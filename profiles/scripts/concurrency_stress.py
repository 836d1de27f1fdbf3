"""Two host threads, two contexts, one GPU: the matching entry points (tensor-core sweep, brute force,
projection search, ORB extraction) on one thread while bundle adjustments of every solver path
(dense, privatised, work lists + cooperative Cholesky) run on the other.  Every result must equal
the one the same call gave alone.
    python profiles/scripts/concurrency_stress.py [rounds]"""
import faulthandler
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lorb_slam_b200 import capi, synth  # noqa: E402

faulthandler.dump_traceback_later(240, exit=True)  # a dead-lock shows where it sits
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
pattern = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden",
                               "orb_golden.npz"))["orb/pattern"].astype(np.int32)
bank = synth.kf_bank(48, 1500, seed=3)
pa, pb = synth.all_pairs(48)
rng = np.random.default_rng(1)
q, t = synth.descriptors_uniform(1800, rng), synth.descriptors_uniform(2100, rng)
fr = synth.make_frame(2000, 4, stereo=True, claimed_frac=0.1)
pts = synth.make_proj_points(fr, 5000, 4, nobs=(0, 1, 2))
img = synth.make_orb_image(5)
ba = [synth.make_ba_problem(1, C=10, P=3000), synth.make_ba_problem(2, C=14, P=800, obs_per_point=(4, 5, 9)),
      synth.make_ba_problem(3, C=40, P=3000, obs_per_point=(5, 6, 7), traj_len=12.0)]
opt = capi.ba_options(max_num_iterations=8)


def match_work(ctx):
    s = ctx.match_sweep(bank, pa, pb)
    b = ctx.match_bf_crosscheck(q, t)
    p = ctx.search_proj_points(fr, pts, 15.0)
    o = ctx.orb_extract(img, pattern)
    return [s[0], s[1], s[2], b["q"], b["t"], b["dist"], p["kp_for_point"], p["point_for_kp"], o["x"], o["desc"]]


def ba_work(ctx):
    out = []
    for pbm in ba:
        c, p, sm = ctx.ba_local(pbm, opt)
        out += [c, p, np.array([sm["iterations"], sm["termination"]])]
    return out


errors = []


def runner(name, work, exact):
    try:
        with capi.Context(0) as ctx:
            want = work(ctx)
            start.wait()
            for r in range(rounds):
                got = work(ctx)
                for i, (g, w) in enumerate(zip(got, want)):
                    ok = np.array_equal(g, w) if exact else np.allclose(g, w, rtol=1e-9, atol=1e-11)
                    if not ok:
                        errors.append("%s round %d output %d differs" % (name, r, i))
    except Exception as e:  # noqa: BLE001
        errors.append("%s: %r" % (name, e))


start = threading.Barrier(2)
th = [threading.Thread(target=runner, args=("matching", match_work, True)),
      threading.Thread(target=runner, args=("bundle adjustment", ba_work, False))]
for x in th:
    x.start()
for x in th:
    x.join()
faulthandler.cancel_dump_traceback_later()
print("concurrency stress: %d rounds per thread, %d problems" % (rounds, len(errors)))
for e in errors[:10]:
    print("  ", e)
