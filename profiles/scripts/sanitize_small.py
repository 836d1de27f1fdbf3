"""Small pass over every C-ABI entry point (for compute-sanitizer memcheck)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
rng = np.random.default_rng(0)
q, t = synth.descriptors_uniform(300, rng), synth.descriptors_uniform(421, rng)
c.match_bf_crosscheck(q, t); c.match_bf_crosscheck(q, t, 1); c.match_knn2(q, t)
bank = synth.kf_bank(5, 333, seed=1); pa, pb = synth.all_pairs(5)
c.match_sweep(bank, pa, pb)
fr = synth.make_frame(700, 1, stereo=True, claimed_frac=0.1); pts = synth.make_proj_points(fr, 900, 1, nobs=(0, 1))
c.search_proj_points(fr, pts, 1.0); c.search_proj_points(fr, pts, 15.0)
cur, last = synth.make_frame_pair(600, 1); c.search_proj_frame(cur, last, 15.0)
c.frustum_project(synth.make_frustum_points(1000, 1))
c.stereo_matches(synth.make_stereo_pair(300, 1))
offs = np.array([0, 3, 3, 10, 30], np.int32); c.compute_descriptors(offs, synth.descriptors_uniform(30, rng))
pattern = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "orb_golden.npz"))["orb/pattern"].astype(np.int32)
c.orb_describe(synth.make_orb_inputs(300, 1), pattern)
img = synth.make_orb_image(1, 413, 307); c.orb_extract(img, pattern, nfeatures=300); c.orb_stages(img)
c.orb_selftest(np.linspace(0, 6.3, 1000), np.linspace(-5, 5, 1000))
c.profile(True)
po = synth.make_pose_only(1, 200); c.ba_pose_only(po["xw"], po["uv"], po["K"], po["rt"])
opt = capi.ba_options(max_num_iterations=4)
c.ba_local(synth.make_ba_problem(1, C=4, P=150, obs_per_point=(3, 4), fixed_frac=0.1), opt)   # dense path
c.ba_local(synth.make_ba_problem(2, C=14, P=300, obs_per_point=(4, 5, 9)), opt)                # privatised path
c.ba_local(synth.make_ba_problem(3, C=20, P=400, obs_per_point=(4, 5, 11), traj_len=6.0), opt)  # work lists + dataflow
pbs = [synth.make_ba_problem(10 + i, C=3 + i, P=60 + 10 * i, obs_per_point=(3,)) for i in range(3)]
c.ba_local_batched(synth.batch_windows(pbs), opt)
print(c.profile_read(0), c.microbench_fp64(0, 16), c.microbench_fp64(1, 16))
c.close()
print("SANITIZE_PASS_DONE")
