"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU
paths — every unit is owned exactly once and the point-sharded
BA pieces add up to the whole problem."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lorb_slam_b200 import capi, sharding, synth
from oracle import ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # keyframe blocks: disjoint cover
        owned = torch.zeros(32, dtype=torch.int64)
        owned[sharding.sweep_blocks(rank, world, 32)] = 1
        dist.all_reduce(owned)
        assert bool((owned == 1).all())
        # tiles of the keyframe grid (C-ABI helper, host logic only): every unordered pair of a
        # ragged bank is in exactly one tile of exactly one rank
        n_kf, blk = 37, 8
        cover = torch.zeros(n_kf * (n_kf - 1) // 2, dtype=torch.int64)
        for bi, bj in capi.sweep_rank_tiles(n_kf, blk, rank, world):
            for a in range(bi * blk, min(n_kf, (bi + 1) * blk)):
                for b in range(bj * blk, min(n_kf, (bj + 1) * blk)):
                    if bi != bj or a < b:
                        cover[capi.sweep_pair_index(n_kf, a, b)] += 1
        dist.all_reduce(cover)
        assert bool((cover == 1).all())
        # windows: contiguous disjoint cover, sizes differ by at most one
        lo, hi = sharding.window_slice(rank, world, 513)
        cnt = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(cnt)
        assert int(cnt) == 513
        # point-sharded BA: per-shard costs add up to the global cost (what the
        # all-reduced LM scalars rely on), observations are split without loss
        pb = synth.make_ba_problem(21, C=5, P=301, obs_per_point=(3, 4, 5), fixed_frac=0.1)
        sh = sharding.shard_ba_by_point(pb, rank, world)
        t = torch.tensor([ref.ba_local_cost(sh), float(sh["O"]), float(len(sh["fix_pt"]))],
                         dtype=torch.float64)
        dist.all_reduce(t)
        full = ref.ba_local_cost(pb)
        assert abs(float(t[0]) - full) < 1e-9 * full
        assert int(t[1]) == pb["O"] and int(t[2]) == pb["F"]
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharding_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]


def test_sweep_tile_grid():
    assert capi.sweep_rank_tiles(10, 4, 0, 1) == [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    # a last block of one keyframe has no diagonal tile
    assert (2, 2) not in capi.sweep_rank_tiles(9, 4, 0, 1)
    idx = sorted(capi.sweep_pair_index(6, a, b) for a in range(6) for b in range(a + 1, 6))
    assert idx == list(range(15)) and capi.sweep_pair_index(6, 4, 1) == capi.sweep_pair_index(6, 1, 4)


def test_ba_point_shards_are_a_balanced_banded_partition():
    """lorb_ba_shard_points: every point in exactly one shard, observation counts level, and a shard
    meets far fewer camera pairs than an arbitrary split would (what shortens its work lists)."""
    pb = synth.make_ba_problem(4, C=120, P=4000, obs_per_point=(5, 6, 7), traj_len=120.0)
    world = 4
    ids = [capi.ba_shard_points(pb["P"], pb["obs_cam"], pb["obs_pt"], r, world) for r in range(world)]
    assert sorted(np.concatenate(ids).tolist()) == list(range(pb["P"]))
    cnt = np.bincount(pb["obs_pt"], minlength=pb["P"])
    per_rank = [int(cnt[i].sum()) for i in ids]
    assert max(per_rank) - min(per_rank) <= 2 * cnt.max()

    def pairs(point_ids):
        keep = np.isin(pb["obs_pt"], point_ids)
        oc, op = pb["obs_cam"][keep], pb["obs_pt"][keep]
        seen = set()
        for p in np.unique(op):
            cs = oc[op == p]
            seen.update((a, b) for a in cs for b in cs if a <= b)
        return len(seen)
    banded = pairs(ids[1])
    arbitrary = pairs(np.arange(pb["P"])[1::world])
    assert banded < 0.7 * arbitrary, (banded, arbitrary)
    # degenerate inputs
    assert len(capi.ba_shard_points(0, np.zeros(0, np.int32), np.zeros(0, np.int32), 0, 2)) == 0
    one = [capi.ba_shard_points(3, np.zeros(0, np.int32), np.zeros(0, np.int32), r, 2) for r in range(2)]
    assert sorted(np.concatenate(one).tolist()) == [0, 1, 2]


def test_window_slice_edges():
    assert sharding.window_slice(0, 1, 7) == (0, 7)
    assert [sharding.window_slice(r, 4, 2) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharding.sweep_blocks(3, 8, 32) == [3, 11, 19, 27]
