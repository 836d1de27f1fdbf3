"""How the hot path shards over the GPUs of one box.

  keyframe-pair sweep   independent units: keyframe blocks round-robin over ranks,
                        bank replicated, no collective during compute
  batched local BA      independent windows: contiguous slice per rank
  large BA              points (with all their observations) split contiguously
                        over ranks, cameras replicated; the reduced camera system
                        is all-reduced every LM iteration (dist.cu)
Single frame-pair matching, projection search and a single small BA do not
shard: `--gpus N` runs replicas.
"""
import numpy as np


def sweep_tiles(rank, world, n_kf, block_kf):
    """(block row, block column) tiles of the keyframe grid owned by `rank` (lorb_sweep_rank_tiles)."""
    from . import capi
    return capi.sweep_rank_tiles(n_kf, block_kf, rank, world)


def sweep_blocks(rank, world, n_blocks):
    """Blocks of the keyframe bank owned by `rank`."""
    return list(range(rank, n_blocks, world))


def window_slice(rank, world, n_windows):
    """Contiguous [lo, hi) slice of the independent windows owned by `rank` (the C ABI's
    lorb_shard_range, so that Python and C++ hosts split identically)."""
    from . import capi
    return capi.shard_range(n_windows, rank, world)


def shard_ba_by_point(pb, rank, world):
    """Sub-problem of `rank`: a contiguous range of points with ALL their
    observations (so H_pp, its inverse and every Schur product are rank-local),
    every camera replicated.  Point indices are renumbered from 0."""
    P = len(pb["pts"])
    lo, hi = window_slice(rank, world, P)
    sel = (pb["obs_pt"] >= lo) & (pb["obs_pt"] < hi)
    out = dict(pb)
    out["pts"] = np.ascontiguousarray(pb["pts"][lo:hi])
    out["obs_cam"] = np.ascontiguousarray(pb["obs_cam"][sel])
    out["obs_pt"] = np.ascontiguousarray(pb["obs_pt"][sel] - lo)
    out["obs_uv"] = np.ascontiguousarray(pb["obs_uv"][sel])
    if len(pb.get("fix_pt", [])):
        fs = (pb["fix_pt"] >= lo) & (pb["fix_pt"] < hi)
        out["fix_pt"] = np.ascontiguousarray(pb["fix_pt"][fs] - lo)
        out["fix_uv"] = np.ascontiguousarray(pb["fix_uv"][fs])
        out["fix_rt"] = np.ascontiguousarray(pb["fix_rt"][fs])
    out["P"] = hi - lo
    out["O"] = int(sel.sum())
    out["point_range"] = (lo, hi)
    return out
