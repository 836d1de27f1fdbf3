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
    """Sub-problem of `rank`: its points (lorb_ba_shard_points: ordered by the lowest camera that
    observes them, cut into runs of equal observation count) with ALL their observations (so H_pp, its
    inverse and every Schur product are rank-local), every camera replicated.  Points are renumbered
    in shard order; out["point_ids"] maps them back."""
    from . import capi
    P = len(pb["pts"])
    ids = capi.ba_shard_points(P, pb["obs_cam"], pb["obs_pt"], rank, world)
    local = np.full(P, -1, np.int64)
    local[ids] = np.arange(len(ids))
    lo = local[pb["obs_pt"]]
    sel = np.flatnonzero(lo >= 0)
    sel = sel[np.argsort(lo[sel], kind="stable")]  # grouped by local point, input order inside a point
    out = dict(pb)
    out["pts"] = np.ascontiguousarray(pb["pts"][ids])
    out["obs_cam"] = np.ascontiguousarray(pb["obs_cam"][sel])
    out["obs_pt"] = np.ascontiguousarray(lo[sel].astype(np.int32))
    out["obs_uv"] = np.ascontiguousarray(pb["obs_uv"][sel])
    if len(pb.get("fix_pt", [])):
        fl = local[pb["fix_pt"]]
        fs = np.flatnonzero(fl >= 0)
        out["fix_pt"] = np.ascontiguousarray(fl[fs].astype(np.int32))
        out["fix_uv"] = np.ascontiguousarray(pb["fix_uv"][fs])
        out["fix_rt"] = np.ascontiguousarray(pb["fix_rt"][fs])
    out["P"] = len(ids)
    out["O"] = len(sel)
    out["point_ids"] = ids
    return out
