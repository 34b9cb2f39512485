"""Multi-GPU MSM (BASELINE north_star: "a single NCCL allreduce over NVLink of the per-window bucket
accumulators").  One process per GPU (torchrun).  The (point, scalar) array is split contiguously across the ranks;
every rank accumulates its shard into the full W x B bucket array with the window size of the WHOLE MSM; window w is
owned by rank  w % world : its dense bucket array travels to the owner (ncclSend / ncclRecv issued by libnmsm.so on its
own stream, overlapping the accumulation of the remaining windows), the owner folds the partial buckets (EC point
addition is not an ncclRedOp_t, hence exchange + fold kernel instead of ncclAllReduce), reduces that ONE window and
applies its weight 2^(c w); one small ncclAllGather of the weighted window sums and a fold give every rank the result.
MSM is linear in its term set (curve.ts:863: sum_i s_i * P_i), so nothing else has to be exchanged.

torch.distributed is only used for the rendezvous (NCCL unique id, shard sizes); the data path is inside the library.
"""
from __future__ import annotations

import ctypes

from . import _lib

_dist_ready = False


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced split of n terms: the first n % world ranks get one extra term."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def window_owner(w: int, world: int) -> int:
    """Rank that reduces bucket window w (engine.cuh submit_msm: own_rank = w % world)."""
    return w % world


def window_slot(w: int, world: int) -> int:
    """Index of window w among its owner's windows (its place in the owner's gather block: w // world)."""
    return w // world


def slots_per_rank(windows: int, world: int) -> int:
    return (windows + world - 1) // world


def shard_layout(n_local: int, group=None):
    """(n_total, offset of this rank's shard, rank, world) from every rank's local term count."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return n_local, 0, 0, 1
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(n_local), group=group)
    return sum(sizes), sum(sizes[:rank]), rank, world


def init(group=None) -> None:
    """Create the library's NCCL communicator: rank 0 makes the id, torch.distributed carries it to the others."""
    global _dist_ready
    import torch.distributed as dist

    if _dist_ready:
        return
    _lib.ensure_init()
    lib = _lib.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ident = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.check(lib.nmsm_dist_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
    if world > 1:
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = ctypes.create_string_buffer(box[0], 128)
    _lib.check(lib.nmsm_dist_init(rank, world, ctypes.cast(ident, ctypes.c_void_p)))
    _dist_ready = True


def msm_sharded(curve_id: int, local_pts, local_scalars, n_local: int, group=None, layout=None):
    """Collective: every rank passes ITS shard (uint8 CUDA tensors in the C-ABI packing, or None when n_local == 0) and
    gets the full MSM result (xy bytes, is_inf).  `layout` = (n_total, offset) skips the size exchange."""
    init(group)
    lib = _lib.load()
    if layout is None:
        n_total, offset, _, _ = shard_layout(n_local, group)
    else:
        n_total, offset = layout
    pb = lib.nmsm_point_bytes(curve_id)
    out = ctypes.create_string_buffer(pb)
    inf = ctypes.c_int(0)
    if n_local:
        # the library's streams are not ordered against torch's: whatever produced the shard must have finished
        # (include/nmsm.h "Stream ordering")
        import torch

        torch.cuda.current_stream(local_pts.device).synchronize()
    rc = lib.nmsm_msm_sharded(curve_id, local_pts.data_ptr() if n_local else None,
                              local_scalars.data_ptr() if n_local else None, n_local, n_total, offset, 1,
                              ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf))
    _lib.check(rc)
    return out.raw, inf.value
