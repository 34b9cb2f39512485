"""Multi-GPU MSM: the term array is split across ranks, each rank reduces its shard to one raw
accumulator on its GPU, ONE all-gather exchanges the accumulators (NCCL over NVLink), and every rank
folds them.  MSM is linear in its term set (curve.ts:863: sum_i s_i*P_i), so no other exchange exists.

EC point addition is not an NCCL reduction operator, hence all-gather + a fold kernel (k_fold) rather
than ncclAllReduce.  Payload: world x nmsm_acc_bytes (192 B per GPU for BLS12-381 G1).
"""
from __future__ import annotations

import ctypes

from . import _lib


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced split of n terms: the first n % world ranks get one extra term."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CudaBackend:
    """Shard reduction and fold on the local GPU through the C ABI."""

    def __init__(self):
        _lib.ensure_init()
        self.lib = _lib.load()

    def acc_bytes(self, curve_id: int) -> int:
        return self.lib.nmsm_acc_bytes(curve_id)

    def partial(self, curve_id: int, pts, scalars, n: int):
        import torch

        acc = torch.empty(self.acc_bytes(curve_id), dtype=torch.uint8, device=pts.device)
        _lib.check(self.lib.nmsm_msm_partial_device(curve_id, pts.data_ptr() if n else None,
                                                    scalars.data_ptr() if n else None, n, acc.data_ptr()))
        return acc

    def fold(self, curve_id: int, accs, count: int):
        pb = self.lib.nmsm_point_bytes(curve_id)
        out = ctypes.create_string_buffer(pb)
        inf = ctypes.c_int(0)
        _lib.check(self.lib.nmsm_fold_partials_device(curve_id, accs.data_ptr(), count,
                                                      ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
        return out.raw, inf.value


def msm_sharded(curve_id: int, local_pts, local_scalars, n_local: int, group=None, backend=None):
    """Every rank passes ITS shard (uint8 tensors in the C-ABI packing) and gets the full MSM result."""
    import torch
    import torch.distributed as dist

    backend = backend or CudaBackend()
    acc = backend.partial(curve_id, local_pts, local_scalars, n_local)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return backend.fold(curve_id, acc, 1)
    gathered = torch.empty(acc.numel() * world, dtype=torch.uint8, device=acc.device)
    dist.all_gather_into_tensor(gathered, acc, group=group)
    if gathered.is_cuda:
        torch.cuda.current_stream(gathered.device).synchronize()
    return backend.fold(curve_id, gathered, world)
