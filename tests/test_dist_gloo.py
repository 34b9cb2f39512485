"""world_size-2 gloo test of the N>1 host path (nmsm.dist.msm_sharded): shard bounds, the single
all-gather of raw accumulators and the fold.  The GPU backend is replaced by the host-emulation harness
(test infrastructure) so this runs on CPU; the CUDA backend is exercised by bench.py --gpus N."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import noble_ref as R


class EmuBackend:
    def __init__(self, name):
        self.name = name
        self.lib = H.hostemu()
        self.acc_words = {"bls12_381_G1": 48, "ed25519": 32}[name]

    def partial(self, curve_id, pts, scalars, n):
        acc = np.zeros(self.acc_words, np.uint32)
        p = pts.numpy().view(np.uint32) if n else np.zeros(4, np.uint32)
        s = scalars.numpy().view(np.uint32) if n else np.zeros(8, np.uint32)
        cp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        assert self.lib.emu_msm_partial(curve_id, cp(p), cp(s), n, cp(acc)) == 0
        return torch.from_numpy(acc.view(np.uint8).copy())

    def fold(self, curve_id, accs, count):
        a = accs.numpy().view(np.uint32).copy()
        cb = H.FP_BYTES[self.name] * H.PARTS[self.name]
        out = np.zeros(2 * cb // 4, np.uint32)
        inf = np.zeros(1, np.uint32)
        cp = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        assert self.lib.emu_fold(curve_id, cp(a), count, cp(out), cp(inf)) == 0
        return out.tobytes(), int(inf[0])


def _worker(rank, world, port, name, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nmsm import dist as nd

    P, pts, scalars, total = H.soak_inputs(name, n)
    lo, hi = nd.shard_bounds(n, world, rank)
    pb = H.pack_points(name, pts[lo:hi])
    sb = H.pack_scalars(scalars[lo:hi])
    tp = torch.frombuffer(bytearray(pb), dtype=torch.uint8) if hi > lo else torch.zeros(0, dtype=torch.uint8)
    ts = torch.frombuffer(bytearray(sb), dtype=torch.uint8) if hi > lo else torch.zeros(0, dtype=torch.uint8)
    xy, inf = nd.msm_sharded(H.CURVE_IDS[name], tp, ts, hi - lo, backend=EmuBackend(name))
    x, y = H.unpack_point(name, xy)
    exp = H.expected_tuple(name, H.expected_from_total(P, total))
    q.put((rank, (x, y, inf) == exp))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("name,n", [("bls12_381_G1", 37), ("ed25519", 24), ("bls12_381_G1", 1)])
def test_sharded_msm_world2(name, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from nmsm import dist as nd

    for n in (0, 1, 7, 8, 9, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [nd.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
