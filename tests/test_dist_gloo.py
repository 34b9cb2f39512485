"""world_size-2 (and 3) gloo tests of the N>1 path on CPU: the shard layout helpers of nmsm.dist (shard bounds, offsets,
window ownership, gather slots) and the bucket-exchange scheme itself — every rank turns its shard into dense per-window
buckets, window w's buckets travel to rank w % world, the owner folds, reduces and weights the window, an all-gather of the
weighted window sums and a fold give the result (engine.cuh submit_msm with shard != nullptr).  The GPU kernels' bodies run
under the host-emulation harness (test infrastructure) and torch.distributed/gloo carries the exchange, so the partition
logic is validated without a GPU; the NCCL/CUDA implementation is exercised by tests/test_gpu_dist.py and bench.py --gpus N."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _cp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def emu_exchange_msm(name, pts_b, sc_b, n_local):
    """The sharded MSM of engine.cuh restated over torch.distributed collectives + hostemu bodies."""
    from nmsm import dist as nd

    lib = H.hostemu()
    cid = H.CURVE_IDS[name]
    n_total, offset, rank, world = nd.shard_layout(n_local)
    plan = np.zeros(4, np.uint32)
    err = np.zeros(2, np.uint32)
    assert lib.emu_shard_buckets(cid, None, None, n_local, n_total, None, _cp(plan), _cp(err)) == 0
    c, W, B, ACC = (int(v) for v in plan)
    pts = H.u32(pts_b) if n_local else np.zeros(4, np.uint32)
    sc = H.u32(sc_b) if n_local else np.zeros(8, np.uint32)
    buckets = np.zeros(W * B * ACC, np.uint32)
    assert lib.emu_shard_buckets(cid, _cp(pts), _cp(sc), n_local, n_total, _cp(buckets), _cp(plan), _cp(err)) == 0
    WB = B * ACC
    slots = nd.slots_per_rank(W, world)
    # identity accumulator = fold of zero partials is not available: take it from an owner call on an all-identity window
    gather_block = np.zeros(slots * ACC + 4, np.uint32)
    ident = np.zeros(ACC, np.uint32)
    assert lib.emu_msm_partial(cid, _cp(np.zeros(4, np.uint32)), _cp(np.zeros(8, np.uint32)), 0, _cp(ident)) == 0
    for j in range(slots):
        gather_block[j * ACC:(j + 1) * ACC] = ident
    for w in range(W - 1, -1, -1):  # top window first, like the GPU path
        owner = nd.window_owner(w, world)
        mine = torch.from_numpy(buckets[w * WB:(w + 1) * WB].view(np.int32).copy())
        if owner == rank:
            recv = []
            for r in range(world):
                if r != rank:
                    t = torch.empty(WB, dtype=torch.int32)
                    dist.recv(t, src=r)
                    recv.append(t.numpy().view(np.uint32))
            own = buckets[w * WB:(w + 1) * WB].copy()
            rc = np.concatenate(recv) if recv else np.zeros(4, np.uint32)
            acc = np.zeros(ACC, np.uint32)
            assert lib.emu_owner_window(cid, n_total, w, _cp(own), _cp(rc), len(recv), _cp(acc)) == 0
            s = nd.window_slot(w, world)
            gather_block[s * ACC:(s + 1) * ACC] = acc
        else:
            dist.send(mine, dst=owner)
    gather_block[slots * ACC:] = [err[0], err[1], offset & 0xFFFFFFFF, offset >> 32]
    blocks = [torch.empty(len(gather_block), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(blocks, torch.from_numpy(gather_block.view(np.int32).copy()))
    accs = np.concatenate([b.numpy().view(np.uint32)[: slots * ACC] for b in blocks])
    cb = H.FP_BYTES[name] * H.PARTS[name]
    out = np.zeros(2 * cb // 4, np.uint32)
    inf = np.zeros(1, np.uint32)
    assert lib.emu_fold(cid, _cp(accs), world * slots, _cp(out), _cp(inf)) == 0
    errs = [b.numpy().view(np.uint32)[slots * ACC:] for b in blocks]
    bad_pt = min([int(e[2]) + (int(e[3]) << 32) + int(e[0]) for e in errs if e[0] != 0xFFFFFFFF], default=None)
    bad_sc = min([int(e[2]) + (int(e[3]) << 32) + int(e[1]) for e in errs if e[1] != 0xFFFFFFFF], default=None)
    x, y = H.unpack_point(name, out.tobytes())
    return (x, y, int(inf[0])), bad_pt, bad_sc, (c, W, B)


def _worker(rank, world, port, name, n, q, bad_scalar_at):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nmsm import dist as nd
    from oracle import noble_ref as R

    P, pts, scalars, total = H.soak_inputs(name, n)
    if bad_scalar_at is not None:
        scalars[bad_scalar_at] = P.Fn.ORDER
    lo, hi = nd.shard_bounds(n, world, rank)
    got, bad_pt, bad_sc, plan = emu_exchange_msm(name, H.pack_points(name, pts[lo:hi]), H.pack_scalars(scalars[lo:hi]), hi - lo)
    if bad_scalar_at is not None:
        ok = bad_pt is None and bad_sc == bad_scalar_at
    else:
        ok = got == H.expected_tuple(name, H.expected_from_total(P, total)) and bad_pt is None and bad_sc is None
    q.put((rank, ok))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, name, n, bad_scalar_at=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, n, q, bad_scalar_at)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


@pytest.mark.parametrize("world,name,n", [(2, "bls12_381_G1", 37), (2, "ed25519", 24), (2, "bls12_381_G1", 1), (3, "secp256k1", 50),
                                          (2, "bls12_381_G2", 9)])  # G2: four sub-terms per term (psi split)
def test_bucket_exchange_msm_gloo(world, name, n):
    _run(world, name, n)


def test_bucket_exchange_reports_global_error_index():
    _run(2, "bls12_381_G1", 40, bad_scalar_at=31)  # lives on rank 1 at local index 11


def test_shard_layout_helpers():
    from nmsm import dist as nd

    for n in (0, 1, 7, 8, 9, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [nd.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    for W in (1, 3, 8, 10, 16, 33):
        for world in (1, 2, 3, 4, 8):
            slots = nd.slots_per_rank(W, world)
            seen = set()
            for w in range(W):
                o, s = nd.window_owner(w, world), nd.window_slot(w, world)
                assert 0 <= o < world and 0 <= s < slots and (o, s) not in seen
                seen.add((o, s))
    assert nd.shard_layout(5) == (5, 0, 0, 1)  # no process group: single rank
