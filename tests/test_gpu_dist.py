"""GPU tests of the multi-GPU MSM (bucket exchange over NCCL inside libnmsm.so).

  * world = 1 in-process: the sharded code path (dense bucket finalisation, dense reduction, weighted window sums,
    gather layout, combine) on one GPU against the oracle
  * world = 2 under torch.distributed.run (skipped with fewer than 2 devices): tests/dist_worker.py"""
import os
import subprocess
import sys

import pytest

import helpers as H
from oracle import noble_ref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_path_world1():
    import ctypes

    import torch

    import nmsm
    from nmsm import _lib, dist as nd

    nmsm.init(0)
    lib = _lib.load()
    ident = ctypes.create_string_buffer(128)
    _lib.check(lib.nmsm_dist_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
    _lib.check(lib.nmsm_dist_init(0, 1, ctypes.cast(ident, ctypes.c_void_p)))
    r, w, v = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(0)
    _lib.check(lib.nmsm_dist_info(ctypes.byref(r), ctypes.byref(w), ctypes.byref(v)))
    assert (r.value, w.value) == (0, 1) and v.value > 20000
    nd._dist_ready = True
    dev = torch.device("cuda", 0)
    for name, n in (("bls12_381_G1", 777), ("ed25519", 130), ("bn254_G2", 40), ("secp256k1", 2049)):
        P, pts, scalars, total = H.soak_inputs(name, n)
        exp = H.expected_tuple(name, H.expected_from_total(P, total))
        pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
        tp = torch.frombuffer(bytearray(pb), dtype=torch.uint8).to(dev)
        ts = torch.frombuffer(bytearray(sb), dtype=torch.uint8).to(dev)
        torch.cuda.synchronize()
        try:
            for groups in (0, 8):  # bulk form (default) and one launch group per window
                nmsm.set_window_groups(groups)
                out, inf = nd.msm_sharded(H.CURVE_IDS[name], tp, ts, n, layout=(n, 0))
                assert (*H.unpack_point(name, out), inf) == exp, (name, groups)
        finally:
            nmsm.set_window_groups(0)
    # empty MSM and an invalid point
    out, inf = nd.msm_sharded(4, None, None, 0, layout=(0, 0))
    assert inf == 1
    bad = bytearray(pb)
    bad[7 * 64:7 * 64 + 32] = b"\xff" * 32
    tp = torch.frombuffer(bad, dtype=torch.uint8).to(dev)
    torch.cuda.synchronize()
    with pytest.raises(_lib.NmsmError, match="invalid point at index 7"):
        nd.msm_sharded(0, tp, ts, n, layout=(n, 0))


def test_sharded_msm_two_ranks_nccl():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "DIST_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
