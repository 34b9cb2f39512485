"""Single-GPU run of the SHARDED pipeline (world = 1: dense buckets, dense reduction, weighted window sums, gather layout)
for per-kernel ncu timing of the kernels a multi-GPU rank runs: python tools/gpu/shard1_profile.py [logn]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "noble-curves_b200")); sys.path.insert(0, ROOT)
import nmsm
import torch
import bench as B
from nmsm import _lib, dist as nd
nmsm.init(0)
lib = _lib.load()
ident = ctypes.create_string_buffer(128)
_lib.check(lib.nmsm_dist_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
_lib.check(lib.nmsm_dist_init(0, 1, ctypes.cast(ident, ctypes.c_void_p)))
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 19)
pts_b, sc_b, total = B.make_terms(nmsm, n, 1000)
exp_xy, exp_inf = B.expected_point(nmsm, total)
dev = torch.device("cuda", 0)
d_pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).to(dev); d_sc = torch.frombuffer(bytearray(sc_b), dtype=torch.uint8).to(dev)
out = ctypes.create_string_buffer(96); inf = ctypes.c_int(0)
for i in range(3):
    _lib.check(lib.nmsm_msm_sharded(4, d_pts.data_ptr(), d_sc.data_ptr(), n, n, 0, 1, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
    assert out.raw == exp_xy and inf.value == exp_inf
print("ok", nmsm.last_timing()[0]["total"])
