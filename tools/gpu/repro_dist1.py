import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "noble-curves_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import torch
import helpers as H
import nmsm
from nmsm import _lib, dist as nd
nmsm.init(0)
lib = _lib.load()
ident = ctypes.create_string_buffer(128)
_lib.check(lib.nmsm_dist_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
_lib.check(lib.nmsm_dist_init(0, 1, ctypes.cast(ident, ctypes.c_void_p)))
nd._dist_ready = True
dev = torch.device("cuda", 0)
for name, n in (("bls12_381_G1", 777), ("ed25519", 130), ("bn254_G2", 40), ("secp256k1", 2049)):
    P, pts, scalars, total = H.soak_inputs(name, n)
    exp = H.expected_tuple(name, H.expected_from_total(P, total))
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    tp = torch.frombuffer(bytearray(pb), dtype=torch.uint8).to(dev)
    ts = torch.frombuffer(bytearray(sb), dtype=torch.uint8).to(dev)
    torch.cuda.synchronize()
    print("running", name, n, flush=True)
    out, inf = nd.msm_sharded(H.CURVE_IDS[name], tp, ts, n, layout=(n, 0))
    print(name, (*H.unpack_point(name, out), inf) == exp, flush=True)
