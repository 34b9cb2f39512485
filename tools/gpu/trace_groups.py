import ctypes, os, sys, time
os.environ["NMSM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "noble-curves_b200")); sys.path.insert(0, ROOT)
import torch
import nmsm
import bench as B
nmsm.init(0)
lib = nmsm._lib.load()
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pts_b, sc_b, total = B.make_terms(nmsm, n, 1000)
dev = torch.device("cuda", 0)
d_pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).to(dev)
d_sc = torch.frombuffer(bytearray(sc_b), dtype=torch.uint8).to(dev)
out = ctypes.create_string_buffer(96); inf = ctypes.c_int(0)
for groups in (8, 4, 2):
    nmsm.set_window_groups(groups)
    print("groups", groups, file=sys.stderr, flush=True)
    for _ in range(4):
        nmsm._lib.check(lib.nmsm_msm_device(4, d_pts.data_ptr(), d_sc.data_ptr(), n, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
