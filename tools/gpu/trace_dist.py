"""Trace of the sharded MSM pipeline (run under torch.distributed.run): NMSM_TRACE prints per-window event times."""
import ctypes, os, sys, time
os.environ["NMSM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "noble-curves_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import nmsm
from nmsm import dist as nd
import bench as B
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
nmsm.init(local); dist.init_process_group("nccl", device_id=dev); nd.init()
lib = nmsm._lib.load()
n_total = 1 << 20
lo, hi = nd.shard_bounds(n_total, world, rank); n = hi - lo
pts_b, sc_b, total = B.make_terms(nmsm, n, 1000 + rank)
d_pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).to(dev); d_sc = torch.frombuffer(bytearray(sc_b), dtype=torch.uint8).to(dev)
out = ctypes.create_string_buffer(96); inf = ctypes.c_int(0)
for i in range(5):
    dist.barrier(); torch.cuda.synchronize()
    nmsm._lib.check(lib.nmsm_msm_sharded(4, d_pts.data_ptr(), d_sc.data_ptr(), n, n_total, lo, 1, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
dist.destroy_process_group()
