#!/usr/bin/env python3
"""Exploration: single-MSM latency (device-resident BLS12-381 G1, 2^logn terms) vs accumulate segment length L, reduce
chunk K and window-group count — each configuration in a fresh process (the plan parameters are read from the environment)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.join(%r, "noble-curves_b200")); sys.path.insert(0, %r)
import torch, nmsm, bench as B
nmsm.init(0); lib = nmsm._lib.load()
logn = int(sys.argv[1]); groups = int(sys.argv[2]); curve = int(sys.argv[3]); n = 1 << logn
pts_b, sc_b, total = B.make_terms(nmsm, n, 1000)
exp_xy, exp_inf = B.expected_point(nmsm, total)
dev = torch.device("cuda", 0)
d_pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).to(dev); d_sc = torch.frombuffer(bytearray(sc_b), dtype=torch.uint8).to(dev)
out = ctypes.create_string_buffer(96); inf = ctypes.c_int(0)
nmsm.set_window_groups(groups)
def run():
    nmsm._lib.check(lib.nmsm_msm_device(curve, d_pts.data_ptr(), d_sc.data_ptr(), n, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
for _ in range(3): run()
assert out.raw == exp_xy and inf.value == exp_inf
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): run()
torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 10
nmsm.set_profiling(True)
for _ in range(3): run()
ms, info = nmsm.last_timing()
print(json.dumps({"logn": logn, "curve": curve, "groups": groups, "L": os.environ.get("NMSM_L"), "K": os.environ.get("NMSM_K"), "wall_ms": round(el*1e3, 4),
                  "linear": {k: round(v, 3) for k, v in ms.items()}, "c": info.c, "starts": info.bucket_starts}))
''' % (ROOT, ROOT)
def main():
    logn = sys.argv[1] if len(sys.argv) > 1 else "20"
    cfgs = []
    for L in (None, "40", "48", "56", "64"):
        for K in (None, "4", "16"):
            cfgs.append((L, K, "1", "4"))
    cfgs += [(None, None, "8", "4"), ("56", None, "8", "4"), ("56", "4", "8", "4"), (None, None, "1", "6"), ("56", None, "1", "6")]
    for L, K, groups, curve in cfgs:
        env = dict(os.environ)
        if L: env["NMSM_L"] = L
        if K: env["NMSM_K"] = K
        r = subprocess.run([sys.executable, "-c", CHILD, logn, groups, curve], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or json.dumps({"L": L, "K": K, "error": r.stderr[-300:]}), flush=True)
if __name__ == "__main__":
    main()
