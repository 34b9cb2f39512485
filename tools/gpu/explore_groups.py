#!/usr/bin/env python3
"""Exploration (not the bench): single-MSM latency of the device-resident BLS12-381 G1 MSM vs window-group count,
and throughput with several MSMs in flight.  Usage: python tools/gpu/explore_groups.py [logn ...]"""
import ctypes, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "noble-curves_b200")); sys.path.insert(0, ROOT)
import torch
import nmsm
import bench as B

def main():
    logns = [int(a) for a in sys.argv[1:]] or [20]
    nmsm.init(0)
    lib = nmsm._lib.load()
    for curve in (4, 6):
      for logn in logns:
        n = 1 << logn
        pts_b, sc_b, total = B.make_terms(nmsm, n, 1000)
        exp_xy, exp_inf = B.expected_point(nmsm, total)
        dev = torch.device("cuda", 0)
        d_pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).to(dev)
        d_sc = torch.frombuffer(bytearray(sc_b), dtype=torch.uint8).to(dev)
        out = ctypes.create_string_buffer(96); inf = ctypes.c_int(0)
        for groups in (1, 2, 4, 8, 0):
            nmsm.set_window_groups(groups)
            for _ in range(3):
                nmsm._lib.check(lib.nmsm_msm_device(curve, d_pts.data_ptr(), d_sc.data_ptr(), n, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
            assert out.raw == exp_xy and inf.value == exp_inf
            torch.cuda.synchronize(); t0 = time.perf_counter(); dev_ms = []
            K = 10
            for _ in range(K):
                nmsm._lib.check(lib.nmsm_msm_device(curve, d_pts.data_ptr(), d_sc.data_ptr(), n, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
                dev_ms.append(nmsm.last_timing()[0]["total"])
            torch.cuda.synchronize(); el = (time.perf_counter() - t0) / K
            info = nmsm.last_timing()[1]
            print(json.dumps({"curve": curve, "logn": logn, "groups_req": groups, "groups": info.window_groups, "c": info.c, "W": info.windows,
                              "wall_ms": round(el * 1e3, 4), "device_ms": round(sum(dev_ms) / K, 4), "launches": info.launches}), flush=True)
        # in flight
        for groups in (1, 0):
            nmsm.set_window_groups(groups)
            for NF in (2, 4):
                outs = [ctypes.create_string_buffer(96) for _ in range(NF)]; infs = [ctypes.c_int(0) for _ in range(NF)]
                def run(steps):
                    for i in range(steps):
                        s = i % NF
                        if i >= NF:
                            nmsm._lib.check(lib.nmsm_msm_collect(s, ctypes.cast(outs[s], ctypes.c_void_p), ctypes.byref(infs[s])))
                        nmsm._lib.check(lib.nmsm_msm_submit(curve, d_pts.data_ptr(), d_sc.data_ptr(), n, 1, s))
                    for j in range(max(0, steps - NF), steps):
                        nmsm._lib.check(lib.nmsm_msm_collect(j % NF, ctypes.cast(outs[j % NF], ctypes.c_void_p), ctypes.byref(infs[j % NF])))
                run(2 * NF + 1)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                run(12)
                torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 12
                assert outs[0].raw == exp_xy
                print(json.dumps({"curve": curve, "logn": logn, "groups_req": groups, "in_flight": NF, "ms_per_step": round(el * 1e3, 4)}), flush=True)
        nmsm.set_window_groups(0)

if __name__ == "__main__":
    main()
