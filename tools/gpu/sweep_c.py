#!/usr/bin/env python3
"""Calibration data for make_plan's time model (msm_body.cuh): device time of ONE MSM per (curve, size, window bits c), with the
per-kernel breakdown of a profiled pass.  Inputs are k_i * G built on the GPU; every result is checked against (sum k_i s_i) * G.
Also times nmsm_mul_batch over batch sizes (the lane-parallel form for small batches vs the one-thread-per-item form)."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "noble-curves_b200")); sys.path.insert(0, ROOT)
import nmsm

def pack(v, nbytes):
    return v.to_bytes(nbytes, "little") if isinstance(v, int) else b"".join(c.to_bytes(nbytes, "little") for c in v)

def terms(C, n, seed):
    order, cid, fb = C.Fn.ORDER, C.CURVE_ID, C.FP_BYTES
    rnd = random.Random(seed)
    ks = [rnd.randrange(1, order) for _ in range(n)]
    sc = [rnd.randrange(order) for _ in range(n)]
    g = pack(C.BASE.x, fb) + pack(C.BASE.y, fb)
    scb = b"".join(s.to_bytes(32, "little") for s in sc)
    pts, _ = nmsm.mul_batch_packed(cid, g * n, b"".join(k.to_bytes(32, "little") for k in ks), n, False)
    total = sum(k * s for k, s in zip(ks, sc)) % order
    exp, _ = nmsm.mul_batch_packed(cid, g, total.to_bytes(32, "little"), 1, True)
    return pts, scb, exp

def sweep_msm(name, logn, cs, ids=None):
    C = nmsm.CURVES[name]
    n = 1 << logn
    pts, scb, exp = terms(C, n, 7 + logn)
    for cid in (ids or [C.CURVE_ID]):
        for c in cs:
            nmsm.set_window_bits(c)
            nmsm.set_profiling(False)
            ok = True
            for _ in range(2):
                o = nmsm.msm_packed(cid, pts, scb, n)
                ok = ok and o[0] == exp and o[1] == 0
            best = 1e9
            for _ in range(5):
                nmsm.msm_packed(cid, pts, scb, n)
                ms, info = nmsm.last_timing()
                best = min(best, ms["total"])
            nmsm.set_profiling(True)
            nmsm.msm_packed(cid, pts, scb, n)
            nmsm.msm_packed(cid, pts, scb, n)
            lin, info = nmsm.last_timing()
            print(json.dumps({"curve": name, "id": cid, "logn": logn, "c_req": c, "c": info.c, "W": info.windows, "L": info.entries_per_thread,
                              "ms_device": round(best, 4), "ok": ok, "linear": {k: round(v, 3) for k, v in lin.items() if v}}), flush=True)
    nmsm.set_window_bits(0)
    nmsm.set_profiling(False)

def sweep_mul(name, sizes):
    C = nmsm.CURVES[name]
    order, cid, fb = C.Fn.ORDER, C.CURVE_ID, C.FP_BYTES
    rnd = random.Random(5)
    g = pack(C.BASE.x, fb) + pack(C.BASE.y, fb)
    b0 = rnd.randrange(1, order)
    pk, _ = nmsm.mul_batch_packed(cid, g, b0.to_bytes(32, "little"), 1, False)
    for n in sizes:
        ks = [rnd.randrange(1, order) for _ in range(n)]
        ksb = b"".join(k.to_bytes(32, "little") for k in ks)
        chk = b"".join(((k * b0) % order).to_bytes(32, "little") for k in ks)
        want, _ = nmsm.mul_batch_packed(cid, g * n, chk, n, True)
        for form, qmax in (("quad", str(1 << 30)), ("serial", "0")):
            os.environ["NMSM_MUL_QUAD_MAX"] = qmax
            got, _ = nmsm.mul_batch_packed(cid, pk * n, ksb, n, False)
            best = 1e9
            for _ in range(6):
                t0 = time.perf_counter()
                nmsm.mul_batch_packed(cid, pk * n, ksb, n, False)
                best = min(best, time.perf_counter() - t0)
            nmsm.set_profiling(True)
            nmsm.mul_batch_packed(cid, pk * n, ksb, n, False)
            ms, _ = nmsm.last_timing()
            nmsm.set_profiling(False)
            print(json.dumps({"mul_batch": name, "n": n, "form": form, "ms_host_buffers": round(best * 1e3, 4), "ms_kernel": round(ms["total"], 4), "ok": got == want}), flush=True)
        del os.environ["NMSM_MUL_QUAD_MAX"]

def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    nmsm.init(0)
    if what in ("all", "mul"):
        sweep_mul("secp256k1", [1024, 8192, 65536])
        sweep_mul("bls12_381_G1", [1024, 16384, 65536])
        sweep_mul("bls12_381_G2", [1024])
    if what == "secp":
        sweep_msm("secp256k1", 16, [0])
        sweep_msm("secp256k1", 20, [0])
        sweep_msm("ed25519", 17, [0])
    if what == "g2quick":
        sweep_msm("bls12_381_G2", 18, [0], ids=[5, 7])
        sweep_msm("bls12_381_G2", 14, [0], ids=[5])
        sweep_msm("bn254_G2", 18, [0])
        sweep_mul("bls12_381_G2", [1024, 32768])
    if what == "g2":
        sweep_msm("bls12_381_G2", 18, [0, 13, 14, 15, 16], ids=[5, 7])
        sweep_msm("bls12_381_G2", 14, [0, 11, 13, 16], ids=[5])
        sweep_msm("bls12_381_G2", 20, [0], ids=[5])
    if what in ("all", "msm"):
        sweep_msm("bls12_381_G1", 14, [9, 10, 11, 12, 13, 14])
        sweep_msm("bls12_381_G1", 16, [0, 10, 11, 12, 13, 14, 15, 16])
        sweep_msm("bls12_381_G1", 18, [0, 13, 14, 15, 16])
        sweep_msm("bls12_381_G2", 18, [0, 11, 12, 13, 14, 15, 16])
        sweep_msm("bls12_381_G2", 14, [0, 9, 10, 11, 12, 13])
        sweep_msm("bn254_G1", 20, [0, 14, 15, 16])
        sweep_msm("secp256k1", 16, [0, 11, 12, 13, 14, 16])
        sweep_msm("ed25519", 17, [0, 11, 12, 13, 14, 16])

if __name__ == "__main__":
    main()
