#!/usr/bin/env python3
"""Headline benchmark: BLS12-381 G1 MSM points/sec at 2^20 scalars (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (default N=1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" = one complete MSM (curve.ts:863-905 `pippenger` semantics) over one batch of synthetic
(point, scalar) terms:  points P_i = k_i*G, scalars uniform in [0, n).  At N GPUs the term array is
sharded across ranks, each rank reduces its shard to one raw accumulator, the accumulators are
exchanged with one NCCL all-gather and folded (MSM is linear in its term set).  Default "scaling": "weak":
every GPU holds 2^20 terms of ONE MSM with N*2^20 terms; `--scaling strong` splits 2^20 terms over the GPUs.

Timed numbers:
  value  — whole-job points/s with inputs already resident in HBM (CUDA path via nmsm_msm_device)
  e2e    — same metric through the host-buffer entry point nmsm_msm (pinned host inputs, H2D inside)
  roofline — modmul-bound integer roofline of the dominant kernel (k_accumulate): executed field
             multiplications / CUDA-event time, against the register-resident Montgomery-multiply
             microbenchmark measured in the same run; HBM figures alongside
  cpu_baseline — the reference algorithm restated for the CPU (oracle/), timed on a bounded sample
"""
import argparse
import ctypes
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "noble-curves_b200"))
sys.path.insert(0, ROOT)

BLS_G1 = 4
POINT_BYTES = 96
BLS_N = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
BLS_GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--logn", type=int, default=20, help="log2 of the MSM size (headline: 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fixed-base", action="store_true", help="skip the fixed-base table companion measurement")
    ap.add_argument("--window", type=int, default=0, help="force window bits c (0 = cost model)")
    ap.add_argument("--in-flight", type=int, default=4, choices=[1, 2, 3, 4], help="MSMs kept in flight in the timed region")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = 2^logn terms PER GPU (one MSM of N*2^logn terms); strong = 2^logn terms in total")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------
def make_terms(nmsm, n, seed):
    """n random (point, scalar) terms; points are generated ON THE GPU as k_i*G (nmsm_mul_batch).
    Returns (points bytes, scalars bytes, expected total scalar sum k_i*s_i mod n)."""
    rnd = random.Random(seed)
    ks = [rnd.randrange(1, BLS_N) for _ in range(n)]
    sc = [rnd.randrange(BLS_N) for _ in range(n)]
    g = BLS_GX.to_bytes(48, "little") + BLS_GY.to_bytes(48, "little")
    pts, infs = nmsm.mul_batch_packed(BLS_G1, g * n, b"".join(k.to_bytes(32, "little") for k in ks), n, False)
    total = sum(k * s for k, s in zip(ks, sc)) % BLS_N
    return pts, b"".join(s.to_bytes(32, "little") for s in sc), total


def expected_point(nmsm, total):
    """(sum k_i s_i)*G computed by the GPU scalar-mult path (itself parity-tested against the oracle)."""
    if total == 0:
        return bytes(96), 1
    g = BLS_GX.to_bytes(48, "little") + BLS_GY.to_bytes(48, "little")
    out, infs = nmsm.mul_batch_packed(BLS_G1, g, total.to_bytes(32, "little"), 1, True)
    return out, infs[0]


# --------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on host cores (oracle/ is allowed here only)
# --------------------------------------------------------------------------------------------
def cpu_reference(n_sample, seed, max_seconds=30.0):
    """Times the CPU restatement of noble's pippenger (same window rule, unsigned windows, complete
    RCB additions) on a bounded sample.  Prefers the C port (oracle/ref_msm.c, all host threads);
    falls back to the Python-int oracle on a smaller sample."""
    from oracle import cpu_baseline

    return cpu_baseline.time_bls_g1_msm(n_sample, seed, max_seconds)


def run_reference(args, real_stdout):
    """`--impl reference`: the reference's algorithm on this box's host cores (rank 0 only under torchrun).
    Inputs are generated once; each step is one timed run of the C port on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as CB

    n_full = 1 << args.logn
    n_sample = CB.choose_sample(30.0, args.logn)
    w = CB.Workload(n_sample, 1)
    times = [w.run() for _ in range(args.warmup + args.steps)][args.warmup:]
    res = CB.describe(n_sample, sum(times) / len(times), args.logn)
    pts_per_s = res["points_per_s_at_full_size"]
    line = {
        "impl": "reference",
        "metric": "BLS12-381 G1 MSM points/sec at 2^%d scalars" % args.logn,
        "value": pts_per_s, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n_full / pts_per_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64-limb integer (381-bit Fp, Montgomery)", "data": "synthetic",
        "config": {"workload": "BLS12-381 G1 MSM, 2^%d terms, reference algorithm (curve.ts:863-905) on host cores; "
                               "host-only: the same single-box workload at every --gpus" % args.logn},
        "cpu_baseline": {"value": pts_per_s, "unit": "points/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": pts_per_s, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(real_stdout, line)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries write there too (NCCL prints its version banner
    to fd 1 when the launch environment sets NCCL_DEBUG=VERSION/WARN).  Point fd 1 at stderr for the whole run and keep
    the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout_fd, line: dict):
    sys.stdout.flush()
    os.write(real_stdout_fd, (json.dumps(line) + "\n").encode())


def main():
    args = parse_args()
    real_stdout = _claim_stdout()
    if args.impl == "reference":
        run_reference(args, real_stdout)
        return

    import torch
    import torch.distributed as dist

    import nmsm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MSM path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nmsm.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = nmsm._lib.load()
    if args.window:
        nmsm.set_window_bits(args.window)

    # weak scaling (default): every GPU holds a 2^logn-term shard of ONE MSM with world*2^logn terms, so per-GPU
    # work is fixed; strong: the 2^logn terms are split across the GPUs.  Either way: one all-gather + fold.
    scaling = args.scaling if world > 1 else "weak"
    n_local = (1 << args.logn) if scaling == "weak" else (1 << args.logn) // world
    n_total = n_local * world
    # every rank generates only its own shard; seeds make shards disjoint and reproducible
    pts_b, sc_b, total_local = make_terms(nmsm, n_local, 1000 + rank)
    dev = torch.device("cuda", local_rank)
    d_pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).to(dev)
    d_sc = torch.frombuffer(bytearray(sc_b), dtype=torch.uint8).to(dev)
    acc_bytes = lib.nmsm_acc_bytes(BLS_G1)
    d_acc = torch.zeros(acc_bytes, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(acc_bytes * world, dtype=torch.uint8, device=dev)
    out = ctypes.create_string_buffer(POINT_BYTES)
    inf = ctypes.c_int(0)

    if world > 1:
        totals = [None] * world
        dist.all_gather_object(totals, total_local)
        total = sum(totals) % BLS_N
    else:
        total = total_local
    exp_xy, exp_inf = expected_point(nmsm, total)

    def step_device():
        if world == 1:
            nmsm._lib.check(lib.nmsm_msm_device(BLS_G1, d_pts.data_ptr(), d_sc.data_ptr(), n_local,
                                                ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
        else:
            nmsm._lib.check(lib.nmsm_msm_partial_device(BLS_G1, d_pts.data_ptr(), d_sc.data_ptr(), n_local,
                                                        d_acc.data_ptr()))
            dist.all_gather_into_tensor(d_all, d_acc)
            torch.cuda.current_stream().synchronize()
            nmsm._lib.check(lib.nmsm_fold_partials_device(BLS_G1, d_all.data_ptr(), world,
                                                          ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: (1) serial steps -> latency + per-kernel CUDA-event times -------------
    nmsm.set_profiling(True)
    for _ in range(max(args.warmup, 3)):
        step_device()
    assert out.raw == exp_xy and inf.value == exp_inf, "MSM result does not match (sum k_i s_i)*G"
    acc_ms, tot_ms, kern_ms = [], [], {}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
        ms, info = nmsm.last_timing()
        acc_ms.append(ms["accumulate"])
        tot_ms.append(ms["total"])
        for k, v in ms.items():
            kern_ms[k] = kern_ms.get(k, 0.0) + v / args.steps
    barrier()
    serial_elapsed = time.perf_counter() - t0
    assert out.raw == exp_xy and inf.value == exp_inf
    nmsm.set_profiling(False)

    # ---- (2) the timed region: K steps, --in-flight MSMs in flight (submit/collect round-robin over the slots) -----------
    # At N > 1 the shard reduction of step i overlaps the all-gather + fold of step i-1 the same way.
    pipelined = True
    NF = args.in_flight
    d_accs = [torch.zeros(acc_bytes, dtype=torch.uint8, device=dev) for _ in range(NF)]

    def run_pipelined(steps, submit):
        outs = [ctypes.create_string_buffer(POINT_BYTES) for _ in range(NF)]
        infs = [ctypes.c_int(0) for _ in range(NF)]
        def take(s):
            collect(s, outs[s], infs[s])
            assert outs[s].raw == exp_xy and infs[s].value == exp_inf

        for i in range(steps):
            s = i % NF
            if i >= NF:  # the slot still holds step i - NF
                take(s)
            submit(s)
        for j in range(max(0, steps - NF), steps):
            take(j % NF)

    def submit_device(slot, pts_t=None, sc_t=None):
        pts_t = d_pts if pts_t is None else pts_t
        sc_t = d_sc if sc_t is None else sc_t
        if world == 1:
            nmsm._lib.check(lib.nmsm_msm_submit(BLS_G1, pts_t.data_ptr(), sc_t.data_ptr(), n_local, 1, slot))
        else:
            nmsm._lib.check(lib.nmsm_msm_submit_partial(BLS_G1, pts_t.data_ptr(), sc_t.data_ptr(), n_local,
                                                        d_accs[slot].data_ptr(), slot))

    def collect(slot, o, f):
        if world == 1:
            nmsm._lib.check(lib.nmsm_msm_collect(slot, ctypes.cast(o, ctypes.c_void_p), ctypes.byref(f)))
        else:  # shard done -> one all-gather of the raw accumulators -> fold
            nmsm._lib.check(lib.nmsm_msm_collect(slot, None, None))
            dist.all_gather_into_tensor(d_all, d_accs[slot])
            torch.cuda.current_stream().synchronize()
            nmsm._lib.check(lib.nmsm_fold_partials_device(BLS_G1, d_all.data_ptr(), world, ctypes.cast(o, ctypes.c_void_p),
                                                          ctypes.byref(f)))

    sampler = ClockSampler(local_rank)
    if pipelined:
        run_pipelined(2 * NF + 1, submit_device)  # touches every slot (workspace allocation) before the timed region
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    if pipelined:
        run_pipelined(args.steps, submit_device)
    else:
        for _ in range(args.steps):
            step_device()
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = n_total * args.steps / elapsed

    # ---- end-to-end through the host-buffer entry point (pinned host inputs) --------------
    h_pts = lib.nmsm_host_alloc(len(pts_b))
    h_sc = lib.nmsm_host_alloc(len(sc_b))
    ctypes.memmove(h_pts, pts_b, len(pts_b))
    ctypes.memmove(h_sc, sc_b, len(sc_b))
    e2e_steps = max(3, args.steps)  # the same K steps as the device-resident region
    e2e_bufs = ([(d_pts, d_sc)] + [(torch.empty_like(d_pts), torch.empty_like(d_sc)) for _ in range(NF - 1)]) if world > 1 else None

    def step_e2e():
        if world == 1:
            nmsm._lib.check(lib.nmsm_msm(BLS_G1, h_pts, h_sc, n_local, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
        else:
            d_pts.copy_(torch.frombuffer((ctypes.c_uint8 * len(pts_b)).from_address(h_pts), dtype=torch.uint8), non_blocking=True)
            d_sc.copy_(torch.frombuffer((ctypes.c_uint8 * len(sc_b)).from_address(h_sc), dtype=torch.uint8), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            step_device()

    def submit_host(slot):
        if world == 1:
            nmsm._lib.check(lib.nmsm_msm_submit(BLS_G1, h_pts, h_sc, n_local, 0, slot))
        else:  # multi-GPU e2e: H2D on torch's stream into this slot's device buffers, then the sharded step
            tp, ts = e2e_bufs[slot]
            tp.copy_(torch.frombuffer((ctypes.c_uint8 * len(pts_b)).from_address(h_pts), dtype=torch.uint8), non_blocking=True)
            ts.copy_(torch.frombuffer((ctypes.c_uint8 * len(sc_b)).from_address(h_sc), dtype=torch.uint8), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            submit_device(slot, tp, ts)

    step_e2e()
    if pipelined:
        run_pipelined(NF + 1, submit_host)
    barrier()
    t0 = time.perf_counter()
    if pipelined:
        run_pipelined(e2e_steps, submit_host)
    else:
        for _ in range(e2e_steps):
            step_e2e()
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    assert out.raw == exp_xy and inf.value == exp_inf
    if world > 1:
        t = torch.tensor([e2e_elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = n_total * e2e_steps / e2e_elapsed
    lib.nmsm_host_free(h_pts)
    lib.nmsm_host_free(h_sc)
    main_timing = nmsm.last_timing()  # plan + kernel times of the general MSM, before the fixed-base pass below

    # ---- companion number (not the headline): the same MSM over a device-resident point set with a fixed-base
    # table (nmsm_points_precompute): scalars resident on the device, same pipelined loop ----------
    fixed = None
    if world == 1 and not args.no_fixed_base:
        h, tc, lv = ctypes.c_uint64(0), ctypes.c_int(0), ctypes.c_int(0)
        nmsm._lib.check(lib.nmsm_points_upload(BLS_G1, ctypes.cast(ctypes.c_char_p(pts_b), ctypes.c_void_p), n_local, ctypes.byref(h)))
        t0 = time.perf_counter()
        nmsm._lib.check(lib.nmsm_points_precompute(h, 0, ctypes.byref(tc), ctypes.byref(lv)))
        t_pre = time.perf_counter() - t0

        def submit_fixed(slot):
            nmsm._lib.check(lib.nmsm_msm_points_submit(h, d_sc.data_ptr(), n_local, 1, slot))

        nmsm.set_profiling(True)
        for _ in range(3):
            submit_fixed(0)
            collect(0, out, inf)
            assert out.raw == exp_xy and inf.value == exp_inf
        fms, finfo = nmsm.last_timing()
        nmsm.set_profiling(False)
        run_pipelined(2 * NF + 1, submit_fixed)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps, submit_fixed)
        barrier()
        f_el = time.perf_counter() - t0
        fixed = {"value": n_local * args.steps / f_el, "unit": "points/s", "ms_per_step": 1e3 * f_el / args.steps,
                 "latency_ms_single_msm": fms["total"], "in_flight": NF,
                 "table": {"window_bits": tc.value, "levels": lv.value, "bytes": lv.value * 2 * n_local * 96,
                           "precompute_ms": t_pre * 1e3},
                 "plan": {"c": finfo.c, "windows": finfo.windows, "sorted_entries": finfo.sorted_entries},
                 "kernel_ms_breakdown": {k: round(v, 4) for k, v in fms.items()},
                 "note": "device-resident point set + table 2^(c*j)*P (nmsm_points_precompute); companion to `value`, "
                         "which stays the general MSM with points passed per call"}
        lib.nmsm_points_free(h)

    # ---- second companion: the same MSM under NMSM_BLS12_381_G1_ANY (no subgroup assumption, no GLV: 16 windows) ----
    any_point = None
    if world == 1 and not args.no_fixed_base:
        ANY = 6
        for _ in range(2):
            nmsm._lib.check(lib.nmsm_msm_device(ANY, d_pts.data_ptr(), d_sc.data_ptr(), n_local,
                                                ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
        assert out.raw == exp_xy and inf.value == exp_inf
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            nmsm._lib.check(lib.nmsm_msm_device(ANY, d_pts.data_ptr(), d_sc.data_ptr(), n_local,
                                                ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
        torch.cuda.synchronize()
        a_el = (time.perf_counter() - t0) / 5
        any_point = {"latency_ms_single_msm": 1e3 * a_el, "value_serial": n_local / a_el, "unit": "points/s",
                     "note": "curve id NMSM_BLS12_381_G1_ANY: valid for every on-curve point, not only the prime-order "
                             "subgroup the headline id assumes (the bench points k_i*G are in it); serial calls"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------
    ms, info = main_timing
    peak = 0.0
    for (bps, thr, ilp) in ((4, 128, 1), (8, 128, 1), (4, 256, 1), (4, 128, 2), (8, 128, 2), (2, 256, 2)):
        peak = max(peak, nmsm.bench_modmul(1, bps, thr, 3000, ilp))
    acc_t = sum(acc_ms) / len(acc_ms) * 1e-3
    madd_modmuls = info.sorted_entries * 10  # madd-2008-s: 8M + 2S per mixed addition
    achieved = madd_modmuls / acc_t
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # algorithmic bytes of k_accumulate per launch: every sorted entry gathers one 96-byte affine point + its
    # 4-byte index; every bucket (and partial) is written once (192 B)
    acc_bytes_alg = info.sorted_entries * (96 + 4) + (info.windows * info.buckets_per_window) * 192
    roofline = {
        "bound": "int-modmul", "kernel": "k_accumulate<BLS12-381 G1>",
        "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "Gmodmul/s (381-bit Montgomery)",
        "frac": achieved / peak if peak > 0 else None,
        "peak_source": "nmsm_bench_modmul: register-resident mont_mul<FpBls381> microbenchmark, same run",
        "modmul_per_launch": madd_modmuls, "kernel_ms": acc_t * 1e3,
        "whole_msm": {"modmul_equiv": info.modmul_equiv, "ms": sum(tot_ms) / len(tot_ms),
                      "frac": (info.modmul_equiv / (sum(tot_ms) / len(tot_ms) * 1e-3)) / peak if peak > 0 else None,
                      # same work against the pipelined step time of this GPU (MSMs overlapped on separate streams)
                      "frac_pipelined": (info.modmul_equiv / (elapsed / args.steps)) / peak if peak > 0 else None},
        "hbm": {"achieved_gbs": acc_bytes_alg / acc_t / 1e9, "peak_gbs": hbm_peak,
                "frac": acc_bytes_alg / acc_t / 1e9 / hbm_peak, "algorithmic_bytes": acc_bytes_alg,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"},
        "traffic": None,
        "plan": {"c": info.c, "windows": info.windows, "buckets_per_window": info.buckets_per_window,
                 "entries_per_thread": info.entries_per_thread, "sorted_entries": info.sorted_entries},
        "kernel_ms_breakdown": {k: round(v, 4) for k, v in kern_ms.items()},
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic_k_accumulate.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            r = cpu_reference(None, 1)
            cpu = {"value": r["points_per_s_at_full_size"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample"]}
        except Exception as e:  # never lose the GPU line because the CPU leg failed
            cpu = {"value": None, "unit": "points/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    line = {
        "metric": "BLS12-381 G1 MSM points/sec at 2^%d scalars" % args.logn,
        "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * elapsed / args.steps, "latency_ms_single_msm": 1e3 * serial_elapsed / args.steps,
        "in_flight": NF, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "u32-limb integer (381-bit Fp, Montgomery)", "data": "synthetic",
        "config": {"workload": ("BLS12-381 G1 Pippenger MSM, 2^%d random terms (points k_i*G, uniform scalars) per GPU; "
                                "at N GPUs one MSM of N*2^%d terms" % (args.logn, args.logn)) if scaling == "weak" else
                               "BLS12-381 G1 Pippenger MSM, 2^%d terms in total split over the GPUs" % args.logn,
                   "terms": n_total, "terms_per_gpu": n_local, "parallelism": "term-sharded x%d, 1 all-gather of raw accumulators" % world,
                   "l2": "inputs+workspace (>=450 MB/GPU at N=2^20) exceed the 126 MB L2; no flush needed",
                   "pipelining": ("timed steps keep several MSMs in flight (see in_flight) on separate CUDA streams (nmsm_msm_submit/collect): the "
                                  "latency-bound tail of one overlaps the copy + wide kernels of the next; "
                                  "latency_ms_single_msm and the roofline block come from a serial pass")
                   + ("; at N>1 the all-gather + fold of step i-1 overlaps the shard reduction of step i" if world > 1 else "")},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": len(pts_b) + len(sc_b),
                "d2h_bytes_per_step": POINT_BYTES + 20, "steps": e2e_steps},
        "gpu_launches": info.launches * args.steps * world,  # kernels of the timed K steps (serial pass not counted)
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "fixed_base": fixed,
        "any_point": any_point,
    }
    _emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
