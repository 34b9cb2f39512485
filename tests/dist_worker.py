"""Worker of tests/test_gpu_dist.py (run under torch.distributed.run, one rank per GPU, NCCL): the sharded MSM with
bucket exchange (nmsm.dist.msm_sharded -> nmsm_msm_sharded) against the oracle on every rank."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "noble-curves_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import nmsm  # noqa: F401  (sets CUDA / NCCL environment defaults before torch creates a context)
import torch
import torch.distributed as dist

import helpers as H
import nmsm
from nmsm import dist as nd
from oracle import noble_ref as R


def shard_tensors(pb, sb, pbytes, lo, hi, dev):
    if hi == lo:
        return None, None
    tp = torch.frombuffer(bytearray(pb[lo * pbytes:hi * pbytes]), dtype=torch.uint8).to(dev)
    ts = torch.frombuffer(bytearray(sb[lo * 32:hi * 32]), dtype=torch.uint8).to(dev)
    return tp, ts


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nmsm.init(local)
    dist.init_process_group("nccl", device_id=dev)
    nd.init()
    ok = True
    # both forms of the sharded pipeline: bulk (one group: default) and one group per window (exchange overlapped)
    for groups in (0, 8):
      nmsm.set_window_groups(groups)
      ok = run_cases(rank, world, dev, groups) and ok
    nmsm.set_window_groups(0)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_WORKER_OK" if int(flag.item()) == 1 else "DIST_WORKER_FAILED", flush=True)
    dist.destroy_process_group()


def run_cases(rank, world, dev, groups):
    ok = True
    # 1. oracle-sized cases on every curve family; ragged shards incl. an empty one
    for name, n in (("bls12_381_G1", 301), ("secp256k1", 97), ("ed25519", 64), ("bls12_381_G2", 33), ("bn254_G1", 5), ("bls12_381_G1", 1)):
        P, pts, scalars, total = H.soak_inputs(name, n)
        exp = H.expected_tuple(name, R.pippenger(P, pts, scalars))
        pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
        cid = H.CURVE_IDS[name]
        pbytes = len(pb) // n
        for bounds in (nd.shard_bounds(n, world, rank), ((0, n) if rank == world - 1 else (0, 0))):
            lo, hi = bounds
            tp, ts = shard_tensors(pb, sb, pbytes, lo, hi, dev)
            torch.cuda.synchronize()
            out, inf = nd.msm_sharded(cid, tp, ts, hi - lo)
            got = (*H.unpack_point(name, out), inf)
            if got != exp:
                ok = False
                print(f"[rank {rank}] MISMATCH {name} n={n} bounds={bounds} groups={groups}", flush=True)
    # 2. 2^16 BLS12-381 G1 terms (points k_i*G made on the GPU), (sum k_i s_i)*G identity; degenerate: all scalars equal
    name, n = "bls12_381_G1", 1 << 16
    P = R.CURVES[name]
    order = P.Fn.ORDER
    rnd = random.Random(99)
    ks = [rnd.randrange(1, order) for _ in range(n)]
    pts_b, _ = nmsm.mul_batch_packed(4, H.point_bytes(name, P.BASE) * n, H.pack_scalars(ks), n, False)
    lo, hi = nd.shard_bounds(n, world, rank)
    for sc in ([rnd.randrange(order) for _ in range(n)], [0x1D3F5A7C9B2E4F60718293A4B5C6D7E8F9 % order] * n):
        exp = H.expected_tuple(name, H.expected_from_total(P, sum(k * s for k, s in zip(ks, sc)) % order))
        tp, ts = shard_tensors(pts_b, H.pack_scalars(sc), 96, lo, hi, dev)
        torch.cuda.synchronize()
        for cid in (4, 6):
            out, inf = nd.msm_sharded(cid, tp, ts, hi - lo)
            if (*H.unpack_point(name, out), inf) != exp:
                ok = False
                print(f"[rank {rank}] MISMATCH large cid={cid} groups={groups}", flush=True)
    # 3. an invalid scalar on the last rank is reported everywhere with its global index
    sc = [rnd.randrange(order) for _ in range(n)]
    sc[n - 5] = order
    tp, ts = shard_tensors(pts_b, H.pack_scalars(sc), 96, lo, hi, dev)
    torch.cuda.synchronize()
    try:
        nd.msm_sharded(4, tp, ts, hi - lo)
        ok = False
        print(f"[rank {rank}] invalid scalar not reported", flush=True)
    except nmsm.NmsmError as e:
        if f"invalid scalar at index {n - 5}" not in str(e):
            ok = False
            print(f"[rank {rank}] wrong error: {e}", flush=True)
    return ok


if __name__ == "__main__":
    main()
