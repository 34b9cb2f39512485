"""GPU parity tests added in round 2 (all through the C ABI):

  * drop-in defaults: nmsm.pippenger with DEFAULT arguments equals the oracle's pippenger on BLS12-381 G1 points
    outside the prime-order subgroup (the reference's pippenger is the plain group law, curve.ts:863-905; its Point
    constructor / fromAffine do not validate, weierstrass.ts:695-718)
  * the shard -> fold path of the multi-GPU split, emulated on one GPU (construction of
    test/slow-curves.test.ts:185-252)
  * window-group pipelining (engine.cuh submit_msm): every forced group count gives the oracle's result
  * mulAddUnsafe incl. allowOversized (curve.ts:820-836; test/point.test.ts:656-683)
  * strict / ZIP-215 Edwards decoding (edwards.ts:405-436), curve-equation check, Y = 0 rejection
"""
import ctypes
import random

import pytest

import helpers as H
from conftest import load_golden
from oracle import noble_ref as R

pytestmark = pytest.mark.gpu

ALL = ["secp256k1", "ed25519", "bn254_G1", "bn254_G2", "bls12_381_G1", "bls12_381_G2"]


@pytest.fixture(scope="module")
def nmsm():
    import nmsm as m

    m.init(0)
    return m


def as_tuple(p):
    return p.x, p.y, 1 if p.is0() else 0


# ------------------------------------------------------------------------------------------------
# 1. contract-safe defaults
# ------------------------------------------------------------------------------------------------
def test_pippenger_default_arguments_on_non_subgroup_points(nmsm):
    P = R.CURVES["bls12_381_G1"]
    C = nmsm.CURVES["bls12_381_G1"]
    rnd = random.Random(2024)
    bad = H.bls_g1_non_subgroup_points(30, seed=9)
    good = [P.BASE.multiplyUnsafe(rnd.randrange(1, P.Fn.ORDER)) for _ in range(30)]
    pts = R.normalizeZ(P, bad + good)
    rnd.shuffle(pts)
    scalars = [rnd.randrange(P.Fn.ORDER) for _ in pts]
    exp = H.expected_tuple("bls12_381_G1", R.pippenger(P, pts, scalars))
    cpts = [C.fromAffine(p.toAffine()) for p in pts]  # fromAffine: unvalidated handles, like the reference
    assert not any(p._valid for p in cpts)
    got = nmsm.pippenger(C, cpts, scalars)  # DEFAULT arguments
    assert as_tuple(got) == exp
    assert not got._valid
    # the same through the other reference-shaped entry points
    assert as_tuple(nmsm.mulAddUnsafe(C, cpts[:4], scalars[:4])) == H.expected_tuple(
        "bls12_381_G1", R.pippenger(P, pts[:4], scalars[:4]))
    run = nmsm.interleavedMSMUnsafe(C, cpts, 4)
    assert as_tuple(run(scalars)) == exp
    assert as_tuple(nmsm.aggregate_points(C, cpts)) == H.expected_tuple("bls12_381_G1", R.pippenger(P, pts, [1] * len(pts)))
    # add / double of non-subgroup handles
    assert as_tuple(cpts[0].add(cpts[1])) == H.expected_tuple("bls12_381_G1", pts[0].add(pts[1]))
    assert as_tuple(cpts[0].double()) == H.expected_tuple("bls12_381_G1", pts[0].double())
    # validated inputs take the endomorphism id and still agree; validity is inherited by results
    gpts = R.normalizeZ(P, good)
    vpts = [C.fromAffine(p.toAffine()) for p in gpts]
    for v in vpts:
        v.assertValidity()
    assert all(v._valid for v in vpts)
    r = nmsm.pippenger(C, vpts, scalars[: len(vpts)])
    assert as_tuple(r) == H.expected_tuple("bls12_381_G1", R.pippenger(P, gpts, scalars[: len(vpts)])) and r._valid
    assert C.BASE.multiply(5)._valid and C.BASE.add(C.BASE)._valid
    # assertValidity rejects exactly like the reference (weierstrass.ts:752-771)
    b0 = C.fromAffine(bad[0].toAffine())
    with pytest.raises(ValueError, match="not in prime-order subgroup"):
        b0.assertValidity()
    off = C.fromAffine({"x": 5, "y": 7})
    with pytest.raises(ValueError, match="equation left != right"):
        off.assertValidity()


def test_bls12_381_g2_psi_split_and_default_arguments(nmsm):
    """BLS12-381 G2: id 5 splits every term four ways along psi (valid on the prime-order subgroup), id 7 is the plain
    schedule.  Both against the oracle on subgroup points incl. scalars that isolate each psi power; twist points outside
    G2 through id 7 and through nmsm.pippenger with DEFAULT arguments (unvalidated handles -> id 7, validated -> id 5)."""
    name = "bls12_381_G2"
    P = R.CURVES[name]
    C = nmsm.CURVES[name]
    r = P.Fn.ORDER
    z = 0xD201000000010000
    half = z // 2
    rnd = random.Random(77)
    n = 300
    ks = [rnd.randrange(1, r) for _ in range(n)]
    pts_b, _ = nmsm.mul_batch_packed(5, H.point_bytes(name, P.BASE) * n, H.pack_scalars(ks), n, False)
    edge = [0, 1, z, z * z, z**3, r - 1, r - z, half, half + 1, (half + 1) * z**3 % r, z**3 + z * z + z + 1, z**3 * (z - 1)]
    scalars = edge + [rnd.randrange(r) for _ in range(n - len(edge))]
    total = sum(k * s for k, s in zip(ks, scalars)) % r
    exp = H.expected_tuple(name, P.BASE.multiplyUnsafe(total))
    for cid in (5, 7):
        for c in (0, 7, 13):
            nmsm.set_window_bits(c)
            try:
                out, inf = nmsm.msm_packed(cid, pts_b, H.pack_scalars(scalars), n)
            finally:
                nmsm.set_window_bits(0)
            x, y = H.unpack_point(name, out)
            assert (x, y, inf) == exp, (cid, c)
    # one term at a time: k * P for the scalars that use a single psi power
    one = pts_b[:192]
    p0 = P.BASE.multiplyUnsafe(ks[0])
    for k in edge[1:]:
        out, inf = nmsm.msm_packed(5, one, H.pack_scalars([k]), 1)
        assert (*H.unpack_point(name, out), inf) == H.expected_tuple(name, p0.multiplyUnsafe(k)), hex(k)
    # fixed-base table over the split set
    ps = nmsm.PointSet(5, pts_b, n)
    ps.precompute(0)
    out, inf = ps.msm(H.pack_scalars(scalars), n)
    assert (*H.unpack_point(name, out), inf) == exp
    # outside the subgroup
    bad = H.bls_g2_non_subgroup_points(6)
    good = [P.BASE.multiplyUnsafe(rnd.randrange(1, r)) for _ in range(6)]
    pts = R.normalizeZ(P, bad + good)
    rnd.shuffle(pts)
    sc = [rnd.randrange(r) for _ in pts]
    exp_bad = H.expected_tuple(name, R.pippenger(P, pts, sc))
    out, inf = nmsm.msm_packed(7, H.pack_points(name, pts), H.pack_scalars(sc), len(pts))
    assert (*H.unpack_point(name, out), inf) == exp_bad
    cpts = [C.fromAffine(p.toAffine()) for p in pts]
    assert not any(p._valid for p in cpts)
    got = nmsm.pippenger(C, cpts, sc)  # DEFAULT arguments
    assert as_tuple(got) == exp_bad and not got._valid
    assert as_tuple(nmsm.interleavedMSMUnsafe(C, cpts, 4)(sc)) == exp_bad
    gpts = R.normalizeZ(P, good)
    vpts = [C.fromAffine(p.toAffine()) for p in gpts]
    for v in vpts:
        v.assertValidity()
    res = nmsm.pippenger(C, vpts, sc[: len(vpts)])
    assert as_tuple(res) == H.expected_tuple(name, R.pippenger(P, gpts, sc[: len(vpts)])) and res._valid
    with pytest.raises(ValueError, match="not in prime-order subgroup"):
        C.fromAffine(bad[0].toAffine()).assertValidity()


@pytest.mark.parametrize("name", ALL)
def test_points_on_curve(nmsm, name):
    P, pts, scalars, _ = H.soak_inputs(name, 40)
    cid = H.CURVE_IDS[name]
    pb = bytearray(H.pack_points(name, pts))
    n = len(pts)
    cb = H.FP_BYTES[name] * H.PARTS[name]
    # corrupt y of point 3 (still in range), put an out-of-range x into point 5, make point 7 the identity encoding
    pb[3 * 2 * cb + cb] ^= 1
    pb[5 * 2 * cb:5 * 2 * cb + H.FP_BYTES[name]] = b"\xff" * H.FP_BYTES[name]
    ident = (bytes(cb) + (1).to_bytes(H.FP_BYTES[name], "little") + bytes(cb - H.FP_BYTES[name])) if name == "ed25519" else bytes(2 * cb)
    pb[7 * 2 * cb:8 * 2 * cb] = ident
    flags = nmsm.points_on_curve(cid, bytes(pb), n)
    assert list(flags) == [0 if i in (3, 5) else 1 for i in range(n)]


def test_constructor_rejects_y_zero_and_oversized_mul_add(nmsm):
    C = nmsm.CURVES["secp256k1"]
    with pytest.raises(ValueError, match="bad point coordinate y"):
        C(5, 0)
    assert C.ZERO.is0()
    for name in ("secp256k1", "bls12_381_G1", "ed25519"):
        P = R.CURVES[name]
        Cn = nmsm.CURVES[name]
        n = P.Fn.ORDER
        G = Cn.BASE
        G2 = G.double()
        # test/point.test.ts:656-683 'mulAddUnsafe: dense grid, edges, oversized, invalid inputs'
        for s1, s2 in ((0, 0), (1, 1), (5, 97), (n - 1, n - 1), (12345, 0)):
            want = H.expected_tuple(name, H.expected_from_total(P, (s1 + 2 * s2) % n))
            assert as_tuple(nmsm.mulAddUnsafe(Cn, [G, G2], [s1, s2])) == want, (name, s1, s2)
        assert nmsm.mulAddUnsafe(Cn, [], []).is0()
        assert as_tuple(nmsm.mulAddUnsafe(Cn, [G, Cn.ZERO], [3, 5])) == H.expected_tuple(name, P.BASE.multiplyUnsafe(3))
        so = n ** 3 + 12345
        assert as_tuple(nmsm.mulAddUnsafe(Cn, [G], [so], True)) == H.expected_tuple(name, H.expected_from_total(P, so % n))
        with pytest.raises(ValueError, match="invalid scalar at index 0"):
            nmsm.mulAddUnsafe(Cn, [G], [n])
        with pytest.raises(ValueError, match="invalid scalar at index 0"):
            nmsm.mulAddUnsafe(Cn, [G], [n ** 4], True)
        with pytest.raises(ValueError, match="invalid scalar at index 0"):
            nmsm.mulAddUnsafe(Cn, [G], [-1], True)
        with pytest.raises(ValueError, match="equal length"):
            nmsm.mulAddUnsafe(Cn, [G], [1, 2])
        with pytest.raises(ValueError, match="invalid point at index 0"):
            nmsm.mulAddUnsafe(Cn, [object()], [1])
    # oversized scalars are NOT reduced mod n: on a point outside the subgroup n*P != O (curve.ts:806-808)
    P = R.CURVES["bls12_381_G1"]
    Cn = nmsm.CURVES["bls12_381_G1"]
    bad = R.normalizeZ(P, H.bls_g1_non_subgroup_points(2, seed=5))
    n = P.Fn.ORDER
    b0 = Cn.fromAffine(bad[0].toAffine())
    nP = bad[0].multiplyUnsafe(n - 1).add(bad[0])
    assert not nP.is0()
    assert as_tuple(nmsm.mulAddUnsafe(Cn, [b0], [n], True)) == H.expected_tuple("bls12_381_G1", nP)
    big = n * n + 3 * n + 17
    ref = nP.multiplyUnsafe(n - 1).add(nP)  # n^2 * P
    ref = ref.add(nP.multiplyUnsafe(3)).add(bad[0].multiplyUnsafe(17))
    assert as_tuple(nmsm.mulAddUnsafe(Cn, [b0], [big], True)) == H.expected_tuple("bls12_381_G1", ref)


def test_ed25519_strict_and_zip215_decoding(nmsm):
    g = load_golden("ed25519.json")
    p = R.ED25519_CURVE["p"]
    encs = [bytes.fromhex(v["pk"]) for v in g["vectors"][:16]]
    encs += [bytes.fromhex(v["vk_bytes"]) for v in g["zip215"][:80]] + [bytes.fromhex(v["sig_bytes"])[:32] for v in g["zip215"][:80]]
    # hand-made non-canonical encodings: y = p (== 0), y = p + 1, x = 0 with the sign bit set (y = 1 and y = p - 1)
    encs += [p.to_bytes(32, "little"), (p + 1).to_bytes(32, "little"), ((1 << 255) | 1).to_bytes(32, "little"),
             ((1 << 255) | (p - 1)).to_bytes(32, "little"), (1).to_bytes(32, "little"), bytes([0xFF] * 32)]
    blob = b"".join(encs)
    C = nmsm.CURVES["ed25519"]
    for zip215 in (False, True):
        pts, st = nmsm.points_decode(1, blob, len(encs), zip215=zip215)
        n_ok = 0
        for i, e in enumerate(encs):
            try:
                a = R.ed25519_point_from_bytes(e, zip215).toAffine()
            except ValueError:
                assert st[i] == 0, (zip215, e.hex())
                assert pts[i * 64:(i + 1) * 64] == bytes(64)
                with pytest.raises(ValueError):
                    C.fromBytes(e, zip215)
                continue
            n_ok += 1
            assert st[i] == 1, (zip215, e.hex())
            assert H.unpack_point("ed25519", pts[i * 64:(i + 1) * 64]) == (a["x"] % p, a["y"] % p)
            q = C.fromBytes(e, zip215)
            assert (q.x, q.y) == (a["x"] % p, a["y"] % p)
        assert n_ok > 16
    # the two modes really differ on this list
    _, st_strict = nmsm.points_decode(1, blob, len(encs))
    _, st_zip = nmsm.points_decode(1, blob, len(encs), zip215=True)
    assert sum(st_zip) > sum(st_strict)


# ------------------------------------------------------------------------------------------------
# 2. shard -> fold (multi-GPU path emulated on one GPU)
# ------------------------------------------------------------------------------------------------
def _shard_fold(nmsm, name, pts_b, sc_b, n, cuts, use_slots):
    """Split [0, n) at `cuts`, reduce every shard to a raw accumulator on the GPU, fold them."""
    import torch

    from nmsm import _lib

    lib = _lib.load()
    cid = H.CURVE_IDS[name]
    pbytes = lib.nmsm_point_bytes(cid)
    ab = lib.nmsm_acc_bytes(cid)
    dev = torch.device("cuda", 0)
    bounds = [0] + list(cuts) + [n]
    G = len(bounds) - 1
    accs = torch.zeros(G * ab, dtype=torch.uint8, device=dev)
    keep = []
    for g in range(G):
        lo, hi = bounds[g], bounds[g + 1]
        m = hi - lo
        dp = torch.frombuffer(bytearray(pts_b[lo * pbytes:hi * pbytes] or bytes(16)), dtype=torch.uint8).to(dev)
        ds = torch.frombuffer(bytearray(sc_b[lo * 32:hi * 32] or bytes(16)), dtype=torch.uint8).to(dev)
        keep.append((dp, ds))
        torch.cuda.synchronize()
        out_ptr = accs.data_ptr() + g * ab
        if use_slots:
            slot = g % 4
            _lib.check(lib.nmsm_msm_submit_partial(cid, dp.data_ptr() if m else None, ds.data_ptr() if m else None, m, out_ptr, slot))
            _lib.check(lib.nmsm_msm_collect(slot, None, None))
        else:
            _lib.check(lib.nmsm_msm_partial_device(cid, dp.data_ptr() if m else None, ds.data_ptr() if m else None, m, out_ptr))
    out = ctypes.create_string_buffer(pbytes)
    inf = ctypes.c_int(0)
    _lib.check(lib.nmsm_fold_partials_device(cid, accs.data_ptr(), G, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
    x, y = H.unpack_point(name, out.raw)
    return x, y, inf.value


@pytest.mark.parametrize("name", ALL)
def test_shard_fold_path_matches_oracle(nmsm, name):
    n = 257
    P, pts, scalars, total = H.soak_inputs(name, n)
    # an all-ZERO shard (points 100..139 are the identity) and an empty shard (cut repeated)
    for i in range(100, 140):
        pts[i] = P.ZERO
    exp = H.expected_tuple(name, R.pippenger(P, pts, scalars))
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    for cuts in ([128], [100, 140], [7, 7, 30, 100, 140, 200, 256]):  # G = 2, 3, 8 (incl. an empty and an all-ZERO shard)
        for use_slots in (False, True):
            assert _shard_fold(nmsm, name, pb, sb, n, cuts, use_slots) == exp, (name, cuts, use_slots)


@pytest.mark.parametrize("name", ALL)
def test_normalize_accs_matches_oracle(nmsm, name):
    """nmsm_accs_normalize, the device form of normalizeZ (curve.ts:311-326): raw accumulators of 70 one-term MSMs (among
    them zero scalars and the identity point, which normalise to ZERO) -> canonical affine, from a device pointer and
    from a host copy; compared with the oracle's s_i * P_i."""
    import torch

    from nmsm import _lib

    lib = _lib.load()
    cid = H.CURVE_IDS[name]
    pbytes, ab = lib.nmsm_point_bytes(cid), lib.nmsm_acc_bytes(cid)
    n = 70 if "G2" not in name else 36
    P, pts, scalars, _ = H.soak_inputs(name, n)
    pts[5] = P.ZERO
    dev = torch.device("cuda", 0)
    accs = torch.zeros(n * ab, dtype=torch.uint8, device=dev)
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    keep = []
    for i in range(n):
        dp = torch.frombuffer(bytearray(pb[i * pbytes:(i + 1) * pbytes]), dtype=torch.uint8).to(dev)
        ds = torch.frombuffer(bytearray(sb[i * 32:(i + 1) * 32]), dtype=torch.uint8).to(dev)
        keep.append((dp, ds))
        torch.cuda.synchronize()
        _lib.check(lib.nmsm_msm_partial_device(cid, dp.data_ptr(), ds.data_ptr(), 1, accs.data_ptr() + i * ab))
    want = [H.expected_tuple(name, p.multiplyUnsafe(s) if s else P.ZERO) for p, s in zip(pts, scalars)]
    assert sum(w[2] for w in want) >= 4  # every 17th scalar is zero, point 5 is the identity
    for src, on_dev in ((accs.data_ptr(), True), (bytes(accs.cpu().numpy()), False)):
        out, infs = nmsm.normalize_accs(cid, src, n, on_device=on_dev)
        got = [(*H.unpack_point(name, out[i * pbytes:(i + 1) * pbytes]), infs[i]) for i in range(n)]
        assert got == want, (name, on_dev)
    assert nmsm.normalize_accs(cid, b"", 0) == (b"", b"")


def test_shard_fold_large_bls12_381_g1(nmsm):
    """2^17 terms in 8 uneven shards: sum s_i*(k_i*G) = (sum k_i s_i)*G (test/slow-curves.test.ts:204-233)."""
    name = "bls12_381_G1"
    P = R.CURVES[name]
    n = 1 << 17
    order = P.Fn.ORDER
    rnd = random.Random(5)
    ks = [rnd.randrange(1, order) for _ in range(n)]
    sc = [0 if i % 17 == 0 else rnd.randrange(order) for i in range(n)]
    pts_b, _ = nmsm.mul_batch_packed(4, H.point_bytes(name, P.BASE) * n, H.pack_scalars(ks), n, False)
    total = sum(k * s for k, s in zip(ks, sc)) % order
    exp = H.expected_tuple(name, H.expected_from_total(P, total))
    cuts = [5, 20000, 20001, 50000, 90000, 100000, 131071]
    assert _shard_fold(nmsm, name, pts_b, H.pack_scalars(sc), n, cuts, True) == exp


# ------------------------------------------------------------------------------------------------
# 3. window-group pipelining
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_window_groups_give_identical_results(nmsm, name):
    n = 3000
    P, pts, scalars, total = H.soak_inputs(name, n, seed_offset=3)
    exp = H.expected_tuple(name, H.expected_from_total(P, total))
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    cid = H.CURVE_IDS[name]
    try:
        for groups in (1, 2, 3, 5, 8):
            nmsm.set_window_groups(groups)
            out, inf = nmsm.msm_packed(cid, pb, sb, n)
            assert (*H.unpack_point(name, out), inf) == exp, (name, groups)
            _, info = nmsm.last_timing()
            assert 1 <= info.window_groups <= groups
        # degenerate inputs across groups: all scalars equal (one bucket per window), P + (-P), all-ZERO points
        nmsm.set_window_groups(8)
        s = (P.Fn.ORDER - 1) // 3
        ks_total = 0
        # points are (k0 + i*ks)*G: sum of them times s
        e2 = R.pippenger(P, pts[:400], [s] * 400)
        out, inf = nmsm.msm_packed(cid, H.pack_points(name, pts[:400]), H.pack_scalars([s] * 400), 400)
        assert (*H.unpack_point(name, out), inf) == H.expected_tuple(name, e2)
        pm = [pts[1], pts[1].negate(), P.ZERO, pts[2]]
        out, inf = nmsm.msm_packed(cid, H.pack_points(name, pm), H.pack_scalars([7, 7, 9, 0]), 4)
        assert inf == 1
    finally:
        nmsm.set_window_groups(0)


def test_window_groups_large_bls12_381_g1(nmsm):
    """2^18 terms: the default (one group) and forced group counts agree with (sum k_i s_i)*G; giant buckets (all
    scalars equal) cross the per-window segment ranges."""
    name = "bls12_381_G1"
    P = R.CURVES[name]
    n = 1 << 18
    order = P.Fn.ORDER
    rnd = random.Random(6)
    ks = [rnd.randrange(1, order) for _ in range(n)]
    sc = [rnd.randrange(order) for _ in range(n)]
    pts_b, _ = nmsm.mul_batch_packed(4, H.point_bytes(name, P.BASE) * n, H.pack_scalars(ks), n, False)
    sb = H.pack_scalars(sc)
    exp = H.expected_tuple(name, H.expected_from_total(P, sum(k * s for k, s in zip(ks, sc)) % order))
    try:
        for groups in (0, 1, 4, 8):
            nmsm.set_window_groups(groups)
            for cid in (4, 6):
                out, inf = nmsm.msm_packed(cid, pts_b, sb, n)
                assert (*H.unpack_point(name, out), inf) == exp, (groups, cid)
            _, info = nmsm.last_timing()
            assert info.window_groups == (groups if groups else 1)
        nmsm.set_window_groups(8)
        s = 0x1D3F5A7C9B2E4F60718293A4B5C6D7E8F9 % order
        exp2 = H.expected_tuple(name, H.expected_from_total(P, sum(ks) * s % order))
        out, inf = nmsm.msm_packed(4, pts_b, H.pack_scalars([s] * n), n)
        assert (*H.unpack_point(name, out), inf) == exp2
    finally:
        nmsm.set_window_groups(0)
