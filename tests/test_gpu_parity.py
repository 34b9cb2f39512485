"""GPU parity tests: the CUDA path (through the C ABI, via the `nmsm` host mirror) vs the CPU oracle.

Bit-exact on canonical affine coordinates (test/point.test.ts:36-44 comparison idiom).  Mirrors
test/point.test.ts:264-305,685-722,825-862, test/slow-curves.test.ts:128-152,185-252,
test/secp256k1.test.ts:59-76, test/bls12-381.test.ts:1463-1533, test/bn254.test.ts:859-887,
test/ed25519.test.ts:50-78, benchmark/msm_timings.ts:20-65.  At BASELINE.json's full sizes the oracle
is too slow, so the scalar-in-exponent identity  sum s_i*(k_i*G) = (sum k_i*s_i mod n)*G  is used
(test/slow-curves.test.ts:204-233).
"""
import hashlib
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


def gpu_msm(nmsm, name, pts_b, sc_b, n):
    out, inf = nmsm.msm_packed(H.CURVE_IDS[name], pts_b, sc_b, n)
    x, y = H.unpack_point(name, out)
    return x, y, inf


@pytest.mark.parametrize("name", ALL)
def test_msm_boundary_soak(nmsm, name):
    sizes = [1, 2, 31, 32, 33, 127, 128, 129, 511, 512, 513, 2047, 2048, 2049]
    if "G2" in name:
        sizes = [1, 31, 33, 129, 513]
    nmax = sizes[-1]
    P, pts, scalars, _ = H.soak_inputs(name, nmax)
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    step = len(pb) // nmax
    for size in sizes:
        exp = H.expected_tuple(name, R.pippenger(P, pts[:size], scalars[:size]))
        got = gpu_msm(nmsm, name, pb[: size * step], sb[: size * 32], size)
        assert got == exp, (name, size)


@pytest.mark.parametrize("name", ALL)
def test_msm_window_sweep(nmsm, name):
    """Every window size c and several segment shapes give the same point."""
    n = 300 if "G2" not in name else 64
    P, pts, scalars, total = H.soak_inputs(name, n, seed_offset=99)
    exp = H.expected_tuple(name, H.expected_from_total(P, total))
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    try:
        for c in ([2, 3, 5, 8, 11, 13, 16] if "G2" not in name else [3, 9, 14]):
            nmsm.set_window_bits(c)
            assert gpu_msm(nmsm, name, pb, sb, n) == exp, (name, c)
    finally:
        nmsm.set_window_bits(0)


@pytest.mark.parametrize("name", ALL)
def test_msm_basic_and_degenerate(nmsm, name):
    P = R.CURVES[name]
    G = P.BASE
    cases = [
        ([G], [0]),
        ([P.ZERO], [123]),
        ([G], [123]),
        ([G, G.double(), G.double().double(), G.double().double().double()], [3, 5, 7, 11]),
        ([G, G.negate()], [5, 5]),
        ([G, G], [P.Fn.ORDER - 1, 1]),
        ([G], [P.Fn.ORDER - 1]),
        ([P.ZERO, G, P.ZERO], [7, 9, 0]),
    ]
    for pts, sc in cases:
        pts = R.normalizeZ(P, pts)
        exp = H.expected_tuple(name, R.pippenger(P, pts, sc))
        assert gpu_msm(nmsm, name, H.pack_points(name, pts), H.pack_scalars(sc), len(pts)) == exp, (name, sc)
    # empty input -> identity (curve.ts:878)
    exp0 = H.expected_tuple(name, P.ZERO)
    assert gpu_msm(nmsm, name, b"", b"", 0) == exp0
    # 2048 copies of G, all scalars 2^10-1 (point.test.ts:842-853): equal-operand additions everywhere
    n = 2048
    s = 2**10 - 1
    exp = H.expected_tuple(name, G.multiply((n * s) % P.Fn.ORDER))
    assert gpu_msm(nmsm, name, H.point_bytes(name, G) * n, H.pack_scalars([s] * n), n) == exp
    # all scalars identical, distinct points (benchmark/msm_timings.ts:45-63): every term in one bucket per window
    P2, pts2, _, _ = H.soak_inputs(name, 200 if "G2" not in name else 40, zero_every=0)
    k = 0xDEADBEEFCAFEBABE1234567
    acc = P2.ZERO
    for p in pts2:
        acc = acc.add(p)
    exp = H.expected_tuple(name, acc.multiplyUnsafe(k))
    assert gpu_msm(nmsm, name, H.pack_points(name, pts2), H.pack_scalars([k] * len(pts2)), len(pts2)) == exp


def test_msm_validation_errors(nmsm):
    name = "bls12_381_G1"
    P, pts, scalars, _ = H.soak_inputs(name, 20)
    sc = list(scalars)
    sc[7] = P.Fn.ORDER
    with pytest.raises(ValueError, match="invalid scalar at index 7"):
        nmsm.msm_packed(H.CURVE_IDS[name], H.pack_points(name, pts), H.pack_scalars(sc), 20)
    pb = bytearray(H.pack_points(name, pts))
    pb[96 * 5: 96 * 5 + 48] = (P.Fp.ORDER).to_bytes(48, "little")
    with pytest.raises(ValueError, match="invalid point at index 5"):
        nmsm.msm_packed(H.CURVE_IDS[name], bytes(pb), H.pack_scalars(sc), 20)
    with pytest.raises(ValueError, match="equal length"):
        nmsm.msm_packed(H.CURVE_IDS[name], bytes(pb), H.pack_scalars(sc[:19]), 20)


@pytest.mark.parametrize("name", ALL)
def test_object_api_matches_reference_semantics(nmsm, name):
    """pippenger(c, points, scalars) / Point.multiply / multiplyUnsafe with the reference's error text."""
    C = nmsm.CURVES[name]
    P = R.CURVES[name]
    G = C.BASE
    assert nmsm.pippenger(C, [], []).equals(C.ZERO)
    assert nmsm.pippenger(C, [G], [0]).equals(C.ZERO)
    assert nmsm.pippenger(C, [C.ZERO], [123]).equals(C.ZERO)
    g123 = nmsm.pippenger(C, [G], [123])
    assert (g123.x, g123.y) == R.affine_tuple(P, P.BASE.multiply(123))
    assert g123.equals(G.multiply(123)) and g123.equals(G.multiplyUnsafe(123))
    pts = [G, G.double(), G.double().double(), G.double().double().double()]
    assert nmsm.pippenger(C, pts, [3, 5, 7, 11]).equals(G.multiply(129))
    assert G.add(G).equals(G.double()) and G.add(G.negate()).is0() and G.subtract(G).equals(C.ZERO)
    assert (G.double().x, G.double().y) == R.affine_tuple(P, P.BASE.double())
    with pytest.raises(ValueError, match="invalid scalar at index 0"):
        nmsm.pippenger(C, [G], [C.Fn.ORDER])
    with pytest.raises(ValueError, match="arrays of points and scalars must have equal length"):
        nmsm.pippenger(C, [G], [1, 2])
    with pytest.raises(ValueError, match="invalid point at index 1"):
        nmsm.pippenger(C, [G, 5], [1, 2])
    with pytest.raises(ValueError, match="array of scalars expected"):
        nmsm.pippenger(C, [G], 5)
    with pytest.raises(ValueError, match="invalid scalar"):
        G.multiply(0)
    assert G.multiplyUnsafe(0).equals(C.ZERO)
    with pytest.raises(ValueError, match="invalid scalar"):
        G.multiplyUnsafe(C.Fn.ORDER)
    # (N-1)*G + G = O  (point.test.ts:69-205 group laws)
    assert G.multiply(C.Fn.ORDER - 1).add(G).is0()
    a, b = 0x1234567890ABCDEF1234, 0xFEDCBA09876543211
    assert G.multiply(a).multiply(b).equals(G.multiply(a * b % C.Fn.ORDER))


@pytest.mark.parametrize("form", ["quad", "serial"])
@pytest.mark.parametrize("name", ALL)
def test_mul_batch_vs_oracle(nmsm, name, form, monkeypatch):
    """Both forms of k_mul_batch: one item per quad of lanes (small batches, the default here) and one item per thread."""
    if form == "serial":
        monkeypatch.setenv("NMSM_MUL_QUAD_MAX", "0")
    P = R.CURVES[name]
    n_order = P.Fn.ORDER
    rng = R.Xorshift64(0xDEADBEEF)
    count = 40 if "G2" not in name else 10
    scalars = [1, 2, 3, n_order - 1, n_order - 2, 2**128 - 1, 2**128, 2**64 + 1, (1 << 200) - 1]
    scalars += [rng.rndBelow(n_order - 1) + 1 for _ in range(count)]
    scalars = [s % n_order or 1 for s in scalars]
    ks = [rng.rndBelow(n_order - 1) + 1 for _ in scalars]
    pts = R.normalizeZ(P, [P.BASE.multiplyUnsafe(k) for k in ks[:-2]] + [P.BASE, P.ZERO])
    out, infs = nmsm.mul_batch_packed(H.CURVE_IDS[name], H.pack_points(name, pts), H.pack_scalars(scalars),
                                      len(scalars), False)
    pb = len(out) // len(scalars)
    for i, (p, s) in enumerate(zip(pts, scalars)):
        x, y = H.unpack_point(name, out[i * pb:(i + 1) * pb])
        assert (x, y, infs[i]) == H.expected_tuple(name, p.multiplyUnsafe(s)), (name, i)


def test_golden_secp256k1_privates(nmsm):
    """test/secp256k1.test.ts:59-76 through the GPU: k*G for the 45 vectors of privates-2.txt."""
    g = load_golden("secp256k1.json")["privates2"]
    P = R.CURVES["secp256k1"]
    n = len(g)
    out, infs = nmsm.mul_batch_packed(0, H.point_bytes("secp256k1", P.BASE) * n, H.pack_scalars([int(k) for k, _, _ in g]),
                                      n, False)
    for i, (_, x, y) in enumerate(g):
        assert H.unpack_point("secp256k1", out[i * 64:(i + 1) * 64]) == (int(x, 16), int(y, 16))
    for v in load_golden("secp256k1.json")["endomorphism"]:
        a = P.fromAffine({"x": int(v["ax"]), "y": int(v["ay"])})
        out, _ = nmsm.mul_batch_packed(0, H.point_bytes("secp256k1", a), H.pack_scalars([int(v["scalar"])]), 1, True)
        assert H.unpack_point("secp256k1", out) == (int(v["cx"]), int(v["cy"]))


def test_golden_bls12_381_multiples(nmsm):
    """test/bls12-381.test.ts:1463-1533 through the GPU: i*G (G1: i<1000, G2: i<256) vs the zkcrypto tables."""
    g = load_golden("bls12_381.json")
    G1, G2 = R.CURVES["bls12_381_G1"], R.CURVES["bls12_381_G2"]
    n = 999
    out, infs = nmsm.mul_batch_packed(4, H.point_bytes("bls12_381_G1", G1.BASE) * n, H.pack_scalars(range(1, n + 1)), n, False)
    for i in range(1, n + 1):
        b = bytearray(bytes.fromhex(g["G1_Uncompressed"][i]))
        b[0] &= 0x1F
        exp = (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big"))
        assert H.unpack_point("bls12_381_G1", out[(i - 1) * 96: i * 96]) == exp, i
    n = 255
    out, infs = nmsm.mul_batch_packed(5, H.point_bytes("bls12_381_G2", G2.BASE) * n, H.pack_scalars(range(1, n + 1)), n, False)
    for i in range(1, n + 1):
        b = bytearray(bytes.fromhex(g["G2_Uncompressed"][i]))
        b[0] &= 0x1F
        x1, x0, y1, y0 = (int.from_bytes(b[j * 48:(j + 1) * 48], "big") for j in range(4))
        assert H.unpack_point("bls12_381_G2", out[(i - 1) * 192: i * 192]) == ((x0, x1), (y0, y1)), i
    # MSM of the whole table against scalars: sum_i s_i*(i*G) = (sum i*s_i)*G
    rng = random.Random(7)
    sc = [rng.randrange(G1.Fn.ORDER) for _ in range(999)]
    def g1_le(h):
        b = bytearray(bytes.fromhex(h))
        b[0] &= 0x1F
        return bytes(b[:48][::-1]) + bytes(b[48:][::-1])

    pts_b = b"".join(g1_le(g["G1_Uncompressed"][i]) for i in range(1, 1000))
    tot = sum(i * s for i, s in zip(range(1, 1000), sc)) % G1.Fn.ORDER
    assert gpu_msm(nmsm, "bls12_381_G1", pts_b, H.pack_scalars(sc), 999) == H.expected_tuple(
        "bls12_381_G1", G1.BASE.multiplyUnsafe(tot))


def test_golden_bn254_and_ed25519(nmsm):
    g = load_golden("bn254.json")
    BN = R.CURVES["bn254_G1"]
    for t in g["seda_mul"]:
        s = int(t["scalar"], 16) % BN.Fn.ORDER
        if s == 0:
            continue
        pt = int(t["x"], 16).to_bytes(32, "little") + int(t["y"], 16).to_bytes(32, "little")
        out, _ = nmsm.mul_batch_packed(2, pt, H.pack_scalars([s]), 1, False)
        assert H.unpack_point("bn254_G1", out) == (int(t["result"][:64], 16), int(t["result"][64:], 16))
    for t in g["seda_add"]:
        pts = b"".join(int(t[k], 16).to_bytes(32, "little") for k in ("x1", "y1", "x2", "y2"))
        x, y, _ = gpu_msm(nmsm, "bn254_G1", pts, H.pack_scalars([1, 1]), 2)
        assert (x, y) == (int(t["result"][:64], 16), int(t["result"][64:], 16))
    # RFC 8032 public keys: pk = compress(clamp(sha512(sk)) * B)   (test/ed25519.test.ts:50-78)
    ED = R.CURVES["ed25519"]
    vec = load_golden("ed25519.json")["vectors"]
    scalars = []
    for v in vec:
        h = bytearray(hashlib.sha512(bytes.fromhex(v["sk"])).digest()[:32])
        h[0] &= 248
        h[31] &= 127
        h[31] |= 64
        scalars.append(int.from_bytes(bytes(h), "little") % ED.Fn.ORDER)
    n = len(vec)
    out, _ = nmsm.mul_batch_packed(1, H.point_bytes("ed25519", ED.BASE) * n, H.pack_scalars(scalars), n, False)
    for i, v in enumerate(vec):
        x, y = H.unpack_point("ed25519", out[i * 64:(i + 1) * 64])
        enc = bytearray(y.to_bytes(32, "little"))
        if x & 1:
            enc[31] |= 0x80
        assert bytes(enc).hex() == v["pk"], i


# ------------------------------------------------------------------------------------------------
# BASELINE.json sizes: scalar-in-exponent identity, points generated on the GPU as k_i*G
# ------------------------------------------------------------------------------------------------
def _large_case(nmsm, name, n, seed, zero_every=0):
    P = R.CURVES[name]
    order = P.Fn.ORDER
    rnd = random.Random(seed)
    ks = [rnd.randrange(1, order) for _ in range(n)]
    sc = [0 if (zero_every and i % zero_every == 0) else rnd.randrange(order) for i in range(n)]
    cid = H.CURVE_IDS[name]
    pts_b, infs = nmsm.mul_batch_packed(cid, H.point_bytes(name, P.BASE) * n, H.pack_scalars(ks), n, False)
    assert not any(infs)
    total = sum(k * s for k, s in zip(ks, sc)) % order
    exp = H.expected_tuple(name, H.expected_from_total(P, total))
    got = gpu_msm(nmsm, name, pts_b, H.pack_scalars(sc), n)
    assert got == exp, (name, n)
    return pts_b, sc, exp


def test_config_bls12_381_g1_2p16(nmsm):
    """configs[1]: BLS12-381 G1, 2^16 terms; also cross-checked against the reference algorithm's C port."""
    _large_case(nmsm, "bls12_381_G1", 1 << 16, 1, zero_every=17)


def test_config_bls12_381_g1_2p20_headline(nmsm):
    """The headline metric's workload: BLS12-381 G1, 2^20 terms."""
    pts_b, sc, exp = _large_case(nmsm, "bls12_381_G1", 1 << 20, 2)
    # linearity: MSM(first half) + MSM(second half) = MSM(all)
    h = 1 << 19
    a = nmsm.msm_packed(4, pts_b[: h * 96], H.pack_scalars(sc[:h]), h)
    b = nmsm.msm_packed(4, pts_b[h * 96:], H.pack_scalars(sc[h:]), h)
    s = nmsm.msm_packed(4, a[0] + b[0], H.pack_scalars([1, 1]), 2)
    x, y = H.unpack_point("bls12_381_G1", s[0])
    assert (x, y, s[1]) == exp


def test_config_bn254_g1_2p20(nmsm):
    _large_case(nmsm, "bn254_G1", 1 << 20, 3)


def test_config_bls12_381_g2_2p18(nmsm):
    _large_case(nmsm, "bls12_381_G2", 1 << 18, 4)


def test_config_ed25519_131073_terms(nmsm):
    """configs[4] core: Edwards MSM with 2*2^16+1 terms (the batch-verify equation's shape)."""
    _large_case(nmsm, "ed25519", 2 * (1 << 16) + 1, 5)


@pytest.mark.parametrize("name,n", [("bls12_381_G1", 100003), ("bn254_G1", 77777), ("secp256k1", 54321),
                                    ("bls12_381_G1_any", 40001), ("bn254_G2", 9973), ("ed25519", 65537)])
def test_msm_ragged_sizes(nmsm, name, n):
    """Sizes that are not powers of two (odd tails in every kernel's grid), every 13th scalar zero."""
    if name == "bls12_381_G1_any":
        P = R.CURVES["bls12_381_G1"]
        rnd = random.Random(n)
        ks = [rnd.randrange(1, P.Fn.ORDER) for _ in range(n)]
        sc = [0 if i % 13 == 0 else rnd.randrange(P.Fn.ORDER) for i in range(n)]
        pts_b, _ = nmsm.mul_batch_packed(4, H.point_bytes("bls12_381_G1", P.BASE) * n, H.pack_scalars(ks), n, False)
        total = sum(k * s for k, s in zip(ks, sc)) % P.Fn.ORDER
        out, inf = nmsm.msm_packed(6, pts_b, H.pack_scalars(sc), n)
        assert (*H.unpack_point("bls12_381_G1", out), inf) == H.expected_tuple("bls12_381_G1", H.expected_from_total(P, total))
    else:
        _large_case(nmsm, name, n, 1000 + n, zero_every=13)


def test_config_secp256k1_multiply_1024(nmsm):
    """configs[0] GPU counterpart: Point.multiply on 1024 random scalars of a non-base point vs the oracle."""
    P = R.CURVES["secp256k1"]
    rnd = random.Random(11)
    base = R.normalizeZ(P, [P.BASE.multiplyUnsafe(rnd.randrange(1, P.Fn.ORDER))])[0]
    ks = [rnd.randrange(1, P.Fn.ORDER) for _ in range(1024)]
    out, infs = nmsm.mul_batch_packed(0, H.point_bytes("secp256k1", base) * 1024, H.pack_scalars(ks), 1024, False)
    for i in range(0, 1024, 16):  # oracle spot-check on 64 of them (fixedWindowCT path, curve.ts:707-729)
        assert H.unpack_point("secp256k1", out[i * 64:(i + 1) * 64]) == R.affine_tuple(P, base.multiply(ks[i]))
    # all 1024: sum_i k_i*P == (sum k_i)*P via the MSM path
    tot = sum(ks) % P.Fn.ORDER
    x, y, inf = gpu_msm(nmsm, "secp256k1", out, H.pack_scalars([1] * 1024), 1024)
    assert (x, y) == R.affine_tuple(P, base.multiplyUnsafe(tot))


@pytest.mark.parametrize("name", ["secp256k1", "ed25519", "bls12_381_G1"])
def test_fixed_point_set_handle(nmsm, name):
    """interleavedMSMUnsafe (curve.ts:937-959) over a device-resident point set: several scalar vectors,
    fewer scalars than points, too many scalars, invalid point at upload."""
    C = nmsm.CURVES[name]
    P, pts, scalars, _ = H.soak_inputs(name, 129)
    cpts = [C.fromAffine(p.toAffine()) for p in pts]
    msm = nmsm.interleavedMSMUnsafe(C, cpts, 5)
    for size in (129, 33, 1):
        exp = H.expected_tuple(name, R.pippenger(P, pts[:size], scalars[:size]))
        got = msm(scalars[:size])
        assert (got.x, got.y, 1 if got.is0() else 0) == exp, (name, size)
    rnd = random.Random(3)
    sc2 = [rnd.randrange(P.Fn.ORDER) for _ in range(129)]
    exp = H.expected_tuple(name, R.pippenger(P, pts, sc2))
    got = msm(sc2)
    assert (got.x, got.y, 1 if got.is0() else 0) == exp
    with pytest.raises(ValueError, match="must not be larger"):
        msm(sc2 + [1])
    pb = bytearray(H.pack_points(name, pts))
    fb = H.FP_BYTES[name]
    pb[3 * 2 * fb: 3 * 2 * fb + fb] = P.Fp.ORDER.to_bytes(fb, "little")
    with pytest.raises(ValueError, match="invalid point at index 3"):
        nmsm.PointSet(H.CURVE_IDS[name], bytes(pb), 129)


@pytest.mark.parametrize("name", ALL)
def test_fixed_base_table_small(nmsm, name):
    """nmsm_points_precompute: the table route must give the plain route's answer for every table
    window size, including c > 16 (single window of up to 2^21 buckets, folded by k_reduce2 twice), with identity
    points, cancelling pairs and n-1 scalars in the set."""
    n = 200 if "G2" not in name else 60
    P, pts, scalars, _ = H.soak_inputs(name, n, seed_offset=9)
    pts[3] = P.ZERO
    pts[8] = pts[7].negate()
    scalars[8] = scalars[7]
    scalars[9] = P.Fn.ORDER - 1
    pts = R.normalizeZ(P, pts)
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    exp = H.expected_tuple(name, R.pippenger(P, pts, scalars))
    cid = H.CURVE_IDS[name]
    for c in (0, 4, 11, 16, 19, 22):
        ps = nmsm.PointSet(cid, pb, n)
        got_c, levels = ps.precompute(c)
        assert (c == 0 or got_c <= c) and levels >= 1  # c bounds the digit width; widths are balanced
        for _ in range(2):  # the table is reused
            out, inf = ps.msm(sb, n)
            assert (*H.unpack_point(name, out), inf) == exp, (name, c)
        zero = ps.msm(H.pack_scalars([0] * n), n)
        assert zero[1] == 1
        with pytest.raises(Exception, match="already carries a table"):
            ps.precompute(c)
        ps.close()
    ps = nmsm.PointSet(cid, pb, n)
    with pytest.raises(Exception, match="table window bits"):
        ps.precompute(3)
    ps.close()


@pytest.mark.parametrize("name", ALL)
def test_point_table_multiply(nmsm, name):
    """nmsm_point_table_* (Point.precompute + cached multiply, curve.ts:532-606): table route equals
    the oracle's multiply and the generic nmsm_mul_batch for edge and random scalars; range errors as multiply's."""
    P = R.CURVES[name]
    C = nmsm.CURVES[name]
    order = P.Fn.ORDER
    rnd = random.Random(77)
    base_o = R.normalizeZ(P, [P.BASE.multiplyUnsafe(rnd.randrange(1, order))])[0]
    n = 300
    ks = [1, 2, 32767, 32768, 32769, 65535, 65536, order - 1, order - 2, 2**128 - 1, 2**128,
          (1 << (order.bit_length() - 1)) - 1, 0x5555555555555555 << 60]
    ks = [k % order or 1 for k in ks] + [rnd.randrange(1, order) for _ in range(n - len(ks))]
    tbl = nmsm.PointTable(H.CURVE_IDS[name], H.point_bytes(name, base_o))
    out, infs = tbl.mul_batch(H.pack_scalars(ks), n, False)
    ref, rinfs = nmsm.mul_batch_packed(H.CURVE_IDS[name], H.point_bytes(name, base_o) * n, H.pack_scalars(ks), n, False)
    assert out == ref and infs == rinfs
    pb = len(out) // n
    for i in list(range(13)) + [50, 299]:
        assert (*H.unpack_point(name, out[i * pb:(i + 1) * pb]), infs[i]) == H.expected_tuple(name, base_o.multiply(ks[i]))
    with pytest.raises(ValueError, match="invalid scalar"):
        tbl.mul_batch(H.pack_scalars([5, 0]), 2, False)
    out0, inf0 = tbl.mul_batch(H.pack_scalars([5, 0]), 2, True)
    assert inf0[1] == 1
    with pytest.raises(ValueError, match="invalid scalar"):
        tbl.mul_batch(H.pack_scalars([order]), 1, True)
    tbl.close()
    # object API: precompute() marks the point, multiply_many / multiply then use the table
    bp = C.fromAffine(base_o.toAffine()).precompute(8, False)
    got = nmsm.multiply_many(C, [bp] * 5, ks[7:12])
    for g, k in zip(got, ks[7:12]):
        assert (g.x, g.y, 1 if g.is0() else 0) == H.expected_tuple(name, base_o.multiply(k))
    one = bp.multiply(ks[20])
    assert (one.x, one.y) == H.expected_tuple(name, base_o.multiply(ks[20]))[:2]
    with pytest.raises(ValueError, match="invalid point"):
        bad = bytearray(H.point_bytes(name, base_o))
        fb = H.FP_BYTES[name]
        bad[:fb] = b"\xff" * fb  # first base-field component >= p
        nmsm.PointTable(H.CURVE_IDS[name], bytes(bad))


def test_point_table_getpublickey_batch(nmsm):
    """BASE.multiply at rate (getPublicKey, weierstrass.ts:1168 / ed25519 RFC 8032 vectors): 2^16 secp256k1 keys through
    the table against the generic batch; the 128 RFC 8032 public keys from their clamped secret scalars."""
    P = R.CURVES["secp256k1"]
    n = 1 << 16
    rnd = random.Random(5)
    ks = [rnd.randrange(1, P.Fn.ORDER) for _ in range(n)]
    sb = H.pack_scalars(ks)
    gb = H.point_bytes("secp256k1", P.BASE)
    tbl = nmsm.PointTable(0, gb)
    out, infs = tbl.mul_batch(sb, n, False)
    ref, _ = nmsm.mul_batch_packed(0, gb * n, sb, n, False)
    assert out == ref and not any(infs)
    for i in (0, 1, 777, n - 1):
        assert H.unpack_point("secp256k1", out[i * 64:(i + 1) * 64]) == R.affine_tuple(P, P.BASE.multiply(ks[i]))
    tbl.close()
    ED = R.CURVES["ed25519"]
    vec = load_golden("ed25519.json")["vectors"]
    scalars = []
    for v in vec:
        h = bytearray(hashlib.sha512(bytes.fromhex(v["sk"])).digest()[:32])
        h[0] &= 248
        h[31] &= 127
        h[31] |= 64
        scalars.append(int.from_bytes(bytes(h), "little") % ED.Fn.ORDER)
    tbl = nmsm.PointTable(1, H.point_bytes("ed25519", ED.BASE))
    out, _ = tbl.mul_batch(H.pack_scalars(scalars), len(vec), False)
    for i, v in enumerate(vec):
        x, y = H.unpack_point("ed25519", out[i * 64:(i + 1) * 64])
        enc = bytearray(y.to_bytes(32, "little"))
        if x & 1:
            enc[31] |= 0x80
        assert bytes(enc).hex() == v["pk"], i
    tbl.close()


@pytest.mark.parametrize("name,logn", [("bls12_381_G1", 20), ("bls12_381_G2", 16), ("ed25519", 17)])
def test_fixed_base_table_large(nmsm, name, logn):
    """Fixed-base table at BASELINE sizes: same result as (sum k_i s_i) * G and as the plain MSM."""
    n = 1 << logn
    pts_b, sc, exp = _large_case(nmsm, name, n, 40 + logn)
    ps = nmsm.PointSet(H.CURVE_IDS[name], pts_b, n)
    c, levels = ps.precompute(0)
    assert 8 <= c <= 22
    sb = H.pack_scalars(sc)
    out, inf = ps.msm(sb, n)
    assert (*H.unpack_point(name, out), inf) == exp
    ms, info = nmsm.last_timing()
    assert info.windows == 1 and info.c == c
    ps.close()


# ------------------------------------------------------------------------------------------------
# ed25519 batch verification: batch accepts <=> every individual reference verify accepts
# ------------------------------------------------------------------------------------------------
def test_ed25519_batch_verify_small(nmsm):
    g = load_golden("ed25519.json")
    vec = g["vectors"]
    sigs = [bytes.fromhex(v["sig"]) for v in vec]
    msgs = [bytes.fromhex(v["msg"]) for v in vec]
    pks = [bytes.fromhex(v["pk"]) for v in vec]
    assert all(R.ed25519_verify(s, m, p) for s, m, p in zip(sigs[:16], msgs[:16], pks[:16]))
    assert nmsm.ed25519_verify_batch(sigs, msgs, pks) == (True, -1)
    assert nmsm.ed25519_verify_batch([], [], []) == (True, -1)
    assert nmsm.ed25519_verify_batch(sigs[:1], msgs[:1], pks[:1]) == (True, -1)
    for idx, byte in ((5, 40), (77, 3), (127, 63)):
        bad = list(sigs)
        b = bytearray(bad[idx])
        b[byte] ^= 1
        bad[idx] = bytes(b)
        indiv = R.ed25519_verify(bad[idx], msgs[idx], pks[idx])
        ok, _ = nmsm.ed25519_verify_batch(bad, msgs, pks)
        assert ok == indiv is False or ok == indiv
    m2 = list(msgs)
    m2[9] = m2[9] + b"\x00"
    assert nmsm.ed25519_verify_batch(sigs, m2, pks)[0] is False
    p2 = list(pks)
    p2[3], p2[4] = p2[4], p2[3]
    assert nmsm.ed25519_verify_batch(sigs, msgs, p2)[0] is False
    b = bytearray(sigs[3])
    b[63] |= 0xF0  # s >= l
    bad = list(sigs)
    bad[3] = bytes(b)
    assert nmsm.ed25519_verify_batch(bad, msgs, pks) == (False, 3)


def test_ed25519_batch_verify_zip215(nmsm):
    """The 196 ZIP-215 vectors (test/ed25519.test.ts:392-405): each one alone must match the reference's verdict,
    and all accepted ones together must pass as one batch (small-order and non-canonical encodings)."""
    zs = load_golden("ed25519.json")["zip215"]
    for v in zs:
        sig, pk = bytes.fromhex(v["sig_bytes"]), bytes.fromhex(v["vk_bytes"])
        assert R.ed25519_verify(sig, b"Zcash", pk) == v["valid_zip215"]
        assert nmsm.ed25519_verify_batch([sig], [b"Zcash"], [pk])[0] == v["valid_zip215"], v
    valid = [v for v in zs if v["valid_zip215"]]
    assert nmsm.ed25519_verify_batch([bytes.fromhex(v["sig_bytes"]) for v in valid], [b"Zcash"] * len(valid),
                                     [bytes.fromhex(v["vk_bytes"]) for v in valid])[0] is True


def test_config_ed25519_batch_verify_2p16(nmsm):
    """configs[4]: 2^16 signatures (the RFC 8032 vectors tiled), one Edwards MSM of 2*2^16+1 terms."""
    import time

    vec = load_golden("ed25519.json")["vectors"]
    reps = (1 << 16) // len(vec)
    sigs = [bytes.fromhex(v["sig"]) for v in vec] * reps
    msgs = [bytes.fromhex(v["msg"]) for v in vec] * reps
    pks = [bytes.fromhex(v["pk"]) for v in vec] * reps
    assert len(sigs) == 1 << 16
    t0 = time.perf_counter()
    assert nmsm.ed25519_verify_batch(sigs, msgs, pks) == (True, -1)
    dt = time.perf_counter() - t0
    print("ed25519 batch verify 2^16 signatures: %.1f ms end-to-end through the Python binding" % (dt * 1e3))
    bad = list(sigs)
    b = bytearray(bad[40000])
    b[33] ^= 0x10
    bad[40000] = bytes(b)
    assert nmsm.ed25519_verify_batch(bad, msgs, pks)[0] is False


@pytest.mark.parametrize("name", ["bls12_381_G1", "bls12_381_G2", "ed25519"])
def test_aggregate_points_and_small_helpers(nmsm, name):
    """f3: aggregation = sum of points (abstract/bls.ts:860,870), incl. 4096 terms in one bucket."""
    C = nmsm.CURVES[name]
    P = R.CURVES[name]
    n = 4096 if name == "bls12_381_G1" else 300
    _, pts, _, _ = H.soak_inputs(name, 40, zero_every=0)
    reps = [pts[i % 40] for i in range(n)]
    acc = P.ZERO
    for p in pts:
        acc = acc.add(p)
    # sum of n points cycling through 40 distinct ones = (n // 40) * S + partial
    exp = P.ZERO
    full, rem = divmod(n, 40)
    if full:
        exp = acc.multiplyUnsafe(full) if full > 1 else acc
    for p in pts[:rem]:
        exp = exp.add(p)
    cpts = [C.fromAffine(p.toAffine()) for p in pts]
    got = nmsm.aggregate_points(C, [cpts[i % 40] for i in range(n)])
    assert (got.x, got.y, 1 if got.is0() else 0) == H.expected_tuple(name, exp)
    assert nmsm.aggregate_points(C, []).equals(C.ZERO)
    assert nmsm.normalizeZ(C, cpts[:3])[2].equals(cpts[2])
    assert cpts[0].precompute(6) is cpts[0]
    if name == "ed25519":
        assert cpts[0].clearCofactor().equals(cpts[0].multiplyUnsafe(8)) and not cpts[0].isSmallOrder()


def test_two_msms_in_flight(nmsm):
    """nmsm_msm_submit / nmsm_msm_collect: two slots, different curves and sizes interleaved, errors on collect."""
    import ctypes

    from nmsm import _lib

    lib = _lib.load()
    cases = []
    for name, n, seed in (("bls12_381_G1", 700, 1), ("ed25519", 333, 2), ("bls12_381_G1", 64, 3), ("secp256k1", 1000, 4)):
        P, pts, scalars, total = H.soak_inputs(name, n, seed_offset=seed)
        cases.append((name, n, H.pack_points(name, pts), H.pack_scalars(scalars), H.expected_tuple(name, H.expected_from_total(P, total))))
    keep = {}

    def submit(i, slot):
        name, n, pb, sb, _ = cases[i]
        keep[slot] = (ctypes.create_string_buffer(pb, len(pb)), ctypes.create_string_buffer(sb, len(sb)))
        _lib.check(lib.nmsm_msm_submit(H.CURVE_IDS[name], ctypes.cast(keep[slot][0], ctypes.c_void_p),
                                       ctypes.cast(keep[slot][1], ctypes.c_void_p), n, 0, slot))

    def collect(i, slot):
        name = cases[i][0]
        out = ctypes.create_string_buffer(lib.nmsm_point_bytes(H.CURVE_IDS[name]))
        inf = ctypes.c_int(0)
        _lib.check(lib.nmsm_msm_collect(slot, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))
        x, y = H.unpack_point(name, out.raw)
        assert (x, y, inf.value) == cases[i][4], (i, name)

    submit(0, 0)
    submit(1, 1)
    collect(0, 0)
    submit(2, 0)
    collect(1, 1)
    submit(3, 1)
    collect(2, 0)
    collect(3, 1)
    # busy slot / empty slot / invalid scalar reported by collect
    submit(0, 0)
    with pytest.raises(_lib.NmsmError):
        submit(1, 0)
    collect(0, 0)
    with pytest.raises(_lib.NmsmError):
        collect(0, 0)
    name, n, pb, sb, _ = cases[2]
    bad = bytearray(sb)
    bad[5 * 32:6 * 32] = R.CURVES[name].Fn.ORDER.to_bytes(32, "little")
    keep[0] = (ctypes.create_string_buffer(pb, len(pb)), ctypes.create_string_buffer(bytes(bad), len(bad)))
    _lib.check(lib.nmsm_msm_submit(H.CURVE_IDS[name], ctypes.cast(keep[0][0], ctypes.c_void_p),
                                   ctypes.cast(keep[0][1], ctypes.c_void_p), n, 0, 0))
    out = ctypes.create_string_buffer(96)
    inf = ctypes.c_int(0)
    with pytest.raises(_lib.NmsmError, match="invalid scalar at index 5"):
        _lib.check(lib.nmsm_msm_collect(0, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf)))


@pytest.mark.parametrize("name", ["bls12_381_G1", "bn254_G1"])
def test_msm_giant_buckets_large(nmsm, name):
    """All scalars equal at N = 2^17 (benchmark/msm_timings.ts:45-63 shape): every window has ONE bucket holding
    every term, i.e. thousands of accumulate segments per bucket -> both tile-sum levels of the stitching."""
    n = 1 << 17
    P = R.CURVES[name]
    order = P.Fn.ORDER
    rnd = random.Random(77)
    ks = [rnd.randrange(1, order) for _ in range(n)]
    cid = H.CURVE_IDS[name]
    pts_b, infs = nmsm.mul_batch_packed(cid, H.point_bytes(name, P.BASE) * n, H.pack_scalars(ks), n, False)
    for s in (0x1D3F5A7C9B2E4F60718293A4B5C6D7E8F9 % order, 1, order - 1):
        total = (sum(ks) * s) % order
        exp = H.expected_tuple(name, H.expected_from_total(P, total))
        assert gpu_msm(nmsm, name, pts_b, H.pack_scalars([s] * n), n) == exp, (name, hex(s))
    # half of the scalars equal, the rest random: a giant bucket next to ordinary ones
    sc = [0xABCDEF0123456789 if i % 2 else rnd.randrange(order) for i in range(n)]
    total = sum(k * s for k, s in zip(ks, sc)) % order
    assert gpu_msm(nmsm, name, pts_b, H.pack_scalars(sc), n) == H.expected_tuple(name, H.expected_from_total(P, total))


def test_point_decoders_gpu(nmsm):
    """nmsm_points_decode: 1000 compressed BLS12-381 G1 encodings and the secp256k1 isPoint list, then an MSM on the
    decoded points (bytes in -> MSM out without host bigints)."""
    gb = load_golden("bls12_381.json")
    encs = [bytes.fromhex(c) for c in gb["G1_Compressed"]]
    pts, st = nmsm.points_decode(4, b"".join(encs), len(encs))
    assert st[0] == 2 and all(s == 1 for s in st[1:])
    for i in (1, 2, 500, 999):
        assert H.unpack_point("bls12_381_G1", pts[i * 96:(i + 1) * 96]) == R.bls12_381_g1_decode(encs[i])
    rnd = random.Random(9)
    G1 = R.CURVES["bls12_381_G1"]
    sc = [rnd.randrange(G1.Fn.ORDER) for _ in encs]
    tot = sum(i * s for i, s in enumerate(sc)) % G1.Fn.ORDER
    assert gpu_msm(nmsm, "bls12_381_G1", pts, H.pack_scalars(sc), len(encs)) == H.expected_tuple(
        "bls12_381_G1", G1.BASE.multiplyUnsafe(tot))
    # G2, 96-byte compressed (Fp2 square roots on the GPU), then subgroup checks and an MSM over the decoded points
    encs2 = [bytes.fromhex(c) for c in gb["G2_Compressed"]]
    bad2 = [bytes([0x80]) + bytes(94) + bytes([k]) for k in range(1, 9)]  # x = (k, 0): some on the twist, never in G2
    pts2, st2 = nmsm.points_decode(5, b"".join(encs2 + bad2), len(encs2) + len(bad2))
    assert st2[0] == 2 and all(s == 1 for s in st2[1:len(encs2)])
    for i in (1, 2, 100, 255):
        x, y = R.bls12_381_g2_decode(encs2[i])
        assert H.unpack_point("bls12_381_G2", pts2[i * 192:(i + 1) * 192]) == (x, y)
    for j, e in enumerate(bad2):
        try:
            exp_pt = R.bls12_381_g2_decode(e)
            assert st2[len(encs2) + j] == 1
            assert H.unpack_point("bls12_381_G2", pts2[(len(encs2) + j) * 192:(len(encs2) + j + 1) * 192]) == exp_pt
        except ValueError:
            assert st2[len(encs2) + j] == 0
    G2 = R.CURVES["bls12_381_G2"]
    ok_idx = [i for i, s in enumerate(st2) if s != 0]
    packed = b"".join(pts2[i * 192:(i + 1) * 192] for i in ok_idx)
    flags = nmsm.torsion_free_packed(5, packed, len(ok_idx))
    for k, i in enumerate(ok_idx):
        assert flags[k] == (1 if i < len(encs2) else 0), i  # i*G2 is in the subgroup; the small-x twist points are not
    sc2 = [rnd.randrange(G2.Fn.ORDER) for _ in encs2]
    tot2 = sum(i * s for i, s in enumerate(sc2)) % G2.Fn.ORDER
    assert gpu_msm(nmsm, "bls12_381_G2", pts2[: len(encs2) * 192], H.pack_scalars(sc2), len(encs2)) == H.expected_tuple(
        "bls12_381_G2", G2.BASE.multiplyUnsafe(tot2))
    C2 = nmsm.CURVES["bls12_381_G2"]
    assert C2.fromBytes(encs2[5]).equals(C2.BASE.multiply(5))
    s = load_golden("secp256k1.json")["isPoint33"]
    pts, st = nmsm.points_decode(0, b"".join(bytes.fromhex(e) for e, _ in s), len(s))
    for i, (e, exp) in enumerate(s):
        assert (st[i] == 1) == exp, e
        if exp and i % 50 == 0:
            assert H.unpack_point("secp256k1", pts[i * 64:(i + 1) * 64]) == R.secp256k1_decode_sec1(bytes.fromhex(e))
    ed = load_golden("ed25519.json")["vectors"]
    pts, st = nmsm.points_decode(1, b"".join(bytes.fromhex(v["pk"]) for v in ed), len(ed))
    assert all(x == 1 for x in st)
    a = R.ed25519_point_from_bytes(bytes.fromhex(ed[7]["pk"]), True).toAffine()
    assert H.unpack_point("ed25519", pts[7 * 64:8 * 64]) == (a["x"], a["y"])
    with pytest.raises(ValueError, match="no decoder"):
        nmsm.points_decode(2, b"", 0) if False else nmsm.points_decode(2, bytes(64), 1)


def test_point_codec_roundtrip_object_api(nmsm):
    """Point.toBytes / Point.fromBytes of the host mirror against the reference's encodings."""
    gb = load_golden("bls12_381.json")
    C = nmsm.CURVES["bls12_381_G1"]
    for i in (0, 1, 2, 77, 999):
        enc = bytes.fromhex(gb["G1_Compressed"][i])
        Pt = C.fromBytes(enc)
        assert Pt.toBytes(True) == enc and Pt.toBytes(False).hex() == gb["G1_Uncompressed"][i]
        assert Pt.equals(C.BASE.multiply(i)) if i else Pt.is0()
    S = nmsm.CURVES["secp256k1"]
    for k, x, y in load_golden("secp256k1.json")["privates2"][:10]:
        Pt = S.BASE.multiply(int(k))
        enc = Pt.toBytes(True)
        assert enc[1:].hex() == x and S.fromBytes(enc).equals(Pt)
        assert Pt.toBytes(False).hex() == "04" + x + y
    with pytest.raises(ValueError, match="bad point"):
        S.fromBytes(bytes([2]) + S.Fp.ORDER.to_bytes(32, "big"))
    E = nmsm.CURVES["ed25519"]
    for v in load_golden("ed25519.json")["vectors"][:8]:
        pk = bytes.fromhex(v["pk"])
        assert E.fromBytes(pk).toBytes() == pk
    # a point of the full curve group that is NOT in G1's prime-order subgroup must be rejected (assertValidity)
    p = R.BLS12_381_G1_CURVE["p"]
    x = 3
    while True:
        y2 = (x**3 + 4) % p
        y = pow(y2, (p + 1) // 4, p)
        if y * y % p == y2:
            cand = R.CURVES["bls12_381_G1"].fromAffine({"x": x, "y": y})
            if not cand.multiplyUnsafe(R.CURVES["bls12_381_G1"].Fn.ORDER - 1).add(cand).is0():
                break
        x += 1
    xb = bytearray(x.to_bytes(48, "big"))
    xb[0] |= 0x80 | (0x20 if (y * 2) // p else 0)
    with pytest.raises(ValueError, match="subgroup"):
        C.fromBytes(bytes(xb))


def test_bls12_381_g1_points_outside_the_subgroup(nmsm):
    """pippenger / multiply are the plain group law in the reference.  Points of E(Fp) outside the prime-order
    subgroup: NMSM_BLS12_381_G1_ANY matches the oracle (also through a fixed-base table), nmsm_mul_batch matches it
    for both ids (isTorsionFree / clearCofactor depend on that), and on torsion-free points both ids agree."""
    P = R.CURVES["bls12_381_G1"]
    C = nmsm.CURVES["bls12_381_G1"]
    rnd = random.Random(12)
    bad = H.bls_g1_non_subgroup_points(40)
    good = [P.BASE.multiplyUnsafe(rnd.randrange(1, P.Fn.ORDER)) for _ in range(24)]
    pts = R.normalizeZ(P, bad + good)
    rnd.shuffle(pts)
    scalars = [rnd.randrange(P.Fn.ORDER) for _ in pts]
    scalars[3] = P.Fn.ORDER - 1
    n = len(pts)
    exp = H.expected_tuple("bls12_381_G1", R.pippenger(P, pts, scalars))
    pb, sb = H.pack_points("bls12_381_G1", pts), H.pack_scalars(scalars)
    out, inf = nmsm.msm_packed(6, pb, sb, n)
    assert (*H.unpack_point("bls12_381_G1", out), inf) == exp
    ps = nmsm.PointSet(6, pb, n)
    ps.precompute(0)
    out, inf = ps.msm(sb, n)
    assert (*H.unpack_point("bls12_381_G1", out), inf) == exp
    ps.close()
    cpts = [C.fromAffine(p.toAffine()) for p in pts]
    got = nmsm.pippenger(C, cpts, scalars, assume_torsion_free=False)
    assert (got.x, got.y, 1 if got.is0() else 0) == exp
    # torsion-free inputs: the GLV id and the plain id give the same (reference) answer
    gpts = R.normalizeZ(P, good)
    gsc = scalars[: len(gpts)]
    gexp = H.expected_tuple("bls12_381_G1", R.pippenger(P, gpts, gsc))
    for cid in (4, 6):
        o, f = nmsm.msm_packed(cid, H.pack_points("bls12_381_G1", gpts), H.pack_scalars(gsc), len(gpts))
        assert (*H.unpack_point("bls12_381_G1", o), f) == gexp, cid
    # Point.multiply on non-subgroup points
    ks = [P.Fn.ORDER - 1, 0x396C8C005555E1568C00AAAB0000AAAB, 5, rnd.randrange(P.Fn.ORDER)]
    bp = R.normalizeZ(P, bad[:4])
    for cid in (4, 6):
        o, f = nmsm.mul_batch_packed(cid, H.pack_points("bls12_381_G1", bp), H.pack_scalars(ks), 4, False)
        for i in range(4):
            assert (*H.unpack_point("bls12_381_G1", o[i * 96:(i + 1) * 96]), f[i]) == \
                H.expected_tuple("bls12_381_G1", bp[i].multiplyUnsafe(ks[i])), (cid, i)
    # batch isTorsionFree (nmsm_points_torsion_free)
    flags = nmsm.torsion_free_packed(4, pb, n)
    assert list(flags) == [1 if p.multiplyUnsafe(P.Fn.ORDER - 1).add(p).is0() else 0 for p in pts]
    assert sum(flags) == len(good) and nmsm.torsion_free_packed(6, pb, n) == flags
    b0 = C.fromAffine(bad[0].toAffine())
    assert not b0.isTorsionFree() and b0.clearCofactor().isTorsionFree()
    assert C.fromAffine(good[0].toAffine()).isTorsionFree()


# ------------------------------------------------------------------------------------------------
# NTT over Fr (SURVEY §8 f4 companion): bit-exact against the oracle restatement of fft.ts
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field", ["bls12_381", "bn254"])
def test_ntt_matches_oracle_all_layouts(nmsm, field):
    """FFT.direct / FFT.inverse for every boundary layout (fft.ts:538-575), generator 7 (the reference tests') and the
    default generator, sizes 1 .. 2^13 (one tile, one partial tile, two passes), edge values 0 and r-1."""
    from nmsm import fft as GF
    from oracle import noble_fft as OF

    p = OF.FR[field]
    rnd = random.Random(99)
    for gen in (7, None):
        oracle = OF.FFT(OF.RootsOfUnity(p, gen))
        gpu = GF.FFT(GF.rootsOfUnity(field, gen))
        for bits in (0, 1, 2, 5, 10, 11, 12, 13):
            n = 1 << bits
            a = [rnd.randrange(p) for _ in range(n)]
            a[0] = p - 1
            if n > 2:
                a[2] = 0
            for bi in (False, True):
                for bo in (False, True):
                    assert gpu.direct(a, bi, bo) == oracle.direct(a, bi, bo), (field, gen, bits, bi, bo, "direct")
                    assert gpu.inverse(a, bi, bo) == oracle.inverse(a, bi, bo), (field, gen, bits, bi, bo, "inverse")
    g = load_golden("fft.json")
    ones = GF.FFT(GF.rootsOfUnity(field, 7)).direct([0, 1, 0, 0, 0, 0, 0, 0])  # evaluates x at the 8 roots: the root table
    assert [str(x) for x in ones] == g["%s_roots3" % field]


@pytest.mark.parametrize("field,bits", [("bls12_381", 20), ("bn254", 22)])
def test_ntt_large_properties(nmsm, field, bits):
    """BASELINE-size transforms through size-independent properties (test/fft.test.ts:545-617): inverse(direct(a)) == a,
    direct of a constant, additivity, and spot values against the DFT definition a(omega^k)."""
    import numpy as np

    from nmsm import fft as GF
    from oracle import noble_fft as OF

    p = OF.FR[field]
    n = 1 << bits
    rs = np.random.RandomState(7)
    raw = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x0F  # < 2^252 < r
    a_b = raw.tobytes()
    d_b = GF.ntt_packed(field, a_b, bits, generator=7)
    assert GF.ntt_packed(field, d_b, bits, inverse=True, generator=7) == a_b
    # brp layouts at full size: direct(a, out=brp) then inverse(in=brp) returns a
    d_brp = GF.ntt_packed(field, a_b, bits, brp_output=True, generator=7)
    assert GF.ntt_packed(field, d_brp, bits, inverse=True, brp_input=True, generator=7) == a_b
    # DFT definition on a sparse polynomial: a = c0 + c1 x + c5 x^5 (+ zeros) at a few roots
    coeffs = {0: 12345, 1: p - 2, 5: 1 << 200, n - 1: 77}
    sp = bytearray(n * 32)
    for i, c in coeffs.items():
        sp[i * 32:(i + 1) * 32] = c.to_bytes(32, "little")
    out = GF.ntt_packed(field, bytes(sp), bits, generator=7)
    w = OF.RootsOfUnity(p, 7).omega(bits)
    for k in (0, 1, 2, n // 2, n - 1, 123457 % n):
        wk = pow(w, k, p)
        exp = sum(c * pow(wk, i, p) for i, c in coeffs.items()) % p
        assert int.from_bytes(out[k * 32:(k + 1) * 32], "little") == exp, k
    const = (5).to_bytes(32, "little") * n
    oc = GF.ntt_packed(field, const, bits, generator=7)
    assert int.from_bytes(oc[:32], "little") == 5 * n % p and not any(oc[32:])


def test_ntt_errors(nmsm):
    from nmsm import fft as GF
    from oracle import noble_fft as OF

    p = OF.FR["bn254"]
    f = GF.FFT(GF.rootsOfUnity("bn254", 7))
    with pytest.raises(ValueError, match="power of two"):
        f.direct([1, 2, 3])
    bad = [1, 2, p, 4]
    with pytest.raises(ValueError, match="invalid field element at index 2"):
        f.direct(bad)
    assert f.direct([1, 2, 3, 4]) == OF.FFT(OF.RootsOfUnity(p, 7)).direct([1, 2, 3, 4])  # still usable afterwards
    import ctypes

    lib = nmsm._lib.load()
    dummy = ctypes.create_string_buffer(64)
    assert lib.nmsm_ntt(2, ctypes.cast(dummy, ctypes.c_void_p), 29, 7, 0, 0, 0) != 0  # bn254 Fr has 2-adicity 28
    assert b"wrong bits 29 powerOfTwo=28" in lib.nmsm_last_error()
    assert lib.nmsm_ntt(0, ctypes.cast(dummy, ctypes.c_void_p), 1, 7, 0, 0, 0) != 0   # secp256k1: no NTT field


def test_ntt_output_feeds_msm_on_device(nmsm):
    """The prover-loop composition without host round trips: coefficients -> nmsm_ntt_device (evaluations, canonical
    32-byte scalars) -> nmsm_msm_device over the same device buffer; equals the oracle's pippenger over the oracle's FFT."""
    import torch

    from nmsm import fft as GF
    from oracle import noble_fft as OF

    G1 = R.CURVES["bls12_381_G1"]
    p = OF.FR["bls12_381"]
    n, bits = 256, 8
    rnd = random.Random(21)
    coeffs = [rnd.randrange(p) for _ in range(n)]
    evals = OF.FFT(OF.RootsOfUnity(p, 7)).direct(coeffs)
    P, pts, _, _ = H.soak_inputs("bls12_381_G1", n, seed_offset=3)
    exp = H.expected_tuple("bls12_381_G1", R.pippenger(P, pts, evals))
    d_sc = torch.frombuffer(bytearray(b"".join(c.to_bytes(32, "little") for c in coeffs)), dtype=torch.uint8).cuda()
    d_pts = torch.frombuffer(bytearray(H.pack_points("bls12_381_G1", pts)), dtype=torch.uint8).cuda()
    GF.ntt_device("bls12_381", d_sc.data_ptr(), bits, generator=7)
    assert bytes(d_sc.cpu().numpy().tobytes()) == b"".join(e.to_bytes(32, "little") for e in evals)
    out, inf = nmsm.msm_device(4, d_pts.data_ptr(), d_sc.data_ptr(), n)
    assert (*H.unpack_point("bls12_381_G1", out), inf) == exp
