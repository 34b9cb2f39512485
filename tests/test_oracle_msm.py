"""Oracle self-consistency on the MSM path, restating the reference's differential MSM tests.

Mirrors test/point.test.ts:264-305 (pippenger basics), :825-862 (secp256k1 L=7, 2048 x G) and
test/slow-curves.test.ts:185-252 (boundary soak, scalar-in-exponent expectation).
"""
import pytest

from oracle import noble_ref as R

NAMES = ["secp256k1", "ed25519", "bn254_G1", "bls12_381_G1", "bn254_G2", "bls12_381_G2"]
# curve index in test/slow-curves.test.ts:186-194 (secp256r1 has index 1 and is out of scope)
SOAK_INDEX = {"secp256k1": 0, "ed25519": 2, "bls12_381_G1": 3, "bls12_381_G2": 4, "bn254_G1": 5, "bn254_G2": 6}


def soak_inputs(name, max_n):
    """test/slow-curves.test.ts:199-222."""
    P = R.CURVES[name]
    order = P.Fn.ORDER
    rng = R.Xorshift64(0x6D736D0000000000 + SOAK_INDEX[name])
    start = rng.rndBelow(order - 1) + 1
    step = rng.rndBelow(order - 1) + 1
    step_point = P.BASE.multiplyUnsafe(step)
    points, scalars, totals = [], [], []
    point = P.BASE.multiplyUnsafe(start)
    ps = start
    total = 0
    for i in range(max_n):
        s = 0 if i % 17 == 0 else rng.rndBelow(order)
        points.append(point)
        scalars.append(s)
        total = (total + ps * s) % order
        totals.append(total)
        point = point.add(step_point)
        ps = (ps + step) % order
    return P, points, scalars, totals


@pytest.mark.parametrize("name", NAMES)
def test_pippenger_basic(name):
    P = R.CURVES[name]
    G = P.BASE
    assert R.pippenger(P, [G], [0]).equals(P.ZERO)
    assert R.pippenger(P, [], []).equals(P.ZERO)
    assert R.pippenger(P, [P.ZERO], [123]).equals(P.ZERO)
    assert R.pippenger(P, [G], [123]).equals(G.multiply(123))
    pts = [G, G.double(), G.double().double(), G.double().double().double()]
    assert R.pippenger(P, pts, [3, 5, 7, 11]).equals(G.multiply(129))
    with pytest.raises(ValueError, match="invalid scalar at index 0"):
        R.pippenger(P, [G], [P.Fn.ORDER])
    with pytest.raises(ValueError, match="equal length"):
        R.pippenger(P, [G], [1, 2])
    with pytest.raises(ValueError, match="invalid point at index 1"):
        R.pippenger(P, [G, 5], [1, 2])


@pytest.mark.parametrize("name", ["secp256k1", "ed25519", "bn254_G1", "bls12_381_G1"])
def test_pippenger_soak_small(name):
    P, points, scalars, totals = soak_inputs(name, 129)
    for size in (31, 32, 33, 127, 128, 129):
        exp = P.BASE.multiplyUnsafe(totals[size - 1]) if totals[size - 1] else P.ZERO
        assert R.pippenger(P, points[:size], scalars[:size]).equals(exp)
        if size == 33:
            assert R.interleavedMSMUnsafe(P, points[:size], 5)(scalars[:size]).equals(exp)


@pytest.mark.parametrize("name", ["bn254_G2", "bls12_381_G2"])
def test_pippenger_soak_g2(name):
    P, points, scalars, totals = soak_inputs(name, 33)
    for size in (31, 33):
        exp = P.BASE.multiplyUnsafe(totals[size - 1])
        assert R.pippenger(P, points[:size], scalars[:size]).equals(exp)


def test_pippenger_secp256k1_same_point():
    """test/point.test.ts:842-853: 2048 x G with scalar 2^10 - 1 (scaled down to 256 for CPU time)."""
    P = R.CURVES["secp256k1"]
    n = 256
    s = 2**10 - 1
    exp = P.BASE.multiply((n * s) % P.Fn.ORDER)
    assert R.pippenger(P, [P.BASE] * n, [s] * n).equals(exp)


@pytest.mark.parametrize("name", ["secp256k1", "ed25519", "bls12_381_G1"])
def test_multiply_variants_agree(name):
    """test/slow-curves.test.ts:128-152 (sampled): multiply == multiplyUnsafe == mulCT == naive."""
    P = R.CURVES[name]
    rng = R.Xorshift64(0xDEADBEEF)
    pt = P.BASE.multiplyUnsafe(rng.rndBelow(P.Fn.ORDER - 1) + 1)
    for _ in range(4):
        k = rng.rndBelow(P.Fn.ORDER - 1) + 1
        exp = R.affine_tuple(P, R.naive_mul(P, pt, k))
        assert R.affine_tuple(P, pt.multiply(k)) == exp
        assert R.affine_tuple(P, pt.multiplyUnsafe(k)) == exp
        assert R.affine_tuple(P, P.wnaf.mulCT(pt, k)[0]) == exp
        assert R.affine_tuple(P, P.BASE.multiply(k)) == R.affine_tuple(P, R.naive_mul(P, P.BASE, k))
    with pytest.raises(ValueError):
        pt.multiply(0)
    assert pt.multiplyUnsafe(0).equals(P.ZERO)
    with pytest.raises(ValueError):
        pt.multiplyUnsafe(P.Fn.ORDER)
