"""Host-side logic of the reference-facing mirror (noble-curves_b200/nmsm) that runs before any GPU call: the argument
validation and error text of pippenger / multiply / FFT (curve.ts:390-404,863-878; weierstrass.ts:900-928;
edwards.ts:555-577; fft.ts:518-575), the wire encodings of toBytes, packing helpers.  No device is needed: every case
here must be decided on the host, and anything that reaches the library must fail loudly (no CPU fallback)."""
import pytest

import nmsm
from nmsm import fft as GF
from oracle import noble_ref as R

from conftest import load_golden


@pytest.mark.parametrize("name", ["secp256k1", "ed25519", "bls12_381_G1", "bn254_G2"])
def test_pippenger_validation_messages(name):
    C = nmsm.CURVES[name]
    G = C.BASE
    n = C.Fn.ORDER
    with pytest.raises(ValueError, match="arrays of points and scalars must have equal length"):
        nmsm.pippenger(C, [G, G], [1])
    with pytest.raises(ValueError, match="invalid point at index 1"):
        nmsm.pippenger(C, [G, object()], [1, 2])
    with pytest.raises(ValueError, match="invalid scalar at index 2"):
        nmsm.pippenger(C, [G, G, G], [1, 2, n])
    with pytest.raises(ValueError, match="invalid scalar at index 0"):
        nmsm.pippenger(C, [G], [-1])
    with pytest.raises(ValueError, match="array of scalars expected"):
        nmsm.pippenger(C, [G], 5)
    with pytest.raises(TypeError):
        nmsm.pippenger(C, "GG", [1, 2])
    # points are validated before scalars, as in the reference (curve.ts:871-872)
    with pytest.raises(ValueError, match="invalid point at index 0"):
        nmsm.pippenger(C, [None], [n])
    assert nmsm.pippenger(C, [], []).is0()  # curve.ts:878: empty input is the identity, no device needed


@pytest.mark.parametrize("name", ["secp256k1", "ed25519", "bls12_381_G1"])
def test_multiply_range_messages(name):
    C = nmsm.CURVES[name]
    n = C.Fn.ORDER
    edw = name == "ed25519"
    msg1 = "invalid scalar: expected 1 <= sc < curve.n" if edw else "invalid scalar: out of range"
    msg0 = "invalid scalar: expected 0 <= sc < curve.n" if edw else "invalid scalar: out of range"
    for bad in (0, n, -3):
        with pytest.raises(ValueError, match=msg1.replace("(", "\\(").replace(")", "\\)")):
            C.BASE.multiply(bad)
    for bad in (n, -1):
        with pytest.raises(ValueError, match=msg0):
            C.BASE.multiplyUnsafe(bad)
    with pytest.raises(ValueError, match="invalid window size"):
        C.BASE.precompute(0)
    assert nmsm.multiply_many(C, [], []) == []


def test_wire_encodings_match_reference_vectors_on_the_host():
    """toBytes is pure byte packing on the host (weierstrass.ts:541-563, bls12-381.ts:377-402): check it against the
    reference's encodings without touching the device."""
    g = load_golden("bls12_381.json")
    C = nmsm.CURVES["bls12_381_G1"]
    O = R.CURVES["bls12_381_G1"]
    for i in (1, 2, 77, 999):
        a = O.BASE.multiplyUnsafe(i).toAffine()
        P = C.fromAffine(a)
        assert P.toBytes(True).hex() == g["G1_Compressed"][i] and P.toBytes(False).hex() == g["G1_Uncompressed"][i]
    assert C.ZERO.toBytes(True).hex() == g["G1_Compressed"][0]
    S = nmsm.CURVES["secp256k1"]
    for k, x, y in load_golden("secp256k1.json")["privates2"][:5]:
        P = S.fromAffine({"x": int(x, 16), "y": int(y, 16)})
        assert P.toBytes(False).hex() == "04" + x + y and P.toBytes(True)[1:].hex() == x
    E = nmsm.CURVES["ed25519"]
    OE = R.CURVES["ed25519"]
    v = load_golden("ed25519.json")["vectors"][0]
    pk = bytes.fromhex(v["pk"])
    a = R.ed25519_point_from_bytes(pk, True).toAffine()
    assert E.fromAffine(a).toBytes() == pk and OE is not None


def test_fft_wrapper_validation():
    f = GF.FFT(GF.rootsOfUnity("bls12_381", 7))
    with pytest.raises(ValueError, match="power of two"):
        f.direct([1, 2, 3])
    with pytest.raises(ValueError, match="power of two"):
        f.inverse([])
    with pytest.raises(ValueError, match="scalar fields of bn254 and bls12_381 only"):
        GF.rootsOfUnity("secp256k1")
    with pytest.raises(TypeError):
        GF.rootsOfUnity("bn254", 7.0)
    r = GF.rootsOfUnity("bn254")
    assert r.info["powerOfTwo"] == 28 and r.info["G"] == 5 and (r.info["oddFactor"] << 28) + 1 == GF.FR_ORDER["bn254"]
    with pytest.raises(ValueError, match="wrong Polynomial length"):
        GF.ntt_packed("bn254", b"\x00" * 31, 0)


def test_compute_entry_points_refuse_without_a_device():
    """There is no CPU fallback: on this box (no GPU) everything that needs the device raises."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    C = nmsm.CURVES["secp256k1"]
    with pytest.raises(Exception, match="CUDA|cuda|device"):
        nmsm.pippenger(C, [C.BASE], [5])
    with pytest.raises(Exception, match="CUDA|cuda|device"):
        GF.FFT(GF.rootsOfUnity("bn254", 7)).direct([1, 2, 3, 4])
    with pytest.raises(Exception, match="CUDA|cuda|device"):
        nmsm.points_decode(0, bytes(33), 1)


def test_round2_host_validation_without_device():
    """Constructor / mulAddUnsafe / fromBytes argument checks that never reach the GPU (weierstrass.ts:698-701,
    curve.ts:826-832, edwards.ts:405-410)."""
    import nmsm

    C = nmsm.CURVES["secp256k1"]
    with pytest.raises(ValueError, match="bad point coordinate y"):
        C(1, 0)
    assert C.BASE._valid and C.ZERO._valid and not C(1, 1)._valid
    assert C.BASE.negate()._valid
    G2 = nmsm.CURVES["bls12_381_G2"]
    with pytest.raises(ValueError, match="bad point coordinate y"):
        G2((1, 2), (0, 0))
    E = nmsm.CURVES["ed25519"]
    assert E(0, 1, True).is0()  # Edwards: y = 0 is not special, identity is (0, 1)
    E(5, 0)
    n = C.Fn.ORDER
    with pytest.raises(ValueError, match="invalid scalar at index 1"):
        nmsm.mulAddUnsafe(C, [C.BASE, C.BASE], [1, n])
    with pytest.raises(ValueError, match="invalid scalar at index 0"):
        nmsm.mulAddUnsafe(C, [C.BASE], [n ** 4], True)
    with pytest.raises(TypeError):
        nmsm.mulAddUnsafe(C, [C.BASE], [1], 1)
    with pytest.raises(ValueError, match="equal length"):
        nmsm.mulAddUnsafe(C, [C.BASE], [1, 2], True)
    assert nmsm.mulAddUnsafe(C, [], [], True).is0()
    with pytest.raises(TypeError):
        E.fromBytes(bytes(32), 1)
    # curve id selection for BLS12-381 G1 (id 4 only for known-valid inputs)
    B = nmsm.CURVES["bls12_381_G1"]
    unv = B.fromAffine({"x": B.BASE.x, "y": B.BASE.y})
    assert nmsm._curve_id_for(B, [B.BASE, B.BASE.negate()]) == 4
    assert nmsm._curve_id_for(B, [B.BASE, unv]) == 6
    assert nmsm._curve_id_for(B, [unv], True) == 4 and nmsm._curve_id_for(B, [B.BASE], False) == 6
    assert nmsm._curve_id_for(C, [C(1, 1)]) == 0
    # same rule for BLS12-381 G2 (id 5 = psi split, id 7 = plain windows)
    unv2 = G2.fromAffine({"x": G2.BASE.x, "y": G2.BASE.y})
    assert nmsm._curve_id_for(G2, [G2.BASE]) == 5 and nmsm._curve_id_for(G2, [G2.BASE, unv2]) == 7


def test_cold_path_validators():
    """validatePointCons / validateW / validateTableBytes (curve.ts:259-272, :328-346, :947, :776-781): typed errors and
    the reference's messages, all before anything reaches the GPU."""
    import nmsm

    C = nmsm.CURVES["secp256k1"]
    nmsm.validatePointCons(C)
    for bad in (None, 5, "Point", C.BASE):
        with pytest.raises(TypeError, match="expected constructor"):
            nmsm.validatePointCons(bad)

    class Fake:
        pass

    with pytest.raises(TypeError, match="Point.fromAffine"):
        nmsm.validatePointCons(Fake)
    with pytest.raises(TypeError, match="expected constructor"):
        nmsm.pippenger(None, [], [])
    with pytest.raises(TypeError, match="Point.fromAffine"):
        nmsm.mulAddUnsafe(Fake, [], [])
    with pytest.raises(TypeError, match="Point.fromAffine"):
        nmsm.interleavedMSMUnsafe(Fake, [], 4)
    bits = C.Fn.BITS
    for w in (1, 0, -3, bits + 1, 2.5, True):
        with pytest.raises(ValueError, match=r"invalid window size, expected \[2\.\.%d\]" % bits):
            nmsm.interleavedMSMUnsafe(C, [], w)
    # 2^20 points at W = 16: 2^34 table entries of (4 * 32 + 128) bytes -> the reference refuses
    pts = [C.BASE] * 8
    with pytest.raises(ValueError, match=r"invalid window size: table would need ~\d+ MiB, max 2048 MiB"):
        nmsm.interleavedMSMUnsafe(C, pts, 24)
    with pytest.raises(ValueError, match=r"invalid window size, expected \[1\.\.%d\]" % bits):
        C.BASE.precompute(0)
    with pytest.raises(ValueError, match="table would need"):
        C.BASE.precompute(30)
    assert nmsm.interleavedMSMUnsafe(C, [], 4)([]).is0()  # empty set: nothing to upload
    assert C.fromHex is not None and C.BASE.toHex() == C.BASE.toBytes().hex()
    with pytest.raises(TypeError):
        C.fromHex(b"02")
    with pytest.raises(ValueError):
        C.fromHex("zz")
    assert C.BASE.hasEvenY() == (C.BASE.y % 2 == 0) and C.BASE.negate().hasEvenY() != C.BASE.hasEvenY()
    with pytest.raises(ValueError, match="isOdd"):
        nmsm.CURVES["bls12_381_G2"].BASE.hasEvenY()


def test_packed_entry_points_validate_lengths_before_the_gpu():
    import nmsm

    with pytest.raises(ValueError, match="packed arrays do not match n"):
        nmsm.ed25519_verify_batch_packed(bytes(64), bytes(32), b"", bytes(8), 1, bytes(16))  # n + 1 offsets expected
    with pytest.raises(ValueError, match="packed arrays do not match n"):
        nmsm.ed25519_verify_batch_packed(bytes(63), bytes(32), b"", bytes(16), 1, bytes(16))
    with pytest.raises(ValueError, match="z must hold 16 bytes per signature"):
        nmsm.ed25519_verify_batch([bytes(64)], [b"m"], [bytes(32)], bytes(15))
    with pytest.raises(ValueError, match="signature expected 64 bytes"):
        nmsm.ed25519_verify_batch([bytes(63)], [b"m"], [bytes(32)])
