"""Pin the CPU oracle (oracle/noble_ref.py) against the reference's own golden vectors.

Fixtures: tests/golden/*.json, extracted from /root/reference/test/vectors by
tests/golden/make_golden.py.  Mirrors test/secp256k1.test.ts:59-127, test/bls12-381.test.ts:1463-1533,
test/bn254.test.ts:750-778,859-887, test/ed25519.test.ts:50-78, test/nist.test.ts:551.
"""
import pytest

from conftest import load_golden
from oracle import noble_ref as R

SECP = R.CURVES["secp256k1"]
ED = R.CURVES["ed25519"]
BN = R.CURVES["bn254_G1"]
G1 = R.CURVES["bls12_381_G1"]
G2 = R.CURVES["bls12_381_G2"]


def sec1_decode(hexstr):
    """SEC1 (weierstrass.ts:541-605): 02/03 compressed, 04 uncompressed."""
    b = bytes.fromhex(hexstr)
    p = SECP.Fp.ORDER
    if b[0] == 4:
        return SECP.fromAffine({"x": int.from_bytes(b[1:33], "big"), "y": int.from_bytes(b[33:], "big")})
    x = int.from_bytes(b[1:], "big")
    y = pow((x * x * x + 7) % p, (p + 1) // 4, p)
    assert (y * y - x * x * x - 7) % p == 0
    if (y & 1) != (b[0] & 1):
        y = p - y
    return SECP.fromAffine({"x": x, "y": y})


def sec1_encode(P, compressed=True):
    a = P.toAffine()
    if compressed:
        return (bytes([2 + (a["y"] & 1)]) + a["x"].to_bytes(32, "big")).hex()
    return (b"\x04" + a["x"].to_bytes(32, "big") + a["y"].to_bytes(32, "big")).hex()


def test_secp256k1_privates2():
    g = load_golden("secp256k1.json")
    assert len(g["privates2"]) == 45
    for k, x, y in g["privates2"]:
        k = int(k)
        for pt in (SECP.BASE.multiply(k), SECP.BASE.multiplyUnsafe(k), R.naive_mul(SECP, SECP.BASE, k)):
            assert R.affine_tuple(SECP, pt) == (int(x, 16), int(y, 16))


def test_secp256k1_points_json():
    g = load_golden("secp256k1.json")
    for P, d, exp in g["pointMultiply"]:
        p = sec1_decode(P)
        d = int(d, 16)
        assert sec1_encode(p.multiply(d), len(exp) == 66) == exp
        assert sec1_encode(p.multiplyUnsafe(d), len(exp) == 66) == exp
    for d, exp in g["pointFromScalar"]:
        assert sec1_encode(SECP.BASE.multiply(int(d, 16)), len(exp) == 66) == exp
    n_add = 0
    for P, Q, exp in g["pointAdd"]:
        if exp is None:
            assert sec1_decode(P).add(sec1_decode(Q)).is0()
        else:
            assert sec1_encode(sec1_decode(P).add(sec1_decode(Q)), len(exp) == 66) == exp
        n_add += 1
    assert n_add == 118


def test_secp256k1_endomorphism():
    g = load_golden("secp256k1.json")
    for v in g["endomorphism"]:
        a = SECP.fromAffine({"x": int(v["ax"]), "y": int(v["ay"])})
        c = a.multiplyUnsafe(int(v["scalar"]))
        assert R.affine_tuple(SECP, c) == (int(v["cx"]), int(v["cy"]))
        c2 = a.multiply(int(v["scalar"]))
        assert R.affine_tuple(SECP, c2) == (int(v["cx"]), int(v["cy"]))


def g1_decode_uncompressed(hexstr):
    b = bytearray(bytes.fromhex(hexstr))
    assert len(b) == 96
    flags = b[0] >> 5
    b[0] &= 0x1F
    if flags & 2:
        return G1.ZERO
    return G1.fromAffine({"x": int.from_bytes(b[:48], "big"), "y": int.from_bytes(b[48:], "big")})


def g2_decode_uncompressed(hexstr):
    b = bytearray(bytes.fromhex(hexstr))
    assert len(b) == 192
    flags = b[0] >> 5
    b[0] &= 0x1F
    if flags & 2:
        return G2.ZERO
    x1, x0, y1, y0 = (int.from_bytes(b[i * 48:(i + 1) * 48], "big") for i in range(4))
    return G2.fromAffine({"x": (x0, x1), "y": (y0, y1)})


def test_bls12_381_g1_multiples():
    g = load_golden("bls12_381.json")["G1_Uncompressed"]
    assert len(g) == 1000
    acc = G1.ZERO
    for i, h in enumerate(g):
        exp = g1_decode_uncompressed(h)
        assert acc.equals(exp), i  # running sum i*G via complete add (incl. O+G, G+G)
        if 0 < i < 200 or i % 97 == 0 and i:
            assert R.affine_tuple(G1, G1.BASE.multiply(i)) == R.affine_tuple(G1, exp)
            assert R.affine_tuple(G1, G1.BASE.multiplyUnsafe(i)) == R.affine_tuple(G1, exp)
        acc = acc.add(G1.BASE)


def test_bls12_381_g2_multiples():
    g = load_golden("bls12_381.json")["G2_Uncompressed"]
    acc = G2.ZERO
    for i, h in enumerate(g):
        exp = g2_decode_uncompressed(h)
        assert acc.equals(exp), i
        if 0 < i < 24 or i in (100, 255):
            assert R.affine_tuple(G2, G2.BASE.multiply(i)) == R.affine_tuple(G2, exp)
            assert R.affine_tuple(G2, G2.BASE.multiplyUnsafe(i)) == R.affine_tuple(G2, exp)
        acc = acc.add(G2.BASE)


def _eth_nums(inp, count):
    """test/bn254.test.ts:740-748 ethNums: zero-extended tape of 32-byte words."""
    if inp.startswith("0x"):
        inp = inp[2:]
    if not inp:
        return [0] * count
    if len(inp) % 64:
        inp += "0" * (64 - len(inp) % 64)
    res = [int(inp[i:i + 64], 16) for i in range(0, len(inp), 64)]
    while len(res) < count:
        res.append(0)
    return res


def _on_curve_bn(x, y):
    p = BN.Fp.ORDER
    return (x == 0 and y == 0) or (y * y - x * x * x - 3) % p == 0


def test_bn254_eth_dump_and_seda():
    g = load_golden("bn254.json")
    n_ok = 0
    for inp, outp in g["eth_mul"]:
        Cx, Cy = _eth_nums(outp, 2)
        Ax, Ay, scalar = _eth_nums(inp, 3)[:3]
        try:
            if not (BN.Fp.isValid(Ax) and BN.Fp.isValid(Ay) and _on_curve_bn(Ax, Ay)):
                raise ValueError
            A = BN.fromAffine({"x": Ax, "y": Ay})
            s = scalar % BN.Fn.ORDER
            res = R.affine_tuple(BN, A.multiply(s))
            assert res == R.affine_tuple(BN, A.multiplyUnsafe(s))
            n_ok += 1
        except ValueError:
            res = (0, 0)
        assert res == (Cx, Cy)
    assert n_ok > 20
    for inp, outp in g["eth_add"]:
        Cx, Cy = _eth_nums(outp, 2)
        Ax, Ay, Bx, By = _eth_nums(inp, 4)[:4]
        try:
            for x, y in ((Ax, Ay), (Bx, By)):
                if not (BN.Fp.isValid(x) and BN.Fp.isValid(y) and _on_curve_bn(x, y)):
                    raise ValueError
            res = R.affine_tuple(BN, BN.fromAffine({"x": Ax, "y": Ay}).add(BN.fromAffine({"x": Bx, "y": By})))
        except ValueError:
            res = (0, 0)
        assert res == (Cx, Cy)
    for t in g["seda_add"]:
        A = BN.fromAffine({"x": int(t["x1"], 16), "y": int(t["y1"], 16)})
        B = BN.fromAffine({"x": int(t["x2"], 16), "y": int(t["y2"], 16)})
        assert R.affine_tuple(BN, A.add(B)) == (int(t["result"][:64], 16), int(t["result"][64:], 16))
    for t in g["seda_mul"]:
        A = BN.fromAffine({"x": int(t["x"], 16), "y": int(t["y"], 16)})
        s = int(t["scalar"], 16) % BN.Fn.ORDER
        exp = (int(t["result"][:64], 16), int(t["result"][64:], 16))
        if s == 0:
            continue
        assert R.affine_tuple(BN, A.multiply(s)) == exp
        assert R.affine_tuple(BN, A.multiplyUnsafe(s)) == exp


def test_ed25519_rfc8032_public_keys():
    g = load_golden("ed25519.json")["vectors"]
    assert len(g) == 128
    for v in g[:64]:
        assert R.ed25519_public_key(bytes.fromhex(v["sk"])).hex() == v["pk"]


def test_ed25519_verify_oracle_against_reference_vectors():
    """EdDSA verify restatement (edwards.ts:942-989) vs test/vectors/ed25519: RFC 8032 signatures,
    the 196 ZIP-215 cases (test/ed25519.test.ts:392-405) and the strict-mode edge cases (:189-196)."""
    g = load_golden("ed25519.json")
    for v in g["vectors"][:48]:
        sig, msg, pk = bytes.fromhex(v["sig"]), bytes.fromhex(v["msg"]), bytes.fromhex(v["pk"])
        assert R.ed25519_verify(sig, msg, pk) is True
        bad = bytearray(sig)
        bad[5] ^= 1
        assert R.ed25519_verify(bytes(bad), msg, pk) is False
        assert R.ed25519_verify(sig, msg + b"x", pk) is False
    assert len(g["zip215"]) == 196
    for v in g["zip215"]:
        got = R.ed25519_verify(bytes.fromhex(v["sig_bytes"]), b"Zcash", bytes.fromhex(v["vk_bytes"]))
        assert got == v["valid_zip215"], v
    for i in (0, 1, 6, 7, 8, 9, 10, 11):
        v = g["edge_cases"][i]
        assert R.ed25519_verify(bytes.fromhex(v["signature"]), bytes.fromhex(v["message"]), bytes.fromhex(v["pub_key"]),
                                zip215=False) is False
    # s = l + 1 is rejected (test/ed25519.test.ts:406-414)
    sig = bytes.fromhex("5866666666666666666666666666666666666666666666666666666666666666"
                        "eed3f55c1a631258d69cf7a2def9de1400000000000000000000000000000010")
    assert R.ed25519_verify(sig, b"Zcash", bytes.fromhex(g["zip215"][0]["vk_bytes"])) is False


def test_codec_oracle_against_reference_vectors():
    """Decode restatements vs test/vectors: zkcrypto compressed G1 i*G (test/bls12-381.test.ts:1463-1500) and the
    secp256k1 isPoint list (test/secp256k1.test.ts:96-104, 33-byte encodings)."""
    g = load_golden("bls12_381.json")
    for i, (c, u) in enumerate(zip(g["G1_Compressed"], g["G1_Uncompressed"])):
        assert R.bls12_381_g1_decode(bytes.fromhex(c)) == R.bls12_381_g1_decode(bytes.fromhex(u)), i
        if i:
            assert R.bls12_381_g1_decode(bytes.fromhex(c)) == R.affine_tuple(G1, g1_decode_uncompressed(u))
    # G2: 96-byte compressed (c1 || c0, sort bit over [y.c1, y.c0]) against the uncompressed list, incl. Fp2 sqrt
    for i, (c, u) in enumerate(zip(g["G2_Compressed"], g["G2_Uncompressed"])):
        if i % 4 and i > 16:
            continue
        dc = R.bls12_381_g2_decode(bytes.fromhex(c))
        assert dc == R.bls12_381_g2_decode(bytes.fromhex(u)), i
        if i:
            assert dc == R.affine_tuple(G2, g2_decode_uncompressed(u))
    F2 = R.Field2(R.Field(G1.Fp.ORDER))
    for v in ((4, 0), (0, 9), (5, 0), (3, 7)):
        sq = F2.sqr(v)
        r = F2.sqrt(sq)
        assert F2.sqr(r) == sq and r in (v, F2.neg(v))
    with pytest.raises(ValueError):
        F2.sqrt((1, 1))  # (1 + u) is a non-residue in Fp2 (the sextic-twist constant's base)
    s = load_golden("secp256k1.json")
    assert len(s["isPoint33"]) > 1000
    for enc, exp in s["isPoint33"]:
        try:
            x, y = R.secp256k1_decode_sec1(bytes.fromhex(enc))
            ok = (y * y - x * x * x - 7) % SECP.Fp.ORDER == 0
        except ValueError:
            ok = False
        assert ok == exp, enc
