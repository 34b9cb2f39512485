"""nmsm — host-side mirror of noble-curves' scalar-multiplication / MSM surface over libnmsm.so.

The reference's host language is TypeScript (no Node / N-API headers exist in this image, see
INTEGRATION.md for the addon a maintainer would add); this Python layer mirrors the same names,
argument meaning and error behaviour so the parity tests read like the reference's own tests:

    pippenger(c, points, scalars)          /root/reference/src/abstract/curve.ts:863-905
    mulAddUnsafe(c, points, scalars)       src/abstract/curve.ts:820-836 (same value, computed as an MSM)
    Point.multiply / Point.multiplyUnsafe  src/abstract/weierstrass.ts:900-928, src/abstract/edwards.ts:555-577
    Point.BASE / ZERO / fromAffine / toAffine / add / double / negate / equals / is0
    secp256k1.Point, ed25519.Point, bn254.G1/G2.Point, bls12_381.G1/G2.Point

plus the packed fast paths the reference lacks: msm_packed, msm_device, multiply_many.
Every curve operation runs on the GPU through the C ABI; the host only marshals bytes and validates
argument shapes.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

from . import _lib
from ._lib import NmsmError, PlanInfo, TIMING_NAMES, init  # noqa: F401

SECP256K1, ED25519, BN254_G1, BN254_G2, BLS12_381_G1, BLS12_381_G2 = range(6)


class _Field:
    """Shape-compatible stand-in for the IField members the MSM surface reads (modular.ts:429-607)."""

    def __init__(self, order: int, is_le: bool = False):
        self.ORDER = order
        self.BITS = order.bit_length()
        self.BYTES = (self.BITS + 7) // 8
        self.isLE = is_le

    def isValid(self, num) -> bool:
        return isinstance(num, int) and not isinstance(num, bool) and 0 <= num < self.ORDER

    def isValidNot0(self, num) -> bool:
        return self.isValid(num) and num != 0


class _Field2:
    def __init__(self, fp: _Field):
        self.Fp = fp
        self.ORDER = fp.ORDER * fp.ORDER
        self.BITS = self.ORDER.bit_length()
        self.BYTES = 2 * fp.BYTES

    def isValid(self, num) -> bool:
        return isinstance(num, tuple) and len(num) == 2 and self.Fp.isValid(num[0]) and self.Fp.isValid(num[1])


def _coord_to_bytes(v, nbytes: int, parts: int) -> bytes:
    if parts == 1:
        return v.to_bytes(nbytes, "little")
    return v[0].to_bytes(nbytes, "little") + v[1].to_bytes(nbytes, "little")


def _coord_from_bytes(b: bytes, nbytes: int, parts: int):
    if parts == 1:
        return int.from_bytes(b[:nbytes], "little")
    return (int.from_bytes(b[:nbytes], "little"), int.from_bytes(b[nbytes:2 * nbytes], "little"))


def _make_point_class(name: str, curve_id: int, p: int, n: int, h: int, Gx, Gy, parts: int, edwards: bool):
    base_fp = _Field(p, is_le=edwards)
    Fp = base_fp if parts == 1 else _Field2(base_fp)
    Fn = _Field(n, is_le=edwards)
    fp_bytes = 48 if p.bit_length() > 256 else 32
    zero_coord = 0 if parts == 1 else (0, 0)
    one_coord = 1 if parts == 1 else (1, 0)

    class Point:
        """Affine host handle of a curve point; every group operation runs on the GPU."""

        __slots__ = ("x", "y", "_inf", "_valid")
        CURVE_ID = curve_id
        NAME = name

        def __init__(self, x, y, _inf: bool = False, _valid: bool = False):
            """Like the reference's constructor (weierstrass.ts:695-704, edwards.ts:377-385) this does NOT validate
            curve or subgroup membership.  `_valid` is the host mirror of the reference's validityCache
            (weierstrass.ts:760-771): set by assertValidity / fromBytes / BASE / ZERO and inherited by the results
            of group operations on valid inputs.  pippenger only takes the endomorphism schedule on BLS12-381 G1
            when every input carries it (phi(P) = lambda*P needs the prime-order subgroup)."""
            if not Fp.isValid(x):
                raise ValueError("bad point coordinate x")
            # weierstrass.ts:698-701: a finite point with y = 0 is 2-torsion and is rejected by the constructor
            if not Fp.isValid(y) or (not edwards and not _inf and y == zero_coord):
                raise ValueError("bad point coordinate y")
            self.x, self.y, self._inf, self._valid = x, y, bool(_inf), bool(_valid)

        # -- noble's projective accessors (weierstrass.ts:687-704 / edwards.ts:370-385), Z = 1 ----
        @property
        def X(self):
            return self.x

        @property
        def Y(self):
            return one_coord if (self._inf and not edwards) else self.y

        @property
        def Z(self):
            return zero_coord if (self._inf and not edwards) else one_coord

        @staticmethod
        def fromAffine(p):
            """weierstrass.ts:711-718 / edwards.ts:396-402."""
            if not isinstance(p, dict) or not Fp.isValid(p.get("x")) or not Fp.isValid(p.get("y")):
                raise ValueError("invalid affine point")
            x, y = p["x"], p["y"]
            if not edwards and x == zero_coord and y == zero_coord:
                return Point.ZERO
            if edwards and x == 0 and y == 1:
                return Point.ZERO
            return Point(x, y)

        def toAffine(self):
            """weierstrass.ts:951-969 (ZERO -> (0,0)) / edwards.ts:595-609 (ZERO -> (0,1))."""
            return {"x": self.x, "y": self.y}

        def is0(self) -> bool:
            return self._inf

        def equals(self, other) -> bool:
            if not isinstance(other, Point):
                raise TypeError("Weierstrass Point expected" if not edwards else "EdwardsPoint expected")
            return self._inf == other._inf and self.x == other.x and self.y == other.y

        def negate(self):
            if self._inf:
                return self
            if edwards:  # -(x, y) = (-x, y)   edwards.ts:497-500
                return Point((-self.x) % p, self.y, False, self._valid)
            if parts == 1:  # weierstrass.ts:785-787
                return Point(self.x, (-self.y) % p, False, self._valid)
            return Point(self.x, ((-self.y[0]) % p, (-self.y[1]) % p), False, self._valid)

        def add(self, other):
            if not isinstance(other, Point):
                raise TypeError("Weierstrass Point expected" if not edwards else "EdwardsPoint expected")
            return _msm_points(Point, [self, other], [1, 1])

        def subtract(self, other):
            return self.add(other.negate())

        def double(self):
            return _msm_points(Point, [self], [2])

        def hasEvenY(self) -> bool:
            """weierstrass.ts:768-772 (prime base fields only: Fp2 has no isOdd)"""
            if edwards or parts != 1:
                raise ValueError("Field doesn't support isOdd")
            return (self.toAffine()["y"] & 1) == 0

        def assertValidity(self) -> None:
            """weierstrass.ts:752-771 / edwards.ts:461-480: on-curve + prime-order-subgroup check (both on the GPU:
            the curve equation through nmsm_points_on_curve, n*P == O through nmsm_points_torsion_free)."""
            if self._inf:
                if name.startswith("bls12_381") or edwards:  # allowInfinityPoint (bls12-381.ts) / Edwards identity
                    return
                raise ValueError("bad point: ZERO")
            if self._valid:
                return
            if points_on_curve(curve_id, self.to_packed(), 1)[0] != 1:
                raise ValueError("bad point: equation left != right")
            if h != 1 and torsion_free_packed(_ANY_POINT_ID.get(curve_id, curve_id), self.to_packed(), 1)[0] != 1:
                raise ValueError("bad point: not in prime-order subgroup")
            self._valid = True

        def multiply(self, scalar):
            """weierstrass.ts:900-907 / edwards.ts:555-564: 1 <= scalar < n."""
            if not Fn.isValidNot0(scalar):
                raise ValueError(
                    "invalid scalar: expected 1 <= sc < curve.n" if edwards else "invalid scalar: out of range"
                )
            return multiply_many(Point, [self], [scalar], unsafe=False)[0]

        def multiplyUnsafe(self, scalar):
            """weierstrass.ts:915-928 / edwards.ts:571-577: 0 <= scalar < n."""
            if not Fn.isValid(scalar):
                raise ValueError(
                    "invalid scalar: expected 0 <= sc < curve.n" if edwards else "invalid scalar: out of range"
                )
            return multiply_many(Point, [self], [scalar], unsafe=True)[0]

        def precompute(self, windowSize: int = 8, isLazy: bool = True):
            """weierstrass.ts:740-745 / edwards.ts:449-453: a cache hint in the reference; results never depend on
            it.  Here it marks the point for a device-resident multiplication table (nmsm_point_table_create; the
            GPU picks its own 16-bit windows, `windowSize` is only validated).  Like the reference's WeakMap cache
            (curve.ts:412,532) the table is built at the first multiply (isLazy) or now, and multiply_many /
            multiply / multiplyUnsafe of this point then take the table route."""
            _validate_w(windowSize, Fn.BITS)
            # curve.ts:776-781 setWindowSize sizes against the blinded path: ceil((bits + BLIND_BITS) / W) + 1 windows
            _validate_table_bytes((-(-(Fn.BITS + 128) // windowSize) + 1) * 2 ** (windowSize - 1), Fp.BYTES)
            _TABLE_MARKS.add((Point.CURVE_ID, self.x, self.y))
            if not isLazy:
                _table_for(Point, self)
            return self

        def clearCofactor(self):
            """weierstrass.ts:977-982 / edwards.ts:611-618 ([h]P; identity map for cofactor-1 curves)."""
            if h == 1:
                return self
            if h >= n:
                raise NotImplementedError("cofactor >= group order: outside the accelerated scalar range")
            return self.multiplyUnsafe(h)

        def isSmallOrder(self) -> bool:
            return self.is0() if h == 1 else self.clearCofactor().is0()

        def isTorsionFree(self) -> bool:
            """weierstrass.ts:971-975 / edwards.ts:584-586: n*P == O, evaluated as (n-1)*P + P on the GPU."""
            if self._inf:
                return True
            return self.multiplyUnsafe(n - 1).add(self).is0()

        def toBytes(self, isCompressed: bool = True) -> bytes:
            """Wire encodings of the reference: SEC1 (weierstrass.ts:541-563), Zcash flags for BLS12-381
            (bls12-381.ts:377-402), RFC 8032 for ed25519 (edwards.ts:620-628).  Pure byte packing on the host."""
            if edwards:
                b = bytearray(self.y.to_bytes(32, "little"))
                if self.x & 1:
                    b[31] |= 0x80
                return bytes(b)
            if name.startswith("bls12_381"):
                if parts != 1:
                    raise NotImplementedError("G2 wire format is not part of the accelerated path")
                if self._inf:
                    return bytes([0xC0 if isCompressed else 0x40]) + bytes(47 if isCompressed else 95)
                xb = bytearray(self.x.to_bytes(48, "big"))
                if isCompressed:
                    xb[0] |= 0x80 | (0x20 if (self.y * 2) // p else 0)
                    return bytes(xb)
                return bytes(xb) + self.y.to_bytes(48, "big")
            if parts != 1:
                raise NotImplementedError("Fp2 curves: wire format is not part of the accelerated path")
            if self._inf:
                raise ValueError("bad point: ZERO")
            xb = self.x.to_bytes(fp_bytes, "big")
            if isCompressed:
                return bytes([3 if self.y & 1 else 2]) + xb
            return b"\x04" + xb + self.y.to_bytes(fp_bytes, "big")

        @staticmethod
        def fromBytes(b: bytes, zip215: bool = False):
            """Point.fromBytes (weierstrass.ts:720-724; edwards.ts:405-436 `fromBytes(bytes, zip215 = false)`): decode on
            the GPU (nmsm_points_decode_ex), then the reference's validity check (subgroup membership where the
            cofactor is not 1).  Edwards: the default is the strict RFC 8032 decoding (y < p, no x = 0 with the sign
            bit set) exactly like the reference; zip215=True accepts the non-canonical encodings ZIP-215 allows.
            Like the reference, Edwards fromBytes does not run a subgroup check (edwards.ts:436)."""
            enc_len = {"secp256k1": 33, "bls12_381_G1": 48, "bls12_381_G2": 96, "ed25519": 32}.get(name)
            if enc_len is None or len(b) != enc_len:
                raise ValueError("bad point: got length %d, expected compressed=%s" % (len(b), enc_len))
            if not isinstance(zip215, bool):
                raise TypeError('"zip215" expected boolean')
            out, st = points_decode(curve_id, bytes(b), 1, zip215=bool(zip215) and edwards)
            if st[0] == 0:
                raise ValueError("bad point: is not on curve" if not edwards else "bad point: invalid y coordinate")
            P_ = Point.from_packed(out, 1 if st[0] == 2 else 0)
            if name in ("bls12_381_G1", "bls12_381_G2") and not P_.isTorsionFree():
                raise ValueError("bad point: not in prime-order subgroup")
            if not edwards:
                P_._valid = True  # decodePoint checks the equation, assertValidity the subgroup (weierstrass.ts:720-724)
            return P_

        @staticmethod
        def fromHex(hex_: str, zip215: bool = False):
            """weierstrass.ts:726-728 / edwards.ts:438-440: fromBytes(hexToBytes(hex))"""
            if not isinstance(hex_, str):
                raise TypeError("hex string expected, got " + type(hex_).__name__)
            try:
                raw = bytes.fromhex(hex_)
            except ValueError as e:
                raise ValueError("hex string expected, got non-hex character") from e
            return Point.fromBytes(raw, zip215) if edwards else Point.fromBytes(raw)

        def toHex(self, *args) -> str:
            return self.toBytes(*args).hex()

        def to_packed(self) -> bytes:
            return _coord_to_bytes(self.x, fp_bytes, parts) + _coord_to_bytes(self.y, fp_bytes, parts)

        @staticmethod
        def from_packed(b: bytes, is_inf: int, valid: bool = False):
            cb = fp_bytes * parts
            if is_inf:
                return Point.ZERO
            return Point(_coord_from_bytes(b[:cb], fp_bytes, parts), _coord_from_bytes(b[cb:2 * cb], fp_bytes, parts),
                         False, valid)

        def __repr__(self):
            return f"<{name}.Point {'ZERO' if self._inf else self.toAffine()}>"

    Point.__name__ = Point.__qualname__ = f"{name}_Point"
    Point.Fp = Fp
    Point.Fn = Fn
    Point.FP_BYTES = fp_bytes
    Point.PARTS = parts
    Point.POINT_BYTES = 2 * fp_bytes * parts
    Point.IS_EDWARDS = edwards
    Point.cofactor = h
    Point.BASE = Point(Gx, Gy, False, True)
    Point.ZERO = Point(0, 1, True, True) if edwards else Point(zero_coord, zero_coord, True, True)
    return Point


# Parameter blocks: SURVEY §8 a17 (src/secp256k1.ts:48-56, src/ed25519.ts:49-63, src/bn254.ts:80-90,207-223,
# src/bls12-381.ts:134-148,321-345).  tests/test_abi.py cross-checks them against the oracle's copies.
_BLS_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_BLS_N = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_BN_P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_BN_N = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


secp256k1 = _NS(
    Point=_make_point_class(
        "secp256k1", SECP256K1,
        2**256 - 2**32 - 977,
        0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141, 1,
        0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
        0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8, 1, False,
    )
)
ed25519 = _NS(
    Point=_make_point_class(
        "ed25519", ED25519,
        2**255 - 19,
        2**252 + 27742317777372353535851937790883648493, 8,
        0x216936D3CD6E53FEC0A4E231FDD6DC5C692CC7609525A7B2C9562D608F25D51A,
        0x6666666666666666666666666666666666666666666666666666666666666658, 1, True,
    )
)
bn254 = _NS(
    G1=_NS(Point=_make_point_class("bn254_G1", BN254_G1, _BN_P, _BN_N, 1, 1, 2, 1, False)),
    G2=_NS(
        Point=_make_point_class(
            "bn254_G2", BN254_G2, _BN_P, _BN_N,
            0x30644E72E131A029B85045B68181585E06CEECDA572A2489345F2299C0F9FA8D,
            (10857046999023057135944570762232829481370756359578518086990519993285655852781,
             11559732032986387107991004021392285783925812861821192530917403151452391805634),
            (8495653923123431417604973247489272438418190587263600148770280649306958101930,
             4082367875863433681332203403145435568316851327593401208105741076214120093531),
            2, False,
        )
    ),
)
bls12_381 = _NS(
    G1=_NS(
        Point=_make_point_class(
            "bls12_381_G1", BLS12_381_G1, _BLS_P, _BLS_N, 0x396C8C005555E1568C00AAAB0000AAAB,
            0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
            0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
            1, False,
        )
    ),
    G2=_NS(
        Point=_make_point_class(
            "bls12_381_G2", BLS12_381_G2, _BLS_P, _BLS_N,
            0x5D543A95414E7F1091D50792876A202CD91DE4547085ABAA68A205B2E5A7DDFA628F1CB4D9E82EF21537E293A6691AE1616EC6E786F0C70CF1C38E31C7238E5,
            (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
             0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
            (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
             0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
            2, False,
        )
    ),
)

CURVES = {
    "secp256k1": secp256k1.Point,
    "ed25519": ed25519.Point,
    "bn254_G1": bn254.G1.Point,
    "bn254_G2": bn254.G2.Point,
    "bls12_381_G1": bls12_381.G1.Point,
    "bls12_381_G2": bls12_381.G2.Point,
}


# ----------------------------------------------------------------------------------------------
# validation (exact messages of curve.ts:390-404, :875)
# ----------------------------------------------------------------------------------------------
TABLE_BYTES_MAX = 2 ** 31  # curve.ts:37


def validatePointCons(c) -> None:
    """curve.ts:259-272: the generic helpers dereference fromAffine / fromBytes / fromHex / BASE / ZERO / Fp / Fn of the
    point constructor; fail with a typed error up front instead of an attribute error later."""
    if not isinstance(c, type):
        raise TypeError('"Point" expected constructor, got type=' + type(c).__name__)
    for fn_name in ("fromAffine", "fromBytes", "fromHex"):
        if not callable(getattr(c, fn_name, None)):
            raise TypeError('"Point.%s" expected function' % fn_name)
    for obj_name in ("BASE", "ZERO"):
        if getattr(c, obj_name, None) is None:
            raise TypeError('"Point.%s" expected object' % obj_name)
    for f_name in ("Fp", "Fn"):
        f = getattr(c, f_name, None)
        if f is None or not isinstance(getattr(f, "ORDER", None), int) or not isinstance(getattr(f, "BITS", None), int):
            raise TypeError('"Point.%s" expected a field' % f_name)


def _validate_w(W, bits: int, lo: int = 1) -> None:
    """curve.ts:328-331 validateW"""
    if not (isinstance(W, int) and not isinstance(W, bool) and lo <= W <= bits):
        raise ValueError("invalid window size, expected [%d..%d], got W=%s" % (lo, bits, W))


def _validate_table_bytes(num_points: int, fp_bytes: int) -> None:
    """curve.ts:336-346 validateTableBytes: the reference refuses window sizes whose tables would need more than ~2 GiB
    of heap (4 coordinates of Fp.BYTES + 128 bytes of object overhead per point); same rule, same message."""
    nbytes = num_points * (4 * fp_bytes + 128)
    if nbytes > TABLE_BYTES_MAX:
        raise ValueError("invalid window size: table would need ~%d MiB, max %d MiB"
                         % (-(-nbytes // 2 ** 20), TABLE_BYTES_MAX // 2 ** 20))


def _validate_msm_points(points, c) -> None:
    if not isinstance(points, (list, tuple)):
        raise TypeError('"points" expected Array, got type=' + type(points).__name__)
    for i, p in enumerate(points):
        if not isinstance(p, c):
            raise ValueError("invalid point at index " + str(i))


def _validate_msm_scalars(scalars, fn) -> None:
    if not isinstance(scalars, (list, tuple)):
        raise ValueError("array of scalars expected")
    for i, s in enumerate(scalars):
        if not fn.isValid(s):
            raise ValueError("invalid scalar at index " + str(i))


def _pack_points(points: Sequence) -> bytes:
    return b"".join(p.to_packed() for p in points)


def _pack_scalars(scalars: Sequence[int]) -> bytes:
    return b"".join(s.to_bytes(32, "little") for s in scalars)


def _raise_mapped(e: NmsmError):
    # the C ABI already carries the reference's message text for point/scalar errors
    raise ValueError(str(e)) from e


# include/nmsm.h: NMSM_BLS12_381_G1 (4) assumes torsion-free points (GLV); NMSM_BLS12_381_G1_ANY (6) does not
# likewise NMSM_BLS12_381_G2 (5, psi split) / NMSM_BLS12_381_G2_ANY (7)
_ANY_POINT_ID = {4: 6, 5: 7}


def _all_valid(points) -> bool:
    return all(p._valid for p in points)


def _curve_id_for(c, points, assume_torsion_free=None) -> int:
    """Curve id of the C ABI for an MSM over `points`.  BLS12-381 G1 has a cofactor, and the endomorphism schedule of
    id 4 is only the group law on the prime-order subgroup; the reference's pippenger is the plain group law on ANY
    Point instance (its constructor / fromAffine do not validate, weierstrass.ts:695-718).  So id 4 is taken only
    when every input is known valid (assertValidity / fromBytes / BASE / results of arithmetic on such points), or
    when the caller vouches for it; otherwise the plain-window id 6."""
    if c.CURVE_ID not in _ANY_POINT_ID:
        return c.CURVE_ID
    ok = _all_valid(points) if assume_torsion_free is None else bool(assume_torsion_free)
    return c.CURVE_ID if ok else _ANY_POINT_ID[c.CURVE_ID]


def _msm_points(c, points, scalars, assume_torsion_free=None):
    cid = _curve_id_for(c, points, assume_torsion_free)
    out_xy, inf = msm_packed(cid, _pack_points(points), _pack_scalars(scalars), len(points))
    # the group generated by valid points consists of valid points
    return c.from_packed(out_xy, inf, _all_valid(points))


# ----------------------------------------------------------------------------------------------
# public surface
# ----------------------------------------------------------------------------------------------
def msm_packed(curve_id: int, pts: bytes, scalars: bytes, n: int):
    """Packed fast path: canonical little-endian affine points + 32-byte LE scalars -> (xy bytes, is_inf)."""
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    if pb <= 0:
        raise ValueError("unknown curve id")
    if len(pts) != n * pb or len(scalars) != n * 32:
        raise ValueError("arrays of points and scalars must have equal length")
    out = ctypes.create_string_buffer(pb)
    inf = ctypes.c_int(0)
    pbuf = ctypes.c_char_p(bytes(pts)) if n else None
    sbuf = ctypes.c_char_p(bytes(scalars)) if n else None
    rc = lib.nmsm_msm(curve_id, ctypes.cast(pbuf, ctypes.c_void_p), ctypes.cast(sbuf, ctypes.c_void_p), n,
                      ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw, inf.value


def msm_device(curve_id: int, d_pts: int, d_scalars: int, n: int):
    """Inputs already resident in device memory (raw device pointers, e.g. torch `tensor.data_ptr()`)."""
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    out = ctypes.create_string_buffer(pb)
    inf = ctypes.c_int(0)
    rc = lib.nmsm_msm_device(curve_id, ctypes.c_void_p(d_pts), ctypes.c_void_p(d_scalars), n,
                             ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw, inf.value


def msm_host_ptr(curve_id: int, h_pts: int, h_scalars: int, n: int):
    """End-to-end path on raw HOST pointers (e.g. pinned buffers): H2D + MSM + D2H inside the call."""
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    out = ctypes.create_string_buffer(pb)
    inf = ctypes.c_int(0)
    rc = lib.nmsm_msm(curve_id, ctypes.c_void_p(h_pts), ctypes.c_void_p(h_scalars), n,
                      ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw, inf.value


def pippenger(c, points, scalars, assume_torsion_free=None):
    """Drop-in for noble's pippenger (curve.ts:863-905): validates like the reference, runs on the GPU.
    With default arguments the result equals the reference's for EVERY Point instance: on BLS12-381 G1 the
    endomorphism schedule (id 4) is used only when all inputs are known members of the prime-order subgroup (the
    `_valid` bit: assertValidity / fromBytes / BASE / arithmetic on valid points), otherwise plain windows (id 6).
    `assume_torsion_free=True` lets a caller who knows its fromAffine-built points are subgroup members (e.g. an
    SRS) opt into the fast schedule; False forces the plain one."""
    validatePointCons(c)
    _validate_msm_points(points, c)
    _validate_msm_scalars(scalars, c.Fn)
    if len(points) != len(scalars):
        raise ValueError("arrays of points and scalars must have equal length")
    if len(points) == 0:
        return c.ZERO
    return _msm_points(c, points, scalars, assume_torsion_free)


def normalizeZ(c, points):
    """curve.ts:311-326: batch projective -> affine.  Host handles are already canonical affine (every
    GPU result is normalised on the device — k_mul_batch / k_table_mul share ONE inversion per warp, the device form
    of FpInvertBatch modular.ts:734-760), so this validates and returns fresh equal points.  Un-normalised values only
    exist as raw accumulators on the device side of the C ABI; their batch normalisation is normalize_accs
    (nmsm_accs_normalize)."""
    validatePointCons(c)
    _validate_msm_points(points, c)
    return [c.from_packed(p.to_packed(), 1 if p.is0() else 0, p._valid) for p in points]


def aggregate_points(c, points):
    """sum_i P_i (BLS aggregatePublicKeys / aggregateSignatures are exactly this, abstract/bls.ts:860,870):
    an MSM with unit scalars; every term lands in one bucket, which the balanced-segment accumulate and the
    tile stitching handle in parallel."""
    _validate_msm_points(points, c)
    if not points:
        return c.ZERO
    return _msm_points(c, points, [1] * len(points))


def mulAddUnsafe(c, points, scalars, allowOversized: bool = False):
    """curve.ts:820-836 — same value as the Strauss–Shamir walk, evaluated as a (small) MSM on the GPU.
    allowOversized replaces the `s < Fn.ORDER` check by the `Fn.ORDER^4` cap (curve.ts:829-830).  Oversized scalars
    must NOT be reduced mod ORDER (the reference uses them for torsion checks `ORDER*P == O` on points that may lie
    outside the prime-order subgroup), so s is cut into digits of Fn.BITS-1 bits, s = sum_j d_j 2^(b j), and the term
    becomes sum_j d_j * (2^(b j) * P) with every d_j and 2^b in the accelerated range — exact for any point."""
    validatePointCons(c)
    _validate_msm_points(points, c)
    if not isinstance(allowOversized, bool):
        raise TypeError('"allowOversized" expected boolean')
    if not allowOversized:
        _validate_msm_scalars(scalars, c.Fn)
    else:
        if not isinstance(scalars, (list, tuple)):
            raise ValueError("array of scalars expected")
        cap = c.Fn.ORDER ** 4
        for i, s_ in enumerate(scalars):
            if not (isinstance(s_, int) and not isinstance(s_, bool) and 0 <= s_ < cap):
                raise ValueError("invalid scalar at index " + str(i))
    if len(points) != len(scalars):
        raise ValueError("arrays of points and scalars must have equal length")
    if len(points) == 0:
        return c.ZERO
    if all(s_ < c.Fn.ORDER for s_ in scalars):
        return _msm_points(c, points, scalars)
    b = c.Fn.BITS - 1
    step, mask = 1 << b, (1 << b) - 1
    pts2, sc2 = [], []
    for P_, s_ in zip(points, scalars):
        Q = P_
        while True:
            pts2.append(Q)
            sc2.append(s_ & mask)
            s_ >>= b
            if s_ == 0:
                break
            Q = multiply_many(c, [Q], [step], unsafe=True)[0]  # 2^b * Q, plain double-and-add on any point
    return _msm_points(c, pts2, sc2)


def torsion_free_packed(curve_id: int, pts: bytes, n: int) -> bytes:
    """Batch Point.isTorsionFree (nmsm_points_torsion_free): one byte per point, 1 = n*P == O."""
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    if len(pts) != n * pb:
        raise ValueError("expected %d bytes per point" % pb)
    out = ctypes.create_string_buffer(max(1, n))
    rc = lib.nmsm_points_torsion_free(curve_id, ctypes.cast(ctypes.c_char_p(bytes(pts)), ctypes.c_void_p), n,
                                      ctypes.cast(out, ctypes.c_void_p))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw[:n]


class PointTable:
    """Device-resident multiplication table of ONE point (nmsm_point_table_create): d * 2^(16 j) * P."""

    def __init__(self, curve_id: int, point_xy: bytes):
        _lib.ensure_init()
        self._lib = _lib.load()
        self.curve_id = curve_id
        h = ctypes.c_uint64(0)
        try:
            _lib.check(self._lib.nmsm_point_table_create(curve_id, ctypes.cast(ctypes.c_char_p(bytes(point_xy)), ctypes.c_void_p),
                                                         ctypes.byref(h)))
        except NmsmError as e:
            _raise_mapped(e)
        self.handle = h.value

    def mul_batch(self, scalars: bytes, n: int, allow_zero: bool):
        pb = self._lib.nmsm_point_bytes(self.curve_id)
        if len(scalars) != n * 32:
            raise ValueError("expected 32 bytes per scalar")
        out = ctypes.create_string_buffer(max(1, n * pb))
        infs = ctypes.create_string_buffer(max(1, n))
        rc = self._lib.nmsm_point_table_mul_batch(self.handle, ctypes.cast(ctypes.c_char_p(bytes(scalars)), ctypes.c_void_p), n,
                                                  1 if allow_zero else 0, ctypes.cast(out, ctypes.c_void_p),
                                                  ctypes.cast(infs, ctypes.c_void_p))
        try:
            _lib.check(rc)
        except NmsmError as e:
            _raise_mapped(e)
        return out.raw[: n * pb], infs.raw[:n]

    def close(self):
        if self.handle:
            self._lib.nmsm_point_table_free(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# points marked by Point.precompute(), and the few most recently used device tables (36-107 MB each)
_TABLE_MARKS: set = set()
_TABLES: dict = {}
_TABLES_MAX = 4


def _table_for(c, point) -> PointTable:
    key = (c.CURVE_ID, point.x, point.y)
    t = _TABLES.pop(key, None)
    if t is None:
        t = PointTable(c.CURVE_ID, _pack_points([point]))
        while len(_TABLES) >= _TABLES_MAX:
            _TABLES.pop(next(iter(_TABLES))).close()
    _TABLES[key] = t  # most recently used last
    return t


def multiply_many(c, points, scalars, unsafe: bool = False) -> List:
    """Batch form of Point.multiply (unsafe=False: 1 <= k < n) / multiplyUnsafe (unsafe=True: 0 <= k < n).
    When every entry of `points` is the same precomputed point (Point.precompute) the device table route is used."""
    _validate_msm_points(points, c)
    if len(points) != len(scalars):
        raise ValueError("arrays of points and scalars must have equal length")
    for s in scalars:
        ok = c.Fn.isValid(s) if unsafe else c.Fn.isValidNot0(s)
        if not ok:
            if c.IS_EDWARDS:
                raise ValueError("invalid scalar: expected %d <= sc < curve.n" % (0 if unsafe else 1))
            raise ValueError("invalid scalar: out of range")
    n = len(points)
    if n == 0:
        return []
    p0 = points[0]
    if (c.CURVE_ID, p0.x, p0.y) in _TABLE_MARKS and not p0.is0() and all(q is p0 or q.equals(p0) for q in points):
        out_xy, infs = _table_for(c, p0).mul_batch(_pack_scalars(scalars), n, unsafe)
    else:
        out_xy, infs = mul_batch_packed(c.CURVE_ID, _pack_points(points), _pack_scalars(scalars), n, unsafe)
    pb = c.POINT_BYTES
    return [c.from_packed(out_xy[i * pb:(i + 1) * pb], infs[i], points[i]._valid) for i in range(n)]


def normalize_accs(curve_id: int, accs, n: int, on_device: bool = False):
    """nmsm_accs_normalize — the device form of normalizeZ (curve.ts:311-326): n raw accumulators (bytes, or a device
    pointer with on_device=True; nmsm_acc_bytes each) -> (packed canonical affine points, infinity flags), one field
    inversion per 32 points."""
    _lib.ensure_init()
    lib = _lib.load()
    pb, ab = lib.nmsm_point_bytes(curve_id), lib.nmsm_acc_bytes(curve_id)
    if pb <= 0:
        raise ValueError("unknown curve id")
    if not on_device and len(accs) != n * ab:
        raise ValueError("expected %d bytes per accumulator" % ab)
    out = ctypes.create_string_buffer(max(1, n * pb))
    infs = ctypes.create_string_buffer(max(1, n))
    src = ctypes.c_void_p(accs) if on_device else ctypes.cast(ctypes.c_char_p(bytes(accs)), ctypes.c_void_p)
    _lib.check(lib.nmsm_accs_normalize(curve_id, src, 1 if on_device else 0, n, ctypes.cast(out, ctypes.c_void_p),
                                       ctypes.cast(infs, ctypes.c_void_p)))
    return out.raw[: n * pb], infs.raw[:n]


def mul_batch_packed(curve_id: int, pts: bytes, scalars: bytes, n: int, allow_zero: bool):
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    if len(pts) != n * pb or len(scalars) != n * 32:
        raise ValueError("arrays of points and scalars must have equal length")
    out = ctypes.create_string_buffer(max(1, n * pb))
    infs = ctypes.create_string_buffer(max(1, n))
    rc = lib.nmsm_mul_batch(curve_id, ctypes.cast(ctypes.c_char_p(bytes(pts)), ctypes.c_void_p),
                            ctypes.cast(ctypes.c_char_p(bytes(scalars)), ctypes.c_void_p), n, 1 if allow_zero else 0,
                            ctypes.cast(out, ctypes.c_void_p), ctypes.cast(infs, ctypes.c_void_p))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw[: n * pb], infs.raw[:n]


class PointSet:
    """Device-resident, validated and prepared point array (nmsm_points_upload): the fixed-base fast path."""

    def __init__(self, curve_id: int, pts: bytes, n: int):
        _lib.ensure_init()
        self._lib = _lib.load()
        self.curve_id, self.n = curve_id, n
        h = ctypes.c_uint64(0)
        try:
            _lib.check(self._lib.nmsm_points_upload(curve_id, ctypes.cast(ctypes.c_char_p(bytes(pts)), ctypes.c_void_p),
                                                    n, ctypes.byref(h)))
        except NmsmError as e:
            _raise_mapped(e)
        self.handle = h.value
        self.table_bits, self.table_levels = 0, 1

    def precompute(self, window_bits: int = 0):
        """Build the fixed-base table 2^(c*j) * P_i on the device (nmsm_points_precompute; the analogue of
        Point.precompute, curve.ts:532-577, for a whole set).  Returns (window bits, levels)."""
        c, lv = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(self._lib.nmsm_points_precompute(self.handle, int(window_bits), ctypes.byref(c), ctypes.byref(lv)))
        self.table_bits, self.table_levels = c.value, lv.value
        return c.value, lv.value

    def msm(self, scalars: bytes, n: int):
        pb = self._lib.nmsm_point_bytes(self.curve_id)
        out = ctypes.create_string_buffer(pb)
        inf = ctypes.c_int(0)
        rc = self._lib.nmsm_msm_points(self.handle, ctypes.cast(ctypes.c_char_p(bytes(scalars)), ctypes.c_void_p), n,
                                       ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf))
        try:
            _lib.check(rc)
        except NmsmError as e:
            _raise_mapped(e)
        return out.raw, inf.value

    def close(self):
        if self.handle:
            self._lib.nmsm_points_free(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interleavedMSMUnsafe(c, points, windowSize: int = 4, precompute: bool = True, assume_torsion_free=None):
    """curve.ts:937-959: captures a FIXED point set once and returns `scalars -> sum s_i*P_i`.  Here the
    captured state is the device-resident prepared array plus (precompute=True, like the reference's per-point
    tables built at capture time) the fixed-base table of nmsm_points_precompute; `windowSize` is accepted for
    signature compatibility (the GPU schedule picks its own window).  Fewer scalars than points are zero-padded.
    BLS12-381 G1: the endomorphism id is used only for sets of known-valid points (see pippenger)."""
    validatePointCons(c)
    _validate_w(windowSize, c.Fn.BITS, 2)
    _validate_msm_points(points, c)
    _validate_table_bytes(len(points) * 2 ** (windowSize - 2), c.Fp.BYTES)  # curve.ts:947
    n = len(points)
    valid = _all_valid(points)
    ps = PointSet(_curve_id_for(c, points, assume_torsion_free), _pack_points(points), n) if n else None
    if ps is not None and precompute:
        ps.precompute(0)

    def run(scalars):
        _validate_msm_scalars(scalars, c.Fn)
        if len(scalars) > n:
            raise ValueError("array of scalars must not be larger than array of points")
        if n == 0:
            return c.ZERO
        padded = list(scalars) + [0] * (n - len(scalars))
        out, inf = ps.msm(_pack_scalars(padded), n)
        return c.from_packed(out, inf, valid)

    return run


def ed25519_verify_batch(signatures, messages, public_keys, z: bytes | None = None):
    """Batch form of ed25519.verify (edwards.ts:942-989, ZIP-215 decoding as ed25519's default): returns
    (ok, bad_index).  ok is True iff every signature would be accepted individually (up to 2^-128);
    bad_index is the first signature rejected without the group equation (undecodable point / s >= l), else -1.
    `z`: optional n*16 bytes of caller randomness (default os.urandom)."""
    import os
    import struct

    n = len(signatures)
    if len(messages) != n or len(public_keys) != n:
        raise ValueError("signatures, messages and public keys must have equal length")
    for s, p in zip(signatures, public_keys):
        if len(s) != 64 or len(p) != 32:
            raise ValueError("signature expected 64 bytes, publicKey 32 bytes")
    if z is None:
        z = os.urandom(16 * n)
    if len(z) != 16 * n:
        raise ValueError("z must hold 16 bytes per signature")
    offs = [0]
    for m in messages:
        offs.append(offs[-1] + len(m))
    blob = b"".join(bytes(m) for m in messages)
    return ed25519_verify_batch_packed(b"".join(signatures), b"".join(public_keys), blob,
                                       struct.pack("<%dQ" % (n + 1), *offs), n, bytes(z))


def ed25519_verify_batch_packed(sigs: bytes, pks: bytes, msgs: bytes, msg_off: bytes, n: int, z: bytes):
    """The C-ABI call itself (nmsm_ed25519_verify_batch) on caller-packed arrays: n*64 signature bytes, n*32 key bytes,
    the concatenated messages with n+1 little-endian u64 offsets, n*16 bytes of randomness."""
    if len(sigs) != 64 * n or len(pks) != 32 * n or len(msg_off) != 8 * (n + 1) or len(z) != 16 * n:
        raise ValueError("packed arrays do not match n")
    _lib.ensure_init()
    lib = _lib.load()
    ok = ctypes.c_int(0)
    bad = ctypes.c_longlong(-1)
    cp = lambda b: ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p)  # noqa: E731
    rc = lib.nmsm_ed25519_verify_batch(cp(sigs), cp(pks), cp(msgs or b"\0"), cp(msg_off), n, cp(z), ctypes.byref(ok),
                                       ctypes.byref(bad))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return bool(ok.value), int(bad.value)


DECODE_ZIP215 = 1  # include/nmsm.h NMSM_DECODE_ZIP215


def points_on_curve(curve_id: int, pts: bytes, n: int) -> bytes:
    """Batch curve-equation check (nmsm_points_on_curve; weierstrass.ts:617-624 isValidXY, edwards.ts:461-480): one
    byte per point, 1 = coordinates in range and on the curve (the affine identity encoding counts as on-curve)."""
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    if len(pts) != n * pb:
        raise ValueError("expected %d bytes per point" % pb)
    out = ctypes.create_string_buffer(max(1, n))
    rc = lib.nmsm_points_on_curve(curve_id, ctypes.cast(ctypes.c_char_p(bytes(pts)), ctypes.c_void_p), n,
                                  ctypes.cast(out, ctypes.c_void_p))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw[:n]


def points_decode(curve_id: int, encodings: bytes, n: int, zip215: bool = False):
    """Batched `fromBytes` decode step on the GPU (secp256k1 SEC1-33, BLS12-381 G1 Zcash-48 / G2 Zcash-96, ed25519-32):
    returns (packed points, status bytes: 0 invalid / 1 point / 2 infinity).  ed25519: strict RFC 8032 by default
    (the reference's `fromBytes(bytes, zip215 = false)`), ZIP-215 acceptance with zip215=True."""
    _lib.ensure_init()
    lib = _lib.load()
    pb = lib.nmsm_point_bytes(curve_id)
    out = ctypes.create_string_buffer(max(1, n * pb))
    st = ctypes.create_string_buffer(max(1, n))
    rc = lib.nmsm_points_decode_ex(curve_id, ctypes.cast(ctypes.c_char_p(bytes(encodings)), ctypes.c_void_p), n,
                                   DECODE_ZIP215 if zip215 else 0,
                                   ctypes.cast(out, ctypes.c_void_p), ctypes.cast(st, ctypes.c_void_p))
    try:
        _lib.check(rc)
    except NmsmError as e:
        _raise_mapped(e)
    return out.raw[: n * pb], st.raw[:n]


def last_timing():
    """(dict of per-kernel ms, PlanInfo) of the last MSM call when profiling is enabled."""
    lib = _lib.load()
    ms = (ctypes.c_float * _lib.TIMING_SLOTS)()
    info = PlanInfo()
    lib.nmsm_last_timing(ms, ctypes.byref(info))
    return {k: float(ms[i]) for i, k in enumerate(TIMING_NAMES)}, info


def set_profiling(enabled: bool) -> None:
    _lib.load().nmsm_set_profiling(1 if enabled else 0)


def set_window_bits(c: int) -> int:
    return _lib.load().nmsm_set_window_bits(c)


def set_window_groups(groups: int) -> int:
    """Force the number of window groups the MSM pipeline overlaps (0 = automatic); results never depend on it."""
    return _lib.load().nmsm_set_window_groups(groups)


def bench_modmul(field: int, blocks_per_sm: int = 8, threads: int = 128, iters: int = 2000, ilp: int = 1) -> float:
    _lib.ensure_init()
    return float(_lib.load().nmsm_bench_modmul(field, blocks_per_sm, threads, iters, ilp))
