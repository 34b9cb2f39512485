"""CPU ORACLE (test infrastructure only) — Python-int restatement of noble-curves' scalar-mult / MSM path.

THIS IS NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import it.  The product path
(`noble-curves_b200/nmsm`) never imports anything under `oracle/`.

Parity status: PINNED.  All arithmetic on this path lives in `/root/reference/src` on native
BigInt (no third-party dependency), and this restatement is checked against the reference's own
golden vectors (tests/golden/*.json, extracted by tests/golden/make_golden.py from
`/root/reference/test/vectors`): secp256k1 privates-2 / points.json / endomorphism.json,
BLS12-381 zkcrypto i*G tables (G1+G2), bn254 EIP-196 dumps + seda vectors, ed25519 RFC 8032
`sk:pk` vectors.  See tests/test_oracle_golden.py.

Python `int` has the semantics of JS `BigInt` for `* + - >> &`; JS `%` truncates, which is why the
reference wraps it in `mod()` (src/abstract/modular.ts:50-54); Python `%` is already floor-mod.

Every function cites the reference file:line (relative to /root/reference/) it follows.
"""
from __future__ import annotations

import hashlib
import os
from typing import Callable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# utils.ts
# --------------------------------------------------------------------------------------


def bit_len(n: int) -> int:
    """src/utils.ts:659 bitLen."""
    return n.bit_length()


def bit_mask(n: int) -> int:
    """src/utils.ts:725 bitMask."""
    return (1 << n) - 1


# --------------------------------------------------------------------------------------
# modular.ts — prime field (plain `(a*b) % p`, no Montgomery: modular.ts:888-1038)
# --------------------------------------------------------------------------------------


def invert(number: int, modulo: int) -> int:
    """src/abstract/modular.ts:159-182 — extended Euclid; throws on 0 / non-invertible."""
    if number == 0:
        raise ValueError("invert: expected non-zero number")
    if modulo <= 0:
        raise ValueError("invert: expected positive modulus, got " + str(modulo))
    a = number % modulo
    b = modulo
    x, y, u, v = 0, 1, 1, 0
    while a != 0:
        q = b // a
        r = b - a * q
        m = x - u * q
        n = y - v * q
        b, a, x, y, u, v = a, r, u, v, m, n
    if b != 1:
        raise ValueError("invert: does not exist")
    return x % modulo


class Field:
    """src/abstract/modular.ts:888-1038 `_Field` — only the ops on the hot path."""

    def __init__(self, order: int, bits: Optional[int] = None, is_le: bool = False):
        self.ORDER = order
        self.BITS = bits if bits is not None else order.bit_length()
        self.BYTES = (self.BITS + 7) // 8
        self.isLE = is_le
        self.ZERO = 0
        self.ONE = 1

    def create(self, num: int) -> int:
        return num % self.ORDER

    def isValid(self, num) -> bool:  # modular.ts:925
        return isinstance(num, int) and not isinstance(num, bool) and 0 <= num < self.ORDER

    def isValidNot0(self, num) -> bool:
        return self.isValid(num) and num != 0

    def is0(self, num: int) -> bool:
        return num == 0

    def eql(self, a: int, b: int) -> bool:
        return a == b

    def neg(self, num: int) -> int:  # modular.ts:940
        return (-num) % self.ORDER

    def sqr(self, num: int) -> int:  # modular.ts:947
        return (num * num) % self.ORDER

    def add(self, a: int, b: int) -> int:  # modular.ts:950
        return (a + b) % self.ORDER

    def sub(self, a: int, b: int) -> int:  # modular.ts:953
        return (a - b) % self.ORDER

    def mul(self, a: int, b: int) -> int:  # modular.ts:956
        return (a * b) % self.ORDER

    # the N-variants are unreduced in the reference (modular.ts:967-978); the reduced results of
    # every consumer are identical, so the oracle keeps them reduced.
    addN = add
    subN = sub
    mulN = mul

    def inv(self, num: int) -> int:  # modular.ts:980 -> :159
        return invert(num, self.ORDER)

    def pow(self, num: int, power: int) -> int:  # modular.ts:72-116 (result only)
        if power < 0:
            raise ValueError("invalid exponent, negatives unsupported")
        return pow(num, power, self.ORDER)

    def toBytes(self, num: int) -> bytes:
        return num.to_bytes(self.BYTES, "little" if self.isLE else "big")

    def fromBytes(self, b: bytes) -> int:
        n = int.from_bytes(b, "little" if self.isLE else "big")
        if not self.isValid(n):
            raise ValueError("invalid field element: outside of range 0..ORDER")
        return n


def FpInvertBatch(F, nums: Sequence, pass_zero: bool = False) -> list:
    """src/abstract/modular.ts:734-760 — Montgomery's trick; zeros are skipped (→ None / 0)."""
    inverted: list = [F.ZERO if pass_zero else None] * len(nums)
    acc = F.ONE
    for i, num in enumerate(nums):
        if F.is0(num):
            continue
        inverted[i] = acc
        acc = F.mul(acc, num)
    inv_acc = F.inv(acc)
    for i in range(len(nums) - 1, -1, -1):
        num = nums[i]
        if F.is0(num):
            continue
        inverted[i] = F.mul(inv_acc, inverted[i])
        inv_acc = F.mul(inv_acc, num)
    return inverted


# --------------------------------------------------------------------------------------
# tower.ts — Fp2 = Fp[u]/(u^2+1) (both bn254 and bls12-381 use u^2 = -1)
# --------------------------------------------------------------------------------------


class Field2:
    """src/abstract/tower.ts:305-561 `_Field2`. Elements are (c0, c1) tuples."""

    def __init__(self, Fp: Field):
        self.Fp = Fp
        self.ORDER = Fp.ORDER * Fp.ORDER
        self.BITS = self.ORDER.bit_length()
        self.BYTES = 2 * Fp.BYTES
        self.ZERO = (0, 0)
        self.ONE = (1, 0)

    def isValid(self, num) -> bool:
        return (
            isinstance(num, tuple) and len(num) == 2 and self.Fp.isValid(num[0]) and self.Fp.isValid(num[1])
        )

    def is0(self, num) -> bool:
        return num[0] == 0 and num[1] == 0

    def eql(self, a, b) -> bool:
        return a[0] == b[0] and a[1] == b[1]

    def neg(self, a):
        return (self.Fp.neg(a[0]), self.Fp.neg(a[1]))

    def add(self, a, b):  # tower.ts:393-404
        return (self.Fp.add(a[0], b[0]), self.Fp.add(a[1], b[1]))

    def sub(self, a, b):  # tower.ts:405-418
        return (self.Fp.sub(a[0], b[0]), self.Fp.sub(a[1], b[1]))

    def mul(self, a, rhs):  # tower.ts:420-431
        Fp = self.Fp
        if isinstance(rhs, int):
            return (Fp.mul(a[0], rhs), Fp.mul(a[1], rhs))
        c0, c1 = a
        r0, r1 = rhs
        t1 = Fp.mul(c0, r0)
        t2 = Fp.mul(c1, r1)
        o0 = Fp.sub(t1, t2)
        o1 = Fp.sub(Fp.mul(Fp.add(c0, c1), Fp.add(r0, r1)), Fp.add(t1, t2))
        return (o0, o1)

    def sqr(self, a):  # tower.ts:432-438
        Fp = self.Fp
        c0, c1 = a
        x = Fp.add(c0, c1)
        y = Fp.sub(c0, c1)
        z = Fp.add(c0, c0)
        return (Fp.mul(x, y), Fp.mul(z, c1))

    addN = add
    subN = sub
    mulN = mul

    def inv(self, num):  # tower.ts:458-475
        Fp = self.Fp
        a, b = num
        factor = Fp.inv(Fp.create(a * a + b * b))
        return (Fp.mul(factor, Fp.create(a)), Fp.mul(factor, Fp.create(-b)))

    def pow(self, num, power: int):
        res = self.ONE
        base = num
        while power > 0:
            if power & 1:
                res = self.mul(res, base)
            base = self.sqr(base)
            power >>= 1
        return res

    def sqrt(self, num):
        """tower.ts:476-498 (complex method, Fp_NONRESIDUE = -1); raises ValueError('Cannot find square root').
        The base-field root is modular.ts sqrt3mod4 (p = 3 mod 4 for BLS12-381 and bn254)."""
        Fp = self.Fp
        p = Fp.ORDER
        assert p % 4 == 3

        def fsqrt(n):
            r = pow(n, (p + 1) // 4, p)
            if r * r % p != n % p:
                raise ValueError("Cannot find square root")
            return r

        def legendre(n):
            t = pow(n, (p - 1) // 2, p)
            return -1 if t == p - 1 else t

        nonres = p - 1
        div2 = pow(2, -1, p)
        c0, c1 = num
        if c1 == 0:
            if legendre(c0) == 1:
                return (fsqrt(c0), 0)
            return (0, fsqrt(c0 * pow(nonres, -1, p) % p))
        a = fsqrt((c0 * c0 - c1 * c1 * nonres) % p)
        d = (a + c0) * div2 % p
        if legendre(d) == -1:
            d = (d - a) % p
        a0 = fsqrt(d)
        cand = (a0, c1 * div2 % p * pow(a0, -1, p) % p)
        if self.sqr(cand) != (c0 % p, c1 % p):
            raise ValueError("Cannot find square root")
        x1, x2 = cand, self.neg(cand)
        (re1, im1), (re2, im2) = x1, x2
        if im1 > im2 or (im1 == im2 and re1 > re2):
            return x1
        return x2


# --------------------------------------------------------------------------------------
# curve.ts — recoding helpers and the generic algorithms
# --------------------------------------------------------------------------------------

FW_WINDOW = 5  # src/abstract/curve.ts:33
BLIND_BITS = 128  # src/abstract/curve.ts:29
BLIND_BYTES = 16


def validate_msm_points(points, c) -> None:
    """src/abstract/curve.ts:390-395."""
    if not isinstance(points, (list, tuple)):
        raise TypeError('"points" expected Array')
    for i, p in enumerate(points):
        if not isinstance(p, c):
            raise ValueError("invalid point at index " + str(i))


def validate_msm_scalars(scalars, field: Field, max_scalar: Optional[int] = None) -> None:
    """src/abstract/curve.ts:398-404."""
    if not isinstance(scalars, (list, tuple)):
        raise ValueError("array of scalars expected")
    for i, s in enumerate(scalars):
        if max_scalar is None:
            ok = field.isValid(s)
        else:
            ok = isinstance(s, int) and not isinstance(s, bool) and 0 <= s < max_scalar
        if not ok:
            raise ValueError("invalid scalar at index " + str(i))


def normalizeZ(c, points: list) -> list:
    """src/abstract/curve.ts:311-326."""
    validate_msm_points(points, c)
    inverted = FpInvertBatch(c.Fp, [p.Z for p in points])
    return [c.fromAffine(p.toAffine(inverted[i])) for i, p in enumerate(points)]


def odd_multiples(p, size: int) -> list:
    """src/abstract/curve.ts:420-425."""
    dbl = p.double()
    t = [p]
    for j in range(1, size):
        t.append(t[j - 1].add(dbl))
    return t


def wnaf_digits(n: int, W: int) -> List[int]:
    """src/abstract/curve.ts:431-447."""
    size = 2**W
    half = size // 2
    mask = size - 1
    d: List[int] = []
    while n > 0:
        w = 0
        if n & 1:
            w = n & mask
            if w >= half:
                w -= size
            n -= w
        d.append(w)
        n >>= 1
    return d


def signed_window_digits(n: int, W: int, windows: int) -> List[int]:
    """src/abstract/curve.ts:454-472."""
    size = 2**W
    half = size // 2
    mask = size - 1
    d: List[int] = []
    for _ in range(windows):
        v = n & mask
        n >>= W
        if v > half:
            v -= size
            n += 1
        d.append(v)
    if n != 0:
        raise ValueError("invalid wnaf")
    return d


def wnaf_walk(zero, tables: list, digits: List[List[int]]):
    """src/abstract/curve.ts:479-498."""
    mx = 0
    for d in digits:
        mx = max(mx, len(d))
    acc = zero
    for bit in range(mx - 1, -1, -1):
        if bit != mx - 1:
            acc = acc.double()
        for i in range(len(digits)):
            w = digits[i][bit] if bit < len(digits[i]) else 0
            if w:
                item = tables[i][(abs(w) - 1) >> 1]
                acc = acc.add(item.negate() if w < 0 else item)
    return acc


def mulAddUnsafe(c, points: list, scalars: list, allow_oversized: bool = False):
    """src/abstract/curve.ts:820-836 — Strauss–Shamir, width-4 wNAF."""
    validate_msm_points(points, c)
    validate_msm_scalars(scalars, c.Fn, c.Fn.ORDER**4 if allow_oversized else None)
    if len(points) != len(scalars):
        raise ValueError("arrays of points and scalars must have equal length")
    tables = [odd_multiples(p, 4) for p in points]
    digits = [wnaf_digits(n, 4) for n in scalars]
    return wnaf_walk(c.ZERO, tables, digits)


def pippenger_window(n_points: int) -> int:
    """src/abstract/curve.ts:879-883 — window size rule."""
    wbits = bit_len(n_points)
    window = 1
    if wbits > 12:
        window = wbits - 3
    elif wbits > 4:
        window = wbits - 2
    elif wbits > 0:
        window = 2
    return window


def pippenger(c, points: list, scalars: list):
    """src/abstract/curve.ts:863-905 — unsigned-window bucket MSM, MSB→LSB, running-sum reduce."""
    fieldN = c.Fn
    validate_msm_points(points, c)
    validate_msm_scalars(scalars, fieldN)
    plength = len(points)
    slength = len(scalars)
    if plength != slength:
        raise ValueError("arrays of points and scalars must have equal length")
    zero = c.ZERO
    if plength == 0:
        return zero
    window = pippenger_window(plength)
    MASK = bit_mask(window)
    buckets = [zero] * (MASK + 1)
    last_bits = ((fieldN.BITS - 1) // window) * window
    total = zero
    for i in range(last_bits, -1, -window):
        for k in range(len(buckets)):
            buckets[k] = zero
        for j in range(slength):
            wb = (scalars[j] >> i) & MASK
            buckets[wb] = buckets[wb].add(points[j])
        resI = zero
        sumI = zero
        for j in range(len(buckets) - 1, 0, -1):
            sumI = sumI.add(buckets[j])
            resI = resI.add(sumI)
        total = total.add(resI)
        if i != 0:
            for _ in range(window):
                total = total.double()
    return total


def interleavedMSMUnsafe(c, points: list, window_size: int) -> Callable:
    """src/abstract/curve.ts:937-959."""
    fieldN = c.Fn
    if not (isinstance(window_size, int) and 2 <= window_size <= fieldN.BITS):
        raise ValueError("invalid window size")
    validate_msm_points(points, c)
    tables = [odd_multiples(p, 2 ** (window_size - 2)) for p in points]

    def run(scalars: list):
        validate_msm_scalars(scalars, fieldN)
        if len(scalars) > len(points):
            raise ValueError("array of scalars must not be larger than array of points")
        return wnaf_walk(c.ZERO, tables, [wnaf_digits(n, window_size) for n in scalars])

    return run


class ScalarMultiplier:
    """src/abstract/curve.ts:527-790 (the arithmetic paths; WeakMap caches become dicts keyed by id)."""

    def __init__(self, Point, random_bytes: Optional[Callable[[int], bytes]] = os.urandom):
        self.Point = Point
        self.BASE = Point.BASE
        self.ZERO = Point.ZERO
        self.randomBytes = random_bytes
        self.bits = Point.Fn.BITS
        self._window_sizes: dict = {}
        self._precomputes: dict = {}
        self._base_can_be_blinded: Optional[bool] = None

    # curve.ts:414-417 / :776-790
    def get_window_size(self, P) -> int:
        return self._window_sizes.get(id(P), 1)

    def setWindowSize(self, point, W: int) -> None:
        if not (isinstance(W, int) and 1 <= W <= self.bits):
            raise ValueError("invalid window size")
        self._window_sizes[id(point)] = W
        self._precomputes.pop(id(point), None)

    def hasWindowSize(self, point) -> bool:
        return self.get_window_size(point) != 1

    def build_wnaf_table(self, point, W: int, bits: int) -> dict:
        """curve.ts:560-577."""
        windows = -(-bits // W) + 1
        half = 2 ** (W - 1)
        comp = []
        base = point
        for _ in range(windows):
            acc = base
            for _i in range(half):
                comp.append(acc)
                acc = acc.add(base)
            base = comp[-1].double()
        return {"W": W, "bits": bits, "windows": windows, "comp": comp}

    def wnaf_cached_ct(self, pre: dict, n: int):
        """curve.ts:588-606 — returns (p, f)."""
        W, windows, comp = pre["W"], pre["windows"], pre["comp"]
        half = 2 ** (W - 1)
        digits = signed_window_digits(n, W, windows)
        p = self.ZERO
        f = self.BASE
        for w in range(windows):
            digit = digits[w]
            start = w * half
            idx = abs(digit) - 1
            sel = comp[start]
            for i in range(1, half):
                if i == idx:
                    sel = comp[start + i]
            neg = sel.negate()
            if digit == 0:
                f = f.add(comp[start])
            else:
                p = p.add(neg if digit < 0 else sel)
        return p, f

    def get_wnaf_precomputes(self, W: int, point, bits: int, transform=None) -> dict:
        """curve.ts:611-629."""
        entries = self._precomputes.setdefault(id(point), [])
        for e in entries:
            if e["W"] == W and e["bits"] == bits:
                return e
        comp = self.build_wnaf_table(point, W, bits)
        if transform is not None:
            comp = dict(comp, comp=transform(comp["comp"]))
        entries.append(comp)
        return comp

    def validate_mul_input(self, point, scalar) -> None:
        """curve.ts:636-641."""
        if not isinstance(point, self.Point):
            raise TypeError('"point" expected Point instance')
        if not (isinstance(scalar, int) and 1 <= scalar < self.Point.Fn.ORDER):
            raise ValueError("invalid scalar")

    def run_ct(self, point, n: int, bits: int, transform=None):
        """curve.ts:647-656."""
        W = self.get_window_size(point)
        if W == 1:
            return self.fixed_window_ct(point, n, bits)
        return self.wnaf_cached_ct(self.get_wnaf_precomputes(W, point, bits, transform), n)

    def mulCT(self, point, scalar: int, transform=None):
        """curve.ts:658-661."""
        self.validate_mul_input(point, scalar)
        return self.run_ct(point, scalar, self.bits, transform)

    def mulCTBlinded(self, point, scalar: int, transform=None, blind_bytes: Optional[bytes] = None):
        """curve.ts:663-690."""
        self.validate_mul_input(point, scalar)
        if self.randomBytes is None:
            raise ValueError("randomBytes is required for scalar blinding")
        bits = self.Point.Fn.BITS + BLIND_BITS
        blind = bytearray(blind_bytes if blind_bytes is not None else self.randomBytes(BLIND_BYTES))
        if len(blind) != BLIND_BYTES:
            raise ValueError("randomBytes returned invalid byte array")
        blind[0] = (blind[0] & 0x3F) | 0x80
        n = scalar + int.from_bytes(bytes(blind), "big") * self.Point.Fn.ORDER
        return self.run_ct(point, n, bits, transform)

    def fixed_window_ct(self, point, n: int, bits: int):
        """curve.ts:707-729."""
        W = FW_WINDOW
        size = 1 << W
        mask = bit_mask(W)
        table = [None] * size
        table[0] = self.ZERO
        for i in range(1, size):
            table[i] = table[i - 1].add(point)
        windows = -(-bits // W)
        acc = self.ZERO
        for window in range(windows - 1, -1, -1):
            if window != windows - 1:
                for _ in range(W):
                    acc = acc.double()
            digit = (n >> (window * W)) & mask
            sel = table[0]
            for i in range(1, size):
                if i == digit:
                    sel = table[i]
            acc = acc.add(sel)
        return acc, acc

    def should_blind(self, point, cofactor: int) -> bool:
        """curve.ts:731-739."""
        if self.randomBytes is None:
            return False
        if cofactor == 1:
            return True
        if point is not self.BASE:
            return False
        if self._base_can_be_blinded is None:
            self._base_can_be_blinded = self.mulUnsafe(self.BASE, self.Point.Fn.ORDER).is0()
        return self._base_can_be_blinded

    def mulSecret(self, point, scalar: int, cofactor: int, transform=None):
        """curve.ts:741-750."""
        if self.should_blind(point, cofactor):
            return self.mulCTBlinded(point, scalar, transform)
        return self.mulCT(point, scalar, transform)

    def mulUnsafe(self, point, scalar: int, transform=None):
        """curve.ts:752-770."""
        if not isinstance(point, self.Point):
            raise TypeError('"point" expected Point instance')
        if not (isinstance(scalar, int) and scalar >= 0):
            raise ValueError("invalid scalar")
        W = self.get_window_size(point)
        if W == 1 or scalar >= self.Point.Fn.ORDER:
            return mulAddUnsafe(self.Point, [point], [scalar], True)
        pre = self.get_wnaf_precomputes(W, point, self.bits, transform)
        return self.wnaf_cached_ct(pre, scalar)[0]


# --------------------------------------------------------------------------------------
# weierstrass.ts — short Weierstrass, homogeneous projective, RCB complete formulas
# --------------------------------------------------------------------------------------


def div_nearest(num: int, den: int) -> int:
    """src/abstract/weierstrass.ts:106 — JS bigint `/` truncates toward zero."""
    half = den // 2  # den > 0 in all callers; JS: (num>=0 ? den : -den)/2n truncates → ±(den//2)
    adj = num + (half if num >= 0 else -half)
    q = abs(adj) // den
    return q if adj >= 0 else -q


def split_endo_scalar(k: int, basis, n: int):
    """src/abstract/weierstrass.ts:121-148 `_splitEndoScalar`."""
    if not (0 <= k < n):
        raise ValueError("expected valid scalar: 0 <= n < " + str(n))
    (a1, b1), (a2, b2) = basis
    c1 = div_nearest(b2 * k, n)
    c2 = div_nearest(-b1 * k, n)
    k1 = k - c1 * a1 - c2 * a2
    k2 = -c1 * b1 - c2 * b2
    k1neg = k1 < 0
    k2neg = k2 < 0
    if k1neg:
        k1 = -k1
    if k2neg:
        k2 = -k2
    MAX_NUM = bit_mask(-(-bit_len(n) // 2)) + 1
    if k1 < 0 or k1 >= MAX_NUM or k2 < 0 or k2 >= MAX_NUM:
        raise ValueError("splitScalar (endomorphism): failed for k")
    return k1neg, k1, k2neg, k2


def weierstrass(CURVE: dict, Fp, Fn: Field, endo: Optional[dict] = None, random_bytes=os.urandom):
    """src/abstract/weierstrass.ts:501-1022 — returns the per-curve Point class."""
    cofactor = CURVE["h"]
    a_is0 = Fp.is0(CURVE["a"])
    b3 = Fp.mul(CURVE["b"], 3)  # weierstrass.ts:612

    def mulA(x):  # weierstrass.ts:613
        return Fp.ZERO if a_is0 else Fp.mul(CURVE["a"], x)

    def acoord(title, n, ban_zero=False):  # weierstrass.ts:640-643
        if not Fp.isValid(n) or (ban_zero and Fp.is0(n)):
            raise ValueError("bad point coordinate " + title)
        return n

    class Point:
        __slots__ = ("X", "Y", "Z")

        def __init__(self, X, Y, Z):  # weierstrass.ts:696-704
            self.X = acoord("x", X)
            self.Y = acoord("y", Y, True)
            self.Z = acoord("z", Z)

        @staticmethod
        def CURVE():
            return CURVE

        @staticmethod
        def fromAffine(p):  # weierstrass.ts:711-718
            if not isinstance(p, dict) or not Fp.isValid(p.get("x")) or not Fp.isValid(p.get("y")):
                raise ValueError("invalid affine point")
            x, y = p["x"], p["y"]
            if Fp.is0(x) and Fp.is0(y):
                return Point.ZERO
            return Point(x, y, Fp.ONE)

        def equals(self, other):  # weierstrass.ts:775-782
            if not isinstance(other, Point):
                raise TypeError("Weierstrass Point expected")
            U1 = Fp.eql(Fp.mul(self.X, other.Z), Fp.mul(other.X, self.Z))
            U2 = Fp.eql(Fp.mul(self.Y, other.Z), Fp.mul(other.Y, self.Z))
            return U1 and U2

        def negate(self):  # weierstrass.ts:785-787
            return Point(self.X, Fp.neg(self.Y), self.Z)

        def double(self):  # weierstrass.ts:793-828 (RCB alg. 3)
            X1, Y1, Z1 = self.X, self.Y, self.Z
            t0 = Fp.mul(X1, X1)
            t1 = Fp.mul(Y1, Y1)
            t2 = Fp.mul(Z1, Z1)
            t3 = Fp.mul(X1, Y1)
            t3 = Fp.add(t3, t3)
            Z3 = Fp.mul(X1, Z1)
            Z3 = Fp.add(Z3, Z3)
            X3 = mulA(Z3)
            Y3 = Fp.mul(b3, t2)
            Y3 = Fp.add(X3, Y3)
            X3 = Fp.sub(t1, Y3)
            Y3 = Fp.add(t1, Y3)
            Y3 = Fp.mul(X3, Y3)
            X3 = Fp.mul(t3, X3)
            Z3 = Fp.mul(b3, Z3)
            t2 = mulA(t2)
            t3 = Fp.sub(t0, t2)
            t3 = mulA(t3)
            t3 = Fp.add(t3, Z3)
            Z3 = Fp.add(t0, t0)
            t0 = Fp.add(Z3, t0)
            t0 = Fp.add(t0, t2)
            t0 = Fp.mul(t0, t3)
            Y3 = Fp.add(Y3, t0)
            t2 = Fp.mul(Y1, Z1)
            t2 = Fp.add(t2, t2)
            t0 = Fp.mul(t2, t3)
            X3 = Fp.sub(X3, t0)
            Z3 = Fp.mul(t2, t1)
            Z3 = Fp.add(Z3, Z3)
            Z3 = Fp.add(Z3, Z3)
            return Point(X3, Y3, Z3)

        def add(self, other):  # weierstrass.ts:834-880 (RCB alg. 1)
            if not isinstance(other, Point):
                raise TypeError("Weierstrass Point expected")
            X1, Y1, Z1 = self.X, self.Y, self.Z
            X2, Y2, Z2 = other.X, other.Y, other.Z
            t0 = Fp.mul(X1, X2)
            t1 = Fp.mul(Y1, Y2)
            t2 = Fp.mul(Z1, Z2)
            t3 = Fp.add(X1, Y1)
            t4 = Fp.add(X2, Y2)
            t3 = Fp.mul(t3, t4)
            t4 = Fp.add(t0, t1)
            t3 = Fp.sub(t3, t4)
            t4 = Fp.add(X1, Z1)
            t5 = Fp.add(X2, Z2)
            t4 = Fp.mul(t4, t5)
            t5 = Fp.add(t0, t2)
            t4 = Fp.sub(t4, t5)
            t5 = Fp.add(Y1, Z1)
            X3 = Fp.add(Y2, Z2)
            t5 = Fp.mul(t5, X3)
            X3 = Fp.add(t1, t2)
            t5 = Fp.sub(t5, X3)
            Z3 = mulA(t4)
            X3 = Fp.mul(b3, t2)
            Z3 = Fp.add(X3, Z3)
            X3 = Fp.sub(t1, Z3)
            Z3 = Fp.add(t1, Z3)
            Y3 = Fp.mul(X3, Z3)
            t1 = Fp.add(t0, t0)
            t1 = Fp.add(t1, t0)
            t2 = mulA(t2)
            t4 = Fp.mul(b3, t4)
            t1 = Fp.add(t1, t2)
            t2 = Fp.sub(t0, t2)
            t2 = mulA(t2)
            t4 = Fp.add(t4, t2)
            t0 = Fp.mul(t1, t4)
            Y3 = Fp.add(Y3, t0)
            t0 = Fp.mul(t5, t4)
            X3 = Fp.mul(t3, X3)
            X3 = Fp.sub(X3, t0)
            t0 = Fp.mul(t3, t1)
            Z3 = Fp.mul(t5, Z3)
            Z3 = Fp.add(Z3, t0)
            return Point(X3, Y3, Z3)

        def subtract(self, other):
            return self.add(other.negate())

        def is0(self):  # weierstrass.ts:889-891
            return self.equals(Point.ZERO)

        def multiply(self, scalar):  # weierstrass.ts:900-907
            if not Fn.isValidNot0(scalar):
                raise ValueError("invalid scalar: out of range")
            p, f = wnaf.mulSecret(self, scalar, cofactor, normalize)
            return normalize([p, f])[0]

        def multiplyUnsafe(self, sc):  # weierstrass.ts:915-928
            if not Fn.isValid(sc):
                raise ValueError("invalid scalar: out of range")
            if sc == 0 or self.is0():
                return Point.ZERO
            if sc == 1:
                return self
            if wnaf.hasWindowSize(self):
                return wnaf.mulUnsafe(self, sc, normalize)
            points: list = []
            scalars: list = []
            push_wnaf_pair(points, scalars, self, sc)
            return mulAddUnsafe(Point, points, scalars)

        def mulAddUnsafe(self, a, other, b):  # weierstrass.ts:937-944
            points: list = []
            scalars: list = []
            push_wnaf_pair(points, scalars, self, a)
            push_wnaf_pair(points, scalars, other, b)
            return mulAddUnsafe(Point, points, scalars)

        def toAffine(self, invertedZ=None):  # weierstrass.ts:951-969
            iz = invertedZ
            X, Y, Z = self.X, self.Y, self.Z
            if Fp.eql(Z, Fp.ONE):
                return {"x": X, "y": Y}
            is0 = self.is0()
            if iz is None:
                iz = Fp.ONE if is0 else Fp.inv(Z)
            x = Fp.mul(X, iz)
            y = Fp.mul(Y, iz)
            zz = Fp.mul(Z, iz)
            if is0:
                return {"x": Fp.ZERO, "y": Fp.ZERO}
            if not Fp.eql(zz, Fp.ONE):
                raise ValueError("invZ was invalid")
            return {"x": x, "y": y}

        def precompute(self, windowSize=8, isLazy=True):  # weierstrass.ts:759-763
            wnaf.setWindowSize(self, windowSize)
            if not isLazy:
                self.multiply(3)
            return self

        def __repr__(self):
            return "<Point %s>" % (self.toAffine(),)

    def push_wnaf_pair(points, scalars, p, k):  # weierstrass.ts:660-671
        if not Fn.isValid(k):
            raise ValueError("invalid scalar: out of range")
        if endo:
            k1neg, k1, k2neg, k2 = split_endo_scalar(k, endo["basises"], Fn.ORDER)
            psi = Point(Fp.mul(p.X, endo["beta"]), p.Y, p.Z)
            points.append(p.negate() if k1neg else p)
            points.append(psi.negate() if k2neg else psi)
            scalars.append(k1)
            scalars.append(k2)
        else:
            points.append(p)
            scalars.append(k)

    def normalize(points):  # weierstrass.ts:1012-1014
        return normalizeZ(Point, points)

    Point.Fp = Fp
    Point.Fn = Fn
    Point.BASE = Point(CURVE["Gx"], CURVE["Gy"], Fp.ONE)
    Point.ZERO = Point(Fp.ZERO, Fp.ONE, Fp.ZERO)  # weierstrass.ts:687
    Point.cofactor = cofactor
    Point.endo = endo
    wnaf = ScalarMultiplier(Point, random_bytes)
    Point.wnaf = wnaf
    return Point


# --------------------------------------------------------------------------------------
# edwards.ts — twisted Edwards, extended coordinates
# --------------------------------------------------------------------------------------


def edwards(CURVE: dict, Fp: Field, Fn: Field, random_bytes=os.urandom):
    """src/abstract/edwards.ts:297-654 — returns the per-curve Point class."""
    cofactor = CURVE["h"]
    a = CURVE["a"]
    d = CURVE["d"]
    if Fp.eql(a, Fp.neg(Fp.ONE)):  # edwards.ts:347-350
        mulA = Fp.neg
    elif Fp.eql(a, Fp.ONE):
        mulA = lambda x: x  # noqa: E731
    else:
        mulA = lambda x: Fp.mul(a, x)  # noqa: E731

    class Point:
        __slots__ = ("X", "Y", "Z", "T")

        def __init__(self, X, Y, Z, T):  # edwards.ts:379-385
            self.X, self.Y, self.Z, self.T = X, Y, Z, T

        @staticmethod
        def CURVE():
            return CURVE

        @staticmethod
        def fromAffine(p):  # edwards.ts:396-402
            x, y = p["x"], p["y"]
            if not (Fp.isValid(x) and Fp.isValid(y)):
                raise ValueError("invalid affine point")
            return Point(x, y, Fp.ONE, Fp.mul(x, y))

        def equals(self, other):  # edwards.ts:482-491
            if not isinstance(other, Point):
                raise TypeError("EdwardsPoint expected")
            X1Z2 = Fp.mul(self.X, other.Z)
            X2Z1 = Fp.mul(other.X, self.Z)
            Y1Z2 = Fp.mul(self.Y, other.Z)
            Y2Z1 = Fp.mul(other.Y, self.Z)
            return Fp.eql(X1Z2, X2Z1) and Fp.eql(Y1Z2, Y2Z1)

        def is0(self):
            return self.equals(Point.ZERO)

        def negate(self):  # edwards.ts:497-500
            return Point(Fp.neg(self.X), self.Y, self.Z, Fp.neg(self.T))

        def double(self):  # edwards.ts:505-521 (dbl-2008-hwcd)
            X1, Y1, Z1 = self.X, self.Y, self.Z
            A = Fp.sqr(X1)
            B = Fp.sqr(Y1)
            C = Fp.mul(Fp.sqr(Z1), 2)
            D = mulA(A)
            x1y1 = Fp.add(X1, Y1)
            E = Fp.sub(Fp.sub(Fp.sqr(x1y1), A), B)
            G = Fp.add(D, B)
            F = Fp.sub(G, C)
            H = Fp.sub(D, B)
            return Point(Fp.mul(E, F), Fp.mul(G, H), Fp.mul(F, G), Fp.mul(E, H))

        def add(self, other):  # edwards.ts:526-545 (add-2008-hwcd)
            if not isinstance(other, Point):
                raise TypeError("EdwardsPoint expected")
            X1, Y1, Z1, T1 = self.X, self.Y, self.Z, self.T
            X2, Y2, Z2, T2 = other.X, other.Y, other.Z, other.T
            A = Fp.mul(X1, X2)
            B = Fp.mul(Y1, Y2)
            C = Fp.mul(Fp.mul(T1, d), T2)
            D = Fp.mul(Z1, Z2)
            E = Fp.sub(Fp.sub(Fp.mul(Fp.add(X1, Y1), Fp.add(X2, Y2)), A), B)
            F = Fp.sub(D, C)
            G = Fp.add(D, C)
            H = Fp.sub(B, mulA(A))
            return Point(Fp.mul(E, F), Fp.mul(G, H), Fp.mul(F, G), Fp.mul(E, H))

        def subtract(self, other):
            return self.add(other.negate())

        def multiply(self, scalar):  # edwards.ts:555-564
            if not Fn.isValidNot0(scalar):
                raise ValueError("invalid scalar: expected 1 <= sc < curve.n")
            p, f = wnaf.mulSecret(self, scalar, cofactor, normalize)
            return normalize([p, f])[0]

        def multiplyUnsafe(self, scalar):  # edwards.ts:571-577
            if not Fn.isValid(scalar):
                raise ValueError("invalid scalar: expected 0 <= sc < curve.n")
            if scalar == 0:
                return Point.ZERO
            if self.is0() or scalar == 1:
                return self
            return wnaf.mulUnsafe(self, scalar, normalize)

        def toAffine(self, invertedZ=None):  # edwards.ts:595-609
            iz = invertedZ
            X, Y, Z = self.X, self.Y, self.Z
            is0 = self.is0()
            if iz is None:
                iz = Fp.create(8) if is0 else Fp.inv(Z)
            x = Fp.mul(X, iz)
            y = Fp.mul(Y, iz)
            zz = Fp.mul(Z, iz)
            if is0:
                return {"x": Fp.ZERO, "y": Fp.ONE}
            if not Fp.eql(zz, Fp.ONE):
                raise ValueError("invZ was invalid")
            return {"x": x, "y": y}

        def clearCofactor(self):  # edwards.ts:611-618
            if cofactor == 1:
                return self
            if cofactor == 2:
                return self.double()
            if cofactor == 4:
                return self.double().double()
            if cofactor == 8:
                return self.double().double().double()
            return self.multiplyUnsafe(cofactor)

        def precompute(self, windowSize=8, isLazy=True):
            wnaf.setWindowSize(self, windowSize)
            if not isLazy:
                self.multiply(3)
            return self

        def toBytes(self) -> bytes:  # edwards.ts:620-628 (RFC 8032 encoding)
            aff = self.toAffine()
            b = bytearray(aff["y"].to_bytes(Fp.BYTES, "little"))
            if aff["x"] & 1:
                b[-1] |= 0x80
            return bytes(b)

        def __repr__(self):
            return "<EdPoint %s>" % (self.toAffine(),)

    def normalize(points):
        return normalizeZ(Point, points)

    Point.Fp = Fp
    Point.Fn = Fn
    Point.BASE = Point(CURVE["Gx"], CURVE["Gy"], Fp.ONE, Fp.mul(CURVE["Gx"], CURVE["Gy"]))
    Point.ZERO = Point(Fp.ZERO, Fp.ONE, Fp.ONE, Fp.ZERO)  # edwards.ts:370
    Point.cofactor = cofactor
    wnaf = ScalarMultiplier(Point, random_bytes)
    Point.wnaf = wnaf
    return Point


# --------------------------------------------------------------------------------------
# Curve instantiations (parameter blocks: SURVEY §8 a17)
# --------------------------------------------------------------------------------------

# src/secp256k1.ts:48-64
SECP256K1_CURVE = dict(
    p=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    h=1,
    a=0,
    b=7,
    Gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    Gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
)
SECP256K1_ENDO = dict(
    beta=0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE,
    basises=(
        (0x3086D221A7D46BCDE86C90E49284EB15, -0xE4437ED6010E88286F547FA90ABFE4C3),
        (0x114CA50F7A8E2F3F657C1108D9D44CFD8, 0x3086D221A7D46BCDE86C90E49284EB15),
    ),
)

# src/ed25519.ts:49-63
ED25519_CURVE = dict(
    p=2**255 - 19,
    n=0x1000000000000000000000000000000014DEF9DEA2F79CD65812631A5CF5D3ED,
    h=8,
    a=2**255 - 20,  # Fp.create(-1)
    d=0x52036CEE2B6FFE738CC740797779E89800700A4D4141D8AB75EB4DCA135978A3,
    Gx=0x216936D3CD6E53FEC0A4E231FDD6DC5C692CC7609525A7B2C9562D608F25D51A,
    Gy=0x6666666666666666666666666666666666666666666666666666666666666658,
)

# src/bn254.ts:80-90
BN254_G1_CURVE = dict(
    p=0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    n=0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    h=1,
    a=0,
    b=3,
    Gx=1,
    Gy=2,
)
# src/bn254.ts:103-106, 207-223
BN254_G2_CURVE = dict(
    n=BN254_G1_CURVE["n"],
    h=0x30644E72E131A029B85045B68181585E06CEECDA572A2489345F2299C0F9FA8D,
    a=(0, 0),
    b=(
        19485874751759354771024239261021720505790618469301721065564631296452457478373,
        266929791119991161246907387137283842545076965332900288569378510910307636690,
    ),
    Gx=(
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ),
    Gy=(
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ),
)

# src/bls12-381.ts:134-148
BLS12_381_G1_CURVE = dict(
    p=0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    n=0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    h=0x396C8C005555E1568C00AAAB0000AAAB,
    a=0,
    b=4,
    Gx=0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    Gy=0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
# src/bls12-381.ts:321-345
BLS12_381_G2_CURVE = dict(
    n=BLS12_381_G1_CURVE["n"],
    h=0x5D543A95414E7F1091D50792876A202CD91DE4547085ABAA68A205B2E5A7DDFA628F1CB4D9E82EF21537E293A6691AE1616EC6E786F0C70CF1C38E31C7238E5,
    a=(0, 0),
    b=(4, 4),
    Gx=(
        0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
    ),
    Gy=(
        0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
    ),
)


def _build():
    out = {}
    # secp256k1 (src/secp256k1.ts:96-101)
    Fp = Field(SECP256K1_CURVE["p"])
    Fn = Field(SECP256K1_CURVE["n"])
    out["secp256k1"] = weierstrass(SECP256K1_CURVE, Fp, Fn, endo=SECP256K1_ENDO)
    # ed25519 (src/ed25519.ts:123; isLE per curve.ts:1036)
    Fp = Field(ED25519_CURVE["p"], is_le=True)
    Fn = Field(ED25519_CURVE["n"], is_le=True)
    out["ed25519"] = edwards(ED25519_CURVE, Fp, Fn)
    # bn254 (src/bn254.ts:226-242)
    Fp = Field(BN254_G1_CURVE["p"])
    Fn = Field(BN254_G1_CURVE["n"])
    out["bn254_G1"] = weierstrass(BN254_G1_CURVE, Fp, Fn)
    out["bn254_G2"] = weierstrass(BN254_G2_CURVE, Field2(Fp), Fn)
    # bls12-381 (src/bls12-381.ts:552-619; G1 base point gets W=4: abstract/bls.ts:935)
    Fp = Field(BLS12_381_G1_CURVE["p"])
    Fn = Field(BLS12_381_G1_CURVE["n"])
    out["bls12_381_G1"] = weierstrass(BLS12_381_G1_CURVE, Fp, Fn)
    out["bls12_381_G2"] = weierstrass(BLS12_381_G2_CURVE, Field2(Fp), Fn)
    # Default base-point window sizes: W=6 (weierstrass.ts:1018, edwards.ts:650), W=4 for bls G1.
    for name, P in out.items():
        P.wnaf.setWindowSize(P.BASE, 4 if name == "bls12_381_G1" else 6)
    return out


CURVES = _build()
CURVE_NAMES = ["secp256k1", "ed25519", "bn254_G1", "bn254_G2", "bls12_381_G1", "bls12_381_G2"]


# --------------------------------------------------------------------------------------
# Test helpers restated from the reference's tests
# --------------------------------------------------------------------------------------


class Xorshift64:
    """test/point.test.ts:536-558 `makeRng`."""

    MASK64 = (1 << 64) - 1

    def __init__(self, seed: int):
        self.seed = seed

    def rnd64(self) -> int:
        s = self.seed
        s = (s ^ (s << 13)) & self.MASK64
        s ^= s >> 7
        s = (s ^ (s << 17)) & self.MASK64
        self.seed = s
        return s

    def rndBig(self, bits: int) -> int:
        r = 0
        for _ in range(0, bits, 64):
            r = (r << 64) | self.rnd64()
        return r & ((1 << bits) - 1)

    def rndBelow(self, n: int) -> int:
        bits = n.bit_length()
        while True:
            r = self.rndBig(bits)
            if r < n:
                return r


def naive_mul(Point, p, k: int):
    """test/point.test.ts naiveMul — double-and-add with only add/double."""
    acc = Point.ZERO
    base = p
    while k > 0:
        if k & 1:
            acc = acc.add(base)
        base = base.double()
        k >>= 1
    return acc


def affine_tuple(Point, p) -> Tuple:
    """Canonical comparison form (test/point.test.ts:36-44): affine (x, y)."""
    a = p.toAffine()
    return (a["x"], a["y"])


def ed25519_public_key(sk: bytes) -> bytes:
    """RFC 8032 §5.1.5 as implemented by src/abstract/edwards.ts:861-893 + src/ed25519.ts:125-135."""
    h = hashlib.sha512(sk).digest()
    head = bytearray(h[:32])
    head[0] &= 248
    head[31] &= 127
    head[31] |= 64
    P = CURVES["ed25519"]
    scalar = int.from_bytes(bytes(head), "little") % P.Fn.ORDER
    return P.BASE.multiply(scalar).toBytes()


# --------------------------------------------------------------------------------------
# EdDSA verification for ed25519 (next-row f1: batch verification is checked against this)
# --------------------------------------------------------------------------------------
ED25519_SQRT_M1 = 19681161376707505956807079304988542015446066515923890162744021073123829784752  # src/ed25519.ts:100-102


def ed25519_uv_ratio(u: int, v: int):
    """src/ed25519.ts:104-121 `uvRatio`: sqrt(u/v) -> (isValid, value)."""
    P = ED25519_CURVE["p"]
    v3 = (v * v * v) % P
    v7 = (v3 * v3 * v) % P
    pw = pow((u * v7) % P, (P - 5) // 8, P)
    x = (u * v3 * pw) % P
    vx2 = (v * x * x) % P
    root1 = x
    root2 = (x * ED25519_SQRT_M1) % P
    use1 = vx2 == u
    use2 = vx2 == (-u) % P
    no_root = vx2 == (-u * ED25519_SQRT_M1) % P
    if use1:
        x = root1
    if use2 or no_root:
        x = root2
    if x & 1:  # isNegativeLE
        x = (-x) % P
    return (use1 or use2), x


def ed25519_point_from_bytes(b: bytes, zip215: bool = False):
    """src/abstract/edwards.ts:405-436 `Point.fromBytes` for ed25519."""
    if len(b) != 32:
        raise ValueError("point expected 32 bytes")
    P = ED25519_CURVE["p"]
    d = ED25519_CURVE["d"]
    last = b[31]
    normed = bytearray(b)
    normed[31] = last & 0x7F
    y = int.from_bytes(bytes(normed), "little")
    mx = (1 << 256) if zip215 else P
    if not (0 <= y < mx):
        raise ValueError("point.y out of range")
    y2 = (y * y) % P
    u = (y2 - 1) % P
    v = (d * y2 + 1) % P
    ok, x = ed25519_uv_ratio(u, v)
    if not ok:
        raise ValueError("bad point: invalid y coordinate")
    is_x_odd = (x & 1) == 1
    is_last_odd = (last & 0x80) != 0
    if not zip215 and x == 0 and is_last_odd:
        raise ValueError("bad point: x=0 and x_0=1")
    if is_last_odd != is_x_odd:
        x = (-x) % P
    Pt = CURVES["ed25519"]
    return Pt(x, y % P, 1, (x * y) % P)


def ed25519_verify(sig: bytes, msg: bytes, public_key: bytes, zip215: bool = True) -> bool:
    """src/abstract/edwards.ts:942-989 `verify` (cofactored; ed25519's default is zip215=true, ed25519.ts:168)."""
    Pt = CURVES["ed25519"]
    L = Pt.Fn.ORDER
    if len(sig) != 64 or len(public_key) != 32:
        raise ValueError("bad lengths")
    r = sig[:32]
    s = int.from_bytes(sig[32:], "little")
    try:
        A = ed25519_point_from_bytes(public_key, zip215)
        Rp = ed25519_point_from_bytes(r, zip215)
        SB = Pt.BASE.multiplyUnsafe(s)  # raises when s >= l
    except ValueError:
        return False
    if not zip215 and A.clearCofactor().is0():
        return False
    k = int.from_bytes(hashlib.sha512(r + public_key + msg).digest(), "little") % L
    RkA = Rp.add(A.multiplyUnsafe(k))
    return RkA.subtract(SB).clearCofactor().is0()


# --------------------------------------------------------------------------------------
# Point codecs (next-row f2): the decode step of `fromBytes`
# --------------------------------------------------------------------------------------
def secp256k1_decode_sec1(b: bytes):
    """src/abstract/weierstrass.ts:565-597 `pointFromBytes` for secp256k1 -> (x, y); raises ValueError."""
    P = SECP256K1_CURVE["p"]
    head, tail = b[0], b[1:]
    if len(b) == 33 and head in (2, 3):
        x = int.from_bytes(tail, "big")
        if not (0 <= x < P):
            raise ValueError("bad point: is not on curve, wrong x")
        y2 = (x * x * x + 7) % P
        y = pow(y2, (P + 1) // 4, P)
        if (y * y) % P != y2:
            raise ValueError("bad point: is not on curve, sqrt error")
        if ((head & 1) == 1) != ((y & 1) == 1):
            y = (-y) % P
        return x, y
    if len(b) == 65 and head == 4:
        x, y = int.from_bytes(tail[:32], "big"), int.from_bytes(tail[32:], "big")
        if not (0 <= x < P and 0 <= y < P) or (y * y - x * x * x - 7) % P:
            raise ValueError("bad point: is not on curve")
        return x, y
    raise ValueError("bad point: got length %d" % len(b))


def bls12_381_g2_decode(b: bytes):
    """src/bls12-381.ts:377-468 `coder('G2').decode` (allowUncompressed): x, y in Fp2 as (c0, c1) tuples, wire order
    c1 || c0 (:354-367), sort bit over [y.c1, y.c0] (:347-352,:488-491); ((0,0),(0,0)) for infinity."""
    P = BLS12_381_G1_CURVE["p"]
    F2 = Field2(Field(P))
    mask = b[0] & 0xE0
    compressed, infinity, sort = bool(mask >> 7 & 1), bool(mask >> 6 & 1), bool(mask >> 5 & 1)
    if (not compressed and not infinity and sort) or (not compressed and infinity and sort) or (compressed and infinity and sort):
        raise ValueError("invalid encoding flag")
    v = bytes([b[0] & 0x1F]) + b[1:]
    ln = 96 if compressed else 192
    if len(v) != ln:
        raise ValueError("invalid G2 point: expected %d bytes" % ln)
    if infinity:
        if any(v):
            raise ValueError("invalid G2 point: non-canonical zero")
        return (0, 0), (0, 0)

    def dec(chunk):
        c1, c0 = int.from_bytes(chunk[:48], "big"), int.from_bytes(chunk[48:], "big")
        if not (0 <= c0 < P and 0 <= c1 < P):
            raise ValueError("invalid field element")
        return (c0, c1)

    x = dec(v[:96])
    if compressed:
        rhs = F2.add(F2.mul(F2.sqr(x), x), BLS12_381_G2_CURVE["b"])
        try:
            y = F2.sqrt(rhs)
        except ValueError:
            raise ValueError("invalid G2 point: compressed")
        parts = [y[1], y[0]]
        bit = False
        for part in parts:
            if part != 0:
                bit = bool((part * 2) // P)
                break
        if bit != sort:
            y = F2.neg(y)
    else:
        y = dec(v[96:])
        if x == (0, 0) and y == (0, 0):
            raise ValueError("invalid G2 point: uncompressed")
    return x, y


def bls12_381_g1_decode(b: bytes):
    """src/bls12-381.ts:377-468 `coder('G1').decode` (allowUncompressed) -> (x, y) with (0, 0) for infinity."""
    P = BLS12_381_G1_CURVE["p"]
    mask = b[0] & 0xE0
    compressed, infinity, sort = bool(mask >> 7 & 1), bool(mask >> 6 & 1), bool(mask >> 5 & 1)
    if (not compressed and not infinity and sort) or (not compressed and infinity and sort) or (compressed and infinity and sort):
        raise ValueError("invalid encoding flag")
    v = bytes([b[0] & 0x1F]) + b[1:]
    ln = 48 if compressed else 96
    if len(v) != ln:
        raise ValueError("invalid G1 point: expected %d bytes" % ln)
    if infinity:
        if any(v):
            raise ValueError("invalid G1 point: non-canonical zero")
        return 0, 0
    x = int.from_bytes(v[:48], "big")
    if not (0 <= x < P):
        raise ValueError("invalid field element")
    if compressed:
        y2 = (pow(x, 3, P) + 4) % P
        y = pow(y2, (P + 1) // 4, P)
        if (y * y) % P != y2:
            raise ValueError("invalid G1 point: compressed")
        if bool((y * 2) // P) != sort:
            y = (-y) % P
    else:
        y = int.from_bytes(v[48:], "big")
        if not (0 <= y < P):
            raise ValueError("invalid field element")
        if x == 0 and y == 0:
            raise ValueError("invalid G1 point: uncompressed")
    return x, y
