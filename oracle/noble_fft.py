"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's radix-2 NTT (SURVEY §8 f4 companion of the MSM).

Follows /root/reference/src/abstract/fft.ts with Python ints (identical semantics to JS BigInt for + - * %):
  reverse_bits / bit_reversal_permutation   fft.ts:93-109,136-173
  RootsOfUnity                               fft.ts:230-312  (findGenerator :175-180)
  fft_core                                   fft.ts:422-480  (DIT / DIF loops, brp boundary)
  FFT.direct / FFT.inverse                   fft.ts:518-575

Parity status: PINNED — tests/test_oracle_fft.py checks this module against the fixed root tables the reference's
own test holds for bls12_381.fields.Fr and bn254.fields.Fr with generator 7 (test/fft.test.ts:155-200, committed as
tests/golden/fft.json) and against the DFT definition direct(a)[k] == a(omega^k) (test/fft.test.ts:620-632).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.
"""

FR = {
    # src/bn254.ts / src/bls12-381.ts scalar-field orders (same numbers as oracle/noble_ref.py CURVES[...].Fn.ORDER)
    "bn254": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    "bls12_381": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
}


def is_power_of_two(x: int) -> bool:
    return x != 0 and (x & (x - 1)) == 0


def reverse_bits(n: int, bits: int) -> int:
    """fft.ts:93-109"""
    r = 0
    for _ in range(bits):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def bit_reversal_permutation(values):
    """fft.ts:136-173 (returns a new list)"""
    n = len(values)
    if not is_power_of_two(n):
        raise ValueError("expected positive power-of-two length, got %d" % n)
    bits = n.bit_length() - 1
    out = list(values)
    for i in range(n):
        j = reverse_bits(i, bits)
        if i < j:
            out[i], out[j] = out[j], out[i]
    return out


def find_generator(p: int) -> int:
    """fft.ts:175-180: the smallest G >= 2 that is not a quadratic residue."""
    g = 2
    while pow(g, p >> 1, p) == 1:
        g += 1
    return g


class RootsOfUnity:
    """fft.ts:230-312"""

    def __init__(self, p: int, generator=None):
        self.p = p
        odd, two = p - 1, 0
        while odd & 1 == 0:
            odd >>= 1
            two += 1
        self.G = generator if generator is not None else find_generator(p)
        self.power_of_two, self.odd_factor = two, odd
        om = [0] * (two + 1)
        om[two] = pow(self.G, odd, p)
        for i in range(two, 0, -1):
            om[i - 1] = om[i] * om[i] % p
        self._omegas = om
        self._cache = {}

    def _check(self, bits):
        if bits > 31 or bits > self.power_of_two or bits < 0:
            raise ValueError("rootsOfUnity: wrong bits %d powerOfTwo=%d" % (bits, self.power_of_two))
        return bits

    def omega(self, bits: int) -> int:
        return self._omegas[self._check(bits)]

    def roots(self, bits: int):
        self._check(bits)
        if bits not in self._cache:
            w, cur, out = self._omegas[bits], 1, []
            for _ in range(1 << bits):
                out.append(cur)
                cur = cur * w % self.p
            self._cache[bits] = out
        return self._cache[bits]

    def brp(self, bits: int):
        return bit_reversal_permutation(self.roots(bits))

    def inverse(self, bits: int):
        r = self.roots(bits)
        return [r[0]] + r[1:][::-1]


def fft_core(p: int, values, roots, dit: bool, brp: bool = True):
    """fft.ts:422-480 with invertButterflies = false, skipStages = 0 (the flavours FFT() uses); mutates a copy."""
    n = len(values)
    if not is_power_of_two(n):
        raise ValueError("FFT: Polynomial size should be power of two")
    if len(roots) != n:
        raise ValueError("FFT: wrong roots length: expected %d, got %d" % (n, len(roots)))
    bits = n.bit_length() - 1
    v = list(values)
    if dit and brp:
        v = bit_reversal_permutation(v)
    for i in range(bits):
        s = i + 1 if dit else bits - i
        m = 1 << s
        m2 = m >> 1
        stride = n >> s
        for k in range(0, n, m):
            for j in range(m2):
                omega = roots[j * stride]
                i0, i1 = k + j, k + j + m2
                a, b = v[i0], v[i1]
                if dit:
                    t = b * omega % p
                    v[i0] = (a + t) % p
                    v[i1] = (a - t) % p
                else:
                    v[i0] = (a + b) % p
                    v[i1] = (a - b) * omega % p
    if not dit and brp:
        v = bit_reversal_permutation(v)
    return v


class FFT:
    """fft.ts:518-575 over a prime field (opts = the field itself)."""

    def __init__(self, roots: RootsOfUnity):
        self.roots, self.p = roots, roots.p

    def _loop(self, values, table, brp_input, brp_output):
        if brp_input and brp_output:
            return fft_core(self.p, bit_reversal_permutation(values), table, dit=False, brp=False)
        if brp_input:
            return fft_core(self.p, values, table, dit=True, brp=False)
        if brp_output:
            return fft_core(self.p, values, table, dit=False, brp=False)
        return fft_core(self.p, values, table, dit=True, brp=True)

    def direct(self, values, brp_input=False, brp_output=False):
        n = len(values)
        if not is_power_of_two(n):
            raise ValueError("FFT: Polynomial size should be power of two")
        return self._loop(values, self.roots.roots(n.bit_length() - 1), brp_input, brp_output)

    def inverse(self, values, brp_input=False, brp_output=False):
        n = len(values)
        if not is_power_of_two(n):
            raise ValueError("FFT: Polynomial size should be power of two")
        res = self._loop(values, self.roots.inverse(n.bit_length() - 1), brp_input, brp_output)
        ivm = pow(n, -1, self.p)
        return [x * ivm % self.p for x in res]


def eval_poly(p: int, coeffs, x: int) -> int:
    """Horner evaluation (monomial basis), used to state the DFT definition."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % p
    return acc
