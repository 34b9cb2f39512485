"""Pin the NTT oracle (oracle/noble_fft.py) against the reference's fixed root tables and the DFT definition.

Mirrors test/fft.test.ts: 'table structure' (:120-148), 'cache and fixed vectors' (:149-215), 'random and algebra
properties' (:545-617), 'DFT semantics' (:620-632).
"""
import random

import pytest

from conftest import load_golden
from oracle import noble_fft as F

FIELDS = ["bls12_381", "bn254"]


@pytest.mark.parametrize("name", FIELDS)
def test_fixed_root_tables(name):
    g = load_golden("fft.json")
    r = F.RootsOfUnity(F.FR[name], 7)
    assert [str(x) for x in r.roots(3)] == g["%s_roots3" % name]
    assert [str(x) for x in r.brp(3)] == g["%s_brp3" % name]


@pytest.mark.parametrize("name", FIELDS)
def test_table_structure(name):
    p = F.FR[name]
    for gen in (7, None):
        r = F.RootsOfUnity(p, gen)
        for bits in (1, 5, 9):
            om, n = r.roots(bits), 1 << bits
            assert len(om) == n and om[0] == 1
            assert r.inverse(bits) == [pow(x, -1, p) for x in om]
            assert r.roots(bits - 1) == om[::2]
            w = r.omega(bits)
            assert w == (om[1] if n > 1 else 1) and pow(w, n, p) == 1 and pow(w, n // 2, p) == p - 1
    small = F.RootsOfUnity(17)
    assert small.G == 3 and small.power_of_two == 4
    with pytest.raises(ValueError):
        F.RootsOfUnity(p, 7).roots(40)


@pytest.mark.parametrize("name", FIELDS)
def test_dft_semantics_and_roundtrips(name):
    p = F.FR[name]
    roots = F.RootsOfUnity(p, 7)
    fft = F.FFT(roots)
    rnd = random.Random(1)
    for bits in (0, 1, 3, 6):
        n = 1 << bits
        a = [rnd.randrange(1, p) for _ in range(n)]
        d = fft.direct(a)
        assert d == [F.eval_poly(p, a, w) for w in roots.roots(bits)]  # direct(a)[k] == a(omega^k)
        assert fft.inverse(d) == a and fft.direct(fft.inverse(a)) == a
        # the four boundary layouts agree up to the bit-reversal permutation (fft.ts:538-550)
        brp = F.bit_reversal_permutation
        assert fft.direct(a, False, True) == brp(d)
        assert fft.direct(brp(a), True, False) == d
        assert fft.direct(brp(a), True, True) == brp(d)
        assert fft.inverse(brp(d), True, False) == a
        assert fft.inverse(d, False, True) == brp(a)
        b = [rnd.randrange(p) for _ in range(n)]
        c = rnd.randrange(1, p)
        assert fft.direct([(x + y) % p for x, y in zip(a, b)]) == [(x + y) % p for x, y in zip(d, fft.direct(b))]
        assert fft.direct([x * c % p for x in a]) == [x * c % p for x in d]
    out = fft.direct([5] * 256)
    assert out[0] == 5 * 256 % p and not any(out[1:])
    with pytest.raises(ValueError):
        fft.direct([1, 2, 3])
