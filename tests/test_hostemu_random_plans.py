"""Randomised differential test of the device MSM bodies (host emulation) against the oracle's pippenger: random sizes,
forced window sizes and segment lengths, scalars with the patterns that stress the signed recoding (all-ones runs,
n - 1, powers of two, zeros) and repeated / negated / identity points.  Deterministic (hypothesis derandomize)."""
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import helpers as H
from oracle import noble_ref as R

_POOL = {}


def pool(name):
    """A small fixed pool of affine points per curve: multiples of G, their negatives, the identity."""
    if name not in _POOL:
        P = R.CURVES[name]
        pts = [P.BASE.multiplyUnsafe(k) for k in (1, 2, 3, 7, 1 << 70, P.Fn.ORDER - 1, 0xDEADBEEF, 12345678901234567890)]
        pts += [p.negate() for p in pts[:3]] + [P.ZERO]
        _POOL[name] = R.normalizeZ(P, pts)
    return _POOL[name]


Z = 0xD201000000010000  # |x| of BLS12-381: the base of the four-way psi split on G2 (msm_body.cuh gls_split)


def scalar_strategy(order):
    special = [0, 1, 2, order - 1, order - 2, (1 << 128) - 1, 1 << 128, (1 << 127) + 1, 0x8000_8000_8000_8000, 0xFFFF_0000_FFFF,
               (1 << (order.bit_length() - 1)) - 1, order >> 1, (order >> 1) + 1,
               Z - 1, Z, Z + 1, Z * Z, Z**3, Z**3 - 1, Z // 2, Z // 2 + 1, (Z // 2 + 1) * Z**3, (Z // 2) * (1 + Z + Z * Z + Z**3)]
    return st.one_of(st.sampled_from([s % order for s in special]), st.integers(min_value=0, max_value=order - 1))


@pytest.mark.parametrize("name", ["secp256k1", "ed25519", "bls12_381_G1", "bn254_G1", "bls12_381_G2"])
def test_random_plans_match_pippenger(name):
    P = R.CURVES[name]
    pts_pool = pool(name)

    @settings(max_examples=60 if "G2" not in name else 24, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(data=st.data())
    def run(data):
        n = data.draw(st.integers(min_value=1, max_value=24 if "G2" not in name else 9))
        idx = data.draw(st.lists(st.integers(0, len(pts_pool) - 1), min_size=n, max_size=n))
        scalars = data.draw(st.lists(scalar_strategy(P.Fn.ORDER), min_size=n, max_size=n))
        c = data.draw(st.sampled_from([0, 2, 3, 5, 8, 11, 16]))
        L = data.draw(st.sampled_from([0, 1, 2, 7, 32]))
        table_c = data.draw(st.sampled_from([0, 0, 6, 9]))
        pts = [pts_pool[i] for i in idx]
        exp = H.expected_tuple(name, R.pippenger(P, pts, scalars))
        got, err, _ = H.emu_msm(name, H.pack_points(name, pts), H.pack_scalars(scalars), n, c if not table_c else 0, L,
                                table_c=table_c)
        assert err == (0xFFFFFFFF, 0xFFFFFFFF)
        assert got == exp, (name, n, idx, scalars, c, L, table_c)

    run()
