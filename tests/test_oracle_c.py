"""Pin the C restatement (oracle/ref_msm.c) to the Python oracle: bit-identical affine MSM results."""
import helpers as H
from oracle import cpu_baseline as CB
from oracle import noble_ref as R


def test_c_port_matches_python_oracle():
    name = "bls12_381_G1"
    P, pts, scalars, _ = H.soak_inputs(name, 513)
    pb, sb = H.pack_points(name, pts), H.pack_scalars(scalars)
    for size in (1, 2, 31, 33, 129, 513):
        exp = H.expected_tuple(name, R.pippenger(P, pts[:size], scalars[:size]))
        for threads in (1, 4):
            assert CB.c_pippenger_bls_g1(pb[: size * 96], sb[: size * 32], size, threads) == exp
    # identity handling: [O]*[123], [G]*[0], P + (-P), 64 x G with equal scalars
    G = P.BASE
    for pts2, sc2 in (([P.ZERO], [123]), ([G], [0]), ([G, G.negate()], [5, 5]), ([G] * 64, [1023] * 64)):
        pts2 = R.normalizeZ(P, pts2)
        exp = H.expected_tuple(name, R.pippenger(P, pts2, sc2))
        assert CB.c_pippenger_bls_g1(H.pack_points(name, pts2), H.pack_scalars(sc2), len(pts2), 2) == exp
    assert CB.c_pippenger_bls_g1(b"", b"", 0, 1) == (0, 0, 1)


def test_c_point_generator_matches_oracle():
    pts, k0, ks = CB.make_points_bls_g1(40, 5)
    P = R.CURVES["bls12_381_G1"]
    for i in (0, 1, 17, 39):
        a = P.BASE.multiplyUnsafe((k0 + i * ks) % P.Fn.ORDER).toAffine()
        assert pts[i * 96:(i + 1) * 96] == a["x"].to_bytes(48, "little") + a["y"].to_bytes(48, "little")


def test_c_port_timing_harness_small():
    r = CB.time_bls_g1_msm(1 << 10, seed=2)
    assert r["kind"] == "port" and r["points_per_s_at_full_size"] > 0
    assert CB.load().ref_pippenger_add_count(1 << 20, 255) == 23592945
    assert CB.load().ref_pippenger_add_count(1 << 16, 255) == 1867757
