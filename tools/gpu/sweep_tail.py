#!/usr/bin/env python3
"""Exploration of the bucket-reduction tail: first level in the lane-parallel (quad) or the one-thread-per-chunk form
(NMSM_QUAD_REDUCE1=2 forces the former) and the chunk length K, per curve / size — each setting in a fresh process
(the library reads the variables once)."""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("bls12_381_G2", 18), ("bls12_381_G2", 14), ("bls12_381_G1", 16), ("bls12_381_G1", 20), ("bn254_G1", 20), ("secp256k1", 16)]
ENVS = [{}, {"NMSM_QUAD_REDUCE1": "2"}, {"NMSM_K": "4"}, {"NMSM_K": "4", "NMSM_QUAD_REDUCE1": "2"}, {"NMSM_K": "16", "NMSM_QUAD_REDUCE1": "2"}]
CHILD = "import sys; sys.path.insert(0, %r); import sweep_c, nmsm; nmsm.init(0); sweep_c.sweep_msm(sys.argv[1], int(sys.argv[2]), [0])" % HERE
for name, logn in CASES:
    for env in ENVS:
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", CHILD, name, str(logn)], env=e, capture_output=True, text=True, timeout=400)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"error": r.stderr[-300:]})
        try:
            d = json.loads(line); d["env"] = env; line = json.dumps(d)
        except Exception:
            pass
        print(line, flush=True)
