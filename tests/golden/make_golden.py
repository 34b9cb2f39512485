#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the scalar-mult / MSM path into small fixtures.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the JSON files written here.

Sources (all under /root/reference/test/vectors, used by the reference tests cited):
  secp256k1/privates-2.txt      test/secp256k1.test.ts:59-76     k:x:y = k*G
  secp256k1/points.json         test/secp256k1.test.ts:96-127    pointMultiply / pointFromScalar / pointAdd
  secp256k1/endomorphism.json   test/nist.test.ts:551            GLV sign cases
  bls12-381/zkcrypto/converted.json  test/bls12-381.test.ts:1463-1533   i*G, i<1000, G1+G2 uncompressed
  bn254/eth-dump.js             test/bn254.test.ts:750-778       EIP-196 ECADD/ECMUL dumps
  bn254/seda.js                 test/bn254.test.ts:859-887       add / mul
  ed25519/vectors.txt           test/ed25519.test.ts:50-78       sk:pk:msg:sig (RFC 8032 / cr.yp.to)
  (test source) test/fft.test.ts:155-215  fixed rootsOfUnity tables (roots(3), brp(3)) of bls12_381.fields.Fr and
                                           bn254.fields.Fr with generator 7 -> fft.json
"""
import json
import os
import re

REF = "/root/reference/test/vectors"
OUT = os.path.dirname(os.path.abspath(__file__))


def dump(name, obj):
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print(name, os.path.getsize(path), "bytes")


def secp256k1():
    rows = []
    for line in open(f"{REF}/secp256k1/privates-2.txt"):
        line = line.strip()
        if not line:
            continue
        k, x, y = line.split(":")
        rows.append([k, x, y])
    pts = json.load(open(f"{REF}/secp256k1/points.json"))["valid"]
    out = {
        "privates2": rows,  # decimal k, hex x, hex y
        "pointMultiply": [[v["P"], v["d"], v["expected"]] for v in pts["pointMultiply"]],
        "pointFromScalar": [[v["d"], v["expected"]] for v in pts["pointFromScalar"]],
        "pointAdd": [[v["P"], v["Q"], v["expected"]] for v in pts["pointAdd"]],
        "endomorphism": json.load(open(f"{REF}/secp256k1/endomorphism.json")),
        # test/secp256k1.test.ts:96-104 isPoint: 33-byte SEC1 encodings only (the decoder under test)
        "isPoint33": [[v["P"], v["expected"]] for v in pts["isPoint"] if len(v["P"]) == 66],
    }
    dump("secp256k1.json", out)


def bls():
    d = json.load(open(f"{REF}/bls12-381/zkcrypto/converted.json"))
    out = {
        # index i holds i*G (i = 0 is the point at infinity, Zcash flag encoding)
        "G1_Uncompressed": d["G1_Uncompressed"][:1000],
        "G2_Uncompressed": d["G2_Uncompressed"][:256],
        "G1_Compressed": d["G1_Compressed"][:1000],  # test/bls12-381.test.ts:1463-1500 (Zcash-flag codec)
        "G2_Compressed": d["G2_Compressed"][:256],   # test/bls12-381.test.ts:1500-1533
    }
    dump("bls12_381.json", out)


def bn254():
    src = open(f"{REF}/bn254/eth-dump.js").read()
    adds, muls = [], []
    for m in re.finditer(r"^NOBLE_DUMP_EC_(ADD|MUL) (\S*) (\S+)$", src, re.M):
        kind, inp, outp = m.group(1), m.group(2), m.group(3)
        (adds if kind == "ADD" else muls).append([inp, outp])
    seda_src = open(f"{REF}/bn254/seda.js").read()
    # seda.js is `const vectors = {...}; export default vectors` with JS object-literal syntax
    body = seda_src[seda_src.index("{"): seda_src.rindex("}") + 1]
    body = re.sub(r"(\w+):", r'"\1":', body)
    body = body.replace("'", '"')
    body = re.sub(r",\s*([}\]])", r"\1", body)
    seda = json.loads(body)
    dump("bn254.json", {"eth_add": adds, "eth_mul": muls, "seda_add": seda["add"], "seda_mul": seda["mul"]})


def ed25519():
    rows = []
    for i, line in enumerate(open(f"{REF}/ed25519/vectors.txt")):
        if i >= 128:
            break
        parts = line.strip().split(":")
        sk_pk, pk, msg, sig_msg = parts[0], parts[1], parts[2], parts[3]
        rows.append({"sk": sk_pk[:64], "pk": pk, "msg": msg, "sig": sig_msg[:128]})
    zip215 = json.load(open(f"{REF}/ed25519/zip215.json"))  # test/ed25519.test.ts:392-405, message = "Zcash"
    edge = json.load(open(f"{REF}/ed25519/edge-cases.json"))  # test/ed25519.test.ts:189-196
    dump("ed25519.json", {"vectors": rows, "zip215": zip215, "edge_cases": edge})


def fft():
    """The reference keeps its NTT known answers inline in test/fft.test.ts ('cache and fixed vectors'): the
    8-entry root tables for generator 7.  Extracted by locating the eql(roots.roots(3) / roots.brp(3), [...]) blocks."""
    src = open("/root/reference/test/fft.test.ts").read()
    out = {}
    for field in ("bls12_381", "bn254"):
        a = src.index("roots = fft.rootsOfUnity(%s.fields.Fr, 7n);" % field)
        seg = src[a:a + 6000]
        for key in ("roots", "brp"):
            m = re.search(r"roots\.%s\(3\),\s*\[(.*?)\]" % key, seg, re.S)
            vals = [int(x) for x in re.findall(r"(\d+)n", m.group(1))]
            assert len(vals) == 8, (field, key, len(vals))
            out["%s_%s3" % (field, key)] = [str(v) for v in vals]
    dump("fft.json", out)


if __name__ == "__main__":
    fft()
    secp256k1()
    bls()
    bn254()
    ed25519()
