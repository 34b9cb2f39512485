"""The C-ABI library loads on a CPU-only box, exports every symbol include/nmsm.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls are made here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nmsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nmsm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nmsm import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTS) == syms


def test_point_and_acc_sizes():
    from nmsm import _lib

    lib = _lib.load()
    assert [lib.nmsm_point_bytes(c) for c in range(6)] == [64, 64, 64, 128, 96, 192]
    assert [lib.nmsm_acc_bytes(c) for c in range(6)] == [128, 128, 128, 256, 192, 384]
    assert lib.nmsm_point_bytes(17) < 0


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import nmsm
    from nmsm import _lib

    with pytest.raises(_lib.NmsmError) as e:
        nmsm.init(0)
    assert e.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(e.value)
    lib = _lib.load()
    import ctypes

    out = ctypes.create_string_buffer(96)
    inf = ctypes.c_int(0)
    rc = lib.nmsm_msm(4, None, None, 0, ctypes.cast(out, ctypes.c_void_p), ctypes.byref(inf))
    assert rc == _lib.ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "noble-curves_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "libhostemu", "libref_msm", "import helpers"):
                    assert needle not in txt, (os.path.join(dirpath, f), needle)


def test_host_mirror_constants_match_oracle():
    import nmsm
    from oracle import noble_ref as R

    for name, C in nmsm.CURVES.items():
        P = R.CURVES[name]
        assert C.Fn.ORDER == P.Fn.ORDER
        a = P.BASE.toAffine()
        assert (C.BASE.x, C.BASE.y) == (a["x"], a["y"])
        base_p = P.Fp.ORDER if not hasattr(P.Fp, "Fp") else P.Fp.Fp.ORDER
        got_p = C.Fp.ORDER if not hasattr(C.Fp, "Fp") else C.Fp.Fp.ORDER
        assert got_p == base_p
        assert C.cofactor == P.cofactor
