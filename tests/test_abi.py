"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/tmolb200.h
declares, and refuses to compute without a CUDA device (there is no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tmolb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tensormol_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tmolb200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.tm_version() >= 100


def test_struct_layouts_match_header():
    from tensormol_b200 import _lib
    # sizes computed from the header's field lists (x86-64 SysV alignment)
    assert C.sizeof(_lib.tm_model_desc) == 4 + 4 * 8 + 4 + 4 * 4
    assert C.sizeof(_lib.tm_params) == 4 * 8 + 3 * 4 + 4 + 7 * 8 + 2 * 4 + 8 + 2 * 8 * 8
    assert C.sizeof(_lib.tm_outputs) == 9 * 8
    assert C.sizeof(_lib.tm_timings) == 9 * 4 + 4 + 5 * 8 + 8


def test_no_cpu_fallback_without_device():
    import torch
    from tensormol_b200 import _lib
    from tensormol_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    params = dict(AN1_r_Rc=4.6, AN1_a_Rc=3.1, AN1_eta=4.0, AN1_zeta=8.0, AN1_num_r_Rs=32, AN1_num_a_Rs=8, AN1_num_a_As=8, EECutoffOn=0.0,
                  EECutoffOff=15.0, Elu_Width=4.6, Poly_Width=4.6, DSFAlpha=0.18, AddEcc=True, sigmoid_alpha=100.0, NeuronType="sigmoid_with_param")
    with pytest.raises(_lib.TMolB200Error, match="no CUDA device"):
        Engine([1, 8], [8], params)


def test_oracle_is_not_imported_by_the_product():
    """The product package must never route through oracle/ (checked statically)."""
    pkg = os.path.join(ROOT, "tensormol_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in txt or f in ("build.py",), f"{f} references oracle/"


def test_c_oracle_matches_numpy_oracle():
    so = os.path.join(ROOT, "oracle", "liboracle_c.so")
    if not os.path.exists(so):
        pytest.skip("oracle/liboracle_c.so not built (make -C oracle)")
    from oracle import oracle_np as onp
    lib = C.CDLL(so)
    lib.tm_oracle_nlist_naive.restype = C.c_int64
    lib.tm_oracle_nlist_naive.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    rng = np.random.default_rng(3)
    for n, nreal, rc, perms in [(400, 400, 4.6, 1), (400, 150, 3.1, 0), (50, 50, 15.0, 1)]:
        x = np.ascontiguousarray(rng.uniform(0, (n / 0.1) ** (1 / 3), (n, 3)))
        off = np.zeros(nreal + 1, np.int64)
        p = C.c_void_p()
        total = lib.tm_oracle_nlist_naive(x.ctypes.data, n, nreal, rc, perms, off.ctypes.data, C.byref(p))
        idx = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(total, 1),))[:total].astype(np.int64)
        for i in range(nreal):
            idx[off[i]:off[i + 1]].sort()
        o_off, o_idx = onp.nlist_csr(x, rc, nreal, perms)
        assert np.array_equal(off, o_off) and np.array_equal(idx, o_idx)
        lib.tm_oracle_free(p)


def test_neuron_type_ids_match_header_enum():
    """The Python mirror's NeuronType -> id table equals the TM_ACT_* enum of include/tmolb200.h."""
    from tensormol_b200 import _lib
    src = open(os.path.join(ROOT, "include", "tmolb200.h")).read()
    ids = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"TM_ACT_([A-Z_]+)\s*=\s*(\d+)", src)}
    assert ids == _lib.TM_ACT and len(ids) == 7


def test_unknown_neuron_type_is_refused():
    from oracle import oracle_graph as og
    from tensormol_b200.engine import Engine
    P = og.default_params()
    P["NeuronType"] = "gaussian"
    with pytest.raises(ValueError, match="not supported"):
        Engine([1, 8], [8], P)
