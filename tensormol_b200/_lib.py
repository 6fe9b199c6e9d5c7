"""ctypes binding of libtmolb200.so (include/tmolb200.h).

There is no CPU fallback: importing this module without the built library raises, and every
compute entry point needs a CUDA device (the library returns TM_ECUDA otherwise).
"""
from __future__ import annotations

import ctypes as C
import os

TM_MAX_ELE = 8
TM_MAX_HIDDEN = 4

TM_NET_CHARGE, TM_NET_ENERGY = 0, 1
TM_ACT = {"sigmoid_with_param": 0, "relu": 1, "softplus": 2, "tanh": 3, "sigmoid": 4, "elu": 5, "selu": 6}
TM_GEMM_FP32, TM_GEMM_TC_SPLIT, TM_GEMM_TC_SPLIT_PAIR, TM_GEMM_TC_SPLIT_N64, TM_GEMM_TC_SPLIT_N128 = 0, 1, 2, 3, 4
TM_GEMM_TC_3XTF32 = TM_GEMM_TC_SPLIT   # name of the same mode before the fp16 split replaced the tf32 split
TM_F_FORCE, TM_F_VDW, TM_F_DESCRIPTORS, TM_F_FOLD_IMAGES, TM_F_REUSE_NLIST = 1, 2, 4, 8, 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmolb200.so")


class tm_model_desc(C.Structure):
    _fields_ = [("n_ele", C.c_int32), ("eles", C.c_int32 * TM_MAX_ELE), ("n_hidden", C.c_int32), ("hidden", C.c_int32 * TM_MAX_HIDDEN)]


class tm_params(C.Structure):
    _fields_ = [("r_Rc", C.c_double), ("a_Rc", C.c_double), ("eta", C.c_double), ("zeta", C.c_double),
                ("num_r_Rs", C.c_int32), ("num_a_Rs", C.c_int32), ("num_a_As", C.c_int32),
                ("ee_cutoff_on", C.c_double), ("ee_cutoff_off", C.c_double), ("elu_width", C.c_double), ("poly_width", C.c_double),
                ("dsf_alpha", C.c_double), ("elu_shift", C.c_double), ("elu_alpha", C.c_double),
                ("add_ecc", C.c_int32), ("activation", C.c_int32), ("sigmoid_alpha", C.c_double),
                ("C6", C.c_double * TM_MAX_ELE), ("Rvdw", C.c_double * TM_MAX_ELE)]


class tm_outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("Etotal", "Ebp", "Ebp_atom", "Ecc", "Evdw", "dipole", "charge", "gradient", "descriptors")]


class tm_timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("total", "h2d", "nlist", "desc", "mlp_fwd", "pair", "mlp_bwd", "force", "d2h")] + \
               [(n, C.c_int64) for n in ("n_centres", "n_slots", "n_rad_pairs", "n_ang_neigh", "n_triples")] + [("launches", C.c_int32)]


# every symbol declared in include/tmolb200.h (tests/test_abi.py checks the library exports them all)
SYMBOLS = ["tm_version", "tm_last_error", "tm_device_count", "tm_create", "tm_destroy", "tm_set_params", "tm_set_weights",
           "tm_set_gemm_mode", "tm_get_gemm_mode", "tm_set_skin", "tm_set_stream", "tm_descriptor_width", "tm_nlist", "tm_pairs_triples_ele",
           "tm_eval", "tm_eval_dev", "tm_eval_images", "tm_eval_lattice", "tm_eval_lattice_dev", "tm_slab_phase_a", "tm_slab_phase_b",
           "tm_slab_phase_c", "tm_slab_p2p_bytes", "tm_slab_p2p_setup", "tm_get_timings", "tm_sync"]

_lib = None


class TMolB200Error(RuntimeError):
    pass


def load():
    """Load the shared library; raise loudly if it has not been built (python -m tensormol_b200.csrc.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TMolB200Error(f"{LIB_PATH} is missing: build it with `python -m tensormol_b200.csrc.build` "
                            "(tensormol_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    P = C.POINTER
    lib.tm_version.restype = i32
    lib.tm_last_error.restype = C.c_char_p
    lib.tm_device_count.restype = i32
    lib.tm_create.restype = vp
    lib.tm_create.argtypes = [i32, P(tm_model_desc), P(tm_params)]
    lib.tm_destroy.argtypes = [vp]
    lib.tm_destroy.restype = None
    lib.tm_set_params.argtypes = [vp, P(tm_params)]
    lib.tm_set_weights.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32]
    lib.tm_set_gemm_mode.argtypes = [vp, i32]
    lib.tm_set_skin.argtypes = [vp, dbl]
    lib.tm_get_gemm_mode.argtypes = [vp]
    lib.tm_set_stream.argtypes = [vp, vp]
    lib.tm_descriptor_width.argtypes = [vp]
    lib.tm_nlist.argtypes = [vp, vp, i64, i64, dbl, i32, P(vp), P(vp)]
    lib.tm_pairs_triples_ele.argtypes = [vp, vp, vp, i64, i64, vp, vp, dbl, dbl, P(i64), P(i64), P(vp), P(vp), P(vp), P(vp)]
    lib.tm_eval.argtypes = [vp, vp, vp, i64, i64, vp, i32, P(tm_outputs)]
    lib.tm_eval_images.argtypes = [vp, vp, vp, i64, i64, i32, P(tm_outputs)]
    lib.tm_eval_lattice.argtypes = [vp, vp, vp, i64, vp, i32, i32, P(tm_outputs)]
    lib.tm_eval_lattice_dev.argtypes = [vp, vp, vp, i64, vp, i32, i32, vp, vp, vp]
    lib.tm_eval_dev.argtypes = [vp, vp, vp, i64, i64, i32, vp, vp, vp]
    lib.tm_slab_phase_a.argtypes = [vp, vp, vp, i64, vp, i32, i32, i32, vp]
    lib.tm_slab_phase_b.argtypes = [vp, vp, vp]
    lib.tm_slab_phase_c.argtypes = [vp, vp, i32, vp]
    lib.tm_slab_p2p_bytes.argtypes = [i32, i64]
    lib.tm_slab_p2p_bytes.restype = i64
    lib.tm_slab_p2p_setup.argtypes = [vp, i32, i32, i64, vp]
    lib.tm_get_timings.argtypes = [vp, P(tm_timings)]
    lib.tm_sync.argtypes = [vp]
    for name in SYMBOLS:
        if name not in ("tm_last_error", "tm_create", "tm_destroy"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().tm_last_error().decode("utf-8", "replace")
        raise TMolB200Error(f"{what} failed (code {rc}): {msg}")
