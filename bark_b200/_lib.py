"""ctypes binding of the C ABI declared in include/bark_b200.h.

There is NO CPU fallback: if libbark_b200.so is missing or a CUDA device is absent the product path raises.
(The library itself loads without a GPU, which is what the `-m "not gpu"` symbol test checks.)"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_PATH

c_void_p, c_int, c_int32, c_int64, c_uint64, c_double, c_size_t = (
    C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_size_t)


class NodesSoA(C.Structure):
    """bark_nodes_soa: eight device pointers."""
    _fields_ = [(n, c_void_p) for n in
                ("is_leaf", "active", "feature", "threshold", "left", "right", "parent", "depth")]


class Params(C.Structure):
    """bark_params (BARKTrainParamsNumba, src/bark/fitting/bark_sampler.py:48-92)."""
    _fields_ = [("alpha", c_double), ("beta", c_double), ("proposal_weights", c_double * 3),
                ("gamma_prior_shape", c_double), ("gamma_prior_rate", c_double),
                ("use_softplus_transform", c_int32), ("sample_scale", c_int32)]


class McmcDims(C.Structure):
    _fields_ = [(n, c_int64) for n in ("chains", "n", "d", "m", "node_limit", "p_cap")]


ST_TREE_OVERFLOW, ST_HYPER_MODE, ST_COL_OVERFLOW, ST_NOT_SPD, ST_TIMEOUT = 1, 2, 4, 8, 16

# name -> (restype, argtypes); mirrors include/bark_b200.h one to one
SIGNATURES = {
    "bark_abi_version": (c_int, []),
    "bark_last_error": (C.c_char_p, []),
    "bark_nodes_unpack": (c_int, [c_void_p, c_int64, NodesSoA, c_void_p]),
    "bark_nodes_pack": (c_int, [NodesSoA, c_int64, c_void_p, c_void_p]),
    "bark_traverse": (c_int, [NodesSoA, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                              c_void_p]),
    "bark_gram_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int32]),
    "bark_gram_umma": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "bark_gram_to_kernel": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_double, c_int,
                                    c_void_p, c_void_p]),
    "bark_mll_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "bark_mll_batched": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "bark_mcmc_workspace_bytes": (c_size_t, [C.POINTER(McmcDims)]),
    "bark_mcmc_max_p_cap": (c_int64, [C.POINTER(McmcDims)]),
    "bark_mcmc_init": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "bark_mcmc_init_ex": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int32, c_void_p]),
    "bark_kinv_scratch_bytes": (c_size_t, [C.POINTER(McmcDims)]),
    "bark_kinv_export": (c_int, [C.POINTER(McmcDims), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bark_mcmc_sweeps": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, C.POINTER(Params), c_int64, c_uint64,
                                 c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "bark_mcmc_sweeps_ex": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, C.POINTER(Params), c_int64, c_uint64,
                                    c_int64, c_int64, c_void_p, c_void_p, c_int32, c_void_p]),
    "bark_mcmc_sweeps_timed3": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, C.POINTER(Params), c_int64, c_uint64,
                                        c_int64, c_int64, C.POINTER(C.c_float), c_void_p]),
    "bark_mcmc_sweeps_timed": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, C.POINTER(Params), c_int64, c_uint64,
                                       c_int64, c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), c_void_p]),
    "bark_mcmc_read": (c_int, [C.POINTER(McmcDims), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "bark_mcmc_export": (c_int, [C.POINTER(McmcDims), c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "bark_prior_sample": (c_int, [NodesSoA, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_double, c_double,
                                  c_uint64, c_void_p, c_void_p]),
    "bark_predict_scratch_bytes": (c_size_t, [C.POINTER(McmcDims), c_int64]),
    "bark_predict": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, c_void_p, c_int64, c_int, c_double, c_double,
                             c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bark_predict_cov_scratch_bytes": (c_size_t, [C.POINTER(McmcDims), c_int64]),
    "bark_predict_cov": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "bark_predict_prep_bytes": (c_size_t, [C.POINTER(McmcDims), c_int32, c_int32]),
    "bark_predict_prepare": (c_int, [C.POINTER(McmcDims), c_void_p, NodesSoA, c_int32, c_int32, c_void_p, c_void_p]),
    "bark_predict_umma": (c_int, [C.POINTER(McmcDims), c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p]),
    "bark_predict_umma_mixture": (c_int, [C.POINTER(McmcDims), c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_double,
                                          c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "bark_predict_mixture": (c_int, [C.POINTER(McmcDims), c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                                     c_int, c_void_p, c_void_p, c_void_p]),
}

INIT_SKIP_NULL = 1

_lib = None


class BarkError(RuntimeError):
    pass


def load():
    """Load libbark_b200.so (raises if it has not been built: run `python -m bark_b200._build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BarkError(
            f"{LIB_PATH} is missing: build it with `python -m bark_b200._build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.bark_abi_version() != 1:
        raise BarkError("libbark_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().bark_last_error()
        raise BarkError(f"libbark_b200 error {rc}: {msg.decode() if msg else ''}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise BarkError("bark_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch
