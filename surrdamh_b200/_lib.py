"""ctypes binding of the C ABI in include/surrdamh_b200.h (libsurrdamh_b200.so, sm_100a).

This is the only place the Python host touches native code.  Device memory is held in
torch CUDA tensors (plumbing); the library sees plain device pointers and a stream handle.
There is NO CPU fallback: `lib()` raises if the shared library is missing, and every compute
wrapper raises if its tensors are not CUDA tensors.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libsurrdamh_b200.so")

# ---- constants of the header ----------------------------------------------------------------
ABI_VERSION = 2
OK, ERR_ARG, ERR_CUDA, ERR_LIMIT, ERR_SINGULAR = 0, 1, 2, 3, 4
MAX_D, MAX_M, MAX_DEGREE = 32, 64, 10
FLAG_PREREJECTED, FLAG_ACCEPTED, FLAG_REJECTED, FLAG_PENDING = 0, 1, 2, 3
MODEL_NONE, MODEL_POLY, MODEL_RBF, MODEL_TOY, MODEL_ZERO = 0, 1, 2, 3, 4
ALG_MH, ALG_DAMH = 0, 1
RNG_PHILOX, RNG_STREAMS = 0, 1
PROP_SHARED, PROP_PER_CHAIN, PROP_FACTOR = 0, 1, 2
PHASE_FUSED, PHASE_PROPOSE, PHASE_DECIDE, PHASE_PREPARE = 0, 1, 2, 3
RBF_SOLVER_DIRECT, RBF_SOLVER_MINRES = 0, 1

_dp = C.c_void_p  # every device pointer travels as void*


class Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("kernel_type", C.c_int32), ("degree", C.c_int32), ("n_terms", C.c_int32),
                ("n_centers", C.c_int64), ("coefs", _dp), ("centers", _dp), ("alpha", _dp), ("herm", _dp),
                ("toy_f", C.c_double), ("toy_L", C.c_double), ("toy_M", C.c_double)]


class Problem(C.Structure):
    _fields_ = [("d", C.c_int32), ("m", C.c_int32), ("prior_mean", _dp), ("prior_full", C.c_int32),
                ("prior_scale", _dp), ("prior_piv", _dp), ("obs", _dp), ("noise_full", C.c_int32),
                ("noise_scale", _dp), ("noise_piv", _dp)]


class ChainState(C.Structure):
    _fields_ = [("n_chains", C.c_int64), ("u", _dp), ("G_u", _dp), ("post_u", _dp), ("pre_u", _dp),
                ("counters", _dp), ("moments", _dp)]


class ChainArgs(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("phase", C.c_int32), ("n_steps", C.c_int32),
                ("lanes_per_chain", C.c_int32), ("refresh_pre", C.c_int32), ("have_G", C.c_int32),
                ("problem", Problem), ("surrogate", Model), ("truth", Model), ("state", ChainState),
                ("prop_kind", C.c_int32), ("prop_scale", _dp),
                ("rng_kind", C.c_int32), ("seed", C.c_uint64), ("stream_id", C.c_uint32),
                ("chain0", C.c_int64), ("step0", C.c_int64), ("z", _dp), ("uni", _dp),
                ("flags", _dp), ("prop", _dp), ("post", _dp), ("pre", _dp),
                ("snap_par", _dp), ("snap_obs", _dp), ("G_prop", _dp), ("pre_cur", _dp), ("coresident", C.c_int32), ("surrogate_f32", C.c_int32),
                ("chain_lo", C.c_int64), ("chain_n", C.c_int64)]


class SdbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("surrdamh_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None

_I, _L, _P = C.c_int32, C.c_int64, C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "sdb_abi_version": (C.c_int, []),
    "sdb_last_error": (C.c_char_p, []),
    "sdb_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "sdb_poly_eval": (C.c_int, [_P, _L, _I, _I, C.POINTER(Model), _P, _P]),
    "sdb_rbf_eval": (C.c_int, [_P, _L, _I, _I, C.POINTER(Model), _P, _P]),
    "sdb_rbf_eval_f32": (C.c_int, [_P, _L, _I, _I, C.POINTER(Model), _P, _P]),
    "sdb_chain_run": (C.c_int, [C.POINTER(ChainArgs), _P]),
    "sdb_chain_plan": (C.c_int, [C.POINTER(ChainArgs)] + [C.POINTER(C.c_int)] * 4),
    "sdb_philox_streams": (C.c_int, [C.c_uint64, C.c_uint32, _L, _L, _I, _I, _L, _P, _P, _P]),
    "sdb_philox_raw": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _L, _P, _P]),
    "sdb_snapshot_compact": (C.c_int, [_P, _P, _P, _I, _L, _I, _I, _P, _P, _L, _P, _P, _P]),
    "sdb_poly_basis": (C.c_int, [_P, _L, _I, C.POINTER(Model), _P, _P]),
    "sdb_poly_fit_work": (C.c_int64, [_L, _I, _I]),
    "sdb_poly_fit": (C.c_int, [_P, _P, _L, _I, _I, C.POINTER(Model), _P, _P, _P, _P]),
    "sdb_rbf_fit_work": (C.c_int64, [_L, _I, _I]),
    "sdb_rbf_fit": (C.c_int, [_P, _P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "sdb_rbf_fit_nullspace_work": (C.c_int64, [_L, _I, _I]),
    "sdb_rbf_fit_nullspace": (C.c_int, [_P, _P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "sdb_rbf_system": (C.c_int, [_P, _P, _L, _I, _I, _I, _P, _P, _P]),
    "sdb_lu_solve": (C.c_int, [_P, _L, _P, _I, _P, _P, _P]),
    "sdb_rbf_thin_work": (C.c_int64, [_L]),
    "sdb_rbf_thin": (C.c_int, [_P, _L, _I, _L, _P, _P, _P]),
    "sdb_rbf_thin_masked": (C.c_int, [_P, _L, _I, _L, _P, _I, _P, _P, _P]),
    "sdb_snapshot_select": (C.c_int, [_P, _P, _P, _I, _L, _I, _I, _P, _L, _P, _P]),
    "sdb_round_candidates": (C.c_int, [_P, _P, _L, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P]),
    "sdb_rows_compact": (C.c_int, [_P, _P, _P, _L, _I, _I, _L, _P, _P, _P, _P]),
    "sdb_minres": (C.c_int, [_P, _L, _P, _P, _P, C.c_double, _L, _P, _P, _P]),
    "sdb_dgemm": (C.c_int, [_I, _I, _I, C.c_double, _P, _L, _L, _P, _L, _L, C.c_double, _P, _L, _I, _P, _P]),
    "sdb_launch_count": (C.c_int64, [_I]),
    "sdb_fp64_probe": (C.c_int, [_L, _I, _P, _P]),
    "sdb_fp32_probe": (C.c_int, [_L, _I, _P, _P]),
    "sdb_fp64_probe3": (C.c_int, [_L, _I, _P, _P]),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    """The loaded shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("%s is not built: run `python -m surrdamh_b200.build` (needs nvcc). "
                              "surrdamh_b200 has no CPU fallback." % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        missing = [name for name in _SIGNATURES if not hasattr(h, name)]
        if missing:
            raise ImportError("%s does not export %s (stale build? run `python -m surrdamh_b200.build --force`)"
                              % (LIB_PATH, ", ".join(missing)))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.sdb_abi_version() != ABI_VERSION:
            raise ImportError("ABI version mismatch: library %d, binding %d" % (h.sdb_abi_version(), ABI_VERSION))
        _lib = h
    return _lib


def check(code):
    if code != OK:
        raise SdbError(code, lib().sdb_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SdbError(ERR_ARG, "expected a CUDA tensor (surrdamh_b200 has no CPU path)")
    if not t.is_contiguous():
        raise SdbError(ERR_ARG, "expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream_handle(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def require_cuda():
    if not torch.cuda.is_available():
        raise SdbError(ERR_CUDA, "no CUDA device: surrdamh_b200 runs on sm_100a GPUs only (no CPU fallback)")


def device_info():
    v = [C.c_int() for _ in range(4)]
    check(lib().sdb_device_info(*[C.byref(x) for x in v]))
    return dict(sm_count=v[0].value, cc=(v[1].value, v[2].value), smem_optin=v[3].value)


def dev(x, dtype=torch.float64, device=None):
    """numpy / list / scalar -> contiguous CUDA tensor."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=device or "cuda", dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device or "cuda").contiguous()
