"""ctypes binding of libtramp_b200.so (the C ABI declared in include/tramp_b200.h).

There is no CPU fallback: if the shared library is missing, or a kernel is
asked to run without a CUDA device, the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtramp_b200.so")

c_double_p = C.c_void_p  # device pointers travel as integers
c_int_p = C.c_void_p


class TrbFactor(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("_pad", C.c_int32),
        ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double), ("p3", C.c_double),
        ("amin", C.c_double), ("amax", C.c_double),
    ]


class TrbSweep(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("M", C.c_int32), ("R", C.c_int32),
        ("ldn", C.c_int32), ("ldm", C.c_int32), ("rank", C.c_int32), ("nslots", C.c_int32),
        ("prior", TrbFactor), ("lik", TrbFactor),
        ("lin_amin", C.c_double), ("lin_amax", C.c_double),
        ("Vt", C.c_void_p), ("strideV", C.c_int64),
        ("Ut", C.c_void_p), ("strideU", C.c_int64),
        ("s", C.c_void_p), ("s2", C.c_void_p), ("stride_s", C.c_int64),
        ("y", C.c_void_p), ("x_true", C.c_void_p),
        ("edge_a", C.c_void_p),
        ("b1", C.c_void_p), ("b3", C.c_void_p), ("b5", C.c_void_p), ("b7", C.c_void_p),
        ("b6_init", C.c_void_p), ("b8_init", C.c_void_p),
        ("damp1", C.c_double), ("damp3", C.c_double), ("damp5", C.c_double), ("damp7", C.c_double),
        ("rx", C.c_void_p), ("rz", C.c_void_p), ("vx", C.c_void_p), ("vz", C.c_void_p),
        ("tz", C.c_void_p), ("tx", C.c_void_p), ("coef", C.c_void_p), ("part", C.c_void_p),
        ("scr_n", C.c_void_p), ("scr_m", C.c_void_p), ("vlin", C.c_void_p), ("stats", C.c_void_p),
        ("active", C.c_void_p), ("flags", C.c_void_p), ("n_iter", C.c_void_p),
        ("rec_mse", C.c_void_p), ("rec_smse", C.c_void_p), ("rec_vx", C.c_void_p),
        ("rec_vz", C.c_void_p), ("rec_tol", C.c_void_p),
        ("max_records", C.c_int32),
        ("es_tol", C.c_double), ("es_max_increase", C.c_double),
        ("es_wait_increase", C.c_int32), ("gemv_impl", C.c_int32), ("es_vars", C.c_int32),
        ("snap_edge_a", C.c_void_p), ("snap_b1", C.c_void_p), ("snap_b3", C.c_void_p),
        ("snap_b5", C.c_void_p), ("snap_b7", C.c_void_p), ("snap_rx", C.c_void_p),
        ("snap_rz", C.c_void_p), ("snap_vx", C.c_void_p), ("snap_vz", C.c_void_p),
        ("snap_tx", C.c_void_p),
        ("R_total", C.c_int32), ("schedule", C.c_int32),
        ("ty", C.c_void_p),
        ("comm", C.c_void_p), ("s_full", C.c_void_p), ("s2_full", C.c_void_p),
        ("es_mode", C.c_int32), ("_pad_es", C.c_int32), ("es_min_variance", C.c_double),
    ]


class TrbQuadrature(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("Q", C.c_int32), ("P", C.c_int32),
        ("kappa", C.c_double),
        ("x2", C.c_void_p), ("w2", C.c_void_p), ("Q2", C.c_int32), ("P2", C.c_int32),
        ("kappa2", C.c_double),
    ]


class TrbSe(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("channel", C.c_int32),
        ("prior", C.c_void_p), ("lik", C.c_void_p),
        ("tau_x", C.c_void_p), ("tau_z", C.c_void_p),
        ("alpha", C.c_void_p), ("mean_spectrum", C.c_void_p),
        ("s2", C.c_void_p), ("stride_s2", C.c_int64),
        ("R", C.c_int32), ("Nz", C.c_int32), ("Nx", C.c_int32), ("rank", C.c_int32),
        ("lin_amin", C.c_double), ("lin_amax", C.c_double),
        ("damp1", C.c_double), ("damp3", C.c_double), ("damp5", C.c_double), ("damp7", C.c_double),
        ("edge_a", C.c_void_p), ("vx", C.c_void_p), ("vz", C.c_void_p),
        ("active", C.c_void_p), ("flags", C.c_void_p), ("n_iter", C.c_void_p),
        ("rec_vx", C.c_void_p), ("rec_vz", C.c_void_p), ("max_records", C.c_int32),
        ("es_tol", C.c_double), ("es_min_variance", C.c_double), ("es_max_increase", C.c_double),
        ("es_wait_increase", C.c_int32), ("es_vars", C.c_int32),
        ("quad", TrbQuadrature),
    ]


# kinds (trb_factor_kind)
GAUSS_BERNOULLI_PRIOR, BINARY_PRIOR, GAUSSIAN_PRIOR = 0, 1, 2
GAUSSIAN_LIKELIHOOD, SGN_LIKELIHOOD, ABS_LIKELIHOOD = 3, 4, 5

(STAGE_PRIOR, STAGE_PROJECT_Z, STAGE_PROJECT_X_INIT, STAGE_RESCALE_FWD, STAGE_EXPAND_X,
 STAGE_Z_UPDATE, STAGE_PROJECT_X, STAGE_RESCALE_BWD, STAGE_EXPAND_Z, STAGE_X_UPDATE,
 STAGE_SNAPSHOT, STAGE_Z_UPDATE_LIGHT, STAGE_TX_RECUR, STAGE_PROJECT_Y) = range(14)

FLAG_NAN_A, FLAG_NAN_B, FLAG_NEG_A, FLAG_CONVERGED, FLAG_DIVERGED, FLAG_RESTORED = 1, 2, 4, 8, 16, 32
FLAG_COMM_TIMEOUT = 64
FLAG_SE_DOMAIN = 128
MEASURE_V, MEASURE_A = 0, 1
SE_MARCHENKO_PASTUR, SE_SPECTRUM = 0, 1
MAX_RANKS = 8

_I, _L, _D, _P = C.c_int, C.c_int64, C.c_double, C.c_void_p
_FP = C.POINTER(TrbFactor)

# name -> (restype, argtypes); must list every function of include/tramp_b200.h
SIGNATURES = {
    "trb_last_error": (C.c_char_p, []),
    "trb_version": (_I, []),
    "trb_sizeof_factor": (C.c_size_t, []),
    "trb_sizeof_sweep": (C.c_size_t, []),
    "trb_device_sm_count": (_I, []),
    "trb_profile_reset": (None, [_I]),
    "trb_profile_launches": (C.c_longlong, [_I]),
    "trb_profile_gemv_ms": (_I, [C.POINTER(C.c_double)]),
    "trb_profile_timeline": (_I, [_P, _P, _I]),
    "trb_factor_posterior": (_I, [_FP, _I, _I, _I, _P, _I, _P, _P, _P, _P, _I, _P]),
    "trb_factor_log_partition": (_I, [_FP, _I, _I, _I, _P, _I, _P, _P, _P, _I, _P]),
    "trb_factor_message": (_I, [_FP, _I, _I, _I, _P, _P, _P, _P, _P, _P, _D, _P, _P, _P, _P]),
    "trb_truncated_normal": (_I, [_I, _P, _P, _D, _D, _P, _P, _P, _P, _P]),
    "trb_posterior_rv": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "trb_lin_project": (_I, [_P, _L, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P]),
    "trb_lin_expand_slots": (_I, [_I, _I]),
    "trb_lin_expand": (_I, [_P, _L, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "trb_lin_project_gemm": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "trb_lin_expand_gemm": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P]),
    "trb_gemm_set_variant": (None, [_I]),
    "trb_message_trial": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _D, _P, _P, _P]),
    "trb_variable_log_partition": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "trb_lin_log_partition": (_I, [_I, _I, _I, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P]),
    "trb_row_dot": (_I, [_I, _I, _I, _P, _P, _P, _P]),
    "trb_message_from_posterior": (_I, [_I, _I, _I, _P, _P, _P, _P, _D, _D, _P, _P, _P]),
    "trb_rows_select": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "trb_jacobi_zsplit": (_I, [_I, _I, _I]),
    "trb_jacobi_set_waves": (None, [_I]),
    "trb_jacobi_set_fused": (None, [_I]),
    "trb_jacobi_sweep": (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _P, _D, _I, _P]),
    "trb_row_norms": (_I, [_P, _L, _I, _I, _I, _I, _P, _P]),
    "trb_rows_gather_scale": (_I, [_P, _L, _I, _P, _P, _I, _I, _I, _P, _L, _I, _P]),
    "trb_lin_reduce_slots": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "trb_lin_rescale": (_I, [_I, _I, _I, _I, _I, _I, _I, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trb_comm_create": (_I, [_I, _I, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "trb_comm_connect": (_I, [_P, C.c_char_p]),
    "trb_comm_destroy": (_I, [_P]),
    "trb_comm_all_reduce": (_I, [_P, _P, C.c_size_t, _P, _P]),
    "trb_set_cuda_graphs": (None, [_I]),
    "trb_set_fused_rescale": (None, [_I]),
    "trb_set_update_kernels": (None, [_I]),
    "trb_set_persistent_sweep": (None, [_I]),
    "trb_sweep_run": (_I, [C.POINTER(TrbSweep), _I, _I, _I, _P]),
    "trb_sweep_stage": (_I, [C.POINTER(TrbSweep), _I, _I, _I, _I, _P]),
    "trb_sizeof_se": (C.c_size_t, []),
    "trb_se_measure": (_I, [_P, _I, _I, _I, _P, _P, C.POINTER(TrbQuadrature), _P, _P, _P]),
    "trb_se_run": (_I, [C.POINTER(TrbSe), _I, _I, _P]),
}

_lib = None


class TrbError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TrbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C tramp_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if (lib.trb_sizeof_factor() != C.sizeof(TrbFactor) or lib.trb_sizeof_sweep() != C.sizeof(TrbSweep)
            or lib.trb_sizeof_se() != C.sizeof(TrbSe)):
        raise TrbError("tramp_b200._lib struct layout is out of sync with include/tramp_b200.h")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().trb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise TrbError(f"tramp_b200 error {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise TrbError("tramp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
