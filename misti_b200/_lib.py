"""ctypes binding of libmisti_b200.so (include/misti_b200.h).

There is no CPU implementation behind this module: if the CUDA library is missing, or no CUDA
device can be opened, the errors raised here propagate to the caller.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MISTI_B200_LIB") or os.path.join(HERE, "libmisti_b200.so")  # the variable is a development knob

ABI_VERSION = 2
MAX_BANDS, MAX_PULSES, MAX_PARAMS = 8, 8, 16

FLAG_CORRECT, FLAG_CPFIT, FLAG_SMOOTH, FLAG_UNFOLDED, FLAG_DEVICE_PTRS = 1, 2, 4, 8, 256
OK, NEGATIVE_PARAM, CORRECTION_FAILED, NONFINITE, INFINITE_COAL_TIME, STIFF, SKIPPED = 0, 1, 2, 3, 4, 5, 6
E_ARG, E_CUDA, E_NODEV = -1, -2, -3

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class ModelDesc(ctypes.Structure):
    _fields_ = [("grid_id", ctypes.c_int32), ("split_t", ctypes.c_int32), ("sample_date", ctypes.c_int32),
                ("n_bands", ctypes.c_int32), ("n_pulses", ctypes.c_int32), ("n_params", ctypes.c_int32),
                ("band_pop", ctypes.c_int32 * MAX_BANDS), ("band_start", ctypes.c_int32 * MAX_BANDS),
                ("band_end", ctypes.c_int32 * MAX_BANDS), ("band_opt", ctypes.c_int32 * MAX_BANDS),
                ("pulse_pop", ctypes.c_int32 * MAX_PULSES), ("pulse_time", ctypes.c_int32 * MAX_PULSES),
                ("pulse_opt", ctypes.c_int32 * MAX_PULSES),
                ("band_val", ctypes.c_double * MAX_BANDS), ("pulse_val", ctypes.c_double * MAX_PULSES)]


class EvalIO(ctypes.Structure):
    _fields_ = [("lc_inject", ctypes.c_void_p), ("jafs", ctypes.c_void_p), ("jafs_raw", ctypes.c_void_p),
                ("lc_out", ctypes.c_void_p), ("pr_out", ctypes.c_void_p), ("status", ctypes.c_void_p),
                ("nfev", ctypes.c_void_p), ("terms", ctypes.c_void_p), ("row_ids", ctypes.c_void_p),
                ("solve_trace", ctypes.c_void_p), ("row_best_llh", ctypes.c_void_p), ("row_best_item", ctypes.c_void_p)]


class FitOpts(ctypes.Structure):
    _fields_ = [("xatol", ctypes.c_double), ("fatol", ctypes.c_double), ("maxiter", ctypes.c_int64), ("maxfev", ctypes.c_int64),
                ("niter", ctypes.c_int32), ("interval", ctypes.c_int32), ("T", ctypes.c_double), ("stepsize", ctypes.c_double),
                ("target_accept_rate", ctypes.c_double), ("stepwise_factor", ctypes.c_double), ("rng_state", ctypes.c_void_p)]


class FitResult(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("fun", ctypes.c_void_p), ("nit", ctypes.c_void_p), ("nfev", ctypes.c_void_p),
                ("status", ctypes.c_void_p), ("accepted", ctypes.c_void_p), ("failures", ctypes.c_void_p),
                ("rounds", ctypes.c_int64), ("points", ctypes.c_int64), ("graph", ctypes.c_int32)]


class MistiLibraryError(RuntimeError):
    pass


# every symbol include/misti_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "misti_abi_version": (ctypes.c_int, []),
    "misti_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "misti_ctx_destroy": (None, [ctypes.c_void_p]),
    "misti_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "misti_ctx_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "misti_ctx_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "misti_ctx_reserve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "misti_add_grid": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p, c_int32_p]),
    "misti_add_model": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ModelDesc), c_int32_p]),
    "misti_clear_models": (ctypes.c_int, [ctypes.c_void_p]),
    "misti_set_data": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p, ctypes.c_int32]),
    "misti_eval_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int32, ctypes.c_uint32, ctypes.c_double, ctypes.c_void_p,
                                        ctypes.POINTER(EvalIO)]),
    "misti_nelder_mead": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_double_p, c_int32_p, c_int32_p,
                                         ctypes.c_uint32, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                         ctypes.c_int64, c_double_p, c_double_p, ctypes.POINTER(ctypes.c_int64),
                                         ctypes.POINTER(ctypes.c_int64), c_int32_p, ctypes.POINTER(ctypes.c_int64)]),
    "misti_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_double_p, c_int32_p, c_int32_p, ctypes.c_uint32,
                                 ctypes.c_double, ctypes.POINTER(FitOpts), ctypes.POINTER(FitResult)]),
    "misti_coalescent_rates": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_double_p, ctypes.c_double,
                                              ctypes.c_double, c_double_p, c_double_p]),
    "misti_score_spectra": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p]),
    "misti_last_kernel_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "misti_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "misti_generator": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, c_double_p]),
    "misti_pulse": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_double, ctypes.c_int32, c_double_p]),
    "misti_ancient_reset": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p]),
    "misti_state_to_jaf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_int32_p]),
}

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MistiLibraryError(
            "%s is missing: build it with `python -m misti_b200._build` (needs nvcc). misti_b200 has no CPU fallback."
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not match the header
        fn.restype = res
        fn.argtypes = args
    if lib.misti_abi_version() != ABI_VERSION:
        raise MistiLibraryError("libmisti_b200.so ABI version mismatch")
    _lib = lib
    return lib
