"""ctypes bindings of include/dqmc_gpu.h.  No torch types cross this boundary."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdqmc_b200.so")

c_i32, c_u32, c_u64, c_f64, c_vp, c_sz = (ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_double,
                                          ctypes.c_void_p, ctypes.c_size_t)


class DqmcError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the C++ shim throws GeneralError instead)."""


class DqmcParams(ctypes.Structure):
    _fields_ = [("model", c_i32), ("opdim", c_i32), ("L", c_i32), ("m", c_i32), ("s", c_i32), ("bc", c_i32),
                ("weakZflux", c_i32), ("delaySteps", c_i32), ("globalShift", c_i32),
                ("globalUpdateInterval", c_i32), ("checkerboard", c_i32), ("denseHopping", c_i32),
                ("dtau", c_f64), ("r", c_f64), ("c", c_f64), ("u", c_f64), ("lambda_", c_f64),
                ("txhor", c_f64), ("txver", c_f64), ("tyhor", c_f64), ("tyver", c_f64),
                ("mux", c_f64), ("muy", c_f64), ("accRatio", c_f64), ("t", c_f64), ("U", c_f64), ("mu", c_f64),
                ("wolffClusterUpdate", c_i32), ("wolffClusterShiftUpdate", c_i32), ("repeatWolffPerSweep", c_i32),
                ("repeatUpdateInSlice", c_i32)]


class ControlData(ctypes.Structure):
    _fields_ = [("phiDelta", c_f64), ("lastAccRatioLocal_phi", c_f64), ("ra_average", c_f64),
                ("ra_samples_added", c_i32), ("ra_count", c_i32), ("ra_values", c_f64 * 100),
                ("acceptedGlobalShifts", c_u32), ("attemptedGlobalShifts", c_u32),
                ("acceptedWolffClusterUpdates", c_u32), ("attemptedWolffClusterUpdates", c_u32),
                ("acceptedWolffClusterShiftUpdates", c_u32), ("attemptedWolffClusterShiftUpdates", c_u32),
                ("addedWolffClusterSize", c_f64)]


# every symbol include/dqmc_gpu.h declares: (name, restype, argtypes)
_P = ctypes.POINTER
SYMBOLS = [
    ("dqmc_create", c_i32, [_P(DqmcParams), c_i32, c_i32, _P(c_vp)]),
    ("dqmc_destroy", None, [c_vp]),
    ("dqmc_last_error", ctypes.c_char_p, [c_vp]),
    ("dqmc_set_stream", c_i32, [c_vp, c_vp]),
    ("dqmc_synchronize", c_i32, [c_vp]),
    ("dqmc_dims", c_i32, [c_vp, _P(c_i32)]),
    ("dqmc_set_option", c_i32, [c_vp, c_i32, c_i32]),
    ("dqmc_download_config_stream", c_i32, [c_vp, c_i32, c_vp]),
    ("dqmc_wolff_cluster_move", c_i32, [c_vp, c_i32, c_vp]),
    ("dqmc_sweep_simple", c_i32, [c_vp, c_i32]),
    ("dqmc_get_fermionic_observables", c_i32, [c_vp, c_i32, c_vp, c_vp]),
    ("dqmc_get_hubbard_observables", c_i32, [c_vp, c_i32, c_vp, c_vp]),
    ("dqmc_set_comm", c_i32, [c_vp, c_vp, c_i32, c_i32]),
    ("dqmc_exchange_payload_len", c_i32, [c_vp, c_i32]),
    ("dqmc_exchange_allgather", c_i32, [c_vp, c_i32, c_vp, c_vp]),
    ("dqmc_rng_look_ahead", c_i32, [c_vp, c_i32, c_vp, c_vp]),
    ("dqmc_rng_set_look_ahead", c_i32, [c_vp, c_i32, c_vp, ctypes.c_size_t]),
    ("dqmc_set_performed_sweeps", c_i32, [c_vp, c_u32]),
    ("dqmc_get_wolff_statistics", c_i32, [c_vp, c_i32, c_vp]),
    ("dqmc_launch_count", c_u64, [c_vp]),
    ("dqmc_rng_seed", c_i32, [c_vp, c_i32, c_u32, c_u32]),
    ("dqmc_rng_set_source", c_i32, [c_vp, c_i32, c_vp, c_vp]),
    ("dqmc_rng_draw", c_i32, [c_vp, c_i32, c_sz, c_vp]),
    ("dqmc_rng_peek", c_i32, [c_vp, c_i32, c_sz, c_vp]),
    ("dqmc_rng_skip", c_i32, [c_vp, c_i32, c_sz]),
    ("dqmc_rng_consumed", c_u64, [c_vp, c_i32]),
    ("dqmc_rng_stream_sample", c_i32, [c_u32, c_u32, c_sz, c_vp]),
    ("dqmc_init_random_fields", c_i32, [c_vp, c_i32]),
    ("dqmc_upload_fields", c_i32, [c_vp, c_i32, c_vp]),
    ("dqmc_download_fields", c_i32, [c_vp, c_i32, c_vp]),
    ("dqmc_download_green", c_i32, [c_vp, c_i32, c_i32, c_vp]),
    ("dqmc_upload_green", c_i32, [c_vp, c_i32, c_i32, c_vp]),
    ("dqmc_set_exchange_parameter", c_i32, [c_vp, c_i32, c_f64]),
    ("dqmc_get_exchange_parameter", c_i32, [c_vp, c_i32, _P(c_f64)]),
    ("dqmc_get_control_data", c_i32, [c_vp, c_i32, _P(ControlData)]),
    ("dqmc_set_control_data", c_i32, [c_vp, c_i32, _P(ControlData)]),
    ("dqmc_get_sweep_state", c_i32, [c_vp, _P(c_i32)]),
    ("dqmc_bmat_mult", c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_u32, c_u32]),
    ("dqmc_bmat_mult_device", c_i32, [c_vp, c_i32, c_i32, c_vp, c_u32, c_u32]),
    ("dqmc_bench_bmat_mult", c_i32, [c_vp, c_i32, c_vp, c_u32, c_u32, c_i32, _P(ctypes.c_float)]),
    ("dqmc_setup_storage", c_i32, [c_vp]),
    ("dqmc_wrap_up", c_i32, [c_vp, c_u32]),
    ("dqmc_wrap_down", c_i32, [c_vp, c_u32]),
    ("dqmc_advance_up", c_i32, [c_vp, c_u32]),
    ("dqmc_advance_down", c_i32, [c_vp, c_u32]),
    ("dqmc_get_green_consistency", c_i32, [c_vp, c_vp]),
    ("dqmc_logdet", c_i32, [c_vp, c_i32, c_i32, _P(c_f64)]),
    ("dqmc_green_for_timeslice", c_i32, [c_vp, c_i32, c_i32, c_u32, c_vp]),
    ("dqmc_green_from_udt_host", c_i32, [c_vp] + [c_vp] * 7 + [_P(c_f64)]),
    ("dqmc_udt_decompose_host", c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    ("dqmc_gemm_host", c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    ("dqmc_update_slice", c_i32, [c_vp, c_u32, c_i32, c_vp]),
    ("dqmc_global_shift_move", c_i32, [c_vp, c_vp]),
    ("dqmc_phi_action", c_i32, [c_vp, c_vp]),
    ("dqmc_sweep", c_i32, [c_vp, c_i32]),
    ("dqmc_rng_preload", c_i32, [c_vp, c_i32]),
    ("dqmc_rng_release", c_i32, [c_vp]),
    ("dqmc_profile_enable", c_i32, [c_vp, c_i32]),
    ("dqmc_profile_get", c_i32, [c_vp, c_vp, c_vp]),
    ("dqmc_profile_name", ctypes.c_char_p, [c_i32]),
    ("dqmc_accepted_total", c_i32, [c_vp, c_vp]),
    ("dqmc_exchange_actions", c_i32, [c_vp, c_vp, c_vp]),
    ("dqmc_exchange_pack", c_i32, [c_vp, c_vp, c_i32]),
    ("dqmc_exchange_apply", c_i32, [c_vp, c_vp, c_vp, c_i32]),
    ("dqmc_exchange_probability", c_f64, [c_f64, c_f64, c_f64, c_f64]),
    ("dqmc_exchange_walk", c_i32, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, _P(c_i32), c_vp]),
]

_lib = None


def load_library():
    """Load libdqmc_b200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python detqmc_b200/build.py` "
                          "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
