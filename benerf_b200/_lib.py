"""ctypes binding of libbenerf_b200.so (include/benerf_b200.h).

The library is the product: if it is missing or fails to load this module raises -- there
is no Python/PyTorch fallback for any arithmetic on the render path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbenerf_b200.so")

OK, ERR_ARG, ERR_DEVICE, ERR_CUDA, ERR_STATE, ERR_NCCL = 0, -1, -2, -3, -4, -5
MLP_TC_FP16X2, MLP_SIMT_FP32, MLP_TC_1CTA, MLP_TC_PAIR_SS = 0, 1, 2, 3
NUM_LINEARS = 12
# order of the 12 linears expected by bnrf_set_weights (reference state-dict order)
LINEAR_NAMES = [f"pts_linears.{i}" for i in range(8)] + ["views_linears.0", "feature_linear", "alpha_linear", "rgb_linear"]


class Cfg(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_importance", C.c_int32), ("channels", C.c_int32), ("ndc", C.c_int32),
                ("near_", C.c_float), ("far_", C.c_float), ("mlp_mode", C.c_int32), ("gemm_mode", C.c_int32)]


class Rng(C.Structure):
    _fields_ = [("t_rand", C.c_void_p), ("noise_c", C.c_void_p), ("u", C.c_void_p), ("noise_f", C.c_void_p),
                ("z_fine", C.c_void_p), ("seed", C.c_uint64), ("offset", C.c_uint64), ("ray_base", C.c_uint64), ("offset_dev", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "sigma", "depth_map", "z_vals")]


class RenderSeg(C.Structure):
    _fields_ = [("poses", C.c_void_p), ("ray_idx", C.c_void_p), ("P", C.c_int32), ("R", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("K", C.c_float * 9), ("remap", C.c_void_p)]


class AdamGroup(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64), ("lr", C.c_float), ("active", C.c_int32)]


class LossCfg(C.Structure):
    _fields_ = [("channels", C.c_int32), ("n_poses", C.c_int32), ("log_mode", C.c_int32), ("event_loss", C.c_int32),
                ("rgb_loss", C.c_int32), ("event_threshold", C.c_float), ("event_coeff_syn", C.c_float),
                ("event_coeff_real", C.c_float), ("rgb_coeff", C.c_float)]


class AdamSchedGroup(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64), ("lr0", C.c_float), ("decay_rate", C.c_float), ("active", C.c_int32)]


class ParamGrads(C.Structure):
    """bnrf_param_grads: 12 weight + 12 bias gradient pointers (PyTorch layouts, accumulated into)."""
    _fields_ = [("weights", C.c_void_p * 12), ("biases", C.c_void_p * 12)]


_P, _I, _L, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
# name -> (restype, argtypes); the loader checks that every symbol of the header is exported
PROTOTYPES = {
    "bnrf_abi_version": (_I, []),
    "bnrf_create": (_I, [C.POINTER(_P), _I, C.POINTER(Cfg)]),
    "bnrf_destroy": (None, [_P]),
    "bnrf_last_error": (C.c_char_p, [_P]),
    "bnrf_set_weights": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_P), _P]),
    "bnrf_set_encoding_weights": (_I, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P]),
    "bnrf_set_weights_pair": (_I, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P]),
    "bnrf_set_sample_grid": (_I, [_P, C.POINTER(C.c_float), _I, _P]),
    "bnrf_spline_poses": (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "bnrf_spline_poses_pair": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "bnrf_spline_poses_pair_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "bnrf_workspace_bytes": (_Z, [_P, _L]),
    "bnrf_render_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(C.c_float), _P, C.POINTER(Rng), C.POINTER(Outputs), _P, _Z, _P]),
    "bnrf_saved_bytes": (_Z, [_P, _L]),
    "bnrf_backward_workspace_bytes": (_Z, [_P, _L]),
    "bnrf_render_forward_train": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(C.c_float), _P, C.POINTER(Rng), C.POINTER(Outputs), _P, _Z, _P, _Z, _P]),
    "bnrf_render_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(C.c_float), _P, _P, _P, _P, _Z, C.POINTER(ParamGrads), C.POINTER(ParamGrads), _P, _P, _Z, _P]),
    "bnrf_render_forward_multi": (_I, [_P, C.POINTER(RenderSeg), _I, C.POINTER(Rng), C.POINTER(Outputs), _P, _Z, _P, _Z, _P]),
    "bnrf_render_backward_multi": (_I, [_P, C.POINTER(RenderSeg), _I, _P, _P, _P, _Z, C.POINTER(ParamGrads), C.POINTER(ParamGrads),
                                        C.POINTER(_P), _P, _Z, _P]),
    "bnrf_spline_poses_backward": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "bnrf_debug_sgemm": (_I, [_P, _I, _I, _L, _I, _L, _P, _L, _P, _L, _P, _L, _I, _P, _L, _P, _L, _P, _P]),
    "bnrf_op_rays": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(C.c_float), _P, _P, _P, _P, _P]),
    "bnrf_op_stratified": (_I, [_P, _P, _L, _I, _P, _P]),
    "bnrf_op_mlp": (_I, [_P, _I, _P, _P, _P, _P, _L, _I, _P, _P]),
    "bnrf_op_composite": (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P]),
    "bnrf_op_resample": (_I, [_P, _P, _P, _P, _L, _I, _I, _P, _P]),
    "bnrf_blur_mean": (_I, [_P, _I, _L, _I, _P, _P]),
    "bnrf_event_logdiff": (_I, [_P, _I, _L, _I, _I, _P, _P]),
    "bnrf_blur_mean_backward": (_I, [_P, _I, _L, _I, _P, _P]),
    "bnrf_event_logdiff_backward": (_I, [_P, _P, _I, _L, _I, _I, _P, _P]),
    "bnrf_accumulate_events": (_I, [_P, _P, _P, _L, _I, _I, _P, _P]),
    "bnrf_accumulate_events_binned": (_I, [_P, _P, _P, _L, _P, _I, _I, _I, _P, _P]),
    "bnrf_training_loss_workspace_bytes": (_Z, [_L]),
    "bnrf_training_loss": (_I, [C.POINTER(LossCfg), _P, _P, _P, _P, _L, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P]),
    "bnrf_training_loss_finish": (_I, [C.POINTER(LossCfg), _P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _P]),
    "bnrf_crf_forward": (_I, [_I, _I, C.POINTER(_P), C.POINTER(_P), _P, _L, _P, _P]),
    "bnrf_crf_backward": (_I, [_I, _I, C.POINTER(_P), C.POINTER(_P), _P, _P, _L, _P, C.POINTER(_P), C.POINTER(_P), _P]),
    "bnrf_adam_step": (_I, [_P, _P, _P, _P, _L, C.POINTER(AdamGroup), _I, _L, C.c_float, C.c_float, C.c_float, C.c_float, _I, _P]),
    "bnrf_adam_step_sched": (_I, [_P, _P, _P, _P, _L, C.POINTER(AdamSchedGroup), _I, _P, C.c_double, C.c_float, C.c_float, C.c_float, C.c_float, _I, _P, _P]),
    "bnrf_step_advance": (_I, [_P, _P]),
    "bnrf_wait_fine_gradients": (_I, [_P, _P]),
    "bnrf_profile": (_I, [_P, _I]),
    "bnrf_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "bnrf_debug_mlp_trace": (_I, [_P, _P]),
    "bnrf_debug_umma_probe": (_I, [_P, _P, _I, _I, _P, _P]),
    "bnrf_debug_tc3_schedule": (_I, [_I, _P, _I, _P]),
    "bnrf_debug_umma_probe_fmt": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "bnrf_debug_umma_ts_probe": (_I, [_P, _P, _I, _I, _P, _P]),
    "bnrf_debug_tile_dgrad": (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P]),
    "bnrf_debug_tile_wgrad": (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library once and type its entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m benerf_b200.build` "
            "(benerf_b200 has no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.bnrf_abi_version() != 2:
        raise RuntimeError("libbenerf_b200.so ABI version mismatch")
    _lib = lib
    return lib


class BnrfError(RuntimeError):
    pass
