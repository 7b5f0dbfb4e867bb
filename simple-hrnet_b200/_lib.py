"""ctypes binding of libhrnet_b200.so (C ABI in include/hrnet_b200.h)."""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LOCK = threading.Lock()


class HrnetError(RuntimeError):
    """Raised when a C-ABI call returns a negative HRNET_E_* code."""


class HrnetDesc(ctypes.Structure):
    _fields_ = [("arch", ctypes.c_int32), ("c", ctypes.c_int32), ("nof_joints", ctypes.c_int32),
                ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("max_batch", ctypes.c_int32),
                ("flags", ctypes.c_uint32), ("tune", ctypes.c_int32 * 24)]


class HrnetParamInfo(ctypes.Structure):
    _fields_ = [("conv_key", ctypes.c_char * 96), ("bn_key", ctypes.c_char * 96),
                ("cout", ctypes.c_int32), ("cin", ctypes.c_int32), ("kh", ctypes.c_int32), ("kw", ctypes.c_int32),
                ("kind", ctypes.c_int32), ("sub_a", ctypes.c_int32), ("sub_b", ctypes.c_int32),
                ("has_bias", ctypes.c_int32), ("w_f32", ctypes.c_int32),
                ("w_offset", ctypes.c_uint64), ("scale_offset", ctypes.c_uint64), ("bias_offset", ctypes.c_uint64)]


ARCH_HRNET, ARCH_POSERESNET = 0, 1
FLAG_FORCE_SIMT, FLAG_NO_GRAPH, FLAG_FUSE_F32, FLAG_SERIAL, FLAG_NO_PATCH, FLAG_PARTITION, FLAG_GROUP = 1, 2, 4, 8, 16, 32, 64
FLAG_NO_CHAIN = 128
TUNE_CHAIN_SHARE0, TUNE_CHAIN_GRID_CAP, TUNE_CHAIN_DEBUG = 0, 4, 5
TUNE_IGEMM_PAIR, TUNE_IGEMM_PAIR_MIN_K, TUNE_PATCH_PAIR_MIN_COUT, TUNE_PATCH_PAIR_MAX_COUT = 6, 7, 8, 9
TUNE_EPILOGUE, TUNE_BPS, TUNE_IGEMM_MMA2, TUNE_PATCH_MMA2, TUNE_PATCH_NACC, TUNE_NO_PDL, TUNE_DEBUG, TUNE_GRID_CAP = 10, 11, 12, 13, 14, 15, 16, 17
TUNE_CHAIN_M2 = 18
TUNE_XUNIT = 19
TUNE_CHAIN_PAIR = 20
TUNE_CHAIN_SKIP = 21
TUNE_CHAIN_STAGES = 22
TUNE_CHAIN_EARLY = 23
TUNE_COUNT = 24
EPI_AUTO, EPI_DIRECT, EPI_TMA, EPI_COAL, EPI_TMA_PATCH, EPI_TMA_IGEMM, EPI_BATCH = 0, 1, 2, 3, 4, 5, 6

# every symbol include/hrnet_b200.h declares: (name, restype, argtypes)
_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
SYMBOLS = {
    "hrnet_plan_create": (_i, [ctypes.POINTER(HrnetDesc), ctypes.POINTER(_vp)]),
    "hrnet_plan_destroy": (None, [_vp]),
    "hrnet_last_error": (ctypes.c_char_p, []),
    "hrnet_plan_workspace_bytes": (_i, [_vp, ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "hrnet_plan_num_params": (_i, [_vp]),
    "hrnet_plan_param_info": (_i, [_vp, _i, ctypes.POINTER(HrnetParamInfo)]),
    "hrnet_plan_describe": (_i, [_vp, ctypes.c_char_p, _sz, ctypes.POINTER(_sz)]),
    "hrnet_plan_bind": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "hrnet_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hrnet_forward_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hrnet_forward_u8": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hrnet_forward_host_u8": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hrnet_forward_host_u8_async": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hrnet_plan_launch_count": (_i, [_vp]),
    "hrnet_profile_ops": (_i, [_vp, _vp, _i, ctypes.POINTER(ctypes.c_float), _i, _vp]),
    "hrnet_debug_set_tune": (None, [ctypes.POINTER(ctypes.c_int32)]),
    "hrnet_conv_bn_act": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "hrnet_fuse": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_i), ctypes.POINTER(_i), _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "hrnet_argmax": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hrnet_final_preds": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hrnet_flip_average": (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_int32), _i, _i, _i, _i, _vp, _vp]),
    "hrnet_resize_cubic_u8": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hrnet_crop_resize_bilinear_u8": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "hrnet_conv_bench": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                              ctypes.POINTER(ctypes.c_float), _vp]),
}


def library_path():
    return os.path.join(HERE, "libhrnet_b200.so")


def load_library(build_if_missing=True):
    """Loads (building in-tree with nvcc if needed) the CUDA library.  Raises if it cannot: the
    product path has no fallback."""
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = library_path()
        if build_if_missing:
            from . import build as _build
            if _build.needs_build():
                _build.build()
        if not os.path.exists(path):
            raise HrnetError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
        return lib


def set_debug_tune(tune=None):
    """Tuning knobs of the single-op entry points ({HRNET_TUNE_* index: value}; None = defaults)."""
    lib = load_library()
    if not tune:
        lib.hrnet_debug_set_tune(None)
        return
    arr = (ctypes.c_int32 * TUNE_COUNT)()
    for k, v in tune.items():
        arr[int(k)] = int(v)
    lib.hrnet_debug_set_tune(arr)


def check(rc, lib=None):
    if rc < 0:
        lib = lib or load_library()
        msg = lib.hrnet_last_error()
        raise HrnetError(f"hrnet_b200 error {rc}: {msg.decode() if msg else '?'}")
    return rc
