"""ctypes binding of libsmow_b200.so (the C ABI declared in include/smow_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this
module raises.  Build it with ``python -m smow_net_b200.build``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmow_b200.so")

ABI_VERSION = 9
F32, BF16 = 0, 1
NCDHW, NDHWC = 0, 1

_vp, _fp, _i, _i64 = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64

# name -> (restype, argtypes); mirrors include/smow_b200.h one to one
SIGNATURES = {
    "smow_abi_version": (_i, []),
    "smow_last_error": (ctypes.c_char_p, []),
    "smow_launch_count": (ctypes.c_uint64, []),
    "smow_set_option": (_i, [ctypes.c_char_p, _i]),
    "smow_get_option": (_i, [ctypes.c_char_p]),
    "smow_warp_stack_fwd": (_i, [_vp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "smow_warp_pair_fwd": (_i, [_vp, _vp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "smow_warp_bwd_workspace_bytes": (_i64, [_i, _i, _i]),
    "smow_warp_stack_bwd": (_i, [_vp, _vp, _fp, _fp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "smow_warp_pair_bwd": (_i, [_vp, _vp, _vp, _fp, _fp, _fp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "smow_tlerp_cat_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp]),
    "smow_tlerp_pair_cat_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp]),
    "smow_tlerp_cat_bwd": (_i, [_vp, _vp, _i, _i, _i, _i64, _i, _i, _vp]),
    "smow_tlerp_pair_cat_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp]),
    "smow_act_tlerp_cat_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, ctypes.c_float, _i, _vp]),
    "smow_act_tlerp_cat_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, ctypes.c_float, _i, _vp]),
    "smow_frame_mix_stats_parts": (_i, [_i, _i, _i, _i64]),
    "smow_frame_mix_apply_tc_stats": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i64, _i64, _i, _i, _vp]),
    "smow_bn_finalize": (_i, [_fp, _i, _i, _i64, _fp, _fp, _fp, _fp, ctypes.c_float, ctypes.c_float, _fp, _vp]),
    "smow_bn_act_tlerp_cat_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i64, _i64, ctypes.c_float, _vp]),
    "smow_bn_act_bwd_workspace_bytes": (_i64, [_i, _i, _i64]),
    "smow_bn_act_bwd_reduce": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i64, ctypes.c_float, _vp, _i64, _vp]),
    "smow_bn_act_tlerp_cat_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i64, _i64, ctypes.c_float, _vp]),
    "smow_tokenizer_workspace_bytes": (_i64, [_i, _i, _i64]),
    "smow_tokenizer_fwd": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _i64, _i, _i, _vp, _i64, _vp]),
    "smow_tokenizer_bwd": (_i, [_fp, _vp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i64, _i, _i, _vp, _i64, _vp]),
    "smow_warp_tokenizer_supported": (_i, [_i]),
    "smow_warp_tokenizer_fwd": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "smow_warp_tokenizer_bwd": (_i, [_fp, _vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _i,
                                     _vp, _i64, _vp]),
    "smow_frame_mix_supported": (_i, [_i]),
    "smow_frame_mix_apply": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i64, _i, _i, _vp]),
    "smow_frame_mix_wgrad_workspace_bytes": (_i64, [_i, _i, _i64]),
    "smow_frame_mix_wgrad": (_i, [_fp, _fp, _fp, _i, _i, _i64, _vp, _i64, _vp]),
    "smow_frame_mix_tc_supported": (_i, [_i, _i]),
    "smow_frame_mix_apply_tc": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i64, _i64, _i, _i, _vp]),
    "smow_frame_mix_wgrad_tc_workspace_bytes": (_i64, [_i, _i, _i, _i64]),
    "smow_frame_mix_wgrad_tc": (_i, [_fp, _fp, _fp, _i, _i, _i, _i64, _i, _i, _vp, _i64, _vp]),
    "smow_flow_head_supported": (_i, [_i, _i, _i, _i, _i]),
    "smow_flow_head_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "smow_flow_head_fwd": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "smow_flow_head_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "smow_net_b200: %s is missing - the CUDA extension is mandatory (no CPU or PyTorch "
            "fallback exists). Build it with `python -m smow_net_b200.build`." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    got = lib.smow_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError("smow_net_b200: ABI version %d, expected %d - rebuild the library" % (got, ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().smow_last_error().decode("utf-8", "replace")
        kind = "argument error" if rc < 0 else "CUDA error"
        raise RuntimeError("%s failed (%s %d): %s" % (what, kind, rc, msg))


def set_option(key, value):
    check(load().smow_set_option(key.encode(), int(value)), "smow_set_option(%s)" % key)


def get_option(key):
    return load().smow_get_option(key.encode())


def launch_count():
    return int(load().smow_launch_count())
