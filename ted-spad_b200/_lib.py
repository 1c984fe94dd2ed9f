"""ctypes binding of the C ABI declared in include/tedspad.h.

The shared library is built in-tree (`make -C ted-spad_b200/csrc`, or `__graft_entry__.build()`)
as `ted-spad_b200/libtedspad.so`.  There is deliberately no fallback: if the library is missing
or a call fails, a Python exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtedspad.so")

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
FEED_AUTO, FEED_FLAT_TMA, FEED_GATHER = 0, 1, 2
RESAMPLE_AA_FLOAT, RESAMPLE_PIL_U8 = 0, 1
SLAB_3X3, SLAB_STEM2D, SLAB_STEM3D, SLAB_3X3_STREAM, SLAB_3X3_PAIR, SLAB_3X3_STREAM_PAIR, SLAB_STEM3D_PAIR = 0, 1, 2, 3, 4, 5, 6
SLAB_3X3_KX_PAIR = 7
SLAB_MAX_MMA = 112
ABI_VERSION = 9


class TensorDesc(C.Structure):
    """struct tedspad_tensor"""
    _fields_ = [("ptr", C.c_void_p)] + [(n, C.c_int32) for n in
                                        ("N", "D", "H", "W", "C", "pd", "ph", "pw", "ld", "coff")]


class ConvDesc(C.Structure):
    """struct tedspad_conv"""
    _fields_ = [("x", TensorDesc), ("y", TensorDesc), ("w", C.c_void_p), ("bias", C.c_void_p),
                ("res", C.c_void_p)] + [(n, C.c_int32) for n in (
                    "res_ld", "res_coff", "Cout", "Cout_pad", "K_pad", "kd", "kh", "kw", "sd", "sh", "sw",
                    "pd", "ph", "pw", "act", "y_fp32", "feed", "n_tile", "max_ctas")] + [
                    ("y2", TensorDesc), ("y3", TensorDesc), ("y2_begin", C.c_int32), ("y3_begin", C.c_int32)]


class ConvSlabDesc(C.Structure):
    """struct tedspad_conv_slab"""
    _fields_ = [("x", TensorDesc), ("y", TensorDesc), ("w_image", C.c_void_p), ("bias", C.c_void_p),
                ("pool", TensorDesc), ("up", TensorDesc), ("oc_w", C.c_void_p), ("oc_b", C.c_void_p), ("oc_planes", C.c_void_p),
                ("oc_frames", C.c_void_p)] + [(n, C.c_int32) for n in (
                    "kind", "Cout", "Cout_pad", "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph", "pw", "act", "tm",
                    "max_ctas", "n_tile", "K_pad")] + [("res", C.c_void_p), ("res_ld", C.c_int32), ("res_coff", C.c_int32), ("oc_clip", TensorDesc), ("oc_T", C.c_int32), ("stack_rows", C.c_int32)]


class SlabPlan(C.Structure):
    """struct tedspad_slab_plan"""
    _fields_ = ([(n, C.c_int32) for n in ("tm", "n_tile", "k_stages", "n_mma", "stages", "tmem_cols", "n_grp", "nk",
                                                  "a_kstep", "b_kstep")] +
                [("box", C.c_int32 * 5), ("tdim", C.c_int32 * 5), ("tstride", C.c_int64 * 4), ("tbase_off", C.c_int64)] +
                [(n, C.c_int32) for n in (
                    "swizzle128", "merged_cw", "slab_bytes", "slab_stride", "w_bytes", "smem_bytes", "a_layout", "a_lbo", "a_sbo",
                    "b_layout", "b_lbo", "b_sbo", "half_a_off", "c_step", "x_step", "x_off", "y_step", "y_off",
                    "z_step", "z_off", "z_kstep", "tiles_x", "tiles_y", "tiles_z", "total_tiles", "b_stream", "b_stages",
                    "b_stride", "cb_n", "cin", "num_n_tiles", "tab_per_stage", "up_cb_first", "stack_hp", "stack_ph", "stack_n", "acc_stages", "pair")] +
                [("tab", C.c_uint32 * (2 * SLAB_MAX_MMA))])


# every symbol include/tedspad.h declares: (name, restype, argtypes)
_TP = C.POINTER(TensorDesc)
_I = C.c_int32
_V = C.c_void_p
SYMBOLS = {
    "tedspad_conv_forward": (C.c_int, [C.POINTER(ConvDesc), _V]),
    "tedspad_conv_slab_plan": (C.c_int, [C.POINTER(ConvSlabDesc), C.POINTER(SlabPlan)]),
    "tedspad_conv_slab_pack": (C.c_int, [_I, _V, _I, _I, _I, _I, _I, _I, _I, _V, C.POINTER(C.c_int64), _V]),
    "tedspad_conv_slab_forward": (C.c_int, [C.POINTER(ConvSlabDesc), _V]),
    "tedspad_planes_to_clip": (C.c_int, [_V, _TP, _I, _V]),
    "tedspad_maxpool": (C.c_int, [_TP, _TP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _V]),
    "tedspad_upsample2x": (C.c_int, [_TP, _TP, _V]),
    "tedspad_upsample2x_nearest": (C.c_int, [_TP, _TP, _V]),
    "tedspad_frames_to_clip": (C.c_int, [_TP, _TP, _I, _I, _V, _V]),
    "tedspad_outconv_sigmoid": (C.c_int, [_TP, _V, _V, _TP, _I, _V, _V]),
    "tedspad_avgpool_features": (C.c_int, [_TP, _I, _V, _V]),
    "tedspad_l2_normalize_rows": (C.c_int, [_V, _I, _I, C.c_float, _V]),
    "tedspad_mgfn_rows": (C.c_int, [_V, _I, _I, _I, _V, _I, _I, _V, _V]),
    "tedspad_preprocess": (C.c_int, [_V, _I, _I, _I, _V, _I, _I, _I, _TP, _I, _V, _V]),
    "tedspad_nchw_to_cl": (C.c_int, [_V, _I, _TP, _V]),
    "tedspad_abi_version": (C.c_int, []),
    "tedspad_abi_layout": (C.c_int, [C.POINTER(C.c_int32), C.c_int32]),
    "tedspad_num_sms": (C.c_int, []),
    "tedspad_last_error": (C.c_char_p, []),
}

_lib = None


def lib():
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.tedspad_abi_version() != ABI_VERSION:
            raise ImportError(f"{LIB_PATH}: ABI version {l.tedspad_abi_version()} != {ABI_VERSION}; rebuild")
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().tedspad_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
