"""ctypes view of the C ABI (include/av1b200.h) and of the symbol-stream structs (csrc/av1b_types.h).

PyTorch / numpy are only plumbing here; all arithmetic is in libav1b200.so.  There is no CPU
fallback: if the shared library is missing, importing fails loudly.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AV1B200_LIB", os.path.join(_HERE, "libav1b200.so"))


class SeqParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32),
                ("enable_cdef", C.c_int32), ("enable_restoration", C.c_int32),
                ("fps_num", C.c_int32), ("fps_den", C.c_int32), ("color_hdr", C.c_int32), ("film_grain_present", C.c_int32),
                ("render_width", C.c_int32), ("render_height", C.c_int32)]


class FrameParams(C.Structure):
    _fields_ = [("frame_type", C.c_int32), ("base_q_idx", C.c_int32), ("disable_cdf_update", C.c_int32),
                ("tile_cols_log2", C.c_int32), ("tile_rows_log2", C.c_int32),
                ("lf_level", C.c_int32 * 4), ("lf_sharpness", C.c_int32),
                ("cdef_damping", C.c_int32), ("cdef_bits", C.c_int32),
                ("cdef_y_strength", C.c_int32 * 8), ("cdef_uv_strength", C.c_int32 * 8),
                ("lr_type", C.c_int32 * 3), ("lr_unit_shift", C.c_int32), ("lr_uv_shift", C.c_int32),
                ("non_reference", C.c_int32), ("grain_scaling", C.c_int32), ("grain_seed", C.c_int32),
                ("using_qmatrix", C.c_int32), ("qm_level", C.c_int32 * 2)]


class Geom(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mi_cols", C.c_int32), ("mi_rows", C.c_int32),
                ("w8", C.c_int32), ("h8", C.c_int32), ("sb_cols", C.c_int32), ("sb_rows", C.c_int32),
                ("stride", C.c_int32 * 3), ("rows", C.c_int32 * 3),
                ("tile_cols", C.c_int32), ("tile_rows", C.c_int32),
                ("tile_cols_log2", C.c_int32), ("tile_rows_log2", C.c_int32),
                ("tile_col_start_sb", C.c_int32 * 65), ("tile_row_start_sb", C.c_int32 * 65)]


BLOCK_INFO_DTYPE = np.dtype([("blk_log2", "u1"), ("y_mode", "u1"), ("uv_mode", "u1"), ("skip", "u1"),
                             ("angle_y", "i1"), ("angle_uv", "i1"), ("tx_type_y", "u1"), ("cfl_alpha_u", "u1"),
                             ("eob", "<u2", (3,)), ("cfl_alpha_v", "u1"), ("is_inter", "u1"),
                             ("mv", "<i2", (2,))])
assert BLOCK_INFO_DTYPE.itemsize == 20

LR_UNIT_DTYPE = np.dtype([("type", "i1"), ("sgr_set", "i1"), ("wiener_v", "i1", (3,)), ("wiener_h", "i1", (3,)),
                          ("sgr_xqd", "i1", (2,)), ("pad", "i1", (6,))])
assert LR_UNIT_DTYPE.itemsize == 16


class FrameSyms(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("coef", C.c_void_p * 3), ("coef_stride", C.c_int32 * 3),
                ("cdef_idx", C.c_void_p), ("lr_units", C.c_void_p * 3),
                ("lr_unit_cols", C.c_int32 * 3), ("lr_unit_rows", C.c_int32 * 3)]


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32),
                ("fps_num", C.c_int32), ("fps_den", C.c_int32), ("crf", C.c_int32), ("preset", C.c_int32),
                ("keyint", C.c_int32), ("lookahead", C.c_int32), ("film_grain", C.c_int32),
                ("enable_qm", C.c_int32), ("qm_min", C.c_int32), ("qm_max", C.c_int32),
                ("tile_cols_log2", C.c_int32), ("tile_rows_log2", C.c_int32), ("device_id", C.c_int32),
                ("hdr", C.c_int32), ("host_threads", C.c_int32), ("frames_in_flight", C.c_int32),
                ("gop_period", C.c_int32), ("tune", C.c_int32 * 7), ("reserved", C.c_int32 * 8)]


class FrameSrc(C.Structure):
    _fields_ = [("planes", C.c_void_p * 3), ("stride", C.c_int32 * 3)]


PACKET_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t, C.c_int64, C.c_int)
PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_double)

_lib = None


def lib():
    """Loads libav1b200.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("av1b200: %s not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.av1b_last_error.restype = C.c_char_p
        L.av1b_host_alloc.restype = C.c_void_p
        L.av1b_host_alloc.argtypes = [C.c_int, C.c_size_t]
        L.av1b_host_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def last_error():
    return lib().av1b_last_error().decode()
