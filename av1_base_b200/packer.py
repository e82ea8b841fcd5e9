"""Host-side mirror of the bitstream packing entry points of the C ABI (av1b_pack_*)."""
import ctypes as C
import numpy as np
from . import abi


def pack_sequence_header(seq):
    out = (C.c_uint8 * 256)()
    n = C.c_size_t(0)
    rc = abi.lib().av1b_pack_sequence_header(C.byref(seq), out, 256, C.byref(n))
    if rc:
        raise RuntimeError("av1b_pack_sequence_header: %d %s" % (rc, abi.last_error()))
    return bytes(out[:n.value])


def make_syms(g, blocks, coef, cdef_idx=None, lr_units=None):
    """blocks: structured array [h8*w8]; coef: 3 int16 arrays with the geometry's strides."""
    s = abi.FrameSyms()
    s.blocks = blocks.ctypes.data
    for p in range(3):
        s.coef[p] = coef[p].ctypes.data
        s.coef_stride[p] = coef[p].strides[0] // 2
    if cdef_idx is None:
        cdef_idx = np.zeros(g.sb_rows * g.sb_cols, np.uint8)
    s.cdef_idx = cdef_idx.ctypes.data
    keep = [blocks, coef, cdef_idx, lr_units]
    if lr_units is not None:
        for p in range(3):
            if lr_units[p] is not None:
                s.lr_units[p] = lr_units[p].ctypes.data
                s.lr_unit_rows[p], s.lr_unit_cols[p] = lr_units[p].shape
    s._keep = keep
    return s


def pack_frame(seq, fp, syms, n_threads=1, with_td=True, cap=None):
    cap = cap or (16 << 20)
    out = np.empty(cap, np.uint8)
    n = C.c_size_t(0)
    rc = abi.lib().av1b_pack_frame(C.byref(seq), C.byref(fp), C.byref(syms), n_threads, int(with_td),
                                   out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n))
    if rc:
        raise RuntimeError("av1b_pack_frame: %d %s" % (rc, abi.last_error()))
    return out[:n.value].tobytes()


def pack_frame_tokens(seq, fp, syms, with_td=True, cap=None):
    """Inter frame through the token path (CPU statement of the device tokenizer). Returns (bytes, n_tokens)."""
    cap = cap or (16 << 20)
    out = np.empty(cap, np.uint8)
    n = C.c_size_t(0)
    nt = C.c_uint64(0)
    rc = abi.lib().av1b_pack_frame_tokens(C.byref(seq), C.byref(fp), C.byref(syms), int(with_td),
                                          out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), C.byref(nt))
    if rc:
        raise RuntimeError("av1b_pack_frame_tokens: %d %s" % (rc, abi.last_error()))
    return out[:n.value].tobytes(), nt.value
