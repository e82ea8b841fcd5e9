"""ctypes mirror of the kernel-suite entry points of the C ABI (include/av1b200.h, av1b_k_*).
numpy is plumbing; each call runs exactly one CUDA kernel on the device and returns its result
(and the mean CUDA-event time per launch when reps > 0)."""
import ctypes as C
import numpy as np
from . import abi


class KernelError(RuntimeError):
    def __init__(self, code):
        super().__init__("av1b200 kernel suite failed with code %d: %s" % (code, abi.last_error()))
        self.code = code


def _ck(rc):
    if rc:
        raise KernelError(rc)


def _p3(planes):
    return (C.c_void_p * 3)(*[p.ctypes.data for p in planes])


def inv_txfm_add(coef, pred, w, h, tx_type, bit_depth, device=0, reps=1):
    """coef: [n, ch, cw] int32 (ch = min(h,32), cw = min(w,32)); pred: [n, h, w] uint16. Returns (recon, ms)."""
    n = coef.shape[0]
    cbuf = np.zeros((n, 1024), np.int32)
    cbuf[:, :coef.shape[1] * coef.shape[2]] = coef.reshape(n, -1)
    dst = np.ascontiguousarray(pred, np.uint16).copy()
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_inv_txfm_add(device, cbuf.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), n, w, h,
                                      tx_type, bit_depth, reps, C.byref(ms)))
    return dst, ms.value


def _stack(planes_list):
    """list over frames of 3 padded planes -> 3 arrays [n, rows, stride]"""
    return [np.ascontiguousarray(np.stack([f[p] for f in planes_list])) for p in range(3)]


def deblock(width, height, bit_depth, blocks, rec, lf_level, sharpness=0, device=0, reps=1):
    """blocks: [n, h8*w8] structured; rec: list (frames) of 3 padded planes. Returns (list of planes, ms)."""
    n = len(rec)
    inp = _stack(rec)
    out = [np.zeros_like(a) for a in inp]
    blocks = np.ascontiguousarray(blocks)
    lv = (C.c_int32 * 4)(*lf_level)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_deblock(device, width, height, bit_depth, n, blocks.ctypes.data_as(C.c_void_p), _p3(inp), _p3(out),
                                 lv, sharpness, reps, C.byref(ms)))
    return [[out[p][i] for p in range(3)] for i in range(n)], ms.value


def cdef(width, height, bit_depth, blocks, fp, rec, src=None, forced_idx=None, device=0, reps=1):
    """Returns (list of planes, cdef_idx [n, n_sb], ms)."""
    n = len(rec)
    inp = _stack(rec)
    out = [np.zeros_like(a) for a in inp]
    blocks = np.ascontiguousarray(blocks)
    srcp = None
    if src is not None:
        s = _stack(src)
        srcp = _p3(s)
    fo = None
    if forced_idx is not None:
        forced_idx = np.ascontiguousarray(forced_idx, np.uint8)
        fo = forced_idx.ctypes.data_as(C.c_void_p)
    nsb = ((height + 63) // 64) * ((width + 63) // 64)
    idx = np.zeros((n, nsb), np.uint8)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_cdef(device, width, height, bit_depth, n, blocks.ctypes.data_as(C.c_void_p), C.byref(fp), _p3(inp),
                              srcp, fo, _p3(out), idx.ctypes.data_as(C.c_void_p), reps, C.byref(ms)))
    return [[out[p][i] for p in range(3)] for i in range(n)], idx, ms.value


def loop_restoration(width, height, bit_depth, fp, cdef_planes, deb_planes, units, device=0, reps=1):
    """units: 3 arrays [n, rows, cols] of abi.LR_UNIT_DTYPE or None. Returns (list of planes, ms)."""
    n = len(cdef_planes)
    c = _stack(cdef_planes)
    d = _stack(deb_planes)
    out = [np.zeros_like(a) for a in c]
    us = [np.ascontiguousarray(u) if u is not None else None for u in units]
    up = (C.c_void_p * 3)(*[u.ctypes.data if u is not None else None for u in us])
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_lr(device, width, height, bit_depth, n, C.byref(fp), _p3(c), _p3(d), up, _p3(out), reps, C.byref(ms)))
    return [[out[p][i] for p in range(3)] for i in range(n)], ms.value


def lr_search(width, height, bit_depth, cand, cdef_planes, deb_planes, src_luma, bias, device=0, reps=1):
    """Per-unit luma restoration decision (NONE / Wiener(cand) / self-guided(cand)) for ONE frame of padded planes.
    Returns (units [rows, cols], sse [3, rows*cols], ms)."""
    c = [np.ascontiguousarray(p, np.uint16) for p in cdef_planes]
    d = [np.ascontiguousarray(p, np.uint16) for p in deb_planes]
    sy = np.ascontiguousarray(src_luma, np.uint16)
    ur, uc = max((height + 32) // 64, 1), max((width + 32) // 64, 1)
    units = np.zeros((ur, uc), abi.LR_UNIT_DTYPE)
    sse = np.zeros((3, ur * uc), np.uint64)
    cand = np.ascontiguousarray(cand)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_lr_search(device, width, height, bit_depth, cand.ctypes.data_as(C.c_void_p), _p3(c), _p3(d),
                                   sy.ctypes.data_as(C.c_void_p), C.c_int64(int(bias)), units.ctypes.data_as(C.c_void_p),
                                   sse.ctypes.data_as(C.c_void_p), reps, C.byref(ms)))
    return units, sse, ms.value


def pyramid(width, height, l0_frames, device=0, reps=1):
    """l0_frames: [n, rows, stride] uint16 padded luma. Returns (l1, l2, ms)."""
    l0 = np.ascontiguousarray(l0_frames, np.uint16)
    n, rows, stride = l0.shape
    l1 = np.zeros((n, rows // 2, stride // 2), np.uint16)
    l2 = np.zeros((n, rows // 4, stride // 4), np.uint16)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_pyramid(device, width, height, n, l0.ctypes.data_as(C.c_void_p), l1.ctypes.data_as(C.c_void_p),
                                 l2.ctypes.data_as(C.c_void_p), reps, C.byref(ms)))
    return l1, l2, ms.value


def fwd_txfm(resid, w, h, tx_type, device=0, reps=1):
    """resid: [n, h, w] int16. Returns (coef [n, min(h,32), min(w,32)] int32, ms)."""
    r = np.ascontiguousarray(resid, np.int16)
    n = r.shape[0]
    co = np.zeros((n, min(h, 32), min(w, 32)), np.int32)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_fwd_txfm(device, r.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), n, w, h, tx_type, reps, C.byref(ms)))
    return co, ms.value


def hme(width, height, cur_l0, ref_l0, lam=0, device=0, reps=1, bd=8):
    """cur_l0 / ref_l0: [n, rows, stride] padded luma; bd: bit depth (the quarter-resolution level compares
    min(v >> (bd - 8), 255)). Returns (mv [n, h8*w8, 2], ms)."""
    cur = np.ascontiguousarray(cur_l0, np.uint16)
    ref = np.ascontiguousarray(ref_l0, np.uint16)
    n = cur.shape[0]
    mv = np.zeros((n, (height // 8) * (width // 8), 2), np.int16)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_hme(device, width, height, bd, n, cur.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p),
                             int(lam), mv.ctypes.data_as(C.c_void_p), reps, C.byref(ms)))
    return mv, ms.value


def hme_sbrd(width, height, cur_l0, ref_l0, lam, lam_s, lam_r, passes=2, device=0, reps=1, bd=8):
    """hme followed by `passes` sweeps of the superblock-level regularisation of the vector field. Returns (mv [n, h8*w8, 2], ms)."""
    cur = np.ascontiguousarray(cur_l0, np.uint16)
    ref = np.ascontiguousarray(ref_l0, np.uint16)
    n = cur.shape[0]
    mv = np.zeros((n, (height // 8) * (width // 8), 2), np.int16)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_hme_sbrd(device, width, height, bd, n, cur.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p),
                                  int(lam), int(lam_s), int(lam_r), int(passes), mv.ctypes.data_as(C.c_void_p), reps, C.byref(ms)))
    return mv, ms.value


def mctf(width, height, bit_depth, cur_padded, nb_padded, nb_mvs, thr_b, thr_p, device=0, reps=1):
    """Temporal filter of one picture: cur_padded = 3 padded planes, nb_padded = list of 3 padded planes per neighbour,
    nb_mvs = list of [h8*w8, 2] vectors. Returns (out[3], ms)."""
    cur = [np.ascontiguousarray(p, np.uint16) for p in cur_padded]
    nbs = [[np.ascontiguousarray(p, np.uint16) for p in nb] for nb in nb_padded]
    mvs = [np.ascontiguousarray(m, np.int16) for m in nb_mvs]
    n = len(nbs)
    pp = (C.c_void_p * max(1, 3 * n))(*[p.ctypes.data for nb in nbs for p in nb])
    mp = (C.c_void_p * max(1, n))(*[m.ctypes.data for m in mvs])
    out = [np.zeros_like(p) for p in cur]
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_mctf(device, width, height, bit_depth, _p3(cur), n, pp, mp, int(thr_b), int(thr_p), _p3(out), reps, C.byref(ms)))
    return out, ms.value


def noise_estimate(width, height, luma_padded, device=0):
    l0 = np.ascontiguousarray(luma_padded, np.uint16)
    out = C.c_int32(0)
    _ck(abi.lib().av1b_k_noise_estimate(device, width, height, l0.ctypes.data_as(C.c_void_p), C.byref(out)))
    return out.value


def partition_smooth(width, height, luma_padded, thr, device=0):
    """Key-frame partition by smoothness: block log2 per 8x8 unit [h8*w8]."""
    l0 = np.ascontiguousarray(luma_padded, np.uint16)
    m = np.zeros((height // 8) * (width // 8), np.uint8)
    _ck(abi.lib().av1b_k_partition_smooth(device, width, height, l0.ctypes.data_as(C.c_void_p), int(thr), m.ctypes.data_as(C.c_void_p)))
    return m


def inter_encode(width, height, bit_depth, base_q_idx, part_map, mvs, src_padded, ref_padded, tb_zero_thr=0,
                 merge_skip=False, device=0, reps=1):
    """Returns (rec[3], coef[3], blocks, ms); planes in the padded layout."""
    src = [np.ascontiguousarray(p, np.uint16) for p in src_padded]
    ref = [np.ascontiguousarray(p, np.uint16) for p in ref_padded]
    rec = [np.zeros_like(p) for p in src]
    coef = [np.zeros(p.shape, np.int16) for p in src]
    pm = np.ascontiguousarray(part_map, np.uint8)
    mv = np.ascontiguousarray(mvs, np.int16)
    blocks = np.zeros(pm.size, abi.BLOCK_INFO_DTYPE)
    ms = C.c_double(0)
    _ck(abi.lib().av1b_k_inter_encode(device, width, height, bit_depth, base_q_idx, pm.ctypes.data_as(C.c_void_p),
                                      mv.ctypes.data_as(C.c_void_p), _p3(src), _p3(ref), _p3(rec), _p3(coef),
                                      blocks.ctypes.data_as(C.c_void_p), int(tb_zero_thr), int(merge_skip), reps, C.byref(ms)))
    return rec, coef, blocks, ms.value
