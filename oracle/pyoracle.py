"""Test infrastructure (NOT product code): ctypes wrapper of liboracle.so (oracle/av1_oracle.cpp)."""
import ctypes as C
import os, subprocess
import numpy as np
from av1_base_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])

def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            build()
        _lib = C.CDLL(p)
    return _lib

def geom(width, height, tile_cols_log2=0, tile_rows_log2=0):
    g = abi.Geom()
    rc = lib().orc_geom_init(C.byref(g), width, height, tile_cols_log2, tile_rows_log2)
    if rc:
        raise ValueError("unsupported size %dx%d" % (width, height))
    return g

def ptr(a):
    return a.ctypes.data_as(C.c_void_p)

def partition_fixed(g, blk_log2):
    m = np.zeros(g.h8 * g.w8, np.uint8)
    lib().orc_partition_fixed(C.byref(g), blk_log2, ptr(m))
    return m

def partition_smooth(g, luma_padded, thr):
    """Key-frame partition by smoothness (orc_partition_smooth): block log2 per 8x8 unit."""
    m = np.zeros(g.h8 * g.w8, np.uint8)
    l0 = np.ascontiguousarray(luma_padded, np.uint16)
    lib().orc_partition_smooth(C.byref(g), ptr(l0), l0.shape[1], int(thr), ptr(m))
    return m


def set_qm(level_y=15, level_uv=15):
    """Quantisation matrix levels of the frames coded from here on (orc_set_qm): 0 steepest .. 14, 15 = flat (off)."""
    lib().orc_set_qm(int(level_y), int(level_uv))

def qm_level(qidx, first, last):
    """aom_get_qmlevel: the level a quantiser index maps to between --qm-min and --qm-max."""
    return int(lib().orc_qm_level(int(qidx), int(first), int(last)))


class IntraResult:
    pass

def encode_intra_frame(g, frame, bit_depth, base_q_idx, part_map, quant_rnd=48):
    """frame: [Y,U,V] uint16. Returns IntraResult with rec[3] (padded), blocks, coef[3] (padded)."""
    Y, U, V = [np.ascontiguousarray(p, dtype=np.uint16) for p in frame]
    r = IntraResult()
    r.rec = [np.zeros((g.rows[p], g.stride[p]), np.uint16) for p in range(3)]
    r.coef = [np.zeros((g.rows[p], g.stride[p]), np.int16) for p in range(3)]
    r.blocks = np.zeros(g.h8 * g.w8, abi.BLOCK_INFO_DTYPE)
    rc = lib().orc_encode_intra_frame(C.byref(g), bit_depth, base_q_idx, quant_rnd, ptr(Y), ptr(U), ptr(V),
                                      Y.shape[1], U.shape[1], ptr(part_map),
                                      ptr(r.rec[0]), ptr(r.rec[1]), ptr(r.rec[2]), ptr(r.blocks),
                                      ptr(r.coef[0]), ptr(r.coef[1]), ptr(r.coef[2]))
    assert rc == 0
    return r

def crop(g, planes):
    return [planes[0][:g.height, :g.width], planes[1][:g.height // 2, :g.width // 2],
            planes[2][:g.height // 2, :g.width // 2]]

def deblock_frame(g, bit_depth, blocks, rec, lf_level, sharpness=0):
    """In-place deblocking of padded planes rec[3]."""
    lv = (C.c_int32 * 4)(*lf_level)
    lib().orc_deblock_frame(C.byref(g), bit_depth, ptr(blocks), ptr(rec[0]), ptr(rec[1]), ptr(rec[2]), lv, sharpness)

def cdef_frame(g, bit_depth, blocks, fp, cdef_idx, rec):
    """Returns new padded planes = CDEF(rec)."""
    out = [np.zeros_like(p) for p in rec]
    lib().orc_cdef_frame(C.byref(g), bit_depth, ptr(blocks), C.byref(fp), ptr(cdef_idx), ptr(rec[0]), ptr(rec[1]),
                         ptr(rec[2]), ptr(out[0]), ptr(out[1]), ptr(out[2]))
    return out

def lr_unit_grid(g, fp, plane):
    us, ur, uc = C.c_int32(), C.c_int32(), C.c_int32()
    lib().orc_lr_unit_grid(C.byref(g), C.byref(fp), plane, C.byref(us), C.byref(ur), C.byref(uc))
    return us.value, ur.value, uc.value

def lr_frame(g, bit_depth, fp, cdef, deblocked, units):
    """units: 3 structured arrays (abi.LR_UNIT_DTYPE, shape [rows, cols]) or None. Returns new planes."""
    out = [np.zeros_like(p) for p in cdef]
    up = [ptr(u) if u is not None else None for u in units]
    lib().orc_lr_frame(C.byref(g), bit_depth, C.byref(fp), ptr(cdef[0]), ptr(cdef[1]), ptr(cdef[2]),
                       ptr(deblocked[0]), ptr(deblocked[1]), ptr(deblocked[2]), ptr(out[0]), ptr(out[1]), ptr(out[2]),
                       up[0], up[1], up[2])
    return out


def cdef_search(g, bit_depth, blocks, fp, rec, src):
    """Encoder-side CDEF preset decision per superblock (rec = deblocked padded planes, src = padded source)."""
    idx = np.zeros(g.sb_rows * g.sb_cols, np.uint8)
    lib().orc_cdef_search(C.byref(g), bit_depth, ptr(blocks), C.byref(fp), ptr(rec[0]), ptr(rec[1]), ptr(rec[2]),
                          ptr(src[0]), ptr(src[1]), ptr(src[2]), ptr(idx))
    return idx


def pad_planes(g, frame):
    """[Y,U,V] picture-sized arrays -> padded planes with the geometry's strides (zeros outside)."""
    out = [np.zeros((g.rows[p], g.stride[p]), np.uint16) for p in range(3)]
    for p in range(3):
        h, w = frame[p].shape
        out[p][:h, :w] = frame[p]
    return out


def inter_predict(ref, x, y, w, h, mv, ss, bit_depth, ref_w=None, ref_h=None):
    """Normative single-reference prediction of one block from plane `ref` (2-D uint16 array)."""
    ref = np.ascontiguousarray(ref, np.uint16)
    out = np.zeros((h, w), np.uint16)
    lib().orc_inter_predict(ptr(ref), ref.shape[1], ref_w or ref.shape[1], ref_h or ref.shape[0], x, y, w, h,
                            int(mv[0]), int(mv[1]), ss, bit_depth, ptr(out), w)
    return out


def pyramid(g, luma_padded):
    """[L0, L1, L2] with strides stride0, stride0/2, stride0/4 (padded layout, picture area filled)."""
    l0 = np.ascontiguousarray(luma_padded, np.uint16)
    l1 = np.zeros((g.rows[0] // 2, g.stride[0] // 2), np.uint16)
    l2 = np.zeros((g.rows[0] // 4, g.stride[0] // 4), np.uint16)
    lib().orc_downscale2(ptr(l0), g.stride[0], g.width, g.height, ptr(l1), g.stride[0] // 2)
    lib().orc_downscale2(ptr(l1), g.stride[0] // 2, g.width // 2, g.height // 2, ptr(l2), g.stride[0] // 4)
    return [l0, l1, l2]


def hme(g, cur_pyr, ref_pyr, lam=0, bd=8):
    """Motion vectors [h8*w8, 2] (row, col) in 1/8 luma samples. lam: vector-deviation cost (SAD units); bd: bit depth of
    the samples (the quarter-resolution search compares min(v >> (bd - 8), 255))."""
    mv = np.zeros((g.h8 * g.w8, 2), np.int16)
    lib().orc_hme(C.byref(g), ptr(cur_pyr[0]), ptr(cur_pyr[1]), ptr(cur_pyr[2]), ptr(ref_pyr[0]), ptr(ref_pyr[1]),
                  ptr(ref_pyr[2]), int(lam), int(bd) - 8, ptr(mv))
    return mv


def me_sbrd(g, cur_pyr, ref_pyr, mv, lam_s, lam_r, passes=2):
    """Superblock-level rate-distortion regularisation of the vector field (orc_me_sbrd): returns the updated [h8*w8, 2] vectors."""
    mv = np.ascontiguousarray(mv, np.int16).copy()
    lib().orc_me_sbrd(C.byref(g), ptr(cur_pyr[0]), ptr(ref_pyr[0]), int(lam_s), int(lam_r), int(passes), ptr(mv))
    return mv


def scene_score(g, cur_luma_padded, prev_luma_padded):
    """Sum of absolute luma differences on the 1/8 x 1/8 grid from (4, 4) (orc_scene_score)."""
    a = np.ascontiguousarray(cur_luma_padded, np.uint16); b = np.ascontiguousarray(prev_luma_padded, np.uint16)
    lib().orc_scene_score.restype = C.c_uint32
    return int(lib().orc_scene_score(C.byref(g), ptr(a), ptr(b), a.shape[1]))


def noise_estimate(g, luma_padded):
    """Lower-quartile block sum of |I * N| (orc_noise_estimate); sigma ~= 0.0010658 * result."""
    l0 = np.ascontiguousarray(luma_padded, np.uint16)
    return int(lib().orc_noise_estimate(C.byref(g), ptr(l0), l0.shape[1]))


def mctf(g, bit_depth, cur_planes, nb_planes, nb_mvs, thr_b, thr_p):
    """Motion-compensated temporal filter (orc_mctf): cur_planes = 3 padded planes, nb_planes = list of 3 padded planes per
    neighbour, nb_mvs = list of [h8*w8, 2] vectors (cur against that neighbour). Returns 3 padded planes."""
    cur = [np.ascontiguousarray(p, np.uint16) for p in cur_planes]
    nbs = [[np.ascontiguousarray(p, np.uint16) for p in nb] for nb in nb_planes]
    mvs = [np.ascontiguousarray(m, np.int16) for m in nb_mvs]
    n = len(nbs)
    pp = (C.c_void_p * max(1, 3 * n))(*[p.ctypes.data for nb in nbs for p in nb])
    mp = (C.c_void_p * max(1, n))(*[m.ctypes.data for m in mvs])
    out = [np.zeros_like(p) for p in cur]
    lib().orc_mctf(C.byref(g), bit_depth, ptr(cur[0]), ptr(cur[1]), ptr(cur[2]), n, pp, mp, int(thr_b), int(thr_p),
                   ptr(out[0]), ptr(out[1]), ptr(out[2]))
    return out


def merge_skip_blocks(g, blocks):
    """In place: merges skipped inter siblings with equal vectors into 32x32 / 64x64 blocks."""
    lib().orc_merge_skip_blocks(C.byref(g), ptr(blocks))
    return blocks


def encode_inter_frame(g, frame, bit_depth, base_q_idx, part_map, mvs, ref_planes, quant_rnd=48, tb_zero_thr=0):
    """frame: [Y,U,V]; ref_planes: 3 padded planes (previous reconstructed frame); mvs: [h8*w8, 2] int16."""
    Y, U, V = [np.ascontiguousarray(p, dtype=np.uint16) for p in frame]
    mvs = np.ascontiguousarray(mvs, np.int16)
    r = IntraResult()
    r.rec = [np.zeros((g.rows[p], g.stride[p]), np.uint16) for p in range(3)]
    r.coef = [np.zeros((g.rows[p], g.stride[p]), np.int16) for p in range(3)]
    r.blocks = np.zeros(g.h8 * g.w8, abi.BLOCK_INFO_DTYPE)
    rc = lib().orc_encode_inter_frame(C.byref(g), bit_depth, base_q_idx, quant_rnd, tb_zero_thr, ptr(Y), ptr(U), ptr(V),
                                      Y.shape[1], U.shape[1], ptr(part_map), ptr(mvs),
                                      ptr(ref_planes[0]), ptr(ref_planes[1]), ptr(ref_planes[2]),
                                      ptr(r.rec[0]), ptr(r.rec[1]), ptr(r.rec[2]), ptr(r.blocks),
                                      ptr(r.coef[0]), ptr(r.coef[1]), ptr(r.coef[2]))
    assert rc == 0
    return r


def lr_candidate(wiener_v=(3, -7, 15), wiener_h=(3, -7, 15), sgr_set=4, sgr_xqd=(-32, 31)):
    c = np.zeros(1, abi.LR_UNIT_DTYPE)
    c["wiener_v"], c["wiener_h"], c["sgr_set"], c["sgr_xqd"] = wiener_v, wiener_h, sgr_set, sgr_xqd
    return c


def lr_search(g, bit_depth, fp, cand, cdef, deblocked, src_luma_padded, bias):
    """Per-unit choice among NONE / WIENER(cand) / SGRPROJ(cand) for the luma plane. Returns (units [rows, cols], sse [3, n])."""
    us, ur, uc = lr_unit_grid(g, fp, 0)
    units = np.zeros((ur, uc), abi.LR_UNIT_DTYPE)
    sse = np.zeros((3, ur * uc), np.uint64)
    lib().orc_lr_search(C.byref(g), bit_depth, C.byref(fp), ptr(cand), ptr(cdef[0]), ptr(cdef[1]), ptr(cdef[2]),
                        ptr(deblocked[0]), ptr(deblocked[1]), ptr(deblocked[2]), ptr(src_luma_padded), C.c_int64(int(bias)),
                        ptr(units), ptr(sse))
    return units, sse
