"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  The CPU oracle's statement of the encoder's whole decision chain for one
closed GOP (what csrc/encoder.cc drives on the device): frame kinds of the one-level hierarchy, quantiser per kind,
key-frame partition, hierarchical motion search + vector-field regularisation against the right reference, inter /
intra encode, in-loop filters.  Tests, smoke() and the rate/quality tools compare the CUDA path with this, frame by frame.
Replaces what the reference delegates to av1an + SVT-AV1 (/root/reference/crates/daemon/src/encode/av1an.rs:126-139)."""
import ctypes as C
import os
import re
import numpy as np
from av1_base_b200 import abi
from . import pyoracle as O

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_tables = {}
DEFAULT_GOP_PERIOD = 6   # csrc/encoder.cc kDefaultGopPeriod


def table(name):
    """Integer table `name` out of csrc/av1_tables.h (quantiser steps, CRF -> quantiser index)."""
    if name not in _tables:
        txt = open(os.path.join(_ROOT, "av1_base_b200", "csrc", "av1_tables.h")).read()
        m = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, txt, re.S)
        _tables[name] = [int(v) for v in re.findall(r"-?\d+", m.group(1))]
    return _tables[name]


def ac_q(bd, qidx):
    return table("av1t_ac_q_%d" % bd)[qidx]


def frame_kind(pos, keyint=240, gop_period=DEFAULT_GOP_PERIOD, intra_only=False):
    """0 key, 1 anchor, 2 non-reference (csrc/encoder.cc frame_kind)."""
    if intra_only:
        return 0
    c = pos % keyint
    if c == 0:
        return 0
    return 1 if (gop_period <= 1 or c % gop_period == 0) else 2


def quantisers(crf, gop_period=DEFAULT_GOP_PERIOD, intra_only=False):
    """(key, anchor, non-reference) quantiser indices for a CRF (csrc/encoder.cc av1b_encoder_create)."""
    q = max(1, table("av1t_quantizer_to_qindex")[crf])
    qkey = q if intra_only else max(1, q * 3 // 4)
    qa = max(1, q - 8) if gop_period > 1 else q
    return qkey, qa, min(255, q + 64)


def class_params(bd, qidx, kind, loop_filters=True, lr=False, tile_log2=(0, 0)):
    fp = abi.FrameParams()
    abi.lib().av1b_select_frame_params(bd, qidx, 0 if kind == 0 else 1, 1 if loop_filters else 0, C.byref(fp))
    fp.non_reference = 1 if kind == 2 else 0
    if kind == 2:   # no CDEF in the frames nobody predicts from (csrc/encoder.cc set_structure)
        fp.cdef_bits = 0
        for i in range(8):
            fp.cdef_y_strength[i] = 0
            fp.cdef_uv_strength[i] = 0
    fp.tile_cols_log2, fp.tile_rows_log2 = tile_log2
    if lr:
        fp.lr_type[0], fp.lr_type[1], fp.lr_type[2] = 3, 0, 0
    return fp


class FrameResult:
    pass


def merged_partition(g, pm16, mvs, tol=0):
    """EXPERIMENT (roadmap measurement, tools/rd_chain.py --kw '{"inter_merge": true}'; not in the product): inter blocks of 32x32 /
    64x64 wherever the four (sixteen) 16x16 blocks of an aligned square that lies inside the picture carry one vector -- coded
    with one prediction and one 32x32 / 64x64 transform instead of four (sixteen) 16x16 ones."""
    pm = np.array(pm16, np.uint8).reshape(g.h8, g.w8).copy()
    mv = mvs.reshape(g.h8, g.w8, 2)
    for bl, n8 in ((5, 4), (6, 8)):
        for y0 in range(0, g.h8 - n8 + 1, n8):
            for x0 in range(0, g.w8 - n8 + 1, n8):
                blk = mv[y0:y0 + n8, x0:x0 + n8].reshape(-1, 2)
                if (pm[y0:y0 + n8, x0:x0 + n8] >= 4).all() and (np.abs(blk.astype(np.int32) - blk[0]) <= tol).all():
                    pm[y0:y0 + n8, x0:x0 + n8] = bl
                    mv[y0:y0 + n8, x0:x0 + n8] = blk[0]     # tol > 0: the block takes the vector of its first unit (mvs is updated in place)
    return pm.reshape(-1)


def choose_structure(g, bd, crf, first_luma_padded):
    """config.gop_period == 0 (csrc/encoder.cc begin_chunk): the P chain where the quantiser is fine enough to code the
    noise of the chunk's first picture, else the one-level hierarchy of the default period.  Returns (gop_period, noise estimate)."""
    q = max(1, table("av1t_quantizer_to_qindex")[crf])
    nb = O.noise_estimate(g, first_luma_padded)
    return (1 if 2 * nb > 15 * ac_q(bd, q) else DEFAULT_GOP_PERIOD), nb


def scene_positions(g, bd, padded, pos0=0, scene_cut=True):
    """Position of every frame in its closed GOP (csrc/encoder.cc launch()): it restarts at a scene change -- a score
    (O.scene_score against the picture before) above 10 per sample in 8-bit units and above three times the running level
    of change plus 2 per sample, at least 12 frames after the last key frame; level = (4 * level + score) // 5, restarted
    at a cut.  Integer arithmetic throughout."""
    unit = (((g.width - 4 + 7) // 8) * ((g.height - 4 + 7) // 8)) << (bd - 8)
    pos, level, out = pos0, -1, []
    for i, pl in enumerate(padded):
        if scene_cut and not (i == 0 and pos0 == 0) and i > 0:
            sc = O.scene_score(g, pl[0], padded[i - 1][0])
            cut = pos >= 12 and sc > 10 * unit and (level < 0 or sc > 3 * level + 2 * unit)
            level = sc if level < 0 else (4 * level + sc) // 5
            if cut:
                level, pos = -1, 0
        out.append(pos)
        pos += 1
    return out


def encode_chain(frames, w, h, bd, crf, keyint=240, gop_period=DEFAULT_GOP_PERIOD, me_smooth=True, key_var_part=True, loop_filters=True,
                 lr=False, intra_only=False, blk_log2=4, tb_zero_thr=0, pos0=0, geom=None, mctf=True, batch=8, lookahead=-1,
                 film_grain=0, mctf_radius=2, mctf_key_fwd=4, scene_cut=True, qm=None, rnd=(48, 48, 48), sbrd_passes=2, inter_merge=False, anchor_boost=None, lam_r_shift=2):
    """Returns one FrameResult per frame: kind, fp, res (blocks / coef / pre-filter rec), fin (padded planes after the
    in-loop filters), cdef_idx, lr_units, mvs.  qm = (qm_min, qm_max): quantisation matrices at the level the frame's quantiser
    index maps to (csrc/encoder.cc set_qm_levels), luma and chroma alike.  sbrd_passes: sweeps of the superblock-level regularisation
    (csrc/encoder.cc: 3 for --preset <= 3, else 2).  rnd: quantiser rounding offsets (/128) of key / anchor / non-reference frames
    (experiments; the product uses 48 throughout)."""
    g = geom if geom is not None else O.geom(w, h, 0, 0)   # key-frame tiling: no intra prediction across tile edges
    if gop_period == 0:
        gop_period, _ = choose_structure(g, bd, crf, O.pad_planes(g, frames[0])[0])
    if gop_period <= 1:
        mctf = False
    qkey, qa, qn = quantisers(crf, gop_period, intra_only)
    qk = {0: qkey, 1: qa, 2: qn}
    lam = ac_q(bd, qa) >> 1
    pm16 = O.partition_fixed(g, 4)
    out = []
    anchor_fin = anchor_pyr = None
    padded = [O.pad_planes(g, fr) for fr in frames]
    pyrs = [O.pyramid(g, pl[0]) for pl in padded]
    n = len(frames)
    gop_pos = scene_positions(g, bd, padded, pos0, scene_cut and not intra_only)
    n_anchor = 0
    try:        # the quantisation matrix levels are oracle state (O.set_qm): never leave them set behind a failure
        for i, fr in enumerate(frames):
            kind = frame_kind(gop_pos[i], keyint, gop_period, intra_only)
            q = qk[kind]
            if anchor_boost and kind == 1:      # EXPERIMENT (not in the product): every anchor_boost[0]-th anchor anchor_boost[1] quantiser steps finer
                n_anchor += 1
                if n_anchor % anchor_boost[0] == 0:
                    q = max(1, q - anchor_boost[1])
            fp = class_params(bd, q, kind, loop_filters, lr)
            fp.qm_level[0] = fp.qm_level[1] = 15      # flat (csrc/encoder.cc set_qm_levels without --enable-qm)
            if qm is not None:
                fp.using_qmatrix = 1
                fp.qm_level[0] = fp.qm_level[1] = O.qm_level(q, qm[0], qm[1])
            O.set_qm(fp.qm_level[0], fp.qm_level[1]) if qm is not None else O.set_qm()
            src = padded[i]
            pyr = pyrs[i]          # the motion search always sees the unfiltered source pictures
            nb = []
            if mctf and kind != 2 and not intra_only:
                # temporal filter of key / anchor sources (csrc/encoder.cc launch()): neighbours inside the closed GOP; ahead of
                # the picture only what the same batch holds (and --lookahead allows), behind it up to mctf_radius pictures
                lo, hi = (0, mctf_key_fwd) if kind == 0 else (-mctf_radius, mctf_radius)
                if lookahead >= 0:
                    hi = min(hi, lookahead)
                in_gop = gop_pos[i] % keyint
                for d in range(lo, hi + 1):
                    j = i + d
                    if d == 0 or j < 0 or j >= n or (d > 0 and j // batch != i // batch):
                        continue
                    if in_gop + d < 0 or in_gop + d >= keyint:
                        continue
                    if d > 0 and gop_pos[j] != gop_pos[i] + d:   # a scene change between the two
                        continue
                    nb.append(j)
                nb = nb[:6]
            if nb:
                mvs_tf = []
                for j in nb:
                    mvs_tf.append(O.hme(g, pyrs[i], pyrs[j], lam, bd))   # no regularisation for the filter's searches
                aq = ac_q(bd, q)
                thr_b = max(1, (aq * aq * (10 + film_grain)) // 2560)
                src = O.mctf(g, bd, padded[i], [padded[j] for j in nb], mvs_tf, thr_b, 3 * thr_b)
                fr = O.crop(g, src)
            r = FrameResult()
            r.kind, r.fp, r.q, r.mvs, r.src, r.filtered_from = kind, fp, q, None, src, nb
            if kind == 0:
                if key_var_part and blk_log2 == 4:
                    pm = O.partition_smooth(g, src[0], min(4 * ac_q(bd, q), 800 << (bd - 8)))
                else:
                    pm = O.partition_fixed(g, blk_log2)
                r.part_map = pm
                r.res = O.encode_intra_frame(g, fr, bd, q, pm, quant_rnd=rnd[0])
            else:
                mvs = O.hme(g, pyr, anchor_pyr, lam, bd)
                if me_smooth:
                    mvs = O.me_sbrd(g, pyr, anchor_pyr, mvs, lam, lam >> lam_r_shift, sbrd_passes)   # lam_r_shift != 2: experiment
                r.mvs = mvs
                pm_inter = merged_partition(g, pm16, mvs, int(inter_merge) - 1) if inter_merge else pm16   # inter_merge = 1 + tolerance (1/8 samples)
                r.res = O.encode_inter_frame(g, fr, bd, q, pm_inter, mvs, anchor_fin, quant_rnd=rnd[kind], tb_zero_thr=tb_zero_thr)
                O.merge_skip_blocks(g, r.res.blocks)
            fin, r.cdef_idx, r.lr_units = r.res.rec, None, None
            if loop_filters:
                O.deblock_frame(g, bd, r.res.blocks, r.res.rec, list(fp.lf_level), fp.lf_sharpness)
                if fp.cdef_bits > 0 or fp.cdef_y_strength[0] > 0 or fp.cdef_uv_strength[0] > 0:
                    r.cdef_idx = O.cdef_search(g, bd, r.res.blocks, fp, r.res.rec, src)
                    fin = O.cdef_frame(g, bd, r.res.blocks, fp, r.cdef_idx, r.res.rec)
                else:
                    r.cdef_idx = np.zeros(g.sb_rows * g.sb_cols, np.uint8)
                    fin = r.res.rec
                if lr:
                    cand = O.lr_candidate((0, 0, 8), (0, 0, 8), 12, (0, 95))
                    aq = ac_q(bd, q)
                    r.lr_units, _ = O.lr_search(g, bd, fp, cand, fin, r.res.rec, src[0], (aq * aq * 5) >> 8)
                    fin = O.lr_frame(g, bd, fp, fin, r.res.rec, [r.lr_units, None, None])
            r.fin = fin
            O.set_qm()
            if kind != 2:
                anchor_fin, anchor_pyr = fin, pyr
            out.append(r)
    finally:
        O.set_qm()
    return g, out
