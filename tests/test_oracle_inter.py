"""Inter frames end to end on the CPU: oracle motion-compensated encode (normative 8-tap prediction,
residual coding) + host entropy coder (inter frame header, reference/mode/motion-vector syntax with the
spec's motion-vector prediction stack) must decode in dav1d 1.5.3 AND libaom 3.13.1 to exactly the
oracle's reconstruction, for random vectors (NEWMV / NEARESTMV / NEARMV / GLOBALMV + DRL paths,
vectors pointing outside the picture), every block size, tiles, 8/10 bit, and for the oracle's
hierarchical motion search."""
import numpy as np
import pytest
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O, decoders as D


def random_mvs(g, pm, rng):
    m = (rng.integers(-6, 7, (g.h8, g.w8, 2)) * 2).astype(np.int16)
    pal = (rng.integers(-40, 41, (5, 2)) * 2).astype(np.int16)
    sel = rng.integers(0, 7, (g.h8, g.w8))
    for k in range(5):
        m[sel == k] = pal[k]
    mv = np.zeros((g.h8, g.w8, 2), np.int16)
    pmm = pm.reshape(g.h8, g.w8)
    for uy in range(g.h8):
        for ux in range(g.w8):
            n8 = 1 << (int(pmm[uy, ux]) - 3)
            mv[uy, ux] = m[uy & ~(n8 - 1), ux & ~(n8 - 1)]
    return mv.reshape(-1, 2)


def encode_clip(w, h, bd, q, nfr, mode, blk, tcl, trl, lf, seed):
    rng = np.random.default_rng(seed)
    g = O.geom(w, h, tcl, trl)
    frames = synth.synth_clip(w, h, bd, nfr, seed=5 + seed, scene_len=100)
    seq = abi.SeqParams(w, h, bd, 0, 0, 30, 1, 0)
    pm = O.partition_fixed(g, blk)
    tus, recs, prev, prev_src = [], [], None, None
    for fi, fr in enumerate(frames):
        fp = abi.FrameParams()
        fp.base_q_idx = q
        fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        fp.cdef_damping = 3
        if lf:
            for i in range(4):
                fp.lf_level[i] = 10 + i
        if fi == 0:
            fp.frame_type = 0
            r = O.encode_intra_frame(g, fr, bd, q, pm)
        else:
            fp.frame_type = 1
            if mode == "zero":
                mvs = np.zeros((g.h8 * g.w8, 2), np.int16)
            elif mode == "rand":
                mvs = random_mvs(g, pm, rng)
            else:
                mvs = O.hme(g, O.pyramid(g, O.pad_planes(g, fr)[0]), O.pyramid(g, O.pad_planes(g, prev_src)[0]), 0, bd)
            r = O.encode_inter_frame(g, fr, bd, q, pm, mvs, prev)
        if lf:
            O.deblock_frame(g, bd, r.blocks, r.rec, list(fp.lf_level), 0)
        sy = packer.make_syms(g, r.blocks, r.coef)
        tus.append(b"\x12\x00" + (packer.pack_sequence_header(seq) if fi == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False))
        recs.append(O.crop(g, r.rec))
        prev, prev_src = r.rec, fr
    return tus, recs


CASES = [
    (64, 64, 8, 120, 2, "zero", 4, 0, 0, False),
    (64, 64, 8, 120, 3, "rand", 4, 0, 0, False),
    (128, 128, 10, 100, 3, "rand", 4, 0, 0, False),
    (200, 136, 10, 60, 3, "rand", 3, 0, 0, False),
    (328, 248, 10, 120, 3, "rand", 5, 1, 1, True),
    (328, 248, 8, 180, 3, "rand", 6, 2, 1, False),
    (328, 248, 8, 120, 3, "hme", 4, 1, 0, True),
]


@pytest.mark.parametrize("w,h,bd,q,nfr,mode,blk,tcl,trl,lf", CASES)
def test_inter_frames_decode_bit_exact(w, h, bd, q, nfr, mode, blk, tcl, trl, lf):
    tus, recs = encode_clip(w, h, bd, q, nfr, mode, blk, tcl, trl, lf, seed=w + q)
    for name, dec in (("dav1d", D.dav1d_decode), ("libaom", D.aom_decode)):
        out = dec(tus)
        assert len(out) == nfr, name
        for fi in range(nfr):
            for p in range(3):
                assert np.array_equal(out[fi][p], recs[fi][p]), (name, fi, p)


def test_inter_predict_integer_and_border():
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 1024, (40, 48)).astype(np.uint16)
    # integer vector: plain copy
    p = O.inter_predict(ref, 8, 8, 16, 16, (16, -24), 0, 10)
    assert np.array_equal(p, ref[10:26, 5:21])
    # far outside: edge replication
    p = O.inter_predict(ref, 0, 0, 8, 8, (-8 * 100, -8 * 100), 0, 10)
    assert np.all(p == ref[0, 0])


def test_hme_finds_global_translation():
    w, h = 256, 192
    g = O.geom(w, h)
    rng = np.random.default_rng(1)
    base = rng.integers(64, 940, (h + 64, w + 64)).astype(np.float32)
    from scipy.ndimage import uniform_filter
    base = uniform_filter(base, 5)
    ref = base[32:32 + h, 32:32 + w].astype(np.uint16)
    cur = base[32 + 5:32 + 5 + h, 32 - 9:32 - 9 + w].astype(np.uint16)     # cur(x,y) = ref(x-9, y+5)
    mv = O.hme(g, O.pyramid(g, O.pad_planes(g, [cur, cur[::2, ::2], cur[::2, ::2]])[0]),
               O.pyramid(g, O.pad_planes(g, [ref, ref[::2, ::2], ref[::2, ::2]])[0]), 0, 10).reshape(g.h8, g.w8, 2)
    inner = mv[4:-4, 4:-4]
    assert np.all(inner[..., 0] == 5 * 8) and np.all(inner[..., 1] == -9 * 8)


def _tables():
    import re, os
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "av1_base_b200", "csrc", "av1_tables.h")).read()

    def tab(name):
        m = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, txt, re.S)
        return np.array([int(v) for v in re.findall(r"-?\d+", m.group(1))])
    return {n: (tab("av1t_scan_default_%dx%d" % (n, n)), tab("av1t_nz_map_ctx_offset_%dx%d" % (n, n))) for n in (4, 8, 16)}


def digest_like_device(g, blocks, coef):
    """Python restatement of what inter_kernel.cu writes with pack_levels = 1: transform blocks whose levels are
    all < 15 become scan-ordered packed symbols (sign | level | br ctx | base ctx), flagged in bit 15 of eob."""
    T = _tables()
    blocks = blocks.copy()
    coef = [c.copy() for c in coef]
    bl = blocks["blk_log2"].reshape(g.h8, g.w8)
    for uy in range(g.h8):
        for ux in range(g.w8):
            b = int(bl[uy, ux]); n8 = 1 << (b - 3)
            if (ux | uy) & (n8 - 1) or b > 4:
                continue
            for p in range(3):
                ss = 1 if p else 0
                n = 1 << (b - ss)
                eob = int(blocks["eob"][uy * g.w8 + ux][p])
                if eob == 0:
                    continue
                x, y = (ux * 8) >> ss, (uy * 8) >> ss
                lsb, lu = 6 - ss, 3 - ss
                uxx, uyy = (x >> lu) & 7, (y >> lu) & 7
                m = (uxx & 1) | ((uyy & 1) << 1) | ((uxx & 2) << 1) | ((uyy & 2) << 2) | ((uxx & 4) << 2) | ((uyy & 4) << 3)
                off = ((((y >> lsb) * g.sb_cols + (x >> lsb)) << (2 * lsb)) + (m << (2 * lu)))
                flat = coef[p].reshape(-1)
                lv = flat[off:off + n * n].astype(np.int32).reshape(n, n)
                if np.abs(lv).max() >= 15:
                    continue
                scan, nzo = T[n]
                a = np.zeros((n + 2, n + 2), np.int32); a[:n, :n] = np.abs(lv)
                words = flat[off:off + n * n].copy().view(np.uint16)
                for i in range(eob):
                    pos = int(scan[i]); r, c = pos // n, pos % n
                    m3 = sum(min(int(v), 3) for v in (a[r, c + 1], a[r + 1, c], a[r + 1, c + 1], a[r, c + 2], a[r + 2, c]))
                    bctx = 0 if pos == 0 else min((m3 + 1) >> 1, 4) + int(nzo[pos])
                    m15 = int(a[r, c + 1] + a[r + 1, c] + a[r + 1, c + 1])
                    brctx = min((m15 + 1) >> 1, 6) + (0 if pos == 0 else (7 if (r < 2 and c < 2) else 14))
                    words[i] = ((1 if lv[r, c] < 0 else 0) << 15) | (int(a[r, c]) << 11) | (brctx << 6) | bctx
                flat[off:off + n * n] = words.view(np.int16)
                for yy in range(n8):
                    for xx in range(n8):
                        blocks["eob"][(uy + yy) * g.w8 + ux + xx][p] = eob | 0x8000
    return blocks, coef


@pytest.mark.parametrize("w,h,bd,q", [(200, 136, 10, 60), (328, 248, 8, 150)])
def test_device_digested_coefficients_pack_to_identical_bytes(w, h, bd, q):
    """The host entropy coder must emit exactly the same bytes from the device-digested symbol form as from
    raster levels (the GPU side of the same statement is tests/test_gpu_inter_parity.py)."""
    rng = np.random.default_rng(q)
    g = O.geom(w, h, 1, 0)
    frames = synth.synth_clip(w, h, bd, 2, seed=9, scene_len=100)
    pm = O.partition_fixed(g, 4)
    r0 = O.encode_intra_frame(g, frames[0], bd, q, pm)
    r1 = O.encode_inter_frame(g, frames[1], bd, q, pm, random_mvs(g, pm, rng), r0.rec)
    seq = abi.SeqParams(w, h, bd, 0, 0, 30, 1, 0)
    fp = abi.FrameParams()
    fp.frame_type, fp.base_q_idx = 1, q
    fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
    fp.cdef_damping = 3
    raster = packer.pack_frame(seq, fp, packer.make_syms(g, r1.blocks, r1.coef))
    b2, c2 = digest_like_device(g, r1.blocks, r1.coef)
    assert (b2["eob"] & 0x8000).any()
    digested = packer.pack_frame(seq, fp, packer.make_syms(g, b2, c2))
    assert raster == digested


def test_non_reference_frames_decode_bit_exact():
    """One-level hierarchy (Av1bFrameParams.non_reference): every second inter frame updates no reference slot and the
    frames after it keep predicting from the last frame that did.  dav1d and libaom must follow the same references."""
    w, h, bd, q = 128, 96, 10, 100
    g = O.geom(w, h, 0, 0)
    frames = synth.synth_clip(w, h, bd, 6, seed=21, scene_len=100)
    seq = abi.SeqParams(w, h, bd, 0, 0, 30, 1, 0)
    pm = O.partition_fixed(g, 4)
    tus, recs, anchor, anchor_src = [], [], None, None
    for fi, fr in enumerate(frames):
        fp = abi.FrameParams()
        fp.base_q_idx = q + (0 if fi % 2 == 0 else 40)
        fp.cdef_damping = 3
        fp.frame_type = 0 if fi == 0 else 1
        fp.non_reference = 1 if (fi > 0 and fi % 2 == 1) else 0
        if fi == 0:
            r = O.encode_intra_frame(g, fr, bd, fp.base_q_idx, pm)
        else:
            mvs = O.hme(g, O.pyramid(g, O.pad_planes(g, fr)[0]), O.pyramid(g, O.pad_planes(g, anchor_src)[0]), 20, bd)
            r = O.encode_inter_frame(g, fr, bd, fp.base_q_idx, pm, mvs, anchor)
        sy = packer.make_syms(g, r.blocks, r.coef)
        tus.append(b"\x12\x00" + (packer.pack_sequence_header(seq) if fi == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False))
        recs.append(O.crop(g, r.rec))
        if not fp.non_reference:
            anchor, anchor_src = r.rec, fr
    for dec in (D.dav1d_decode(tus), D.aom_decode(tus)):
        assert len(dec) == len(frames)
        for i in range(len(frames)):
            for p in range(3):
                assert np.array_equal(dec[i][p], recs[i][p]), (i, p)
