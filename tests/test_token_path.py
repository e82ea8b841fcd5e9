"""Token path of the host entropy coder (csrc/tokens.h): one token per coded symbol, derived per block without any
entropy-coder state (what the device tokenizer does), then range-coded.  Pinned here against the block-walking
tile writer (av1b_pack_frame), whose streams dav1d and libaom decode bit-exactly (tests/test_oracle_inter.py):
identical bytes on inter frames with random / searched vectors, merged skip blocks, picture-edge 8x8 blocks,
several tile layouts, CDEF indices, and levels large enough for Golomb escapes."""
import ctypes as C
import numpy as np
import pytest
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O
from oracle import decoders as D


def random_mvs(g, pm, rng, amp=40):
    mv = np.zeros((g.h8, g.w8, 2), np.int16)
    for by in range(0, g.h8, 2):
        for bx in range(0, g.w8, 2):
            # a few distinct vectors so that NEAREST / NEAR / GLOBAL / NEW all occur
            k = rng.integers(0, 6)
            v = [(0, 0), (8, -16), (8, -16), (-24, 4), tuple(2 * rng.integers(-amp, amp, 2)), (2, 2)][k]
            mv[by:by + 2, bx:bx + 2] = v
    return mv.reshape(-1, 2)


def inter_frame(w, h, bd, q, tcl, trl, seed, mode):
    rng = np.random.default_rng(seed)
    g = O.geom(w, h, tcl, trl)
    frames = synth.synth_clip(w, h, bd, 2, seed=7 + seed, scene_len=100)
    pm = O.partition_fixed(g, 4)
    r0 = O.encode_intra_frame(g, frames[0], bd, q, pm)
    if mode == "rand":
        mvs = random_mvs(g, pm, rng)
    else:
        mvs = O.hme(g, O.pyramid(g, O.pad_planes(g, frames[1])[0]), O.pyramid(g, O.pad_planes(g, frames[0])[0]), 40, bd)
    r1 = O.encode_inter_frame(g, frames[1], bd, q, pm, mvs, r0.rec)
    O.merge_skip_blocks(g, r1.blocks)
    cdef_idx = rng.integers(0, 8, g.sb_rows * g.sb_cols).astype(np.uint8)
    return g, r0, r1, cdef_idx, frames


CASES = [
    (64, 64, 8, 120, 0, 0, "rand"),
    (128, 128, 10, 100, 0, 0, "rand"),
    (200, 136, 10, 60, 0, 0, "rand"),      # 8x8 blocks at the right and bottom edges
    (328, 248, 10, 120, 1, 1, "rand"),     # 2x2 tiles
    (328, 248, 8, 180, 2, 1, "hme"),
    (328, 248, 8, 12, 1, 0, "rand"),       # low quantiser: levels >= 15 (Golomb escapes)
    (640, 360, 10, 150, 1, 1, "hme"),      # mostly skipped, merged 32x32 / 64x64 blocks
    (640, 360, 10, 40, 0, 0, "hme"),
]


@pytest.mark.parametrize("w,h,bd,q,tcl,trl,mode", CASES)
def test_token_path_gives_identical_bytes(w, h, bd, q, tcl, trl, mode):
    g, r0, r1, cdef_idx, _ = inter_frame(w, h, bd, q, tcl, trl, 3, mode)
    seq = abi.SeqParams(w, h, bd, 1, 0, 30, 1, 0)
    fp = abi.FrameParams()
    abi.lib().av1b_select_frame_params(bd, q, 1, 1, C.byref(fp))
    fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
    sy = packer.make_syms(g, r1.blocks, r1.coef, cdef_idx=cdef_idx)
    ref = packer.pack_frame(seq, fp, sy, with_td=False)
    got, n_tok = packer.pack_frame_tokens(seq, fp, sy, with_td=False)
    assert n_tok > 0
    assert got == ref
    if q == 12:
        assert max(int(np.abs(c).max()) for c in r1.coef) >= 15   # the escape path was exercised


def test_token_path_stream_decodes():
    """End to end: key frame by the tile writer, inter frame by the token path, decoded by dav1d == reconstruction."""
    w, h, bd, q = 328, 248, 10, 100
    g, r0, r1, _, frames = inter_frame(w, h, bd, q, 1, 1, 5, "hme")
    seq = abi.SeqParams(w, h, bd, 0, 0, 30, 1, 0)
    tus = []
    for ft, r in enumerate((r0, r1)):
        fp = abi.FrameParams()
        fp.frame_type = ft
        fp.base_q_idx = q
        fp.cdef_damping = 3
        fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        sy = packer.make_syms(g, r.blocks, r.coef)
        body = packer.pack_frame(seq, fp, sy, with_td=False) if ft == 0 else packer.pack_frame_tokens(seq, fp, sy, with_td=False)[0]
        tus.append(b"\x12\x00" + (packer.pack_sequence_header(seq) if ft == 0 else b"") + body)
    dec = D.dav1d_decode(tus)
    for i, r in enumerate((r0, r1)):
        for p in range(3):
            assert np.array_equal(dec[i][p], O.crop(g, r.rec)[p])


@pytest.mark.parametrize("w,h,tcl,trl", [(328, 248, 1, 1), (200, 136, 0, 0), (640, 360, 1, 0)])
def test_token_path_with_restoration_units(w, h, tcl, trl):
    """Luma restoration units (frame type SWITCHABLE, 64x64 units) go through the token path as one symbol plus
    coefficient tokens the host codes against its running references: same bytes as the tile writer."""
    bd, q = 10, 110
    g, r0, r1, cdef_idx, _ = inter_frame(w, h, bd, q, tcl, trl, 11, "hme")
    rng = np.random.default_rng(w)
    seq = abi.SeqParams(w, h, bd, 1, 1, 30, 1, 0)
    fp = abi.FrameParams()
    abi.lib().av1b_select_frame_params(bd, q, 1, 1, C.byref(fp))
    fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
    fp.lr_type[0], fp.lr_type[1], fp.lr_type[2] = 3, 0, 0
    us, ur, uc = O.lr_unit_grid(g, fp, 0)
    assert us == 64
    units = np.zeros((ur, uc), abi.LR_UNIT_DTYPE)
    units["type"] = rng.integers(0, 3, (ur, uc))
    units["sgr_set"] = rng.integers(0, 16, (ur, uc))
    for j, (lo, hi) in enumerate([(-5, 10), (-23, 8), (-17, 46)]):
        units["wiener_v"][..., j] = rng.integers(lo, hi + 1, (ur, uc))
        units["wiener_h"][..., j] = rng.integers(lo, hi + 1, (ur, uc))
    units["sgr_xqd"][..., 0] = rng.integers(-96, 32, (ur, uc))
    units["sgr_xqd"][..., 1] = rng.integers(-32, 96, (ur, uc))
    sy = packer.make_syms(g, r1.blocks, r1.coef, cdef_idx=cdef_idx, lr_units=[units, None, None])
    ref = packer.pack_frame(seq, fp, sy, with_td=False)
    got, _ = packer.pack_frame_tokens(seq, fp, sy, with_td=False)
    assert got == ref
