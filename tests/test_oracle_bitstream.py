"""End-to-end pinning of the CPU oracle + host entropy coder: complete key-frame bitstreams (intra
prediction, all transform sizes, deblock, CDEF, loop restoration, tiles, 8/10 bit) must decode in
dav1d 1.5.3 AND libaom 3.13.1 to exactly the oracle's reconstruction."""
import numpy as np
import pytest
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O, decoders as D

SGR_R = [(2, 1)] * 10 + [(0, 1)] * 4 + [(2, 0)] * 2


def random_lr_units(g, fp, rng):
    tmin, tmax = [-5, -23, -17], [10, 8, 46]
    units = []
    for p in range(3):
        lt = fp.lr_type[p]
        if lt == 0:
            units.append(None)
            continue
        us, ur, uc = O.lr_unit_grid(g, fp, p)
        u = np.zeros((ur, uc), abi.LR_UNIT_DTYPE)
        for a in range(ur):
            for b in range(uc):
                t = int(rng.integers(0, 3)) if lt == 3 else int(rng.integers(0, 2)) * (1 if lt == 1 else 2)
                u[a, b]["type"] = t
                if t == 1:
                    for nm in ("wiener_v", "wiener_h"):
                        for j in range(3):
                            u[a, b][nm][j] = 0 if (p > 0 and j == 0) else int(rng.integers(tmin[j], tmax[j] + 1))
                elif t == 2:
                    st = int(rng.integers(0, 16))
                    r0, r1 = SGR_R[st]
                    x0 = int(rng.integers(-96, 32)) if r0 else 0
                    x1 = int(rng.integers(-32, 96)) if r1 else min(max(128 - x0, -32), 95)
                    u[a, b]["sgr_set"] = st
                    u[a, b]["sgr_xqd"][0] = x0
                    u[a, b]["sgr_xqd"][1] = x1
        units.append(u)
    return units


def random_partition(g, rng, max_log2):
    pm = O.partition_fixed(g, max_log2).reshape(g.h8, g.w8).copy()
    for y in range(0, g.h8, 8):
        for x in range(0, g.w8, 8):
            sub = pm[y:y + 8, x:x + 8]
            sub[...] = np.minimum(sub, int(rng.integers(3, 7)))
            for yy in range(0, 8, 4):
                for xx in range(0, 8, 4):
                    s2 = sub[yy:yy + 4, xx:xx + 4]
                    s2[...] = np.minimum(s2, int(rng.integers(3, 7)))
    return pm.ravel()


CASES = [
    # w, h, bd, q, tcl, trl, lf, cdef_bits, lr_types, unit_shift, uv_shift, adapt
    (64, 64, 8, 120, 0, 0, (0, 0, 0, 0), None, (0, 0, 0), 0, 0, 1),
    (128, 128, 8, 60, 0, 0, (10, 10, 10, 10), 0, (0, 0, 0), 0, 0, 1),
    (200, 136, 10, 160, 0, 0, (20, 14, 9, 30), 2, (1, 1, 1), 0, 1, 1),
    (328, 248, 10, 100, 1, 1, (12, 9, 7, 11), 3, (2, 2, 2), 0, 0, 0),
    (328, 248, 8, 200, 2, 1, (63, 5, 12, 7), 1, (3, 3, 3), 1, 1, 1),
    (640, 360, 10, 140, 2, 2, (17, 33, 8, 21), 3, (3, 1, 2), 2, 1, 1),
]


@pytest.mark.parametrize("w,h,bd,q,tcl,trl,lf,cdef_bits,lrt,ushift,uvshift,adapt", CASES)
def test_full_key_frame_decodes_bit_exact(w, h, bd, q, tcl, trl, lf, cdef_bits, lrt, ushift, uvshift, adapt):
    rng = np.random.default_rng(w + h + q)
    g = O.geom(w, h, tcl, trl)
    frames = synth.synth_clip(w, h, bd, 2, seed=5, scene_len=1)
    seq = abi.SeqParams(w, h, bd, int(cdef_bits is not None), int(any(lrt)), 30, 1, 0)
    tus, recs = [], []
    for fi, fr in enumerate(frames):
        pm = random_partition(g, rng, 6)
        r = O.encode_intra_frame(g, fr, bd, q, pm)
        fp = abi.FrameParams()
        fp.frame_type = 0
        fp.base_q_idx = q
        fp.disable_cdf_update = 0 if adapt else 1
        fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        for i in range(4):
            fp.lf_level[i] = lf[i]
        fp.lf_sharpness = int(rng.integers(0, 8))
        fp.cdef_damping = int(rng.integers(3, 7))
        cidx = np.zeros(g.sb_rows * g.sb_cols, np.uint8)
        O.deblock_frame(g, bd, r.blocks, r.rec, lf, fp.lf_sharpness)
        post = r.rec
        if cdef_bits is not None:
            fp.cdef_bits = cdef_bits
            for i in range(1 << cdef_bits):
                fp.cdef_y_strength[i] = int(rng.integers(0, 64))
                fp.cdef_uv_strength[i] = int(rng.integers(0, 64))
            cidx = rng.integers(0, 1 << cdef_bits, g.sb_rows * g.sb_cols).astype(np.uint8)
            post = O.cdef_frame(g, bd, r.blocks, fp, cidx, r.rec)
        for p in range(3):
            fp.lr_type[p] = lrt[p]
        fp.lr_unit_shift, fp.lr_uv_shift = ushift, uvshift
        units = random_lr_units(g, fp, rng)
        final = O.lr_frame(g, bd, fp, post, r.rec, units) if any(lrt) else post
        sy = packer.make_syms(g, r.blocks, r.coef, cdef_idx=cidx, lr_units=units)
        tu = b"\x12\x00" + (packer.pack_sequence_header(seq) if fi == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False)
        tus.append(tu)
        recs.append(O.crop(g, final))
    for name, dec in (("libaom", D.aom_decode), ("dav1d", D.dav1d_decode)):
        out = dec(tus)
        assert len(out) == len(frames), name
        for fi in range(len(frames)):
            for p in range(3):
                assert np.array_equal(out[fi][p], recs[fi][p]), (name, fi, p)


def test_film_grain_parameters_decode_and_add_grain():
    """Row f-4 (--film-grain): a stream whose sequence header announces film grain and whose frame headers carry our
    parameter set (spec 5.9.30: two luma scaling points, chroma from luma, no auto-regression) must decode in dav1d and
    libaom, and the decoders must add grain of about the signalled strength on top of the encoder's reconstruction."""
    from av1_base_b200 import abi, packer, synth
    from oracle import chain, decoders as D
    w, h, bd, n = 200, 136, 10, 3
    frames = synth.synth_clip(w, h, bd, n, seed=11, scene_len=100, noise=0.0)
    g, res = chain.encode_chain(frames, w, h, bd, 36, gop_period=2)
    for scaling in (0, 96):
        seq = abi.SeqParams(w, h, bd, 1, 0, 30, 1, 0, 1)
        tus = []
        for i, r in enumerate(res):
            fp = r.fp
            fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
            fp.grain_scaling, fp.grain_seed = scaling, 1234 + 77 * i
            sy = packer.make_syms(g, r.res.blocks, r.res.coef, cdef_idx=r.cdef_idx)
            tus.append(b"\x12\x00" + (packer.pack_sequence_header(seq) if i == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False))
        for name, dec in (("dav1d", D.dav1d_decode(tus)), ("dav1d+grain", D.dav1d_decode(tus, apply_grain=True)), ("libaom", D.aom_decode(tus))):
            assert len(dec) == n, name
            for i, r in enumerate(res):
                for p in range(3):
                    rec = O.crop(g, r.fin)[p].astype(np.int64)
                    d = dec[i][p].astype(np.int64) - rec
                    if scaling == 0 or name == "dav1d":
                        # apply_grain = 0, or a decoder asked for the pictures without grain: the plain reconstruction -- the
                        # parameters were parsed exactly, or the tile data behind them would not decode
                        assert not d.any(), (name, i, p)
                    else:             # sigma = scaling / 64 in 8-bit units, x4 at 10 bits (clipping at the range ends aside)
                        sigma = d.std()
                        assert 0.5 * (scaling / 64.0) * 4 < sigma < 1.6 * (scaling / 64.0) * 4, (name, i, p, sigma)
                        assert abs(d.mean()) < 1.0, (name, i, p, d.mean())
