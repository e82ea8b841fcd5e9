"""GPU parity of the whole key + inter frame path (through the C ABI): per frame the CUDA encoder must
reproduce the CPU oracle pipeline bit for bit -- hierarchical motion vectors, block side info,
quantised levels, reconstruction after deblock + CDEF -- and the produced chunk must decode in dav1d
AND libaom to exactly the encoder's reconstruction."""
import numpy as np
import pytest
from av1_base_b200 import encoder, synth
from oracle import pyoracle as O, decoders as D

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, bd, crf, tcl, trl, loop_filters, n_frames, frames_in_flight, keyint, preset (<= 5: loop restoration on)
    (64, 64, 8, 30, 0, 0, False, 4, 2, 240, 6),
    (200, 136, 10, 30, 0, 0, True, 5, 2, 240, 6),
    (328, 248, 8, 45, 1, 1, True, 6, 4, 4, 6),
    (640, 360, 10, 25, 2, 1, True, 5, 3, 240, 6),
    (200, 136, 10, 40, 0, 0, True, 5, 2, 3, 4),
    (328, 248, 8, 30, 1, 1, True, 7, 3, 240, 3),
    (640, 360, 10, 35, -1, -1, True, 4, 4, 240, 5),
]
# every case also through the device range coder (pack_path 4); 0 = automatic placement
CASES = [c + (pp,) for c in CASES for pp in (0, 4)]


def oracle_filters(g, bd, fp, res, frame, lr, acq):
    O.deblock_frame(g, bd, res.blocks, res.rec, list(fp.lf_level), fp.lf_sharpness)
    src = O.pad_planes(g, frame)
    idx = O.cdef_search(g, bd, res.blocks, fp, res.rec, src)
    fin, units = O.cdef_frame(g, bd, res.blocks, fp, idx, res.rec), None
    if lr:
        cand = O.lr_candidate((0, 0, 8), (0, 0, 8), 12, (0, 95))
        units, _ = O.lr_search(g, bd, fp, cand, fin, res.rec, src[0], (acq * acq * 5) >> 8)
        fin = O.lr_frame(g, bd, fp, fin, res.rec, [units, None, None])
    return fin, idx, units


def ac_q(bd, qidx):
    import re, os
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "av1_base_b200", "csrc", "av1_tables.h")).read()
    m = re.search(r"av1t_ac_q_%d\[\d+\] = \{(.*?)\};" % bd, txt, re.S)
    return [int(v) for v in re.findall(r"-?\d+", m.group(1))][qidx]


@pytest.mark.parametrize("w,h,bd,crf,tcl,trl,lf,nfr,fif,keyint,preset,pack_path", CASES)
def test_chunk_parity(w, h, bd, crf, tcl, trl, lf, nfr, fif, keyint, preset, pack_path):
    frames = synth.synth_clip(w, h, bd, nfr, seed=w + bd, scene_len=100)
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, tile_cols_log2=tcl, tile_rows_log2=trl,
                          frames_in_flight=fif, loop_filters=lf, keyint=keyint, preset=preset, pack_path=pack_path)
    lr = lf and preset <= 5
    tus = enc.encode_chunk(frames)
    assert len(tus) == nfr
    g = enc.geom
    q = enc.stats()["base_q_idx"]
    fp_key, fp_inter = enc.frame_params(), enc.inter_frame_params()
    pm = O.partition_fixed(g, 4)
    dec_d = D.dav1d_decode(tus)
    dec_a = D.aom_decode(tus)
    assert len(dec_d) == nfr and len(dec_a) == nfr
    prev_fin, prev_pyr = None, None
    for i, fr in enumerate(frames):
        key = i % keyint == 0
        assert enc.frame_is_key(i) == key
        pyr = O.pyramid(g, O.pad_planes(g, fr)[0])
        if key:
            ref = O.encode_intra_frame(g, fr, bd, fp_key.base_q_idx, pm)
        else:
            mvs = O.hme(g, pyr, prev_pyr, enc.me_lambda())
            ref = O.encode_inter_frame(g, fr, bd, q, pm, mvs, prev_fin)
            O.merge_skip_blocks(g, ref.blocks)
        blocks, coef = enc.frame_syms(i)
        for f in ("blk_log2", "y_mode", "uv_mode", "skip", "eob", "is_inter", "mv"):
            assert np.array_equal(blocks[f], ref.blocks[f]), (f, i)
        fin = ref.rec
        if lf:
            fpf = fp_key if key else fp_inter
            fin, idx, units = oracle_filters(g, bd, fpf, ref, fr, lr, ac_q(bd, fpf.base_q_idx))
            if lr:
                assert (fpf.lr_type[0], fpf.lr_type[1], fpf.lr_type[2]) == (3, 0, 0)
                assert enc.lr_units(i).tobytes() == units.tobytes(), ("restoration units", i)
        rec = enc.recon(i)
        orc = O.crop(g, fin)
        for p in range(3):
            hh, ww = (g.height, g.width) if p == 0 else (g.height // 2, g.width // 2)
            assert np.array_equal(coef[p], ref.coef[p]), ("coef", i, p)
            assert np.array_equal(rec[p], orc[p]), ("recon vs oracle", i, p)
            assert np.array_equal(dec_d[i][p], rec[p]), ("dav1d", i, p)
            assert np.array_equal(dec_a[i][p], rec[p]), ("libaom", i, p)
        prev_fin, prev_pyr = fin, pyr
    enc.close()


@pytest.mark.parametrize("w,h,bd,crf,tcl,trl", [(200, 136, 10, 20, -1, -1), (328, 248, 8, 40, 1, 1), (640, 360, 10, 30, -1, -1),
                                                (328, 248, 10, 2, 0, 1), (1920, 1080, 10, 35, -1, -1)])
def test_token_path_gives_identical_streams(w, h, bd, crf, tcl, trl):
    """Production mode tokenizes inter frames on the device (one token per coded symbol, tokens.h) and range-codes
    the token lists there too (one warp per tile, rc_kernel.cu); the host range coder over the same tokens is the
    CPU statement.  The bitstream must equal, byte for byte, what the host block walker writes from
    raster levels and from in-place packed symbols, and it must decode to the reconstruction.  CRF 2 exercises
    the Golomb escapes, 1080p the picture-edge 8x8 blocks and several tiles."""
    nfr = 5 if w < 1000 else 3
    frames = synth.synth_clip(w, h, bd, nfr, seed=3, scene_len=100)
    kw = dict(crf=crf, frames_in_flight=2, tile_cols_log2=tcl, tile_rows_log2=trl, tile_sb=6)
    a = encoder.Encoder(w, h, bd, pack_path=1, **kw)
    b = encoder.Encoder(w, h, bd, pack_path=4, **kw)    # device tokenizer + device range coder
    d = encoder.Encoder(w, h, bd, pack_path=2, **kw)
    h3 = encoder.Encoder(w, h, bd, pack_path=3, **kw)   # device tokenizer, range coder on the host
    ta, tb, td, th = a.encode_chunk(frames), b.encode_chunk(frames), d.encode_chunk(frames), h3.encode_chunk(frames)
    assert b.stats()["tokens"] > 0 and a.stats()["tokens"] == 0
    assert th == tb                                      # device range coder == host range coder
    assert [len(x) for x in ta] == [len(x) for x in tb]
    assert ta == tb
    assert td == tb
    c = encoder.Encoder(w, h, bd, keep_debug=True, pack_path=4, **kw)
    tc = c.encode_chunk(frames)
    assert tc == tb
    assert b.stats()["rc_ms"] > 0 and h3.stats()["rc_ms"] == 0
    dec = D.dav1d_decode(tb)
    for i in range(len(frames)):
        for p in range(3):
            assert np.array_equal(dec[i][p], c.recon(i)[p])
    for e in (a, b, c, d, h3):
        e.close()
