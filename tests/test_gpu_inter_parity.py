"""GPU parity of the whole key + inter frame path (through the C ABI): per frame the CUDA encoder must
reproduce the CPU oracle pipeline bit for bit -- hierarchical motion vectors, block side info,
quantised levels, reconstruction after deblock + CDEF -- and the produced chunk must decode in dav1d
AND libaom to exactly the encoder's reconstruction."""
import numpy as np
import pytest
from av1_base_b200 import encoder, synth
from oracle import pyoracle as O, decoders as D, chain

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, bd, crf, tcl, trl, loop_filters, n_frames, frames_in_flight, keyint, preset (<= 5: loop restoration on), gop_period
    (64, 64, 8, 30, 0, 0, False, 4, 2, 240, 6, 1),
    (200, 136, 10, 30, 0, 0, True, 5, 2, 240, 6, 1),
    (328, 248, 8, 45, 1, 1, True, 6, 4, 4, 6, 0),
    (640, 360, 10, 25, 2, 1, True, 10, 8, 240, 6, 0),
    (200, 136, 10, 40, 0, 0, True, 7, 2, 6, 4, 3),
    (328, 248, 8, 30, 1, 1, True, 9, 3, 240, 3, 0),
    (640, 360, 10, 35, -1, -1, True, 6, 4, 240, 5, 2),
    (328, 248, 10, 50, 0, 0, True, 13, 8, 240, 6, 0),
    (640, 360, 8, 38, -1, -1, True, 18, 8, 240, 6, 0),
    (328, 248, 10, 14, 0, 0, True, 6, 4, 240, 6, 0),      # fine quantiser on a noisy source: the automatic structure is the P chain
    (328, 248, 10, 14, 0, 0, True, 6, 4, 240, 6, 4),
]
# every case also through the device range coder (pack_path 4); 0 = automatic placement
CASES = [c + (pp,) for c in CASES for pp in (0, 4)]


@pytest.mark.parametrize("w,h,bd,crf,tcl,trl,lf,nfr,fif,keyint,preset,gop,pack_path", CASES)
def test_chunk_parity(w, h, bd, crf, tcl, trl, lf, nfr, fif, keyint, preset, gop, pack_path):
    """The whole decision chain of a closed GOP against oracle/chain.py: frame kinds of the hierarchy (anchors /
    non-reference frames, each predicted from the right picture at its own quantiser), key-frame partition, motion
    vectors after regularisation, side info, levels, reconstruction after the in-loop filters; dav1d and libaom decode
    the stream to the same pictures."""
    frames = synth.synth_clip(w, h, bd, nfr, seed=w + bd, scene_len=100)
    _check_chunk(frames, w, h, bd, crf, tcl, trl, lf, nfr, fif, keyint, preset, gop, pack_path)


@pytest.mark.parametrize("w,h,bd,crf,gop,qm,kvp,pack_path", [
    (200, 136, 10, 30, 3, (1, 15), True, 0),      # the daemon's flags (av1an.rs:14); 8x8 / 4x4 blocks at the picture edge
    (328, 248, 8, 44, 2, (0, 0), False, 4),       # steepest level; fixed 16x16 key-frame blocks (intra_fast.cu)
    (640, 360, 10, 12, 1, (1, 15), True, 0),      # fine quantiser, P chain
    (328, 248, 10, 36, 0, (5, 9), True, 4),
    (200, 136, 8, 55, 2, (15, 15), True, 0),      # level 15 = flat: signalled, nothing weighted
])
def test_chunk_parity_with_quantisation_matrices(w, h, bd, crf, gop, qm, kvp, pack_path):
    """Row f-4, --enable-qm 1 --qm-min A --qm-max B: the level follows each frame kind's quantiser index, every encode kernel
    (intra_encode_kernel, intra_recon_kernel, inter_encode_kernel<true>) weights the step of every position with the level's
    matrix (spec 7.12.3) exactly as the oracle chain does; dav1d and libaom decode to the reconstruction."""
    nfr = 7
    frames = synth.synth_clip(w, h, bd, nfr, seed=w + crf, scene_len=100, noise=0.3)
    kinds = _check_chunk(frames, w, h, bd, crf, -1, -1, True, nfr, 4, 240, 6, gop, pack_path, qm=qm, key_var_part=kvp)
    assert kinds[0] == 0 and 1 in kinds


@pytest.mark.parametrize("w,h,bd", [(202, 132, 10), (197, 131, 8)])
def test_sources_that_are_not_multiples_of_8(w, h, bd):
    """The library pads such a source to the coded size (edge replication, in the staging buffer), codes the padded frame like
    any other (CUDA == oracle chain on the padded pictures == both decoders) and signals the source size as render_size."""
    from tests.test_oracle_chain import unaligned_clip
    n, crf = 6, 34
    src, padded, cw, ch = unaligned_clip(w, h, bd, n, seed=w)
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, frames_in_flight=4, gop_period=2)
    assert (enc.geom.width, enc.geom.height) == (cw, ch)
    tus = enc.encode_chunk(src)
    assert len(tus) == n and [D.render_size_in_tu(t) for t in tus] == [(w, h)] * n
    g, want = chain.encode_chain(padded, cw, ch, bd, crf, gop_period=2, geom=enc.geom, batch=4)
    dec_d, dec_a = D.dav1d_decode(tus), D.aom_decode(tus)
    for i, r in enumerate(want):
        rec, orc = enc.recon(i), O.crop(g, r.fin)
        blocks, coef = enc.frame_syms(i)
        assert np.array_equal(blocks["mv"], r.res.blocks["mv"]) and np.array_equal(blocks["eob"], r.res.blocks["eob"]), i
        for p in range(3):
            assert np.array_equal(rec[p], orc[p]), ("recon vs oracle", i, p)
            assert np.array_equal(dec_d[i][p], rec[p]) and np.array_equal(dec_a[i][p], rec[p]), ("decoders", i, p)
    enc.close()


@pytest.mark.parametrize("noise,pack_path", [(1.0, 0), (0.05, 4)])
def test_daemon_settings_parity(noise, pack_path):
    """The daemon's fixed settings (av1an.rs:14: --crf 8 --preset 3 --film-grain 20 --enable-qm 1 --qm-min 1 --qm-max 15 --keyint 240
    --lookahead 40) all at once through the library: structure from the noise level (the P chain on the noisy clip, the
    hierarchy on the clean one), loop restoration, three regularisation sweeps, film-grain-strength temporal filter,
    quantisation matrices -- CUDA == oracle chain == dav1d (== libaom where no grain is signalled)."""
    w, h, bd, nfr = 328, 248, 10, 8
    frames = synth.synth_clip(w, h, bd, nfr, seed=21, scene_len=100, noise=noise)
    _check_chunk(frames, w, h, bd, 8, -1, -1, True, nfr, 8, 240, 3, 0, pack_path, qm=(1, 15), film_grain=20, lookahead=40)


def test_scene_change_inside_a_chunk_becomes_a_key_frame():
    """Row f-3: scene scores computed on the GPU as the pictures arrive (scene_score_kernel on the upload stream) restart the
    structure inside a chunk -- key frame at the cut, no motion search or temporal filter across it; frame kinds, vectors,
    levels and reconstruction equal the oracle chain's, which takes the same integer decision from orc_scene_score."""
    w, h, bd, nfr = 328, 248, 10, 30
    frames = synth.synth_clip(w, h, bd, nfr, seed=77, scene_len=13)      # cuts at 13 and 26: both at least 12 frames after a key frame
    kinds = _check_chunk(frames, w, h, bd, 36, 0, 0, True, nfr, 8, 240, 6, 0, 0)
    assert [i for i, k in enumerate(kinds) if k == 0] == [0, 13, 26]
    # the same clip with the detection off: one key frame
    enc = encoder.Encoder(w, h, bd, crf=36, keep_debug=True, frames_in_flight=8, scene_cut=False)
    enc.encode_chunk(frames)
    assert [i for i in range(nfr) if enc.frame_kind(i) == 0] == [0]
    enc.close()


def _check_chunk(frames, w, h, bd, crf, tcl, trl, lf, nfr, fif, keyint, preset, gop, pack_path, qm=None, key_var_part=True,
                 film_grain=0, lookahead=-1):
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, tile_cols_log2=tcl, tile_rows_log2=trl,
                          frames_in_flight=fif, loop_filters=lf, keyint=keyint, preset=preset, pack_path=pack_path, gop_period=gop,
                          qm=qm, key_var_part=key_var_part, film_grain=film_grain, lookahead=lookahead)
    lr = lf and preset <= 5
    tus = enc.encode_chunk(frames)
    assert len(tus) == nfr
    g, want = chain.encode_chain(frames, w, h, bd, crf, keyint=keyint, gop_period=gop, loop_filters=lf, lr=lr, geom=enc.geom, batch=fif,
                                 qm=qm, key_var_part=key_var_part, sbrd_passes=3 if preset <= 3 else 2, film_grain=film_grain,
                                 lookahead=lookahead)
    info = enc.chunk_info()
    if gop == 0:   # structure chosen from the noise level of the first picture
        gop, nb = chain.choose_structure(g, bd, crf, O.pad_planes(g, frames[0])[0])
        assert info["noise_b"] == nb and info["auto"]
    assert info["gop_period"] == gop
    assert enc.me_lambda() == chain.ac_q(bd, chain.quantisers(crf, gop)[1]) >> 1
    dec_d = D.dav1d_decode(tus)
    dec_a = D.aom_decode(tus)
    assert len(dec_d) == nfr and len(dec_a) == nfr
    kinds = set()
    for i, r in enumerate(want):
        kinds.add(r.kind)
        assert enc.frame_kind(i) == r.kind and enc.frame_is_key(i) == (r.kind == 0)
        fpe = enc.class_params(r.kind)
        assert fpe.base_q_idx == r.q and fpe.non_reference == r.fp.non_reference
        assert list(fpe.lf_level) == list(r.fp.lf_level) and list(fpe.cdef_y_strength) == list(r.fp.cdef_y_strength)
        assert fpe.using_qmatrix == r.fp.using_qmatrix and list(fpe.qm_level) == list(r.fp.qm_level)
        blocks, coef = enc.frame_syms(i)
        for f in ("blk_log2", "y_mode", "uv_mode", "skip", "eob", "is_inter", "mv"):
            assert np.array_equal(blocks[f], r.res.blocks[f]), (f, i)
        if lf:
            assert np.array_equal(enc.cdef_idx(i), r.cdef_idx), ("cdef_idx", i)
        if lr:
            assert (fpe.lr_type[0], fpe.lr_type[1], fpe.lr_type[2]) == (3, 0, 0)
            assert enc.lr_units(i).tobytes() == r.lr_units.tobytes(), ("restoration units", i)
        rec = enc.recon(i)
        orc = O.crop(g, r.fin)
        for p in range(3):
            assert np.array_equal(coef[p], r.res.coef[p]), ("coef", i, p)
            assert np.array_equal(rec[p], orc[p]), ("recon vs oracle", i, p)
            assert np.array_equal(dec_d[i][p], rec[p]), ("dav1d", i, p)
            if not film_grain:   # libaom always applies the grain a stream signals (dav1d above was asked not to)
                assert np.array_equal(dec_a[i][p], rec[p]), ("libaom", i, p)
    if gop > 1 and nfr > gop and keyint > gop:
        assert kinds == {0, 1, 2}
    assert enc.stats()["mctf_frames"] == sum(1 for r in want if r.filtered_from)
    enc.close()
    return [r.kind for r in want]


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
