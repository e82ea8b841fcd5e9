"""CPU checks of the encoder-side decisions added with the one-level hierarchy (oracle/chain.py is what the CUDA
path is compared with on the B200): streams of the whole chain -- key frame with 64x64 / 32x32 / 16x16 blocks,
anchors, non-reference frames at their own quantisers, regularised vector fields -- must decode in dav1d AND libaom
to the oracle's reconstruction; the partition and the regularisation are pinned against numpy restatements."""
import hashlib
import json
import os
import numpy as np
import pytest
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O, decoders as D, chain

HERE = os.path.dirname(os.path.abspath(__file__))


def pack_chain(w, h, bd, want, g, lr=False, render=(0, 0)):
    seq = abi.SeqParams(w, h, bd, 1, 1 if lr else 0, 30, 1, 0, 0, render[0], render[1])
    tus = []
    for i, r in enumerate(want):
        sy = packer.make_syms(g, r.res.blocks, r.res.coef, cdef_idx=r.cdef_idx, lr_units=[r.lr_units, None, None] if lr else None)
        body = packer.pack_frame(seq, r.fp, sy, with_td=False)
        if r.kind != 0:
            tok, _ = packer.pack_frame_tokens(seq, r.fp, sy, with_td=False)
            assert tok == body, "token path != block walker (frame %d)" % i
        tus.append(b"\x12\x00" + (packer.pack_sequence_header(seq) if r.kind == 0 else b"") + body)
    return tus


@pytest.mark.parametrize("w,h,bd,crf,n,keyint,gop,lr", [(200, 136, 10, 30, 10, 240, 4, False), (328, 248, 8, 48, 7, 5, 3, False),
                                                        (256, 192, 10, 38, 6, 240, 2, True), (192, 136, 8, 24, 5, 240, 1, False)])
def test_chain_streams_decode_to_oracle_reconstruction(w, h, bd, crf, n, keyint, gop, lr):
    frames = synth.synth_clip(w, h, bd, n, seed=w + crf, scene_len=100, noise=0.6)
    g, want = chain.encode_chain(frames, w, h, bd, crf, keyint=keyint, gop_period=gop, lr=lr)
    kinds = [r.kind for r in want]
    assert kinds == [chain.frame_kind(i, keyint, gop) for i in range(n)]
    if gop > 1 and n > gop and keyint > gop:
        assert set(kinds) == {0, 1, 2}
        assert want[1].q > want[gop].q            # non-reference frames are coded coarser than anchors
    tus = pack_chain(w, h, bd, want, g, lr)
    for dec in (D.dav1d_decode, D.aom_decode):
        out = dec(tus)
        assert len(out) == n
        for i in range(n):
            for p in range(3):
                assert np.array_equal(out[i][p], O.crop(g, want[i].fin)[p]), (dec.__name__, i, p)


@pytest.mark.parametrize("w,h,bd,crf,n,gop,qm", [(200, 136, 10, 30, 7, 3, (1, 15)), (328, 248, 8, 44, 5, 2, (0, 0)),
                                                 (256, 192, 10, 12, 4, 1, (1, 15)), (192, 136, 8, 55, 4, 2, (15, 15)),
                                                 (320, 192, 10, 36, 3, 1, (5, 9))])
def test_quantisation_matrices_decode(w, h, bd, crf, n, gop, qm):
    """Row f-4, --enable-qm 1 --qm-min A --qm-max B (av1an.rs:14): frame headers signal using_qmatrix with the level the
    frame's quantiser index maps to, the quantiser / dequantiser weight every position's step with Quantizer_Matrix (spec
    7.12.3; csrc/av1_qm_tables.h, read out of the libaom binary).  dav1d (its own copy of the constants) and libaom must
    decode the chain's streams to the oracle's reconstruction: that pins the tables, the header syntax and the arithmetic.
    Smooth clips make the key frames use 64x64 / 32x32 transforms (the 32x32 matrix) besides 16x16 / 8x8 / 4x4."""
    frames = synth.synth_clip(w, h, bd, n, seed=w + crf, scene_len=100, noise=0.3)
    g, want = chain.encode_chain(frames, w, h, bd, crf, gop_period=gop, qm=qm)
    g0, flat = chain.encode_chain(frames, w, h, bd, crf, gop_period=gop)
    lv = [O.qm_level(r.q, qm[0], qm[1]) for r in want]
    for r, l in zip(want, lv):
        assert r.fp.using_qmatrix == 1 and list(r.fp.qm_level) == [l, l] and qm[0] <= l <= qm[1]
    if qm[0] < 15:
        assert {5, 6} & set(np.unique(want[0].res.blocks["blk_log2"]).tolist())      # large transforms are in the test
        # the matrices change what is coded (steps of the high frequencies grow), level 15 is the flat quantiser
        assert any(not np.array_equal(a.res.coef[0], b.res.coef[0]) for a, b in zip(want, flat))
    else:
        assert all(np.array_equal(a.res.coef[p], b.res.coef[p]) for a, b in zip(want, flat) for p in range(3))
    tus = pack_chain(w, h, bd, want, g)
    for dec in (D.dav1d_decode, D.aom_decode):
        out = dec(tus)
        assert len(out) == n
        for i in range(n):
            for p in range(3):
                assert np.array_equal(out[i][p], O.crop(g, want[i].fin)[p]), (dec.__name__, i, p)


def unaligned_clip(w, h, bd, n, seed):
    """n pictures of w x h (any size; 4:2:0 chroma rounded up) and the same pictures padded to multiples of 8 by edge
    replication: what csrc/encoder.cc stage() makes of a source that is not a multiple of 8."""
    cw, ch = (w + 7) & ~7, (h + 7) & ~7
    big = synth.synth_clip(cw, ch, bd, n, seed=seed, scene_len=100, noise=0.4)
    src = [[np.ascontiguousarray(f[0][:h, :w]), np.ascontiguousarray(f[1][:(h + 1) // 2, :(w + 1) // 2]),
            np.ascontiguousarray(f[2][:(h + 1) // 2, :(w + 1) // 2])] for f in big]
    padded = [[np.pad(f[0], ((0, ch - h), (0, cw - w)), mode="edge"),
               np.pad(f[1], ((0, ch // 2 - (h + 1) // 2), (0, cw // 2 - (w + 1) // 2)), mode="edge"),
               np.pad(f[2], ((0, ch // 2 - (h + 1) // 2), (0, cw // 2 - (w + 1) // 2)), mode="edge")] for f in src]
    return src, padded, cw, ch


@pytest.mark.parametrize("w,h,bd", [(202, 132, 10), (197, 131, 8)])
def test_render_size_for_sources_that_are_not_multiples_of_8(w, h, bd):
    """1920x804-like sources: the coded frame is the source padded to multiples of 8, every frame header carries the source size
    as render_size (spec 5.9.6).  Both decoders parse the headers (the tile data behind them decodes to the reconstruction of
    the padded frame) and the bits at the header positions hold the source size."""
    n = 4
    src, padded, cw, ch = unaligned_clip(w, h, bd, n, seed=w)
    assert (cw, ch) != (w, h)
    g, want = chain.encode_chain(padded, cw, ch, bd, 34, gop_period=2)
    tus = pack_chain(cw, ch, bd, want, g, render=(w, h))
    assert [D.render_size_in_tu(t) for t in tus] == [(w, h)] * n
    assert [D.render_size_in_tu(t) for t in pack_chain(cw, ch, bd, want, g)] == [None] * n
    for dec in (D.dav1d_decode, D.aom_decode):
        out = dec(tus)
        assert len(out) == n
        for i in range(n):
            assert out[i][0].shape == (ch, cw)
            for p in range(3):
                assert np.array_equal(out[i][p], O.crop(g, want[i].fin)[p]), (dec.__name__, i, p)
            # inside the source area the coded picture is the source's reconstruction; the rest is padding
            assert D.psnr(out[i][0][:h, :w], src[i][0], bd) > 30


@pytest.mark.parametrize("noise", [1.0, 0.05])
def test_daemon_settings_chain_decodes(noise):
    """The daemon's fixed settings (av1an.rs:14: --crf 8 --preset 3 --film-grain 20 --enable-qm 1 --qm-min 1 --qm-max 15 --keyint 240
    --lookahead 40) all at once: structure from the noise level, loop restoration (preset <= 5), three regularisation sweeps
    (preset <= 3), film-grain-strength temporal filter, quantisation matrices.  Both decoders reproduce the reconstruction."""
    w, h, bd, n = 328, 248, 10, 8
    frames = synth.synth_clip(w, h, bd, n, seed=21, scene_len=100, noise=noise)
    g, want = chain.encode_chain(frames, w, h, bd, 8, keyint=240, gop_period=0, lr=True, film_grain=20, lookahead=40, qm=(1, 15), sbrd_passes=3)
    gop, _ = chain.choose_structure(g, bd, 8, O.pad_planes(g, frames[0])[0])
    assert gop == (1 if noise >= 1.0 else chain.DEFAULT_GOP_PERIOD)      # a fine quantiser on a noisy source: the P chain
    tus = pack_chain(w, h, bd, want, g, lr=True)
    for dec in (D.dav1d_decode, D.aom_decode):
        out = dec(tus)
        assert len(out) == n
        for i in range(n):
            for p in range(3):
                assert np.array_equal(out[i][p], O.crop(g, want[i].fin)[p]), (dec.__name__, i, p)
            assert D.psnr(out[i][0], frames[i][0], bd) > (44 if noise < 1 else 40)


def test_random_configurations_decode():
    """A seeded sample of tools/fuzz_chain.py (6000 random configurations were run in round 2 without a mismatch): sizes that are and
    are not multiples of 8, bit depths, quantisers, structures, loop restoration, quantisation matrices -- both decoders reproduce
    the chain's reconstruction."""
    import subprocess, sys
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_chain.py"), "424242", "16"], capture_output=True, text=True, cwd=root)
    assert r.returncode == 0 and "FINISHED 16 bad 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_qm_level_mapping():
    """aom_get_qmlevel (SVT-AV1 and libaom map the quantiser index to a level the same way)."""
    assert O.qm_level(0, 1, 15) == 1 and O.qm_level(255, 1, 15) == 15 and O.qm_level(128, 0, 15) == 8
    assert [O.qm_level(q, 8, 15) for q in (0, 31, 32, 255)] == [8, 8, 9, 15]


def np_partition_smooth(g, Y, thr):
    """numpy restatement: least information possible shared with the C++ (box sums via reshape, plane via mgrid)."""
    pm = O.partition_fixed(g, 4).reshape(g.h8, g.w8).copy()
    h, w = g.height, g.width
    B = Y[:h, :w].astype(np.int64).reshape(h // 4, 4, w // 4, 4).sum((1, 3))
    for bl, N in ((6, 64), (5, 32)):
        n, n8 = N // 4, N // 8
        for y0 in range(0, h - N + 1, N):
            for x0 in range(0, w - N + 1, N):
                if pm[y0 // 8, x0 // 8] >= 6:
                    continue
                b = B[y0 // 4:y0 // 4 + n, x0 // 4:x0 // 4 + n]
                S = b.sum(); gx = 2 * (b[:, n // 2:].sum() - b[:, :n // 2].sum()); gy = 2 * (b[n // 2:, :].sum() - b[:n // 2, :].sum())
                i, j = np.mgrid[0:n, 0:n]
                if np.abs(n ** 3 * b - (n * S + (2 * j - n + 1) * gx + (2 * i - n + 1) * gy)).max() <= thr * n ** 3:
                    pm[y0 // 8:y0 // 8 + n8, x0 // 8:x0 // 8 + n8] = bl
    return pm.reshape(-1)


@pytest.mark.parametrize("w,h,bd", [(328, 248, 8), (640, 360, 10), (200, 136, 10)])
def test_partition_smooth_vs_numpy(w, h, bd):
    g = O.geom(w, h, 0, 0)
    fr = synth.synth_clip(w, h, bd, 1, seed=w, scene_len=100, noise=0.5)[0]
    l0 = O.pad_planes(g, fr)[0]
    seen = set()
    for thr in (0, 60 << (bd - 8), 300 << (bd - 8), 800 << (bd - 8), 1 << 20):
        got = O.partition_smooth(g, l0, thr)
        assert np.array_equal(got, np_partition_smooth(g, l0, thr)), thr
        seen |= set(np.unique(got).tolist())
    assert {4, 6} <= seen
    # a perfect plane is smooth at threshold 0, wherever whole 64x64 blocks fit
    yy, xx = np.mgrid[0:g.rows[0], 0:g.stride[0]]
    plane = (100 + 2 * xx + 3 * yy).astype(np.uint16) % 1024
    m = O.partition_smooth(g, plane[:, :] * 0 + (100 + xx // 4 * 4).astype(np.uint16), 0).reshape(g.h8, g.w8)
    assert m[0, 0] == 6


def test_mv_dominant_and_smoothing_properties():
    w, h, bd = 328, 248, 10
    g = O.geom(w, h, 0, 0)
    frames = synth.synth_clip(w, h, bd, 2, seed=9, scene_len=100)
    pyr = [O.pyramid(g, O.pad_planes(g, f)[0]) for f in frames]
    mv = O.hme(g, pyr[1], pyr[0], 80, bd)
    sm0 = O.me_sbrd(g, pyr[1], pyr[0], mv, 1, 1, 1)
    n1x, n1y = (w + 15) // 16, (h + 15) // 16
    f0 = mv.reshape(g.h8, g.w8, 2)[::2, ::2].reshape(-1, 2)
    f1 = sm0.reshape(g.h8, g.w8, 2)[::2, ::2].reshape(-1, 2)
    assert f0.shape[0] == n1x * n1y
    # a strong cost makes the field more uniform, never less
    sm = O.me_sbrd(g, pyr[1], pyr[0], mv, 100000, 100000, 3)
    fs = sm.reshape(g.h8, g.w8, 2)[::2, ::2].reshape(-1, 2)
    assert len(np.unique(fs, axis=0)) <= len(np.unique(f0, axis=0))
    # dominant vector: numpy restatement of the hashed histogram
    import ctypes as C
    dom = np.zeros(2, np.int16)
    O.lib().orc_mv_dominant(O.ptr(np.ascontiguousarray(f0)), f0.shape[0], O.ptr(dom))
    keys = ((f0[:, 0].astype(np.int64) & 0xFFFF) << 16) | (f0[:, 1].astype(np.int64) & 0xFFFF)
    bins = ((keys * 2654435761) & 0xFFFFFFFF) >> 22
    cnt = np.bincount(bins, minlength=1024)
    b = int(np.argmax(cnt))
    cand = f0[bins == b]
    best = max(map(tuple, cand.tolist()))
    assert (int(dom[0]), int(dom[1])) == best
    assert len(f1) == len(f0)


def test_noise_estimate_and_structure_choice():
    """orc_noise_estimate against a numpy restatement; the estimate follows the noise, ignores noiseless (saturated /
    letterbox) blocks, and the structure choice turns to the P chain only where the quantiser is fine against it."""
    w, h, bd = 328, 248, 10
    g = O.geom(w, h, 0, 0)
    est = []
    for noise in (0.0, 0.25, 1.0, 2.0):
        fr = synth.synth_clip(w, h, bd, 1, seed=2, scene_len=100, noise=noise)[0]
        l0 = O.pad_planes(g, fr)[0]
        I = l0[:h, :w].astype(np.int64)
        L = np.abs(I[:-2, :-2] - 2 * I[:-2, 1:-1] + I[:-2, 2:] - 2 * I[1:-1, :-2] + 4 * I[1:-1, 1:-1] - 2 * I[1:-1, 2:] + I[2:, :-2] - 2 * I[2:, 1:-1] + I[2:, 2:])
        Lp = np.zeros((h, w), np.int64); Lp[1:-1, 1:-1] = L
        B = np.array([[Lp[by * 16 + 1:by * 16 + 15, bx * 16 + 1:bx * 16 + 15].sum() for bx in range(w // 16)] for by in range(h // 16)]).ravel()
        bins = np.minimum(B >> 4, 4095)
        bins = np.sort(bins[bins > 0])
        want = 0 if len(bins) == 0 else int(bins[(len(bins) + 3) // 4 - 1]) * 16 + 8
        got = O.noise_estimate(g, l0)
        assert got == want, noise
        est.append(got)
    assert est[0] < est[1] < est[2] < est[3]
    # a letterboxed picture (noiseless bars) keeps the estimate of its active area
    fr = synth.synth_clip(w, h, bd, 1, seed=2, scene_len=100, noise=1.0)[0]
    l0 = O.pad_planes(g, fr)[0]
    full = O.noise_estimate(g, l0)
    l0[:64, :] = 64; l0[h - 64:h, :] = 64
    assert abs(O.noise_estimate(g, l0) - full) <= 0.15 * full
    noisy = O.pad_planes(g, synth.synth_clip(w, h, bd, 1, seed=2, scene_len=100, noise=1.0)[0])[0]
    clean = O.pad_planes(g, synth.synth_clip(w, h, bd, 1, seed=2, scene_len=100, noise=0.1)[0])[0]
    assert chain.choose_structure(g, bd, 14, noisy)[0] == 1 and chain.choose_structure(g, bd, 44, noisy)[0] == chain.DEFAULT_GOP_PERIOD
    assert chain.choose_structure(g, bd, 14, clean)[0] == chain.DEFAULT_GOP_PERIOD


def test_temporal_filter_properties():
    """orc_mctf against a numpy restatement built on pyoracle.inter_predict (the normative predictor pinned elsewhere),
    and what it is for: independent noise averages out."""
    w, h, bd = 200, 136, 10
    g = O.geom(w, h, 0, 0)
    clean = synth.synth_clip(w, h, bd, 3, seed=5, scene_len=100, noise=0.0)
    noisy = synth.synth_clip(w, h, bd, 3, seed=5, scene_len=100, noise=1.0)
    padded = [O.pad_planes(g, f) for f in noisy]
    pyr = [O.pyramid(g, p[0]) for p in padded]
    mvs = [O.hme(g, pyr[1], pyr[j], 100, bd) for j in (0, 2)]
    thr_b, thr_p = 900, 2700
    got = O.mctf(g, bd, padded[1], [padded[0], padded[2]], mvs, thr_b, thr_p)
    # numpy restatement
    num = [p.astype(np.int64) * 256 for p in padded[1]]
    den = [np.full(p.shape, 256, np.int64) for p in padded[1]]
    for nb, mv in zip((padded[0], padded[2]), mvs):
        m = mv.reshape(g.h8, g.w8, 2)
        for by in range((h + 15) // 16):
            for bx in range((w + 15) // 16):
                v = m[by * 2, bx * 2]
                bw, bh = min(16, w - bx * 16), min(16, h - by * 16)
                preds = [O.inter_predict(nb[p], (bx * 16) >> (p > 0), (by * 16) >> (p > 0), 16 >> (p > 0), 16 >> (p > 0), v, int(p > 0), bd,
                                         ref_w=w >> (p > 0), ref_h=h >> (p > 0)).astype(np.int64) for p in range(3)]
                cy = padded[1][0][by * 16:by * 16 + bh, bx * 16:bx * 16 + bw].astype(np.int64)
                mse = int(((preds[0][:bh, :bw] - cy) ** 2).sum()) // (bw * bh)
                wb = min(16, max(0, 16 - (16 * mse) // thr_b))
                if wb == 0:
                    continue
                for p in range(3):
                    ss = int(p > 0)
                    y0, x0, hh, ww = (by * 16) >> ss, (bx * 16) >> ss, bh >> ss, bw >> ss
                    pr = preds[p][:hh, :ww]
                    d = pr - padded[1][p][y0:y0 + hh, x0:x0 + ww].astype(np.int64)
                    wgt = wb * np.clip(16 - (16 * d * d) // thr_p, 0, 16)
                    num[p][y0:y0 + hh, x0:x0 + ww] += wgt * pr
                    den[p][y0:y0 + hh, x0:x0 + ww] += wgt
    for p in range(3):
        assert np.array_equal(got[p], ((num[p] + den[p] // 2) // den[p]).astype(np.uint16)), p
    cy = clean[1][0].astype(np.float64)
    e_in = ((noisy[1][0] - cy) ** 2).mean(); e_out = ((O.crop(g, got)[0] - cy) ** 2).mean()
    assert e_out < 0.7 * e_in, (e_in, e_out)


def test_chain_golden():
    """Committed digest of a chain run (tests/golden/chain_digest.json, written by tests/golden/make_golden.py): pins
    the oracle's decisions -- any change to the hierarchy, the partition rule or the regularisation shows up here."""
    from tests.golden import make_golden as G
    want = json.load(open(os.path.join(HERE, "golden", "chain_digest.json")))
    assert G.chain_digest() == want
