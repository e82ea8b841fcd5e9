"""Kernel bit-exact suite (BASELINE.json config 2): each CUDA kernel, called through the C ABI, against
libaom 3.13.1's own C functions (inverse transforms) and against the CPU oracle, which is itself
pinned against libaom's C line filters and both decoders (deblock / CDEF / loop restoration)."""
import ctypes as C
import numpy as np
import pytest
from av1_base_b200 import abi, kernels, synth
from oracle import aomsym, pyoracle as O

pytestmark = pytest.mark.gpu

SIZES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8), (16, 32), (32, 16),
         (32, 64), (64, 32), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16)]   # (w, h)
I32P = C.POINTER(C.c_int32)


def legal_types(w, h):
    m = max(w, h)
    return [0] if m == 64 else ([0, 9] if m == 32 else list(range(16)))


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("bd", [8, 10])
def test_inv_txfm_add_vs_libaom_and_oracle(w, h, bd):
    f = aomsym.func("av1_inv_txfm2d_add_%dx%d_c" % (w, h), None, [I32P, C.c_void_p, C.c_int, C.c_int, C.c_int])
    rng = np.random.default_rng(w * 131 + h * 7 + bd)
    cw, ch = min(w, 32), min(h, 32)
    lim = 1 << (bd + 7)
    n = 4096    # 4096 blocks per size and type
    for tx in legal_types(w, h):
        co = rng.integers(-lim, lim, (n, ch, cw))
        co[n // 3:2 * n // 3] = rng.integers(-300, 300, (n // 3, ch, cw)) * (rng.random((n // 3, ch, cw)) < 0.15)
        co[2 * n // 3:] = rng.choice([-lim, lim - 1], (n - 2 * n // 3, ch, cw))
        co[0] = 0
        co = np.ascontiguousarray(co, np.int32)
        pred = rng.integers(0, 1 << bd, (n, h, w)).astype(np.uint16)
        got, _ = kernels.inv_txfm_add(co, pred, w, h, tx, bd)
        mine = pred.copy()                       # every block against the oracle
        O.lib().orc_inv_txfm2d_add_batch(n, O.ptr(co), O.ptr(mine), w, h, tx, bd)
        assert np.array_equal(got, mine), ("oracle", w, h, tx, np.nonzero((got != mine).any(axis=(1, 2)))[0][:4])
        for b in range(0, n, 37):                # and a sample of them against libaom's C function
            theirs = pred[b].copy()
            buf = np.zeros(64 * 64, np.int32)
            ci = np.ascontiguousarray(co[b].T, np.int32)
            buf[:ci.size] = ci.ravel()
            f(buf.ctypes.data_as(I32P), O.ptr(theirs), w, tx, bd)
            assert np.array_equal(got[b], theirs), ("libaom", w, h, tx, b)


@pytest.mark.parametrize("w,h", SIZES)
def test_fwd_txfm_every_size_and_type_vs_oracle(w, h):
    """Row E4: the forward transform of every size (squares and 2:1 / 4:1 rectangles) and every legal type (DCT, ADST,
    flipADST, identity) -- 1024 blocks each bit-exact against the oracle's integer matrix form, and forward followed by the
    NORMATIVE inverse (the kernel pinned against libaom above) gives the residual back to within the transform's rounding."""
    rng = np.random.default_rng(w * 17 + h)
    n = 1024
    cw, ch = min(w, 32), min(h, 32)
    for tx in legal_types(w, h):
        resid = rng.integers(-1023, 1024, (n, h, w)).astype(np.int16)
        resid[:n // 4] = (rng.integers(-40, 41, (n // 4, h, w)) * (rng.random((n // 4, h, w)) < 0.3)).astype(np.int16)
        resid[0] = 0
        got, _ = kernels.fwd_txfm(resid, w, h, tx)
        want = np.zeros((n, ch, cw), np.int32)
        O.lib().orc_fwd_txfm2d_batch(n, O.ptr(resid), O.ptr(want), w, h, tx)
        assert np.array_equal(got, want), (w, h, tx, np.nonzero((got != want).any(axis=(1, 2)))[0][:4])
        if max(w, h) <= 32:      # (64-point transforms drop everything beyond 32x32: no round trip)
            small = np.ascontiguousarray(resid[:64])
            co, _ = kernels.fwd_txfm(small, w, h, tx)
            base = np.full((64, h, w), 512, np.uint16)
            rec, _ = kernels.inv_txfm_add(np.ascontiguousarray(co), base, w, h, tx, 10)
            back = rec.astype(np.int64) - 512
            inside = np.abs(small.astype(np.int64)) <= 500       # (the reconstruction clips at the sample range)
            assert np.abs(back - small)[inside].max() <= 2, (w, h, tx, np.abs(back - small)[inside].max())


def random_partition(g, rng):
    pm = O.partition_fixed(g, 6).reshape(g.h8, g.w8).copy()
    for y in range(0, g.h8, 8):
        for x in range(0, g.w8, 8):
            sub = pm[y:y + 8, x:x + 8]
            sub[...] = np.minimum(sub, int(rng.integers(3, 7)))
            for yy in range(0, 8, 4):
                for xx in range(0, 8, 4):
                    s2 = sub[yy:yy + 4, xx:xx + 4]
                    s2[...] = np.minimum(s2, int(rng.integers(3, 7)))
    return pm.ravel()


def encoded_frames(w, h, bd, q, n, seed):
    """n frames encoded by the oracle with random partitions: realistic blocking + side info."""
    rng = np.random.default_rng(seed)
    g = O.geom(w, h, 0, 0)
    frames = synth.synth_clip(w, h, bd, n, seed=seed, scene_len=1)
    res = [O.encode_intra_frame(g, fr, bd, q, random_partition(g, rng)) for fr in frames]
    return g, frames, res, rng


FRAME_CASES = [(64, 64, 8, 150), (200, 136, 10, 180), (328, 248, 8, 220), (640, 360, 10, 120)]


@pytest.mark.parametrize("w,h,bd,q", FRAME_CASES)
def test_deblock_vs_oracle(w, h, bd, q):
    g, frames, res, rng = encoded_frames(w, h, bd, q, 2, w + q)
    for trial in range(3):
        lf = [int(rng.integers(0, 64)) for _ in range(4)] if trial else [0, 0, 9, 9]
        if trial == 2:
            lf[1] = 0
        sharp = int(rng.integers(0, 8))
        blocks = np.stack([r.blocks for r in res])
        got, _ = kernels.deblock(w, h, bd, blocks, [r.rec for r in res], lf, sharp)
        for i, r in enumerate(res):
            ref = [p.copy() for p in r.rec]
            O.deblock_frame(g, bd, r.blocks, ref, lf, sharp)
            for p in range(3):
                assert np.array_equal(O.crop(g, got[i])[p], O.crop(g, ref)[p]), (trial, i, p, lf, sharp)


def cdef_params(rng, bits):
    fp = abi.FrameParams()
    fp.cdef_damping = int(rng.integers(3, 7))
    fp.cdef_bits = bits
    for i in range(1 << bits):
        fp.cdef_y_strength[i] = int(rng.integers(0, 64))
        fp.cdef_uv_strength[i] = int(rng.integers(0, 64))
    return fp


@pytest.mark.parametrize("w,h,bd,q", FRAME_CASES)
def test_cdef_filter_and_decision_vs_oracle(w, h, bd, q):
    g, frames, res, rng = encoded_frames(w, h, bd, q, 2, w * 3 + q)
    for r in res:
        O.deblock_frame(g, bd, r.blocks, r.rec, [12, 12, 8, 8], 0)
    blocks = np.stack([r.blocks for r in res])
    for bits in (0, 2, 3):
        fp = cdef_params(rng, bits)
        if bits == 2:
            fp.cdef_y_strength[0] = 0; fp.cdef_uv_strength[0] = 0
        # normative filter with forced presets
        forced = rng.integers(0, 1 << bits, (len(res), g.sb_rows * g.sb_cols)).astype(np.uint8)
        got, _, _ = kernels.cdef(w, h, bd, blocks, fp, [r.rec for r in res], forced_idx=forced)
        for i, r in enumerate(res):
            ref = O.cdef_frame(g, bd, r.blocks, fp, forced[i], r.rec)
            for p in range(3):
                assert np.array_equal(O.crop(g, got[i])[p], O.crop(g, ref)[p]), ("forced", bits, i, p)
        # decision against the source
        srcs = [O.pad_planes(g, fr) for fr in frames]
        got, idx, _ = kernels.cdef(w, h, bd, blocks, fp, [r.rec for r in res], src=srcs)
        for i, r in enumerate(res):
            want = O.cdef_search(g, bd, r.blocks, fp, r.rec, srcs[i])
            sb_live = np.zeros(g.sb_rows * g.sb_cols, bool)
            sk = r.blocks["skip"].reshape(g.h8, g.w8)
            for sr in range(g.sb_rows):
                for sc in range(g.sb_cols):
                    sb_live[sr * g.sb_cols + sc] = not sk[sr * 8:sr * 8 + 8, sc * 8:sc * 8 + 8].all()
            assert np.array_equal(idx[i][sb_live], want[sb_live]), ("decision", bits, i)
            ref = O.cdef_frame(g, bd, r.blocks, fp, idx[i], r.rec)
            for p in range(3):
                assert np.array_equal(O.crop(g, got[i])[p], O.crop(g, ref)[p]), ("decided", bits, i, p)


SGR_R = [(2, 1)] * 10 + [(0, 1)] * 4 + [(2, 0)] * 2


def random_lr_units(g, fp, rng, n):
    tmin, tmax = [-5, -23, -17], [10, 8, 46]
    units = []
    for p in range(3):
        lt = fp.lr_type[p]
        if lt == 0:
            units.append(None)
            continue
        us, ur, uc = O.lr_unit_grid(g, fp, p)
        u = np.zeros((n, ur, uc), abi.LR_UNIT_DTYPE)
        for f in range(n):
            for a in range(ur):
                for b in range(uc):
                    t = int(rng.integers(0, 3)) if lt == 3 else int(rng.integers(0, 2)) * (1 if lt == 1 else 2)
                    u[f, a, b]["type"] = t
                    if t == 1:
                        for nm in ("wiener_v", "wiener_h"):
                            for j in range(3):
                                u[f, a, b][nm][j] = 0 if (p > 0 and j == 0) else int(rng.integers(tmin[j], tmax[j] + 1))
                    elif t == 2:
                        st = int(rng.integers(0, 16))
                        r0, r1 = SGR_R[st]
                        x0 = int(rng.integers(-96, 32)) if r0 else 0
                        x1 = int(rng.integers(-32, 96)) if r1 else min(max(128 - x0, -32), 95)
                        u[f, a, b]["sgr_set"] = st
                        u[f, a, b]["sgr_xqd"][0] = x0
                        u[f, a, b]["sgr_xqd"][1] = x1
        units.append(u)
    return units


@pytest.mark.parametrize("w,h,bd,q", FRAME_CASES)
@pytest.mark.parametrize("lrt,ushift,uvshift", [((1, 1, 1), 0, 0), ((2, 2, 2), 0, 1), ((3, 3, 3), 1, 1), ((3, 0, 2), 2, 0)])
def test_loop_restoration_vs_oracle(w, h, bd, q, lrt, ushift, uvshift):
    g, frames, res, rng = encoded_frames(w, h, bd, q, 2, w * 5 + q + ushift)
    fp = cdef_params(rng, 1)
    for p in range(3):
        fp.lr_type[p] = lrt[p]
    fp.lr_unit_shift, fp.lr_uv_shift = ushift, uvshift
    deb, cdf = [], []
    for r in res:
        O.deblock_frame(g, bd, r.blocks, r.rec, [20, 20, 10, 10], 0)
        deb.append(r.rec)
        cdf.append(O.cdef_frame(g, bd, r.blocks, fp, rng.integers(0, 2, g.sb_rows * g.sb_cols).astype(np.uint8), r.rec))
    units = random_lr_units(g, fp, rng, len(res))
    got, _ = kernels.loop_restoration(w, h, bd, fp, cdf, deb, units)
    for i in range(len(res)):
        ref = O.lr_frame(g, bd, fp, cdf[i], deb[i], [u[i] if u is not None else None for u in units])
        for p in range(3):
            assert np.array_equal(O.crop(g, got[i])[p], O.crop(g, ref)[p]), (i, p)


@pytest.mark.parametrize("w,h,bd,q", FRAME_CASES)
def test_lr_search_vs_oracle(w, h, bd, q):
    """Encoder-side restoration decision: squared errors of NONE / Wiener / self-guided per unit and the choice."""
    g, frames, res, rng = encoded_frames(w, h, bd, q, 1, w * 3 + q)
    fp = cdef_params(rng, 1)
    fp.lr_type[0], fp.lr_type[1], fp.lr_type[2] = 3, 0, 0
    r = res[0]
    O.deblock_frame(g, bd, r.blocks, r.rec, [20, 20, 10, 10], 0)
    cdf = O.cdef_frame(g, bd, r.blocks, fp, rng.integers(0, 2, g.sb_rows * g.sb_cols).astype(np.uint8), r.rec)
    src = O.pad_planes(g, frames[0])
    for cand, bias in ((O.lr_candidate((0, 0, 8), (0, 0, 8), 12, (0, 95)), 3000), (O.lr_candidate(), 0),
                       (O.lr_candidate((2, -5, 20), (-3, 4, -10), 3, (-40, 70)), 100)):
        ref_units, ref_sse = O.lr_search(g, bd, fp, cand, cdf, r.rec, src[0], bias)
        units, sse, _ = kernels.lr_search(w, h, bd, cand, cdf, r.rec, src[0], bias)
        assert np.array_equal(sse, ref_sse)
        assert units.tobytes() == ref_units.tobytes()


ME_CASES = [(64, 64, 8), (200, 136, 10), (328, 248, 8), (640, 360, 10)]


@pytest.mark.parametrize("w,h,bd", ME_CASES)
def test_pyramid_and_hme_vs_oracle(w, h, bd):
    g = O.geom(w, h, 0, 0)
    frames = synth.synth_clip(w, h, bd, 4, seed=w + bd, scene_len=100)
    l0 = np.stack([O.pad_planes(g, fr)[0] for fr in frames])
    l1, l2, _ = kernels.pyramid(w, h, l0)
    pyr = [O.pyramid(g, l0[i]) for i in range(len(frames))]
    for i in range(len(frames)):
        assert np.array_equal(l1[i][:h // 2, :w // 2], pyr[i][1][:h // 2, :w // 2])
        assert np.array_equal(l2[i][:h // 4, :w // 4], pyr[i][2][:h // 4, :w // 4])
    for lam in (0, 40 << (bd - 8), 300 << (bd - 8)):
        mv, _ = kernels.hme(w, h, l0[1:], l0[:-1], lam=lam, bd=bd)
        for i in range(1, len(frames)):
            want = O.hme(g, pyr[i], pyr[i - 1], lam, bd)
            assert np.array_equal(mv[i - 1], want), (i, lam)


@pytest.mark.parametrize("w,h,bd", ME_CASES + [(1920, 1080, 10)])
def test_vector_field_regularisation_vs_oracle(w, h, bd):
    """hme + superblock-level rate-distortion sweeps (dominant vector histogram, bilinear quarter-sample SAD table per
    superblock, relaxation, 64x64 / 32x32 merge tests, checkerboard order)."""
    g = O.geom(w, h, 0, 0)
    n = 3 if w < 1000 else 2
    frames = synth.synth_clip(w, h, bd, n, seed=w + bd + 1, scene_len=100)
    l0 = np.stack([O.pad_planes(g, fr)[0] for fr in frames])
    pyr = [O.pyramid(g, l0[i]) for i in range(n)]
    for lam, lam_s, lam_r, passes in ((40 << (bd - 8), 40 << (bd - 8), 10 << (bd - 8), 2), (300 << (bd - 8), 150 << (bd - 8), 75 << (bd - 8), 3),
                                      (8, 2000, 1, 1), (8, 1, 3000, 1)):
        if w > 1000 and passes != 2:
            continue
        mv, _ = kernels.hme_sbrd(w, h, l0[1:], l0[:1].repeat(n - 1, 0), lam, lam_s, lam_r, passes, bd=bd)   # every frame against frame 0
        for i in range(1, n):
            want = O.me_sbrd(g, pyr[i], pyr[0], O.hme(g, pyr[i], pyr[0], lam, bd), lam_s, lam_r, passes)
            assert np.array_equal(mv[i - 1], want), (i, lam, lam_s, lam_r, passes)


@pytest.mark.parametrize("w,h,bd", ME_CASES + [(1920, 1080, 8), (3840, 2160, 10)])
def test_noise_estimate_vs_oracle(w, h, bd):
    g = O.geom(w, h, 0, 0)
    vals = []
    for noise in (0.0, 0.3, 1.0):
        fr = synth.synth_clip(w, h, bd, 1, seed=w + bd, scene_len=100, noise=noise, hdr=(w > 3000))[0]
        l0 = O.pad_planes(g, fr)[0]
        got = kernels.noise_estimate(w, h, l0)
        assert got == O.noise_estimate(g, l0), noise
        vals.append(got)
    assert vals[0] < vals[1] < vals[2]


@pytest.mark.parametrize("w,h,bd", ME_CASES + [(1920, 1080, 8), (3840, 2160, 10)])
def test_partition_smooth_vs_oracle(w, h, bd):
    g = O.geom(w, h, 0, 0)
    fr = synth.synth_clip(w, h, bd, 1, seed=w + bd, scene_len=100, noise=0.5)[0]
    l0 = O.pad_planes(g, fr)[0]
    for thr in (0, 100 << (bd - 8), 800 << (bd - 8), 1 << 20):
        got = kernels.partition_smooth(w, h, l0, thr)
        want = O.partition_smooth(g, l0, thr)
        assert np.array_equal(got, want), thr
    if w >= 640:
        sizes = set(np.unique(kernels.partition_smooth(w, h, l0, 400 << (bd - 8))).tolist())
        assert len(sizes & {5, 6}) >= 1 and 4 in sizes, sizes


@pytest.mark.parametrize("w,h,bd", ME_CASES + [(1920, 1080, 10)])
def test_temporal_filter_vs_oracle(w, h, bd):
    """mctf_kernel: normative interpolation of every neighbour + block / sample weights + weighted mean, all planes."""
    g = O.geom(w, h, 0, 0)
    n = 4 if w < 1000 else 3
    frames = synth.synth_clip(w, h, bd, n, seed=w + bd + 2, scene_len=100)
    padded = [O.pad_planes(g, fr) for fr in frames]
    pyr = [O.pyramid(g, p[0]) for p in padded]
    lam = 60 << (bd - 8)
    mvs = [O.me_sbrd(g, pyr[1], pyr[j], O.hme(g, pyr[1], pyr[j], lam, bd), lam, lam >> 2, 2) for j in range(n) if j != 1]
    nbs = [padded[j] for j in range(n) if j != 1]
    for thr_b in (3 << (2 * (bd - 8)), 40 << (2 * (bd - 8)), 4000 << (2 * (bd - 8))):
        got, _ = kernels.mctf(w, h, bd, padded[1], nbs, mvs, thr_b, 3 * thr_b)
        want = O.mctf(g, bd, padded[1], nbs, mvs, thr_b, 3 * thr_b)
        for p in range(3):
            assert np.array_equal(got[p], want[p]), (thr_b, p)
    assert not np.array_equal(want[0], padded[1][0])          # the strong filter changes the picture
    got, _ = kernels.mctf(w, h, bd, padded[1], [], [], 10, 30)   # no neighbours: identity
    for p in range(3):
        assert np.array_equal(got[p], padded[1][p])


def random_mvs(g, pm, rng, integer):
    step = 8 if integer else 2
    m = (rng.integers(-6, 7, (g.h8, g.w8, 2)) * step).astype(np.int16)
    pal = (rng.integers(-40, 41, (5, 2)) * step).astype(np.int16)
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


@pytest.mark.parametrize("w,h,bd", ME_CASES)
@pytest.mark.parametrize("q", [40, 140, 230])
def test_inter_encode_vs_oracle(w, h, bd, q):
    rng = np.random.default_rng(w + q)
    g = O.geom(w, h, 0, 0)
    frames = synth.synth_clip(w, h, bd, 2, seed=w + q, scene_len=100)
    pm = O.partition_fixed(g, 4)
    ref = O.encode_intra_frame(g, frames[0], bd, q, pm).rec
    src = O.pad_planes(g, frames[1])
    for integer, thr, merge in ((True, 0, False), (False, 0, True), (True, 3, True)):
        mvs = random_mvs(g, pm, rng, integer)
        if merge:   # large uniform regions so that 32x32 / 64x64 merges happen
            mvs = mvs.reshape(g.h8, g.w8, 2).copy()
            mvs[:g.h8 // 2] = mvs[0, 0]
            mvs = mvs.reshape(-1, 2)
        want = O.encode_inter_frame(g, frames[1], bd, q, pm, mvs, ref, tb_zero_thr=thr)
        if merge:
            O.merge_skip_blocks(g, want.blocks)
        rec, coef, blocks, _ = kernels.inter_encode(w, h, bd, q, pm, mvs, src, ref, tb_zero_thr=thr, merge_skip=merge)
        for f in ("blk_log2", "skip", "eob", "is_inter", "mv", "tx_type_y"):
            assert np.array_equal(blocks[f], want.blocks[f]), (f, integer)
        for p in range(3):
            hh, ww = (h, w) if p == 0 else (h // 2, w // 2)
            assert np.array_equal(coef[p], want.coef[p]), ("coef", p, integer)
            assert np.array_equal(rec[p][:hh, :ww], want.rec[p][:hh, :ww]), ("rec", p, integer)
