"""Pins the oracle's normative kernels against libaom 3.13.1's own C reference functions
(BASELINE.json config 2: "inverse DCT/ADST 4x4-64x64 ... vs libaom C on random coeffs/pixels").
libaom's internal `_c` functions are reached through the ELF .symtab of the bundled shared object."""
import ctypes as C
import numpy as np
import pytest
from oracle import aomsym, pyoracle as O

I32P = C.POINTER(C.c_int32)
SIZES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8), (16, 32), (32, 16),
         (32, 64), (64, 32), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16)]   # (w, h)
TX_NAMES = ["DCT_DCT", "ADST_DCT", "DCT_ADST", "ADST_ADST", "FLIPADST_DCT", "DCT_FLIPADST", "FLIPADST_FLIPADST",
            "ADST_FLIPADST", "FLIPADST_ADST", "IDTX", "V_DCT", "H_DCT", "V_ADST", "H_ADST", "V_FLIPADST", "H_FLIPADST"]


def legal_types(w, h):
    m = max(w, h)
    if m == 64:
        return [0]
    if m == 32:
        return [0, 9]          # DCT_DCT, IDTX
    return list(range(16))


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("bd", [8, 10])
def test_inv_txfm2d_add_vs_libaom(w, h, bd):
    f = aomsym.func("av1_inv_txfm2d_add_%dx%d_c" % (w, h), None, [I32P, C.c_void_p, C.c_int, C.c_int, C.c_int])
    rng = np.random.default_rng(w * 100 + h + bd)
    cw, ch = min(w, 32), min(h, 32)
    lim = 1 << (bd + 7)
    for tx in legal_types(w, h):
        for case in range(6):
            if case == 0:
                co = rng.integers(-lim, lim, (ch, cw))
            elif case == 1:
                co = np.zeros((ch, cw), np.int64); co[0, 0] = rng.integers(-lim, lim)
            elif case == 2:
                co = np.zeros((ch, cw), np.int64)
                k = min(10, ch * cw)
                co.flat[rng.choice(min(ch * cw, 64), k, replace=False)] = rng.integers(-lim // 4, lim // 4, k)
            elif case == 3:
                co = rng.choice([-lim, lim - 1], (ch, cw))
            else:
                co = rng.integers(-200, 200, (ch, cw))
            co = np.ascontiguousarray(co, np.int32)
            pred = rng.integers(0, 1 << bd, (h, w)).astype(np.uint16)
            mine = pred.copy()
            O.lib().orc_inv_txfm2d_add(O.ptr(co), cw, O.ptr(mine), w, w, h, tx, bd)
            # libaom >= 3.5 keeps coefficients transposed: input[col * ch + row]; 64-point sizes take
            # the 32x32 (or 32xN) coded area only
            theirs = pred.copy()
            ci = np.ascontiguousarray(co.T, np.int32)
            buf = np.zeros(64 * 64, np.int32); buf[:ci.size] = ci.ravel()
            f(buf.ctypes.data_as(I32P), O.ptr(theirs), w, tx, bd)
            assert np.array_equal(mine, theirs), (w, h, TX_NAMES[tx], case)


PRED = {0: "dc", 1: "v", 2: "h", 9: "smooth", 10: "smooth_v", 11: "smooth_h", 12: "paeth"}


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("bd", [8, 10])
def test_intra_predictors_vs_libaom(n, bd):
    rng = np.random.default_rng(n + bd)
    for trial in range(4):
        edge_a = rng.integers(0, 1 << bd, 2 * n + 17).astype(np.uint16)
        edge_l = rng.integers(0, 1 << bd, 2 * n + 17).astype(np.uint16)
        edge_l[15] = edge_a[15]   # shared top-left
        above = edge_a[16:]; left = edge_l[16:]
        for mode, nm in PRED.items():
            f = aomsym.func("aom_highbd_%s_predictor_%dx%d_c" % (nm, n, n), None,
                            [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_int])
            a = np.zeros((n, n), np.uint16); b = np.zeros((n, n), np.uint16)
            f(O.ptr(a), n, above.ctypes.data, left.ctypes.data, bd)
            O.lib().orc_intra_predict(O.ptr(b), n, n, n, C.c_void_p(above.ctypes.data), C.c_void_p(left.ctypes.data),
                                      mode, 0, 1, 1, bd)
            assert np.array_equal(a, b), (nm, n)
        # directional: z1 (<90), z2 (90..180), z3 (>180) with every angle delta
        z1 = aomsym.func("av1_highbd_dr_prediction_z1_c", None, [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_void_p,
                                                                 C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
        z2 = aomsym.func("av1_highbd_dr_prediction_z2_c", None, [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_void_p,
                                                                 C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int])
        z3 = aomsym.func("av1_highbd_dr_prediction_z3_c", None, [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_void_p,
                                                                 C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
        from oracle.aomsym import load
        _, _, es = load()
        dr = np.frombuffer(es.read("dr_intra_derivative"), np.int16)
        base_angle = {1: 90, 2: 180, 3: 45, 4: 135, 5: 113, 6: 157, 7: 203, 8: 67}
        for mode, ang in base_angle.items():
            for delta in range(-3, 4):
                p = ang + 3 * delta
                if p in (90, 180) or p <= 0 or p >= 270:
                    continue
                a = np.zeros((n, n), np.uint16); b = np.zeros((n, n), np.uint16)
                if p < 90:
                    z1(O.ptr(a), n, n, n, above.ctypes.data, left.ctypes.data, 0, int(dr[p]), 1, bd)
                elif p < 180:
                    z2(O.ptr(a), n, n, n, above.ctypes.data, left.ctypes.data, 0, 0, int(dr[180 - p]), int(dr[p - 90]), bd)
                else:
                    z3(O.ptr(a), n, n, n, above.ctypes.data, left.ctypes.data, 0, 1, int(dr[270 - p]), bd)
                O.lib().orc_intra_predict(O.ptr(b), n, n, n, C.c_void_p(above.ctypes.data),
                                          C.c_void_p(left.ctypes.data), mode, delta, 1, 1, bd)
                assert np.array_equal(a, b), (mode, delta, n)


@pytest.mark.parametrize("n", [8, 16])
@pytest.mark.parametrize("bd", [8, 10])
def test_motion_search_sad_vs_libaom(n, bd):
    """the block SAD the motion search is built on against libaom's aom_highbd_sad{8x8,16x16}_c on the same
    samples (high-bit-depth buffers are passed as CONVERT_TO_BYTEPTR pointers: the address shifted right by one)."""
    f = aomsym.func("aom_highbd_sad%dx%d_c" % (n, n), C.c_uint, [C.c_void_p, C.c_int, C.c_void_p, C.c_int])
    rng = np.random.default_rng(n + bd)
    w, h, stride = 96, 80, 128
    cur = np.zeros((h, stride), np.uint16); ref = np.zeros((h, stride), np.uint16)
    cur[:, :w] = rng.integers(0, 1 << bd, (h, w)); ref[:, :w] = rng.integers(0, 1 << bd, (h, w))
    for _ in range(200):
        bx, by = int(rng.integers(8, w - n - 8)), int(rng.integers(8, h - n - 8))
        dx, dy = int(rng.integers(-8, 9)), int(rng.integers(-8, 9))
        mine = O.lib().orc_sad_block(O.ptr(cur), O.ptr(ref), stride, w, h, bx, by, n, dx, dy)
        a = cur.ctypes.data + 2 * (by * stride + bx)
        b = ref.ctypes.data + 2 * ((by + dy) * stride + bx + dx)
        assert a % 2 == 0 and b % 2 == 0
        theirs = f(C.c_void_p(a >> 1), stride, C.c_void_p(b >> 1), stride)
        assert mine == theirs, (bx, by, dx, dy)
