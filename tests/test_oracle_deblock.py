"""Deblocking filter: oracle line filters vs libaom's aom_highbd_lpf_*_c, and frame-level filtering
validated through both decoders."""
import ctypes as C
import numpy as np
import pytest
from oracle import aomsym, pyoracle as O


def limits(lvl, sharp):
    shift = 2 if sharp > 4 else (1 if sharp > 0 else 0)
    limit = min(max(lvl >> shift, 1), 9 - sharp) if sharp > 0 else max(1, lvl >> shift)
    return 2 * (lvl + 2) + limit, limit, lvl >> 4


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("size,name", [(4, "4"), (6, "6"), (8, "8"), (16, "14")])
def test_lpf_lines_vs_libaom(bd, size, name):
    rng = np.random.default_rng(size + bd)
    U8P = C.POINTER(C.c_uint8)
    fv = aomsym.func("aom_highbd_lpf_vertical_%s_c" % name, None, [C.c_void_p, C.c_int, U8P, U8P, U8P, C.c_int])
    fh = aomsym.func("aom_highbd_lpf_horizontal_%s_c" % name, None, [C.c_void_p, C.c_int, U8P, U8P, U8P, C.c_int])
    for trial in range(300):
        lvl = int(rng.integers(1, 64)); sharp = int(rng.integers(0, 8))
        bl, li, th = limits(lvl, sharp)
        b = (C.c_uint8 * 16)(*([bl] * 16)); l = (C.c_uint8 * 16)(*([li] * 16)); t = (C.c_uint8 * 16)(*([th] * 16))
        # smooth-ish content so that the flat / wide branches are exercised too
        kind = trial % 3
        base = int(rng.integers(20, (1 << bd) - 20))
        if kind == 0:
            img = rng.integers(0, 1 << bd, (16, 16))
        elif kind == 1:
            img = base + rng.integers(-2 << (bd - 8), (2 << (bd - 8)) + 1, (16, 16))
            img[:, 8:] += int(rng.integers(-6, 7)) << (bd - 8)
        else:
            img = base + rng.integers(-1, 2, (16, 16)) * (1 << (bd - 8))
            img[8:, :] += int(rng.integers(-3, 4)) << (bd - 8)
        img = np.clip(img, 0, (1 << bd) - 1).astype(np.uint16)
        # vertical edge at column 8, 4 rows
        a = img.copy(); m = img.copy()
        fv(C.c_void_p(a.ctypes.data + (4 * 16 + 8) * 2), 16, b, l, t, bd)
        for i in range(4):
            O.lib().orc_lf_filter_line(C.c_void_p(m.ctypes.data + ((4 + i) * 16 + 8) * 2), 1, size, lvl, sharp, bd)
        assert np.array_equal(a, m), ("v", size, lvl, sharp, kind)
        a = img.copy(); m = img.copy()
        fh(C.c_void_p(a.ctypes.data + (8 * 16 + 4) * 2), 16, b, l, t, bd)
        for i in range(4):
            O.lib().orc_lf_filter_line(C.c_void_p(m.ctypes.data + (8 * 16 + 4 + i) * 2), 16, size, lvl, sharp, bd)
        assert np.array_equal(a, m), ("h", size, lvl, sharp, kind)
