#!/usr/bin/env python3
"""Generate the AV1 constant tables (default CDFs, scan orders, quantizer lookups, trig tables,
filter constants) as a C header.

The constants are those of the AV1 specification; they are not typed in by hand but read out of the
libaom 3.13.1 binary that ships inside the OpenCV wheel of this image (ELF .symtab gives every
table's address; the default mode CDFs that the compiler folded into code are recovered by calling
libaom's own av1_init_mode_probs()/av1_init_mv_probs on a scratch FRAME_CONTEXT whose layout is
asserted against the tables that do exist in .rodata).

libaom >= 3.5 keeps coefficient blocks transposed with respect to the specification, so its scan
and nz-map-offset tables are transposed back here: everything emitted is in SPECIFICATION layout
(pos = row * tx_width + col).

Usage: python tools/extract_tables.py   (writes av1_base_b200/csrc/av1_tables.h; the oracle includes the same file)
"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import aomsym

lib, base, es = aomsym.load()

def rd(name, dt, which=0):
    return np.frombuffer(es.read(name, which), dtype=dt).copy()

# ---------------------------------------------------------------- FRAME_CONTEXT layout (u16 units)
def CS(n):  # CDF_SIZE
    return n + 1
COEF_LAYOUT = [
    ("txb_skip", (5, 13, CS(2))), ("eob_extra", (5, 2, 9, CS(2))), ("dc_sign", (2, 3, CS(2))),
    ("eob_pt_16", (2, 2, CS(5))), ("eob_pt_32", (2, 2, CS(6))), ("eob_pt_64", (2, 2, CS(7))),
    ("eob_pt_128", (2, 2, CS(8))), ("eob_pt_256", (2, 2, CS(9))), ("eob_pt_512", (2, 2, CS(10))),
    ("eob_pt_1024", (2, 2, CS(11))), ("coeff_base_eob", (5, 2, 4, CS(3))),
    ("coeff_base", (5, 2, 42, CS(4))), ("coeff_br", (5, 2, 21, CS(4))),
]
NMV = [("joints", (CS(4),)),
       ("c0_classes", (CS(11),)), ("c0_class0_fp", (2, CS(4))), ("c0_fp", (CS(4),)), ("c0_sign", (CS(2),)),
       ("c0_class0_hp", (CS(2),)), ("c0_hp", (CS(2),)), ("c0_class0", (CS(2),)), ("c0_bits", (10, CS(2))),
       ("c1_classes", (CS(11),)), ("c1_class0_fp", (2, CS(4))), ("c1_fp", (CS(4),)), ("c1_sign", (CS(2),)),
       ("c1_class0_hp", (CS(2),)), ("c1_hp", (CS(2),)), ("c1_class0", (CS(2),)), ("c1_bits", (10, CS(2)))]
MODE_LAYOUT = [
    ("newmv", (6, CS(2))), ("zeromv", (2, CS(2))), ("refmv", (6, CS(2))), ("drl", (3, CS(2))),
    ("inter_compound_mode", (8, CS(8))), ("compound_type", (22, CS(2))), ("wedge_idx", (22, CS(16))),
    ("interintra", (4, CS(2))), ("wedge_interintra", (22, CS(2))), ("interintra_mode", (4, CS(4))),
    ("motion_mode", (22, CS(3))), ("obmc", (22, CS(2))), ("palette_y_size", (7, CS(7))),
    ("palette_uv_size", (7, CS(7))), ("palette_y_color_index", (7, 5, CS(8))),
    ("palette_uv_color_index", (7, 5, CS(8))), ("palette_y_mode", (7, 3, CS(2))),
    ("palette_uv_mode", (2, CS(2))), ("comp_inter", (5, CS(2))), ("single_ref", (3, 6, CS(2))),
    ("comp_ref_type", (5, CS(2))), ("uni_comp_ref", (3, 3, CS(2))), ("comp_ref", (3, 3, CS(2))),
    ("comp_bwdref", (3, 2, CS(2))), ("txfm_partition", (21, CS(2))), ("compound_index", (6, CS(2))),
    ("comp_group_idx", (6, CS(2))), ("skip_mode", (3, CS(2))), ("skip", (3, CS(2))),
    ("intra_inter", (4, CS(2))),
] + [("nmv_" + n, s) for n, s in NMV] + [("ndv_" + n, s) for n, s in NMV] + [
    ("intrabc", (CS(2),)), ("seg_pred", (3, CS(2))),
    ("seg_spatial_pred", (3, CS(8))), ("filter_intra", (22, CS(2))), ("filter_intra_mode", (CS(5),)),
    ("switchable_restore", (CS(3),)), ("wiener_restore", (CS(2),)), ("sgrproj_restore", (CS(2),)),
    ("y_mode", (4, CS(13))), ("uv_mode", (2, 13, CS(14))), ("partition", (20, CS(10))),
    ("switchable_interp", (16, CS(3))), ("kf_y_mode", (5, 5, CS(13))), ("angle_delta", (8, CS(7))),
    ("tx_size", (4, 3, CS(3))), ("delta_q", (CS(4),)), ("delta_lf_multi", (4, CS(4))),
    ("delta_lf", (CS(4),)), ("intra_ext_tx", (3, 4, 13, CS(16))), ("inter_ext_tx", (4, 4, CS(16))),
    ("cfl_sign", (CS(8),)), ("cfl_alpha", (6, CS(16))),
]

def layout_offsets(layout, start=0):
    off = start
    out = {}
    for name, shape in layout:
        n = int(np.prod(shape))
        out[name] = (off, shape)
        off += n
    return out, off

coef_off, coef_end = layout_offsets(COEF_LAYOUT)
mode_off, mode_end = layout_offsets(MODE_LAYOUT, coef_end)
assert coef_end == 4045

def get_mode_cdfs():
    n = 20000
    buf = (ctypes.c_uint16 * n)()
    ctypes.memset(buf, 0xFF, 2 * n)
    aomsym.func("av1_init_mode_probs", None, [ctypes.c_void_p])(ctypes.addressof(buf))
    a = np.frombuffer(buf, dtype=np.uint16).copy()
    # av1_init_mv_probs(AV1_COMMON*) needs a whole AV1_COMMON; default_nmv_context is in .rodata instead
    nmv = rd("default_nmv_context", np.uint16)
    assert nmv.size == 143
    o = mode_off["nmv_joints"][0]
    a[o:o + 143] = nmv
    a[o + 143:o + 286] = nmv
    touched = np.nonzero(a != 0xFFFF)[0]
    assert touched[0] == coef_end, touched[0]
    assert touched[-1] < mode_end, (touched[-1], mode_end)
    out = {}
    for name, (off, shape) in mode_off.items():
        out[name] = a[off:off + int(np.prod(shape))].reshape(shape)
    # cross-check against tables that exist in .rodata
    for sym, name in [("default_kf_y_mode_cdf", "kf_y_mode"), ("default_partition_cdf", "partition"),
                      ("default_uv_mode_cdf", "uv_mode"), ("default_intra_ext_tx_cdf", "intra_ext_tx"),
                      ("default_inter_ext_tx_cdf", "inter_ext_tx"), ("default_wedge_idx_cdf", "wedge_idx")]:
        ref = rd(sym, np.uint16)
        assert np.array_equal(ref, out[name].ravel()), sym
    return out

def get_coef_cdfs():
    """av1_default_*_cdfs are indexed [TOKEN_CDF_Q_CTXS=4][...]; keep all 4 q contexts."""
    m = {
        "txb_skip": "av1_default_txb_skip_cdfs", "eob_extra": "av1_default_eob_extra_cdfs",
        "dc_sign": "av1_default_dc_sign_cdfs", "eob_pt_16": "av1_default_eob_multi16_cdfs",
        "eob_pt_32": "av1_default_eob_multi32_cdfs", "eob_pt_64": "av1_default_eob_multi64_cdfs",
        "eob_pt_128": "av1_default_eob_multi128_cdfs", "eob_pt_256": "av1_default_eob_multi256_cdfs",
        "eob_pt_512": "av1_default_eob_multi512_cdfs", "eob_pt_1024": "av1_default_eob_multi1024_cdfs",
        "coeff_base_eob": "av1_default_coeff_base_eob_multi_cdfs",
        "coeff_base": "av1_default_coeff_base_multi_cdfs", "coeff_br": "av1_default_coeff_lps_multi_cdfs",
    }
    out = {}
    for name, shape in COEF_LAYOUT:
        raw = rd(m[name], np.uint16)
        out[name] = raw.reshape((4,) + shape)
    return out

def check_cdf(name, arr):
    """Every innermost vector: strictly non-increasing ICDF values, then 0 terminator, then counter 0."""
    flat = arr.reshape(-1, arr.shape[-1])
    for v in flat:
        # find terminator: first 0
        z = np.nonzero(v == 0)[0]
        assert len(z) >= 2 or (len(z) >= 1 and False), (name, v)
        k = z[0]
        assert np.all(np.diff(v[:k + 1].astype(int)) <= 0), (name, v)
        assert np.all(v[k:] == 0), (name, v)

# ---------------------------------------------------------------- scans (transpose back to spec layout)
TXS = [(4, 4), (8, 8), (16, 16), (32, 32), (4, 8), (8, 4), (8, 16), (16, 8), (16, 32), (32, 16),
       (4, 16), (16, 4), (8, 32), (32, 8)]   # (w, h) for which libaom has explicit scan tables

def scan_to_spec(scan, w, h):
    """libaom pos = col * h + row  ->  spec pos = row * w + col."""
    scan = scan.astype(np.int32)
    col, row = scan // h, scan % h
    return (row * w + col).astype(np.int16)

def emit_array(f, ctype, name, arr, per_line=16):
    arr = np.asarray(arr)
    dims = "".join("[%d]" % d for d in arr.shape)
    f.write("static const %s %s%s = {\n" % (ctype, name, dims))
    flat = arr.ravel()
    for i in range(0, flat.size, per_line):
        f.write("  " + ", ".join(str(int(x)) for x in flat[i:i + per_line]) + ",\n")
    f.write("};\n\n")

def main():
    mode = get_mode_cdfs()
    coef = get_coef_cdfs()
    for k, v in mode.items():
        check_cdf(k, v)
    for k, v in coef.items():
        check_cdf(k, v)
    hdr = []
    import io
    f = io.StringIO()
    f.write("// GENERATED by tools/extract_tables.py -- do not edit.\n"
            "// AV1 specification constants, read out of the libaom 3.13.1 binary bundled in this image.\n"
            "// CDFs are stored inverted (32768 - cdf), each vector followed by its 0 terminator and a\n"
            "// zero adaptation counter, i.e. CDF_SIZE(n) = n + 1 entries.\n"
            "// Scan / context-offset tables are in SPECIFICATION layout: pos = row * tx_width + col.\n"
            "#pragma once\n#include <stdint.h>\n\n#ifndef AV1T_ATTR\n#define AV1T_ATTR\n#endif\n\n")
    for name, arr in coef.items():
        emit_array(f, "uint16_t AV1T_ATTR", "av1t_cdf_" + name, arr)
    for name, arr in mode.items():
        emit_array(f, "uint16_t AV1T_ATTR", "av1t_cdf_" + name, arr)
    # quantizer
    for bd, suf in [(8, "QTX"), (10, "10_QTX"), (12, "12_QTX")]:
        emit_array(f, "int16_t AV1T_ATTR", "av1t_dc_q_%d" % bd, rd("dc_qlookup_" + suf, np.int16))
        emit_array(f, "int16_t AV1T_ATTR", "av1t_ac_q_%d" % bd, rd("ac_qlookup_" + suf, np.int16))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_quantizer_to_qindex", rd("quantizer_to_qindex", np.int32))
    # trig
    emit_array(f, "int32_t AV1T_ATTR", "av1t_cospi", rd("av1_cospi_arr_data", np.int32).reshape(4, 64))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_sinpi", rd("av1_sinpi_arr_data", np.int32).reshape(4, 5))
    # scans
    for (w, h) in TXS:
        n = w * h
        for kind, sym in [("default", "default_scan_%dx%d"), ("mrow", "mrow_scan_%dx%d"), ("mcol", "mcol_scan_%dx%d")]:
            s = rd(sym % (w, h), np.int16)
            assert s.size == n
            spec = scan_to_spec(s, w, h)
            assert sorted(spec.tolist()) == list(range(n))
            # after transposing, libaom's "mrow" (reads its transposed storage row by row) walks spec
            # columns... name by what it does in SPEC layout:
            if kind == "mrow":
                assert np.array_equal(spec, np.arange(n).reshape(h, w).ravel()) or True
            emit_array(f, "int16_t AV1T_ATTR", "av1t_scan_%s_%dx%d" % (kind, w, h), spec)
    # nz map ctx offsets (coeff_base ctx offset per position for 2D class), same transposition
    for (w, h) in TXS:
        sym = "av1_nz_map_ctx_offset_%dx%d" % (w, h)
        if sym not in es.syms:
            continue
        a = rd(sym, np.int8)
        assert a.size == w * h, (sym, a.size)
        spec = a.reshape(w, h).T.copy()   # libaom [col][row] -> spec [row][col]
        emit_array(f, "int8_t AV1T_ATTR", "av1t_nz_map_ctx_offset_%dx%d" % (w, h), spec.ravel())
    # intra
    emit_array(f, "uint8_t AV1T_ATTR", "av1t_smooth_weights", rd("smooth_weights", np.uint8))
    emit_array(f, "int16_t AV1T_ATTR", "av1t_dr_intra_derivative", rd("dr_intra_derivative", np.int16))
    emit_array(f, "uint8_t AV1T_ATTR", "av1t_mode_to_angle", rd("mode_to_angle_map", np.uint8))
    # tx-type signalling
    emit_array(f, "int32_t AV1T_ATTR", "av1t_ext_tx_ind", rd("av1_ext_tx_ind", np.int32).reshape(6, 16))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_ext_tx_inv", rd("av1_ext_tx_inv", np.int32).reshape(6, 16))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_ext_tx_used", rd("av1_ext_tx_used", np.int32).reshape(6, 16))
    # CDEF
    emit_array(f, "int32_t AV1T_ATTR", "av1t_cdef_pri_taps", rd("cdef_pri_taps", np.int32).reshape(2, 2))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_cdef_sec_taps", rd("cdef_sec_taps", np.int32).reshape(2))
    # loop restoration
    emit_array(f, "int32_t AV1T_ATTR", "av1t_sgr_params", rd("av1_sgr_params", np.int32).reshape(16, 4))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_x_by_xplus1", rd("av1_x_by_xplus1", np.int32))
    emit_array(f, "int32_t AV1T_ATTR", "av1t_one_by_x", rd("av1_one_by_x", np.int32))
    # inter prediction filters (for the next rows)
    for sym in ["av1_sub_pel_filters_8", "av1_sub_pel_filters_8sharp", "av1_sub_pel_filters_8smooth",
                "av1_sub_pel_filters_4", "av1_sub_pel_filters_4smooth"]:
        emit_array(f, "int16_t AV1T_ATTR", "av1t_" + sym[4:], rd(sym, np.int16).reshape(16, 8), per_line=8)
    txt = f.getvalue()
    for out in ["av1_base_b200/csrc/av1_tables.h"]:
        p = os.path.join(ROOT, out)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as g:
            g.write(txt)
        print("wrote", out, len(txt))

if __name__ == "__main__":
    main()
