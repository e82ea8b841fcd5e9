#!/usr/bin/env python3
"""Per-kernel timing through the kernel-suite C ABI on realistic 4K / 1080p data (oracle-free: the inputs
are produced by the CUDA encoder itself).  Usage: tools/kbench.py [--size 4k|1080p] [--only NAME] [--reps N]
Prints one JSON line per kernel: ms per launch, algorithmic bytes, achieved GB/s, fraction of the measured HBM peak."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from av1_base_b200 import abi, encoder, kernels, synth
import ctypes as C

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="4k")
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--frames", type=int, default=4)
a = ap.parse_args()
w, h = (3840, 2160) if a.size == "4k" else (1920, 1080)
bd, n = 10, a.frames
peak = 6447.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
frames = synth.synth_clip(w, h, bd, n, seed=4, scene_len=100, hdr=(a.size == "4k"))
# realistic reconstruction + side info: encode with the in-loop filters off and keep everything
enc = encoder.Encoder(w, h, bd, crf=30, keep_debug=True, frames_in_flight=n, loop_filters=False)
enc.encode_chunk(frames)
g = enc.geom
recs, blocks = [], []
for i in range(n):
    b, _ = enc.frame_syms(i)
    blocks.append(b)
    r = enc.recon(i)
    pad = [np.zeros((g.rows[p], g.stride[p]), np.uint16) for p in range(3)]
    for p in range(3):
        pad[p][:r[p].shape[0], :r[p].shape[1]] = r[p]
    recs.append(pad)
enc.close()
blocks = np.stack(blocks)
srcs = []
for fr in frames:
    pad = [np.zeros((g.rows[p], g.stride[p]), np.uint16) for p in range(3)]
    for p in range(3):
        pad[p][:fr[p].shape[0], :fr[p].shape[1]] = fr[p]
    srcs.append(pad)
S = int(1.5 * w * h * 2)
Y = w * h * 2
fp = abi.FrameParams()
abi.lib().av1b_select_frame_params(bd, 120, 1, 1, C.byref(fp))


def report(name, ms, bytes_per_launch):
    gbs = bytes_per_launch / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "size": "%dx%d" % (w, h), "frames_per_launch": n, "ms_per_launch": round(ms, 4),
                      "algorithmic_bytes": bytes_per_launch, "achieved_gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4)}), flush=True)


want = lambda k: (not a.only) or a.only in k
deb = None
if want("deblock") or want("cdef") or want("lr"):
    deb, ms = kernels.deblock(w, h, bd, blocks, recs, list(fp.lf_level), 0, reps=a.reps)
    if want("deblock"):
        report("deblock_kernel", ms, 2 * S * n)
if want("cdef") or want("lr"):
    cd, idx, ms = kernels.cdef(w, h, bd, blocks, fp, deb, src=srcs, reps=a.reps)
    if want("cdef"):
        report("cdef_kernel (8-preset decision + filter)", ms, 3 * S * n)
        _, _, ms2 = kernels.cdef(w, h, bd, blocks, fp, deb, forced_idx=idx, reps=a.reps)
        report("cdef_kernel (filter only)", ms2, 2 * S * n)
if want("lr"):
    for lt, nm in ((1, "wiener"), (2, "self-guided")):
        fp2 = abi.FrameParams()
        C.memmove(C.byref(fp2), C.byref(fp), C.sizeof(fp))
        units = []
        for p in range(3):
            fp2.lr_type[p] = lt
        fp2.lr_unit_shift, fp2.lr_uv_shift = 0, 0
        for p in range(3):
            us = 64
            ph, pw = (h, w) if p == 0 else (h // 2, w // 2)
            ur, uc = max((ph + 32) // 64, 1), max((pw + 32) // 64, 1)
            u = np.zeros((n, ur, uc), abi.LR_UNIT_DTYPE)
            u["type"] = lt
            u["wiener_v"] = [3, -7, 15]; u["wiener_h"] = [3, -7, 15]
            if p:
                u["wiener_v"][..., 0] = 0; u["wiener_h"][..., 0] = 0
            u["sgr_set"] = 4; u["sgr_xqd"] = [-32, 31]
            units.append(u)
        _, ms = kernels.loop_restoration(w, h, bd, fp2, cd, deb, units, reps=a.reps)
        report("lr_kernel (%s)" % nm, ms, int(2.125 * S) * n)
if want("pyramid") or want("hme"):
    l0 = np.stack([s[0] for s in srcs])
    if want("pyramid"):
        _, _, ms = kernels.pyramid(w, h, l0, reps=a.reps)
        report("pyramid_kernel", ms, int(1.3125 * Y) * n)
    if want("hme"):
        _, ms = kernels.hme(w, h, l0[1:], l0[:-1], lam=280, reps=a.reps, bd=bd)
        report("hme_l2_kernel + hme_refine_kernel", ms, int(2.625 * Y) * (n - 1))
if want("inter"):
    mv, _ = kernels.hme(w, h, np.stack([srcs[1][0]]), np.stack([srcs[0][0]]), lam=280, bd=bd)
    pm = np.full(g.h8 * g.w8, 4, np.uint8)
    pmm = pm.reshape(g.h8, g.w8)
    if g.h8 & 1:
        pmm[-1, :] = 3
    if g.w8 & 1:
        pmm[:, -1] = 3
    _, _, _, ms = kernels.inter_encode(w, h, bd, 120, pm, mv[0], srcs[1], recs[0], reps=a.reps)
    report("inter_encode_kernel", ms, 4 * S)
if want("txfm"):
    rng = np.random.default_rng(0)
    for N in (8, 16, 32, 64):
        nb = (w * h) // (N * N)
        cn = min(N, 32)
        co = rng.integers(-2000, 2000, (nb, cn, cn)).astype(np.int32)
        pred = rng.integers(0, 1024, (nb, N, N)).astype(np.uint16)
        _, ms = kernels.inv_txfm_add(co, pred, N, N, 0, bd, reps=a.reps)
        report("inv_txfm_add_kernel<%d,%d> (incl. device copy of the prediction)" % (N, N), ms, nb * (4 * cn * cn + 4 * N * N))
