#!/usr/bin/env python3
"""Scratch bench for encoder-side ideas on the CPU oracle chain (test infrastructure, not product): numpy
prototypes of pre-filters / structures are tried here before they are restated in oracle/av1_oracle.cpp and in CUDA.
Usage: tools/rd_lab.py [--noise 1.0] [--hier P,A,O] [--mctf R,S] ..."""
import argparse, json, os, sys
import ctypes as C
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from rd_oracle import table


def mc_bilinear(ref, mv, g, ss):
    """whole-plane motion compensation with per-8x8-unit vectors (1/8 luma samples), bilinear; numpy prototype"""
    h, w = (g.height >> ss), (g.width >> ss)
    u = 8 >> ss
    mvr = np.repeat(np.repeat(mv[:, 0].reshape(g.h8, g.w8), u, 0), u, 1)[:h, :w].astype(np.int32)
    mvc = np.repeat(np.repeat(mv[:, 1].reshape(g.h8, g.w8), u, 0), u, 1)[:h, :w].astype(np.int32)
    yy, xx = np.mgrid[0:h, 0:w]
    y16 = yy * 16 + ((2 * mvr) >> ss); x16 = xx * 16 + ((2 * mvc) >> ss)
    iy, ix, fy, fx = y16 >> 4, x16 >> 4, y16 & 15, x16 & 15
    r = ref[:h, :w].astype(np.int64)
    def px(y, x):
        return r[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)]
    a = px(iy, ix) * (16 - fx) + px(iy, ix + 1) * fx
    b = px(iy + 1, ix) * (16 - fx) + px(iy + 1, ix + 1) * fx
    return (a * (16 - fy) + b * fy + 128) >> 8


def mctf(O, g, frames, pyrs, i, radius, strength, bd, lam):
    """temporal filter of frame i over its neighbours: per 16x16 block weight from the block's mean squared error"""
    cur = frames[i]
    num = [p.astype(np.int64) * 256 for p in cur]
    den = [np.full(p.shape, 256, np.int64) for p in cur]
    for j in range(max(0, i - radius), min(len(frames), i + radius + 1)):
        if j == i:
            continue
        mv = O.hme(g, pyrs[i], pyrs[j], lam, bd)
        for p in range(3):
            ss = 1 if p else 0
            pred = mc_bilinear(np.asarray(frames[j][p]), mv, g, ss)
            c = cur[p].astype(np.int64)
            d2 = (pred - c) ** 2
            bs = 16 >> ss
            H, W = d2.shape
            Hp, Wp = (H + bs - 1) // bs * bs, (W + bs - 1) // bs * bs
            pad = np.zeros((Hp, Wp)); pad[:H, :W] = d2
            cnt = np.zeros((Hp, Wp)); cnt[:H, :W] = 1
            mse = pad.reshape(Hp // bs, bs, Wp // bs, bs).sum((1, 3)) / np.maximum(1, cnt.reshape(Hp // bs, bs, Wp // bs, bs).sum((1, 3)))
            scale = (1 << (bd - 8)) ** 2
            wb = np.exp(-mse / (strength * scale))            # block weight
            wpx = np.exp(-d2 / (3.0 * strength * scale))       # pixel weight
            wgt = (np.repeat(np.repeat(wb, bs, 0), bs, 1)[:H, :W] * wpx * 256).astype(np.int64)
            num[p] += wgt * pred
            den[p] += wgt
    return [((n + d // 2) // d).astype(np.uint16) for n, d in zip(num, den)]


def encode(args):
    w, h, bd, nfr, seed, noise, crf, opts = args
    from av1_base_b200 import abi, packer, synth
    from oracle import pyoracle as O
    from oracle import decoders as D
    frames = synth.synth_clip(w, h, bd, nfr, seed=seed, scene_len=1000, noise=noise)
    qidx = max(1, table("av1t_quantizer_to_qindex")[crf])
    qkey = max(1, min(255, qidx + opts.get("dkey", -(qidx // 4))))
    acq = table("av1t_ac_q_%d" % bd)[qidx]
    g = O.geom(w, h, 0, 0)
    pm = O.partition_fixed(g, 4)
    seq = abi.SeqParams(w, h, bd, 1, 0, 30, 1, 0)
    hier = opts.get("hier")
    pyrs = [O.pyramid(g, O.pad_planes(g, fr)[0]) for fr in frames]
    enc_src = list(frames)
    if opts.get("mctf"):
        rad, strength = opts["mctf"]
        for i in range(nfr):
            if i == 0 or (hier and i % hier[0] == 0) or not hier:
                enc_src[i] = mctf(O, g, frames, pyrs, i, int(rad), strength, bd, acq >> 1)
    if opts.get("mctf2"):
        rad, kb, kp = opts["mctf2"]
        thr_b = max(1, int(kb * acq * acq / 256)); thr_p = max(1, int(kp * thr_b))
        padded = [O.pad_planes(g, fr) for fr in frames]
        for i in range(nfr):
            if i == 0 or (hier and i % hier[0] == 0) or not hier:
                lo, hi = (0, int(opts.get("keyfwd", 6))) if i == 0 else (-int(rad), int(rad))
                nb = [j for j in range(i + lo, i + hi + 1) if 0 <= j < nfr and j != i]
                mvs = []
                for j in nb:
                    mv = O.hme(g, pyrs[i], pyrs[j], acq >> 1, bd)
                    if opts.get("smooth") and not opts.get("tf_nosmooth"):
                        mv = O.me_sbrd(g, pyrs[i], pyrs[j], mv, int((acq >> 1) * opts["smooth"][0]), int((acq >> 3) * opts["smooth"][0]), int(opts["smooth"][1]))
                    mvs.append(mv)
                enc_src[i] = O.crop(g, O.mctf(g, bd, padded[i], [padded[j] for j in nb], mvs, thr_b, thr_p))
    nbytes, psnr = 0, []
    per_frame = []
    anchor_fin = anchor_pyr = prev_fin = prev_pyr = None
    for i, fr in enumerate(enc_src):
        src = O.pad_planes(g, fr)
        pyr = O.pyramid(g, src[0]) if opts.get("me_filtered") else pyrs[i]
        if hier and i > 0:
            prev_fin, prev_pyr = anchor_fin, anchor_pyr
        fp = abi.FrameParams()
        if i == 0:
            abi.lib().av1b_select_frame_params(bd, qkey, 0, 1, C.byref(fp))
            pmk = pm
            if opts.get("varpart"):
                acqk = table("av1t_ac_q_%d" % bd)[qkey]
                pmk = O.partition_smooth(g, src[0], min(4 * acqk, 800 << (bd - 8)))
            r = O.encode_intra_frame(g, fr, bd, qkey, pmk)
        else:
            mv = O.hme(g, pyr, prev_pyr, acq >> 1, bd)
            if opts.get("smooth"):
                k, it = opts["smooth"]
                mv = O.me_sbrd(g, pyr, prev_pyr, mv, int((acq >> 1) * k), int((acq >> 3) * k), int(it))
            qf = qidx
            if hier:
                qf = max(1, min(255, qidx + (hier[1] if i % hier[0] == 0 else hier[2])))
            abi.lib().av1b_select_frame_params(bd, qf, 1, 1, C.byref(fp))
            fp.non_reference = 1 if (hier and i % hier[0] != 0) else 0
            r = O.encode_inter_frame(g, fr, bd, qf, pm, mv, prev_fin, quant_rnd=opts.get("rnd", 48), tb_zero_thr=opts.get("thr", 0))
            O.merge_skip_blocks(g, r.blocks)
        fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        O.deblock_frame(g, bd, r.blocks, r.rec, list(fp.lf_level), fp.lf_sharpness)
        idx = O.cdef_search(g, bd, r.blocks, fp, r.rec, src)
        fin = O.cdef_frame(g, bd, r.blocks, fp, idx, r.rec)
        sy = packer.make_syms(g, r.blocks, r.coef, cdef_idx=idx)
        tu = b"\x12\x00" + (packer.pack_sequence_header(seq) if i == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False)
        nbytes += len(tu)
        per_frame.append(len(tu))
        psnr.append(D.psnr(O.crop(g, fin)[0], frames[i][0], bd))
        prev_fin, prev_pyr = fin, pyr
        if i == 0 or (hier and i % hier[0] == 0):
            anchor_fin, anchor_pyr = fin, pyr
    return dict(crf=crf, kbps=nbytes * 8 * 30.0 / nfr / 1000, psnr_y=float(np.mean(psnr)), key_bytes=per_frame[0],
                frame_bytes=per_frame[1:9], psnr_first=[round(x, 2) for x in psnr[:9]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="960x544")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--noise", type=float, default=1.0)
    ap.add_argument("--crfs", default="20,28,36,44,52")
    ap.add_argument("--hier", default="")
    ap.add_argument("--mctf", default="", help="radius,strength")
    ap.add_argument("--dkey", type=int, default=None)
    ap.add_argument("--mctf2", default="", help="radius,kb,kp: oracle temporal filter, thr_b = kb acq^2 / 256, thr_p = kp thr_b")
    ap.add_argument("--keyfwd", type=int, default=6)
    ap.add_argument("--varpart", action="store_true")
    ap.add_argument("--tf-nosmooth", action="store_true")
    ap.add_argument("--me-filtered", action="store_true")
    ap.add_argument("--smooth", default="", help="k,iters: vector-field regularisation with lam_s = k * lambda")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    opts = dict(hier=tuple(map(int, a.hier.split(','))) if a.hier else None,
                mctf=tuple(map(float, a.mctf.split(','))) if a.mctf else None, me_filtered=a.me_filtered,
                smooth=tuple(map(float, a.smooth.split(','))) if a.smooth else None,
                mctf2=tuple(map(float, a.mctf2.split(','))) if a.mctf2 else None, keyfwd=a.keyfwd, varpart=a.varpart, tf_nosmooth=a.tf_nosmooth)
    if a.dkey is not None:
        opts["dkey"] = a.dkey
    jobs = [(w, h, a.bd, a.frames, a.seed, a.noise, crf, opts) for crf in map(int, a.crfs.split(","))]
    with ProcessPoolExecutor(min(8, len(jobs))) as ex:
        res = list(ex.map(encode, jobs))
    for r in res:
        print(json.dumps(r))
    from bdrate import bd_rate
    refs = {1.0: ["profiles/r01g_bdrate_noise1.json", "profiles/r01l_bdrate_vs_libaom_lag0.json"],
            0.25: ["profiles/r01g_bdrate_noise025.json"]}.get(a.noise, [])
    for f in refs:
        d = json.load(open(os.path.join(ROOT, f)))
        for name, pts in (("libaom", d["libaom_cpu6"]), ("ours_r01", d["ours"])):
            bdv = bd_rate([x["kbps"] for x in pts], [x["psnr_y"] for x in pts], [x["kbps"] for x in res], [x["psnr_y"] for x in res])
            print("BD-rate (PSNR-Y) vs %s of %s: %s" % (name, os.path.basename(f), "%.1f %%" % bdv if bdv is not None else "no overlap"))


if __name__ == "__main__":
    main()
