#!/usr/bin/env python3
"""CPU-only rate/quality sweep of the ORACLE restatement of the encode path (test infrastructure): used to
tune encoder-side decisions (quantiser rounding, transform-block drop threshold, ...) without a GPU.
The device path must then match the oracle bit for bit, so the curve measured here is the product's curve.
Compares with stored libaom points of the same clip (profiles/r01g_bdrate_*.json, profiles/r01l_*.json).
Usage: tools/rd_oracle.py [--noise 1.0] [--frames 30] [--crfs 20,28,36,44,52] [--rnd R] [--thr T] [--lr] [--qmod P,L,H] [--hier P,A,O]"""
import argparse, json, os, re, sys
import ctypes as C
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def table(name):
    txt = open(os.path.join(ROOT, "av1_base_b200", "csrc", "av1_tables.h")).read()
    m = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, txt, re.S)
    return [int(v) for v in re.findall(r"-?\d+", m.group(1))]


def encode(args):
    w, h, bd, nfr, seed, noise, crf, opts = args
    from av1_base_b200 import abi, packer, synth
    from oracle import pyoracle as O
    from oracle import decoders as D
    frames = synth.synth_clip(w, h, bd, nfr, seed=seed, scene_len=1000, noise=noise)
    qidx = max(1, table("av1t_quantizer_to_qindex")[crf])
    qkey = max(1, qidx * 3 // 4)
    acq = table("av1t_ac_q_%d" % bd)[qidx]
    g = O.geom(w, h, 0, 0)
    pm = O.partition_fixed(g, 4)
    seq = abi.SeqParams(w, h, bd, 1, 1 if opts.get("lr") else 0, 30, 1, 0)
    nlr = np.zeros(3, np.int64)
    fps = []
    for ft in (0, 1):
        fp = abi.FrameParams()
        abi.lib().av1b_select_frame_params(bd, qidx if ft else qkey, ft, 1, C.byref(fp))
        fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        fps.append(fp)
    prev_fin = prev_pyr = None
    nbytes, psnr, nskip = 0, [], 0
    tus, fins = [], []
    hier = opts.get("hier")
    anchor_fin = anchor_pyr = None
    for i, fr in enumerate(frames):
        src = O.pad_planes(g, fr)
        pyr = O.pyramid(g, src[0])
        if hier and i > 0:
            per, da, dn = hier
            is_anchor = i % per == 0
            prev_fin, prev_pyr = anchor_fin, anchor_pyr
        if i == 0:
            r, fp = O.encode_intra_frame(g, fr, bd, qkey, pm), fps[0]
        else:
            mv = O.hme(g, pyr, prev_pyr, acq >> 1, bd)
            qf = qidx
            if hier:
                qf = max(1, min(255, qidx + (hier[1] if i % hier[0] == 0 else hier[2])))
            if opts.get("qmod"):   # experiment: periodic quantiser modulation inside the P chain
                per, lo, hi = opts["qmod"]
                qf = max(1, min(255, qidx + (lo if i % per == 0 else hi)))
            fp = abi.FrameParams()
            abi.lib().av1b_select_frame_params(bd, qf, 1, 1, C.byref(fp))
            fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
            fp.non_reference = 1 if (hier and i % hier[0] != 0) else 0
            r = O.encode_inter_frame(g, fr, bd, qf, pm, mv, prev_fin, quant_rnd=opts.get("rnd", 48),
                                     tb_zero_thr=opts.get("thr", 0))
            O.merge_skip_blocks(g, r.blocks)
            nskip += int(np.count_nonzero(r.blocks["skip"]))
        O.deblock_frame(g, bd, r.blocks, r.rec, list(fp.lf_level), fp.lf_sharpness)
        idx = O.cdef_search(g, bd, r.blocks, fp, r.rec, src)
        fin = O.cdef_frame(g, bd, r.blocks, fp, idx, r.rec)
        lr_units = None
        if opts.get("lr"):
            fp.lr_type[0], fp.lr_type[1], fp.lr_type[2] = 3, 0, 0
            q_acq = table("av1t_ac_q_%d" % bd)[fp.base_q_idx]
            cand = O.lr_candidate(sgr_set=opts.get("sgr_set", 4), wiener_v=opts["wv"], wiener_h=opts["wv"], sgr_xqd=opts["xqd"])
            units, sse = O.lr_search(g, bd, fp, cand, fin, r.rec, src[0], (q_acq * q_acq * 5) >> 8)
            fin = O.lr_frame(g, bd, fp, fin, r.rec, [units, None, None])
            lr_units = [units, None, None]
            nlr += np.bincount(units["type"].ravel().astype(np.int64), minlength=3)
        sy = packer.make_syms(g, r.blocks, r.coef, cdef_idx=idx, lr_units=lr_units)
        tu = b"\x12\x00" + (packer.pack_sequence_header(seq) if i == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False)
        tus.append(tu)
        nbytes += len(tu)
        psnr.append(D.psnr(O.crop(g, fin)[0], fr[0], bd))
        prev_fin, prev_pyr = fin, pyr
        if i == 0 or (hier and i % hier[0] == 0):
            anchor_fin, anchor_pyr = fin, pyr
        fins.append(fin)
    if opts.get("verify"):
        dec = D.dav1d_decode(tus)
        for i in (0, 1, len(dec) - 1):
            assert np.array_equal(dec[i][0], O.crop(g, fins[i])[0]), "decode != recon (frame %d)" % i
    return dict(crf=crf, kbps=nbytes * 8 * 30.0 / nfr / 1000, psnr_y=float(np.mean(psnr)),
                skip_frac=nskip / max(1, (nfr - 1) * g.h8 * g.w8), lr_types=[int(v) for v in nlr])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="960x544")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--noise", type=float, default=1.0)
    ap.add_argument("--crfs", default="20,28,36,44,52")
    ap.add_argument("--rnd", type=int, default=48)
    ap.add_argument("--thr", type=int, default=0)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--lr", action="store_true", help="loop restoration decision on (preset <= 5)")
    ap.add_argument("--sgr-set", type=int, default=4)
    ap.add_argument("--qmod", default="", help="period,delta_low,delta_high: quantiser index offsets inside the P chain (experiment)")
    ap.add_argument("--hier", default="", help="period,delta_anchor,delta_other: one-level hierarchy -- every period-th frame is an anchor (the only frames "
                                               "that update the reference), the others predict from the last anchor and are not referenced (experiment)")
    ap.add_argument("--wv", default="3,-7,15")
    ap.add_argument("--xqd", default="-32,31")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    opts = dict(rnd=a.rnd, thr=a.thr, verify=a.verify, lr=a.lr, sgr_set=a.sgr_set, wv=tuple(map(int, a.wv.split(','))), xqd=tuple(map(int, a.xqd.split(','))), qmod=tuple(map(int, a.qmod.split(','))) if a.qmod else None, hier=tuple(map(int, a.hier.split(','))) if a.hier else None)
    jobs = [(w, h, a.bd, a.frames, a.seed, a.noise, crf, opts) for crf in map(int, a.crfs.split(","))]
    with ProcessPoolExecutor(min(8, len(jobs))) as ex:
        res = list(ex.map(encode, jobs))
    for r in res:
        print(json.dumps(r))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bdrate import bd_rate
    refs = {1.0: ["profiles/r01g_bdrate_noise1.json", "profiles/r01l_bdrate_vs_libaom_lag0.json"],
            0.25: ["profiles/r01g_bdrate_noise025.json"]}.get(a.noise, [])
    for f in refs:
        d = json.load(open(os.path.join(ROOT, f)))
        if d["clip"]["frames"] != a.frames or (d["clip"]["w"], d["clip"]["h"]) != (w, h):
            continue
        for name, pts in (("libaom", d["libaom_cpu6"]), ("ours_r01", d["ours"])):
            bdv = bd_rate([x["kbps"] for x in pts], [x["psnr_y"] for x in pts], [x["kbps"] for x in res], [x["psnr_y"] for x in res])
            print("BD-rate (PSNR-Y) vs %s of %s: %s" % (name, f, "%.1f %%" % bdv if bdv is not None else "no overlap"))


if __name__ == "__main__":
    main()
