#!/usr/bin/env python3
"""CPU-only rate/quality sweep of the ORACLE's statement of the whole decision chain (oracle/chain.py; test infrastructure):
the device path matches the chain bit for bit (tests/test_gpu_inter_parity.py), so the curve measured here is the product's
curve.  Compares with the libaom points of the same clip stored in a bench line (profiles/r02a_bench_4k10.json `bd_rate`).
Usage: tools/rd_chain.py [--crfs 28,36,40,44,52] [--frames 30] [--noise 1.0] [--gop 0] [--ref profiles/r02a_bench_4k10.json]"""
import argparse, json, os, sys
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def encode(args):
    w, h, bd, nfr, seed, noise, crf, gop, kw = args
    from av1_base_b200 import abi, packer, synth
    from oracle import pyoracle as O, decoders as D, chain
    frames = synth.synth_clip(w, h, bd, nfr, seed=seed, scene_len=1000, noise=noise)
    g, res = chain.encode_chain(frames, w, h, bd, crf, gop_period=gop, **kw)
    seq = abi.SeqParams(w, h, bd, 1, 0, 30, 1, 0)
    tot, ps, byk = 0, [], {0: 0, 1: 0, 2: 0}
    for i, r in enumerate(res):
        sy = packer.make_syms(g, r.res.blocks, r.res.coef, cdef_idx=r.cdef_idx)
        fp = r.fp
        fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        tu = packer.pack_frame(seq, fp, sy, with_td=False)
        ps.append(D.psnr(O.crop(g, r.fin)[0], frames[i][0], bd))
        tot += len(tu) + 2
        byk[r.kind] += len(tu)
    return dict(crf=crf, kbps=tot * 8 * 30.0 / nfr / 1000, psnr_y=float(np.mean(ps)), bytes=tot, by_kind=byk)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="960x544")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--noise", type=float, default=1.0)
    ap.add_argument("--crfs", default="20,30,36,40,44,48,52,58")
    ap.add_argument("--gop", type=int, default=0)
    ap.add_argument("--ref", default=os.path.join(ROOT, "profiles", "r02a_bench_4k10.json"))
    ap.add_argument("--kw", default="{}", help="JSON of extra encode_chain keyword arguments (experiments)")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    kw = json.loads(a.kw)
    jobs = [(w, h, a.bd, a.frames, a.seed, a.noise, crf, a.gop, kw) for crf in map(int, a.crfs.split(","))]
    with ProcessPoolExecutor(min(os.cpu_count() or 1, len(jobs))) as ex:
        res = list(ex.map(encode, jobs))
    for r in res:
        print(json.dumps(r))
    from bdrate import bd_rate, bd_rate_pchip
    if a.ref and os.path.exists(a.ref) and a.noise == 1.0 and a.frames == 30 and (w, h) == (960, 544):
        d = json.load(open(a.ref))["bd_rate"]
        for name in ("libaom_cpu6", "libaom_cpu6_lag0"):
            pts = d[name]
            v = bd_rate([x["kbps"] for x in pts], [x["psnr_y"] for x in pts], [x["kbps"] for x in res], [x["psnr_y"] for x in res])
            vp = bd_rate_pchip([x["kbps"] for x in pts], [x["psnr_y"] for x in pts], [x["kbps"] for x in res], [x["psnr_y"] for x in res])
            print("BD-rate (PSNR-Y) vs %s: cubic %s, pchip %s" % (name, "%.1f %%" % v if v is not None else "no overlap", "%.1f %%" % vp if vp is not None else "no overlap"))


if __name__ == "__main__":
    main()
