#!/usr/bin/env python3
"""Compression-efficiency check (BASELINE.json north_star, correctness part 3): rate/quality points of the
B200 encoder against libaom 3.13.1 (cpu-used 6, constant quality, all host cores) on the same synthetic
clip, and the Bjontegaard rate difference between the two curves.  The reference's own SVT-AV1 cannot
run in this image, libaom cpu-used=6 is the stand-in SURVEY.md 8d names.
Every B200 stream is decoded with dav1d and compared with the encoder's reconstruction on the way.
Usage (GPU box): tools/bdrate.py [--size 960x544] [--frames 30] [--bd 10] [--out profiles/x.json]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from av1_base_b200 import encoder, synth
from oracle import decoders as D


def ssim_y(a, b, bd):
    from scipy.ndimage import uniform_filter
    a = a.astype(np.float64); b = b.astype(np.float64)
    L = (1 << bd) - 1
    c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    ma, mb = uniform_filter(a, 8), uniform_filter(b, 8)
    va = uniform_filter(a * a, 8) - ma * ma
    vb = uniform_filter(b * b, 8) - mb * mb
    cov = uniform_filter(a * b, 8) - ma * mb
    s = ((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2))
    return float(s.mean())


def quality(frames, dec, bd):
    py = np.mean([D.psnr(d[0], f[0], bd) for d, f in zip(dec, frames)])
    pu = np.mean([D.psnr(d[1], f[1], bd) for d, f in zip(dec, frames)])
    pv = np.mean([D.psnr(d[2], f[2], bd) for d, f in zip(dec, frames)])
    ss = np.mean([ssim_y(d[0], f[0], bd) for d, f in zip(dec, frames)])
    return dict(psnr_y=float(py), psnr_u=float(pu), psnr_v=float(pv), psnr_avg=float((6 * py + pu + pv) / 8), ssim_y=float(ss))


def bd_rate(r1, q1, r2, q2):
    """Bjontegaard delta rate (%) of curve 2 against curve 1 (cubic fit of log-rate over quality)."""
    l1, l2 = np.log(r1), np.log(r2)
    p1, p2 = np.polyfit(q1, l1, 3), np.polyfit(q2, l2, 3)
    lo, hi = max(min(q1), min(q2)), min(max(q1), max(q2))
    if hi <= lo:
        return None
    i1, i2 = np.polyint(p1), np.polyint(p2)
    a1 = (np.polyval(i1, hi) - np.polyval(i1, lo)) / (hi - lo)
    a2 = (np.polyval(i2, hi) - np.polyval(i2, lo)) / (hi - lo)
    return float((np.exp(a2 - a1) - 1) * 100)


def bd_rate_pchip(r1, q1, r2, q2):
    """Bjontegaard delta rate (%) of curve 2 against curve 1 with piecewise cubic Hermite interpolation of log-rate over
    quality (the AOM common-test-conditions form: no overshoot between the measured points, unlike one cubic through all)."""
    from scipy.interpolate import PchipInterpolator
    o1, o2 = np.argsort(q1), np.argsort(q2)
    q1, l1 = np.asarray(q1, float)[o1], np.log(np.asarray(r1, float)[o1])
    q2, l2 = np.asarray(q2, float)[o2], np.log(np.asarray(r2, float)[o2])
    lo, hi = max(q1[0], q2[0]), min(q1[-1], q2[-1])
    if hi <= lo:
        return None
    a1 = PchipInterpolator(q1, l1).integrate(lo, hi) / (hi - lo)
    a2 = PchipInterpolator(q2, l2).integrate(lo, hi) / (hi - lo)
    return float((np.exp(a2 - a1) - 1) * 100)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="960x544")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--crfs", default="20,28,36,44,52")
    ap.add_argument("--cqs", default="24,32,40,48,56")
    ap.add_argument("--out", default="")
    ap.add_argument("--noise", type=float, default=1.0, help="scale of the synthetic sensor noise")
    ap.add_argument("--tb-zero-thr", type=int, default=0)
    ap.add_argument("--skip-libaom", action="store_true")
    ap.add_argument("--lag", type=int, default=19, help="libaom lag_in_frames (0 = low delay, no alt-ref filtering)")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    frames = synth.synth_clip(w, h, a.bd, a.frames, seed=a.seed, scene_len=1000, noise=a.noise)
    fps = 30.0
    res = {"clip": {"w": w, "h": h, "bit_depth": a.bd, "frames": a.frames, "seed": a.seed, "noise": a.noise}, "ours": [], "libaom_cpu6": []}
    for crf in map(int, a.crfs.split(",")):
        enc = encoder.Encoder(w, h, a.bd, crf=crf, keyint=240, keep_debug=True, tb_zero_thr=a.tb_zero_thr)
        t0 = time.perf_counter()
        tus = enc.encode_chunk(frames)
        dt = time.perf_counter() - t0
        dec = D.dav1d_decode(tus)
        for i in range(a.frames):
            rec = enc.recon(i)
            for p in range(3):
                assert np.array_equal(dec[i][p], rec[p]), "decode != reconstruction (crf %d frame %d plane %d)" % (crf, i, p)
        enc.close()
        q = quality(frames, dec, a.bd)
        q.update(crf=crf, kbps=sum(map(len, tus)) * 8 * fps / a.frames / 1000, enc_fps=a.frames / dt, decode_matches_recon=True)
        res["ours"].append(q)
        print(json.dumps({"ours": q}), flush=True)
    cores = os.cpu_count() or 1
    for cq in ([] if a.skip_libaom else list(map(int, a.cqs.split(",")))):
        t0 = time.perf_counter()
        tus = D.aom_encode(frames, a.bd, cq_level=cq, cpu_used=6, threads=cores, lag=a.lag)
        dt = time.perf_counter() - t0
        dec = D.dav1d_decode(tus)
        q = quality(frames, dec, a.bd)
        q.update(cq=cq, kbps=sum(map(len, tus)) * 8 * fps / a.frames / 1000, enc_fps=a.frames / dt, cores=cores)
        res["libaom_cpu6"].append(q)
        print(json.dumps({"libaom_cpu6": q}), flush=True)
    res["libaom_lag_in_frames"] = a.lag
    for m in (() if a.skip_libaom else ("psnr_y", "psnr_avg", "ssim_y")):
        r1 = [x["kbps"] for x in res["libaom_cpu6"]]; q1 = [x[m] for x in res["libaom_cpu6"]]
        r2 = [x["kbps"] for x in res["ours"]]; q2 = [x[m] for x in res["ours"]]
        res["bd_rate_vs_libaom_cpu6_%s_pct" % m] = bd_rate(r1, q1, r2, q2)
    print(json.dumps({k: v for k, v in res.items() if k.startswith("bd_rate")}), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
