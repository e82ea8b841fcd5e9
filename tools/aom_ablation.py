#!/usr/bin/env python3
"""TEST / ANALYSIS TOOL (CPU only, no product code involved): which of libaom's tools carry its rate / quality lead on the clip the
bench's `bd_rate` is measured on (960x544 10-bit, 30 frames, synth seed 4, noise 1.0)?  Runs libaom 3.13.1 `cpu-used=6` (the
stand-in SURVEY.md 8d names) at five quality levels with tools switched off one at a time and prints each variant's BD-rate
(PSNR-Y, cubic) against the full encoder -- the roadmap numbers of DESIGN.md section 6 / 8b.  Every stream is decoded by dav1d.
Usage: tools/aom_ablation.py [--cqs 24,32,40,48,56] [--frames 30] [--out FILE]"""
import argparse, json, os, sys
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

VARIANTS = [
    ("full (lag 19)", 19, ()),
    ("lag_in_frames = 0 (low delay: no look-ahead, no alt-ref)", 0, ()),
    ("no alt-ref temporal filter (arnr-maxframes=0)", 19, (("arnr-maxframes", "0"),)),
    ("no key-frame filtering (enable-keyframe-filtering=0)", 19, (("enable-keyframe-filtering", "0"),)),
    ("no temporal filter at all (arnr-maxframes=0, enable-keyframe-filtering=0)", 19, (("arnr-maxframes", "0"), ("enable-keyframe-filtering", "0"))),
    ("no TPL model (enable-tpl-model=0)", 19, (("enable-tpl-model", "0"),)),
    ("one-level pyramid (gf-max-pyr-height=1)", 19, (("gf-max-pyr-height", "1"),)),
    ("flat golden-frame group (gf-max-pyr-height=0)", 19, (("gf-max-pyr-height", "0"),)),
    ("no masked / weighted / inter-intra compound", 19, (("enable-masked-comp", "0"), ("enable-dist-wtd-comp", "0"), ("enable-interintra-comp", "0"),
                                                        ("enable-diff-wtd-comp", "0"), ("enable-onesided-comp", "0"))),
    ("fixed 16x16 blocks (min = max partition size 16)", 19, (("min-partition-size", "16"), ("max-partition-size", "16"))),
    ("no OBMC / warped / global motion", 19, (("enable-obmc", "0"), ("enable-warped-motion", "0"), ("enable-global-motion", "0"))),
    ("DCT only (enable-flip-idtx=0, use-intra-dct-only=1, use-inter-dct-only=1)", 19, (("enable-flip-idtx", "0"), ("use-intra-dct-only", "1"), ("use-inter-dct-only", "1"))),
    ("three reference frames (max-reference-frames=3)", 19, (("max-reference-frames", "3"),)),
]


def run(job):
    name, lag, extra, cq, w, h, bd, n = job
    import numpy as np
    from av1_base_b200 import synth
    from oracle import decoders as D
    frames = synth.synth_clip(w, h, bd, n, seed=4, scene_len=1000, noise=1.0)
    try:
        tus = D.aom_encode(frames, bd, cq_level=cq, cpu_used=6, threads=2, lag=lag, extra=extra)
    except RuntimeError as e:
        return name, cq, None, None, str(e)
    dec = D.dav1d_decode(tus)
    ps = float(np.mean([D.psnr(dec[i][0], frames[i][0], bd) for i in range(n)]))
    return name, cq, sum(map(len, tus)) * 8 * 30.0 / n / 1000, ps, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cqs", default="24,32,40,48,56")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--size", default="960x544")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    cqs = list(map(int, a.cqs.split(",")))
    jobs = [(name, lag, extra, cq, w, h, 10, a.frames) for name, lag, extra in VARIANTS for cq in cqs]
    with ProcessPoolExecutor(max(1, (os.cpu_count() or 2) // 2)) as ex:
        res = list(ex.map(run, jobs))
    from bdrate import bd_rate
    pts = {}
    for name, cq, kbps, ps, err in res:
        if err:
            pts.setdefault(name, {"error": err})
        else:
            pts.setdefault(name, {"points": []})["points"].append({"cq": cq, "kbps": kbps, "psnr_y": ps})
    base = pts[VARIANTS[0][0]]["points"]
    out = {"clip": "%dx%d 10-bit, %d frames, synth seed 4 noise 1.0" % (w, h, a.frames), "encoder": "libaom 3.13.1 cpu-used=6, end-usage=q", "variants": []}
    for name, lag, extra in VARIANTS:
        v = pts[name]
        row = {"variant": name, "options": dict(extra), "lag_in_frames": lag}
        if "error" in v:
            row["error"] = v["error"]
        else:
            p = v["points"]
            row["points"] = p
            row["bd_rate_vs_full_psnr_y_pct"] = bd_rate([x["kbps"] for x in base], [x["psnr_y"] for x in base], [x["kbps"] for x in p], [x["psnr_y"] for x in p])
        out["variants"].append(row)
        print("%-80s %s" % (name, row.get("error") or ("%+.1f %%" % row["bd_rate_vs_full_psnr_y_pct"] if row["bd_rate_vs_full_psnr_y_pct"] is not None else "no overlap")), flush=True)
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
