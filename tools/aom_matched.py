#!/usr/bin/env python3
"""TEST / ANALYSIS TOOL (CPU only): this encoder (oracle chain = the device's decisions, bit for bit) against libaom 3.13.1
`cpu-used=6` restricted to a MATCHING TOOLSET -- the tools this encoder does not have switched off in libaom: fixed 16x16 blocks,
a one-level pyramid, no TPL model, three reference frames, DCT only, no OBMC / warped / global motion, no masked / weighted /
inter-intra compound -- on the clip of the bench line's `bd_rate` (960x544 10-bit, 30 frames, synth seed 4, noise 1.0).  Separates
"which tools are missing" (tools/aom_ablation.py) from "how good are the decisions with the tools that are there".
Usage: tools/aom_matched.py [OUT.json]      Record: profiles/r02z_vs_libaom_matched_tools_960x544.json (+4.8 % on these 30 frames -- a
quarter of which is the key frame, where the restriction hurts libaom most -- and +60 % over a whole 150-frame chunk, `over_150_frames`)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
from concurrent.futures import ProcessPoolExecutor
import numpy as np
EXTRA = (("min-partition-size", "16"), ("max-partition-size", "16"), ("gf-max-pyr-height", "1"), ("enable-tpl-model", "0"),
         ("max-reference-frames", "3"), ("enable-flip-idtx", "0"), ("use-intra-dct-only", "1"), ("use-inter-dct-only", "1"),
         ("enable-obmc", "0"), ("enable-warped-motion", "0"), ("enable-global-motion", "0"), ("enable-masked-comp", "0"),
         ("enable-dist-wtd-comp", "0"), ("enable-interintra-comp", "0"), ("enable-diff-wtd-comp", "0"), ("enable-onesided-comp", "0"))
def aom(cq):
    from av1_base_b200 import synth
    from oracle import decoders as D
    frames = synth.synth_clip(960, 544, 10, 30, seed=4, scene_len=1000, noise=1.0)
    tus = D.aom_encode(frames, 10, cq_level=cq, cpu_used=6, threads=2, lag=19, extra=EXTRA)
    dec = D.dav1d_decode(tus)
    return dict(cq=cq, kbps=sum(map(len, tus)) * 8 * 30.0 / 30 / 1000, psnr_y=float(np.mean([D.psnr(dec[i][0], frames[i][0], 10) for i in range(30)])))
def ours(crf):
    from rd_chain import encode
    return encode((960, 544, 10, 30, 4, 1.0, crf, 0, {}))
if __name__ == "__main__":
    with ProcessPoolExecutor(8) as ex:
        fa = [ex.submit(aom, cq) for cq in (16, 24, 32, 40, 48, 56)]
        fo = [ex.submit(ours, crf) for crf in (20, 30, 36, 40, 44, 48, 52, 58)]
        ra = [f.result() for f in fa]; ro = [f.result() for f in fo]
    from bdrate import bd_rate, bd_rate_pchip
    full = json.load(open(os.path.join(ROOT, 'profiles', 'r02w_libaom_tool_ablation_960x544.json')))["variants"][0]["points"]
    out = {"clip": "960x544 10-bit, 30 frames, synth seed 4 noise 1.0",
           "libaom_restricted_options": dict(EXTRA), "libaom_restricted": ra, "ours_oracle_chain": ro,
           "ours_vs_libaom_restricted_psnr_y_pct": bd_rate([x["kbps"] for x in ra], [x["psnr_y"] for x in ra], [x["kbps"] for x in ro], [x["psnr_y"] for x in ro]),
           "ours_vs_libaom_restricted_psnr_y_pchip_pct": bd_rate_pchip([x["kbps"] for x in ra], [x["psnr_y"] for x in ra], [x["kbps"] for x in ro], [x["psnr_y"] for x in ro]),
           "libaom_restricted_vs_full_psnr_y_pct": bd_rate([x["kbps"] for x in full], [x["psnr_y"] for x in full], [x["kbps"] for x in ra], [x["psnr_y"] for x in ra])}
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else '/tmp/aom_matched.json', 'w'), indent=1)
    print({k: v for k, v in out.items() if k.endswith("_pct")})
    for r in ra: print("aom", r)
    for r in ro: print("ours", r["crf"], r["kbps"], r["psnr_y"])
