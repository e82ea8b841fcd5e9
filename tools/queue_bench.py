#!/usr/bin/env python3
"""Scaled-down BASELINE config C5 ("daemon queue: mixed 1080p / 4K files scheduled across the GPUs"): K synthetic
Y4M files are pushed through the av1an-compatible executable by a pool of J concurrent jobs, each `--workers 1`,
exactly the way the daemon's JobExecutor would (one exec per job, max_concurrent_jobs permits,
job_executor.rs:185-194).  The executable's per-device flock leases spread the jobs over the GPUs.
Reports whole-box frames/s and checks every output with dav1d (frame count).
Usage: tools/queue_bench.py [--files 6] [--jobs 2] [--frames-1080p 96] [--frames-4k 32]"""
import argparse, json, os, subprocess, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from av1_base_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=6)
ap.add_argument("--jobs", type=int, default=2)
ap.add_argument("--frames-1080p", type=int, default=96)
ap.add_argument("--frames-4k", type=int, default=32)
ap.add_argument("--verify", type=int, default=1)
a = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="av1b_queue_")
CLI = os.path.join(ROOT, "av1_base_b200", "av1an")
jobs = []
for k in range(a.files):
    big = k & 1
    w, h, bd, n = (3840, 2160, 10, a.frames_4k) if big else (1920, 1080, 8, a.frames_1080p)
    base = synth.synth_clip(w, h, bd, 8, seed=100 + k, scene_len=1000, hdr=bool(big))
    pal = base + base[::-1]
    y4m = os.path.join(tmp, "in%02d.y4m" % k)
    with open(y4m, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C%s\n" % (w, h, "420p10" if bd > 8 else "420jpeg")).encode())
        for i in range(n):
            f.write(b"FRAME\n")
            for p in pal[i % len(pal)]:
                f.write(p.astype("<u2").tobytes() if bd > 8 else p.astype(np.uint8).tobytes())
    jobs.append(dict(k=k, y4m=y4m, out=os.path.join(tmp, "out%02d.mkv" % k), frames=n, w=w, h=h, bd=bd))


def run(j):
    cmd = [CLI, "-i", j["y4m"], "-o", j["out"], "--encoder", "svt-av1", "--pix-format", "yuv420p10le",
           "--video-params", "--crf 30 --preset 6 --keyint 240 --lookahead 40", "--audio-params", "-c:a copy",
           "--workers", "1", "--temp", os.path.join(tmp, "chunks_%02d" % j["k"]), "--quiet"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    return dict(k=j["k"], rc=r.returncode, seconds=round(time.perf_counter() - t0, 2), err=r.stderr[-200:],
                bytes=os.path.getsize(j["out"]) if os.path.exists(j["out"]) else 0)


t0 = time.perf_counter()
with ThreadPoolExecutor(a.jobs) as ex:
    res = list(ex.map(run, jobs))
dt = time.perf_counter() - t0
ok = all(r["rc"] == 0 and r["bytes"] > 0 for r in res)
verified = None
if a.verify and ok:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cli import mkv_blocks
    from oracle import decoders as D
    verified = True
    for j in jobs[:2]:
        _, _, blocks = mkv_blocks(open(j["out"], "rb").read())
        dec = D.dav1d_decode([b"\x12\x00" + b for b in blocks])
        verified = verified and len(dec) == j["frames"] and dec[0][0].shape == (j["h"], j["w"])
total = sum(j["frames"] for j in jobs)
print(json.dumps({"what": "queue of mixed 1080p8 / 4K10 files through the av1an-compatible executable", "files": a.files,
                  "concurrent_jobs": a.jobs, "frames": total, "seconds": round(dt, 2), "whole_box_fps": round(total / dt, 1),
                  "all_ok": ok, "decoded_ok": verified, "jobs": res}))
for j in jobs:
    for fn in (j["y4m"], j["out"]):
        try:
            os.remove(fn)
        except OSError:
            pass
