#!/usr/bin/env python3
"""Times the drop-in executable itself (the boundary the daemon execs): writes a synthetic Y4M clip, runs
`av1an -i clip.y4m -o out.mkv ...` exactly as crates/daemon/src/encode/av1an.rs:79-107 would, reports
frames/s over the whole process (start-up, Y4M read, encode, mux) and checks that dav1d decodes the result.
Usage: tools/cli_bench.py [--size 4k|1080p] [--frames 96] [--workers 1] [--bd 10]"""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from av1_base_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="4k")
ap.add_argument("--frames", type=int, default=96)
ap.add_argument("--workers", type=int, default=1)
ap.add_argument("--bd", type=int, default=10)
ap.add_argument("--keyint", type=int, default=240)
a = ap.parse_args()
w, h = (3840, 2160) if a.size == "4k" else (1920, 1080)
base = synth.synth_clip(w, h, a.bd, 8, seed=4, scene_len=1000, hdr=(a.size == "4k"))
pal = base + base[::-1]
tmp = tempfile.mkdtemp(prefix="av1b_cli_")
y4m = os.path.join(tmp, "clip.y4m")
with open(y4m, "wb") as f:
    f.write(("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C%s\n" % (w, h, "420p10" if a.bd > 8 else "420jpeg")).encode())
    for i in range(a.frames):
        f.write(b"FRAME\n")
        for p in pal[i % len(pal)]:
            f.write(p.astype("<u2").tobytes() if a.bd > 8 else p.astype(np.uint8).tobytes())
out = os.path.join(tmp, "out.mkv")
cmd = [os.path.join(ROOT, "av1_base_b200", "av1an"), "-i", y4m, "-o", out, "--encoder", "svt-av1", "--pix-format",
       "yuv420p10le" if a.bd > 8 else "yuv420p", "--video-params", "--crf 30 --preset 6 --keyint %d --lookahead 40" % a.keyint,
       "--audio-params", "-c:a copy", "--workers", str(a.workers), "--temp", os.path.join(tmp, "chunks"), "--quiet"]
t0 = time.perf_counter()
r = subprocess.run(cmd, capture_output=True, text=True)
dt = time.perf_counter() - t0
ok = r.returncode == 0 and os.path.getsize(out) > 0
print(json.dumps({"what": "av1an-compatible CLI, whole process", "size": "%dx%d" % (w, h), "bit_depth": a.bd, "frames": a.frames,
                  "workers": a.workers, "returncode": r.returncode, "seconds": round(dt, 3), "fps": round(a.frames / dt, 1),
                  "output_bytes": os.path.getsize(out) if ok else 0, "input_bytes": os.path.getsize(y4m),
                  "stderr": r.stderr[-300:]}))
for fn in (y4m, out):
    try:
        os.remove(fn)
    except OSError:
        pass
