#!/usr/bin/env python3
"""BASELINE config C4: a 4K 10-bit HDR synthetic clip with a scene cut every 150 frames, scene-chunked across the GPUs of the
box THROUGH THE DROP-IN EXECUTABLE (av1_base_b200/av1an --workers N, the argv of crates/daemon/src/encode/av1an.rs:79-107).
The Y4M file is written to --dir (default /dev/shm; the frame count is cut to what fits there): every scene walks back and
forth over --distinct pictures of synth_clip.  Reports whole-process frames/s per --workers value, the number of chunks the
reader cut, and checks the Matroska file: one block per frame, dav1d decodes the first pictures of the first two chunks.
Usage (GPU box): tools/c4_run.py [--frames 2400] [--workers 1,2,4,8] [--size 3840x2160] [--out profiles/x.json]"""
import argparse, json, os, shutil, subprocess, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scene_pictures(args):
    w, h, bd, scene, scene_len, distinct, hdr = args
    from av1_base_b200 import synth
    fr = synth.synth_clip(w, h, bd, distinct, seed=4, scene_len=scene_len, hdr=hdr, start=scene * scene_len)
    return [b"".join(p.astype("<u2").tobytes() if bd > 8 else p.astype(np.uint8).tobytes() for p in f) for f in fr]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2400)
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--scene-len", type=int, default=150)
    ap.add_argument("--distinct", type=int, default=30)
    ap.add_argument("--workers", default="1,2,4,8")
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--crf", type=int, default=30)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    fbytes = w * h * 3 // 2 * (2 if a.bd > 8 else 1)
    free = shutil.disk_usage(a.dir).free
    frames = min(a.frames, int(free * 0.8 // (fbytes + 6)) // a.scene_len * a.scene_len)
    if frames <= 0:
        raise SystemExit("no room in %s" % a.dir)
    tmp = os.path.join(a.dir, "av1b_c4_%d" % os.getpid())
    os.makedirs(tmp, exist_ok=True)
    y4m = os.path.join(tmp, "clip.y4m")
    n_scenes = (frames + a.scene_len - 1) // a.scene_len
    t0 = time.perf_counter()
    with ProcessPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        scenes = list(ex.map(scene_pictures, [(w, h, a.bd, s, a.scene_len, a.distinct, w > 3000) for s in range(n_scenes)]))
    hdr = ("YUV4MPEG2 W%d H%d F60:1 Ip A1:1 C%s\n" % (w, h, "420p10" if a.bd > 8 else "420jpeg")).encode()
    with open(y4m, "wb") as f:
        f.write(hdr)
        f.truncate(len(hdr) + frames * (6 + fbytes))
    fd = os.open(y4m, os.O_WRONLY)

    def put(i):
        s_, k = divmod(i, a.scene_len)
        k %= 2 * a.distinct - 2
        os.pwrite(fd, b"FRAME\n" + scenes[s_][k if k < a.distinct else 2 * a.distinct - 2 - k], len(hdr) + i * (6 + fbytes))

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as tp:      # tmpfs takes parallel writers
        list(tp.map(put, range(frames)))
    os.close(fd)
    del scenes
    t_gen = time.perf_counter() - t0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cli import mkv_blocks
    from oracle import decoders as D   # dav1d is the checker here
    runs = []
    for wk in map(int, a.workers.split(",")):
        out = os.path.join(tmp, "out_w%d.mkv" % wk)
        cmd = [os.path.join(ROOT, "av1_base_b200", "av1an"), "-i", y4m, "-o", out, "--encoder", "svt-av1", "--pix-format",
               "yuv420p10le" if a.bd > 8 else "yuv420p", "--video-params", "--crf %d --preset 6 --keyint 240 --lookahead 40" % a.crf,
               "--audio-params", "-c:a copy", "--workers", str(wk), "--temp", os.path.join(tmp, "chunks_w%d" % wk), "--quiet"]
        t0 = time.perf_counter()
        # (the executable would by itself use fewer workers for so short a clip: the scaling run asks for all of them)
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, AV1B_MIN_FRAMES_PER_WORKER="1"))
        dt = time.perf_counter() - t0
        rec = dict(workers=wk, returncode=r.returncode, seconds=round(dt, 2), fps=round(frames / dt, 1), stderr=r.stderr[-700:])
        if r.returncode == 0:
            data = open(out, "rb").read()
            codec, _, blocks = mkv_blocks(data)
            keys = [i for i, b in enumerate(blocks) if b[:2] == b"\x0a\x0b" or b[:1] == b"\x0a"]   # sequence header OBU first = a chunk starts
            rec.update(output_bytes=len(data), codec=codec, blocks=len(blocks), chunks=len(keys), one_block_per_frame=len(blocks) == frames,
                       chunk_starts=keys[:20])
            ok = True
            for c in keys[:2]:
                dec = D.dav1d_decode([b"\x12\x00" + b for b in blocks[c:c + 4]])
                ok = ok and len(dec) == 4 and dec[0][0].shape == (h, w)
            rec["dav1d_decodes_chunk_starts"] = bool(ok)
            os.remove(out)
        runs.append(rec)
        print(json.dumps(rec), flush=True)
    res = dict(config="C4: %dx%d %d-bit HDR synthetic, %d frames, scene cut every %d frames, through av1an --workers N" % (w, h, a.bd, frames, a.scene_len),
               frames=frames, y4m_bytes=os.path.getsize(y4m), generate_seconds=round(t_gen, 1), host_cores=os.cpu_count(), runs=runs)
    shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
