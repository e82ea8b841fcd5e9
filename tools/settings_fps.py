#!/usr/bin/env python3
"""Resident-clip throughput of the encoder at the settings the daemon really passes (av1an.rs:14) beside the bench's
(BASELINE.json configs: CRF 30, preset 6), measured in one process on one GPU: the same 4K 10-bit clip as bench.py (synth
seed 4, scene length 150, `--distinct` pictures walked back and forth), closed 150-frame chunks out of HBM
(av1b_stage_clip / av1b_encode_clip), wall clock around the calls (host entropy coding included).  No torch: a short run.
Usage: tools/settings_fps.py [--distinct 30] [--chunks 8] [--warmup 3] [--size 3840x2160] [--out FILE]"""
import argparse, json, os, sys, time
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _gen(job):
    from av1_base_b200 import synth
    w, h, bd, i = job
    return synth.synth_clip(w, h, bd, 1, seed=4, scene_len=150, hdr=(bd == 10 and w >= 3840), start=i)[0]


def chunk_order(n_distinct, n_frames):
    period = list(range(n_distinct)) + list(range(n_distinct - 2, 0, -1))
    return [period[i % len(period)] for i in range(n_frames)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--distinct", type=int, default=30)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunk-len", type=int, default=150)
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", type=int, default=-1, help="run only this row of the settings table")
    ap.add_argument("--pack-paths", default="0", help="comma list: 0 automatic placement of the range coder, 3 host, 4 device")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    bd = a.bd
    t0 = time.perf_counter()
    with ProcessPoolExecutor(min(a.distinct, max(1, (os.cpu_count() or 2) // 2), 16)) as ex:
        frames = list(ex.map(_gen, [(w, h, bd, i) for i in range(a.distinct)]))
    t_synth = time.perf_counter() - t0
    from av1_base_b200 import encoder
    order = chunk_order(a.distinct, a.chunk_len)
    daemon = dict(preset=3, film_grain=20, qm=(1, 15), lookahead=40)
    configs = [("--crf 30 --preset 6 (bench.py, BASELINE.json configs)", dict(crf=30, preset=6)),
               ("--crf 30 --preset 3 --film-grain 20 --enable-qm 1 --qm-min 1 --qm-max 15 --lookahead 40", dict(crf=30, **daemon)),
               ("--crf 8 --preset 3 --film-grain 20 --enable-qm 1 --qm-min 1 --qm-max 15 --lookahead 40 (av1an.rs:14)", dict(crf=8, **daemon))]
    rows = []
    if a.only >= 0:
        configs = [configs[a.only]]
    configs = [(name + (" [range coder: %s]" % {0: "automatic", 3: "host", 4: "device"}[pp] if a.pack_paths != "0" else ""), dict(kw, pack_path=pp))
               for name, kw in configs for pp in map(int, a.pack_paths.split(","))]
    for name, kw in configs:
        enc = encoder.Encoder(w, h, bd, hdr=(bd == 10 and w >= 3840), frames_in_flight=8, keyint=240, **kw)
        enc.stage_clip(frames)
        for _ in range(a.warmup):                    # warm-up chunks (clocks, lazily loaded kernels, staging buffers)
            enc.encode_clip(order)
        per = []
        t0 = time.perf_counter()
        for c in range(a.chunks):
            t1 = time.perf_counter()
            enc.encode_clip(order, accumulate=c > 0)
            per.append(round(1e3 * (time.perf_counter() - t1), 1))
        dt = time.perf_counter() - t0
        st, info = enc.stats(), enc.chunk_info()
        n = a.chunks * a.chunk_len
        rows.append({"settings": name, "fps": round(n / dt, 1), "ms_per_chunk": round(1e3 * dt / a.chunks, 2),
                     "kernel_ms_per_chunk": round(st["kernel_ms"] / a.chunks, 2), "pack_ms_per_chunk": round(st["pack_ms"] / a.chunks, 2), "rc_ms_per_chunk": round(st.get("rc_ms", 0) / a.chunks, 2),
                     "bytes_per_frame": round(st["bytes_out"] / max(1, st["frames_done"]), 1), "gop_period": info["gop_period"],
                     "q_key_anchor_nonref": [info["q_key"], info["q_anchor"], info["q_nonref"]], "temporal_filter": info["mctf"], "ms_of_each_chunk": per})
        enc.close()
    out = {"what": "resident 150-frame chunks, one B200, %dx%d %d-bit, %d distinct pictures" % (w, h, bd, a.distinct),
           "timed_chunks": a.chunks, "warmup_chunks": a.warmup, "synth_s": round(t_synth, 1), "host_cores": os.cpu_count(), "rows": rows}
    txt = json.dumps(out, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
