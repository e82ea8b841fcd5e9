#!/usr/bin/env python3
"""Times the host entropy coder (av1b_pack_frame, one thread) on oracle-produced symbol streams of a
1080p key frame and inter frame.  CPU only.  Usage: tools/bench_pack.py [w h]"""
import sys, time, os, pickle
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O
import ctypes as C

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
bd, q = 10, 120
cache = "/tmp/bench_pack_%dx%d.pkl" % (w, h)
g = O.geom(w, h, 3, 3)
if os.path.exists(cache):
    res = pickle.load(open(cache, "rb"))
else:
    frames = synth.synth_clip(w, h, bd, 2, seed=4, scene_len=100)
    pm = O.partition_fixed(g, 4)
    r0 = O.encode_intra_frame(g, frames[0], bd, q, pm)
    mv = O.hme(g, O.pyramid(g, O.pad_planes(g, frames[1])[0]), O.pyramid(g, O.pad_planes(g, frames[0])[0]), 280, bd)
    r1 = O.encode_inter_frame(g, frames[1], bd, q, pm, mv, r0.rec)
    O.merge_skip_blocks(g, r1.blocks)
    res = [(r.blocks, r.coef) for r in (r0, r1)]
    pickle.dump(res, open(cache, "wb"))
seq = abi.SeqParams(w, h, bd, 1, 0, 30, 1, 0)
for ft, (blocks, coef) in enumerate(res):
    fp = abi.FrameParams()
    abi.lib().av1b_select_frame_params(bd, q, ft, 1, C.byref(fp))
    fp.tile_cols_log2, fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
    sy = packer.make_syms(g, blocks, coef)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        out = packer.pack_frame(seq, fp, sy, n_threads=1)
        best = min(best, time.perf_counter() - t0)
    nz = sum(int(np.count_nonzero(c)) for c in coef)
    print("%s frame: %d bytes, %d nonzero levels, pack %.2f ms (1 thread) = %.1f ns/byte" %
          ("key" if ft == 0 else "inter", len(out), nz, best * 1e3, best * 1e9 / len(out)))
    if ft == 1:
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            out2, ntok = packer.pack_frame_tokens(seq, fp, sy)
            best = min(best, time.perf_counter() - t0)
        assert out2 == out
        print("inter frame, token path (CPU tokenizer + range coder): %d tokens, %.2f ms" % (ntok, best * 1e3))
