#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Randomised sweep of the oracle's decision chain (oracle/chain.py, what the CUDA path is
compared with) through the host bitstream writer and BOTH decoders: picture sizes 16..334 (multiples of 8 and not: padded +
render_size), 8 / 10 bits, CRF 1..63, P chain / hierarchies / automatic structure, key frames every 3 / 5 / 240 frames and at scene
cuts, loop restoration, quantisation matrix ranges, two or three regularisation sweeps, film-grain-strength temporal filter, fixed or
smoothness-driven key-frame partition, up to 4 x 4 tiles, 1..7 frames.  dav1d and libaom must reproduce the chain's reconstruction of every frame.
CPU only.  Usage: tools/fuzz_chain.py SEED ITERATIONS [big]   (run several seeds side by side; about 8 configurations per second and core)
Record: seeds 1000-4000 x 150, 5000-10000 x 500 (single tile) and 11000-16000 x 400 (random tilings) = 6000 configurations, plus seeds 21000-27000 x 250 `big` = 1750 larger
ones, no mismatch (round 2)."""
import sys, os, random, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O, decoders as D, chain
from tests.test_oracle_chain import pack_chain, unaligned_clip
seed0 = int(sys.argv[1]); n_iter = int(sys.argv[2])
BIG = len(sys.argv) > 3 and sys.argv[3] == "big"      # larger pictures (up to 648 x 366), up to 12 frames: about 1 configuration per second and core
rng = random.Random(seed0)
bad = 0
for it in range(n_iter):
    w = rng.choice([16, 24, 40, 64, 72, 104, 136, 200, 264, 328]) + rng.choice([0, 0, 0, 2, 3, 5])
    h = rng.choice([16, 24, 40, 64, 88, 136, 184, 248]) + rng.choice([0, 0, 0, 1, 4, 6])
    if BIG:
        w = rng.choice([328, 400, 512, 640]) + rng.choice([0, 0, 4, 7, 8]); h = rng.choice([184, 248, 288, 360]) + rng.choice([0, 0, 2, 6])
    bd = rng.choice([8, 10]); crf = rng.randint(1, 63); n = rng.randint(1, 12 if BIG else 7)
    gop = rng.choice([0, 1, 2, 3, 6]); lr = rng.random() < 0.4
    qm = None if rng.random() < 0.4 else tuple(sorted((rng.randint(0, 15), rng.randint(0, 15))))
    noise = rng.choice([0.0, 0.3, 1.0, 2.0]); keyint = rng.choice([240, 3, 5])
    passes = rng.choice([2, 3]); fg = rng.choice([0, 20]); kvp = rng.random() < 0.7
    cfg = dict(seed=seed0, it=it, w=w, h=h, bd=bd, crf=crf, n=n, gop=gop, lr=lr, qm=qm, noise=noise, keyint=keyint, passes=passes, fg=fg, kvp=kvp)
    try:
        src, padded, cw, ch = unaligned_clip(w, h, bd, n, seed=it + seed0)
        if noise != 0.4:
            big = synth.synth_clip(cw, ch, bd, n, seed=it + seed0, scene_len=rng.choice([100, 3]), noise=noise)
            padded = [[np.pad(f[0][:h, :w], ((0, ch - h), (0, cw - w)), mode="edge"),
                       np.pad(f[1][:(h + 1) // 2, :(w + 1) // 2], ((0, ch // 2 - (h + 1) // 2), (0, cw // 2 - (w + 1) // 2)), mode="edge"),
                       np.pad(f[2][:(h + 1) // 2, :(w + 1) // 2], ((0, ch // 2 - (h + 1) // 2), (0, cw // 2 - (w + 1) // 2)), mode="edge")] for f in big]
        tiles = (rng.choice([0, 0, 1, 2]), rng.choice([0, 0, 1, 2]))      # uniform tiling of every frame (clamped to what the size allows)
        g0 = O.geom(cw, ch, tiles[0], tiles[1])
        g, want = chain.encode_chain(padded, cw, ch, bd, crf, keyint=keyint, gop_period=gop, lr=lr, qm=qm, sbrd_passes=passes, film_grain=fg, key_var_part=kvp, geom=g0)
        for r in want:
            r.fp.tile_cols_log2, r.fp.tile_rows_log2 = g.tile_cols_log2, g.tile_rows_log2
        tus = pack_chain(cw, ch, bd, want, g, lr=lr, render=(w, h) if (cw, ch) != (w, h) else (0, 0))
        for dec in (D.dav1d_decode, D.aom_decode):
            out = dec(tus)
            assert len(out) == n, (dec.__name__, len(out))
            for i in range(n):
                for p in range(3):
                    if not np.array_equal(out[i][p], O.crop(g, want[i].fin)[p]):
                        raise AssertionError("mismatch %s frame %d plane %d" % (dec.__name__, i, p))
    except Exception as e:
        bad += 1
        print("FAIL", json.dumps(cfg), repr(e)[:300], flush=True)
        traceback.print_exc()
    if it % 20 == 19:
        print("done", it + 1, "bad", bad, flush=True)
print("FINISHED", n_iter, "bad", bad)
