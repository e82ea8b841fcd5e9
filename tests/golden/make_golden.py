#!/usr/bin/env python3
"""Generates the committed golden fixtures: small key + inter frame streams produced by the CPU oracle
+ host entropy coder, together with the reconstruction the oracle holds for every frame.
The reference repository has no golden vectors for this path, so these pin OUR
definition: tests/test_golden.py checks that (a) the current oracle + packer reproduce the committed
bytes, (b) dav1d and libaom decode the committed bytes to the committed reconstruction.
Usage: python tests/golden/make_golden.py   (rewrites tests/golden/*.npz)"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ctypes as C
from av1_base_b200 import abi, packer, synth
from oracle import pyoracle as O

# name -> (w, h, bit depth, quantiser index, frames[, quantisation matrix level (--enable-qm, spec 7.12.3)])
CASES = {"key_inter_96x64_8bit": (96, 64, 8, 110, 3), "key_inter_72x88_10bit": (72, 88, 10, 140, 3),
         "key_inter_qm4_96x64_10bit": (96, 64, 10, 90, 3, 4)}


def encode(w, h, bd, q, n, qm=None):
    try:
        O.set_qm(*((qm, qm) if qm is not None else (15, 15)))
        return _encode(w, h, bd, q, n, qm)
    finally:
        O.set_qm()


def _encode(w, h, bd, q, n, qm):
    g = O.geom(w, h, 0, 0)
    frames = synth.synth_clip(w, h, bd, n, seed=w + h, scene_len=100)
    seq = abi.SeqParams(w, h, bd, 1, 0, 30, 1, 0)
    pm = O.partition_fixed(g, 4)
    acq = None
    tus, recs, prev, prev_pyr = [], [], None, None
    for i, fr in enumerate(frames):
        fp = abi.FrameParams()
        abi.lib().av1b_select_frame_params(bd, q, 0 if i == 0 else 1, 1, C.byref(fp))
        if qm is not None:
            fp.using_qmatrix, fp.qm_level[0], fp.qm_level[1] = 1, qm, qm
        pyr = O.pyramid(g, O.pad_planes(g, fr)[0])
        if i == 0:
            r = O.encode_intra_frame(g, fr, bd, q, pm)
        else:
            r = O.encode_inter_frame(g, fr, bd, q, pm, O.hme(g, pyr, prev_pyr, 40 << (bd - 8), bd), prev)
            O.merge_skip_blocks(g, r.blocks)
        O.deblock_frame(g, bd, r.blocks, r.rec, list(fp.lf_level), fp.lf_sharpness)
        idx = O.cdef_search(g, bd, r.blocks, fp, r.rec, O.pad_planes(g, fr))
        fin = O.cdef_frame(g, bd, r.blocks, fp, idx, r.rec)
        sy = packer.make_syms(g, r.blocks, r.coef, cdef_idx=idx)
        tus.append(b"\x12\x00" + (packer.pack_sequence_header(seq) if i == 0 else b"") + packer.pack_frame(seq, fp, sy, with_td=False))
        recs.append(O.crop(g, fin))
        prev, prev_pyr = fin, pyr
    return tus, recs


def chain_digest():
    """Digest of the oracle's whole decision chain (oracle/chain.py) on a small clip: frame kinds, quantisers,
    block-size histogram of the key frame, vector checksum and SHA-256 of every reconstructed frame."""
    import hashlib
    from oracle import chain
    w, h, bd, crf, n = 200, 136, 10, 34, 9
    frames = synth.synth_clip(w, h, bd, n, seed=21, scene_len=100, noise=0.5)
    g, res = chain.encode_chain(frames, w, h, bd, crf)
    out = {"kinds": [r.kind for r in res], "q": [r.q for r in res],
           "key_blocks": np.bincount(res[0].res.blocks["blk_log2"], minlength=7).tolist(),
           "mv_sum": [int(np.abs(r.mvs.astype(np.int64)).sum()) if r.mvs is not None else 0 for r in res],
           "sha": [hashlib.sha256(b"".join(np.ascontiguousarray(pl).tobytes() for pl in O.crop(g, r.fin))).hexdigest() for r in res]}
    return out


def main():
    import __graft_entry__ as ge
    ge.build()
    import json
    json.dump(chain_digest(), open(os.path.join(HERE, "chain_digest.json"), "w"), indent=1)
    for name, case in CASES.items():
        n = case[4]
        tus, recs = encode(*case)
        arrs = {"n": np.array([n])}
        for i in range(n):
            arrs["tu%d" % i] = np.frombuffer(tus[i], np.uint8)
            for p in range(3):
                arrs["rec%d_%d" % (i, p)] = recs[i][p]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        print(name, [len(t) for t in tus])


if __name__ == "__main__":
    main()
