"""One small invocation of the hot path on cuda:0, checked against the CPU oracle and both decoders."""
import numpy as np


def run():
    from . import encoder, synth
    from oracle import pyoracle as O, decoders as D   # the oracle is the checker here, never the product path
    if encoder.device_count() < 1:
        raise RuntimeError("smoke(): no CUDA device (av1b200 has no CPU fallback)")
    w, h, bd, crf = 192, 136, 10, 30
    frames = synth.synth_clip(w, h, bd, 2, seed=7)
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, frames_in_flight=2)
    tus = enc.encode_chunk(frames)
    g = enc.geom
    q = enc.stats()["base_q_idx"]
    pm = O.partition_fixed(g, 4)
    dec = D.dav1d_decode(tus)
    assert len(dec) == len(frames)
    for i, fr in enumerate(frames):
        ref = O.encode_intra_frame(g, fr, bd, q, pm)
        fp = enc.frame_params()
        O.deblock_frame(g, bd, ref.blocks, ref.rec, list(fp.lf_level), fp.lf_sharpness)
        ref.rec = O.cdef_frame(g, bd, ref.blocks, fp, O.cdef_search(g, bd, ref.blocks, fp, ref.rec, O.pad_planes(g, fr)), ref.rec)
        rec = enc.recon(i)
        orc = O.crop(g, ref.rec)
        for p in range(3):
            assert np.array_equal(rec[p], orc[p]), "CUDA recon != oracle recon (frame %d plane %d)" % (i, p)
            assert np.array_equal(dec[i][p], rec[p]), "dav1d decode != encoder recon (frame %d plane %d)" % (i, p)
    print("smoke OK: %d frames %dx%d %d-bit, %d bytes, CUDA == oracle == dav1d" % (len(frames), w, h, bd, sum(map(len, tus))))
