"""One small invocation of the hot path on cuda:0 (key frame + inter frames), checked against the CPU
oracle and the dav1d decoder."""
import numpy as np


def run():
    from . import encoder, synth
    from oracle import pyoracle as O, decoders as D   # the oracle is the checker here, never the product path
    if encoder.device_count() < 1:
        raise RuntimeError("smoke(): no CUDA device (av1b200 has no CPU fallback)")
    w, h, bd, crf = 192, 136, 10, 30
    frames = synth.synth_clip(w, h, bd, 3, seed=7)
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, frames_in_flight=2)
    tus = enc.encode_chunk(frames)
    g = enc.geom
    q = enc.stats()["base_q_idx"]
    pm = O.partition_fixed(g, 4)
    dec = D.dav1d_decode(tus)
    assert len(dec) == len(frames)
    prev_fin = prev_pyr = None
    for i, fr in enumerate(frames):
        pyr = O.pyramid(g, O.pad_planes(g, fr)[0])
        if i == 0:
            ref, fp = O.encode_intra_frame(g, fr, bd, enc.frame_params().base_q_idx, pm), enc.frame_params()
        else:
            ref, fp = O.encode_inter_frame(g, fr, bd, q, pm, O.hme(g, pyr, prev_pyr, enc.me_lambda()), prev_fin), enc.inter_frame_params()
            O.merge_skip_blocks(g, ref.blocks)
        O.deblock_frame(g, bd, ref.blocks, ref.rec, list(fp.lf_level), fp.lf_sharpness)
        fin = O.cdef_frame(g, bd, ref.blocks, fp, O.cdef_search(g, bd, ref.blocks, fp, ref.rec, O.pad_planes(g, fr)), ref.rec)
        rec = enc.recon(i)
        orc = O.crop(g, fin)
        for p in range(3):
            assert np.array_equal(rec[p], orc[p]), "CUDA recon != oracle recon (frame %d plane %d)" % (i, p)
            assert np.array_equal(dec[i][p], rec[p]), "dav1d decode != encoder recon (frame %d plane %d)" % (i, p)
        prev_fin, prev_pyr = fin, pyr
    print("smoke OK: %d frames (1 key + %d inter) %dx%d %d-bit, %d bytes, CUDA == oracle == dav1d"
          % (len(frames), len(frames) - 1, w, h, bd, sum(map(len, tus))))
