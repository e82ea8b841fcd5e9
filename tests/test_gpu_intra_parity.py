"""GPU parity (through the C ABI): the CUDA intra encode path must reproduce the CPU oracle bit for
bit (reconstruction, block side info, quantised coefficients) and its bitstreams must decode in
dav1d AND libaom to exactly the encoder's reconstruction."""
import numpy as np
import pytest
from av1_base_b200 import encoder, synth
from oracle import pyoracle as O, decoders as D

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, bd, crf, blk_log2, tile_cols_log2, tile_rows_log2, loop_filters
    (64, 64, 8, 30, 6, 0, 0, True),
    (64, 64, 10, 30, 3, 0, 0, False),
    (128, 128, 8, 20, 5, 0, 0, True),
    (200, 136, 8, 35, 4, 0, 0, True),
    (200, 136, 10, 10, 3, 0, 0, True),
    (328, 248, 10, 30, 6, 1, 1, True),
    (328, 248, 8, 55, 5, 2, 1, False),
    (328, 248, 8, 63, 5, 2, 1, True),
    (640, 360, 10, 30, 4, 2, 2, True),
]


@pytest.mark.parametrize("w,h,bd,crf,blk,tcl,trl,lf", CASES)
def test_intra_frame_parity(w, h, bd, crf, blk, tcl, trl, lf):
    frames = synth.synth_clip(w, h, bd, 3, seed=w + h + bd, scene_len=2)
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, blk_log2=blk, tile_cols_log2=tcl, tile_rows_log2=trl,
                          frames_in_flight=2, loop_filters=lf, intra_only=True)
    fp = enc.frame_params()
    tus = enc.encode_chunk(frames)
    assert len(tus) == len(frames)
    g = enc.geom
    q = enc.stats()["base_q_idx"]
    pm = O.partition_fixed(g, blk)
    dec_d = D.dav1d_decode(tus)
    dec_a = D.aom_decode(tus)
    assert len(dec_d) == len(frames) and len(dec_a) == len(frames)
    for i, fr in enumerate(frames):
        ref = O.encode_intra_frame(g, fr, bd, q, pm)
        blocks, coef = enc.frame_syms(i)
        for f in ("blk_log2", "y_mode", "uv_mode", "skip", "eob", "tx_type_y"):
            assert np.array_equal(blocks[f], ref.blocks[f]), (f, i)
        rec = enc.recon(i)
        if lf:   # in-loop filters: deblock -> CDEF preset decision -> CDEF, as the oracle defines them
            O.deblock_frame(g, bd, ref.blocks, ref.rec, list(fp.lf_level), fp.lf_sharpness)
            src = O.pad_planes(g, fr)
            want_idx = O.cdef_search(g, bd, ref.blocks, fp, ref.rec, src)
            got_idx = enc.cdef_idx(i)
            live = ~ref.blocks["skip"].reshape(g.h8, g.w8).astype(bool)
            for sr in range(g.sb_rows):
                for sc in range(g.sb_cols):
                    if live[sr * 8:sr * 8 + 8, sc * 8:sc * 8 + 8].any():
                        assert got_idx[sr * g.sb_cols + sc] == want_idx[sr * g.sb_cols + sc], ("cdef_idx", i, sr, sc)
            ref.rec = O.cdef_frame(g, bd, ref.blocks, fp, got_idx, ref.rec)
        orc = O.crop(g, ref.rec)
        for p in range(3):
            hh, ww = (g.height, g.width) if p == 0 else (g.height // 2, g.width // 2)
            assert np.array_equal(coef[p], ref.coef[p]), ("coef", i, p)
            assert np.array_equal(rec[p], orc[p]), ("recon vs oracle", i, p)
            assert np.array_equal(dec_d[i][p], rec[p]), ("dav1d", i, p)
            assert np.array_equal(dec_a[i][p], rec[p]), ("libaom", i, p)
    enc.close()


def test_no_device_id_out_of_range():
    with pytest.raises(encoder.EncodeError):
        encoder.Encoder(64, 64, 8, device_id=99)
