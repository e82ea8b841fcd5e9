"""GPU parity (through the C ABI): the CUDA intra encode path must reproduce the CPU oracle bit for
bit (reconstruction, block side info, quantised coefficients) and its bitstreams must decode in
dav1d AND libaom to exactly the encoder's reconstruction."""
import numpy as np
import pytest
from av1_base_b200 import encoder, synth
from oracle import pyoracle as O, decoders as D

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, bd, crf, blk_log2, tile_cols_log2, tile_rows_log2
    (64, 64, 8, 30, 6, 0, 0),
    (64, 64, 10, 30, 3, 0, 0),
    (128, 128, 8, 20, 5, 0, 0),
    (200, 136, 8, 35, 4, 0, 0),
    (200, 136, 10, 10, 3, 0, 0),
    (328, 248, 10, 30, 6, 1, 1),
    (328, 248, 8, 55, 5, 2, 1),
    (640, 360, 10, 30, 4, 2, 2),
]


@pytest.mark.parametrize("w,h,bd,crf,blk,tcl,trl", CASES)
def test_intra_frame_parity(w, h, bd, crf, blk, tcl, trl):
    frames = synth.synth_clip(w, h, bd, 3, seed=w + h + bd, scene_len=2)
    enc = encoder.Encoder(w, h, bd, crf=crf, keep_debug=True, blk_log2=blk, tile_cols_log2=tcl, tile_rows_log2=trl,
                          frames_in_flight=2)
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
        orc = O.crop(g, ref.rec)
        for p in range(3):
            hh, ww = (g.height, g.width) if p == 0 else (g.height // 2, g.width // 2)
            assert np.array_equal(coef[p][:hh, :ww], ref.coef[p][:hh, :ww]), ("coef", i, p)
            assert np.array_equal(rec[p], orc[p]), ("recon vs oracle", i, p)
            assert np.array_equal(dec_d[i][p], rec[p]), ("dav1d", i, p)
            assert np.array_equal(dec_a[i][p], rec[p]), ("libaom", i, p)
    enc.close()


def test_no_device_id_out_of_range():
    with pytest.raises(encoder.EncodeError):
        encoder.Encoder(64, 64, 8, device_id=99)
