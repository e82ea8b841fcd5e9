"""BASELINE.json's full sizes (configs C1/C3/C4): the oracle is too slow there, so the check is the
size-independent property the domain offers -- every produced stream must decode in dav1d AND libaom to
exactly the encoder's own reconstruction (frame for frame, all planes), key and inter frames."""
import numpy as np
import pytest
from av1_base_b200 import encoder, synth
from oracle import decoders as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,bd,n,hdr", [(1920, 1080, 8, 5, False), (1920, 1080, 10, 5, False), (3840, 2160, 10, 4, True)])
def test_decode_matches_reconstruction_at_full_size(w, h, bd, n, hdr):
    frames = synth.synth_clip(w, h, bd, n, seed=3, scene_len=100, hdr=hdr)
    enc = encoder.Encoder(w, h, bd, crf=30, keep_debug=True, frames_in_flight=3, hdr=hdr)
    tus = enc.encode_chunk(frames)
    assert len(tus) == n
    assert enc.frame_is_key(0) and not enc.frame_is_key(1)
    dec_d = D.dav1d_decode(tus)
    dec_a = D.aom_decode(tus) if w <= 1920 else None      # libaom's decoder is slow at 4K; dav1d covers it
    assert len(dec_d) == n
    for i in range(n):
        rec = enc.recon(i)
        for p in range(3):
            assert np.array_equal(dec_d[i][p], rec[p]), ("dav1d", i, p)
            if dec_a is not None:
                assert np.array_equal(dec_a[i][p], rec[p]), ("libaom", i, p)
        psnr = D.psnr(rec[0], frames[i][0], bd)
        assert psnr > 34, (i, psnr)
    st = enc.stats()
    assert st["key_frames"] == 1 and st["inter_launches"] == n - 1
    enc.close()


def test_empty_and_tiny_chunks():
    """Edge cases: one-frame chunk (key only), smallest legal picture, chunk shorter than a batch."""
    for w, h, n in [(16, 16, 1), (16, 16, 3), (64, 24, 2), (24, 64, 2)]:
        frames = synth.synth_clip(w, h, 10, n, seed=1, scene_len=100)
        enc = encoder.Encoder(w, h, 10, crf=40, keep_debug=True, frames_in_flight=2)
        tus = enc.encode_chunk(frames)
        dec = D.dav1d_decode(tus)
        assert len(dec) == n
        for i in range(n):
            for p in range(3):
                assert np.array_equal(dec[i][p], enc.recon(i)[p]), (w, h, i, p)
        enc.close()
    with pytest.raises(encoder.EncodeError):
        encoder.Encoder(12, 16, 10)          # below 16 x 16 (sizes that are not multiples of 8 are padded: test_gpu_inter_parity.py)
    with pytest.raises(encoder.EncodeError):
        encoder.Encoder(8200, 64, 10)        # wider than 8192
    with pytest.raises(encoder.EncodeError):
        encoder.Encoder(64, 64, 12)          # unsupported bit depth


def test_page_locked_sources_are_read_in_place_and_give_the_same_stream():
    """av1b_host_alloc: page-locked source planes skip the encoder's staging copy (stats: staged_direct) and the
    stream is byte-identical to the one from pageable sources; a batch mixing both kinds takes the staged path."""
    w, h, bd, n = 328, 248, 10, 7
    frames = synth.synth_clip(w, h, bd, n, seed=6, scene_len=100)
    enc = encoder.Encoder(w, h, bd, crf=32, frames_in_flight=3)
    ref = enc.encode_chunk(frames)
    assert enc.stats()["staged_direct"] == 0
    pinned = encoder.PinnedFrames(frames)
    got = enc.encode_chunk(list(pinned))
    assert enc.stats()["staged_direct"] == n
    assert got == ref
    mixed = [pinned[i] if i % 2 else frames[i] for i in range(n)]
    got = enc.encode_chunk(mixed)
    assert enc.stats()["staged_direct"] == 0
    assert got == ref
    # strided page-locked planes (a wider allocation): rows are gathered by the copy engine
    wide = encoder.PinnedFrames([[np.pad(p, ((0, 0), (0, 24))) for p in fr] for fr in frames])
    views = [[p[:, :p.shape[1] - 24] for p in fr] for fr in wide]
    got = enc.encode_chunk_strided(views)
    assert enc.stats()["staged_direct"] == n
    assert got == ref
    pinned.close(); wide.close()
    enc.close()


def test_film_grain_is_signalled_and_decoders_add_it():
    """--film-grain (row f-4): the temporal filter takes the noise out of the anchors, the headers carry film grain
    parameters of three quarters of the measured noise strength; without grain dav1d still decodes to the encoder's
    reconstruction bit for bit, with grain dav1d and libaom put noise of about that strength back."""
    w, h, bd, n = 328, 248, 10, 8
    frames = synth.synth_clip(w, h, bd, n, seed=5, scene_len=100, noise=1.0)
    enc = encoder.Encoder(w, h, bd, crf=40, keep_debug=True, film_grain=20)
    tus = enc.encode_chunk(frames)
    info = enc.chunk_info()
    assert info["gop_period"] > 1 and info["noise_b"] > 0
    plain = D.dav1d_decode(tus)
    grainy = D.dav1d_decode(tus, apply_grain=True)
    aom = D.aom_decode(tus)
    assert len(plain) == n and len(grainy) == n and len(aom) == n
    want = 0.75 * 0.0010658 * info["noise_b"]          # sigma in samples at this bit depth
    for i in range(n):
        rec = enc.recon(i)
        for p in range(3):
            assert np.array_equal(plain[i][p], rec[p]), ("dav1d without grain", i, p)
        for name, dec in (("dav1d", grainy), ("libaom", aom)):
            d = dec[i][0].astype(np.int64) - rec[0].astype(np.int64)
            assert 0.5 * want < d.std() < 1.6 * want, (name, i, d.std(), want)
    # the same clip without --film-grain: no grain parameters, the decoders agree with the reconstruction as they are
    enc2 = encoder.Encoder(w, h, bd, crf=40, keep_debug=True)
    tus2 = enc2.encode_chunk(frames)
    dec2 = D.aom_decode(tus2)
    for i in range(n):
        assert np.array_equal(dec2[i][0], enc2.recon(i)[0]), i
    enc.close(); enc2.close()
