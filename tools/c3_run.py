#!/usr/bin/env python3
"""BASELINE config C3: a 1080p 10-bit synthetic clip on ONE B200, every produced frame decoded with dav1d and compared,
by SHA-256 of its three planes, with the encoder's own reconstruction (av1b_get_recon).  The clip is coded the way the
CLI codes it: closed chunks of --keyint frames (scene_len = keyint, so every chunk starts at a scene cut).
Usage (GPU box): tools/c3_run.py [--frames 1200] [--size 1920x1080] [--bd 10] [--crf 30] [--keyint 240] [--out profiles/x.json]"""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(w, h, bd, frames, crf=30, keyint=240, device_id=0):
    from av1_base_b200 import encoder, synth
    from oracle import decoders as D   # dav1d is the checker here, never the product path
    enc = encoder.Encoder(w, h, bd, crf=crf, keyint=keyint, device_id=device_id, keep_debug=True)
    done, nbytes, mism, kinds = 0, 0, 0, {0: 0, 1: 0, 2: 0}
    t_enc = t_dec = t_gen = 0.0
    digest = hashlib.sha256()
    while done < frames:
        n = min(keyint, frames - done)
        t0 = time.perf_counter()
        clip = synth.synth_clip(w, h, bd, n, seed=3, scene_len=keyint, start=done)
        t1 = time.perf_counter()
        tus = enc.encode_chunk(clip)
        t2 = time.perf_counter()
        dec = D.dav1d_decode(tus)
        t3 = time.perf_counter()
        t_gen += t1 - t0; t_enc += t2 - t1; t_dec += t3 - t2
        if len(dec) != n:
            raise RuntimeError("dav1d returned %d frames for a chunk of %d" % (len(dec), n))
        for i in range(n):
            rec = enc.recon(i)
            hd = hashlib.sha256(b"".join(np.ascontiguousarray(p).tobytes() for p in dec[i])).digest()
            hr = hashlib.sha256(b"".join(np.ascontiguousarray(p).tobytes() for p in rec)).digest()
            mism += hd != hr
            digest.update(hr)
            kinds[enc.frame_kind(i)] += 1
        nbytes += sum(map(len, tus))
        done += n
    enc.close()
    return dict(config="C3: %dx%d %d-bit, %d frames, one B200, CRF %d, closed chunks of %d" % (w, h, bd, frames, crf, keyint),
                frames=frames, frames_hashed=done, hash_mismatches=int(mism), decode_matches_recon=bool(mism == 0),
                key_frames=kinds[0], anchor_frames=kinds[1], non_reference_frames=kinds[2], bytes=nbytes,
                kbps_at_60fps=nbytes * 8 * 60.0 / frames / 1000, sha256_of_recon_hashes=digest.hexdigest(),
                seconds=dict(generate=round(t_gen, 2), encode_incl_debug_downloads=round(t_enc, 2), dav1d_decode=round(t_dec, 2)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1200)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--bd", type=int, default=10)
    ap.add_argument("--crf", type=int, default=30)
    ap.add_argument("--keyint", type=int, default=240)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    res = run(w, h, a.bd, a.frames, a.crf, a.keyint)
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    sys.exit(0 if res["decode_matches_recon"] else 1)


if __name__ == "__main__":
    main()
