#!/usr/bin/env python3
"""Encodes a few synthetic frames once (for ncu captures).  Usage: tools/encode_once.py [4k|1080p] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from av1_base_b200 import encoder, synth
size = sys.argv[1] if len(sys.argv) > 1 else "4k"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w, h = (3840, 2160) if size == "4k" else (1920, 1080)
frames = synth.synth_clip(w, h, 10, n, seed=4, scene_len=100, hdr=(size == "4k"))
enc = encoder.Encoder(w, h, 10, crf=30, frames_in_flight=n)
tus = enc.encode_chunk(frames)
print("encoded", len(tus), "frames", sum(map(len, tus)), "bytes", enc.stats())
