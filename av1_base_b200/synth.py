"""Deterministic synthetic YUV 4:2:0 clips: per scene a smooth gradient, textured
rectangles translating at constant sub-pel velocity, a global pan and band-limited noise; chroma is a
low-passed function of luma.  Values stay in video range.  Used by bench.py and the tests."""
import numpy as np


def _smooth_noise(rng, h, w, sigma):
    n = rng.standard_normal((h // 4 + 2, w // 4 + 2))
    n = np.kron(n, np.ones((4, 4)))[:h, :w]
    return n * sigma


def synth_clip(width, height, bit_depth, n_frames, seed=1, scene_len=80, hdr=False, noise=1.0, start=0):
    """Returns a list of [Y, U, V] uint16 arrays: frames start .. start + n_frames - 1 of the clip (every frame is a pure
    function of (seed, frame number), so slices of a clip can be made independently)."""
    scale = 1 << (bit_depth - 8)
    lo, hi = 16 * scale, 235 * scale
    frames = []
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    scene = None
    for f in range(start, start + n_frames):
        sc = f // scene_len
        if scene is None or scene["id"] != sc:
            rng = np.random.default_rng(seed * 1000 + sc)
            nrect = int(rng.integers(3, 7))
            scene = dict(id=sc, rng=rng, gx=rng.uniform(-0.08, 0.08), gy=rng.uniform(-0.08, 0.08),
                         base=rng.uniform(60, 160), pan=rng.uniform(-3, 3, 2),
                         rects=[dict(x=rng.uniform(0, width), y=rng.uniform(0, height),
                                     w=rng.uniform(width / 12, width / 3), h=rng.uniform(height / 12, height / 3),
                                     vx=rng.uniform(-24, 24) / 4, vy=rng.uniform(-24, 24) / 4,
                                     lum=rng.uniform(30, 220), fx=rng.uniform(0.05, 0.6), fy=rng.uniform(0.05, 0.6),
                                     amp=rng.uniform(4, 30)) for _ in range(nrect)])
        t = f - sc * scene_len
        px, py = scene["pan"] * t
        img = scene["base"] + scene["gx"] * (xx + px) + scene["gy"] * (yy + py)
        img = img + 10 * np.sin((xx + px) * 0.021) * np.cos((yy + py) * 0.017)
        for r in scene["rects"]:
            x0 = (r["x"] + r["vx"] * t) % width
            y0 = (r["y"] + r["vy"] * t) % height
            m = (xx >= x0) & (xx < x0 + r["w"]) & (yy >= y0) & (yy < y0 + r["h"])
            tex = r["lum"] + r["amp"] * np.sin((xx - x0) * r["fx"]) * np.sin((yy - y0) * r["fy"])
            img = np.where(m, tex, img)
        nrng = np.random.default_rng(seed * 100000 + f)
        img = img + (_smooth_noise(nrng, height, width, 1.5) + nrng.standard_normal((height, width)) * 1.0) * noise
        if hdr:   # PQ-like: compress most codes into the lower half, sparse highlights
            img = 16 + (np.clip(img, 16, 235) - 16) ** 1.35 / (219 ** 0.35)
        Y = np.clip(np.rint(img * scale), lo, hi).astype(np.uint16)
        ch, cw = (height + 1) // 2, (width + 1) // 2
        sub = img[0:ch * 2:2, 0:cw * 2:2]
        U = np.clip(np.rint((128 + 0.25 * (sub - 128) + 6 * np.sin(xx[:ch, :cw] * 0.05)) * scale), lo, 240 * scale)
        V = np.clip(np.rint((128 - 0.18 * (sub - 128) + 6 * np.cos(yy[:ch, :cw] * 0.04)) * scale), lo, 240 * scale)
        frames.append([Y, U.astype(np.uint16), V.astype(np.uint16)])
    return frames
