"""Third-party validation of the containers the drop-in executable writes (jobs.rs:187 makes `.mkv` the daemon's output):
FFmpeg's own Matroska and IVF demuxers -- libavformat 62 inside the OpenCV wheel of this image, reached through
cv2.VideoCapture in raw mode -- must find the track, its size, frame rate and frame count, and hand back, packet for
packet, the temporal units that went in.  CPU test: the streams are coded by the oracle chain and muxed by the executable's
container writers alone (`av1an --mux-packets`, which needs no GPU); dav1d then decodes the demuxed packets to the oracle's
reconstruction.  (The FFmpeg build in the wheel has no software AV1 decoder, so decoding stays with dav1d / libaom.)"""
import os, struct, subprocess
import numpy as np
import pytest
from av1_base_b200 import synth
from oracle import pyoracle as O, decoders as D, chain
from tests.test_oracle_chain import pack_chain, unaligned_clip

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "av1_base_b200", "av1an")


def write_packet_files(d, tus, keys, per_chunk):
    """What the CLI's workers leave under --temp: chunk_NNNNNN.pkt = records of (size u32 LE, key flag, temporal unit)."""
    os.makedirs(d, exist_ok=True)
    for c in range(0, len(tus), per_chunk):
        with open(os.path.join(d, "chunk_%06d.pkt" % (c // per_chunk)), "wb") as f:
            for i in range(c, min(c + per_chunk, len(tus))):
                f.write(struct.pack("<IB", len(tus[i]), int(keys[i])) + tus[i])


def mux(d, out, cw, ch, w, h, fps, bd):
    r = subprocess.run([CLI, "--mux-packets", d, "-o", out, "--mux-format", "%dx%d,%dx%d,%d:%d,%d" % (cw, ch, w, h, fps[0], fps[1], bd), "--quiet"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(out + ".part")


def demux(path):
    """(properties, packets) as FFmpeg's demuxer sees the file."""
    cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG, [cv2.CAP_PROP_FORMAT, -1])
    assert cap.isOpened(), path
    props = dict(w=int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), h=int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), fps=cap.get(cv2.CAP_PROP_FPS),
                 n=int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), fourcc=int(cap.get(cv2.CAP_PROP_FOURCC)).to_bytes(4, "little"))
    pk = []
    while True:
        ok, p = cap.read()
        if not ok:
            break
        pk.append(p.tobytes())
    cap.release()
    return props, pk


@pytest.mark.parametrize("w,h,bd,fps", [(200, 136, 8, (30, 1)), (328, 248, 10, (24000, 1001))])
def test_ffmpeg_demuxes_our_matroska_and_ivf(tmp_path, w, h, bd, fps):
    n, keyint = 7, 3
    frames = synth.synth_clip(w, h, bd, n, seed=w, scene_len=100, noise=0.3)
    g, want = chain.encode_chain(frames, w, h, bd, 32, keyint=keyint, gop_period=2)
    tus = pack_chain(w, h, bd, want, g)
    d = str(tmp_path / "pk")
    write_packet_files(d, tus, [r.kind == 0 for r in want], keyint)
    for ext in (".mkv", ".ivf"):
        out = str(tmp_path / ("out" + ext))
        mux(d, out, w, h, w, h, fps, bd)
        props, pk = demux(out)
        assert (props["w"], props["h"], props["n"], props["fourcc"]) == (w, h, n, b"AV01"), (ext, props)
        assert abs(props["fps"] - fps[0] / fps[1]) < 0.01, (ext, props)
        assert len(pk) == n
        # Matroska blocks carry the temporal units without the delimiter OBU (the AV1-in-Matroska mapping), IVF frames whole
        assert pk == ([t[2:] for t in tus] if ext == ".mkv" else tus), ext
        dec = D.dav1d_decode([b"\x12\x00" + p if ext == ".mkv" else p for p in pk])
        assert len(dec) == n
        for i in range(n):
            for p in range(3):
                assert np.array_equal(dec[i][p], O.crop(g, want[i].fin)[p]), (ext, i, p)
    # a raw .obu file is the concatenation
    out = str(tmp_path / "out.obu")
    mux(d, out, w, h, w, h, fps, bd)
    assert open(out, "rb").read() == b"".join(tus)


def _elems(b, lo, hi):
    """(id, payload start, payload size, element start) of the EBML elements in b[lo:hi]."""
    i = lo
    while i < hi:
        n = 1
        while not (b[i] & (0x80 >> (n - 1))):
            n += 1
        eid = int.from_bytes(b[i:i + n], "big")
        first, m = b[i + n], 1
        while not (first & (0x80 >> (m - 1))):
            m += 1
        size = int.from_bytes(bytes([first & (0xFF >> m)]) + b[i + n + 1:i + n + m], "big")
        yield eid, i + n + m, size, i
        i += n + m + size


def test_matroska_index_and_seeking(tmp_path):
    """What players seek by: the SeekHead points at Info, Tracks and Cues, every CuePoint at a Cluster that starts with a key
    frame and carries that cluster's timestamp; FFmpeg seeks to the key frames and returns their packets."""
    w, h, bd, n, keyint = 200, 136, 8, 12, 3
    frames = synth.synth_clip(w, h, bd, n, seed=3, scene_len=100, noise=0.3)
    g, want = chain.encode_chain(frames, w, h, bd, 32, keyint=keyint, gop_period=2)
    tus = pack_chain(w, h, bd, want, g)
    d, out = str(tmp_path / "pk"), str(tmp_path / "out.mkv")
    write_packet_files(d, tus, [r.kind == 0 for r in want], keyint)
    mux(d, out, w, h, w, h, (30, 1), bd)
    data = open(out, "rb").read()
    seg = [e for e in _elems(data, 0, len(data)) if e[0] == 0x18538067][0]
    assert seg[1] + seg[2] == len(data)                                   # the patched Segment size
    kids = list(_elems(data, seg[1], seg[1] + seg[2]))
    assert [k[0] for k in kids] == [0x114D9B74, 0x1549A966, 0x1654AE6B] + [0x1F43B675] * (n // keyint) + [0x1C53BB6B]
    targets = {}
    for s in _elems(data, kids[0][1], kids[0][1] + kids[0][2]):           # SeekHead -> Seek { SeekID, SeekPosition }
        f = {e[0]: int.from_bytes(data[e[1]:e[1] + e[2]], "big") for e in _elems(data, s[1], s[1] + s[2])}
        targets[f[0x53AB]] = seg[1] + f[0x53AC]
    assert targets == {k[0]: k[3] for k in kids if k[0] in (0x1549A966, 0x1654AE6B, 0x1C53BB6B)}
    clusters = [k for k in kids if k[0] == 0x1F43B675]
    cues = []
    for cp in _elems(data, kids[-1][1], kids[-1][1] + kids[-1][2]):       # Cues -> CuePoint { CueTime, CueTrackPositions { .. } }
        f = {}
        for e in _elems(data, cp[1], cp[1] + cp[2]):
            if e[0] == 0xB3:
                f["t"] = int.from_bytes(data[e[1]:e[1] + e[2]], "big")
            elif e[0] == 0xB7:
                f.update({q[0]: int.from_bytes(data[q[1]:q[1] + q[2]], "big") for q in _elems(data, e[1], e[1] + e[2])})
        cues.append((f["t"], f[0xF7], seg[1] + f[0xF1]))
    assert cues == [(round(1000 * c * keyint / 30), 1, clusters[c][3]) for c in range(n // keyint)]
    cap = cv2.VideoCapture(out, cv2.CAP_FFMPEG, [cv2.CAP_PROP_FORMAT, -1])
    for k in (6, 3, 9, 0):
        assert cap.set(cv2.CAP_PROP_POS_FRAMES, k)
        ok, p = cap.read()
        assert ok and p.tobytes() == tus[k][2:], k
    cap.release()


def test_ffmpeg_demuxes_a_cropped_track(tmp_path):
    """Sources that are not multiples of 8: PixelCrop / DisplayWidth / DisplayHeight in the track do not disturb the demuxer; the
    track's pixel size is the coded size and the packets come back unchanged."""
    w, h, bd, n = 202, 132, 10, 4
    src, padded, cw, ch = unaligned_clip(w, h, bd, n, seed=1)
    g, want = chain.encode_chain(padded, cw, ch, bd, 34, gop_period=2)
    tus = pack_chain(cw, ch, bd, want, g, render=(w, h))
    d = str(tmp_path / "pk")
    write_packet_files(d, tus, [r.kind == 0 for r in want], n)
    out = str(tmp_path / "out.mkv")
    mux(d, out, cw, ch, w, h, (30, 1), bd)
    props, pk = demux(out)
    assert (props["w"], props["h"], props["n"]) == (cw, ch, n), props
    assert pk == [t[2:] for t in tus]


def test_mux_mode_fails_cleanly(tmp_path):
    r = subprocess.run([CLI, "--mux-packets", str(tmp_path / "none"), "-o", str(tmp_path / "o.mkv"), "--mux-format", "64x64,64x64,30:1,8"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and not os.path.exists(str(tmp_path / "o.mkv"))
    r = subprocess.run([CLI, "--mux-packets", str(tmp_path), "-o", str(tmp_path / "o.mkv"), "--mux-format", "nonsense"], capture_output=True, text=True)
    assert r.returncode == 2
