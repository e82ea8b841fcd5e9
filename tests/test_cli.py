"""The av1an-compatible executable (the boundary the daemon execs: av1an.rs:79-139, startup.rs:98-116)."""
import os, subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "av1_base_b200", "av1an")


def write_y4m(path, frames, bd, fps=(30, 1)):
    h, w = frames[0][0].shape
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F%d:%d Ip A1:1 C%s\n" % (w, h, fps[0], fps[1], "420p10" if bd > 8 else "420jpeg")).encode())
        for fr in frames:
            f.write(b"FRAME\n")
            for p in fr:
                f.write(p.astype("<u2").tobytes() if bd > 8 else p.astype(np.uint8).tobytes())


def test_version_exits_zero():
    r = subprocess.run([CLI, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and "av1b200" in r.stdout


def test_bad_arguments_exit_nonzero(tmp_path):
    assert subprocess.run([CLI], capture_output=True).returncode == 2
    assert subprocess.run([CLI, "-i", "a", "-o", "b", "--encoder", "x264"], capture_output=True).returncode == 2
    assert subprocess.run([CLI, "-i", "a", "-o", "b", "--frobnicate"], capture_output=True).returncode == 2


def test_argv_contract_holds_no_forbidden_hardware_substring():
    """The daemon's startup policy rejects argv that contains any of FORBIDDEN_HW_FLAGS (startup.rs:13-15: nvenc, qsv, vaapi,
    cuda, amf, vce, qsvenc).  Nothing the daemon would put on our command line -- the executable's name, the flags of
    build_av1an_command (av1an.rs:79-107) and the flags our --help lists -- contains one."""
    forbidden = ("nvenc", "qsv", "vaapi", "cuda", "amf", "vce", "qsvenc")
    help_text = subprocess.run([CLI, "--help"], capture_output=True, text=True).stdout
    flags = [w.strip("[]()") for w in help_text.split() if w.lstrip("[").startswith("-")]
    argv = [os.path.basename(CLI), "-i", "-o", "--encoder", "svt-av1", "--pix-format", "yuv420p10le", "--video-params",
            "--crf 8 --preset 3 --film-grain 20 --enable-qm 1 --qm-min 1 --qm-max 15 --keyint 240 --lookahead 40",
            "--audio-params", "-c:a copy", "--workers", "--temp"] + flags
    assert len(flags) >= 8
    for a in argv:
        assert not any(f in a.lower() for f in forbidden), a


def test_gpu_plan_without_a_gpu_is_an_error():
    """`av1an --gpu-plan` (what replaces concurrency.rs:67-84) answers only where there is a GPU; exit 4 keeps the
    daemon on its core-count plan."""
    from av1_base_b200 import abi
    if abi.lib().av1b_device_count() > 0:
        pytest.skip("a GPU is visible")
    r = subprocess.run([CLI, "--gpu-plan"], capture_output=True, text=True)
    assert r.returncode == 4 and "no CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_gpu_plan():
    import json
    from av1_base_b200 import abi
    r = subprocess.run([CLI, "--gpu-plan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    plan = json.loads(r.stdout)
    n = abi.lib().av1b_device_count()
    assert plan["gpus"] == n and plan["av1an_workers"] == n and plan["max_concurrent_jobs"] == n
    assert plan["host_cores"] >= 1 and plan["host_threads_per_gpu"] == max(1, plan["host_cores"] // n)


def read_ebml_size(b, i):
    first = b[i]
    n = 1
    while n <= 8 and not (first & (0x80 >> (n - 1))):
        n += 1
    v = first & (0xFF >> n)
    for k in range(1, n):
        v = (v << 8) | b[i + k]
    return v, i + n


def mkv_blocks(data):
    """Walks our Matroska file: returns (codec_id, codec_private, [block payloads])."""
    out, info = [], {}

    def walk(b, lo, hi):
        i = lo
        while i < hi:
            first = b[i]
            n = 1
            while not (first & (0x80 >> (n - 1))):
                n += 1
            eid = int.from_bytes(b[i:i + n], "big")
            size, j = read_ebml_size(b, i + n)
            if eid in (0x18538067, 0x1654AE6B, 0xAE, 0x1F43B675):
                walk(b, j, j + size)
            elif eid == 0x86:
                info["codec"] = bytes(b[j:j + size]).decode()
            elif eid == 0x63A2:
                info["private"] = bytes(b[j:j + size])
            elif eid == 0xA3:
                out.append(bytes(b[j + 4:j + size]))
            i = j + size
    walk(data, 0, len(data))
    return info.get("codec"), info.get("private"), out


def mkv_video_fields(data):
    """Unsigned elements of the Video master of our Matroska file: {element id: value}."""
    out = {}

    def walk(b, lo, hi, inside):
        i = lo
        while i < hi:
            first = b[i]
            n = 1
            while not (first & (0x80 >> (n - 1))):
                n += 1
            eid = int.from_bytes(b[i:i + n], "big")
            size, j = read_ebml_size(b, i + n)
            if eid in (0x18538067, 0x1654AE6B, 0xAE, 0xE0):
                walk(b, j, j + size, eid == 0xE0)
            elif inside:
                out[eid] = int.from_bytes(b[j:j + size], "big")
            i = j + size
    walk(data, 0, len(data), False)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bd", [(202, 132, 10), (197, 131, 8)])
def test_cli_codes_sizes_that_are_not_multiples_of_8(tmp_path, w, h, bd):
    """Scope crops (1920x804, 3840x1606, ...) are ordinary daemon jobs: the picture is coded padded to multiples of 8 by edge
    replication, the frame headers carry the source size as render_size and the Matroska track the padding as PixelCrop."""
    from oracle import decoders as D
    from tests.test_oracle_chain import unaligned_clip
    n = 6
    src, padded, cw, ch = unaligned_clip(w, h, bd, n, seed=w)
    y4m = str(tmp_path / "in.y4m")
    write_y4m(y4m, src, bd)
    out = str(tmp_path / "out.mkv")
    r = subprocess.run([CLI, "-i", y4m, "-o", out, "--encoder", "svt-av1", "--pix-format", "yuv420p10le" if bd > 8 else "yuv420p",
                        "--video-params", "--crf 30 --preset 6 --keyint 240", "--workers", "1", "--temp", str(tmp_path / "tmp"), "--quiet"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    data = open(out, "rb").read()
    v = mkv_video_fields(data)
    assert (v[0xB0], v[0xBA]) == (cw, ch) and (v[0x54B0], v[0x54BA]) == (w, h)
    assert v.get(0x54DD, 0) == cw - w and v.get(0x54AA, 0) == ch - h
    codec, priv, blocks = mkv_blocks(data)
    tus = [b"\x12\x00" + b for b in blocks]
    assert codec == "V_AV1" and len(tus) == n
    assert [D.render_size_in_tu(t) for t in tus] == [(w, h)] * n
    for dec in (D.dav1d_decode(tus), D.aom_decode(tus)):
        assert len(dec) == n
        for i in range(n):
            assert dec[i][0].shape == (ch, cw) and dec[i][1].shape == (ch // 2, cw // 2)
            assert D.psnr(dec[i][0][:h, :w], src[i][0], bd) > 30
            # the padding is the edge, coded like any other part of the picture
            assert D.psnr(dec[i][0], padded[i][0], bd) > 30


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_cli_encodes_chunks_and_containers(tmp_path, bd):
    from av1_base_b200 import encoder, synth
    from oracle import decoders as D
    w, h, n = 200, 136, 7
    frames = synth.synth_clip(w, h, bd, n, seed=11, scene_len=100)
    y4m = str(tmp_path / "in.y4m")
    write_y4m(y4m, frames, bd)
    outs = {}
    for ext, workers in ((".obu", 1), (".ivf", 2), (".mkv", 3)):
        out = str(tmp_path / ("out" + ext))
        tmp = str(tmp_path / ("tmp" + ext))
        env = dict(os.environ, AV1B_SHARE_GPU="1")
        r = subprocess.run([CLI, "-i", y4m, "-o", out, "--encoder", "svt-av1", "--pix-format", "yuv420p10le" if bd > 8 else "yuv420p",
                            "--video-params", "--crf 32 --preset 6 --keyint 3 --lookahead 40 --film-grain 0",
                            "--audio-params", "-c:a copy", "--workers", str(workers), "--temp", tmp],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert os.path.getsize(out) > 0 and os.path.exists(os.path.join(tmp, "progress.json"))
        assert not os.path.exists(out + ".part")
        outs[ext] = open(out, "rb").read()
    # the same encode through the library API: chunks of keyint frames, concatenated
    enc = encoder.Encoder(w, h, bd, crf=32, keyint=3, fps=(30, 1))
    want = []
    for c in range(0, n, 3):
        want += enc.encode_chunk(frames[c:c + 3])
    enc.close()
    assert outs[".obu"] == b"".join(want)
    # IVF: 32-byte header, then (size, pts) + TU
    ivf, tus, i = outs[".ivf"], [], 32
    assert ivf[:4] == b"DKIF" and ivf[8:12] == b"AV01"
    while i < len(ivf):
        sz = int.from_bytes(ivf[i:i + 4], "little")
        tus.append(ivf[i + 12:i + 12 + sz])
        i += 12 + sz
    assert tus == want
    codec, priv, blocks = mkv_blocks(outs[".mkv"])
    assert codec == "V_AV1" and priv[0] == 0x81 and len(blocks) == n
    assert [b"\x12\x00" + b for b in blocks] == want          # blocks = temporal units without the delimiter
    # FFmpeg's Matroska / IVF demuxers (libavformat inside the OpenCV wheel) read the same packets out of the files
    from tests.test_mux import demux
    props, pk = demux(str(tmp_path / "out.mkv"))
    assert (props["w"], props["h"], props["n"], props["fourcc"]) == (w, h, n, b"AV01") and abs(props["fps"] - 30) < 0.01
    assert [b"\x12\x00" + b for b in pk] == want
    assert demux(str(tmp_path / "out.ivf"))[1] == want
    dec = D.dav1d_decode(want)
    assert len(dec) == n
    for i in range(n):
        mse = np.mean((dec[i][0].astype(np.float64) - frames[i][0]) ** 2)
        psnr = 10 * np.log10(((1 << bd) - 1) ** 2 / mse)
        assert psnr > 30, (i, psnr)


@pytest.mark.gpu
def test_cli_failure_leaves_no_output(tmp_path):
    out = str(tmp_path / "o.mkv")
    r = subprocess.run([CLI, "-i", str(tmp_path / "missing.y4m"), "-o", out], capture_output=True)
    assert r.returncode == 3 and not os.path.exists(out)
    bad = str(tmp_path / "bad.y4m")
    open(bad, "wb").write(b"YUV4MPEG2 W64 H64 F30:1 C420jpeg\nFRAME\n" + b"\0" * 100)
    r = subprocess.run([CLI, "-i", bad, "-o", out], capture_output=True)
    assert r.returncode != 0 and not os.path.exists(out)


@pytest.mark.gpu
def test_cli_splits_chunks_at_scene_cuts(tmp_path):
    """E0: a hard scene cut starts a new closed GOP (chunk) even though --keyint is far away."""
    from av1_base_b200 import encoder, synth
    w, h, bd, n = 200, 136, 10, 9
    frames = synth.synth_clip(w, h, bd, n, seed=21, scene_len=4)        # cuts before frames 4 and 8
    y4m = str(tmp_path / "in.y4m")
    write_y4m(y4m, frames, bd)
    out = str(tmp_path / "out.obu")
    r = subprocess.run([CLI, "-i", y4m, "-o", out, "--video-params", "--crf 32 --keyint 240", "--workers", "2",
                        "--min-scene-len", "2", "--quiet"], capture_output=True, text=True, env=dict(os.environ, AV1B_SHARE_GPU="1"))
    assert r.returncode == 0, r.stderr
    enc = encoder.Encoder(w, h, bd, crf=32, keyint=240)
    want = enc.encode_chunk(frames[0:4]) + enc.encode_chunk(frames[4:8]) + enc.encode_chunk(frames[8:9])
    assert open(out, "rb").read() == b"".join(want)
    # without detection the whole clip is one chunk
    out2 = str(tmp_path / "out2.obu")
    r = subprocess.run([CLI, "-i", y4m, "-o", out2, "--video-params", "--crf 32 --keyint 240", "--no-scene-detection", "--quiet"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out2, "rb").read() == b"".join(enc.encode_chunk(frames))
    enc.close()


@pytest.mark.gpu
def test_cli_progress_fields(tmp_path):
    """f-1: progress.json carries the JobMetrics fields the reference leaves at zero (metrics.rs:12-30, job_executor.rs:117-137)."""
    import json
    from av1_base_b200 import synth
    w, h, bd, n = 200, 136, 10, 10
    frames = synth.synth_clip(w, h, bd, n, seed=5, scene_len=100)
    y4m = str(tmp_path / "in.y4m")
    write_y4m(y4m, frames, bd)
    out, tmp = str(tmp_path / "o.mkv"), str(tmp_path / "t")
    r = subprocess.run([CLI, "-i", y4m, "-o", out, "--video-params", "--crf 30 --keyint 5", "--temp", tmp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pr = json.load(open(os.path.join(tmp, "progress.json")))
    assert pr["done"] is True and pr["frames_encoded"] == n and pr["total_frames"] == n and pr["progress"] == 1.0
    size_kbps = os.path.getsize(out) * 8 / 1000 * 30 / n
    assert 0.7 * size_kbps < pr["bitrate_kbps"] <= size_kbps           # payload of the packets; the file adds the container (SeekHead, Cues, block headers)
    assert 30 < pr["psnr"] < 70 and 0.8 < pr["ssim"] <= 1.0 and pr["est_remaining_secs"] == 0
    # the chunks' packet files are gone, only progress.json stays under --temp
    assert sorted(os.listdir(tmp)) == ["progress.json"]
    # the same numbers from a decode of the stream
    from oracle import decoders as D
    _, _, blocks = mkv_blocks(open(out, "rb").read())
    dec = D.dav1d_decode([b"\x12\x00" + b for b in blocks])
    ps = [10 * np.log10(1023.0 ** 2 / np.mean((dec[i][0].astype(np.float64) - frames[i][0]) ** 2)) for i in range(n)]
    assert abs(np.mean(ps) - pr["psnr"]) < 0.01


def _fake_ffmpeg(dirpath, body):
    p = os.path.join(dirpath, "ffmpeg")
    open(p, "w").write("#!/bin/bash\n[ \"$1\" = \"-version\" ] && exit 0\n" + body)
    os.chmod(p, 0o755)
    return dict(os.environ, PATH=dirpath + ":" + os.environ["PATH"])


@pytest.mark.gpu
def test_cli_failed_decoder_or_audio_copy_is_a_failed_job(tmp_path):
    """A decode pipe that dies mid-stream or an audio copy that fails must not leave a file at -o
    (the daemon replaces the source with whatever it finds there: replacer.rs / size_gate.rs)."""
    from av1_base_b200 import synth
    w, h, bd, n = 200, 136, 8, 4
    frames = synth.synth_clip(w, h, bd, n, seed=3, scene_len=100)
    y4m = str(tmp_path / "src.y4m")
    write_y4m(y4m, frames, bd)
    src = str(tmp_path / "in.mp4")                      # not Y4M: goes through "ffmpeg"
    open(src, "wb").write(b"not a real mp4")
    out = str(tmp_path / "o.mkv")
    bindir = str(tmp_path / "bin"); os.mkdir(bindir)
    # 1. decoder delivers whole frames, then dies with a non-zero status
    env = _fake_ffmpeg(bindir, 'cat "%s"; exit 1\n' % y4m)
    r = subprocess.run([CLI, "-i", src, "-o", out, "--quiet"], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and not os.path.exists(out) and not os.path.exists(out + ".part"), r.stderr
    # 2. decoder fine, audio mux (the call that has two -i) fails
    env = _fake_ffmpeg(bindir, 'n=0; for a in "$@"; do [ "$a" = "-i" ] && n=$((n+1)); done\n'
                               'if [ $n -ge 2 ]; then exit 1; fi\ncat "%s"\n' % y4m)
    r = subprocess.run([CLI, "-i", src, "-o", out, "--quiet", "--audio-params", "-c:a copy"], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and not os.path.exists(out) and not os.path.exists(out + ".part"), r.stderr
    # 3. both fine (the mux "copies" the video-only file): success
    env = _fake_ffmpeg(bindir, 'n=0; for a in "$@"; do [ "$a" = "-i" ] && n=$((n+1)); last="$a"; done\n'
                               'if [ $n -ge 2 ]; then cp "$5" "$last"; exit 0; fi\ncat "%s"\n' % y4m)
    r = subprocess.run([CLI, "-i", src, "-o", out, "--quiet", "--audio-params", "-c:a copy"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and os.path.getsize(out) > 0, r.stderr


@pytest.mark.gpu
def test_cli_rejects_y4m_with_frame_parameters(tmp_path):
    bad = str(tmp_path / "p.y4m")
    with open(bad, "wb") as f:
        f.write(b"YUV4MPEG2 W64 H64 F30:1 C420jpeg\n")
        for _ in range(2):
            f.write(b"FRAME Ip\n" + b"\x80" * (64 * 64 * 3 // 2))
    out = str(tmp_path / "o.obu")
    r = subprocess.run([CLI, "-i", bad, "-o", out, "--quiet"], capture_output=True, text=True)
    assert r.returncode != 0 and not os.path.exists(out)


@pytest.mark.gpu
def test_cli_with_the_daemons_exact_command_line(tmp_path):
    """argv exactly as build_av1an_command makes it (av1an.rs:79-107) with SVT_PARAMS as they stand (av1an.rs:14) and the
    daemon's `.mkv` output (jobs.rs:187): exit 0, a Matroska file FFmpeg's demuxer reads, every frame decodes, the quality is what
    CRF 8 stands for, the progress file ends with done = true and the JobMetrics fields."""
    import json
    from av1_base_b200 import synth
    from oracle import decoders as D
    from tests.test_mux import demux
    w, h, bd, n = 328, 248, 10, 12
    frames = synth.synth_clip(w, h, bd, n, seed=9, scene_len=100, noise=0.5)
    y4m = str(tmp_path / "in.y4m")
    write_y4m(y4m, frames, bd)
    out, tmp = str(tmp_path / "out.mkv"), str(tmp_path / "chunks")
    svt_params = "--crf 8 --preset 3 --film-grain 20 --enable-qm 1 --qm-min 1 --qm-max 15 --keyint 240 --lookahead 40"
    r = subprocess.run([CLI, "-i", y4m, "-o", out, "--encoder", "svt-av1", "--pix-format", "yuv420p10le", "--video-params", svt_params,
                        "--audio-params", "-c:a copy", "--workers", "8", "--temp", tmp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    props, pk = demux(out)
    assert (props["w"], props["h"], props["n"]) == (w, h, n) and len(pk) == n
    tus = [b"\x12\x00" + p for p in pk]
    dec = D.dav1d_decode(tus)
    assert len(dec) == n
    for i in range(n):
        assert D.psnr(dec[i][0], frames[i][0], bd) > 42, i
    prog = json.load(open(os.path.join(tmp, "progress.json")))
    assert prog["done"] is True and prog["frames_encoded"] == n and prog["total_frames"] == n
    assert prog["bitrate_kbps"] > 0 and prog["psnr"] > 42 and 0.9 < prog["ssim"] <= 1.0
    assert not os.path.exists(out + ".part")


@pytest.mark.gpu
def test_cli_film_grain_and_lookahead_reach_the_encoder(tmp_path):
    """Row f-4 through the drop-in executable: `--film-grain 20` (av1an.rs:14) makes the stream carry film grain parameters
    (a decoder that applies grain gives other pictures than one that does not; without the flag both agree) and
    `--lookahead 0` changes the stream (the temporal filter of the key picture may no longer look ahead); `--enable-qm 1
    --qm-min 1 --qm-max 15` codes with quantisation matrices."""
    from av1_base_b200 import synth
    from oracle import decoders as D
    w, h, bd, n = 328, 248, 10, 8
    frames = synth.synth_clip(w, h, bd, n, seed=5, scene_len=100, noise=1.0)
    y4m = str(tmp_path / "in.y4m")
    write_y4m(y4m, frames, bd)
    outs = {}
    for name, vp in (("plain", "--crf 40 --preset 6 --keyint 240"), ("grain", "--crf 40 --preset 6 --keyint 240 --film-grain 20"),
                     ("nolook", "--crf 40 --preset 6 --keyint 240 --lookahead 0"),
                     ("qm", "--crf 40 --preset 6 --keyint 240 --enable-qm 1 --qm-min 1 --qm-max 15")):
        out = str(tmp_path / (name + ".obu"))
        r = subprocess.run([CLI, "-i", y4m, "-o", out, "--encoder", "svt-av1", "--pix-format", "yuv420p10le", "--video-params", vp,
                            "--workers", "1", "--temp", str(tmp_path / ("tmp_" + name)), "--quiet"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[name] = open(out, "rb").read()
    def tus_of(stream):   # walk the OBUs (header, optional extension byte, leb128 size): a temporal delimiter starts a unit
        starts, i = [], 0
        while i < len(stream):
            hdr = stream[i]
            if (hdr >> 3) & 15 == 2:
                starts.append(i)
            j = i + 1 + ((hdr >> 2) & 1)
            size, shift = 0, 0
            while True:
                b = stream[j]; j += 1
                size |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            i = j + size
        return [stream[a:b] for a, b in zip(starts, starts[1:] + [len(stream)])]
    for name in ("plain", "grain"):
        tus = tus_of(outs[name])
        assert len(tus) == n
        a = D.dav1d_decode(tus)
        b = D.dav1d_decode(tus, apply_grain=True)
        assert len(a) == n and len(b) == n
        same = all(np.array_equal(a[i][0], b[i][0]) for i in range(n))
        assert same == (name == "plain"), name
    assert outs["nolook"] != outs["plain"]
    # --enable-qm 1 --qm-min 1 --qm-max 15 (av1an.rs:14): quantisation matrices are signalled and used; both decoders agree on the pictures
    assert outs["qm"] != outs["plain"]
    tus = tus_of(outs["qm"])
    a, b = D.dav1d_decode(tus), D.aom_decode(tus)
    assert len(a) == n and len(b) == n and all(np.array_equal(a[i][p], b[i][p]) for i in range(n) for p in range(3))
