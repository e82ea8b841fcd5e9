"""Test infrastructure (NOT product code): normative-decoder oracle.

Two independent AV1 decoders bundled in this image are driven through ctypes:
  * dav1d 1.5.3   (exported from Pillow's libavif)
  * libaom 3.13.1 (OpenCV wheel; decoder and, for the CPU baseline, encoder)
A stream produced by the B200 encoder is accepted only if BOTH decoders reproduce the encoder's
own reconstruction bit for bit.  ABI offsets follow SURVEY.md section 8c (verified in this image).
"""
import ctypes, os
import numpy as np
from . import aomsym

_aom = None
_dav1d = None

def aom():
    global _aom
    if _aom is None:
        _aom = ctypes.CDLL(aomsym.find_lib(aomsym.LIBAOM_GLOB))
        _aom.aom_codec_av1_dx.restype = ctypes.c_void_p
        _aom.aom_codec_av1_cx.restype = ctypes.c_void_p
        _aom.aom_codec_get_frame.restype = ctypes.c_void_p
        _aom.aom_codec_get_cx_data.restype = ctypes.c_void_p
        _aom.aom_img_alloc.restype = ctypes.c_void_p
        _aom.aom_codec_error_detail.restype = ctypes.c_char_p
        _aom.aom_codec_error.restype = ctypes.c_char_p
        _aom.aom_codec_err_to_string.restype = ctypes.c_char_p
    return _aom

def dav1d():
    global _dav1d
    if _dav1d is None:
        _dav1d = ctypes.CDLL(aomsym.find_lib(aomsym.LIBAVIF_GLOB))
        _dav1d.dav1d_data_create.restype = ctypes.c_void_p
        _dav1d.dav1d_version.restype = ctypes.c_char_p
    return _dav1d

def _img_to_planes(img_ptr):
    """aom_image_t -> [Y,U,V] uint16 arrays cropped to d_w x d_h."""
    raw = (ctypes.c_uint8 * 160).from_address(img_ptr)
    u32 = np.frombuffer(raw, dtype=np.uint32, count=40)
    fmt = int(u32[0]); d_w = int(u32[10]); d_h = int(u32[11])
    xs = int(u32[14]); ys = int(u32[15])
    planes = np.frombuffer(raw, dtype=np.uint64, count=3, offset=64)
    strides = np.frombuffer(raw, dtype=np.int32, count=3, offset=88)
    hbd = bool(fmt & 0x800)
    out = []
    for p in range(3):
        w = d_w if p == 0 else (d_w + xs) >> xs
        h = d_h if p == 0 else (d_h + ys) >> ys
        st = int(strides[p])
        buf = (ctypes.c_uint8 * (st * h)).from_address(int(planes[p]))
        a = np.frombuffer(buf, dtype=np.uint8).reshape(h, st)
        if hbd:
            a = a.view(np.uint16)[:, :w]
        else:
            a = a[:, :w].astype(np.uint16)
        out.append(np.array(a, dtype=np.uint16, order="C", copy=True))   # the decoder reuses / frees its buffers
    return out

def aom_decode(temporal_units):
    """temporal_units: list of bytes (low-overhead OBU TUs). Returns list of [Y,U,V] per shown frame."""
    A = aom()
    ctx = ctypes.create_string_buffer(256)
    rc = A.aom_codec_dec_init_ver(ctx, ctypes.c_void_p(A.aom_codec_av1_dx()), None, 0, 22)
    if rc:
        raise RuntimeError("aom dec init %d" % rc)
    frames = []
    try:
        for tu in temporal_units:
            rc = A.aom_codec_decode(ctx, tu, ctypes.c_size_t(len(tu)), None)
            if rc:
                det = A.aom_codec_error_detail(ctx)
                raise RuntimeError("aom_codec_decode rc=%d: %s / %s" % (
                    rc, A.aom_codec_error(ctx), det))
            it = ctypes.c_void_p(0)
            while True:
                img = A.aom_codec_get_frame(ctx, ctypes.byref(it))
                if not img:
                    break
                frames.append(_img_to_planes(img))
    finally:
        A.aom_codec_destroy(ctx)
    return frames

def dav1d_decode(temporal_units, apply_grain=False):
    """apply_grain: let dav1d add the film grain a stream signals (the reconstruction checks decode without it)."""
    D = dav1d()
    settings = ctypes.create_string_buffer(512)
    D.dav1d_default_settings(settings)
    s32 = ctypes.cast(settings, ctypes.POINTER(ctypes.c_int32))
    s32[0] = 1; s32[1] = 1; s32[2] = 1 if apply_grain else 0      # n_threads, max_frame_delay, apply_grain
    c = ctypes.c_void_p(0)
    rc = D.dav1d_open(ctypes.byref(c), settings)
    if rc:
        raise RuntimeError("dav1d_open %d" % rc)
    frames = []

    def drain():
        while True:
            pic = ctypes.create_string_buffer(2048)
            r = D.dav1d_get_picture(c, pic)
            if r < 0:
                return r
            raw = np.frombuffer(pic, dtype=np.uint8)
            data = np.frombuffer(pic, dtype=np.uint64, count=3, offset=16)
            stride = np.frombuffer(pic, dtype=np.int64, count=2, offset=40)
            w, h, layout, bpc = np.frombuffer(pic, dtype=np.int32, count=4, offset=56)
            assert layout == 1, layout
            out = []
            for p in range(3):
                pw = int(w) if p == 0 else (int(w) + 1) >> 1
                ph = int(h) if p == 0 else (int(h) + 1) >> 1
                st = int(stride[0 if p == 0 else 1])
                buf = (ctypes.c_uint8 * (st * ph)).from_address(int(data[p]))
                a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, st)
                if bpc > 8:
                    a = a.view(np.uint16)[:, :pw]
                else:
                    a = a[:, :pw].astype(np.uint16)
                out.append(np.ascontiguousarray(a).copy())
            frames.append(out)
            D.dav1d_picture_unref(pic)

    try:
        for tu in temporal_units:
            d = ctypes.create_string_buffer(256)
            p = D.dav1d_data_create(d, ctypes.c_size_t(len(tu)))
            if not p:
                raise RuntimeError("dav1d_data_create")
            ctypes.memmove(p, tu, len(tu))
            while True:
                r = D.dav1d_send_data(c, d)
                if r == -11:
                    drain()
                    continue
                if r < 0:
                    raise RuntimeError("dav1d_send_data rc=%d" % r)
                break
            r = drain()
            if r not in (-11,):
                raise RuntimeError("dav1d_get_picture rc=%d" % r)
        # flush
        drain()
    finally:
        D.dav1d_close(ctypes.byref(c))
    return frames

def aom_encode(frames, bit_depth, cq_level=30, cpu_used=6, threads=1, lag=0, tile_cols_log2=0,
               tile_rows_log2=0, kf_max_dist=None, extra=()):
    """CPU baseline / stream source: libaom encoder, constant-quality mode. frames: list of [Y,U,V]
    uint16 arrays (4:2:0). Returns list of temporal units (bytes)."""
    A = aom()
    h, w = frames[0][0].shape
    iface = ctypes.c_void_p(A.aom_codec_av1_cx())
    cfg = ctypes.create_string_buffer(4096)
    rc = A.aom_codec_enc_config_default(iface, cfg, 0)
    assert rc == 0, rc
    c32 = ctypes.cast(cfg, ctypes.POINTER(ctypes.c_uint32))
    c32[1] = threads          # g_threads
    c32[3] = w; c32[4] = h
    c32[8] = bit_depth; c32[9] = bit_depth
    c32[10] = 1; c32[11] = 30  # timebase
    c32[14] = lag             # g_lag_in_frames
    c32[24] = 3               # rc_end_usage = AOM_Q
    hbd = bit_depth > 8
    ctx = ctypes.create_string_buffer(256)
    rc = A.aom_codec_enc_init_ver(ctx, iface, cfg, 0x40000 if hbd else 0, 25)
    if rc:
        raise RuntimeError("aom enc init %d %s" % (rc, A.aom_codec_err_to_string(rc)))
    opts = [("cpu-used", str(cpu_used)), ("cq-level", str(cq_level)), ("row-mt", "1"),
            ("tile-columns", str(tile_cols_log2)), ("tile-rows", str(tile_rows_log2))] + list(extra)
    if kf_max_dist is not None:
        opts.append(("kf-max-dist", str(kf_max_dist)))
    for k, v in opts:
        rc = A.aom_codec_set_option(ctx, k.encode(), v.encode())
        if rc:
            raise RuntimeError("set_option %s=%s rc=%d" % (k, v, rc))
    fmt = 0x102 | (0x800 if hbd else 0)
    img = A.aom_img_alloc(None, fmt, w, h, 32)
    raw = (ctypes.c_uint8 * 160).from_address(img)
    planes = np.frombuffer(raw, dtype=np.uint64, count=3, offset=64)
    strides = np.frombuffer(raw, dtype=np.int32, count=3, offset=88)
    tus = []

    def pull():
        it = ctypes.c_void_p(0)
        while True:
            pkt = A.aom_codec_get_cx_data(ctx, ctypes.byref(it))
            if not pkt:
                break
            kind = ctypes.c_int.from_address(pkt).value
            if kind != 0:
                continue
            buf = ctypes.c_void_p.from_address(pkt + 8).value
            sz = ctypes.c_size_t.from_address(pkt + 16).value
            tus.append(ctypes.string_at(buf, sz))
    try:
        for i, fr in enumerate(frames):
            for p in range(3):
                ph, pw = fr[p].shape
                st = int(strides[p])
                dst = (ctypes.c_uint8 * (st * ph)).from_address(int(planes[p]))
                d = np.frombuffer(dst, dtype=np.uint8).reshape(ph, st)
                if hbd:
                    d.view(np.uint16)[:, :pw] = fr[p]
                else:
                    d[:, :pw] = fr[p].astype(np.uint8)
            rc = A.aom_codec_encode(ctx, ctypes.c_void_p(img), ctypes.c_int64(i), ctypes.c_ulong(1), ctypes.c_long(0))
            if rc:
                raise RuntimeError("aom_codec_encode rc=%d %s" % (rc, A.aom_codec_error_detail(ctx)))
            pull()
        while True:
            n0 = len(tus)
            rc = A.aom_codec_encode(ctx, None, ctypes.c_int64(0), ctypes.c_ulong(1), ctypes.c_long(0))
            pull()
            if len(tus) == n0:
                break
    finally:
        A.aom_img_free(ctypes.c_void_p(img))
        A.aom_codec_destroy(ctx)
    return tus

class AomStream:
    """libaom encoder driven frame by frame (CPU baseline of bench.py: frames are pushed while a clock runs; packets
    come out lag_in_frames later).  Constant-quality mode, as aom_encode."""

    def __init__(self, w, h, bit_depth, cq_level=30, cpu_used=6, threads=1, lag=19, tile_cols_log2=0, tile_rows_log2=0):
        A = aom()
        self.A, self.hbd = A, bit_depth > 8
        iface = ctypes.c_void_p(A.aom_codec_av1_cx())
        cfg = ctypes.create_string_buffer(4096)
        rc = A.aom_codec_enc_config_default(iface, cfg, 0)
        assert rc == 0, rc
        c32 = ctypes.cast(cfg, ctypes.POINTER(ctypes.c_uint32))
        c32[1] = threads
        c32[3] = w; c32[4] = h
        c32[8] = bit_depth; c32[9] = bit_depth
        c32[10] = 1; c32[11] = 30
        c32[14] = lag
        c32[24] = 3               # rc_end_usage = AOM_Q
        self.ctx = ctypes.create_string_buffer(256)
        rc = A.aom_codec_enc_init_ver(self.ctx, iface, cfg, 0x40000 if self.hbd else 0, 25)
        if rc:
            raise RuntimeError("aom enc init %d" % rc)
        for k, v in (("cpu-used", cpu_used), ("cq-level", cq_level), ("row-mt", 1), ("tile-columns", tile_cols_log2),
                     ("tile-rows", tile_rows_log2)):
            rc = A.aom_codec_set_option(self.ctx, k.encode(), str(v).encode())
            if rc:
                raise RuntimeError("set_option %s rc=%d" % (k, rc))
        self.img = A.aom_img_alloc(None, 0x102 | (0x800 if self.hbd else 0), w, h, 32)
        raw = (ctypes.c_uint8 * 160).from_address(self.img)
        self.planes = np.frombuffer(raw, dtype=np.uint64, count=3, offset=64)
        self.strides = np.frombuffer(raw, dtype=np.int32, count=3, offset=88)
        self.pts, self.packets, self.bytes_out = 0, 0, 0

    def _pull(self):
        it = ctypes.c_void_p(0)
        while True:
            pkt = self.A.aom_codec_get_cx_data(self.ctx, ctypes.byref(it))
            if not pkt:
                break
            if ctypes.c_int.from_address(pkt).value == 0:
                self.packets += 1
                self.bytes_out += ctypes.c_size_t.from_address(pkt + 16).value

    def push(self, fr):
        """Hands one frame ([Y,U,V] uint16) to the encoder; returns the number of packets produced so far."""
        for p in range(3):
            ph, pw = fr[p].shape
            st = int(self.strides[p])
            dst = (ctypes.c_uint8 * (st * ph)).from_address(int(self.planes[p]))
            d = np.frombuffer(dst, dtype=np.uint8).reshape(ph, st)
            if self.hbd:
                d.view(np.uint16)[:, :pw] = fr[p]
            else:
                d[:, :pw] = fr[p].astype(np.uint8)
        rc = self.A.aom_codec_encode(self.ctx, ctypes.c_void_p(self.img), ctypes.c_int64(self.pts), ctypes.c_ulong(1), ctypes.c_long(0))
        if rc:
            raise RuntimeError("aom_codec_encode rc=%d" % rc)
        self.pts += 1
        self._pull()
        return self.packets

    def flush(self):
        while True:
            n0 = self.packets
            self.A.aom_codec_encode(self.ctx, None, ctypes.c_int64(0), ctypes.c_ulong(1), ctypes.c_long(0))
            self._pull()
            if self.packets == n0:
                break
        return self.packets

    def close(self):
        if self.img:
            self.A.aom_img_free(ctypes.c_void_p(self.img))
            self.A.aom_codec_destroy(self.ctx)
            self.img = None


def psnr(a, b, bit_depth):
    a = a.astype(np.float64); b = b.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return 100.0
    peak = (1 << bit_depth) - 1
    return 10 * np.log10(peak * peak / mse)


def obus(tu):
    """(type, payload) of every OBU of a temporal unit (low-overhead format with size fields)."""
    out, i = [], 0
    while i < len(tu):
        hdr = tu[i]
        j = i + 1 + ((hdr >> 2) & 1)
        size, shift = 0, 0
        while True:
            b = tu[j]; j += 1
            size |= (b & 0x7F) << shift; shift += 7
            if not b & 0x80:
                break
        out.append(((hdr >> 3) & 15, bytes(tu[j:j + size])))
        i = j + size
    return out


def render_size_in_tu(tu):
    """render_size() of the frame header in a temporal unit as THIS encoder writes headers (spec 5.9.2 with no order hints,
    no frame ids, no superres, frame_size_override_flag = 0): (render_width, render_height), or None when the header says
    render_and_frame_size_different = 0.  Test-side bit reader; the decoders check everything behind these bits."""
    for typ, pl in obus(tu):
        if typ not in (3, 6):
            continue
        bits = "".join("{:08b}".format(b) for b in pl[:16])
        assert bits[0] == "0"                        # show_existing_frame
        frame_type, show = int(bits[1:3], 2), bits[3]
        if frame_type == 0 and show == "1":
            pos = 6                                  # disable_cdf_update, frame_size_override_flag
        else:
            assert frame_type == 1                   # error_resilient_mode, disable_cdf_update, frame_size_override_flag,
            pos = 7 + 3 + 8 + 21                     # primary_ref_frame, refresh_frame_flags, ref_frame_idx[7]
        if bits[pos] == "0":
            return None
        return int(bits[pos + 1:pos + 17], 2) + 1, int(bits[pos + 17:pos + 33], 2) + 1
    raise ValueError("no frame header in the temporal unit")
