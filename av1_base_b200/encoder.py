"""Host-side mirror of the reference's encode-job interface over the C ABI.

Reference interface mirrored (same names / meaning / error behaviour where they exist):
  /root/reference/crates/daemon/src/encode/av1an.rs:36-62   Av1anEncodeParams  -> EncodeParams
  /root/reference/crates/daemon/src/encode/av1an.rs:18-30   EncodeError        -> EncodeError
  /root/reference/crates/daemon/src/encode/av1an.rs:126-139 run_av1an          -> Encoder.encode_chunk
numpy is plumbing only; every pixel operation happens in libav1b200.so on the GPU.
"""
import ctypes as C
import numpy as np
from . import abi


class EncodeError(RuntimeError):
    """Mirror of EncodeError::{Av1anFailed(code), Io}: carries the negative C-ABI return code."""

    def __init__(self, code, msg):
        super().__init__("av1b200 failed with code %d: %s" % (code, msg))
        self.code = code


def _check(rc):
    if rc != 0:
        raise EncodeError(rc, abi.last_error())


class Encoder:
    def __init__(self, width, height, bit_depth=10, crf=30, preset=6, keyint=240, fps=(30, 1), device_id=0,
                 tile_cols_log2=-1, tile_rows_log2=-1, hdr=False, host_threads=0, frames_in_flight=0,
                 keep_debug=False, blk_log2=0, loop_filters=True, intra_only=False, tb_zero_thr=0, raster_levels=False,
                 pack_path=0, lr_off=False, tile_sb=0, gop_period=0, me_smooth=True, key_var_part=True, mctf=True,
                 lookahead=-1, film_grain=0, scene_cut=True, qm=None):
        L = abi.lib()
        cfg = abi.Config()
        L.av1b_config_default(C.byref(cfg))
        cfg.width, cfg.height, cfg.bit_depth = width, height, bit_depth
        cfg.crf, cfg.preset, cfg.keyint = crf, preset, keyint
        cfg.fps_num, cfg.fps_den = fps
        cfg.device_id = device_id
        cfg.tile_cols_log2, cfg.tile_rows_log2 = tile_cols_log2, tile_rows_log2
        cfg.hdr = int(hdr)
        cfg.host_threads = host_threads
        cfg.frames_in_flight = frames_in_flight
        cfg.reserved[0] = int(keep_debug)
        cfg.reserved[1] = blk_log2
        cfg.reserved[2] = 0 if loop_filters else 1
        cfg.reserved[3] = 1 if intra_only else 0
        cfg.reserved[4] = tb_zero_thr
        # inter-frame entropy coding path: 0 device tokenizer, range coder where it is faster (default), 3 device tokenizer +
        # host range coder, 4 device tokenizer + device range coder, 1 host block walker over raster levels, 2 host block walker over in-place packed symbols
        cfg.reserved[5] = 1 if raster_levels else pack_path
        cfg.reserved[6] = int(lr_off)
        cfg.reserved[7] = tile_sb          # inter-frame tile size in superblocks (0 = default)
        cfg.gop_period = gop_period        # 0 = default (6, or a P chain where the quantiser codes the noise), 1 = plain P chain
        cfg.tune[0] = 0 if me_smooth else 1
        cfg.tune[1] = 0 if key_var_part else 1
        cfg.tune[2] = 0 if mctf else 1     # temporal filter of key / anchor sources
        cfg.tune[4] = 0 if scene_cut else 1   # key frame at a scene change inside a chunk (scores from the GPU)
        cfg.lookahead = lookahead          # -1 = default; frames the temporal filter may look ahead
        cfg.film_grain = film_grain
        if qm is not None:                 # (qm_min, qm_max): --enable-qm 1 --qm-min .. --qm-max ..
            cfg.enable_qm, cfg.qm_min, cfg.qm_max = 1, qm[0], qm[1]
        self.cfg = cfg
        self._h = C.c_void_p()
        _check(L.av1b_encoder_create(C.byref(cfg), C.byref(self._h)))
        self.geom = abi.Geom()
        _check(L.av1b_get_geom(self._h, C.byref(self.geom)))

    def close(self):
        if self._h:
            abi.lib().av1b_encoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode_chunk_strided(self, frames, progress=None):
        """Like encode_chunk, but row-strided uint16 views are handed over as they are (plane pointer + stride)."""
        return self.encode_chunk(frames, progress, _strided=True)

    def encode_chunk(self, frames, progress=None, _strided=False):
        """frames: list of [Y,U,V] uint16 arrays. Returns the list of temporal units (bytes)."""
        n = len(frames)
        srcs = (abi.FrameSrc * n)()
        keep = []
        for i, fr in enumerate(frames):
            for p in range(3):
                a = fr[p]
                if not (_strided and a.dtype == np.uint16 and a.strides[1] == 2 and a.strides[0] % 2 == 0):
                    a = np.ascontiguousarray(a, dtype=np.uint16)
                keep.append(a)
                srcs[i].planes[p] = a.ctypes.data
                srcs[i].stride[p] = a.strides[0] // 2
        out = []

        def on_packet(user, data, size, idx, is_key):
            out.append(C.string_at(data, size))
            return 0

        def on_progress(user, done, total, fps):
            if progress:
                progress(done, total, fps)

        cb = abi.PACKET_CB(on_packet)
        pcb = abi.PROGRESS_CB(on_progress)
        _check(abi.lib().av1b_encode_chunk(self._h, srcs, n, cb, pcb, None))
        return out

    def recon(self, frame):
        g = self.geom
        planes = [np.zeros((g.height, g.width), np.uint16), np.zeros((g.height // 2, g.width // 2), np.uint16),
                  np.zeros((g.height // 2, g.width // 2), np.uint16)]
        dst = (C.c_void_p * 3)(*[p.ctypes.data for p in planes])
        st = (C.c_int32 * 3)(*[p.shape[1] for p in planes])
        _check(abi.lib().av1b_get_recon(self._h, frame, dst, st))
        return planes

    def frame_syms(self, frame):
        g = self.geom
        blocks = np.zeros(g.w8 * g.h8, abi.BLOCK_INFO_DTYPE)
        coef = [np.zeros((g.rows[p], g.stride[p]), np.int16) for p in range(3)]
        cp = (C.c_void_p * 3)(*[c.ctypes.data for c in coef])
        _check(abi.lib().av1b_get_frame_syms(self._h, frame, blocks.ctypes.data_as(C.c_void_p), cp))
        return blocks, coef

    def stats(self):
        s = (C.c_double * 24)()
        _check(abi.lib().av1b_get_stats(self._h, s, 24))
        return dict(h2d_ms=s[0], kernel_ms=s[1], d2h_ms=s[2], pack_ms=s[3], kernel_launches=int(s[4]),
                    base_q_idx=int(s[5]), intra_ms=s[6], intra_launches=int(s[7]), frames_done=int(s[8]),
                    bytes_out=int(s[9]), deblock_ms=s[10], cdef_ms=s[11], inter_ms=s[12], me_ms=s[13],
                    inter_launches=int(s[14]), key_frames=int(s[15]), staged_direct=int(s[16]), tok_ms=s[17],
                    tokens=int(s[18]), d2h_bytes=int(s[19]), lr_ms=s[20], rc_ms=s[21], mctf_ms=s[22], mctf_frames=int(s[23]))

    def lr_units(self, frame):
        """Luma restoration units [rows, cols] of a kept frame (preset <= 5, keep_debug=True)."""
        r, c = C.c_int32(0), C.c_int32(0)
        _check(abi.lib().av1b_get_lr_units(self._h, frame, None, C.byref(r), C.byref(c)))
        u = np.zeros((r.value, c.value), abi.LR_UNIT_DTYPE)
        _check(abi.lib().av1b_get_lr_units(self._h, frame, u.ctypes.data_as(C.c_void_p), None, None))
        return u

    def frame_params(self):
        fp = abi.FrameParams()
        _check(abi.lib().av1b_get_frame_params(self._h, C.byref(fp)))
        return fp

    def inter_frame_params(self):
        fp = abi.FrameParams()
        _check(abi.lib().av1b_get_inter_frame_params(self._h, C.byref(fp)))
        return fp

    def class_params(self, kind):
        """Frame-level parameters of a frame kind: 0 key, 1 anchor, 2 non-reference."""
        fp = abi.FrameParams()
        _check(abi.lib().av1b_get_class_params(self._h, kind, C.byref(fp)))
        return fp

    def chunk_info(self):
        """Structure of the chunk coded last (av1b_get_chunk_info)."""
        v = (C.c_int32 * 8)()
        _check(abi.lib().av1b_get_chunk_info(self._h, v))
        return dict(gop_period=v[0], q_key=v[1], q_anchor=v[2], q_nonref=v[3], mctf=bool(v[4]), noise_b=v[5], q_nominal=v[6], auto=bool(v[7]))

    def frame_kind(self, pos):
        return int(abi.lib().av1b_get_frame_kind(self._h, C.c_int64(pos)))

    def me_lambda(self):
        """Vector-deviation cost the encoder hands to the motion search (SAD units)."""
        return int(abi.lib().av1b_get_me_lambda(self._h))

    def frame_is_key(self, frame):
        rc = abi.lib().av1b_get_frame_is_key(self._h, frame)
        if rc < 0:
            raise EncodeError(rc, "frame not kept")
        return bool(rc)

    def cdef_idx(self, frame):
        g = self.geom
        idx = np.zeros(g.sb_rows * g.sb_cols, np.uint8)
        _check(abi.lib().av1b_get_cdef_idx(self._h, frame, idx.ctypes.data_as(C.c_void_p)))
        return idx

    def _srcs(self, frames):
        n = len(frames)
        srcs = (abi.FrameSrc * n)()
        keep = []
        for i, fr in enumerate(frames):
            for p in range(3):
                a = np.ascontiguousarray(fr[p], dtype=np.uint16)
                keep.append(a)
                srcs[i].planes[p] = a.ctypes.data
                srcs[i].stride[p] = a.shape[1]
        return srcs, keep

    def stage_frames(self, slot, frames):
        """Upload frames into device slot 0/1 (inputs resident in HBM for encode_resident)."""
        srcs, keep = self._srcs(frames)
        _check(abi.lib().av1b_stage_frames(self._h, slot, srcs, len(frames)))

    def stage_clip(self, frames):
        """Upload a clip into HBM once (resident clip); encode_clip then codes closed chunks made of its pictures."""
        srcs, keep = self._srcs(frames)
        _check(abi.lib().av1b_stage_clip(self._h, srcs, len(frames)))

    def encode_clip(self, order, accumulate=False, collect=False):
        """One closed chunk: frame i = resident clip picture order[i]."""
        out = []

        def on_packet(user, data, size, idx, is_key):
            out.append(C.string_at(data, size))
            return 0

        cb = abi.PACKET_CB(on_packet) if collect else None
        o = np.ascontiguousarray(order, np.uint32)
        _check(abi.lib().av1b_encode_clip(self._h, o.ctypes.data_as(C.c_void_p), len(o), int(accumulate), cb, None))
        return out

    def encode_resident(self, n_steps, collect=False):
        out = []

        def on_packet(user, data, size, idx, is_key):
            if collect:
                out.append(C.string_at(data, size))
            return 0

        cb = abi.PACKET_CB(on_packet) if collect else None   # no per-packet trip through the interpreter when nobody listens
        _check(abi.lib().av1b_encode_resident(self._h, n_steps, cb, None))
        return out


class PinnedFrames:
    """Source frames held in page-locked host memory (av1b_host_alloc): the encoder's copy engine reads them
    where they lie instead of going through its pageable-source staging copy.  Behaves like a list of
    [Y, U, V] uint16 arrays; the memory is released by close() / garbage collection."""

    def __init__(self, frames, device_id=0):
        self._ptr = None
        total = sum(int(np.asarray(pl).size) for fr in frames for pl in fr) * 2
        ptr = abi.lib().av1b_host_alloc(device_id, total)
        if not ptr:
            raise EncodeError(-1, "av1b_host_alloc(%d)" % total)
        self._ptr = ptr
        buf = (C.c_uint8 * total).from_address(ptr)
        flat = np.frombuffer(buf, dtype=np.uint16)
        self.frames, off = [], 0
        for fr in frames:
            planes = []
            for pl in fr:
                pl = np.asarray(pl, dtype=np.uint16)
                v = flat[off:off + pl.size].reshape(pl.shape)
                v[...] = pl
                planes.append(v)
                off += pl.size
            self.frames.append(planes)

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        return self.frames[i]

    def __iter__(self):
        return iter(self.frames)

    def close(self):
        if self._ptr:
            self.frames = []
            abi.lib().av1b_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    return abi.lib().av1b_device_count()


def version():
    b = C.create_string_buffer(128)
    abi.lib().av1b_version(b, 128)
    return b.value.decode()
