#!/usr/bin/env python3
"""bench.py -- AV1 encode fps of the B200 backend (BASELINE.json metric), one rank per GPU.

A "step" is two closed chunks of the workload's clip (C4: 2 x 150 frames of 3840x2160 10-bit, a chunk = one scene of the
scene_len-150 generator = what the chunker cuts; 1 key frame + 149 inter frames each) through the hot path, one chunk stream per GPU
(chunks share no state, so ranks never communicate on the data path; weak scaling).
  value : frames/s with the clip resident in HBM (av1b_stage_clip + av1b_encode_clip): device kernels + symbol
          download + host entropy coding, pipelined; whole job over all ranks.
  e2e   : frames/s through av1b_encode_chunk with HOST buffers in page-locked memory (H2D -> kernels ->
          D2H -> entropy coding -> packets), the call a reference-side binding makes.
  roofline : the dominant kernel by device time, algorithmic bytes / CUDA-event duration
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline : libaom 3.13.1 cpu-used=6 (the stand-in SURVEY.md 8d names for the reference's av1an + SVT-AV1, which
             cannot run in this image) on the same clip, av1an style: nproc/2 chunk workers; bounded sample.
  bd_rate : Bjontegaard rate difference of this encoder against libaom cpu-used=6 on the 960x544 clip of tools/bdrate.py.
--impl reference times that libaom stand-in alone (no product library is loaded), streaming the same chunk through
nproc/2 encoder instances at steady state.
"""
import argparse, json, os, statistics, subprocess, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, bit_depth, description, scene_len of the generator, frames of one chunk); a step is
    # CHUNKS_PER_STEP closed chunks, so that the driver's 20 steps give a timed region of more than 2 s
    "4k10": (3840, 2160, 10, "C4: 3840x2160 10-bit HDR 4:2:0 synthetic, CRF 30, one chunk stream per GPU", 150, 150),
    "1080p10": (1920, 1080, 10, "C3: 1920x1080 10-bit 4:2:0 synthetic, CRF 30", 240, 240),
    "1080p8": (1920, 1080, 8, "C1: 1920x1080 8-bit 4:2:0 synthetic, CRF 30", 80, 80),
}


CHUNKS_PER_STEP = 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region: NVML in a thread every few milliseconds
    (the timed region of a short run is shorter than one nvidia-smi start-up), nvidia-smi -lms as a fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, uuid=None):
        self.idx = gpu_index
        self.uuid = uuid
        self.proc = None
        self.lines = []
        self.samples = []      # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.nvml = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:      # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by UUID
                    h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(self.uuid))
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml = (pynvml, h)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((mhz, mask))
            except Exception:
                pass
            self.stop_flag.wait(0.004)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.th.join(timeout=1)
            sm = [s[0] for s in self.samples]
            reasons = set()
            for _, mask in self.samples:
                for bit, nm in self.BITS.items():
                    if mask & bit:
                        reasons.add(nm)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def _gen(args):
    w, h, bd, seed, scene_len, hdr, start = args
    from av1_base_b200 import synth
    return synth.synth_clip(w, h, bd, 1, seed=seed, scene_len=scene_len, hdr=hdr, start=start)[0]


def make_clip(w, h, bd, n, hdr, scene_len):
    """The first n frames of the workload's clip (SURVEY.md 8d generator, seed 4), generated by the host cores in parallel."""
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(w, h, bd, 4, scene_len, hdr, i) for i in range(n)]
    with ProcessPoolExecutor(min(len(jobs), max(1, (os.cpu_count() or 2) // 2), 16)) as ex:
        return list(ex.map(_gen, jobs))


def chunk_order(n_distinct, n_frames):
    """A closed chunk of n_frames made of n_distinct pictures: forwards, then backwards without repeating the end points
    (continuous motion, no frame equals its predecessor)."""
    period = list(range(n_distinct)) + list(range(n_distinct - 2, 0, -1))
    return [period[i % len(period)] for i in range(n_frames)]


def workload_config(args):
    w, h, bd, desc, scene_len, chunk = WORKLOADS[args.workload]
    return {"workload": desc, "clip": "synth_clip seed 4, scene_len %d" % scene_len, "frames_per_step": CHUNKS_PER_STEP * chunk,
            "chunk": "%d closed chunks (scenes) per step, each 1 key frame + %d inter frames" % (CHUNKS_PER_STEP, chunk - 1),
            "distinct_frames": args.distinct, "crf": args.crf, "preset": args.preset, "keyint": args.keyint}


LIBAOM_LAG = 19


def libaom_stream_fps(frames, order, w, h, bd, cq, workers, threads, fpw, warm_steps, steps):
    """av1an-style CPU stand-in: `workers` libaom encoder instances (cpu-used 6, constant quality, `threads` threads each)
    each code the same closed chunk; every step hands fpw more frames to every worker.  Before the clock starts each
    worker is fed until its first packet appears (lag_in_frames + 1 frames: the key frame and the look-ahead are behind it)
    and warm_steps more steps: the timed steps run at steady state, one frame coded per frame handed in.
    Returns (fps, seconds, seconds per step list)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import decoders as D
    tiles = (1, 1) if w > 2000 else (1, 0)
    encs = [D.AomStream(w, h, bd, cq_level=cq, cpu_used=6, threads=threads, lag=LIBAOM_LAG, tile_cols_log2=tiles[0],
                        tile_rows_log2=tiles[1]) for _ in range(workers)]
    pos = [0] * workers

    def feed(k, n, until_packet=False):
        e = encs[k]
        for _ in range(n):
            got = e.push(frames[order[pos[k] % len(order)]])
            pos[k] += 1
            if until_packet and got > 0:
                break
        return n

    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(lambda k: feed(k, LIBAOM_LAG + 8, True), range(workers)))
        for _ in range(warm_steps):
            list(ex.map(lambda k: feed(k, fpw), range(workers)))
        per = []
        t0 = time.perf_counter()
        for _ in range(steps):
            ts = time.perf_counter()
            list(ex.map(lambda k: feed(k, fpw), range(workers)))
            per.append(time.perf_counter() - ts)
        dt = time.perf_counter() - t0
    for e in encs:
        e.close()
    return steps * workers * fpw / dt, dt, per


def libaom_plan(w):
    cores = os.cpu_count() or 2
    workers = max(1, cores // 2)
    threads = max(1, cores // workers)
    fpw = 2 if w > 2000 else 6        # frames per worker and step: a bounded sample (a step of the 4K clip takes seconds)
    return cores, workers, threads, fpw


def run_reference(args, rank, world):
    """CPU arm: rank 0 only.  Loads libaom and numpy only: no product library, no device."""
    if rank != 0:
        return
    w, h, bd, desc, scene_len, chunk = WORKLOADS[args.workload]
    cores, workers, threads, fpw = libaom_plan(w)
    nd = args.distinct
    frames = make_clip(w, h, bd, nd, args.workload == "4k10", scene_len)
    order = chunk_order(nd, chunk)
    fps, dt, per = libaom_stream_fps(frames, order, w, h, bd, args.crf, workers, threads, fpw, args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": "AV1 encode fps", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "stand-in-libaom",
                         "encoder": "libaom 3.13.1 cpu-used=6, end-usage=q, cq-level=%d, lag_in_frames=%d, row-mt, %d threads per worker" % (args.crf, LIBAOM_LAG, threads),
                         "workers": workers,
                         "sample": "%d workers (av1an --workers style, nproc/2) each stream the step's closed chunk through their own encoder: "
                                   "%d frames per worker and step at steady state (the key frame and the look-ahead fill are before the clock), "
                                   "%d timed frames in all; av1an + SVT-AV1 itself cannot run in this image (BASELINE.md section 2)"
                                   % (workers, fpw, args.steps * workers * fpw)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "svt_av1_fps": None,
    }
    print(json.dumps(line), flush=True)


def port_fps(frames, w, h, bd, crf, threads):
    """Second CPU figure: the scalar oracle restatement of THIS encoder's path (oracle/chain.py), one 2-frame chunk per thread."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import chain
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: chain.encode_chain(frames[k % (len(frames) - 1):][:2], w, h, bd, crf), range(threads)))
    return 2 * threads / (time.perf_counter() - t0)


def bd_rate_check(device_id):
    """Rate/quality of this encoder against libaom cpu-used=6 on the 960x544 10-bit clip of tools/bdrate.py (same run, same
    clip): five points each, PSNR-Y on the dav1d-decoded streams, Bjontegaard cubic fit; every B200 stream must also
    decode to the encoder's reconstruction."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bdrate import bd_rate, bd_rate_pchip
    from av1_base_b200 import encoder, synth
    from oracle import decoders as D
    w, h, bd, n = 960, 544, 10, 30
    frames = synth.synth_clip(w, h, bd, n, seed=4, scene_len=1000, noise=1.0)

    def point(tus):
        dec = D.dav1d_decode(tus)
        return sum(map(len, tus)) * 8 * 30.0 / n / 1000, float(np.mean([D.psnr(d[0], f[0], bd) for d, f in zip(dec, frames)])), dec

    ours, match = [], True
    for crf in (6, 10, 14, 20, 28, 36, 44, 52):
        enc = encoder.Encoder(w, h, bd, crf=crf, device_id=device_id, keep_debug=True)
        kbps, ps, dec = point(enc.encode_chunk(frames))
        for i in (0, 1, 4, n - 1):
            match = match and all(np.array_equal(dec[i][p], enc.recon(i)[p]) for p in range(3))
        enc.close()
        ours.append((kbps, ps))
    out = {"clip": "960x544 10-bit, 30 frames, synth seed 4 noise 1.0", "ours": [{"kbps": k, "psnr_y": p} for k, p in ours],
           "decode_matches_recon": bool(match)}
    cores = os.cpu_count() or 1
    for name, lag in (("libaom_cpu6", LIBAOM_LAG), ("libaom_cpu6_lag0", 0)):
        pts = []
        for cq in (24, 32, 40, 48, 56):
            k, p, _ = point(D.aom_encode(frames, bd, cq_level=cq, cpu_used=6, threads=cores, lag=lag))
            pts.append((k, p))
        out[name] = [{"kbps": k, "psnr_y": p} for k, p in pts]
        out["vs_%s_psnr_y_pct" % name] = bd_rate([x[0] for x in pts], [x[1] for x in pts], [x[0] for x in ours], [x[1] for x in ours])
        out["vs_%s_psnr_y_pchip_pct" % name] = bd_rate_pchip([x[0] for x in pts], [x[1] for x in pts], [x[0] for x in ours], [x[1] for x in ours])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k10", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-in-flight", type=int, default=8, help="frames per device batch")
    ap.add_argument("--distinct", type=int, default=30, help="distinct pictures generated for the chunk (it walks them back and forth)")
    ap.add_argument("--crf", type=int, default=30)
    ap.add_argument("--keyint", type=int, default=240)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bd-rate", action="store_true")
    ap.add_argument("--tile-sb", type=int, default=0, help="inter-frame tile size in superblocks (0 = encoder default)")
    ap.add_argument("--pack-path", type=int, default=0, help="0 auto, 3 device tokenizer + host range coder, 4 device tokenizer + device range coder")
    ap.add_argument("--preset", type=int, default=6, help="<= 5 adds loop restoration (the BASELINE configs name preset 6)")
    ap.add_argument("--gop-period", type=int, default=0)
    ap.add_argument("--no-mctf", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from av1_base_b200 import encoder
    if encoder.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the B200 backend has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    w, h, bd, desc, scene_len, chunk = WORKLOADS[args.workload]
    F = args.frames_in_flight
    hdr = args.workload == "4k10"
    frames = make_clip(w, h, bd, args.distinct, hdr, scene_len)
    order = chunk_order(args.distinct, chunk)
    if rank:     # other ranks code other chunks: the same pictures, started elsewhere
        order = [(o + 3 * rank) % args.distinct for o in order]
    from av1_base_b200 import sharding
    # ranks share the node's host cores: each rank entropy-codes with its share of them
    enc = encoder.Encoder(w, h, bd, crf=args.crf, device_id=local_rank, hdr=hdr, frames_in_flight=F, keyint=args.keyint, preset=args.preset, tile_sb=args.tile_sb, pack_path=args.pack_path,
                          host_threads=sharding.host_threads_per_rank(world), gop_period=args.gop_period, mctf=not args.no_mctf)
    g = enc.geom
    frame_bytes = sum(g.stride[p] * (h if p == 0 else h // 2) * 2 for p in range(3))
    S = int(1.5 * w * h * 2)                             # bytes of one 4:2:0 frame at 2 B/sample

    # ---------------- value: the clip resident in HBM ----------------
    enc.stage_clip(frames)
    for _ in range(max(args.warmup, 3)):
        enc.encode_clip(order)
    try:
        dev_uuid = torch.cuda.get_device_properties(local_rank).uuid
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        for c in range(CHUNKS_PER_STEP):
            enc.encode_clip(order, accumulate=(k > 0 or c > 0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = enc.stats()
    barrier()
    clocks = sampler.stop()
    tmax = sharding.max_over_ranks(dt, dist, "cuda")
    value = world * args.steps * CHUNKS_PER_STEP * chunk / tmax

    # ---------------- e2e: host buffers through av1b_encode_chunk ----------------
    # the host copies of the sources live in page-locked memory (av1b_host_alloc), as a capture / decode front
    # end would hand them over; the timed region holds their H2D copies, the kernels, the D2H of the
    # symbol streams, host entropy coding and packet delivery
    pinned = encoder.PinnedFrames(frames, device_id=local_rank)
    host_chunk = [pinned[o] for o in order]
    enc.encode_chunk(host_chunk[:4 * F])                 # warm-up
    e2e_steps = max(1, min(args.steps, 8))
    barrier()
    t0 = time.perf_counter()
    h2d_ms = kern_ms = d2h_ms = pack_ms = 0.0
    d2h_bytes = staged_direct = 0
    for _ in range(e2e_steps * CHUNKS_PER_STEP):
        tus = enc.encode_chunk(host_chunk)
        se = enc.stats()
        h2d_ms += se["h2d_ms"]; kern_ms += se["kernel_ms"]; d2h_ms += se["d2h_ms"]; pack_ms += se["pack_ms"]
        d2h_bytes += se["d2h_bytes"]; staged_direct += se["staged_direct"]
    dte = time.perf_counter() - t0
    dte = sharding.max_over_ranks(dte, dist, "cuda")
    e2e = world * e2e_steps * CHUNKS_PER_STEP * chunk / dte

    if rank != 0:
        return
    pk, pk_src = peaks()
    n_inter, n_key = max(1, st["inter_launches"]), max(1, st["key_frames"])
    nf = st["frames_done"]
    n_batches = args.steps * CHUNKS_PER_STEP * ((chunk + F - 1) // F)
    n_tf = max(1, st["mctf_frames"])
    # CDEF runs on the key and anchor frames only (the non-reference frames signal none)
    n_cdef = max(1, args.steps * CHUNKS_PER_STEP * sum(1 for pos in range(chunk) if enc.frame_kind(pos) != 2))
    # per-kernel CUDA-event time per unit (frame; ME and tokenizer per batch) and algorithmic bytes per unit
    kern = {
        "inter_encode_kernel": (st["inter_ms"] / n_inter, 4 * S),
        "intra_encode_kernel": (st["intra_ms"] / n_key, 3 * S),
        "deblock_kernel": (st["deblock_ms"] / max(1, nf), 2 * S),
        "cdef_kernel": (st["cdef_ms"] / n_cdef, 3 * S),
        "pyramid+hme+regularisation (per batch)": ((st["me_ms"] - st["mctf_ms"]) / n_batches, int((1.3125 + 0.625 + 2.0 + 2 * 2.0) * (w * h * 2)) * F),
        # temporal filter of one key / anchor picture: 4 searches (2.625 Y each) + the filter itself ((2 + 4) S)
        "temporal filter: hme + mctf_kernel (per filtered picture)": (st["mctf_ms"] / n_tf, int(4 * 2.625 * (w * h * 2)) + 6 * S),
        # tokenizer: reads the block info (20 B per 8x8 unit) twice (count + emit), writes 4 B per token
        "tokenizer (per batch)": (st["tok_ms"] / n_batches, 2 * 20 * g.w8 * g.h8 * F + 4 * st["tokens"] // n_batches),
    }
    share = {"inter_encode_kernel": st["inter_ms"], "intra_encode_kernel": st["intra_ms"], "deblock_kernel": st["deblock_ms"],
             "cdef_kernel": st["cdef_ms"], "pyramid+hme+regularisation (per batch)": st["me_ms"] - st["mctf_ms"],
             "temporal filter: hme + mctf_kernel (per filtered picture)": st["mctf_ms"], "tokenizer (per batch)": st["tok_ms"]}
    dom = max(share, key=lambda k: share[k])
    dom_ms, dom_bytes = kern[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    cfg = workload_config(args)
    line = {
        "metric": "AV1 encode fps", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1000 * tmax / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32", "data": "synthetic",
        "config": cfg,
        "encoder": {"base_q_idx_anchor": st["base_q_idx"], "frames_in_flight": F, "key_frames": st["key_frames"], "inter_frames": st["inter_launches"],
                    "temporally_filtered_frames": st["mctf_frames"], "cdef_frames": n_cdef, "gop_period": enc.chunk_info()["gop_period"], "tiles_key_frames": "%dx%d" % (g.tile_cols, g.tile_rows),
                    "tile_sb_inter": args.tile_sb, "pack_path": args.pack_path, "host_threads": sharding.host_threads_per_rank(world),
                    "l2": "inputs larger than L2 (%.0f MB of distinct source pictures, %.0f MB working set per batch)" % (args.distinct * frame_bytes / 1e6, 5 * F * frame_bytes / 1e6),
                    "timing": "wall clock between synchronize+barrier pairs (host entropy coding is part of the step); "
                              "kernel times from CUDA events on the encoder stream"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": CHUNKS_PER_STEP * chunk * frame_bytes, "d2h_bytes_per_step": d2h_bytes // e2e_steps,
                "steps": e2e_steps, "bitrate_bytes_per_frame": sum(map(len, tus)) / len(tus),
                "host_buffers": "page-locked (av1b_host_alloc), %d of %d frames read in place by the copy engine" % (staged_direct, e2e_steps * CHUNKS_PER_STEP * chunk),
                "breakdown_ms_per_step": {"h2d_ms": h2d_ms / e2e_steps, "kernel_ms": kern_ms / e2e_steps, "d2h_ms": d2h_ms / e2e_steps, "pack_ms": pack_ms / e2e_steps}},
        "gpu_launches": st["kernel_launches"],
        "breakdown_ms_per_step": {k: st[k] / args.steps for k in ("kernel_ms", "me_ms", "mctf_ms", "intra_ms", "inter_ms", "deblock_ms",
                                                                   "cdef_ms", "lr_ms", "tok_ms", "rc_ms", "d2h_ms", "pack_ms")},
        "tokens_per_frame": st["tokens"] / max(1, st["inter_launches"]),
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_src,
                     "ms_per_launch": dom_ms, "algorithmic_bytes_per_launch": dom_bytes,
                     "kernels": {k: {"ms_per_launch": v[0], "algorithmic_bytes": v[1],
                                     "frac": (v[1] / (v[0] * 1e-3) / 1e9 / pk["hbm_gbs"]) if v[0] > 0 else None}
                                 for k, v in kern.items()}},
        "clocks": clocks,
        "svt_av1_fps": None,
    }
    pinned.close()
    if world == 1:
        enc.close()
        if not args.no_cpu_baseline:
            cores, workers, threads, fpw = libaom_plan(w)
            steps = 12 if w > 2000 else 8
            fps, dtl, _ = libaom_stream_fps(frames, order, w, h, bd, args.crf, workers, threads, fpw, 1, steps)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "stand-in-libaom", "workers": workers,
                                    "encoder": "libaom 3.13.1 cpu-used=6, end-usage=q, cq-level=%d, lag_in_frames=%d, row-mt, %d threads per worker" % (args.crf, LIBAOM_LAG, threads),
                                    "sample": "%d workers x %d frames x %d steps of the same chunk at steady state (%.1f s)" % (workers, fpw, steps, dtl),
                                    "port": {"value": port_fps(frames, w, h, bd, args.crf, min(cores, 8)), "unit": "frames/s", "cores": min(cores, 8),
                                             "kind": "port", "sample": "oracle/chain.py (scalar restatement of this encoder), one 2-frame chunk per thread"}}
        if not args.no_bd_rate:
            line["bd_rate"] = bd_rate_check(local_rank)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
