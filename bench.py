#!/usr/bin/env python3
"""bench.py -- AV1 encode fps of the B200 backend (BASELINE.json metric), one rank per GPU.

A "step" is one pass of the hot path over one batch of F synthetic frames (one chunk stream per GPU,
SURVEY.md 8e: chunks share no state, so ranks never communicate on the data path; weak scaling).
  value : frames/s with the source frames already resident in HBM (av1b_encode_resident): device
          kernels + symbol download + host entropy coding, pipelined; whole job over all ranks.
  e2e   : frames/s through av1b_encode_chunk with HOST buffers in page-locked memory (H2D -> kernels ->
          D2H -> entropy coding -> packets), the call a reference-side binding makes.
  roofline : the dominant kernel by device time, algorithmic bytes / CUDA-event duration
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle port of the same path on the host cores (bounded sample).
--impl reference times the CPU path (oracle port; the reference's av1an + SVT-AV1 cannot run here,
BASELINE.md section 2) with all host threads on the same workload.
"""
import argparse, json, os, statistics, subprocess, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, bit_depth, description)
    "4k10": (3840, 2160, 10, "C4: 3840x2160 10-bit HDR 4:2:0 synthetic, CRF 30, one chunk stream per GPU"),
    "1080p10": (1920, 1080, 10, "C3: 1920x1080 10-bit 4:2:0 synthetic, CRF 30"),
    "1080p8": (1920, 1080, 8, "C1: 1920x1080 8-bit 4:2:0 synthetic, CRF 30"),
}


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


def make_frames(w, h, bd, n, hdr):
    from av1_base_b200 import synth
    # two scenes so that the batch is not one static picture; deterministic
    return synth.synth_clip(w, h, bd, n, seed=4, scene_len=max(1, n // 2), hdr=hdr)


def abi_tables_acq(bd, qidx):
    """AC quantiser step of a qindex (read from the generated table header: no device needed)."""
    import re
    txt = open(os.path.join(ROOT, "av1_base_b200", "csrc", "av1_tables.h")).read()
    m = re.search(r"av1t_ac_q_%d\[\d+\] = \{(.*?)\};" % bd, txt, re.S)
    return [int(v) for v in re.findall(r"-?\d+", m.group(1))][qidx]


def oracle_chunk(frames, w, h, bd, qidx):
    """CPU port of the same path (oracle/av1_oracle.cpp) for one chunk: key frame, then inter frames
    (hierarchical ME on the source pyramid, motion-compensated residual coding), deblock + CDEF
    decision + CDEF after every frame.  Scalar code, one thread per chunk."""
    from av1_base_b200 import abi
    from oracle import pyoracle as O
    import ctypes as C
    g = O.geom(w, h, 0, 0)
    pm = O.partition_fixed(g, 4)
    lam = (abi_tables_acq(bd, qidx)) >> 1
    fps = []
    for ft in (0, 1):
        fp = abi.FrameParams()
        abi.lib().av1b_select_frame_params(bd, qidx if ft else max(1, qidx * 3 // 4), ft, 1, C.byref(fp))
        fps.append(fp)
    prev_fin = prev_pyr = None
    for i, fr in enumerate(frames):
        pyr = O.pyramid(g, O.pad_planes(g, fr)[0])
        if i == 0:
            r, fp = O.encode_intra_frame(g, fr, bd, max(1, qidx * 3 // 4), pm), fps[0]
        else:
            r, fp = O.encode_inter_frame(g, fr, bd, qidx, pm, O.hme(g, pyr, prev_pyr, lam), prev_fin), fps[1]
            O.merge_skip_blocks(g, r.blocks)
        O.deblock_frame(g, bd, r.blocks, r.rec, list(fp.lf_level), fp.lf_sharpness)
        src = O.pad_planes(g, fr)
        prev_fin = O.cdef_frame(g, bd, r.blocks, fp, O.cdef_search(g, bd, r.blocks, fp, r.rec, src), r.rec)
        prev_pyr = pyr
    return len(frames)


def oracle_fps(frames, w, h, bd, qidx, threads, chunk_len):
    """threads independent chunks of chunk_len frames each (av1an-style chunk parallelism; ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    chunks = [[frames[(c + i) % len(frames)] for i in range(chunk_len)] for c in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        n = sum(ex.map(lambda ch: oracle_chunk(ch, w, h, bd, qidx), chunks))
    return n / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """CPU arm: rank 0 only."""
    if rank != 0:
        return
    w, h, bd, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    from av1_base_b200 import abi  # noqa: F401  (tables only; no device use)
    qidx = 120   # CRF 30 -> quantizer_to_qindex[30]
    workers = max(1, min(cores, 32))
    chunk_len = 2 if w > 2000 else 3   # bounded sample: key + inter frame(s) per worker and step
    per_step = workers * chunk_len
    frames = make_frames(w, h, bd, 4, args.workload == "4k10")
    for _ in range(min(args.warmup, 1)):
        oracle_fps(frames, w, h, bd, qidx, workers, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_fps(frames, w, h, bd, qidx, workers, chunk_len)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": "AV1 encode fps", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step": per_step, "crf": 30, "chunk_len": chunk_len},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": workers, "kind": "port",
                         "sample": "%d steps x %d chunks of %d frames (key + inter) of the workload, one chunk per host thread, "
                                   "scalar oracle/av1_oracle.cpp (av1an+SVT-AV1 itself cannot run in this image: BASELINE.md "
                                   "section 2)" % (args.steps, workers, chunk_len)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "svt_av1_fps": None,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k10", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=8)
    ap.add_argument("--crf", type=int, default=30)
    ap.add_argument("--keyint", type=int, default=240)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile-sb", type=int, default=0, help="inter-frame tile size in superblocks (0 = encoder default)")
    ap.add_argument("--pack-path", type=int, default=0, help="0 auto, 3 device tokenizer + host range coder, 4 device tokenizer + device range coder")
    ap.add_argument("--preset", type=int, default=6, help="<= 5 adds loop restoration (the BASELINE configs name preset 6)")
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

    w, h, bd, desc = WORKLOADS[args.workload]
    F = args.frames_per_step
    hdr = args.workload == "4k10"
    frames = make_frames(w, h, bd, F, hdr)               # F consecutive frames of one scene
    if rank:
        frames = frames[rank % len(frames):] + frames[:rank % len(frames)]
    from av1_base_b200 import sharding
    # ranks share the node's host cores: each rank entropy-codes with its share of them
    enc = encoder.Encoder(w, h, bd, crf=args.crf, device_id=local_rank, hdr=hdr, frames_in_flight=F, keyint=args.keyint, preset=args.preset, tile_sb=args.tile_sb, pack_path=args.pack_path,
                          host_threads=sharding.host_threads_per_rank(world))
    g = enc.geom
    frame_bytes = sum(g.stride[p] * (h if p == 0 else h // 2) * 2 for p in range(3))
    S = int(1.5 * w * h * 2)                             # bytes of one 4:2:0 frame at 2 B/sample

    # ---------------- value: inputs resident in HBM ----------------
    # slot 0 = frames 0..F-1, slot 1 = the same frames in reverse order: the resident sequence is a
    # palindrome (0..F-1, F-1..0, 0..F-1, ...), i.e. continuous motion without artificial scene cuts
    enc.stage_frames(0, frames)
    enc.stage_frames(1, frames[::-1])
    enc.encode_resident(max(args.warmup, 3))
    try:
        dev_uuid = torch.cuda.get_device_properties(local_rank).uuid
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    enc.encode_resident(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = enc.stats()
    barrier()
    clocks = sampler.stop()
    tmax = sharding.max_over_ranks(dt, dist, "cuda")
    value = world * args.steps * F / tmax

    # ---------------- e2e: host buffers through av1b_encode_chunk ----------------
    # the host copies of the sources live in page-locked memory (av1b_host_alloc), as a capture / decode front
    # end would hand them over; the timed region holds their H2D copies, the kernels, the D2H of the
    # symbol streams, host entropy coding and packet delivery
    pinned = encoder.PinnedFrames(frames, device_id=local_rank)
    pal = list(pinned) + list(pinned)[::-1]
    chunk = [pal[i % len(pal)] for i in range(args.steps * F)]
    enc.encode_chunk(chunk[:2 * F])                      # warm-up
    barrier()
    t0 = time.perf_counter()
    tus = enc.encode_chunk(chunk)
    dte = time.perf_counter() - t0
    st_e = enc.stats()
    dte = sharding.max_over_ranks(dte, dist, "cuda")
    e2e = world * len(chunk) / dte
    d2h_per_step = st_e["d2h_bytes"] // args.steps   # counted by the encoder: token lists + offsets, key-frame levels / block info

    if rank != 0:
        return
    pk, pk_src = peaks()
    n_inter, n_key = max(1, st["inter_launches"]), max(1, st["key_frames"])
    nf = st["frames_done"]
    # per-kernel CUDA-event time per launch (one launch = one frame, except ME = one batch) and
    # algorithmic bytes per launch
    kern = {
        "inter_encode_kernel": (st["inter_ms"] / n_inter, 4 * S),
        "intra_encode_kernel": (st["intra_ms"] / n_key, 3 * S),
        "deblock_kernel": (st["deblock_ms"] / max(1, nf), 2 * S),
        "cdef_kernel": (st["cdef_ms"] / max(1, nf), 3 * S),
        "pyramid+hme (per batch)": (st["me_ms"] / max(1, args.steps), int((1.3125 + 0.625 + 2.0) * (w * h * 2)) * F),
        # tokenizer: reads the block info (20 B per 8x8 unit) twice (count + emit), writes 4 B per token
        "tokenizer (per batch)": (st["tok_ms"] / max(1, args.steps), 2 * 20 * g.w8 * g.h8 * F + 4 * st["tokens"] // max(1, args.steps)),
    }
    share = {"inter_encode_kernel": st["inter_ms"], "intra_encode_kernel": st["intra_ms"], "deblock_kernel": st["deblock_ms"],
             "cdef_kernel": st["cdef_ms"], "pyramid+hme (per batch)": st["me_ms"], "tokenizer (per batch)": st["tok_ms"]}
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
    line = {
        "metric": "AV1 encode fps", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1000 * tmax / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step": F, "crf": args.crf, "preset": args.preset, "base_q_idx": st["base_q_idx"],
                   "keyint": args.keyint, "key_frames": st["key_frames"], "inter_frames": st["inter_launches"],
                   "tiles_key_frames": "%dx%d" % (g.tile_cols, g.tile_rows), "tile_sb_inter": args.tile_sb, "pack_path": args.pack_path, "host_threads": sharding.host_threads_per_rank(world),
                   "l2": "inputs larger than L2 (%.0f MB working set per step)" % (5 * F * frame_bytes / 1e6),
                   "timing": "wall clock between synchronize+barrier pairs (host entropy coding is part of the step); "
                             "kernel times from CUDA events on the encoder stream"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": F * frame_bytes, "d2h_bytes_per_step": d2h_per_step,
                "bitrate_bytes_per_frame": sum(map(len, tus)) / len(tus),
                "host_buffers": "page-locked (av1b_host_alloc), %d of %d frames read in place by the copy engine" % (st_e["staged_direct"], len(chunk)),
                "breakdown_ms_per_step": {k: st_e[k] / args.steps for k in ("h2d_ms", "kernel_ms", "d2h_ms", "pack_ms")}},
        "gpu_launches": st["kernel_launches"],
        "breakdown_ms_per_step": {k: st[k] / args.steps for k in ("kernel_ms", "me_ms", "intra_ms", "inter_ms", "deblock_ms",
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
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        workers = max(1, min(cores, 32))
        chunk_len = 2 if w > 2000 else 3
        fps = oracle_fps(frames, w, h, bd, st["base_q_idx"], workers, chunk_len)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": workers, "kind": "port",
                                "sample": "%d chunks of %d frames (key + inter) of the workload, one chunk per host thread, scalar "
                                          "oracle/av1_oracle.cpp" % (workers, chunk_len)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
