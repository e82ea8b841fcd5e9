#!/usr/bin/env python3
"""bench.py -- AV1 encode fps of the B200 backend (BASELINE.json metric), one rank per GPU.

A "step" is one pass of the hot path over one batch of F synthetic frames (one chunk stream per GPU,
SURVEY.md 8e: chunks share no state, so ranks never communicate on the data path; weak scaling).
  value : frames/s with the source frames already resident in HBM (av1b_encode_resident): device
          kernels + symbol download + host entropy coding, pipelined; whole job over all ranks.
  e2e   : frames/s through av1b_encode_chunk with HOST buffers (pageable -> pinned staging -> H2D ->
          kernels -> D2H -> entropy coding -> packets), the call a reference-side binding makes.
  roofline : the dominant kernel (intra_encode_kernel), algorithmic bytes / CUDA-event duration
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
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(w, h, bd, n, hdr):
    from av1_base_b200 import synth
    # two scenes so that the batch is not one static picture; deterministic
    return synth.synth_clip(w, h, bd, n, seed=4, scene_len=max(1, n // 2), hdr=hdr)


def oracle_fps(frames, w, h, bd, qidx, threads, blk_log2=4):
    """CPU port of the same path (oracle/av1_oracle.cpp), one frame per thread (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as O
    g = O.geom(w, h, 0, 0)
    pm = O.partition_fixed(g, blk_log2)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda fr: O.encode_intra_frame(g, fr, bd, qidx, pm), frames))
    return len(frames) / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """CPU arm: rank 0 only."""
    if rank != 0:
        return
    w, h, bd, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    from av1_base_b200 import abi  # noqa: F401  (tables only; no device use)
    qidx = 120   # CRF 30 -> quantizer_to_qindex[30]
    per_step = max(1, min(cores, 16))
    frames = make_frames(w, h, bd, min(per_step, 4), args.workload == "4k10")
    frames = [frames[i % len(frames)] for i in range(per_step)]
    for _ in range(args.warmup):
        oracle_fps(frames[:max(1, per_step // 2)], w, h, bd, qidx, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_fps(frames, w, h, bd, qidx, cores)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": "AV1 encode fps", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step": per_step, "crf": 30, "all_intra": True},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d frames of the workload, one frame per host thread, oracle/av1_oracle.cpp "
                                   "(av1an+SVT-AV1 itself cannot run in this image: BASELINE.md section 2)" % (args.steps, per_step)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "svt_av1_fps": None,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k10", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=8)
    ap.add_argument("--crf", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    frames = make_frames(w, h, bd, 2 * F, hdr)          # two device slots of distinct frames (seeded per rank below)
    if rank:
        frames = frames[rank % len(frames):] + frames[:rank % len(frames)]
    enc = encoder.Encoder(w, h, bd, crf=args.crf, device_id=local_rank, hdr=hdr, frames_in_flight=F)
    g = enc.geom
    frame_bytes = sum(g.stride[p] * (h if p == 0 else h // 2) * 2 for p in range(3))
    alg_bytes_per_frame = 3 * int(1.5 * w * h * 2)     # src read + recon write + int16 levels write

    # ---------------- value: inputs resident in HBM ----------------
    enc.stage_frames(0, frames[:F])
    enc.stage_frames(1, frames[F:2 * F])
    enc.encode_resident(max(args.warmup, 3))
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    enc.encode_resident(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = enc.stats()
    barrier()
    clocks = sampler.stop()
    tmax = dt
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tmax = float(t.item())
    value = world * args.steps * F / tmax

    # ---------------- e2e: host buffers through av1b_encode_chunk ----------------
    chunk = [frames[i % len(frames)] for i in range(args.steps * F)]
    enc.encode_chunk(chunk[:2 * F])                      # warm-up
    barrier()
    t0 = time.perf_counter()
    tus = enc.encode_chunk(chunk)
    dte = time.perf_counter() - t0
    st_e = enc.stats()
    if dist is not None:
        t = torch.tensor([dte], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dte = float(t.item())
    e2e = world * len(chunk) / dte
    d2h_per_step = F * (frame_bytes + g.w8 * g.h8 * 16)

    if rank != 0:
        return
    pk, pk_src = peaks()
    intra_ms = st["intra_ms"] / max(1, st["intra_launches"])
    achieved = (alg_bytes_per_frame * F) / (intra_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "intra_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    line = {
        "metric": "AV1 encode fps", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1000 * tmax / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step": F, "crf": args.crf, "base_q_idx": st["base_q_idx"],
                   "all_intra": True, "tiles": "%dx%d" % (g.tile_cols, g.tile_rows),
                   "l2": "inputs larger than L2 (%.0f MB per step)" % (3 * F * frame_bytes / 1e6),
                   "timing": "wall clock between synchronize+barrier pairs (host entropy coding is part of the step); "
                             "kernel times from CUDA events on the encoder stream"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": F * frame_bytes, "d2h_bytes_per_step": d2h_per_step,
                "bitrate_bytes_per_frame": sum(map(len, tus)) / len(tus),
                "breakdown_ms_per_step": {k: st_e[k] / args.steps for k in ("h2d_ms", "kernel_ms", "d2h_ms", "pack_ms")}},
        "gpu_launches": st["kernel_launches"],
        "breakdown_ms_per_step": {k: st[k] / args.steps for k in ("kernel_ms", "d2h_ms", "pack_ms")},
        "roofline": {"kernel": "intra_encode_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_src,
                     "ms_per_launch": intra_ms, "algorithmic_bytes_per_launch": alg_bytes_per_frame * F},
        "clocks": clocks,
        "svt_av1_fps": None,
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        nsample = max(1, min(cores, F))
        fps = oracle_fps(frames[:nsample], w, h, bd, st["base_q_idx"], cores)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d frames of the workload, one per host thread (oracle/av1_oracle.cpp)" % nsample}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
