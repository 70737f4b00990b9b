#!/usr/bin/env python3
"""bench.py -- 1080p MPEG-1 I-frame encode throughput (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps K --warmup W            our CUDA path
  torchrun ... bench.py --gpus N ...                       one rank per GPU, frame-range sharding
  python bench.py --impl reference ...                     the reference's own CPU functions

A "step" is one pass of the hot path over one batch of synthetic pictures (configs[1]:
300 frames of 1920x1080 RGB, quality 12) that are already resident in HBM.  The batch (1.87 GB)
is far larger than L2 (126 MB), so no flush is needed between steps.  Timing: CUDA events on the
launching stream, barrier + synchronize on both sides, max over ranks.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, QUALITY, FRAMES_PER_STEP, SEED = 1920, 1080, 12, 300, 12345
METRIC, UNIT = "1080p I-frame encode frames/sec", "frames/s"
WORKLOAD = "synthetic 1920x1080 RGB, 300 frames per GPU per step, quality 12 (BASELINE configs[1]; N>1 = configs[3] frame-range sharding)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# clocks: sampled with NVML from a thread while the timed region runs
# --------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), threading.Event(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self.thread = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.samples.clear(); self.reasons.clear(); self.stop_flag.clear()
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread:
            self.stop_flag.set(); self.thread.join(); self.thread = None
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own functions (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------------
def _ref_worker(args):
    first, count, opt = args
    import numpy as np
    import oracle
    ref, port = oracle.Ref(opt), oracle.Port()
    frames = np.stack([port.synth_rgb(SEED, first + i, W, H, oracle.SYNTH_NATURAL) for i in range(count)])
    secs, nbytes = ref.time_pictures(frames, QUALITY)
    return secs, nbytes


def cpu_reference_fps(frames_per_worker: int, workers: int, opt: str, first: int = 0):
    """Whole-sample fps with `workers` processes, each running the single-threaded reference
    functions on its own frames (the reference leaks ~31 MB per 1080p frame, so every task runs in
    a fresh process).  Wall clock includes only the encode calls' span (inputs generated first)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    tasks = [(first + w * frames_per_worker, frames_per_worker, opt) for w in range(workers)]
    if workers == 1:
        with ctx.Pool(1, maxtasksperchild=1) as pool:
            secs, _ = pool.map(_ref_worker, tasks)[0]
        return frames_per_worker / secs, secs
    t0 = time.perf_counter()
    with ctx.Pool(workers, maxtasksperchild=1) as pool:
        res = pool.map(_ref_worker, tasks, chunksize=1)
    wall = time.perf_counter() - t0
    busy = max(s for s, _ in res)               # workers run concurrently: the slowest one bounds the sample
    return workers * frames_per_worker / busy, wall


def run_reference(args):
    import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not oracle.Ref.available("O2"):
        # the oracle always exists: fall back to our C port of the same algorithm
        kind, opt = "port", None
    else:
        kind, opt = "reference", "O2"
    cores = len(os.sched_getaffinity(0))
    workers = max(1, cores)
    fpw = 2
    vals = []
    for i in range(args.warmup + args.steps):
        if kind == "reference":
            fps, _ = cpu_reference_fps(fpw, workers, opt, first=i * workers * fpw)
        else:
            fps = _port_fps(workers * fpw)
        if i >= args.warmup:
            vals.append(fps)
    fps = statistics.mean(vals)
    sample = (f"{workers} processes x {fpw} frames of the 1920x1080 workload per step, each process running the "
              f"reference's single-threaded functions (oracle/_ref, gcc -{opt}, -ffp-contract=off) under oracle/ref_driver.c"
              if kind == "reference" else "oracle port, single thread")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * workers * fpw / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "quality": QUALITY, "width": W, "height": H},
            "megapixels_per_s": fps * W * H / 1e6,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def _port_fps(nframes):
    import numpy as np
    import oracle
    port = oracle.Port()
    frames = [port.synth_rgb(SEED, i, W, H, oracle.SYNTH_NATURAL) for i in range(nframes)]
    t0 = time.perf_counter()
    for f in frames:
        port.encode_picture(f, QUALITY, oracle.MODE_FULL)
    return nframes / (time.perf_counter() - t0)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from ec504_imageencoder_b200 import M1Encoder, MODE_FULL, SYNTH_NATURAL
    from ec504_imageencoder_b200.distributed import PeerGather

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the encode path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL kernels on a high-priority stream: the gather must not queue behind the encode grid
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    n = args.frames
    enc = M1Encoder(W, H, 3, MODE_FULL, QUALITY, max_frames=n, device=local)
    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local)
    hbm_peak, peak_src = measured_peaks()

    with torch.cuda.stream(stream):
        first = rank * n                                            # contiguous frame range per rank
        rgb = enc.synth_rgb(SEED, first, n, SYNTH_NATURAL)          # resident in HBM before timing
        # N > 1 (PeerGather, staged): after a step's encode, a small copy kernel on a high-priority side
        # stream pushes the rank's payload bytes into its region of rank 0's memory (NVLink peer stores),
        # followed by one small all_gather of the frame sizes/offsets = completion fence; both overlap the
        # next step's encode.  Two slots; the timed region ends after the last fence has completed on
        # every rank.  (M1_PEER_DIRECT=1: k_stitch writes remotely instead, no push, not overlappable.)
        pg = PeerGather(enc, n, slots=2, staged=os.environ.get("M1_PEER_DIRECT", "0") != "1") if world > 1 else None
        bufs = [pg.batch(0), pg.batch(1)] if world > 1 else [enc.alloc_outputs(n)]
        res = bufs[0]
        comm = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
        done = [torch.cuda.Event() for _ in bufs]
        sent = [torch.cuda.Event() for _ in bufs]
        state = {"i": 0, "pending": None}
        enc.enable_timing(True)

        def launch_gather(j):
            comm.wait_event(done[j])
            with torch.cuda.stream(comm):
                pg.finish(j)
                sent[j].record(comm)

        def step():
            j = state["i"] % len(bufs)
            state["i"] += 1
            if world > 1:
                stream.wait_event(sent[j])                       # buffer j's previous gather has been sent
            enc.encode_device(rgb, res=bufs[j], check=False)
            if world > 1:
                done[j].record(stream)
                if state["pending"] is not None:
                    launch_gather(state["pending"])              # overlaps the encode just launched
                state["pending"] = j

        def drain():
            if world > 1:
                if state["pending"] is not None:
                    launch_gather(state["pending"])
                    state["pending"] = None
                stream.wait_stream(comm)

        def fence():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        for _ in range(max(args.warmup, 3)):
            step()
        drain()
        enc.check()
        enc.kernel_times()
        fence()
        l0 = enc.launches
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        drain()
        e1.record(stream)
        fence()
        clocks = sampler.stop()
        ms_total = e0.elapsed_time(e1)
        launches = enc.launches - l0
        kms, kn = enc.kernel_times()
        enc.check()
        if clocks["samples"] < 5:
            # the timed region is only a few ms: sample the clocks over ~1 s of the same steps
            sampler.start()
            t_end = time.perf_counter() + 1.0
            while time.perf_counter() < t_end:
                enc.encode_device(rgb, res=res, check=False)
                torch.cuda.synchronize(dev)
            clocks = sampler.stop()
            clocks["note"] = "timed region shorter than the NVML sampling period; sampled over 1 s of identical steps right after it"

        payload_bytes = int(res.frame_bytes.to(torch.int64).sum().item())

        # ---- e2e: the same metric through the host-buffer C-ABI call (pinned host memory in,
        # host payload out), H2D and D2H inside the timed region
        host_rgb = torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=True)
        host_rgb.copy_(rgb)
        torch.cuda.synchronize(dev)
        out_np = np.empty(enc.typical_out_bytes(n), np.uint8)
        e2e_steps = max(1, min(args.steps, 3))
        enc.encode_host(host_rgb, out=out_np)                      # warm-up (allocations)
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            payloads, _ = enc.encode_host(host_rgb, out=out_np, copy=False)
        fence()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        # context for the e2e number: what a plain pinned host->device copy of the same input achieves
        dcopy = torch.empty_like(rgb)
        dcopy.copy_(host_rgb, non_blocking=True)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        dcopy.copy_(host_rgb, non_blocking=True)
        torch.cuda.synchronize(dev)
        h2d_gbs = host_rgb.numel() / (time.perf_counter() - t1) / 1e9
        del dcopy
        d2h = sum(len(p) for p in payloads) + 4 * n + 8 * (n + 1)

    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = t.tolist()
    ms_step = ms_total / args.steps
    fps = world * n / (ms_step * 1e-3)
    e2e_fps = world * n / e2e_s

    line = None
    if rank == 0:
        alg_bytes_frame = 3 * W * H + payload_bytes / n + 4
        enc_launch_ms = kms[0] / max(1, kn[0])
        frames_per_launch = n * args.steps / max(1, kn[0])
        achieved = alg_bytes_frame * frames_per_launch / (enc_launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:                                                        # DRAM bytes from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            traffic = tj["dram_bytes_per_frame"] * frames_per_launch
            traffic_src = "profiles/r1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per frame x frames per launch)"
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "quality": QUALITY, "frames_per_gpu_per_step": n,
                       "l2": "per-step input (%.2f GB per GPU) exceeds L2; no flush needed" % (n * 3 * W * H / 1e9),
                       "timer": "CUDA events on the launching stream, max over ranks",
                       "multi_gpu": ("contiguous frame ranges per rank; inside the timed region every rank pushes its payload bytes into "
                                     "its region of rank 0's memory over NVLink (CUDA IPC peer mapping, copy kernel on a side stream), then one "
                                     "all_gather of the per-frame sizes/offsets = completion fence; both overlap the next step's encode")
                       if world > 1 else "single GPU"},
            "megapixels_per_s": fps * W * H / 1e6,
            "payload_bytes_per_frame": payload_bytes / n,
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": n * 3 * W * H, "d2h_bytes_per_step": d2h,
                    "timer": "host wall clock around m1cu_encode_host (pinned input, pageable output), max over ranks",
                    "steps": e2e_steps, "plain_h2d_copy_gbs": h2d_gbs,
                    "note": "bound by the host->device copy of the RGB input (PCIe); plain_h2d_copy_gbs is a bare pinned copy of the same bytes"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_encode_chunks", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": alg_bytes_frame, "frames_per_launch": frames_per_launch,
                         "launch_ms": enc_launch_ms,
                         "note": ("the dominant kernel is not HBM-bound: it sits at the issue ceiling of its mix of two-cycle "
                                  "FP64 / ALU / IMAD instructions (DESIGN.md section 7, profiles/r1s2_ubench_issue_mix.txt)"),
                         "kernel_ms_per_step": {"k_encode_chunks": kms[0] / args.steps, "k_layout": kms[1] / args.steps,
                                                "k_stitch": kms[2] / args.steps}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
    if world > 1:
        pg.close()
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def cpu_baseline():
    """The reference's own single-threaded functions (as shipped: gcc -O0) on a bounded sample of
    the same workload, in a fresh process (the reference leaks ~31 MB per 1080p frame)."""
    import oracle
    cores = len(os.sched_getaffinity(0))
    if oracle.Ref.available("O0"):
        nfr = 24
        fps, _ = cpu_reference_fps(nfr, 1, "O0")
        out = {"value": fps, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"{nfr} frames of the same 1920x1080 q=12 synthetic workload, reference functions compiled as shipped "
                         f"(gcc -g -O0) under oracle/ref_driver.c, 1 thread of {cores} host cores"}
        if oracle.Ref.available("O2"):
            fps2, _ = cpu_reference_fps(nfr, 1, "O2")
            out["value_O2"] = fps2
        return out
    nfr = 24
    return {"value": _port_fps(nfr), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{nfr} frames, oracle/m1_oracle.c (gcc -O2), 1 thread of {cores} host cores"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
