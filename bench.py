#!/usr/bin/env python3
"""bench.py -- 1080p MPEG-1 I-frame encode throughput (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps K --warmup W            our CUDA path
  torchrun ... bench.py --gpus N ...                       one rank per GPU, frame-range sharding
  python bench.py --impl reference ...                     the reference's own CPU functions

Workloads (BASELINE.json `configs`):
  N = 1   configs[1]: 300 synthetic 1920x1080 frames, quality 12, resident in HBM; one step = one pass.
          `other_configs` adds configs[0] (SIF x 30), configs[2] (4K x 300 at quality 5 / 12 / 50), one
          GPU's share of configs[4] (8K x 15) and the content worst cases (noise at quality 50, grey, r == g).
  N > 1   configs[3]: 8000 frames of 1920x1080 split into contiguous frame ranges (rank k encodes
          frame_range(k, N, 8000)); one step = one pass over all 8000 frames INCLUDING the hand-over of every
          rank's compressed segments and per-frame sizes to rank 0 (`scaling: "strong"`).  `weak` repeats the
          round-1 measurement (300 frames per GPU per step), `other_configs` adds configs[4] (8K x 120 split).
Each input (>= 1.8 GB per GPU) is far larger than L2 (126 MB), so no flush is needed between steps.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
After the timed region every rank hashes the bytes it produced and rank 0 hashes what landed in that
rank's region (`gather_verified`), and every rank re-encodes two of its timed frames with the oracle
(`parity`); neither is inside a timed region.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import hashlib
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
SHARDED_TOTAL = 8000                       # configs[3]
METRIC, UNIT = "1080p I-frame encode frames/sec", "frames/s"


def make_config(world: int, frames_local: int) -> dict:
    """The `config` object of the JSON line; both arms (ours and --impl reference) print the same one."""
    if world == 1:
        workload = ("synthetic 1920x1080 RGB, 300 frames, quality 12, 1 B200 (BASELINE configs[1]); one step = one pass "
                    "over the 300 resident frames")
        multi = "single GPU"
    else:
        workload = (f"synthetic 1920x1080 RGB, {SHARDED_TOTAL} frames sharded by contiguous frame range across {world} B200 "
                    f"(BASELINE configs[3]), quality 12; one step = one pass over all {SHARDED_TOTAL} frames")
        multi = ("rank k encodes frame_range(k, N, 8000); inside the timed region every rank pushes its payload bytes into "
                 "its region of rank 0's memory over NVLink (CUDA IPC peer mapping, copy kernel on a side stream), then one "
                 "all_gather of the per-frame sizes/offsets = completion fence; both overlap the next step's encode")
    return {"workload": workload, "width": W, "height": H, "quality": QUALITY,
            "frames_per_gpu_per_step": frames_local,
            "l2": "per-step input (%.2f GB per GPU) exceeds L2; no flush needed" % (frames_local * 3 * W * H / 1e9),
            "timer": "CUDA events on the launching stream, max over ranks", "multi_gpu": multi}


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
    world = int(os.environ.get("WORLD_SIZE", str(max(1, args.gpus))))
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
              f"reference's single-threaded functions (oracle/_ref, gcc -{opt}, -ffp-contract=off) under oracle/ref_driver.c; "
              f"fps = frames / the slowest worker's encode seconds (pool start skew is not charged: this flatters the reference)"
              if kind == "reference" else "oracle port, single thread")
    n_local = FRAMES_PER_STEP if world == 1 else -(-SHARDED_TOTAL // world)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * workers * fpw / fps,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": make_config(world, n_local),
            "megapixels_per_s": fps * W * H / 1e6,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def _port_fps(nframes):
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
class Bench:
    """One process per GPU.  measure() times one workload (device resident, gather inside the timed
    region when world > 1) and verifies it afterwards."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the encode path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL kernels on a high-priority stream: the gather must not queue behind the encode grid
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=self.dev, pg_options=opts)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1) if self.world > 1 else None
        self.sampler = ClockSampler(self.local)
        self.hbm_peak, self.peak_src = measured_peaks()

    def fence(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    # -- one workload ---------------------------------------------------------------------------------
    def measure(self, width, height, total_frames, quality, kind, steps, warmup, sharded, sample_clocks=False,
                check_frames=2):
        """total_frames: the whole job when `sharded` (split by frame_range), else frames per GPU (weak)."""
        torch, dist = self.torch, self.dist
        from ec504_imageencoder_b200 import M1Encoder, MODE_FULL
        from ec504_imageencoder_b200.distributed import PeerGather, frame_range
        world, rank = self.world, self.rank
        if sharded:
            lo, hi = frame_range(rank, world, total_frames)
            counts = [frame_range(k, world, total_frames)[1] - frame_range(k, world, total_frames)[0] for k in range(world)]
        else:
            lo, hi = rank * total_frames, (rank + 1) * total_frames
            counts = [total_frames] * world
        n, nmax = hi - lo, max(counts)
        job_frames = sum(counts)
        enc = M1Encoder(width, height, 3, MODE_FULL, quality, max_frames=nmax, device=self.local)
        stream, comm = self.stream, self.comm
        with torch.cuda.stream(stream):
            rgb = enc.synth_rgb(SEED, lo, n, kind)                      # resident in HBM before timing
            # N > 1 (PeerGather, staged): after a step's encode, a small copy kernel on a high-priority side
            # stream pushes the rank's payload bytes into its region of rank 0's memory (NVLink peer stores),
            # followed by one small all_gather of the frame sizes/offsets = completion fence; both overlap the
            # next step's encode.  Two slots; the timed region ends after the last fence has completed on
            # every rank.
            pg = PeerGather(enc, nmax, slots=2, staged=True) if world > 1 else None
            bufs = [pg.batch(0), pg.batch(1)] if world > 1 else [enc.alloc_outputs(n)]
            done = [torch.cuda.Event() for _ in bufs]
            sent = [torch.cuda.Event() for _ in bufs]
            state = {"i": 0, "pending": None, "last": None, "gathered": None}
            enc.enable_timing(True)

            def launch_gather(j):
                comm.wait_event(done[j])
                with torch.cuda.stream(comm):
                    state["gathered"] = pg.finish(j, counts)
                    sent[j].record(comm)

            def step():
                j = state["i"] % len(bufs)
                state["i"] += 1
                if world > 1:
                    stream.wait_event(sent[j])                       # buffer j's previous gather has been sent
                enc.encode_device(rgb, res=bufs[j], check=False)
                state["last"] = j
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

            for _ in range(max(warmup, 3)):
                step()
            drain()
            enc.check()
            enc.kernel_times()
            self.fence()
            l0 = enc.launches
            if sample_clocks:
                self.sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                step()
            drain()
            e1.record(stream)
            self.fence()
            clocks = self.sampler.stop() if sample_clocks else None
            ms_total = e0.elapsed_time(e1)
            launches = enc.launches - l0
            kms, kn = enc.kernel_times()
            enc.check()
            res = bufs[state["last"]]
            if sample_clocks and clocks["samples"] < 5:
                # the timed region is only a few ms: sample the clocks over ~1 s of the same encode
                scratch = enc.alloc_outputs(n)
                self.sampler.start()
                t_end = time.perf_counter() + 1.0
                while time.perf_counter() < t_end:
                    enc.encode_device(rgb, res=scratch, check=False)
                    torch.cuda.synchronize(self.dev)
                clocks = self.sampler.stop()
                clocks["note"] = "timed region shorter than the NVML sampling period; sampled over 1 s of identical steps right after it"
                del scratch
            payload_local = int(res.frame_bytes[:n].to(torch.int64).sum().item())

            # ---- after the timed region: what did rank 0 receive, and is it what the reference computes?
            verify = self._verify_gather(res, n, counts, state["gathered"]) if world > 1 else None
            parity = self._parity(enc, rgb, res, n, lo, width, height, quality, kind, check_frames)

        (ms_total,) = self.max_over_ranks(ms_total)
        tot = torch.tensor([payload_local, parity["frames_checked"], parity["identical"]], dtype=torch.int64, device=self.dev)
        if world > 1:
            dist.all_reduce(tot)
        payload_job, checked, identical = tot.tolist()
        ms_step = ms_total / steps
        fps = job_frames / (ms_step * 1e-3)
        alg_frame = 3 * width * height + payload_job / job_frames + 4
        launch_ms = kms[0] / max(1, kn[0])
        frames_per_launch = n * steps / max(1, kn[0])
        achieved = alg_frame * frames_per_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
        out = {
            "fps": fps, "ms_per_step": ms_step, "frames_per_step": job_frames, "frames_local": n, "launches": launches,
            "payload_bytes_per_frame": payload_job / job_frames, "alg_bytes_per_frame": alg_frame,
            "launch_ms": launch_ms, "frames_per_launch": frames_per_launch, "achieved_gbs": achieved,
            "frac": achieved / self.hbm_peak, "whole_step_frac": alg_frame * fps / 1e9 / self.hbm_peak / world,
            "kernel_ms_per_step": {"k_encode_chunks": kms[0] / steps, "k_layout": kms[1] / steps, "k_stitch": kms[2] / steps},
            "clocks": clocks, "gather": verify,
            "parity": {"frames_checked": checked, "identical": identical, "ties": 0,
                       "pct_identical": 100.0 * identical / max(1, checked), "checker": parity["checker"],
                       "what": "quantised zigzag levels and payload bytes of the first and last frame each rank timed"},
            "flat_blocks": parity.get("flat"),
            "enc": enc, "rgb": rgb,
        }
        if pg is not None:
            pg.close()
        return out

    def _verify_gather(self, res, n, counts, gathered):
        """Every rank hashes the payload bytes [0, end) it produced for the last timed step; rank 0 hashes
        what arrived in that rank's region and compares sizes and offsets too."""
        torch, dist = self.torch, self.dist
        end = int(res.frame_offsets[n].item()) if n else 0
        mine = hashlib.sha256(res.out[:end].cpu().numpy().tobytes()).hexdigest() if res.out_ptr is None else None
        sizes = res.frame_bytes[:n].cpu().tolist()
        offs = res.frame_offsets[:n + 1].cpu().tolist()
        local = {"end": end, "sha": mine, "sizes_sha": hashlib.sha256(repr((sizes, offs)).encode()).hexdigest()}
        allv = [None] * self.world
        dist.all_gather_object(allv, local)
        ok = None
        if self.rank == 0:
            ok = True
            for k in range(self.world):
                if counts[k] == 0:
                    continue
                s = gathered.sizes[k].cpu().tolist()
                o = gathered.offsets[k].cpu().tolist()
                got_meta = hashlib.sha256(repr((s, o)).encode()).hexdigest()
                got = hashlib.sha256(gathered.segments[k][:allv[k]["end"]].cpu().numpy().tobytes()).hexdigest()
                ok = ok and got_meta == allv[k]["sizes_sha"] and got == allv[k]["sha"]
        flag = [ok]
        dist.broadcast_object_list(flag, src=0)
        return {"gather_verified": bool(flag[0]), "ranks": self.world,
                "bytes_into_rank0_per_step": sum(v["end"] for v in allv[1:]),
                "what": "sha256 of every rank's payload segment and of its sizes/offsets, sender vs rank 0's region, last timed step"}

    def _parity(self, enc, rgb, res, n, lo, width, height, quality, kind, check_frames):
        """Re-encodes up to `check_frames` of the frames this rank just timed with the oracle (the unmodified
        reference functions when oracle/_ref travelled, else the C port) and compares levels and bytes."""
        import numpy as np
        import oracle
        torch = self.torch
        use_ref = oracle.Ref.available("O2")
        chk = oracle.Ref("O2") if use_ref else oracle.Port()
        idx = sorted({0, n - 1})[:check_frames] if n else []
        identical = 0
        if idx:
            sub = rgb[idx].contiguous()
            small = enc.encode_device(sub, want_levels=True)
            pays = small.payloads()
            sizes = res.frame_bytes.cpu().tolist()
            offs = res.frame_offsets.cpu().tolist()
            for j, f in enumerate(idx):
                host = sub[j].cpu().numpy()
                want_pay, want_lev = chk.encode_picture(host, quality, oracle.MODE_FULL, want_levels=True)
                timed = (res.out[offs[f]:offs[f] + sizes[f]].cpu().numpy().tobytes() if res.out_ptr is None else pays[j])
                same = np.array_equal(small.levels[j].cpu().numpy(), want_lev) and pays[j] == want_pay and timed == want_pay
                identical += bool(same)
        return {"frames_checked": len(idx), "identical": identical,
                "checker": "unmodified reference functions (oracle/_ref, -O2)" if use_ref else "oracle C port",
                "flat": self._flat_share(enc, rgb[idx[0]].cpu().numpy() if idx else None)}

    @staticmethod
    def _flat_share(enc, host):
        """Share of the first timed frame's 8x8 blocks whose samples span at most m1cu_flat_range (those skip the DCT in
        k_encode_chunks): computed on the host from the oracle's planes, outside every timed region."""
        import numpy as np
        import oracle
        R = int(getattr(enc, "flat_range", -1))
        if host is None or R < 0:
            return {"range": R, "blocks_skipping_dct_pct": 0.0}
        port = oracle.Port()
        H, W = host.shape[:2]
        planes = [np.asarray(p).reshape(H, W).astype(np.int32) for p in port.rgb_to_ycbcr(host)]
        pad = lambda p: np.pad(p, ((0, (-p.shape[0]) % 16), (0, (-p.shape[1]) % 16)), mode="edge")
        y, cb, cr = (pad(p) for p in planes)
        sub = lambda p: (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]) // 4

        def spans(p):
            b = p.reshape(p.shape[0] // 8, 8, p.shape[1] // 8, 8)
            return (b.max(axis=(1, 3)) - b.min(axis=(1, 3))).ravel()
        sp = np.concatenate([spans(y), spans(sub(cb)), spans(sub(cr))])
        return {"range": R, "blocks_skipping_dct_pct": round(100.0 * float((sp <= R).mean()), 2)}

    # -- e2e through the host-buffer C-ABI call ------------------------------------------------------
    def e2e(self, enc, rgb, n):
        import numpy as np
        torch = self.torch
        host_rgb = torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=True)
        host_rgb.copy_(rgb[:n])
        torch.cuda.synchronize(self.dev)
        out_np = np.empty(enc.typical_out_bytes(n), np.uint8)
        e2e_steps = max(1, min(self.args.steps, 3))
        enc.encode_host(host_rgb, out=out_np)                      # warm-up (allocations)
        self.fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            payloads, _ = enc.encode_host(host_rgb, out=out_np, copy=False)
        self.fence()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        # context for the e2e number: what a plain pinned host->device copy of the same input achieves
        # when all ranks copy at the same time (the host memory system / PCIe roots are shared)
        dcopy = torch.empty_like(rgb[:n])
        dcopy.copy_(host_rgb, non_blocking=True)
        self.fence()
        t1 = time.perf_counter()
        dcopy.copy_(host_rgb, non_blocking=True)
        torch.cuda.synchronize(self.dev)
        h2d_s = time.perf_counter() - t1
        del dcopy
        d2h = sum(len(p) for p in payloads) + 4 * n + 8 * (n + 1)
        nbytes = host_rgb.numel()
        # the same two measurements with the input in WRITE-COMBINED pinned memory (m1cu_pinned_alloc_wc)
        wc_e2e_s = wc_h2d_s = float("nan")
        import ctypes as C
        lib = enc.lib
        wc = lib.m1cu_pinned_alloc_wc(nbytes)
        if wc:
            try:
                lib.m1cu_memcpy_d2h(wc, rgb.data_ptr(), nbytes)                      # filled by DMA writes only
                view = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(wc)).reshape(n, H, W, 3)
                dst = torch.empty_like(rgb[:n])
                lib.m1cu_memcpy_h2d(dst.data_ptr(), wc, nbytes)
                self.fence()
                t2 = time.perf_counter()
                lib.m1cu_memcpy_h2d(dst.data_ptr(), wc, nbytes)
                wc_h2d_s = time.perf_counter() - t2
                del dst
                enc.encode_host(view, out=out_np, copy=False)
                self.fence()
                t3 = time.perf_counter()
                for _ in range(e2e_steps):
                    enc.encode_host(view, out=out_np, copy=False)
                self.fence()
                wc_e2e_s = (time.perf_counter() - t3) / e2e_steps
                del view
            finally:
                lib.m1cu_pinned_free(wc)
        e2e_s, h2d_s, wc_e2e_s, wc_h2d_s = self.max_over_ranks(e2e_s, h2d_s, wc_e2e_s, wc_h2d_s)
        return e2e_s, h2d_s, d2h, e2e_steps, nbytes, wc_e2e_s, wc_h2d_s


def run_ours(args):
    from ec504_imageencoder_b200 import SYNTH_NATURAL, SYNTH_NOISE, SYNTH_GREY, SYNTH_RG_EQUAL, SYNTH_SCATTERED
    b = Bench(args)
    world, rank = b.world, b.rank
    steps, warmup = args.steps, max(args.warmup, 3)

    if world == 1:
        head = b.measure(W, H, args.frames, QUALITY, SYNTH_NATURAL, steps, warmup, sharded=False, sample_clocks=True)
        n_local = args.frames
    else:
        head = b.measure(W, H, args.total_frames, QUALITY, SYNTH_NATURAL, steps, warmup, sharded=True, sample_clocks=True)
        n_local = head["frames_local"]

    # e2e on a bounded sample of the same frames (300 per rank: 1.87 GB of pinned host memory per rank)
    n_e2e = min(n_local, FRAMES_PER_STEP)
    e2e_s, h2d_s, d2h, e2e_steps, h2d_bytes, wc_e2e_s, wc_h2d_s = b.e2e(head["enc"], head["rgb"], n_e2e)
    head.pop("enc").close()
    head.pop("rgb")
    b.torch.cuda.empty_cache()

    others, weak = [], None
    osteps = max(3, min(steps, 10))
    if not args.no_other_configs:
        if world == 1:
            cases = [("configs[0] SIF x 30", 352, 240, 30, 12, SYNTH_NATURAL),
                     ("configs[2] 4K x 300, quality 5", 3840, 2160, 300, 5, SYNTH_NATURAL),
                     ("configs[2] 4K x 300, quality 12", 3840, 2160, 300, 12, SYNTH_NATURAL),
                     ("configs[2] 4K x 300, quality 50", 3840, 2160, 300, 50, SYNTH_NATURAL),
                     ("configs[4] 8K, one GPU's share of 120 frames over 8 (15 frames)", 7680, 4320, 15, 12, SYNTH_NATURAL),
                     ("1080p x 300 noise, quality 12", W, H, 300, 12, SYNTH_NOISE),
                     ("1080p x 300 noise, quality 50 (VLC stress)", W, H, 300, 50, SYNTH_NOISE),
                     ("1080p x 300 scattered (a quarter of the 8x8 pixel tiles are noise: busy and flat blocks in every warp), quality 12",
                      W, H, 300, 12, SYNTH_SCATTERED),
                     ("1080p x 300 grey (every pixel an exact-quotient case of the colour arithmetic), quality 12", W, H, 300, 12, SYNTH_GREY),
                     ("1080p x 300 r == g (every pixel an exact-quotient case of Cb), quality 12", W, H, 300, 12, SYNTH_RG_EQUAL)]
            for name, w, h, nfr, q, kind in cases:
                r = b.measure(w, h, nfr, q, kind, osteps, 3, sharded=False, check_frames=1 if w * h > 3000 * 2000 else 2)
                r.pop("enc").close(); r.pop("rgb")
                b.torch.cuda.empty_cache()
                others.append(_other(name, w, h, q, r))
        else:
            r = b.measure(7680, 4320, 120, 12, SYNTH_NATURAL, osteps, 3, sharded=True, check_frames=1)
            r.pop("enc").close(); r.pop("rgb")
            b.torch.cuda.empty_cache()
            others.append(_other(f"configs[4] 8K x 120 across {world} B200 ({r['frames_local']} per GPU), gather to rank 0 inside the step",
                                 7680, 4320, 12, r))
            wk = b.measure(W, H, FRAMES_PER_STEP, QUALITY, SYNTH_NATURAL, osteps, 3, sharded=False)
            wk.pop("enc").close(); wk.pop("rgb")
            b.torch.cuda.empty_cache()
            weak = {"value": wk["fps"], "unit": UNIT, "frames_per_gpu_per_step": FRAMES_PER_STEP, "ms_per_step": wk["ms_per_step"],
                    "gather_verified": wk["gather"]["gather_verified"], "parity": wk["parity"],
                    "note": "round-1 workload: every GPU encodes its own 300 frames per step (weak scaling), same hand-over"}

    line = None
    if rank == 0:
        traffic, traffic_src = None, None
        for name in ("r2_traffic.json", "r1_traffic.json"):                 # DRAM bytes from the committed ncu capture
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic = tj["dram_bytes_per_frame"] * head["frames_per_launch"]
                traffic_src = f"profiles/{name} (ncu dram__bytes_read.sum + dram__bytes_write.sum per frame x frames per launch)"
                break
            except Exception:
                pass
        e2e_fps = world * n_e2e / e2e_s
        line = {
            "metric": METRIC, "value": head["fps"], "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": make_config(world, n_local),
            "megapixels_per_s": head["fps"] * W * H / 1e6,
            "payload_bytes_per_frame": head["payload_bytes_per_frame"],
            "clocks": head["clocks"],
            "parity": head["parity"],
            "flat_blocks": dict(head.get("flat_blocks") or {}, note=(
                "k_encode_chunks computes only the DC coefficient of a block whose samples span at most `range` grey levels: a "
                "rigorous bound (csrc/m1cu_quant.h, tests/test_block_host.py::test_flat_block_bound) proves every AC level of such "
                "a block zero at this quality; throughput therefore depends on content -- see the noise rows of other_configs, "
                "where no block qualifies; rank 0's first timed frame")),
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                    "frames_per_gpu_per_step": n_e2e,
                    "timer": "host wall clock around m1cu_encode_host (pinned input, pageable output), max over ranks",
                    "steps": e2e_steps,
                    "plain_h2d_copy_gbs": h2d_bytes / h2d_s / 1e9,
                    "plain_h2d_aggregate_gbs": world * h2d_bytes / h2d_s / 1e9,
                    "e2e_input_gbs": world * h2d_bytes / e2e_s / 1e9,
                    "frac_of_plain_h2d": h2d_s / e2e_s,
                    "write_combined_input": {"value": world * n_e2e / wc_e2e_s, "unit": UNIT,
                                             "plain_h2d_aggregate_gbs": world * h2d_bytes / wc_h2d_s / 1e9,
                                             "note": "same call with the input in cudaHostAllocWriteCombined memory (m1cu_pinned_alloc_wc)"},
                    "note": ("bound by the host->device copy of the RGB input: plain_h2d_* is a bare pinned copy of the same bytes "
                             "issued by all ranks at the same time (per rank / summed), e2e_input_gbs the RGB bytes per second the "
                             "whole call sustains, frac_of_plain_h2d their ratio; "
                             + ("a bounded sample of each rank's range (300 frames)" if world > 1 else "the whole 300-frame step"))},
            "gpu_launches": head["launches"],
            "roofline": {"bound": "hbm", "kernel": "k_encode_chunks", "achieved": head["achieved_gbs"], "peak": b.hbm_peak,
                         "unit": "GB/s", "frac": head["frac"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": b.peak_src,
                         "algorithmic_bytes_per_frame": head["alg_bytes_per_frame"], "frames_per_launch": head["frames_per_launch"],
                         "launch_ms": head["launch_ms"], "whole_step_frac": head["whole_step_frac"],
                         "note": ("the dominant kernel is instruction-issue bound, not HBM bound: the reference's exact double-precision "
                                  "colour chain + int32 DCT + VLC cost 59 issue slots per pixel where every block needs its DCT and about 48 "
                                  "on the default content (flat blocks get their DC only, see flat_blocks), issued at 0.71 - 0.77 per cycle and "
                                  "sub-partition (two-cycle FP64 / ALU / IMAD instructions), DRAM at 18 % of peak; an integer colour path, "
                                  "a mixed one, a barrier-free warp-per-chunk kernel and a queue-fed persistent form were built, are "
                                  "bit-exact and measured slower (DESIGN.md section 7, profiles/r2_*)"),
                         "kernel_ms_per_step": head["kernel_ms_per_step"]},
        }
        line["memcheck"] = ("compute-sanitizer is closed on this pool: no memcheck / racecheck run of the kernels exists; evidence is the "
                            "bit-exact parity suite (every geometry, both modes, odd sizes, multi-window chunks) and, for the host C, "
                            "ASan + UBSan (profiles/r2_host_sanitizers.txt: clean)")
        if head["gather"] is not None:
            line["gather_verified"] = head["gather"]["gather_verified"]
            line["ranks"] = head["gather"]["ranks"]
            line["gather"] = head["gather"]
        if others:
            line["other_configs"] = others
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
    if world > 1:
        b.dist.barrier()
        b.dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def _other(name, w, h, q, r):
    o = {"workload": name, "width": w, "height": h, "quality": q, "frames_per_step": r["frames_per_step"],
         "value": r["fps"], "unit": UNIT, "megapixels_per_s": r["fps"] * w * h / 1e6, "ms_per_step": r["ms_per_step"],
         "payload_bytes_per_frame": r["payload_bytes_per_frame"],
         "roofline": {"kernel": "k_encode_chunks", "frac": r["frac"], "achieved": r["achieved_gbs"], "launch_ms": r["launch_ms"],
                      "frames_per_launch": r["frames_per_launch"], "whole_step_frac": r["whole_step_frac"]},
         "parity": r["parity"], "flat_blocks": r.get("flat_blocks")}
    if r["gather"] is not None:
        o["gather_verified"] = r["gather"]["gather_verified"]
        o["ranks"] = r["gather"]["ranks"]
    return o


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
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="N=1: frames per step (configs[1]: 300)")
    ap.add_argument("--total-frames", type=int, default=SHARDED_TOTAL, help="N>1: frames of the whole job (configs[3]: 8000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
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
