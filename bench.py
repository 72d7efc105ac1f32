#!/usr/bin/env python
"""bench.py — Mpix/s warped on B200 for BASELINE.json's configs, with roofline + CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU loop (oracle port), rank 0 only
    python bench.py --workload piecewise4 ...                # one secondary workload alone (profiling)

Headline at every N: config 2 of BASELINE.json — projective 4-point warp of 1920x1080 RGBA8 frames into their
1728x1080 output window, the reference's per-frame protocol (test/benchmark.js:96-113): every frame solves its own
transform from its own destiny points on the device, then runs the pixel loop.  One step = one batch of FRAMES
independent frames (distinct source + distinct output buffer per frame: a ring far larger than the 126 MB L2, so every
step streams from / to HBM).  Frames are independent, so N GPUs each warp their own batch (weak scaling, no data-path
collective).  The same line carries a `secondary` list: BASELINE configs 3, 4 and 5 and the forward loops, each with its
own parity / checksum gate (config 5 broadcasts its shared source image over NCCL when N > 1).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_PIXEL = 8  # 4 B RGBA8 read + 4 B RGBA8 write per output pixel (SURVEY §8d / north_star)


def measured_traffic(kernel: str, frames_per_launch: int):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), scaled to this
    run's frames per launch (the capture's bytes are proportional to the frames it covered), with where it came from."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return int(t["dram_bytes_per_launch"] * frames_per_launch / t["frames_per_launch"]), \
            {"file": "profiles/" + t["capture"], "commit": t["commit"], "frames_per_launch_captured": t["frames_per_launch"]}
    except Exception:
        return None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample_once(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                     "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample_once()
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        self.sample_once()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# the headline workload: config 2 with per-frame destiny points

def headline_points(wl, n_frames: int, first: int = 0):
    """Destiny points of frames first .. first+n-1.  Frame 0 is exactly BASELINE config 2; the others move the two right-hand
    corners vertically by a multiple of 1/64 pixel (< 1 pixel): a different transform per frame, the same 1728x1080 output
    window (the extremes that define it, H.js:1503-1527, are the untouched left corners and the right edge)."""
    d = np.tile(np.asarray(wl["dst"], np.float64), (n_frames, 1))
    if wl["kind"] == 1:
        j = ((np.arange(first, first + n_frames) % 64) / 64.0)
        d[:, 5] += j      # (w, h/4)  -> y + j
        d[:, 7] -= j      # (w, 3h/4) -> y - j
    return d


def config_dict(wl):
    """Identical in both arms (the driver compares it)."""
    return {"workload": wl["name"],
            "protocol": "per frame: calculateTransformMatrix(dst_f, src) + _inverseGeometricWarp (test/benchmark.js:96-113)",
            "frame_points": "frame 0 = BASELINE config 2; frame f moves the right-hand corners by (f mod 64)/64 px"}


def cpu_port_rate(wl, threads: int, budget_s: float, frames_per_round: int = 4, seed: int = 2):
    """The reference's loop (oracle port, C, -O2, unfused doubles) on the host cores: Mpix/s over a bounded sample."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (wl["H"], wl["W"], 4), dtype=np.uint8)
    kind = "projective" if wl["kind"] else "affine"
    px = 0
    t0 = time.perf_counter()
    n = 0
    while True:
        pts = headline_points(wl, frames_per_round, n)
        for k in range(frames_per_round):
            inv = O.calculate_transform_matrix(kind, pts[k], wl["src"])   # per-frame solve, like _inverseGeometricWarp
            O.warp_inverse_geometric(img, wl["W"], wl["H"], inv, wl["x_off"], wl["y_off"], wl["o_w"], wl["o_h"], threads=threads)
            px += wl["o_w"] * wl["o_h"]
            n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return px / dt / 1e6, n, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Node.js is not in this image and
    the reference is JavaScript (nothing gcc could compile into oracle/_ref), so this is the oracle PORT of
    _inverseGeometricWarp with all host threads (the real reference is single-threaded; see cpu_baseline.single_thread)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import homography_js_b200 as hg
    from oracle import oracle as O
    O.build()
    wl = hg.workloads.projective_1080p()
    threads = os.cpu_count() or 1
    frames = args.ref_frames
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (wl["H"], wl["W"], 4), dtype=np.uint8)
    state = {"n": 0}

    def step():
        pts = headline_points(wl, frames, state["n"])
        state["n"] += frames
        for k in range(frames):
            inv = O.projective_from_squares(pts[k], wl["src"])
            O.warp_inverse_geometric(img, wl["W"], wl["H"], inv, wl["x_off"], wl["y_off"], wl["o_w"], wl["o_h"], threads=threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    px = frames * wl["o_w"] * wl["o_h"] * args.steps
    val = px / dt / 1e6
    one_t, _, _ = cpu_port_rate(wl, 1, 1.5, 1)
    line = {
        "impl": "reference", "metric": "Mpix/s warped", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(wl),
        "step": {"frames_per_step": frames,
                 "note": "oracle port of Homography.js _inverseGeometricWarp (C, unfused doubles); Node.js absent from the image"},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "port",
                         "sample": f"{frames} frames x {args.steps} steps of 1728x1080, per-frame solve, OpenMP over output rows",
                         "single_thread": one_t},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# process environment shared by every workload of one bench process

class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import homography_js_b200 as hg
        self.args, self.torch, self.dist, self.hg = args, torch, dist, hg
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:   # frames are independent: NCCL carries the barrier, the reductions below and config 5's broadcast
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = hg.Context(self.local_rank)
        self.peak, self.peak_src = measured_peak_gbs()
        self._oracle = None

    @property
    def O(self):
        if self._oracle is None:
            from oracle import oracle as O
            O.build()
            self._oracle = O
        return self._oracle

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v: float, op: str) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def reduce_u64_sum(self, v: int) -> int:
        """sum mod 2^64 over ranks (two 32-bit halves through an int64 all-reduce)."""
        if self.world == 1:
            return v & 0xFFFFFFFFFFFFFFFF
        t = self.torch.tensor(list(self.hg.workloads.u64_to_halves(v)), dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return self.hg.workloads.halves_to_u64(int(t[0].item()), int(t[1].item()))

    def oracle_threads(self) -> int:
        return max(1, (os.cpu_count() or 1) // max(self.world, 1))

    def timed(self, fn, steps: int, warmup: int):
        """W warm-up steps, then exactly `steps` steps between barriers, CUDA events on the context stream; returns the max
        over ranks of the elapsed ms, the per-kernel profile (pixel-loop kernels) and the launches counted."""
        ctx = self.ctx
        for _ in range(warmup):
            fn()
        ctx.synchronize()
        sampler = ClockSampler(self.local_rank)
        self.barrier()
        sampler.start()
        l0 = ctx.launch_count()
        ctx.profile_enable(True)
        t0 = time.perf_counter()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms_dev = ctx.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        sampler.sample_once()
        self.barrier()
        sampler.stop()
        kms, kn = ctx.profile_read()
        ctx.profile_enable(False)
        return {"ms": self.reduce(ms_dev, "max"), "ms_local": ms_dev, "wall_ms": self.reduce(wall, "max"), "kernel_ms": kms,
                "kernels": kn, "launches": int(ctx.launch_count() - l0), "clocks": sampler.summary()}

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()
        self.ctx.close()


def frac_of_peak(env, npix: float, ms: float, gpus: int = 1) -> float:
    return ALG_BYTES_PER_PIXEL * npix / (ms * 1e-3) / 1e9 / (env.peak * gpus) if ms > 0 else 0.0


# ----------------------------------------------------------------------------------------------------------------
# secondary workloads (each returns one dict; every rank runs it, rank 0 reports)

def sec_piecewise_batch(env, which: str, frames: int, steps: int, warmup: int):
    """BASELINE config 3 (10x10 grid, 162 triangles) or the config-4 mesh (64x64 grid, 7,938 triangles) at 3840x2160 through
    hg_warp_piecewise_inverse_batch: F frames per step per GPU, each with its own source image and its own phase of the
    sinusoid (per-frame matrices, triangle map and pixel loop, H.js:1033-1058)."""
    torch, hg, ctx = env.torch, env.hg, env.ctx
    w, h = 3840, 2160
    nx = ny = 10 if which == "piecewise3" else 64
    F = frames
    src, _, tris = hg.workloads.piecewise_sinusoid(nx, ny, w, h)
    ctx.piecewise_set_mesh(src, tris)
    g = torch.Generator(device=env.dev)
    g.manual_seed(3 + env.rank)
    n_src = min(F, 8)
    src_ring = torch.randint(0, 256, (n_src, h * w * 4), dtype=torch.uint8, device=env.dev, generator=g)
    dsts, fr, outs, npix = [], [], [], 0
    lo, _ = hg.workloads.shard_range(F * env.world, env.rank, env.world)
    for f in range(F):
        phase = 2 * np.pi * (lo + f) / max(F * env.world, 1)
        _, dst, _ = hg.workloads.piecewise_sinusoid(nx, ny, w, h, phase=phase)
        xo, yo, oW, oH = hg.workloads.piecewise_extent(dst)
        o = torch.zeros(oW * oH * 4, dtype=torch.uint8, device=env.dev)
        outs.append(o)
        dsts.append(dst)
        fr.append(hg.HgFrame(src_ring[f % n_src].data_ptr(), o.data_ptr(), w, h, xo, yo, oW, oH))
        npix += oW * oH
    dst_all = np.stack(dsts)
    arr = (hg.HgFrame * F)(*fr)
    torch.cuda.synchronize()
    res = {}
    for mode in ("fused", "general") if env.args.general else ("fused",):
        ctx.debug_force_general(mode == "general")
        t = env.timed(lambda: ctx.warp_piecewise_inverse_batch(dst_all, arr, 0, 0), steps, warmup)
        step_ms = max(t["ms"], t["wall_ms"]) / steps     # the call ends with a status read-back: device and wall agree
        res[mode] = {"ms_per_step": step_ms, "pixel_kernel_ms_per_step": t["kernel_ms"] / steps, "launches": t["launches"],
                     "clocks": t["clocks"]}
    ctx.debug_force_general(False)
    # parity: frame 0 and the last frame of this rank pixel by pixel, every frame by checksum of its neighbours' kind
    O = env.O
    ok = True
    for f in sorted({0, F - 1}):
        xo, yo, oW, oH = fr[f].x_off, fr[f].y_off, fr[f].o_w, fr[f].o_h
        fwd = O.piecewise_matrices(src, dsts[f], tris)
        imap = O.build_index_map(dsts[f], tris, oW, yo, oW * oH)
        want = O.warp_inverse_piecewise(src_ring[f % n_src].cpu().numpy(), w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, 0, 0,
                                        threads=env.oracle_threads())
        ok = ok and bool(np.array_equal(outs[f].cpu().numpy(), want))
    ok = env.reduce(1.0 if ok else 0.0, "min") == 1.0
    npix_all = env.reduce(npix, "sum")
    fu = res["fused"]
    out = {"name": which, "workload": f"piecewiseaffine {nx}x{ny} grid ({len(tris)} tris), 3840x2160 RGBA8, inverse path, "
                                       f"{F} frames/step per GPU (own source + own sinusoid phase per frame)",
           "value": npix_all / (fu["ms_per_step"] * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": env.world, "steps": steps,
           "ms_per_step": fu["ms_per_step"], "frames_per_step_per_gpu": F, "parity_gate": ok,
           "roofline_frac_whole_step": frac_of_peak(env, npix_all, fu["ms_per_step"], env.world),
           "roofline_frac_pixel_kernel": frac_of_peak(env, npix, fu["pixel_kernel_ms_per_step"]),
           "pixel_kernel_ms_per_step": fu["pixel_kernel_ms_per_step"], "gpu_launches": fu["launches"], "clocks": fu["clocks"]}
    if "general" in res:
        out["general_path_ms_per_step"] = res["general"]["ms_per_step"]
    del src_ring, outs
    torch.cuda.empty_cache()
    return out


def _stream_verify(env, run_chunk, n_frames, n_slots, slot_bytes, ring_ptr, oracle_frame, sample):
    """Checksum gate of a streamed config: re-runs the stream in ring-sized calls (no slot is reused inside a call), takes
    the 64-bit checksum of EVERY frame on the device, compares the frames in `sample` with the oracle's checksum (and the
    first of them pixel by pixel), and returns (all ok, sum of all checksums mod 2^64)."""
    hg, ctx = env.hg, env.ctx
    sums = np.zeros(n_frames, np.uint64)
    first_pixels = None
    for f0 in range(0, n_frames, n_slots):
        nf = min(n_slots, n_frames - f0)
        info = run_chunk(f0, nf)
        frames = [hg.HgFrame(None, ring_ptr + I.slot * slot_bytes, 0, 0, I.x_off, I.y_off, I.o_w, I.o_h) for I in info]
        sums[f0:f0 + nf] = ctx.checksum_frames(frames)
        if first_pixels is None and sample and f0 <= sample[0] < f0 + nf:
            I = info[sample[0] - f0]
            buf = np.empty(I.o_w * I.o_h * 4, np.uint8)
            ctx.memcpy_d2h(buf.ctypes.data, ring_ptr + I.slot * slot_bytes, buf.size)
            ctx.synchronize()
            first_pixels = ((I.x_off, I.y_off, I.o_w, I.o_h), buf)
    ok = True
    for k, f in enumerate(sample):
        win, want = oracle_frame(f)
        ok = ok and int(sums[f]) == hg._abi.checksum_reference(want)
        if k == 0 and first_pixels is not None:
            ok = ok and first_pixels[0] == tuple(win) and bool(np.array_equal(first_pixels[1], want))
    total = 0
    for v in sums:
        total = (total + int(v)) & 0xFFFFFFFFFFFFFFFF
    return ok, total


def sec_config4(env, frames: int, steps: int, warmup: int):
    """BASELINE config 4: piecewise sinusoid over a 64x64 grid (7,938 triangles), 3840x2160, `frames` frames per GPU per step
    (4096 over 8 GPUs), frame f of the whole job using phase 2 pi f / total (distinct destiny points, matrices and triangle
    map per frame, H.js:1033-1038), streamed through hg_warp_piecewise_stream: output windows on the device, sources from a
    ring of 8 distinct images per GPU, outputs into a ring of 128 slots (4.8 GB: chunks of 64 frames per launch chain)."""
    torch, hg, ctx = env.torch, env.hg, env.ctx
    w, h, nx = 3840, 2160, 64
    total = frames * env.world
    lo, hi = hg.workloads.shard_range(total, env.rank, env.world)
    F = hi - lo
    src, _, tris = hg.workloads.piecewise_sinusoid(nx, nx, w, h)
    A = h / 20.0
    xs = src[:, 0].astype(np.float64)
    ph = 2 * np.pi * np.arange(lo, hi) / total
    dst_all = np.repeat(src[None, :, :], F, axis=0).copy()
    dst_all[:, :, 1] = (A + src[None, :, 1].astype(np.float64) + A * np.sin(2 * math.pi * 2 * xs[None, :] / w + ph[:, None])).astype(np.float32)
    ctx.piecewise_set_mesh(src, tris)
    n_src, n_slots = 8, env.args.c4_slots
    max_w, max_h = w + 8, int(h + 2 * A) + 16
    slot = ctx.stream_slot_bytes(max_w, max_h)
    g = torch.Generator(device=env.dev)
    g.manual_seed(4 + env.rank)
    src_ring = torch.randint(0, 256, (n_src, h * w * 4), dtype=torch.uint8, device=env.dev, generator=g)
    ring = torch.zeros(n_slots * slot, dtype=torch.uint8, device=env.dev)
    torch.cuda.synchronize()
    info = (hg.HgStreamInfo * F)()

    def step():
        ctx.warp_piecewise_stream(dst_all, lo, 0, 0, ring.data_ptr(), n_slots, max_w, max_h, src_ring_dev=src_ring.data_ptr(),
                                  n_src=n_src, src_w=w, src_h=h, info=info)

    t = env.timed(step, steps, warmup)
    npix = sum(I.o_w * I.o_h for I in info)
    skipped = sum(1 for I in info if I.status != 0)
    step_ms = max(t["ms"], t["wall_ms"]) / steps
    O = env.O
    host_src = {}

    def oracle_frame(f):
        k = (lo + f) % n_src
        if k not in host_src:
            host_src[k] = src_ring[k].cpu().numpy()
        xo, yo, oW, oH = hg.workloads.piecewise_extent(dst_all[f])
        fwd = O.piecewise_matrices(src, dst_all[f], tris)
        imap = O.build_index_map(dst_all[f], tris, oW, yo, oW * oH)
        return (xo, yo, oW, oH), O.warp_inverse_piecewise(host_src[k], w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, 0, 0,
                                                          threads=env.oracle_threads())

    def run_chunk(f0, nf):
        return ctx.warp_piecewise_stream(dst_all[f0:f0 + nf], lo + f0, 0, 0, ring.data_ptr(), n_slots, max_w, max_h,
                                         src_ring_dev=src_ring.data_ptr(), n_src=n_src, src_w=w, src_h=h)

    ok, total_cs = _stream_verify(env, run_chunk, F, n_slots, slot, ring.data_ptr(), oracle_frame, sorted({0, F // 2, F - 1}))
    ok = env.reduce(1.0 if (ok and skipped == 0) else 0.0, "min") == 1.0
    npix_all = env.reduce(npix, "sum")
    out = {"name": "config4", "workload": f"piecewiseaffine sinusoid 64x64 grid ({len(tris)} tris), 3840x2160 RGBA8, {total} frames "
                                          f"sharded over {env.world} GPU(s) ({F} per GPU per step), per-frame phase, streamed "
                                          f"(windows on the device, ring of {n_src} sources, {n_slots} output slots)",
           "value": npix_all / (step_ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": env.world, "steps": steps, "ms_per_step": step_ms,
           "frames_per_step_per_gpu": F, "frames_per_step_total": total,
           "checksum_gate": ok, "checksum_gate_note": "64-bit checksum of every frame on the device; 3 frames per rank against the "
                                                      "oracle's checksum and one of them pixel by pixel; all ranks must pass",
           "checksum_of_checksums": "%016x" % env.reduce_u64_sum(total_cs),
           "roofline_frac_whole_step": frac_of_peak(env, npix_all, step_ms, env.world),
           "roofline_frac_pixel_kernel": frac_of_peak(env, npix, t["kernel_ms"] / steps),
           "pixel_kernel_ms_per_step": t["kernel_ms"] / steps, "gpu_launches": t["launches"], "clocks": t["clocks"]}
    del src_ring, ring
    torch.cuda.empty_cache()
    return out


def sec_config5(env, frames: int, steps: int, warmup: int):
    """BASELINE config 5: video stream, 1920x1080, 30-point piecewise mesh, new destiny points every frame (a different
    output window every frame), `frames` frames per GPU per step (100,000 over 8 GPUs), ONE source image shared by all
    GPUs: rank 0 owns it and it is broadcast over NCCL (NVLink) — the one collective of the design.  Per frame, on the
    device and inside the timed step: output window (H.js:706-710), placement in the output ring, per-triangle matrices
    and inverses, triangle map, pixel loop."""
    torch, dist, hg, ctx = env.torch, env.dist, env.hg, env.ctx
    w, h = 1920, 1080
    total = frames * env.world
    src_pts, dst_total, tris = hg.workloads.video_stream(total, w, h)
    lo, hi = hg.workloads.shard_range(total, env.rank, env.world)
    F = hi - lo
    dst_all = np.ascontiguousarray(dst_total[lo:hi])
    del dst_total
    image = torch.zeros(h * w * 4, dtype=torch.uint8, device=env.dev)
    if env.rank == 0:
        g = torch.Generator(device=env.dev)
        g.manual_seed(5)
        image = torch.randint(0, 256, (h * w * 4,), dtype=torch.uint8, device=env.dev, generator=g)
    bcast = None
    if env.world > 1:
        dummy = torch.zeros(1024, dtype=torch.uint8, device=env.dev)
        dist.broadcast(dummy, src=0)                       # communicator warm-up, not timed
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        dist.broadcast(image, src=0)                      # W*H*4 bytes from rank 0 to every GPU
        e1.record()
        torch.cuda.synchronize()
        ms = env.reduce(e0.elapsed_time(e1), "max")
        bcast = {"ms": ms, "bytes": h * w * 4, "GB/s": h * w * 4 / (ms * 1e-3) / 1e9}
    ctx.image_set_device(image.data_ptr(), w, h)
    ctx.piecewise_set_mesh(src_pts, tris)
    smm = [int(np.floor(v + 0.5)) for v in (float(src_pts[:, 0].min()), float(src_pts[:, 1].min()))]
    # output ring: 2,048 slots of 9.6 MB by default = chunks of 1,024 frames per launch chain (measured, 12,500 frames per step:
    # 256 / 512 / 2,048 slots -> 0.499 / 0.519 / 0.528 of the HBM bound)
    n_slots = env.args.c5_slots
    max_w, max_h = int(w * 1.07) + 8, int(h * 1.07) + 8      # points move by +-3 % of the frame
    slot = ctx.stream_slot_bytes(max_w, max_h)
    ring = torch.zeros(n_slots * slot, dtype=torch.uint8, device=env.dev)
    torch.cuda.synchronize()
    info = (hg.HgStreamInfo * F)()
    f0s, g0s = ctx.debug_piecewise_stats()

    def step():
        ctx.warp_piecewise_stream(dst_all, lo, smm[0], smm[1], ring.data_ptr(), n_slots, max_w, max_h, info=info)

    t = env.timed(step, steps, warmup)
    f1s, g1s = ctx.debug_piecewise_stats()
    npix = sum(I.o_w * I.o_h for I in info)
    skipped = sum(1 for I in info if I.status != 0)
    windows = len({(I.o_w, I.o_h) for I in info})
    step_ms = max(t["ms"], t["wall_ms"]) / steps
    O = env.O
    host_img = image.cpu().numpy()

    def oracle_frame(f):
        xo, yo, oW, oH = hg.workloads.piecewise_extent(dst_all[f])
        fwd = O.piecewise_matrices(src_pts, dst_all[f], tris)
        imap = O.build_index_map(dst_all[f], tris, oW, yo, oW * oH)
        return (xo, yo, oW, oH), O.warp_inverse_piecewise(host_img, w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, smm[0], smm[1],
                                                          threads=env.oracle_threads())

    def run_chunk(c0, nf):
        return ctx.warp_piecewise_stream(dst_all[c0:c0 + nf], lo + c0, smm[0], smm[1], ring.data_ptr(), n_slots, max_w, max_h)

    ok, total_cs = _stream_verify(env, run_chunk, F, n_slots, slot, ring.data_ptr(), oracle_frame,
                                  sorted({0, F // 3, (2 * F) // 3, F - 1}))
    ok = env.reduce(1.0 if (ok and skipped == 0) else 0.0, "min") == 1.0
    npix_all = env.reduce(npix, "sum")
    out = {"name": "config5", "workload": f"video: 1920x1080 RGBA8, 30-pt piecewise ({len(tris)} tris), per-frame dstPoints, {total} frames "
                                          f"sharded over {env.world} GPU(s) ({F} per GPU per step), one shared source image, "
                                          f"{n_slots}-slot output ring, {windows} distinct output windows on this rank",
           "value": npix_all / (step_ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": env.world, "steps": steps, "ms_per_step": step_ms,
           "frames_per_step_per_gpu": F, "frames_per_step_total": total, "frames_per_s": total / (step_ms * 1e-3),
           "checksum_gate": ok, "checksum_gate_note": "64-bit checksum of every frame on the device; 4 frames per rank against the "
                                                      "oracle's checksum and one of them pixel by pixel; all ranks must pass",
           "checksum_of_checksums": "%016x" % env.reduce_u64_sum(total_cs),
           "frames_fused_vs_general": [int(f1s - f0s), int(g1s - g0s)],
           "nccl_broadcast": bcast, "nccl_broadcast_ms": None if bcast is None else bcast["ms"],
           "in_timed_step": "per-frame extent (device) + ring placement + A6/A8 solves + triangle map + pixel loop; dstPoints H2D",
           "roofline_frac_whole_step": frac_of_peak(env, npix_all, step_ms, env.world),
           "roofline_frac_pixel_kernel": frac_of_peak(env, npix, t["kernel_ms"] / steps),
           "gpu_launches": t["launches"], "clocks": t["clocks"]}
    # host -> host: the same stream with results landing in pinned host memory (points in, pixels out), pipelined
    Fe = min(F, 64)
    pipe = hg.Pipe.piecewise(ctx, w, h, max_w, max_h, depth=4)
    h_out = [ctx.pinned_array(max_w * max_h * 4) for _ in range(4)]

    def e2e_step():
        px = 0
        for f in range(Fe):
            _, win = pipe.submit_piecewise(None, dst_all[f], smm[0], smm[1], h_out[f % 4].ctypes.data)
            px += win[2] * win[3]
        pipe.flush()
        return px

    for _ in range(2):
        e2e_px = e2e_step()
    env.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(steps, 10))
    for _ in range(e2e_steps):
        e2e_step()
    e2e_ms = env.reduce((time.perf_counter() - t0) * 1e3, "max") / e2e_steps
    env.barrier()
    # the last frame that landed in host memory is the oracle's
    win, want = oracle_frame(Fe - 1)
    e2e_ok = bool(np.array_equal(h_out[(Fe - 1) % 4][: win[2] * win[3] * 4], want))
    e2e_ok = env.reduce(1.0 if e2e_ok else 0.0, "min") == 1.0
    out["e2e"] = {"value": env.reduce(e2e_px, "sum") / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "parity_gate": e2e_ok,
                  "api": "hg_pipe_submit_piecewise per frame (dstPoints H2D + window + warp + D2H of the result, 4 frames in flight), "
                         "image resident like the reference's setImage-once video loop",
                  "frames_per_step_per_gpu": Fe, "d2h_bytes_per_step": int(e2e_px) * 4, "h2d_bytes_per_step": Fe * dst_all[0].size * 4}
    pipe.close()
    del ring, image
    torch.cuda.empty_cache()
    return out


def sec_forward(env, which: str, frames: int, steps: int, warmup: int):
    """The forward loops (H.js:911 / H.js:948) as one batch per step, device-resident rings > L2:
       affine_forward          _geometricWarp, 1080p translation by (100, 50) — the case warp() dispatches to the forward loop
                               (oW == W, oH == H; test.js:167-192 shape): a lattice-preserving map, one gather pass
       affine_forward_general  the same frames under a 1.5 degree rotation: collisions and holes, the deterministic
                               scatter (atomic max of the loop-order key) + gather passes
       piecewise_forward       _piecewiseAffineWarp: 1920x1080, 10x10 grid, destiny points shrunk to 0.9x (+ a per-frame
                               wobble) so that warp() picks the forward loop (output within [W/1.2, W] x [H/1.2, H], H.js:421)
                               — the reference's own 400x400 -> 400x400 piecewise benchmark shape at 1080p"""
    torch, hg, ctx = env.torch, env.hg, env.ctx
    O = env.O
    W, H, F = 1920, 1080, frames
    g = torch.Generator(device=env.dev)
    g.manual_seed(7 + env.rank)
    src_ring = torch.randint(0, 256, (F, H * W * 4), dtype=torch.uint8, device=env.dev, generator=g)
    if which in ("affine_forward", "affine_forward_general"):
        if which == "affine_forward":
            src_pts = np.array([0, 0, 0, H, W, 0], np.float64)
            fwd = ctx.solve_affine(src_pts, src_pts + np.array([100, 50] * 3, np.float64))   # [1,0,0,1,100,50]
            win = (100, 50, W, H)
        else:
            a = math.radians(1.5)
            fwd = np.array([math.cos(a), math.sin(a), -math.sin(a), math.cos(a), 30.0, 10.0], np.float32)
            lim = O.transform_limits(fwd, W, H)
            win = (int(lim[0]), int(lim[1]), int(lim[2]), int(lim[3]))
        oW, oH = win[2], win[3]
        out_ring = torch.zeros((F, oW * oH * 4), dtype=torch.uint8, device=env.dev)
        fr = (hg.HgFrame * F)(*[hg.HgFrame(src_ring[f].data_ptr(), out_ring[f].data_ptr(), W, H, *win) for f in range(F)])
        mats = np.tile(fwd, (F, 1))
        torch.cuda.synchronize()
        step = lambda: ctx.warp_forward_batch(0, mats, fr)
        step()
        ctx.synchronize()
        ok = True
        for f in sorted({0, F - 1}):
            want = O.warp_forward_geometric(src_ring[f].cpu().numpy(), W, H, fwd, *win)
            ok = ok and bool(np.array_equal(out_ring[f].cpu().numpy(), want))
        name = ("affine forward loop (translation by (100,50): lattice plan, one gather pass)" if which == "affine_forward"
                else "affine forward loop (1.5 degree rotation: scatter of loop-order keys + gather)") + \
               f", 1920x1080 RGBA8 -> {oW}x{oH}, {F} frames/step per GPU"
        npix = F * oW * oH
    else:
        src, tris = hg.workloads.grid_mesh(10, 10, W, H)
        ctx.piecewise_set_mesh(src, tris)
        smm = [int(v) for v in O.minmax_xy(src)]
        rng = np.random.default_rng(11 + env.rank)
        dsts, wins = [], []
        for f in range(F):
            wob = 6.0 * np.sin(0.3 * f + np.arange(len(src)))[:, None] * np.array([1.0, 0.6])
            interior = ((src[:, 0] > 0) & (src[:, 0] < W) & (src[:, 1] > 0) & (src[:, 1] < H))[:, None]
            dst = (src.astype(np.float64) * 0.9 + 20 + wob * interior).astype(np.float32)
            dsts.append(dst)
            mm = O.minmax_xy(dst)
            wins.append((int(mm[0]), int(mm[1]), int(mm[2] - mm[0]), int(mm[3] - mm[1])))
        outs = [torch.zeros(wn[2] * wn[3] * 4, dtype=torch.uint8, device=env.dev) for wn in wins]
        fr = (hg.HgFrame * F)(*[hg.HgFrame(src_ring[f].data_ptr(), outs[f].data_ptr(), W, H, *wins[f]) for f in range(F)])
        dst_all = np.stack(dsts)
        torch.cuda.synchronize()
        step = lambda: ctx.warp_piecewise_forward_batch(dst_all, fr, smm[0], smm[1], smm[2], smm[3])
        step()
        ctx.synchronize()
        mw = smm[2] - smm[0]
        fmap = O.build_index_map(src, tris, mw, smm[1], mw * (smm[3] - smm[1]))
        ok = True
        for f in sorted({0, F - 1}):
            want = O.warp_forward_piecewise(src_ring[f].cpu().numpy(), W, H, fmap, O.piecewise_matrices(src, dsts[f], tris), *wins[f],
                                            smm[0], smm[1], smm[2], smm[3])
            ok = ok and bool(np.array_equal(outs[f].cpu().numpy(), want))
        name = f"piecewise forward loop (_piecewiseAffineWarp), 1920x1080 RGBA8, 10x10 grid ({len(tris)} tris) -> ~{wins[0][2]}x{wins[0][3]}, " \
               f"{F} frames/step per GPU"
        npix = sum(wn[2] * wn[3] for wn in wins)
    ok = env.reduce(1.0 if ok else 0.0, "min") == 1.0
    t = env.timed(step, steps, warmup)
    npix_all = env.reduce(npix, "sum")
    ms = t["ms"] / steps
    out = {"name": which, "workload": name, "value": npix_all / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": env.world,
           "steps": steps, "ms_per_step": ms, "frames_per_step_per_gpu": F, "parity_gate": ok,
           "roofline_frac_whole_step": frac_of_peak(env, npix_all, ms, env.world), "gpu_launches": t["launches"],
           "clocks": t["clocks"]}
    del src_ring
    torch.cuda.empty_cache()
    return out


def sec_bilinear(env, frames: int, steps: int, warmup: int):
    """Config 2 through the bilinear extension (the reference has Math.round sampling only); gate: <= 1 LSB per channel
    against the oracle's f64 definition."""
    torch, hg, ctx = env.torch, env.hg, env.ctx
    O = env.O
    wl = hg.workloads.projective_1080p()
    W, H, oW, oH, F = wl["W"], wl["H"], wl["o_w"], wl["o_h"], frames
    g = torch.Generator(device=env.dev)
    g.manual_seed(8 + env.rank)
    inv = ctx.solve_projective(wl["dst"], wl["src"])
    src_ring = torch.randint(0, 256, (F, H * W * 4), dtype=torch.uint8, device=env.dev, generator=g)
    out_ring = torch.zeros((F, oW * oH * 4), dtype=torch.uint8, device=env.dev)
    mats = np.tile(inv, (F, 1))
    fr = [hg.HgFrame(src_ring[f].data_ptr(), out_ring[f].data_ptr(), W, H, wl["x_off"], wl["y_off"], oW, oH) for f in range(F)]
    ctx.set_sampling(hg._abi.HG_BILINEAR)
    torch.cuda.synchronize()
    try:
        step = lambda: ctx.warp_inverse_batch(1, mats, fr)
        step()
        ctx.synchronize()
        want = O.warp_inverse_geometric_bilinear(src_ring[0].cpu().numpy(), W, H, O.projective_from_squares(wl["dst"], wl["src"]),
                                                 wl["x_off"], wl["y_off"], oW, oH, threads=env.oracle_threads())
        got = out_ring[0].cpu().numpy()
        ok = bool(np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1)
        ok = env.reduce(1.0 if ok else 0.0, "min") == 1.0
        t = env.timed(step, steps, warmup)
    finally:
        ctx.set_sampling(hg._abi.HG_NEAREST)
    npix = F * oW * oH
    ms = t["ms"] / steps
    out = {"name": "projective_bilinear", "workload": f"config 2 with BILINEAR sampling (extension), {F} frames/step per GPU",
           "value": env.reduce(npix, "sum") / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": env.world, "steps": steps, "ms_per_step": ms,
           "frames_per_step_per_gpu": F, "parity_gate": ok, "parity_tolerance": "<= 1 LSB per channel",
           "roofline_frac_whole_step": frac_of_peak(env, env.reduce(npix, "sum"), ms, env.world), "gpu_launches": t["launches"],
           "clocks": t["clocks"]}
    del src_ring, out_ring
    torch.cuda.empty_cache()
    return out


def run_secondary(env, names, steps, warmup):
    a = env.args
    out = []
    for n in names:
        try:
            if n in ("piecewise3", "piecewise4"):
                r = sec_piecewise_batch(env, n, a.pw_frames, steps, warmup)
            elif n == "config4":
                r = sec_config4(env, a.c4_frames, steps, warmup)
            elif n == "config5":
                r = sec_config5(env, a.c5_frames, steps, warmup)
            elif n in ("affine_forward", "affine_forward_general", "piecewise_forward"):
                r = sec_forward(env, n, a.fwd_frames, steps, warmup)
            elif n == "projective_bilinear":
                r = sec_bilinear(env, a.fwd_frames, steps, warmup)
            else:
                raise ValueError(n)
        except Exception as e:  # a failing secondary must not take the headline line down with it — but it is reported
            if env.world > 1:
                raise
            r = {"name": n, "error": f"{type(e).__name__}: {e}"[:400]}
        out.append(r)
    return out


SECONDARY_DEFAULT = ["piecewise3", "config4", "config5", "affine_forward", "affine_forward_general", "piecewise_forward",
                     "projective_bilinear"]


# ----------------------------------------------------------------------------------------------------------------
def run_headline(env, wl_name: str, with_secondary: bool):
    args, torch, hg, ctx = env.args, env.torch, env.hg, env.ctx
    rank, world = env.rank, env.world
    wl = {"projective": hg.workloads.projective_1080p, "affine": hg.workloads.affine_1080p,
          "projective_generic": hg.workloads.projective_1080p_generic,
          "affine_rot90": hg.workloads.affine_1080p_rot90}[wl_name]()
    KIND = wl["kind"]
    kind_name = "projective" if KIND else "affine"
    W, H, oW, oH = wl["W"], wl["H"], wl["o_w"], wl["o_h"]
    F = args.frames
    npix_frame = oW * oH

    # ---- device-resident rings (torch = device-memory plumbing only)
    g = torch.Generator(device=env.dev)
    g.manual_seed(2 + rank)
    src_ring = torch.randint(0, 256, (F, H * W * 4), dtype=torch.uint8, device=env.dev, generator=g)
    out_ring = torch.zeros((F, npix_frame * 4), dtype=torch.uint8, device=env.dev)
    lo, _ = hg.workloads.shard_range(F * world, rank, world)
    dst_pts = headline_points(wl, F, lo)                 # this rank's frames of the job
    src_pts = np.tile(np.asarray(wl["src"], np.float64), (F, 1))
    frames = (hg.HgFrame * F)(*[hg.HgFrame(src_ring[f].data_ptr(), out_ring[f].data_ptr(), W, H, wl["x_off"], wl["y_off"], oW, oH)
                                for f in range(F)])
    torch.cuda.synchronize()

    def step():
        # per frame: inverse matrix = calculateTransformMatrix(kind, dst_f, src) on the device (H.js:994), then the pixel loop
        ctx.warp_inverse_points_batch(KIND, dst_pts, src_pts, frames)

    # ---- parity gate inside the run: first, second and last frame of the batch against the oracle (not timed)
    step()
    ctx.synchronize()
    parity = None
    if not args.no_cpu_baseline:
        O = env.O
        ok = True
        for f in sorted({0, 1 % F, F - 1}):
            inv = O.calculate_transform_matrix(kind_name, dst_pts[f], wl["src"])
            want = O.warp_inverse_geometric(src_ring[f].cpu().numpy(), W, H, inv, wl["x_off"], wl["y_off"], oW, oH,
                                            threads=env.oracle_threads())
            ok = ok and bool(np.array_equal(out_ring[f].cpu().numpy(), want))
        parity = env.reduce(1.0 if ok else 0.0, "min") == 1.0
        if not parity:
            raise SystemExit("parity gate failed: CUDA output differs from the oracle")

    # ---- device-resident timing
    t = env.timed(step, args.steps, args.warmup)
    ms = t["ms"]
    value = world * F * npix_frame * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (warp_inverse_geo_kernel), live CUDA events around each launch
    px_per_launch = F * npix_frame
    avg_kernel_s = (t["kernel_ms"] / max(t["kernels"], 1)) * 1e-3
    achieved = ALG_BYTES_PER_PIXEL * px_per_launch / avg_kernel_s / 1e9
    kname = "warp_inverse_geo_kernel<%s>" % kind_name
    traffic, traffic_src = measured_traffic(kname, F)
    # the same batch with EVERY frame on the exact config-2 points (round 1's workload): ten whole output columns then sit
    # exactly on rounding boundaries in every frame, the kernel's worst case (exact-resolution queue)
    exact = None
    if wl_name == "projective":
        dst_exact = np.tile(np.asarray(wl["dst"], np.float64), (F, 1))
        t2 = env.timed(lambda: ctx.warp_inverse_points_batch(KIND, dst_exact, src_pts, frames), max(3, args.steps // 2), 3)
        k2 = (t2["kernel_ms"] / max(t2["kernels"], 1)) * 1e-3
        exact = {"frac": ALG_BYTES_PER_PIXEL * px_per_launch / k2 / 1e9 / env.peak, "avg_kernel_ms": k2 * 1e3,
                 "value": world * F * npix_frame * t2["kernels"] / (t2["ms"] * 1e-3) / 1e6,
                 "note": "every frame = the exact BASELINE config-2 points (round 1's batch): ten output columns on rounding boundaries"}
        step()   # leave the rings holding the per-frame-points result
        ctx.synchronize()

    # ---- end to end.  (a) the class surface a user of the reference calls: setImage + setDestinyPoints + warp per frame,
    #      synchronous, frame images in pinned host memory, results in pinned host memory (H2D + solve + limits + warp + D2H)
    Fe = args.e2e_frames
    h_src = [ctx.pinned_array(H * W * 4, write_combined=args.wc_src) for _ in range(Fe)]
    rs = np.random.default_rng(100 + rank)
    for a in h_src:
        a[:] = rs.integers(0, 256, a.size, dtype=np.uint8)
    e_pts = headline_points(wl, Fe, 0)
    hom = hg.Homography(kind_name, context=ctx, pinned_output=True)
    hom.setSourcePoints(np.asarray(wl["src"], np.float64).copy(), None, W, H, False)
    last = {}

    def e2e_class_step():
        for f in range(Fe):
            hom.setImage(hg.ImageData(h_src[f], W, H))
            hom.setDestinyPoints(e_pts[f].copy(), False)
            last["out"] = hom.warp()

    # (b) the streaming C ABI underneath: hg_pipe_submit per frame, 4 frames in flight
    h_out = [ctx.pinned_array(npix_frame * 4) for _ in range(Fe)]
    pipe = hg.Pipe(ctx, KIND, W, H, oW, oH, depth=4)

    def e2e_pipe_step():
        for f in range(Fe):
            pipe.submit(h_src[f].ctypes.data, e_pts[f], wl["src"], wl["x_off"], wl["y_off"], oW, oH, h_out[f].ctypes.data)
        pipe.flush()

    def time_e2e(fn, steps):
        for _ in range(3):
            fn()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ctx.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        env.barrier()
        return env.reduce(wall_ms, "max")

    e2e_steps = args.steps
    class_ms = time_e2e(e2e_class_step, e2e_steps)
    class_val = world * Fe * npix_frame * e2e_steps / (class_ms * 1e-3) / 1e6
    pipe_ms = time_e2e(e2e_pipe_step, e2e_steps)
    pipe_val = world * Fe * npix_frame * e2e_steps / (pipe_ms * 1e-3) / 1e6
    e2e_ok = None
    if not args.no_cpu_baseline:
        O = env.O
        inv = O.calculate_transform_matrix(kind_name, e_pts[Fe - 1], wl["src"])
        want = O.warp_inverse_geometric(h_src[Fe - 1], W, H, inv, wl["x_off"], wl["y_off"], oW, oH, threads=env.oracle_threads())
        o = last["out"]
        ok = (o.width, o.height) == (oW, oH) and bool(np.array_equal(np.asarray(o.data), want)) and \
            bool(np.array_equal(h_out[Fe - 1], want))
        e2e_ok = env.reduce(1.0 if ok else 0.0, "min") == 1.0
        if not e2e_ok:
            raise SystemExit("parity gate failed: host-to-host output differs from the oracle")
    pipe.close()
    # (c) the ceiling of the host link, all ranks probing at the same moment: raw pinned copies, both directions at once
    ctx.pcie_probe(64 << 20, 1)       # the context pins the probe's buffers once: the timed call below starts copying at once
    env.barrier()
    p_h2d, p_d2h, p_bi = ctx.pcie_probe(64 << 20, 6)
    env.barrier()
    link = {"h2d_GBps_per_gpu_min": env.reduce(p_h2d, "min"), "d2h_GBps_per_gpu_min": env.reduce(p_d2h, "min"),
            "bidir_GBps_all_gpus": env.reduce(p_bi, "sum"), "how": "hg_pcie_probe: 6 x 64 MiB pinned copies per direction, "
            "alone and both directions at once, every rank at the same time"}
    bytes_per_px = (H * W * 4 + npix_frame * 4) / npix_frame

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v_all, n_all, dt_all = cpu_port_rate(wl, threads, args.cpu_budget * 0.8)
        v_one, n_one, dt_one = cpu_port_rate(wl, 1, args.cpu_budget * 0.2, 1)
        cpu = {"value": v_all, "unit": "Mpix/s", "cores": threads, "kind": "port",
               "sample": f"{n_all} frames of 1728x1080 in {dt_all:.1f} s, per-frame solve, OpenMP over output rows "
                         "(oracle port of Homography.js _inverseGeometricWarp; Node.js absent from the image)",
               "single_thread": v_one}

    del src_ring, out_ring
    torch.cuda.empty_cache()
    secondary = None
    if with_secondary:
        secondary = run_secondary(env, [s for s in SECONDARY_DEFAULT if s not in args.skip],
                                  max(3, min(args.steps, args.secondary_steps)), max(3, min(args.warmup, 5)))

    if rank == 0:
        line = {
            "metric": "Mpix/s warped", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl),
            "step": {"frames_per_step_per_gpu": F, "launches_per_step": t["launches"] // max(args.steps, 1),
                     "l2": f"ring of {F} distinct sources + {F} distinct outputs per GPU "
                           f"({F * (H * W * 4 + npix_frame * 4) / 1e6:.0f} MB) > 126 MB L2",
                     "timed_region_ms": ms},
            "parity_gate": parity,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": env.peak, "unit": "GB/s", "frac": achieved / env.peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": kname,
                         "alg_bytes_per_launch": ALG_BYTES_PER_PIXEL * px_per_launch,
                         "avg_kernel_ms": avg_kernel_s * 1e3, "kernels_timed": t["kernels"], "peak_source": env.peak_src,
                         "same_points_every_frame": exact},
            "cpu_baseline": cpu,
            "e2e": {"value": class_val, "unit": "Mpix/s", "h2d_bytes_per_step": Fe * H * W * 4,
                    "d2h_bytes_per_step": Fe * npix_frame * 4, "frames_per_step_per_gpu": Fe, "steps": e2e_steps,
                    "ms_per_step": class_ms / e2e_steps, "timer": "host wall clock, max over ranks", "parity_gate": e2e_ok,
                    "api": "Homography.setImage(frame) + setDestinyPoints(points) + warp() per frame — the reference's class surface "
                           "(H.js:290/337/408), synchronous; frames and results in pinned host memory",
                    "pipe": {"value": pipe_val, "unit": "Mpix/s", "ms_per_step": pipe_ms / e2e_steps,
                             "api": "hg_pipe_submit per frame (H2D image + solve + warp + D2H result, 4 frames in flight) + hg_pipe_flush"},
                    "host_link": link,
                    "frac_of_host_link": (class_val * 1e6 * bytes_per_px / 1e9) / link["bidir_GBps_all_gpus"]
                    if link["bidir_GBps_all_gpus"] > 0 else None,
                    "pipe_frac_of_host_link": (pipe_val * 1e6 * bytes_per_px / 1e9) / link["bidir_GBps_all_gpus"]
                    if link["bidir_GBps_all_gpus"] > 0 else None},
            "gpu_launches": int(t["launches"]),
            "clocks": t["clocks"],
        }
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="headline frames per step (per GPU)")
    ap.add_argument("--e2e-frames", type=int, default=16, help="frames per end-to-end step (per GPU)")
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the CPU reference arm")
    ap.add_argument("--pw-frames", type=int, default=128, help="frames per step per GPU of the piecewise3 / piecewise4 batches")
    ap.add_argument("--c4-frames", type=int, default=512, help="config 4: frames per GPU per step (4096 / 8)")
    ap.add_argument("--c5-frames", type=int, default=12500, help="config 5: frames per GPU per step (100000 / 8)")
    ap.add_argument("--c4-slots", type=int, default=128, help="config 4: slots of the output ring (a launch chain covers half of it)")
    ap.add_argument("--c5-slots", type=int, default=2048,
                    help="config 5: slots of the output ring (19.6 GB; a launch chain covers half of it, at most 1,024 frames)")
    ap.add_argument("--fwd-frames", type=int, default=64, help="frames per step per GPU of the forward / bilinear lines")
    ap.add_argument("--secondary-steps", type=int, default=10, help="timed steps of each secondary workload (at most --steps)")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle legs (parity gates + cpu_baseline)")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--wc-src", action="store_true", help="end-to-end source frames in write-combined pinned memory (A/B)")
    ap.add_argument("--skip", nargs="*", default=[], help="secondary workloads to leave out")
    ap.add_argument("--general", action="store_true", help="piecewise3/4: also time the general map-based path")
    ap.add_argument("--workload", default="projective",
                    choices=["projective", "affine", "projective_generic", "affine_rot90", "piecewise3", "piecewise4", "config4",
                             "config5", "video5", "affine_forward", "affine_forward_general", "piecewise_forward", "projective_bilinear"],
                    help="projective = BASELINE config 2 (the headline, with the secondary list); the others run alone")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return
    env = Env(args)
    try:
        if args.workload in ("projective", "affine", "projective_generic", "affine_rot90"):
            run_headline(env, args.workload, with_secondary=(args.workload == "projective" and not args.no_secondary))
        else:
            name = "config5" if args.workload == "video5" else args.workload
            r = run_secondary(env, [name], args.steps, args.warmup)[0]
            if env.rank == 0:
                r.setdefault("metric", "Mpix/s warped")
                print(json.dumps(r), flush=True)
    finally:
        env.close()


if __name__ == "__main__":
    main()
