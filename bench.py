#!/usr/bin/env python
"""bench.py — Mpix/s warped on B200 for BASELINE.json's headline config, with roofline + CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU loop (oracle port), rank 0 only

Workload at every N: config 2 of BASELINE.json — projective 4-point warp of 1920x1080 RGBA8 frames into their
1728x1080 output window.  One step = one batch of FRAMES independent frames (distinct source + distinct output
buffer per frame: a ring far larger than the 126 MB L2, so every step streams from / to HBM).  Frames are
independent, so N GPUs each warp their own batch (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_PIXEL = 8  # 4 B RGBA8 read + 4 B RGBA8 write per output pixel (SURVEY §8d / north_star)


def measured_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(kernel)
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
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


def _inv_matrix(O, wl):
    return O.calculate_transform_matrix("projective" if wl["kind"] else "affine", wl["dst"], wl["src"])


def cpu_port_rate(wl, threads: int, budget_s: float, frames_per_round: int = 4, seed: int = 2):
    """The reference's loop (oracle port, C, -O2, unfused doubles) on the host cores: Mpix/s over a bounded sample."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (wl["H"], wl["W"], 4), dtype=np.uint8)
    px = 0
    t0 = time.perf_counter()
    n = 0
    while True:
        for _ in range(frames_per_round):
            inv = _inv_matrix(O, wl)  # per-frame solve, like _inverseGeometricWarp
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

    def step():
        for _ in range(frames):
            inv = O.projective_from_squares(wl["dst"], wl["src"])
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
        "config": {"workload": wl["name"], "frames_per_step": frames,
                   "note": "oracle port of Homography.js _inverseGeometricWarp (C, unfused doubles); Node.js absent"},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "port",
                         "sample": f"{frames} frames x {args.steps} steps of 1728x1080, OpenMP over output rows",
                         "single_thread": one_t},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_piecewise(args, which):
    """Secondary bench lines (not the headline): BASELINE configs 3 / 4 through hg_warp_piecewise_inverse_batch."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import homography_js_b200 as hg
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:   # frames are independent: block-partitioned over ranks, NCCL only for the barrier / max / sum below
        dist.init_process_group("nccl", device_id=dev)
    ctx = hg.Context(local_rank)
    w, h = 3840, 2160
    nx = ny = 10 if which == "piecewise3" else 64
    F = args.frames if args.frames != 64 else 16      # frames per step PER GPU (weak scaling)
    src, _, tris = hg.workloads.piecewise_sinusoid(nx, ny, w, h)
    ctx.piecewise_set_mesh(src, tris)
    g = torch.Generator(device=dev)
    g.manual_seed(3 + rank)
    src_ring = torch.randint(0, 256, (F, h * w * 4), dtype=torch.uint8, device=dev, generator=g)
    dsts, frames, outs, npix = [], [], [], 0
    lo, _ = hg.workloads.shard_range(F * world, rank, world)
    for f in range(F):
        _, dst, _ = hg.workloads.piecewise_sinusoid(nx, ny, w, h, phase=2 * np.pi * (lo + f) / max(F * world, 1))
        xo, yo, oW, oH = hg.workloads.piecewise_extent(dst)
        o = torch.zeros(oW * oH * 4, dtype=torch.uint8, device=dev)
        outs.append(o)
        dsts.append(dst)
        frames.append(hg.HgFrame(src_ring[f].data_ptr(), o.data_ptr(), w, h, xo, yo, oW, oH))
        npix += oW * oH
    dst_all = np.stack(dsts)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    npix_all = reduce(npix, dist.ReduceOp.SUM)
    res = {}
    for mode in ("fused", "general"):
        ctx.debug_force_general(mode == "general")
        for _ in range(args.warmup):
            ctx.warp_piecewise_inverse_batch(dst_all, frames, 0, 0)
        ctx.synchronize()
        l0 = ctx.launch_count()
        barrier()
        ctx.profile_enable(True)
        t0 = time.perf_counter()
        ctx.timer_start()
        for _ in range(args.steps):
            ctx.warp_piecewise_inverse_batch(dst_all, frames, 0, 0)
        ms = ctx.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        kms, kn = ctx.profile_read()
        ctx.profile_enable(False)
        barrier()
        step_ms = reduce(max(ms, wall), dist.ReduceOp.MAX)       # the slowest rank sets the job's time
        res[mode] = {"Mpix/s": npix_all * args.steps / (step_ms * 1e-3) / 1e6, "ms_per_step": step_ms / args.steps,
                     "pixel_kernel_ms_per_step": kms / args.steps, "pixel_kernels": kn, "launches": ctx.launch_count() - l0}
    ctx.debug_force_general(False)
    if rank == 0:
        # parity of frame 0 against the oracle
        from oracle import oracle as O
        O.build()
        xo, yo, oW, oH = frames[0].x_off, frames[0].y_off, frames[0].o_w, frames[0].o_h
        fwd = O.piecewise_matrices(src, dsts[0], tris)
        imap = O.build_index_map(dsts[0], tris, oW, yo, oW * oH)
        want = O.warp_inverse_piecewise(src_ring[0].cpu().numpy(), w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, 0, 0,
                                        threads=os.cpu_count() or 1)
        parity = bool(np.array_equal(outs[0].cpu().numpy(), want))
        peak, _ = measured_peak_gbs()
        fr = res["fused"]
        print(json.dumps({"metric": "Mpix/s warped", "n_gpus": world, "scaling": "weak",
                          "workload": f"piecewiseaffine {nx}x{ny} grid ({len(tris)} tris), 3840x2160, {F} frames/step per GPU, sharded over {world} GPU(s)",
                          "value": fr["Mpix/s"], "unit": "Mpix/s", "parity_gate": parity, "fused": fr, "general": res["general"],
                          # rank 0's pixel kernel against ONE GPU's peak; the whole job against the peak of all of them
                          "roofline_frac_pixel_kernel": ALG_BYTES_PER_PIXEL * npix / (fr["pixel_kernel_ms_per_step"] * 1e-3) / 1e9 / peak,
                          "roofline_frac_whole_step": ALG_BYTES_PER_PIXEL * npix_all / (fr["ms_per_step"] * 1e-3) / 1e9 / (peak * world)}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def run_secondary(args, which):
    """Secondary bench lines for the remaining pixel kernels (not the headline), 1 GPU, device-resident rings > L2:
       affine_forward       _geometricWarp (H.js:911): 1080p translation by (100, 50) — the case warp() dispatches to the
                            forward loop (oW == W, oH == H) — one hg_warp_forward_matrix per frame = winner-plane
                            memset + forward_scatter_kernel + forward_gather_kernel
       projective_bilinear  config 2 through the bilinear extension (warp_inverse_geo_bilinear_kernel), one batched
                            launch per step; gate: <= 1 LSB per channel against the oracle's f64 definition"""
    import torch
    import homography_js_b200 as hg
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = hg.Context(0)
    from oracle import oracle as O
    O.build()
    F = args.frames
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    peak, peak_src = measured_peak_gbs()
    if which == "affine_forward":
        W, H, xo, yo = 1920, 1080, 100, 50
        oW, oH = W, H
        src_pts = np.array([0, 0, 0, H, W, 0], np.float64)
        fwd = ctx.solve_affine(src_pts, src_pts + np.array([xo, yo] * 3, np.float64))   # [1,0,0,1,100,50] (test.js:167-192 shape)
        src_ring = torch.randint(0, 256, (F, H * W * 4), dtype=torch.uint8, device=dev, generator=g)
        out_ring = torch.zeros((F, oW * oH * 4), dtype=torch.uint8, device=dev)

        def step():
            for f in range(F):
                ctx.image_set_device(src_ring[f].data_ptr(), W, H)
                ctx.warp_forward_matrix(fwd, xo, yo, oW, oH, to_host=False, out_dev=out_ring[f].data_ptr())

        torch.cuda.synchronize()   # the rings were filled on torch's stream; the context has its own
        step()
        ctx.synchronize()
        want = O.warp_forward_geometric(src_ring[0].cpu().numpy(), W, H, fwd, xo, yo, oW, oH)
        parity = bool(np.array_equal(out_ring[0].cpu().numpy(), want))
        name = "affine forward scatter (translation), 1920x1080 RGBA8 -> 1920x1080"
        kernel = "forward_gather_kernel (timed); whole step = memset + forward_scatter_kernel + forward_gather_kernel"
    else:
        wl = hg.workloads.projective_1080p()
        W, H, oW, oH = wl["W"], wl["H"], wl["o_w"], wl["o_h"]
        inv = ctx.solve_projective(wl["dst"], wl["src"])
        src_ring = torch.randint(0, 256, (F, H * W * 4), dtype=torch.uint8, device=dev, generator=g)
        out_ring = torch.zeros((F, oW * oH * 4), dtype=torch.uint8, device=dev)
        mats = np.tile(inv, (F, 1))
        frames = [hg.HgFrame(src_ring[f].data_ptr(), out_ring[f].data_ptr(), W, H, wl["x_off"], wl["y_off"], oW, oH) for f in range(F)]
        ctx.set_sampling(hg._abi.HG_BILINEAR)

        def step():
            ctx.warp_inverse_batch(1, mats, frames)

        torch.cuda.synchronize()
        step()
        ctx.synchronize()
        want = O.warp_inverse_geometric_bilinear(src_ring[0].cpu().numpy(), W, H, O.projective_from_squares(wl["dst"], wl["src"]),
                                                 wl["x_off"], wl["y_off"], oW, oH, threads=os.cpu_count() or 1)
        got = out_ring[0].cpu().numpy()
        parity = bool(np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1)
        name = "projective 4-point warp, BILINEAR sampling (extension), 1920x1080 RGBA8 -> 1728x1080"
        kernel = "warp_inverse_geo_bilinear2_kernel<projective>"
    if not parity:
        raise SystemExit("parity gate failed: CUDA output differs from the oracle")
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = ctx.launch_count()
    ctx.profile_enable(True)
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    sampler.sample_once()
    sampler.stop()
    kms, kn = ctx.profile_read()
    ctx.profile_enable(False)
    npix = F * oW * oH
    val = npix * args.steps / (ms * 1e-3) / 1e6
    print(json.dumps({"metric": "Mpix/s warped", "value": val, "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms / args.steps, "workload": name, "frames_per_step": F, "parity_gate": parity,
                      "l2": f"ring of {F} distinct sources + {F} distinct outputs ({(src_ring.numel() + out_ring.numel()) / 1e6:.0f} MB) > 126 MB L2",
                      "timed_kernel": kernel, "timed_kernel_ms_per_step": kms / args.steps, "timed_kernels": kn,
                      "roofline_frac_timed_kernel": ALG_BYTES_PER_PIXEL * npix / (kms / args.steps * 1e-3) / 1e9 / peak,
                      "roofline_frac_whole_step": ALG_BYTES_PER_PIXEL * npix / (ms / args.steps * 1e-3) / 1e9 / peak,
                      "peak_source": peak_src, "gpu_launches": int(ctx.launch_count() - l0), "clocks": sampler.summary()}), flush=True)
    ctx.close()


def run_video5(args):
    """Secondary line: BASELINE config 5 — video stream, 1920x1080, 30-point piecewise mesh, per-frame destiny points
    (a different output window every frame), ONE source image shared by all GPUs: rank 0 owns it and it is broadcast
    once over NCCL (NVLink) before the timed region; frames are block-partitioned over ranks, no other collective."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import homography_js_b200 as hg
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = hg.Context(local_rank)
    w, h = 1920, 1080
    n_total = args.frames * world if args.frames != 64 else 256 * world   # frames per step, whole job
    src_pts, dst_all, tris = hg.workloads.video_stream(n_total, w, h)
    lo, hi = hg.workloads.shard_range(n_total, rank, world)
    image = torch.zeros(h * w * 4, dtype=torch.uint8, device=dev)
    if rank == 0:
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        image = torch.randint(0, 256, (h * w * 4,), dtype=torch.uint8, device=dev, generator=g)
    bcast_ms = None
    if world > 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.broadcast(image, src=0)          # the one collective of the design: W*H*4 bytes, once
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
    ctx.image_set_device(image.data_ptr(), w, h)
    ctx.piecewise_set_mesh(src_pts, tris)
    smm = [int(np.floor(v + 0.5)) for v in (src_pts[:, 0].min(), src_pts[:, 1].min())]
    frames, outs, npix = [], [], 0
    for f in range(lo, hi):
        xo, yo, oW, oH = hg.workloads.piecewise_extent(dst_all[f])
        o = torch.empty(oW * oH * 4, dtype=torch.uint8, device=dev)
        outs.append(o)
        frames.append(hg.HgFrame(None, o.data_ptr(), 0, 0, xo, yo, oW, oH))
        npix += oW * oH
    arr = (hg.HgFrame * len(frames))(*frames)
    mine = np.ascontiguousarray(dst_all[lo:hi])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ctx.warp_piecewise_inverse_batch(mine, arr, smm[0], smm[1])
    f0, g0 = ctx.debug_piecewise_stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.warp_piecewise_inverse_batch(mine, arr, smm[0], smm[1])   # ends with a stream sync (status read-back)
    ms = (time.perf_counter() - t0) * 1e3
    barrier()
    f1, g1 = ctx.debug_piecewise_stats()
    tot = torch.tensor([float(npix), ms], dtype=torch.float64, device=dev)
    if world > 1:
        px_all = tot[:1].clone()
        dist.all_reduce(px_all, op=dist.ReduceOp.SUM)
        ms_max = tot[1:].clone()
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)
        npix_all, ms = float(px_all.item()), float(ms_max.item())
    else:
        npix_all = float(npix)
    parity = None
    if rank == 0:
        from oracle import oracle as O
        O.build()
        xo, yo, oW, oH = frames[0].x_off, frames[0].y_off, frames[0].o_w, frames[0].o_h
        fwd = O.piecewise_matrices(src_pts, dst_all[lo], tris)
        imap = O.build_index_map(dst_all[lo], tris, oW, yo, oW * oH)
        want = O.warp_inverse_piecewise(image.cpu().numpy(), w, h, imap, O.inverse_matrices(fwd), xo, yo, oW, oH, smm[0], smm[1],
                                        threads=os.cpu_count() or 1)
        parity = bool(np.array_equal(outs[0].cpu().numpy(), want))
        peak, _ = measured_peak_gbs()
        val = npix_all * args.steps / (ms * 1e-3) / 1e6
        print(json.dumps({"metric": "Mpix/s warped", "value": val, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                          "ms_per_step": ms / args.steps, "scaling": "weak", "parity_gate": parity,
                          "workload": f"video: 1920x1080, 30-pt piecewise ({len(tris)} tris), per-frame dstPoints, {n_total} frames/step sharded over {world} GPU(s)",
                          "frames_fused_vs_general": [int(f1 - f0), int(g1 - g0)],
                          "nccl_broadcast_ms": bcast_ms, "roofline_frac_whole_step": ALG_BYTES_PER_PIXEL * val * 1e6 / 1e9 / (peak * world)}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step (per GPU)")
    ap.add_argument("--e2e-frames", type=int, default=16, help="frames per end-to-end step (per GPU)")
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the CPU reference arm")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="projective", choices=["projective", "affine", "projective_generic", "affine_rot90", "piecewise3", "piecewise4", "video5",
                             "affine_forward", "projective_bilinear"],
                    help="projective = BASELINE config 2 (the headline); affine = same sizes through the affine kernel")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "video5":
        run_video5(args)
        return
    if args.workload.startswith("piecewise"):
        run_piecewise(args, args.workload)
        return
    if args.workload in ("affine_forward", "projective_bilinear"):
        run_secondary(args, args.workload)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    import homography_js_b200 as hg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = hg.Context(local_rank)
    wl = {"projective": hg.workloads.projective_1080p, "affine": hg.workloads.affine_1080p,
          "projective_generic": hg.workloads.projective_1080p_generic,
          "affine_rot90": hg.workloads.affine_1080p_rot90}[args.workload]()
    KIND = wl["kind"]
    W, H, oW, oH = wl["W"], wl["H"], wl["o_w"], wl["o_h"]
    F = args.frames
    npix_frame = oW * oH

    # ---- device-resident rings (torch = device-memory plumbing only)
    g = torch.Generator(device=dev)
    g.manual_seed(2 + rank)
    src_ring = torch.randint(0, 256, (F, H * W * 4), dtype=torch.uint8, device=dev, generator=g)
    out_ring = torch.zeros((F, npix_frame * 4), dtype=torch.uint8, device=dev)
    # inverse matrix: calculateTransformMatrix('projective', dst, src) on the device (K5)
    inv = ctx.solve_projective(wl["dst"], wl["src"]) if KIND == 1 else ctx.solve_affine(wl["dst"], wl["src"])
    mats = np.tile(inv, (F, 1))
    frames = [hg.HgFrame(src_ring[f].data_ptr(), out_ring[f].data_ptr(), W, H, wl["x_off"], wl["y_off"], oW, oH)
              for f in range(F)]
    torch.cuda.synchronize()

    def step():
        ctx.warp_inverse_batch(KIND, mats, frames)

    # ---- parity gate inside the run: frame 0 of the batch against the oracle (not timed)
    step()
    ctx.synchronize()
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        want = O.warp_inverse_geometric(src_ring[0].cpu().numpy(), W, H, O.calculate_transform_matrix("projective" if KIND else "affine", wl["dst"], wl["src"]),
                                        wl["x_off"], wl["y_off"], oW, oH, threads=os.cpu_count() or 1)
        parity = bool(np.array_equal(out_ring[0].cpu().numpy(), want))
        if not parity:
            raise SystemExit("parity gate failed: CUDA output differs from the oracle")

    # ---- device-resident timing
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    ctx.profile_enable(True)
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms_local = ctx.timer_stop()
    sampler.sample_once()
    barrier()
    sampler.stop()
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms_local)
    value = world * F * npix_frame * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (warp_inverse_geo_kernel<projective>), live CUDA events
    peak, peak_src = measured_peak_gbs()
    px_per_launch = F * npix_frame
    avg_kernel_s = (kern_ms / max(kern_n, 1)) * 1e-3
    achieved = ALG_BYTES_PER_PIXEL * px_per_launch / avg_kernel_s / 1e9

    # ---- end to end through the host-buffer ABI (the calls a binding makes per frame):
    #      hg_image_set (H2D from pinned host memory) + hg_warp_inverse_points (solve + warp + D2H)
    Fe = args.e2e_frames
    h_src = torch.randint(0, 256, (Fe, H * W * 4), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((Fe, npix_frame * 4), dtype=torch.uint8).pin_memory()

    def e2e_step_sync():
        for f in range(Fe):
            ctx.image_set_host_ptr(h_src[f].data_ptr(), W, H)
            ctx.warp_inverse_points(KIND, wl["dst"], wl["src"], wl["x_off"], wl["y_off"], oW, oH,
                                    out_host_ptr=h_out[f].data_ptr())

    pipe = hg.Pipe(ctx, KIND, W, H, oW, oH, depth=4)

    def e2e_step():
        # every frame: H2D of its image from pinned host memory, solve + warp, D2H of its result; the step ends when
        # the last result byte is in host memory
        for f in range(Fe):
            pipe.submit(h_src[f].data_ptr(), wl["dst"], wl["src"], wl["x_off"], wl["y_off"], oW, oH, h_out[f].data_ptr())
        pipe.flush()

    def time_e2e(fn, steps):
        for _ in range(3):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ctx.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        return max_over_ranks(wall_ms)

    e2e_steps = args.steps
    e2e_ms = time_e2e(e2e_step, e2e_steps)
    e2e_val = world * Fe * npix_frame * e2e_steps / (e2e_ms * 1e-3) / 1e6
    sync_steps = max(3, min(args.steps, 20))
    e2e_sync_ms = time_e2e(e2e_step_sync, sync_steps)
    e2e_sync_val = world * Fe * npix_frame * sync_steps / (e2e_sync_ms * 1e-3) / 1e6
    if rank == 0 and not args.no_cpu_baseline:
        # the last pipelined frame that landed in host memory equals the oracle's answer
        from oracle import oracle as O
        want = O.warp_inverse_geometric(h_src[Fe - 1].numpy(), W, H, _inv_matrix(O, wl), wl["x_off"], wl["y_off"], oW, oH,
                                        threads=os.cpu_count() or 1)
        e2e_step()
        if not np.array_equal(h_out[Fe - 1].numpy(), want):
            raise SystemExit("parity gate failed: pipelined host-to-host output differs from the oracle")
    pipe.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v_all, n_all, dt_all = cpu_port_rate(wl, threads, args.cpu_budget * 0.8)
        v_one, n_one, dt_one = cpu_port_rate(wl, 1, args.cpu_budget * 0.2, 1)
        cpu = {"value": v_all, "unit": "Mpix/s", "cores": threads, "kind": "port",
               "sample": f"{n_all} frames of 1728x1080 in {dt_all:.1f} s, OpenMP over output rows "
                         "(oracle port of Homography.js _inverseGeometricWarp; Node.js absent from the image)",
               "single_thread": v_one}

    if rank == 0:
        line = {
            "metric": "Mpix/s warped", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "frames_per_step_per_gpu": F,
                       "l2": f"ring of {F} distinct sources + {F} distinct outputs per GPU "
                             f"({(src_ring.numel() + out_ring.numel()) / 1e6:.0f} MB) > 126 MB L2",
                       "parity_gate": parity},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic("warp_inverse_geo_kernel<%s>" % ("projective" if KIND else "affine")) if F == 64 else None,
                         "kernel": "warp_inverse_geo_kernel<%s>" % ("projective" if KIND else "affine"),
                         "alg_bytes_per_launch": ALG_BYTES_PER_PIXEL * px_per_launch,
                         "avg_kernel_ms": avg_kernel_s * 1e3, "kernels_timed": kern_n, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": Fe * H * W * 4,
                    "d2h_bytes_per_step": Fe * npix_frame * 4, "frames_per_step_per_gpu": Fe, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps, "timer": "host wall clock around submit..flush, max over ranks",
                    "api": "hg_pipe_submit per frame (H2D image + solve + warp + D2H result, 4 frames in flight) + hg_pipe_flush, pinned host buffers",
                    "sync_single_frame": {"value": e2e_sync_val, "unit": "Mpix/s",
                                          "api": "hg_image_set + hg_warp_inverse_points (blocking, what Homography.warp() does)"}},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
