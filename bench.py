#!/usr/bin/env python
"""bench.py -- novel views/sec (256x256) of the pixelsynth_b200 hot path, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Contract (see DESIGN.md "Measurement"): one JSON line on stdout from rank 0.
  step      = one pass of the hot path over one batch of B synthetic 256x256 source images (RGB U[-1,1],
              depth U[min_z,max_z], circle-translation target cameras of z_buffermodel.py:214), each -> one view.
  value     = views/s with inputs resident in HBM, CUDA events on the launching stream, max over ranks.
  e2e       = the same through the reference-facing call (PtsManipulator.forward_justpts) with HOST pinned
              buffers: H2D of depth+features+cameras and D2H of the image+mask inside the timed region.
  roofline  = the dominant kernel (fine_kernel) timed alone with CUDA events; algorithmic bytes per launch
              = 69 009 408 B/view x B over the measured HBM peak (MEASURED_PEAKS.json).
  cpu_baseline = the CPU oracle (port of the reference algorithm) on a bounded sample, rank 0, N=1 only.
--impl reference times the CPU oracle with all host threads on the same workload (rank 0 only).
Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W = 256
K_PP = 128
RADIUS = 4.0
C = 3
BYTES_PER_VIEW_MAPS = 4 * W * W + 4 * C * W * W + 4 * C * W * W + W * W + 2 * 4 * K_PP * W * W  # 69 009 408
BYTES_PER_VIEW_FUSED = 4 * W * W + 4 * C * W * W + 4 * C * W * W + W * W                        # 1 900 544
METRIC = "novel views/sec (256x256)"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="views per step per GPU")
    ap.add_argument("--cpu-sample", type=int, default=16, help="views in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "z-buffer splat stage (SURVEY 8a S1-S5: unproject -> camera transform -> K-nearest rasterise -> "
                    "alpha composite -> background mask) of BASELINE configs[1], 256x256, P=65536, K=128, radius 4 px; "
                    "value/roofline with idx+zbuf maps emitted (69.0 MB/view, the bit-exact parity surface), e2e "
                    "through PtsManipulator.forward_justpts (image+mask only). Depth Unet / VQ-VAE / lmconv / "
                    "refinement stages are not in this number yet.",
        "views_per_step_per_gpu": args.batch,
        "global_views_per_step": args.batch * world,
        "image": "256x256", "points_per_pixel": K_PP, "radius_px": RADIUS,
        "target_cameras": "translation circle n=rank%8 (create_nerf_like_circles.py:14)",
        "l2_policy": "outputs per step (%.0f MB) exceed the 126 MB L2; no explicit flush" %
                     (BYTES_PER_VIEW_MAPS * args.batch / 1e6),
        "parallelism": "views sharded across ranks, one NCCL broadcast of the sources at job start" if world > 1
                       else "single GPU",
    }


def make_inputs(B, view, seed=0):
    import numpy as np
    from util import demo_cameras, pack_mats

    rng = np.random.default_rng(seed)
    depth = rng.uniform(0.5, 10.0, (B, 1, W, W)).astype(np.float32)
    feat = rng.uniform(-1, 1, (B, C, W, W)).astype(np.float32)
    cams = demo_cameras(B, "translate", seed, views=[view] * B)
    return depth, feat, cams, pack_mats(*cams)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_oracle_views_per_s(n_views, threads):
    """Times the CPU oracle (oracle/splat_oracle.c, a port of the reference algorithm) on n_views views."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import splat_ref

    splat_ref.build()
    depth, feat, cams, mats = make_inputs(n_views, 0)
    splat_ref.splat(depth[:1], feat[:1], mats[:1], W, K=K_PP, radius_px=RADIUS)  # warm-up

    def one(i):
        return splat_ref.splat(depth[i:i + 1], feat[i:i + 1], mats[i:i + 1], W, K=K_PP, radius_px=RADIUS)["out"].sum()

    t0 = time.perf_counter()
    if threads == 1:
        for i in range(n_views):
            one(i)
    else:
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the C oracle
            list(ex.map(one, range(n_views)))
    dt = time.perf_counter() - t0
    return n_views / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(threads, 8)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_oracle_views_per_s(per_step, threads)
        if i >= args.warmup:
            vals.append((v, dt))
        if sum(d for _, d in vals) > 150:
            break
    tot_t = sum(d for _, d in vals)
    value = per_step * len(vals) / tot_t
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / len(vals), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d views per step x %d steps of the same 256x256 K=128 splat workload, CPU oracle "
                                   "(oracle/splat_oracle.c; PyTorch3D, the reference's rasteriser, is not installable "
                                   "offline), %d host threads" % (per_step, len(vals), threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    devs = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=devs)

    import pixelsynth_b200.ops as ops  # registers torch.ops.pixelsynth_b200 (fails loudly if the .so is missing)
    from pixelsynth_b200 import _lib
    from pixelsynth_b200.models.projection.z_buffer_manipulator import PtsManipulator
    import types

    L = _lib.lib()
    B = args.batch
    view = rank % 8
    depth, feat, cams, mats = make_inputs(B, view)

    # ---- job start: sources live on rank 0 and are broadcast once over NCCL ----
    d_depth = torch.empty((B, 1, W, W), device=devs)
    d_feat = torch.empty((B, C, W, W), device=devs)
    bcast_ms = 0.0
    if rank == 0:
        d_depth.copy_(torch.from_numpy(depth))
        d_feat.copy_(torch.from_numpy(feat))
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dist.broadcast(d_depth, 0)
        dist.broadcast(d_feat, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    d_mats = torch.from_numpy(mats).to(devs)

    def step_maps():
        return torch.ops.pixelsynth_b200.splat(d_depth, d_feat, d_mats, W, W, K_PP, RADIUS, 1.0, 2, 0, 13, 1e-2, True,
                                               False)

    def step_fused():
        return torch.ops.pixelsynth_b200.splat(d_depth, d_feat, d_mats, W, W, K_PP, RADIUS, 1.0, 2, 0, 13, 1e-2, False,
                                               False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.ps_launch_count_reset()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = L.ps_launch_count()
        if world > 1:
            t = torch.tensor([ms], device=devs)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_maps, launches = timed(step_maps, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_fused, _ = timed(step_fused, args.steps, max(3, args.warmup))

    # ---- e2e: reference-facing call with host buffers ----
    opt = types.SimpleNamespace(splatter="xyblending", learn_default_feature=True, radius=RADIUS, pp_pixel=K_PP,
                                rad_pow=2, tau=1.0, accumulation="alphacomposite", background_smoothing_kernel_size=13)
    pm = PtsManipulator(W, C=C, opt=opt).to(devs)
    h_depth = torch.from_numpy(depth).pin_memory()
    h_feat = torch.from_numpy(feat).pin_memory()
    h_cams = [torch.from_numpy(np.ascontiguousarray(m)).pin_memory() for m in cams]
    h_out = torch.empty((B, C, W, W)).pin_memory()
    h_bg = torch.empty((B, W, W), dtype=torch.bool).pin_memory()

    def step_e2e():
        dd = h_depth.to(devs, non_blocking=True)
        ff = h_feat.to(devs, non_blocking=True)
        cc = [m.to(devs, non_blocking=True) for m in h_cams]
        gen_fs, bg = pm.forward_justpts(ff, dd, *cc)
        h_out.copy_(gen_fs, non_blocking=True)
        h_bg.copy_(bg, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the result every step

    ms_e2e, _ = timed(step_e2e, args.steps, max(3, args.warmup))
    h2d = h_depth.numel() * 4 + h_feat.numel() * 4 + sum(m.numel() * 4 for m in h_cams)
    d2h = h_out.numel() * 4 + h_bg.numel()

    # ---- roofline: the dominant kernel alone (ps_splat_points on pre-projected points), CUDA events ----
    pts, _ = torch.ops.pixelsynth_b200.project_pts(d_depth, d_mats, W, 1e-2, False)
    f3 = d_feat.reshape(B, C, -1)

    def step_points():
        return torch.ops.pixelsynth_b200.splat_points(pts, f3, W, K_PP, RADIUS, 1.0, 2, 0, 13, True, False)

    for _ in range(3):
        step_maps()
    torch.cuda.synchronize()
    _lib.kernel_time_ms(None)
    L.ps_timing_enable(1)  # CUDA events around fine_kernel on the launching stream, same step as `value`
    for _ in range(args.steps):
        step_maps()
    torch.cuda.synchronize()
    L.ps_timing_enable(0)
    kt_total, kt_n = _lib.kernel_time_ms("fine_kernel")
    _lib.kernel_time_ms(None)
    ms_points, _ = timed(step_points, args.steps, 3)
    peak, peak_src = measured_peak()
    kt = kt_total / max(kt_n, 1)
    kernel_name = "fine_kernel"
    achieved = BYTES_PER_VIEW_MAPS * B / (kt * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    views_per_step = B * world
    value = views_per_step * args.steps / (ms_maps * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_maps / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": views_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "PtsManipulator.forward_justpts (maps suppressed)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": kernel_name, "ms_per_launch": kt, "share_of_step": kt / (ms_maps / args.steps),
                     "rasterise_call_ms": ms_points / args.steps,
                     "algorithmic_bytes_per_launch": BYTES_PER_VIEW_MAPS * B, "peak_source": peak_src},
        "splat_hbm_gbs_whole_step": BYTES_PER_VIEW_MAPS * views_per_step * args.steps / (ms_maps * 1e-3) / 1e9 / world,
        "splat_fused_views_per_s": views_per_step * args.steps / (ms_fused * 1e-3),
        "broadcast_ms": bcast_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_oracle_views_per_s(args.cpu_sample, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d views of the same workload (256x256, K=128) through oracle/splat_oracle.c, "
                                          "single thread, %.1f s" % (args.cpu_sample, dt)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
