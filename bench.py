#!/usr/bin/env python
"""bench.py -- novel views/sec (256x256) of the pixelsynth_b200 hot path, with rooflines and a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--view V]
                    [--in-flight D] [--sampler-sms S] [--workload views|scene]

One JSON line on stdout from rank 0 (contract in DESIGN.md section 7).
  workload  = BASELINE.json configs[1]: the demo path `ZbufferModelPts.forward` (models/z_buffermodel.py:291-419) --
              depth U-Net -> z-buffer splat -> order/masks -> VQ-VAE-2 encode -> lmconv autoregressive outpaint ->
              VQ-VAE-2 decode -> refinement decoder -- on synthetic 256x256 images, one novel view per image, batched
              B images per step (the path is embarrassingly parallel over images).  Seeded random weights of the
              reference's architecture (no checkpoint is reachable offline).
  value     = views/s with the inputs resident in HBM; the K steps go through pixelsynth_b200.pipeline.ViewPipeline
              with `--in-flight` batches in flight (the sampler of step k+1 on its own SM partition beside the decoder
              of step k) and all complete inside the timed region; CUDA events; max over ranks.
              `one_step_at_a_time` = the same model called serially on the whole device (the latency of a step).
  e2e       = views/s through ViewPipeline.submit(ZbufferModelPts.forward) + BaseModel's rescale
              (models/base_model.py:93-103) with HOST pinned inputs: H2D of images + cameras and D2H of PredImg every
              step, the host waiting for each step's pixels.
  roofline  = the dominant kernel of the step (by summed device time, measured with CUDA events the library records
              on the launching stream during the one-step-at-a-time pass of the same run); `rooflines` lists splat
              (HBM; also with maps emitted on the U-Net's and on SURVEY 8d's uniform depth), sampler (step and
              BASELINE config 3) and convolutions (tensor).
  --workload scene = BASELINE configs[4]: batched gen_scene sweeps (models/z_buffermodel.py:421-592).
  cpu_baseline / --impl reference = the CPU oracle of the same path (fp32 torch restatement of the reference's
              modules + the C splat oracle); the sampler is timed reference-style (one full forward per token,
              models/lmconv/sample.py:54-66) on a few tokens and extrapolated linearly -- stated in `sample`.
Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W = 256
K_PP = 128
RADIUS = 4.0
BYTES_PER_VIEW_SPLAT_FUSED = 4 * W * W + 2 * 4 * 3 * W * W + W * W  # depth + feat in, image + mask out = 1 900 544
BYTES_PER_VIEW_SPLAT_MAPS = BYTES_PER_VIEW_SPLAT_FUSED + 2 * 4 * K_PP * W * W  # + idx and z maps = 69 009 408
FLOP_PER_CELL = 11.163e6       # lmconv column, SURVEY.md 8d
FLOP_DECODER = 128.22e9        # ResNetDecoder per image
FLOP_UNET, FLOP_VQ_ENC, FLOP_VQ_DEC = 5.53e9, 3.66e9 + 0.07e9, 2.58e9
METRIC = "novel views/sec (256x256)"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64,
                    help="images (= views) per step per GPU (BASELINE configs[3]: batch 64).  The sampler's serial levels "
                         "cost the same at any batch, so throughput grows with it")
    ap.add_argument("--view", type=int, default=-1,
                    help="-1 (default): slot i of rank r renders circle view (i + r) mod 8 of image (i + r) mod B, so every rank carries the same "
                         "mix of all 8 views; 0..7: every image renders that one view")
    ap.add_argument("--workload", default="views", choices=["views", "scene"],
                    help="views (default, the headline): one novel view per image, BASELINE configs[1]/[3]; scene: BASELINE "
                         "configs[4], every image sweeps a scene (gen_scene, refinement decoder included) -- a step is the whole "
                         "sweep of --batch images per GPU (default 32 = 256 images over 8 GPUs)")
    ap.add_argument("--directions", nargs="+", default=["R", "L"], help="scene sweep directions (scripts/demo_scene.sh uses 10)")
    ap.add_argument("--num_split", type=int, default=2, help="scene sweep splits per direction (scripts/demo_scene.sh: 32)")
    ap.add_argument("--cpu-tokens", type=int, default=16, help="sampler tokens timed per step for the CPU baseline")
    ap.add_argument("--view-offset", type=int, default=0, help="developer aid: run rank r's view mix on one GPU (offset r)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=3,
                    help="batches in flight per GPU (pixelsynth_b200.pipeline.ViewPipeline); 1 = one forward at a time")
    ap.add_argument("--sampler-sms", type=int, default=24, help="SMs of the sampler's green-context partition (in-flight > 1)")
    return ap.parse_args()


def make_opt(**kw):
    o = dict(W=256, splatter="xyblending", learn_default_feature=True, radius=RADIUS, pp_pixel=K_PP, rad_pow=2, tau=1.0,
             accumulation="alphacomposite", background_smoothing_kernel_size=13, min_z=0.5, max_z=10.0,
             use_rgb_features=True, use_gt_depth=False, use_inverse_depth=False, depth_predictor_type="unet",
             no_outpainting=False, vqvae=True, num_samples=1, temperature=0.7, model_setting="gen_paired_img",
             direction="L", rotation=0.6, homography=False, seed=0, normalize_image=True, predict_residual=True,
             normalize_before_residual=False, refine_model_type="resnet_256W8UpDown3", ngf=64, norm_G="sync:spectral_batch")
    o.update(kw)
    return types.SimpleNamespace(**o)


def workload_config(args, world):
    mix = ("slot i of rank r = image (i + r) mod B rendering circle view (i + r) mod 8: every rank carries the same (image, "
           "view) pairs, rotated"
           if args.view < 0 else "every image renders circle view %d" % args.view)
    return {
        "workload": "BASELINE configs[3] per GPU = configs[1] batched (demo path: depth Unet -> splat -> order/masks -> "
                    "VQ-VAE-2 encode -> lmconv outpaint -> VQ-VAE-2 decode -> refinement decoder), one novel 256x256 view "
                    "per synthetic image on the 8-view translation circle (create_nerf_like_circles.py:14), num_samples 1, "
                    "temperature 0.7, %d images per step per GPU" % args.batch,
        "views_per_step_per_gpu": args.batch, "global_views_per_step": args.batch * world,
        "view_mix": mix,
        "weights": "seeded random init of the reference architecture (pixelsynth_b200/synthetic.py)",
        "l2_policy": "activations + sampler cache per step (> 1 GB at batch 32) exceed the 126 MB L2; no explicit flush",
        "parallelism": "images sharded across ranks; one NCCL broadcast of the source images at job start" if world > 1
                       else "single GPU",
        "in_flight": max(1, args.in_flight),
        "schedule": ("%d steps in flight per GPU (pixelsynth_b200.pipeline.ViewPipeline): step k+1's sampler launch runs on a "
                     "%d-SM green-context partition beside step k's refinement decoder on the other SMs; all K steps complete "
                     "inside the timed region" % (args.in_flight, args.sampler_sms)) if args.in_flight > 1
                    else "one step at a time on the whole device",
    }


def batch_views(args, rank):
    return [(i + rank + args.view_offset) % 8 for i in range(args.batch)] if args.view < 0 else [args.view] * args.batch


def make_batch(B, view, seed=0):
    import torch
    from util import demo_cameras
    from pixelsynth_b200 import synthetic

    views = list(view) if isinstance(view, (list, tuple)) else [view] * B   # per-image circle view, or one for all
    K, Kinv, RT1, RT1inv, RT2, RT2inv = [torch.from_numpy(m) for m in demo_cameras(B, "translate", seed, views=views)]
    img = synthetic.synth_image(B, seed)
    return {"images": [img, img.clone()],
            "cameras": [{"K": K, "Kinv": Kinv, "P": RT1, "Pinv": RT1inv}, {"K": K, "Kinv": Kinv, "P": RT2, "Pinv": RT2inv}]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
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


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json: copy GB/s, sustained bf16 TF/s)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle of the same path (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_view_seconds(n_tokens, threads, seed=0, view=0):
    """One view through the CPU oracle; the sampler runs `n_tokens` reference-style steps (one full network forward per
    token, models/lmconv/sample.py:54-66) and is extrapolated to the view's sampled-cell count.
    Returns (extrapolated seconds per view, dict of stage seconds, sampled cells, measured wall seconds)."""
    import numpy as np
    import torch
    from oracle import lmconv_ref, nets_ref, splat_ref
    from pixelsynth_b200 import synthetic
    from util import demo_cameras, pack_mats

    torch.set_num_threads(threads)
    sds = {n: synthetic.make_state(n, 0) for n in ("unet", "vqvae", "lmconv", "decoder")}
    img = synthetic.synth_image(1, seed)
    cams = demo_cameras(1, "translate", seed, views=[view])
    st = {}
    t_wall = time.perf_counter()
    with torch.no_grad():
        t = time.perf_counter()
        depth = nets_ref.unet_depth(sds["unet"], img, 0.5, 10.0)
        st["unet"] = time.perf_counter() - t
        t = time.perf_counter()
        sp = splat_ref.splat(depth.numpy(), img.numpy(), pack_mats(*cams), W, K=K_PP, radius_px=RADIUS)
        st["splat"] = time.perf_counter() - t
        gen_fs, bg = torch.from_numpy(sp["out"]), torch.from_numpy(sp["bg"])
        t = time.perf_counter()
        _, orders, words, smask = lmconv_ref.glue_from_background(bg)
        st["glue"] = time.perf_counter() - t
        t = time.perf_counter()
        ids, _ = nets_ref.vqvae_encode_top(sds["vqvae"], gen_fs)
        st["vq_encode"] = time.perf_counter() - t
        n_sampled = int(smask.sum())
        t = time.perf_counter()
        g = torch.Generator().manual_seed(1)
        steps = max(1, min(n_tokens, n_sampled))
        lmconv_ref.sample_reference_style(sds["lmconv"], ids, orders, words, smask.numpy(), torch.rand(1, 1024, generator=g), 0.7,
                                          max_steps=steps)
        per_token = (time.perf_counter() - t) / steps
        st["lmconv_per_token"] = per_token
        t = time.perf_counter()
        ar = nets_ref.vqvae_decode_code(sds["vqvae"], ids)
        st["vq_decode"] = time.perf_counter() - t
        t = time.perf_counter()
        comb = gen_fs * (~bg)[:, None].float() + ar * bg[:, None].float()
        nets_ref.decoder_forward(sds["decoder"], comb, bg, [torch.randn(1, 20, generator=g) for _ in range(16)])
        st["decoder"] = time.perf_counter() - t
    wall = time.perf_counter() - t_wall
    st["lmconv_tokens_timed"] = steps
    total = sum(v for k, v in st.items() if k not in ("lmconv_per_token", "lmconv_tokens_timed")) + per_token * n_sampled
    return total, st, n_sampled, wall


def run_reference(args):
    """--impl reference: the CPU oracle of the same path on all host cores, rank 0 only.  One step = ONE view (a bounded
    sample of the arm's 64-views-per-step workload): every stage runs in full except the sampler, which is timed the
    reference's way (a full network forward per token) on --cpu-tokens tokens and scaled to the view's sampled cells.
    `ms_per_step` is the wall time the step really took; `value` is the extrapolated views/s (labelled)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    secs, walls = [], []
    info = None
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        view = i % 8 if args.view < 0 else args.view        # the arm's view mix, one view per step
        s, st, n_sampled, wall = cpu_oracle_view_seconds(args.cpu_tokens, threads, view=view)
        if i >= args.warmup:
            secs.append(s)
            walls.append(wall)
            info = (st, n_sampled)
        if time.perf_counter() - t_start > 240 and secs:
            break
    value = len(secs) / sum(secs)
    st, n_sampled = info
    sample = ("1 view per step (views cycle over the 8-view circle) through the CPU oracle = torch fp32 restatement of the "
              "reference modules + C splat oracle (the reference itself needs PyTorch3D CUDA ops and hard-coded .cuda()), "
              "%d host threads; every stage measured in full except the sampler: %d tokens per step timed reference-style "
              "(one full forward per token, %.3f s/token) and scaled to the view's sampled cells (last view: %d). "
              "value = 1 / extrapolated seconds per view; ms_per_step = measured wall per step (%.0f%% of a view's "
              "extrapolated cost was actually executed)"
              % (threads, args.cpu_tokens, st["lmconv_per_token"], n_sampled, 100.0 * sum(walls) / sum(secs)))
    cfg = workload_config(args, world)
    cfg["reference_sample"] = "1 view per step, sampler extrapolated from %d tokens" % args.cpu_tokens
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "value_kind": "extrapolated from a bounded sample (see cpu_baseline.sample)",
        "unit": UNIT, "n_gpus": args.gpus, "steps": len(secs),
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls), "extrapolated_ms_per_view": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "stage_seconds": st},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_scene(args):
    """BASELINE configs[4]: batched scene sweeps sharded by image over the GPUs (no data-path collective).  One step =
    every image of the rank's batch rendered along all `--directions` x `--num_split` poses over its growing point
    cloud; value = rendered views/s of the whole job (inputs resident in HBM), e2e adds the H2D of the images / cameras
    and the D2H of every rendered view."""
    import torch
    import torch.distributed as dist
    from pixelsynth_b200 import _lib
    from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts

    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=3))
        os.environ.setdefault("PS_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    B = args.batch if any(a.startswith("--batch") for a in sys.argv[1:]) else 32
    opt = make_opt(model_setting="gen_scene", directions=list(args.directions), num_split=args.num_split,
                   sequential_outpainting=False)
    model = ZbufferModelPts(opt, device=dev)
    from pixelsynth_b200 import synthetic
    from util import demo_cameras
    K, Kinv, RT1, RT1inv, _, _ = [torch.from_numpy(m) for m in demo_cameras(B, "identity", 0)]
    img = synthetic.synth_image(B, 100 + rank)                      # every rank owns different images
    host = {"images": [img], "cameras": [{"K": K, "Kinv": Kinv, "P": RT1, "Pinv": RT1inv}]}
    resident = {"images": [img.to(dev)], "cameras": [{k: v.to(dev) for k, v in host["cameras"][0].items()}]}
    pinned = {"images": [img.pin_memory()], "cameras": [{k: v.pin_memory() for k, v in host["cameras"][0].items()}]}
    g = torch.Generator().manual_seed(1 + rank)
    noise, uniforms = torch.randn(16, B, 20, generator=g).to(dev), torch.rand(B, 1024, generator=g)
    views = sum(model._splits(d) + 1 for d in args.directions)
    h_out = torch.empty((views, B, 3, W, W)).pin_memory()

    def step(batch, fetch):
        _, out = model.forward(batch, noise=noise, uniforms=uniforms)
        if fetch:
            for i, k in enumerate(k for k in out if k.startswith("PredImg_")):
                h_out[i].copy_(out[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.lib().ps_launch_count_reset()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.check_wedge("bench.py")
        ms, launches = e0.elapsed_time(e1), _lib.lib().ps_launch_count()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, launches = timed(lambda: step(resident, False))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(lambda: step(pinned, True))
    cloud = int(model.last_scene[-1]["cloud"].shape[2])
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    total = views * B * world * args.steps
    print(json.dumps({
        "metric": METRIC, "value": total / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[4]: scene sweep (gen_scene, z_buffermodel.py:421-592) with the refinement decoder, "
                               "%d images per GPU in lock step, directions %s, num_split %d = %d dependent views per image; "
                               "sharded by image, no collective" % (B, " ".join(args.directions), args.num_split, views),
                   "views_per_image": views, "images_per_gpu": B, "final_cloud_points": cloud,
                   "weights": "seeded random init of the reference architecture"},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(img.numel() * 4 + 4 * 64 * B),
                "d2h_bytes_per_step": int(h_out.numel() * 4)},
        "gpu_launches": int(launches), "clocks": clocks,
    }), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "scene":
        return run_scene(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a rank that dies must not leave the others in a collective for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=3))
        # one process per GPU share the host: the native order / mask / level code gets cores / world worker threads
        os.environ.setdefault("PS_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // world)))

    from pixelsynth_b200 import _lib
    from pixelsynth_b200.models.base_model import BaseModel
    from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts
    from pixelsynth_b200.parallel import broadcast_sources

    L = _lib.lib()
    B = args.batch
    opt = make_opt()
    model = ZbufferModelPts(opt, device=dev)
    host_batch = make_batch(B, batch_views(args, rank))

    # ---- job start: the source images live on rank 0 and are broadcast once over NCCL.  The communicator
    # is created by a warm-up collective first, so `broadcast_ms` is the transfer, not NCCL's initialisation ----
    src = host_batch["images"][0].to(dev) if rank == 0 else torch.empty((B, 3, W, W), device=dev)
    if world > 1:
        warm = torch.zeros(1, device=dev)
        dist.broadcast(warm, 0)
        torch.cuda.synchronize()
    bcast_ms = broadcast_sources(src, world)
    # rank r holds the SAME (image, view) pairs as rank 0, rotated by r slots (slot i = image (i + r) mod B with view
    # (i + r) mod 8): every GPU does identical work, so the 1 -> N curve measures the system and not which rank drew the
    # images with the longest sampling chains (the same pairs in rank order gave 23.9 .. 25.3 ms/step on one GPU)
    shift = (rank + args.view_offset) % B
    if shift:
        src = torch.roll(src, -shift, 0)
        host_batch["images"] = [torch.roll(t, -shift, 0) for t in host_batch["images"]]
    dev_batch = {"images": [src, src],
                 "cameras": [{k: v.to(dev) for k, v in c.items()} for c in host_batch["cameras"]]}
    g = torch.Generator().manual_seed(1 + rank)
    noise = torch.randn(16, B, 20, generator=g).to(dev)
    uniforms = torch.rand(B, 1024, generator=g)

    # `--in-flight` batches per GPU: batch k+1's sampler (a latency-bound chain on a few SMs) runs on its own SM partition
    # beside batch k's refinement decoder (pixelsynth_b200/pipeline.py).  Every step's work is inside the timed region:
    # the loop is drained before the closing event.
    from pixelsynth_b200.pipeline import ViewPipeline
    pipe = ViewPipeline(model, depth=max(1, args.in_flight), sampler_sms=args.sampler_sms, partition="auto")
    pending = []

    def step_resident():
        pending.append(pipe.submit(dev_batch, noise=noise, uniforms=uniforms))
        if len(pending) >= pipe.depth:
            pipe.result(pending.pop(0), host_wait=False)

    def drain():
        while pending:
            pipe.result(pending.pop(0), host_wait=False)

    def step_serial():
        return model.forward(dev_batch, noise=noise, uniforms=uniforms)[1]["PredImg"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain=lambda: None):
        for _ in range(warmup):
            fn()
        drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.ps_launch_count_reset()
        e0.record()
        for _ in range(steps):
            fn()
        drain()                                # the current stream now waits for every batch in flight
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        _lib.check_wedge("bench.py")          # a wedged kernel's numbers are not numbers
        launches = L.ps_launch_count()
        per_rank = [ms]
        if world > 1:
            t = torch.zeros(world, device=dev)
            t[rank] = ms
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            per_rank = [float(x) for x in t.tolist()]
            ms = max(per_rank)
        return ms, launches, per_rank

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, launches, ms_res_ranks = timed(step_resident, args.steps, args.warmup, drain)
    clocks = sampler.stop() if rank == 0 else None
    ms_serial, _, _ = timed(step_serial, args.steps, 1) if pipe.depth > 1 else (ms_res, 0, 0)   # one forward at a time: the step's latency

    # ---- e2e through the reference-facing wrapper with host buffers ----
    bm = BaseModel(model, opt)
    pin = lambda t: t.pin_memory()
    pinned = {"images": [pin(t) for t in host_batch["images"]],
              "cameras": [{k: pin(v) for k, v in c.items()} for c in host_batch["cameras"]]}
    h_outs = [torch.empty((B, 3, W, W)).pin_memory() for _ in range(pipe.depth)]
    h_out = h_outs[0]
    model_kw = dict(noise=noise, uniforms=uniforms)
    e2e_n = [0]

    def step_e2e():
        # every step: H2D of its images + cameras from pinned memory (process_batch, on the batch's stream), forward,
        # BaseModel's rescale (base_model.py:96-99), D2H of the result into pinned memory; the host then waits for the
        # OLDEST batch in flight and owns its pixels
        dst = h_outs[e2e_n[0] % pipe.depth]
        e2e_n[0] += 1
        pending.append(pipe.submit(pinned, then=lambda loss, out: dst.copy_(0.5 * out["PredImg"] + 0.5, non_blocking=True),
                                   **model_kw))
        if len(pending) >= pipe.depth:
            pipe.result(pending.pop(0), host_wait=True)

    def drain_host():
        while pending:
            pipe.result(pending.pop(0), host_wait=True)

    ms_e2e, _, _ = timed(step_e2e, args.steps, max(3, args.warmup), drain_host)
    h2d = sum(t.numel() * 4 for t in pinned["images"]) + sum(v.numel() * 4 for c in pinned["cameras"] for v in c.values())
    d2h = h_out.numel() * 4

    # ---- per-kernel device time over the same step (CUDA events recorded by the library on the launching stream) ----
    # One step at a time on the whole device: with steps in flight an event-bracketed duration would include the time a
    # kernel's CTAs queue behind the other step's kernels, so the kernels are timed where their durations are their own.
    step_serial()
    torch.cuda.synchronize()
    _lib.kernel_time_ms(None)
    L.ps_timing_enable(1)
    for _ in range(args.steps):
        step_serial()
    torch.cuda.synchronize()
    L.ps_timing_enable(0)
    kt = {n: _lib.kernel_time_ms(n) for n in ("fine_kernel", "fine_big_kernel", "conv_igemm_kernel", "lmconv_tc_kernel")}
    _lib.kernel_time_ms(None)
    last = pipe.last(0)
    # the splat in its map-emitting mode (idx + z maps: the bit-exact parity surface, 69.0 MB/view): the HBM-bound
    # configuration SURVEY 8d defines the splat roofline on.  Same depth / cameras as the step, 16 views per launch.
    from pixelsynth_b200.ops import pack_mats as pack_mats_t
    nb = min(B, 16)
    cam0, cam1 = dev_batch["cameras"][0], dev_batch["cameras"][1]
    mats = pack_mats_t(cam0["K"][:nb], cam0["Kinv"][:nb], cam0["P"][:nb], cam0["Pinv"][:nb], cam1["P"][:nb], cam1["Pinv"][:nb])
    depth16, feat16 = last["depth"][:nb].contiguous(), src[:nb].contiguous()
    run_maps = lambda: torch.ops.pixelsynth_b200.splat(depth16, feat16, mats, W, W, K_PP, RADIUS, 1.0, 2, 0, 13, 1e-2, True, False)
    for _ in range(3):
        run_maps()
    torch.cuda.synchronize()
    L.ps_timing_enable(1)
    for _ in range(5):
        run_maps()
    torch.cuda.synchronize()
    L.ps_timing_enable(0)
    maps_ms = _lib.kernel_time_ms("fine_kernel")[0] / 5
    _lib.kernel_time_ms(None)
    # the same launch on SURVEY 8d's own distribution: depth ~ U[min_z, max_z] (seed 0) instead of the U-Net's output
    depth_u = (torch.rand(nb, 1, W, W, generator=torch.Generator().manual_seed(0)) * 9.5 + 0.5).to(dev)
    run_maps_u = lambda: torch.ops.pixelsynth_b200.splat(depth_u, feat16, mats, W, W, K_PP, RADIUS, 1.0, 2, 0, 13, 1e-2, True, False)
    for _ in range(3):
        run_maps_u()
    torch.cuda.synchronize()
    L.ps_timing_enable(1)
    for _ in range(5):
        run_maps_u()
    torch.cuda.synchronize()
    L.ps_timing_enable(0)
    maps_u_ms = _lib.kernel_time_ms("fine_kernel")[0] / 5
    _lib.kernel_time_ms(None)
    cells_processed = int(model.outpaint2.last_levels[-1])   # rows the sampler pushed through the network this step
    levels = len(model.outpaint2.last_levels) - 1
    cells_sampled = int(last["sample_mask"].sum())
    levels_prefix = int(model.outpaint2.last_first_b)

    # ---- BASELINE configs[2]: batch 32, right half of the 32x32 code grid masked (512 cells / image), T = 0.7,
    # uniforms seed 1.  tokens/s = 32 * 512 / device time of the sampler launch (prefix included; order /
    # mask / level construction on the host excluded and reported) ----
    import numpy as np
    from pixelsynth_b200 import lmconv as lm
    bg3 = torch.zeros(32, W, W, dtype=torch.bool)
    bg3[:, :, W // 2:] = True
    t0 = time.perf_counter()
    _, order3, words3, smask3 = lm.glue_host(bg3)
    prep3 = model.outpaint2.prepare(order3, words3, smask3, 0)
    host3_ms = 1e3 * (time.perf_counter() - t0)
    codes3 = torch.randint(0, 512, (32, 32, 32), generator=torch.Generator().manual_seed(0)).to(dev)
    uni3 = torch.rand(32, 1024, generator=torch.Generator().manual_seed(1)).to(dev)
    for _ in range(3):
        model.outpaint2.sample(codes3, order3, words3, smask3, uni3, 0.7, prepared=prep3)
    torch.cuda.synchronize()
    L.ps_timing_enable(1)
    for _ in range(5):
        model.outpaint2.sample(codes3, order3, words3, smask3, uni3, 0.7, prepared=prep3)
    torch.cuda.synchronize()
    L.ps_timing_enable(0)
    c3_ms = _lib.kernel_time_ms("lmconv_tc_kernel")[0] / 5
    _lib.kernel_time_ms(None)
    c3_cells, c3_levels, c3_first_b = int(prep3["offs"][-1]), len(prep3["offs"]) - 1, int(prep3["first_b"])
    _lib.check_wedge("bench.py")

    # whole-job sampler counts: every rank's cells (ranks carry different images)
    counts = [cells_sampled, cells_processed]
    if world > 1:
        t = torch.tensor(counts, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        counts = [int(x) for x in t.tolist()]
    job_cells_sampled, job_cells_processed = counts

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tf_peak, peak_src = peaks()
    step_ms = ms_res / args.steps
    per_step = {n: (v[0] / args.steps, v[1] // args.steps) for n, v in kt.items()}
    conv_flops = B * (FLOP_DECODER + FLOP_UNET + FLOP_VQ_ENC + FLOP_VQ_DEC)
    rl = {
        "splat fine_kernel": {"bound": "hbm", "achieved": BYTES_PER_VIEW_SPLAT_FUSED * B / (per_step["fine_kernel"][0] * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s", "ms_per_step": per_step["fine_kernel"][0],
                              "overflow_tiles_ms_per_step": per_step["fine_big_kernel"][0],
                              "note": "maps suppressed inside the pipeline (1.90 MB/view: compute bound); the map-emitting "
                                      "mode (69.0 MB/view) is the line below, measured in the same run; DESIGN.md section 4"},
        "lmconv_tc_kernel": {"bound": "tensor", "achieved": FLOP_PER_CELL * cells_processed / (per_step["lmconv_tc_kernel"][0] * 1e-3) / 1e12,
                                 "peak": tf_peak, "unit": "TFLOP/s", "ms_per_step": per_step["lmconv_tc_kernel"][0],
                                 "cells_processed_per_step": cells_processed, "cells_sampled_per_step": cells_sampled,
                                 "dependency_levels": levels, "prefix_levels_pipelined": levels_prefix,
                                 "sampled_levels_serial": levels - levels_prefix,
                                 "note": "one launch; the sampled levels are a chain of dependent 34-layer columns "
                                         "(a token feeds the next level's first layer), so the step time is "
                                         "sampled_levels x one tile's latency and the tensor pipe idles in between"},
        "splat fine_kernel (maps emitted)": {
            "bound": "hbm", "achieved": BYTES_PER_VIEW_SPLAT_MAPS * nb / (maps_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "ms_per_step": maps_ms, "views_per_launch": nb,
            "note": "not part of the timed step: the map-emitting parity configuration on the U-Net's depth, measured in this run"},
        "splat fine_kernel (maps emitted, uniform depth)": {
            "bound": "hbm", "achieved": BYTES_PER_VIEW_SPLAT_MAPS * nb / (maps_u_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "ms_per_step": maps_u_ms, "views_per_launch": nb,
            "note": "same launch on SURVEY 8d's distribution: depth ~ U[0.5, 10], seed 0"},
        "lmconv_tc_kernel (config 3)": {
            "bound": "tensor", "achieved": FLOP_PER_CELL * c3_cells / (c3_ms * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
            "ms_per_step": c3_ms, "tokens_per_s": 32 * 512 / (c3_ms * 1e-3), "cells_processed": c3_cells,
            "cells_sampled": 32 * 512, "dependency_levels": c3_levels, "prefix_levels": c3_first_b,
            "host_order_masks_levels_ms": host3_ms,
            "note": "BASELINE configs[2]: batch 32, right half of the code grid masked, T 0.7; not part of the timed step"},
        "conv_igemm_kernel": {"bound": "tensor", "achieved": conv_flops / (per_step["conv_igemm_kernel"][0] * 1e-3) / 1e12,
                              "peak": tf_peak, "unit": "TFLOP/s", "ms_per_step": per_step["conv_igemm_kernel"][0],
                              "launches_per_step": per_step["conv_igemm_kernel"][1]},
    }
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    n_small, n_big = pipe.sm_counts if pipe.sm_counts else (0, 0)
    serial_step_ms = ms_serial / args.steps
    for k, v in rl.items():
        v["frac"] = v["achieved"] / v["peak"]
        v["share_of_step"] = 0.0 if ("maps" in k or "config 3" in k) else v["ms_per_step"] / serial_step_ms
    dom = max((k for k in rl if "maps" not in k and "config 3" not in k), key=lambda k: rl[k]["ms_per_step"])
    roof = dict(rl[dom])
    roof.update({"timed_in": "the one-step-at-a-time pass of this run (whole device; `one_step_at_a_time`): kernel durations "
                             "are unambiguous there, while steps in flight queue behind each other's kernels",
                 "kernel": dom, "traffic": traffic.get(dom), "traffic_source": traffic.get("source"), "peak_source": peak_src})
    views_per_step = B * world
    line = {
        "metric": METRIC, "value": views_per_step * args.steps / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": views_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "ViewPipeline.submit(ZbufferModelPts.forward) + BaseModel rescale, host pinned in/out; the host waits for "
                       "each step's pixels"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "rooflines": rl,
        "lmconv_tokens_per_s": job_cells_sampled / (per_step["lmconv_tc_kernel"][0] * 1e-3),
        "lmconv_cells_per_s": job_cells_processed / (per_step["lmconv_tc_kernel"][0] * 1e-3),
        "lmconv_config3_tokens_per_s": 32 * 512 / (c3_ms * 1e-3),
        "ms_per_step_per_rank": [m / args.steps for m in ms_res_ranks],
        "one_step_at_a_time": {"value": views_per_step * args.steps / (ms_serial * 1e-3), "unit": UNIT,
                               "ms_per_step": ms_serial / args.steps,
                               "note": "same model, ZbufferModelPts.forward called serially on the whole device (the latency of a step)"},
        "sm_partition": {"sampler": n_small, "rest": n_big} if pipe.sm_counts else
                        {"unavailable": pipe.partition_error or "one step at a time"},
        "broadcast_ms": bcast_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v0 = 0 if args.view < 0 else args.view
        s, st, n_sampled, wall = cpu_oracle_view_seconds(args.cpu_tokens, cores, view=v0)
        line["cpu_baseline"] = {"value": 1.0 / s, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 view (circle view %d) through the CPU oracle on all %d host threads (%.1f s of CPU "
                                          "wall); sampler timed reference-style on %d tokens (%.3f s/token) and extrapolated "
                                          "to the view's %d sampled cells" %
                                          (v0, cores, wall, args.cpu_tokens, st["lmconv_per_token"], n_sampled),
                                "stage_seconds": st}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
