"""GPU: the drop-in entry points (ZbufferModelPts.forward, BaseModel.__call__) end to end on seeded weights.
Stage parity is covered by test_splat_gpu / test_nets_gpu / test_lmconv_gpu; here the stages are chained the way
models/z_buffermodel.py:291-419 chains them and checked against the oracle at the seams that stay comparable
(the splat of the model's own predicted depth; the decoder on the model's own combined image)."""
import types

import numpy as np
import pytest
import torch

from util import demo_cameras

pytestmark = pytest.mark.gpu


def make_opt(**kw):
    o = dict(W=256, splatter="xyblending", learn_default_feature=True, radius=4.0, pp_pixel=128, rad_pow=2, tau=1.0,
             accumulation="alphacomposite", background_smoothing_kernel_size=13, min_z=0.5, max_z=10.0,
             use_rgb_features=True, use_gt_depth=False, use_inverse_depth=False, depth_predictor_type="unet",
             no_outpainting=False, vqvae=True, num_samples=1, temperature=0.7, model_setting="gen_paired_img",
             direction="L", rotation=0.6, homography=False, seed=0, normalize_image=True, predict_residual=True,
             normalize_before_residual=False, refine_model_type="resnet_256W8UpDown3", ngf=64, norm_G="sync:spectral_batch")
    o.update(kw)
    return types.SimpleNamespace(**o)


def make_batch(B, kind="translate", seed=0):
    from pixelsynth_b200 import synthetic

    K, Kinv, RT1, RT1inv, RT2, RT2inv = [torch.from_numpy(m) for m in demo_cameras(B, kind, seed)]
    img = synthetic.synth_image(B, seed)
    cam0 = {"K": K, "Kinv": Kinv, "P": RT1, "Pinv": RT1inv}
    cam1 = {"K": K, "Kinv": Kinv, "P": RT2, "Pinv": RT2inv}
    return {"images": [img, img.clone()], "cameras": [cam0, cam1]}


@pytest.fixture(scope="module")
def model():
    from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts

    return ZbufferModelPts(make_opt())


def test_gen_paired_img_end_to_end(model, oracle):
    from oracle import nets_ref
    from pixelsynth_b200 import synthetic
    from pixelsynth_b200.models.base_model import BaseModel

    B = 2
    batch = make_batch(B)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(16, B, 20, generator=g)
    uniforms = torch.rand(B, 1024, generator=g)
    loss, out = model.forward(batch, noise=noise, uniforms=uniforms)
    torch.cuda.synchronize()
    for k in ("InputImg", "PredImg", "PredDepthImg", "ForegroundImg", "FeaturesImg", "OutputImg"):
        assert k in out
    assert out["PredImg"].shape == (B, 3, 256, 256) and torch.isfinite(out["PredImg"]).all()
    assert out["PredImg"].abs().max() <= 1.0 + 1e-6
    assert out["ForegroundImg"].shape == (B, B, 256, 256)        # the reference's repeat quirk (z_buffermodel.py:389)
    last = model.last
    assert last["sample_mask"].any()                             # the translated view really has something to outpaint
    # seam 1: the splat of the model's own depth equals the oracle's splat of that depth
    cams = demo_cameras(B, "translate", 0)
    from util import pack_mats
    ref = oracle.splat(last["depth"].cpu().numpy(), batch["images"][0].numpy(), pack_mats(*cams), 256, K=128, radius_px=4.0)
    np.testing.assert_allclose(last["gen_fs"].cpu().numpy(), ref["out"], rtol=0, atol=2e-6)
    assert np.array_equal(last["background_mask"].cpu().numpy(), ref["bg"])
    # seam 2: the sampled cells are exactly the all-background cells (every one of the cell's 64 pixels is background,
    # z_buffermodel.py:646-669), and the sampler leaves every other cell at the encoder's code
    bg = last["background_mask"].cpu()
    cell_all_bg = bg.view(B, 32, 8, 32, 8).permute(0, 1, 3, 2, 4).reshape(B, 32, 32, 64).all(-1).numpy()
    assert np.array_equal(last["sample_mask"], cell_all_bg)
    sampled = model.outpaint2.sample(last["codes"], last["order"], last["words"], last["sample_mask"], uniforms, 0.7).cpu()
    keep = ~torch.from_numpy(last["sample_mask"])
    assert torch.equal(sampled[keep], last["codes"].cpu()[keep])
    # seam 3: deterministic under injected noise / uniforms
    _, out2 = model.forward(batch, noise=noise, uniforms=uniforms)
    assert torch.equal(out["PredImg"], out2["PredImg"])
    # BaseModel rescales every *Img* output to [0,1] (base_model.py:96-99)
    bm = BaseModel(model, model.opt)
    _, o3, b3 = bm(batch, isval=True, return_batch=True)
    assert o3["PredImg"].min() >= -1e-6 and o3["PredImg"].max() <= 1 + 1e-6 and b3 is batch


def test_end_to_end_tolerance_teacher_forced(model, oracle):
    """ONE stated PredImg tolerance for the whole chain: the fp32 CPU oracle chain, teacher-forced
    on the GPU's discrete decisions (it consumes the GPU's depth for the splat, and the GPU's codes / sampled tokens
    for the decode), must reproduce PredImg within 0.02 rms / 0.10 max on [-1,1] (bf16 convolutions); the continuous
    stages feeding those decisions are bounded separately: depth <= 1.5% rms of range, <= 5% VQ code flips, sampled
    tokens >= 98.5% equal to the oracle's draws."""
    from oracle import lmconv_ref, nets_ref
    from pixelsynth_b200 import synthetic
    from util import pack_mats

    B = 2
    batch = make_batch(B)
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(16, B, 20, generator=g)
    uniforms = torch.rand(B, 1024, generator=g)
    _, out = model.forward(batch, noise=noise, uniforms=uniforms)
    torch.cuda.synchronize()
    last = model.last
    sds = {n: synthetic.make_state(n, 0) for n in ("unet", "vqvae", "lmconv", "decoder")}
    img = batch["images"][0]
    with torch.no_grad():
        # D1: depth
        depth_ref = nets_ref.unet_depth(sds["unet"], img, 0.5, 10.0)
        derr = (last["depth"].cpu() - depth_ref)
        assert derr.pow(2).mean().sqrt().item() <= 0.015 * 9.5
        # S: oracle splat of the GPU depth (bit-exact maps => same image / mask)
        ref = oracle.splat(last["depth"].cpu().numpy(), img.numpy(), pack_mats(*demo_cameras(B, "translate", 0)), 256, K=128,
                           radius_px=4.0)
        gen_fs, bgm = torch.from_numpy(ref["out"]), torch.from_numpy(ref["bg"])
        assert np.array_equal(last["background_mask"].cpu().numpy(), ref["bg"])
        # V1: codes
        ids_ref, _ = nets_ref.vqvae_encode_top(sds["vqvae"], gen_fs)
        flips = (ids_ref != last["codes"].cpu()).float().mean().item()
        assert flips <= 0.05, flips
        # L: tokens, teacher-forced on the GPU's own result
        sampled = model.outpaint2.sample(last["codes"], last["order"], last["words"], last["sample_mask"], uniforms, 0.7).cpu()
        data = torch.nn.functional.one_hot(sampled, 512).permute(0, 3, 1, 2).float()
        words, order, smask = last["words"], last["order"], last["sample_mask"]
        mf = [torch.cat([lmconv_ref.masks_to_float(words[b, k]) for b in range(B)]) for k in range(3)]
        lg = lmconv_ref.lmconv_logits(sds["lmconv"], data, *mf)
        bad = tot = 0
        for b in range(B):
            k = 0
            for cell in order[b]:
                r, c = divmod(int(cell), 32)
                if smask[b, r, c]:
                    bad += int(lmconv_ref.draw(lg[b, :, r, c], 0.7, float(uniforms[b, k])) != int(sampled[b, r, c]))
                    tot += 1
                    k += 1
        assert tot > 0 and bad <= 0.015 * tot + 1, (bad, tot)
        # V2 + C1 + R: decode the GPU's tokens, combine, refine with the same noise
        ar = nets_ref.vqvae_decode_code(sds["vqvae"], sampled)
        comb = gen_fs * (~bgm)[:, None].float() + ar * bgm[:, None].float()
        pred_ref = nets_ref.decoder_forward(sds["decoder"], comb, bgm, [noise[i] for i in range(16)])
    err = out["PredImg"].cpu() - pred_ref
    rms, mx = err.pow(2).mean().sqrt().item(), err.abs().max().item()
    print("end to end (teacher-forced): PredImg rms %.4f max %.4f; depth rms %.4f; code flips %.2f%%; token flips %d/%d"
          % (rms, mx, derr.pow(2).mean().sqrt().item(), 100 * flips, bad, tot))
    assert rms <= 0.02 and mx <= 0.10


def test_gen_img_direction(model):
    model.opt.model_setting = "gen_img"
    try:
        batch = make_batch(1, "identity")
        loss, out = model.forward(batch)
        assert out["PredImg"].shape == (1, 3, 256, 256) and "OutputImg" not in out
        # direction L, rotation 0.6: pure rotation leaves a band of background to outpaint
        assert model.last["sample_mask"].sum() > 50
    finally:
        model.opt.model_setting = "gen_paired_img"


def test_gen_scene_sweep(model, oracle):
    """gen_scene (z_buffermodel.py:421-592): a short sweep over the growing point cloud.  Keys and shapes follow the
    reference; the last view's cumulative splat is recomputed by the oracle from what the model recorded going in."""
    from util import pack_mats

    model.opt.model_setting = "gen_scene"
    model.opt.directions, model.opt.num_split, model.opt.sequential_outpainting = ["L"], 2, False
    try:
        batch = make_batch(1, "identity")
        g = torch.Generator().manual_seed(7)
        _, out = model.forward(batch, noise=torch.randn(16, 1, 20, generator=g), uniforms=torch.rand(1, 1024, generator=g))
        torch.cuda.synchronize()
        # far view first, then the views in between back to the input pose
        assert [s["num"] for s in model.last_scene] == [2, 1, 0]
        for k in ("InputImg", "PredImg_L_2", "PredImg_L_1", "PredImg_L_0", "FeaturesImg_L_2", "FeaturesImg_L_0",
                  "PredDepthImg_L_2", "ForegroundImg_L_2"):
            assert k in out, k
        for i in range(3):
            img = out["PredImg_L_%d" % i]
            assert img.shape == (1, 3, 256, 256) and torch.isfinite(img).all() and img.abs().max() <= 1 + 1e-6
        # the cloud grows by exactly the pixels the previous view had to outpaint
        P = 256 * 256
        sizes = [s["cloud"].shape[2] for s in model.last_scene]
        bgs = [int(s["background_mask"].sum()) for s in model.last_scene]
        assert sizes[0] == P and sizes[1] == P + bgs[0] and sizes[2] == P + bgs[0] + bgs[1] and bgs[0] > 0
        # last view through the oracle: new pixels (masked by the previous view's background) + the prior cloud
        s = model.last_scene[-1]
        f = lambda t: t.detach().float().cpu().numpy()
        K = f(batch["cameras"][0]["K"])
        Kinv = f(batch["cameras"][0]["Kinv"])
        mats = pack_mats(K, Kinv, f(s["src_rt"]), f(s["src_inv"]), f(s["dst_rt"]), f(s["dst_inv"]))
        pts_n, _ = oracle.project(f(s["depth"]), mats, 256, want_xyproj=True)
        sel = f(s["prior_bg"]).reshape(-1).astype(bool)
        mats3 = np.ascontiguousarray(np.stack([K.reshape(1, 16), f(s["dst_rt"]).reshape(1, 16), f(s["prior_out_inv"]).reshape(1, 16)], 1))
        pts_o, _ = oracle.project_cloud(f(s["prior_cloud"]), mats3)
        pts_c = np.concatenate([pts_n[:, sel], pts_o], 1)
        feat_c = np.concatenate([f(s["src"]).reshape(1, 3, -1)[:, :, sel], f(s["prior_feats"])], 2)
        radius = 4.0 / 256 * 2.0
        idx, _, d2 = oracle.rasterize(pts_c, 256, 128, radius)
        np.testing.assert_allclose(f(s["gen_fs"]), oracle.composite(idx, d2, feat_c, radius), rtol=0, atol=2e-6)
        assert np.array_equal(s["background_mask"].cpu().numpy(), oracle.bgmask(idx, 13))
    finally:
        model.opt.model_setting = "gen_paired_img"


def test_soak_batch64_all_views():
    """Round 1's intermittent launch failure (profiles/r02_launch_failure_rootcause.txt) showed up as a CUDA fault or as
    silently different tokens at batch 64-128.  30 full steps at batch 64 with all eight circle views in the batch: no
    fault, no wedged barrier, and bit-identical output every step."""
    from bench import make_batch as bench_batch
    from pixelsynth_b200 import _lib
    from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts

    B = 64
    m = ZbufferModelPts(make_opt())
    hb = bench_batch(B, [i % 8 for i in range(B)])
    db = {"images": [t.cuda() for t in hb["images"]], "cameras": [{k: v.cuda() for k, v in c.items()} for c in hb["cameras"]]}
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(16, B, 20, generator=g).cuda()
    uniforms = torch.rand(B, 1024, generator=g)
    ref = None
    for step in range(30):
        out = m.forward(db, noise=noise, uniforms=uniforms)[1]["PredImg"]
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(out, ref), "step %d differs from step 0" % step
    torch.cuda.synchronize()
    _lib.check_wedge("soak")
    assert torch.isfinite(ref).all()


def test_gen_scene_batched_equals_single(model):
    """BASELINE config 5: a BATCH of images sweeps a scene in lock step (the reference renders scenes at batch 1 only).
    The clouds are ragged per image -- the two images outpaint different numbers of pixels -- and zero-padded; image b
    of the batched sweep must equal the batch-1 sweep of image b: splat outputs bit-exact, refined images equal."""
    from pixelsynth_b200 import synthetic

    model.opt.model_setting = "gen_scene"
    # 'S' = the translation circle: what a view uncovers depends on the image's depth, so the clouds really are ragged
    model.opt.directions, model.opt.num_split, model.opt.sequential_outpainting = ["S"], 1, False
    try:
        B = 2
        batch = make_batch(B, "identity")
        batch["images"][0] = synthetic.synth_image(B, 3)           # two different images
        g = torch.Generator().manual_seed(11)
        noise, uniforms = torch.randn(16, B, 20, generator=g), torch.rand(B, 1024, generator=g)
        _, out = model.forward(batch, noise=noise, uniforms=uniforms)
        torch.cuda.synchronize()
        scene_b = model.last_scene
        keys = [k for k in out if k.startswith(("PredImg_", "FeaturesImg_"))]
        assert len(keys) == 2 * len(scene_b) == 6 and all(out[k].shape[0] == B for k in keys)
        n_bg = [int(s["background_mask"][b].sum()) for s in scene_b[:1] for b in range(B)]
        assert n_bg[0] != n_bg[1]                                    # really ragged
        # every view appends max-over-images(newly outpainted pixels) columns (shorter images are zero-padded)
        grown = sum(max(int(s["background_mask"][b].sum()) for b in range(B)) for s in scene_b[:-1])
        assert scene_b[-1]["cloud"].shape[2] == 256 * 256 + grown
        for b in range(B):
            one = {"images": [t[b:b + 1] for t in batch["images"]],
                   "cameras": [{k: v[b:b + 1] for k, v in c.items()} for c in batch["cameras"]]}
            _, o1 = model.forward(one, noise=noise[:, b:b + 1], uniforms=uniforms[b:b + 1])
            torch.cuda.synchronize()
            for k in keys:
                if k.startswith("FeaturesImg_"):
                    assert torch.equal(out[k][b], o1[k][0]), (k, b)
                else:
                    assert (out[k][b] - o1[k][0]).abs().max().item() <= 1e-5, (k, b)
    finally:
        model.opt.model_setting = "gen_paired_img"


def test_num_samples_ranking_matches_oracle_scores(oracle):
    """z_buffermodel.py:244-276 with num_samples = 3: the candidates come out of ONE sampler launch / decode / refinement
    batch, are scored on the GPU in one batch, and the GPU's scores and choice agree with the fp32 oracle's scorers run
    on the same candidate images (discriminator D_Fake within 2%, entropy within 0.03 nat, same rank fusion)."""
    from oracle import nets_ref
    from pixelsynth_b200 import ranking, synthetic
    from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts

    m = ZbufferModelPts(make_opt(num_samples=3))
    B = 1
    batch = make_batch(B)
    g = torch.Generator().manual_seed(21)
    uniforms = torch.rand(3, B, 1024, generator=g)           # a different draw per candidate
    noise = torch.randn(16, 3 * B, 20, generator=g)
    _, out = m.forward(batch, noise=noise, uniforms=uniforms)
    torch.cuda.synchronize()
    assert out["PredImg"].shape == (B, 3, 256, 256)
    d, e = m.ranker.last_scores
    assert len(d) == len(e) == 3 and len(set(np.round(d, 6))) == 3     # three different candidates
    # the same candidates through the oracle's scorers
    n = 3
    m.opt.num_samples = 1
    cands = []
    for i in range(n):
        _, o = m.forward(batch, noise=noise[:, i * B:(i + 1) * B], uniforms=uniforms[i])
        cands.append(o["PredImg"].cpu())
    sdD, sdC = synthetic.make_state("netD", 0), synthetic.make_state("resnet18", 0)
    with torch.no_grad():
        d_ref = [float(nets_ref.d_fake(nets_ref.discriminator_forward(sdD, c))) for c in cands]
        e_ref = [float(nets_ref.entropy(nets_ref.resnet18_logits(sdC, nets_ref.classifier_input(c[0])))) for c in cands]
    print("D_Fake gpu", d, "oracle", d_ref, "| entropy gpu", e, "oracle", e_ref)
    np.testing.assert_allclose(d, d_ref, rtol=2e-2, atol=1e-3)
    np.testing.assert_allclose(e, e_ref, atol=3e-2)
    if min(abs(a - b) for i, a in enumerate(d_ref) for b in d_ref[i + 1:]) > 0.05 * max(d_ref) and \
            min(abs(a - b) for i, a in enumerate(e_ref) for b in e_ref[i + 1:]) > 0.1:
        assert m.last_best == ranking.rank_fusion(d_ref, e_ref)
    assert torch.equal(out["PredImg"].cpu(), cands[m.last_best])     # the folded batch equals the per-candidate calls


@pytest.mark.parametrize("partition,depth", [(True, 2), (False, 2), (True, 3)])
def test_view_pipeline_two_in_flight_equals_serial(model, partition, depth):
    """pixelsynth_b200.pipeline.ViewPipeline: two batches in flight, the sampler on its own SM partition (CUDA green
    contexts), give bit-identical images to one forward at a time -- five different batches, different views."""
    from pixelsynth_b200 import _lib
    from pixelsynth_b200.pipeline import ViewPipeline

    B = 6
    g = torch.Generator().manual_seed(11)
    jobs = []
    for i in range(5):
        jobs.append((make_batch(B, "translate" if i % 2 == 0 else "rotate", seed=i),
                     torch.randn(16, B, 20, generator=g), torch.rand(B, 1024, generator=g)))
    want = []
    for b, n, u in jobs:
        want.append(model.forward(b, noise=n, uniforms=u)[1]["PredImg"].clone())
    torch.cuda.synchronize()
    with ViewPipeline(model, depth=depth, sampler_sms=24, partition=partition) as pipe:
        if partition:
            ns, nb = pipe.sm_counts
            assert ns >= 24 and ns + nb == torch.cuda.get_device_properties(0).multi_processor_count
            L = _lib.lib()
            assert L.ps_stream_sm_count(pipe.sampler_stream.cuda_stream) == ns
            assert L.ps_stream_sm_count(pipe.streams[0].cuda_stream) == nb
            assert L.ps_stream_sm_count(pipe.front_streams[depth - 1].cuda_stream) == nb
            assert L.ps_stream_sm_count(torch.cuda.current_stream().cuda_stream) == ns + nb
        tickets = [pipe.submit(b, noise=n, uniforms=u) for b, n, u in jobs]
        got = [pipe.result(t)[1]["PredImg"] for t in tickets]
        for w, x in zip(want, got):
            assert torch.equal(w, x)
        # the generator form, and the `then` hook (runs on the slot's stream)
        outs = list(pipe.map([j[0] for j in jobs[:3]], noise=jobs[0][1], uniforms=jobs[0][2]))
        assert torch.equal(outs[0][1]["PredImg"], want[0])
        t = pipe.submit(jobs[1][0], then=lambda loss, o: (0.5 * o["PredImg"] + 0.5).cpu(), noise=jobs[1][1], uniforms=jobs[1][2])
        assert torch.equal(pipe.result(t), (0.5 * want[1] + 0.5).cpu())


def test_use_inverse_depth_option(model):
    """z_buffermodel.py:311-315: depth = 1 / (sigmoid(Unet(x)) * 10 + 0.01) for long-tailed depth distributions.
    Against the fp32 oracle U-Net at the depth bar of DESIGN.md section 2, taken on the sigmoid the branch inverts."""
    from oracle import nets_ref
    from pixelsynth_b200 import synthetic

    model.opt.use_inverse_depth = True
    try:
        batch = make_batch(1)
        loss, out = model.forward(batch)
        depth = model.last["depth"].cpu()
        with torch.no_grad():
            ref = 1.0 / (torch.sigmoid(nets_ref.unet_features(synthetic.make_state("unet", 0), batch["images"][0])) * 10 + 0.01)
        assert depth.shape == ref.shape and 0.0999 <= float(depth.min()) and float(depth.max()) <= 100.0
        sig = lambda d: (1.0 / d - 0.01) / 10                      # back to the sigmoid: the U-Net bar is 1.5 % rms of its range
        assert float((sig(depth) - sig(ref)).pow(2).mean().sqrt()) <= 0.015 and torch.isfinite(out["PredImg"]).all()
        assert torch.allclose(out["PredDepthImg"].cpu(), depth / 5 - 1, rtol=1e-6, atol=1e-6)
    finally:
        model.opt.use_inverse_depth = False
