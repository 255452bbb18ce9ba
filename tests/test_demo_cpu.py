"""CPU: the demo front end, checkpoint ingestion, target-camera synthesis and sample ranking
(8f-3, 8f-4) against vectors produced by the reference's own source (tests/golden/make_demo_golden.py cuts
process_demo_data, get_rt_from_rot and the rank fusion out of /root/reference and runs them unmodified)."""
import argparse
import os
import types

import numpy as np
import pytest
import torch

from pixelsynth_b200 import demo, ranking
from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts, strip_parallel_prefixes


@pytest.fixture(scope="module")
def gold(golden_dir):
    g = np.load(os.path.join(golden_dir, "demo_front.npz"))
    return {k: g[k] for k in g.files}


def test_process_demo_data_matches_reference(gold, golden_dir):
    opts = types.SimpleNamespace(W=256, demo_img_name="demo_input.png")
    batch = demo.process_demo_data(opts, folder=golden_dir)
    assert len(batch["images"]) == 1 and len(batch["cameras"]) == 1
    img = batch["images"][0]
    assert img.dtype == torch.float32 and tuple(img.shape) == (1, 3, 256, 256)
    assert np.array_equal(img.numpy(), gold["image"])                      # same PIL resize, same scaling: bit-exact
    for k in ("P", "Pinv", "K", "Kinv"):
        t = batch["cameras"][0][k]
        assert t.dtype == torch.float32 and np.array_equal(t.numpy(), gold["cam_" + k]), k
    assert tuple(batch["cameras"][0]["OrigP"].shape) == (1, 3, 4)


def test_get_rt_from_rot_matches_reference(gold):
    """All eight rotation directions (with and without --homography), the translation circle 'S' and the rotation
    circle 'C' of scene mode: target pose and its inverse equal to the reference's float32 results."""
    P = torch.from_numpy(gold["cam_P"])
    worst = 0.0
    for case, rt, rtinv in zip(gold["rt_cases"], gold["rt"], gold["rtinv"]):
        setting, d, rot, hom, num, den = str(case).split("|")
        shim = types.SimpleNamespace(opt=types.SimpleNamespace(model_setting=setting, rotation=float(rot),
                                                               homography=bool(int(hom))))
        num, den = int(num), int(den)
        inv, out = ZbufferModelPts.get_rt_from_rot(shim, d, P.clone(), None if num < 0 else num, None if den < 0 else den)
        assert out.dtype == torch.float32 and tuple(out.shape) == (1, 4, 4)
        assert np.array_equal(out.numpy(), rt), case                      # same float64 -> float32 path: bit-exact
        np.testing.assert_allclose(inv.numpy(), rtinv, rtol=0, atol=1e-6, err_msg=str(case))  # LAPACK inverse
        worst = max(worst, float(np.abs(inv.numpy() - rtinv).max()))
    assert len(gold["rt_cases"]) == 93 and worst <= 1e-6


def test_rank_fusion_matches_reference(gold):
    for d, e, (n, best) in zip(gold["rank_d"], gold["rank_e"], gold["rank_best"]):
        assert ranking.rank_fusion(list(d[:n]), list(e[:n])) == int(best)
    with pytest.raises(ValueError):
        ranking.rank_fusion([0.1], [])
    # the preferred candidate: lowest classifier entropy, highest D_Fake
    assert ranking.rank_fusion([0.0, 1.0, 0.5], [3.0, 1.0, 2.0]) == 1


def test_ranker_scores_and_hinge_loss():
    imgs = [torch.full((1, 3, 8, 8), v) for v in (0.1, 0.9, 0.5)]
    r = ranking.Ranker(discriminator=lambda fake, real: fake.mean(),
                       classifier=lambda im: torch.tensor([10.0 * float(im.mean()), 0.0, 0.0]))
    assert r(imgs, imgs[0]) == 1                     # largest D_Fake and the most peaked class distribution
    assert ranking.Ranker()(imgs, imgs[0]) == 0      # no scorers: all ranks tie, the first candidate is kept
    p = torch.tensor([0.7, 0.2, 0.1])
    assert abs(ranking.entropy_of_logits(p.log()) + float((p * p.log()).sum())) < 1e-6
    x0, x1 = torch.tensor([[-2.0, 0.0, 3.0]]), torch.tensor([[-0.5]])
    want = 0.5 * ((0.0 + 1.0 + 4.0) / 3 + 0.5)       # mean(max(x + 1, 0)) per scale, averaged over the two scales
    assert abs(float(ranking.hinge_d_fake([[x0 * 0, x0], [x1]])) - want) < 1e-6


def _fake_checkpoint(tmp_path):
    from pixelsynth_b200 import synthetic

    sd = {}
    for prefix, net in (("pts_regressor.", "unet"), ("vqvae.", "vqvae"), ("outpaint2.", "lmconv"), ("projector.", "decoder")):
        for k, shape in synthetic.shapes()[net].items():
            sd["model.module." + prefix + k] = torch.zeros(shape[0], dtype=getattr(torch, shape[1]))
    sd["model.module.pts_transformer.xyzs"] = torch.zeros(1, 4, 65536)
    sd["model.module.pts_transformer.splatter.ones"] = torch.zeros(1)
    sd["netD.discriminator_0.model0.0.weight"] = torch.zeros(64, 3, 4, 4)
    opts = argparse.Namespace(W=256, radius=4.0, pp_pixel=128, tau=1.0, rad_pow=2, accumulation="alphacomposite",
                              min_z=1.0, max_z=100.0, dataset="realestate", norm_G="sync:spectral_batch", num_samples=7,
                              temperature=0.1, use_rgb_features=True)
    path = str(tmp_path / "pixelsynth.pth")
    torch.save({"state_dict": sd, "opts": opts}, path)
    return path, sd


def test_checkpoint_ingestion_and_option_merge(tmp_path):
    """demo.py:198-208 + utils/opts_helper.py:3-55 on a checkpoint with the reference's layout (BaseModel over
    DataParallel: `model.module.` prefixes, `xyzs` / `ones` buffers, a discriminator, a pickled Namespace)."""
    from pixelsynth_b200 import synthetic

    path, sd = _fake_checkpoint(tmp_path)
    args = demo.build_parser().parse_args(
        ["--vqvae", "--use_fixed_testset", "--model_setting", "gen_img", "--old_model", path, "--gpu", "0,1",
         "--demo_img_name", "1011.png", "--result_folder", "demo/1011", "--temperature=.7", "--num_samples", "50",
         "--direction", "L", "--rotation", ".6"])                       # scripts/demo_image.sh verbatim
    assert args.gpu_ids == "0,1" and args.rotation == 0.6
    state, ck_opts = demo.load_checkpoint(path)
    assert not any("xyzs" in k or "ones" in k for k in state) and len(state) == len(sd) - 2
    state = demo.assemble_state(args, state)
    for prefix, net in (("pts_regressor.", "unet"), ("vqvae.", "vqvae"), ("outpaint2.", "lmconv"), ("projector.", "decoder")):
        sub = {k[len(prefix):] for k in state if k.startswith(prefix)}
        assert sub == set(synthetic.shapes()[net]), prefix               # every tensor the kernels' packers read
    opts = demo.opts_helper(args, ck_opts)
    assert opts.num_samples == 50 and opts.temperature == 0.7 and opts.direction == "L" and opts.rotation == 0.6
    assert opts.model_setting == "gen_img" and opts.vqvae is True and opts.min_z == 1.0 and opts.max_z == 100.0
    assert opts.background_smoothing_kernel_size == 13 and opts.normalize_before_residual is False
    assert opts.homography is False and opts.no_outpainting is False and opts.gpu_ids == "0,1" and opts.num_split == 1
    assert opts.dataset == "realestate" and opts.train_depth is False and opts.isTrain is True

    mp3d = demo.build_parser().parse_args(["--old_model", "modelcheckpoints/mp3d/pixelsynth.pth"])
    assert demo.opts_helper(mp3d, ck_opts).normalize_before_residual is True   # opts_helper.py:48-50

    with pytest.raises(KeyError):
        bad = str(tmp_path / "bad.pth")
        torch.save({"weights": {}}, bad)
        demo.load_checkpoint(bad)


def test_separately_trained_overlays(tmp_path):
    """--load_vqvae strips `module.` (demo.py:210-216); --load_autoregressive reads `model_state_dict` (:219-221)."""
    path, _ = _fake_checkpoint(tmp_path)
    vq, ar = str(tmp_path / "vqvae_150.pt"), str(tmp_path / "ar.pt")
    torch.save({"module.quantize_t.embed": torch.ones(64, 512), "module.enc_b.blocks.0.bias": torch.ones(64)}, vq)
    torch.save({"model_state_dict": {"nin_out.lin_a.bias": torch.full((512,), 2.0)}}, ar)
    args = demo.build_parser().parse_args(["--old_model", path, "--load_vqvae", "--vqvae_path", vq,
                                           "--load_autoregressive", "--autoregressive", ar])
    state = demo.assemble_state(args, demo.load_checkpoint(path)[0])
    assert {k for k in state if k.startswith("vqvae.")} == {"vqvae.quantize_t.embed", "vqvae.enc_b.blocks.0.bias"}
    assert float(state["vqvae.quantize_t.embed"].sum()) == 64 * 512
    assert float(state["outpaint2.nin_out.lin_a.bias"][0]) == 2.0
    assert strip_parallel_prefixes({"module.a": 1, "model.module.b": 2, "c": 3}) == {"a": 1, "b": 2, "c": 3}


def test_default_options_without_a_checkpoint():
    args = demo.build_parser().parse_args(["--model_setting", "gen_scene", "--directions", "R", "L", "U", "--num_split", "4",
                                           "--seed", "5"])
    opts = demo.opts_helper(args, None)
    assert opts.seed == 5 and opts.directions == ["R", "L", "U"] and opts.num_split == 4 and opts.W == 256
    assert opts.pp_pixel == 128 and opts.radius == 4.0 and opts.accumulation == "alphacomposite"


def test_output_files(tmp_path):
    """File names of demo.py:100-178, pixel values of torchvision.utils.save_image (x*255 + .5, clamp, truncate)."""
    from PIL import Image

    folder = str(tmp_path / "out")
    ops_ = types.SimpleNamespace(result_folder=folder, direction="L", rotation=0.6, directions=["R", "U", "S"], num_split=2)
    g = torch.Generator().manual_seed(0)
    img = torch.rand(1, 3, 16, 16, generator=g)
    demo.save_img({"PredImg": img, "FeaturesImg": img * 0.5}, ops_)
    assert sorted(os.listdir(folder)) == ["input_fs_image_L_0.png", "output_image_L_0.png"]
    back = np.asarray(Image.open(os.path.join(folder, "output_image_L_0.png")))
    want = img[0].mul(255).add(0.5).clamp(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
    assert np.array_equal(back, want)
    preds = {"PredImg_%s_%d" % (d, i): img for d in ("R", "U", "S") for i in range(0, 5)}
    demo.save_scene(preds, ops_)
    assert sorted(os.listdir(os.path.join(folder, "scene"))) == ["output_image_R_0001.png", "output_image_R_0002.png",
                                                                  "output_image_U_0001.png"]
    keys = demo.video_frames(2)
    assert keys[:5] == ["PredImg_R_0", "PredImg_R_1", "PredImg_R_1", "PredImg_R_0", "PredImg_L_1"]
    assert len(keys) == 1 + 2 * (1 + 2) + 4 * 3
    grid = str(tmp_path / "grid.png")
    demo.save_image(torch.ones(3, 3, 4, 4), grid)
    assert np.asarray(Image.open(grid)).shape == (4 + 2 * 2, 3 * 4 + 4 * 2, 3)


def test_demo_refuses_to_run_without_a_gpu(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        demo.main(["--model_setting", "gen_img", "--direction", "L", "--demo_img_name", "x.png", "--result_folder", str(tmp_path)])


def test_reference_import_paths_resolve_to_the_mirrors():
    """north_star: models.z_buffermodel / BaseModel.forward / demo.py stay drop-in -- the reference's import paths
    (demo.py:9-14, options/options.py:3-18) resolve to the sm_100a mirrors."""
    import importlib

    import pixelsynth_b200.models.base_model as bm
    import pixelsynth_b200.models.z_buffermodel as zb

    assert importlib.import_module("models.z_buffermodel").ZbufferModelPts is zb.ZbufferModelPts
    assert importlib.import_module("models.base_model").BaseModel is bm.BaseModel
    assert importlib.import_module("models.projection.z_buffer_manipulator").PtsManipulator is not None
    assert importlib.import_module("models.layers.z_buffer_layers").RasterizePointsXYsBlending is not None
    top = importlib.import_module("demo")
    assert top.main is demo.main


def test_ragged_cloud_compaction_pads_with_zero_points():
    """PtsManipulator._compact (batched gen_scene): image b's selected columns first, in order, zero columns behind."""
    from pixelsynth_b200.models.projection.z_buffer_manipulator import PtsManipulator

    x = torch.arange(2 * 3 * 6, dtype=torch.float32).view(2, 3, 6) + 1
    sel = torch.tensor([[1, 0, 1, 1, 0, 0], [0, 1, 0, 0, 0, 0]], dtype=torch.bool)
    out = PtsManipulator._compact(sel, x)
    assert out.shape == (2, 3, 3)
    assert torch.equal(out[0], x[0][:, [0, 2, 3]])
    assert torch.equal(out[1, :, 0], x[1][:, 1]) and (out[1, :, 1:] == 0).all()
    # equal counts (the reference's only working case): exactly the boolean-mask view
    sel2 = torch.tensor([[1, 0, 1, 0, 0, 0], [0, 1, 0, 0, 0, 1]], dtype=torch.bool)
    ref = x[sel2.unsqueeze(1).repeat(1, 3, 1)].view(2, 3, -1)
    assert torch.equal(PtsManipulator._compact(sel2, x), ref)


def test_pil_resize_table_reproduces_pil_bit_for_bit():
    """The classifier's input is resized by PIL (z_buffermodel.py:105-110).  The product's integer filter table
    (nets.pil_bilinear_table, consumed by classifier_input_kernel) applied the way the kernel applies it -- horizontal
    pass, vertical pass, (2^21 + sum kk p) >> 22 clipped to uint8 -- equals PIL's own resize exactly."""
    from PIL import Image
    from pixelsynth_b200.nets import pil_bilinear_table

    t0, kk = [t.numpy().astype(np.int64) for t in pil_bilinear_table()]
    a = np.random.default_rng(3).integers(0, 256, (256, 256, 3)).astype(np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((224, 224), Image.BILINEAR)).astype(np.int64)
    x = a.astype(np.int64)
    h = np.stack([np.clip(((1 << 21) + sum(kk[ox, k] * x[:, t0[ox] + k] for k in range(4) if kk[ox, k])) >> 22, 0, 255)
                  for ox in range(224)], 1)
    v = np.stack([np.clip(((1 << 21) + sum(kk[oy, k] * h[t0[oy] + k] for k in range(4) if kk[oy, k])) >> 22, 0, 255)
                  for oy in range(224)], 0)
    assert np.array_equal(v, ref)
