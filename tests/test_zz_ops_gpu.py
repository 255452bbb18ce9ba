"""GPU: the torch.ops.pixelsynth_b200.* seams of SURVEY.md 8b that the class mirrors do not already exercise --
rasterize_points_zbuf (the PyTorch3D-shaped seam), splat_cumulative, the network ops behind weight handles,
gen_order_masks, lmconv_sample / lmconv_logits, combine.  Each op is checked against the oracle where one exists
(bit-exact for maps / masks / orders) and against the class mirror's own result otherwise (same kernels, so equal)."""
import os
import types

import numpy as np
import pytest
import torch

from util import demo_cameras, pack_mats, synthetic_view

pytestmark = pytest.mark.gpu
ATOL_OUT = 2e-6


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    import pixelsynth_b200.ops as o

    return o


@pytest.fixture(scope="module")
def handles(ops):
    from pixelsynth_b200 import synthetic

    h = {k: ops.register_weights(k, synthetic.make_state(n, 0))
         for k, n in (("depth_unet", "unet"), ("vqvae", "vqvae"), ("refine", "decoder"), ("lmconv", "lmconv"))}
    yield h
    for v in h.values():
        ops.release_weights(v)


def test_rasterize_points_zbuf_matches_oracle(ops, oracle):
    W, K = 64, 32
    depth, _, mats = synthetic_view(2, W, kind="rotate", seed=5)
    pts = oracle.project(depth, mats, W)
    idx, zbuf, d2 = torch.ops.pixelsynth_b200.rasterize_points_zbuf(dev(pts), W, 4.0, K, True)
    ridx, rz, rd2 = oracle.rasterize(pts, W, K, 4.0 / W * 2.0)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (2, W, W, K)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(zbuf.cpu().numpy(), rz)
    assert np.array_equal(d2.cpu().numpy(), rd2)
    with pytest.raises(RuntimeError):
        torch.ops.pixelsynth_b200.rasterize_points_zbuf(dev(pts)[:, :, :2], W, 4.0, K, True)
    with pytest.raises(RuntimeError):  # K beyond the compiled maximum (PyTorch3D: "Must have num_closest <= 150")
        torch.ops.pixelsynth_b200.rasterize_points_zbuf(dev(pts), W, 4.0, 129, False)


def test_splat_cumulative_op_two_views(ops, oracle):
    """Same scenario as test_splat_gpu.test_cumulative_cloud_two_views, through the op."""
    W, B, Kpp = 64, 1, 128
    radius = 4.0 / W * 2.0
    depth1, feat1, _ = synthetic_view(B, W, kind="translate", seed=21, depth_mode="smooth")
    depth2, feat2, _ = synthetic_view(B, W, kind="translate", seed=22, depth_mode="smooth")
    K, Kinv, RT1, RT1inv, RT2, RT2inv = demo_cameras(B, "translate", 21, views=[2])
    _, _, _, _, RT3, RT3inv = demo_cameras(B, "translate", 21, views=[3])
    RT2[:, 0, 3] += 0.6
    RT2inv = np.linalg.inv(RT2).astype(np.float32)
    op = torch.ops.pixelsynth_b200.splat_cumulative
    cfg = (W, W, Kpp, 4.0, 1.0, 2, 0, 13, 1e-2)

    m1 = pack_mats(K, Kinv, RT1, RT1inv, RT2, RT2inv)
    res1, bg1, cloud1, src1 = op(dev(depth1), dev(feat1), dev(m1), None, None, None, None, *cfg)
    pts_a, xyp_a = oracle.project(depth1, m1, W, want_xyproj=True)
    idx_a, _, d2_a = oracle.rasterize(pts_a, W, Kpp, radius)
    bg_a = oracle.bgmask(idx_a, 13)
    assert bg1.dtype == torch.bool and np.array_equal(bg1.cpu().numpy(), bg_a) and 0 < bg_a.sum() < bg_a.size
    assert np.array_equal(cloud1.cpu().numpy(), xyp_a)
    np.testing.assert_allclose(res1.cpu().numpy(), oracle.composite(idx_a, d2_a, feat1.reshape(B, 3, -1), radius), rtol=0,
                               atol=ATOL_OUT)

    m2 = pack_mats(K, Kinv, RT2, RT2inv, RT3, RT3inv)
    res2, bg2, cloud2, src2 = op(dev(depth2), dev(feat2), dev(m2), cloud1, src1, bg1, dev(RT2inv), *cfg)
    sel = bg_a.reshape(B, -1)[0]
    pts_n, xyp_n = oracle.project(depth2, m2, W, want_xyproj=True)
    mats3 = np.ascontiguousarray(np.stack([K.reshape(B, 16), RT3.reshape(B, 16), RT2inv.reshape(B, 16)], 1))
    pts_o, xyp_o = oracle.project_cloud(xyp_a, mats3)
    pts_c = np.concatenate([pts_n[:, sel], pts_o], 1)
    feat_c = np.concatenate([feat2.reshape(B, 3, -1)[:, :, sel], feat1.reshape(B, 3, -1)], 2)
    idx_c, _, d2_c = oracle.rasterize(pts_c, W, Kpp, radius)
    assert np.array_equal(cloud2.cpu().numpy(), np.concatenate([xyp_n[:, :, sel], xyp_o], 2))
    assert np.array_equal(src2.cpu().numpy(), feat_c)
    assert np.array_equal(bg2.cpu().numpy(), oracle.bgmask(idx_c, 13))
    np.testing.assert_allclose(res2.cpu().numpy(), oracle.composite(idx_c, d2_c, feat_c, radius), rtol=0, atol=ATOL_OUT)
    with pytest.raises(RuntimeError):  # a prior feature set without its cloud
        op(dev(depth2), dev(feat2), dev(m2), None, src1, bg1, dev(RT2inv), *cfg)


def test_network_ops_behind_handles(ops, handles):
    """depth_unet / vqvae_* / refine_decode / combine / gen_order_masks / lmconv_*: one pass of the demo path written
    with ops only, equal to the class mirror (ZbufferModelPts) on the same weights, noise and uniforms."""
    from pixelsynth_b200 import synthetic
    from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts
    from test_pipeline_gpu import make_batch, make_opt

    P = torch.ops.pixelsynth_b200
    B = 2
    batch = make_batch(B, "translate", 3)
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(16, B, 20, generator=g)
    uniforms = torch.rand(B, 1024, generator=g)
    model = ZbufferModelPts(make_opt())
    _, ref = model.forward(batch, noise=noise, uniforms=uniforms)
    last = model.last

    img = batch["images"][0].cuda()
    cam0, cam1 = batch["cameras"]
    depth = P.depth_unet(img, handles["depth_unet"], 0.5, 10.0)
    assert torch.equal(depth, last["depth"])
    mats = ops.pack_mats(cam0["K"], cam0["Kinv"], cam0["P"], cam0["Pinv"], cam1["P"], cam1["Pinv"]).cuda()
    gen_fs, bg, _, _, _ = P.splat(depth, img, mats, 256, 256, 128, 4.0, 1.0, 2, 0, 13, 1e-2, False, False)
    assert torch.equal(gen_fs, last["gen_fs"]) and torch.equal(bg, last["background_mask"])
    dist, order, words, smask = P.gen_order_masks(bg)
    assert order.dtype == torch.int32 and tuple(words.shape) == (B, 3, 1024) and smask.dtype == torch.bool
    assert np.array_equal(order.numpy(), last["order"]) and np.array_equal(words.numpy().view(np.uint16), last["words"])
    assert np.array_equal(smask.numpy(), last["sample_mask"]) and tuple(dist.shape) == (B, 32, 32)
    codes = P.vqvae_encode_top(gen_fs, handles["vqvae"])
    assert codes.dtype == torch.int64 and torch.equal(codes, last["codes"])
    sampled = P.lmconv_sample(codes, order, words, smask, uniforms.cuda(), 0.7, handles["lmconv"])
    sm = smask.cuda()
    assert torch.equal(sampled[~sm], codes[~sm]) and (sampled[sm] != codes[sm]).float().mean() > 0.9
    ar = P.vqvae_decode_code(sampled, handles["vqvae"])
    comb = P.combine(gen_fs, ar, bg)
    assert torch.equal(comb, torch.where(bg[:, None], ar, gen_fs))
    out = P.refine_decode(comb, bg, noise.cuda(), handles["refine"])
    assert torch.equal(out, ref["PredImg"])

    lg = P.lmconv_logits(sampled, order, words, handles["lmconv"])
    assert tuple(lg.shape) == (B, 512, 32, 32) and torch.isfinite(lg).all()

    with pytest.raises(RuntimeError):   # a handle of the wrong kind
        P.depth_unet(img, handles["vqvae"], 0.5, 10.0)
    with pytest.raises(RuntimeError):   # released / unknown handle
        P.vqvae_encode_top(img, 10 ** 6)
    with pytest.raises(RuntimeError):   # reference asserts 4-D NCHW input
        P.refine_decode(comb[:, :2], bg, None, handles["refine"])


@pytest.mark.parametrize("setting", ["gen_img", "gen_scene"])
def test_demo_command_line(tmp_path, golden_dir, setting):
    """pixelsynth_b200.demo with the reference's flags (scripts/demo_image.sh / demo_scene.sh), seeded weights: the
    files demo.py writes exist, are 256x256 RGB, and the novel view differs from the input."""
    from PIL import Image

    from pixelsynth_b200 import demo

    out = str(tmp_path / "res")
    argv = ["--vqvae", "--use_fixed_testset", "--model_setting", setting, "--gpu", "0", "--demo_img_name", "demo_input.png",
            "--demo_folder", golden_dir, "--result_folder", out, "--temperature=.7", "--num_samples", "1"]
    argv += ["--direction", "L", "--rotation", ".6"] if setting == "gen_img" else ["--directions", "R", "--num_split", "1"]
    assert demo.main(argv) == 0
    names = ["input_image_.png"] + (["output_image_L_0.png", "input_fs_image_L_0.png"] if setting == "gen_img"
                                     else ["scene/output_image_R_0001.png"])
    imgs = {}
    for n in names:
        im = np.asarray(Image.open(os.path.join(out, n)))
        assert im.shape == (256, 256, 3) and im.dtype == np.uint8, n
        imgs[n] = im.astype(np.int32)
    assert np.abs(imgs[names[1]] - imgs[names[0]]).mean() > 1.0
