"""GPU parity: the sm_100a splat (through torch.ops -> C ABI) against the CPU oracle.

Bar: projected points, idx, zbuf, dist2 and the background mask bit-exact; composited features within
2e-6 absolute (fp32; the kernel sums the K terms with a warp-shuffle tree, the oracle front to back)."""
import types

import numpy as np
import pytest
import torch

from util import golden_cases, load_golden, synthetic_view

pytestmark = pytest.mark.gpu
ATOL_OUT = 2e-6


@pytest.fixture(scope="module")
def ops():
    import pixelsynth_b200.ops as o  # registers torch.ops.pixelsynth_b200.*

    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_splat(depth, feat, mats, W, S, K, radius_px, tau=1.0, acc=0, ksize=13, maps=True):
    r = torch.ops.pixelsynth_b200.splat(dev(depth), dev(feat), dev(mats), W, S, K, radius_px, tau, 2, acc, ksize, 1e-2,
                                        maps, maps)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in r]


def check_against_oracle(oracle, depth, feat, mats, W, S, K, radius_px, tau=1.0, accumulation="alphacomposite",
                         ksize=13):
    from pixelsynth_b200.ops import ACCUMULATION
    ref = oracle.splat(depth, feat, mats, W, S=S, K=K, radius_px=radius_px, tau=tau, bg_ksize=ksize,
                       accumulation=accumulation)
    out, bg, idx, zbuf, d2 = run_splat(depth, feat, mats, W, S, K, radius_px, tau, ACCUMULATION[accumulation], ksize)
    assert np.array_equal(idx, ref["idx"]), f"idx mismatches: {(idx != ref['idx']).sum()}"
    assert np.array_equal(zbuf, ref["zbuf"])
    assert np.array_equal(d2, ref["dist2"])
    assert np.array_equal(bg, ref["bg"])
    np.testing.assert_allclose(out, ref["out"], rtol=0, atol=ATOL_OUT)
    # production mode (maps suppressed) gives the same image and mask
    out2, bg2, *_ = run_splat(depth, feat, mats, W, S, K, radius_px, tau, ACCUMULATION[accumulation], ksize, maps=False)
    assert np.array_equal(out2, out) and np.array_equal(bg2, bg)
    return ref


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("splat_")[-1][:-4])
def test_golden_fixtures(ops, oracle, path):
    g = load_golden(path)
    W, K = g["W"], g["K"]
    B = g["depth"].shape[0]
    acc = ops.ACCUMULATION[g["accumulation"]]
    # projection stage against the oracle (bit-exact) and the reference's torch.bmm (2 ulp)
    pts, xyp = torch.ops.pixelsynth_b200.project_pts(dev(g["depth"]), dev(g["mats"]), W, 1e-2, True)
    opts, oxyp = oracle.project(g["depth"], g["mats"], W, want_xyproj=True)
    assert np.array_equal(pts.cpu().numpy(), opts)
    assert np.array_equal(xyp.cpu().numpy(), oxyp)
    np.testing.assert_allclose(pts.cpu().numpy(), g["ref_pts"], rtol=3e-7, atol=1e-6)
    # rasterise + composite from the reference's own points: maps bit-exact vs the fixture
    feat = g["feat"].reshape(B, -1, W * W)
    out, bg, idx, zbuf, d2 = torch.ops.pixelsynth_b200.splat_points(
        dev(g["ref_pts"]), dev(feat), W, K, g["radius_px"], g["tau"], 2, acc, g["ksize"], True, True)
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    assert np.array_equal(zbuf.cpu().numpy(), g["zbuf"])
    assert np.array_equal(d2.cpu().numpy(), g["dist2"])
    assert np.array_equal(bg.cpu().numpy(), g["ref_bg"])
    np.testing.assert_allclose(out.cpu().numpy(), g["ref_gen_fs"], rtol=0, atol=ATOL_OUT)


@pytest.mark.parametrize("kind,depth_mode,W,K,radius", [
    ("translate", "uniform", 64, 128, 4.0),
    ("rotate", "uniform", 64, 32, 4.0),
    ("behind", "uniform", 64, 16, 2.0),
    ("identity", "const", 48, 128, 4.0),     # pure tie-break ordering, 49+ hits per pixel
    ("translate", "quant", 80, 8, 4.0),      # K much smaller than hits: selection, many z ties
    ("translate", "smooth", 96, 128, 4.0),
    ("translate", "uniform", 50, 5, 1.5),    # ragged: S not a multiple of the tile, K not a multiple of 4
    ("rotate", "smooth", 37, 33, 3.0),
])
def test_seeded_parity(ops, oracle, kind, depth_mode, W, K, radius):
    depth, feat, mats = synthetic_view(2, W, kind=kind, seed=W + K, depth_mode=depth_mode)
    check_against_oracle(oracle, depth, feat, mats, W, W, K, radius)


@pytest.mark.parametrize("accumulation,tau", [("wsum", 1.0), ("wsumnorm", 2.0), ("alphacomposite", 0.5)])
def test_accumulation_modes(ops, oracle, accumulation, tau):
    depth, feat, mats = synthetic_view(2, 48, kind="translate", seed=11)
    check_against_oracle(oracle, depth, feat, mats, 48, 48, 24, 3.0, tau=tau, accumulation=accumulation, ksize=5)


def test_many_feature_channels(ops, oracle):
    # C = 64 is the reference's non-rgb feature width (PtsManipulator C=64): generic-C path
    depth, feat, mats = synthetic_view(1, 32, C=64, kind="translate", seed=2)
    check_against_oracle(oracle, depth, feat, mats, 32, 32, 16, 4.0)


def test_overflow_tiles_dense_cloud(ops, oracle):
    # a whole cloud collapsed onto few pixels: every tile list overflows the shared-memory fast path
    rng = np.random.default_rng(0)
    P = 6000
    pts = np.empty((1, P, 3), np.float32)
    pts[0, :, :2] = rng.uniform(-0.2, 0.2, (P, 2))
    pts[0, :, 2] = np.round(rng.uniform(0.5, 3.0, P) * 8) / 8
    feat = rng.uniform(-1, 1, (1, 3, P)).astype(np.float32)
    S, K, rp = 32, 128, 4.0
    radius = rp / S * 2.0
    idx, zbuf, d2 = oracle.rasterize(pts, S, K, radius)
    ref_out = oracle.composite(idx, d2, feat, radius)
    out, bg, gi, gz, gd = torch.ops.pixelsynth_b200.splat_points(dev(pts), dev(feat), S, K, rp, 1.0, 2, 0, 13, True, True)
    assert np.array_equal(gi.cpu().numpy(), idx)
    assert np.array_equal(gz.cpu().numpy(), zbuf)
    assert np.array_equal(gd.cpu().numpy(), d2)
    np.testing.assert_allclose(out.cpu().numpy(), ref_out, rtol=0, atol=ATOL_OUT)
    assert np.array_equal(bg.cpu().numpy(), oracle.bgmask(idx, 13))


def test_empty_inputs(ops):
    z = torch.zeros((1, 0, 3), device="cuda")
    f = torch.zeros((1, 3, 0), device="cuda")
    out, bg, idx, zbuf, d2 = torch.ops.pixelsynth_b200.splat_points(z, f, 16, 8, 2.0, 1.0, 2, 0, 13, True, True)
    assert (out == 0).all() and bg.all() and (idx == -1).all() and (zbuf == -1).all() and (d2 == -1).all()
    # every point behind the camera / NaN
    p = torch.tensor([[[0.0, 0.0, -1.0], [float("nan"), 0.0, 1.0], [0.0, 0.0, float("nan")]]], device="cuda")
    out, bg, idx, *_ = torch.ops.pixelsynth_b200.splat_points(p, torch.ones((1, 3, 3), device="cuda"), 16, 8, 2.0, 1.0,
                                                             2, 0, 13, True, False)
    assert (out == 0).all() and bg.all() and (idx == -1).all()


def test_full_size_256_k128(ops, oracle):
    """BASELINE config shape: 256x256, K=128, radius 4, circle-translation target (oracle ~1 s/view)."""
    depth, feat, mats = synthetic_view(2, 256, kind="translate", seed=0)
    ref = check_against_oracle(oracle, depth, feat, mats, 256, 256, 128, 4.0)
    assert (ref["idx"] >= 0).sum(-1).mean() > 20  # the case really exercises deep z-buffers


def test_full_size_properties(ops):
    """Size-independent properties at B=8 full size: sortedness, padding, index/feature consistency,
    permutation invariance of the input cloud (idx map relabels, z map unchanged)."""
    depth, feat, mats = synthetic_view(8, 256, kind="translate", seed=4, depth_mode="smooth")
    B, W, K = 8, 256, 128
    pts, _ = torch.ops.pixelsynth_b200.project_pts(dev(depth), dev(mats), W, 1e-2, False)
    f = dev(feat).reshape(B, 3, -1)
    out, bg, idx, zbuf, d2 = torch.ops.pixelsynth_b200.splat_points(pts, f, W, K, 4.0, 1.0, 2, 0, 13, True, True)
    valid = idx >= 0
    assert ((zbuf == -1) == ~valid).all() and ((d2 == -1) == ~valid).all()
    assert (valid[..., 1:] <= valid[..., :-1]).all()                      # padding only at the tail
    both = valid[..., 1:]
    dz = zbuf[..., 1:] - zbuf[..., :-1]
    assert (dz[both] >= 0).all()                                          # ascending z
    tie = both & (dz == 0)
    assert (idx[..., 1:][tie] > idx[..., :-1][tie]).all()                 # ties ascending packed index
    b_of = torch.arange(B, device="cuda").view(B, 1, 1, 1).expand_as(idx)
    assert ((idx // (W * W))[valid] == b_of[valid]).all()                 # packed index stays in its cloud
    z_of = pts.reshape(-1, 3)[:, 2][idx.clamp(min=0).long()]
    assert (z_of[valid] == zbuf[valid]).all()
    assert (d2[valid] < (4.0 / W * 2) ** 2).all() and (d2[valid] >= 0).all()
    # permuting the cloud relabels idx but leaves the z map, dist2 map (no ties here), image and mask unchanged
    perm = torch.randperm(W * W, device="cuda")
    out2, bg2, idx2, zbuf2, d22 = torch.ops.pixelsynth_b200.splat_points(
        pts[:, perm].contiguous(), f[:, :, perm].contiguous(), W, K, 4.0, 1.0, 2, 0, 13, True, True)
    assert (zbuf2 == zbuf).all() and (bg2 == bg).all()
    notie = ~(tie.any(-1))
    assert (out2[notie.unsqueeze(1).expand_as(out2)] - out[notie.unsqueeze(1).expand_as(out)]).abs().max() <= ATOL_OUT


def test_module_mirror_forward_justpts(ops, oracle):
    """The reference-shaped seam: PtsManipulator.forward_justpts (z_buffer_manipulator.py:85-107)."""
    from pixelsynth_b200.models.projection.z_buffer_manipulator import PtsManipulator
    from util import demo_cameras

    W, B = 64, 2
    opt = types.SimpleNamespace(splatter="xyblending", learn_default_feature=True, radius=4.0, pp_pixel=128, rad_pow=2,
                                tau=1.0, accumulation="alphacomposite", background_smoothing_kernel_size=13)
    pm = PtsManipulator(W, C=3, opt=opt).cuda()
    depth, feat, mats = synthetic_view(B, W, kind="translate", seed=9)
    cams = [dev(m) for m in demo_cameras(B, "translate", 9)]
    gen_fs, bg = pm.forward_justpts(dev(feat), dev(depth), *cams)
    ref = oracle.splat(depth, feat, mats, W, K=128, radius_px=4.0)
    assert gen_fs.shape == (B, 3, W, W) and bg.dtype == torch.bool and bg.shape == (B, W, W)
    np.testing.assert_allclose(gen_fs.cpu().numpy(), ref["out"], rtol=0, atol=ATOL_OUT)
    assert np.array_equal(bg.cpu().numpy(), ref["bg"])
    sampler = pm.project_pts(dev(depth).view(B, 1, -1), *cams)
    assert np.array_equal(sampler.permute(0, 2, 1).cpu().numpy(), ref["pts"])


def test_cumulative_cloud_two_views(ops, oracle):
    """S2c: PtsManipulator.forward_justpts_cumulative (z_buffer_manipulator.py:184-266) over two consecutive views of
    one image, as forward_scene drives it: the first view splats the source grid and returns the pre-division cloud;
    the second appends only the pixels the first view left as background (newly outpainted content) to the prior
    cloud re-expressed in the new camera.  Clouds and masks bit-exact, image within ATOL_OUT."""
    from pixelsynth_b200.models.projection.z_buffer_manipulator import PtsManipulator
    from util import demo_cameras, pack_mats

    W, B, Kpp = 64, 1, 128
    radius = 4.0 / W * 2.0
    opt = types.SimpleNamespace(splatter="xyblending", learn_default_feature=True, radius=4.0, pp_pixel=Kpp, rad_pow=2,
                                tau=1.0, accumulation="alphacomposite", background_smoothing_kernel_size=13)
    pm = PtsManipulator(W, C=3, opt=opt).cuda()
    depth1, feat1, _ = synthetic_view(B, W, kind="translate", seed=21, depth_mode="smooth")
    depth2, feat2, _ = synthetic_view(B, W, kind="translate", seed=22, depth_mode="smooth")
    K, Kinv, RT1, RT1inv, RT2, RT2inv = demo_cameras(B, "translate", 21, views=[2])
    _, _, _, _, RT3, RT3inv = demo_cameras(B, "translate", 21, views=[3])
    RT2[:, 0, 3] += 0.6  # push part of the first view out of frame so it leaves a background band
    RT2inv = np.linalg.inv(RT2).astype(np.float32)

    # ---- view 1: no prior cloud ----
    g = [dev(m) for m in (K, Kinv, RT1, RT1inv, RT2, RT2inv)]
    res1, bg1, cloud1, src1 = pm.forward_justpts_cumulative(dev(feat1), dev(depth1), *g, None, None, None, None)
    pts_a, xyp_a = oracle.project(depth1, pack_mats(K, Kinv, RT1, RT1inv, RT2, RT2inv), W, want_xyproj=True)
    idx_a, _, d2_a = oracle.rasterize(pts_a, W, Kpp, radius)
    out_a = oracle.composite(idx_a, d2_a, feat1.reshape(B, 3, -1), radius)
    bg_a = oracle.bgmask(idx_a, 13)
    assert np.array_equal(cloud1.cpu().numpy(), xyp_a)
    assert np.array_equal(bg1.cpu().numpy(), bg_a) and 0 < bg_a.sum() < bg_a.size
    np.testing.assert_allclose(res1.cpu().numpy(), out_a, rtol=0, atol=ATOL_OUT)
    assert np.array_equal(src1.cpu().numpy(), feat1.reshape(B, 3, -1))

    # ---- view 2: the "outpainted" image (feat2, depth2) seen from camera 2, moved to camera 3 ----
    g2 = [dev(m) for m in (K, Kinv, RT2, RT2inv, RT3, RT3inv)]
    res2, bg2, cloud2, src2 = pm.forward_justpts_cumulative(dev(feat2), dev(depth2), *g2, cloud1, src1, bg1, dev(RT2inv))
    sel = bg_a.reshape(B, -1)[0]
    pts_n, xyp_n = oracle.project(depth2, pack_mats(K, Kinv, RT2, RT2inv, RT3, RT3inv), W, want_xyproj=True)
    mats3 = np.ascontiguousarray(np.stack([K.reshape(B, 16), RT3.reshape(B, 16), RT2inv.reshape(B, 16)], 1))
    pts_o, xyp_o = oracle.project_cloud(xyp_a, mats3)
    pts_c = np.concatenate([pts_n[:, sel], pts_o], 1)
    xyp_c = np.concatenate([xyp_n[:, :, sel], xyp_o], 2)
    feat_c = np.concatenate([feat2.reshape(B, 3, -1)[:, :, sel], feat1.reshape(B, 3, -1)], 2)
    idx_c, _, d2_c = oracle.rasterize(pts_c, W, Kpp, radius)
    out_c = oracle.composite(idx_c, d2_c, feat_c, radius)
    assert cloud2.shape == (B, 4, int(sel.sum()) + W * W)
    assert np.array_equal(cloud2.cpu().numpy(), xyp_c)
    assert np.array_equal(src2.cpu().numpy(), feat_c)
    assert np.array_equal(bg2.cpu().numpy(), oracle.bgmask(idx_c, 13))
    np.testing.assert_allclose(res2.cpu().numpy(), out_c, rtol=0, atol=ATOL_OUT)


def test_cumulative_chain_matches_reference_fixture(golden_dir):
    """S2c on the GPU against the REFERENCE's own forward_justpts_cumulative (three chained views, fixture written by
    tests/golden/make_splat_golden.py::run_cumulative): stored cloud, concatenated features (order: newly outpainted
    pixels first, prior cloud after), image and background mask of every view."""
    import os
    import types

    from pixelsynth_b200.models.projection.z_buffer_manipulator import PtsManipulator

    g = np.load(os.path.join(golden_dir, "cumul_w32_k16.npz"))
    W, K, nv = int(g["W"]), int(g["K"]), int(g["n_views"])
    opt = types.SimpleNamespace(splatter="xyblending", learn_default_feature=True, radius=float(g["radius_px"]), pp_pixel=K,
                                rad_pow=2, tau=1.0, accumulation="alphacomposite",
                                background_smoothing_kernel_size=int(g["ksize"]))
    pm = PtsManipulator(W, C=3, opt=opt).cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    Km, Kinv = t(g["K_mat"]), t(g["Kinv_mat"])
    prior, feats, last_bg, last_out_inv = None, None, None, None
    for v in range(nv):
        f = lambda k: g["v%d_%s" % (v, k)]
        gen_fs, bg, cloud, src = pm.forward_justpts_cumulative(
            t(f("feat")), t(f("depth")), Km, Kinv, t(f("src_rt")), t(f("src_inv")), t(f("dst_rt")), t(f("dst_inv")),
            prior, feats, last_bg, last_out_inv)
        assert tuple(cloud.shape) == f("cloud").shape and tuple(src.shape) == f("src").shape
        np.testing.assert_allclose(cloud.cpu().numpy(), f("cloud"), rtol=3e-7, atol=1e-6)   # bmm contraction: 2 ulp
        assert np.array_equal(src.cpu().numpy(), f("src"))
        # the image can differ where a 2-ulp point flips a membership test; such pixels are a handful
        diff = np.abs(gen_fs.cpu().numpy() - f("gen_fs")).max(1)
        assert (diff > 2e-6).mean() < 2e-3, (diff > 2e-6).mean()
        assert (bg.cpu().numpy() != f("bg")).mean() < 2e-3
        # chain on the REFERENCE's state so one flipped pixel cannot snowball through the views
        prior, feats, last_bg, last_out_inv = t(f("cloud")), t(f("src")), t(f("bg")), t(f("dst_inv"))
