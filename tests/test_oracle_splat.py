"""CPU: the splat oracle (oracle/splat_oracle.c) against the golden fixtures produced by the reference's
own PtsManipulator / RasterizePointsXYsBlending code (tests/golden/make_splat_golden.py), and against
itself (binned vs brute force, C vs numpy)."""
import numpy as np
import pytest

from util import golden_cases, load_golden, synthetic_view


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("splat_")[-1][:-4])
def test_oracle_matches_reference_glue(oracle, path):
    g = load_golden(path)
    W, K = g["W"], g["K"]
    radius = g["radius_px"] / W * 2.0
    B = g["depth"].shape[0]
    # stage 1-2: projection.  The reference's torch.bmm may contract to FMA / reorder the 4-term dot
    # products, so general rotations agree to 2 ulp; axis-aligned cases are bit-exact.
    pts = oracle.project(g["depth"], g["mats"], W)
    np.testing.assert_allclose(pts, g["ref_pts"], rtol=3e-7, atol=1e-6)
    # stage 3: from the reference's own points the z-buffer maps are bit-exact
    idx, zbuf, d2 = oracle.rasterize(g["ref_pts"], W, K, radius)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(zbuf, g["zbuf"])
    assert np.array_equal(d2, g["dist2"])
    # stage 4-5
    out = oracle.composite(idx, d2, g["feat"].reshape(B, -1, W * W), radius, 2, g["tau"], g["accumulation"])
    np.testing.assert_allclose(out, g["ref_gen_fs"], rtol=0, atol=2e-6)
    assert np.array_equal(oracle.bgmask(idx, g["ksize"]), g["ref_bg"])


def test_xyzs_grid_matches_reference_buffer(oracle):
    g = load_golden(golden_cases()[0])
    W = g["W"]
    # unit depth + identity matrices => project() returns the grid itself (x, -(-y), z=+1)
    B = 1
    eye = np.tile(np.eye(4, dtype=np.float32).reshape(1, 1, 16), (B, 6, 1))
    pts, xyp = oracle.project(np.ones((B, W * W), np.float32), eye, W, want_xyproj=True)
    assert np.array_equal(xyp[0], g["xyzs"][0])


@pytest.mark.parametrize("kind", ["translate", "rotate", "behind"])
def test_binned_equals_bruteforce(oracle, kind):
    depth, feat, mats = synthetic_view(2, 40, kind=kind, seed=3)
    pts = oracle.project(depth, mats, 40)
    a = oracle.rasterize(pts, 40, 24, 4.0 / 40 * 2, naive=True)
    b = oracle.rasterize(pts, 40, 24, 4.0 / 40 * 2, naive=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_c_equals_numpy_restatement(oracle):
    depth, feat, mats = synthetic_view(1, 24, kind="translate", seed=5, depth_mode="quant")
    pts = oracle.project(depth, mats, 24)
    radius = 3.0 / 24 * 2
    idx, zbuf, d2 = oracle.rasterize(pts, 24, 16, radius)
    neg = pts * np.array([-1, -1, 1], np.float32)  # z_buffer_layers.py:71-72
    nidx, nz, nd2 = oracle.np_rasterize_points(neg, 24, radius, 16)
    assert np.array_equal(idx, nidx.astype(np.int32))
    assert np.array_equal(zbuf, nz)
    assert np.array_equal(d2, nd2)


def test_upsampled_target_and_empty(oracle):
    # S != W (ResNet feature splat path, z_buffer_layers.py:57-62 uses the src width as image size)
    depth, feat, mats = synthetic_view(1, 16, kind="identity", seed=1)
    r = oracle.splat(depth, feat, mats, 16, S=32, K=8, radius_px=2.0)
    assert r["idx"].shape == (1, 32, 32, 8)
    # a cloud entirely behind the camera: everything empty, bg all ones, out zeros
    pts = np.zeros((1, 10, 3), np.float32)
    pts[..., 2] = -1
    idx, zbuf, d2 = oracle.rasterize(pts, 8, 4, 0.5)
    assert (idx == -1).all() and (zbuf == -1).all() and (d2 == -1).all()
    assert oracle.bgmask(idx, 3).all()
    # zero points
    idx, _, _ = oracle.rasterize(np.zeros((1, 0, 3), np.float32), 8, 4, 0.5)
    assert (idx == -1).all()


def _cumul():
    import os
    from util import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "cumul_w32_k16.npz"))
    return {k: g[k] for k in g.files}


def test_oracle_cumulative_cloud_matches_reference():
    """S2c pinned: the oracle's projection + prior-cloud transform + concatenation order against the reference's
    own forward_justpts_cumulative / project_pts_cumulative (z_buffer_manipulator.py:184-266) run on three chained
    views (tests/golden/make_splat_golden.py::run_cumulative)."""
    from oracle import splat_ref as oracle
    g = _cumul()
    W, K, nv = int(g["W"]), int(g["K"]), int(g["n_views"])
    radius = float(g["radius_px"]) / W * 2.0
    prior, feats, last_bg, last_out_inv = None, None, None, None
    for v in range(nv):
        f = lambda k: g["v%d_%s" % (v, k)]
        mats = oracle.pack_mats(g["K_mat"], g["Kinv_mat"], f("src_rt"), f("src_inv"), f("dst_rt"), f("dst_inv"))
        pts, xyp = oracle.project(f("depth"), mats, W, want_xyproj=True)
        feat = f("feat").reshape(1, 3, -1)
        if prior is not None:
            sel = last_bg.reshape(-1)
            pts, xyp, feat = pts[:, sel], xyp[:, :, sel], feat[:, :, sel]          # only newly outpainted pixels (:201-202,228)
            mats3 = np.stack([g["K_mat"].reshape(1, 16), f("dst_rt").reshape(1, 16), last_out_inv.reshape(1, 16)], 1)
            pts2, xyp2 = oracle.project_cloud(prior, mats3)
            pts, xyp = np.concatenate([pts, pts2], 1), np.concatenate([xyp, xyp2], 2)   # new first, prior after (:206,247)
            feat = np.concatenate([feat, feats], 2)
        # the stored cloud (pre-division xy_proj with EPS written through the view, :250-266)
        np.testing.assert_allclose(xyp, f("cloud"), rtol=3e-7, atol=1e-6)
        assert np.array_equal(feat, f("src"))
        # the oracle's own points give the same maps up to the 2-ulp projection difference ...
        idx, zbuf, d2 = oracle.rasterize(pts, W, K, radius)
        same = idx == f("idx")
        assert same.mean() > 0.999
        np.testing.assert_allclose(zbuf[same], f("zbuf")[same], rtol=3e-7, atol=1e-6)
        # ... and from the reference's own cloud (its division + flip, z_buffer_manipulator.py:258-264) they are bit-exact
        c = f("cloud")
        zs = c[:, 2]
        ref_pts = np.stack([c[:, 0] / -zs, -(c[:, 1] / -zs), -zs], 2).astype(np.float32)
        idx, zbuf, d2 = oracle.rasterize(ref_pts, W, K, radius)
        assert np.array_equal(idx, f("idx")) and np.array_equal(zbuf, f("zbuf"))
        out = oracle.composite(idx, d2, feat, radius, 2, 1.0, "alphacomposite")
        np.testing.assert_allclose(out, f("gen_fs"), rtol=0, atol=2e-6)
        assert np.array_equal(oracle.bgmask(idx, int(g["ksize"])), f("bg"))
        prior, feats, last_bg, last_out_inv = f("cloud"), f("src"), f("bg"), f("dst_inv")
