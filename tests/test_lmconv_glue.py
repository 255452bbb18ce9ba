"""CPU: generation order + locally-masked-convolution masks.
  * oracle/lmconv_ref.py against tests/golden/lmconv.npz, whose orders / mask words were asserted identical to the
    reference's get_generation_order_idx('custom') + get_unfolded_masks when the fixture was made;
  * the native host glue of the product (ps_lmconv_glue_host, csrc/glue.cu) against both, incl. cv2's transform."""
import os
import sys

import numpy as np
import pytest
import torch

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(ROOT, "tests", "golden", "lmconv.npz"))


def native_glue(bgs):
    from pixelsynth_b200 import _lib

    B = bgs.shape[0]
    m = np.ascontiguousarray(bgs.numpy().astype(np.uint8))
    d = np.zeros((B, 32, 32), np.int32)
    o = np.zeros((B, 1024), np.int32)
    w = np.zeros((B, 3, 1024), np.uint16)
    sm = np.zeros((B, 32, 32), np.uint8)
    rc = _lib.lib().ps_lmconv_glue_host(m.ctypes.data, B, 256, d.ctypes.data, o.ctypes.data, w.ctypes.data, sm.ctypes.data)
    assert rc == 0
    return d, o, w, sm.astype(bool)


def test_oracle_matches_reference_fixture(fx):
    from oracle import lmconv_ref
    import make_lmconv_golden as mk

    dist, orders, words, smask = lmconv_ref.glue_from_background(mk.background_cases())
    assert np.array_equal(dist, fx["dist"]) and np.array_equal(orders, fx["orders"]) and np.array_equal(words, fx["words"])
    assert smask[0].sum() == 512 and smask[0, :, 16:].all()     # BASELINE config 3: right half of the grid


def test_native_glue_matches_reference_fixture(fx):
    import make_lmconv_golden as mk

    d, o, w, sm = native_glue(mk.background_cases())
    assert np.array_equal(d, fx["dist"])
    assert np.array_equal(o, fx["orders"][:, :, 0] * 32 + fx["orders"][:, :, 1])
    assert np.array_equal(w, fx["words"])


def test_native_glue_random_masks_and_degenerate():
    from oracle import lmconv_ref

    g = torch.Generator().manual_seed(11)
    bgs = [torch.nn.functional.interpolate(torch.rand(1, 1, 4 + i % 13, 4 + i % 13, generator=g), size=256,
                                           mode="bilinear")[0, 0] > 0.3 + 0.015 * i for i in range(24)]
    bgs += [torch.zeros(256, 256, dtype=torch.bool), torch.ones(256, 256, dtype=torch.bool)]   # nothing / everything to sample
    bgs = torch.stack(bgs)
    dist, orders, words, smask = lmconv_ref.glue_from_background(bgs)
    d, o, w, sm = native_glue(bgs)
    assert np.array_equal(d, dist) and np.array_equal(w, words) and np.array_equal(sm, smask.numpy())
    assert np.array_equal(o, orders[:, :, 0] * 32 + orders[:, :, 1])
    # every order is a permutation and type-A masks never read the centre
    assert all(sorted(o[i].tolist()) == list(range(1024)) for i in range(o.shape[0]))
    assert ((w[:, 0] >> 4) & 1 == 0).all() and ((w[:, 1] >> 4) & 1 == 1).all()


def test_lmconv_logits_oracle_matches_reference_fixture(fx):
    from oracle import lmconv_ref, weights

    sd = weights.make_state("lmconv", 0)
    codes = torch.from_numpy(fx["codes"])
    B = codes.shape[0]
    smask = torch.from_numpy(((fx["dist"][:B] * 0) == 1))  # placeholder, replaced below
    import make_lmconv_golden as mk
    _, _, words, smask = lmconv_ref.glue_from_background(mk.background_cases())
    data = torch.nn.functional.one_hot(codes, 512).permute(0, 3, 1, 2).float() * (~smask[:B])[:, None].float()
    mf = [torch.cat([lmconv_ref.masks_to_float(words[b, k]) for b in range(B)]) for k in range(3)]
    with torch.no_grad():
        out = lmconv_ref.lmconv_logits(sd, data, *mf).numpy().reshape(-1)
    np.testing.assert_allclose(out[::257][:4096], fx["logits_sample"], rtol=0, atol=2e-4 * float(fx["logits_absmax"]))
