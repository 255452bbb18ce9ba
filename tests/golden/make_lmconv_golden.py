"""Pins oracle/lmconv_ref.py to the REFERENCE's own lmconv code (build container only; needs /root/reference):
  * OurPixelCNN.forward (models/lmconv/model.py) with the seeded weights of oracle/weights.py loaded strictly,
  * get_generation_order_idx('custom') = get_custom_order.pyx (executed as Python: the shipped .so is cpython-37)
    and get_unfolded_masks (models/lmconv/masking.py) on distance maps produced the way get_masks_for_batch does.
Writes tests/golden/lmconv.npz (orders, mask words, a strided sample of the reference logits)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import lmconv_ref, weights  # noqa: E402


import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _ref_import import use_reference_models  # noqa: E402  (called from __main__ only: importing this file must not rebind `models`)

def install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    gco = types.ModuleType("models.lmconv.get_custom_order")
    src = open("/root/reference/models/lmconv/get_custom_order.pyx").read()
    exec(compile(src, "get_custom_order.pyx", "exec"), gco.__dict__)
    sys.modules["models.lmconv.get_custom_order"] = gco
    import models.lmconv

    models.lmconv.get_custom_order = gco


def background_cases():
    bgs = torch.zeros(4, 256, 256, dtype=torch.bool)
    bgs[0, :, 128:] = True                      # right half (BASELINE config 3)
    bgs[1, :60, :] = True
    bgs[1, :, 200:] = True                      # L-shaped band
    yy, xx = torch.meshgrid(torch.arange(256), torch.arange(256), indexing="ij")
    bgs[2] = ((yy - 100) ** 2 + (xx - 180) ** 2) < 70 ** 2   # a hole
    g = torch.Generator().manual_seed(3)
    bgs[3] = torch.nn.functional.interpolate(torch.rand(1, 1, 8, 8, generator=g), size=256, mode="bilinear")[0, 0] > 0.55
    return bgs


def main():
    install_stubs()
    from models.lmconv import masking
    from models.lmconv.layers import PONO
    from models.lmconv.model import OurPixelCNN

    bgs = background_cases()
    dist, orders, words, smask = lmconv_ref.glue_from_background(bgs)
    for i in range(bgs.shape[0]):
        ref_order = masking.get_generation_order_idx("custom", 32, 32, dist[i].copy(), np.array([0, 0]))
        assert np.array_equal(np.asarray(ref_order), orders[i]), f"order {i}"
        for k, (dil, mt) in enumerate(((1, "A"), (1, "B"), (2, "B"))):
            ref_m = masking.get_unfolded_masks(ref_order, 32, 32, k=3, dilation=dil, mask_type=mt)
            assert torch.equal(ref_m, lmconv_ref.masks_to_float(words[i, k])), f"mask {i} {k}"
    print("orders and masks identical to the reference for", bgs.shape[0], "backgrounds; sampled cells:",
          smask.flatten(1).sum(1).tolist())

    sd = weights.make_state("lmconv", 0)
    m = OurPixelCNN(nr_resnet=2, nr_filters=80, input_channels=512, nr_logistic_mix=10, kernel_size=(3, 3),
                    max_dilation=2, weight_norm=False, feature_norm_op=lambda c: PONO(), dropout_prob=0,
                    conv_bias=True, conv_mask_weight=False, rematerialize=False, binarize=False).eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(0)
    B = 2
    codes = torch.randint(0, 512, (B, 32, 32), generator=g)
    data = torch.nn.functional.one_hot(codes, 512).permute(0, 3, 1, 2).float()
    data = data * (~smask[:B])[:, None].float()      # sampled cells zeroed, as sample.py:47
    mf = [torch.cat([lmconv_ref.masks_to_float(words[b, k]) for b in range(B)]) for k in range(3)]
    rep = [mf[0].repeat_interleave(513, 0), mf[1].repeat_interleave(160, 0), mf[2].repeat_interleave(80, 0)]
    with torch.no_grad():
        ref = m([data, rep[0], rep[1], rep[2]], sample=True)
        mine = lmconv_ref.lmconv_logits(sd, data, *mf)
    err = (ref - mine).abs().max().item()
    print("lmconv logits max|ref-oracle| =", err, "logit std", ref.std().item())
    assert err <= 2e-4 * ref.abs().max().item()
    flat = ref.numpy().reshape(-1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lmconv.npz"), dist=dist, orders=orders, words=words,
                        codes=codes.numpy(), logits_sample=flat[::257][:4096].copy(), logits_absmax=np.abs(flat).max(),
                        logits_std=flat.std())


if __name__ == "__main__":
    use_reference_models()
    main()
