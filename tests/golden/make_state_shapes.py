"""Dumps parameter/buffer names and shapes of the reference's inference networks to pixelsynth_b200/data/state_shapes.json
(run in the build container only; needs /root/reference).  The seeded weight factory (oracle/weights.py) fills
these shapes, so the same state dicts load -- strictly -- into the reference's own modules."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, "/root/reference")
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))


import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _ref_import import use_reference_models  # noqa: E402  (called from __main__ only: importing this file must not rebind `models`)

class Opt(dict):
    __getattr__ = dict.get


def main():
    from models.networks.architectures import ResNetDecoder, Unet
    from models.vqvae2.vqvae import VQVAETop
    from models.lmconv.layers import PONO
    from models.lmconv.model import OurPixelCNN

    o = Opt(norm_G="sync:spectral_batch", refine_model_type="resnet_256W8UpDown3", ngf=64, predict_residual=True,
            normalize_before_residual=False)
    nets = {
        "unet": Unet(channels_in=3, channels_out=1, opt=o, num_filters=32),            # z_buffermodel.py:41-42
        "decoder": ResNetDecoder(o, channels_in=4, channels_out=3),                    # utilities.py:26-35
        "vqvae": VQVAETop(),                                                           # z_buffermodel.py:82
        "lmconv": OurPixelCNN(nr_resnet=2, nr_filters=80, input_channels=512, nr_logistic_mix=10, kernel_size=(3, 3),
                              max_dilation=2, weight_norm=False, feature_norm_op=lambda c: PONO(), dropout_prob=0,
                              conv_bias=True, conv_mask_weight=False, rematerialize=False, binarize=False),  # :62-74
    }
    import torchvision
    from models.networks.discriminators import MultiscaleDiscriminator
    od = Opt(ndf=64, norm_D="spectralinstance", output_nc=3, no_ganFeat_loss=False, isTrain=False)
    nets["netD"] = MultiscaleDiscriminator(od)                                         # gan_loss.py:125-127, discriminators.py:142
    nets["resnet18"] = torchvision.models.resnet18(num_classes=365)                    # z_buffermodel.py:88 (places365 classifier)
    out = {k: {n: [list(t.shape), str(t.dtype).replace("torch.", "")] for n, t in m.state_dict().items()}
           for k, m in nets.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "pixelsynth_b200", "data", "state_shapes.json")
    json.dump(out, open(path, "w"), indent=0, sort_keys=False)
    for k, v in out.items():
        print(k, len(v), "tensors", sum(int(torch.tensor(s[0]).prod()) if s[0] else 1 for s in v.values()), "elements")


if __name__ == "__main__":
    use_reference_models()
    main()
