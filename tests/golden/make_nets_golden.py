"""Pins oracle/nets_ref.py to the REFERENCE's own modules (run in the build container; needs /root/reference).

For each network the seeded state dict of oracle/weights.py is loaded -- strict=True -- into the unmodified
reference class, both are run on the same seeded CPU inputs, agreement is asserted, and a compact fixture
(tests/golden/nets_<name>.npz: inputs' seeds, output statistics and a strided sample of the output) is written for
the CPU tests that run where the reference is absent."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import nets_ref, weights  # noqa: E402

import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _ref_import import use_reference_models  # noqa: E402  (called from __main__ only: importing this file must not rebind `models`)

OUT = os.path.dirname(os.path.abspath(__file__))


class Opt(dict):
    __getattr__ = dict.get


def fixture(name, ref_out, **extra):
    a = ref_out.detach().numpy()
    flat = a.reshape(-1)
    np.savez_compressed(os.path.join(OUT, f"nets_{name}.npz"), shape=np.array(a.shape), mean=flat.mean(), std=flat.std(),
                        absmax=np.abs(flat).max(), sample=flat[::max(1, flat.size // 4096)][:4096].copy(), **extra)


def main():
    torch.manual_seed(0)
    o = Opt(norm_G="sync:spectral_batch", refine_model_type="resnet_256W8UpDown3", ngf=64, predict_residual=True,
            normalize_before_residual=False)
    from models.networks.architectures import ResNetDecoder, Unet
    from models.vqvae2.vqvae import VQVAETop

    g = torch.Generator().manual_seed(0)
    x = weights.synth_image(1, 0)
    with torch.no_grad():
        # ---- Unet ----
        sd = weights.make_state("unet", 0)
        m = Unet(channels_in=3, channels_out=1, opt=o, num_filters=32).eval()
        m.load_state_dict(sd, strict=True)
        ref = m(x)
        mine = nets_ref.unet_features(sd, x)
        err = (ref - mine).abs().max().item()
        print("unet  max|ref-oracle| =", err, "out std", ref.std().item())
        assert err <= 1e-4 * ref.abs().max().item()
        fixture("unet", ref)

        # ---- VQ-VAE ----
        sd = weights.make_state("vqvae", 0)
        m = VQVAETop().eval()
        m.load_state_dict(sd, strict=True)
        ids_ref = m.encode(x.clone())[3]
        ids, z = nets_ref.vqvae_encode_top(sd, x)
        print("vqvae id mismatches:", (ids != ids_ref).sum().item(), "of", ids.numel(), "distinct codes", ids.unique().numel())
        assert (ids != ids_ref).sum().item() == 0
        dec_ref = m.decode_code(ids_ref)
        dec = nets_ref.vqvae_decode_code(sd, ids)
        err = (dec_ref - dec).abs().max().item()
        print("vqvae decode max err", err, "std", dec_ref.std().item())
        assert err <= 1e-4 * dec_ref.abs().max().item()
        fixture("vqvae", dec_ref, ids=ids_ref.numpy())

        # ---- ResNetDecoder (noise injected by patching torch.randn for the reference's LinearNoiseLayer) ----
        sd = weights.make_state("decoder", 0)
        m = ResNetDecoder(o, channels_in=4, channels_out=3).eval()
        m.load_state_dict(sd, strict=True)
        xs = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1   # fully convolutional: 64x64 keeps the fixture run short
        bg = torch.rand(1, 64, 64, generator=g) < 0.3
        noise = [torch.randn(1, 20, generator=g) for _ in range(16)]
        queue = list(noise)
        real_randn = torch.randn
        torch.randn = lambda *a, **k: queue.pop(0)
        try:
            ref = m(xs, bg)
        finally:
            torch.randn = real_randn
        assert not queue
        mine = nets_ref.decoder_forward(sd, xs, bg, noise)
        err = (ref - mine).abs().max().item()
        print("decoder max err", err, "pre-tanh saturation: frac |out|>0.99 =", (ref.abs() > 0.99).float().mean().item())
        assert err <= 1e-4
        fixture("decoder", ref)

        # ---- sample ranking networks ----
        from models.networks.discriminators import MultiscaleDiscriminator
        import torchvision
        od = Opt(ndf=64, norm_D="spectralinstance", output_nc=3, no_ganFeat_loss=False, isTrain=False)
        sd = weights.make_state("netD", 0)
        m = MultiscaleDiscriminator(od).eval()
        m.load_state_dict(sd, strict=True)
        xs = weights.synth_image(2, 4)
        ref = m(xs)                                   # list (scale) of lists (layer outputs)
        mine = nets_ref.discriminator_forward(sd, xs)
        for r, o_ in zip(ref, mine):
            err = (r[-1] - o_).abs().max().item()
            print("netD last map", tuple(o_.shape), "max err", err, "std", r[-1].std().item())
            assert err <= 1e-4 * max(1.0, r[-1].abs().max().item())
        # D_Fake through the reference's own GANLoss (gan_loss.py:101-116, hinge)
        from models.losses.gan_loss import GANLoss
        crit = GANLoss("hinge", tensor=torch.FloatTensor, opt=od)
        dref = crit(ref, False, for_discriminator=True)
        dmine = nets_ref.d_fake(mine)
        print("D_Fake", float(dref), float(dmine))
        assert abs(float(dref) - float(dmine)) <= 1e-5
        fixture("netD", torch.cat([r[-1].reshape(-1) for r in ref]), d_fake=float(dref))

        sd = weights.make_state("resnet18", 0)
        m = torchvision.models.resnet18(num_classes=365).eval()
        m.load_state_dict(sd, strict=True)
        xc = nets_ref.classifier_input(xs[0])
        ref = m(xc)
        mine = nets_ref.resnet18_logits(sd, xc)
        err = (ref - mine).abs().max().item()
        print("resnet18 logits max err", err, "std", ref.std().item(), "entropy", float(nets_ref.entropy(ref)))
        assert err <= 1e-4 * max(1.0, ref.abs().max().item())
        fixture("resnet18", ref, entropy=float(nets_ref.entropy(ref)))


if __name__ == "__main__":
    use_reference_models()
    main()
