"""CPU: oracle/nets_ref.py (functional fp32 restatement of Unet / VQVAETop / ResNetDecoder) against fixtures written
by tests/golden/make_nets_golden.py, which ran the reference's own modules on the same seeded weights and inputs
(and asserted max|ref - oracle| <= 1e-4 there).  Also the weight factory's determinism."""
import os

import numpy as np
import pytest
import torch

from util import ROOT


def load(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"nets_{name}.npz"))


def check(out, fx, tol=2e-4):
    a = out.detach().numpy().reshape(-1)
    assert list(out.shape) == list(fx["shape"])
    s = a[::max(1, a.size // 4096)][:4096]
    scale = float(fx["absmax"])
    np.testing.assert_allclose(s, fx["sample"], rtol=0, atol=tol * scale)
    assert abs(a.mean() - float(fx["mean"])) <= tol * scale
    assert abs(a.std() - float(fx["std"])) <= tol * scale


@pytest.fixture(scope="module")
def mods():
    from oracle import nets_ref, weights

    return nets_ref, weights


def test_factory_is_deterministic_and_complete(mods):
    _, weights = mods
    for net in ("unet", "decoder", "vqvae", "lmconv", "netD", "resnet18"):
        a, b = weights.make_state(net, 0), weights.make_state(net, 0)
        assert list(a) == list(weights.shapes()[net])
        assert all(torch.equal(a[k], b[k]) for k in a)
        assert all(list(a[k].shape) == weights.shapes()[net][k][0] for k in a)
    assert not torch.equal(weights.make_state("unet", 0)["conv1.bias"], weights.make_state("unet", 1)["conv1.bias"])


def test_unet_matches_reference_fixture(mods):
    nets_ref, weights = mods
    with torch.no_grad():
        out = nets_ref.unet_features(weights.make_state("unet", 0), weights.synth_image(1, 0))
    check(out, load("unet"))
    assert out.std() > 0.5  # the depth head is not degenerate


def test_vqvae_matches_reference_fixture(mods):
    nets_ref, weights = mods
    sd = weights.make_state("vqvae", 0)
    fx = load("vqvae")
    with torch.no_grad():
        ids, _ = nets_ref.vqvae_encode_top(sd, weights.synth_image(1, 0))
        assert np.array_equal(ids.numpy(), fx["ids"])          # config 1 of BASELINE.json: id_t exact on CPU
        assert len(np.unique(fx["ids"])) > 100
        check(nets_ref.vqvae_decode_code(sd, ids), fx)


def test_decoder_matches_reference_fixture(mods):
    nets_ref, weights = mods
    sd = weights.make_state("decoder", 0)
    g = torch.Generator().manual_seed(0)
    xs = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    bg = torch.rand(1, 64, 64, generator=g) < 0.3
    noise = [torch.randn(1, 20, generator=g) for _ in range(16)]
    with torch.no_grad():
        check(nets_ref.decoder_forward(sd, xs, bg, noise), load("decoder"))


def test_discriminator_matches_reference_fixture(mods):
    """the multiscale PatchGAN's last feature maps and D_Fake, against the reference's own
    MultiscaleDiscriminator + GANLoss run by make_nets_golden.py on the same seeded weights."""
    nets_ref, weights = mods
    sd = weights.make_state("netD", 0)
    fx = load("netD")
    with torch.no_grad():
        outs = nets_ref.discriminator_forward(sd, weights.synth_image(2, 4))
    assert [tuple(o.shape) for o in outs] == [(2, 1, 35, 35), (2, 1, 19, 19)]
    check(torch.cat([o.reshape(-1) for o in outs]), fx)
    assert abs(float(nets_ref.d_fake(outs)) - float(fx["d_fake"])) <= 1e-5


def test_classifier_matches_torchvision_fixture(mods):
    """The places365 classifier (torchvision resnet18, 365 classes) on the reference's scrambled, PIL-resized input
    (z_buffermodel.py:256-261), and the entropy that ranks the candidates."""
    nets_ref, weights = mods
    sd = weights.make_state("resnet18", 0)
    fx = load("resnet18")
    with torch.no_grad():
        xc = nets_ref.classifier_input(weights.synth_image(2, 4)[0])
        assert tuple(xc.shape) == (1, 3, 224, 224)
        out = nets_ref.resnet18_logits(sd, xc)
    check(out, fx)
    e = float(nets_ref.entropy(out))
    assert abs(e - float(fx["entropy"])) <= 1e-4 and 1.0 < e < np.log(365)   # a distribution that can rank candidates
