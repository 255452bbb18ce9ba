"""Seeded random initialisation of the reference's inference networks, and synthetic inputs.

No checkpoint of the reference is reachable offline (pixelsynth.pth, the VQ-VAE and lmconv
checkpoints are downloads), so smoke runs, the parity tests and the bench all run on seeded random weights of the
reference's architecture.  make_state(net, seed) fills the parameter/buffer shapes of data/state_shapes.json --
dumped from the reference's own modules by tests/golden/make_state_shapes.py -- so the result loads strictly into
models.networks.architectures.{Unet,ResNetDecoder}, models.vqvae2.vqvae.VQVAETop and models.lmconv.model.OurPixelCNN.
Values are chosen so activations neither vanish nor saturate: fan-in scaled weights, spectral-norm u/v vectors
converged by power iteration (sigma close to the true spectral norm, as after training), batch-norm statistics away
from (0, 1)."""
import json
import math
import os

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHAPES = None
_CONV_TRANSPOSE = ("dec_t.blocks.4.weight", "upsample_t.weight", "dec.blocks.4.weight", "dec.blocks.6.weight")


def shapes():
    global _SHAPES
    if _SHAPES is None:
        _SHAPES = json.load(open(os.path.join(_HERE, "data", "state_shapes.json")))
    return _SHAPES


def _power_iteration(w2d, g, iters=8):
    u = torch.randn(w2d.shape[0], generator=g)
    u = u / u.norm()
    v = None
    for _ in range(iters):
        v = w2d.t() @ u
        v = v / (v.norm() + 1e-12)
        u = w2d @ v
        u = u / (u.norm() + 1e-12)
    return u, v


def make_state(net, seed=0, gain=1.0):
    g = torch.Generator().manual_seed(1000 * seed + sum(map(ord, net)))
    sd = {}
    table = shapes()[net]
    for name, (shape, dtype) in table.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf in ("weight_u", "weight_v"):
            continue  # filled with their weight_orig
        if dtype == "int64":
            sd[name] = torch.zeros(shape, dtype=torch.int64)
            continue
        if leaf in ("weight", "weight_orig") and len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if name in _CONV_TRANSPOSE:
                fan_in = shape[0] * 4  # ConvTranspose2d (Cin,Cout,4,4), stride 2: four taps reach each output pixel
            w = torch.randn(shape, generator=g) * (gain * (2.0 / fan_in) ** 0.5)
            sd[name] = w
            if leaf == "weight_orig":
                u, v = _power_iteration(w.reshape(shape[0], -1), g)
                pre = name[:-len("weight_orig")]
                sd[pre + "weight_u"], sd[pre + "weight_v"] = u, v
        elif leaf == "weight_g":  # weight-normed nin: g = |v| row norms by default scale
            sd[name] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf == "weight_v":
            sd[name] = torch.randn(shape, generator=g) / shape[1] ** 0.5
        elif leaf == "weight":  # BatchNorm2d gamma
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif leaf == "bias":
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif leaf in ("running_mean", "stored_mean"):
            sd[name] = torch.randn(shape, generator=g) * 0.2
        elif leaf in ("running_var", "stored_var"):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif leaf == "accumulation_counter":
            sd[name] = torch.zeros(shape)
        elif leaf in ("embed", "embed_avg"):
            sd[name] = torch.randn(shape, generator=g)
        elif leaf == "cluster_size":
            sd[name] = torch.zeros(shape)
        else:
            raise KeyError(f"no fill rule for {net}.{name} {shape}")
    # weight_v of weight-normed linears was skipped above together with spectral-norm v: fill the missing ones
    for name, (shape, dtype) in table.items():
        if name not in sd:
            sd[name] = torch.randn(shape, generator=g) / max(shape[-1], 1) ** 0.5
    sd = {k: sd[k] for k in table}
    if net == "vqvae":
        _calibrate_codebook(sd, g)
    if net == "lmconv":
        # a trained prior is peaked; with unit-scale logits over 512 classes every draw would sit on a near-uniform
        # CDF where a 1e-3 logit error already moves the token.  Sharpen the output layer (logit std ~6).
        sd["nin_out.lin_a.weight_g"] = sd["nin_out.lin_a.weight_g"] * 6.0
    if net == "resnet18":
        # a trained classifier's logits have unit-ish scale; with fan-in init through 18 layers they come out at std ~50
        # and every softmax is one-hot (entropy 0 for every candidate, nothing to rank)
        sd["fc.weight"] = sd["fc.weight"] * 0.04
    if net == "unet":
        # the depth head sees sigmoid(): widen its pre-activation (std ~1.5) so predicted depth really varies.
        # weight_orig / sigma is scale invariant, so the knob is the stored u vector (sigma = u . W v).
        sd["dconv8.weight_u"] = sd["dconv8.weight_u"] / 36.0
    return sd


def synth_image(n, seed, W=256):
    """Image-like synthetic RGB in [-1,1]: a few random low-frequency sinusoids per channel plus 20% uniform noise
    (pure noise images make every 8x8 latent cell statistically identical, which collapses the VQ codes)."""
    g = torch.Generator().manual_seed(7919 * seed + 17)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, W), torch.linspace(0, 1, W), indexing="ij")
    img = torch.zeros(n, 3, W, W)
    for i in range(n):
        for c in range(3):
            acc = torch.zeros(W, W)
            for _ in range(6):
                fx, fy = (torch.rand(2, generator=g) * 10 - 5).tolist()
                ph = torch.rand(1, generator=g).item() * 2 * math.pi
                acc += torch.rand(1, generator=g).item() * torch.sin(2 * math.pi * (fx * xx + fy * yy) + ph)
            img[i, c] = acc / acc.abs().max()
    return (img * 0.8 + 0.2 * (torch.rand(n, 3, W, W, generator=g) * 2 - 1)).clamp(-1, 1)


def _vq_res(sd, p, r):
    t = F.relu(F.conv2d(r, sd[p + "conv.1.weight"], sd[p + "conv.1.bias"], padding=1))
    return F.relu(F.conv2d(t, sd[p + "conv.3.weight"], sd[p + "conv.3.bias"]) + r)


def _vq_pre_quant_cpu(sd, x):
    """fp32 CPU pass of enc_b / enc_t / quantize_conv_t (vqvae.py:98-126,280-286), used only to place the codebook."""
    c = lambda k, t, **kw: F.conv2d(t, sd[k + ".weight"], sd[k + ".bias"], **kw)
    h = F.relu(c("enc_b.blocks.0", x, stride=2, padding=1))
    h = F.relu(c("enc_b.blocks.2", h, stride=2, padding=1))
    h = F.relu(c("enc_b.blocks.4", h, padding=1))
    h = _vq_res(sd, "enc_b.blocks.6.", _vq_res(sd, "enc_b.blocks.5.", h))
    h = F.relu(c("enc_t.blocks.0", h, stride=2, padding=1))
    h = F.relu(c("enc_t.blocks.2", h, padding=1))
    h = _vq_res(sd, "enc_t.blocks.4.", _vq_res(sd, "enc_t.blocks.3.", h))
    return c("quantize_conv_t", h)


def _vq_decode_cpu(sd, ids):
    ct = lambda k, t: F.conv_transpose2d(t, sd[k + ".weight"], sd[k + ".bias"], stride=2, padding=1)
    h = ct("upsample_t", F.embedding(ids, sd["quantize_t.embed"].t()).permute(0, 3, 1, 2))
    h = F.relu(F.conv2d(h, sd["dec.blocks.0.weight"], sd["dec.blocks.0.bias"], padding=1))
    h = _vq_res(sd, "dec.blocks.2.", _vq_res(sd, "dec.blocks.1.", h))
    return ct("dec.blocks.6", F.relu(ct("dec.blocks.4", h)))


def _calibrate_codebook(sd, g):
    """A trained codebook tiles the encoder's output distribution; emulate that by drawing the 512 codes around
    the per-channel statistics of the pre-quantisation tensor on a fixed calibration image."""
    with torch.no_grad():
        z = _vq_pre_quant_cpu(sd, synth_image(2, 12345))
    mu, sigma = z.mean(dim=(0, 2, 3)), z.std(dim=(0, 2, 3))
    e = mu[:, None] + sigma[:, None] * torch.randn(sd["quantize_t.embed"].shape, generator=g) * 0.7
    sd["quantize_t.embed"] = e
    sd["quantize_t.embed_avg"] = e.clone()
    with torch.no_grad():  # scale the last decoder layer so decoded images have image-like amplitude (std 0.5)
        flat = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        d = flat.pow(2).sum(1, keepdim=True) - 2 * flat @ e + e.pow(2).sum(0, keepdim=True)
        ids = (-d).max(1)[1].view(z.shape[0], z.shape[2], z.shape[3])
        k = 0.5 / _vq_decode_cpu(sd, ids).std().item()
    sd["dec.blocks.6.weight"] = sd["dec.blocks.6.weight"] * k
    sd["dec.blocks.6.bias"] = sd["dec.blocks.6.bias"] * k
