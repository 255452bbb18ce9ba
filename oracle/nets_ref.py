"""CPU fp32 restatement of the reference's dense inference networks, as pure functions of a reference-keyed
state dict (TEST INFRASTRUCTURE ONLY -- the product never imports this module).

  unet_depth          models/networks/architectures.py:174-279 (Unet) + models/z_buffermodel.py:304-308 (sigmoid range)
  decoder_forward     models/networks/architectures.py:126-167 (ResNetDecoder), models/layers/blocks.py:33-74
                      (ResNet_Block), models/layers/normalization.py:21-47,114-178 (LinearNoiseLayer, bn, fused_bn)
  vqvae_encode_top    models/vqvae2/vqvae.py:280-297 (only id_t is consumed: z_buffermodel.py:345), :41-48 (Quantize)
  vqvae_decode_code   models/vqvae2/vqvae.py:299-312
Pinned against the reference's own modules by tests/golden/make_nets_golden.py (run where /root/reference
exists), which loads the same seeded state dicts into the real classes, asserts agreement and commits fixtures.
The noise that LinearNoiseLayer draws with torch.randn in every forward (normalization.py:40) is an explicit
argument here: noise[i] is the (N,20) draw of the i-th LinearNoiseLayer in execution order."""
import torch
import torch.nn.functional as F


def sn_weight(sd, prefix):
    """torch.nn.utils.spectral_norm in eval mode: weight_orig / (u . W v), no power iteration."""
    w = sd[prefix + "weight_orig"]
    sigma = torch.dot(sd[prefix + "weight_u"], w.reshape(w.shape[0], -1) @ sd[prefix + "weight_v"])
    return w / sigma


def _bn_eval(sd, prefix, x, eps=1e-5):
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                        sd[prefix + "bias"], False, 0.0, eps)


def unet_features(sd, x):
    """Unet.forward: 8 stride-2 4x4 convs down to 1x1, 8 bilinear-upsample + 3x3 convs back up."""
    def down(i, t):
        return F.conv2d(t, sn_weight(sd, f"conv{i}."), sd[f"conv{i}.bias"], stride=2, padding=1)

    def up(i, t):
        t = F.interpolate(F.relu(t), scale_factor=2, mode="bilinear", align_corners=False)
        return F.conv2d(t, sn_weight(sd, f"dconv{i}."), sd[f"dconv{i}.bias"], padding=1)

    enc_bn = [None, None, "batch_norm2_0.", "batch_norm4_0.", "batch_norm8_0.", "batch_norm8_1.", "batch_norm8_2.",
              "batch_norm8_3.", None]
    dec_bn = [None, "batch_norm8_4.", "batch_norm8_5.", "batch_norm8_6.", "batch_norm8_7.", "batch_norm4_1.",
              "batch_norm2_1.", "batch_norm.", None]
    e = [None, down(1, x)]
    for i in range(2, 9):
        t = down(i, F.leaky_relu(e[-1], 0.2))
        e.append(_bn_eval(sd, enc_bn[i], t) if enc_bn[i] else t)
    d = e[8]
    for i in range(1, 9):
        t = up(i, d)
        if dec_bn[i]:
            t = _bn_eval(sd, dec_bn[i], t)
        d = torch.cat((t, e[8 - i]), 1) if i < 8 else t
    return d


def unet_depth(sd, x, min_z, max_z):
    return torch.sigmoid(unet_features(sd, x)) * (max_z - min_z) + min_z


DECODER_CHANNELS = [4, 64, 128, 256, 256, 128, 128, 128, 3]          # configs.py:221-231 with ngf = 64
DECODER_RESAMPLE = [None, "Down", "Down", None, "Up", "Up", None, None]  # configs.py:232-241


def _noise_bn(sd, prefix, x, z, eps=1e-5):
    gain = 1 + z @ sn_weight(sd, prefix + "gain.").t()
    bias = z @ sn_weight(sd, prefix + "bias.").t()
    scale = torch.rsqrt(sd[prefix + "bn.stored_var"] + eps)[None, :] * gain
    shift = sd[prefix + "bn.stored_mean"][None, :] * scale - bias
    return x * scale[:, :, None, None] - shift[:, :, None, None]


def _resample(kind, t):
    if kind == "Down":
        return F.avg_pool2d(t, 3, 2, 1)
    if kind == "Up":
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    return t


def decoder_forward(sd, x, background_mask, noise, normalize_before_residual=False):
    """ResNetDecoder.forward with predict_residual: tanh(eblocks(cat(x, ~bg)) + x)."""
    h = torch.cat((x, (~background_mask).unsqueeze(1).float()), 1)
    zi = 0
    for b in range(8):
        p = f"eblocks.{b}."
        cin, cout, kind = DECODER_CHANNELS[b], DECODER_CHANNELS[b + 1], DECODER_RESAMPLE[b]
        a = F.relu(_noise_bn(sd, p + "ch_a.0.", h, noise[zi]))
        a = F.conv2d(a, sn_weight(sd, p + "ch_a.2."), sd[p + "ch_a.2.bias"], padding=1)
        a = F.relu(_noise_bn(sd, p + "ch_a.3.", a, noise[zi + 1]))
        a = F.conv2d(a, sn_weight(sd, p + "ch_a.5."), sd[p + "ch_a.5.bias"], padding=1)
        zi += 2
        a = _resample(kind, a)
        if kind or cin != cout:
            s = _resample(kind, F.conv2d(h, sn_weight(sd, p + "ch_b.0."), sd[p + "ch_b.0.bias"]))
        else:
            s = h
        h = a + s
    return torch.tanh(h) + x if normalize_before_residual else torch.tanh(h + x)


def _vq_resblock(sd, prefix, r):
    """ResBlock with its in-place first ReLU: the skip adds relu(input) (vqvae.py:84-95); r is already relu'd."""
    t = F.relu(F.conv2d(r, sd[prefix + "conv.1.weight"], sd[prefix + "conv.1.bias"], padding=1))
    return F.conv2d(t, sd[prefix + "conv.3.weight"], sd[prefix + "conv.3.bias"]) + r


def vqvae_pre_quant(sd, x):
    """enc_b, enc_t and quantize_conv_t: the (N,64,32,32) tensor whose nearest codes are id_t."""
    h = F.relu(F.conv2d(x, sd["enc_b.blocks.0.weight"], sd["enc_b.blocks.0.bias"], stride=2, padding=1))
    h = F.relu(F.conv2d(h, sd["enc_b.blocks.2.weight"], sd["enc_b.blocks.2.bias"], stride=2, padding=1))
    h = F.relu(F.conv2d(h, sd["enc_b.blocks.4.weight"], sd["enc_b.blocks.4.bias"], padding=1))
    h = F.relu(_vq_resblock(sd, "enc_b.blocks.5.", h))
    h = F.relu(_vq_resblock(sd, "enc_b.blocks.6.", h))
    h = F.relu(F.conv2d(h, sd["enc_t.blocks.0.weight"], sd["enc_t.blocks.0.bias"], stride=2, padding=1))
    h = F.relu(F.conv2d(h, sd["enc_t.blocks.2.weight"], sd["enc_t.blocks.2.bias"], padding=1))
    h = F.relu(_vq_resblock(sd, "enc_t.blocks.3.", h))
    h = F.relu(_vq_resblock(sd, "enc_t.blocks.4.", h))
    return F.conv2d(h, sd["quantize_conv_t.weight"], sd["quantize_conv_t.bias"])


def vq_distances(z, embed):
    flat = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
    return flat.pow(2).sum(1, keepdim=True) - 2 * flat @ embed + embed.pow(2).sum(0, keepdim=True)


def vqvae_encode_top(sd, x):
    z = vqvae_pre_quant(sd, x)
    ids = (-vq_distances(z, sd["quantize_t.embed"])).max(1)[1]
    return ids.view(z.shape[0], z.shape[2], z.shape[3]), z


def vqvae_decode_code(sd, ids):
    q = F.embedding(ids, sd["quantize_t.embed"].t()).permute(0, 3, 1, 2)
    h = F.conv_transpose2d(q, sd["upsample_t.weight"], sd["upsample_t.bias"], stride=2, padding=1)
    h = F.relu(F.conv2d(h, sd["dec.blocks.0.weight"], sd["dec.blocks.0.bias"], padding=1))
    h = F.relu(_vq_resblock(sd, "dec.blocks.1.", h))
    h = F.relu(_vq_resblock(sd, "dec.blocks.2.", h))
    h = F.relu(F.conv_transpose2d(h, sd["dec.blocks.4.weight"], sd["dec.blocks.4.bias"], stride=2, padding=1))
    return F.conv_transpose2d(h, sd["dec.blocks.6.weight"], sd["dec.blocks.6.bias"], stride=2, padding=1)


# ---------------------------------------------------------------------------------------------------------------
# Sample ranking networks: the multiscale PatchGAN discriminator and the places365 classifier
# ---------------------------------------------------------------------------------------------------------------
def _instance_norm(x, eps=1e-5):
    return F.instance_norm(x, eps=eps)


def discriminator_forward(sd, x):
    """MultiscaleDiscriminator.forward (models/networks/discriminators.py:196-207) of two NLayerDiscriminators
    (:78-140: n_layers_D = 4, ndf = 64, norm_D = spectralinstance, kernel 4, padding 2): returns the LAST feature map
    of each scale -- all that D_Fake reads (models/losses/gan_loss.py:104-116).  The second scale sees
    avg_pool2d(3, 2, 1, count_include_pad=False) of the input (:170-177)."""
    outs = []
    for d in range(2):
        p = f"discriminator_{d}."
        h = F.leaky_relu(F.conv2d(x, sd[p + "model0.0.weight"], sd[p + "model0.0.bias"], stride=2, padding=2), 0.2)
        for n, stride in ((1, 2), (2, 2), (3, 1)):
            h = F.conv2d(h, sn_weight(sd, p + f"model{n}.0.0."), None, stride=stride, padding=2)  # bias removed (normalization.py:69-73)
            h = F.leaky_relu(_instance_norm(h), 0.2)
        outs.append(F.conv2d(h, sd[p + "model4.0.weight"], sd[p + "model4.0.bias"], stride=1, padding=2))
        x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
    return outs


def d_fake(preds):
    """D_Fake (gan_loss.py:172-181 -> GANLoss.__call__ :101-116 -> hinge, fake, for the discriminator :80-88):
    mean over scales of -mean(min(-x - 1, 0)) over the whole fake batch."""
    return sum(-torch.mean(torch.min(-x - 1, torch.zeros_like(x))) for x in preds) / len(preds)


def classifier_input(gen_img0):
    """z_buffermodel.py:256-257 with :105-110: the (3,256,256) image is RESHAPED (not permuted) to (256,256,3) --
    a reinterpretation of the CHW buffer as HWC, which scrambles it; the reference does exactly this -- scaled to
    uint8 (truncation), resized to 224x224 with PIL's antialiased bilinear filter, divided by 255 and normalised with
    the ImageNet statistics.  PIL is test infrastructure here; the product restates the filter (csrc/elementwise.cu)."""
    import numpy as np
    from PIL import Image

    a = ((gen_img0.detach().cpu().reshape(256, 256, 3).numpy() * .5 + .5) * 255).astype(np.uint8)
    im = Image.fromarray(a).resize((224, 224), Image.BILINEAR)
    t = torch.from_numpy(np.asarray(im).astype(np.float32) / 255.0).permute(2, 0, 1)
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    return ((t - mean[:, None, None]) / std[:, None, None]).unsqueeze(0)


def _bn(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, eps)


def resnet18_logits(sd, x):
    """torchvision.models.resnet18(num_classes=365).forward in eval mode (the classifier of z_buffermodel.py:88,258)."""
    h = F.relu(_bn(sd, "bn1.", F.conv2d(x, sd["conv1.weight"], None, stride=2, padding=3)))
    h = F.max_pool2d(h, 3, 2, 1)
    for layer, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for blk in range(2):
            p = f"layer{layer}.{blk}."
            s = stride if blk == 0 else 1
            idt = h
            o = F.relu(_bn(sd, p + "bn1.", F.conv2d(h, sd[p + "conv1.weight"], None, stride=s, padding=1)))
            o = _bn(sd, p + "bn2.", F.conv2d(o, sd[p + "conv2.weight"], None, padding=1))
            if (p + "downsample.0.weight") in sd:
                idt = _bn(sd, p + "downsample.1.", F.conv2d(h, sd[p + "downsample.0.weight"], None, stride=s))
            h = F.relu(o + idt)
    h = F.adaptive_avg_pool2d(h, 1).flatten(1)
    return F.linear(h, sd["fc.weight"], sd["fc.bias"])


def entropy(logits):
    """-sum p log p of softmax(logits) per row (z_buffermodel.py:259-261)."""
    p = torch.softmax(logits.double(), 1)
    return -(p * torch.log(p)).sum(1)
