"""Runners of the reference's dense inference networks on the sm_100a kernels (csrc/conv.cu, csrc/elementwise.cu).

Each class takes a state dict with the REFERENCE's key names (so a real checkpoint's sub-dicts load unchanged),
folds what is constant at inference (spectral norm sigma, eval batch-norm statistics) into packed bf16 weights and
fp32 per-channel scale/shift vectors, and replays the module's forward as a fixed sequence of C-ABI launches:
  UnetB200            models/networks/architectures.py:174-279 + depth range models/z_buffermodel.py:304-308
  VQVAETopB200        models/vqvae2/vqvae.py:240-312 (encode -> id_t only, decode_code)
  ResNetDecoderB200   models/networks/architectures.py:126-167, models/layers/blocks.py:33-74,
                      models/layers/normalization.py:21-47,146-171
Activations are NHWC bf16 between kernels; network inputs/outputs are the reference's NCHW fp32 tensors.
"""
import ctypes

import torch

from . import _lib
from ._lib import check
from .conv import ConvOutput, Out, PackedConv, conv_igemm, round_up

MODE = {"identity": 0, "avgpool": 1, "bilinear": 2, "avgpool_nopad": 3, "maxpool": 4}
_ACT = {"none": 0, "relu": 1, "leaky": 2, "tanh": 3, "sigmoid_affine": 4, "elu": 5}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def nchw_to_nhwc(x, cstride, mask=None):
    n, c, h, w = x.shape
    x = x.contiguous().float()
    out = torch.empty((n, h, w, cstride), dtype=torch.bfloat16, device=x.device)
    m = None if mask is None else mask.contiguous().view(torch.uint8)
    check(_lib.lib().ps_nchw_to_nhwc_bf16(_p(x), n, c, h, w, _p(m), _p(out), cstride, _stream()), "ps_nchw_to_nhwc_bf16")
    return out


def _fill_out(co, o):
    co.ptr, co.cstride, co.coffset = o.buf.data_ptr(), o.buf.shape[3], o.coffset
    co.act, co.per_sample = _ACT[o.act], int(o.per_sample)
    co.scale, co.shift = _p(o.scale), _p(o.shift)


def resample(x, mode, out0, out1=None, C=None):
    """x NHWC bf16 -> out0 (and out1): Out(buf, act, scale, shift, per_sample, coffset)."""
    n, h, w, cs = x.shape
    C = cs if C is None else C
    a, b = ConvOutput(), ConvOutput()
    _fill_out(a, out0)
    if out1 is not None:
        _fill_out(b, out1)
    check(_lib.lib().ps_resample(_p(x), n, h, w, C, cs, MODE[mode], ctypes.byref(a), ctypes.byref(b) if out1 else None,
                                 _stream()), "ps_resample")


def spectral_fold(sd, prefix):
    w = sd[prefix + "weight_orig"].float()
    sigma = torch.dot(sd[prefix + "weight_u"].float(), w.reshape(w.shape[0], -1) @ sd[prefix + "weight_v"].float())
    return w / sigma


def _bn_fold(sd, prefix, device, eps=1e-5):
    scale = sd[prefix + "weight"].float() / torch.sqrt(sd[prefix + "running_var"].float() + eps)
    shift = sd[prefix + "bias"].float() - sd[prefix + "running_mean"].float() * scale
    return scale.to(device).contiguous(), shift.to(device).contiguous()


def _buf(n, h, w, c, device):
    return torch.empty((n, h, w, c), dtype=torch.bfloat16, device=device)


class UnetB200:
    ENC_BN = [None, None, "batch_norm2_0.", "batch_norm4_0.", "batch_norm8_0.", "batch_norm8_1.", "batch_norm8_2.",
              "batch_norm8_3.", None]
    DEC_BN = [None, "batch_norm8_4.", "batch_norm8_5.", "batch_norm8_6.", "batch_norm8_7.", "batch_norm4_1.",
              "batch_norm2_1.", "batch_norm.", None]

    def __init__(self, sd, device="cuda"):
        self.device = device
        self.down, self.up, self.ebn, self.dbn = [None], [None], [None], [None]
        for i in range(1, 9):
            self.down.append(PackedConv.conv2d(spectral_fold(sd, f"conv{i}."), sd[f"conv{i}.bias"], padding=1, device=device))
            self.up.append(PackedConv.conv2d(spectral_fold(sd, f"dconv{i}."), sd[f"dconv{i}.bias"], padding=1, device=device))
            self.ebn.append(_bn_fold(sd, self.ENC_BN[i], device) if self.ENC_BN[i] else (None, None))
            self.dbn.append(_bn_fold(sd, self.DEC_BN[i], device) if self.DEC_BN[i] else (None, None))
        self.ec = [None] + [self.down[i].Cout for i in range(1, 9)]
        self.dc = [None] + [self.up[i].Cout for i in range(1, 9)]

    def forward(self, x, min_z, max_z):
        """x (N,3,S,S) f32 in [-1,1] -> depth (N,1,S,S) f32 = sigmoid(Unet(x)) * (max_z - min_z) + min_z."""
        n, _, S, _ = x.shape
        dev = x.device
        cur = nchw_to_nhwc(x, 8)
        # concat buffers of the decoder: cat[j] = [relu(d_j_) | relu(e_{8-j})] at resolution 2^j (S = 256)
        cat = [None] + [_buf(n, S >> (8 - j), S >> (8 - j), self.dc[j] + self.ec[8 - j], dev) for j in range(1, 8)]
        for i in range(1, 9):
            r = S >> i
            sc, sh = self.ebn[i]
            outs = []
            if i < 8:
                nxt = _buf(n, r, r, self.ec[i], dev)
                outs = [Out(nxt, "leaky", sc, sh), Out(cat[8 - i], "relu", sc, sh, coffset=self.dc[8 - i])]
            else:
                nxt = _buf(n, r, r, self.ec[i], dev)
                outs = [Out(nxt, "relu", sc, sh)]
            conv_igemm(cur, self.down[i], outs, stride=2, Hout=r, Wout=r)
            cur = nxt
        src = cur  # relu(e8), 1x1
        depth = torch.empty((n, 1, S, S), dtype=torch.float32, device=dev)
        for j in range(1, 9):
            r = S >> (8 - j)
            up = _buf(n, r, r, src.shape[3], dev)
            resample(src, "bilinear", Out(up))
            sc, sh = self.dbn[j]
            if j < 8:
                conv_igemm(up, self.up[j], [Out(cat[j], "relu", sc, sh)])
                src = cat[j]
            else:
                conv_igemm(up, self.up[j], [Out(None, "sigmoid_affine")], out_f32=depth,
                           act_param=(float(max_z - min_z), float(min_z)))
        return depth


class VQVAETopB200:
    def __init__(self, sd, device="cuda"):
        self.device = device
        c2 = lambda k, pad: PackedConv.conv2d(sd[k + ".weight"].float(), sd[k + ".bias"], padding=pad, device=device)
        self.eb0, self.eb2, self.eb4 = c2("enc_b.blocks.0", 1), c2("enc_b.blocks.2", 1), c2("enc_b.blocks.4", 1)
        self.eb_res = [(c2(f"enc_b.blocks.{i}.conv.1", 1), c2(f"enc_b.blocks.{i}.conv.3", 0)) for i in (5, 6)]
        self.et0, self.et2 = c2("enc_t.blocks.0", 1), c2("enc_t.blocks.2", 1)
        self.et_res = [(c2(f"enc_t.blocks.{i}.conv.1", 1), c2(f"enc_t.blocks.{i}.conv.3", 0)) for i in (3, 4)]
        self.qconv = c2("quantize_conv_t", 0)
        self.embed = sd["quantize_t.embed"].float().to(device).contiguous()  # (64, 512)
        ct = lambda k: [[PackedConv.conv_transpose_4x4_s2_phase(sd[k + ".weight"].float(), py, px, sd[k + ".bias"], device)
                         for px in range(2)] for py in range(2)]
        self.up_t = ct("upsample_t")
        self.d0 = c2("dec.blocks.0", 1)
        self.d_res = [(c2(f"dec.blocks.{i}.conv.1", 1), c2(f"dec.blocks.{i}.conv.3", 0)) for i in (1, 2)]
        self.d4, self.d6 = ct("dec.blocks.4"), ct("dec.blocks.6")

    @staticmethod
    def _resblock(r, pcs, dev):
        """r = relu(block input).  relu(conv1x1(relu(conv3x3(r))) + r): the in-place ReLU of the reference makes
        the skip add relu(input) (vqvae.py:84-95), and every consumer of a block output applies ReLU first."""
        n, h, w, _ = r.shape
        t = _buf(n, h, w, pcs[0].Cout, dev)
        conv_igemm(r, pcs[0], [Out(t, "relu")])
        o = _buf(n, h, w, pcs[1].Cout, dev)
        conv_igemm(t, pcs[1], [Out(o, "relu")], residual=r)
        return o

    @staticmethod
    def _conv_t(x, phases, out_buf=None, act="none", out_f32=None):
        n, h, w, _ = x.shape
        for py in range(2):
            for px in range(2):
                conv_igemm(x, phases[py][px], [Out(out_buf, act)], Hout=h, Wout=w, out_f32=out_f32,
                           geometry=(2 * h, 2 * w, 2, 2, py, px))

    def pre_quant(self, x):
        n, _, S, _ = x.shape
        dev = x.device
        a = nchw_to_nhwc(x, 8)
        b = _buf(n, S // 2, S // 2, 64, dev)
        conv_igemm(a, self.eb0, [Out(b, "relu")], stride=2)
        c = _buf(n, S // 4, S // 4, 128, dev)
        conv_igemm(b, self.eb2, [Out(c, "relu")], stride=2)
        r = _buf(n, S // 4, S // 4, 128, dev)
        conv_igemm(c, self.eb4, [Out(r, "relu")])
        for pcs in self.eb_res:
            r = self._resblock(r, pcs, dev)
        t = _buf(n, S // 8, S // 8, 64, dev)
        conv_igemm(r, self.et0, [Out(t, "relu")], stride=2)
        r = _buf(n, S // 8, S // 8, 128, dev)
        conv_igemm(t, self.et2, [Out(r, "relu")])
        for pcs in self.et_res:
            r = self._resblock(r, pcs, dev)
        z = torch.empty((n, 64, S // 8, S // 8), dtype=torch.float32, device=dev)
        conv_igemm(r, self.qconv, [Out(None)], out_f32=z)
        return z

    def encode_top(self, x):
        """VQVAETop.encode(x)[3]: (N,3,S,S) f32 -> id_t (N,S/8,S/8) int64."""
        z = self.pre_quant(x)
        n, d, h, w = z.shape
        ids = torch.empty((n, h, w), dtype=torch.int64, device=z.device)
        check(_lib.lib().ps_vq_argmin(_p(z), n, d, h * w, _p(self.embed), self.embed.shape[1], _p(ids), _stream()),
              "ps_vq_argmin")
        return ids

    def decode_code(self, ids):
        """VQVAETop.decode_code: (N,h,w) int64 -> (N,3,8h,8w) f32."""
        n, h, w = ids.shape
        dev = ids.device
        ids = ids.contiguous()
        q = _buf(n, h, w, 64, dev)
        check(_lib.lib().ps_embed_codes(_p(ids), n * h * w, 64, _p(self.embed), self.embed.shape[1], _p(q), _stream()),
              "ps_embed_codes")
        u = _buf(n, 2 * h, 2 * w, 64, dev)
        self._conv_t(q, self.up_t, u)
        r = _buf(n, 2 * h, 2 * w, 128, dev)
        conv_igemm(u, self.d0, [Out(r, "relu")])
        for pcs in self.d_res:
            r = self._resblock(r, pcs, dev)
        v = _buf(n, 4 * h, 4 * w, 64, dev)
        self._conv_t(r, self.d4, v, "relu")
        out = torch.empty((n, 3, 8 * h, 8 * w), dtype=torch.float32, device=dev)
        self._conv_t(v, self.d6, None, "none", out_f32=out)
        return out


class ResNetDecoderB200:
    CH = [4, 64, 128, 256, 256, 128, 128, 128, 3]                        # configs.py:221-231, ngf = 64
    RESAMPLE = [None, "avgpool", "avgpool", None, "bilinear", "bilinear", None, None]  # configs.py:232-241

    def __init__(self, sd, device="cuda", normalize_before_residual=False):
        self.device = device
        self.nbr = bool(normalize_before_residual)
        self.blocks = []
        f = lambda t: t.float().to(device).contiguous()
        for b in range(8):
            p = f"eblocks.{b}."
            cin, cout = self.CH[b], self.CH[b + 1]
            blk = {}
            for tag, k in (("n1", "ch_a.0."), ("n2", "ch_a.3.")):
                blk[tag] = dict(Wg=f(spectral_fold(sd, p + k + "gain.")), Wb=f(spectral_fold(sd, p + k + "bias.")),
                                mean=f(sd[p + k + "bn.stored_mean"]), var=f(sd[p + k + "bn.stored_var"]))
            blk["aa"] = PackedConv.conv2d(spectral_fold(sd, p + "ch_a.2."), sd[p + "ch_a.2.bias"], padding=1, device=device)
            w_ab = spectral_fold(sd, p + "ch_a.5.")
            has_b = (p + "ch_b.0.weight_orig") in sd
            taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
            wt = [w_ab[:, :, ky, kx] for ky in range(3) for kx in range(3)]
            bias = sd[p + "ch_a.5.bias"].float()
            if has_b:
                # the 1x1 skip convolution accumulates into the same tile: tenth weight block, read from the raw input
                wb = spectral_fold(sd, p + "ch_b.0.")[:, :, 0, 0]
                kpad = max(cout, round_up(cin, 8))
                wt = [torch.nn.functional.pad(w, (0, kpad - w.shape[1])) for w in wt]
                wt.append(torch.nn.functional.pad(wb, (0, kpad - wb.shape[1])))
                bias = bias + sd[p + "ch_b.0.bias"].float()
                pc = PackedConv(wt, taps + [(0, 0)], bias, device)
                pc.taps = taps
                pc.Cin = cout
            else:
                pc = PackedConv(wt, taps, bias, device)
            blk["ab"], blk["has_b"] = pc, has_b
            self.blocks.append(blk)

    def _affine(self, z, nb, C, cpad):
        n = z.shape[0]
        scale = torch.empty((n, cpad), dtype=torch.float32, device=z.device)
        shift = torch.empty((n, cpad), dtype=torch.float32, device=z.device)
        check(_lib.lib().ps_noise_affine(_p(z), n, z.shape[1], _p(nb["Wg"]), _p(nb["Wb"]), _p(nb["mean"]), _p(nb["var"]),
                                         1e-5, C, cpad, _p(scale), _p(shift), _stream()), "ps_noise_affine")
        return scale, shift

    def forward(self, x, background_mask, noise=None):
        """x (N,3,S,S) f32, background_mask (N,S,S) bool, noise (16,N,20) f32 (drawn with torch.randn when None, as
        LinearNoiseLayer does on every forward) -> tanh(eblocks(cat(x, ~bg)) + x), (N,3,S,S) f32."""
        n, _, S, _ = x.shape
        dev = x.device
        x = x.contiguous().float()
        if noise is None:
            noise = torch.randn(16, n, 20, device=dev)
        noise = noise.to(dev).float().contiguous()
        aff = []
        for b in range(8):
            cin, cout = self.CH[b], self.CH[b + 1]
            aff.append((self._affine(noise[2 * b], self.blocks[b]["n1"], cin, round_up(cin, 8)),
                        self._affine(noise[2 * b + 1], self.blocks[b]["n2"], cout, cout)))
        raw = nchw_to_nhwc(x, 8, background_mask)        # cat(x, float(~bg)), channels 4..7 zero
        a0 = _buf(n, S, S, 8, dev)
        resample(raw, "identity", Out(a0, "relu", aff[0][0][0], aff[0][0][1], per_sample=True))
        r = S
        v32 = torch.empty((n, 3, S, S), dtype=torch.float32, device=dev)
        for b in range(8):
            blk = self.blocks[b]
            cin, cout, kind = self.CH[b], self.CH[b + 1], self.RESAMPLE[b]
            cpo = round_up(cout, 8)
            s2, t2 = aff[b][1]
            a1 = _buf(n, r, r, cpo, dev)
            if cpo != cout:
                a1.zero_()
            conv_igemm(a0, blk["aa"], [Out(a1, "relu", s2, t2, per_sample=True)])
            kw = {}
            if blk["has_b"]:
                kw = dict(x2=raw, pc2_taps=[(0, 0)], pc2_wrow=[9 * blk["ab"].cout_pad], cin2=round_up(cin, 8))
            else:
                kw = dict(residual=raw)
            last = b == 7
            nxt_aff = None if last else aff[b + 1][0]
            if last:
                conv_igemm(a1, blk["ab"], [Out(None)], out_f32=v32, cin=cpo, **kw)
                break
            if kind is None:
                nraw, na0 = _buf(n, r, r, cpo, dev), _buf(n, r, r, cpo, dev)
                conv_igemm(a1, blk["ab"], [Out(nraw), Out(na0, "relu", nxt_aff[0], nxt_aff[1], per_sample=True)], cin=cpo, **kw)
            else:
                t = _buf(n, r, r, cpo, dev)
                conv_igemm(a1, blk["ab"], [Out(t)], cin=cpo, **kw)
                r = r // 2 if kind == "avgpool" else r * 2
                nraw, na0 = _buf(n, r, r, cpo, dev), _buf(n, r, r, cpo, dev)
                resample(t, kind, Out(nraw), Out(na0, "relu", nxt_aff[0], nxt_aff[1], per_sample=True))
            raw, a0 = nraw, na0
        out = torch.empty_like(x)
        check(_lib.lib().ps_tanh_residual(_p(v32), _p(x), x.numel(), int(self.nbr), _p(out), _stream()), "ps_tanh_residual")
        return out


# ---------------------------------------------------------------------------------------------------------------
# Sample ranking: ZbufferModelPts.get_best_sample scores every candidate with the multiscale PatchGAN's
# D_Fake and with the entropy of a places365 resnet18 (models/z_buffermodel.py:244-276)
# ---------------------------------------------------------------------------------------------------------------
def pil_bilinear_table(in_size=256, out_size=224):
    """First tap and the (at most three, stored as four) 22-bit fixed-point weights of PIL's antialiased bilinear resize
    per output coordinate -- the filter torchvision's Resize applies to the PIL image at z_buffermodel.py:105-110.
    Restates Pillow's precompute_coeffs + normalize_coeffs_8bpc (support = scale, triangle kernel, weights normalised
    in double, rounded to (int)(0.5 + w * 2^22)); with these integers the device kernel reproduces PIL bit for bit
    (tests/test_demo_cpu.py checks the table against PIL itself)."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    tap0, kk = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) / fs)) for x in range(xmax)]
        tot = 0.0
        for v in k:
            tot += v
        ki = [int(0.5 + v / tot * (1 << 22)) for v in k]
        while ki and ki[-1] == 0:
            ki.pop()
        assert len(ki) <= 4, "PIL bilinear 256->224 never needs more than four taps"
        tap0.append(xmin)
        kk.append(ki + [0] * (4 - len(ki)))
    return torch.tensor(tap0, dtype=torch.int32), torch.tensor(kk, dtype=torch.int32)


class MultiscaleDiscriminatorB200:
    """models/networks/discriminators.py:142-207 (two NLayerDiscriminators :78-140, n_layers_D 4, ndf 64, spectralinstance)
    on conv_igemm_kernel; InstanceNorm2d = instnorm_stats_kernel + a per-sample affine fused with LeakyReLU in the
    resampling kernel.  forward -> the last feature map of each scale (fp32), d_fake -> the D_Fake score per candidate."""

    def __init__(self, sd, device="cuda"):
        self.device = device
        self.scales = []
        for d in range(2):
            p = f"discriminator_{d}."
            first = PackedConv.conv2d(sd[p + "model0.0.weight"].float(), sd[p + "model0.0.bias"], padding=2, device=device)
            mid = [PackedConv.conv2d(spectral_fold(sd, p + f"model{n}.0.0."), None, padding=2, device=device) for n in (1, 2, 3)]
            last = PackedConv.conv2d(sd[p + "model4.0.weight"].float(), sd[p + "model4.0.bias"], padding=2, device=device)
            self.scales.append((first, mid, last))

    def _instnorm_leaky(self, raw):
        n, h, w, c = raw.shape
        scale = torch.empty((n, c), dtype=torch.float32, device=raw.device)
        shift = torch.empty((n, c), dtype=torch.float32, device=raw.device)
        check(_lib.lib().ps_instance_norm_stats(_p(raw), n, h * w, c, c, 1e-5, _p(scale), _p(shift), _stream()),
              "ps_instance_norm_stats")
        out = _buf(n, h, w, c, raw.device)
        resample(raw, "identity", Out(out, "leaky", scale, shift, per_sample=True))
        return out

    def forward(self, x):
        """x (M,3,S,S) f32 -> [ (M,1,h0,w0), (M,1,h1,w1) ] f32."""
        m, _, S, _ = x.shape
        dev = x.device
        cur = nchw_to_nhwc(x, 8)
        outs = []
        for d, (first, mid, last) in enumerate(self.scales):
            r = cur.shape[1]
            o = r // 2 + 1                                           # kernel 4, stride 2, padding 2
            h = _buf(m, o, o, 64, dev)
            conv_igemm(cur, first, [Out(h, "leaky")], stride=2, Hout=o, Wout=o)
            for pc, stride in zip(mid, (2, 2, 1)):
                o = h.shape[1] // 2 + 1 if stride == 2 else h.shape[1] + 1
                raw = _buf(m, o, o, pc.Cout, dev)
                conv_igemm(h, pc, [Out(raw)], stride=stride, Hout=o, Wout=o)
                h = self._instnorm_leaky(raw)
            o = h.shape[1] + 1
            y = torch.empty((m, 1, o, o), dtype=torch.float32, device=dev)
            conv_igemm(h, last, [Out(None)], Hout=o, Wout=o, out_f32=y)
            outs.append(y)
            if d == 0:
                nxt = _buf(m, (r + 1) // 2, (r + 1) // 2, 8, dev)
                resample(cur, "avgpool_nopad", Out(nxt))
                cur = nxt
        return outs

    def d_fake(self, x, groups):
        """D_Fake (gan_loss.py:172-181, hinge) of `groups` candidates stacked along the batch: (groups,) f32."""
        outs = self.forward(x)
        per = [torch.relu(o.view(groups, -1) + 1).mean(1) for o in outs]      # -min(-x - 1, 0) = max(x + 1, 0)
        return sum(per) / len(per)


class ResNet18B200:
    """torchvision.models.resnet18(num_classes=365) in eval mode (the places365 classifier, z_buffermodel.py:88,258) on
    conv_igemm_kernel: batch norm folded into weights and bias, the 7x7 stem as two launches (49 taps > 32 per launch),
    the strided 1x1 downsample convolutions as their own launches, the residual add + ReLU in the epilogue."""

    def __init__(self, sd, device="cuda"):
        self.device = device

        def fold(conv, bn):
            s = sd[bn + "weight"].float() / torch.sqrt(sd[bn + "running_var"].float() + 1e-5)
            return sd[conv + "weight"].float() * s[:, None, None, None], sd[bn + "bias"].float() - sd[bn + "running_mean"].float() * s

        w, b = fold("conv1.", "bn1.")
        taps = [(ky - 3, kx - 3) for ky in range(7) for kx in range(7)]
        wt = [w[:, :, ky, kx] for ky in range(7) for kx in range(7)]
        self.stem_a = PackedConv(wt[:32], taps[:32], None, device)
        self.stem_a.taps = taps[:16]
        self.stem_b = PackedConv(wt[32:], taps[32:], b, device)
        self.stem_b.taps = taps[32:48]
        self.blocks = []
        for layer, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
            for blk in range(2):
                p = f"layer{layer}.{blk}."
                w1, b1 = fold(p + "conv1.", p + "bn1.")
                w2, b2 = fold(p + "conv2.", p + "bn2.")
                ds = None
                if (p + "downsample.0.weight") in sd:
                    wd, bd = fold(p + "downsample.0.", p + "downsample.1.")
                    ds = PackedConv.conv2d(wd, bd, padding=0, device=device)
                self.blocks.append((stride if blk == 0 else 1, PackedConv.conv2d(w1, b1, padding=1, device=device),
                                    PackedConv.conv2d(w2, b2, padding=1, device=device), ds))
        self.fc = PackedConv.conv2d(sd["fc.weight"].float()[:, :, None, None], sd["fc.bias"], padding=0, device=device)
        self.tap0, self.wtab = [t.to(device).contiguous() for t in pil_bilinear_table()]

    def classifier_input(self, imgs):
        """imgs (M,B,3,256,256) or (M,3,256,256) f32: image 0 of each candidate -> (M,224,224,8) NHWC bf16."""
        imgs = imgs.contiguous().float()
        m = imgs.shape[0]
        stride = imgs[0].numel()
        out = torch.empty((m, 224, 224, 8), dtype=torch.bfloat16, device=imgs.device)
        check(_lib.lib().ps_classifier_input(_p(imgs), stride, m, _p(self.tap0), _p(self.wtab), _p(out), _stream()),
              "ps_classifier_input")
        return out

    def logits_nhwc(self, x):
        m = x.shape[0]
        dev = x.device
        a = _buf(m, 112, 112, 64, dev)
        conv_igemm(x, self.stem_a, [Out(a)], stride=2, Hout=112, Wout=112, x2=x, pc2_taps=[(ky - 3, kx - 3) for ky in range(7)
                   for kx in range(7)][16:32], pc2_wrow=self.stem_a.wrow[16:32])
        h = _buf(m, 112, 112, 64, dev)
        conv_igemm(x, self.stem_b, [Out(h, "relu")], stride=2, Hout=112, Wout=112, x2=x, pc2_taps=[(3, 3)],
                   pc2_wrow=self.stem_b.wrow[16:17], residual=a)
        p = _buf(m, 56, 56, 64, dev)
        resample(h, "maxpool", Out(p))
        h = p
        for stride, c1, c2, ds in self.blocks:
            r = h.shape[1] // stride
            t = _buf(m, r, r, c1.Cout, dev)
            conv_igemm(h, c1, [Out(t, "relu")], stride=stride, Hout=r, Wout=r)
            idt = h
            if ds is not None:
                idt = _buf(m, r, r, ds.Cout, dev)
                conv_igemm(h, ds, [Out(idt)], stride=stride, Hout=r, Wout=r)
            o = _buf(m, r, r, c2.Cout, dev)
            conv_igemm(t, c2, [Out(o, "relu")], residual=idt)
            h = o
        pooled = h.float().mean((1, 2)).to(torch.bfloat16).view(m, 1, 1, 512).contiguous()   # AdaptiveAvgPool2d(1)
        out = torch.empty((m, 365, 1, 1), dtype=torch.float32, device=dev)
        conv_igemm(pooled, self.fc, [Out(None)], out_f32=out)
        return out.view(m, 365)

    def entropy(self, imgs):
        """-sum p log p of the class distribution of each candidate's image 0 (z_buffermodel.py:256-261): (M,) f64."""
        lg = self.logits_nhwc(self.classifier_input(imgs)).double()
        p = torch.softmax(lg, 1)
        return -(p * torch.log(p)).sum(1)
