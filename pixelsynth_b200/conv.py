"""Host side of the TMA + tcgen05 implicit-GEMM convolution (csrc/conv.cu, C ABI ps_conv_igemm).

Activations are NHWC bf16 tensors whose channel count is padded to a multiple of 8; weights are packed once into
the [taps * cout_pad, cin_pad] bf16 matrix the kernel's weight tensor map reads.  There is no fallback: every
call goes through libpixelsynth_b200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import check

ACT = {"none": 0, "relu": 1, "leaky": 2, "tanh": 3, "sigmoid_affine": 4, "elu": 5}


class ConvInput(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int), ("C", ctypes.c_int),
                ("cstride", ctypes.c_int), ("ntaps", ctypes.c_int), ("dy", ctypes.c_int * 16), ("dx", ctypes.c_int * 16),
                ("wrow", ctypes.c_int * 16)]


class ConvOutput(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("per_sample", ctypes.c_int), ("act", ctypes.c_int), ("cstride", ctypes.c_int), ("coffset", ctypes.c_int)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("inp", ConvInput * 2), ("weights", ctypes.c_void_p), ("w_rows", ctypes.c_int),
                ("w_cin_pad", ctypes.c_int), ("N", ctypes.c_int), ("Hout", ctypes.c_int), ("Wout", ctypes.c_int),
                ("Cout", ctypes.c_int), ("cout_pad", ctypes.c_int), ("stride", ctypes.c_int), ("bias", ctypes.c_void_p),
                ("residual", ctypes.c_void_p), ("res_cstride", ctypes.c_int), ("out", ConvOutput * 2),
                ("out_f32_nchw", ctypes.c_void_p), ("act_param", ctypes.c_float * 2), ("out_H", ctypes.c_int),
                ("out_W", ctypes.c_int), ("out_sy", ctypes.c_int), ("out_sx", ctypes.c_int), ("out_py", ctypes.c_int),
                ("out_px", ctypes.c_int)]


def round_up(x, m):
    return (x + m - 1) // m * m


def to_nhwc_bf16(x, cpad=8):
    """(N,C,H,W) float -> (N,H,W,Cp) bf16 contiguous, channels zero-padded to a multiple of `cpad`."""
    n, c, h, w = x.shape
    cp = round_up(c, cpad)
    out = torch.zeros((n, h, w, cp), dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out


def from_nhwc(x, c):
    return x[..., :c].permute(0, 3, 1, 2).float().contiguous()


class Out:
    """One NHWC bf16 output of a convolution: y = act(v * scale + shift) written at channel `coffset` of `buf`."""

    def __init__(self, buf, act="none", scale=None, shift=None, per_sample=False, coffset=0):
        self.buf, self.act, self.scale, self.shift, self.per_sample, self.coffset = buf, act, scale, shift, per_sample, coffset


class PackedConv:
    """Weights of one convolution in the kernel's layout.

    taps: list of (dy, dx) input offsets relative to out*stride; weight[t] is (Cout, Cin)."""

    def __init__(self, weight_taps, taps, bias=None, device="cuda"):
        cout, cin = weight_taps[0].shape
        self.Cout, self.Cin = cout, cin
        # the kernel takes the whole padded Cout as one UMMA N when it fits (<= 256), else 128-wide column blocks
        self.cout_pad = round_up(cout, 16) if cout <= 256 else round_up(cout, 128)
        self.cin_pad = round_up(cin, 64)
        self.taps = list(taps)
        T = len(taps)
        w = torch.zeros((T, self.cout_pad, self.cin_pad), dtype=torch.float32)
        for t, wt in enumerate(weight_taps):
            w[t, :cout, :cin] = wt.detach().float().cpu()
        self.w = w.reshape(T * self.cout_pad, self.cin_pad).to(device=device, dtype=torch.bfloat16).contiguous()
        self.bias = None if bias is None else bias.detach().float().to(device).contiguous()
        self.wrow = [t * self.cout_pad for t in range(T)]

    @staticmethod
    def conv2d(weight, bias=None, padding=0, device="cuda"):
        """nn.Conv2d weight (Cout,Cin,kh,kw): tap (ky,kx) reads input (oy*s + ky - pad, ox*s + kx - pad)."""
        cout, cin, kh, kw = weight.shape
        taps = [(ky - padding, kx - padding) for ky in range(kh) for kx in range(kw)]
        wt = [weight[:, :, ky, kx] for ky in range(kh) for kx in range(kw)]
        return PackedConv(wt, taps, bias, device)

    @staticmethod
    def conv_transpose_4x4_s2_phase(weight, py, px, bias=None, device="cuda"):
        """nn.ConvTranspose2d(k=4, s=2, p=1) weight (Cin,Cout,4,4), output phase (oy&1, ox&1) = (py,px):
        out[2y+py, 2x+px] = sum_{ky,kx} in[y + (py+1-ky)/2, x + (px+1-kx)/2] w[:, :, ky, kx] over ky = py+1 (mod 2)."""
        taps, wt = [], []
        for ky in range(4):
            if (py + 1 - ky) % 2:
                continue
            for kx in range(4):
                if (px + 1 - kx) % 2:
                    continue
                taps.append(((py + 1 - ky) // 2, (px + 1 - kx) // 2))
                wt.append(weight[:, :, ky, kx].t())
        return PackedConv(wt, taps, bias, device)


def _fill_input(ci, x, C, taps, wrow):
    n, h, w, cs = x.shape
    ci.ptr, ci.H, ci.W, ci.C, ci.cstride, ci.ntaps = x.data_ptr(), h, w, C, cs, len(taps)
    for t, (dy, dx) in enumerate(taps):
        ci.dy[t], ci.dx[t], ci.wrow[t] = dy, dx, wrow[t]


def conv_igemm(x, pc, outs, stride=1, Hout=None, Wout=None, x2=None, pc2_taps=None, pc2_wrow=None, residual=None,
               out_f32=None, act_param=(0.0, 0.0), geometry=None, cin=None, cin2=None):
    """Launches one implicit-GEMM convolution.

    x: NHWC bf16 input; pc: PackedConv; outs: list of up to two Out.  x2 with (pc2_taps, pc2_wrow): a second
    input accumulated into the same output (its weights live in pc.w at rows pc2_wrow).
    geometry = (out_H, out_W, sy, sx, py, px) places output pixel (oy,ox) at (oy*sy+py, ox*sx+px)."""
    assert x.dtype == torch.bfloat16 and x.is_cuda and x.is_contiguous() and x.dim() == 4
    n, h, w, cs = x.shape
    if Hout is None:
        Hout, Wout = (h, w) if stride == 1 else (h // 2, w // 2)
    d = ConvDesc()
    _fill_input(d.inp[0], x, cin if cin is not None else min(cs, round_up(pc.Cin, 8)), pc.taps, pc.wrow)
    keep = [x]
    if x2 is not None:
        assert x2.dtype == torch.bfloat16 and x2.is_contiguous()
        _fill_input(d.inp[1], x2, cin2 if cin2 is not None else x2.shape[3], pc2_taps, pc2_wrow)
        keep.append(x2)
    d.weights, d.w_rows, d.w_cin_pad = pc.w.data_ptr(), pc.w.shape[0], pc.w.shape[1]
    d.N, d.Hout, d.Wout, d.Cout, d.cout_pad, d.stride = n, Hout, Wout, pc.Cout, pc.cout_pad, stride
    d.bias = None if pc.bias is None else pc.bias.data_ptr()
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous()
        d.residual, d.res_cstride = residual.data_ptr(), residual.shape[3]
    for i, o in enumerate(outs):
        if o is None:
            continue
        co = d.out[i]
        if o.buf is not None:
            assert o.buf.dtype == torch.bfloat16 and o.buf.is_contiguous()
            co.ptr, co.cstride, co.coffset = o.buf.data_ptr(), o.buf.shape[3], o.coffset
        co.act, co.per_sample = ACT[o.act], int(o.per_sample)
        co.scale = None if o.scale is None else o.scale.data_ptr()
        co.shift = None if o.shift is None else o.shift.data_ptr()
        keep += [o.buf, o.scale, o.shift]
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        d.out_f32_nchw = out_f32.data_ptr()
    d.act_param[0], d.act_param[1] = act_param
    if geometry is not None:
        d.out_H, d.out_W, d.out_sy, d.out_sx, d.out_py, d.out_px = geometry
    with torch.cuda.device(x.device):
        check(_lib.lib().ps_conv_igemm(ctypes.byref(d), torch.cuda.current_stream().cuda_stream), "ps_conv_igemm")
    return keep
