"""Host side of the activation-cached lmconv sampler (csrc/lmconv.cu) and of the native order/mask glue (csrc/glue.cu).

LmconvB200 takes the reference's OurPixelCNN state dict (models/lmconv/model.py:61-108, instantiated as at
models/z_buffermodel.py:62-74), packs every layer as [tap][cin][cout] bf16 and exposes
  sample(codes, order, words, sample_mask, uniforms, temperature)   ~ models/lmconv/sample.py:8-73
  logits(codes, order, words)                                        ~ OurPixelCNN.forward (teacher forced)
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check

U0, DS0, DS1, US0, US1 = 0, 13, 14, 31, 32


class _Op(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("kind", "og", "a", "mid", "out", "w_in", "b_in", "w_skip", "b_skip", "w_out",
                                            "b_out")]


class _Weights(ctypes.Structure):
    _fields_ = [("weights", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("w_uinit", ctypes.c_int),
                ("b_uinit", ctypes.c_int), ("w_nin", ctypes.c_int), ("b_nin", ctypes.c_int), ("ops", _Op * 18)]


def glue_host(background_mask):
    """ZbufferModelPts.get_masks_for_batch (z_buffermodel.py:641-701) in native host code.
    background_mask (B,256,256) bool tensor (any device) -> numpy dist (B,32,32) i32, order (B,1024) i32,
    words (B,3,1024) u16, sample_mask (B,32,32) bool."""
    m = np.ascontiguousarray(background_mask.detach().to("cpu").numpy().astype(np.uint8))
    B, S, _ = m.shape
    dist = np.zeros((B, 32, 32), np.int32)
    order = np.zeros((B, 1024), np.int32)
    words = np.zeros((B, 3, 1024), np.uint16)
    smask = np.zeros((B, 32, 32), np.uint8)
    check(_lib.lib().ps_lmconv_glue_host(m.ctypes.data, B, S, dist.ctypes.data, order.ctypes.data, words.ctypes.data,
                                         smask.ctypes.data), "ps_lmconv_glue_host")
    return dist, order, words, smask.astype(bool)


class LmconvB200:
    def __init__(self, sd, device="cuda"):
        self.device = device
        ws, bs = [], []
        self._wn = self._bn = 0

        def add_w(t):  # t: (rows, cout) fp32
            off = self._wn
            flat = t.reshape(-1).float()
            pad = (-flat.numel()) % 8          # keep every block 16-byte aligned for the kernel's uint4 loads
            ws.append(torch.cat([flat, torch.zeros(pad)]))
            self._wn += flat.numel() + pad
            return off

        def add_b(t):
            off = self._bn
            bs.append(t.reshape(-1).float())
            self._bn += t.numel()
            return off

        def conv(prefix):  # (Cout,Cin,3,3) -> [tap][cin][cout]
            w = sd[prefix + "weight"].float()
            return add_w(w.permute(2, 3, 1, 0).reshape(9 * w.shape[1], w.shape[0])), add_b(sd[prefix + "bias"])

        def nin(prefix):   # weight-normed Linear: g * v / |v| per output row -> [cin][cout]
            v = sd[prefix + "lin_a.weight_v"].float()
            w = sd[prefix + "lin_a.weight_g"].float() * v / v.norm(dim=1, keepdim=True)
            return add_w(w.t().contiguous()), add_b(sd[prefix + "lin_a.bias"])

        self.w = _Weights()
        self.w.w_uinit, self.w.b_uinit = conv("u_init.")
        ops = []

        def resnet(prefix, og, a, mid, out):
            o = _Op()
            o.kind, o.og, o.a, o.mid, o.out = 0, og, a, mid, out
            o.w_in, o.b_in = conv(prefix + "conv_input.")
            if a >= 0:
                o.w_skip, o.b_skip = nin(prefix + "nin_skip.")
            o.w_out, o.b_out = conv(prefix + "conv_out.")
            ops.append(o)

        def dilated(prefix, src, dst):
            o = _Op()
            o.kind, o.og, o.a, o.mid, o.out = 1, src, -1, -1, dst
            o.w_in, o.b_in = conv(prefix)
            ops.append(o)

        # up pass (model.py:130-141); u_list = [U0, 2, 4, DS0, 6, 8, DS1, 10, 12]
        resnet("up_layers.0.u_stream.0.", U0, -1, 1, 2)
        resnet("up_layers.0.u_stream.1.", 2, -1, 3, 4)
        dilated("downsize_u_stream.0.", 4, DS0)
        resnet("up_layers.1.u_stream.0.", DS0, -1, 5, 6)
        resnet("up_layers.1.u_stream.1.", 6, -1, 7, 8)
        dilated("downsize_u_stream.1.", 8, DS1)
        resnet("up_layers.2.u_stream.0.", DS1, -1, 9, 10)
        resnet("up_layers.2.u_stream.1.", 10, -1, 11, 12)
        # down pass (model.py:145-151): u = 12; skips pop 10, DS1 | 8, 6, DS0 | 4, 2, U0
        resnet("down_layers.0.u_stream.0.", 12, 10, 15, 16)
        resnet("down_layers.0.u_stream.1.", 16, DS1, 17, 18)
        dilated("upsize_u_stream.0.", 18, US0)
        resnet("down_layers.1.u_stream.0.", US0, 8, 19, 20)
        resnet("down_layers.1.u_stream.1.", 20, 6, 21, 22)
        resnet("down_layers.1.u_stream.2.", 22, DS0, 23, 24)
        dilated("upsize_u_stream.1.", 24, US1)
        resnet("down_layers.2.u_stream.0.", US1, 4, 25, 26)
        resnet("down_layers.2.u_stream.1.", 26, 2, 27, 28)
        resnet("down_layers.2.u_stream.2.", 28, U0, 29, 30)
        assert len(ops) == 18
        for i, o in enumerate(ops):
            self.w.ops[i] = o
        self.w.w_nin, self.w.b_nin = nin("nin_out.")
        self.W = torch.cat(ws).to(device=device, dtype=torch.bfloat16).contiguous()
        self.bias = torch.cat(bs).to(device).contiguous()
        self.w.weights, self.w.bias = self.W.data_ptr(), self.bias.data_ptr()
        self._cache = None

    def _run(self, codes, order, words, sample_mask, uniforms, temperature, nsteps, sample, want_logits):
        dev = self.device
        B = codes.shape[0]
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(device=dev, dtype=dt).contiguous()
        codes_d = t(codes, torch.int64).reshape(B, 1024).clone()
        order_d = t(order, torch.int32).reshape(B, 1024)
        wn = words.detach().cpu().numpy() if torch.is_tensor(words) else np.asarray(words)
        words_d = torch.from_numpy(np.ascontiguousarray(wn).astype(np.uint16).view(np.int16).reshape(B, 3, 1024)).to(dev)
        smask_d = t(sample_mask, torch.uint8).reshape(B, 1024)
        nsteps_d = t(nsteps, torch.int32).reshape(B)
        uni_d = None if uniforms is None else t(uniforms, torch.float32).reshape(B, -1)
        logits = torch.empty((B, 1024, 512), dtype=torch.float32, device=dev) if want_logits else None
        nbytes = _lib.lib().ps_lmconv_cache_bytes(B)
        if self._cache is None or self._cache.numel() < nbytes:
            self._cache = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(self.W.device):
            check(_lib.lib().ps_lmconv_sample(
                ctypes.byref(self.w), B, order_d.data_ptr(), words_d.data_ptr(), smask_d.data_ptr(), codes_d.data_ptr(),
                None if uni_d is None else uni_d.data_ptr(), 0 if uni_d is None else uni_d.shape[1], float(temperature),
                nsteps_d.data_ptr(), int(sample), None if logits is None else logits.data_ptr(), self._cache.data_ptr(),
                nbytes, torch.cuda.current_stream().cuda_stream), "ps_lmconv_sample")
        return codes_d.view(B, 32, 32), logits

    @staticmethod
    def steps_needed(order, sample_mask):
        """cells of the generation order up to and including the last sampled one (0 when nothing is sampled)."""
        sm = np.asarray(sample_mask).reshape(len(order), -1)
        out = np.zeros(len(order), np.int32)
        for b in range(len(order)):
            hit = np.nonzero(sm[b][np.asarray(order[b])])[0]
            out[b] = hit[-1] + 1 if hit.size else 0
        return out

    def sample(self, codes, order, words, sample_mask, uniforms, temperature=1.0):
        """codes (B,32,32) int64 with the known cells; returns codes with the sample_mask cells drawn in generation
        order (the argmax of sample.py's one-hot `data`, as z_buffermodel.py:249 takes it)."""
        smn = sample_mask.detach().cpu().numpy() if torch.is_tensor(sample_mask) else np.asarray(sample_mask)
        ordn = order.detach().cpu().numpy() if torch.is_tensor(order) else np.asarray(order)
        nsteps = self.steps_needed(ordn, smn)
        out, _ = self._run(codes, order, words, smn.astype(np.uint8), uniforms, temperature, nsteps, 1, False)
        return out

    def logits(self, codes, order, words):
        """Teacher-forced logits of every cell given all codes: (B,512,32,32) like OurPixelCNN.forward."""
        B = codes.shape[0]
        _, lg = self._run(codes, order, words, np.zeros((B, 1024), np.uint8), None, 1.0, np.full(B, 1024, np.int32), 0, True)
        return lg.view(B, 32, 32, 512).permute(0, 3, 1, 2).contiguous()
