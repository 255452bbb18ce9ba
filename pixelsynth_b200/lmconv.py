"""Host side of the wavefront tensor-core lmconv sampler (csrc/lmconv_tc.cu) and of the native order/mask glue
(csrc/glue.cu).

LmconvB200 takes the reference's OurPixelCNN state dict (models/lmconv/model.py:61-108, instantiated as at
models/z_buffermodel.py:62-74), packs every layer as the kernel's K-chunk schedule of pre-swizzled fp16 weight tiles
(ps_lmconv_plan, include/pixelsynth_b200.h) and exposes
  sample(codes, order, words, sample_mask, uniforms, temperature)   ~ models/lmconv/sample.py:8-73
  logits(codes, order, words)                                        ~ OurPixelCNN.forward (teacher forced)
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check

U0, DS0, DS1, US0, US1 = 0, 13, 14, 31, 32
MAX_GEMMS = 40
MAX_LEVELS = 2050
A_GATHER, A_CENTRE, A_EPILOGUE, A_TMEM, A_REUSE = 0, 1, 2, 3, 4


class _Op(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("kind", "og", "a", "mid", "out", "w_in", "b_in", "w_skip", "b_skip", "w_out",
                                            "b_out")]


class _Chunk(ctypes.Structure):
    _fields_ = [("w_off16", ctypes.c_uint32), ("w_rows", ctypes.c_uint16), ("a_kind", ctypes.c_uint8),
                ("a_tensor", ctypes.c_uint8), ("mask", ctypes.c_uint8), ("cin8", ctypes.c_uint8), ("kc", ctypes.c_uint8),
                ("ch_off8", ctypes.c_uint8), ("d_col", ctypes.c_uint16), ("flags", ctypes.c_uint8), ("gemm", ctypes.c_uint8)]


class _Row(ctypes.Structure):
    _fields_ = [("bc", ctypes.c_int32), ("w01", ctypes.c_uint32), ("w2_flags", ctypes.c_uint32), ("uidx", ctypes.c_int32)]


class _Plan(ctypes.Structure):
    _fields_ = [("wblob", ctypes.c_void_p), ("chunks", ctypes.c_void_p), ("n_chunks_body", ctypes.c_int),
                ("n_chunks_total", ctypes.c_int), ("epi_first", ctypes.c_int * MAX_GEMMS), ("w_uinit", ctypes.c_void_p),
                ("bias", ctypes.c_void_p), ("b_uinit", ctypes.c_int), ("b_nin", ctypes.c_int), ("ops", _Op * 18),
                ("raw_mask", ctypes.c_uint64),
                # the same schedule split for the sampled levels: `chain` = what depends on the row's own column (nin_skip,
                # centre tap, nin_out), `halo` = the gathered neighbour taps, whose per-row partial sums other CTAs compute
                ("chunks_chain", ctypes.c_void_p), ("n_chain_body", ctypes.c_int), ("n_chain_total", ctypes.c_int),
                ("chunks_halo", ctypes.c_void_p), ("halo_first", ctypes.c_int * 33), ("part_col", ctypes.c_int * 33)]


assert ctypes.sizeof(_Chunk) == 16 and ctypes.sizeof(_Row) == 16

NONCENTRE_TAPS = [0, 1, 2, 3, 5, 6, 7, 8]


def swizzle_tiles(mat):
    """(rows, K) fp32 with rows % 8 == 0 and K % 64 == 0 -> list of K/64 uint8 arrays, each the fp16 [rows][64] K-major
    tile in the 128-byte-swizzle image tcgen05 reads: 16-byte group j of row r is stored at group j ^ (r & 7)."""
    rows, K = mat.shape
    assert rows % 8 == 0 and K % 64 == 0
    bits = mat.to(torch.float16).contiguous().view(torch.int16).numpy().reshape(rows, K // 64, 8, 8)
    src = np.arange(8)[None, :] ^ (np.arange(rows)[:, None] & 7)          # stored group j holds logical group j ^ (r & 7)
    out = np.take_along_axis(bits, src[:, None, :, None], axis=2)          # (rows, K/64, 8, 8)
    return [np.ascontiguousarray(out[:, kc]).view(np.uint8).reshape(-1) for kc in range(K // 64)]


def _pad_k(mat, mult=64):
    pad = (-mat.shape[1]) % mult
    return mat if pad == 0 else torch.cat([mat, torch.zeros(mat.shape[0], pad)], 1)


def glue_host(background_mask):
    """ZbufferModelPts.get_masks_for_batch (z_buffermodel.py:641-701) in native host code.
    background_mask (B,256,256) bool tensor (any device) -> numpy dist (B,32,32) i32, order (B,1024) i32,
    words (B,3,1024) u16, sample_mask (B,32,32) bool."""
    if isinstance(background_mask, np.ndarray):
        m = np.ascontiguousarray(background_mask.astype(np.uint8, copy=False))
    else:
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
        blobs, chunks, bs, epi_first = [], [], [], []
        self._woff = self._bn = 0

        def add_b(t):
            off = self._bn
            bs.append(t.reshape(-1).float())
            self._bn += t.numel()
            return off

        def add_chunks(tiles, rows, a_kind, tensor, mask, cin8, ch_off8, d_col, first_acc, last_bar=None, gemm=0):
            """one chunk per tile; accumulate flag off only for the very first chunk when first_acc is False"""
            for kc, tile in enumerate(tiles):
                c = _Chunk()
                c.w_off16, c.w_rows, c.a_kind, c.a_tensor, c.mask, c.cin8 = self._woff // 16, rows, a_kind, tensor, mask, cin8
                c.kc, c.ch_off8, c.d_col, c.gemm = kc, ch_off8, d_col, gemm
                c.flags = 1 if (first_acc or kc > 0) else 0
                if last_bar is not None and kc == len(tiles) - 1:
                    c.flags |= 2 | (last_bar << 2)
                if a_kind == A_TMEM and kc == 0:
                    c.flags |= 16   # the issuer waits here for the previous epilogue's operand
                chunks.append(c)
                blobs.append(tile)
                self._woff += tile.size

        def conv_mats(prefix):  # (Cout,Cin,3,3) -> non-centre (Cout, 8*Cin) in tap-slot order, centre (Cout, Cin)
            w = sd[prefix + "weight"].float()
            cout, cin = w.shape[:2]
            w9 = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
            return w9[:, NONCENTRE_TAPS].reshape(cout, 8 * cin), w9[:, 4], add_b(sd[prefix + "bias"])

        def nin_mat(prefix):   # weight-normed Linear: g * v / |v| per output row -> (Cout, Cin)
            v = sd[prefix + "lin_a.weight_v"].float()
            return sd[prefix + "lin_a.weight_g"].float() * v / v.norm(dim=1, keepdim=True), add_b(sd[prefix + "lin_a.bias"])

        self.gemm = 0
        self.raw_mask = 0

        def masked_gemm(prefix, src, mask, raw, skip=None):
            """chunks of one masked 3x3 conv reading cached tensor `src`: gathered non-centre taps, [nin_skip of tensor
            skip[0] into the next 80 accumulator columns,] then the centre tap written by the previous epilogue"""
            wn, wc, b = conv_mats(prefix)
            cout, cin = wc.shape
            col = (self.gemm & 1) * 160
            add_chunks(swizzle_tiles(wn), cout, A_GATHER, src, mask, cin // 8, 20 if raw else 0, col, False, gemm=self.gemm)
            if raw:
                self.raw_mask |= 1 << src
            bskip = -1
            if skip is not None:
                ws, bskip = nin_mat(skip[1])
                add_chunks(swizzle_tiles(_pad_k(ws)), 80, A_CENTRE, skip[0], 0, 20, 0, col + 80, False, gemm=self.gemm)
            epi_first.append(len(chunks))
            add_chunks(swizzle_tiles(_pad_k(wc)), cout, A_TMEM, src, mask, cin // 8, 0, col, True, last_bar=self.gemm & 1,
                       gemm=self.gemm)
            self.gemm += 1
            return b, bskip

        w = sd["u_init.weight"].float()                       # (80, 513, 3, 3) -> [tap][cin][cout]
        self.w_uinit = w.permute(2, 3, 1, 0).reshape(9, 513, 80).contiguous().to(device=device, dtype=torch.float16)
        self.plan = _Plan()
        self.plan.b_uinit = add_b(sd["u_init.bias"])
        ops = []

        def resnet(prefix, og, a, mid, out):
            o = _Op()
            o.kind, o.og, o.a, o.mid, o.out = 0, og, a, mid, out
            o.b_in, o.b_skip = masked_gemm(prefix + "conv_input.", og, 1, False, None if a < 0 else (a, prefix + "nin_skip."))
            o.b_out, _ = masked_gemm(prefix + "conv_out.", mid, 1, False)
            ops.append(o)

        def dilated(prefix, src, dst):
            o = _Op()
            o.kind, o.og, o.a, o.mid, o.out = 1, src, -1, -1, dst
            o.b_in, _ = masked_gemm(prefix, src, 2, True)
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
        assert len(ops) == 18 and self.gemm == 32
        for i, o in enumerate(ops):
            self.plan.ops[i] = o
        self.plan.n_chunks_body = len(chunks)
        wno, self.plan.b_nin = nin_mat("nin_out.")            # (512, 80): four 128-class quarters, K padded to 128
        for q in range(4):
            epi_first.append(len(chunks))
            add_chunks(swizzle_tiles(_pad_k(wno[128 * q:128 * q + 128])), 128, A_EPILOGUE if q == 0 else A_REUSE, 30, 0, 10, 0,
                       128 * q, False, last_bar=2 if q == 3 else None, gemm=32 + q)
        self.plan.n_chunks_total = len(chunks)
        self.plan.raw_mask = self.raw_mask
        assert len(epi_first) <= MAX_GEMMS
        for i, v in enumerate(epi_first):
            self.plan.epi_first[i] = v
        self.wblob = torch.from_numpy(np.concatenate(blobs)).to(device)
        self.chunks = torch.from_numpy(np.frombuffer(b"".join(bytes(c) for c in chunks), dtype=np.uint8).copy()).to(device)
        self.bias = torch.cat(bs).to(device).contiguous()
        # ---- chain / halo split of the same chunk list (csrc/lmconv_tc.cu, sampled levels) ----
        def clone(c):
            d = _Chunk()
            ctypes.memmove(ctypes.byref(d), ctypes.byref(c), ctypes.sizeof(_Chunk))
            return d

        chain, halo, halo_first, part_col = [], [], [], []
        col = 0
        for g in range(32):
            mine = [c for c in chunks[:self.plan.n_chunks_body] if c.gemm == g]
            gath = [clone(c) for c in mine if c.a_kind == A_GATHER]
            halo_first.append(len(halo))
            gath[-1].flags |= 2 | ((g & 1) << 2)          # the last gathered chunk completes the halo accumulator
            halo += gath
            part_col.append(col)
            col += mine[0].w_rows                          # 80 or 160 partial-sum columns per row
            for c in mine:
                if c.a_kind == A_GATHER:
                    continue
                d = clone(c)
                if d.a_kind == A_TMEM and d.kc == 0:
                    d.flags &= ~1                          # nothing was accumulated before the centre tap any more
                chain.append(d)
        halo_first.append(len(halo))
        part_col.append(col)
        self.part_cols = col
        self.plan.n_chain_body = len(chain)
        chain += [clone(c) for c in chunks[self.plan.n_chunks_body:]]
        self.plan.n_chain_total = len(chain)
        assert self.plan.n_chain_body % 2 == 0 and self.plan.n_chain_total % 2 == 0 and all(h % 2 == 0 for h in halo_first)
        for i in range(33):
            self.plan.halo_first[i], self.plan.part_col[i] = halo_first[i], part_col[i]
        pack = lambda cs: torch.from_numpy(np.frombuffer(b"".join(bytes(c) for c in cs), dtype=np.uint8).copy()).to(device)
        self.chunks_chain, self.chunks_halo = pack(chain), pack(halo)
        self.plan.chunks_chain, self.plan.chunks_halo = self.chunks_chain.data_ptr(), self.chunks_halo.data_ptr()
        self.plan.wblob, self.plan.chunks = self.wblob.data_ptr(), self.chunks.data_ptr()
        self.plan.w_uinit, self.plan.bias = self.w_uinit.data_ptr(), self.bias.data_ptr()
        self._cache = None
        self._rows_pin = self._rows_evt = None
        self.last_levels = None

    @staticmethod
    def levels_host(order, words, sample_mask, mode):
        """ps_lmconv_levels_host: -> (rows uint8 array of 16-byte records, level offsets, index of the first level of
        phase B = sampled cells and their descendants; the levels before it are the known prefix)."""
        order = np.ascontiguousarray(np.asarray(order).reshape(-1, 1024).astype(np.int32))
        B = order.shape[0]
        words = np.ascontiguousarray(np.asarray(words).reshape(B, 3, 1024).astype(np.uint16))
        sm = None if sample_mask is None else np.ascontiguousarray(np.asarray(sample_mask).reshape(B, 1024).astype(np.uint8))
        rows = np.zeros((B * 1024, 16), np.uint8)
        offs = np.zeros(MAX_LEVELS + 1, np.int32)
        n, first_b = ctypes.c_int(0), ctypes.c_int(0)
        check(_lib.lib().ps_lmconv_levels_host(order.ctypes.data, words.ctypes.data, None if sm is None else sm.ctypes.data, B,
                                               int(mode), rows.ctypes.data, offs.ctypes.data, MAX_LEVELS, ctypes.byref(n),
                                               ctypes.byref(first_b)),
              "ps_lmconv_levels_host")
        offs = offs[:n.value + 1].copy() if n.value else np.zeros(1, np.int32)
        return rows[:int(offs[-1])], offs, first_b.value

    def prepare(self, order, words, sample_mask, mode=0):
        """Host half of a sampler call: dependency levels of (order, words, sample_mask) and the upload of the row
        records (pinned staging, asynchronous).  The result can be passed to sample()/logits() as `prepared`, so the
        host work overlaps whatever the GPU is doing in between."""
        wn = words.detach().cpu().numpy() if torch.is_tensor(words) else np.asarray(words)
        ordn = order.detach().cpu().numpy() if torch.is_tensor(order) else np.asarray(order)
        smn = None
        if sample_mask is not None:
            smn = sample_mask.detach().cpu().numpy() if torch.is_tensor(sample_mask) else np.asarray(sample_mask)
            smn = smn.astype(np.uint8)
        rows, offs, first_b = self.levels_host(ordn, wn, smn, mode)
        rows_d = None
        if len(offs) > 1:
            n = rows.shape[0]
            if self._rows_pin is None or self._rows_pin.shape[0] < n:
                self._rows_pin = torch.empty((max(n, 1024), 16), dtype=torch.uint8).pin_memory()
            if self._rows_evt is not None:
                self._rows_evt.synchronize()      # the previous upload has left the staging buffer
            self._rows_pin[:n].numpy()[:] = rows
            # every prepare() owns its device rows: two prepared calls can be in flight without sharing a buffer
            rows_d = torch.empty((n, 16), dtype=torch.uint8, device=self.device)
            rows_d.copy_(self._rows_pin[:n], non_blocking=True)
            self._rows_evt = torch.cuda.Event()
            self._rows_evt.record()
        # the kernel reads uniforms[b, k] for the k-th sampled cell of image b: the widest image sets the minimum width
        need_u = 0 if smn is None else int(smn.reshape(smn.shape[0], -1).astype(bool).sum(1).max(initial=0))
        return dict(rows=rows_d, offs=offs, first_b=first_b, mode=mode, need_uniforms=need_u,
                    batch=int(np.asarray(ordn).reshape(-1, 1024).shape[0]))

    def _run(self, codes, prepared, uniforms, temperature):
        dev = self.device
        B = codes.shape[0]
        offs, first_b, mode, rows_d = prepared["offs"], prepared["first_b"], prepared["mode"], prepared["rows"]
        if prepared.get("batch", B) != B:
            raise ValueError(f"prepared levels are for {prepared['batch']} images, codes hold {B}")
        self.last_levels, self.last_first_b = offs, first_b
        codes_d = torch.as_tensor(codes).to(device=dev, dtype=torch.int64).reshape(B, 1024).clone()
        logits = torch.empty((B, 1024, 512), dtype=torch.float32, device=dev) if mode == 1 else None
        if len(offs) > 1:
            uni_d = None
            if uniforms is not None:
                u = torch.as_tensor(uniforms)
                # an asynchronous copy from pageable memory may read its source after the caller has dropped it
                uni_d = u.to(device=dev, dtype=torch.float32, non_blocking=u.is_cuda or u.is_pinned()).reshape(B, -1).contiguous()
                if uni_d.shape[1] < prepared.get("need_uniforms", 0):
                    raise RuntimeError(f"uniforms: {uni_d.shape[1]} columns, but an image samples {prepared['need_uniforms']} "
                                     "cells (one uniform number per sampled cell, in generation order)")
            elif mode == 0:
                raise RuntimeError("sampling needs the uniform numbers")
            nbytes = _lib.lib().ps_lmconv_tc_cache_bytes(B)
            if self._cache is None or self._cache.numel() < nbytes:
                self._cache = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(self.wblob.device):
                check(_lib.lib().ps_lmconv_tc_run(
                    ctypes.byref(self.plan), B, rows_d.data_ptr(), offs.ctypes.data, len(offs) - 1, first_b, codes_d.data_ptr(),
                    None if uni_d is None else uni_d.data_ptr(), 0 if uni_d is None else uni_d.shape[1], float(temperature),
                    None if logits is None else logits.data_ptr(), self._cache.data_ptr(), nbytes,
                    torch.cuda.current_stream().cuda_stream), "ps_lmconv_tc_run")
        return codes_d.view(B, 32, 32), logits

    def sample(self, codes, order, words, sample_mask, uniforms, temperature=1.0, prepared=None):
        """codes (B,32,32) int64 with the known cells; returns codes with the sample_mask cells drawn in generation
        order (the argmax of sample.py's one-hot `data`, as z_buffermodel.py:249 takes it)."""
        if prepared is None:
            prepared = self.prepare(order, words, sample_mask, 0)
        out, _ = self._run(codes, prepared, uniforms, temperature)
        return out

    def logits(self, codes, order, words):
        """Teacher-forced logits of every cell given all codes: (B,512,32,32) like OurPixelCNN.forward."""
        B = codes.shape[0]
        _, lg = self._run(codes, self.prepare(order, words, None, 1), None, 1.0)
        return lg.view(B, 32, 32, 512).permute(0, 3, 1, 2).contiguous()
