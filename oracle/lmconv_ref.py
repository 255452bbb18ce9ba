"""CPU fp32 restatement of the locally-masked-convolution PixelCNN, its generation order / masks and its sampling
loop (TEST INFRASTRUCTURE ONLY -- the product never imports this module).

  order_from_distances   models/lmconv/get_custom_order.pyx:4-124 (custom_idx: frontier heap keyed (-distance, [r,c]))
  masks_from_order       models/lmconv/masking.py:287-349 (kernel_masks / get_unfolded_masks), as 9-bit words
  glue_from_background   models/z_buffermodel.py:641-701 (get_masks_for_batch: AvgPool8 -> uint8 -> cv2 distance
                         transforms -> int(fd - bd) -> order + three masks per image)
  lmconv_logits          models/lmconv/model.py:110-155 (OurPixelCNN.forward), layers.py:136-163 (gated_resnet),
                         :20-38 (nin), :224-243 (PONO), utils.py:31-35 (concat_elu),
                         locally_masked_convolution.py:11-50 (unfold . mask . matmul)
  sample_reference_style models/lmconv/sample.py:8-73 (one FULL forward per sampled token), device-agnostic, with
                         the categorical draw made explicit: token = first j with cumsum(softmax(l/T))_j > u.
                         (torch.multinomial's Philox stream cannot be reproduced outside torch's CUDA generator;
                         SURVEY.md section 0 fact 8.)
Pinned to the reference's own OurPixelCNN / masking code by tests/golden/make_lmconv_golden.py."""
import heapq

import numpy as np
import torch
import torch.nn.functional as F

TAPS = [(dr, dc) for dr in (-1, 0, 1) for dc in (-1, 0, 1)]  # tap t = (dr+1)*3 + (dc+1), as F.unfold orders them


def order_from_distances(distances):
    """distances (rows, rows) int -> (rows*rows, 2) generation order.  Start at the first maximum (row-major);
    repeatedly pop the frontier cell with the largest distance, ties to the lexicographically smallest [r, c];
    a cell enters the frontier when a 4-neighbour (Up, Down, Left, Right) is generated."""
    d = np.asarray(distances).astype(np.int64) * 10000
    rows = d.shape[0]
    am = int(np.argmax(d))
    c = am % rows
    r = (am - c) // rows
    order = [(r, c)]
    seen = np.zeros_like(d, dtype=bool)
    seen[r, c] = True
    heap = []
    while len(order) < rows * rows:
        for rr, cc in ((r - 1, c), (r + 1, c), (r, c - 1), (r, c + 1)):
            if 0 <= rr < rows and 0 <= cc < rows and not seen[rr, cc]:
                seen[rr, cc] = True
                heapq.heappush(heap, (-int(d[rr, cc]), rr, cc))
        _, r, c = heapq.heappop(heap)
        order.append((r, c))
    return np.array(order, dtype=np.int64)


def masks_from_order(order, rows=32, cols=32):
    """-> (3, rows*cols) uint16: nine-bit words [A dil 1, B dil 1, B dil 2], bit t set iff tap t may be read."""
    rank = np.full((rows, cols), -1, dtype=np.int64)
    for i, (r, c) in enumerate(order):
        rank[r, c] = i
    out = np.zeros((3, rows * cols), dtype=np.uint16)
    for r in range(rows):
        for c in range(cols):
            for mi, (dil, centre) in enumerate(((1, 0), (1, 1), (2, 1))):
                w = 0
                for t, (dr, dc) in enumerate(TAPS):
                    if dr == 0 and dc == 0:
                        w |= centre << t
                        continue
                    rr, cc = r + dr * dil, c + dc * dil
                    if 0 <= rr < rows and 0 <= cc < cols and rank[rr, cc] < rank[r, c]:
                        w |= 1 << t
                out[mi, r * cols + c] = w
    return out


def masks_to_float(words):
    """(L,) nine-bit words -> (1, 9, L) float mask in the layout get_unfolded_masks returns."""
    w = np.asarray(words).astype(np.int64)
    return torch.tensor(((w[None, :] >> np.arange(9)[:, None]) & 1).astype(np.float32))[None]


def glue_from_background(background_mask):
    """background_mask (B,S,S) bool at image resolution -> distances (B,32,32) int, orders (B,1024,2),
    mask words (B,3,1024) uint16, sample mask (B,32,32) bool (cells whose 64 pixels are all background)."""
    import cv2

    bg = F.avg_pool2d(background_mask.float()[:, None], 8)[:, 0]
    fg = F.avg_pool2d((~background_mask).float()[:, None], 8)[:, 0]
    bin_fg = fg.numpy().astype(np.uint8)
    bin_bg = bg.numpy().astype(np.uint8)
    B = bg.shape[0]
    dist = np.zeros((B, 32, 32), dtype=np.int64)
    orders, words = [], []
    for i in range(B):
        fd = cv2.distanceTransform(bin_fg[i], distanceType=cv2.DIST_L2, maskSize=5)
        bd = cv2.distanceTransform(bin_bg[i], distanceType=cv2.DIST_L2, maskSize=5)
        # OpenCV 4.2 (the reference's pin, docs/INSTALL.md:60) saturates a transform with no zero pixel at
        # (UINT_MAX - LONG_DIST) / 65536; newer builds return FLT_MAX there, which does not survive astype(int)
        sat = np.float32((0xFFFFFFFF - round(2.1969 * 65536)) / 65536.0)
        fd, bd = np.minimum(fd, sat), np.minimum(bd, sat)
        dist[i] = (fd.astype(np.float64) - bd.astype(np.float64)).astype(int)
        orders.append(order_from_distances(dist[i]))
        words.append(masks_from_order(orders[-1]))
    return dist, np.stack(orders), np.stack(words), torch.from_numpy(bin_bg.astype(bool))


def _masked_conv(x, mask, w, b, dil):
    B, Cin, H, W = x.shape
    unf = F.unfold(x, (3, 3), dilation=dil, padding=dil)           # (B, Cin*9, L)
    unf = (unf.view(B, Cin, 9, H * W) * mask[:, None]).view(B, Cin * 9, H * W)
    out = w.view(w.shape[0], -1) @ unf + b[None, :, None]
    return out.view(B, -1, H, W)


def _pono(x):
    mean = x.mean(dim=1, keepdim=True)
    std = x.var(dim=1, keepdim=True).add(1e-5).sqrt()   # unbiased variance, as torch.var defaults
    return (x - mean) / std


def _celu(x):
    return F.elu(torch.cat([x, -x], dim=1))


def _nin(sd, p, x):
    v = sd[p + "lin_a.weight_v"]
    w = sd[p + "lin_a.weight_g"] * v / v.norm(dim=1, keepdim=True)   # torch weight_norm, dim=0
    return torch.einsum("oc,bchw->bohw", w, x) + sd[p + "lin_a.bias"][None, :, None, None]


def _resnet(sd, p, og, a, mask, trace=None):
    x = _masked_conv(_celu(og), mask, sd[p + "conv_input.weight"], sd[p + "conv_input.bias"], 1)
    x = _pono(x)
    if a is not None:
        x = x + _nin(sd, p + "nin_skip.", _celu(a))
    if trace is not None:
        trace.append(x)
    y = _masked_conv(_celu(x), mask, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], 1)
    ya, yb = torch.chunk(y, 2, dim=1)
    return og + _pono(ya) * torch.sigmoid(yb)


def lmconv_logits(sd, data, m_init, m_undil, m_dil, trace=None):
    """data (B,512,32,32) one-hot (zeros at not-yet-generated cells); masks (B,9,1024) float -> logits (B,512,32,32).
    trace (a list) receives every intermediate tensor in execution order: u_init's output, then per gated resnet its
    mid and output tensors, per dilated conv its output."""
    def keep(t):
        if trace is not None:
            trace.append(t)
        return t

    x = torch.cat((data, torch.ones_like(data[:, :1])), 1)
    u_list = [keep(_pono(_masked_conv(x, m_init, sd["u_init.weight"], sd["u_init.bias"], 1)))]
    for i in range(2):
        for j in range(2):
            u_list.append(keep(_resnet(sd, f"up_layers.{i}.u_stream.{j}.", u_list[-1], None, m_undil, trace)))
        u_list.append(keep(_pono(_masked_conv(u_list[-1], m_dil, sd[f"downsize_u_stream.{i}.weight"],
                                              sd[f"downsize_u_stream.{i}.bias"], 2))))
    for j in range(2):
        u_list.append(keep(_resnet(sd, f"up_layers.2.u_stream.{j}.", u_list[-1], None, m_undil, trace)))
    u = u_list.pop()
    for i, n in enumerate((2, 3, 3)):
        for j in range(n):
            u = keep(_resnet(sd, f"down_layers.{i}.u_stream.{j}.", u, u_list.pop(), m_undil, trace))
        if i < 2:
            u = keep(_pono(_masked_conv(u, m_dil, sd[f"upsize_u_stream.{i}.weight"], sd[f"upsize_u_stream.{i}.bias"], 2)))
    return _nin(sd, "nin_out.", F.elu(u))


def draw(logits, temperature, u):
    """Explicit categorical draw: first j with cumsum(softmax(logits / T))_j > u (clamped to the last class)."""
    p = torch.softmax(logits.double() / temperature, -1)
    return int(min((torch.cumsum(p, -1) > u).float().argmax().item() if (torch.cumsum(p, -1) > u).any() else 511, 511))


def sample_reference_style(sd, codes, orders, words, sample_mask, uniforms, temperature, max_steps=None):
    """sample.py:8-73: cells of sample_mask are zeroed, then filled one per step in generation order, each step running
    the full network.  codes (B,32,32) int64; uniforms (B, steps).  Returns codes with the sampled cells filled and
    the logits used at each step (B, steps, 512)."""
    B = codes.shape[0]
    data = F.one_hot(codes, 512).permute(0, 3, 1, 2).float()
    m = [torch.cat([masks_to_float(words[b, k]) for b in range(B)]) for k in range(3)]
    seq = []
    for b in range(B):
        cells = [(int(r), int(c)) for r, c in orders[b] if sample_mask[b, r, c]]
        seq.append(cells)
        for r, c in cells:
            data[b, :, r, c] = 0
    steps = min(len(s) for s in seq) if max_steps is None else min(max_steps, min(len(s) for s in seq))
    out = codes.clone()
    used = torch.zeros(B, steps, 512)
    for k in range(steps):
        logits = lmconv_logits(sd, data, *m)
        for b in range(B):
            r, c = seq[b][k]
            used[b, k] = logits[b, :, r, c]
            tok = draw(logits[b, :, r, c], temperature, float(uniforms[b, k]))
            out[b, r, c] = tok
            data[b, :, r, c] = 0
            data[b, tok, r, c] = 1
    return out, used
