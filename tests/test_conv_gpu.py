"""GPU: the TMA + tcgen05 implicit-GEMM convolution against torch's fp32 convolution of the same bf16-rounded
operands.  Tolerance: products are exact in fp32, accumulation order differs and the output is rounded to bf16
(rel 2^-8): |err| <= 1e-2 * max|ref| for bf16 outputs, 2e-3 * max|ref| for the fp32 output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cv():
    import pixelsynth_b200.conv as c

    return c


def rb(x):
    return x.to(torch.bfloat16).float()


def close(a, ref, rel):
    err = (a - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= rel * scale + 1e-6, f"max err {err:.4g} vs scale {scale:.4g}"


@pytest.mark.parametrize("N,C,Cout,H,W,k,stride,pad", [
    (2, 64, 128, 32, 32, 3, 1, 1),
    (1, 128, 128, 48, 40, 3, 1, 1),    # ragged tiles
    (2, 4, 64, 32, 32, 3, 1, 1),       # tiny channel count (zero-filled K)
    (2, 256, 256, 16, 16, 3, 1, 1),    # BN = 256
    (3, 128, 64, 16, 16, 1, 1, 0),     # 1x1
    (2, 3, 32, 64, 64, 4, 2, 1),       # stride-2 4x4 (Unet conv1 / VQ-VAE encoder)
    (2, 64, 128, 32, 32, 4, 2, 1),
    (5, 256, 256, 4, 4, 4, 2, 1),      # small spatial, several images per tile
    (3, 256, 256, 2, 2, 3, 1, 1),
    (2, 512, 256, 8, 8, 3, 1, 1),      # Unet decoder, K = 9*512
    (1, 80, 24, 16, 16, 3, 1, 1),      # channel counts that are not multiples of 64 / 16
])
def test_conv2d(cv, N, C, Cout, H, W, k, stride, pad):
    g = torch.Generator().manual_seed(N * 1000 + C + H)
    x = rb(torch.randn(N, C, H, W, generator=g))
    w = rb(torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, stride=stride, padding=pad)
    xd = cv.to_nhwc_bf16(x.cuda())
    pc = cv.PackedConv.conv2d(w, b, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    out = torch.zeros((N, Ho, Wo, cv.round_up(Cout, 8)), dtype=torch.bfloat16, device="cuda")
    o32 = torch.zeros((N, Cout, Ho, Wo), dtype=torch.float32, device="cuda")
    cv.conv_igemm(xd, pc, [cv.Out(out)], stride=stride, Hout=Ho, Wout=Wo, out_f32=o32)
    torch.cuda.synchronize()
    close(o32.cpu(), ref, 2e-3)
    close(cv.from_nhwc(out, Cout).cpu(), ref, 1e-2)


def test_epilogue_dual_output_scale_shift_residual(cv):
    g = torch.Generator().manual_seed(7)
    N, C, Cout, H, W = 3, 64, 96, 16, 16
    x = rb(torch.randn(N, C, H, W, generator=g))
    w = rb(torch.randn(Cout, C, 3, 3, generator=g) / 24)
    b = torch.randn(Cout, generator=g)
    res = rb(torch.randn(N, Cout, H, W, generator=g))
    sc = torch.rand(N, Cout, generator=g) + 0.5
    sh = torch.randn(N, Cout, generator=g)
    sc2 = torch.rand(Cout, generator=g) + 0.5
    v = F.conv2d(x, w, b, padding=1) + res
    ref0 = F.relu(v * sc[:, :, None, None] + sh[:, :, None, None])
    ref1 = F.leaky_relu(v * sc2[None, :, None, None], 0.2)
    pc = cv.PackedConv.conv2d(w, b, padding=1)
    buf0 = torch.zeros((N, H, W, 96), dtype=torch.bfloat16, device="cuda")
    cat = torch.zeros((N, H, W, 256), dtype=torch.bfloat16, device="cuda")  # output 1 lands at channel 128 of a concat buffer
    cv.conv_igemm(cv.to_nhwc_bf16(x.cuda()), pc,
                  [cv.Out(buf0, "relu", sc.cuda(), sh.cuda(), per_sample=True),
                   cv.Out(cat, "leaky", sc2.cuda(), None, coffset=128)],
                  residual=cv.to_nhwc_bf16(res.cuda()))
    torch.cuda.synchronize()
    close(cv.from_nhwc(buf0, Cout).cpu(), ref0, 1e-2)
    close(cat[..., 128:128 + Cout].permute(0, 3, 1, 2).float().cpu(), ref1, 1e-2)
    assert (cat[..., :128] == 0).all() and (cat[..., 224:] == 0).all()


def test_fused_skip_conv(cv):
    """3x3 conv over h plus 1x1 conv over x accumulated in one tile (ResNet_Block ch_a + ch_b, blocks.py:55-73)."""
    g = torch.Generator().manual_seed(11)
    N, Ca, Cb, Cout, H, W = 2, 128, 64, 128, 32, 32
    h = rb(torch.randn(N, Ca, H, W, generator=g))
    x = rb(torch.randn(N, Cb, H, W, generator=g))
    wa = rb(torch.randn(Cout, Ca, 3, 3, generator=g) / 34)
    wb = rb(torch.randn(Cout, Cb, 1, 1, generator=g) / 8)
    ba, bb = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    ref = F.conv2d(h, wa, ba, padding=1) + F.conv2d(x, wb, bb)
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    wt = [wa[:, :, ky, kx] for ky in range(3) for kx in range(3)]
    wb_pad = torch.zeros(Cout, Ca)
    wb_pad[:, :Cb] = wb[:, :, 0, 0]
    pc = cv.PackedConv(wt + [wb_pad], taps + [(0, 0)], ba + bb)
    pc.taps = taps  # the 10th weight block belongs to the second input
    out = torch.zeros((N, H, W, Cout), dtype=torch.bfloat16, device="cuda")
    cv.conv_igemm(cv.to_nhwc_bf16(h.cuda()), pc, [cv.Out(out)], x2=cv.to_nhwc_bf16(x.cuda()), pc2_taps=[(0, 0)],
                  pc2_wrow=[9 * pc.cout_pad])
    torch.cuda.synchronize()
    close(cv.from_nhwc(out, Cout).cpu(), ref, 1e-2)


@pytest.mark.parametrize("N,C,Cout,H", [(2, 64, 64, 16), (1, 128, 64, 32), (2, 64, 3, 32)])
def test_conv_transpose_4x4_s2(cv, N, C, Cout, H):
    g = torch.Generator().manual_seed(C + Cout)
    x = rb(torch.randn(N, C, H, H, generator=g))
    w = rb(torch.randn(C, Cout, 4, 4, generator=g) / (4 * C) ** 0.5)
    b = torch.randn(Cout, generator=g)
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    xd = cv.to_nhwc_bf16(x.cuda())
    out = torch.zeros((N, 2 * H, 2 * H, cv.round_up(Cout, 8)), dtype=torch.bfloat16, device="cuda")
    for py in range(2):
        for px in range(2):
            pc = cv.PackedConv.conv_transpose_4x4_s2_phase(w, py, px, b)
            cv.conv_igemm(xd, pc, [cv.Out(out)], Hout=H, Wout=H, geometry=(2 * H, 2 * H, 2, 2, py, px))
    torch.cuda.synchronize()
    close(cv.from_nhwc(out, Cout).cpu(), ref, 1e-2)


def test_activations_fp32_only(cv):
    g = torch.Generator().manual_seed(3)
    x = rb(torch.randn(2, 64, 16, 16, generator=g))
    w = rb(torch.randn(1, 64, 3, 3, generator=g) / 24)
    ref = torch.sigmoid(F.conv2d(x, w, None, padding=1)) * 9.5 + 0.5   # depth head: sigmoid*(max_z-min_z)+min_z
    pc = cv.PackedConv.conv2d(w, None, padding=1)
    o32 = torch.zeros((2, 1, 16, 16), dtype=torch.float32, device="cuda")
    cv.conv_igemm(cv.to_nhwc_bf16(x.cuda()), pc, [cv.Out(None, "sigmoid_affine")], out_f32=o32, act_param=(9.5, 0.5))
    torch.cuda.synchronize()
    close(o32.cpu(), ref, 2e-3)


@pytest.mark.parametrize("mode,shape", [("bilinear", (2, 5, 7, 16)), ("bilinear", (3, 16, 16, 64)), ("bilinear", (1, 1, 1, 8)),
                                        ("avgpool", (2, 9, 6, 16)), ("avgpool_nopad", (2, 8, 8, 8)), ("maxpool", (1, 7, 7, 8))])
def test_resample_modes_against_torch(mode, shape):
    """ps_resample (nn.Upsample(bilinear) / nn.AvgPool2d(3,2,1) of blocks.py:46-48, the discriminator's and resnet18's
    pools) against torch on the same bf16 inputs, fp32 arithmetic, one bf16 rounding at the end; the second output
    carries the per-sample affine + ReLU the decoder fuses here."""
    import torch.nn.functional as F
    from pixelsynth_b200 import nets
    from pixelsynth_b200.conv import Out

    n, h, w, c = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).cuda()
    scale = (torch.rand(n, c, generator=g) + 0.5).cuda()
    shift = torch.randn(n, c, generator=g).cuda()
    xf = x.float().permute(0, 3, 1, 2)
    if mode == "bilinear":
        ref = F.interpolate(xf, scale_factor=2, mode="bilinear", align_corners=False)
    elif mode == "avgpool":
        ref = F.avg_pool2d(xf, 3, 2, 1)
    elif mode == "avgpool_nopad":
        ref = F.avg_pool2d(xf, 3, 2, 1, count_include_pad=False)
    else:
        ref = F.max_pool2d(xf, 3, 2, 1)
    ho, wo = ref.shape[2:]
    o0 = torch.empty(n, ho, wo, c, dtype=torch.bfloat16, device="cuda")
    o1 = torch.empty_like(o0)
    nets.resample(x, mode, Out(o0), Out(o1, "relu", scale, shift, per_sample=True))
    ref1 = torch.relu(ref * scale[:, :, None, None] + shift[:, :, None, None])
    for got, want in ((o0, ref), (o1, ref1)):
        got = got.float().permute(0, 3, 1, 2)
        err = (got - want).abs()
        assert (err <= 2.0 ** -7 * want.abs() + 1e-6).all(), float(err.max())   # within one bf16 rounding of fp32
