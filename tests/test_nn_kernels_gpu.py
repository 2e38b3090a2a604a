"""GPU parity of the dense kernels (through the C ABI) against plain fp32 PyTorch on the same fp16-rounded inputs.

Tolerances: operands are fp16 with fp32 accumulation, so against an fp32 reference computed from the SAME
fp16-rounded inputs the only error is the final fp16 rounding of the output (2^-11 relative) plus accumulation
order; relative-L2 bounds of 1e-3 are used throughout.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rnd(*shape, dev, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().to(dev)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 320, 320), (300, 200, 136), (4096, 1280, 2560),
                                   (77, 640, 1024), (5, 64, 72),
                                   # 128 x 256 tiles (N % 256 == 0, K >= 1024, >= 148 tiles): ragged M, K tail, several
                                   # tiles per CTA; and the same N / K below the tile-count threshold
                                   (19000, 512, 1160), (40000, 256, 1088), (6400, 768, 1024), (300, 512, 1024),
                                   (256, 1280, 11520)])
def test_gemm_plain(cuda_device, M, N, K):
    from scaledreamer_b200 import nn_ops as O

    a, b = rnd(M, K, dev=cuda_device, seed=1), rnd(N, K, dev=cuda_device, seed=2)
    out = O.gemm(a, b)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T
    assert rel(out, ref) < 1e-3


@pytest.mark.parametrize("M,N,K", [(512, 320, 640), (19072, 256, 1152)])
def test_gemm_epilogue_and_fp32_out(cuda_device, M, N, K):
    from scaledreamer_b200 import nn_ops as O

    a, b = rnd(M, K, dev=cuda_device, seed=1, scale=0.5), rnd(N, K, dev=cuda_device, seed=2, scale=0.1)
    bias, res = rnd(N, dev=cuda_device, seed=3), rnd(M, N, dev=cuda_device, seed=4)
    rowbias = torch.randn(M // 128, N, device=cuda_device)
    ref = a.float() @ b.float().T * 0.7 + bias.float() + rowbias.repeat_interleave(128, 0)
    out = O.gemm(a, b, bias=bias, rowbias=rowbias, rows_per_group=128, residual=res, alpha=0.7, act="silu", out_fp32=True)
    assert rel(out, F.silu(ref) + res.float()) < 1e-4
    out = O.gemm(a, b, bias=bias, act="gelu")
    assert rel(out, F.gelu(a.float() @ b.float().T + bias.float())) < 1e-3


@pytest.mark.parametrize("M,N,K", [(4096, 512, 4608),     # 128 tiles of 72 k-blocks: fewer tiles than CTA slots
                                   (20480, 320, 2880),    # 320 tiles on 296 slots: a nearly empty second wave
                                   (1280, 1280, 11520),   # 80 tiles, long K: several CTAs per tile
                                   (5000, 600, 2000),     # ragged M, N and K tails
                                   (320, 1280, 11520)])   # 24 tiles cut into 296 pieces: ~12 contributors per tile
def test_gemm_uneven_tile_counts(cuda_device, M, N, K):
    """Shapes whose tiles fill the CTA slots unevenly (fewer tiles than slots, a nearly empty last wave, long K with few
    tiles -> split-K, ragged tails): the persistent tile walk, both accumulator buffers, the split-K planes and the TMA
    store epilogue. Results match fp32 torch, are bitwise reproducible, and every epilogue feature still applies. (These
    are also the shapes the stream-K experiment of round 2 was measured on, profiles/r2h_*.)"""
    from scaledreamer_b200 import nn_ops as O

    a, b = rnd(M, K, dev=cuda_device, seed=1, scale=0.5), rnd(N, K, dev=cuda_device, seed=2, scale=0.1)
    bias, res = rnd(N, dev=cuda_device, seed=3), rnd(M, N, dev=cuda_device, seed=4)
    ref = a.float() @ b.float().T
    out = O.gemm(a, b)
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-3
    for _ in range(3):
        assert torch.equal(O.gemm(a, b), out)
    out2 = O.gemm(a, b, bias=bias, residual=res, alpha=0.7, act="silu", out_fp32=True)
    assert rel(out2, F.silu(ref * 0.7 + bias.float()) + res.float()) < 1e-4
    # back-to-back launches
    outs = [O.gemm(a, b, bias=bias) for _ in range(4)]
    torch.cuda.synchronize()
    assert all(torch.equal(o, outs[0]) for o in outs) and rel(outs[0], ref + bias.float()) < 1e-3


def test_conv3x3_fewer_tiles_than_slots(cuda_device):
    from scaledreamer_b200 import nn_ops as O

    x = rnd(5, 32, 32, 640, dev=cuda_device, seed=1)
    w = rnd(640, 3, 3, 640, dev=cuda_device, seed=2, scale=0.02)
    bias = rnd(640, dev=cuda_device, seed=3)
    out = O.conv3x3(x, w, bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias.float(), padding=1).permute(0, 2, 3, 1)
    assert rel(out, ref) < 1e-3
    assert torch.equal(O.conv3x3(x, w, bias=bias), out)


def test_gemm_batched(cuda_device):
    from scaledreamer_b200 import nn_ops as O

    a, b = rnd(6, 200, 128, dev=cuda_device, seed=1), rnd(6, 96, 128, dev=cuda_device, seed=2)
    out = O.gemm(a, b)
    assert rel(out, torch.einsum("zmk,znk->zmn", a.float(), b.float())) < 1e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(1, 16, 16, 64, 128), (2, 32, 32, 128, 320), (5, 8, 8, 320, 320),
                                            (3, 4, 4, 128, 64), (1, 128, 128, 64, 128), (1, 256, 256, 128, 128),
                                            (1, 256, 256, 128, 256), (2, 16, 16, 256, 512), (5, 8, 8, 1280, 1280)])
def test_conv3x3(cuda_device, N, H, W, Cin, Cout):
    from scaledreamer_b200 import nn_ops as O

    x = rnd(N, H, W, Cin, dev=cuda_device, seed=1)
    w = rnd(Cout, 3, 3, Cin, dev=cuda_device, seed=2, scale=1 / math.sqrt(9 * Cin))
    bias = rnd(Cout, dev=cuda_device, seed=3)
    rowbias = torch.randn(N, Cout, device=cuda_device)
    res = rnd(N, H, W, Cout, dev=cuda_device, seed=5)
    out = O.conv3x3(x, w, bias=bias, rowbias=rowbias, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias.float(), padding=1)
    ref = (ref + rowbias[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    assert rel(out, ref) < 1e-3


@pytest.mark.parametrize("cin,cout,fp32in", [(3, 128, True), (4, 320, False), (8, 512, False), (320, 4, False),
                                             (512, 8, False), (128, 3, False)])
def test_conv3x3_small(cuda_device, cin, cout, fp32in):
    from scaledreamer_b200 import nn_ops as O

    x = rnd(2, 12, 21, cin, dev=cuda_device, seed=1)  # width not a multiple of the 2 / 4 pixels a thread owns
    if fp32in:
        x = x.float()
    w = rnd(cout, 3, 3, cin, dev=cuda_device, seed=2, scale=0.1)
    bias = rnd(cout, dev=cuda_device, seed=3)
    out = O.conv3x3_small(x, w, bias, out_fp32=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias.float(), padding=1)
    assert rel(out, ref.permute(0, 2, 3, 1)) < 1e-4


@pytest.mark.parametrize("H,W,cout", [(363, 365, 3), (363, 365, 8), (725, 727, 3), (725, 727, 4)])
def test_conv3x3_small_cout_large_images(cuda_device, H, W, cout):
    """The thread-per-2/4-pixels kernels that take over from the warp-per-pixel one above 128 k / 512 k pixels
    (VAE conv_in data gradient at 512 x 512); odd widths leave a ragged last pixel group."""
    from scaledreamer_b200 import nn_ops as O

    x = rnd(1, H, W, 16, dev=cuda_device, seed=1)
    w = rnd(cout, 3, 3, 16, dev=cuda_device, seed=2, scale=0.1)
    bias = rnd(cout, dev=cuda_device, seed=3)
    out = O.conv3x3_small(x, w, bias, out_fp32=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias.float(), padding=1)
    assert rel(out, ref.permute(0, 2, 3, 1)) < 1e-4


@pytest.mark.parametrize("B,heads,Lq,Lk", [(2, 5, 256, 256), (3, 10, 64, 77), (1, 1, 128, 1024), (2, 20, 16, 77),
                                           (1, 5, 4096, 4096)])
def test_attention(cuda_device, B, heads, Lq, Lk):
    from scaledreamer_b200 import nn_ops as O

    q = rnd(B, Lq, heads * 64, dev=cuda_device, seed=1)
    k = rnd(B, Lk, heads * 64, dev=cuda_device, seed=2)
    v = rnd(B, Lk, heads * 64, dev=cuda_device, seed=3)
    out = O.attention(q, k, v, heads)
    sp = lambda t, L: t.float().view(B, L, heads, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q, Lq), sp(k, Lk), sp(v, Lk)).transpose(1, 2).reshape(B, Lq, heads * 64)
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("B,heads,Lq,Lk", [(2, 5, 256, 256), (1, 5, 4096, 4096), (2, 20, 64, 64), (3, 10, 64, 77),
                                           (1, 3, 200, 333), (1, 10, 1024, 1024)])
def test_flash_attention(cuda_device, B, heads, Lq, Lk):
    """Fused attention against SDPA, with q/k/v taken as column slices of one fused projection (as the UNet does),
    ragged lengths (not multiples of the 128-row tiles) and the single-block case."""
    from scaledreamer_b200 import nn_ops as O

    L = max(Lq, Lk)
    qkv = rnd(B, L, 3 * heads * 64, dev=cuda_device, seed=7)
    if Lq == Lk:
        q, k, v = qkv[..., :heads * 64], qkv[..., heads * 64:2 * heads * 64], qkv[..., 2 * heads * 64:]
    else:
        q = rnd(B, Lq, heads * 64, dev=cuda_device, seed=1)
        k = rnd(B, Lk, heads * 64, dev=cuda_device, seed=2)
        v = rnd(B, Lk, heads * 64, dev=cuda_device, seed=3)
    out = O.flash_attention(q, k, v, heads)
    sp = lambda t, Ln: t.float().reshape(B, Ln, heads, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q, Lq), sp(k, Lk), sp(v, Lk)).transpose(1, 2).reshape(B, Lq, heads * 64)
    assert rel(out, ref) < 2e-3
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("N,HW,C,silu", [(2, 64, 320, True), (1, 4096, 128, True), (3, 256, 1920, False),
                                         (2, 16, 2560, True), (1, 1024, 512, True)])
def test_groupnorm_forward_backward(cuda_device, N, HW, C, silu):
    from scaledreamer_b200 import nn_ops as O

    x = (rnd(N, HW, C, dev=cuda_device, seed=1) + 0.3)
    gamma, beta = rnd(C, dev=cuda_device, seed=2) * 0.2 + 1, rnd(C, dev=cuda_device, seed=3) * 0.2
    y, stats = O.groupnorm(x, gamma, beta, 32, 1e-5, silu)
    xr = x.float().transpose(1, 2).requires_grad_(True)
    ref = F.group_norm(xr, 32, gamma.float(), beta.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    assert rel(y, ref.transpose(1, 2)) < 1e-3
    dy = rnd(N, HW, C, dev=cuda_device, seed=4)
    ref.backward(dy.float().transpose(1, 2))
    dx = O.groupnorm_backward(x, gamma, beta, stats, dy, 32, 1e-5, silu)
    assert rel(dx, xr.grad.transpose(1, 2)) < 2e-3


def test_layernorm_geglu_upsample(cuda_device):
    from scaledreamer_b200 import nn_ops as O

    # register-resident rows (C <= 1280, incl. widths that leave lanes idle) and the strided fallback (wider / C % 8 != 0)
    for C in (640, 320, 1280, 8, 264, 2560, 1284):
        x = rnd(301, C, dev=cuda_device, seed=1) + 0.5
        g, b = rnd(C, dev=cuda_device, seed=2), rnd(C, dev=cuda_device, seed=3)
        assert rel(O.layernorm(x, g, b), F.layer_norm(x.float(), (C,), g.float(), b.float())) < 1e-3, C
    xg = rnd(100, 2 * 1280, dev=cuda_device, seed=4)
    a, gate = xg.float().chunk(2, -1)
    assert rel(O.geglu(xg), a * F.gelu(gate)) < 1e-3
    im = rnd(2, 8, 8, 64, dev=cuda_device, seed=5)
    up = F.interpolate(im.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(O.upsample2x(im).float(), up)


@pytest.mark.parametrize("pad_lo", [0, 1])
def test_strided_conv_via_im2col_and_its_dgrad(cuda_device, pad_lo):
    from scaledreamer_b200 import nn_ops as O

    N, H, W, Cin, Cout = 2, 16, 16, 64, 128
    x = rnd(N, H, W, Cin, dev=cuda_device, seed=1)
    w = rnd(Cout, 3, 3, Cin, dev=cuda_device, seed=2, scale=0.05)
    col = O.im2col3x3s2(x, pad_lo)
    out = O.gemm(col.view(-1, 9 * Cin), w.view(Cout, -1)).view(N, H // 2, W // 2, Cout)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    xp = xr if pad_lo else F.pad(xr, (0, 1, 0, 1))
    ref = F.conv2d(xp, w.float().permute(0, 3, 1, 2), stride=2, padding=pad_lo)
    assert rel(out, ref.permute(0, 2, 3, 1)) < 1e-3
    dy = rnd(N, H // 2, W // 2, Cout, dev=cuda_device, seed=3)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    wt = w.view(Cout, 9 * Cin).t().contiguous()  # [9*Cin, Cout]
    dcol = O.gemm(dy.view(-1, Cout), wt).view(N, H // 2, W // 2, 9 * Cin)
    dx = O.col2im3x3s2(dcol, H, W, pad_lo)
    assert rel(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-3
