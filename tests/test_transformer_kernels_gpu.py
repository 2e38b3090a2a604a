"""fp32 / tf32 kernels of the native Triplane-Transformer against plain torch fp32 (float64 for the GEMM) on the same
seeded inputs. tf32 keeps 10 mantissa bits of each operand: products are compared at 2e-3 of the result's scale
(|a|.|b| accumulated), everything else is fp32 arithmetic and compared at 1e-5."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from scaledreamer_b200 import transformer_ops as T
    return T


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 200, 96), (3072, 768, 768), (77, 48, 1024), (3072, 48, 77),
                                   (256, 3072, 48), (1, 5, 3)])
def test_gemm_tf32_plain(M, N, K):
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    lda, ldb, ldc = (K + 3) // 4 * 4, (K + 3) // 4 * 4, (N + 3) // 4 * 4
    A = torch.randn(M, lda, device="cuda", generator=g)
    B = torch.randn(N, ldb, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, ldc, device="cuda", generator=g)
    out = torch.full((M, ldc), float("nan"), device="cuda")
    T.gemm(T.Operand(A, lda), T.Operand(B, ldb), M, N, K, T.Operand(out, ldc), bias=bias, residual=T.Operand(res, ldc),
           alpha=0.5, act=T.ACT_GELU)
    ref = torch.nn.functional.gelu(0.5 * (A[:, :K].double() @ B[:, :K].double().T) + bias.double()) + res[:, :N].double()
    assert _rel(out[:, :N], ref) < 2e-3
    if ldc > N:
        assert torch.isnan(out[:, N:]).all()  # padding columns are not touched


def test_gemm_tf32_heads_batched():
    """The attention products: per-(prompt, head) views of [B, L, heads*d] activations, K = d = 48 (one and a half k-blocks),
    scores written as [B*heads][Lq][lds]."""
    T = _ops()
    Bn, H, d, Lq, Lk = 2, 4, 48, 200, 77
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(Bn, Lq, H * d, device="cuda", generator=g)
    k = torch.randn(Bn, Lk, H * d, device="cuda", generator=g)
    lds = 80
    S = torch.zeros(Bn * H, Lq, lds, device="cuda")
    T.gemm(T.Operand(q, H * d, Lq * H * d, d), T.Operand(k, H * d, Lk * H * d, d), Lq, Lk, d,
           T.Operand(S, lds, H * Lq * lds, Lq * lds), batch=Bn * H, zdiv=H, alpha=0.25)
    ref = 0.25 * torch.einsum("bqhd,bkhd->bhqk", q.view(Bn, Lq, H, d).double(), k.view(Bn, Lk, H, d).double())
    assert _rel(S[:, :, :Lk].view(Bn, H, Lq, Lk), ref) < 2e-3
    # P V with V^T produced by the operand-swapped projection: out written back in the head-interleaved layout
    vT = torch.randn(Bn, H * d, lds, device="cuda", generator=g)
    O = torch.zeros(Bn, Lq, H * d, device="cuda")
    T.gemm(T.Operand(S, lds, H * Lq * lds, Lq * lds), T.Operand(vT, lds, H * d * lds, d * lds), Lq, d, Lk,
           T.Operand(O, H * d, Lq * H * d, d), batch=Bn * H, zdiv=H)
    ref_o = torch.einsum("bhqk,bhdk->bqhd", S[:, :, :Lk].view(Bn, H, Lq, Lk).double(),
                         vT[:, :, :Lk].view(Bn, H, d, Lk).double()).reshape(Bn, Lq, H * d)
    assert _rel(O, ref_o) < 2e-3


@pytest.mark.parametrize("M,N,K,Z", [(128, 48, 64, 1), (200, 48, 300, 3), (77, 48, 3072, 2), (3072, 48, 77, 2)])
def test_gemm_tf32_mn_major_a(M, N, K, Z):
    """A handed over transposed ([K rows][M], M contiguous): out = A^T-free product through MN-major shared-memory tiles
    (dV = P^T dO, dK = dS^T Q of the attention backward read P / dS [q][k] as they are)."""
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    ldm = (M + 3) // 4 * 4
    At = torch.randn(Z, K, ldm, device="cuda", generator=g)  # A^T: [K][M]
    B = torch.randn(Z, N, (K + 3) // 4 * 4, device="cuda", generator=g)
    out = torch.zeros(Z, M, N, device="cuda")
    T.gemm(T.Operand(At, ldm, K * ldm), T.Operand(B, B.shape[-1], N * B.shape[-1]), M, N, K, T.Operand(out, N, M * N),
           batch=Z, a_mn_major=True)
    ref = torch.einsum("zkm,znk->zmn", At[:, :, :M].double(), B[:, :, :K].double())
    assert _rel(out, ref) < 2e-3


def test_gemm_tf32_shared_weight_batched_output():
    """V^T[b] = W_v ctx[b]^T: A shared by the batch, B batched, one output matrix per prompt."""
    T = _ops()
    Bn, Lk, Cin, Co = 3, 77, 1024, 768
    g = torch.Generator(device="cuda").manual_seed(6)
    W = torch.randn(Co, Cin, device="cuda", generator=g) * 0.03
    ctx = torch.randn(Bn, Lk, Cin, device="cuda", generator=g)
    ldo = 80
    out = torch.zeros(Bn, Co, ldo, device="cuda")
    T.gemm(T.mat(W), T.Operand(ctx, Cin, Lk * Cin), Co, Lk, Cin, T.Operand(out, ldo, Co * ldo), batch=Bn)
    ref = torch.einsum("oc,bkc->bok", W.double(), ctx.double())
    assert _rel(out[:, :, :Lk], ref) < 2e-3


def test_transpose_layernorm_softmax_gelu_colsum_shuffle():
    T = _ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(3, 70, 45, device="cuda", generator=g)
    xt = T.transpose(x, 70, 45, batch=3, round_out=False)
    assert xt.shape == (3, 45, 72) and torch.equal(xt[:, :, :70], x.transpose(1, 2))

    C = 768
    h = torch.randn(1000, C, device="cuda", generator=g) * 2 + 0.3
    gamma, beta = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    y, mean, rstd = T.layernorm_forward(h, gamma, beta, 1e-6)
    hr = h.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(hr, (C,), gr, br, 1e-6)
    assert _rel(y, yr) < 1e-5
    dy = torch.randn_like(y)
    skip = torch.randn_like(y)
    yr.backward(dy)
    dx, dg, db = T.layernorm_backward(h, gamma, mean, rstd, dy, skip)
    assert _rel(dx, hr.grad + skip) < 1e-5 and _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4

    for rows, cols, ld in ((64, 3072, 3072), (100, 77, 80), (9, 3073, 3076)):
        s = torch.randn(rows, ld, device="cuda", generator=g) * 3
        ref = torch.softmax(s[:, :cols], -1)
        ref_lse = torch.logsumexp(s[:, :cols], -1)
        p = s.clone()
        lse = T.softmax_forward_(p, rows, cols, ld)
        assert _rel(p[:, :cols], ref) < 1e-5 and _rel(lse, ref_lse) < 1e-5
        # statistics form of the backward, row-indexed
        dP = torch.randn(rows, ld, device="cuda", generator=g)
        delta = (dP[:, :cols] * ref).sum(-1)
        X, Y = s.clone(), dP.clone()
        T.softmax_backward_stats_(X, Y, 1, rows, cols, ld, lse, delta.contiguous(), by_col=False)
        ref_ds = ref * (dP[:, :cols] - delta[:, None])
        assert _rel(X[:, :cols], ref) < 1e-5 and _rel(Y[:, :cols], ref_ds) < 1e-4
        X, Y = s.clone(), dP.clone()
        dl = T.softmax_backward_rows_(X, Y, rows, cols, ld, lse)
        assert _rel(dl, delta) < 1e-5 and _rel(X[:, :cols], ref) < 1e-5 and _rel(Y[:, :cols], ref_ds) < 1e-4
        assert float(Y[:, :cols].sum(-1).abs().max()) < 1e-5  # rows of dS sum to zero
        X2, Y2 = s.clone(), dP.clone()
        T.softmax_backward_rows_(X2, Y2, rows, cols, ld, lse, write_p=False)
        assert torch.equal(X2, s) and torch.equal(Y2, Y)
    # column-indexed statistics on the transposed scores
    Z, R, Cc = 2, 40, 64
    s = torch.randn(Z, R, Cc, device="cuda", generator=g)
    ref = torch.softmax(s, -1)
    lse = torch.logsumexp(s, -1)
    dP = torch.randn(Z, R, Cc, device="cuda", generator=g)
    delta = (dP * ref).sum(-1)
    XT, YT = s.transpose(1, 2).contiguous(), dP.transpose(1, 2).contiguous()
    T.softmax_backward_stats_(XT, YT, Z, Cc, R, R, lse.contiguous(), delta.contiguous(), by_col=True)
    assert _rel(XT, ref.transpose(1, 2)) < 1e-5 and _rel(YT, (ref * (dP - delta[..., None])).transpose(1, 2)) < 1e-4

    hh = torch.randn(128, 3072, device="cuda", generator=g)
    hg = hh.clone().requires_grad_(True)
    gref = torch.nn.functional.gelu(hg)
    assert _rel(T.gelu_forward(hh), gref) < 1e-6
    dgo = torch.randn_like(hh)
    gref.backward(dgo)
    assert _rel(T.gelu_backward_(hh, dgo.clone()), hg.grad) < 1e-5

    m = torch.randn(12288, 96, device="cuda", generator=g)
    assert _rel(T.colsum(m, 12288, 96), m.double().sum(0)) < 1e-5
    assert _rel(T.colsum(m[:3].contiguous(), 3, 96), m[:3].double().sum(0)) < 1e-6

    # nearest-tf32 rounding: 10 mantissa bits, ties away from zero; transposes round on the way by default
    r = T.round_tf32(m)
    assert (r.view(torch.int32) & 0x1FFF).eq(0).all() and float((r - m).abs().max() / m.abs().max()) < 2.0 ** -11
    assert torch.equal(T.transpose(x, 70, 45, batch=3)[:, :, :70], T.round_tf32(x).transpose(1, 2))

    pe = torch.randn(50, 64, device="cuda", generator=g)
    assert torch.equal(T.broadcast(pe, 3), pe.expand(3, 50, 64))

    planes, H, W, D = 6, 8, 8, 32
    t = torch.randn(planes * H * W, 4 * D, device="cuda", generator=g)
    p = T.deconv_shuffle(t, planes, H, W, D, inverse=False)
    ref_p = t.view(planes, H, W, D, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(planes, 2 * H, 2 * W, D)
    assert torch.equal(p, ref_p)
    assert torch.equal(T.deconv_shuffle(p, planes, H, W, D, inverse=True), t)

    Bn, Lq, H2, d = 2, 33, 4, 48
    dO, O = torch.randn(Bn, Lq, H2 * d, device="cuda", generator=g), torch.randn(Bn, Lq, H2 * d, device="cuda", generator=g)
    dl = T.attn_delta(dO, O, Bn, Lq, H2, d)
    ref_dl = (dO * O).view(Bn, Lq, H2, d).sum(-1).permute(0, 2, 1).reshape(-1)
    assert _rel(dl, ref_dl) < 1e-5
