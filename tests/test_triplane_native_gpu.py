"""The native Triplane-Transformer (triplane_native.py: tf32 tensor-core GEMMs + fp32 kernels) against the same network
in plain torch fp32 (TriplaneTransformer.forward_torch, itself pinned to the reference module by
tests/test_amortized_golden_cpu.py): planes and EVERY parameter gradient. Tolerance: planes within north_star's 1e-3
relative L2; gradients within 3e-3, or twice what torch's OWN tf32 path (allow_tf32 = True: cuBLAS tf32 GEMMs on the same
tensors) deviates from fp32 on that tensor, whichever is larger. Measured: planes 5.3e-4 / 6.1e-4, worst gradient
1.1e-3 / 2.0e-3 (query / key projections, where softmax gradients cancel) against 9.5e-4 / 1.5e-3 for torch-tf32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _build(cfg, dev, seed=0):
    from scaledreamer_b200.amortized import TriplaneTransformer

    torch.manual_seed(seed)
    gen = TriplaneTransformer(**cfg).to(dev)
    with torch.no_grad():  # non-trivial norms / biases so their gradients are exercised
        for n, p in gen.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
    return gen


def _compare(cfg, n_prompts, dev, tol_planes, tol_grad):
    gen = _build(cfg, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    emb = torch.randn(n_prompts, 77, cfg["condition_dim"], device=dev, generator=g)
    low = cfg["triplane_low_res"]
    w = torch.randn(n_prompts, 3, cfg["triplane_dim"], 2 * low, 2 * low, device=dev, generator=g)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = gen.forward_torch(emb)
        (ref * w).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    ref_grads = {n: p.grad.clone() for n, p in gen.named_parameters()}
    for p in gen.parameters():
        p.grad = None
    torch.backends.cuda.matmul.allow_tf32 = True  # calibration: how far cuBLAS tf32 is from fp32 on these tensors
    try:
        (gen.forward_torch(emb) * w).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    tf32_err = {n: _rel(p.grad, ref_grads[n]) for n, p in gen.named_parameters()}
    for p in gen.parameters():
        p.grad = None
    from scaledreamer_b200 import lib as L

    n0 = L.launch_count()
    out = gen(emb)
    assert out.shape == ref.shape
    # the planes come out channels-last: the consumer's permute(0, 1, 3, 4, 2).contiguous() is a no-op
    assert out.permute(0, 1, 3, 4, 2).is_contiguous()
    (out * w).sum().backward()
    assert L.launch_count() - n0 > 50 * cfg["num_layers"]
    e = _rel(out, ref)
    print(f"planes rel_l2 {e:.2e}")
    assert e < tol_planes
    worst = ("", 0.0, 0.0)
    for n, p in gen.named_parameters():
        assert p.grad is not None, n
        ge = _rel(p.grad, ref_grads[n])
        if ge > worst[1]:
            worst = (n, ge, tf32_err[n])
        assert ge < max(tol_grad, 2.0 * tf32_err[n]), (n, ge, tf32_err[n])
    print(f"worst gradient {worst[0]} rel_l2 {worst[1]:.2e} (torch tf32 on the same tensor: {worst[2]:.2e}; "
          f"largest torch-tf32 error {max(tf32_err.values()):.2e})")


def test_small_generator_forward_backward(cuda_device):
    cfg = {"inner_dim": 128, "condition_dim": 1024, "triplane_low_res": 8, "triplane_high_res": 16, "triplane_dim": 32,
           "num_layers": 2, "num_heads": 4, "flash_attention": False, "local_text": True}
    _compare(cfg, 2, cuda_device, 1e-3, 3e-3)


def test_c5_width_generator_forward_backward(cuda_device):
    """The C5 widths (768 channels, 16 heads of 48, 3072 tokens, 77 x 1024 text tokens) at two blocks and two prompts."""
    cfg = {"inner_dim": 768, "condition_dim": 1024, "triplane_low_res": 32, "triplane_high_res": 64, "triplane_dim": 32,
           "num_layers": 2, "num_heads": 16, "flash_attention": False, "local_text": True}
    _compare(cfg, 2, cuda_device, 1e-3, 3e-3)


def test_native_generator_matches_reference_module_golden(cuda_device):
    """The native path against planes computed by the reference's own module (tests/golden/triplane_generator_golden.pt,
    see make_triplane_generator_golden.py): 64 channels, 4 heads of 16, 48 tokens, 7 x 48 text tokens, two blocks."""
    import os

    from scaledreamer_b200.amortized import TriplaneTransformer

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "triplane_generator_golden.pt"))["local_text"]
    gen = TriplaneTransformer(**g["cfg"]).to(cuda_device)
    gen.load_state_dict(g["state_dict"], strict=True)
    with torch.no_grad():
        out = gen(g["text_embed"].to(cuda_device))
    e = _rel(out.cpu(), g["planes"])
    print(f"planes vs reference module rel_l2 {e:.2e}")
    assert e < 1e-3


def test_generator_step_is_bitwise_reproducible(cuda_device):
    """No atomics anywhere in the generator: every reduction (LayerNorm / bias / position-embedding gradients, split
    weight-gradient products, softmax statistics) has a fixed order, so two runs on the same inputs agree bit for bit."""
    cfg = {"inner_dim": 128, "condition_dim": 1024, "triplane_low_res": 8, "triplane_high_res": 16, "triplane_dim": 32,
           "num_layers": 2, "num_heads": 4, "flash_attention": False, "local_text": True}
    gen = _build(cfg, cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    emb = torch.randn(3, 77, 1024, device=cuda_device, generator=g)
    d = torch.randn(3, 3, 32, 16, 16, device=cuda_device, generator=g)
    runs = []
    for _ in range(2):
        for p in gen.parameters():
            p.grad = None
        out = gen(emb)
        out.backward(d)
        runs.append((out.detach().clone(), [p.grad.clone() for p in gen.parameters()]))
    assert torch.equal(runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)


def test_no_torch_kernels_between_embeddings_and_planes(cuda_device):
    """Every kernel of the generator's forward + backward is this library's: the profiler sees no at:: / cuBLAS / SDPA
    kernel apart from the loss the test itself builds."""
    from torch.profiler import ProfilerActivity, profile

    cfg = {"inner_dim": 128, "condition_dim": 1024, "triplane_low_res": 8, "triplane_high_res": 16, "triplane_dim": 32,
           "num_layers": 1, "num_heads": 4, "flash_attention": False, "local_text": True}
    gen = _build(cfg, cuda_device)
    emb = torch.randn(2, 77, 1024, device=cuda_device)
    out = gen(emb)
    out.backward(torch.ones_like(out))  # warm-up: lazy initialisation outside the profile
    torch.cuda.synchronize()
    for p in gen.parameters():
        p.grad = None  # else autograd's own accumulation (at::add into .grad) shows up in the profile
    d = torch.randn(2, 3, 16, 16, 32, device=cuda_device).permute(0, 1, 4, 2, 3)  # channels-last gradient, as the sampler sends it
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = gen(emb)
        out.backward(d)
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    foreign = [n for n in names if not any(k in n for k in ("dense::", "Memset", "Memcpy"))]
    assert names and not foreign, foreign[:10]
