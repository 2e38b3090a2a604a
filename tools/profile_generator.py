"""One forward + backward of the C5 Triplane-Transformer generator at the C5 widths (768 channels, 16 heads, 3072 tokens,
4 prompts) with 2 of the 12 blocks, bracketed by cudaProfilerStart/Stop (use with `ncu --profile-from-start off ...`).
Never a benchmark: numbers under a profiler are not bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scaledreamer_b200.amortized import TriplaneTransformer

cfg = {"inner_dim": 768, "condition_dim": 1024, "triplane_low_res": 32, "triplane_high_res": 64, "triplane_dim": 32,
       "num_layers": 2, "num_heads": 16, "flash_attention": False, "local_text": True}
torch.manual_seed(0)
gen = TriplaneTransformer(**cfg).cuda()
emb = torch.randn(4, 77, 1024, device="cuda")
d = torch.randn(4, 3, 64, 64, 32, device="cuda").permute(0, 1, 4, 2, 3)
for _ in range(2):
    for p in gen.parameters():
        p.grad = None
    gen(emb).backward(d)
for p in gen.parameters():
    p.grad = None
torch.cuda.synchronize()
torch.cuda.profiler.start()
gen(emb).backward(d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
