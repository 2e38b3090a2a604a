"""Times GroupNorm(+SiLU) forward / backward at the VAE and UNet shapes of C2 and prints achieved HBM GB/s.
Algorithmic bytes: forward = read x twice + write y (3 x tensor); backward = read x, dy twice + write dx (5 x tensor).
Usage (GPU box): python tools/time_gn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scaledreamer_b200 import nn_ops as N  # noqa: E402

SHAPES = [("vae L0", 1, 512 * 512, 128), ("vae L1", 1, 256 * 256, 256), ("vae L2", 1, 128 * 128, 512),
          ("vae L3", 1, 64 * 64, 512), ("unet 320@64", 5, 64 * 64, 320), ("unet 640@64", 5, 64 * 64, 640),
          ("unet 960@64", 5, 64 * 64, 960), ("unet 640@32", 5, 32 * 32, 640), ("unet 1280@16", 5, 16 * 16, 1280),
          ("unet 2560@8", 5, 8 * 8, 2560)]


def timeit(fn, iters=int(os.environ.get('GN_ITERS', '20'))):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def main():
    dev = "cuda"
    for name, n, hw, c in SHAPES:
        x = torch.randn(n, hw, c, device=dev, dtype=torch.float16)
        dy = torch.randn_like(x)
        g, b = torch.randn(c, device=dev, dtype=torch.float16), torch.randn(c, device=dev, dtype=torch.float16)
        y, stats = N.groupnorm(x, g, b, 32, 1e-6, True)
        mb = x.numel() * 2 / 1e6
        tf = timeit(lambda: N.groupnorm(x, g, b, 32, 1e-6, True))
        tb = timeit(lambda: N.groupnorm_backward(x, g, b, stats, dy, 32, 1e-6, True))
        print(f"{name:14s} {mb:7.1f} MB  fwd {tf * 1e3:7.1f} us {3 * mb / tf / 1e3:6.2f} TB/s   "
              f"bwd {tb * 1e3:7.1f} us {5 * mb / tb / 1e3:6.2f} TB/s")


if __name__ == "__main__":
    main()
