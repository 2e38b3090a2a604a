"""Adan with global-norm clipping (max_grad_norm > 0) from the REFERENCE's own optimizer class.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_adan_clip_golden.py
threestudio/systems/optimizers.py is plain torch and is executed as it is (runpy). Two parameter tensors in two groups
with different learning rates, six steps; the gradients are scaled so that some steps clip and some do not.
Output: tests/golden/adan_clip_golden.pt (a few kB).
"""
import os
import runpy

import torch
import torch.nn as nn

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adan_clip_golden.pt")


def main():
    adan = runpy.run_path("/root/reference/threestudio/systems/optimizers.py")["Adan"]
    g = torch.Generator().manual_seed(7)
    cases = {}
    for tag, kw in (("prox", dict(weight_decay=0.02, no_prox=False)), ("no_prox", dict(weight_decay=0.02, no_prox=True))):
        ps = [nn.Parameter(torch.randn(131, generator=g)), nn.Parameter(torch.randn(7, 9, generator=g))]
        p0 = [p.detach().clone() for p in ps]
        opt = adan([{"params": [ps[0]], "lr": 1e-2}, {"params": [ps[1]], "lr": 3e-3}], lr=1e-2, betas=(0.98, 0.92, 0.99),
                   eps=1e-8, max_grad_norm=5.0, foreach=False, **kw)
        grads, traj, norms = [], [], []
        for step in range(6):
            scale = (0.1, 1.0, 3.0, 0.2, 2.0, 0.5)[step]  # ||g|| around 1.4 .. 42: steps 1, 2, 4 clip at 5.0
            gs = [torch.randn(p.shape, generator=g) * scale for p in ps]
            norms.append(float(torch.sqrt(sum(x.pow(2).sum() for x in gs))))
            for p, x in zip(ps, gs):
                p.grad = x.clone()
            opt.step()
            grads.append(gs)
            traj.append([p.detach().clone() for p in ps])
        cases[tag] = {"p0": p0, "grads": grads, "params": traj, "norms": norms, "lrs": (1e-2, 3e-3),
                      "betas": (0.98, 0.92, 0.99), "eps": 1e-8, "max_grad_norm": 5.0, **kw}
    torch.save(cases, OUT)
    print("wrote", OUT, {k: [round(n, 2) for n in v["norms"]] for k, v in cases.items()})


if __name__ == "__main__":
    main()
