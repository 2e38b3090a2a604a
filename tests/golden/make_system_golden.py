"""Golden loss aggregation produced by the REFERENCE's own training_step methods.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_system_golden.py
StableDreamer.training_step (threestudio/systems/scaledreamer.py:48-170) and
MultipromptRadienceFieldGeneratorSystem.training_step (custom/amortized/systems/multiprompt_radience_field_generator.py:
127-216) are taken out of the files with `ast` and bound to a stand-in object: the renderer output `out` and the guidance
output are recorded tensors, `self.C` is the reference's own C() at a given (epoch, global_step), `self.log` records.
Output: tests/golden/system_golden.pt (about 30 kB).
"""
import ast
import math
import os

import torch
import torch.nn.functional as F

ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "system_golden.pt")


def top(path, names, ns):
    s = open(path).read()
    for node in ast.parse(s).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.get_source_segment(s, node), path, "exec"), ns)


def method(path, cls_name, name, ns):
    s = open(path).read()
    cls = next(n for n in ast.parse(s).body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
    body = "\n".join(s.splitlines()[fn.lineno - 1:fn.end_lineno])
    exec(compile("\n".join(line[4:] for line in body.splitlines()), path, "exec"), ns)
    return ns[name]


class AttrDict(dict):  # OmegaConf-like: cfg.loss.lambda_x and cfg.loss["lambda_x"]
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def scene(seed, with_sdf):
    g = torch.Generator().manual_seed(seed)
    B, H, W, N = 2, 8, 8, 500
    opacity = torch.rand(B, H, W, 1, generator=g)
    opacity[0, :2] = 0.0
    opacity[1, -1] = 1.0
    out = {"comp_rgb": torch.rand(B, H, W, 3, generator=g), "opacity": opacity, "weights": torch.rand(N, 1, generator=g),
           "normal": F.normalize(torch.randn(N, 3, generator=g), dim=-1),
           "t_dirs": F.normalize(torch.randn(N, 3, generator=g), dim=-1),
           "z_variance": torch.rand(B, H, W, 1, generator=g)}
    if with_sdf:
        out["sdf_grad"] = torch.randn(N, 3, generator=g) * 0.3 + F.normalize(torch.randn(N, 3, generator=g), dim=-1)
        out["inv_std"] = torch.tensor(29.96)
    return out


def toy_model(seed=0):
    """Module tree with the attribute paths the C2 optimizer block addresses (asd_sd_nerf.yaml:110-125)."""
    import torch.nn as nn

    g = torch.Generator().manual_seed(seed)

    class Enc(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.params = nn.Parameter(torch.randn(n, generator=g) * 0.1)

    class Net(nn.Module):
        def __init__(self, i, o):
            super().__init__()
            self.layers = nn.Sequential(nn.Linear(i, 8, bias=False), nn.ReLU(), nn.Linear(8, o, bias=False))
            for p in self.parameters():
                p.data = torch.randn(p.shape, generator=g) * 0.3

    class Part(nn.Module):
        def __init__(self, with_feature):
            super().__init__()
            self.encoding = Enc(37)
            self.network = Net(4, 3)
            if with_feature:
                self.density_network, self.feature_network = Net(4, 1), Net(4, 3)

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.geometry, self.background = Part(True), Part(False)

    return Model()


def optimizer_case():
    """parse_optimizer (threestudio/systems/utils.py:19-53) -> torch.optim.AdamW with the C2 groups, five steps."""
    import types as _t

    import torch.nn as nn

    ns = {"torch": torch, "nn": nn, "threestudio": _t.SimpleNamespace(debug=lambda *a, **k: None)}
    top(f"{ROOT}/threestudio/systems/utils.py", ["getattr_recursive", "get_parameters", "parse_optimizer"], ns)
    cfg_dict = {"name": "AdamW", "args": {"betas": [0.0, 0.99], "eps": 1e-15},
                "params": {"geometry.encoding": {"lr": 0.01}, "geometry.density_network": {"lr": 0.001},
                           "geometry.feature_network": {"lr": 0.001}, "background.encoding": {"lr": 0.01},
                           "background.network": {"lr": 0.001}}}
    cfg = AttrDict(name=cfg_dict["name"], args=cfg_dict["args"], params=cfg_dict["params"])
    model = toy_model()
    init = {k: v.clone() for k, v in model.state_dict().items()}
    opt = ns["parse_optimizer"](cfg, model)
    groups = [{"name": gr["name"], "lr": gr["lr"], "n": sum(p.numel() for p in gr["params"])} for gr in opt.param_groups]
    g = torch.Generator().manual_seed(4)
    grads = []
    for _ in range(5):
        step = {}
        for k, p in model.named_parameters():
            p.grad = torch.randn(p.shape, generator=g)
            step[k] = p.grad.clone()
        grads.append(step)
        opt.step()
    return {"config": cfg_dict, "init": init, "groups": groups, "grads": grads,
            "final": {k: v.clone() for k, v in model.state_dict().items()},
            "untouched": ["geometry.network.layers.0.weight", "geometry.network.layers.2.weight"]}


def main():
    ns = {"torch": torch, "F": F, "math": math, "config_to_primitive": lambda v: list(v) if isinstance(v, (list, tuple)) else v,
          "Any": object}
    top(f"{ROOT}/threestudio/utils/misc.py", ["C"], ns)
    top(f"{ROOT}/threestudio/utils/ops.py", ["dot", "binary_cross_entropy"], ns)
    systems = {
        "scaledreamer": method(f"{ROOT}/threestudio/systems/scaledreamer.py", "StableDreamer", "training_step", dict(ns)),
        "multiprompt": method(f"{ROOT}/custom/amortized/systems/multiprompt_radience_field_generator.py",
                              "MultipromptRadienceFieldGeneratorSystem", "training_step", dict(ns)),
    }
    losses = {
        "c2": dict(lambda_asd=1.0, lambda_orient=0.0, lambda_sparsity=30, lambda_opaque=[10000, 0.0, 100.0, 10001],
                   lambda_z_variance=0.0),
        "c1": dict(lambda_asd=1.0, lambda_orient=[0, 10.0, 1000.0, 5000], lambda_sparsity=30,
                   lambda_opaque=[10000, 0.0, 100.0, 10001], lambda_z_variance=0.0),
        "c4": dict(lambda_asd=1.0, lambda_orient=0.0, lambda_sparsity=20.0, lambda_opaque=[40000, 0.0, 100.0, 40001],
                   lambda_z_variance=0.0, lambda_eikonal=[0, 100.0, 1.0, 5000]),
        "c5": dict(lambda_asd=1.0, lambda_orient=0.0, lambda_sparsity=0.0, lambda_opaque=0.0, lambda_z_variance=0.0,
                   lambda_eikonal=0.01),
    }
    gold = []
    for sysname, lossname, steps, with_sdf in (("scaledreamer", "c2", (0, 10000, 10001, 20000), False),
                                               ("scaledreamer", "c1", (0, 2500, 9000), False),
                                               ("multiprompt", "c4", (0, 2500, 40001), True),
                                               ("multiprompt", "c5", (0, 77), True)):
        for step in steps:
            out = scene(step + len(lossname), with_sdf)
            logged = {}

            class System:
                cfg = AttrDict(stage="coarse", rgb_as_latents=False, loss=AttrDict(losses[lossname]))
                prompt_utils = None

                def __call__(self, batch):
                    return out

                def C(self, v):
                    return ns["C"](v, 0, step)

                def log(self, name, value, **kw):
                    logged[name] = float(value)

                def guidance(self, rgb, prompt_utils, **kw):
                    assert rgb is out["comp_rgb"] and kw["rgb_as_latents"] is False
                    return {"loss_asd": torch.tensor(1.25 + step * 1e-4), "grad_norm": torch.tensor(3.0), "min_step": 20,
                            "max_step": 980}

            System.training_step = systems[sysname]
            res = System().training_step({"elevation": torch.zeros(2)}, 0)
            gold.append({"system": sysname, "loss_cfg": losses[lossname], "global_step": step, "out": out,
                         "loss": res["loss"].detach().clone(), "logged": logged})
    gold = {"training_step": gold, "optimizer": optimizer_case()}
    torch.save(gold, OUT)
    print("wrote", OUT, len(gold["training_step"]), "cases;", list(gold["optimizer"]))


if __name__ == "__main__":
    main()
