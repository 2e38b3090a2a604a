"""Loss aggregation of the two systems against the reference's own training_step methods (tests/golden/
make_system_golden.py): lambda schedules through C(), orientation / sparsity / opaque / eikonal terms, logging names."""
import os

import pytest
import torch

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "system_golden.pt"))["training_step"]


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_training_step_losses_match_reference(i):
    from scaledreamer_b200 import core
    from scaledreamer_b200.amortized import MultipromptRadienceFieldGeneratorSystem
    from scaledreamer_b200.systems import StableDreamer

    c = GOLD[i]
    out = {k: (v.clone().requires_grad_(True) if k == "normal" else v) for k, v in c["out"].items()}
    logged = {}

    class System:
        cfg = type("Cfg", (), {"loss": dict(c["loss_cfg"]), "visualize_samples": False})()
        prompt_utils = None

        def __call__(self, batch):
            return out

        def C(self, v):
            return core.C(v, 0, c["global_step"])

        def log(self, name, value, **kw):
            logged[name] = float(value.detach()) if torch.is_tensor(value) else float(value)

        def guidance(self, rgb, prompt_utils, **kw):
            assert rgb is out["comp_rgb"] and kw["rgb_as_latents"] is False
            return {"loss_asd": torch.tensor(1.25 + c["global_step"] * 1e-4), "grad_norm": torch.tensor(3.0),
                    "min_step": 20, "max_step": 980}

    cls = StableDreamer if c["system"] == "scaledreamer" else MultipromptRadienceFieldGeneratorSystem
    System.training_step = cls.training_step
    res = System().training_step({"elevation": torch.zeros(2)}, 0)
    torch.testing.assert_close(res["loss"].detach(), c["loss"], atol=0, rtol=1e-6)
    assert set(logged) == set(c["logged"]), (sorted(logged), sorted(c["logged"]))
    for k, v in c["logged"].items():
        assert abs(logged[k] - v) <= 1e-6 * max(1.0, abs(v)), k


def toy_model():
    """The module tree of tests/golden/make_system_golden.py (attribute paths of the C2 optimizer block)."""
    import torch.nn as nn

    class Enc(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.params = nn.Parameter(torch.zeros(n))

    class Net(nn.Module):
        def __init__(self, i, o):
            super().__init__()
            self.layers = nn.Sequential(nn.Linear(i, 8, bias=False), nn.ReLU(), nn.Linear(8, o, bias=False))

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


def test_optimizer_groups_match_reference_parse_optimizer():
    """parse_optimizer (threestudio/systems/utils.py:19-53): one group per dotted module path, in config order, with
    its own lr; parameters not named by any group are left out."""
    from scaledreamer_b200.systems import parse_optimizer

    o = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "system_golden.pt"))["optimizer"]
    model = toy_model()
    model.load_state_dict(o["init"])
    opt = parse_optimizer(o["config"], model)
    got = [{"name": g["name"], "lr": g["lr"], "n": sum(p.numel() for p in g["params"])} for g in opt.param_groups]
    assert got == o["groups"]
    named = {id(p) for g in opt.param_groups for p in g["params"]}
    for k in o["untouched"]:
        assert id(dict(model.named_parameters())[k]) not in named
    assert opt.defaults["betas"] == (0.0, 0.99) or tuple(opt.param_groups[0]["betas"]) == (0.0, 0.99)


def test_multiprompt_evaluation_steps_and_saved_views(tmp_path):
    """validation_step / test_step of the multi-prompt system (multiprompt_radience_field_generator.py:218-385): views
    filed under the sanitised prompt (test: under `name` when the fix-prompt data set supplies one), depth normalised
    per view; launch.save_views writes <name>/<index>.png rows of rgb | normal | opacity | depth."""
    import sys

    from PIL import Image

    from scaledreamer_b200.amortized import MultipromptRadienceFieldGeneratorSystem as Sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import launch

    B, H, W = 3, 5, 7
    g = torch.Generator().manual_seed(0)
    out = {"comp_rgb": torch.rand(B, H, W, 3, generator=g), "comp_normal": torch.rand(B, H, W, 3, generator=g),
           "opacity": torch.rand(B, H, W, 1, generator=g), "depth": torch.rand(B, H, W, 1, generator=g) * 4 + 1}

    class Stub:
        cfg = type("Cfg", (), {"visualize_samples": False})()
        _eval_images = Sys._eval_images
        validation_step = Sys.validation_step
        test_step = Sys.test_step

        def __call__(self, batch):
            return out

    batch = {"prompt": ["a red apple, on a table."], "index": torch.arange(B)}
    v = Stub().validation_step(batch, 0)
    assert v["name"] == "a_red_apple_on_a_table" and v["index"].tolist() == [0, 1, 2]
    for i in range(B):
        d = out["depth"][i, :, :, 0]
        torch.testing.assert_close(v["depth"][i], (d - d.min()) / (d.max() - d.min()))
    t = Stub().test_step({**batch, "name": ["a corgi_to_a cat"]}, 0)
    assert t["name"] == "a_corgi_to_a_cat"
    assert Stub().test_step(batch, 0)["name"] == v["name"]  # library test set: no `name`, the prompt is used
    launch.save_views([v, t], str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["a_corgi_to_a_cat", "a_red_apple_on_a_table"]
    assert sorted(os.listdir(tmp_path / "a_red_apple_on_a_table")) == ["0.png", "1.png", "2.png"]
    assert Image.open(tmp_path / "a_red_apple_on_a_table" / "2.png").size == (4 * W, H)


@pytest.mark.parametrize("tag", ["prox", "no_prox"])
def test_adan_clip_coefficient_matches_reference_trajectory(tag):
    """FusedAdan(max_grad_norm > 0): the clip coefficient (threestudio/systems/optimizers.py:108-128) folded into the
    kernel's gradient scale. The reference's own Adan produced the trajectory (tests/golden/make_adan_clip_golden.py,
    two groups, some steps clipped); here the coefficient comes from FusedAdan.clip_coefficient() and the per-element
    update is adan_kernel's arithmetic written in torch (the kernel itself: tests/test_zz_multiprompt_eval_gpu.py)."""
    import math

    from scaledreamer_b200.systems import FusedAdan

    c = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adan_clip_golden.pt"))[tag]
    ps = [torch.nn.Parameter(p.clone()) for p in c["p0"]]
    opt = FusedAdan([{"params": [ps[0]], "lr": c["lrs"][0]}, {"params": [ps[1]], "lr": c["lrs"][1]}], lr=1e-2,
                    betas=c["betas"], eps=c["eps"], weight_decay=c["weight_decay"], no_prox=c["no_prox"],
                    max_grad_norm=c["max_grad_norm"])
    b1, b2, b3 = c["betas"]
    state = [dict(m=torch.zeros_like(p), n=torch.zeros_like(p), d=torch.zeros_like(p), prev=torch.zeros_like(p)) for p in ps]
    clipped = 0
    for step in range(1, 7):
        for p, g in zip(ps, c["grads"][step - 1]):
            p.grad = g.clone()
        coef = opt.clip_coefficient()
        assert coef == pytest.approx(min(1.0, c["max_grad_norm"] / (c["norms"][step - 1] + c["eps"])), rel=1e-6)
        clipped += coef < 1.0
        for p, s, lr in zip(ps, state, c["lrs"]):
            with torch.no_grad():
                gi = p.grad * coef
                diff = torch.zeros_like(gi) if step == 1 else gi - s["prev"]
                s["m"] = b1 * s["m"] + (1 - b1) * gi
                s["d"] = b2 * s["d"] + (1 - b2) * diff
                u = b2 * diff + gi
                s["n"] = b3 * s["n"] + (1 - b3) * u * u
                s["prev"] = gi
                denom = s["n"].sqrt() / math.sqrt(1 - b3 ** step) + c["eps"]
                if c["no_prox"]:
                    p.mul_(1 - lr * c["weight_decay"])
                p.sub_((lr / (1 - b1 ** step)) * (s["m"] / denom) + (lr * b2 / (1 - b2 ** step)) * (s["d"] / denom))
                if not c["no_prox"]:
                    p.div_(1 + lr * c["weight_decay"])
        for p, ref in zip(ps, c["params"][step - 1]):
            torch.testing.assert_close(p.detach(), ref, atol=1e-6, rtol=1e-5)
    assert clipped == 4  # norms 13.9, 43.1, 29.6, 6.9 against max_grad_norm 5
    # data-parallel runs hand the optimizer SUMMED gradients and grad_scale = 1 / world: the norm is that of the mean
    opt.grad_scale = 0.5
    for p, g in zip(ps, c["grads"][1]):
        p.grad = g.clone() * 2.0
    assert opt.clip_coefficient() == pytest.approx(min(1.0, c["max_grad_norm"] / (c["norms"][1] + c["eps"])), rel=1e-6)
    assert FusedAdan([torch.nn.Parameter(torch.zeros(3))]).clip_coefficient() == 1.0
