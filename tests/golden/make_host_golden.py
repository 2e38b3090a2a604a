"""Golden vectors for the small host-side / glue pieces of the path, produced by the REFERENCE's own definitions.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_host_golden.py
The reference modules cannot be imported (pytorch-lightning, tinycudann, igl ... at module level), so the definitions
are taken out of the files by name with `ast`, compiled unchanged and run against a namespace that holds only torch /
math and inert stand-ins for the type annotations:
    threestudio/utils/misc.py                       C  (scheduled scalars)
    threestudio/utils/config.py                     C_max and the OmegaConf resolver lambdas
    threestudio/utils/ops.py                        get_ray_directions, get_rays, get_projection_matrix, get_mvp_matrix,
                                                    shifted_expotional_decay
    threestudio/models/prompt_processors/base.py    DirectionConfig, PromptProcessorOutput, shift_azimuth_deg, the
                                                    `self.directions = [...]` list and the Perp-Neg defaults of
                                                    PromptProcessor.Config
    custom/amortized/models/prompt_processors/base.py   MultiPromptProcessorOutput (per-sample prompt tables)
    threestudio/models/guidance/stable_diffusion_asd_guidance.py   the methods __call__, get_latents, get_t_plus, get_eps
                                                    of the SD ASD guidance, bound to a stand-in object whose UNet is a
                                                    recorded random tensor and whose scheduler.add_noise is q_sample on the
                                                    LDM linear beta schedule (extern/mvdream/ldm/.../util.py:37-40)
    threestudio/models/guidance/mvdream_asd_guidance.py   __call__, get_latents, get_t_plus, get_camera_cond of the
                                                    multi-view guidance with extern/mvdream/camera_utils.py normalize_camera,
                                                    same stand-ins (model.q_sample = q_sample, model.apply_model = recorded)
Output: tests/golden/host_golden.pt (about 0.5 MB).
"""
import ast
import math
import os
import types
from dataclasses import dataclass
from typing import Callable, Dict, List, Tuple

import torch
import torch.nn.functional as F

REF = "/root/reference/threestudio"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.pt")


class _Any:
    def __getitem__(self, item):
        return self


def _src(path):
    s = open(path).read()
    return s, ast.parse(s)


def top_level(path, names, ns):
    s, tree = _src(path)
    found = set()
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.get_source_segment(s, node) if not node.decorator_list else
                         "\n".join(s.splitlines()[node.decorator_list[0].lineno - 1:node.end_lineno]), path, "exec"), ns)
            found.add(node.name)
    assert found == set(names), set(names) - found


def base_namespace():
    return {"torch": torch, "F": F, "math": math, "Float": _Any(), "Tensor": torch.Tensor, "Tuple": Tuple, "List": List,
            "Dict": Dict, "Callable": Callable, "Union": _Any(), "Optional": _Any(), "Any": object, "dataclass": dataclass,
            "config_to_primitive": lambda v: list(v) if isinstance(v, (list, tuple)) else v}


def prompt_pieces(ns):
    """DirectionConfig / PromptProcessorOutput / shift_azimuth_deg, the second `self.directions = [...]` list
    ("{s}, front view" wording, view_dependent_prompt_front = False) and the Perp-Neg coefficient defaults."""
    path = f"{REF}/models/prompt_processors/base.py"
    top_level(path, ["DirectionConfig", "PromptProcessorOutput", "shift_azimuth_deg"], ns)
    s, tree = _src(path)
    pp = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PromptProcessor")
    cfg_cls = next(n for n in pp.body if isinstance(n, ast.ClassDef) and n.name == "Config")
    defaults = {}
    for n in cfg_cls.body:
        if isinstance(n, ast.AnnAssign) and n.value is not None and isinstance(n.target, ast.Name):
            try:
                defaults[n.target.id] = ast.literal_eval(n.value)
            except Exception:
                pass
    lists = [n for n in ast.walk(pp) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute)
             and n.targets[0].attr == "directions"]
    assert len(lists) == 2
    expr = ast.get_source_segment(s, lists[1].value)
    return defaults, expr


def methods(path, cls_name, names):
    s, tree = _src(path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    out = {}
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            body = "\n".join(s.splitlines()[n.lineno - 1:n.end_lineno])  # without decorators (autocast wrappers)
            out[n.name] = "\n".join(line[4:] for line in body.splitlines())
    assert set(out) == set(names), set(names) - set(out)
    return out


class _RecordingTorch:
    """`torch` as the extracted methods see it: every random draw is recorded so the tests can replay it."""

    def __init__(self):
        self.draws = {}

    def __getattr__(self, name):
        return getattr(torch, name)

    def randn_like(self, x):
        self.draws["noise"] = torch.randn_like(x)
        return self.draws["noise"]

    def randint(self, *a, **k):
        self.draws["t"] = torch.randint(*a, **k)
        return self.draws["t"]

    def rand(self, *a, **k):
        self.draws["u"] = torch.rand(*a, **k)
        return self.draws["u"]


def guidance_case(ns, prompt_utils, elevation, azimuth, dist):
    """SDTimestepShiftedScoreDistillationGuidance.__call__ (:211-292) with rgb_as_latents=True on 64x64x4 inputs."""
    gns = dict(ns)
    rec = _RecordingTorch()
    gns["torch"] = rec
    gns["nn"] = torch.nn
    gns["PromptProcessorOutput"] = ns["PromptProcessorOutput"]
    top_level(f"{REF}/utils/ops.py", ["perpendicular_component"], gns)
    src = methods(f"{REF}/models/guidance/stable_diffusion_asd_guidance.py",
                  "SDTimestepShiftedScoreDistillationGuidance", ["__call__", "get_latents", "get_t_plus", "get_eps"])
    for code in src.values():
        exec(compile(code, "stable_diffusion_asd_guidance.py", "exec"), gns)
    import numpy as np

    lns = {"torch": torch, "np": np}
    top_level("/root/reference/extern/mvdream/ldm/modules/diffusionmodules/util.py", ["make_beta_schedule"], lns)
    betas = torch.as_tensor(lns["make_beta_schedule"]("linear", 1000, linear_start=0.00085, linear_end=0.0120))
    alphas = torch.cumprod(1.0 - betas, dim=0).float()
    B = elevation.shape[0]
    gen = torch.Generator().manual_seed(21)
    unet_out = torch.randn(5 * B, 4, 64, 64, generator=gen)
    seen = {}

    class Guidance:
        pass

    for name in src:
        setattr(Guidance, name, gns[name])

    def forward_unet(self, unet, latents, t, encoder_hidden_states):
        seen.update(unet_x=latents.clone(), unet_t=t.clone(), ctx=encoder_hidden_states.clone())
        return unet_out

    Guidance.forward_unet = forward_unet
    g = Guidance()
    g.cfg = types.SimpleNamespace(guidance_scale=7.5, guidance_perp_neg=-0.5, plus_ratio=0.1, plus_random=True,
                                  weighting_strategy="sds", view_dependent_prompting=True)
    g.use_perp_neg, g.device, g.unet = True, torch.device("cpu"), None
    g.num_train_timesteps, g.min_step, g.max_step, g.alphas, g.grad_clip_val = 1000, 20, 980, alphas, None
    g.scheduler = types.SimpleNamespace(add_noise=lambda x, n, t: alphas[t].sqrt().view(-1, 1, 1, 1) * x
                                        + (1 - alphas[t]).sqrt().view(-1, 1, 1, 1) * n)
    latents = (torch.randn(B, 64, 64, 4, generator=gen) * 0.8).requires_grad_(True)  # "rgb" in BHWC, used as latents
    torch.manual_seed(77)
    out = g(latents, prompt_utils, elevation, azimuth, dist, rgb_as_latents=True)
    out["loss_asd"].backward()
    # the UNet stand-in output is reproducible from its seed (torch.Generator().manual_seed(21), drawn before the
    # latents); a strided sample of it and of the UNet input batch is kept to check the replay
    return {"latents_bhwc": latents.detach(), "noise": rec.draws["noise"], "t": rec.draws["t"], "u": rec.draws["u"],
            "t_plus": seen["unet_t"][-B:].long(), "unet_t": seen["unet_t"], "ctx": seen["ctx"],
            "unet_x_sample": seen["unet_x"].flatten()[::97].clone(), "unet_out_seed": 21,
            "unet_out_sample": unet_out.flatten()[::97].clone(), "alphas_cumprod": alphas,
            "loss_asd": out["loss_asd"].detach(), "grad_norm": out["grad_norm"].detach(),
            "grad_bhwc": latents.grad.clone(), "min_step": 20, "max_step": 980, "guidance_scale": 7.5,
            "guidance_perp_neg": -0.5, "plus_ratio": 0.1, "elevation": elevation, "azimuth": azimuth}


def mv_guidance_case(ns, prompt_utils, c2w3):
    """MVDreamTimestepShiftedScoreDistillationGuidance.__call__ (mvdream_asd_guidance.py:167-304): 4 views of one
    object, one shared timestep, plain CFG, camera conditioning."""
    gns = dict(ns)
    rec = _RecordingTorch()
    gns["torch"] = rec
    gns["PromptProcessorOutput"] = ns["PromptProcessorOutput"]
    import numpy as np

    cns = {"torch": torch, "np": np}
    top_level("/root/reference/extern/mvdream/camera_utils.py", ["normalize_camera"], cns)
    gns["normalize_camera"] = cns["normalize_camera"]
    src = methods(f"{REF}/models/guidance/mvdream_asd_guidance.py", "MVDreamTimestepShiftedScoreDistillationGuidance",
                  ["__call__", "get_latents", "get_t_plus", "get_camera_cond"])
    for code in src.values():
        exec(compile(code, "mvdream_asd_guidance.py", "exec"), gns)
    lns = {"torch": torch, "np": np}
    top_level("/root/reference/extern/mvdream/ldm/modules/diffusionmodules/util.py", ["make_beta_schedule"], lns)
    betas = torch.as_tensor(lns["make_beta_schedule"]("linear", 1000, linear_start=0.00085, linear_end=0.0120))
    alphas = torch.cumprod(1.0 - betas, dim=0).float()
    B = 4
    gen = torch.Generator().manual_seed(33)
    unet_out = torch.randn(3 * B, 4, 32, 32, generator=gen)
    seen = {}

    class Guidance:
        pass

    for name in src:
        setattr(Guidance, name, gns[name])
    g = Guidance()
    g.cfg = types.SimpleNamespace(guidance_scale=7.5, plus_ratio=0.1, plus_random=True, weighting_strategy="sds",
                                  view_dependent_prompting=False, camera_condition_type="rotation", n_view=4)
    g.device, g.num_train_timesteps, g.min_step, g.max_step, g.alphas, g.grad_clip_val = \
        torch.device("cpu"), 1000, 20, 980, alphas, None

    def apply_model(x, t, context):
        seen.update(unet_x=x.clone(), unet_t=t.clone(), ctx=context["context"].clone(), camera=context["camera"].clone(),
                    num_frames=context["num_frames"])
        return unet_out

    g.model = types.SimpleNamespace(
        q_sample=lambda x, t, noise: alphas[t].sqrt().view(-1, 1, 1, 1) * x + (1 - alphas[t]).sqrt().view(-1, 1, 1, 1) * noise,
        apply_model=apply_model)
    c2w = torch.cat([c2w3, c2w3[:1].flip(-1)], 0).clone()
    c2w[:, 3, :] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    c2w[:, :3, 3] *= torch.tensor([0.7, 1.3, 2.0, 1.0])[:, None]
    latents = (torch.randn(B, 32, 32, 4, generator=gen) * 0.8).requires_grad_(True)
    el, az, dist = torch.tensor([15.0] * 4), torch.tensor([-100.0, -10.0, 80.0, 170.0]), torch.full((4,), 1.1)
    torch.manual_seed(78)
    out = g(latents, prompt_utils, el, az, dist, c2w.clone(), rgb_as_latents=True)
    out["loss_asd"].backward()
    return {"latents_bhwc": latents.detach(), "noise": rec.draws["noise"], "t": rec.draws["t"], "u": rec.draws["u"],
            "unet_t": seen["unet_t"], "ctx": seen["ctx"], "camera": seen["camera"], "num_frames": seen["num_frames"],
            "c2w": c2w, "unet_x_sample": seen["unet_x"].flatten()[::53].clone(), "unet_out_seed": 33,
            "unet_out_sample": unet_out.flatten()[::53].clone(), "alphas_cumprod": alphas,
            "loss_asd": out["loss_asd"].detach(), "grad_norm": out["grad_norm"].detach(),
            "grad_bhwc": latents.grad.clone(), "min_step": 20, "max_step": 980, "guidance_scale": 7.5, "plus_ratio": 0.1}


def main():
    gold = {}
    g = torch.Generator().manual_seed(3)

    # ---- C(): threestudio/utils/misc.py:66-101
    ns = base_namespace()
    top_level(f"{REF}/utils/misc.py", ["C"], ns)
    cases = []
    specs = [0.5, 3, [0, 0.5, 0.02, 25000], [0.98, 0.5, 25000], [10000, 0.0, 100.0, 10001], [0, 100.0, 1.0, 5000],
             [0, 1.0, 0.1, 2.0], [0, 0.1, 1.0, 100, 2.0, 300, 0.5, 1000], [50, 1e-3, 1e-1, 150]]
    for spec in specs:
        for epoch, step in ((0, 0), (0, 1), (1, 77), (1, 150), (3, 5000), (7, 12500), (9, 10000), (9, 10001), (20, 40000)):
            for interp in ("linear", "exp"):
                if interp == "exp" and (not isinstance(spec, list) or min(v for v in (spec[-3:-1] if len(spec) != 8 else [1]) ) <= 0):
                    continue
                if interp == "exp" and isinstance(spec, list) and len(spec) == 8:
                    continue
                cases.append({"value": spec, "epoch": epoch, "global_step": step, "interpolation": interp,
                              "out": float(ns["C"](spec, epoch, step, interp))})
    gold["C"] = cases

    # ---- OmegaConf resolvers and C_max: threestudio/utils/config.py:10-49 (the lambdas handed to
    # OmegaConf.register_new_resolver are taken out of the call expressions)
    rns = dict(ns, os=os)
    top_level(f"{REF}/utils/config.py", ["C_max"], rns)
    src_cfg, tree_cfg = _src(f"{REF}/utils/config.py")
    resolvers = {}
    for node in ast.walk(tree_cfg):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "register_new_resolver":
            resolvers[node.args[0].value] = eval(compile(ast.Expression(node.args[1]), "config.py", "eval"), rns)
    sched = [0, 0.0, 0.5, 100, 2.0, 300, -1.0, 1000]
    calls = [("calc_exp_lr_decay_rate", (0.1, 25000)), ("add", (3, 4.5)), ("sub", (3, 4.5)), ("mul", (3, 4.5)),
             ("div", (7, 2)), ("idiv", (7, 2)), ("basename", ("a/b/c.yaml",)), ("rmspace", ("a DSLR photo", "_")),
             ("tuple2", ("1.5",)), ("gt0", (0.0,)), ("gt0", (0.1,)), ("cmaxgt0", ([10000, 0.0, 100.0, 10001],)),
             ("cmaxgt0", (0.0,)), ("cmaxgt0", ([0.0, 0.0, 5],)), ("cmaxgt0", (sched,)), ("not", (True,)), ("not", (0,)),
             ("cmaxgt0orcmaxgt0", (0.0, [0, 0.0, 1.0, 10])), ("cmaxgt0orcmaxgt0", (0.0, 0.0))]
    gold["resolvers"] = {"names": sorted(resolvers), "calls": [(n, a, resolvers[n](*a)) for n, a in calls],
                         "c_max": [(v, rns["C_max"](v)) for v in (0.5, 3, [0, 0.5, 0.02, 25000], [0.98, 0.5, 25000], sched,
                                                                  [0, 1.0, 0.1, 10, 4.0, 20])]}

    # ---- rays / projection: threestudio/utils/ops.py:183-300, used as data/uncond.py:302-328 uses them
    ns = base_namespace()
    top_level(f"{REF}/utils/ops.py", ["get_ray_directions", "get_rays", "get_projection_matrix", "get_mvp_matrix",
                                      "shifted_expotional_decay"], ns)
    B, H, W = 3, 12, 20
    q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    c2w = torch.zeros(B, 4, 4)
    c2w[:, :3, :3] = q
    c2w[:, :3, 3] = torch.randn(B, 3, generator=g)
    c2w[:, 3, 3] = 1.0
    fovy = torch.deg2rad(torch.tensor([40.0, 55.0, 70.0]))
    focal = 0.5 * H / torch.tan(0.5 * fovy)
    directions = ns["get_ray_directions"](H=H, W=W, focal=1.0)[None].repeat(B, 1, 1, 1)
    directions[:, :, :, :2] = directions[:, :, :, :2] / focal[:, None, None, None]
    rays_o, rays_d = ns["get_rays"](directions, c2w, keepdim=True, normalize=True)
    proj = ns["get_projection_matrix"](fovy, W / H, 0.1, 1000.0)
    gold["rays"] = {"c2w": c2w, "fovy": fovy, "H": H, "W": W, "rays_o": rays_o.contiguous(), "rays_d": rays_d,
                    "proj_mtx": proj, "mvp_mtx": ns["get_mvp_matrix"](c2w, proj)}

    # ---- view-dependent / Perp-Neg text embeddings: prompt_processors/base.py:37-167
    defaults, dir_expr = prompt_pieces(ns)
    cfg = types.SimpleNamespace(front_threshold=30.0, back_threshold=30.0, overhead_threshold=60.0)
    directions_list = eval(compile(dir_expr, "base.py", "eval"), dict(ns, self=types.SimpleNamespace(cfg=cfg)))
    T, D = 5, 16
    half = lambda t: t.half().float()
    tables = {k: half(torch.randn(*shape, generator=g)) for k, shape in
              (("text_embeddings", (1, T, D)), ("uncond_text_embeddings", (1, T, D)),
               ("text_embeddings_vd", (4, T, D)), ("uncond_text_embeddings_vd", (4, T, D)))}
    coeff = {k: tuple(defaults[k]) for k in ("perp_neg_f_sb", "perp_neg_f_fsb", "perp_neg_f_fs", "perp_neg_f_sf")}
    out = ns["PromptProcessorOutput"](directions=directions_list,
                                      direction2idx={d.name: i for i, d in enumerate(directions_list)},
                                      use_perp_neg=True, prompt="p", prompts_vd=["a", "b", "c", "d"], **tables, **coeff)
    elevation = torch.tensor([10.0, 0.0, 35.0, -5.0, 75.0, 61.0, 20.0, 59.0, 15.0, 0.0, 5.0, 45.0, 30.0, 12.0])
    azimuth = torch.tensor([0.0, 17.5, -29.0, 45.0, 10.0, -170.0, 89.0, 91.0, 135.0, 179.0, -179.0, -120.0, 200.0, -95.0])
    dist = torch.full_like(elevation, 1.2)
    pn, w = out.get_text_embeddings_perp_neg(elevation, azimuth, dist, True)
    gold["prompt"] = {**tables, **coeff, "front_threshold": 30.0, "back_threshold": 30.0, "overhead_threshold": 60.0,
                      "elevation": elevation, "azimuth": azimuth,
                      "vd": out.get_text_embeddings(elevation, azimuth, dist, True),
                      "global": out.get_text_embeddings(elevation, azimuth, dist, False).contiguous(),
                      "perp_neg": pn, "neg_weights": w.float(),
                      "decay_check": float(ns["shifted_expotional_decay"](1.0, 0.5, -0.606, torch.tensor(0.25)))}
    # ---- multi-prompt batches: custom/amortized/models/prompt_processors/base.py:409-568
    mns = dict(ns)
    top_level("/root/reference/custom/amortized/models/prompt_processors/base.py", ["MultiPromptProcessorOutput"], mns)
    P, Bm = 3, 6
    vd_table = half(torch.randn(P, 4, T, D, generator=g))
    local_table = half(torch.randn(P, T, D, generator=g))
    global_table = half(torch.randn(P, D, generator=g))
    prompt_idx = torch.tensor([2, 0, 1, 1, 0, 2])
    mout = mns["MultiPromptProcessorOutput"](
        global_text_embeddings=[global_table[i] for i in prompt_idx], local_text_embeddings=[local_table[i] for i in prompt_idx],
        uncond_text_embeddings=tables["uncond_text_embeddings"][0], text_embeddings_vd=[vd_table[i] for i in prompt_idx],
        uncond_text_embeddings_vd=tables["uncond_text_embeddings_vd"], directions=directions_list,
        direction2idx={d.name: i for i, d in enumerate(directions_list)}, use_perp_neg=True, device="cpu", **coeff)
    el_m, az_m = elevation[[0, 4, 8, 3, 10, 12]], azimuth[[0, 4, 8, 3, 10, 12]]
    pn_m, w_m = mout.get_text_embeddings_perp_neg(el_m, az_m, dist[:Bm], True)
    gold["multi_prompt"] = {"vd_table": vd_table, "local_table": local_table, "global_table": global_table,
                            "prompt_idx": prompt_idx, "elevation": el_m, "azimuth": az_m,
                            "uncond": tables["uncond_text_embeddings"][0], "uncond_vd": tables["uncond_text_embeddings_vd"],
                            "vd": mout.get_text_embeddings(el_m, az_m, dist[:Bm], True),
                            "local": mout.get_text_embeddings(el_m, az_m, dist[:Bm], False).contiguous(),
                            "global": mout.get_global_text_embeddings(), "perp_neg": pn_m, "neg_weights": w_m.float(),
                            "perp_neg_scaled": mout.get_text_embeddings_perp_neg(el_m, az_m, dist[:Bm], True, 0.5)[1].float()}
    gold["guidance"] = guidance_case(ns, out, elevation[[3, 8]], azimuth[[3, 8]], dist[[3, 8]])
    gold["mv_guidance"] = mv_guidance_case(ns, out, gold["rays"]["c2w"])
    torch.save(gold, OUT)
    print("wrote", OUT, len(cases), "C cases;", "perp-neg", tuple(pn.shape), tuple(w.shape), coeff)


if __name__ == "__main__":
    main()
