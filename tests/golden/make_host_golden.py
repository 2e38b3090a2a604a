"""Golden vectors for the small host-side / glue pieces of the path, produced by the REFERENCE's own definitions.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_host_golden.py
The reference modules cannot be imported (pytorch-lightning, tinycudann, igl ... at module level), so the definitions
are taken out of the files by name with `ast`, compiled unchanged and run against a namespace that holds only torch /
math and inert stand-ins for the type annotations:
    threestudio/utils/misc.py                       C  (scheduled scalars)
    threestudio/utils/ops.py                        get_ray_directions, get_rays, get_projection_matrix, get_mvp_matrix,
                                                    shifted_expotional_decay
    threestudio/models/prompt_processors/base.py    DirectionConfig, PromptProcessorOutput, shift_azimuth_deg, the
                                                    `self.directions = [...]` list and the Perp-Neg defaults of
                                                    PromptProcessor.Config
Output: tests/golden/host_golden.pt (a few tens of kB).
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
    torch.save(gold, OUT)
    print("wrote", OUT, len(cases), "C cases;", "perp-neg", tuple(pn.shape), tuple(w.shape), coeff)


if __name__ == "__main__":
    main()
