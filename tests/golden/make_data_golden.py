"""Golden camera batches produced by the REFERENCE's own data-set classes under fixed seeds.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_data_golden.py
threestudio/data/uncond.py (RandomCameraDataModuleConfig, RandomCameraIterableDataset, RandomCameraDataset) and
threestudio/data/uncond_multiview.py (RandomMultiviewCameraDataModuleConfig, RandomMultiviewCameraIterableDataset) are
taken out of the files by name with `ast` and executed unchanged (the modules themselves import pytorch-lightning / cv2);
their helpers come from threestudio/utils/ops.py the same way. `random.seed(s); torch.manual_seed(s)` precede every
`collate`, so a port that consumes the two generators in the same order reproduces the batch bit for bit.
Output: tests/golden/data_golden.pt (rays are kept as strided samples; about 150 kB).
"""
import ast
import bisect
import math
import os
import random
import types
from dataclasses import dataclass, field
from typing import Any, Dict, List, Tuple

import torch
import torch.nn.functional as F

REF = "/root/reference/threestudio"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_golden.pt")


class _Any:
    def __getitem__(self, item):
        return self


def pieces(path, names, ns):
    s = open(path).read()
    found = set()
    lines = s.splitlines()
    for node in ast.parse(s).body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            start = node.decorator_list[0].lineno - 1 if node.decorator_list else node.lineno - 1
            exec(compile("\n".join(lines[start:node.end_lineno]), path, "exec"), ns)
            found.add(node.name)
    assert found == set(names), set(names) - found


def namespace():
    ns = {"torch": torch, "F": F, "math": math, "random": random, "bisect": bisect, "dataclass": dataclass, "field": field,
          "Any": Any, "Dict": Dict, "List": List, "Tuple": Tuple, "Float": _Any(), "Tensor": torch.Tensor, "Union": _Any(),
          "Optional": _Any(), "IterableDataset": type("IterableDataset", (), {}), "Dataset": type("Dataset", (), {}),
          "Updateable": type("Updateable", (), {}),
          "threestudio": types.SimpleNamespace(debug=lambda *a, **k: None, warn=lambda *a, **k: None)}
    pieces(f"{REF}/utils/ops.py", ["get_ray_directions", "get_rays", "get_projection_matrix", "get_mvp_matrix"], ns)
    pieces(f"{REF}/data/uncond.py", ["RandomCameraDataModuleConfig", "RandomCameraIterableDataset", "RandomCameraDataset"], ns)
    pieces(f"{REF}/data/uncond_multiview.py", ["RandomMultiviewCameraDataModuleConfig",
                                               "RandomMultiviewCameraIterableDataset"], ns)
    return ns


def keep(batch):
    out = {}
    for k, v in batch.items():
        if k in ("rays_o", "rays_d"):
            out[k + "_sample"] = v.reshape(-1, 3)[::37].clone()
        elif torch.is_tensor(v):
            out[k] = v.clone()
        else:
            out[k] = v
    return out


def main():
    ns = namespace()
    gold = {}
    c2 = dict(batch_size=[1, 1], width=[64, 256], height=[64, 256], resolution_milestones=[10000],
              camera_distance_range=[1.0, 1.5], fovy_range=[40, 70], elevation_range=[-10, 45], camera_perturb=0.0,
              center_perturb=0.0, up_perturb=0.0, eval_camera_distance=1.2, eval_fovy_deg=70.0, n_val_views=30)
    generic = dict(batch_size=5, width=24, height=16, light_sample_strategy="magic3d", batch_uniform_azimuth=False,
                   progressive_until=100)
    cases = []
    for name, cfg_kw, cls, cfg_cls, steps in (
            ("c2", c2, "RandomCameraIterableDataset", "RandomCameraDataModuleConfig", (0, 10000)),
            ("generic", generic, "RandomCameraIterableDataset", "RandomCameraDataModuleConfig", (0, 40, 500)),
            ("mv", dict(batch_size=[8, 4], n_view=4, width=[16, 32], height=[16, 32], resolution_milestones=[5000],
                        camera_distance_range=[0.8, 1.0], fovy_range=[15, 60], elevation_range=[0, 30],
                        camera_perturb=0.0, center_perturb=0.0, up_perturb=0.0, n_val_views=4, eval_camera_distance=3.0,
                        eval_fovy_deg=40.0, relative_radius=True),
             "RandomMultiviewCameraIterableDataset", "RandomMultiviewCameraDataModuleConfig", (0, 5000)),
            ("mv_zoom", dict(batch_size=4, n_view=2, width=16, height=16, relative_radius=False, zoom_range=[0.8, 1.2],
                             light_sample_strategy="magic3d"),
             "RandomMultiviewCameraIterableDataset", "RandomMultiviewCameraDataModuleConfig", (0,))):
        cfg = ns[cfg_cls](**cfg_kw)
        ds = ns[cls](cfg)
        batches = []
        for step in steps:
            ds.update_step(0, step)
            for seed in (11, 12, 13, 14):
                random.seed(seed)
                torch.manual_seed(seed)
                batches.append({"step": step, "seed": seed, "batch": keep(ds.collate({}))})
        cases.append(name)
        gold[name] = {"config": cfg_kw, "batches": batches}
    # multi-prompt variants (custom/amortized/data/multiprompt.py:61-83, multiview_multiprompt.py:51-77): base collate +
    # generator noise + prompts drawn with random.choices / random.sample
    pieces("/root/reference/custom/amortized/data/multiprompt.py",
           ["MultipromptRandomCameraDataModuleConfig", "MultipromptRandomCameraIterableDataset"], ns)
    pieces("/root/reference/custom/amortized/data/multiview_multiprompt.py",
           ["MultiviewMultipromptRandomCameraDataModuleConfig", "MultiviewMultipromptRandomCameraIterableDataset"], ns)
    library = {"train": [f"prompt number {i}" for i in range(5)]}
    for name, cfg_cls, cls, cfg_kw in (
            ("multiprompt", "MultipromptRandomCameraDataModuleConfig", "MultipromptRandomCameraIterableDataset",
             # batch_size 3 is avoided on purpose: the reference calls torch.cross(lookat, up) without `dim`, which picks the
             # FIRST axis of size 3 -- for a [3, 3] batch that is the batch axis, and the cameras come out wrong
             dict(batch_size=4, width=16, height=16, dim_gaussian=8)),
            ("multiprompt_more_than_library", "MultipromptRandomCameraDataModuleConfig",
             "MultipromptRandomCameraIterableDataset", dict(batch_size=7, width=16, height=16, dim_gaussian=4)),
            ("multiview_multiprompt", "MultiviewMultipromptRandomCameraDataModuleConfig",
             "MultiviewMultipromptRandomCameraIterableDataset",
             dict(batch_size=8, n_view=4, width=16, height=16, dim_gaussian=8, camera_distance_range=[0.8, 1.0],
                  fovy_range=[15, 60], elevation_range=[0, 30], camera_perturb=0.0, center_perturb=0.0, up_perturb=0.0))):
        ds = ns[cls](ns[cfg_cls](**cfg_kw), library)
        batches = []
        for seed in (21, 22, 23):
            random.seed(seed)
            torch.manual_seed(seed)
            batches.append({"step": 0, "seed": seed, "batch": keep(ds.collate({}))})
        gold[name] = {"config": cfg_kw, "library": library, "batches": batches}
    # evaluation orbit
    cfg = ns["RandomCameraDataModuleConfig"](eval_height=20, eval_width=28, n_val_views=5, n_test_views=7,
                                             eval_elevation_deg=15.0, eval_camera_distance=1.2, eval_fovy_deg=70.0)
    for split in ("val", "test"):
        ds = ns["RandomCameraDataset"](cfg, split)
        gold[f"eval_{split}"] = {"config": dict(eval_height=20, eval_width=28, n_val_views=5, n_test_views=7,
                                                 eval_elevation_deg=15.0, eval_camera_distance=1.2, eval_fovy_deg=70.0),
                                 "items": [keep(ds[i]) for i in range(len(ds))]}
    torch.save(gold, OUT)
    print("wrote", OUT, {k: len(v.get("batches", v.get("items"))) for k, v in gold.items()})


if __name__ == "__main__":
    main()
