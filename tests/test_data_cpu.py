"""Host-side logic of the camera data modules (no GPU): the evaluation orbit of threestudio/data/uncond.py:347-467 and the
loader plumbing around it. Ray generation itself runs on the device (tests/test_render_gpu.py::test_raygen_parity)."""
import math

import torch

from oracle import render_oracle as ro


def _dm(**over):
    import scaledreamer_b200 as sd

    cfg = {"eval_height": 48, "eval_width": 64, "n_val_views": 4, "n_test_views": 7, "eval_elevation_deg": 15.0,
           "eval_camera_distance": 1.2, "eval_fovy_deg": 70.0, "eval_batch_size": 1}
    cfg.update(over)
    return sd.find("random-camera-datamodule")(cfg)


def test_eval_orbit_azimuths_and_batches():
    dm = _dm()
    dm.setup(None)
    val, test = dm.val_dataset, dm.test_dataset
    assert len(val) == 4 and len(test) == 7
    # val: first and last view differ; test: the loop closes (uncond.py:358-363)
    torch.testing.assert_close(val.azimuth_deg, torch.tensor([0.0, 90.0, 180.0, 270.0]))
    torch.testing.assert_close(test.azimuth_deg, torch.linspace(0, 360.0, 7))
    batches = list(dm.val_dataloader())
    assert len(batches) == 4
    b = batches[2]
    assert b["c2w"].shape == (1, 4, 4) and b["mvp_mtx"].shape == (1, 4, 4) and int(b["index"][0]) == 2
    assert b["height"] == 48 and b["width"] == 64
    assert float(b["elevation"][0]) == 15.0 and abs(float(b["camera_distances"][0]) - 1.2) < 1e-6
    assert abs(float(b["fovy"][0]) - math.radians(70.0)) < 1e-6
    assert len(list(_dm(eval_batch_size=3).test_dataloader())) == 3  # 7 views in batches of 3: 3 + 3 + 1


def test_eval_cameras_look_at_origin_and_match_oracle():
    dm = _dm()
    dm.setup("validate")
    ds = dm.val_dataset
    c2w = ds.c2w
    R, t = c2w[:, :3, :3], c2w[:, :3, 3]
    torch.testing.assert_close(R @ R.transpose(1, 2), torch.eye(3).expand(4, 3, 3), atol=1e-6, rtol=0)
    torch.testing.assert_close(t.norm(dim=-1), torch.full((4,), 1.2), atol=1e-6, rtol=0)
    # the optical axis (-z column) points at the origin; +z is up
    torch.testing.assert_close(-R[:, :, 2], -t / t.norm(dim=-1, keepdim=True), atol=1e-6, rtol=0)
    assert (R[:, 2, 1] > 0).all()
    ref = ro.look_at_c2w(ds.elevation_deg, ds.azimuth_deg, ds.camera_distances)
    torch.testing.assert_close(c2w, ref, atol=1e-6, rtol=0)
    torch.testing.assert_close(ds.light_positions, t)


import os

import pytest

DATA_GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_golden.pt"))
TENSOR_KEYS = ("mvp_mtx", "camera_positions", "c2w", "light_positions", "elevation", "azimuth", "camera_distances",
               "fovy", "proj_mtx")


@pytest.mark.parametrize("case,plugin", [("c2", "random-camera-datamodule"), ("generic", "random-camera-datamodule"),
                                          ("mv", "mvdream-random-multiview-camera-datamodule"),
                                          ("mv_zoom", "mvdream-random-multiview-camera-datamodule")])
def test_training_batches_match_reference_datasets_bit_for_bit(case, plugin):
    """Same `random` / torch generator consumption as RandomCameraIterableDataset.collate (uncond.py:143-344) and the
    multi-view variant (uncond_multiview.py:41-255): with the seeds of tests/golden/make_data_golden.py every camera
    tensor of the batch equals the reference's, incl. resolution milestones, progressive view ranges, both light
    strategies, relative radius and zoom."""
    import random

    import scaledreamer_b200 as sd

    g = DATA_GOLD[case]
    dm = sd.find(plugin)(dict(g["config"]))
    dm.setup("fit")
    ds = dm.train_dataset
    last_step = None
    for rec in g["batches"]:
        if rec["step"] != last_step:
            ds.update_step(0, rec["step"])
            last_step = rec["step"]
        random.seed(rec["seed"])
        torch.manual_seed(rec["seed"])
        b = ds.collate({})
        ref = rec["batch"]
        assert b["height"] == ref["height"] and b["width"] == ref["width"], (case, rec["step"])
        for k in TENSOR_KEYS:
            if k not in ref:  # the multi-view batch carries no proj_mtx
                continue
            torch.testing.assert_close(b[k], ref[k], atol=0, rtol=0, msg=lambda m: f"{case} step {rec['step']} {k}: {m}")
        assert set(ref) - {"rays_o_sample", "rays_d_sample"} <= set(b), set(ref) - set(b)


def test_eval_orbit_matches_reference_dataset():
    import scaledreamer_b200 as sd

    for split in ("val", "test"):
        g = DATA_GOLD[f"eval_{split}"]
        dm = sd.find("random-camera-datamodule")(dict(g["config"]))
        dm.setup(None)
        ds = dm.val_dataset if split == "val" else dm.test_dataset
        assert len(ds) == len(g["items"])
        for i, ref in enumerate(g["items"]):
            it = ds[i]
            for k in ("mvp_mtx", "c2w", "camera_positions", "light_positions", "elevation", "azimuth",
                      "camera_distances", "fovy", "proj_mtx"):
                torch.testing.assert_close(it[k], ref[k], atol=0, rtol=0, msg=lambda m: f"{split} {i} {k}: {m}")
            assert it["index"] == ref["index"] and it["height"] == ref["height"] and it["width"] == ref["width"]


@pytest.mark.parametrize("case", ["multiprompt", "multiprompt_more_than_library", "multiview_multiprompt"])
def test_multiprompt_batches_match_reference_datasets(case):
    """custom/amortized/data/multiprompt.py:61-83 and multiview_multiprompt.py:51-77: camera block as above plus the
    generator noise and the prompts drawn from the library (random.sample, or random.choices when the batch is larger
    than the library)."""
    import random

    from scaledreamer_b200 import amortized as A
    from scaledreamer_b200.core import parse_structured

    g = DATA_GOLD[case]
    mv = case.startswith("multiview")
    cfg_cls = A.MultiviewMultipromptRandomCameraDataModuleConfig if mv else A.MultipromptRandomCameraDataModuleConfig
    ds_cls = A.MultiviewMultipromptRandomCameraIterableDataset if mv else A.MultipromptRandomCameraIterableDataset
    ds = ds_cls(parse_structured(cfg_cls, dict(g["config"])), g["library"])
    for rec in g["batches"]:
        random.seed(rec["seed"])
        torch.manual_seed(rec["seed"])
        b = ds.collate({})
        ref = rec["batch"]
        assert b["prompt"] == ref["prompt"]
        torch.testing.assert_close(b["noise"], ref["noise"], atol=0, rtol=0)
        for k in TENSOR_KEYS:
            if k in ref:
                torch.testing.assert_close(b[k], ref[k], atol=0, rtol=0, msg=lambda m: f"{case} {k}: {m}")


def test_launch_save_views_writes_rgb_opacity_depth_rows(tmp_path):
    """launch.save_views: one PNG per evaluation view, rgb | opacity | depth side by side."""
    import sys

    from PIL import Image

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import launch

    H, W = 6, 10
    outs = [{"index": torch.tensor([i]), "comp_rgb": torch.rand(1, H, W, 3), "opacity": torch.rand(1, H, W, 1),
             "depth": torch.rand(H, W)} for i in (0, 3)]
    outs[1].pop("depth")
    launch.save_views(outs, str(tmp_path / "val"))
    assert sorted(os.listdir(tmp_path / "val")) == ["0.png", "3.png"]
    a = Image.open(tmp_path / "val" / "0.png")
    assert a.size == (3 * W, H)
    assert Image.open(tmp_path / "val" / "3.png").size == (2 * W, H)
    import numpy as np

    px = torch.from_numpy(np.asarray(a).copy()).float() / 255.0
    torch.testing.assert_close(px[:, :W], outs[0]["comp_rgb"][0], atol=0.5 / 255 + 1e-6, rtol=0)
    torch.testing.assert_close(px[:, W:2 * W, 0], outs[0]["opacity"][0, :, :, 0], atol=0.5 / 255 + 1e-6, rtol=0)
    torch.testing.assert_close(px[:, 2 * W:, 1], outs[0]["depth"], atol=0.5 / 255 + 1e-6, rtol=0)


EVAL_GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_data_golden.pt"))
EVAL_KEYS = ("mvp_mtx", "c2w", "camera_positions", "light_positions", "elevation", "azimuth", "camera_distances", "fovy",
             "proj_mtx", "noise", "ratio")


@pytest.mark.parametrize("i", range(len(EVAL_GOLD["cases"])))
def test_multiprompt_evaluation_batches_match_reference_datasets(i, tmp_path, monkeypatch):
    """custom/amortized/data/multiprompt.py:85-164 through the data modules' val / test loaders: per-prompt batches that
    hold the whole orbit (library split), or one view per batch for `eval_prompt` (with `target_prompt` interpolation
    ratios and `eval_fix_camera`, including the reference's `if self.fix_camera` treatment of camera 0)."""
    import json

    import scaledreamer_b200 as sd

    c = EVAL_GOLD["cases"][i]
    os.makedirs(tmp_path / "load")
    json.dump(EVAL_GOLD["library"], open(tmp_path / "load" / "lib.json", "w"))
    monkeypatch.chdir(tmp_path)
    for name in ("multiprompt-camera-datamodule", "multiprompt-multiview-camera-datamodule"):
        dm = sd.find(name)(dict(c["config"], prompt_library="lib"))
        if "seed" in c:
            torch.manual_seed(c["seed"])
        dm.setup("validate" if c["split"] == "val" else "test")
        loader = dm.val_dataloader() if c["split"] == "val" else dm.test_dataloader()
        batches = list(loader)
        assert len(batches) == len(c["batches"])
        for b, ref in zip(batches, c["batches"]):
            for k in ("prompt", "prompt_target", "name"):
                assert b.get(k) == ref.get(k), k
            assert b["index"].tolist() == ref["index"].tolist()
            assert b["height"] == int(ref["height"][0]) and b["width"] == int(ref["width"][0])
            for k in EVAL_KEYS:
                assert (k in b) == (k in ref), k
                if k in ref:
                    torch.testing.assert_close(b[k], ref[k], atol=0, rtol=0, msg=lambda m: f"{c['kind']} {k}: {m}")
