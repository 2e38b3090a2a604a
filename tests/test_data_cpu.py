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
