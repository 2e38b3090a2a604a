"""Random-camera data modules (threestudio names and Config keys).

  "random-camera-datamodule"                     threestudio/data/uncond.py:470 (collate :143-344)
  "mvdream-random-multiview-camera-datamodule"   threestudio/data/uncond_multiview.py:258 (collate :41-255)

Camera scalars are sampled on the host with the reference's draw order (Python `random` for the elevation branch,
then the CPU torch generator), so a seeded run consumes the same random stream. The per-pixel rays are NOT built
on the host: `to_device()` uploads the ~100-byte camera block and sdb_raygen writes rays_o / rays_d
(utils/ops.py:183-269) straight into HBM -- 1.5 MB of H2D per 256x256 view in the reference.
"""
from __future__ import annotations

import bisect
import math
import random
from dataclasses import dataclass, field
from typing import Any, Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import lib as L
from .core import Updateable, parse_structured, register


@dataclass
class RandomCameraDataModuleConfig:
    height: Any = 64
    width: Any = 64
    batch_size: Any = 1
    resolution_milestones: List[int] = field(default_factory=lambda: [])
    eval_height: int = 512
    eval_width: int = 512
    eval_batch_size: int = 1
    n_val_views: int = 1
    n_test_views: int = 120
    elevation_range: Tuple[float, float] = (-10, 90)
    azimuth_range: Tuple[float, float] = (-180, 180)
    camera_distance_range: Tuple[float, float] = (1, 1.5)
    fovy_range: Tuple[float, float] = (40, 70)
    camera_perturb: float = 0.1
    center_perturb: float = 0.2
    up_perturb: float = 0.02
    light_position_perturb: float = 1.0
    light_distance_range: Tuple[float, float] = (0.8, 1.5)
    eval_elevation_deg: float = 15.0
    eval_camera_distance: float = 1.5
    eval_fovy_deg: float = 70.0
    light_sample_strategy: str = "dreamfusion"
    batch_uniform_azimuth: bool = True
    progressive_until: int = 0
    rays_d_normalize: bool = True


@dataclass
class RandomMultiviewCameraDataModuleConfig(RandomCameraDataModuleConfig):
    relative_radius: bool = True
    n_view: int = 1
    zoom_range: Tuple[float, float] = (1.0, 1.0)


def get_projection_matrix(fovy: torch.Tensor, aspect_wh: float, near: float, far: float) -> torch.Tensor:
    b = fovy.shape[0]
    p = torch.zeros(b, 4, 4, dtype=torch.float32)
    p[:, 0, 0] = 1.0 / (torch.tan(fovy / 2.0) * aspect_wh)
    p[:, 1, 1] = -1.0 / torch.tan(fovy / 2.0)
    p[:, 2, 2] = -(far + near) / (far - near)
    p[:, 2, 3] = -2.0 * far * near / (far - near)
    p[:, 3, 2] = -1.0
    return p


def get_mvp_matrix(c2w: torch.Tensor, proj_mtx: torch.Tensor) -> torch.Tensor:
    w2c = torch.zeros(c2w.shape[0], 4, 4).to(c2w)
    w2c[:, :3, :3] = c2w[:, :3, :3].permute(0, 2, 1)
    w2c[:, :3, 3:] = -c2w[:, :3, :3].permute(0, 2, 1) @ c2w[:, :3, 3:]
    w2c[:, 3, 3] = 1.0
    return proj_mtx @ w2c


def look_at(camera_positions, center, up):
    lookat = F.normalize(center - camera_positions, dim=-1)
    right = F.normalize(torch.cross(lookat, up, dim=-1), dim=-1)
    up = F.normalize(torch.cross(right, lookat, dim=-1), dim=-1)
    c2w3x4 = torch.cat([torch.stack([right, up, -lookat], dim=-1), camera_positions[:, :, None]], dim=-1)
    c2w = torch.cat([c2w3x4, torch.zeros_like(c2w3x4[:, :1])], dim=1)
    c2w[:, 3, 3] = 1.0
    return c2w


def rays_on_device(c2w: torch.Tensor, fovy: torch.Tensor, height: int, width: int, device):
    """c2w [B,4,4], fovy [B] radians (host or device) -> rays_o, rays_d [B,H,W,3] on device via sdb_raygen."""
    c2w_d = c2w.to(device, torch.float32, non_blocking=True).contiguous()
    fovy_d = fovy.to(device, torch.float32, non_blocking=True).contiguous()
    B = c2w_d.shape[0]
    rays_o = torch.empty(B, height, width, 3, device=device)
    rays_d = torch.empty(B, height, width, 3, device=device)
    L.check(L.load().sdb_raygen(L.ptr(c2w_d), L.ptr(fovy_d), B, height, width, L.ptr(rays_o), L.ptr(rays_d),
                                L.stream_ptr()), "sdb_raygen")
    return rays_o, rays_d, c2w_d, fovy_d


class RandomCameraIterableDataset(Updateable):
    config_cls = RandomCameraDataModuleConfig

    def __init__(self, cfg: Any) -> None:
        super().__init__()
        self.cfg = cfg
        as_list = lambda v: [v] if isinstance(v, int) else list(v)
        self.heights, self.widths, self.batch_sizes = as_list(cfg.height), as_list(cfg.width), as_list(cfg.batch_size)
        assert len(self.heights) == len(self.widths) == len(self.batch_sizes)
        if len(self.heights) == 1:
            self.resolution_milestones = [-1]
        else:
            assert len(self.heights) == len(cfg.resolution_milestones) + 1
            self.resolution_milestones = [-1] + list(cfg.resolution_milestones)
        self.height, self.width, self.batch_size = self.heights[0], self.widths[0], self.batch_sizes[0]
        self.elevation_range = list(cfg.elevation_range)
        self.azimuth_range = list(cfg.azimuth_range)
        self.camera_distance_range = list(cfg.camera_distance_range)
        self.fovy_range = list(cfg.fovy_range)

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        i = bisect.bisect_right(self.resolution_milestones, global_step) - 1
        self.height, self.width, self.batch_size = self.heights[i], self.widths[i], self.batch_sizes[i]
        r = min(1.0, global_step / (self.cfg.progressive_until + 1))  # progressive view ranges (uncond.py:122-131)
        e0 = self.cfg.eval_elevation_deg
        self.elevation_range = [(1 - r) * e0 + r * self.cfg.elevation_range[0],
                                (1 - r) * e0 + r * self.cfg.elevation_range[1]]
        self.azimuth_range = [r * self.cfg.azimuth_range[0], r * self.cfg.azimuth_range[1]]

    def __iter__(self):
        while True:
            yield {}

    # ---- sampling pieces shared with the multi-view variant ----
    def _light(self, camera_positions, n_draw, rep):
        cfg = self.cfg
        light_distances = (torch.rand(n_draw) * (cfg.light_distance_range[1] - cfg.light_distance_range[0])
                           + cfg.light_distance_range[0]).repeat_interleave(rep, dim=0)
        if cfg.light_sample_strategy == "dreamfusion":
            d = F.normalize(camera_positions + torch.randn(n_draw, 3).repeat_interleave(rep, dim=0)
                            * cfg.light_position_perturb, dim=-1)
            return d * light_distances[:, None]
        if cfg.light_sample_strategy == "magic3d":
            local_z = F.normalize(camera_positions, dim=-1)
            local_x = F.normalize(torch.stack([local_z[:, 1], -local_z[:, 0], torch.zeros_like(local_z[:, 0])], -1), dim=-1)
            local_y = F.normalize(torch.cross(local_z, local_x, dim=-1), dim=-1)
            rot = torch.stack([local_x, local_y, local_z], dim=-1)
            if rep == 1:
                az = torch.rand(n_draw) * math.pi * 2 - math.pi
            else:  # the multi-view variant draws rand*pi - 2*pi (uncond_multiview.py:185-187)
                az = (torch.rand(n_draw) * math.pi - 2 * math.pi).repeat_interleave(rep, dim=0)
            el = (torch.rand(n_draw) * math.pi / 3 + math.pi / 6).repeat_interleave(rep, dim=0)
            local = torch.stack([light_distances * torch.cos(el) * torch.cos(az),
                                 light_distances * torch.cos(el) * torch.sin(az), light_distances * torch.sin(el)], -1)
            return (rot @ local[:, :, None])[:, :, 0]
        raise ValueError(f"Unknown light sample strategy: {cfg.light_sample_strategy}")

    def collate(self, batch=None) -> Dict[str, Any]:
        cfg, B = self.cfg, self.batch_size
        if random.random() < 0.5:  # uniform in elevation (biased towards the poles)
            elevation_deg = torch.rand(B) * (self.elevation_range[1] - self.elevation_range[0]) + self.elevation_range[0]
            elevation = elevation_deg * math.pi / 180
        else:  # uniform on the sphere
            lo, hi = self.elevation_range[0] / 180.0 * math.pi, self.elevation_range[1] / 180.0 * math.pi
            elevation = torch.asin(torch.rand(B) * (math.sin(hi) - math.sin(lo)) + math.sin(lo))
            elevation_deg = elevation / math.pi * 180.0
        if cfg.batch_uniform_azimuth:
            azimuth_deg = (torch.rand(B) + torch.arange(B)) / B * (self.azimuth_range[1] - self.azimuth_range[0]) \
                + self.azimuth_range[0]
        else:
            azimuth_deg = torch.rand(B) * (self.azimuth_range[1] - self.azimuth_range[0]) + self.azimuth_range[0]
        azimuth = azimuth_deg * math.pi / 180
        camera_distances = torch.rand(B) * (self.camera_distance_range[1] - self.camera_distance_range[0]) \
            + self.camera_distance_range[0]
        camera_positions = torch.stack([camera_distances * torch.cos(elevation) * torch.cos(azimuth),
                                        camera_distances * torch.cos(elevation) * torch.sin(azimuth),
                                        camera_distances * torch.sin(elevation)], dim=-1)
        center = torch.zeros_like(camera_positions)
        up = torch.as_tensor([0, 0, 1], dtype=torch.float32)[None, :].repeat(B, 1)
        camera_positions = camera_positions + (torch.rand(B, 3) * 2 * cfg.camera_perturb - cfg.camera_perturb)
        center = center + torch.randn(B, 3) * cfg.center_perturb
        up = up + torch.randn(B, 3) * cfg.up_perturb
        fovy_deg = torch.rand(B) * (self.fovy_range[1] - self.fovy_range[0]) + self.fovy_range[0]
        fovy = fovy_deg * math.pi / 180
        light_positions = self._light(camera_positions, B, 1)
        c2w = look_at(camera_positions, center, up)
        proj_mtx = get_projection_matrix(fovy, self.width / self.height, 0.01, 100.0)
        return {"mvp_mtx": get_mvp_matrix(c2w, proj_mtx), "camera_positions": camera_positions, "c2w": c2w,
                "light_positions": light_positions, "elevation": elevation_deg, "azimuth": azimuth_deg,
                "camera_distances": camera_distances, "height": self.height, "width": self.width, "fovy": fovy,
                "proj_mtx": proj_mtx, "_fovy_rad": fovy}

    def to_device(self, batch: Dict[str, Any], device) -> Dict[str, Any]:
        """Uploads the camera block and generates the rays on device (adds rays_o / rays_d). Every floating-point tensor of
        the batch is packed into ONE pinned staging buffer and crosses the bus as ONE asynchronous copy (the reference
        moves ten small pageable tensors plus 1.5 MB of rays per view); the device-side tensors are views of that block."""
        if not self.cfg.rays_d_normalize:
            raise NotImplementedError("rays_d_normalize=false is not supported by the device ray generator")
        batch = dict(batch)
        fovy_rad = batch.pop("_fovy_rad")
        out = _stage_to_device({**batch, "_fovy_rad": fovy_rad}, torch.device(device))
        fovy_d = out.pop("_fovy_rad")
        rays_o, rays_d, c2w_d, _ = rays_on_device(out["c2w"], fovy_d, batch["height"], batch["width"], device)
        out["c2w"], out["rays_o"], out["rays_d"] = c2w_d, rays_o, rays_d
        return out


_STAGE_SLOTS = 4  # the host may run a few batches ahead of the stream: a slot is reused only after its copy has completed
_stage: Dict[Any, Any] = {}


def _stage_to_device(batch: Dict[str, Any], device: torch.device) -> Dict[str, Any]:
    """fp32 tensors -> one pinned ring slot -> one H2D copy -> views; everything else (ints, strings, index tensors) as is."""
    floats = [(k, v) for k, v in batch.items() if torch.is_tensor(v) and v.dtype == torch.float32 and not v.is_cuda]
    out = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()
           if not (torch.is_tensor(v) and v.dtype == torch.float32 and not v.is_cuda)}
    if not floats or device.type != "cuda":
        out.update({k: v.to(device) for k, v in floats})
        return out
    pad4 = lambda k: (k + 3) & ~3  # every tensor starts on a 16-byte boundary of the block
    n = sum(pad4(v.numel()) for _, v in floats)
    st = _stage.setdefault(device, {"bufs": [None] * _STAGE_SLOTS, "events": [None] * _STAGE_SLOTS, "next": 0})
    i = st["next"]
    st["next"] = (i + 1) % _STAGE_SLOTS
    if st["bufs"][i] is None or st["bufs"][i].numel() < n:
        st["bufs"][i] = torch.empty(max(n, 1024), dtype=torch.float32).pin_memory()
        st["events"][i] = torch.cuda.Event()
    else:
        st["events"][i].synchronize()  # the copy that last used this slot has finished
    host = st["bufs"][i]
    off = 0
    for _, v in floats:
        host[off:off + v.numel()].copy_(v.reshape(-1))
        off += pad4(v.numel())
    dev = torch.empty(n, dtype=torch.float32, device=device)
    dev.copy_(host[:n], non_blocking=True)
    st["events"][i].record()
    off = 0
    for k, v in floats:
        out[k] = dev[off:off + v.numel()].view(v.shape)
        off += pad4(v.numel())
    return out


class RandomMultiviewCameraIterableDataset(RandomCameraIterableDataset):
    config_cls = RandomMultiviewCameraDataModuleConfig

    def __init__(self, cfg):
        super().__init__(cfg)
        self.zoom_range = list(cfg.zoom_range)

    def collate(self, batch=None) -> Dict[str, Any]:
        cfg, B, V = self.cfg, self.batch_size, self.cfg.n_view
        assert B % V == 0, f"batch_size ({B}) must be dividable by n_view ({V})!"
        R = B // V
        rep = lambda t: t.repeat_interleave(V, dim=0)
        if random.random() < 0.5:
            elevation_deg = rep(torch.rand(R) * (self.elevation_range[1] - self.elevation_range[0]) + self.elevation_range[0])
            elevation = elevation_deg * math.pi / 180
        else:
            lo, hi = (self.elevation_range[0] + 90.0) / 180.0, (self.elevation_range[1] + 90.0) / 180.0
            elevation = rep(torch.asin(2 * (torch.rand(R) * (hi - lo) + lo) - 1.0))
            elevation_deg = elevation / math.pi * 180.0
        azimuth_deg = (torch.rand(R).reshape(-1, 1) + torch.arange(V).reshape(1, -1)).reshape(-1) / V \
            * (self.azimuth_range[1] - self.azimuth_range[0]) + self.azimuth_range[0]
        azimuth = azimuth_deg * math.pi / 180
        fovy_deg = rep(torch.rand(R) * (self.fovy_range[1] - self.fovy_range[0]) + self.fovy_range[0])
        fovy = fovy_deg * math.pi / 180
        camera_distances = rep(torch.rand(R) * (self.camera_distance_range[1] - self.camera_distance_range[0])
                               + self.camera_distance_range[0])
        if cfg.relative_radius:
            camera_distances = (1 / torch.tan(0.5 * fovy)) * camera_distances  # the reference's rounding order
        zoom = rep(torch.rand(R) * (self.zoom_range[1] - self.zoom_range[0]) + self.zoom_range[0])
        fovy, fovy_deg = fovy * zoom, fovy_deg * zoom
        camera_positions = torch.stack([camera_distances * torch.cos(elevation) * torch.cos(azimuth),
                                        camera_distances * torch.cos(elevation) * torch.sin(azimuth),
                                        camera_distances * torch.sin(elevation)], dim=-1)
        center = torch.zeros_like(camera_positions)
        up = torch.as_tensor([0, 0, 1], dtype=torch.float32)[None, :].repeat(B, 1)
        camera_positions = camera_positions + rep(torch.rand(R, 3) * 2 * cfg.camera_perturb - cfg.camera_perturb)
        center = center + rep(torch.randn(R, 3) * cfg.center_perturb)
        up = up + rep(torch.randn(R, 3) * cfg.up_perturb)
        light_positions = self._light(camera_positions, R, V)
        c2w = look_at(camera_positions, center, up)
        proj_mtx = get_projection_matrix(fovy, self.width / self.height, 0.1, 1000.0)
        return {"mvp_mtx": get_mvp_matrix(c2w, proj_mtx), "camera_positions": camera_positions, "c2w": c2w,
                "light_positions": light_positions, "elevation": elevation_deg, "azimuth": azimuth_deg,
                "camera_distances": camera_distances, "height": self.height, "width": self.width, "fovy": fovy_deg,
                "_fovy_rad": fovy}


class RandomCameraDataset:
    """Evaluation orbit (uncond.py:347-467): n_val_views / n_test_views cameras at eval_elevation_deg,
    eval_camera_distance, eval_fovy_deg; `val` spreads the azimuths so that the first and last view differ
    (linspace(0, 360, n+1)[:n]), `test` closes the loop (linspace(0, 360, n)). Camera scalars live on the host; rays are
    generated on device by `to_device` like the training batches."""

    def __init__(self, cfg: Any, split: str) -> None:
        self.cfg, self.split = cfg, split
        n = self.n_views = cfg.n_val_views if split == "val" else cfg.n_test_views
        azimuth_deg = torch.linspace(0, 360.0, n + 1)[:n] if split == "val" else torch.linspace(0, 360.0, n)
        elevation_deg = torch.full_like(azimuth_deg, cfg.eval_elevation_deg)
        camera_distances = torch.full_like(azimuth_deg, cfg.eval_camera_distance)
        elevation, azimuth = elevation_deg * math.pi / 180, azimuth_deg * math.pi / 180
        camera_positions = torch.stack([camera_distances * torch.cos(elevation) * torch.cos(azimuth),
                                        camera_distances * torch.cos(elevation) * torch.sin(azimuth),
                                        camera_distances * torch.sin(elevation)], dim=-1)
        up = torch.as_tensor([0, 0, 1], dtype=torch.float32)[None, :].repeat(n, 1)
        self.c2w = look_at(camera_positions, torch.zeros_like(camera_positions), up)
        self.fovy = torch.full_like(azimuth_deg, cfg.eval_fovy_deg) * math.pi / 180
        self.proj_mtx = get_projection_matrix(self.fovy, cfg.eval_width / cfg.eval_height, 0.01, 100.0)
        self.mvp_mtx = get_mvp_matrix(self.c2w, self.proj_mtx)
        self.camera_positions = self.light_positions = camera_positions
        self.elevation_deg, self.azimuth_deg, self.camera_distances = elevation_deg, azimuth_deg, camera_distances

    def __len__(self) -> int:
        return self.n_views

    def __getitem__(self, index: int) -> Dict[str, Any]:
        return {"index": index, "mvp_mtx": self.mvp_mtx[index], "c2w": self.c2w[index],
                "camera_positions": self.camera_positions[index], "light_positions": self.light_positions[index],
                "elevation": self.elevation_deg[index], "azimuth": self.azimuth_deg[index],
                "camera_distances": self.camera_distances[index], "height": self.cfg.eval_height,
                "width": self.cfg.eval_width, "fovy": self.fovy[index], "proj_mtx": self.proj_mtx[index]}

    def collate(self, items: List[Dict[str, Any]]) -> Dict[str, Any]:
        out = {k: (torch.stack([it[k] for it in items]) if torch.is_tensor(items[0][k])
                   else torch.as_tensor([it[k] for it in items])) for k in items[0] if k not in ("height", "width")}
        out.update(height=self.cfg.eval_height, width=self.cfg.eval_width, _fovy_rad=out["fovy"])
        return out

    def to_device(self, batch: Dict[str, Any], device) -> Dict[str, Any]:
        return RandomCameraIterableDataset.to_device(self, batch, device)


class _CameraDataModule:
    dataset_cls = RandomCameraIterableDataset
    eval_dataset_cls = RandomCameraDataset

    def __init__(self, cfg=None) -> None:
        self.cfg = parse_structured(self.dataset_cls.config_cls, cfg)
        self.train_dataset = self.val_dataset = self.test_dataset = None

    def setup(self, stage=None) -> None:
        if stage in (None, "fit"):
            self.train_dataset = self.dataset_cls(self.cfg)
        if stage in (None, "fit", "validate"):
            self.val_dataset = self.eval_dataset_cls(self.cfg, "val")
        if stage in (None, "test", "predict"):
            self.test_dataset = self.eval_dataset_cls(self.cfg, "test")

    def _eval_loader(self, ds):
        bs = int(self.cfg.eval_batch_size)
        for i in range(0, len(ds), bs):
            yield ds.collate([ds[j] for j in range(i, min(i + bs, len(ds)))])

    def val_dataloader(self):
        if self.val_dataset is None:
            self.setup("validate")
        return self._eval_loader(self.val_dataset)

    def test_dataloader(self):
        if self.test_dataset is None:
            self.setup("test")
        return self._eval_loader(self.test_dataset)

    def train_dataloader(self):
        """Generator of host batches (num_workers=0, batch_size=None in the reference: uncond.py:489-502)."""
        if self.train_dataset is None:
            self.setup("fit")
        ds = self.train_dataset
        while True:
            yield ds.collate({})


@register("random-camera-datamodule")
class RandomCameraDataModule(_CameraDataModule):
    dataset_cls = RandomCameraIterableDataset


@register("mvdream-random-multiview-camera-datamodule")
class RandomMultiviewCameraDataModule(_CameraDataModule):
    dataset_cls = RandomMultiviewCameraIterableDataset
