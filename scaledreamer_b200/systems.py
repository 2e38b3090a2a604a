"""Training system + fit loop for the ASD step (threestudio's "scaledreamer-system" contract without Lightning).

  StableDreamer            threestudio/systems/scaledreamer.py:14-170 (training_step :48-170)
  BaseLift3DSystem         threestudio/systems/base.py:205-303 (plugin instantiation by registry name)
  parse_optimizer          threestudio/systems/utils.py:25-53 (per-module parameter groups by dotted path)
  fit loop hook order      on_train_batch_start -> training_step -> backward -> optimizer step -> on_train_batch_end
                           (pytorch-lightning 2.0.0 as used by launch.py:233-249; systems/base.py:120-202)
Data parallelism (launch.py:233-240 -> Lightning DDP): one process per GPU, generator gradients packed in one flat
buffer and averaged with a single NCCL all-reduce per optimizer step.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import core, lib as L
from .core import C, Updateable, find, parse_structured, register


# ------------------------------------------------------------------------------------------------ optimizer
def getattr_recursive(m, attr: str):
    for name in attr.split("."):
        m = getattr(m, name)
    return m


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW / Adam semantics (decoupled weight decay; Adam == weight_decay 0 here), one
    sdb_adamw_step launch per parameter tensor. `grad_scale` multiplies every gradient (1/world_size after the
    all-reduce sum)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **unused):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        lib = L.load()
        st = L.stream_ptr()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                L.check(lib.sdb_adamw_step(L.ptr(p.data), L.ptr(g), L.ptr(state["exp_avg"]), L.ptr(state["exp_avg_sq"]),
                                           p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                           float(group["weight_decay"]), int(state["step"]), float(self.grad_scale), st),
                        "sdb_adamw_step")
        return None


class FusedAdan(torch.optim.Optimizer):
    """Adan (threestudio/systems/optimizers.py:23-315; the Triplane-Transformer configs), one sdb_adan_step launch per
    parameter tensor. max_grad_norm > 0 clips by the global gradient norm as the reference does (optimizers.py:108-128: one
    device reduction and one host read per step); the coefficient is folded into the kernel's gradient scale, so the
    stored previous gradient is the clipped one, like the reference's in-place `grad.mul_(clip)`."""

    def __init__(self, params, lr=1e-3, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0,
                 no_prox=False, foreach=True, **unused):
        if not 0.0 <= max_grad_norm:
            raise ValueError("Invalid Max grad norm: {}".format(max_grad_norm))
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, no_prox=no_prox,
                                      max_grad_norm=max_grad_norm))
        self.grad_scale = 1.0

    @torch.no_grad()
    def clip_coefficient(self) -> float:
        """clamp(max_grad_norm / (||g|| + eps), max=1) over every gradient of every group, g being the gradient the
        update uses (`grad_scale` included: the reference sees DDP-averaged gradients). 1.0 when clipping is off."""
        max_norm = self.defaults["max_grad_norm"]
        if max_norm <= 0:
            return 1.0
        grads = [p.grad for group in self.param_groups for p in group["params"] if p.grad is not None]
        if not grads:
            return 1.0
        total = torch.zeros(1, device=grads[0].device)
        for g in grads:
            total.add_(g.float().pow(2).sum())
        norm = torch.sqrt(total) * self.grad_scale
        return float(torch.clamp(max_norm / (norm + self.param_groups[-1]["eps"]), max=1.0).item())

    @torch.no_grad()
    def step(self, closure=None):
        lib = L.load()
        st = L.stream_ptr()
        scale = self.grad_scale * self.clip_coefficient()
        for group in self.param_groups:
            b1, b2, b3 = group["betas"]
            group["step"] = group.get("step", 0) + 1
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if not state:
                    for k in ("exp_avg", "exp_avg_sq", "exp_avg_diff"):
                        state[k] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    # optimizers.py:277-281: a parameter's first update sees neg_pre_grad = -grad, i.e. a zero gradient
                    # difference, whatever the group's step count is by then
                    state["prev_grad"] = (g * scale).to(torch.float32).clone(memory_format=torch.preserve_format)
                L.check(lib.sdb_adan_step(L.ptr(p.data), L.ptr(g), L.ptr(state["exp_avg"]), L.ptr(state["exp_avg_sq"]),
                                          L.ptr(state["exp_avg_diff"]), L.ptr(state["prev_grad"]), p.numel(),
                                          float(group["lr"]), float(b1), float(b2), float(b3), float(group["eps"]),
                                          float(group["weight_decay"]), int(group["step"]), float(scale),
                                          int(bool(group["no_prox"])), st), "sdb_adan_step")
        return None


    # The reference keeps MINUS the previous gradient under `neg_pre_grad` (optimizers.py:277-281, 307); checkpoints use
    # its name and sign so that either side resumes the other's optimizer state.
    def state_dict(self):
        sd = super().state_dict()
        sd["state"] = {k: {("neg_pre_grad" if n == "prev_grad" else n): (-v if n == "prev_grad" else v)
                           for n, v in st.items()} for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        sd["state"] = {k: {("prev_grad" if n == "neg_pre_grad" else n): (-v if n == "neg_pre_grad" else v)
                           for n, v in st.items()} for k, st in state_dict["state"].items()}
        super().load_state_dict(sd)


def parse_optimizer(config: dict, model: nn.Module) -> torch.optim.Optimizer:
    name = config["name"]
    args = dict(config.get("args", {}))
    if "params" in config:
        params = []
        for pname, pargs in config["params"].items():
            mod = getattr_recursive(model, pname)
            plist = list(mod.parameters()) if isinstance(mod, nn.Module) else ([mod] if isinstance(mod, nn.Parameter) else [])
            params.append({"params": plist, "name": pname, **pargs})
    else:
        params = list(model.parameters())
    if name in ("AdamW", "Adam"):
        if name == "Adam":
            args.setdefault("weight_decay", 0.0)
        return FusedAdamW(params, **args)
    if name == "Adan":
        return FusedAdan(params, **args)
    if hasattr(torch.optim, name):
        return getattr(torch.optim, name)(params, **args)
    raise NotImplementedError(f"optimizer {name}")


def parse_scheduler(config: dict, optimizer: torch.optim.Optimizer) -> Dict[str, Any]:
    """threestudio/systems/utils.py:74-104: `{name, args, interval}` -> torch.optim.lr_scheduler instance; `SequentialLR`
    (with `schedulers`, `milestones`) and `ChainedScheduler` (with `schedulers`) nest. interval is "epoch" (default) or
    "step"."""
    from torch.optim import lr_scheduler

    interval = config.get("interval", "epoch")
    assert interval in ["epoch", "step"]
    name = config["name"]
    if name == "SequentialLR":
        sched = lr_scheduler.SequentialLR(optimizer, [parse_scheduler(c, optimizer)["scheduler"] for c in config["schedulers"]],
                                          milestones=list(config["milestones"]))
    elif name == "ChainedScheduler":
        sched = lr_scheduler.ChainedScheduler([parse_scheduler(c, optimizer)["scheduler"] for c in config["schedulers"]])
    elif hasattr(lr_scheduler, name):
        sched = getattr(lr_scheduler, name)(optimizer, **dict(config.get("args", {})))
    else:
        raise NotImplementedError(f"scheduler {name}")
    return {"scheduler": sched, "interval": interval}


# ------------------------------------------------------------------------------------------------ system
def binary_cross_entropy(input, target):
    """threestudio/utils/ops.py:365-369"""
    return -(target * torch.log(input) + (1 - target) * torch.log(1 - input)).mean()


class BaseSystem(nn.Module, Updateable):
    @dataclass
    class Config:
        loggers: dict = field(default_factory=dict)
        loss: dict = field(default_factory=dict)
        optimizer: dict = field(default_factory=dict)
        scheduler: Optional[dict] = None
        weights: Optional[str] = None
        weights_ignore_modules: Optional[List[str]] = None
        cleanup_after_validation_step: bool = False
        cleanup_after_test_step: bool = False

    cfg: Config

    def __init__(self, cfg, resumed=False) -> None:
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self._resumed = resumed
        self.true_global_step = 0
        self.true_current_epoch = 0
        self.logged: Dict[str, Any] = {}
        self.configure()
        self.cfg.weights = core.find_last_path(self.cfg.weights)  # systems/base.py:250-251
        if self.cfg.weights is not None:
            sd, epoch, step = core.load_module_weights(self.cfg.weights, ignore_modules=self.cfg.weights_ignore_modules)
            self.load_state_dict(sd, strict=False)
            self.do_update_step(epoch, step, on_load_weights=True)

    def configure(self) -> None:
        pass

    def C(self, value: Any) -> float:
        return C(value, self.true_current_epoch, self.true_global_step)

    def log(self, name: str, value, **kwargs) -> None:
        self.logged[name] = value  # device scalars stay on device; the trainer reads them back lazily

    def configure_optimizers(self):
        return parse_optimizer(self.cfg.optimizer, self)

    def configure_scheduler(self, optimizer) -> Optional[Dict[str, Any]]:
        """The `lr_scheduler` half of the reference's configure_optimizers (systems/base.py:101-112)."""
        return parse_scheduler(self.cfg.scheduler, optimizer) if self.cfg.scheduler is not None else None

    def on_fit_start(self) -> None:
        pass


@register("scaledreamer-system")
class StableDreamer(BaseSystem):
    @dataclass
    class Config(BaseSystem.Config):
        geometry_type: str = ""
        geometry: dict = field(default_factory=dict)
        geometry_convert_from: Optional[str] = None
        geometry_convert_inherit_texture: bool = False
        geometry_convert_override: dict = field(default_factory=dict)
        material_type: str = ""
        material: dict = field(default_factory=dict)
        background_type: str = ""
        background: dict = field(default_factory=dict)
        renderer_type: str = ""
        renderer: dict = field(default_factory=dict)
        guidance_type: str = ""
        guidance: dict = field(default_factory=dict)
        prompt_processor_type: str = ""
        prompt_processor: dict = field(default_factory=dict)
        exporter_type: str = "mesh-exporter"
        exporter: dict = field(default_factory=dict)
        stage: str = "coarse"
        visualize_samples: bool = False
        validation_via_video: bool = False

    cfg: Config

    def configure(self) -> None:
        if self.cfg.geometry_convert_from:
            raise NotImplementedError("geometry_convert_from (coarse -> refine hand-over) is outside the ASD hot path")
        if self.cfg.stage != "coarse":
            raise NotImplementedError(f"stage '{self.cfg.stage}' is not implemented (coarse only)")
        dev = core.get_device()
        self.geometry = find(self.cfg.geometry_type)(self.cfg.geometry).to(dev)
        self.material = find(self.cfg.material_type)(self.cfg.material).to(dev)
        self.background = find(self.cfg.background_type)(self.cfg.background).to(dev)
        self.renderer = find(self.cfg.renderer_type)(self.cfg.renderer, geometry=self.geometry, material=self.material,
                                                     background=self.background).to(dev)

    def forward(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        return {**self.renderer(**batch)}

    def on_fit_start(self) -> None:
        self.prompt_processor = find(self.cfg.prompt_processor_type)(self.cfg.prompt_processor)
        self.guidance = find(self.cfg.guidance_type)(self.cfg.guidance)
        self.prompt_utils = self.prompt_processor()

    def training_step(self, batch, batch_idx):
        # the fused renderer evaluates the orientation term itself (no per-sample normal / weights / t_dirs tensors)
        if getattr(self, "renderer", None) is not None:
            self.renderer.orient_loss = self.C(self.cfg.loss.get("lambda_orient", 0.0)) > 0
        out = self(batch)
        guidance_out = self.guidance(out["comp_rgb"], self.prompt_utils, **batch, rgb_as_latents=False)
        loss = 0.0
        lam = self.cfg.loss
        for name, value in guidance_out.items():
            self.log(f"train/{name}", value)
            if name.startswith("loss_"):
                loss = loss + value * self.C(lam[name.replace("loss_", "lambda_")])
        if self.C(lam.get("lambda_orient", 0.0)) > 0:
            if "orient" in out:  # fused renderer: per-ray sums of w relu(n . d)^2 (csrc/render_orient.cu)
                loss_orient = out["orient"].sum() / (out["opacity"] > 0).sum()
            else:
                if "normal" not in out:
                    raise ValueError("Normal is required for orientation loss, no normal is found in the output.")
                cos = (out["normal"] * out["t_dirs"]).sum(-1, keepdim=True)  # scaledreamer.py:74-83
                loss_orient = (out["weights"].detach() * cos.clamp_min(0.0) ** 2).sum() / (out["opacity"] > 0).sum()
            self.log("train/loss_orient", loss_orient)
            loss = loss + loss_orient * self.C(lam["lambda_orient"])
        if self.C(lam.get("lambda_sparsity", 0.0)) > 0:
            loss_sparsity = (out["opacity"] ** 2 + 0.01).sqrt().mean()
            self.log("train/loss_sparsity", loss_sparsity)
            loss = loss + loss_sparsity * self.C(lam["lambda_sparsity"])
        if self.C(lam.get("lambda_opaque", 0.0)) > 0:
            op = out["opacity"].clamp(1.0e-3, 1.0 - 1.0e-3)
            loss_opaque = binary_cross_entropy(op, op)
            self.log("train/loss_opaque", loss_opaque)
            loss = loss + loss_opaque * self.C(lam["lambda_opaque"])
        if self.C(lam.get("lambda_z_variance", 0.0)) > 0:  # scaledreamer.py:93-102 (HiFA)
            if not out["z_variance"].requires_grad:
                raise NotImplementedError("lambda_z_variance > 0 needs the tape renderer (fused geometry, "
                                          "packed_capacity 0): this renderer's z_variance carries no gradient")
            loss_z_variance = out["z_variance"][out["opacity"] > 0.5].mean()
            self.log("train/loss_z_variance", loss_z_variance)
            loss = loss + loss_z_variance * self.C(lam["lambda_z_variance"])
        if self.C(lam.get("lambda_eikonal", 0.0)) > 0:
            raise ValueError("sdf is required for eikonal loss, no sdf is found in the output.")
        return {"loss": loss}


    def _eval_images(self, batch) -> Dict[str, Any]:
        """validation_step / test_step (scaledreamer.py:172-300) up to the image grid: the rendered view, the opacity
        map and the min-max normalised depth the reference writes to `it{step}-val/{index}.png`. Writing files / videos
        is the saver's job and stays outside this package; the tensors are returned instead."""
        out = self(batch)
        res = {"index": batch["index"], "comp_rgb": out["comp_rgb"], "opacity": out["opacity"]}
        if "comp_normal" in out:
            res["comp_normal"] = out["comp_normal"]
        if "depth" in out:
            d = out["depth"][0, :, :, 0]
            res["depth"] = (d - d.min()) / (d.max() - d.min())
        return res

    def validation_step(self, batch, batch_idx):
        if self.cfg.visualize_samples:
            raise NotImplementedError
        return self._eval_images(batch)

    def test_step(self, batch, batch_idx):
        return self._eval_images(batch)


# ------------------------------------------------------------------------------------------------ trainer
class Trainer:
    """Minimal fit loop with Lightning's hook order; optional data-parallel gradient averaging over NCCL."""

    def __init__(self, max_steps: int = 1, log_every_n_steps: int = 50, accumulate_grad_batches: int = 1,
                 distributed: Optional[bool] = None, ckpt_dir: Optional[str] = None, checkpoint: Optional[dict] = None,
                 val_check_interval: Optional[int] = None, on_validation=None, **unused) -> None:
        self.max_steps = int(max_steps)
        # Lightning's `val_check_interval: N` (int): the validation loop runs after every N training batches; the views
        # go to `on_validation(outs, global_step)` (launch.py writes them where the reference's savers do). Without a
        # consumer the loop is skipped: rendering 120 views nobody looks at is not part of a training step.
        self.val_check_interval = int(val_check_interval) if val_check_interval and on_validation else 0
        self.on_validation = on_validation
        self.log_every_n_steps = int(log_every_n_steps)
        self.accumulate = int(accumulate_grad_batches)
        # launch.py:201-206 of the reference: ModelCheckpoint(dirpath=<trial_dir>/ckpts, **cfg.checkpoint) with
        # `save_last`, `every_n_train_steps` (the yaml keys of configs/*/asd_*.yaml `checkpoint:`)
        self.ckpt_dir = ckpt_dir
        ck = dict(checkpoint or {})
        self.ckpt_every = int(ck.get("every_n_train_steps") or 0)
        self.ckpt_save_last = bool(ck.get("save_last", False))
        self._resume: Optional[Dict[str, Any]] = None
        self._last_saved_step = -1
        self._scheduler: Optional[Dict[str, Any]] = None
        import torch.distributed as dist

        self.dist = dist if (distributed if distributed is not None else dist.is_initialized()) else None
        self.world_size = self.dist.get_world_size() if self.dist else 1
        self.global_step = 0
        self.history: List[Dict[str, float]] = []
        self._flat = None
        self._ar_stream = None

    def _allreduce_grads(self, params: List[torch.Tensor]) -> None:
        """One collective per optimizer step on a flat fp32 buffer (replaces DDP's bucketed reducer). After the first
        step every `.grad` IS a view of that buffer (autograd accumulates into it in place, zero_grad(set_to_none=False)
        clears it in place), so no gather / scatter copies remain around the all-reduce."""
        with_grad = [p for p in params if p.grad is not None]
        if not with_grad:
            return
        n = sum(p.grad.numel() for p in with_grad)
        if self._flat is None or self._flat.numel() != n:
            self._flat = torch.empty(n, device=with_grad[0].grad.device, dtype=torch.float32)
        off = 0
        for p in with_grad:
            g, k = p.grad, p.grad.numel()
            view = self._flat[off:off + k].view_as(g)
            if g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
            off += k
        # The collective runs on a side stream: it starts when the backward's last kernel has finished and the optimizer
        # kernels wait for it, while the compute stream stays free for work that does not read the gradients (the
        # update hooks and the next batch's upload / ray generation are enqueued behind the optimizer only).
        if with_grad[0].is_cuda:
            if self._ar_stream is None:
                self._ar_stream = torch.cuda.Stream()
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self._ar_stream):
                self._ar_stream.wait_event(ready)
                self.dist.all_reduce(self._flat, op=self.dist.ReduceOp.SUM)
                done = torch.cuda.Event()
                done.record()
            torch.cuda.current_stream().wait_event(done)
        else:
            self.dist.all_reduce(self._flat, op=self.dist.ReduceOp.SUM)

    def sync_initial_state(self, system: nn.Module) -> None:
        """What DistributedDataParallel does when Lightning wraps the module (launch.py:233-240 of the reference): rank
        0's parameters and buffers (occupancy grid included, through the state dict) overwrite every other rank's, so
        replicas seeded with `seed + rank` start from one generator. A no-op without a process group."""
        if self.dist is None or self.world_size == 1:
            return
        with torch.no_grad():
            tensors = [t for t in list(system.parameters()) + list(system.buffers()) if t.numel() > 0]
            for t in tensors:
                self.dist.broadcast(t.data, src=0)
            occ = getattr(getattr(system, "renderer", None), "occ", None)
            if occ is not None:
                for t in (occ.occs, occ.bits, occ.mean):
                    self.dist.broadcast(t, src=0)

    # ---- checkpoints in the layout Lightning 2.0 writes (what `resume=` / `system.weights=` of the reference read)
    def checkpoint_dict(self, system: BaseSystem, optimizer: Optional[torch.optim.Optimizer] = None) -> Dict[str, Any]:
        ckpt = {"epoch": system.true_current_epoch, "global_step": self.global_step,
                "pytorch-lightning_version": "2.0.0", "state_dict": system.state_dict()}
        if optimizer is not None:
            ckpt["optimizer_states"] = [optimizer.state_dict()]
            ckpt["lr_schedulers"] = [self._scheduler["scheduler"].state_dict()] if self._scheduler else []
        return ckpt

    def save_checkpoint(self, path: str, system: BaseSystem, optimizer: Optional[torch.optim.Optimizer] = None) -> None:
        """Rank 0 writes (every rank holds the same parameters after the all-reduce); atomic rename so an interrupted
        write never leaves a truncated `last.ckpt`."""
        if core.get_rank() != 0:
            return
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".part"
        torch.save(self.checkpoint_dict(system, optimizer), tmp)
        os.replace(tmp, path)

    def load_checkpoint(self, path: str, system: BaseSystem) -> Dict[str, Any]:
        """`trainer.fit(ckpt_path=cfg.resume)` of the reference (launch.py:246-248): restores the module state, the step
        counters, re-runs the step-dependent schedules (`on_load_weights`) and keeps the optimizer state for `fit`."""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        sd = ckpt.get("state_dict", ckpt)
        missing, unexpected = system.load_state_dict(sd, strict=False)
        own = {k.split(".")[0] for k in system.state_dict()}
        # entries of modules that are not ours (guidance / prompt processor) and empty tensors (tcnn encodings without
        # parameters, e.g. SphericalHarmonics `params` of size 0) carry nothing to restore
        unexpected = [k for k in unexpected if k.split(".")[0] in own and sd[k].numel() > 0]
        if missing or unexpected:
            raise RuntimeError(f"checkpoint {path} does not match the system: missing {missing[:8]}, "
                               f"unexpected {unexpected[:8]}")
        self.global_step = int(ckpt.get("global_step", 0))
        system.true_global_step = self.global_step
        system.true_current_epoch = int(ckpt.get("epoch", 0))
        system.do_update_step(system.true_current_epoch, system.true_global_step, on_load_weights=True)
        self._resume = ckpt
        return ckpt

    def _maybe_checkpoint(self, system: BaseSystem, optimizer, last: bool = False) -> None:
        if not self.ckpt_dir:
            return
        periodic = bool(self.ckpt_every) and self.global_step % self.ckpt_every == 0
        if periodic and not last:
            name = f"epoch={system.true_current_epoch}-step={self.global_step}.ckpt"
            self.save_checkpoint(os.path.join(self.ckpt_dir, name), system, optimizer)
        if self.ckpt_save_last and (periodic or last) and self._last_saved_step != self.global_step:
            self.save_checkpoint(os.path.join(self.ckpt_dir, "last.ckpt"), system, optimizer)
            self._last_saved_step = self.global_step

    def _evaluate(self, system: BaseSystem, datamodule, split: str) -> List[Dict[str, Any]]:
        """Lightning's validate / test loop for the evaluation orbit: eval mode (no jitter, no random background, no
        tape), no autograd, one `*_step` per camera batch."""
        device = core.get_device()
        loader = datamodule.val_dataloader() if split == "val" else datamodule.test_dataloader()
        dataset = datamodule.val_dataset if split == "val" else datamodule.test_dataset
        was_training = system.training
        system.eval()
        hook = getattr(system, "on_validation_start" if split == "val" else "on_test_start", None)
        if hook is not None:
            hook()
        outs = []
        with torch.no_grad():
            for i, host_batch in enumerate(loader):
                # on_validation_batch_start / on_test_batch_start (systems/base.py:186-196): the step-dependent state
                # (finite-difference eps, progressive bands, ...) is refreshed before every evaluation batch; in eval
                # mode the renderers leave their occupancy grids alone
                if hasattr(dataset, "update_step"):
                    dataset.update_step(system.true_current_epoch, system.true_global_step)
                system.do_update_step(system.true_current_epoch, system.true_global_step)
                batch = dataset.to_device(host_batch, device)
                outs.append(system.validation_step(batch, i) if split == "val" else system.test_step(batch, i))
        system.train(was_training)
        return outs

    def validate(self, system: BaseSystem, datamodule) -> List[Dict[str, Any]]:
        return self._evaluate(system, datamodule, "val")

    def test(self, system: BaseSystem, datamodule) -> List[Dict[str, Any]]:
        return self._evaluate(system, datamodule, "test")

    def fit(self, system: BaseSystem, datamodule) -> None:
        device = core.get_device()
        datamodule.setup("fit")
        dataset = datamodule.train_dataset
        loader = datamodule.train_dataloader()
        system.train()
        system.on_fit_start()
        self.sync_initial_state(system)
        optimizer = system.configure_optimizers()
        # gradients arrive SUMMED over ranks and micro-batches; Lightning / DDP hand the optimizer their mean. The fused
        # optimizers take the factor as a kernel argument, any other optimizer gets its gradients scaled in place.
        mean_scale = 1.0 / (self.world_size * self.accumulate)
        fused = isinstance(optimizer, (FusedAdamW, FusedAdan))
        if fused:
            optimizer.grad_scale = mean_scale
        params = [p for g in optimizer.param_groups for p in g["params"]]
        # Lightning's order: schedulers are built with the optimizer (their constructors already rewrite the learning
        # rates), THEN the saved optimizer state (moments and the learning rates in force) and scheduler state are loaded
        configure_scheduler = getattr(system, "configure_scheduler", None)
        self._scheduler = configure_scheduler(optimizer) if configure_scheduler is not None else None
        if self._resume is not None and self._resume.get("optimizer_states"):
            optimizer.load_state_dict(self._resume["optimizer_states"][0])  # moments land on each parameter's device
        if self._scheduler and self._resume is not None and self._resume.get("lr_schedulers"):
            self._scheduler["scheduler"].load_state_dict(self._resume["lr_schedulers"][0])
        self._resume = None
        micro = 0
        while self.global_step < self.max_steps:
            batch = next(loader)
            # on_train_batch_start (systems/base.py:180-184)
            dataset.update_step(system.true_current_epoch, system.true_global_step)
            system.do_update_step(system.true_current_epoch, system.true_global_step)
            batch = dataset.to_device(batch, device)
            out = system.training_step(batch, micro)
            out["loss"].backward()
            micro += 1
            if micro % self.accumulate == 0:
                if self.dist is not None:
                    self._allreduce_grads(params)
                if not fused and mean_scale != 1.0:
                    torch._foreach_mul_([p.grad for p in params if p.grad is not None], mean_scale)
                optimizer.step()
                optimizer.zero_grad(set_to_none=False)
                # Lightning steps "step"-interval schedulers after every optimizer step; "epoch"-interval ones at epoch
                # end, which an endless camera stream (IterableDataset without a length) never reaches
                if self._scheduler and self._scheduler["interval"] == "step":
                    self._scheduler["scheduler"].step()
                self.global_step += 1
                system.true_global_step = self.global_step
                self._maybe_checkpoint(system, optimizer)
            system.do_update_step_end(system.true_current_epoch, system.true_global_step)
            if self.log_every_n_steps and self.global_step % self.log_every_n_steps == 0 and micro % self.accumulate == 0:
                rec = {k: (float(v.detach()) if torch.is_tensor(v) else v) for k, v in system.logged.items()}
                rec["step"] = self.global_step
                self.history.append(rec)
            if self.val_check_interval and micro % self.val_check_interval == 0:
                self.on_validation(self.validate(system, datamodule), self.global_step)
        self._maybe_checkpoint(system, optimizer, last=True)
