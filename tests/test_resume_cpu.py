"""Checkpoint compatibility with the reference, on the CPU:
  * state-dict keys / shapes / dtypes of the C2 and C4 systems equal what the reference's own `configure` methods
    register (tests/golden/state_keys_golden.json, made by tests/golden/make_state_keys_golden.py);
  * the occupancy grid travels under nerfacc.OccGridEstimator's buffer names and is restored bit for bit;
  * the Trainer writes Lightning's checkpoint layout where the reference's `checkpoint:` section says
    (launch.py:201-206 there) and `resume=` continues a run exactly (parameters, optimizer moments, step counters).
Modules are only constructed / (de)serialised here; nothing is rendered without the GPU."""
import json
import os

import pytest
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = json.load(open(os.path.join(HERE, "golden", "state_keys_golden.json")))


@pytest.fixture
def cpu_system(monkeypatch, tmp_path):
    import scaledreamer_b200 as sd
    from scaledreamer_b200 import core

    monkeypatch.setattr(core, "get_device", lambda: torch.device("cpu"))
    monkeypatch.chdir(tmp_path)
    os.makedirs(tmp_path / "load")
    json.dump({"train": ["a", "b"], "val": ["a"], "test": ["a"]}, open(tmp_path / "load" / "lib.json", "w"))

    def make(which):
        name, cli = {"C2": ("asd_sd_nerf.yaml", ["system.prompt_processor.prompt=a hamburger"]),
                     "C4": ("asd_sd_hyper_iNGP.yaml", ["system.prompt_processor.prompt_library=lib"])}[which]
        cfg = sd.load_config(os.path.join(HERE, "configs", name), cli_args=cli)
        return sd.find(cfg.system_type)(cfg.system)

    return make


@pytest.mark.parametrize("which", ["C2", "C4"])
def test_state_dict_matches_the_reference_layout(cpu_system, which):
    system = cpu_system(which)
    ours = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in system.state_dict().items()}
    assert ours == KEYS[which], (sorted(set(ours) ^ set(KEYS[which])))


def test_occupancy_grid_travels_in_the_state_dict(cpu_system):
    system = cpu_system("C2")
    g = torch.Generator().manual_seed(3)
    occs = torch.rand(32 ** 3, generator=g)
    binaries = (occs > 0.6).reshape(1, 32, 32, 32)
    sd = system.state_dict()
    assert not sd["renderer.estimator.binaries"].any() and sd["renderer.estimator.occs"].abs().sum() == 0
    assert sd["renderer.estimator.aabbs"].tolist() == [[-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]]
    sd["renderer.estimator.occs"], sd["renderer.estimator.binaries"] = occs, binaries
    assert system.load_state_dict(sd, strict=True).missing_keys == []
    occ = system.renderer.occ
    assert torch.equal(occ.occs, occs) and torch.equal(occ.binaries().reshape(1, 32, 32, 32), binaries)
    assert abs(float(occ.mean) - float(occs.mean())) < 1e-7  # the alpha threshold min(0.01, mean) survives the reload
    # cell (x, y, z) sits at bit (x*32 + y)*32 + z of the packed field the march kernel reads (field.cuh)
    x, y, z = 5, 17, 30
    i = (x * 32 + y) * 32 + z
    assert bool((occ.bits[i // 32] >> (i % 32)) & 1) == bool(binaries[0, x, y, z])
    again = system.state_dict()
    assert torch.equal(again["renderer.estimator.binaries"], binaries) and torch.equal(again["renderer.estimator.occs"], occs)
    # a grid of another resolution is refused, not silently mis-indexed
    sd["renderer.estimator.occs"] = torch.zeros(16 ** 3)
    sd["renderer.estimator.binaries"] = torch.zeros(1, 16, 16, 16, dtype=torch.bool)
    with pytest.raises(RuntimeError, match="occupancy grid"):
        system.load_state_dict(sd, strict=True)


def test_flat_table_key_of_earlier_builds_still_loads(cpu_system):
    system = cpu_system("C2")
    sd = system.state_dict()
    table = torch.randn_like(sd["geometry.encoding.encoding.encoding.params"])
    old = {k.replace("encoding.encoding.encoding.params", "encoding.encoding.params"): v for k, v in sd.items()}
    old["geometry.encoding.encoding.params"] = table
    system.load_state_dict(old, strict=True)
    assert torch.equal(system.geometry.encoding.table.detach(), table)
    assert system.geometry.encoding.encoding.params is system.geometry.encoding.table  # shortcut used by the kernels


# ------------------------------------------------------------------------------------------------ trainer
def _toy(tmp_path, max_steps, **trainer_kw):
    """A system with the BaseSystem contract whose step is plain torch (the fit loop is what is under test)."""
    from scaledreamer_b200 import core
    from scaledreamer_b200.systems import BaseSystem, Trainer

    class Toy(BaseSystem):
        def configure(self):
            torch.manual_seed(0)
            self.geometry = nn.Linear(4, 3)
            self.seen_steps = []

        def update_step(self, epoch, global_step, on_load_weights=False):
            self.seen_steps.append((global_step, on_load_weights))

        def training_step(self, batch, batch_idx):
            loss = (self.geometry(batch["x"]) ** 2).mean() * (1.0 + 0.1 * self.true_global_step)
            self.log("train/loss", loss)
            return {"loss": loss}

    class Data:
        def setup(self, stage):
            pass

        train_dataset = property(lambda self: self)

        def train_dataloader(self):
            g = torch.Generator().manual_seed(5)
            while True:
                yield {"x": torch.randn(8, 4, generator=g)}

        def update_step(self, epoch, step):
            pass

        def to_device(self, batch, device):
            return batch

    system = Toy({"optimizer": {"name": "Adam", "args": {"lr": 0.05}, "params": {"geometry": {"lr": 0.05}}}})
    system.configure_optimizers = lambda: torch.optim.Adam(system.geometry.parameters(), lr=0.05)
    tr = Trainer(max_steps=max_steps, log_every_n_steps=1, distributed=False, ckpt_dir=str(tmp_path / "ckpts"),
                 **trainer_kw)
    return core, system, Data(), tr


def test_trainer_writes_lightning_layout_and_resumes_exactly(tmp_path, monkeypatch):
    from scaledreamer_b200 import core

    monkeypatch.setattr(core, "get_device", lambda: torch.device("cpu"))
    ck = {"save_last": True, "save_top_k": -1, "every_n_train_steps": 3}
    # one run of six steps ...
    _, full, data, tr = _toy(tmp_path / "a", 6, checkpoint=ck)
    tr.fit(full, data)
    assert sorted(os.listdir(tmp_path / "a" / "ckpts")) == ["epoch=0-step=3.ckpt", "epoch=0-step=6.ckpt", "last.ckpt"]
    last = torch.load(tmp_path / "a" / "ckpts" / "last.ckpt", weights_only=False)
    assert {"epoch", "global_step", "pytorch-lightning_version", "state_dict", "optimizer_states", "lr_schedulers"} <= set(last)
    assert last["global_step"] == 6 and set(last["state_dict"]) == {"geometry.weight", "geometry.bias"}
    assert last["optimizer_states"][0]["state"][0]["exp_avg"].shape == (3, 4)
    # ... equals three steps, a checkpoint, and three more steps from it (the data stream restarts in both halves, so
    # the second half is fed the batches 4..6 of the full run by skipping three)
    _, first, data1, tr1 = _toy(tmp_path / "b", 3, checkpoint=ck)
    tr1.fit(first, data1)
    _, second, data2, tr2 = _toy(tmp_path / "c", 6, checkpoint=ck)
    with torch.no_grad():
        second.geometry.weight.zero_()  # whatever it held is replaced by the checkpoint
    loader = data2.train_dataloader()
    [next(loader) for _ in range(3)]
    data2.train_dataloader = lambda: loader
    tr2.load_checkpoint(str(tmp_path / "b" / "ckpts" / "last.ckpt"), second)
    assert tr2.global_step == 3 and second.true_global_step == 3 and second.seen_steps[-1] == (3, True)
    tr2.fit(second, data2)
    assert tr2.global_step == 6
    for k, v in full.state_dict().items():
        torch.testing.assert_close(second.state_dict()[k], v, rtol=0, atol=0)
    assert [r["step"] for r in tr2.history] == [4, 5, 6]
    assert tr2.history[-1]["train/loss"] == tr.history[-1]["train/loss"]


def test_resume_refuses_a_checkpoint_of_another_model(tmp_path, monkeypatch):
    from scaledreamer_b200 import core

    monkeypatch.setattr(core, "get_device", lambda: torch.device("cpu"))
    _, system, _, tr = _toy(tmp_path, 1)
    torch.save({"state_dict": {"geometry.weight": torch.zeros(3, 4)}, "global_step": 2}, tmp_path / "bad.ckpt")
    with pytest.raises(RuntimeError, match="missing"):
        tr.load_checkpoint(str(tmp_path / "bad.ckpt"), system)
    # entries of modules the system does not own (guidance, prompt processor) and empty tensors are ignored
    sd = dict(system.state_dict())
    sd.update({"guidance.unet.w": torch.zeros(2), "geometry.encoding.params": torch.zeros(0)})
    torch.save({"state_dict": sd, "global_step": 2, "epoch": 0}, tmp_path / "ok.ckpt")
    tr.load_checkpoint(str(tmp_path / "ok.ckpt"), system)
    assert tr.global_step == 2


def test_lr_scheduler_follows_torch_and_survives_a_resume(tmp_path, monkeypatch):
    """`system.scheduler` (threestudio/systems/utils.py:74-104, systems/base.py:101-112): nested SequentialLR with
    interval "step" drives the optimizer's learning rate exactly like the torch schedulers built by hand, its state is
    written to `lr_schedulers` and a resumed run continues the curve; an "epoch" scheduler never steps on the endless
    camera stream."""
    from torch.optim import lr_scheduler

    from scaledreamer_b200 import core
    from scaledreamer_b200.systems import parse_scheduler

    monkeypatch.setattr(core, "get_device", lambda: torch.device("cpu"))
    sched_cfg = {"name": "SequentialLR", "interval": "step", "milestones": [3],
                 "schedulers": [{"name": "LinearLR", "interval": "step", "args": {"start_factor": 0.1, "total_iters": 3}},
                                {"name": "ExponentialLR", "interval": "step", "args": {"gamma": 0.8}}]}

    def run(path, steps, resume=None):
        _, system, data, tr = _toy(path, steps, checkpoint={"save_last": True})
        system.cfg.scheduler = sched_cfg
        lrs = []
        step = system.training_step

        def spy(batch, i):
            lrs.append(tr._opt.param_groups[0]["lr"])
            return step(batch, i)

        system.training_step = spy
        make_opt = system.configure_optimizers
        system.configure_optimizers = lambda: setattr(tr, "_opt", make_opt()) or tr._opt
        if resume:
            tr.load_checkpoint(resume, system)
        tr.fit(system, data)
        return lrs, tr

    lrs, tr = run(tmp_path / "a", 7)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=0.05)
    ref = lr_scheduler.SequentialLR(opt, [lr_scheduler.LinearLR(opt, start_factor=0.1, total_iters=3),
                                          lr_scheduler.ExponentialLR(opt, gamma=0.8)], milestones=[3])
    expect = []
    for _ in range(7):
        expect.append(opt.param_groups[0]["lr"])
        opt.step()
        ref.step()
    assert lrs == pytest.approx(expect, rel=1e-12) and len(set(lrs)) == 7
    ck = torch.load(tmp_path / "a" / "ckpts" / "last.ckpt", weights_only=False)
    assert len(ck["lr_schedulers"]) == 1 and ck["lr_schedulers"][0]["last_epoch"] == 7
    first, _ = run(tmp_path / "b", 4)
    second, tr2 = run(tmp_path / "c", 7, resume=str(tmp_path / "b" / "ckpts" / "last.ckpt"))
    assert first + second == pytest.approx(expect, rel=1e-12) and tr2.global_step == 7
    # "epoch" interval: constructed, never stepped
    p2 = torch.optim.SGD([p], lr=1.0)
    assert parse_scheduler({"name": "ExponentialLR", "args": {"gamma": 0.5}}, p2)["interval"] == "epoch"
    with pytest.raises(NotImplementedError):
        parse_scheduler({"name": "NoSuchLR"}, p2)


def test_val_check_interval_and_gradient_accumulation(tmp_path, monkeypatch):
    """trainer.val_check_interval counts training BATCHES (Lightning), trainer.accumulate_grad_batches optimizer-steps every
    k-th batch (C5: 2): with k = 2 and an interval of 4 the validation views arrive at optimizer steps 2 and 4, the module
    is back in training mode afterwards, and max_steps counts optimizer steps."""
    from scaledreamer_b200 import core
    from scaledreamer_b200.systems import Trainer

    monkeypatch.setattr(core, "get_device", lambda: torch.device("cpu"))
    _, system, data, _ = _toy(tmp_path, 5)
    batches, seen = [], []
    step = system.training_step
    system.training_step = lambda b, i: (batches.append((i, system.true_global_step, system.training)), step(b, i))[1]
    system.validation_step = lambda b, i: {"index": b["index"], "mode": system.training}
    type(data).val_dataset = property(lambda self: self)
    type(data).val_dataloader = lambda self: iter([{"index": torch.tensor([0])}, {"index": torch.tensor([1])}])
    tr = Trainer(max_steps=5, log_every_n_steps=1, distributed=False, accumulate_grad_batches=2, val_check_interval=4,
                 on_validation=lambda outs, s: seen.append((s, [int(o["index"][0]) for o in outs], [o["mode"] for o in outs])))
    tr.fit(system, data)
    assert tr.global_step == 5 and len(batches) == 10
    assert [b[1] for b in batches] == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4] and all(b[2] for b in batches)
    assert seen == [(2, [0, 1], [False, False]), (4, [0, 1], [False, False])] and system.training
    assert [r["step"] for r in tr.history] == [1, 2, 3, 4, 5]
    # without a consumer for the views the loop is skipped
    assert Trainer(max_steps=1, val_check_interval=4, distributed=False).val_check_interval == 0


def test_accumulated_gradients_reach_a_plain_optimizer_as_their_mean(tmp_path, monkeypatch):
    """accumulate_grad_batches = k with an optimizer that is not one of the fused kernels: k micro-batch gradients are
    summed by autograd and divided by k before the step (Lightning divides the loss by k), so one SGD step over two
    batches equals a step on the mean gradient."""
    from scaledreamer_b200 import core
    from scaledreamer_b200.systems import Trainer

    monkeypatch.setattr(core, "get_device", lambda: torch.device("cpu"))
    _, system, data, _ = _toy(tmp_path, 1)
    system.configure_optimizers = lambda: torch.optim.SGD(system.geometry.parameters(), lr=0.1)
    w0 = system.geometry.weight.detach().clone()
    loader = data.train_dataloader()
    grads = []
    for _ in range(2):
        x = next(loader)["x"]
        w = w0.clone().requires_grad_(True)
        ((x @ w.T + system.geometry.bias.detach()) ** 2).mean().backward()
        grads.append(w.grad)
    tr = Trainer(max_steps=1, log_every_n_steps=1, distributed=False, accumulate_grad_batches=2)
    tr.fit(system, data)
    torch.testing.assert_close(system.geometry.weight.detach(), w0 - 0.1 * (grads[0] + grads[1]) / 2, rtol=1e-6, atol=1e-7)


def test_system_weights_and_ignore_modules(cpu_system, tmp_path):
    """`system.weights=<ckpt>` with `system.weights_ignore_modules` (systems/base.py:54-60, utils/misc.py:33-63): the named
    modules keep their fresh initialisation, everything else (occupancy grid included) comes from the checkpoint, and the
    step-dependent state is replayed at the checkpoint's step with on_load_weights."""
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    donor = cpu_system("C2")
    with torch.no_grad():
        for p in donor.parameters():
            p.add_(1.0)
    donor_sd = donor.state_dict()
    donor_sd["renderer.estimator.binaries"] = torch.ones(1, 32, 32, 32, dtype=torch.bool)
    donor_sd["renderer.estimator.occs"] = torch.full((32 ** 3,), 0.5)
    donor.load_state_dict(donor_sd)
    tr = Trainer(max_steps=0, distributed=False)
    tr.global_step = 1234
    donor.true_global_step = 1234
    path = str(tmp_path / "donor.ckpt")
    tr.save_checkpoint(path, donor)
    cfg = sd.load_config(os.path.join(HERE, "configs", "asd_sd_nerf.yaml"),
                         cli_args=["system.prompt_processor.prompt=a hamburger", f"system.weights={path}",
                                   "system.weights_ignore_modules=[background]"])
    system = sd.find(cfg.system_type)(cfg.system)
    mine, ref = system.state_dict(), donor.state_dict()
    for k in ref:
        same = torch.equal(mine[k], ref[k])
        assert same != k.startswith("background."), k
    assert bool(system.renderer.occ.binaries().all())


def test_find_last_path_picks_the_newest_trial(tmp_path):
    """threestudio/utils/misc.py:143-161, used for `system.weights` (systems/base.py:250-251)."""
    from scaledreamer_b200.core import find_last_path

    for stamp in ("20240101-000000", "20240301-120000", "20240201-000000"):
        os.makedirs(tmp_path / "exp" / f"a_corgi@{stamp}" / "ckpts")
        open(tmp_path / "exp" / f"a_corgi@{stamp}" / "ckpts" / "last.ckpt", "w").close()
    os.makedirs(tmp_path / "exp" / "zebra@20250101-000000")
    got = find_last_path(str(tmp_path / "exp" / "a corgi@LAST" / "ckpts" / "last.ckpt"))
    assert got == str(tmp_path / "exp" / "a_corgi@20240301-120000" / "ckpts" / "last.ckpt")
    assert find_last_path(None) is None and find_last_path("x/y.ckpt") == "x/y.ckpt"
    with pytest.raises(FileNotFoundError):
        find_last_path(str(tmp_path / "exp" / "a_corgi@LAST" / "ckpts" / "missing.ckpt"))
