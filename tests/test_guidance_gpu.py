"""GPU parity of the whole guidance call (resize -> VAE -> q-sample -> UNet x5 -> Perp-Neg CFG -> loss and its
gradient w.r.t. the rendered image) against the CPU oracle, at a reduced image size so the oracle finishes in
seconds; plus an end-to-end training step through the plugin API with the reference-schema yaml.

Tolerances (relative L2 against the fp32 oracle): eps-pred 5e-3 (fp16 activations over ~100 layers; north_star's
1e-3 is stated for the reference's own fp16 path, which is not runnable here). The score gradient amplifies that
error: grad = w(t) (e_u + 7.5 (e_c - e_u + perp terms) - e_second) takes 7.5x DIFFERENCES of nearly equal
predictions, so its bound is 2e-2 (the reference's fp16 UNet has the same amplification); d(loss)/d(rgb)
additionally crosses the fp16 VAE backward: 3e-2.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "asd_sd_nerf.yaml")


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm())


def test_sd_asd_guidance_matches_oracle(cuda_device):
    import scaledreamer_b200 as sd
    from oracle import ldm_oracle as lo
    from scaledreamer_b200 import nets

    torch.manual_seed(0)
    pp = sd.find("stable-diffusion-prompt-processor")({"prompt": "a hamburger", "use_perp_neg": True,
                                                        "front_threshold": 30.0, "back_threshold": 30.0})
    g = sd.find("stable-diffusion-asynchronous-score-distillation-guidance")(
        {"guidance_scale": 7.5, "plus_ratio": 0.1, "plus_random": True, "guidance_perp_neg": -0.5,
         "min_step_percent": 0.02, "max_step_percent": 0.98})
    g.image_size = 128  # latents 16x16: keeps the CPU oracle fast; the kernels are size-generic
    B, S, hw = 1, 128, 256
    gen = torch.Generator().manual_seed(3)
    rgb = torch.rand(B, 32, 32, 3, generator=gen)
    rng = dict(noise=torch.randn(B, hw, 4, generator=gen), eps_post=torch.randn(B, hw, 4, generator=gen),
               t=torch.tensor([431]), u=torch.tensor([0.73]))
    elev, azim, dist = torch.tensor([12.0]), torch.tensor([-57.0]), torch.tensor([1.2])
    rgb_d = rgb.to(cuda_device).requires_grad_(True)
    out = g(rgb_d, pp(), elev, azim, dist, _rng={k: v.to(cuda_device) for k, v in rng.items()})
    out["loss_asd"].backward()
    torch.cuda.synchronize()

    # ---- oracle on the CPU with the same weights and draws
    sd_u = {k: v.half().float() for k, v in nets.random_state_dict(g.unet.specs, g.weights_seed).items()}
    sd_v = {k: v.half().float() for k, v in nets.random_state_dict(g.vae.specs, g.weights_seed + 1).items()}
    qw, qb = g.quant_w.cpu(), g.quant_b.cpu()
    x = rgb.clone().requires_grad_(True)
    img = F.interpolate(x.permute(0, 3, 1, 2), (S, S), mode="bilinear", align_corners=False) * 2 - 1
    h = lo.vae_encoder_forward(sd_v, img)
    nchw = lambda t_: t_.view(B, 16, 16, 4).permute(0, 3, 1, 2)
    z = lo.sample_latents(h, qw, qb, nchw(rng["eps_post"]))
    ac = lo.alphas_cumprod()
    t = rng["t"]
    tp = lo.t_plus(t, rng["u"], 0.1, g.min_step)
    qs = lambda tt: ac[tt].sqrt().view(-1, 1, 1, 1) * z + (1 - ac[tt]).sqrt().view(-1, 1, 1, 1) * nchw(rng["noise"])
    ctx, neg_w = pp().get_text_embeddings_perp_neg(elev, azim, dist, True)
    ctx = torch.cat([ctx, ctx[:B]], 0).float().cpu()
    neg_w = neg_w.cpu() * -1 * -0.5
    with torch.no_grad():
        x_in = torch.cat([qs(t)] * 4 + [qs(tp)], 0).half().float()
        eps = lo.unet_forward(sd_u, x_in, torch.cat([t] * 4 + [tp]).float(), ctx)
    grad, loss, gnorm = lo.asd_grad(eps, z, t, ac, B, 7.5, neg_w)
    loss.backward()

    e_eps = rel(g.buf["eps"].permute(0, 3, 1, 2), eps)
    e_grad = rel(g.buf["grad"].view(B, 16, 16, 4).permute(0, 3, 1, 2), grad)
    e_drgb = rel(rgb_d.grad, x.grad)
    print(f"guidance parity: eps {e_eps:.2e} grad {e_grad:.2e} loss {float(out['loss_asd'].detach()):.4f}/{float(loss.detach()):.4f} "
          f"d_rgb {e_drgb:.2e}")
    assert int(g.buf["t_plus"][0]) == int(tp[0])
    assert e_eps < 5e-3 and e_grad < 2e-2
    assert abs(float(out["loss_asd"]) - float(loss)) / float(loss) < 1e-2
    assert abs(float(out["grad_norm"]) - float(gnorm)) / float(gnorm) < 5e-3
    assert e_drgb < 3e-2


def test_training_step_end_to_end_with_reference_schema_yaml(cuda_device):
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    torch.manual_seed(1)
    cfg = sd.load_config(CFG, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger",
                                        "data.width=[64,64]", "data.height=[64,64]", "trainer.max_steps=3",
                                        "trainer.log_every_n_steps=1"])
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    before = {k: v.detach().clone() for k, v in system.state_dict().items() if v.numel() > 0}
    assert {"geometry.encoding.encoding.encoding.params", "geometry.density_network.layers.0.weight",
            "geometry.feature_network.layers.2.weight", "background.encoding.encoding.encoding.params",
            "background.network.layers.4.weight"} <= set(before)
    tr = Trainer(**cfg.trainer)
    tr.fit(system, dm)
    torch.cuda.synchronize()
    assert tr.global_step == 3 and len(tr.history) == 3
    last = tr.history[-1]
    assert all(k in last for k in ("train/loss_asd", "train/grad_norm", "train/min_step", "train/max_step"))
    assert last["train/loss_asd"] > 0 and last["train/loss_asd"] == last["train/loss_asd"]
    after = system.state_dict()
    assert (after["geometry.encoding.encoding.encoding.params"] != before["geometry.encoding.encoding.encoding.params"]).any()
    assert (after["geometry.feature_network.layers.2.weight"] != before["geometry.feature_network.layers.2.weight"]).any()


def test_mvdream_training_step_end_to_end(cuda_device):
    """C3: multi-view camera batches (4 views of one object) -> fused render -> MVDream guidance (multi-view UNet with
    camera conditioning, one shared timestep, plain CFG) -> backward -> AdamW, from the reference-schema yaml."""
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    torch.manual_seed(2)
    cfg_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "asd_mv_nerf.yaml")
    cfg = sd.load_config(cfg_path, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger",
                                             "data.width=[64,64]", "data.height=[64,64]", "trainer.max_steps=2",
                                             "trainer.log_every_n_steps=1"])
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    before = system.geometry.encoding.table.detach().clone()
    tr = Trainer(**cfg.trainer)
    tr.fit(system, dm)
    torch.cuda.synchronize()
    assert tr.global_step == 2
    last = tr.history[-1]
    assert last["train/loss_asd"] > 0 and last["train/loss_asd"] == last["train/loss_asd"]
    assert (system.geometry.encoding.table.detach() != before).any()
    g = system.guidance
    assert g.unet_cfg["num_frames"] == 4 and g.buf["unet_x"].shape[0] == 12  # (cond, uncond, t+dt) x 4 views
    assert int(g._last["t"].unique().numel()) == 1                            # one timestep for the whole view batch


def test_mvdream_yaml_survives_its_orientation_schedule(cuda_device):
    """configs/single-prompt_benchmark/asd_mv_nerf.yaml:100-102 switch lambda_orient / lambda_opaque from 0 to 100 at step
    10001. The fit loop is started at step 10000 and runs across that boundary on the FUSED renderer: loss_orient is
    logged, finite and positive, and training keeps moving the density network (round 1 raised NotImplementedError)."""
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    torch.manual_seed(3)
    cfg_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "asd_mv_nerf.yaml")
    cfg = sd.load_config(cfg_path, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger",
                                             "data.width=[64,64]", "data.height=[64,64]", "trainer.max_steps=10003",
                                             "trainer.log_every_n_steps=1"])
    assert cfg.system["loss"]["lambda_orient"] == [10000, 0.0, 100.0, 10001]
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    assert system.geometry.fusable
    with torch.no_grad():
        system.geometry.encoding.table.mul_(2000.0)  # 1e-4 init -> 0.2: the hash features shape the density
    tr = Trainer(**cfg.trainer)
    tr.global_step = system.true_global_step = 10000
    tr.fit(system, dm)
    torch.cuda.synchronize()
    assert tr.global_step == 10003
    by_step = {int(r["step"]): r for r in tr.history}
    assert "train/loss_orient" not in by_step[10001]          # the batch of step 10000 still ran with lambda 0
    for st in (10002, 10003):
        lo_ = by_step[st]["train/loss_orient"]
        assert lo_ == lo_ and lo_ > 0, (st, lo_)
        assert by_step[st]["train/loss_opaque"] == by_step[st]["train/loss_opaque"]
    assert system.geometry.density_network.layers[0].weight.grad is not None


def test_validation_and_test_loops(cuda_device):
    """Evaluation orbit through the system (scaledreamer.py:172-300): eval-mode renders are deterministic (no jitter, no
    random background, nothing taped), match a direct eval-mode renderer call on the same camera, and leave the module
    in training mode afterwards."""
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    torch.manual_seed(3)
    cfg = sd.load_config(CFG, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger",
                                        "data.eval_height=40", "data.eval_width=56", "data.n_val_views=3",
                                        "data.n_test_views=5"])
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    system.train()
    system.do_update_step(0, 0)  # warm-up occupancy refresh
    tr = Trainer(**cfg.trainer)
    v1, v2 = tr.validate(system, dm), tr.validate(system, dm)
    assert len(v1) == 3 and system.training
    for a, b in zip(v1, v2):
        assert a["comp_rgb"].shape == (1, 40, 56, 3) and a["opacity"].shape == (1, 40, 56, 1)
        assert a["depth"].shape == (40, 56) and float(a["depth"].min()) == 0.0 and float(a["depth"].max()) == 1.0
        assert torch.equal(a["comp_rgb"], b["comp_rgb"]) and torch.equal(a["opacity"], b["opacity"])
        assert not a["comp_rgb"].requires_grad
    assert [int(o["index"][0]) for o in v1] == [0, 1, 2]
    assert not torch.equal(v1[0]["comp_rgb"], v1[1]["comp_rgb"])  # different azimuths
    t = tr.test(system, dm)
    assert len(t) == 5
    # test view 0 and the closing view (azimuth 360) see the same image; val view 0 is the same camera
    assert (t[0]["comp_rgb"] - t[4]["comp_rgb"]).abs().max() < 2e-3
    assert torch.equal(t[0]["comp_rgb"], v1[0]["comp_rgb"])
    system.eval()
    ds = dm.val_dataset
    batch = ds.to_device(ds.collate([ds[1]]), cuda_device)
    with torch.no_grad():
        direct = system.renderer(**batch)
    assert torch.equal(direct["comp_rgb"], v1[1]["comp_rgb"])
    assert system.renderer.randomized is False


def test_launch_validate_writes_the_evaluation_views(cuda_device, tmp_path):
    """`launch.py --config ... --validate` (the reference's CLI contract): renders the evaluation orbit in eval mode and
    writes one rgb | opacity | depth PNG per view under <trial_dir>/save/val/."""
    import subprocess
    import sys

    from PIL import Image

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "launch.py"), "--config", CFG, "--validate",
                        "system.prompt_processor.prompt=a DSLR photo of a hamburger", "data.eval_height=32",
                        "data.eval_width=48", "data.n_val_views=2", f"exp_root_dir={tmp_path}"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    # the trial directory carries the reference's timestamp suffix (threestudio/utils/config.py:86-101)
    import glob

    trials = glob.glob(os.path.join(str(tmp_path), "asd_sd_nerf", "a_DSLR_photo_of_a_hamburger@*"))
    assert len(trials) == 1, trials
    out_dir = os.path.join(trials[0], "save", "val")
    files = sorted(os.listdir(out_dir))
    assert files == ["0.png", "1.png"], files
    im = Image.open(os.path.join(out_dir, "0.png"))
    assert im.size == (3 * 48, 32) and im.mode == "RGB"


def test_checkpoint_resume_on_the_device(cuda_device, tmp_path):
    """Two optimizer steps write <ckpt_dir>/last.ckpt in Lightning's layout with the reference's state-dict keys; a fresh
    system resumed from it holds the same hash table, MLPs, occupancy grid (bit for bit) and AdamW moments, renders the
    same evaluation view, and continues from step 2."""
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    cli = ["system.prompt_processor.prompt=a DSLR photo of a hamburger", "data.width=[64,64]", "data.height=[64,64]",
           "data.eval_height=32", "data.eval_width=32", "data.n_val_views=1", "trainer.max_steps=2",
           "trainer.log_every_n_steps=1"]
    torch.manual_seed(2)
    cfg = sd.load_config(CFG, cli_args=cli)
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    tr = Trainer(**cfg.trainer, ckpt_dir=str(tmp_path), checkpoint={"save_last": True, "every_n_train_steps": 2})
    tr.fit(system, dm)
    torch.cuda.synchronize()
    assert sorted(os.listdir(tmp_path)) == ["epoch=0-step=2.ckpt", "last.ckpt"]
    ck = torch.load(tmp_path / "last.ckpt", map_location="cpu", weights_only=False)
    assert ck["global_step"] == 2 and ck["state_dict"]["renderer.estimator.binaries"].shape == (1, 32, 32, 32)
    assert ck["state_dict"]["renderer.estimator.binaries"].any()  # the warm-up refresh marked the blob
    view = tr.validate(system, dm)[0]["comp_rgb"]

    torch.manual_seed(99)
    fresh = sd.find(cfg.system_type)(cfg.system)
    tr2 = Trainer(**{**cfg.trainer, "max_steps": 3}, ckpt_dir=str(tmp_path / "second"), checkpoint={})
    tr2.load_checkpoint(str(tmp_path / "last.ckpt"), fresh)
    assert tr2.global_step == 2 and fresh.true_global_step == 2
    a, b = system.state_dict(), fresh.state_dict()
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(fresh.renderer.occ.bits, system.renderer.occ.bits)
    assert torch.equal(tr2.validate(fresh, dm)[0]["comp_rgb"], view)
    tr2.fit(fresh, dm)
    torch.cuda.synchronize()
    assert tr2.global_step == 3 and tr2.history[-1]["step"] == 3
    assert not os.path.exists(tmp_path / "second")  # an empty `checkpoint:` section writes nothing
    table = "geometry.encoding.encoding.encoding.params"
    assert (fresh.state_dict()[table] != a[table]).any()
