"""Command-line entry with the reference's contract (launch.py:267-299 there):
    python launch.py --config configs/x.yaml --train [--gpu 0] key=value ...
--train runs the ASD hot path; --validate / --test render the evaluation orbit (scaledreamer.py:172-300) and write the
views as PNG files under <trial_dir>/save/ (rgb | opacity | depth side by side, like the reference's image grid);
--export (mesh extraction) is outside the scope of this repo. Checkpoints follow the reference's `checkpoint:` yaml
section (<trial_dir>/ckpts/last.ckpt, epoch=E-step=N.ckpt; Lightning's dict layout and the reference's state-dict keys);
`resume=<ckpt>` restores module state, occupancy grid, step counters and optimizer moments. During --train the evaluation
orbit is rendered every `trainer.val_check_interval` batches into <trial_dir>/save/it{step}-val/, and the test orbit once
after the last step into save/it{step}-test/ (the reference's --train ends with trainer.test).
"""
import argparse
import os
import sys


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--gpu", default="0")
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument("--train", action="store_true")
    g.add_argument("--validate", action="store_true")
    g.add_argument("--test", action="store_true")
    g.add_argument("--export", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args, extras = ap.parse_known_args()
    if args.export:
        raise NotImplementedError("--export (mesh extraction) is not implemented")
    if "LOCAL_RANK" not in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    import torch

    import scaledreamer_b200 as sd
    from scaledreamer_b200.core import get_rank
    from scaledreamer_b200.systems import Trainer

    n_gpus = int(os.environ.get("WORLD_SIZE", "1"))
    if n_gpus > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl")
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    cfg = sd.load_config(args.config, cli_args=extras, n_gpus=n_gpus)
    import random

    import numpy as np

    seed = cfg.seed + get_rank()  # launch.py:171 of the reference
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    save_dir = os.path.join(cfg.trial_dir, "save")

    def on_validation(outs, step):  # scaledreamer.py:172-233: it{step}-val/...; multi-prompt ranks hold different prompts
        if get_rank() == 0 or (outs and "name" in outs[0]):
            save_views(outs, os.path.join(save_dir, f"it{step}-val"))

    trainer = Trainer(**cfg.trainer, ckpt_dir=os.path.join(cfg.trial_dir, "ckpts"), checkpoint=cfg.checkpoint,
                      on_validation=on_validation)
    if getattr(cfg, "resume", None):
        trainer.load_checkpoint(cfg.resume, system)  # module state, step counters, schedules; optimizer state in fit()
    if args.train:
        trainer.fit(system, dm)
        if get_rank() == 0 and trainer.history:
            print(trainer.history[-1])
        outs = trainer.test(system, dm)  # the reference's --train ends with trainer.test (launch.py:247-248)
        if get_rank() == 0 or (outs and "name" in outs[0]):
            save_views(outs, os.path.join(save_dir, f"it{trainer.global_step}-test"))
        return
    outs = trainer.validate(system, dm) if args.validate else trainer.test(system, dm)
    if get_rank() == 0 or (outs and "name" in outs[0]):
        save_views(outs, os.path.join(save_dir, "val" if args.validate else "test"))


def save_views(outs, out_dir: str) -> None:
    """One PNG per evaluation view: rgb | [normal |] opacity | normalised depth (the reference's save_image_grid row). The
    multi-prompt systems return several views per batch and a `name`: those go to <out_dir>/<name>/<index>.png
    (multiprompt_radience_field_generator.py:236-240)."""
    import numpy as np
    import torch
    from PIL import Image

    n = 0
    for o in outs:
        sub = os.path.join(out_dir, o["name"]) if "name" in o else out_dir
        os.makedirs(sub, exist_ok=True)
        for v in range(o["comp_rgb"].shape[0]):
            rgb = o["comp_rgb"][v].clamp(0, 1)
            gray = lambda t: t.reshape(rgb.shape[0], rgb.shape[1], 1).clamp(0, 1).expand(-1, -1, 3)
            cells = [rgb]
            if "comp_normal" in o:
                cells.append(o["comp_normal"][v].clamp(0, 1))
            cells.append(gray(o["opacity"][v]))
            if "depth" in o:
                cells.append(gray(o["depth"] if o["depth"].dim() == 2 else o["depth"][v]))
            img = (torch.cat(cells, dim=1) * 255.0).round().to(torch.uint8).cpu().numpy()
            Image.fromarray(np.ascontiguousarray(img)).save(os.path.join(sub, f"{int(o['index'][v])}.png"))
            n += 1
    print(f"wrote {n} views to {out_dir}")


if __name__ == "__main__":
    sys.exit(main())
