"""Command-line entry with the reference's contract (launch.py:267-299 there):
    python launch.py --config configs/x.yaml --train [--gpu 0] key=value ...
Only --train is implemented (the ASD hot path); validate / test / export are outside the scope of this repo.
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
    if not args.train:
        raise NotImplementedError("only --train is implemented")
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
    trainer = Trainer(**cfg.trainer)
    trainer.fit(system, dm)
    if get_rank() == 0 and trainer.history:
        print(trainer.history[-1])


if __name__ == "__main__":
    sys.exit(main())
