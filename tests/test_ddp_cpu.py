"""World-size-2 `gloo` coverage (CPU) of the multi-GPU host logic: ONE all-reduce of the flat
generator-gradient buffer per optimizer step (systems.Trainer._allreduce_grads, replacing Lightning DDP's bucketed
reducer) and the rank-strided prompt sharding of the multi-prompt data module / prompt processor
(custom/amortized/data/multiprompt.py:180-186). The NCCL path differs only in the backend string."""
import json
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, tmp: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scaledreamer_b200.amortized import _world, load_prompt_library, MultipromptRandomCameraDataModuleConfig
        from scaledreamer_b200.systems import Trainer

        # ---- gradient averaging: identical parameters, rank-dependent gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(5)),
                  torch.nn.Parameter(torch.randn(2, 2, 2))]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        params[1].grad = None if False else params[1].grad
        tr = Trainer(max_steps=0, distributed=True)
        assert tr.world_size == world
        tr._allreduce_grads(params)
        for i, p in enumerate(params):
            expect = sum(float(r + 1) * (i + 1) for r in range(world))  # SUM; 1/world is folded into the optimizer
            assert torch.allclose(p.grad, torch.full_like(p, expect)), (rank, i)
        assert tr._flat.numel() == sum(p.numel() for p in params)      # one flat buffer, one collective
        # a parameter without gradient is skipped consistently on every rank
        params[2].grad = None
        tr._flat = None
        tr._allreduce_grads(params)
        assert tr._flat.numel() == params[0].numel() + params[1].numel()

        # ---- initial state: replicas seeded with seed + rank (launch.py) must start from rank 0's generator
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
        net[1].running_mean.fill_(float(rank))
        before = [t.clone() for t in list(net.parameters()) + list(net.buffers())]
        tr.sync_initial_state(net)
        mine = torch.cat([t.detach().float().reshape(-1) for t in list(net.parameters()) + list(net.buffers())])
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1]), "parameters / buffers differ across ranks after sync_initial_state"
        if rank == 0:
            assert all(torch.equal(a, b) for a, b in zip(before, list(net.parameters()) + list(net.buffers())))

        # ---- checkpoints: every rank calls save, rank 0 alone writes (parameters are identical after the all-reduce)
        model = torch.nn.Linear(3, 2)
        model.true_current_epoch = 0
        tr.global_step = 5
        tr.save_checkpoint(os.path.join(tmp, f"ckpt_rank{rank}", "last.ckpt"), model)
        dist.barrier()
        assert os.path.exists(os.path.join(tmp, "ckpt_rank0", "last.ckpt"))
        assert not os.path.exists(os.path.join(tmp, "ckpt_rank1"))

        # ---- prompt sharding: rank r keeps library[r::world] of every split
        assert _world() == (rank, world)
        cfg = MultipromptRandomCameraDataModuleConfig(prompt_library="lib", prompt_library_dir=os.path.join(tmp, "load"))
        lib = load_prompt_library(cfg, *_world())
        allp = [f"prompt {i}" for i in range(7)]
        assert lib["train"] == allp[rank::world] and lib["val"] == allp[:3][rank::world]
        gathered = [None] * world
        dist.all_gather_object(gathered, lib["train"])
        assert sorted(sum(gathered, [])) == sorted(allp)               # disjoint cover of the library
    finally:
        dist.destroy_process_group()


def test_allreduce_and_prompt_sharding_world_size_2(tmp_path):
    os.makedirs(tmp_path / "load")
    allp = [f"prompt {i}" for i in range(7)]
    json.dump({"train": allp, "val": allp[:3], "test": allp[:1]}, open(tmp_path / "load" / "lib.json", "w"))
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
