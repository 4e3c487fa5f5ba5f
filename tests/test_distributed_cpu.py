"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding, parameter broadcast, the
flat-gradient all-reduce between backward and clip_grad_norm_ (train.py:246-247), and a PPO update that
keeps the replicas bit-identical while each rank trains on its own shard."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppo_radiotherapy_b200.networks import PPO
from ppo_radiotherapy_b200.train import (DEFAULTS, FlatGradAllReduce, broadcast_parameters, finalize_config,
                                         load_config, ppo_update, shard_range)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                       # different init per rank ...
        agent = PPO((9,), (6,), 64)
        broadcast_parameters(agent)                         # ... made identical here
        flat0 = torch.cat([p.detach().reshape(-1) for p in agent.parameters()])
        gathered = [torch.zeros_like(flat0) for _ in range(world)]
        dist.all_gather(gathered, flat0)
        assert all(torch.equal(g, gathered[0]) for g in gathered)

        # flat gradient all-reduce == mean of the per-rank gradients
        sync = FlatGradAllReduce(agent.parameters())
        x = torch.randn(32, 9, generator=torch.Generator().manual_seed(rank))
        agent.get_value(x).sum().backward()                 # actor head gets no gradient: exercises grad=None
        local = [None if p.grad is None else p.grad.clone() for p in agent.parameters()]
        sync()
        for p, g in zip(agent.parameters(), local):
            g = torch.zeros_like(p) if g is None else g
            both = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(both, g)
            assert torch.allclose(p.grad, sum(both) / world, rtol=0, atol=1e-7)

        # a PPO update on rank-local shards keeps the replicas identical
        cfg = SimpleNamespace(**DEFAULTS)
        cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.update_epochs, cfg.total_timesteps = 8, 16, 4, 2, 1024
        finalize_config(cfg, world)
        lo, hi = shard_range(cfg.num_envs, world, rank)
        n = (hi - lo) * cfg.num_steps
        g = torch.Generator().manual_seed(7 + rank)
        b_obs, b_act = torch.randn(n, 9, generator=g), torch.randn(n, 6, generator=g)
        with torch.no_grad():
            _, lp, _, v = agent.get_action_and_value(b_obs, b_act)
        adv, ret = torch.randn(n, generator=g), torch.randn(n, generator=g)
        opt = torch.optim.Adam(agent.parameters(), lr=3e-4, eps=1e-5)
        stats = ppo_update(agent, opt, cfg, b_obs, b_act, lp, adv, ret, v.reshape(-1), sync,
                           generator=torch.Generator().manual_seed(5))
        assert set(stats) >= {"pg_loss", "v_loss", "entropy", "approx_kl", "clipfrac"}
        flat1 = torch.cat([p.detach().reshape(-1) for p in agent.parameters()])
        assert not torch.equal(flat0, flat1)
        gathered = [torch.zeros_like(flat1) for _ in range(world)]
        dist.all_gather(gathered, flat1)
        assert all(torch.equal(gg, gathered[0]) for gg in gathered)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_sync(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_range_and_config(tmp_path):
    assert [shard_range(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 65536)]
    covered = sorted(i for r in range(4) for i in range(*shard_range(16, 4, r)))
    assert covered == list(range(16))
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
    # derived sizes of train.py:292-297 from the reference's template values
    p = tmp_path / "c.yaml"
    p.write_text("num_envs: 16\nnum_steps: 2048\nnum_minibatches: 32\ntotal_timesteps: 10000000\nnum_saves: 5\n")
    cfg = load_config(str(p))
    assert (cfg.batch_size, cfg.minibatch_size, cfg.num_iterations, cfg.save_frequency_iterations) == (32768, 1024, 305, 61)
    assert cfg.gamma == 0.99 and cfg.visionless is True
    with pytest.raises(ValueError):
        finalize_config(cfg, world_size=3)


def test_reference_checkpoint_layout():
    keys = set(PPO((9,), (6,), 64).state_dict())
    assert keys == {"actor_logstd"} | {f"{h}.{i}.{w}" for h in ("critic", "actor_mean") for i in (0, 2, 4)
                                       for w in ("weight", "bias")}
    n = sum(p.numel() for p in PPO((9,), (6,), 64).parameters())
    assert n == 10061
