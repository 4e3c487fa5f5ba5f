#!/usr/bin/env python
"""Tiny multi-GPU PPO run (torchrun): env shards per rank, NCCL flat-gradient all-reduce; prints SPS and
checks that the replicas stay identical."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200.train import load_config, train

world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
vision = len(sys.argv) > 2 and sys.argv[2] == "vision"          # voxel observations + C3D, compressed rollout records
steps = int(sys.argv[3]) if len(sys.argv) > 3 else (16 if vision else 128)
cfg = load_config(None, num_envs=envs * world, num_steps=steps, num_minibatches=4, update_epochs=1 if vision else 2,
                  total_timesteps=envs * world * steps * (3 if vision else 10), num_saves=0, save_model=False, seed=1, visionless=not vision)
torch.manual_seed(1 + local)
t0 = time.time()
agent = train(cfg, None, dev, None, "smoke", log=(print if local == 0 else None))
torch.cuda.synchronize()
flat = torch.cat([p.detach().reshape(-1) for p in agent.parameters()])
if world > 1:
    ref = flat.clone(); dist.broadcast(ref, 0)
    same = torch.equal(ref, flat)
    t = torch.tensor([int(same)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if local == 0: print("replicas identical:", bool(t.item()))
if local == 0:
    h = agent.history
    print(f"world={world} envs/rank={envs} steps={h[-1]['global_step']} wall={time.time()-t0:.1f}s sps(cumulative)={h[-1]['sps']:.0f} "
          f"sps(steady, median of last iterations)={sorted(r['iter_sps'] for r in h[len(h)//2:])[len(h[len(h)//2:])//2]:.0f} "
          f"return {h[0].get('episodic_return', float('nan')):.2f} -> {h[-1].get('episodic_return', float('nan')):.2f}")
if world > 1: dist.destroy_process_group()
