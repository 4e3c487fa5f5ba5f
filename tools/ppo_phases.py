#!/usr/bin/env python
"""Where a PPO iteration spends its time (visionless MLP policy): rollout / GAE / update, CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import train as T

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
cfg = T.load_config(None, num_envs=envs, num_steps=128, num_minibatches=4, update_epochs=2,
                    total_timesteps=envs * 128 * 4, num_saves=0, save_model=False, seed=1)
marks = []
orig_update, orig_gae = T.ppo_update, T.compute_gae

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def timed_update(*a, **k):
    e0 = ev(); r = orig_update(*a, **k); e1 = ev(); marks.append(("update", e0, e1)); return r

def timed_gae(*a, **k):
    e0 = ev(); r = orig_gae(*a, **k); e1 = ev(); marks.append(("gae", e0, e1)); return r

T.ppo_update, T.compute_gae = timed_update, timed_gae
t0 = time.time()
agent = T.train(cfg, None, dev, None, "p", log=None)
torch.cuda.synchronize()
wall = time.time() - t0
tot = {}
for name, e0, e1 in marks:
    tot[name] = tot.get(name, 0.0) + e0.elapsed_time(e1)
iters = len([m for m in marks if m[0] == "update"])
print(f"envs={envs} iterations={iters} wall={wall:.2f}s  sps(last)={agent.history[-1]['sps']:.0f}")
for k, v in tot.items():
    print(f"  {k:8s} {v/iters:8.2f} ms / iteration")
print(f"  steps per iteration {envs*128}: update+gae {sum(tot.values())/iters:.1f} ms; the rest of {wall/iters*1e3:.1f} ms wall is the rollout (+ set-up amortised)")
