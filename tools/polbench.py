#!/usr/bin/env python
"""rt_ppo_act alone (for ncu): tools/polbench.py [envs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppo_radiotherapy_b200 as rt

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
agent = rt.PPO((9,), (6,), 64).to(dev)
fr = rt.FusedRollout(agent, envs, 8, seed=1)
obs = torch.rand((envs, 9), device=dev) * 2 - 1
for _ in range(5):
    fr.act(obs)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(100):
            fr.act(obs)
    g.replay(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); g.replay(); e1.record(s); s.synchronize()
print(f"rt_ppo_act, {envs} envs: {e0.elapsed_time(e1) * 10:.2f} us per launch (graph of 100)")
