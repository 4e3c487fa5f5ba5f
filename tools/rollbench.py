#!/usr/bin/env python
"""PPO rollout of T steps on n envs (train.py:138-161): one rt_rollout launch against T x (rt_ppo_act, rt_step,
rt_ppo_record) replayed from a CUDA graph.   python tools/rollbench.py [n=8192] [T=128]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppo_radiotherapy_b200 as rt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
agent = rt.PPO((9,), (6,), 64).to(dev)
eng = rt.BatchedEpisodes(n, device=dev, seed=1); eng.reset()
fr = rt.FusedRollout(agent, n, T, seed=1)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def one_launch():
    fr.begin_iteration(); fr.rollout(eng, T)
ms = timed(one_launch)
print(f"rt_rollout            n={n} T={T}: {ms:8.3f} ms  {ms / T * 1e3:7.2f} us/step  {n * T / ms / 1e3:8.1f} M env-steps/s")

def step():
    fr.act(eng.obs); eng.step(fr.action, want_info=True); fr.record(eng); fr.advance()
s = torch.cuda.Stream(dev)
with torch.cuda.stream(s):
    for _ in range(3): step()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    def per_step():
        fr.begin_iteration()
        for _ in range(T): g.replay()
    ms = timed(per_step)
print(f"act + step + record   n={n} T={T}: {ms:8.3f} ms  {ms / T * 1e3:7.2f} us/step  {n * T / ms / 1e3:8.1f} M env-steps/s (CUDA graph per step)")
