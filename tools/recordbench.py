#!/usr/bin/env python
"""Compressed observation records: pack and render bandwidth (rt_pack_observations / rt_render_observations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0"); n = 256
eng = rt.BatchedEpisodes(n, device=dev, seed=1); eng.reset()
a = torch.rand((n, 6), device=dev) * 2 - 1
for _ in range(30): eng.step(a, want_info=False)
store = eng.observation_store(8 * n)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
k = [0]
def pack():
    eng.pack_observations(store, (k[0] % 8) * n); k[0] += 1
s = timed(pack)
V = eng.nvox
print(f"pack   {n} envs: {s*1e6:8.1f} us  {n*V*6/s/1e9:7.0f} GB/s (4 B read + 2 B written per voxel)")
idx = torch.randperm(8 * n, device=dev)[:n].contiguous()
out = torch.empty((n, 4) + eng.grid, device=dev)
s = timed(lambda: eng.render_observations(store, idx, out=out))
print(f"render {n} records: {s*1e6:8.1f} us  {n*V*18/s/1e9:7.0f} GB/s (2 B read + 16 B written per voxel)")
