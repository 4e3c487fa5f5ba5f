#!/usr/bin/env python
"""Per-stage clock64() profile of the step kernel (rt_set_stage_clock)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
eng = rt.BatchedEpisodes(n, device=dev); eng.reset()
g = torch.Generator(device=dev).manual_seed(0)
acts = torch.rand((40, n, 6), device=dev, generator=g) * 2 - 1
for i in range(20): eng.step(acts[i], want_info=False)
stamps = torch.zeros((n, 12), dtype=torch.int64, device=dev)
nat.check(nat.lib().rt_set_stage_clock(eng._h, C.c_void_p(stamps.data_ptr())))
res = []
for i in range(20, 40):
    eng.step(acts[i], want_info=False); torch.cuda.synchronize()
    s = stamps.cpu().numpy().astype(np.float64)
    res.append(s)
s = np.stack(res)            # [iters, n, 12]
order = [0, 8, 9, 10, 1, 2, 3, 4, 6, 11, 7]
ok = (s[:, :, order] > 0).all(axis=2)          # envs that walked a beam this step (all stamps written)
rel = s - s[:, :, 0:1]                        # clock64 is per SM: only differences within an env's block are meaningful
names = {0: "scalar warp start", 8: "scalar: state loaded, translated", 9: "scalar: pose updated", 10: "scalar: beam set up",
         1: "scalar: walk done", 2: "env warp: tumour + distance done", 3: "env warp: beams seen (mbarrier)",
         4: "env warp: cells of all passes back", 6: "env warp: all passes stored",
         11: "scalar: past barrier 2", 7: "end"}
print(f"n={n}: cycles since the block's scalar-warp start (mean / p50 / p99 over {int(ok.sum())} env-steps)")
for k in order:
    v = rel[:, :, k][ok]
    print(f"  {k:2d} {names[k]:30s} {v.mean():9.0f} {np.percentile(v,50):9.0f} {np.percentile(v,99):9.0f}")
eng.close()
# the block that finishes last decides the kernel: per step, the largest "end" over all envs
end = np.where(ok, rel[:, :, 7], 0.0)
mx = end.max(axis=1)
am = end.argmax(axis=1)
print(f"  slowest block per step (cycles): mean {mx.mean():.0f}  min {mx.min():.0f}  max {mx.max():.0f}")
for k in order:
    v = np.array([rel[i, am[i], k] for i in range(rel.shape[0])])
    print(f"     slowest env's stamp {k:2d} {names[k]:30s} mean {v.mean():9.0f}")
