#!/usr/bin/env python
"""What makes an env's deposition slow?  Per-env duration (stage clocks of the instrumented 28-env kernel) against the
beam it deposited: voxels hit, slabs, warp index in the block."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import _native as nat

n = 4096
dev = torch.device("cuda:0")
eng = rt.BatchedEpisodes(n, device=dev); eng.reset()
g = torch.Generator(device=dev).manual_seed(0)
acts = torch.rand((60, n, 6), device=dev, generator=g) * 2 - 1
for i in range(40): eng.step(acts[i], want_info=False)
stamps = torch.zeros((n, 12), dtype=torch.int64, device=dev)
nat.check(nat.lib().rt_set_stage_clock(eng._h, C.c_void_p(stamps.data_ptr())))
D, CNT, SL, LE, FR = [], [], [], [], []
for i in range(40, 60):
    eng.step(acts[i], want_info=False); torch.cuda.synchronize()
    s = stamps.cpu().numpy().astype(np.float64)
    pose = eng.pose()
    idx, w, count = rt.beam_voxels_batch(pose[:, :3].contiguous(), pose[:, 3:].contiguous())
    idx = idx.cpu().numpy(); count = count.cpu().numpy()
    ok = (s > 0).all(axis=1)
    D.append((s[:, 6] - s[:, 3])[ok]); CNT.append(count[ok]); LE.append((np.arange(n) % 28)[ok])
    sec = [len(set((idx[e, :count[e]] >> 3).tolist())) for e in np.nonzero(ok)[0]]
    SL.append(np.array(sec))
D, CNT, LE, SL = map(np.concatenate, (D, CNT, LE, SL))
print(f"{len(D)} env-steps: deposition (barrier 1 -> all passes done) mean {D.mean():.0f} p50 {np.percentile(D,50):.0f} p90 {np.percentile(D,90):.0f} p99 {np.percentile(D,99):.0f}")
print("corr(duration, voxels hit) %.2f   corr(duration, distinct sectors) %.2f" % (np.corrcoef(D, CNT)[0, 1], np.corrcoef(D, SL)[0, 1]))
for lo, hi in ((0, 40), (40, 80), (80, 120), (120, 160), (160, 200), (200, 400)):
    m = (CNT >= lo) & (CNT < hi)
    if m.sum(): print(f"  voxels {lo:3d}-{hi:3d}: {m.mean()*100:5.1f}% of env-steps, duration mean {D[m].mean():7.0f} p90 {np.percentile(D[m],90):7.0f}")
for lo, hi in ((0, 20), (20, 40), (40, 60), (60, 80), (80, 100), (100, 200)):
    m = (SL >= lo) & (SL < hi)
    if m.sum(): print(f"  sectors {lo:3d}-{hi:3d}: {m.mean()*100:5.1f}% of env-steps, duration mean {D[m].mean():7.0f} p90 {np.percentile(D[m],90):7.0f}")
print("  by warp index mod 4 (scheduler):", " ".join(f"{D[(LE + 1) % 4 == q].mean():.0f}" for q in range(4)))
eng.close()
