#!/usr/bin/env python
"""Algorithmic bytes per env-step of the sparse visionless step, measured on the
bench workload with the CPU oracle: uniform(-1,1) actions, tumour id (i*7919) mod 1000.

  U      distinct voxels hit by the step's beam
  W      in-bounds splat writes of the beam (before merging duplicates)
  Sec    distinct 32-byte dose sectors hit by the beam (dose volume 128-byte aligned per env)
  Sec_ep distinct sectors hit over the whole 100-step episode
  first  sectors hit for the first time in the episode (no HBM read needed with the sector-valid bitmap)

  payload  P = 202 + 8 U + 4 W
  sector   S = 256 + 64 Sec + 32 Sec_ep / 100          (the reference dataflow: RMW every sector + sparse clear)
  design   D = 256 + 32 Sec + 32 (Sec - first)          (this design: write every sector, read re-touched ones)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O

def main(n=96, T=100, kind="uniform"):
    ph = O.Phantom()
    rng = np.random.default_rng(0)
    acts = (rng.uniform(-1, 1, (T, n, 6)) if kind == "uniform" else rng.standard_normal((T, n, 6))).astype(np.float32)
    U, W, SEC, FIRST, SEC_EP, SLABS = [], [], [], [], [], []
    for i in range(n):
        env = O.OracleEnv(ph, (i * 7919) % 1000)
        seen = set()
        for t in range(T):
            env.step(acts[t, i])
            p = env.pose
            idx, w, ns = O.beam_trace(p[:3], p[3:])
            sec = set((idx >> 3).tolist())
            U.append(len(set(idx.tolist()))); W.append(len(idx)); SEC.append(len(sec)); SLABS.append(ns)
            FIRST.append(len(sec - seen)); seen |= sec
        SEC_EP.append(len(seen))
    U, W, SEC, FIRST, SEC_EP, SLABS = map(np.array, (U, W, SEC, FIRST, SEC_EP, SLABS))
    P = 202 + 8 * U.mean() + 4 * W.mean()
    S = 256 + 64 * SEC.mean() + 32 * SEC_EP.mean() / T
    D = 256 + 32 * SEC.mean() + 32 * (SEC.mean() - FIRST.mean())
    print(f"{kind} actions, {n} envs x {T} steps")
    print(f"  slabs/beam mean {SLABS.mean():.1f} max {SLABS.max()};  U mean {U.mean():.1f} max {U.max()};  W mean {W.mean():.1f} max {W.max()}")
    print(f"  Sec mean {SEC.mean():.1f} max {SEC.max()};  first-touch {FIRST.mean():.1f} ({100*FIRST.mean()/SEC.mean():.0f}%);  Sec_ep mean {SEC_EP.mean():.0f}")
    print(f"  payload P = {P:.0f} B/env-step;  sector-granular S = {S:.0f} B/env-step;  this design D = {D:.0f} B/env-step")

if __name__ == "__main__":
    main(kind="uniform")
    main(kind="normal")
