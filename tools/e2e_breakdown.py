#!/usr/bin/env python
"""Where the host-buffer step (rt_step_host) spends its 30 us: launch + sync floor, device-pointer step, host-buffer step."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import _native as nat

n = 4096
dev = torch.device("cuda:0")
eng = rt.BatchedEpisodes(n, device=dev); eng.reset()
act = torch.rand((n, 6), device=dev) * 2 - 1
h_act = torch.empty((n, 6), dtype=torch.float32, pin_memory=True); h_act.copy_(act)
h_obs = torch.empty((n, 9), dtype=torch.float32, pin_memory=True)
h_rew = torch.empty(n, dtype=torch.float64, pin_memory=True)
h_term = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_trunc = torch.empty(n, dtype=torch.uint8, pin_memory=True)
call = eng.bind_step_host(h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_term.numpy(), h_trunc.numpy())


def wall(fn, reps=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


s = torch.cuda.current_stream(dev)
z = torch.zeros(1, device=dev)
print(f"tiny torch kernel + stream sync      {wall(lambda: (z.add_(1), s.synchronize())):6.1f} us")
print(f"rt_step (device buffers) + sync      {wall(lambda: (eng.step(act, want_info=False), s.synchronize())):6.1f} us")
print(f"rt_step (device buffers), no sync    {wall(lambda: eng.step(act, want_info=False)):6.1f} us per call (device-bound)")
print(f"rt_step_host (pinned host buffers)   {wall(call):6.1f} us")
eng.close()
