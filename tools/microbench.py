#!/usr/bin/env python
"""Device-time the individual entry points (CUDA events, many repetitions) — where does a step go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppo_radiotherapy_b200 as rt

dev = torch.device("cuda:0")
def timeit(fn, reps=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
        g.replay(); s.synchronize()
        e0.record(s); g.replay(); e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us

for m in (4096, 32768):
    g = torch.Generator(device=dev).manual_seed(0)
    pos = torch.rand((m, 3), device=dev, generator=g, dtype=torch.float64) * torch.tensor([67., 43., 70.], device=dev, dtype=torch.float64)
    d = torch.randn((m, 3), device=dev, generator=g, dtype=torch.float64); d /= d.norm(dim=1, keepdim=True)
    a = torch.rand((m, 6), device=dev, generator=g) * 2 - 1
    print(f"m={m}")
    print("  pose_update      %8.2f us" % timeit(lambda: rt.pose_update_batch(pos, d, a)))
    print("  beam_voxels      %8.2f us" % timeit(lambda: rt.beam_voxels_batch(pos, d)))
    eng = rt.BatchedEpisodes(m, device=dev)
    eng.reset()
    print("  step             %8.2f us" % timeit(lambda: eng.step(a, want_info=False)))
    print("  step+info        %8.2f us" % timeit(lambda: eng.step(a, want_info=True)))
    r = torch.randn((128, m), device=dev); 
    print("  gae T=128        %8.2f us" % timeit(lambda: rt.compute_gae(r, r, torch.zeros_like(r), r[0], r[1], 0.99, 0.95), reps=20))
    eng.close()
