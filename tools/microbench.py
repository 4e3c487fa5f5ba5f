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
    zz = torch.zeros_like(r); oa = torch.empty_like(r); ob = torch.empty_like(r)
    us = timeit(lambda: rt.compute_gae(r, r, zz, r[0], r[1], 0.99, 0.95, out=(oa, ob)), reps=20)
    print("  gae T=128        %8.2f us  -> %.0f GB/s (20 B per element)" % (us, 128 * m * 20 / us / 1e3))
    eng.close()
    if m <= 4096:
        nd = 1024
        de = rt.BatchedEpisodes(nd, device=dev, dense=True); de.reset()
        ad = a[:nd].contiguous()
        us = timeit(lambda: de.step(ad, want_info=False), reps=20)
        byt = nd * 2 * de.phantom.nvox * 4
        print("  dense step n=%d  %8.2f us  -> %.0f GB/s (R+W of the dose volumes)" % (nd, us, byt / us / 1e3))
        de.close()
        ve = rt.BatchedEpisodes(256, device=dev); ve.reset()
        for i in range(10): ve.step(a[:256].contiguous(), want_info=False)
        out = torch.empty((256, 4) + ve.grid, dtype=torch.float32, device=dev)
        us = timeit(lambda: ve.volumes(0, 256, out=out), reps=20)
        print("  volumes n=256     %8.2f us  -> %.0f GB/s (5 x V x 4 B per env)" % (us, 256 * 5 * ve.nvox * 4 / us / 1e3))
        ve.close()
