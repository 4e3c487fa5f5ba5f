#!/usr/bin/env python
"""Device-time the sparse step for a list of env counts, the way bench.py does (steady state incl. autoreset calls,
pool of action batches, CUDA graphs of 50 steps, CUDA events); `nopdl:` in front of a count switches programmatic
dependent launch off (rt_set_pdl), `eager:` launches every step from Python instead of replaying graphs.

    python tools/stepbench.py 4096 nopdl:4096 eager:4096 8192 65536
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

S_BYTES = 5637.0
PEAK = 6460.5


def run(mode, n, steps=600, warm=150, pool=104, chunk=50):
    import ppo_radiotherapy_b200 as rt
    from ppo_radiotherapy_b200 import _native as nat
    dev = torch.device("cuda:0")
    eng = rt.BatchedEpisodes(n, device=dev)
    if "nopdl" in mode:
        nat.check(nat.lib().rt_set_pdl(eng._h, 0))
    eng.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    acts = torch.rand((pool, n, 6), device=dev, generator=g) * 2 - 1
    s = torch.cuda.Stream()
    graphs = []
    with torch.cuda.stream(s):
        for i in range(warm):
            eng.step(acts[i % pool], want_info=False)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = steps // chunk
        if "eager" in mode:
            e0.record(s)
            for i in range(reps * chunk):
                eng.step(acts[i % pool], want_info=False)
            e1.record(s)
        else:
            for gi in range(max(1, pool // chunk)):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s):
                    for j in range(chunk):
                        eng.step(acts[(gi * chunk + j) % pool], want_info=False)
                graphs.append(gr)
            for gr in graphs:
                gr.replay()
            s.synchronize()
            e0.record(s)
            for i in range(reps):
                graphs[i % len(graphs)].replay()
            e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) / (reps * chunk) * 1e3
    eng.close()
    gbs = S_BYTES * n / us / 1e3
    print(f"{mode or 'graphs+pdl':>12} n={n:>6}: {us:8.2f} us/step  {n / us:8.1f} M env-steps/s  "
          f"{gbs:7.1f} GB/s = {gbs / PEAK:.3f} of peak", flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        mode, _, n = spec.rpartition(":")
        run(mode, int(n))
