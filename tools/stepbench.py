#!/usr/bin/env python
"""Device-time the sparse step for a list of (RT_STEP_KB[:RT_SPLIT_KW], env count) pairs, the way bench.py does
(steady state incl. autoreset calls, pool of action batches, CUDA graphs of 50 steps, CUDA events).

    python tools/stepbench.py 14@4096 -2:8@4096 14@65536 -2:8@65536 -2:16@65536
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

S_BYTES = 5637.0
PEAK = 6460.5


def run(kb, kw, n, steps=600, warm=150, pool=64, chunk=50):
    os.environ["RT_STEP_KB"] = kb
    if kw:
        os.environ["RT_SPLIT_KW"] = kw
    import ppo_radiotherapy_b200 as rt
    dev = torch.device("cuda:0")
    eng = rt.BatchedEpisodes(n, device=dev)
    eng.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    acts = torch.rand((pool, n, 6), device=dev, generator=g) * 2 - 1
    s = torch.cuda.Stream()
    graphs = []
    with torch.cuda.stream(s):
        for i in range(warm):
            eng.step(acts[i % pool], want_info=False)
        s.synchronize()
        for gi in range(max(1, pool // chunk)):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for j in range(chunk):
                    eng.step(acts[(gi * chunk + j) % pool], want_info=False)
            graphs.append(gr)
        for gr in graphs:
            gr.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        reps = steps // chunk
        for i in range(reps):
            graphs[i % len(graphs)].replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) / (reps * chunk) * 1e3
    eng.close()
    gbs = S_BYTES * n / us / 1e3
    print(f"RT_STEP_KB={kb:>3} kw={kw or '-':>2} n={n:>6}: {us:8.2f} us/step  {n / us:8.1f} M env-steps/s  "
          f"{gbs:7.1f} GB/s = {gbs / PEAK:.3f} of peak", flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        v, n = spec.split("@")
        kb, _, kw = v.partition(":")
        run(kb, kw, int(n))
