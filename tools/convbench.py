#!/usr/bin/env python
"""Fused Conv3d(4->16,k3)+bias+ReLU+MaxPool3d block (rt_conv1_relu_pool) timed alone: median / min over 20 launches."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0"); n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.rand((n, 4, 67, 43, 70), device=dev)
fe = rt.FeaturesExtractor3D((4, 67, 43, 70), 64, compute_dtype=torch.bfloat16).to(dev)
with torch.no_grad():
    for _ in range(10): fe._fused_first_block(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = fe._fused_first_block(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
ms, mn = float(np.median(ts)), float(np.min(ts))
print(f"fused conv1 block n={n}: median {ms:.3f} ms (min {mn:.3f})  {ms*1e3/n:.2f} us/sample  "
      f"{n*0.626e9/ms/1e9:.1f} TFLOP/s useful  {n*(3226720 + 16*33*21*34*2)/ms/1e6:.0f} GB/s")
