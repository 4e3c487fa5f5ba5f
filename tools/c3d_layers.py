import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0"); n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fe = rt.FeaturesExtractor3D((4, 67, 43, 70), 64).to(dev)
x = torch.rand((n, 4, 67, 43, 70), device=dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): y = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mode in ("bf16_cl3d", "bf16", "fp32"):
    with torch.no_grad():
        h = x
        if mode == "bf16_cl3d": h = x.contiguous(memory_format=torch.channels_last_3d)
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "fp32" else torch.autocast("cuda", enabled=False)
        with ctx:
            print(mode)
            for i, layer in enumerate(fe.cnn):
                ms = timed(lambda: layer(h))
                h = layer(h)
                print(f"   {i} {layer.__class__.__name__:10s} {ms:8.3f} ms  out {tuple(h.shape)} {h.dtype}")
