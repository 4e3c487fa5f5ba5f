import sys; sys.path.insert(0, '/root/repo')
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0")
def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) * 1e-3 / reps
g = torch.Generator(device=dev).manual_seed(7)
def dense():
    n = 1024
    de = rt.BatchedEpisodes(n, device=dev, dense=True, seed=3); de.reset()
    a = torch.rand((n, 6), device=dev, generator=g) * 2 - 1
    s = timed(lambda: de.step(a, want_info=False), 10)
    s2 = timed(lambda: de.step(a, want_info=False), 30)
    print("dense us", s*1e6, s2*1e6, "GB/s", n*2*de.nvox*4/s2/1e9); de.close()
dense()
n = 65536
se = rt.BatchedEpisodes(n, device=dev, seed=11); se.reset()
a = torch.rand((n, 6), device=dev, generator=g) * 2 - 1
print("big us", timed(lambda: se.step(a, want_info=False), 20)*1e6); se.close()
dense()
torch.cuda.empty_cache()
dense()
