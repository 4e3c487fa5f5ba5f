import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0")
ve = rt.BatchedEpisodes(256, device=dev); ve.reset()
a = torch.rand((256, 6), device=dev) * 2 - 1
for i in range(10): ve.step(a, want_info=False)
out = torch.empty((256, 4) + ve.grid, dtype=torch.float32, device=dev)
for i in range(5): ve.volumes(0, 256, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for i in range(10): ve.volumes(0, 256, out=out)
e1.record(); torch.cuda.synchronize()
print("volumes us", e0.elapsed_time(e1) / 10 * 1e3)
