import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0")
n = 1024
de = rt.BatchedEpisodes(n, device=dev, dense=True); de.reset()
a = torch.rand((n, 6), device=dev) * 2 - 1
for i in range(8): de.step(a, want_info=False)
torch.cuda.synchronize(); print("ok")
