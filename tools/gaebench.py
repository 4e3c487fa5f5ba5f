import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0")
T, N = 128, 65536
r = torch.randn((T, N), device=dev); v = torch.randn((T, N), device=dev); d = torch.zeros((T, N), device=dev)
oa = torch.empty_like(r); ob = torch.empty_like(r)
for i in range(8): rt.compute_gae(r, v, d, r[0], d[0], 0.99, 0.95, out=(oa, ob))
torch.cuda.synchronize(); print("ok")
