import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0"); n = 256
fe = rt.FeaturesExtractor3D((4, 67, 43, 70), 64, compute_dtype=torch.bfloat16).to(dev)
x = torch.rand((n, 4, 67, 43, 70), device=dev)
with torch.no_grad():
    for _ in range(3): y = fe(x)
torch.cuda.synchronize(); print("ok", y.shape)
