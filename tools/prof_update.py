import os, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import train as T
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
n = 8192 * 128
cfg = T.load_config(None, num_envs=8192, num_steps=128, num_minibatches=4, update_epochs=1)
T.finalize_config(cfg, 1)
agent = rt.PPO((9,), (6,), 64).to(dev)
opt = torch.optim.Adam(agent.parameters(), lr=3e-4, eps=1e-5)
g = torch.Generator(device=dev).manual_seed(0)
b_obs = torch.rand((n, 9), device=dev, generator=g); b_act = torch.rand((n, 6), device=dev, generator=g)
b_lp = torch.randn(n, device=dev, generator=g); b_adv = torch.randn(n, device=dev, generator=g)
b_ret = torch.randn(n, device=dev, generator=g); b_val = torch.randn(n, device=dev, generator=g)
for _ in range(2): T.ppo_update(agent, opt, cfg, b_obs, b_act, b_lp, b_adv, b_ret, b_val)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); T.ppo_update(agent, opt, cfg, b_obs, b_act, b_lp, b_adv, b_ret, b_val); e1.record(); torch.cuda.synchronize()
print("one epoch (4 minibatches):", e0.elapsed_time(e1), "ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    T.ppo_update(agent, opt, cfg, b_obs, b_act, b_lp, b_adv, b_ret, b_val); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
