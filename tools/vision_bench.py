#!/usr/bin/env python
"""BASELINE configs[3]: voxel observations + FeaturesExtractor3D at 1024 envs.  Times the observation assembly
kernel, the C3D forward (PyTorch/cuDNN, bf16 channels-last-3d vs fp32) and the whole vision rollout step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppo_radiotherapy_b200 as rt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")

def timed(fn, reps=reps, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

envs = rt.RadiotherapyVectorEnv(n, visionless=False, device=dev, seed=1)
obs, _ = envs.reset(options={"backend": "torch"})
a = torch.rand((n, 6), device=dev) * 2 - 1
for _ in range(10): envs.engine.step(a, want_info=False)
s = timed(lambda: envs._volumes())
print(f"obs assembly  n={n}: {s*1e3:8.3f} ms  {n*5*envs.engine.nvox*4/s/1e9:7.0f} GB/s  {n/s:10.0f} env-obs/s")
flops = 0.761e9 * n
for name, dt in (("bf16 autocast", torch.bfloat16), ("fp32", None)):
    agent = rt.PPO_3DCNN(envs.single_observation_space.shape, (6,), 64, compute_dtype=dt).to(dev)
    with torch.no_grad():
        s = timed(lambda: agent.features_extractor(obs))
    print(f"C3D forward   {name:22s}: {s*1e3:8.3f} ms  {flops/s/1e12:6.2f} TFLOP/s  {n/s:10.0f} samples/s")
agent = rt.PPO_3DCNN(envs.single_observation_space.shape, (6,), 64, compute_dtype=torch.bfloat16).to(dev)
def step():
    with torch.no_grad():
        action, logprob, _, value = agent.get_action_and_value(envs._vol)
    envs.step(action)
s = timed(step)
print(f"vision rollout step (policy + env + obs): {s*1e3:8.3f} ms  {n/s:10.0f} env-steps/s")

def step_env():
    with torch.no_grad():
        action, logprob, _, value = agent.get_action_and_value_from_env(envs.engine)
    envs.engine.step(action, want_info=False)
s = timed(step_env)
print(f"vision rollout step, observation never materialised (rt_conv1_from_env): {s*1e3:8.3f} ms  {n/s:10.0f} env-steps/s")
