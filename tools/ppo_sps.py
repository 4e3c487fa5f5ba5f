#!/usr/bin/env python
"""PPO env-steps/s per iteration (visionless MLP agent), fused rollout step on / off:  tools/ppo_sps.py [envs] [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_radiotherapy_b200 import train as T

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for fused in (False, True, False, True):
    cfg = T.load_config(None, num_envs=envs, num_steps=128, num_minibatches=4, update_epochs=2,
                        total_timesteps=envs * 128 * iters, num_saves=0, save_model=False, seed=1, fused_rollout=fused)
    torch.manual_seed(0)
    t0 = time.time()
    agent = T.train(cfg, None, torch.device("cuda:0"), None, "p", log=None)
    torch.cuda.synchronize()
    h = agent.history
    sps = [r["iter_sps"] for r in h]
    print(f"fused_rollout={fused}: wall {time.time() - t0:.2f} s; per-iteration SPS (M): " + " ".join(f"{s / 1e6:.1f}" for s in sps))

# the three kernels of the fused rollout step, each alone (CUDA events, 200 launches)
import ppo_radiotherapy_b200 as rt
dev = torch.device("cuda:0")
eng = rt.BatchedEpisodes(envs, device=dev, seed=1)
eng.reset()
agent = rt.PPO((9,), (6,), 64).to(dev)
fr = rt.FusedRollout(agent, envs, 128, seed=1)


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


with torch.no_grad():
    print(f"envs={envs}: rt_ppo_act {timeit(lambda: fr.act(eng.obs)):.1f} us, rt_step {timeit(lambda: eng.step(fr.action, want_info=True)):.1f} us, "
          f"rt_ppo_record {timeit(lambda: fr.record(eng)):.1f} us; PyTorch get_action_and_value alone "
          f"{timeit(lambda: agent.get_action_and_value(eng.obs)):.1f} us (eager launches)")
