"""The CleanRL-style loop on the device-resident env (GPU): runs, learns something measurable in a few
iterations of a tiny config, writes reference-compatible checkpoints, and the reference's own loop body
works verbatim on top of RadiotherapyVectorEnv (numpy seam and tensor seam)."""
import os

import numpy as np
import pytest
import torch

import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200.train import load_config, train

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_train_tiny(tmp_path):
    cfg = load_config(None, num_envs=64, num_steps=100, num_minibatches=4, update_epochs=2,
                      total_timesteps=64 * 100 * 3, num_saves=1, seed=3)
    torch.manual_seed(0)
    agent = train(cfg, writer=None, device=DEV, output_dir=str(tmp_path), run_name="t", log=None)
    h = agent.history
    assert len(h) == 3 and h[-1]["global_step"] == 64 * 100 * 3
    assert all(np.isfinite(r["v_loss"]) and np.isfinite(r["pg_loss"]) for r in h)
    assert h[0]["episodes"] == 64 and abs(h[0]["episodic_length"] - 100) < 1e-9
    assert -200 < h[0]["episodic_return"] < 200
    path = tmp_path / "models" / "t" / "default_3.model"
    assert path.exists()
    sd = torch.load(path, map_location="cpu", weights_only=True)
    assert set(sd) == set(rt.PPO((9,), (6,), 64).state_dict())


def test_reference_loop_body_verbatim():
    """train.py:138-161 / ppo_eval.py:20-35 line for line against the drop-in vector env (numpy seam)."""
    envs = rt.RadiotherapyVectorEnv(16, visionless=True, device=DEV, seed=1)
    device = torch.device(DEV)
    agent = rt.PPO(envs.single_observation_space.shape, envs.single_action_space.shape, 64).to(device)
    assert isinstance(envs.single_action_space, rt.Box)
    next_obs, _ = envs.reset(seed=1)
    next_obs = torch.Tensor(next_obs).to(device)
    episodic_returns = []
    for step in range(101):
        with torch.no_grad():
            action, logprob, _, value = agent.get_action_and_value(next_obs)
        next_obs, reward, terminations, truncations, infos = envs.step(action.cpu().numpy())
        next_done = np.logical_or(terminations, truncations)
        rewards = torch.tensor(reward).to(device).view(-1)
        next_obs, next_done = torch.Tensor(next_obs).to(device), torch.Tensor(next_done).to(device)
        assert rewards.shape == (16,) and next_obs.shape == (16, 9)
        if "episode" in infos.keys():
            ep_returns = infos["episode"]["r"]
            ep_completions = infos["episode"]["_r"]
            episodic_returns.append(np.mean(ep_returns[ep_completions]))
            for k in ("tumour", "lung", "distance_to_tumour", "total"):
                assert np.isfinite(np.mean(infos["reward_components"][k][ep_completions]))
    assert len(episodic_returns) == 1          # all 16 episodes end at t = 100; call 101 is the autoreset
    envs.close()


def test_vision_agent_step():
    envs = rt.RadiotherapyVectorEnv(4, visionless=False, device=DEV, seed=2)
    agent = rt.PPO_3DCNN(envs.single_observation_space.shape, envs.single_action_space.shape, 64,
                         compute_dtype=torch.bfloat16).to(DEV)
    obs, _ = envs.reset(options={"backend": "torch"})
    assert obs.shape == (4, 4, 67, 43, 70)
    for _ in range(3):
        with torch.no_grad():
            action, logprob, entropy, value = agent.get_action_and_value(obs)
        obs, reward, term, trunc, _ = envs.step(action)
        assert torch.isfinite(value).all() and obs.min() >= 0 and obs.max() <= 1
    envs.close()
