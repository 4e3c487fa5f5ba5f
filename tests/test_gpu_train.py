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


def test_observation_records_render_like_volumes():
    """Compressed voxel-observation records (rt_pack_observations / rt_render_observations): lungs, tumour and
    beam-view planes equal get_volumes bit for bit, the dose plane is the bfloat16-rounded dose; gathering by index
    picks the right records."""
    n = 12
    eng = rt.BatchedEpisodes(n, device=DEV, seed=9)
    eng.reset()
    g = torch.Generator(device=DEV).manual_seed(3)
    store = eng.observation_store(3 * n)
    want = []
    for k in range(3):
        for _ in range(7):
            eng.step(torch.rand((n, 6), device=DEV, generator=g) * 2 - 1, want_info=False)
        eng.pack_observations(store, k * n)
        want.append(eng.volumes().clone())
    want = torch.cat(want)
    got = eng.render_observations(store)
    assert got.shape == want.shape == (3 * n, 4, 67, 43, 70)
    for plane in (0, 1, 3):
        assert torch.equal(got[:, plane], want[:, plane])
    assert torch.equal(got[:, 2], want[:, 2].bfloat16().float())
    assert float(got[:, 2].max()) > 0.05 and float(got[:, 3].max()) > 0.05
    idx = torch.tensor([35, 0, 17, 17, 4], device=DEV, dtype=torch.int64)
    sel = eng.render_observations(store, idx)
    assert torch.equal(sel, got[idx])
    assert store.nbytes < 0.13 * want.numel() * 4
    # a record taken right after the autoreset call is the empty volume of the new episode
    for _ in range(100 - 21 + 1):
        eng.step(torch.rand((n, 6), device=DEV, generator=g) * 2 - 1, want_info=False)
    eng.pack_observations(store, 0)
    fresh = eng.render_observations(store, torch.arange(n, device=DEV))
    assert float(fresh[:, 2].abs().sum()) == 0.0 and torch.equal(fresh, torch.cat([eng.volumes()[:, :2], fresh[:, 2:3], eng.volumes()[:, 3:]], 1))
    eng.close()


def test_train_vision_tiny():
    """Vision-mode PPO with compressed rollout storage and re-rendered micro-batches runs and stays finite."""
    cfg = load_config(None, num_envs=6, num_steps=10, num_minibatches=2, update_epochs=1, visionless=False,
                      total_timesteps=6 * 10 * 2, num_saves=0, save_model=False, seed=5, render_microbatch=8)
    torch.manual_seed(0)
    agent = train(cfg, writer=None, device=DEV, output_dir=None, run_name="v", log=None)
    h = agent.history
    assert len(h) == 2 and all(np.isfinite(r["v_loss"]) and np.isfinite(r["pg_loss"]) for r in h)


def test_evaluate_like_ppo_eval(tmp_path):
    """ppo_eval.py:5-36 on the device-resident env: loads a reference-format checkpoint, returns one mean return per
    batch of finished episodes."""
    from ppo_radiotherapy_b200.ppo_eval import evaluate
    envs = rt.RadiotherapyVectorEnv(32, visionless=True, device=DEV, seed=4)
    torch.manual_seed(1)
    model = rt.PPO(envs.single_observation_space.shape, envs.single_action_space.shape, 64)
    path = tmp_path / "m.model"
    torch.save(model.state_dict(), path)
    out = evaluate(envs, 64, str(path), eval_episodes=2, Model=rt.PPO, device=torch.device(DEV), log=None)
    assert len(out) == 2 and all(np.isfinite(r) and -200 < r < 200 for r in out)
    envs.close()


def test_fused_rollout_policy_matches_torch():
    """rt_ppo_act against `PPO.get_action_and_value` (networks.py:132-147) on the same parameters: value and
    log-probability of the kernel's own action in float32 tolerance, rollout rows written, the noise standard normal
    and fresh every step; a ragged env count (not a multiple of the 64-env tile)."""
    torch.manual_seed(1)
    dev = torch.device(DEV)
    agent = rt.PPO((9,), (6,), 64).to(dev)
    with torch.no_grad():
        agent.actor_logstd.copy_(torch.linspace(-0.7, 0.4, 6, device=dev).reshape(1, 6))
        agent.actor_mean[4].weight.mul_(30.0)
        agent.actor_mean[4].bias.uniform_(-0.3, 0.3)
        agent.critic[4].bias.fill_(0.25)
    n, T = 4099, 3
    fr = rt.FusedRollout(agent, n, T, seed=5)
    zs = []
    for t in range(T):
        obs = torch.rand((n, 9), device=dev) * 2 - 1
        fr.next_done.copy_((torch.rand(n, device=dev) < 0.1).float())
        a = fr.act(obs).clone()
        with torch.no_grad():
            mean = agent.actor_mean(obs)
            _, lp, _, v = agent.get_action_and_value(obs, a)
        assert torch.equal(fr.obs[t], obs) and torch.equal(fr.actions[t], a) and torch.equal(fr.dones[t], fr.next_done)
        torch.testing.assert_close(fr.values[t], v.flatten(), rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(fr.logprobs[t], lp, rtol=1e-5, atol=2e-4)
        zs.append(((a - mean) / torch.exp(agent.actor_logstd)).detach().cpu().numpy())
        fr.advance()
    assert fr.counters.cpu().tolist() == [T, T]
    z = np.stack(zs)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs(np.mean(z ** 3)) < 0.05 and abs(np.mean(z ** 4) - 3.0) < 0.15            # skewness, kurtosis of N(0,1)
    assert np.abs(np.corrcoef(z[0].ravel(), z[1].ravel())[0, 1]) < 0.02                 # fresh noise every step
    assert np.abs(np.corrcoef(z[0][:, 0], z[0][:, 1])[0, 1]) < 0.05                     # ... and per action component
    # the same (seed, env, step) gives the same noise
    fr2 = rt.FusedRollout(agent, n, T, seed=5)
    obs = torch.zeros((n, 9), device=dev)
    assert torch.equal(fr2.act(obs), rt.FusedRollout(agent, n, T, seed=5).act(obs))


def test_fused_rollout_record_and_train():
    """rt_ppo_record against the tensor expressions of the PyTorch path, then the training loop with the fused rollout
    step (CUDA graph) and without it: same env trajectories up to the sampled actions, finite losses, same statistics."""
    from ppo_radiotherapy_b200 import _native as nat
    dev = torch.device(DEV)
    n = 300
    eng = rt.BatchedEpisodes(n, device=dev, seed=2)
    eng.reset()
    agent = rt.PPO((9,), (6,), 64).to(dev)
    fr = rt.FusedRollout(agent, n, 8, seed=1)
    eng.step(torch.rand((n, 6), device=dev) * 2 - 1, want_info=True)
    eng.terminated[::7] = 1                                                   # pretend some envs finished
    fr.counters[0] = 3
    fr.record(eng)
    torch.cuda.synchronize()
    assert torch.equal(fr.rewards[3], eng.reward_f32) and fr.rewards[:3].abs().sum() == 0
    assert torch.equal(fr.next_done, eng.terminated.float())
    f64 = eng.terminated.double()
    cols = [nat.INFO_EPISODE_RETURN, nat.INFO_EPISODE_LENGTH, nat.INFO_REWARD_TUMOUR, nat.INFO_REWARD_LUNG,
            nat.INFO_REWARD_DISTANCE, nat.INFO_REWARD_TOTAL]
    want = torch.cat([f64.sum().reshape(1), (eng.info[:, cols] * f64.unsqueeze(1)).sum(0)])
    torch.testing.assert_close(fr.episode_stats, want, rtol=1e-12, atol=1e-12)
    eng.close()
    hist = {}
    for fused in (True, False):
        cfg = load_config(None, num_envs=128, num_steps=100, num_minibatches=4, update_epochs=2,
                          total_timesteps=128 * 100 * 2, num_saves=0, save_model=False, seed=4, fused_rollout=fused)
        torch.manual_seed(0)
        hist[fused] = train(cfg, writer=None, device=DEV, output_dir=None, run_name="f", log=None).history
    for fused in (True, False):
        h = hist[fused]
        assert len(h) == 2 and all(np.isfinite(r["v_loss"]) and np.isfinite(r["pg_loss"]) for r in h)
        assert h[0]["episodes"] == 128 and abs(h[0]["episodic_length"] - 100) < 1e-9
        assert -200 < h[0]["episodic_return"] < 200
    # first-iteration policy is the same initial network: approx_kl of the first minibatches is tiny in both paths
    assert abs(hist[True][0]["approx_kl"]) < 0.05 and abs(hist[False][0]["approx_kl"]) < 0.05


@pytest.mark.parametrize("n_obs,n_act,n", [(12, 4, 130), (3, 1, 64), (16, 6, 1)])
def test_fused_rollout_other_shapes(n_obs, n_act, n):
    """rt_ppo_act for observation / action sizes other than the radiotherapy env's (9, 6), env counts below and at a tile."""
    torch.manual_seed(2)
    dev = torch.device(DEV)
    agent = rt.PPO((n_obs,), (n_act,), 64).to(dev)
    with torch.no_grad():
        agent.actor_logstd.uniform_(-0.5, 0.5)
        agent.actor_mean[4].weight.mul_(20.0)
    fr = rt.FusedRollout(agent, n, 2, seed=9)
    obs = torch.randn((n, n_obs), device=dev)
    a = fr.act(obs).clone()
    with torch.no_grad():
        _, lp, _, v = agent.get_action_and_value(obs, a)
    torch.testing.assert_close(fr.values[0], v.flatten(), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(fr.logprobs[0], lp, rtol=1e-5, atol=2e-4)
    assert torch.equal(fr.actions[0], a) and torch.equal(fr.obs[0], obs) and a.shape == (n, n_act)


@pytest.mark.parametrize("n", [61, 1200])
def test_rollout_kernel_equals_per_step_kernels(n):
    """rt_rollout (all T steps in one launch, every block keeps its envs) against T x (rt_ppo_act, rt_step, rt_ppo_record)
    on the same Philox stream: observation, done, value, action, log-prob and reward rows bit for bit, through the end of an
    episode, the autoreset call and the start of the next one; same final env state; same episode statistics (atomics: to
    rounding).  Both block shapes (7 and 14 envs per block), ragged last block, two chunks (rows 0..59, 60..T-1)."""
    dev = torch.device(DEV)
    T = 110
    torch.manual_seed(5)
    agent = rt.PPO((9,), (6,), 64).to(dev)
    with torch.no_grad():
        agent.actor_logstd.copy_(torch.linspace(-0.6, 0.3, 6).reshape(1, 6))
    sched = np.stack([(np.arange(n) * 7919) % 1000, (np.arange(n) * 104729 + 17) % 1000]).astype(np.int32)
    a = rt.BatchedEpisodes(n, device=dev); a.set_tumour_schedule(sched); a.reset()
    b = rt.BatchedEpisodes(n, device=dev); b.set_tumour_schedule(sched); b.reset()
    fa = rt.FusedRollout(agent, n, T, seed=77)
    fb = rt.FusedRollout(agent, n, T, seed=77)
    for t in range(T):                                   # per-step kernels (train.py:138-161 one step at a time)
        fa.act(a.obs)
        a.step(fa.action, want_info=True)
        fa.record(a)
        fa.advance()
    fb.rollout(b, 60)                                    # one launch for rows 0..59, one for the rest
    fb.rollout(b)
    torch.cuda.synchronize()
    for name in ("obs", "dones", "values", "actions", "logprobs", "rewards"):
        x, y = getattr(fa, name), getattr(fb, name)
        assert torch.equal(x.view(torch.int32), y.view(torch.int32)), name
    assert torch.equal(fa.next_done, fb.next_done) and torch.equal(a.obs, b.obs)
    assert torch.equal(a.pose(), b.pose()) and torch.equal(a.counters(), b.counters())
    assert torch.equal(a.dose(n - 1), b.dose(n - 1))
    assert fa.episode_stats[0] == n and fb.episode_stats[0] == n           # every env finished one episode (t = 100)
    torch.testing.assert_close(fa.episode_stats, fb.episode_stats, rtol=1e-12, atol=1e-9)
    assert fb._row == T and int(fb.counters[0]) == T and int(fb.counters[1]) == T
    with pytest.raises(nat_err()):
        fb.rollout(b, 1)                                 # past the end of the rollout buffers
    a.close(); b.close()


def nat_err():
    from ppo_radiotherapy_b200 import _native as nat
    return nat.RtError


def test_act_and_record_past_the_last_row_store_nothing():
    """rt_ppo_act / rt_ppo_record take the row count of the rollout buffers; a call whose device-side row
    counter is past the end still produces actions but writes no row."""
    dev = torch.device(DEV)
    n, T = 70, 3
    agent = rt.PPO((9,), (6,), 64).to(dev)
    eng = rt.BatchedEpisodes(n, device=dev, seed=1); eng.reset()
    fr = rt.FusedRollout(agent, n, T, seed=3)
    guard = torch.full((4, n), 7.0, device=dev)
    for t in range(T + 2):
        before = [x.clone() for x in (fr.obs, fr.dones, fr.values, fr.actions, fr.logprobs, fr.rewards)]
        fr.act(eng.obs)
        eng.step(fr.action, want_info=True)
        fr.record(eng)
        fr.advance()
        torch.cuda.synchronize()
        if t >= T:
            for x, y in zip(before, (fr.obs, fr.dones, fr.values, fr.actions, fr.logprobs, fr.rewards)):
                assert torch.equal(x, y)
            assert torch.isfinite(fr.action).all()
    assert (guard == 7.0).all()
    eng.close()
