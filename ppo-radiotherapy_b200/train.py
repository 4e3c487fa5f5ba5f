"""CleanRL-style PPO on the device-resident vector env — the reference's train.py:91-282 with the
host round trips removed.

    python -m torch.distributed.run --nproc-per-node 8 ppo_radiotherapy_b200_train.py --config-file cfg.yaml --output-dir out
    python ppo_radiotherapy_b200_train.py --config-file cfg.yaml --output-dir out      (single GPU)

(`ppo_radiotherapy_b200_train.py` at the repo root is the launcher: the package directory has a hyphen in its name, so
`python -m` cannot address this module.)

What stays as in the reference: config field names (configs/default.yaml.template), the agent classes and
their checkpoint layout, the loss (clipped surrogate, clipped value loss, per-minibatch advantage
normalisation, Adam eps 1e-5, grad-norm clip 0.5), the TensorBoard tags, `agent.state_dict()` checkpoints.
What changes: observations, actions, rewards and dones never leave the GPU (train.py:151-158 did two
copies per step); GAE is one kernel (rt_gae) instead of 2048 tiny launches (train.py:164-181); episode
statistics are reduced on the device and read once per iteration; under torchrun every rank owns
num_envs / world_size envs and gradients are averaged with ONE NCCL all-reduce of the flattened
gradient per optimiser step (between backward and clip_grad_norm_, i.e. train.py:246-247).
Advantage normalisation is per rank and per minibatch (train.py:215-218 semantics on the local shard).
"""
import argparse
import os
import random
import time
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim

from . import _native as nat
from .geometry import compute_gae
from .networks import PPO, PPO_3DCNN
from .rollout import FusedRollout, supported as fused_rollout_supported
from .vector_env import RadiotherapyVectorEnv

DEFAULTS = dict(
    exp_name="default", seed=1, torch_deterministic=True, cuda=True, save_model=True, use_tqdm=False,
    total_timesteps=10_000_000, num_saves=5, learning_rate=3e-4, num_envs=16, num_steps=2048, anneal_lr=False,
    num_minibatches=32, update_epochs=10, gamma=0.99, gae_lambda=0.95, norm_adv=True, clip_coef=0.1,
    clip_vloss=True, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, feature_dim=64, visionless=True,
    cuda_graph=True,     # replay one captured rollout step (policy + env + buffer writes) instead of ~50 launches
    fused_rollout=True,  # MLP agent: rt_ppo_act + rt_step + rt_ppo_record per rollout step (rollout.py) instead of PyTorch ops
    rollout_kernel=True, # ... and all num_steps steps of an iteration in ONE launch (rt_rollout: blocks keep their envs, no grid barrier per step)
    render_microbatch=256,   # vision mode: samples re-rendered from compressed records per gradient micro-batch
)


def load_config(path: Optional[str] = None, **overrides) -> SimpleNamespace:
    """YAML -> namespace with the derived sizes of train.py:292-297."""
    cfg = dict(DEFAULTS)
    if path:
        import yaml
        with open(path) as f:
            cfg.update(yaml.safe_load(f) or {})
    cfg.update(overrides)
    cfg = SimpleNamespace(**cfg)
    finalize_config(cfg)
    return cfg


def finalize_config(cfg: SimpleNamespace, world_size: int = 1) -> SimpleNamespace:
    cfg.batch_size = int(cfg.num_envs * cfg.num_steps)
    cfg.minibatch_size = int(cfg.batch_size // cfg.num_minibatches)
    cfg.num_iterations = int(cfg.total_timesteps // cfg.batch_size)
    cfg.save_frequency_iterations = cfg.num_iterations // cfg.num_saves if cfg.num_saves > 0 else 0
    if cfg.num_envs % world_size:
        raise ValueError(f"num_envs={cfg.num_envs} is not divisible by world_size={world_size}")
    return cfg


# --------------------------------------------------------------------------------------------
# multi-GPU plumbing (exercised on CPU with gloo in tests/test_distributed_cpu.py)
def shard_range(total: int, world_size: int, rank: int):
    """Contiguous env shard of a rank: env i lives on rank i // (total / world_size)."""
    if total % world_size:
        raise ValueError("total must be divisible by world_size")
    per = total // world_size
    return rank * per, (rank + 1) * per


class FlatGradAllReduce:
    """One all-reduce of the flattened gradient per optimiser step, averaged over ranks."""

    def __init__(self, params, group=None, events: Optional[list] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.events = events       # profiling: (start, end) CUDA event pairs around every all_reduce call
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self):
        if self.world == 1:
            return
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if self.events is not None and self.flat.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            e1.record()
            self.events.append((e0, e1))
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(self.world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.flat[off:off + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(self.flat[off:off + n].reshape(p.shape))
            off += n


def broadcast_parameters(module: nn.Module, src: int = 0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


# --------------------------------------------------------------------------------------------
def ppo_update_rendered(agent, optimizer, cfg, render, b_actions, b_logprobs, b_advantages, b_returns, b_values,
                        sync_grads=None, generator: Optional[torch.Generator] = None):
    """The same update (train.py:191-248) when observations are rendered on demand: a minibatch is processed in
    micro-batches of cfg.render_microbatch samples whose gradients are accumulated (the losses are means over the
    minibatch, so each micro-batch is weighted by its share of it; the advantage normalisation uses the statistics
    of the whole minibatch like the reference)."""
    n = b_actions.shape[0]
    dev = b_actions.device
    mb = max(1, n // cfg.num_minibatches)
    micro = max(1, min(int(getattr(cfg, "render_microbatch", 256)), mb))
    clipfracs = []
    stats = {}
    for _ in range(cfg.update_epochs):
        perm = torch.randperm(n, device=dev, generator=generator)
        for start in range(0, n - mb + 1, mb):
            idx = perm[start:start + mb]
            adv_all = b_advantages[idx]
            if cfg.norm_adv:
                adv_all = (adv_all - adv_all.mean()) / (adv_all.std() + 1e-8)
            optimizer.zero_grad()
            acc = torch.zeros(6, device=dev)          # pg, v, entropy, old_kl, kl, clipfrac (minibatch means)
            for m0 in range(0, mb, micro):
                sub = idx[m0:m0 + micro].contiguous()
                share = sub.numel() / mb
                obs = render(sub)
                _, newlogprob, entropy, newvalue = agent.get_action_and_value(obs, b_actions[sub])
                logratio = newlogprob - b_logprobs[sub]
                ratio = logratio.exp()
                adv = adv_all[m0:m0 + micro]
                pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - cfg.clip_coef, 1 + cfg.clip_coef)).mean()
                newvalue = newvalue.view(-1)
                if cfg.clip_vloss:
                    v_unclipped = (newvalue - b_returns[sub]) ** 2
                    v_clipped = b_values[sub] + torch.clamp(newvalue - b_values[sub], -cfg.clip_coef, cfg.clip_coef)
                    v_loss = 0.5 * torch.max(v_unclipped, (v_clipped - b_returns[sub]) ** 2).mean()
                else:
                    v_loss = 0.5 * ((newvalue - b_returns[sub]) ** 2).mean()
                entropy_loss = entropy.mean()
                loss = (pg_loss - cfg.ent_coef * entropy_loss + v_loss * cfg.vf_coef) * share
                loss.backward()
                with torch.no_grad():
                    acc += share * torch.stack([pg_loss, v_loss, entropy_loss, (-logratio).mean(),
                                                ((ratio - 1) - logratio).mean(),
                                                ((ratio - 1.0).abs() > cfg.clip_coef).float().mean()])
                del obs
            if sync_grads is not None:
                sync_grads()
            nn.utils.clip_grad_norm_(agent.parameters(), cfg.max_grad_norm)
            optimizer.step()
            clipfracs.append(acc[5])
            stats = dict(pg_loss=acc[0], v_loss=acc[1], entropy=acc[2], old_approx_kl=acc[3], approx_kl=acc[4])
    stats["clipfrac"] = torch.stack(clipfracs).mean() if clipfracs else torch.zeros((), device=dev)
    return stats


def ppo_update(agent, optimizer, cfg, b_obs, b_actions, b_logprobs, b_advantages, b_returns, b_values,
               sync_grads=None, generator: Optional[torch.Generator] = None):
    """train.py:191-248 on device tensors.  Returns the last minibatch's statistics like the reference.
    `b_obs` is the flattened observation tensor, or a callable idx -> observations (voxel observations are
    re-rendered from compressed records, see ppo_update_rendered)."""
    if callable(b_obs):
        return ppo_update_rendered(agent, optimizer, cfg, b_obs, b_actions, b_logprobs, b_advantages, b_returns,
                                   b_values, sync_grads, generator)
    n = b_obs.shape[0]
    mb = max(1, n // cfg.num_minibatches)
    clipfracs = []
    stats = {}
    for _ in range(cfg.update_epochs):
        perm = torch.randperm(n, device=b_obs.device, generator=generator)
        for start in range(0, n - mb + 1, mb):
            idx = perm[start:start + mb]
            _, newlogprob, entropy, newvalue = agent.get_action_and_value(b_obs[idx], b_actions[idx])
            logratio = newlogprob - b_logprobs[idx]
            ratio = logratio.exp()
            with torch.no_grad():
                old_approx_kl = (-logratio).mean()
                approx_kl = ((ratio - 1) - logratio).mean()
                clipfracs.append(((ratio - 1.0).abs() > cfg.clip_coef).float().mean())
            adv = b_advantages[idx]
            if cfg.norm_adv:
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
            pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - cfg.clip_coef, 1 + cfg.clip_coef)).mean()
            newvalue = newvalue.view(-1)
            if cfg.clip_vloss:
                v_unclipped = (newvalue - b_returns[idx]) ** 2
                v_clipped = b_values[idx] + torch.clamp(newvalue - b_values[idx], -cfg.clip_coef, cfg.clip_coef)
                v_loss = 0.5 * torch.max(v_unclipped, (v_clipped - b_returns[idx]) ** 2).mean()
            else:
                v_loss = 0.5 * ((newvalue - b_returns[idx]) ** 2).mean()
            entropy_loss = entropy.mean()
            loss = pg_loss - cfg.ent_coef * entropy_loss + v_loss * cfg.vf_coef
            optimizer.zero_grad()
            loss.backward()
            if sync_grads is not None:
                sync_grads()
            nn.utils.clip_grad_norm_(agent.parameters(), cfg.max_grad_norm)
            optimizer.step()
            stats = dict(pg_loss=pg_loss.detach(), v_loss=v_loss.detach(), entropy=entropy_loss.detach(),
                         old_approx_kl=old_approx_kl, approx_kl=approx_kl)
    stats["clipfrac"] = torch.stack(clipfracs).mean() if clipfracs else torch.zeros((), device=b_obs.device)
    return stats


def train(cfg, writer=None, device="cuda", output_dir: Optional[str] = None, run_name: str = "run", log=print,
          profile: Optional[dict] = None):
    """The reference's train(cfg, writer, device) (train.py:91) on one rank's env shard.
    `profile` (a dict) asks for CUDA-event timings per iteration: it receives `rollout_ms`, `update_ms`, `iter_ms`
    (lists, one entry per iteration) and `allreduce_us` (one entry per NCCL all-reduce call)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    finalize_config(cfg, world)
    device = torch.device(device)
    lo, hi = shard_range(cfg.num_envs, world, rank)
    n_local = hi - lo

    envs = RadiotherapyVectorEnv(n_local, visionless=cfg.visionless, device=device, seed=cfg.seed * 1_000_003 + rank)
    obs_shape = envs.single_observation_space.shape
    act_shape = envs.single_action_space.shape
    if cfg.visionless:
        agent = PPO(obs_shape, act_shape, cfg.feature_dim).to(device)
    else:
        agent = PPO_3DCNN(obs_shape, act_shape, cfg.feature_dim, compute_dtype=torch.bfloat16).to(device)
    broadcast_parameters(agent)
    optimizer = optim.Adam(agent.parameters(), lr=cfg.learning_rate, eps=1e-5)
    ar_events = [] if profile is not None else None
    sync_grads = FlatGradAllReduce(agent.parameters(), events=ar_events) if world > 1 else None
    iter_events = []

    T = cfg.num_steps
    # visionless: the rollout keeps the observations (train.py:110).  Vision: a float32 voxel observation is 3.2 MB
    # (422 GB for 1024 envs x 128 steps), so the rollout keeps compressed records (bfloat16 dose + pose + tumour
    # id, 1/8 of that) and the update re-renders each micro-batch.
    store = None if cfg.visionless else envs.engine.observation_store(T * n_local)
    obs = torch.zeros((T, n_local) + obs_shape, device=device) if cfg.visionless else None
    actions = torch.zeros((T, n_local) + act_shape, device=device)
    logprobs = torch.zeros((T, n_local), device=device)
    rewards = torch.zeros((T, n_local), device=device)
    dones = torch.zeros((T, n_local), device=device)
    values = torch.zeros((T, n_local), device=device)

    eng = envs.engine
    # visionless MLP agent: the whole rollout step is three hand-written kernels (policy forward + sampling + buffer
    # writes, env step, reward / done / episode statistics); the rollout buffers are the FusedRollout's own
    fused = None
    if cfg.visionless and getattr(cfg, "fused_rollout", True) and fused_rollout_supported(agent):
        fused = FusedRollout(agent, n_local, T, seed=cfg.seed * 7919 + rank)
        obs, actions, logprobs, rewards, dones, values = (fused.obs, fused.actions, fused.logprobs, fused.rewards,
                                                          fused.dones, fused.values)
    if cfg.visionless:
        next_obs, _ = envs.reset(seed=None, options={"backend": "torch"})
        next_obs = next_obs.clone()
    else:
        # vision mode never materialises the float32 observation during the rollout: the policy reads the env state
        # through rt_conv1_from_env, the rollout buffer keeps compressed records
        eng.reset()
        next_obs = None
    next_done = torch.zeros(n_local, device=device)
    step_idx = torch.zeros(1, dtype=torch.long, device=device)
    if fused is not None:
        next_obs, next_done, step_idx = eng.obs, fused.next_done, fused.counters[:1]     # views of the kernels' own state
    # per-iteration episode statistics, reduced on the device: [finished, sum return, sum length,
    # sum of last-step reward components (tumour, lung, distance, total)] (train.py:42-66)
    ep = torch.zeros(7, dtype=torch.float64, device=device) if fused is None else fused.episode_stats
    ep_cols = torch.tensor([nat.INFO_EPISODE_RETURN, nat.INFO_EPISODE_LENGTH, nat.INFO_REWARD_TUMOUR,
                            nat.INFO_REWARD_LUNG, nat.INFO_REWARD_DISTANCE, nat.INFO_REWARD_TOTAL], device=device)

    rollout_pos = [0]

    def rollout_step():
        """train.py:139-158 for the step selected by the device-side counter `step_idx` (so that one captured
        CUDA graph serves every step of the rollout)."""
        if fused is not None:
            fused.act(eng.obs)                                                   # train.py:139-149
            eng.step(fused.action, want_info=True)                               # train.py:151
            fused.record(eng)                                                    # train.py:153-161
            fused.advance()
            return
        if cfg.visionless:
            obs.index_copy_(0, step_idx, next_obs.unsqueeze(0))
        else:                                          # eager only (no CUDA graph in vision mode): host-side slot index
            eng.pack_observations(store, rollout_pos[0] * n_local)
            rollout_pos[0] += 1
        dones.index_copy_(0, step_idx, next_done.unsqueeze(0))
        with torch.no_grad():
            if cfg.visionless:
                action, logprob, _, value = agent.get_action_and_value(next_obs)
            else:
                action, logprob, _, value = agent.get_action_and_value_from_env(eng)
        values.index_copy_(0, step_idx, value.reshape(1, -1))
        actions.index_copy_(0, step_idx, action.unsqueeze(0))
        logprobs.index_copy_(0, step_idx, logprob.unsqueeze(0))
        o, _, term, trunc, info = eng.step(action, want_info=True)              # train.py:151, on the device
        if cfg.visionless:
            next_obs.copy_(o)
        rewards.index_copy_(0, step_idx, eng.reward_f32.unsqueeze(0))
        next_done.copy_((term | trunc).float())
        f64 = term.double()
        ep[0:1] += f64.sum()
        ep[1:7] += (info.index_select(1, ep_cols) * f64.unsqueeze(1)).sum(0)
        step_idx.add_(1)

    # one launch for the whole rollout while the envs fit two waves of blocks (measured on B200: 1.6x the per-step kernels
    # at 4,096 envs, 1.1x at 8,192, slower from ~12,000 envs on, where the per-step launches fill the GPU anyway)
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    one_launch = fused is not None and getattr(cfg, "rollout_kernel", True) and n_local <= 56 * sms
    graph = None
    if getattr(cfg, "cuda_graph", True) and cfg.visionless and not one_launch:
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(min(3, T)):              # warm-up on the side stream (allocator, cuBLAS handles)
                rollout_step()
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                rollout_step()
        torch.cuda.current_stream(device).wait_stream(side)
        # the warm-up steps advanced the envs; start the run from a clean reset
        envs.engine.reset()
        next_obs.copy_(eng.obs)
        next_done.zero_()
        ep.zero_()

    global_step = 0
    start = time.time()
    history = []

    for iteration in range(1, cfg.num_iterations + 1):
        iter_start = time.time()
        if cfg.anneal_lr:
            optimizer.param_groups[0]["lr"] = (1.0 - (iteration - 1.0) / cfg.num_iterations) * cfg.learning_rate
        if fused is not None:
            fused.begin_iteration()                 # row counter (device and host mirror) and episode statistics to zero
        else:
            ep.zero_()
            step_idx.zero_()
        rollout_pos[0] = 0
        if profile is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        if one_launch:
            fused.rollout(eng, T)                   # train.py:138-161, all T steps in one kernel
            global_step += cfg.num_envs * T
        else:
            for step in range(T):
                global_step += cfg.num_envs
                if graph is not None:
                    graph.replay()
                else:
                    rollout_step()

        with torch.no_grad():
            if cfg.visionless:
                next_value = agent.get_value(next_obs).reshape(-1)
            else:
                next_value = agent.critic(agent.features_extractor.forward_from_env(eng)).reshape(-1)
            advantages, returns = compute_gae(rewards, values, dones, next_value, next_done, cfg.gamma, cfg.gae_lambda)
        if profile is not None:
            ev[1].record()

        b_obs = obs.reshape((-1,) + obs_shape) if cfg.visionless else (lambda idx: eng.render_observations(store, idx))
        stats = ppo_update(agent, optimizer, cfg,
                           b_obs, actions.reshape((-1,) + act_shape), logprobs.reshape(-1),
                           advantages.reshape(-1), returns.reshape(-1), values.reshape(-1), sync_grads)
        if profile is not None:
            ev[2].record()
            iter_events.append(ev)

        if world > 1:
            dist.all_reduce(ep, op=dist.ReduceOp.SUM)
        epc = ep.cpu().numpy()
        y_pred, y_true = values.reshape(-1), returns.reshape(-1)
        var_y = torch.var(y_true)
        explained = float("nan") if float(var_y) == 0 else float(1 - torch.var(y_true - y_pred) / var_y)
        rec = dict(iteration=iteration, global_step=global_step, sps=global_step / max(time.time() - start, 1e-9),
                   iter_sps=cfg.num_envs * T / max(time.time() - iter_start, 1e-9),   # this iteration alone (ep.cpu() above synchronised)
                   episodes=int(epc[0]), explained_variance=explained,
                   **{k: float(v) for k, v in stats.items()})
        if epc[0] > 0:
            rec.update(episodic_return=epc[1] / epc[0], episodic_length=epc[2] / epc[0],
                       episodic_tumour_reward=epc[3] / epc[0], episodic_lung_reward=epc[4] / epc[0],
                       episodic_distance_reward=epc[5] / epc[0], episodic_total_reward=epc[6] / epc[0])
        history.append(rec)
        if rank == 0:
            if writer is not None:
                for k in ("episodic_return", "episodic_length", "episodic_tumour_reward", "episodic_lung_reward",
                          "episodic_distance_reward", "episodic_total_reward"):
                    if k in rec:
                        writer.add_scalar(f"charts/{k}", rec[k], global_step)
                writer.add_scalar("charts/learning_rate", optimizer.param_groups[0]["lr"], global_step)
                for tag, k in (("value_loss", "v_loss"), ("policy_loss", "pg_loss"), ("entropy", "entropy"),
                               ("old_approx_kl", "old_approx_kl"), ("approx_kl", "approx_kl"), ("clipfrac", "clipfrac")):
                    writer.add_scalar(f"losses/{tag}", rec[k], global_step)
                writer.add_scalar("losses/explained_variance", explained, global_step)
            if log is not None:
                log(f"iter {iteration}/{cfg.num_iterations} step {global_step} sps {rec['sps']:.0f} "
                    f"return {rec.get('episodic_return', float('nan')):.2f} v_loss {rec['v_loss']:.4f}")
            if cfg.save_model and output_dir and (
                    (cfg.save_frequency_iterations and iteration % cfg.save_frequency_iterations == 0)
                    or iteration == cfg.num_iterations):
                d = os.path.join(output_dir, "models", run_name)
                os.makedirs(d, exist_ok=True)
                path = os.path.join(d, f"{cfg.exp_name}_{iteration}.model")
                torch.save(agent.state_dict(), path)                           # train.py:270-279
                if log is not None:
                    log(f"model saved to {path}")
    if profile is not None:
        torch.cuda.synchronize(device)
        profile["rollout_ms"] = [e[0].elapsed_time(e[1]) for e in iter_events]      # rollout + GAE
        profile["update_ms"] = [e[1].elapsed_time(e[2]) for e in iter_events]
        profile["iter_ms"] = [e[0].elapsed_time(e[2]) for e in iter_events]
        profile["allreduce_us"] = [a.elapsed_time(b) * 1e3 for a, b in (ar_events or [])]
        profile["fused_rollout"] = fused is not None
        profile["rollout_kernel"] = one_launch
    envs.close()
    agent.history = history
    return agent


def main(argv=None):
    ap = argparse.ArgumentParser(description="PPO on the B200 radiotherapy vector env")
    ap.add_argument("--config-file", type=str, default=None, help="YAML with the reference's field names")
    ap.add_argument("--output-dir", type=str, default="runs_b200")
    ap.add_argument("--set", nargs="*", default=[], help="overrides key=value")
    args = ap.parse_args(argv)
    over = {}
    for kv in args.set:
        k, v = kv.split("=", 1)
        over[k] = type(DEFAULTS[k])(v) if k in DEFAULTS and not isinstance(DEFAULTS[k], bool) else (v.lower() == "true")
    cfg = load_config(args.config_file, **over)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("training needs a CUDA device (the environment step has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    random.seed(cfg.seed + rank)
    np.random.seed(cfg.seed + rank)
    torch.manual_seed(cfg.seed + rank)
    torch.backends.cudnn.deterministic = cfg.torch_deterministic
    run_name = f"{cfg.exp_name}_{int(time.time())}"
    writer = None
    if rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(os.path.join(args.output_dir, "tensorboard", run_name))
        except Exception:
            writer = None
    train(cfg, writer, device, args.output_dir, run_name)
    if writer is not None:
        writer.close()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
