"""Fused rollout step of the PPO loop (train.py:139-161 of the reference) for the MLP agent `PPO`:

    rollout.act(obs, next_done)        # obs/dones rows, policy + value forward, Gaussian sample, log-prob, values/actions/
                                       # logprobs rows -> rollout.action             (rt_ppo_act, one kernel)
    engine.step(rollout.action)        # the environment step                        (rt_step)
    rollout.record(engine)             # rewards row, next_done, episode statistics  (rt_ppo_record, one kernel)
    rollout.advance()                  # device-side row / RNG counters

replaces the ~45 small PyTorch launches of `agent.get_action_and_value` + buffer writes per step, and

    rollout.rollout(engine, T)         # all T steps of an iteration in one launch                   (rt_rollout)

runs the whole loop inside one kernel in which every block keeps its envs for all T steps (same rows, bit for bit).  The parameters are
read in place from the agent's tensors (training keeps updating them through PyTorch); there is no CPU path.
"""
import ctypes as C
from typing import Optional

import torch

from . import _native as nat
from .networks import PPO


def supported(agent) -> bool:
    """The kernel covers the reference's MLP agent with feature_dim 64 (configs/default.yaml.template)."""
    if type(agent) is not PPO:
        return False
    try:
        c0, a0, a2 = agent.critic[0], agent.actor_mean[0], agent.actor_mean[4]
    except (IndexError, TypeError):
        return False
    return (c0.out_features == 64 and a0.out_features == 64 and agent.critic[2].out_features == 64
            and agent.actor_mean[2].out_features == 64 and c0.in_features <= 16 and a2.out_features <= 6
            and all(p.dtype == torch.float32 and p.is_cuda and p.is_contiguous() for p in agent.parameters()))


class FusedRollout:
    def __init__(self, agent: PPO, num_envs: int, num_steps: int, seed: int = 0):
        if not supported(agent):
            raise nat.RtError("FusedRollout: the agent is not the reference's MLP `PPO` with feature_dim 64 on a CUDA device")
        self.agent = agent
        self.device = next(agent.parameters()).device
        self.n, self.T = int(num_envs), int(num_steps)
        self.n_obs, self.n_act = agent.critic[0].in_features, agent.actor_mean[4].out_features
        self.seed = int(seed) & (2 ** 64 - 1)
        d = self.device
        self.counters = torch.zeros(2, dtype=torch.int64, device=d)           # [rollout row, RNG step]
        self._row, self._rng = 0, 0                                           # host mirrors of the two counters
        self.action = torch.zeros((self.n, self.n_act), dtype=torch.float32, device=d)
        self.obs = torch.zeros((self.T, self.n, self.n_obs), dtype=torch.float32, device=d)
        self.actions = torch.zeros((self.T, self.n, self.n_act), dtype=torch.float32, device=d)
        self.logprobs = torch.zeros((self.T, self.n), dtype=torch.float32, device=d)
        self.rewards = torch.zeros((self.T, self.n), dtype=torch.float32, device=d)
        self.dones = torch.zeros((self.T, self.n), dtype=torch.float32, device=d)
        self.values = torch.zeros((self.T, self.n), dtype=torch.float32, device=d)
        self.next_done = torch.zeros(self.n, dtype=torch.float32, device=d)
        self.episode_stats = torch.zeros(7, dtype=torch.float64, device=d)
        self._params = self._bind()

    def _bind(self) -> nat.MlpParams:
        a = self.agent
        p = nat.MlpParams()
        for prefix, head in (("critic", a.critic), ("actor", a.actor_mean)):
            for i, layer in enumerate((head[0], head[2], head[4])):
                setattr(p, f"{prefix}_w{i}", layer.weight.data_ptr())
                setattr(p, f"{prefix}_b{i}", layer.bias.data_ptr())
        p.actor_logstd = a.actor_logstd.data_ptr()
        p.n_obs, p.hidden, p.n_act = self.n_obs, 64, self.n_act
        return p

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def act(self, obs: torch.Tensor, next_done: Optional[torch.Tensor] = None, store: bool = True) -> torch.Tensor:
        """train.py:139-149 for the row selected by the device-side counter; returns this step's actions [N][n_act]."""
        nd = self.next_done if next_done is None else next_done
        if obs.shape != (self.n, self.n_obs) or obs.dtype != torch.float32 or not obs.is_contiguous() or obs.device != self.device:
            raise ValueError("FusedRollout.act: obs must be a contiguous float32 [N][n_obs] tensor on the agent's device")
        z = C.c_void_p(0)
        buf = (lambda t: C.c_void_p(t.data_ptr())) if store else (lambda t: z)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().rt_ppo_act(C.byref(self._params), C.c_void_p(obs.data_ptr()), C.c_void_p(nd.data_ptr()), self.n,
                                           C.c_uint64(self.seed), C.c_void_p(self.counters.data_ptr()), self.T, buf(self.obs),
                                           buf(self.dones), buf(self.values), buf(self.actions), buf(self.logprobs),
                                           C.c_void_p(self.action.data_ptr()), self._stream()), "rt_ppo_act")
        return self.action

    def record(self, engine, store: bool = True):
        """train.py:153-161 after engine.step(..., want_info=True)."""
        z = C.c_void_p(0)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().rt_ppo_record(C.c_void_p(engine.reward_f32.data_ptr()), C.c_void_p(engine.terminated.data_ptr()),
                                              C.c_void_p(engine.truncated.data_ptr()), C.c_void_p(engine.info.data_ptr()), self.n,
                                              C.c_void_p(self.counters.data_ptr()), self.T,
                                              C.c_void_p(self.rewards.data_ptr()) if store else z,
                                              C.c_void_p(self.next_done.data_ptr()), C.c_void_p(self.episode_stats.data_ptr()),
                                              self._stream()), "rt_ppo_record")

    def advance(self):
        self.counters.add_(1)
        self._row += 1
        self._rng += 1

    def begin_iteration(self):
        self.counters[:1].zero_()
        self._row = 0
        self.episode_stats.zero_()

    def rollout(self, engine, n_steps: Optional[int] = None) -> None:
        """train.py:138-161 for `n_steps` steps (default: the rest of the iteration) in one launch: rows
        [row, row + n_steps) of every rollout buffer, engine.obs / self.next_done carried across, episode statistics
        accumulated.  Equal, bit for bit, to n_steps x (act, engine.step, record, advance)."""
        n_steps = self.T - self._row if n_steps is None else int(n_steps)
        if engine.num_envs != self.n or engine.device != self.device:
            raise ValueError("FusedRollout.rollout: the engine must hold this rollout's envs on the agent's device")
        with torch.cuda.device(self.device):
            nat.check(nat.lib().rt_rollout(engine._h, C.byref(self._params), n_steps, self._row, self.T, C.c_uint64(self.seed),
                                           self._rng, C.c_void_p(self.obs.data_ptr()), C.c_void_p(self.dones.data_ptr()),
                                           C.c_void_p(self.values.data_ptr()), C.c_void_p(self.actions.data_ptr()),
                                           C.c_void_p(self.logprobs.data_ptr()), C.c_void_p(self.rewards.data_ptr()),
                                           C.c_void_p(engine.obs.data_ptr()), C.c_void_p(self.next_done.data_ptr()),
                                           C.c_void_p(self.episode_stats.data_ptr()), self._stream()), "rt_rollout")
        self.counters.add_(n_steps)
        self._row += n_steps
        self._rng += n_steps
