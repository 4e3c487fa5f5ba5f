"""Vector environment with the surface train.py / ppo_eval.py use from
`gym.vector.SyncVectorEnv([make_env(...)] * N)` (train.py:93-100,123,151-161,281),
backed by N episodes resident on one GPU.

Semantics reproduced from gymnasium 1.0.0 (pinned in the reference's environment.yaml:222;
the package itself is not a dependency here):
  * autoreset mode NEXT_STEP: the step on which an env terminates returns its terminal
    observation and reward; on the following step() its action is ignored, the env is
    reset and it reports reward 0 / terminated False;
  * per-env info dicts are merged into arrays with boolean `_key` masks;
  * RecordEpisodeStatistics (train.py:36) adds info["episode"] = {"r","l","t"} on terminal steps.

Two data paths, chosen by the type of `actions` passed to step():
  numpy  -> host path (rt_step_host): numpy in / numpy out, like the reference;
  torch CUDA tensor -> device path (rt_step): tensors in / tensors out, no host sync; `infos`
           is materialised lazily (touching it costs one device->host copy).
"""
import time
from collections.abc import Mapping
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from .engine import BatchedEpisodes
from .phantom import Phantom

try:  # use gymnasium's Box when it is installed so isinstance checks in user code hold
    from gymnasium.spaces import Box  # type: ignore
except Exception:  # pragma: no cover - gymnasium is absent in the build image
    class Box:
        """Minimal stand-in for gymnasium.spaces.Box (shape / dtype / bounds holder)."""

        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


_COMPONENTS = (("total", nat.INFO_REWARD_TOTAL), ("tumour", nat.INFO_REWARD_TUMOUR),
               ("lung", nat.INFO_REWARD_LUNG), ("distance_to_tumour", nat.INFO_REWARD_DISTANCE))


def build_infos(info: np.ndarray, terminated: np.ndarray, elapsed: float) -> dict:
    """Merge the [N][16] info block into gymnasium-style vector infos (see module docstring)."""
    stepped = info[:, nat.INFO_STEPPED] > 0
    infos = {}
    if stepped.any():
        rc = {}
        for name, col in _COMPONENTS:
            rc[name] = np.where(stepped, info[:, col], 0.0)
            rc["_" + name] = stepped.copy()
        infos["reward_components"] = rc
        infos["_reward_components"] = stepped.copy()
        # info["beam_position"] (pose lists, environment.py:229-232) is served by envs[i].beam_position
        doses = {"tumour": np.where(stepped, info[:, nat.INFO_DOSE_TUMOUR], 0.0),
                 "lung": np.where(stepped, info[:, nat.INFO_DOSE_LUNG], 0.0),
                 "_tumour": stepped.copy(), "_lung": stepped.copy()}
        infos["doses"] = doses
        infos["_doses"] = stepped.copy()
        ov = {"translation": np.where(stepped[:, None], info[:, nat.INFO_OVERSHOOT_T0:nat.INFO_OVERSHOOT_T0 + 3], 0.0),
              "rotation": np.where(stepped, info[:, nat.INFO_OVERSHOOT_R], 0.0),
              "_translation": stepped.copy(), "_rotation": stepped.copy()}
        infos["overshoot"] = ov
        infos["_overshoot"] = stepped.copy()
    fin = terminated.astype(bool) & stepped
    if fin.any():
        infos["episode"] = {
            "r": np.where(fin, info[:, nat.INFO_EPISODE_RETURN], 0.0),
            "l": np.where(fin, info[:, nat.INFO_EPISODE_LENGTH], 0).astype(np.int64),
            "t": np.where(fin, elapsed, 0.0),
            "_r": fin.copy(), "_l": fin.copy(), "_t": fin.copy(),
        }
        infos["_episode"] = fin.copy()
    return infos


class LazyInfos(Mapping):
    """`infos` of the device path: a read-only mapping that copies the info block to the host
    the first time it is inspected (`"episode" in infos`, `infos["reward_components"]`, ...)."""

    def __init__(self, info_dev: torch.Tensor, terminated_dev: torch.Tensor, t0: float):
        # snapshot now (device-side, asynchronous): the engine reuses its output buffers every step
        self._info = info_dev.clone()
        self._term = terminated_dev.clone()
        self._t0 = t0
        self._dict = None

    def _materialise(self) -> dict:
        if self._dict is None:
            self._dict = build_infos(self._info.cpu().numpy(), self._term.cpu().numpy(), time.perf_counter() - self._t0)
        return self._dict

    def __getitem__(self, key):
        return self._materialise()[key]

    def __iter__(self):
        return iter(self._materialise())

    def __len__(self):
        return len(self._materialise())


class EnvView:
    """`envs.envs[i]`: read access to one episode's state with the reference's attribute names
    (environment.py:39-49) — used by ppo_visualize.py:22 and for inspection."""

    def __init__(self, owner: "RadiotherapyVectorEnv", index: int):
        self._o, self._i = owner, index

    @property
    def beam_position(self) -> np.ndarray:
        return self._o.engine.pose()[self._i, :3].cpu().numpy()

    @property
    def beam_direction(self) -> np.ndarray:
        return self._o.engine.pose()[self._i, 3:].cpu().numpy()

    @property
    def dose(self) -> np.ndarray:
        return self._o.engine.dose(self._i).cpu().numpy()

    @property
    def lungs(self) -> np.ndarray:
        return self._o.engine.phantom.lungs_volume()

    @property
    def tumour_id(self) -> int:
        return int(self._o.engine.counters()[self._i, 1].item())

    @property
    def t(self) -> int:
        return int(self._o.engine.counters()[self._i, 0].item())

    @property
    def tumours(self) -> np.ndarray:
        return self._o.engine.phantom.tumour_volume(self.tumour_id)

    @property
    def beams(self):
        b = self._o.engine.beams(self._i).cpu().numpy()
        return [(row[:3].copy(), row[3:].copy()) for row in b]

    def get_volumes(self) -> np.ndarray:
        return self._o.engine.volumes(self._i, 1)[0].cpu().numpy()

    def export_trajectory(self, filename):
        """environment.py:69-75: npz with keys tumours, dose, beams."""
        np.savez_compressed(filename, tumours=self.tumours, dose=self.dose,
                            beams=np.array([np.stack(b) for b in self.beams]))

    def export_animation(self, output_file=None):
        raise NotImplementedError("rendering (graphics.py / trimesh) is outside the environment-step path; "
                                  "use export_trajectory() and the reference's offline tools")


class RadiotherapyVectorEnv:
    """N RadiotherapyEnv episodes behind the SyncVectorEnv API (see module docstring)."""

    metadata = {"render_modes": ["human"], "render_fps": 30, "autoreset_mode": "next_step"}

    def __init__(self, num_envs: int, visionless: bool = True, device="cuda", phantom: Optional[Phantom] = None,
                 seed: int = 0, tumour_ids=None, record_beams: bool = False, dense: bool = False):
        self.num_envs = int(num_envs)
        self.visionless = bool(visionless)
        self.engine = BatchedEpisodes(self.num_envs, device=device, phantom=phantom, record_beams=record_beams,
                                      seed=seed, dense=dense)
        self.device = self.engine.device
        grid = self.engine.grid
        obs_shape = (nat.OBS_SIZE,) if self.visionless else (4,) + grid
        # environment.py:59-65 declares Box(0, 1) although the vector observation spans [-1, 1]
        self.single_observation_space = Box(low=0.0, high=1.0, shape=obs_shape, dtype=np.float32)
        self.single_action_space = Box(low=-1.0, high=1.0, shape=(nat.ACTION_SIZE,), dtype=np.float32)
        self.observation_space = Box(low=0.0, high=1.0, shape=(self.num_envs,) + obs_shape, dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(self.num_envs, nat.ACTION_SIZE), dtype=np.float32)
        self.max_episode_steps = nat.MAX_TIME_STEPS
        if tumour_ids is not None:
            self.engine.set_tumour_schedule(tumour_ids)
        self.envs = [EnvView(self, i) for i in range(self.num_envs)] if self.num_envs <= 4096 else _LazyViews(self)
        self._t0 = time.perf_counter()
        self._closed = False
        n = self.num_envs
        pin = torch.cuda.is_available()
        # page-locked host buffers: the step kernel reads/writes them in place over PCIe
        self._h_actions = torch.empty((n, nat.ACTION_SIZE), dtype=torch.float32, pin_memory=pin)
        self._h_obs = torch.empty((n, nat.OBS_SIZE), dtype=torch.float32, pin_memory=pin)
        self._h_reward = torch.empty(n, dtype=torch.float64, pin_memory=pin)
        self._h_term = torch.empty(n, dtype=torch.uint8, pin_memory=pin)
        self._h_trunc = torch.empty(n, dtype=torch.uint8, pin_memory=pin)
        self._h_info = torch.zeros((n, nat.INFO_SIZE), dtype=torch.float64, pin_memory=pin)
        self._vol = None
        self._last_on_host = False
        self._host_step = self.engine.bind_step_host(self._h_actions.numpy(), self._h_obs.numpy(),
                                                     self._h_reward.numpy(), self._h_term.numpy(),
                                                     self._h_trunc.numpy(), self._h_info.numpy())

    # -- helpers ----------------------------------------------------------------------
    def last_info(self) -> np.ndarray:
        """The [N][16] info block of the most recent reset/step."""
        return self._h_info.numpy() if self._last_on_host else self.engine.info.cpu().numpy()

    def _volumes(self) -> torch.Tensor:
        if self._vol is None:
            self._vol = torch.empty((self.num_envs, 4) + self.engine.grid, dtype=torch.float32, device=self.device)
        return self.engine.volumes(0, self.num_envs, out=self._vol)

    # -- gymnasium vector API ------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """Reset every env.  `seed` reseeds the tumour RNG (the reference ignores it,
        environment.py:77-84); options={"tumour_ids": int[E][N]} pins the tumour schedule;
        options={"backend": "torch"} returns a CUDA tensor instead of numpy."""
        options = options or {}
        if seed is not None:
            self.engine.seed(int(seed))
        if "tumour_ids" in options:
            self.engine.set_tumour_schedule(options["tumour_ids"])
        self._t0 = time.perf_counter()
        obs = self.engine.reset()
        self.engine.info.zero_()
        self._last_on_host = False
        if not self.visionless:
            obs = self._volumes()
        if options.get("backend") == "torch":
            return obs, {}
        return obs.cpu().numpy(), {}

    def step(self, actions):
        if isinstance(actions, torch.Tensor) and actions.is_cuda:
            return self._step_device(actions)
        return self._step_host(np.asarray(actions))

    def _step_device(self, actions: torch.Tensor):
        obs, reward, term, trunc, info = self.engine.step(actions, want_info=True)
        self._last_on_host = False
        if not self.visionless:
            obs = self._volumes()
        return obs, reward, term.bool(), trunc.bool(), LazyInfos(info, term, self._t0)

    def _step_host(self, actions: np.ndarray):
        n = self.num_envs
        if actions.shape != (n, nat.ACTION_SIZE):
            raise ValueError(f"actions must have shape {(n, nat.ACTION_SIZE)}, got {actions.shape}")
        self._h_actions.numpy()[...] = actions            # cast to float32 = the declared action dtype
        self._host_step()
        self._last_on_host = True
        term = self._h_term.numpy().astype(bool)
        infos = build_infos(self._h_info.numpy(), term, time.perf_counter() - self._t0)
        if self.visionless:
            obs = self._h_obs.numpy().copy()
        else:
            obs = self._volumes().cpu().numpy()
        return obs, self._h_reward.numpy().copy(), term, self._h_trunc.numpy().astype(bool), infos

    def close(self):
        if not self._closed:
            self.engine.close()
            self._closed = True


class _LazyViews:
    """envs[i] for very large batches without building N Python objects up front."""

    def __init__(self, owner):
        self._o = owner

    def __len__(self):
        return self._o.num_envs

    def __getitem__(self, i):
        if not -self._o.num_envs <= i < self._o.num_envs:
            raise IndexError(i)
        return EnvView(self._o, i % self._o.num_envs)
