"""Single-episode façade with the reference's class name and method surface
(environment.py:15-303), backed by a one-env batch on the GPU.

    env = RadiotherapyEnv(visionless=True)
    obs, info = env.reset()
    obs, reward, terminated, truncated, info = env.step(action)

Differences that follow from the declared spaces rather than from the arithmetic:
observations are returned as float32 (the dtype of observation_space, and what
SyncVectorEnv hands to train.py) where the reference's vector observation is float64;
actions are cast to float32 (the dtype of action_space).  The tumour is drawn from a
seeded counter-based RNG (or given explicitly with options={"tumour_id": k}) instead of
the global NumPy RNG over an unsorted directory listing (environment.py:28,90).
"""
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from .engine import BatchedEpisodes
from .phantom import Phantom
from .vector_env import Box, EnvView


class RadiotherapyEnv:
    ACTION_SIZE = 6
    MAX_TIME_STEPS = 100
    MIN_ANGLE_Z = np.pi / 4
    BEAM_DOSE = 0.1
    LUNG_DOSE_THRESHOLD = 0.2
    TUMOUR_DOSE_THRESHOLD = 0.9
    LUNG_DOSE_REWARD = -1.0
    TUMOUR_DOSE_REWARD = 10.0
    DISTANCE_TO_TUMOUR_REWARD = -1.0
    MOVEMENT_SPEED = 0.2
    ROTATION_SPEED = 0.5

    metadata = {"render_modes": ["human"], "render_fps": 30}

    def __init__(self, visionless: bool = False, device="cuda", phantom: Optional[Phantom] = None, seed: int = 0,
                 tumour_id: Optional[int] = None):
        self.visionless = visionless
        self.engine = BatchedEpisodes(1, device=device, phantom=phantom, record_beams=True, seed=seed)
        self.LUNG_SHAPE = np.array(self.engine.grid)
        self._view = EnvView(self, 0)
        self._seed = int(seed)
        self._resets = 0
        self.observation_shape = (9,) if visionless else (4,) + self.engine.grid
        self.observation_space = Box(low=0.0, high=1.0, shape=self.observation_shape, dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(self.ACTION_SIZE,), dtype=np.float32)
        self.max_episode_steps = self.MAX_TIME_STEPS
        self.t = 0
        self.done = False
        self.reset(options=None if tumour_id is None else {"tumour_id": tumour_id})   # environment.py:52

    # -- reference attribute names (environment.py:39-49) --------------------------------
    beam_position = property(lambda self: self._view.beam_position)
    beam_direction = property(lambda self: self._view.beam_direction)
    dose = property(lambda self: self._view.dose)
    lungs = property(lambda self: self._view.lungs)
    tumours = property(lambda self: self._view.tumours)
    tumour_id = property(lambda self: self._view.tumour_id)
    beams = property(lambda self: self._view.beams)

    def observation(self):
        return self.get_vector_observation() if self.visionless else self.get_volumes()

    def get_vector_observation(self) -> np.ndarray:
        return self.engine.obs[0].cpu().numpy()

    def get_volumes(self) -> np.ndarray:
        return self.engine.volumes(0, 1)[0].cpu().numpy()

    def reset(self, seed=None, options=None):
        options = options or {}
        if seed is not None:
            self._seed, self._resets = int(seed), 0
        if "tumour_id" in options:
            self.engine.set_tumour_schedule(np.array([[int(options["tumour_id"])]], dtype=np.int32))
        else:
            # rt_reset restarts the episode counter; a new seed per reset gives a fresh draw per episode
            self.engine.set_tumour_schedule(None)
            self.engine.seed(self._seed + self._resets)
        self._resets += 1
        self.engine.reset()
        self.t = 0
        self.done = False
        return self.observation(), {}

    def step(self, action):
        if self.done:
            raise RuntimeError("step() called on a finished episode: call reset() first")
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, self.ACTION_SIZE),
                            device=self.engine.device)
        _, reward, term, _, info = self.engine.step(a)
        row = info[0].cpu().numpy()
        self.t = int(row[nat.INFO_T])
        self.done = bool(term.item())
        pose = self.engine.pose()[0].cpu().numpy()
        out_info = {                                                        # environment.py:222-241
            "reward_components": {
                "total": float(row[nat.INFO_REWARD_TOTAL]),
                "tumour": float(row[nat.INFO_REWARD_TUMOUR]),
                "lung": float(row[nat.INFO_REWARD_LUNG]),
                "distance_to_tumour": float(row[nat.INFO_REWARD_DISTANCE]),
            },
            "beam_position": {"translation": list(pose[:3]), "rotation": list(pose[3:])},
            "doses": {"tumour": float(row[nat.INFO_DOSE_TUMOUR]), "lung": float(row[nat.INFO_DOSE_LUNG])},
            "overshoot": {"translation": list(row[nat.INFO_OVERSHOOT_T0:nat.INFO_OVERSHOOT_T0 + 3]),
                          "rotation": float(row[nat.INFO_OVERSHOOT_R])},
        }
        return self.observation(), float(reward.item()), self.done, False, out_info

    def export_trajectory(self, filename):
        self._view.export_trajectory(filename)

    def export_animation(self, output_file=None):
        self._view.export_animation(output_file)

    def render(self):
        raise NotImplementedError("rendering (graphics.py / trimesh) is outside the environment-step path")

    def close(self):
        self.engine.close()
