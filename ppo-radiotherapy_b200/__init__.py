"""B200-native batched environment step for rmaguado/ppo-radiotherapy.

Public surface (mirrors the reference's module/function names for the hot path):

    RadiotherapyEnv            environment.py:15    single episode, gymnasium reset/step API
    RadiotherapyVectorEnv      train.py:93          SyncVectorEnv-compatible batch resident in HBM
    BatchedEpisodes            —                    the device-resident engine under both
    beam_voxels                draw_line.py:4
    apply_rotation / apply_translation     transforms.py:7 / :62
    compute_gae                train.py:164-181
    PPO, PPO_3DCNN             networks.py:54,107   (PyTorch; checkpoint-compatible)
    train.train / train.main   train.py:91,285      device-resident CleanRL loop, NCCL grad all-reduce
    ppo_eval.evaluate          ppo_eval.py:5        evaluation of a saved policy
    Phantom                    environment.py:28-29,90-97 data, packed

Everything computes in librtenv_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/rt_env.h).  There is no CPU fallback.
"""
from . import _native
from ._native import RtError, build
from .phantom import Phantom, default_phantom
from .engine import BatchedEpisodes, ObservationStore
from .geometry import (apply_rotation, apply_rotation_batch, apply_translation, apply_translation_batch,
                       beam_voxels, beam_voxels_batch, beam_voxels_dense_batch, compute_gae, pose_update_batch)
from .vector_env import Box, RadiotherapyVectorEnv
from .environment import RadiotherapyEnv
from .networks import PPO, PPO_3DCNN, FeaturesExtractor3D
from .rollout import FusedRollout
from . import train as train          # noqa: F401  (ppo_radiotherapy_b200.train.train / main)
from . import ppo_eval as ppo_eval    # noqa: F401

__all__ = [
    "RtError", "build", "Phantom", "default_phantom", "BatchedEpisodes", "ObservationStore", "RadiotherapyEnv",
    "RadiotherapyVectorEnv", "Box", "beam_voxels", "beam_voxels_batch", "beam_voxels_dense_batch",
    "apply_rotation", "apply_rotation_batch", "apply_translation", "apply_translation_batch",
    "pose_update_batch", "compute_gae", "PPO", "PPO_3DCNN", "FeaturesExtractor3D", "FusedRollout", "train",
]
