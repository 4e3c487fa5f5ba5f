"""Device-resident batch of radiotherapy episodes: thin Python over the C ABI.

`BatchedEpisodes` owns one `rt_env` handle (N episodes resident in the HBM of one
GPU) and exchanges data as torch CUDA tensors whose raw pointers are handed to the
library together with torch's current stream.  PyTorch is only the allocator and the
stream owner here; all arithmetic happens in librtenv_b200.so.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from .phantom import Phantom, default_phantom


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise nat.RtError("the environment step runs on CUDA devices only (no CPU fallback)")
    if not torch.cuda.is_available():
        raise nat.RtError("CUDA is not available: the environment step has no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class BatchedEpisodes:
    """N independent RadiotherapyEnv episodes (environment.py:15) stepped by one kernel launch."""

    def __init__(self, num_envs: int, device="cuda", phantom: Optional[Phantom] = None,
                 record_beams: bool = False, seed: int = 0, dense: bool = False):
        self.device = _require_cuda(device)
        self.num_envs = int(num_envs)
        self.phantom = phantom if phantom is not None else default_phantom()
        self.grid = tuple(int(g) for g in self.phantom.grid)
        self.nvox = self.phantom.nvox
        self._lib = nat.lib()
        self._h = C.c_void_p()
        # dense=True: full-volume dose update and from-scratch reductions every step (the reference's own
        # dataflow; BASELINE configs[4] "dose-grid stress") instead of the sparse incremental step
        self.dense = bool(dense)
        flags = (nat.FLAG_RECORD_BEAMS if record_beams else 0) | (nat.FLAG_DENSE if dense else 0)
        desc = self.phantom.desc()
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_create(C.byref(self._h), self.device.index, self.num_envs, flags, C.byref(desc)),
                      "rt_create")
        n, dev = self.num_envs, self.device
        self.obs = torch.empty((n, nat.OBS_SIZE), dtype=torch.float32, device=dev)
        self.reward = torch.empty(n, dtype=torch.float64, device=dev)
        self.reward_f32 = torch.empty(n, dtype=torch.float32, device=dev)
        self.terminated = torch.empty(n, dtype=torch.uint8, device=dev)
        self.truncated = torch.empty(n, dtype=torch.uint8, device=dev)
        self.info = torch.empty((n, nat.INFO_SIZE), dtype=torch.float64, device=dev)
        self.seed(seed)

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            torch.cuda.synchronize(self.device)
            self._lib.rt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(self._lib.rt_device_bytes(self._h))

    # -- tumour choice ------------------------------------------------------------------
    def seed(self, seed: int):
        nat.check(self._lib.rt_seed(self._h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)), "rt_seed")

    def set_tumour_schedule(self, ids):
        """ids int [E][N]: episode e of env i uses tumour ids[min(e, E-1)][i]; None = device RNG."""
        if ids is None:
            nat.check(self._lib.rt_set_tumour_schedule(self._h, None, 0), "rt_set_tumour_schedule")
            return
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if ids.ndim == 1:
            ids = ids[None, :]
        if ids.shape[1] != self.num_envs:
            raise ValueError(f"tumour schedule must have {self.num_envs} columns, got {ids.shape}")
        nat.check(self._lib.rt_set_tumour_schedule(self._h, C.c_void_p(ids.ctypes.data), ids.shape[0]),
                  "rt_set_tumour_schedule")

    # -- reset / step -------------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_reset(self._h, _ptr(mask), _ptr(self.obs), _stream(self.device)), "rt_reset")
        return self.obs

    def step(self, actions: torch.Tensor, want_info: bool = True):
        """actions float32 [N][6] on this device.  Returns views of the persistent output tensors
        (obs, reward f64, terminated u8, truncated u8, info f64 [N][16] or None)."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.shape != (self.num_envs, nat.ACTION_SIZE):
            raise ValueError(f"actions must have shape {(self.num_envs, nat.ACTION_SIZE)}, got {tuple(actions.shape)}")
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_step(self._h, _ptr(actions), _ptr(self.obs), _ptr(self.reward),
                                        _ptr(self.reward_f32), _ptr(self.terminated), _ptr(self.truncated),
                                        _ptr(self.info if want_info else None), _stream(self.device)), "rt_step")
        return self.obs, self.reward, self.terminated, self.truncated, (self.info if want_info else None)

    def step_host(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, terminated: np.ndarray,
                  truncated: np.ndarray, info: Optional[np.ndarray] = None):
        """Host-buffer step (numpy in, numpy out, synchronous): the reference's own seam, train.py:151."""
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_step_host(self._h, C.c_void_p(actions.ctypes.data), C.c_void_p(obs.ctypes.data),
                                             C.c_void_p(reward.ctypes.data), C.c_void_p(terminated.ctypes.data),
                                             C.c_void_p(truncated.ctypes.data),
                                             None if info is None else C.c_void_p(info.ctypes.data)), "rt_step_host")

    def bind_step_host(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, terminated: np.ndarray,
                       truncated: np.ndarray, info: Optional[np.ndarray] = None):
        """Pre-bound host-buffer step for fixed buffers: returns a zero-argument callable (the per-call Python
        cost is one ctypes call; `ndarray.ctypes` alone costs more than the kernel launch)."""
        fn, h = self._lib.rt_step_host, self._h
        args = (h, C.c_void_p(actions.ctypes.data), C.c_void_p(obs.ctypes.data), C.c_void_p(reward.ctypes.data),
                C.c_void_p(terminated.ctypes.data), C.c_void_p(truncated.ctypes.data),
                None if info is None else C.c_void_p(info.ctypes.data))
        keep = (actions, obs, reward, terminated, truncated, info)
        dev_index = self.device.index

        def call(_keep=keep):
            if torch.cuda.current_device() != dev_index:
                torch.cuda.set_device(dev_index)
            rc = fn(*args)
            if rc != 0:
                nat.check(rc, "rt_step_host")
        return call

    def reset_host(self, obs: np.ndarray, mask: Optional[np.ndarray] = None):
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_reset_host(self._h, None if mask is None else C.c_void_p(mask.ctypes.data),
                                              C.c_void_p(obs.ctypes.data)), "rt_reset_host")

    # -- state access -------------------------------------------------------------------
    def pose(self) -> torch.Tensor:
        out = torch.empty((self.num_envs, 6), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_get_pose(self._h, _ptr(out), _stream(self.device)), "rt_get_pose")
        return out

    def set_pose(self, pose: torch.Tensor):
        pose = pose.to(device=self.device, dtype=torch.float64).contiguous()
        if pose.shape != (self.num_envs, 6):
            raise ValueError("pose must have shape [N][6]")
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_set_pose(self._h, _ptr(pose), _stream(self.device)), "rt_set_pose")

    def counters(self) -> torch.Tensor:
        """int32 [N][6]: t, tumour id, lung voxels above threshold, episode index, needs-reset, beams recorded."""
        out = torch.empty((self.num_envs, 6), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_get_counters(self._h, _ptr(out), _stream(self.device)), "rt_get_counters")
        return out

    def dose(self, env_index: int) -> torch.Tensor:
        out = torch.empty(self.grid, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_get_dose(self._h, int(env_index), _ptr(out), _stream(self.device)), "rt_get_dose")
        return out

    def volumes(self, first: int = 0, count: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Voxel observations float32 [count][4][G0][G1][G2] (get_volumes, environment.py:252-257)."""
        count = self.num_envs - first if count is None else count
        shape = (count, 4) + self.grid
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {shape}")
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_assemble_volumes(self._h, int(first), int(count), _ptr(out),
                                                    _stream(self.device)), "rt_assemble_volumes")
        return out

    # ---- compressed voxel-observation records (rollout storage) -----------------------------------------
    def observation_store(self, slots: int) -> "ObservationStore":
        """Caller-owned storage for `slots` compressed voxel observations (403 KB each instead of 3.2 MB)."""
        return ObservationStore(self, slots)

    def pack_observations(self, store: "ObservationStore", slot0: int, first: int = 0, count: Optional[int] = None):
        """Write the records of envs [first, first+count) to slots [slot0, slot0+count) of `store`."""
        count = self.num_envs - first if count is None else count
        if slot0 < 0 or slot0 + count > store.slots:
            raise ValueError("slot range outside the store")
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_pack_observations(self._h, int(first), int(count), int(slot0), _ptr(store.dose),
                                                     _ptr(store.pose), _ptr(store.tumour_id), _stream(self.device)),
                      "rt_pack_observations")

    def render_observations(self, store: "ObservationStore", index: Optional[torch.Tensor] = None,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """float32 [count][4][G0][G1][G2] rebuilt from the records `index` (int64 CUDA tensor) of `store`, or from all
        of them.  Lungs, tumour and beam-view planes equal `volumes()`; the dose plane is the stored bfloat16 value."""
        if index is not None:
            if index.dtype != torch.int64 or not index.is_cuda or not index.is_contiguous():
                raise ValueError("index must be a contiguous int64 CUDA tensor")
            count = int(index.numel())
        else:
            count = store.slots
        shape = (count, 4) + self.grid
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {shape}")
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_render_observations(self._h, _ptr(store.dose), _ptr(store.pose), _ptr(store.tumour_id),
                                                       _ptr(index), int(count), _ptr(out), _stream(self.device)),
                      "rt_render_observations")
        return out

    def beams(self, env_index: int):
        """Recorded (position, direction) pairs of the env's current episode (environment.py:110)."""
        out = torch.empty((nat.MAX_TIME_STEPS, 6), dtype=torch.float64, device=self.device)
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self._lib.rt_get_beams(self._h, int(env_index), _ptr(out), _ptr(n), _stream(self.device)),
                      "rt_get_beams")
        return out[: int(n.item())]


class ObservationStore:
    """`slots` compressed voxel observations: bfloat16 dose volumes, float64 poses, int32 tumour ids."""

    def __init__(self, engine: "BatchedEpisodes", slots: int):
        self.slots = int(slots)
        self.stride = int(engine._lib.rt_observation_record_stride(engine._h))
        self.dose = torch.empty((self.slots, self.stride), dtype=torch.bfloat16, device=engine.device)
        self.pose = torch.zeros((self.slots, 6), dtype=torch.float64, device=engine.device)
        self.tumour_id = torch.zeros(self.slots, dtype=torch.int32, device=engine.device)

    @property
    def nbytes(self) -> int:
        return self.dose.numel() * 2 + self.pose.numel() * 8 + self.tumour_id.numel() * 4
