"""CUDA replacements for the reference's geometry functions, with the reference's signatures.

    beam_voxels(base_matrix, position, direction, epsilon)    draw_line.py:4
    apply_rotation(initial_direction, rotation_vector, min_angle)   transforms.py:7
    apply_translation(position, translation_vector, bounds)   transforms.py:62

The single-call forms take and return numpy arrays like the reference; the `*_batch`
forms keep everything on the device (torch tensors).  All of them run the kernels of
librtenv_b200.so — there is no CPU path.
"""
import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as nat
from .engine import _ptr, _require_cuda, _stream

_EPS = 1e-6


def _grid_arg(shape):
    g = [int(x) for x in shape]
    if len(g) != 3:
        raise ValueError("the volume must be 3-dimensional")
    return (C.c_int32 * 3)(*g)


def beam_voxels_batch(position: torch.Tensor, direction: torch.Tensor, grid=(67, 43, 70)):
    """Distinct voxels hit by m rays and their summed bilinear weights.

    position, direction: float64 [m][3] CUDA tensors.  Returns (idx int32 [m][cap],
    weight float32 [m][cap], count int32 [m]) with cap = max(288, 4*(max(grid)+1)); count == -1 where the direction norm
    is below 1e-6 (the reference raises ValueError there)."""
    device = _require_cuda(position.device)
    pos = position.to(dtype=torch.float64).contiguous().reshape(-1, 3)
    dr = direction.to(device=device, dtype=torch.float64).contiguous().reshape(-1, 3)
    m = pos.shape[0]
    cap = max(nat.BEAM_CAP, 4 * (max(int(g) for g in grid) + 1))
    idx = torch.zeros((m, cap), dtype=torch.int32, device=device)
    w = torch.zeros((m, cap), dtype=torch.float32, device=device)
    count = torch.zeros(m, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        nat.check(nat.lib().rt_beam_voxels(_grid_arg(grid), _ptr(pos), _ptr(dr), m, cap, _ptr(idx),
                                           _ptr(w), _ptr(count), _stream(device)), "rt_beam_voxels")
    return idx, w, count


def beam_voxels_dense_batch(position: torch.Tensor, direction: torch.Tensor, grid=(67, 43, 70)):
    """Dense float32 [m][G0][G1][G2] beam volumes plus int32 status [m] (-1 = direction too small)."""
    device = _require_cuda(position.device)
    pos = position.to(dtype=torch.float64).contiguous().reshape(-1, 3)
    dr = direction.to(device=device, dtype=torch.float64).contiguous().reshape(-1, 3)
    m = pos.shape[0]
    out = torch.empty((m,) + tuple(int(g) for g in grid), dtype=torch.float32, device=device)
    status = torch.zeros(m, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        nat.check(nat.lib().rt_beam_voxels_dense(_grid_arg(grid), _ptr(pos), _ptr(dr), m, _ptr(out), _ptr(status),
                                                 _stream(device)), "rt_beam_voxels_dense")
    return out, status


def beam_voxels(base_matrix, position, direction, epsilon: float = _EPS, device="cuda") -> np.ndarray:
    """draw_line.py:4 — dense float32 volume of bilinear splat weights along the line.

    `base_matrix` is used for its shape only, as in the reference.  Raises
    ValueError("Direction vector magnitude is too small.") like draw_line.py:23-24."""
    if epsilon != _EPS:
        raise ValueError("the CUDA path implements the reference's default epsilon=1e-6 only")
    device = _require_cuda(device)
    pos = torch.as_tensor(np.asarray(position, dtype=np.float64).reshape(1, 3), device=device)
    dr = torch.as_tensor(np.asarray(direction, dtype=np.float64).reshape(1, 3), device=device)
    out, status = beam_voxels_dense_batch(pos, dr, np.shape(base_matrix))
    if int(status.item()) < 0:
        raise ValueError("Direction vector magnitude is too small.")
    return out[0].cpu().numpy()


def apply_rotation_batch(direction: torch.Tensor, rotation_vector: torch.Tensor, min_angle: float):
    device = _require_cuda(direction.device)
    d = direction.to(dtype=torch.float64).contiguous().reshape(-1, 3)
    rv = rotation_vector.to(device=device, dtype=torch.float64).contiguous().reshape(-1, 3)
    m = d.shape[0]
    out = torch.empty_like(d)
    os_r = torch.empty(m, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        nat.check(nat.lib().rt_apply_rotation(_ptr(d), _ptr(rv), m, float(min_angle), _ptr(out), _ptr(os_r),
                                              _stream(device)), "rt_apply_rotation")
    return out, os_r


def apply_rotation(initial_direction, rotation_vector, min_angle: float, device="cuda") -> Tuple[np.ndarray, float]:
    """transforms.py:7 — rotate a direction by a rotation vector, keep it >= min_angle from axis 0."""
    device = _require_cuda(device)
    d = torch.as_tensor(np.asarray(initial_direction, dtype=np.float64).reshape(1, 3), device=device)
    rv = torch.as_tensor(np.asarray(rotation_vector, dtype=np.float64).reshape(1, 3), device=device)
    out, os_r = apply_rotation_batch(d, rv, min_angle)
    return out[0].cpu().numpy(), float(os_r.item())


def apply_translation_batch(position: torch.Tensor, translation: torch.Tensor, bounds):
    device = _require_cuda(position.device)
    p = position.to(dtype=torch.float64).contiguous().reshape(-1, 3)
    t = translation.to(device=device, dtype=torch.float64).contiguous().reshape(-1, 3)
    m = p.shape[0]
    b = (C.c_double * 3)(*[float(x) for x in np.asarray(bounds).reshape(3)])
    out = torch.empty_like(p)
    os_t = torch.empty_like(p)
    with torch.cuda.device(device):
        nat.check(nat.lib().rt_apply_translation(_ptr(p), _ptr(t), m, b, _ptr(out), _ptr(os_t), _stream(device)),
                  "rt_apply_translation")
    return out, os_t


def apply_translation(position, translation_vector, bounds, device="cuda") -> Tuple[np.ndarray, np.ndarray]:
    """transforms.py:62 — clip(position + translation, 0, bounds) and the clipped-off overshoot."""
    device = _require_cuda(device)
    p = torch.as_tensor(np.asarray(position, dtype=np.float64).reshape(1, 3), device=device)
    t = torch.as_tensor(np.asarray(translation_vector, dtype=np.float64).reshape(1, 3), device=device)
    out, os_t = apply_translation_batch(p, t, bounds)
    return out[0].cpu().numpy(), os_t[0].cpu().numpy()


def pose_update_batch(position: torch.Tensor, direction: torch.Tensor, actions: torch.Tensor, grid=(67, 43, 70)):
    """environment.py:196-207 for m independent poses: map the action, translate, rotate."""
    device = _require_cuda(position.device)
    p = position.to(dtype=torch.float64).contiguous().reshape(-1, 3)
    d = direction.to(device=device, dtype=torch.float64).contiguous().reshape(-1, 3)
    a = actions.to(device=device, dtype=torch.float32).contiguous().reshape(-1, 6)
    m = p.shape[0]
    po, do, ot = torch.empty_like(p), torch.empty_like(d), torch.empty_like(p)
    orr = torch.empty(m, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        nat.check(nat.lib().rt_pose_update(_grid_arg(grid), _ptr(p), _ptr(d), _ptr(a), m, _ptr(po), _ptr(do),
                                           _ptr(ot), _ptr(orr), _stream(device)), "rt_pose_update")
    return po, do, ot, orr


def compute_gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, next_value: torch.Tensor,
                next_done: torch.Tensor, gamma: float, gae_lambda: float,
                out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """train.py:164-181 — advantages and returns for float32 [T][N] rollout tensors on the device."""
    device = _require_cuda(rewards.device)
    T, N = rewards.shape
    f = lambda x: x.to(device=device, dtype=torch.float32).contiguous()
    rewards, values, dones = f(rewards), f(values), f(dones)
    next_value = f(next_value).reshape(-1)
    next_done = f(next_done).reshape(-1)
    if values.shape != (T, N) or dones.shape != (T, N) or next_value.numel() != N or next_done.numel() != N:
        raise ValueError("GAE inputs disagree on (T, N)")
    adv, ret = out if out is not None else (torch.empty_like(rewards), torch.empty_like(rewards))
    with torch.cuda.device(device):
        nat.check(nat.lib().rt_gae(_ptr(rewards), _ptr(values), _ptr(dones), _ptr(next_value), _ptr(next_done), T, N,
                                   float(gamma), float(gae_lambda), _ptr(adv), _ptr(ret), _stream(device)), "rt_gae")
    return adv, ret
