"""ctypes front-end of the CPU oracle (oracle/rt_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product package.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_REPO = os.path.dirname(_HERE)
PHANTOM_PATH = os.path.join(_REPO, "ppo-radiotherapy_b200", "data", "phantom.npz")

STEP_OUT = 20  # ORC_STEP_OUT

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rt_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_np_sum_f32.restype = C.c_float
        L.orc_np_sum_f32.argtypes = [_fp, C.c_long]
        L.orc_beam_trace.restype = C.c_int
        L.orc_beam_trace.argtypes = [_dp, _dp, _ip, _ip, _fp, _ip]
        L.orc_beam_voxels.restype = C.c_int
        L.orc_beam_voxels.argtypes = [_dp, _dp, _ip, _fp]
        L.orc_apply_translation.restype = None
        L.orc_apply_translation.argtypes = [_dp, _dp, _dp, _dp, _dp]
        L.orc_apply_rotation.restype = None
        L.orc_apply_rotation.argtypes = [_dp, _dp, C.c_double, _dp, _dp]
        L.orc_beam_batch.restype = None
        L.orc_beam_batch.argtypes = [_dp, _dp, C.c_int, _ip, C.c_int, _ip, _fp, _ip]
        L.orc_pose_batch.restype = None
        L.orc_pose_batch.argtypes = [_dp, _dp, _fp, C.c_int, _ip, _dp, _dp, _dp, _dp]
        L.orc_env_create.restype = C.c_void_p
        L.orc_env_create.argtypes = [_fp, _ip]
        L.orc_env_destroy.restype = None
        L.orc_env_destroy.argtypes = [C.c_void_p]
        L.orc_env_reset.restype = None
        L.orc_env_reset.argtypes = [C.c_void_p, _ip, C.c_int]
        L.orc_env_vector_obs.restype = None
        L.orc_env_vector_obs.argtypes = [C.c_void_p, _dp]
        L.orc_env_volumes.restype = None
        L.orc_env_volumes.argtypes = [C.c_void_p, _fp]
        L.orc_env_step.restype = C.c_int
        L.orc_env_step.argtypes = [C.c_void_p, _fp, _dp]
        L.orc_env_dose.restype = _fp
        L.orc_env_dose.argtypes = [C.c_void_p]
        L.orc_env_pose.restype = None
        L.orc_env_pose.argtypes = [C.c_void_p, _dp]
        L.orc_env_set_pose.restype = None
        L.orc_env_set_pose.argtypes = [C.c_void_p, _dp]
        L.orc_rollout.restype = C.c_long
        L.orc_rollout.argtypes = [_fp, _ip, _ip, _ip, _ip, C.c_int, _fp, C.c_int, C.c_int,
                                  _dp, C.POINTER(C.c_byte), C.c_int, C.c_int]
        L.orc_gae.restype = None
        L.orc_gae.argtypes = [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_double, C.c_double, _fp, _fp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


def _i(a):
    return a.ctypes.data_as(_ip)


GRID = np.array([67, 43, 70], dtype=np.int32)


class Phantom:
    """The packed lungs + tumour table (tools/pack_phantom.py), expanded for the dense oracle."""

    def __init__(self, path: str = PHANTOM_PATH):
        z = np.load(path)
        self.grid = z["grid"].astype(np.int32)
        self.nvox = int(np.prod(self.grid))
        bits = z["lungs_bits"]
        flat = np.unpackbits(bits.view(np.uint8), bitorder="little")[: self.nvox]
        self.lungs = np.ascontiguousarray(flat.astype(np.float32))
        self.names = [str(x) for x in z["names"]]
        self.vox_offsets = np.ascontiguousarray(z["vox_offsets"].astype(np.int32))
        self.vox = np.ascontiguousarray(z["vox"].astype(np.int32))
        self.centroid = z["centroid"]
        self.tumour_sum = z["tumour_sum"]
        self.lung_mask_sum = z["lung_mask_sum"]
        self.n_tumours = len(self.names)

    def tumour_voxels(self, tid: int) -> np.ndarray:
        return self.vox[self.vox_offsets[tid]: self.vox_offsets[tid + 1]]


def beam_trace(pos, direction, grid=GRID):
    """(linear indices, weights, slab count) of the reference's in-bounds splat writes, in order."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    direction = np.ascontiguousarray(direction, dtype=np.float64)
    grid = np.ascontiguousarray(grid, dtype=np.int32)
    cap = 4 * (int(grid.max()) + 2)
    idx = np.empty(cap, dtype=np.int32)
    w = np.empty(cap, dtype=np.float32)
    ns = C.c_int(0)
    n = lib().orc_beam_trace(_d(pos), _d(direction), _i(grid), _i(idx), _f(w), C.byref(ns))
    if n < 0:
        raise ValueError("Direction vector magnitude is too small.")
    return idx[:n].copy(), w[:n].copy(), ns.value


def merge_trace(idx, w):
    """Sum duplicate voxels in write order -> (ascending unique indices, float32 sums)."""
    out = {}
    for i, x in zip(idx.tolist(), w):
        out[i] = np.float32(out.get(i, np.float32(0.0)) + x)
    keys = np.array(sorted(out), dtype=np.int32)
    vals = np.array([out[k] for k in keys.tolist()], dtype=np.float32)
    return keys, vals


BEAM_CAP = 288  # >= 4 * (max(G) + 1) = 284 splat writes


def beam_batch(pos, direction, grid=GRID, cap=BEAM_CAP):
    """Merged traces of m rays: (idx int32 [m,cap], w float32 [m,cap], count int32 [m]; -1 = ValueError)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    direction = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
    grid = np.ascontiguousarray(grid, dtype=np.int32)
    m = pos.shape[0]
    idx = np.zeros((m, cap), dtype=np.int32)
    w = np.zeros((m, cap), dtype=np.float32)
    count = np.zeros(m, dtype=np.int32)
    lib().orc_beam_batch(_d(pos), _d(direction), m, _i(grid), cap, _i(idx), _f(w), _i(count))
    return idx, w, count


def pose_batch(pos, direction, actions, grid=GRID):
    """One environment pose update for m independent poses -> (pos, dir, overshoot_t, overshoot_r)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    direction = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
    actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(-1, 6)
    grid = np.ascontiguousarray(grid, dtype=np.int32)
    m = pos.shape[0]
    po, do, ot, orr = np.empty((m, 3)), np.empty((m, 3)), np.empty((m, 3)), np.empty(m)
    lib().orc_pose_batch(_d(pos), _d(direction), _f(actions), m, _i(grid), _d(po), _d(do), _d(ot), _d(orr))
    return po, do, ot, orr


def beam_voxels(pos, direction, grid=GRID):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    direction = np.ascontiguousarray(direction, dtype=np.float64)
    grid = np.ascontiguousarray(grid, dtype=np.int32)
    out = np.empty(tuple(int(g) for g in grid), dtype=np.float32)
    n = lib().orc_beam_voxels(_d(pos), _d(direction), _i(grid), _f(out))
    if n < 0:
        raise ValueError("Direction vector magnitude is too small.")
    return out


def apply_translation(pos, t, bounds):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    bounds = np.ascontiguousarray(bounds, dtype=np.float64)
    o, ov = np.empty(3), np.empty(3)
    lib().orc_apply_translation(_d(pos), _d(t), _d(bounds), _d(o), _d(ov))
    return o, ov


def apply_rotation(direction, rotvec, min_angle):
    direction = np.ascontiguousarray(direction, dtype=np.float64)
    rotvec = np.ascontiguousarray(rotvec, dtype=np.float64)
    o = np.empty(3)
    ov = C.c_double(0.0)
    lib().orc_apply_rotation(_d(direction), _d(rotvec), float(min_angle), _d(o), C.byref(ov))
    return o, ov.value


def np_sum_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    return np.float32(lib().orc_np_sum_f32(_f(a), a.size))


class OracleEnv:
    """One dense reference-style episode (environment.py:15-273) with an explicit tumour id."""

    def __init__(self, phantom: Phantom, tumour_id: int = 0):
        self.ph = phantom
        self._h = lib().orc_env_create(_f(phantom.lungs), _i(phantom.grid))
        self.reset(tumour_id)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_env_destroy(self._h)
            self._h = None

    def reset(self, tumour_id: int):
        v = self.ph.tumour_voxels(tumour_id)
        lib().orc_env_reset(self._h, _i(v), v.size)
        self.tumour_id = tumour_id
        return self.vector_obs()

    def vector_obs(self):
        o = np.empty(9)
        lib().orc_env_vector_obs(self._h, _d(o))
        return o

    def volumes(self):
        out = np.empty((4,) + tuple(int(g) for g in self.ph.grid), dtype=np.float32)
        lib().orc_env_volumes(self._h, _f(out))
        return out

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32)
        out = np.empty(STEP_OUT)
        done = lib().orc_env_step(self._h, _f(a), _d(out))
        if done < 0:
            raise ValueError("Direction vector magnitude is too small.")
        return out, bool(done)

    @property
    def dose(self):
        p = lib().orc_env_dose(self._h)
        return np.ctypeslib.as_array(p, shape=(self.ph.nvox,)).reshape(tuple(int(g) for g in self.ph.grid))

    @property
    def pose(self):
        p = np.empty(6)
        lib().orc_env_pose(self._h, _d(p))
        return p

    def set_pose(self, pos, direction):
        p = np.ascontiguousarray(np.concatenate([pos, direction]), dtype=np.float64)
        lib().orc_env_set_pose(self._h, _d(p))


def rollout(phantom: Phantom, tumour_ids, actions, threads: int = 1):
    """Vector rollout with gymnasium-1.0.0 NEXT_STEP autoreset.

    tumour_ids: int32 [E][n] (episode k of env i uses tumour_ids[k][i]);
    actions: float32 [T][n][6].  Returns (out float64 [T][n][STEP_OUT], done int8 [T][n]).
    """
    tumour_ids = np.ascontiguousarray(tumour_ids, dtype=np.int32)
    if tumour_ids.ndim == 1:
        tumour_ids = tumour_ids[None, :]
    actions = np.ascontiguousarray(actions, dtype=np.float32)
    T, n, _ = actions.shape
    assert tumour_ids.shape[1] == n
    out = np.zeros((T, n, STEP_OUT), dtype=np.float64)
    done = np.zeros((T, n), dtype=np.int8)
    L = lib()

    def run(lo, hi):
        return L.orc_rollout(_f(phantom.lungs), _i(phantom.grid), _i(phantom.vox_offsets), _i(phantom.vox),
                             _i(tumour_ids), tumour_ids.shape[0], _f(actions), T, n,
                             _d(out), done.ctypes.data_as(C.POINTER(C.c_byte)), lo, hi)

    threads = max(1, min(threads, n))
    if threads == 1:
        run(0, n)
    else:
        # small chunks so threads stay balanced
        chunk = max(1, n // (threads * 4))
        bounds = [(lo, min(n, lo + chunk)) for lo in range(0, n, chunk)]
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda b: run(*b), bounds))
    return out, done


def gae(rewards, values, dones, next_value, next_done, gamma, gae_lambda):
    rewards = np.ascontiguousarray(rewards, dtype=np.float32)
    values = np.ascontiguousarray(values, dtype=np.float32)
    dones = np.ascontiguousarray(dones, dtype=np.float32)
    next_value = np.ascontiguousarray(next_value, dtype=np.float32).reshape(-1)
    next_done = np.ascontiguousarray(next_done, dtype=np.float32).reshape(-1)
    T, n = rewards.shape
    adv = np.empty_like(rewards)
    ret = np.empty_like(rewards)
    lib().orc_gae(_f(rewards), _f(values), _f(dones), _f(next_value), _f(next_done), T, n,
                  float(gamma), float(gae_lambda), _f(adv), _f(ret))
    return adv, ret
