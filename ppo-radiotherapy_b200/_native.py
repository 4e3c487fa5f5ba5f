"""ctypes binding of the C ABI in include/rt_env.h (lib/librtenv_b200.so).

There is no CPU fallback: if the shared library is missing or the call fails, the
caller gets an exception.  `build()` compiles the library in-tree with nvcc for
sm_100a (it cross-compiles on a machine without a GPU).
"""
import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librtenv_b200.so")
HEADER = os.path.join(_REPO, "include", "rt_env.h")

ABI_VERSION = 2
ACTION_SIZE = 6
OBS_SIZE = 9
MAX_TIME_STEPS = 100
INFO_SIZE = 16
BEAM_CAP = 288
FLAG_DENSE = 1
FLAG_RECORD_BEAMS = 2

# columns of the info block (enum in rt_env.h)
INFO_REWARD_TOTAL, INFO_REWARD_TUMOUR, INFO_REWARD_LUNG, INFO_REWARD_DISTANCE = 0, 1, 2, 3
INFO_DOSE_TUMOUR, INFO_DOSE_LUNG = 4, 5
INFO_OVERSHOOT_T0, INFO_OVERSHOOT_R = 6, 9
INFO_EPISODE_RETURN, INFO_EPISODE_LENGTH = 10, 11
INFO_LUNG_COUNT, INFO_STEPPED, INFO_TUMOUR_ID, INFO_T = 12, 13, 14, 15

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                    # the reference never fuses a*b+c (SURVEY §8a addendum)
    "--compiler-options", "-fPIC,-fvisibility=hidden",
    "-shared",
]


class RtError(RuntimeError):
    pass


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into lib/librtenv_b200.so for sm_100a."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RtError("nvcc not found: cannot build librtenv_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RtError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class MlpParams(C.Structure):
    """rt_mlp_params: device pointers to the parameters of the reference's MLP agent (include/rt_env.h)."""
    _fields_ = [(name, C.c_void_p) for name in (
        "critic_w0", "critic_b0", "critic_w1", "critic_b1", "critic_w2", "critic_b2",
        "actor_w0", "actor_b0", "actor_w1", "actor_b1", "actor_w2", "actor_b2", "actor_logstd")] + [
        ("n_obs", C.c_int32), ("hidden", C.c_int32), ("n_act", C.c_int32)]


class PhantomDesc(C.Structure):
    _fields_ = [
        ("grid", C.c_int32 * 3),
        ("lungs_bits", C.POINTER(C.c_uint32)),
        ("n_tumours", C.c_int32),
        ("vox_offsets", C.POINTER(C.c_int32)),
        ("vox", C.POINTER(C.c_int32)),
        ("centroid", C.POINTER(C.c_double)),
        ("tumour_sum", C.POINTER(C.c_float)),
        ("lung_mask_sum", C.POINTER(C.c_float)),
    ]


_vp = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "rt_abi_version": (C.c_int, []),
    "rt_last_error": (C.c_char_p, []),
    "rt_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint32, C.POINTER(PhantomDesc)]),
    "rt_destroy": (C.c_int, [_vp]),
    "rt_num_envs": (C.c_int, [_vp]),
    "rt_device_bytes": (C.c_int64, [_vp]),
    "rt_seed": (C.c_int, [_vp, C.c_uint64]),
    "rt_set_tumour_schedule": (C.c_int, [_vp, _vp, C.c_int]),
    "rt_reset": (C.c_int, [_vp, _vp, _vp, _vp]),
    "rt_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rt_reset_host": (C.c_int, [_vp, _vp, _vp]),
    "rt_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rt_get_pose": (C.c_int, [_vp, _vp, _vp]),
    "rt_set_pose": (C.c_int, [_vp, _vp, _vp]),
    "rt_get_counters": (C.c_int, [_vp, _vp, _vp]),
    "rt_get_dose": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "rt_assemble_volumes": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "rt_get_beams": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "rt_observation_record_stride": (C.c_int, [_vp]),
    "rt_pack_observations": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp]),
    "rt_render_observations": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "rt_beam_voxels": (C.c_int, [C.POINTER(C.c_int32), _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "rt_beam_voxels_dense": (C.c_int, [C.POINTER(C.c_int32), _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rt_pose_update": (C.c_int, [C.POINTER(C.c_int32), _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "rt_apply_rotation": (C.c_int, [_vp, _vp, C.c_int, C.c_double, _vp, _vp, _vp]),
    "rt_apply_translation": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_double), _vp, _vp, _vp]),
    "rt_gae": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp]),
    "rt_conv1_relu_pool": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rt_conv1_relu_pool_grouped": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rt_conv1_from_env": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "rt_conv2_relu_pool": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rt_c3d_tail": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rt_ppo_act": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_uint64, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rt_ppo_record": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "rt_rollout": (C.c_int, [_vp, _vp, C.c_int, C.c_int64, C.c_int, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rt_launch_count": (C.c_int64, []),
    "rt_set_stage_clock": (C.c_int, [_vp, _vp]),
    "rt_set_pdl": (C.c_int, [_vp, C.c_int]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def lib():
    """The loaded shared library; raises RtError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RtError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the environment step)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)       # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if L.rt_abi_version() != ABI_VERSION:
            raise RtError(f"librtenv_b200.so ABI {L.rt_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rt_last_error()
        raise RtError(f"{what or 'rt call'} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().rt_launch_count())
