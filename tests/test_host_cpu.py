"""CPU-side checks: the C-ABI library loads and exports every symbol include/rt_env.h declares,
the host logic (info merging, phantom packing) behaves, and the product refuses to run without
its CUDA extension / device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import _native as nat
from ppo_radiotherapy_b200.vector_env import build_infos

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(REPO, "include", "rt_env.h")).read()
    return re.findall(r"RT_API\s+[\w\s\*]+?\b(rt_\w+)\s*\(", text)


def test_library_exports_every_header_symbol():
    syms = _header_symbols()
    assert len(syms) >= 20 and len(set(syms)) == len(syms)
    lib = ctypes.CDLL(nat.build())
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in rt_env.h but not exported"
    # and the Python binding table covers exactly the header
    assert set(nat.EXPORTS) == set(syms)
    assert nat.lib().rt_abi_version() == nat.ABI_VERSION


def test_header_constants_match_binding():
    text = open(os.path.join(REPO, "include", "rt_env.h")).read()
    defs = dict(re.findall(r"#define\s+(RT_\w+)\s+(\d+)u?\b", text))
    assert int(defs["RT_ACTION_SIZE"]) == nat.ACTION_SIZE
    assert int(defs["RT_OBS_SIZE"]) == nat.OBS_SIZE
    assert int(defs["RT_INFO_SIZE"]) == nat.INFO_SIZE
    assert int(defs["RT_BEAM_CAP"]) == nat.BEAM_CAP
    assert int(defs["RT_MAX_TIME_STEPS"]) == nat.MAX_TIME_STEPS
    assert int(defs["RT_ABI_VERSION"]) == nat.ABI_VERSION


def test_invalid_arguments_are_reported_without_a_gpu():
    L = nat.lib()
    assert L.rt_create(None, 0, 4, 0, None) < 0
    assert b"NULL" in L.rt_last_error()
    grid = (ctypes.c_int32 * 3)(67, 43, 70)
    assert L.rt_beam_voxels(grid, None, None, 1, 288, None, None, None, None) < 0
    bad = (ctypes.c_int32 * 3)(67, 43, 700)
    assert L.rt_beam_voxels(bad, None, None, 1, 288, None, None, None, None) < 0
    assert b"grid" in L.rt_last_error()
    assert L.rt_gae(None, None, None, None, None, 1, 1, 0.99, 0.95, None, None, None) < 0
    # empty batches are valid and need no buffers
    assert L.rt_gae(None, None, None, None, None, 0, 4, 0.99, 0.95, None, None, None) == 0
    assert L.rt_beam_voxels(grid, None, None, 0, 288, None, None, None, None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(rt.RtError, match="no CPU fallback"):
        rt.RadiotherapyVectorEnv(4)
    with pytest.raises(rt.RtError):
        rt.beam_voxels(np.zeros((4, 4, 4), np.float32), [1, 1, 1], [0, 1, 0])
    with pytest.raises(rt.RtError):
        rt.compute_gae(torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2), torch.zeros(2), 0.99, 0.95)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "ppo-radiotherapy_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.lower() or f == "__none__", f"{f} mentions the oracle"


def test_phantom_table():
    ph = rt.default_phantom()
    assert tuple(ph.grid) == (67, 43, 70) and ph.n_tumours == 1000
    lungs = ph.lungs_volume()
    assert lungs.sum() == 74029
    sizes = np.diff(ph.vox_offsets)
    assert sizes.min() == 16 and sizes.max() == 586
    assert ph.names == sorted(ph.names)
    # from_volumes reproduces the packed constants (the reference's own expressions)
    ids = [0, 17, int(np.argmax(sizes))]
    re_ph = rt.Phantom.from_volumes(lungs, [ph.tumour_volume(i) for i in ids])
    for j, i in enumerate(ids):
        assert np.array_equal(re_ph.tumour_voxels(j), ph.tumour_voxels(i))
        assert np.array_equal(re_ph.centroid[j], ph.centroid[i])
        assert re_ph.tumour_sum[j] == ph.tumour_sum[i] and re_ph.lung_mask_sum[j] == ph.lung_mask_sum[i]
    assert np.array_equal(re_ph.lungs_bits, ph.lungs_bits)
    with pytest.raises(ValueError):
        rt.Phantom.from_volumes(lungs, [np.zeros_like(lungs)])


def test_build_infos_merging():
    n = 5
    info = np.zeros((n, nat.INFO_SIZE))
    info[:, nat.INFO_STEPPED] = [1, 1, 0, 1, 1]
    info[:, nat.INFO_REWARD_TOTAL] = [1.0, 2.0, 9.0, 4.0, 5.0]
    info[:, nat.INFO_EPISODE_RETURN] = [10, 20, 30, 40, 50]
    info[:, nat.INFO_EPISODE_LENGTH] = [100, 7, 0, 100, 3]
    term = np.array([1, 0, 0, 1, 0], dtype=np.uint8)
    infos = build_infos(info, term, 1.5)
    rc = infos["reward_components"]
    assert np.array_equal(rc["_total"], [True, True, False, True, True])
    assert rc["total"][2] == 0.0 and rc["total"][3] == 4.0
    ep = infos["episode"]
    assert np.array_equal(ep["_r"], [True, False, False, True, False])
    assert np.array_equal(ep["r"], [10, 0, 0, 40, 0]) and np.array_equal(ep["l"], [100, 0, 0, 100, 0])
    assert ep["l"].dtype == np.int64 and np.array_equal(infos["_episode"], ep["_r"])
    # the reference's logger (train.py:42-66) works on it
    assert np.mean(ep["r"][ep["_r"]]) == 25.0 and np.mean(rc["total"][ep["_r"]]) == 2.5
    # no finished env -> no "episode" key (train.py:160); autoreset-only call -> empty dict
    assert "episode" not in build_infos(info, np.zeros(n, np.uint8), 0.0)
    info[:, nat.INFO_STEPPED] = 0
    assert build_infos(info, term, 0.0) == {}


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU leg the driver times beside the CUDA arm): one JSON line with the same
    metric / unit as the CUDA arm on rank 0, nothing and exit status 0 on the other ranks."""
    import json
    import subprocess
    import sys
    bench = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env-steps/sec" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    from oracle import ref_runtime
    # the unmodified Python reference (oracle/_ref, copied by oracle/build_ref.py) when it is there, else the C port
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_runtime.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert set(line["config"]) == {"workload", "envs_per_gpu", "total_envs", "parallelism", "l2"}
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_mlp_params_struct_matches_header():
    """rt_mlp_params (include/rt_env.h) and its ctypes mirror: same fields in the same order, same size; the fused
    rollout refuses agents the kernel does not cover (and anything on the CPU)."""
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "rt_env.h")).read()
    body = re.search(r"typedef struct rt_mlp_params \{(.*?)\} rt_mlp_params;", header, re.S).group(1)
    ptrs = [n for line in re.findall(r"const float \*([^;]+);", body) for n in re.split(r",\s*\*", line)]
    ints = [n.strip() for n in re.search(r"int32_t ([^;]+);", body).group(1).split(",")]
    assert [f[0] for f in nat.MlpParams._fields_] == ptrs + ints
    assert ctypes.sizeof(nat.MlpParams) == 13 * 8 + 16
    from ppo_radiotherapy_b200 import rollout
    assert not rollout.supported(rt.PPO((9,), (6,), 64))            # CPU tensors: there is no CPU path
    assert not rollout.supported(rt.PPO((9,), (6,), 32))
    assert not rollout.supported(torch.nn.Linear(3, 3))
    with pytest.raises(nat.RtError):
        rt.FusedRollout(rt.PPO((9,), (6,), 64), 8, 4)
