"""Multi-GPU paths on real devices (needs >= 2 GPUs; skipped on a single-GPU box): env shards per rank, the NCCL
flat-gradient all-reduce of the PPO loop (reference train.py:246-247 is where it goes), replicas bit-identical, and
the torchrun launcher through train.main().  The same logic runs on CPU with gloo in tests/test_distributed_cpu.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


def _torchrun(args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=REPO)


@needs2
def test_ppo_two_ranks_nccl_replicas_identical():
    out = _torchrun([os.path.join(REPO, "tools", "train_smoke.py"), "512", "mlp", "32"], 29731)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "replicas identical: True" in out.stdout, out.stdout[-2000:]


@needs2
def test_launcher_main_two_ranks(tmp_path):
    """`ppo_radiotherapy_b200_train.py` (the reference's `python train.py --config-file X --output-dir Y`, train.py:285-324)
    under torchrun: two ranks, NCCL, a checkpoint with the reference's state-dict keys at the end."""
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text("exp_name: t\nnum_envs: 256\nnum_steps: 16\nnum_minibatches: 2\nupdate_epochs: 1\ntotal_timesteps: 8192\n"
                   "num_saves: 1\nvisionless: true\n")
    out = _torchrun([os.path.join(REPO, "ppo_radiotherapy_b200_train.py"), "--config-file", str(cfg), "--output-dir", str(tmp_path)],
                    29732)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    models = [os.path.join(r, f) for r, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".model")]
    assert models, out.stdout[-1000:]
    sd = torch.load(models[0], weights_only=True)
    assert {"critic.0.weight", "actor_mean.4.bias", "actor_logstd"} <= set(sd)


@needs2
def test_bench_two_gpus_reports_ppo_and_dense():
    """bench.py --gpus 2: the headline line plus the PPO (configs[2]) and dense (configs[4]) legs, replicas identical."""
    out = _torchrun([os.path.join(REPO, "bench.py"), "--gpus", "2", "--steps", "100", "--warmup", "10"], 29733, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["value"] > 0
    assert line["ppo"]["replicas_identical"] is True and line["ppo"]["allreduce"]["us_median"] > 0
    assert line["dense"]["value"] > 0 and 0 < line["dense"]["frac"] <= 1.05
