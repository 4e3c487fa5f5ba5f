#!/usr/bin/env python
"""Benchmark of the batched environment step (BASELINE.json metric: env-steps/sec, device-timed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl ours|reference]

Workload (BASELINE.json configs[1]): visionless vector-env, E = 4096 envs per GPU resident in
HBM, tumour id (i*7919) mod 1000 for the first episode then the device RNG, synthetic actions
uniform(-1, 1).  One "step" is one rt_step over the E envs of a rank (E env-steps), steady
state including the NEXT_STEP autoreset calls.  Multi-GPU: one process per GPU (torchrun),
envs are sharded with no data-path collective (weak scaling: E per GPU fixed).

`value`  — inputs resident in HBM, K rt_step launches replayed from CUDA graphs, timed with CUDA
           events on the launching stream, max over ranks.
`e2e`    — the same metric through the host-buffer C-ABI call (rt_step_host): pinned host actions
           in, pinned host obs/reward/flags out, every step, copies inside the timed region.
`roofline` — step kernel: algorithmic bytes (SURVEY.md §8d sector-granular S per env-step x E)
           over the average launch duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / --impl reference — the UNMODIFIED Python reference (oracle/_ref, copied by oracle/build_ref.py;
           one process per host core, one env each) on a bounded sample, `kind: "reference"`; the C port
           (oracle/rt_oracle.c) is timed beside it (`port`).  Without oracle/_ref the port alone, `kind: "port"`.
`ppo`    — BASELINE.json configs[2]: PPO (MLP agent) on 8,192 envs per GPU x 128 steps per iteration, fused rollout,
           NCCL flat-gradient all-reduce per optimiser step when N > 1; steady-state env-steps/s, CUDA-event time per
           all_reduce call, rollout / update split, replicas_identical.
`dense`  — BASELINE.json configs[4]: RT_FLAG_DENSE full-volume dose update at 1,024 envs per GPU, HBM fraction.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

# SURVEY.md §8(d), sparse visionless step under a uniform-random policy (re-measured by
# tools/measure_traffic.py): payload P = 202 + 8*U + 4*W bytes, sector-granular
# S = 256 + 64*Sec + 32*Sec_ep/100 bytes per env-step.
ALGO_BYTES_PAYLOAD = 1690.0     # P = 202 + 8*116.6 + 4*138.7
ALGO_BYTES_SECTOR = 5637.0      # S = 256 + 64*59.1 + 32*4994/100
GRAPH_CHUNK = 50


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 500))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-kernels", action="store_true", help="skip the dense / voxel-obs / GAE side measurements")
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO leg (BASELINE configs[2])")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-mode leg (BASELINE configs[4])")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of CUDA graphs")
    ap.add_argument("--action-pool", type=int, default=128, help="distinct device-resident action batches cycled through")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch from the committed ncu capture
    (profiles/step_traffic.json, written from the --set full report; cold-cache figures, per launch)."""
    p = os.path.join(REPO, "profiles", "step_traffic.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["dram_read_bytes"]) + float(d["dram_write_bytes"])
    return None


def tensor_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        for k in ("bf16_tflops_sustained", "bf16_tflops", "bf16_tflops_burst"):
            if k in d:
                return float(d[k])
    return 2250.0


def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
def cpu_rate(n_envs: int, T: int, threads: int, seed: int = 0):
    """env-steps/s of the CPU oracle port on `threads` host threads over n_envs x T steps."""
    from oracle import oracle as O
    ph = O.Phantom()
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, (T, n_envs, 6)).astype(np.float32)
    tids = ((np.arange(n_envs) * 7919) % 1000).astype(np.int32)[None, :]
    t0 = time.perf_counter()
    O.rollout(ph, tids, acts, threads=threads)
    dt = time.perf_counter() - t0
    return n_envs * T / dt, dt


def workload_config(E: int, world: int, dose_gb=None):
    """The `config` object of the JSON line: the same dict in both arms (ours / reference)."""
    return {
        "workload": f"visionless vector-env, {E} envs/GPU resident in HBM (BASELINE.json configs[1]), "
                    "uniform(-1,1) actions, tumour id (i*7919) mod 1000 then seeded RNG, NEXT_STEP autoreset calls "
                    "counted in autoreset_calls_in_timed_region",
        "envs_per_gpu": E, "total_envs": E * world, "parallelism": f"env-sharded x{world}, no data-path collective",
        "l2": "no flush: the dose state (7.1 GB/GPU at 4096 envs) >> 126 MB L2 and every step touches cells not "
              "touched before in the episode; the 0.5 MB of env records are L2-resident by design",
    }


def port_baseline(target_s: float = 8.0):
    cores = os.cpu_count() or 1
    rate, _ = cpu_rate(cores, 20, cores)                       # calibration
    T = 101
    n = int(max(cores, min(4096, rate * target_s / T)))
    n = (n // cores) * cores
    rate, dt = cpu_rate(n, T, cores, seed=1)
    return {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} envs x {T} calls of configs[1] (oracle/rt_oracle.c, dense volumes as the reference "
                      f"holds them), {dt:.1f} s on {cores} threads"}


def reference_rate(processes: int, n_steps: int, warm: int = 3, envs_per_process: int = 1):
    """The unmodified Python reference: `processes` vector-env loops side by side (SURVEY.md 8d CPU baseline)."""
    from oracle import ref_runtime as R
    rate, wall, resets = R.measure(processes, n_steps, warm=warm, envs_per_process=envs_per_process)
    return rate, wall, resets


def cpu_baseline(target_s: float = 12.0):
    """Bounded CPU sample for the `cpu_baseline` key of our line: the Python reference on every host core (and on
    one), the C port beside it."""
    from oracle import ref_runtime as R
    cores = os.cpu_count() or 1
    port = port_baseline()
    if not R.available():
        port["note"] = "oracle/_ref (Python reference) or SciPy absent on this box: C port only"
        return port
    r1, w1, _ = reference_rate(1, 300)                          # reference single-env CPU path
    n_steps = int(max(101, min(2000, r1 * 0.6 * target_s)))     # per process; all cores are slower per core than one
    rp, wp, resets = reference_rate(cores, n_steps)
    return {"value": rp, "unit": "env-steps/s", "cores": cores, "kind": "reference",
            "sample": f"unmodified environment.py/draw_line.py/transforms.py (oracle/_ref): {cores} processes x 1 env x "
                      f"{n_steps} vector-env calls incl. {resets} autoreset calls, {wp:.1f} s; BLAS threads 1",
            "single_process": {"value": r1, "cores": 1, "sample": f"1 process x 1 env x 300 calls, {w1:.1f} s"},
            "port": port}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_runtime as R
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if R.available():
        # a "step" = one vector-env call over a bounded sample of n = cores x m envs of the 4096-env workload (each
        # process steps its m envs serially, as SyncVectorEnv does); m and the timed steps are sized for ~30 s
        r1, _, _ = reference_rate(1, 200)
        per_proc = r1 * 0.5                                       # all cores busy: slower per core than one process
        m = int(max(1, min(32, per_proc * 30.0 / max(K + W, 1))))
        k_run = int(max(1, min(K, per_proc * 100.0 / m)))
        rate, wall, resets = reference_rate(cores, k_run, warm=max(1, min(W, 20)), envs_per_process=m)
        kind = "reference"
        n = cores * m
        sample = (f"unmodified Python reference (oracle/_ref: environment.py, draw_line.py, transforms.py): {n} of "
                  f"{args.envs_per_gpu} envs per step ({cores} processes, one per host core, x {m} envs), {k_run} of {K} "
                  f"steps timed ({resets} autoreset calls), {wall:.1f} s")
        port = port_baseline()
        dt_per_step = wall / k_run
    else:
        rate0, _ = cpu_rate(cores, 20, cores)
        n = int(max(cores, min(args.envs_per_gpu, rate0 * 120.0 / max(K + W, 1))))
        n = max(cores, (n // cores) * cores)
        cpu_rate(n, max(W, 1), cores, seed=2)                  # warm-up (page-in, thread pool)
        rate, dt = cpu_rate(n, K, cores, seed=3)
        kind, port = "port", None
        sample = (f"{n} of {args.envs_per_gpu} envs per step x {K} steps (oracle/rt_oracle.c port of environment.py/"
                  f"draw_line.py/transforms.py; oracle/_ref or SciPy absent on this box)")
        dt_per_step = dt / K
    cb = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample}
    if port is not None:
        cb["port"] = port
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": dt_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": workload_config(args.envs_per_gpu, world),
        "envs_per_step_sample": n,
        "note": "one host runs this arm whatever N is: at N > 1 the ratio compares N GPUs x 4096 envs with the CPUs of one box",
        "cpu_baseline": cb,
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def other_kernels(rt, dev, peak):
    """The streaming kernels of the path, each timed alone with CUDA events on working sets > L2."""
    import torch

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) * 1e-3 / reps

    out = []
    g = torch.Generator(device=dev).manual_seed(7)
    # voxel observation (BASELINE configs[3]): 4 planes written + dose read, 4,033,400 B per env
    n = 256
    ve = rt.BatchedEpisodes(n, device=dev, seed=4)
    ve.reset()
    a = torch.rand((n, 6), device=dev, generator=g) * 2 - 1
    for _ in range(20):
        ve.step(a, want_info=False)
    vol = torch.empty((n, 4) + ve.grid, dtype=torch.float32, device=dev)
    s = timed(lambda: ve.volumes(0, n, out=vol), 10)
    b = n * 5 * ve.nvox * 4
    moved = n * (4 * ve.nvox * 4 + 1723392)          # 4 planes written + the bricked 8-byte dose cells read
    out.append({"kernel": "rt_volumes_kernel", "workload": f"{n} envs, (4,67,43,70) float32 observations",
                "bytes": b, "us": s * 1e6, "achieved": b / s / 1e9, "unit": "GB/s", "frac": b / s / 1e9 / peak,
                "env_steps_per_s": n / s, "bytes_moved": moved, "moved_gbs": moved / s / 1e9,
                "note": "algorithmic bytes count the dose as 4 bytes per voxel (SURVEY 8d); the sparse-mode state keeps 8-byte cells "
                        "{dose, generation}, so the kernel actually moves bytes_moved"})
    ve.close()
    del vol
    # GAE (train.py:164-181): T=128, N=65536 (BASELINE configs[2] size), 20 B per element, two rotating sets > L2
    T, N = 128, 65536
    sets = [[torch.randn((T, N), device=dev, generator=g) for _ in range(3)] + [torch.empty((T, N), device=dev) for _ in range(2)]
            for _ in range(2)]
    nv, nd = torch.randn(N, device=dev, generator=g), torch.zeros(N, device=dev)
    k = [0]

    def gae():
        r, v, d, oa, ob = sets[k[0] % 2]
        k[0] += 1
        rt.compute_gae(r, v, d, nv, nd, 0.99, 0.95, out=(oa, ob))
    s = timed(gae, 10)
    b = T * N * 20
    out.append({"kernel": "rt_gae_kernel", "workload": f"T={T}, N={N}", "bytes": b, "us": s * 1e6,
                "achieved": b / s / 1e9, "unit": "GB/s", "frac": b / s / 1e9 / peak})
    del sets
    # FeaturesExtractor3D (BASELINE configs[3], 1024 voxel observations): the tcgen05 blocks and the whole forward
    n = 1024
    tpeak = tensor_peak()
    x = torch.rand((n, 4) + (67, 43, 70), device=dev, generator=g)
    fe = rt.FeaturesExtractor3D((4, 67, 43, 70), 64, compute_dtype=torch.bfloat16).to(dev)
    with torch.no_grad():
        s = timed(lambda: fe._fused_first_block(x), 10)
        fl = n * 2 * 313.0e6                                     # conv1 313 MMAC per sample
        out.append({"kernel": "rt_conv1_tc_kernel (tcgen05)", "workload": f"{n} samples, Conv3d(4->16,k3)+ReLU+MaxPool fused",
                    "flops": fl, "us": s * 1e6, "achieved": fl / s / 1e12, "unit": "TFLOP/s", "bound": "tensor",
                    "frac": fl / s / 1e12 / tpeak, "peak": tpeak, "us_per_sample": s * 1e6 / n,
                    "hbm_gbs": n * (4 * 201670 * 4 + 16 * 33 * 21 * 34 * 2) / s / 1e9})
        s = timed(lambda: fe(x), 10)
        fl = n * 0.761e9
        out.append({"kernel": "FeaturesExtractor3D forward (conv1, conv2 on tcgen05; conv3, linear cuDNN/cuBLAS)",
                    "workload": f"{n} samples", "flops": fl, "us": s * 1e6, "achieved": fl / s / 1e12, "unit": "TFLOP/s",
                    "bound": "tensor", "frac": fl / s / 1e12 / tpeak, "peak": tpeak, "us_per_sample": s * 1e6 / n})
    # rollout body of the PPO loop for the MLP agent (train.py:139-149): policy + value forward, sampling, log-prob and
    # rollout-buffer rows in one kernel; float32 FFMA, 2 * 10,061 flop per env against the FFMA peak of the SMs
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fpeak = sms * 128 * 2 * 1.965e9 / 1e12
    agent = rt.PPO((9,), (6,), 64).to(dev)
    for n in (8192, 65536):
        fr = rt.FusedRollout(agent, n, 4, seed=1)
        o = torch.rand((n, 9), device=dev, generator=g) * 2 - 1
        s = timed(lambda: fr.act(o), 50)
        fl = n * 2.0 * 10061
        out.append({"kernel": "rt_ppo_act_kernel", "workload": f"{n} envs, PPO MLP 9-64-64-{{6,1}} forward + Gaussian sample + log-prob",
                    "flops": fl, "us": s * 1e6, "achieved": fl / s / 1e12, "unit": "TFLOP/s", "bound": "fp32",
                    "frac": fl / s / 1e12 / fpeak, "peak": fpeak, "peak_source": "SMs x 128 FFMA x 2 x 1.965 GHz"})
        del fr
    # Last, because the 113 GB handle it allocates and frees leaves the next measurements of the process 20-100 % slower
    # (volumes 581 us against 262 us, 4096-env step 15.2 us against 12.2 us, measured with tools/stepbench.py).
    # the step kernel again at 65,536 envs on this GPU (16 waves): its throughput once launch latency and the
    # per-env dependent chain are amortised ("roofline fraction per kernel at large N")
    n = 65536
    # one distinct action batch per call of an episode cycle: a short repeating pool would walk every beam into the
    # bounds of the volume (a lighter workload than the uniform-random policy the algorithmic bytes were counted on)
    n_pool = 104
    acts = [torch.rand((n, 6), device=dev, generator=g) * 2 - 1 for _ in range(n_pool)]
    b = n * ALGO_BYTES_SECTOR
    for name in ("rt_step_kernel",):
        se = rt.BatchedEpisodes(n, device=dev, seed=11)
        se.reset()
        for i in range(127):                       # into the second episode (timed() adds three more warm-up calls)
            se.step(acts[i % n_pool], want_info=False)
        k = [0]

        def big_step():
            se.step(acts[k[0] % n_pool], want_info=False)
            k[0] += 1
        s = timed(big_step, 101)                   # one full episode cycle incl. the autoreset call: steady-state mix of
                                                   # first-touched and re-touched cells
        out.append({"kernel": name, "workload": f"{n} envs (visionless sparse step)", "bytes": b, "us": s * 1e6,
                    "achieved": b / s / 1e9, "unit": "GB/s", "frac": b / s / 1e9 / peak, "env_steps_per_s": n / s,
                    "state_gb": se.device_bytes / 1e9})
        se.close()
    del acts
    return out


def ppo_leg(rt, torch, dist, dev, world, rank, envs_per_gpu=8192, num_steps=128, iterations=8, steady=5):
    """BASELINE.json configs[2]: CleanRL-style PPO (reference train.py:91-282) with the MLP agent on `envs_per_gpu`
    envs per rank x `num_steps` steps per iteration, fused rollout step, one NCCL all-reduce of the flat gradient per
    optimiser step (reference train.py:246-247 is where it goes).  Steady state = the last `steady` iterations, timed
    with CUDA events from the first rollout step to the last optimiser step, max over ranks."""
    from ppo_radiotherapy_b200.train import load_config, train
    total = envs_per_gpu * world
    cfg = load_config(None, num_envs=total, num_steps=num_steps, num_minibatches=4, update_epochs=2,
                      total_timesteps=total * num_steps * iterations, num_saves=0, save_model=False, seed=1, visionless=True)
    torch.manual_seed(1 + rank)
    prof = {}
    agent = train(cfg, None, dev, None, "bench", log=None, profile=prof)
    flat = torch.cat([p.detach().reshape(-1) for p in agent.parameters()])
    same = True
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, 0)
        t = torch.tensor([int(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        same = bool(t.item())
    it = torch.tensor([sum(prof["iter_ms"][-steady:]), sum(prof["rollout_ms"][-steady:]), sum(prof["update_ms"][-steady:])],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(it, op=dist.ReduceOp.MAX)
    ar = sorted(prof["allreduce_us"][len(prof["allreduce_us"]) // 2:])
    h = agent.history
    out = {
        "workload": f"PPO, MLP agent, {total} envs = {world} x {envs_per_gpu}, {num_steps} steps per iteration, 2 epochs x 4 "
                    f"minibatches (BASELINE.json configs[2]); {steady} steady iterations of {iterations}",
        "value": steady * total * num_steps / (float(it[0]) * 1e-3), "unit": "env-steps/s",
        "iter_ms": float(it[0]) / steady, "rollout_ms": float(it[1]) / steady, "update_ms": float(it[2]) / steady,
        "fused_rollout": bool(prof.get("fused_rollout")), "rollout_one_launch": bool(prof.get("rollout_kernel")),
        "allreduce": None if world == 1 else {
            "calls_per_iteration": len(prof["allreduce_us"]) // max(1, iterations), "elements": int(flat.numel()),
            "bytes": int(flat.numel()) * 4, "us_median": ar[len(ar) // 2] if ar else None,
            "us_p90": ar[int(len(ar) * 0.9)] if ar else None, "backend": "nccl", "timed": "CUDA events around dist.all_reduce"},
        "replicas_identical": same,
        "episodic_return_first_last": [h[0].get("episodic_return"), h[-1].get("episodic_return")],
    }
    del agent
    return out


def dense_leg(rt, torch, dist, dev, world, rank, peak, n=1024, reps=20):
    """BASELINE.json configs[4]: dose-grid stress — RT_FLAG_DENSE handles read and write every dose volume every step
    (the reference's own dataflow, environment.py:107-110, 164-191), random tumours from the bundled table, `n` envs
    per GPU, no collective.  CUDA events, max over ranks."""
    de = rt.BatchedEpisodes(n, device=dev, dense=True, seed=3 + rank)
    de.reset()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    acts = [torch.rand((n, 6), device=dev, generator=g) * 2 - 1 for _ in range(8)]
    for i in range(3):
        de.step(acts[i], want_info=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        de.step(acts[i % 8], want_info=False)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = float(t[0]) * 1e-3 / reps
    b = n * 2 * de.nvox * 4
    de.close()
    return {"workload": f"dense mode (RT_FLAG_DENSE), {n} envs/GPU x {world} GPUs, full-volume dose update + from-scratch "
                        "reductions every step (BASELINE.json configs[4])",
            "value": n * world / s, "unit": "env-steps/s", "us_per_step": s * 1e6, "bytes_per_gpu_step": b,
            "achieved_gbs_per_gpu": b / s / 1e9, "frac": b / s / 1e9 / peak, "working_set": f"{b / 2e9:.2f} GB/GPU >> L2"}


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the CPU arm forks worker processes: run it before this process creates a CUDA context
    cpu_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_line = cpu_baseline()

    import torch
    import torch.distributed as dist
    import ppo_radiotherapy_b200 as rt
    from ppo_radiotherapy_b200 import _native as nat
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    E, K, W = args.envs_per_gpu, args.steps, args.warmup
    eng = rt.BatchedEpisodes(E, device=dev, seed=1234 + rank)
    first = (((np.arange(E) + rank * E) * 7919) % 1000).astype(np.int32)
    eng.set_tumour_schedule(first[None, :])
    eng.reset()
    eng.set_tumour_schedule(None)                 # later episodes: device RNG
    # steps per captured graph: 50, or all K timed steps when fewer are asked for (so that a short run is not measured
    # on eager launches)
    chunk = GRAPH_CHUNK if args.steps >= GRAPH_CHUNK else max(1, args.steps)
    n_act = max(chunk, args.action_pool) if not args.no_graph else max(1, args.action_pool)
    gen = torch.Generator(device=dev).manual_seed(rank)
    act_pool = torch.rand((n_act, E, 6), device=dev, generator=gen) * 2 - 1
    stream = torch.cuda.Stream(dev)
    step_idx = 0

    def eager(k):
        nonlocal step_idx
        for _ in range(k):
            eng.step(act_pool[step_idx % n_act], want_info=False)
            step_idx += 1

    # CUDA graphs of GRAPH_CHUNK consecutive steps (distinct action batches); pool of n_act/chunk graphs
    graphs = []
    with torch.cuda.stream(stream):
        eager(min(W, 8))
        stream.synchronize()
        if not args.no_graph:
            n_graphs = max(1, n_act // chunk)
            for gi in range(n_graphs):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    for j in range(chunk):
                        eng.step(act_pool[(gi * chunk + j) % n_act], want_info=False)
                graphs.append(g)
            # the K % chunk steps left over at the end of the timed region get a graph of their own
            rem_graph = None
            if K % chunk:
                rem_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(rem_graph, stream=stream):
                    for j in range(K % chunk):
                        eng.step(act_pool[j % n_act], want_info=False)
            # first replay of a graph uploads it: do that outside the timed region (these steps add to the warm-up)
            for g in graphs + ([rem_graph] if rem_graph is not None else []):
                g.replay()
            stream.synchronize()

        def run_steps(k, timed=False):
            nonlocal step_idx
            if not graphs:
                eager(k)
                return k
            launched, gi = 0, 0
            while k - launched >= chunk:
                graphs[gi % len(graphs)].replay()
                gi += 1
                launched += chunk
            if timed and rem_graph is not None and k - launched == K % chunk:
                rem_graph.replay()
            else:
                eager(k - launched)
            return k

        run_steps(W)
        barrier()
        episodes0 = int(eng.counters()[:, 3].sum().item())
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        launches0 = nat.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        run_steps(K, timed=True)
        ev1.record(stream)
        stream.synchronize()
        barrier()
        ms = ev0.elapsed_time(ev1)
        eager_launches = nat.launch_count() - launches0
        # every autoreset call of an env starts an episode: calls of the timed region that were resets, per env
        autoreset_calls = (int(eng.counters()[:, 3].sum().item()) - episodes0) / float(E)
        # kernels per rt_step call (1 for the fused step kernel, 2 for the pose + deposit variant), counted by the
        # library on one eager call after the timed region; graph replays re-issue the captured launches
        l0 = nat.launch_count()
        eager(1)
        stream.synchronize()
        per_step = nat.launch_count() - l0
        gpu_launches = K * per_step

        # ---- e2e: host buffers through rt_step_host, every step --------------------------------
        Ke = args.e2e_steps or min(K, 500)
        h_act = torch.empty((n_act, E, 6), dtype=torch.float32, pin_memory=True)
        h_act.copy_(act_pool)
        h_obs = torch.empty((E, 9), dtype=torch.float32, pin_memory=True)
        h_rew = torch.empty(E, dtype=torch.float64, pin_memory=True)
        h_term = torch.empty(E, dtype=torch.uint8, pin_memory=True)
        h_trunc = torch.empty(E, dtype=torch.uint8, pin_memory=True)
        ho, hr, ht, hu = h_obs.numpy(), h_rew.numpy(), h_term.numpy(), h_trunc.numpy()
        calls = [eng.bind_step_host(h_act[i].numpy(), ho, hr, ht, hu) for i in range(n_act)]
        for i in range(5):
            calls[i]()
        barrier()
        t0 = time.perf_counter()
        acc = 0.0
        for i in range(Ke):
            calls[i % n_act]()                      # host actions in, host obs/reward/flags out, synchronous
            acc += float(hr[0])                     # the host reads the step's result
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        barrier()
        clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])

    if rank == 0:
        total_steps = float(K) * E * world
        value = total_steps / (ms_max * 1e-3)
        peak, peak_src = hbm_peak()
        launch_s = ms_max * 1e-3 / K
        achieved = ALGO_BYTES_SECTOR * E / launch_s / 1e9
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32/f64", "data": "synthetic",
            "config": workload_config(E, world),
            "launch": "eager" if not graphs else f"CUDA graphs of {chunk} steps",
            "autoreset_calls_in_timed_region": autoreset_calls,
            "state_gb_per_gpu": eng.device_bytes / 1e9,
            "roofline": {
                "kernel": "rt_step_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic() if E == 4096 else None, "peak_source": peak_src,
                "bytes_per_env_step": ALGO_BYTES_SECTOR, "payload_bytes_per_env_step": ALGO_BYTES_PAYLOAD,
                "avg_launch_us": launch_s * 1e6,
            },
            "e2e": {"value": float(Ke) * E * world / (e2e_ms_max * 1e-3), "unit": "env-steps/s",
                    "h2d_bytes_per_step": E * 6 * 4, "d2h_bytes_per_step": E * (9 * 4 + 8 + 1 + 1),
                    "steps": Ke, "path": "rt_step_host (pinned host buffers mapped into the kernel)"},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
        }
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
    eng.close()
    del act_pool, graphs
    torch.cuda.empty_cache()
    peak_all, _ = hbm_peak()
    # BASELINE.json configs[2] and configs[4] at this N (every rank takes part; rank 0 reports)
    ppo = None if args.no_ppo else ppo_leg(rt, torch, dist, dev, world, rank)
    dense = None if args.no_dense else dense_leg(rt, torch, dist, dev, world, rank, peak_all)
    if rank == 0:
        line["ppo"] = ppo
        line["dense"] = dense
        if world == 1 and not args.no_other_kernels:
            line["other_kernels"] = other_kernels(rt, dev, peak)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
