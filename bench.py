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
`cpu_baseline` / --impl reference — the CPU oracle port of the reference path (oracle/rt_oracle.c,
           the Python reference cannot travel to the GPU box) on all host cores, bounded sample.
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


def cpu_baseline(target_s: float = 15.0):
    cores = os.cpu_count() or 1
    rate, _ = cpu_rate(cores, 20, cores)                       # calibration
    T = 101
    n = int(max(cores, min(4096, rate * target_s / T)))
    n = (n // cores) * cores
    rate, dt = cpu_rate(n, T, cores, seed=1)
    return {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} envs x {T} calls of configs[1] (oracle/rt_oracle.c, dense volumes as the reference "
                      f"holds them), {dt:.1f} s on {cores} threads"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    rate, _ = cpu_rate(cores, 20, cores)
    # K timed "steps", each over a bounded sample of n envs of the 4096-env workload, sized for <= ~120 s
    n = int(max(cores, min(args.envs_per_gpu, rate * 120.0 / max(K + W, 1))))
    n = max(cores, (n // cores) * cores)
    cpu_rate(n, max(W, 1), cores, seed=2)                      # warm-up (page-in, thread pool)
    rate, dt = cpu_rate(n, K, cores, seed=3)
    sample = (f"{n} of {args.envs_per_gpu} envs per step x {K} steps (oracle/rt_oracle.c port of environment.py/"
              f"draw_line.py/transforms.py; the Python reference cannot travel to the GPU box)")
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": "visionless vector-env, 4096 envs/GPU (BASELINE.json configs[1]); CPU port on a "
                               f"bounded sample of {n} envs per step", "envs_per_step_sample": n},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
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
    # the step kernel again at 65,536 envs on this GPU (16 waves): its throughput once launch latency and the
    # per-env dependent chain are amortised ("roofline fraction per kernel at large N")
    n = 65536
    # one distinct action batch per call of an episode cycle: a short repeating pool would walk every beam into the
    # bounds of the volume (a lighter workload than the uniform-random policy the algorithmic bytes were counted on)
    n_pool = 104
    acts = [torch.rand((n, 6), device=dev, generator=g) * 2 - 1 for _ in range(n_pool)]
    b = n * ALGO_BYTES_SECTOR
    # ... and the two-kernel variant of the step (rt_step_split.cuh, RT_STEP_KB=-2: thread-per-env pose kernel +
    # persistent warp-per-env deposit kernel), which exists for this regime
    for name, kb in (("rt_step3_kernel", None), ("rt_split_pose_kernel + rt_split_deposit_kernel (RT_STEP_KB=-2)", "-2")):
        old_kb = os.environ.get("RT_STEP_KB")
        if kb is not None:
            os.environ["RT_STEP_KB"] = kb
        try:
            se = rt.BatchedEpisodes(n, device=dev, seed=11)
        finally:
            if kb is not None:
                if old_kb is None:
                    del os.environ["RT_STEP_KB"]
                else:
                    os.environ["RT_STEP_KB"] = old_kb
        se.reset()
        for i in range(127):                       # into the second episode (timed() adds three more warm-up calls)
            se.step(acts[i % n_pool], want_info=False)
        k = [0]

        def big_step():
            se.step(acts[k[0] % n_pool], want_info=False)
            k[0] += 1
        s = timed(big_step, 101)                   # one full episode cycle incl. the autoreset call: steady-state mix of
                                                   # fresh and re-touched sectors
        out.append({"kernel": name, "workload": f"{n} envs (visionless sparse step)", "bytes": b, "us": s * 1e6,
                    "achieved": b / s / 1e9, "unit": "GB/s", "frac": b / s / 1e9 / peak, "env_steps_per_s": n / s})
        se.close()
    del acts
    # dense-mode step (BASELINE configs[4]): read + write every dose volume, 1,613,360 B per env-step
    n = 1024
    de = rt.BatchedEpisodes(n, device=dev, dense=True, seed=3)
    de.reset()
    a = torch.rand((n, 6), device=dev, generator=g) * 2 - 1
    s = timed(lambda: de.step(a, want_info=False), 10)
    b = n * 2 * de.nvox * 4
    out.append({"kernel": "rt_step_kernel<dense> + rt_dense_kernel", "workload": f"{n} envs, full-volume dose update",
                "bytes": b, "us": s * 1e6, "achieved": b / s / 1e9, "unit": "GB/s", "frac": b / s / 1e9 / peak,
                "env_steps_per_s": n / s})
    de.close()
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
    out.append({"kernel": "rt_volumes_kernel", "workload": f"{n} envs, (4,67,43,70) float32 observations",
                "bytes": b, "us": s * 1e6, "achieved": b / s / 1e9, "unit": "GB/s", "frac": b / s / 1e9 / peak,
                "env_steps_per_s": n / s})
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
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ppo_radiotherapy_b200 as rt
    from ppo_radiotherapy_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
        # kernels per rt_step call (1 for the fused step kernel, 2 for the pose + deposit variant), counted by the
        # library on one eager call after the timed region; graph replays re-issue the captured launches
        l0 = nat.launch_count()
        eager(1)
        stream.synchronize()
        per_step = nat.launch_count() - l0
        gpu_launches = K * per_step
        step_kernel = "rt_step3_kernel" if per_step == 1 else "rt_split_pose_kernel + rt_split_deposit_kernel"

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
            "config": {
                "workload": f"visionless vector-env, {E} envs/GPU resident in HBM (BASELINE.json configs[1]), "
                            "uniform(-1,1) actions, autoreset included",
                "envs_per_gpu": E, "total_envs": E * world, "parallelism": f"env-sharded x{world}, no data-path collective",
                "launch": "eager" if not graphs else f"CUDA graphs of {chunk} steps",
                "l2": f"no flush: dose state {eng.device_bytes / 1e9:.2f} GB/GPU >> 126 MB L2 and every step writes "
                      "sectors not touched before in the episode; the env records (0.5 MB) and the sector-valid "
                      "bitmaps (13 MB) are L2-resident by design",
            },
            "roofline": {
                "kernel": step_kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
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
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
    eng.close()
    if rank == 0:
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
