"""Run the UNMODIFIED Python reference (oracle/_ref/, see oracle/build_ref.py) as the CPU arm of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py (`--impl reference`, `cpu_baseline`) and tests/.
Nothing here is reachable from the product package.

The reference reads ./data relative to the CWD at import time (environment.py:28-29) and on every reset
(environment.py:94-95), so a run gets a temporary root with the modules and the data it needs:
    <tmp>/environment.py ...            copies from oracle/_ref/ (or /root/reference when that exists)
    <tmp>/data/lungs.npy                bool (67,43,70), expanded from the packed phantom table
    <tmp>/data/tumours/<name>.npy       float32 (67,43,70) {0,1} volumes of the tumours the run uses
gymnasium / stable_baselines3 / matplotlib are inert stubs (oracle/ref_harness.py): they take no part in the
arithmetic.  Workers are separate processes (fork), one env each, BLAS threads pinned to 1.
"""
import multiprocessing as mp
import os
import shutil
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_COPY = os.path.join(_HERE, "_ref")
_REPO = os.path.dirname(_HERE)
_PHANTOM = os.path.join(_REPO, "ppo-radiotherapy_b200", "data", "phantom.npz")
MODULES = ("environment.py", "draw_line.py", "transforms.py", "visualize_voxel.py")


def source_dir():
    """Directory holding the reference modules: oracle/_ref/ (travels to the GPU box), else /root/reference."""
    for d in (_REF_COPY, os.environ.get("RT_REFERENCE_SRC", "/root/reference")):
        if all(os.path.isfile(os.path.join(d, m)) for m in MODULES):
            return d
    return None


def available() -> bool:
    if source_dir() is None:
        return False
    try:
        import scipy.spatial.transform  # noqa: F401  (transforms.py:2)
    except ImportError:
        return False
    return True


def prepare_root(tumour_ids, root=None) -> str:
    """Temporary reference root with the modules and the dense data of `tumour_ids` (indices into the sorted
    file-name list of the packed phantom)."""
    src = source_dir()
    if src is None:
        raise RuntimeError("reference modules not found (run oracle/build_ref.py where /root/reference exists)")
    root = root or tempfile.mkdtemp(prefix="rt_ref_")
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m), os.path.join(root, m))
    z = np.load(_PHANTOM)
    grid = tuple(int(g) for g in z["grid"])
    nvox = int(np.prod(grid))
    os.makedirs(os.path.join(root, "data", "tumours"), exist_ok=True)
    lungs = np.unpackbits(z["lungs_bits"].view(np.uint8), bitorder="little")[:nvox].astype(bool).reshape(grid)
    np.save(os.path.join(root, "data", "lungs.npy"), lungs)
    off, vox, names = z["vox_offsets"], z["vox"], z["names"]
    for t in sorted(set(int(t) for t in tumour_ids)):
        v = np.zeros(nvox, dtype=np.float32)
        v[vox[off[t]:off[t + 1]]] = 1.0
        np.save(os.path.join(root, "data", "tumours", str(names[t])), v.reshape(grid))
    return root


def _worker(root, seed, n_steps, warm, m, barrier, out_q):
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    os.environ["RT_REFERENCE_ROOT"] = root
    from oracle import ref_harness as H
    H.REF_ROOT = root
    ns = H.load()
    os.chdir(root)                                   # environment.py:94 loads ./data/tumours/<file> on every reset
    np.random.seed(seed)                             # environment.py:90 draws from the global NumPy RNG
    envs = [ns.environment.RadiotherapyEnv(visionless=True) for _ in range(m)]
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, (warm + n_steps, m, 6)).astype(np.float32)
    done = [False] * m

    def vector_step(t):                              # gymnasium 1.0.0 SyncVectorEnv.step, NEXT_STEP autoreset
        resets = 0
        for i, env in enumerate(envs):
            if done[i]:
                env.reset(); done[i] = False; resets += 1
            else:
                _, _, done[i], _, _ = env.step(acts[t, i])
        return resets

    for t in range(warm):
        vector_step(t)
    barrier.wait()
    t0 = time.perf_counter()
    resets = 0
    for t in range(warm, warm + n_steps):
        resets += vector_step(t)
    out_q.put((time.perf_counter() - t0, resets))


def measure(processes: int, n_steps: int, warm: int = 3, n_tumours: int = 16, seed: int = 0, envs_per_process: int = 1):
    """`processes` independent copies of the reference's vector-env loop over `envs_per_process` envs each (train.py:93
    steps its envs serially in one process), `n_steps` vector-env calls (a call per env is env.step, or env.reset on
    the call after a terminal step, as SyncVectorEnv does).  Returns (env-steps/s aggregate, wall seconds of the
    slowest worker, autoreset calls)."""
    tids = [(i * 7919) % 1000 for i in range(n_tumours)]
    root = prepare_root(tids)
    m = int(envs_per_process)
    try:
        ctx = mp.get_context("fork")
        barrier = ctx.Barrier(processes)
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(root, seed * 1000 + i, n_steps, warm, m, barrier, q)) for i in range(processes)]
        for p in procs:
            p.start()
        res = [q.get() for _ in procs]
        for p in procs:
            p.join()
    finally:
        shutil.rmtree(root, ignore_errors=True)
    wall = max(r[0] for r in res)
    return processes * m * n_steps / wall, wall, sum(r[1] for r in res)
