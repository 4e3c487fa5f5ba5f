#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING the unmodified reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs
/root/reference); the committed .npz files are what travels.  The reference ships
no golden vectors of its own, so these are its outputs under the
container's NumPy 2.3.5 / SciPy 1.18.1 / torch 2.11 (versions are recorded in each
file).  Re-run:  python oracle/gen_golden.py

  beams.npz  (G1) draw_line.beam_voxels on random + edge-case rays
  poses.npz       transforms.apply_translation / apply_rotation chains
  steps.npz  (G2) RadiotherapyEnv reset/step traces for fixed tumours and actions
  tiny.npz        synthetic few-voxel tumours that terminate early (full irradiation)
  resets.npz (G4) reset observation for every tumour
  gae.npz    (G3) train.py:164-181 on random tensors (CPU torch)
  stress.npz      240 synthetic 16-64-voxel tumours with homing actions: the dose ratio crosses the 0.9
                  termination threshold (environment.py:184-191) on most of them, many within a few ulp
  trajectory.npz  RadiotherapyEnv.export_trajectory (environment.py:69-75): keys, shapes, dtypes, content
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_harness  # noqa: E402
from oracle.hashing import volume_hash  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
G = np.array([67, 43, 70])


def versions():
    import scipy
    import torch
    return np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}", f"torch {torch.__version__}"])


def make_rays(rng, m):
    """Ray classes of SURVEY.md §8(a-4): generic, inside, integer/diagonal ties, axis-aligned,
    near-zero components, positions on the upper bound (== shape), far outside, corners."""
    pos = np.empty((m, 3))
    d = np.empty((m, 3))
    for k in range(m):
        c = k % 8
        p = rng.uniform(-5, 75, 3)
        v = rng.standard_normal(3)
        if c == 1:
            p = rng.uniform(0, 1, 3) * G
        elif c == 2:
            p = np.floor(rng.uniform(0, 1, 3) * G)
            v = np.sign(v) * np.array([1.0, 1.0, 1.0])
            if rng.random() < 0.5:
                v[rng.integers(3)] *= 0.5
        elif c == 3:
            v[rng.integers(3)] = 0.0
            if rng.random() < 0.5:
                v[rng.integers(3)] = 0.0
            if not np.any(v):
                v[rng.integers(3)] = 1.0
            p = rng.uniform(0, 1, 3) * (G - 1)
        elif c == 4:
            v[rng.integers(3)] = rng.choice([1e-7, -1e-7, 9.9e-7, 1.01e-6, 2e-6])
            v[rng.integers(3)] *= 1e-3
            p = rng.uniform(0, 1, 3) * G
        elif c == 5:
            p = np.where(rng.random(3) < 0.5, G.astype(float), rng.uniform(0, 1, 3) * G)
        elif c == 6:
            p = rng.uniform(-200, 300, 3)
        elif c == 7:
            p = np.where(rng.random(3) < 0.5, 0.0, (G - 1).astype(float))
            v = (G / 2 - p) + rng.standard_normal(3) * rng.choice([0.0, 1e-3, 1.0])
            if not np.any(v):
                v = np.array([1.0, 1.0, 1.0])
        pos[k], d[k] = p, v
    return pos, d


def gen_beams(ns, m=24000, n_full=256):
    rng = np.random.default_rng(20240601)
    base = np.zeros(tuple(G), dtype=np.float32)
    pos, d = make_rays(rng, m)
    # rays as the environment produces them: in-volume positions, unit directions at >= pi/4 from axis 0
    env_m = m // 4
    pos[:env_m] = rng.uniform(0, 1, (env_m, 3)) * G
    v = rng.standard_normal((env_m, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[:, 0] = np.clip(v[:, 0], -np.cos(np.pi / 4), np.cos(np.pi / 4))
    d[:env_m] = v / np.linalg.norm(v, axis=1, keepdims=True)
    count = np.zeros(m, dtype=np.int32)
    h = np.zeros(m, dtype=np.uint64)
    full_off = [0]
    full_idx, full_w = [], []
    for k in range(m):
        vol = ns.draw_line.beam_voxels(base, pos[k], d[k])
        assert vol.dtype == np.float32
        flat = vol.reshape(-1)
        nz = np.flatnonzero(flat)
        count[k] = nz.size
        h[k] = volume_hash(nz, flat[nz])
        if k % (m // n_full) == 0:
            full_idx.append(nz.astype(np.int32))
            full_w.append(flat[nz])
            full_off.append(full_off[-1] + nz.size)
    full_ids = np.arange(0, m, m // n_full, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "beams.npz"), pos=pos, dir=d, count=count, hash=h,
                        full_ids=full_ids, full_off=np.array(full_off, dtype=np.int64),
                        full_idx=np.concatenate(full_idx), full_w=np.concatenate(full_w),
                        versions=versions())
    print("beams.npz:", m, "rays; empty:", int((count == 0).sum()), "max voxels:", int(count.max()))


def gen_poses(ns, chains=64, steps=100):
    rng = np.random.default_rng(7)
    acts = rng.uniform(-1.3, 1.3, (chains, steps, 6)).astype(np.float32)
    acts[chains // 2:] = rng.standard_normal((chains - chains // 2, steps, 6)).astype(np.float32)
    pos = np.empty((chains, steps, 3))
    dr = np.empty((chains, steps, 3))
    os_t = np.empty((chains, steps, 3))
    os_r = np.empty((chains, steps))
    E = ns.environment.RadiotherapyEnv
    shape = E.LUNG_SHAPE
    for c in range(chains):
        p = np.array(shape) / 2
        v = np.array([0.0, 1.0, 0.0])
        for s in range(steps):
            a = acts[c, s]
            tr = np.clip(a[:3], -1.0, 1.0) * shape * E.MOVEMENT_SPEED          # environment.py:122-125
            rv = np.clip(a[3:6], -1.0, 1.0) * np.pi * E.ROTATION_SPEED         # environment.py:139-141
            p, ot = ns.transforms.apply_translation(p, tr, shape)
            v, orr = ns.transforms.apply_rotation(v, rv, E.MIN_ANGLE_Z)
            pos[c, s], dr[c, s], os_t[c, s], os_r[c, s] = p, v, ot, orr
    np.savez_compressed(os.path.join(OUT, "poses.npz"), actions=acts, pos=pos, dir=dr,
                        overshoot_t=os_t, overshoot_r=os_r, versions=versions())
    print("poses.npz:", chains, "chains x", steps)


def policy_actions(ns, names, episode_len, seed=3):
    """Actions of the shipped saves/20M.model policy driving the reference env (stochastic, seeded)."""
    import torch
    nets = ref_harness.load_networks()
    agent = nets.PPO((9,), (6,), 64)
    sd = torch.load(os.path.join(ref_harness.REF_ROOT, "saves", "20M.model"), map_location="cpu", weights_only=True)
    agent.load_state_dict(sd)
    agent.eval()
    torch.manual_seed(seed)
    acts = np.zeros((len(names), episode_len, 6), dtype=np.float32)
    for i, name in enumerate(names):
        env = ref_harness.RefEnv(visionless=True, tumour_name=name)
        obs, _ = env.reset()
        for t in range(episode_len):
            with torch.no_grad():
                a = agent.get_action_and_value(torch.Tensor(obs).reshape(1, -1))[0]
            a = a.cpu().numpy()[0].astype(np.float32)
            acts[i, t] = a
            obs, _, done, _, _ = env.step(a)
            if done:
                break
    return acts


def gen_steps(ns, phantom_names, per_kind=10, T=100):
    rng = np.random.default_rng(0)
    tids = [(i * 7919) % 1000 for i in range(3 * per_kind)]
    # make sure the largest and smallest tumours are covered
    sizes = np.load(os.path.join(REPO, "ppo-radiotherapy_b200", "data", "phantom.npz"))["vox_offsets"]
    sizes = np.diff(sizes)
    tids[1] = int(np.argmax(sizes))
    tids[2] = int(np.argmin(sizes))
    tids = np.array(tids, dtype=np.int32)
    E = len(tids)
    acts = np.zeros((E, T, 6), dtype=np.float32)
    acts[:per_kind] = rng.uniform(-1, 1, (per_kind, T, 6)).astype(np.float32)
    acts[per_kind:2 * per_kind] = rng.standard_normal((per_kind, T, 6)).astype(np.float32)
    acts[2 * per_kind:] = policy_actions(ns, [phantom_names[t] for t in tids[2 * per_kind:]], T)
    # a slow-moving sequence that re-irradiates the same voxels until the dose clips at 1.0
    acts[0, :, :] *= 0.02

    rec = np.zeros((E, T, 20))
    done = np.zeros((E, T), dtype=np.int8)
    pose = np.zeros((E, T, 6))
    dose_hash = np.zeros((E, T), dtype=np.uint64)
    dose_nnz = np.zeros((E, T), dtype=np.int32)
    vol_hash = np.zeros((E, 3), dtype=np.uint64)
    vol_steps = np.array([0, 1, 40], dtype=np.int32)
    reset_obs = np.zeros((E, 9))
    length = np.zeros(E, dtype=np.int32)
    final_off = [0]
    final_idx, final_val = [], []
    for e in range(E):
        env = ref_harness.RefEnv(visionless=True, tumour_name=phantom_names[tids[e]])
        obs, info = env.reset()
        assert info == {}
        reset_obs[e] = obs
        for t in range(T):
            obs, reward, dn, trunc, info = env.step(acts[e, t])
            assert trunc is False
            r = rec[e, t]
            r[0:9] = obs
            r[9] = reward
            rc = info["reward_components"]
            assert rc["total"] == reward
            r[10], r[11], r[12] = rc["tumour"], rc["lung"], rc["distance_to_tumour"]
            r[13], r[14] = info["doses"]["tumour"], info["doses"]["lung"]
            r[15:18] = info["overshoot"]["translation"]
            r[18] = info["overshoot"]["rotation"]
            inner = env.env
            mask = inner.lungs * (1 - inner.tumours)
            r[19] = np.sum(inner.dose * mask > inner.LUNG_DOSE_THRESHOLD)
            done[e, t] = dn
            pose[e, t, :3] = inner.beam_position
            pose[e, t, 3:] = inner.beam_direction
            flat = inner.dose.reshape(-1)
            nz = np.flatnonzero(flat)
            dose_hash[e, t] = volume_hash(nz, flat[nz])
            dose_nnz[e, t] = nz.size
            for j, vs in enumerate(vol_steps):
                if t == vs:
                    vol = inner.get_volumes()
                    assert vol.dtype == np.float32 and vol.shape == (4, 67, 43, 70)
                    vf = vol.reshape(-1)
                    vnz = np.flatnonzero(vf)
                    vol_hash[e, j] = volume_hash(vnz, vf[vnz])
            length[e] = t + 1
            if dn:
                break
        flat = env.env.dose.reshape(-1)
        nz = np.flatnonzero(flat)
        final_idx.append(nz.astype(np.int32))
        final_val.append(flat[nz])
        final_off.append(final_off[-1] + nz.size)
    np.savez_compressed(os.path.join(OUT, "steps.npz"), tumour_ids=tids, actions=acts, rec=rec, done=done,
                        pose=pose, dose_hash=dose_hash, dose_nnz=dose_nnz, vol_hash=vol_hash,
                        vol_steps=vol_steps, reset_obs=reset_obs, length=length,
                        final_off=np.array(final_off, dtype=np.int64),
                        final_idx=np.concatenate(final_idx), final_val=np.concatenate(final_val),
                        versions=versions())
    print("steps.npz:", E, "episodes; lengths", length.tolist())


def gen_tiny(ns, T=100):
    """Synthetic few-voxel tumours that CAN be fully irradiated, to pin the early-termination
    branch (environment.py:184-191,220), which the bundled tumours never reach.  The reference
    loads ./data/tumours/<name> relative to the CWD, so a scratch data root is used."""
    import tempfile
    lungs = np.load(os.path.join(ref_harness.REF_ROOT, "data", "lungs.npy"))
    cases = {
        "tiny_a": [(30, 20, 30)],
        "tiny_b": [(40, 25, 36), (40, 25, 37)],
        "tiny_c": [(20, 18, 20), (20, 18, 21), (21, 18, 20), (21, 18, 21)],
        "tiny_d": [(50, 10, 50), (50, 11, 50), (50, 12, 50)],
    }
    root = tempfile.mkdtemp(prefix="rt_tiny_")
    os.makedirs(os.path.join(root, "data", "tumours"))
    np.save(os.path.join(root, "data", "lungs.npy"), lungs)
    rec = np.zeros((len(cases), T, 20))
    done = np.zeros((len(cases), T), dtype=np.int8)
    acts = np.zeros((len(cases), T, 6), dtype=np.float32)
    length = np.zeros(len(cases), dtype=np.int32)
    dose_hash = np.zeros((len(cases), T), dtype=np.uint64)
    vox_off = [0]
    vox = []
    old_root = ref_harness.REF_ROOT
    for e, (name, voxels) in enumerate(cases.items()):
        t = np.zeros(tuple(G), dtype=np.float32)
        for v in voxels:
            t[v] = 1.0
        fname = f"0.0_0.0_0.0_0.01_{name}.npy"
        np.save(os.path.join(root, "data", "tumours", fname), t)
        lin = np.flatnonzero(t.reshape(-1)).astype(np.int32)
        vox.append(lin)
        vox_off.append(vox_off[-1] + lin.size)
        ref_harness.REF_ROOT = root
        try:
            env = ref_harness.RefEnv(visionless=True, tumour_name=fname)
            env.reset()
        finally:
            ref_harness.REF_ROOT = old_root
        target = np.mean(np.array(voxels, dtype=np.float64), axis=0)
        if name == "tiny_c":
            target = target  # between voxels: bilinear weights 1/4 each
        for s in range(T):
            p = env.env.beam_position
            a = np.zeros(6, dtype=np.float32)
            a[:3] = np.clip((target - p) / (G * 0.2), -1, 1).astype(np.float32)
            if name == "tiny_d" and s > 3:
                a[:3] = 0.0
            acts[e, s] = a
            obs, reward, dn, _, info = env.step(a)
            r = rec[e, s]
            r[0:9] = obs
            r[9] = reward
            rc = info["reward_components"]
            r[10], r[11], r[12] = rc["tumour"], rc["lung"], rc["distance_to_tumour"]
            r[13], r[14] = info["doses"]["tumour"], info["doses"]["lung"]
            r[15:18] = info["overshoot"]["translation"]
            r[18] = info["overshoot"]["rotation"]
            inner = env.env
            mask = inner.lungs * (1 - inner.tumours)
            r[19] = np.sum(inner.dose * mask > inner.LUNG_DOSE_THRESHOLD)
            done[e, s] = dn
            flat = inner.dose.reshape(-1)
            nz = np.flatnonzero(flat)
            dose_hash[e, s] = volume_hash(nz, flat[nz])
            length[e] = s + 1
            if dn:
                break
    np.savez_compressed(os.path.join(OUT, "tiny.npz"), names=np.array(list(cases)), vox_off=np.array(vox_off, dtype=np.int32),
                        vox=np.concatenate(vox), actions=acts, rec=rec, done=done, length=length,
                        dose_hash=dose_hash, versions=versions())
    print("tiny.npz: lengths", length.tolist())


def stress_tumours(n=240, seed=5):
    """Tube-shaped tumours in the plane of the start beam (axis-0 index 33; the beam starts at (33.5, 21.5, 35)
    pointing along axis 1 and, through the axis quirk of draw_line.py:88-90, deposits on plane 33 only): L columns
    along axis 1 times 2-4 rows along axis 2, some thinned out; 16-64 voxels each."""
    rng = np.random.default_rng(seed)
    out = []
    for e in range(n):
        L = int(rng.integers(8, 33))
        y0 = int(rng.integers(1, 43 - L - 1))
        rows = [35, 36] if e % 3 else [34, 35, 36, 37][:int(rng.integers(2, 5))]
        idx = sorted(set(((33 * 43 + y) * 70 + c) for y in range(y0, y0 + L) for c in rows))
        if e % 5 == 0:
            idx = sorted(set(idx[::2] + idx[1::4]))
        if len(idx) > 64:
            idx = idx[:64]
        out.append(np.array(idx, dtype=np.int32))
    return out


def gen_stress(ns, T=48):
    """The termination threshold under stress: the reference's own `done` flags and rewards on
    tumours of 16-64 voxels whose dose ratio climbs by ~0.05 per beam and crosses 0.9 around beam 18 — for the
    parked beams exactly at the knife edge (ratios 0.8999999 / 0.9 / 0.90000004), where a different summation
    order of np.sum(dose * tumours) (environment.py:166,186) flips the flag."""
    import shutil
    import tempfile
    lungs = np.load(os.path.join(ref_harness.REF_ROOT, "data", "lungs.npy"))
    tumours = stress_tumours()
    E = len(tumours)
    root = tempfile.mkdtemp(prefix="rt_stress_")
    os.makedirs(os.path.join(root, "data", "tumours"))
    np.save(os.path.join(root, "data", "lungs.npy"), lungs)
    rng = np.random.default_rng(17)
    acts = np.zeros((E, T, 6), dtype=np.float32)
    rec = np.zeros((E, T, 20))
    done = np.zeros((E, T), dtype=np.int8)
    length = np.zeros(E, dtype=np.int32)
    ratio = np.zeros((E, T), dtype=np.float32)
    old_root = ref_harness.REF_ROOT
    near = 0
    try:
        for e, lin in enumerate(tumours):
            t = np.zeros(int(np.prod(G)), dtype=np.float32)
            t[lin] = 1.0
            fname = f"0.0_0.0_0.0_0.01_stress{e}.npy"
            np.save(os.path.join(root, "data", "tumours", fname), t.reshape(tuple(G)))
            ref_harness.REF_ROOT = root
            try:
                env = ref_harness.RefEnv(visionless=True, tumour_name=fname)
                env.reset()
            finally:
                ref_harness.REF_ROOT = old_root
            for s in range(T):
                p = env.env.beam_position
                a = np.zeros(6, dtype=np.float32)
                target = 35.0 + rng.uniform(0.05, 0.95)          # stay between rows 35 and 36 of axis 2
                a[2] = np.float32((target - p[2]) / 14.0)
                if e % 2:
                    a[3] = np.float32(rng.uniform(-0.004, 0.004))    # tilt the beam in the plane
                if e % 7 == 3:
                    a[1] = np.float32(rng.uniform(-0.01, 0.01))
                acts[e, s] = a
                obs, reward, dn, _, info = env.step(a)
                r = rec[e, s]
                r[0:9] = obs
                r[9] = reward
                rc = info["reward_components"]
                r[10], r[11], r[12] = rc["tumour"], rc["lung"], rc["distance_to_tumour"]
                r[13], r[14] = info["doses"]["tumour"], info["doses"]["lung"]
                r[15:18] = info["overshoot"]["translation"]
                r[18] = info["overshoot"]["rotation"]
                inner = env.env
                mask = inner.lungs * (1 - inner.tumours)
                r[19] = np.sum(inner.dose * mask > inner.LUNG_DOSE_THRESHOLD)
                ratio[e, s] = np.sum(inner.dose * inner.tumours) / np.sum(inner.tumours)   # environment.py:186-189
                near += int(abs(float(ratio[e, s]) - 0.9) < 2e-6)
                done[e, s] = dn
                length[e] = s + 1
                if dn:
                    break
    finally:
        shutil.rmtree(root, ignore_errors=True)
    off = np.concatenate([[0], np.cumsum([len(v) for v in tumours])]).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "stress.npz"), vox_off=off, vox=np.concatenate(tumours), actions=acts,
                        rec=rec, done=done, length=length, ratio=ratio, versions=versions())
    early = int((done.sum(axis=1) > 0).sum())
    print(f"stress.npz: {E} tumours, {early} terminate before step {T}, {near} steps within 2e-6 of the threshold; "
          f"length histogram {np.bincount(length).tolist()}")


def gen_trajectory(ns, phantom_names, steps=12):
    """environment.py:69-75 export_trajectory after `steps` beams: what the reference writes."""
    import tempfile
    tid = 123
    rng = np.random.default_rng(23)
    acts = rng.uniform(-1, 1, (steps, 6)).astype(np.float32)
    env = ref_harness.RefEnv(visionless=True, tumour_name=phantom_names[tid])
    env.reset()
    for a in acts:
        env.step(a)
    path = os.path.join(tempfile.mkdtemp(prefix="rt_traj_"), "traj.npz")
    env.env.export_trajectory(path)
    z = np.load(path)
    keys = sorted(z.files)
    dose = z["dose"].reshape(-1)
    nz = np.flatnonzero(dose)
    np.savez_compressed(os.path.join(OUT, "trajectory.npz"), tumour_id=np.int32(tid), actions=acts,
                        keys=np.array(keys), shapes=np.array([str(tuple(z[k].shape)) for k in keys]),
                        dtypes=np.array([str(z[k].dtype) for k in keys]), beams=z["beams"],
                        tumours_nz=np.flatnonzero(z["tumours"].reshape(-1)).astype(np.int32),
                        dose_idx=nz.astype(np.int32), dose_val=dose[nz], versions=versions())
    print("trajectory.npz:", keys, [tuple(z[k].shape) for k in keys], [str(z[k].dtype) for k in keys])


def gen_resets(ns, phantom_names):
    obs = np.zeros((len(phantom_names), 9))
    env = ref_harness.RefEnv(visionless=True, tumour_name=phantom_names[0])
    for i, name in enumerate(phantom_names):
        o, _ = env.reset(tumour_name=name)
        obs[i] = o
    np.savez_compressed(os.path.join(OUT, "resets.npz"), obs=obs, versions=versions())
    print("resets.npz:", len(phantom_names))


def gen_gae():
    """train.py:164-181 verbatim (cfg.* replaced by locals) on CPU torch tensors."""
    import torch
    torch.manual_seed(11)
    out = {}
    for tag, (T, N) in {"a": (128, 64), "b": (17, 5), "c": (1, 3)}.items():
        rewards = torch.randn(T, N)
        values = torch.randn(T, N) * 3
        dones = (torch.rand(T, N) < 0.05).float()
        next_done = (torch.rand(N) < 0.2).float()
        next_value = torch.randn(1, N)
        gamma, gae_lambda = 0.99, 0.95
        num_steps = T
        advantages = torch.zeros_like(rewards)
        lastgaelam = 0
        for t in reversed(range(num_steps)):
            if t == num_steps - 1:
                nextnonterminal = 1.0 - next_done
                nextvalues = next_value
            else:
                nextnonterminal = 1.0 - dones[t + 1]
                nextvalues = values[t + 1]
            delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]
            advantages[t] = lastgaelam = delta + gamma * gae_lambda * nextnonterminal * lastgaelam
        returns = advantages + values
        for k, v in dict(rewards=rewards, values=values, dones=dones, next_done=next_done,
                         next_value=next_value, advantages=advantages, returns=returns).items():
            out[f"{tag}_{k}"] = v.numpy()
    out["gamma"] = np.float64(0.99)
    out["gae_lambda"] = np.float64(0.95)
    out["versions"] = versions()
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **out)
    print("gae.npz")


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_harness.load()
    names = [str(x) for x in np.load(os.path.join(REPO, "ppo-radiotherapy_b200", "data", "phantom.npz"))["names"]]
    which = sys.argv[1:] or ["beams", "poses", "steps", "tiny", "resets", "gae", "stress", "trajectory"]
    if "beams" in which:
        gen_beams(ns)
    if "poses" in which:
        gen_poses(ns)
    if "steps" in which:
        gen_steps(ns, names)
    if "tiny" in which:
        gen_tiny(ns)
    if "resets" in which:
        gen_resets(ns, names)
    if "gae" in which:
        gen_gae()
    if "stress" in which:
        gen_stress(ns)
    if "trajectory" in which:
        gen_trajectory(ns, names)


if __name__ == "__main__":
    main()
