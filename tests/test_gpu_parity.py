"""Parity of the CUDA path (through the C ABI) against the golden vectors of the reference and
against the CPU oracle on the same seeded inputs.  Needs a GPU: `pytest -m gpu`.

Stated tolerances (north_star): voxel indices, per-beam weights, dose volumes, lung counts and
done flags BIT-EXACT; float64 pose within 1e-12 absolute of the reference (CUDA sin/cos/acos differ
from glibc by <= 2 ulp) and identical after rounding to float32 — which is all draw_line.py:19-20
consumes; observations (float32) within 2e-7; rewards rtol 1e-6 / atol 1e-7; GAE bit-exact.  The rotation
overshoot (info only) is pi/4 - acos(z): acos amplifies the 1e-16 pose difference by 1/sqrt(1 - z^2), so it is
held to 1e-9 (worst seen over 409,600 steps: 1.3e-11)."""
import numpy as np
import pytest
import torch

import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import _native as nat
from oracle import oracle as O
from oracle.hashing import batch_hash, dense_hash

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
REW_RTOL, REW_ATOL = 1e-6, 1e-7
OBS_ATOL = 2e-7
POSE_ATOL = 1e-12
OVERSHOOT_ATOL = 1e-9
GRID = np.array([67.0, 43.0, 70.0])


def _cuda(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV) if dtype is None else \
        torch.as_tensor(np.ascontiguousarray(a), device=DEV).to(dtype)


# ------------------------------------------------------------------------------------ beams
def test_beams_golden_bit_exact(golden):
    g = golden("beams")
    idx, w, count = rt.beam_voxels_batch(_cuda(g["pos"]), _cuda(g["dir"]))
    idx, w, count = idx.cpu().numpy(), w.cpu().numpy(), count.cpu().numpy()
    assert (count >= 0).all()
    nnz, h = batch_hash(idx, w, count)
    assert np.array_equal(nnz, g["count"])
    assert np.array_equal(h, g["hash"])
    # no voxel appears twice in a trace
    for k in range(0, idx.shape[0], 37):
        assert len(set(idx[k, :count[k]].tolist())) == count[k]


def test_beams_full_traces(golden):
    g = golden("beams")
    ids = g["full_ids"]
    vols, status = rt.beam_voxels_dense_batch(_cuda(g["pos"][ids]), _cuda(g["dir"][ids]))
    vols = vols.cpu().numpy().reshape(len(ids), -1)
    assert (status.cpu().numpy() == 0).all()
    for j in range(len(ids)):
        nz = np.flatnonzero(vols[j])
        lo, hi = g["full_off"][j], g["full_off"][j + 1]
        assert np.array_equal(nz, g["full_idx"][lo:hi])
        assert np.array_equal(vols[j][nz].view(np.uint32), g["full_w"][lo:hi].view(np.uint32))


def test_beams_vs_oracle_random_100k():
    rng = np.random.default_rng(99)
    m = 100_000
    pos = rng.uniform(-3, 73, (m, 3))
    pos[: m // 2] = rng.uniform(0, 1, (m // 2, 3)) * GRID
    d = rng.standard_normal((m, 3))
    d[::7, rng.integers(3)] = 0.0
    oi, ow, oc = O.beam_batch(pos, d)
    idx, w, count = rt.beam_voxels_batch(_cuda(pos), _cuda(d))
    n0, h0 = batch_hash(oi, ow, oc)
    n1, h1 = batch_hash(idx.cpu().numpy(), w.cpu().numpy(), count.cpu().numpy())
    assert np.array_equal(n0, n1)
    assert np.array_equal(h0, h1)


def test_beam_voxels_dropin_and_error():
    base = np.zeros((67, 43, 70), dtype=np.float32)
    pos, d = np.array([33.5, 21.5, 35.0]), np.array([0.3, 1.0, -0.2])
    out = rt.beam_voxels(base, pos, d)
    assert out.dtype == np.float32 and out.shape == base.shape
    assert np.array_equal(out.view(np.uint32), O.beam_voxels(pos, d).view(np.uint32))
    with pytest.raises(ValueError, match="too small"):
        rt.beam_voxels(base, pos, np.array([0.0, 1e-9, 0.0]))
    # empty cases: parallel to an axis and outside the slab; pointing away is still a line (not a ray)
    assert not rt.beam_voxels(base, np.array([-5.0, 3.0, 3.0]), np.array([0.0, 1.0, 0.0])).any()
    assert rt.beam_voxels(base, np.array([-5.0, 3.0, 3.0]), np.array([-1.0, 0.0, 0.0])).any()


def test_beam_other_grid_shapes():
    rng = np.random.default_rng(3)
    for grid in [(8, 9, 10), (95, 4, 17), (2, 2, 2), (31, 64, 33)]:
        g = np.array(grid)
        pos = rng.uniform(-1, 1, (500, 3)) * g + g / 2
        d = rng.standard_normal((500, 3))
        oi, ow, oc = O.beam_batch(pos, d, grid=g, cap=4 * (max(grid) + 2))
        idx, w, count = rt.beam_voxels_batch(_cuda(pos), _cuda(d), grid=grid)
        n0, h0 = batch_hash(oi, ow, oc)
        n1, h1 = batch_hash(idx.cpu().numpy(), w.cpu().numpy(), count.cpu().numpy())
        assert np.array_equal(n0, n1) and np.array_equal(h0, h1), grid


# ------------------------------------------------------------------------------------ pose
def test_pose_golden(golden):
    g = golden("poses")
    acts = g["actions"]
    C = acts.shape[0]
    p_in = np.concatenate([np.broadcast_to(GRID / 2, (C, 1, 3)), g["pos"][:, :-1]], axis=1).reshape(-1, 3)
    d_in = np.concatenate([np.broadcast_to(np.array([0.0, 1.0, 0.0]), (C, 1, 3)), g["dir"][:, :-1]], axis=1).reshape(-1, 3)
    p, d, ot, orr = (x.cpu().numpy() for x in rt.pose_update_batch(_cuda(p_in), _cuda(d_in), _cuda(acts.reshape(-1, 6))))
    assert np.array_equal(p, g["pos"].reshape(-1, 3))
    assert np.array_equal(ot, g["overshoot_t"].reshape(-1, 3))
    np.testing.assert_allclose(d, g["dir"].reshape(-1, 3), rtol=0, atol=POSE_ATOL)
    np.testing.assert_allclose(orr, g["overshoot_r"].reshape(-1), rtol=0, atol=OVERSHOOT_ATOL)
    mism = (d.astype(np.float32) != g["dir"].reshape(-1, 3).astype(np.float32)).any(axis=1).sum()
    assert mism == 0, f"{mism} of {d.shape[0]} directions differ after float32 rounding"


def test_transform_dropins_vs_oracle():
    rng = np.random.default_rng(5)
    for _ in range(20):
        d = rng.standard_normal(3)
        rv = rng.uniform(-1.6, 1.6, 3) * rng.choice([1.0, 1e-4, 0.0])
        ma = rng.choice([np.pi / 4, 0.3, 1.2])
        got, ov = rt.apply_rotation(d, rv, ma)
        want, ov_w = O.apply_rotation(d, rv, ma)
        np.testing.assert_allclose(got, want, rtol=0, atol=POSE_ATOL)
        assert abs(ov - ov_w) < OVERSHOOT_ATOL
        p, t = rng.uniform(0, 70, 3), rng.uniform(-30, 30, 3)
        gp, go = rt.apply_translation(p, t, GRID)
        wp, wo = O.apply_translation(p, t, GRID)
        assert np.array_equal(gp, wp) and np.array_equal(go, wo)


# ------------------------------------------------------------------------------------ step
def _compare_step(info, obs, reward, term, rec, done, mask=None, overshoot_atol=None):
    if mask is not None:
        info, obs, reward, term, rec, done = info[mask], obs[mask], reward[mask], term[mask], rec[mask], done[mask]
    np.testing.assert_allclose(obs, rec[:, 0:9].astype(np.float32), rtol=0, atol=OBS_ATOL)
    np.testing.assert_allclose(reward, rec[:, 9], rtol=REW_RTOL, atol=REW_ATOL)
    np.testing.assert_allclose(info[:, 0:4], rec[:, 9:13], rtol=REW_RTOL, atol=REW_ATOL)
    np.testing.assert_allclose(info[:, 4:6], rec[:, 13:15], rtol=REW_RTOL, atol=REW_ATOL)
    assert np.array_equal(info[:, nat.INFO_OVERSHOOT_T0:nat.INFO_OVERSHOOT_T0 + 3], rec[:, 15:18])
    np.testing.assert_allclose(info[:, nat.INFO_OVERSHOOT_R], rec[:, 18], rtol=0,
                               atol=OVERSHOOT_ATOL if overshoot_atol is None else overshoot_atol)
    assert np.array_equal(info[:, nat.INFO_LUNG_COUNT], rec[:, 19])
    assert np.array_equal(term.astype(np.int8), done)


def test_step_traces_golden(golden):
    """30 reference episodes (uniform, normal and saves/20M.model actions) replayed as one batch."""
    g = golden("steps")
    tids = g["tumour_ids"]
    E, T = g["actions"].shape[:2]
    env = rt.RadiotherapyVectorEnv(E, visionless=True, device=DEV, tumour_ids=tids[None, :])
    obs0, _ = env.reset()
    np.testing.assert_allclose(obs0, g["reset_obs"].astype(np.float32), rtol=0, atol=OBS_ATOL)
    f32_pose_mismatch = 0
    for t in range(T):
        obs, reward, term, trunc, _ = env.step(_cuda(g["actions"][:, t]))
        info = env.engine.info.cpu().numpy()
        _compare_step(info, obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), g["rec"][:, t], g["done"][:, t])
        assert not trunc.any()
        pose = env.engine.pose().cpu().numpy()
        np.testing.assert_allclose(pose, g["pose"][:, t], rtol=0, atol=POSE_ATOL)
        f32_pose_mismatch += int((pose.astype(np.float32) != g["pose"][:, t].astype(np.float32)).any(axis=1).sum())
        if t % 9 == 0 or t == T - 1:
            for e in range(0, E, 3):
                assert dense_hash(env.engine.dose(e).cpu().numpy()) == g["dose_hash"][e, t], (e, t)
    assert f32_pose_mismatch == 0
    # final dose volumes, voxel by voxel
    for e in range(E):
        flat = env.engine.dose(e).cpu().numpy().reshape(-1)
        nz = np.flatnonzero(flat)
        lo, hi = g["final_off"][e], g["final_off"][e + 1]
        assert np.array_equal(nz, g["final_idx"][lo:hi])
        assert np.array_equal(flat[nz].view(np.uint32), g["final_val"][lo:hi].view(np.uint32))
    env.close()


def _tiny_phantom(g):
    base = rt.default_phantom()
    lungs = base.lungs_volume()
    vols = []
    for e in range(len(g["length"])):
        v = np.zeros(lungs.size, dtype=np.float32)
        v[g["vox"][g["vox_off"][e]:g["vox_off"][e + 1]]] = 1.0
        vols.append(v.reshape(lungs.shape))
    return rt.Phantom.from_volumes(lungs, vols, names=[str(x) for x in g["names"]])


def test_tiny_tumours_early_termination_and_autoreset(golden):
    """Few-voxel tumours reach dose ratio >= 0.9 after 9-18 beams (environment.py:184-191): the terminal
    step, the NEXT_STEP autoreset call and the following episode all match the reference / oracle."""
    g = golden("tiny")
    ph = _tiny_phantom(g)
    E = len(g["length"])
    T = 40
    acts = np.zeros((T, E, 6), dtype=np.float32)
    for e in range(E):
        L = int(g["length"][e])
        acts[:min(L, T), e] = g["actions"][e, :min(L, T)]
    rng = np.random.default_rng(8)
    for e in range(E):
        L = int(g["length"][e])
        if L < T:
            acts[L:, e] = rng.uniform(-1, 1, (T - L, 6))
    sched = np.stack([np.arange(E), (np.arange(E) + 1) % E]).astype(np.int32)
    env = rt.RadiotherapyVectorEnv(E, device=DEV, phantom=ph, tumour_ids=sched)
    env.reset()
    oph = O.Phantom()
    oph.vox_offsets = np.ascontiguousarray(g["vox_off"].astype(np.int32))
    oph.vox = np.ascontiguousarray(g["vox"].astype(np.int32))
    ref_out, ref_done = O.rollout(oph, sched, acts)
    saw_reset = 0
    for t in range(T):
        obs, reward, term, trunc, infos = env.step(_cuda(acts[t]))
        info = env.engine.info.cpu().numpy()
        stepped = info[:, nat.INFO_STEPPED] > 0
        _compare_step(info, obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), ref_out[t], ref_done[t], stepped)
        rs = ~stepped
        if rs.any():
            saw_reset += int(rs.sum())
            assert (reward.cpu().numpy()[rs] == 0).all() and not term.cpu().numpy()[rs].any()
            np.testing.assert_allclose(obs.cpu().numpy()[rs], ref_out[t, rs, 0:9].astype(np.float32), rtol=0, atol=OBS_ATOL)
        for e in range(E):   # the reference's own numbers for the first episode
            if t < g["length"][e]:
                np.testing.assert_allclose(reward.cpu().numpy()[e], g["rec"][e, t, 9], rtol=REW_RTOL, atol=REW_ATOL)
                assert bool(term.cpu().numpy()[e]) == bool(g["done"][e, t])
                assert dense_hash(env.engine.dose(e).cpu().numpy()) == g["dose_hash"][e, t]
        if "episode" in infos:
            fin = infos["episode"]["_r"]
            assert np.array_equal(fin, term.cpu().numpy() & stepped)
            assert np.array_equal(infos["episode"]["l"][fin], info[fin, nat.INFO_T].astype(np.int64))
    assert saw_reset >= 3
    env.close()


def test_termination_threshold_under_stress(golden):
    """240 synthetic 16-64-voxel tumours with homing actions, the reference's own traces
    (stress.npz).  The dose ratio crosses 0.9 on 159 of them, 49 steps lie within 2e-6 of the threshold (ratios
    0.8999999 / 0.9 / 0.90000004) where the summation order of np.sum(dose * tumours) (environment.py:166,186)
    decides `done`.  The terminal step index of every env equals the reference's: 0 flips."""
    g = golden("stress")
    E, T = g["actions"].shape[:2]
    lungs = rt.default_phantom().lungs_volume()
    vols = []
    for e in range(E):
        v = np.zeros(lungs.size, dtype=np.float32)
        v[g["vox"][g["vox_off"][e]:g["vox_off"][e + 1]]] = 1.0
        vols.append(v.reshape(lungs.shape))
    ph = rt.Phantom.from_volumes(lungs, vols, names=[f"stress{e}" for e in range(E)])
    env = rt.RadiotherapyVectorEnv(E, device=DEV, phantom=ph, tumour_ids=np.arange(E, dtype=np.int32)[None, :])
    env.reset()
    length = g["length"]
    alive = np.ones(E, dtype=bool)
    flips = 0
    early = 0
    for t in range(T):
        obs, reward, term, trunc, infos = env.step(_cuda(g["actions"][:, t]))
        term = term.cpu().numpy().astype(bool)
        info = env.engine.info.cpu().numpy()
        m = alive & (t < length)
        flips += int((term[m] != g["done"][m, t].astype(bool)).sum())
        np.testing.assert_allclose(reward.cpu().numpy()[m], g["rec"][m, t, 9], rtol=REW_RTOL, atol=REW_ATOL)
        np.testing.assert_allclose(info[m, nat.INFO_DOSE_TUMOUR], g["rec"][m, t, 13], rtol=REW_RTOL, atol=REW_ATOL)
        assert np.array_equal(info[m, nat.INFO_LUNG_COUNT], g["rec"][m, t, 19])
        early += int((term & m).sum())
        alive &= ~(g["done"][:, t].astype(bool) | term)          # past its (reference) terminal step an env is not compared
    print(f"stress: {early} early terminations, {flips} done flips")
    assert flips == 0
    assert early >= 150
    env.close()


def test_reset_observation_all_tumours(golden):
    """reset observation of all 1000 bundled tumours on the GPU against the reference's
    (resets.npz; environment.py:86-97, 145-148, 259-268)."""
    g = golden("resets")
    n = g["obs"].shape[0]
    env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=np.arange(n, dtype=np.int32)[None, :])
    obs, _ = env.reset()
    np.testing.assert_allclose(obs, g["obs"].astype(np.float32), rtol=0, atol=OBS_ATOL)
    assert np.array_equal(obs[:, 6:9], g["obs"][:, 6:9].astype(np.float32))     # centroid: a per-tumour constant, exact
    # ... and through the autoreset path: terminate every env, the next call resets it to the next tumour
    sched = np.stack([np.arange(n), (np.arange(n) + 1) % n]).astype(np.int32)
    env.engine.set_tumour_schedule(sched)
    env.reset()
    a = torch.zeros((n, 6), device=DEV)
    for _ in range(100):
        _, _, term, _, _ = env.step(a)
    assert term.all()
    obs, reward, term, _, _ = env.step(a)
    np.testing.assert_allclose(obs.cpu().numpy(), np.roll(g["obs"], -1, axis=0).astype(np.float32), rtol=0, atol=OBS_ATOL)
    assert (reward == 0).all() and not term.any()
    env.close()


def test_export_trajectory_matches_reference_file(golden, tmp_path):
    """export_trajectory (environment.py:69-75) writes the reference's npz: keys tumours, dose,
    beams with its shapes and dtypes; dose bit-exact, beams as (position, direction) pairs."""
    g = golden("trajectory")
    env = rt.RadiotherapyVectorEnv(1, device=DEV, tumour_ids=np.array([[int(g["tumour_id"])]], dtype=np.int32), record_beams=True)
    env.reset()
    for a in g["actions"]:
        env.step(_cuda(a[None, :]))
    path = str(tmp_path / "traj.npz")
    env.envs[0].export_trajectory(path)
    z = np.load(path)
    keys = sorted(z.files)
    assert keys == [str(k) for k in g["keys"]]
    assert [str(tuple(z[k].shape)) for k in keys] == [str(x) for x in g["shapes"]]
    assert [str(z[k].dtype) for k in keys] == [str(x) for x in g["dtypes"]]
    assert np.array_equal(np.flatnonzero(z["tumours"].reshape(-1)), g["tumours_nz"])
    assert set(np.unique(z["tumours"]).tolist()) == {0.0, 1.0}
    dose = z["dose"].reshape(-1)
    nz = np.flatnonzero(dose)
    assert np.array_equal(nz, g["dose_idx"])
    assert np.array_equal(dose[nz].view(np.uint32), g["dose_val"].view(np.uint32))
    assert np.array_equal(z["beams"][:, 0], g["beams"][:, 0])                   # positions: exact IEEE
    np.testing.assert_allclose(z["beams"][:, 1], g["beams"][:, 1], rtol=0, atol=POSE_ATOL)
    env.close()


@pytest.mark.parametrize("kind", ["uniform", "normal"])
def test_rollout_vs_oracle_256_envs(kind):
    """SURVEY §8d C2 at reduced width: tumour id (i*7919) mod 1000, T = 101 calls (a full episode plus
    the autoreset call) then 20 steps of the second episode."""
    n, T = 256, 121
    rng = np.random.default_rng(0 if kind == "uniform" else 1)
    acts = (rng.uniform(-1, 1, (T, n, 6)) if kind == "uniform" else rng.standard_normal((T, n, 6))).astype(np.float32)
    sched = np.stack([(np.arange(n) * 7919) % 1000, (np.arange(n) * 104729 + 17) % 1000]).astype(np.int32)
    ref_out, ref_done = O.rollout(O.Phantom(), sched, acts, threads=8)
    env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=sched)
    env.reset()
    ep_ret = np.zeros(n)
    for t in range(T):
        obs, reward, term, trunc, infos = env.step(_cuda(acts[t]))
        info = env.engine.info.cpu().numpy()
        stepped = info[:, nat.INFO_STEPPED] > 0
        assert stepped.all() == (t != 100)
        _compare_step(info, obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), ref_out[t], ref_done[t], stepped)
        np.testing.assert_allclose(obs.cpu().numpy(), ref_out[t, :, 0:9].astype(np.float32), rtol=0, atol=OBS_ATOL)
        ep_ret += ref_out[t, :, 9]
        if t == 99:
            assert term.all()
            np.testing.assert_allclose(infos["episode"]["r"], ep_ret, rtol=REW_RTOL, atol=1e-5)
            assert (infos["episode"]["l"] == 100).all()
            ep_ret[:] = 0
    env.close()


def test_dose_volumes_bit_exact_vs_oracle():
    n, T = 16, 100
    rng = np.random.default_rng(4)
    acts = rng.standard_normal((T, n, 6)).astype(np.float32) * 0.3     # slow motion: heavy re-irradiation
    acts[:, : n // 2] *= 0.02                                          # nearly parked beams: the dose clips at 1.0
    tids = (np.arange(n) * 61 + 5).astype(np.int32) % 1000
    env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=tids[None, :])
    env.reset()
    for t in range(T):
        env.step(_cuda(acts[t]))
    ph = O.Phantom()
    clipped = 0
    for e in range(n):
        o = O.OracleEnv(ph, int(tids[e]))
        for t in range(T):
            o.step(acts[t, e])
        d = env.engine.dose(e).cpu().numpy()
        assert np.array_equal(d.view(np.uint32), o.dose.view(np.uint32)), e
        clipped += int((d == 1.0).sum())
    assert clipped > 0        # the clip at 1.0 was exercised
    env.close()


def test_dense_mode_equals_sparse_mode_and_oracle():
    """RT_FLAG_DENSE (BASELINE configs[4]): full-volume update + from-scratch reductions give the same
    episode as the sparse incremental step — dose volumes bit-identical, rewards within tolerance —
    through a termination and the autoreset call."""
    n, T = 24, 104
    rng = np.random.default_rng(21)
    acts = rng.uniform(-1, 1, (T, n, 6)).astype(np.float32)
    acts[:, :4] *= 0.02
    sched = np.stack([(np.arange(n) * 7919) % 1000, (np.arange(n) * 31 + 3) % 1000]).astype(np.int32)
    a = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=sched)
    b = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=sched, dense=True)
    ref_out, ref_done = O.rollout(O.Phantom(), sched, acts, threads=8)
    oa, _ = a.reset()
    ob, _ = b.reset()
    assert np.array_equal(oa, ob)
    for t in range(T):
        o1, r1, d1, _, _ = a.step(_cuda(acts[t]))
        o2, r2, d2, _, _ = b.step(_cuda(acts[t]))
        i1, i2 = a.engine.info.cpu().numpy(), b.engine.info.cpu().numpy()
        assert torch.equal(o1, o2) and torch.equal(d1, d2)
        np.testing.assert_allclose(r2.cpu().numpy(), r1.cpu().numpy(), rtol=REW_RTOL, atol=REW_ATOL)
        assert np.array_equal(i1[:, nat.INFO_LUNG_COUNT], i2[:, nat.INFO_LUNG_COUNT])
        np.testing.assert_allclose(i2[:, :12], i1[:, :12], rtol=REW_RTOL, atol=REW_ATOL)
        stepped = i2[:, nat.INFO_STEPPED] > 0
        _compare_step(i2, o2.cpu().numpy(), r2.cpu().numpy(), d2.cpu().numpy(), ref_out[t], ref_done[t], stepped)
        if t in (0, 50, 99, 100, 103):
            for e in (0, 5, 23):
                assert torch.equal(a.engine.dose(e), b.engine.dose(e))
    assert torch.equal(a.engine.counters(), b.engine.counters())
    a.close()
    b.close()


# ------------------------------------------------------------------------------------ vision
def test_vision_volumes(golden):
    g = golden("steps")
    sel = np.arange(0, len(g["tumour_ids"]), 5)
    tids = g["tumour_ids"][sel]
    env = rt.RadiotherapyVectorEnv(len(sel), visionless=False, device=DEV, tumour_ids=tids[None, :])
    obs0, _ = env.reset()
    assert obs0.shape == (len(sel), 4, 67, 43, 70) and obs0.dtype == np.float32
    steps = g["vol_steps"]
    ph = O.Phantom()
    oenv = [O.OracleEnv(ph, int(t)) for t in tids]
    for t in range(int(steps.max()) + 1):
        obs, *_ = env.step(_cuda(g["actions"][sel, t]))
        for j, e in enumerate(sel):
            oenv[j].step(g["actions"][e, t])
        for k, vs in enumerate(steps):
            if t == vs:
                vol = obs.cpu().numpy()
                for j, e in enumerate(sel):
                    assert dense_hash(vol[j]) == g["vol_hash"][e, k]
                    assert np.array_equal(vol[j].view(np.uint32), oenv[j].volumes().view(np.uint32))
    env.close()


# ------------------------------------------------------------------------------------ GAE
def test_gae_golden_bit_exact(golden):
    g = golden("gae")
    for tag in "abc":
        adv, ret = rt.compute_gae(_cuda(g[f"{tag}_rewards"]), _cuda(g[f"{tag}_values"]), _cuda(g[f"{tag}_dones"]),
                                  _cuda(g[f"{tag}_next_value"]), _cuda(g[f"{tag}_next_done"]),
                                  float(g["gamma"]), float(g["gae_lambda"]))
        assert np.array_equal(adv.cpu().numpy().view(np.uint32), g[f"{tag}_advantages"].view(np.uint32))
        assert np.array_equal(ret.cpu().numpy().view(np.uint32), g[f"{tag}_returns"].view(np.uint32))


def test_gae_vs_oracle_large():
    rng = np.random.default_rng(2)
    T, N = 128, 4096
    r = rng.standard_normal((T, N)).astype(np.float32)
    v = (rng.standard_normal((T, N)) * 5).astype(np.float32)
    d = (rng.random((T, N)) < 0.02).astype(np.float32)
    nv = rng.standard_normal(N).astype(np.float32)
    nd = (rng.random(N) < 0.1).astype(np.float32)
    a0, r0 = O.gae(r, v, d, nv, nd, 0.99, 0.95)
    a1, r1 = rt.compute_gae(_cuda(r), _cuda(v), _cuda(d), _cuda(nv), _cuda(nd), 0.99, 0.95)
    assert np.array_equal(a0.view(np.uint32), a1.cpu().numpy().view(np.uint32))
    assert np.array_equal(r0.view(np.uint32), r1.cpu().numpy().view(np.uint32))


# ------------------------------------------------------------------------------------ host path / API
def test_host_path_equals_device_path():
    n, T = 64, 30
    rng = np.random.default_rng(12)
    acts = rng.uniform(-1, 1, (T, n, 6)).astype(np.float32)
    tids = (np.arange(n) * 7919 % 1000).astype(np.int32)[None, :]
    a = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=tids)
    b = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=tids)
    oa, _ = a.reset()
    ob, _ = b.reset()
    assert np.array_equal(oa, ob)
    for t in range(T):
        o1, r1, d1, tr1, i1 = a.step(acts[t])                       # numpy in -> numpy out
        o2, r2, d2, tr2, i2 = b.step(_cuda(acts[t]))                # tensor in -> tensor out
        assert isinstance(o1, np.ndarray) and o1.dtype == np.float32 and r1.dtype == np.float64 and d1.dtype == bool
        assert np.array_equal(o1, o2.cpu().numpy()) and np.array_equal(r1, r2.cpu().numpy())
        assert np.array_equal(d1, d2.cpu().numpy())
        for k in ("total", "tumour", "lung", "distance_to_tumour"):
            assert np.array_equal(i1["reward_components"][k], i2["reward_components"][k])
    a.close()
    b.close()


def test_single_env_facade():
    env = rt.RadiotherapyEnv(visionless=True, device=DEV, tumour_id=7)
    assert env.observation_space.shape == (9,) and env.action_space.shape == (6,)
    obs, info = env.reset(options={"tumour_id": 7})
    assert info == {} and obs.shape == (9,)
    o = O.OracleEnv(O.Phantom(), 7)
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.uniform(-1, 1, 6).astype(np.float32)
        obs, reward, term, trunc, info = env.step(a)
        want, dn = o.step(a)
        np.testing.assert_allclose(obs, want[0:9].astype(np.float32), rtol=0, atol=OBS_ATOL)
        np.testing.assert_allclose(reward, want[9], rtol=REW_RTOL, atol=REW_ATOL)
        assert term == dn and trunc is False
        assert set(info) == {"reward_components", "beam_position", "doses", "overshoot"}
    assert len(env.beams) == 5 and env.t == 5
    assert np.array_equal(env.dose.view(np.uint32), o.dose.view(np.uint32))
    v = env.get_volumes()
    assert np.array_equal(v.view(np.uint32), o.volumes().view(np.uint32))
    env.close()
    env2 = rt.RadiotherapyEnv(visionless=False, device=DEV)
    assert env2.observation().shape == (4, 67, 43, 70)
    ids = set()
    for _ in range(6):
        env2.reset()
        ids.add(env2.tumour_id)
    assert len(ids) > 1           # a fresh tumour per reset
    env2.close()


def test_rng_tumour_choice_is_seeded_and_spread():
    a = rt.BatchedEpisodes(2048, device=DEV, seed=5)
    b = rt.BatchedEpisodes(2048, device=DEV, seed=5)
    c = rt.BatchedEpisodes(2048, device=DEV, seed=6)
    for e in (a, b, c):
        e.reset()
    ia, ib, ic = (e.counters()[:, 1].cpu().numpy() for e in (a, b, c))
    assert np.array_equal(ia, ib) and not np.array_equal(ia, ic)
    assert ia.min() >= 0 and ia.max() < 1000 and len(np.unique(ia)) > 700
    for e in (a, b, c):
        e.close()


# ------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n", [61, 1501])
def test_step_kernel_block_shapes_vs_oracle(n):
    """The step kernel runs 7 envs per block while one block per SM covers the envs and 14 per block above that
    (picked from the env count in rt_create): both, with a ragged last block, an env that never moves, a full episode,
    the autoreset call and the start of the next episode, against the CPU oracle; final dose volumes bit for bit."""
    T = 106
    rng = np.random.default_rng(11 + n)
    acts = rng.uniform(-1, 1, (T, n, 6)).astype(np.float32)
    acts[:, 7, :] = 0.0                                     # an env that never moves: the same voxels every step
    sched = np.stack([(np.arange(n) * 7919) % 1000, (np.arange(n) * 104729 + 17) % 1000]).astype(np.int32)
    ref_out, ref_done = O.rollout(O.Phantom(), sched, acts, threads=8)
    env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=sched)
    env.reset()
    for t in range(T):
        obs, reward, term, _, _ = env.step(_cuda(acts[t]))
        info = env.engine.info.cpu().numpy()
        stepped = info[:, nat.INFO_STEPPED] > 0
        assert stepped.all() == (t != 100)
        # rotation overshoot (info only) = max(0, pi/4 - acos(z)): for a direction within 1e-3 of the axis the 1e-12
        # pose tolerance is amplified a thousandfold by acos; 150,000 random steps do get that close
        _compare_step(info, obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), ref_out[t], ref_done[t], stepped,
                      overshoot_atol=1e-7)
    # bit-exact dose volumes in the second episode (the first episode's cells belong to another generation)
    for e in (0, 7, n - 1):
        o = O.OracleEnv(O.Phantom(), int(sched[1, e]))
        for t in range(101, T):
            o.step(acts[t, e])
        assert np.array_equal(env.engine.dose(e).cpu().numpy().view(np.uint32), o.dose.view(np.uint32))
    env.close()


@pytest.mark.parametrize("n", [1, 6, 7, 8, 13, 33])
def test_ragged_env_counts(n):
    """Env counts that do not fill the kernel's 7-env blocks (and N = 1) behave like the oracle."""
    T = 12
    rng = np.random.default_rng(n)
    acts = rng.uniform(-1, 1, (T, n, 6)).astype(np.float32)
    tids = ((np.arange(n) * 37 + 11) % 1000).astype(np.int32)[None, :]
    ref_out, ref_done = O.rollout(O.Phantom(), tids, acts)
    env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=tids)
    env.reset()
    for t in range(T):
        obs, reward, term, _, _ = env.step(_cuda(acts[t]))
        _compare_step(env.engine.info.cpu().numpy(), obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(),
                      ref_out[t], ref_done[t])
    env.close()


def test_empty_and_degenerate_inputs():
    z3 = torch.zeros((0, 3), dtype=torch.float64, device=DEV)
    idx, w, count = rt.beam_voxels_batch(z3, z3)
    assert idx.shape[0] == 0 and count.numel() == 0
    p, d, ot, orr = rt.pose_update_batch(z3, z3, torch.zeros((0, 6), device=DEV))
    assert p.shape == (0, 3) and orr.numel() == 0
    e = torch.zeros((0, 4), device=DEV)
    adv, ret = rt.compute_gae(e, e, e, torch.zeros(4, device=DEV), torch.zeros(4, device=DEV), 0.99, 0.95)
    assert adv.shape == (0, 4)
    # T = 1: the last row uses next_value / next_done only
    r, v = torch.tensor([[1.0, 2.0]], device=DEV), torch.tensor([[0.5, 0.25]], device=DEV)
    adv, ret = rt.compute_gae(r, v, torch.zeros_like(r), torch.tensor([4.0, 8.0], device=DEV),
                              torch.tensor([0.0, 1.0], device=DEV), 0.5, 0.5)
    assert adv.cpu().tolist() == [[1.0 + 0.5 * 4.0 - 0.5, 2.0 - 0.25]]
    # a beam that misses the volume entirely deposits nothing and the step still completes
    env = rt.BatchedEpisodes(2, device=DEV)
    env.reset()
    pose = env.pose()
    pose[:, :3] = torch.tensor([80.0, 50.0, 80.0], dtype=torch.float64)         # outside: clipped to the bounds by the step
    env.set_pose(pose)
    obs, reward, term, trunc, info = env.step(torch.zeros((2, 6), device=DEV))
    assert torch.isfinite(reward).all() and not term.any()
    with pytest.raises(ValueError):
        env.step(torch.zeros((3, 6), device=DEV))
    with pytest.raises(ValueError):
        env.set_tumour_schedule(np.zeros((1, 5), dtype=np.int32))
    with pytest.raises(rt.RtError, match="out of range"):
        env.set_tumour_schedule(np.full((1, 2), 5000, dtype=np.int32))
    env.close()
    with pytest.raises(rt.RtError):
        rt.BatchedEpisodes(0, device=DEV)


def test_masked_reset_and_actions_outside_the_box():
    n = 10
    tids = np.arange(n, dtype=np.int32)[None, :]
    env = rt.BatchedEpisodes(n, device=DEV)
    env.set_tumour_schedule(tids)
    env.reset()
    big = torch.full((n, 6), 7.5, device=DEV)                                    # clipped to +-1 (environment.py:122,140)
    one = torch.ones((n, 6), device=DEV)
    env.step(big)
    a = env.pose().clone()
    env.reset()
    env.step(one)
    assert torch.equal(a, env.pose())
    before = env.counters().clone()
    mask = torch.zeros(n, dtype=torch.uint8, device=DEV)
    mask[[2, 5]] = 1
    obs = env.reset(mask).clone()
    after = env.counters()
    keep = torch.ones(n, dtype=torch.bool, device=DEV)
    keep[[2, 5]] = False
    assert torch.equal(after[keep], before[keep]) and (after[~keep][:, 0] == 0).all()
    assert float(env.dose(2).abs().sum()) == 0.0 and float(env.dose(3).abs().sum()) > 0.0
    o = O.OracleEnv(O.Phantom(), 2)
    np.testing.assert_allclose(obs[2].cpu().numpy(), o.vector_obs().astype(np.float32), rtol=0, atol=OBS_ATOL)
    env.close()


# ------------------------------------------------------------------------------------ full size
def test_full_size_properties_4096_envs():
    """BASELINE.json configs[1] width.  Size-independent properties: dose is monotone non-decreasing
    and <= 1; the lung count is monotone; rewards stay in their analytic ranges; every env terminates
    at t = 100 exactly; the step after is a reset whose observation equals the initial one; a
    checksum over all envs is reproducible run to run (determinism)."""
    n, T = 4096, 102
    g = torch.Generator(device=DEV).manual_seed(0)
    acts = torch.rand((T, n, 6), device=DEV, generator=g) * 2 - 1
    tids = ((np.arange(n) * 7919) % 1000).astype(np.int32)[None, :]

    def run():
        env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=tids)
        obs0, _ = env.reset(options={"backend": "torch"})
        obs0 = obs0.clone()
        prev_cnt = torch.zeros(n, dtype=torch.float64, device=DEV)
        prev_dose = env.engine.dose(17).clone()
        csum = torch.zeros((), dtype=torch.float64, device=DEV)
        for t in range(T):
            obs, reward, term, trunc, _ = env.step(acts[t])
            info = env.engine.info
            if t < 100:
                assert bool((info[:, nat.INFO_LUNG_COUNT] >= prev_cnt).all())
                prev_cnt = info[:, nat.INFO_LUNG_COUNT].clone()
                assert bool((info[:, nat.INFO_REWARD_TUMOUR] >= 0).all() and (info[:, nat.INFO_REWARD_TUMOUR] <= 10).all())
                assert bool((info[:, nat.INFO_REWARD_LUNG] <= 0).all() and (info[:, nat.INFO_REWARD_LUNG] >= -1).all())
                assert bool((info[:, nat.INFO_REWARD_DISTANCE] <= 0).all() and (info[:, nat.INFO_REWARD_DISTANCE] >= -1).all())
                assert bool(term.all()) == (t == 99)
                if t % 10 == 0:
                    d = env.engine.dose(17)
                    assert bool((d >= prev_dose).all() and (d <= 1).all())
                    prev_dose = d.clone()
            elif t == 100:
                assert bool((reward == 0).all()) and not bool(term.any())
                assert torch.equal(obs, obs0)
                assert float(env.engine.dose(17).abs().sum()) == 0.0
            csum += reward.sum() + obs.double().sum()
        out = float(csum)
        env.close()
        return out

    assert run() == run()


@pytest.mark.parametrize("kind", ["uniform", "normal"])
def test_full_size_rollout_vs_oracle_4096_envs(kind):
    """SURVEY §8d C2 in full: 4096 envs, tumour id (i*7919) mod 1000, T = 101 calls (one whole episode plus the
    autoreset call) = 409,600 beams, every output of every step against the CPU oracle; uniform(-1,1) actions and
    N(0,1) actions (the initial policy: actor_logstd = 0; a third of the components are clipped at +-1)."""
    n, T = 4096, 101
    rng = np.random.default_rng(0 if kind == "uniform" else 1)
    acts = (rng.uniform(-1, 1, (T, n, 6)) if kind == "uniform" else rng.standard_normal((T, n, 6))).astype(np.float32)
    tids = ((np.arange(n) * 7919) % 1000).astype(np.int32)[None, :]
    import os
    ref_out, ref_done = O.rollout(O.Phantom(), tids, acts, threads=os.cpu_count() or 8)
    env = rt.RadiotherapyVectorEnv(n, device=DEV, tumour_ids=tids)
    env.reset()
    dev_acts = _cuda(acts)
    worst = 0.0
    for t in range(T):
        obs, reward, term, _, _ = env.step(dev_acts[t])
        info = env.engine.info.cpu().numpy()
        stepped = info[:, nat.INFO_STEPPED] > 0
        assert stepped.all() == (t != 100)
        _compare_step(info, obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), ref_out[t], ref_done[t], stepped)
        if stepped.all():
            worst = max(worst, float(np.abs(reward.cpu().numpy() - ref_out[t, :, 9]).max()))
    assert worst < 1e-5
    env.close()


def test_host_buffer_step_matches_device_step():
    """rt_step_host (pinned and pageable host buffers) gives bit for bit what rt_step gives on device buffers: a full
    episode, the autoreset call and the next steps."""
    n, T = 333, 106
    rng = np.random.default_rng(31)
    acts = rng.uniform(-1, 1, (T, n, 6)).astype(np.float32)
    sched = np.stack([(np.arange(n) * 7919) % 1000, (np.arange(n) * 104729 + 17) % 1000]).astype(np.int32)
    a = rt.BatchedEpisodes(n, device=DEV); a.set_tumour_schedule(sched); a.reset()
    b = rt.BatchedEpisodes(n, device=DEV); b.set_tumour_schedule(sched); b.reset()
    pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, pin_memory=True)
    h_act, h_obs, h_rew = pin(n, 6), pin(n, 9), pin(n, dt=torch.float64)
    h_term, h_trunc, h_info = pin(n, dt=torch.uint8), pin(n, dt=torch.uint8), pin(n, nat.INFO_SIZE, dt=torch.float64)
    p_obs, p_rew = np.empty((n, 9), np.float32), np.empty(n, np.float64)          # pageable: staged by the library
    p_term, p_trunc = np.empty(n, np.uint8), np.empty(n, np.uint8)
    for t in range(T):
        obs, rew, term, trunc, info = a.step(_cuda(acts[t]), want_info=True)
        if t % 2 == 0:
            h_act.copy_(torch.from_numpy(acts[t]))
            b.step_host(h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_term.numpy(), h_trunc.numpy(), h_info.numpy())
            got = (h_obs.numpy(), h_rew.numpy(), h_term.numpy(), h_trunc.numpy())
            assert np.array_equal(h_info.numpy().view(np.uint64), info.cpu().numpy().view(np.uint64))
        else:
            b.step_host(acts[t], p_obs, p_rew, p_term, p_trunc)
            got = (p_obs, p_rew, p_term, p_trunc)
        assert np.array_equal(got[0].view(np.uint32), obs.cpu().numpy().view(np.uint32))
        assert np.array_equal(got[1].view(np.uint64), rew.cpu().numpy().view(np.uint64))
        assert np.array_equal(got[2], term.cpu().numpy()) and np.array_equal(got[3], trunc.cpu().numpy())
    assert np.array_equal(a.dose(n - 1).cpu().numpy().view(np.uint32), b.dose(n - 1).cpu().numpy().view(np.uint32))
    a.close(); b.close()


def test_host_call_waits_for_device_side_work():
    """the *_host entry points run on the handle's own stream; a host step issued right after a device-side
    reset / step / set_pose (on any stream, here behind tens of milliseconds of queued work) must see their effect."""
    n = 257
    tids = (np.arange(n) * 7919 % 1000).astype(np.int32)[None, :]
    rng = np.random.default_rng(2)
    acts = rng.uniform(-1, 1, (8, n, 6)).astype(np.float32)
    dacts = _cuda(acts)
    a = rt.BatchedEpisodes(n, device=DEV); a.set_tumour_schedule(tids)       # reference: device calls on one stream
    b = rt.BatchedEpisodes(n, device=DEV); b.set_tumour_schedule(tids)       # device calls on a busy side stream + host calls
    h_obs, h_rew = np.empty((n, 9), np.float32), np.empty(n, np.float64)
    h_term, h_trunc = np.empty(n, np.uint8), np.empty(n, np.uint8)
    big = torch.zeros(1 << 27, dtype=torch.float32, device=DEV)
    side = torch.cuda.Stream(DEV)
    torch.cuda.synchronize()

    def busy():
        for _ in range(20):
            big.add_(1.0)

    a.reset()
    with torch.cuda.stream(side):
        busy()
        b.reset()                                               # queued behind the adds
    for t in range(8):
        a.step(dacts[t], want_info=False)
        if t % 2 == 0:
            with torch.cuda.stream(side):
                busy()
                b.step(dacts[t], want_info=False)               # device-side step, still pending when the next call comes
        else:
            b.step_host(acts[t], h_obs, h_rew, h_term, h_trunc)  # no explicit synchronisation in between
            assert np.array_equal(h_obs.view(np.uint32), a.obs.cpu().numpy().view(np.uint32)), t
            assert np.array_equal(h_rew.view(np.uint64), a.reward.cpu().numpy().view(np.uint64)), t
        if t == 4:
            pose = a.pose()
            pose[:, :3] = torch.flip(pose[:, :3], dims=(0,))    # move the beams: a pose set on the device ...
            torch.cuda.synchronize()
            a.set_pose(pose)
            with torch.cuda.stream(side):
                busy()
                b.set_pose(pose)                                # ... must be seen by the host step that follows
    torch.cuda.synchronize()
    assert torch.equal(a.pose(), b.pose())
    a.close(); b.close()
