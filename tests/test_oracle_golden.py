"""Pin the CPU oracle (oracle/rt_oracle.c) against the golden vectors produced by
executing the unmodified reference (oracle/gen_golden.py).  CPU only.

Tolerances: voxel indices / weights / dose volumes / done flags / lung counts are
bit-exact; float64 pose within 1e-12 absolute (BLAS summation order differs)
and identical after rounding to float32 (what draw_line.py:19-20 consumes); rewards
rtol 1e-6 (north_star "stated float32 tolerance")."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.hashing import batch_hash, dense_hash

RTOL = 1e-6
POSE_ATOL = 1e-12


def test_np_sum_emulation():
    rng = np.random.default_rng(1)
    for n in (1, 7, 8, 9, 127, 128, 129, 1000, 201670):
        a = rng.random(n, dtype=np.float32)
        assert np.sum(a) == O.np_sum_f32(a)


def test_beams_bit_exact(golden):
    g = golden("beams")
    idx, w, count = O.beam_batch(g["pos"], g["dir"])
    assert (count >= 0).all()
    nnz, h = batch_hash(idx, w, count)
    assert np.array_equal(nnz, g["count"])
    assert np.array_equal(h, g["hash"])
    # the batched (merged) form agrees with the dense form
    for k in range(0, idx.shape[0], 997):
        vol = np.zeros(67 * 43 * 70, dtype=np.float32)
        vol[idx[k, :count[k]]] = w[k, :count[k]]
        assert np.array_equal(vol, O.beam_voxels(g["pos"][k], g["dir"][k]).reshape(-1))


def test_beams_full_traces(golden):
    g = golden("beams")
    for j, k in enumerate(g["full_ids"]):
        vol = O.beam_voxels(g["pos"][k], g["dir"][k]).reshape(-1)
        nz = np.flatnonzero(vol)
        lo, hi = g["full_off"][j], g["full_off"][j + 1]
        assert np.array_equal(nz, g["full_idx"][lo:hi])
        assert np.array_equal(vol[nz].view(np.uint32), g["full_w"][lo:hi].view(np.uint32))


def test_beam_error_path():
    with pytest.raises(ValueError, match="too small"):
        O.beam_voxels([1.0, 1.0, 1.0], [0.0, 0.0, 1e-9])


def test_pose_chains(golden):
    g = golden("poses")
    acts = g["actions"]
    C, S = acts.shape[:2]
    G = np.array([67.0, 43.0, 70.0])
    # every step starts from the reference's previous pose, so errors do not compound in the comparison
    p_in = np.concatenate([np.broadcast_to(G / 2, (C, 1, 3)), g["pos"][:, :-1]], axis=1)
    d_in = np.concatenate([np.broadcast_to(np.array([0.0, 1.0, 0.0]), (C, 1, 3)), g["dir"][:, :-1]], axis=1)
    p, d, ot, orr = O.pose_batch(p_in.reshape(-1, 3), d_in.reshape(-1, 3), acts.reshape(-1, 6))
    assert np.array_equal(p, g["pos"].reshape(-1, 3))                 # translation is exact IEEE
    assert np.array_equal(ot, g["overshoot_t"].reshape(-1, 3))
    np.testing.assert_allclose(d, g["dir"].reshape(-1, 3), rtol=0, atol=POSE_ATOL)
    np.testing.assert_allclose(orr, g["overshoot_r"].reshape(-1), rtol=0, atol=1e-11)
    assert np.array_equal(d.astype(np.float32), g["dir"].reshape(-1, 3).astype(np.float32))
    assert (orr > 0).sum() > 100 and (ot > 0).sum() > 100             # both clamps exercised
    # the scalar entry points agree with the batch
    p1, ot1 = O.apply_translation(p_in[3, 7], np.clip(acts[3, 7, :3], -1, 1) * np.array([67, 43, 70]) * 0.2, G)
    d1, or1 = O.apply_rotation(d_in[3, 7], np.clip(acts[3, 7, 3:], -1, 1) * np.pi * 0.5, np.pi / 4)
    k = 3 * S + 7
    assert np.array_equal(p1, p[k]) and np.array_equal(d1, d[k]) and or1 == orr[k]


def _check_episode(env, acts, rec, done, length, dose_hash=None, pose=None):
    for t in range(length):
        out, dn = env.step(acts[t])
        r = rec[t]
        np.testing.assert_allclose(out[0:9], r[0:9], rtol=0, atol=1e-12)
        np.testing.assert_allclose(out[9:15], r[9:15], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(out[15:18], r[15:18], rtol=0, atol=0)
        np.testing.assert_allclose(out[18], r[18], rtol=0, atol=1e-11)
        assert out[19] == r[19]
        assert dn == bool(done[t])
        if dose_hash is not None:
            assert dense_hash(env.dose) == dose_hash[t]
        if pose is not None:
            np.testing.assert_allclose(env.pose, pose[t], rtol=0, atol=POSE_ATOL)


def test_step_traces(golden, phantom):
    g = golden("steps")
    for e, tid in enumerate(g["tumour_ids"]):
        env = O.OracleEnv(phantom, int(tid))
        np.testing.assert_allclose(env.vector_obs(), g["reset_obs"][e], rtol=0, atol=1e-15)
        _check_episode(env, g["actions"][e], g["rec"][e], g["done"][e], int(g["length"][e]),
                       g["dose_hash"][e], g["pose"][e])
        flat = env.dose.reshape(-1)
        nz = np.flatnonzero(flat)
        lo, hi = g["final_off"][e], g["final_off"][e + 1]
        assert np.array_equal(nz, g["final_idx"][lo:hi])
        assert np.array_equal(flat[nz].view(np.uint32), g["final_val"][lo:hi].view(np.uint32))


def test_vision_volumes(golden, phantom):
    g = golden("steps")
    steps = g["vol_steps"]
    for e in range(0, len(g["tumour_ids"]), 5):
        env = O.OracleEnv(phantom, int(g["tumour_ids"][e]))
        for t in range(int(steps.max()) + 1):
            env.step(g["actions"][e, t])
            for j, vs in enumerate(steps):
                if t == vs:
                    assert dense_hash(env.volumes()) == g["vol_hash"][e, j]


def test_tiny_tumours_terminate_early(golden, phantom):
    g = golden("tiny")
    ph = O.Phantom()
    ph.vox_offsets = np.ascontiguousarray(g["vox_off"].astype(np.int32))
    ph.vox = np.ascontiguousarray(g["vox"].astype(np.int32))
    lengths = g["length"]
    assert (lengths < 100).sum() >= 3
    for e in range(len(lengths)):
        env = O.OracleEnv(ph, e)
        _check_episode(env, g["actions"][e], g["rec"][e], g["done"][e], int(lengths[e]), g["dose_hash"][e])
        assert bool(g["done"][e, lengths[e] - 1])


def test_termination_threshold_under_stress(golden, phantom):
    """stress.npz: 240 synthetic 16-64-voxel tumours whose dose ratio crosses 0.9 (environment.py:184-191), 49 steps
    within 2e-6 of the threshold: the oracle's NumPy-order float32 sums reproduce every `done` flag and the float32
    ratio bit for bit."""
    g = golden("stress")
    ph = O.Phantom()
    ph.vox_offsets = np.ascontiguousarray(g["vox_off"].astype(np.int32))
    ph.vox = np.ascontiguousarray(g["vox"].astype(np.int32))
    lengths = g["length"]
    assert (g["done"].sum(axis=1) > 0).sum() >= 150
    near = 0
    for e in range(len(lengths)):
        env = O.OracleEnv(ph, e)
        n_vox = int(g["vox_off"][e + 1] - g["vox_off"][e])
        for t in range(int(lengths[e])):
            out, dn = env.step(g["actions"][e, t])
            assert dn == bool(g["done"][e, t]), (e, t)
            ratio = np.float32(out[13]) / np.float32(n_vox)              # info doses.tumour / sum(tumours)
            assert ratio == g["ratio"][e, t], (e, t)
            near += int(abs(float(ratio) - 0.9) < 2e-6)
            np.testing.assert_allclose(out[9:15], g["rec"][e, t, 9:15], rtol=RTOL, atol=1e-7)
            assert out[19] == g["rec"][e, t, 19]
    assert near >= 40


def test_reset_obs_all_tumours(golden, phantom):
    g = golden("resets")
    env = O.OracleEnv(phantom, 0)
    for tid in range(phantom.n_tumours):
        np.testing.assert_allclose(env.reset(tid), g["obs"][tid], rtol=0, atol=1e-15)
    # packed centroid table == reference's tumour_position()
    G = np.array([67.0, 43.0, 70.0])
    np.testing.assert_array_equal(phantom.centroid / G * 2 - 1, g["obs"][:, 6:9])


def test_gae(golden):
    g = golden("gae")
    for tag in "abc":
        adv, ret = O.gae(g[f"{tag}_rewards"], g[f"{tag}_values"], g[f"{tag}_dones"], g[f"{tag}_next_value"],
                         g[f"{tag}_next_done"], float(g["gamma"]), float(g["gae_lambda"]))
        assert np.array_equal(adv.view(np.uint32), g[f"{tag}_advantages"].view(np.uint32))
        assert np.array_equal(ret.view(np.uint32), g[f"{tag}_returns"].view(np.uint32))


def test_vector_rollout_autoreset(phantom):
    """NEXT_STEP autoreset (gymnasium 1.0.0): the call after a terminal step resets."""
    rng = np.random.default_rng(5)
    T, n = 103, 3
    acts = rng.uniform(-1, 1, (T, n, 6)).astype(np.float32)
    tids = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.int32)
    out, done = O.rollout(phantom, tids, acts, threads=2)
    assert done[99].all() and not done[:99].any()
    assert not done[100].any() and (out[100, :, 9] == 0).all()
    env = O.OracleEnv(phantom, 4)
    np.testing.assert_array_equal(out[100, 0, 0:9], env.vector_obs())
    o, _ = env.step(acts[101, 0])
    np.testing.assert_array_equal(out[101, 0], o)
