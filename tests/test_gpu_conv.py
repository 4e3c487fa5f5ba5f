"""Fused Conv3d(4->16,k3)+bias+ReLU+MaxPool3d tensor-core kernel (rt_conv1_relu_pool) against PyTorch.

Tolerance: the kernel multiplies bf16-rounded inputs and weights exactly and accumulates in float32, so against
a float32 torch convolution of the same bf16-rounded operands only the summation order and the final bf16
rounding of the output differ: |diff| <= 2^-8 relative + 1e-3 absolute."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ppo_radiotherapy_b200 as rt
from ppo_radiotherapy_b200 import _native as nat

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fused(x, w, b):
    n, _, D, H, W = x.shape
    Do, Ho, Wo = D - 2, H - 2, W - 2
    pd, ph = Do % 2, Ho % 2
    out = torch.empty((n, 16, (Do + 2 * pd - 2) // 2 + 1, (Ho + 2 * ph - 2) // 2 + 1, (Wo - 2) // 2 + 1),
                      dtype=torch.bfloat16, device=x.device)
    scratch = torch.empty(4096, dtype=torch.int32, device=x.device)
    rc = nat.lib().rt_conv1_relu_pool(C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), n, D, H, W,
                                      C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return rc, out


def _reference(x, w, b):
    xb, wb = x.bfloat16().float(), w.bfloat16().float()
    y = F.relu(F.conv3d(xb, wb, b))
    pad = tuple((y.shape[i + 2]) % 2 for i in range(3))
    return F.max_pool3d(y, 2, 2, padding=pad)


@pytest.mark.parametrize("shape", [(2, 67, 43, 70), (3, 9, 12, 10), (1, 8, 11, 14), (2, 5, 5, 6), (5, 16, 3, 4)])
def test_conv1_block_matches_torch(shape):
    n, D, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(D * 1000 + H)
    x = torch.rand((n, 4, D, H, W), device=DEV, generator=g)
    x[:, :2] = (x[:, :2] > 0.5).float()                       # lungs / tumour planes are {0, 1}
    w = torch.randn((16, 4, 3, 3, 3), device=DEV, generator=g) * 0.2
    b = torch.randn(16, device=DEV, generator=g) * 0.1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rc, out = _fused(x, w, b)
    assert rc == 0
    want = _reference(x, w, b)
    assert out.shape == want.shape
    got = out.float()
    err = (got - want).abs()
    tol = want.abs() * 2.0 ** -7 + 2e-3
    assert bool((err <= tol).all()), f"max err {float(err.max())} at {int(err.argmax())}"
    assert float(got.max()) > 0.1                             # not trivially zero


def test_conv1_rejects_uncovered_shapes():
    x = torch.rand((1, 4, 8, 8, 9), device=DEV)               # odd width: pool padding 1 on the last axis
    w = torch.randn((16, 4, 3, 3, 3), device=DEV)
    rc, _ = _fused(x.contiguous(), w, torch.zeros(16, device=DEV))
    assert rc == -1


def _grouped_to_ncdhw(a, D, H, W):
    """[n][2][D][H*W][8] -> [n][16][D][H][W]"""
    n = a.shape[0]
    return a.view(n, 2, D, H, W, 8).permute(0, 1, 5, 2, 3, 4).reshape(n, 16, D, H, W)


@pytest.mark.parametrize("shape", [(2, 67, 43, 70), (3, 9, 12, 10), (1, 8, 11, 14)])
def test_conv1_grouped_layout_equals_ncdhw(shape):
    n, D, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(D + 7)
    x = torch.rand((n, 4, D, H, W), device=DEV, generator=g)
    w = torch.randn((16, 4, 3, 3, 3), device=DEV, generator=g) * 0.2
    b = torch.randn(16, device=DEV, generator=g) * 0.1
    rc, ref = _fused(x, w, b)
    assert rc == 0
    _, _, D1, H1, W1 = ref.shape
    act = torch.empty((n, 2, D1, H1 * W1, 8), dtype=torch.bfloat16, device=DEV)
    scratch = torch.empty(4096, dtype=torch.int32, device=DEV)
    rc = nat.lib().rt_conv1_relu_pool_grouped(C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), n, D, H, W,
                                              C.c_void_p(act.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    assert torch.equal(_grouped_to_ncdhw(act, D1, H1, W1), ref)


@pytest.mark.parametrize("shape", [(2, 33, 21, 34), (3, 6, 7, 8), (1, 9, 12, 10), (70, 5, 4, 6)])
def test_conv2_block_matches_torch(shape):
    """Conv3d(16->16, k3, groups=2) + ReLU + MaxPool3d(2, 2) on tcgen05 against float32 torch on the same bf16 operands."""
    n, D, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(D * 100 + W)
    x = torch.rand((n, 16, D, H, W), device=DEV, generator=g).bfloat16()          # post-ReLU activations are >= 0
    w = torch.randn((16, 8, 3, 3, 3), device=DEV, generator=g) * 0.1
    b = torch.randn(16, device=DEV, generator=g) * 0.1
    xg = x.view(n, 2, 8, D, H, W).permute(0, 1, 3, 4, 5, 2).reshape(n, 2, D, H * W, 8).contiguous()
    out = torch.empty((n, 16, (D - 2) // 2, (H - 2) // 2, (W - 2) // 2), dtype=torch.bfloat16, device=DEV)
    scratch = torch.empty(16384, dtype=torch.int32, device=DEV)
    rc = nat.lib().rt_conv2_relu_pool(C.c_void_p(xg.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), n, D, H, W,
                                      C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.backends.cudnn.allow_tf32 = False
    want = F.max_pool3d(F.relu(F.conv3d(x.float(), w.bfloat16().float(), b, groups=2)), 2, 2)
    assert out.shape == want.shape
    err = (out.float() - want).abs()
    tol = want.abs() * 2.0 ** -7 + 2e-3
    assert bool((err <= tol).all()), f"max err {float(err.max())} at {int(err.argmax())}"
    assert float(out.float().max()) > 0.1


@pytest.mark.parametrize("shape", [(3, 15, 9, 16, 64), (2, 6, 5, 8, 10), (130, 4, 4, 4, 7)])
def test_c3d_tail_matches_torch(shape):
    """Conv3d(16->16, k3, groups=4) + ReLU + MaxPool3d(2,2) + Flatten + Linear + ReLU in one kernel."""
    n, D, H, W, Fdim = shape
    g = torch.Generator(device=DEV).manual_seed(D * 10 + W)
    x = torch.rand((n, 16, D, H, W), device=DEV, generator=g).bfloat16()
    w3 = torch.randn((16, 4, 3, 3, 3), device=DEV, generator=g) * 0.1
    b3 = torch.randn(16, device=DEV, generator=g) * 0.1
    n_flat = 16 * ((D - 2) // 2) * ((H - 2) // 2) * ((W - 2) // 2)
    wl = torch.randn((Fdim, n_flat), device=DEV, generator=g) * (1.0 / n_flat ** 0.5)
    bl = torch.randn(Fdim, device=DEV, generator=g) * 0.1
    out = torch.empty((n, Fdim), dtype=torch.float32, device=DEV)
    rc = nat.lib().rt_c3d_tail(C.c_void_p(x.data_ptr()), C.c_void_p(w3.data_ptr()), C.c_void_p(b3.data_ptr()),
                               C.c_void_p(wl.data_ptr()), C.c_void_p(bl.data_ptr()), n, D, H, W, Fdim,
                               C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = F.max_pool3d(F.relu(F.conv3d(x.float(), w3, b3, groups=4)), 2, 2).flatten(1)
    want = F.relu(F.linear(y, wl, bl))
    assert out.shape == want.shape
    assert float((out - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max()))
    assert float(out.max()) > 0.01


def test_conv1_from_env_equals_conv1_on_assembled_volumes():
    """rt_conv1_from_env generates the observation planes in its loader warps: same activation as assembling the
    float32 observation in HBM and convolving it (a voxel crossed by both view beams may differ in the last bit of
    its float sum, hence a tolerance of one bf16 ulp on a handful of outputs)."""
    n = 10
    eng = rt.BatchedEpisodes(n, device=DEV, seed=21)
    eng.reset()
    g = torch.Generator(device=DEV).manual_seed(5)
    for _ in range(25):
        eng.step(torch.rand((n, 6), device=DEV, generator=g) * 2 - 1, want_info=False)
    w = torch.randn((16, 4, 3, 3, 3), device=DEV, generator=g) * 0.2
    b = torch.randn(16, device=DEV, generator=g) * 0.1
    obs = eng.volumes()
    D1, H1, W1 = 33, 21, 34
    scratch = torch.empty(4096, dtype=torch.int32, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    want = torch.empty((n, 2, D1, H1 * W1, 8), dtype=torch.bfloat16, device=DEV)
    assert nat.lib().rt_conv1_relu_pool_grouped(C.c_void_p(obs.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                                n, 67, 43, 70, C.c_void_p(want.data_ptr()), C.c_void_p(scratch.data_ptr()), stream) == 0
    for first, count in ((0, n), (3, 4)):
        got = torch.empty((count, 2, D1, H1 * W1, 8), dtype=torch.bfloat16, device=DEV)
        nat.check(nat.lib().rt_conv1_from_env(eng._h, first, count, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                              C.c_void_p(got.data_ptr()), C.c_void_p(scratch.data_ptr()), stream))
        ref = want[first:first + count].float()
        diff = (got.float() - ref).abs()
        assert bool((diff <= ref.abs() * 2.0 ** -7 + 1e-6).all())
        assert float((diff > 0).float().mean()) < 1e-4
    assert float(want.float().max()) > 0.1
    # whole extractor and the policy heads through the env path
    torch.manual_seed(0)
    agent = rt.PPO_3DCNN((4, 67, 43, 70), (6,), 64, compute_dtype=torch.bfloat16).to(DEV)
    with torch.no_grad():
        f_env = agent.features_extractor.forward_from_env(eng)
        f_obs = agent.features_extractor(obs)
    assert f_env.shape == (n, 64) and float((f_env - f_obs).abs().max()) <= 1e-3 * max(1.0, float(f_obs.abs().max()))
    a, lp, ent, v = agent.get_action_and_value_from_env(eng)
    assert a.shape == (n, 6) and torch.isfinite(v).all()
    eng.close()


def test_features_extractor_fused_path_matches_unfused():
    torch.manual_seed(0)
    fe = rt.FeaturesExtractor3D((4, 67, 43, 70), 64, compute_dtype=torch.bfloat16).to(DEV)
    envs = rt.RadiotherapyVectorEnv(6, visionless=False, device=DEV, seed=5)
    obs, _ = envs.reset(options={"backend": "torch"})
    a = torch.rand((6, 6), device=DEV) * 2 - 1
    for _ in range(5):
        obs, *_ = envs.step(a)
    with torch.no_grad():
        fe.fused_first_block = True
        y1 = fe(obs)                        # both blocks on tcgen05 + fused tail
        fe.fused_tail = False
        y1a = fe(obs)                       # both blocks on tcgen05, cuDNN tail
        fe.fused_second_block = False
        y1b = fe(obs)                       # first block only
        fe.fused_first_block = False
        y2 = fe(obs)
        fe.fused_first_block = fe.fused_second_block = fe.fused_tail = True
    assert y1.shape == (6, 64)
    assert float((y1a - y2).abs().max()) <= 0.05 * float(y2.abs().max()) + 1e-2
    assert float((y1 - y2).abs().max()) <= 0.05 * float(y2.abs().max()) + 1e-2
    assert float((y1b - y2).abs().max()) <= 0.05 * float(y2.abs().max()) + 1e-2
    # with autograd enabled the module takes the differentiable cuDNN path
    y3 = fe(obs)
    assert y3.requires_grad
    envs.close()


def test_features_extractor_against_float32_reference_module():
    """the fused bfloat16 forward (three hand-written kernels) against the reference's
    FeaturesExtractor3D evaluated in plain float32 (networks.py:8-45: the same torch layers and weights, no autocast)
    on real voxel observations.  Stated bound: |y_fused - y_fp32| <= 3e-2 * max|y_fp32| + 5e-3 (operands are rounded to
    bfloat16, 8 significant bits, once per layer; accumulation is float32); the observed error is printed."""
    torch.manual_seed(3)
    fe = rt.FeaturesExtractor3D((4, 67, 43, 70), 64, compute_dtype=torch.bfloat16).to(DEV)
    envs = rt.RadiotherapyVectorEnv(8, visionless=False, device=DEV, seed=9)
    obs, _ = envs.reset(options={"backend": "torch"})
    g = torch.Generator(device=DEV).manual_seed(1)
    for _ in range(30):
        obs, *_ = envs.step(torch.rand((8, 6), device=DEV, generator=g) * 2 - 1)
    with torch.no_grad():
        y = fe(obs)
        old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            y32 = fe.mlp(fe.cnn(obs.float()))                     # the reference forward, float32 end to end
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    err = float((y - y32).abs().max())
    scale = float(y32.abs().max())
    print(f"C3D fused bf16 vs float32 reference: max abs err {err:.3e}, max |y| {scale:.3e}, relative {err / scale:.3e}")
    assert err <= 3e-2 * scale + 5e-3
    envs.close()


def test_features_extractor_refuses_uncovered_shapes():
    """No silent cuDNN fallback under no_grad: a shape the kernels do not cover raises; the
    PyTorch layers run only when asked for (fused_first_block = False) or under autograd."""
    fe = rt.FeaturesExtractor3D((4, 35, 27, 37), 64, compute_dtype=torch.bfloat16).to(DEV)     # odd last extent
    x = torch.rand((2, 4, 35, 27, 37), device=DEV)
    with torch.no_grad():
        with pytest.raises(nat.RtError, match="does not cover"):
            fe(x)
        fe.fused_first_block = False
        assert fe(x).shape == (2, 64)
    fe.fused_first_block = True
    assert fe(x).requires_grad                                   # autograd enabled: the differentiable PyTorch path
