"""Policy / value networks with the reference's interface and checkpoint layout (networks.py:8-147).

PyTorch modules (the north star keeps the networks in PyTorch).  Module names and Sequential
indices match the reference so `saves/*.model` state dicts load unchanged:
    PPO:        critic.{0,2,4}, actor_mean.{0,2,4}, actor_logstd
    PPO_3DCNN:  features_extractor.cnn.{0,3,6}, features_extractor.mlp.0, critic.*, actor_mean.*, actor_logstd
"""
import math
from typing import Sequence

import torch
import torch.nn as nn
from torch.distributions.normal import Normal


def layer_init(layer: nn.Module, std: float = math.sqrt(2.0), bias_const: float = 0.0) -> nn.Module:
    """Orthogonal weights, constant bias (networks.py:48-51)."""
    nn.init.orthogonal_(layer.weight, std)
    nn.init.constant_(layer.bias, bias_const)
    return layer


def _head(n_in: int, hidden: int, n_out: int, out_std: float) -> nn.Sequential:
    """Linear-Tanh-Linear-Tanh-Linear, the shape of both the actor and the critic (networks.py:62-78,113-130)."""
    return nn.Sequential(layer_init(nn.Linear(n_in, hidden)), nn.Tanh(),
                         layer_init(nn.Linear(hidden, hidden)), nn.Tanh(),
                         layer_init(nn.Linear(hidden, n_out), std=out_std))


def _prod(shape: Sequence[int]) -> int:
    return int(math.prod(int(s) for s in shape))


class _GaussianActorCritic(nn.Module):
    """State-independent log-std diagonal Gaussian policy + value head (networks.py:80-97,132-147)."""

    def features(self, x: torch.Tensor) -> torch.Tensor:
        return x

    def get_value(self, x: torch.Tensor) -> torch.Tensor:
        return self.critic(self.features(x))

    def get_action_and_value(self, x: torch.Tensor, action: torch.Tensor = None):
        return self.heads(self.features(x), action)

    def heads(self, f: torch.Tensor, action: torch.Tensor = None):
        """Actor and critic heads on already extracted features (networks.py:88-97,139-147)."""
        mean = self.actor_mean(f)
        # validate_args=False: the argument check synchronises with the host, which a CUDA-graph capture forbids
        dist = Normal(mean, torch.exp(self.actor_logstd.expand_as(mean)), validate_args=False)
        if action is None:
            # dist.sample() == randn * std + mean; written out because torch.normal(mean, std) validates std
            # with a host synchronisation, which a CUDA-graph capture forbids
            with torch.no_grad():
                action = torch.randn_like(mean) * dist.scale + mean
        return action, dist.log_prob(action).sum(1), dist.entropy().sum(1), self.critic(f)


class PPO(_GaussianActorCritic):
    def __init__(self, observation_shape, action_space, feature_dim: int = 64):
        super().__init__()
        self.feature_dim, self.observation_shape, self.action_space = feature_dim, observation_shape, action_space
        n_obs, n_act = _prod(observation_shape), _prod(action_space)
        self.critic = _head(n_obs, feature_dim, 1, 1.0)
        self.actor_mean = _head(n_obs, feature_dim, n_act, 0.01)
        self.actor_logstd = nn.Parameter(torch.zeros(1, n_act))


class FeaturesExtractor3D(nn.Module):
    """C3D-style trunk (networks.py:8-45): Conv3d(C,16,3) ReLU MaxPool(2,2,pad) Conv3d(16,16,3,g=2) ReLU
    MaxPool(2) Conv3d(16,16,3,g=4) ReLU MaxPool(2) Flatten Linear ReLU.  `compute_dtype=torch.bfloat16`
    runs the convolutions under bf16 autocast (cuDNN, NCDHW: measured faster than channels-last-3d for this
    4->16-channel trunk, whose last pooling layer is pathological in channels-last; tools/c3d_layers.py)."""

    def __init__(self, observation_shape, features_dim: int, compute_dtype: torch.dtype = None):
        super().__init__()
        self.observation_shape = observation_shape
        self.compute_dtype = compute_dtype
        self.fused_first_block = True       # rollout (no_grad) path: rt_conv1_relu_pool instead of cuDNN + 4 more kernels
        self.fused_second_block = True      # ... and rt_conv2_relu_pool for cnn[3:6]
        self.fused_tail = True              # ... and rt_c3d_tail for cnn[6:] + mlp: no library kernel left in the rollout
        pad = tuple((int(observation_shape[i + 1]) - 2) % 2 for i in range(3))
        self.cnn = nn.Sequential(
            nn.Conv3d(int(observation_shape[0]), 16, 3), nn.ReLU(), nn.MaxPool3d(2, 2, padding=pad),
            nn.Conv3d(16, 16, 3, groups=2), nn.ReLU(), nn.MaxPool3d(2, 2),
            nn.Conv3d(16, 16, 3, groups=4), nn.ReLU(), nn.MaxPool3d(2, 2),
            nn.Flatten())
        with torch.no_grad():
            n_flat = self.cnn(torch.zeros((1,) + tuple(int(s) for s in observation_shape))).shape[1]
        self.mlp = nn.Sequential(nn.Linear(n_flat, features_dim), nn.ReLU())

    def _fused_first_block(self, observations: torch.Tensor):
        """Conv3d + ReLU + MaxPool3d (cnn[0:3]) as one tensor-core kernel of librtenv_b200.so (inference only:
        no autograd graph).  Returns None for shapes / dtypes the kernel does not cover."""
        import ctypes as C
        from . import _native as nat
        conv = self.cnn[0]
        x = observations
        if (x.dtype != torch.float32 or x.dim() != 5 or x.shape[1] != 4 or conv.out_channels != 16
                or tuple(conv.kernel_size) != (3, 3, 3) or x.shape[4] % 2 or not x.is_contiguous()):
            return None
        n, _, D, H, W = x.shape
        Do, Ho, Wo = D - 2, H - 2, W - 2
        pd, ph = Do % 2, Ho % 2
        shape = (n, 16, (Do + 2 * pd - 2) // 2 + 1, (Ho + 2 * ph - 2) // 2 + 1, (Wo - 2) // 2 + 1)
        out = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
        if getattr(self, "_conv_scratch", None) is None or self._conv_scratch.device != x.device:
            self._conv_scratch = torch.empty(4096, dtype=torch.int32, device=x.device)
        w = conv.weight.detach().float().contiguous()
        b = conv.bias.detach().float().contiguous()
        with torch.cuda.device(x.device):
            rc = nat.lib().rt_conv1_relu_pool(C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                              n, D, H, W, C.c_void_p(out.data_ptr()),
                                              C.c_void_p(self._conv_scratch.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        if rc == -1:                                    # RT_ERR_INVALID: shape not covered
            return None
        nat.check(rc, "rt_conv1_relu_pool")
        return out

    def _fused_two_blocks(self, observations: torch.Tensor):
        """cnn[0:6] — both Conv3d + ReLU + MaxPool3d blocks — as two tcgen05 kernels of librtenv_b200.so: the first
        writes its activation in the grouped channels-last layout the second bulk-copies into shared memory
        (inference only).  Returns None for shapes / dtypes the kernels do not cover."""
        import ctypes as C
        from . import _native as nat
        conv1, pool1, conv2, pool2 = self.cnn[0], self.cnn[2], self.cnn[3], self.cnn[5]
        x = observations
        if (x.dtype != torch.float32 or x.dim() != 5 or x.shape[1] != 4 or conv1.out_channels != 16
                or tuple(conv1.kernel_size) != (3, 3, 3) or x.shape[4] % 2 or not x.is_contiguous()
                or conv2.groups != 2 or conv2.in_channels != 16 or conv2.out_channels != 16
                or tuple(conv2.kernel_size) != (3, 3, 3) or pool2.padding not in (0, (0, 0, 0))):
            return None
        n, _, D, H, W = x.shape
        Do, Ho, Wo = D - 2, H - 2, W - 2
        pd, ph = Do % 2, Ho % 2
        D1, H1, W1 = (Do + 2 * pd - 2) // 2 + 1, (Ho + 2 * ph - 2) // 2 + 1, (Wo - 2) // 2 + 1
        if W1 % 2 or min(D1, H1, W1) < 4:
            return None
        act1 = torch.empty((n, 2, D1, H1 * W1, 8), dtype=torch.bfloat16, device=x.device)
        out = torch.empty((n, 16, (D1 - 2) // 2, (H1 - 2) // 2, (W1 - 2) // 2), dtype=torch.bfloat16, device=x.device)
        if getattr(self, "_conv_scratch2", None) is None or self._conv_scratch2.device != x.device:
            self._conv_scratch = torch.empty(4096, dtype=torch.int32, device=x.device)
            self._conv_scratch2 = torch.empty(16384, dtype=torch.int32, device=x.device)
        w1, b1 = conv1.weight.detach().float().contiguous(), conv1.bias.detach().float().contiguous()
        w2, b2 = conv2.weight.detach().float().contiguous(), conv2.bias.detach().float().contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            rc = nat.lib().rt_conv1_relu_pool_grouped(C.c_void_p(x.data_ptr()), C.c_void_p(w1.data_ptr()), C.c_void_p(b1.data_ptr()),
                                                      n, D, H, W, C.c_void_p(act1.data_ptr()),
                                                      C.c_void_p(self._conv_scratch.data_ptr()), stream)
            if rc == -1:
                return None
            nat.check(rc, "rt_conv1_relu_pool_grouped")
            rc = nat.lib().rt_conv2_relu_pool(C.c_void_p(act1.data_ptr()), C.c_void_p(w2.data_ptr()), C.c_void_p(b2.data_ptr()),
                                              n, D1, H1, W1, C.c_void_p(out.data_ptr()),
                                              C.c_void_p(self._conv_scratch2.data_ptr()), stream)
        if rc == -1:
            return None
        nat.check(rc, "rt_conv2_relu_pool")
        return out

    def forward_from_env(self, engine, first: int = 0, count: int = None) -> torch.Tensor:
        """Features of the current voxel observation of envs [first, first+count) of a BatchedEpisodes engine without
        materialising the observation: the first block generates the four planes inside its loader warps
        (rt_conv1_from_env), then rt_conv2_relu_pool and rt_c3d_tail.  Inference only; equals
        forward(engine.volumes(first, count)) up to the summation order of crossing view beams."""
        import ctypes as C
        from . import _native as nat
        count = engine.num_envs - first if count is None else count
        conv1, conv2 = self.cnn[0], self.cnn[3]
        D, H, W = engine.grid
        Do, Ho, Wo = D - 2, H - 2, W - 2
        pd, ph = Do % 2, Ho % 2
        D1, H1, W1 = (Do + 2 * pd - 2) // 2 + 1, (Ho + 2 * ph - 2) // 2 + 1, (Wo - 2) // 2 + 1
        dev = engine.device
        act1 = torch.empty((count, 2, D1, H1 * W1, 8), dtype=torch.bfloat16, device=dev)
        act2 = torch.empty((count, 16, (D1 - 2) // 2, (H1 - 2) // 2, (W1 - 2) // 2), dtype=torch.bfloat16, device=dev)
        if getattr(self, "_conv_scratch2", None) is None or self._conv_scratch2.device != dev:
            self._conv_scratch = torch.empty(4096, dtype=torch.int32, device=dev)
            self._conv_scratch2 = torch.empty(16384, dtype=torch.int32, device=dev)
        w1, b1 = conv1.weight.detach().float().contiguous(), conv1.bias.detach().float().contiguous()
        w2, b2 = conv2.weight.detach().float().contiguous(), conv2.bias.detach().float().contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            nat.check(nat.lib().rt_conv1_from_env(engine._h, int(first), int(count), C.c_void_p(w1.data_ptr()),
                                                  C.c_void_p(b1.data_ptr()), C.c_void_p(act1.data_ptr()),
                                                  C.c_void_p(self._conv_scratch.data_ptr()), stream), "rt_conv1_from_env")
            nat.check(nat.lib().rt_conv2_relu_pool(C.c_void_p(act1.data_ptr()), C.c_void_p(w2.data_ptr()), C.c_void_p(b2.data_ptr()),
                                                   count, D1, H1, W1, C.c_void_p(act2.data_ptr()),
                                                   C.c_void_p(self._conv_scratch2.data_ptr()), stream), "rt_conv2_relu_pool")
        y = self._fused_tail(act2)
        if y is None:
            raise nat.RtError("FeaturesExtractor3D.forward_from_env: rt_c3d_tail does not cover this layer configuration")
        return y

    def _fused_tail(self, h: torch.Tensor):
        """cnn[6:] + mlp — Conv3d(16,16,3,groups=4) ReLU MaxPool3d(2,2) Flatten Linear ReLU — as one kernel
        (rt_c3d_tail) on the second block's bfloat16 activation.  Returns None when the layers do not match."""
        import ctypes as C
        from . import _native as nat
        conv3, pool3, lin = self.cnn[6], self.cnn[8], self.mlp[0]
        n, c, D, H, W = h.shape
        F = lin.out_features
        if (h.dtype != torch.bfloat16 or c != 16 or not h.is_contiguous() or conv3.groups != 4 or conv3.in_channels != 16
                or conv3.out_channels != 16 or tuple(conv3.kernel_size) != (3, 3, 3) or pool3.padding not in (0, (0, 0, 0))
                or min(D, H, W) < 4 or F > 256 or lin.in_features != 16 * ((D - 2) // 2) * ((H - 2) // 2) * ((W - 2) // 2)):
            return None
        out = torch.empty((n, F), dtype=torch.float32, device=h.device)
        w3, b3 = conv3.weight.detach().float().contiguous(), conv3.bias.detach().float().contiguous()
        wl, bl = lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous()
        with torch.cuda.device(h.device):
            rc = nat.lib().rt_c3d_tail(C.c_void_p(h.data_ptr()), C.c_void_p(w3.data_ptr()), C.c_void_p(b3.data_ptr()),
                                       C.c_void_p(wl.data_ptr()), C.c_void_p(bl.data_ptr()), n, D, H, W, F,
                                       C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(h.device).cuda_stream))
        if rc == -1:
            return None
        nat.check(rc, "rt_c3d_tail")
        return out

    def _not_covered(self, what: str, observations: torch.Tensor):
        from . import _native as nat
        return nat.RtError(
            f"FeaturesExtractor3D: {what} does not cover input shape {tuple(observations.shape)} / this layer configuration "
            "(the hand-written kernels implement networks.py:22-45 of the reference: Conv3d(4->16,k3), groups 2 and 4, even "
            "last extent).  There is no silent fallback: set fused_first_block = False on the module to run the "
            "PyTorch/cuDNN layers explicitly.")

    def forward(self, observations: torch.Tensor) -> torch.Tensor:
        """Inference (no_grad, bfloat16, CUDA): the hand-written kernels, and an error for shapes they do not cover.
        Training (autograd enabled) and modules with fused_first_block = False: the PyTorch layers."""
        if (self.compute_dtype == torch.bfloat16 and observations.is_cuda and not torch.is_grad_enabled()
                and self.fused_first_block):
            if self.fused_second_block:
                h = self._fused_two_blocks(observations)
                if h is None:
                    raise self._not_covered("rt_conv1_relu_pool_grouped + rt_conv2_relu_pool", observations)
                if self.fused_tail:
                    y = self._fused_tail(h)
                    if y is None:
                        raise self._not_covered("rt_c3d_tail", observations)
                    return y
                with torch.autocast("cuda", dtype=torch.bfloat16):       # fused_tail = False: explicit opt-out
                    h = self.cnn[6:](h)
                return self.mlp(h.float())
            h = self._fused_first_block(observations)                    # fused_second_block = False: explicit opt-out
            if h is None:
                raise self._not_covered("rt_conv1_relu_pool", observations)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                h = self.cnn[3:](h)
            return self.mlp(h.float())
        if self.compute_dtype is not None and observations.is_cuda:
            with torch.autocast("cuda", dtype=self.compute_dtype):
                x = self.cnn(observations)
            return self.mlp(x.float())
        return self.mlp(self.cnn(observations))


class PPO_3DCNN(_GaussianActorCritic):
    def __init__(self, observation_shape, action_space, feature_dim: int = 64, compute_dtype: torch.dtype = None):
        super().__init__()
        self.feature_dim, self.observation_shape, self.action_space = feature_dim, observation_shape, action_space
        n_act = _prod(action_space)
        self.features_extractor = FeaturesExtractor3D(observation_shape, feature_dim, compute_dtype)
        self.critic = _head(feature_dim, feature_dim, 1, 1.0)
        self.actor_mean = _head(feature_dim, feature_dim, n_act, 0.01)
        self.actor_logstd = nn.Parameter(torch.zeros(1, n_act))

    def features(self, x: torch.Tensor) -> torch.Tensor:
        return self.features_extractor(x)

    def get_action_and_value_from_env(self, engine, action: torch.Tensor = None):
        """get_action_and_value on the engine's current voxel observations, which are never materialised."""
        with torch.no_grad():
            return self.heads(self.features_extractor.forward_from_env(engine), action)

    def summary(self):
        n = sum(p.numel() for p in self.parameters())
        print(f"Observation shape: {self.observation_shape}\nAction space: {self.action_space}\n"
              f"Feature dim: {self.feature_dim}\nParameters: {n}")
