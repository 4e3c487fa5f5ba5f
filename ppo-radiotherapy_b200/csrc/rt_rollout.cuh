// rt_rollout.cuh — T steps of the PPO rollout (the reference's train.py:138-161) in ONE launch, included by rt_env.cu
// after rt_step.cuh and rt_policy.cuh.
//
// Per rollout step the reference does: store obs / done in the rollout buffers, agent.get_action_and_value(next_obs)
// (networks.py:132-147, no_grad), store value / action / log-prob, envs.step(action), store reward, next_done, episode
// statistics.  Launched per step that is rt_ppo_act + rt_step + rt_ppo_record (rt_policy.cuh): three grid-wide
// barriers per step, and the step kernel's blocks all wait for the slowest env of the GPU before anyone starts the
// next step.  Envs are independent, so here a block keeps its kB envs for all T steps: policy forward (the MLP's
// 10,061 float32 parameters straight from the agent's tensors through L1 / L2), Gaussian sample, log-prob, the env
// step (step_block, the same code rt_step_kernel runs), reward / done bookkeeping — and loops, without ever
// synchronising with another block.
//
// The arithmetic is the per-call kernels': every dot product is the same fmaf chain in the same order, the same
// tanhf / expf, the same Philox4x32-10 stream keyed by (seed, env, rng step), so the T rows a rollout writes equal
// T x (rt_ppo_act, rt_step, rt_ppo_record) bit for bit (tests/test_gpu_train.py); the episode statistics are summed
// with atomics in whatever order the blocks finish episodes, like rt_ppo_record's.
#pragma once

namespace {

struct RolloutArgs {
    PolicyNet critic, actor;
    const float *logstd;
    int n, n_steps;
    long long row0, rng_step0;
    unsigned long long seed;
    float *obs_buf, *dones_buf, *values_buf, *actions_buf, *logprobs_buf, *rewards_buf;   // [rows][n]..., rows row0 .. row0 + n_steps - 1 written
    float *next_obs;          // [n][9]: observation before the first step in, after the last step out
    float *next_done;         // [n]: likewise (train.py:123-124, 155-158)
    double *episode_stats;    // [7] or NULL
};

constexpr int kRollEnvPad = 16;      // activations are [unit][16 envs]: one 16-byte read serves four envs of a unit
constexpr int kRollWStride = kPolHidden + 1;   // transposed hidden-layer weights [k][unit], stride 65: conflict-free both ways
constexpr int kRollPolicyBytes = (2 * kPolHidden * kRollEnvPad + 2 * kPolHidden * kRollWStride) * 4;   // h1 + both w1T

template <int kB>
__global__ void __launch_bounds__((step_block_threads<kB, false>()), 28 / kB)
rt_rollout_kernel(Tables T, Schedule S, EnvRec *rec, uint2 *cells, double *beams, int n_envs, RolloutArgs A)
{
    static_assert(kB <= kRollEnvPad, "activation rows hold 16 envs");
    static_assert(kB < 14 || kRollPolicyBytes <= kB * kMaxPass * 4 * kWarp * 8,
                  "first-layer activations and both hidden-layer weight matrices must fit where the env step keeps its item slots "
                  "(7-env blocks stage no lungs bitmask behind the slots: the launch just asks for more dynamic shared memory)");
    __shared__ StepShared<kB> M;
    __shared__ __align__(16) float s_act[kB * RT_ACTION_SIZE];       // this step's actions, [env][6]
    __shared__ float s_done[kB];
    __shared__ float s_head[kB * kPolOut];                            // actor means 0..5, critic value 6
    __shared__ float s_lp[3 * kB];
    __shared__ float s_sigma[2 * kPolOut];                            // exp(logstd), logstd
    // small parameters, staged once per launch: first layer [net][unit][k], biases, output layer transposed [k][8]
    __shared__ float s_w0[2 * kPolHidden * RT_OBS_SIZE], s_b0[2 * kPolHidden], s_b1[2 * kPolHidden];
    __shared__ float s_w2T[kPolHidden * kPolOut], s_b2[kPolOut];
    __shared__ __align__(16) float h2[2 * kPolHidden * kRollEnvPad];  // second-layer activations [net][unit][16 envs]
    __shared__ __align__(16) float s_xT[RT_OBS_SIZE * kRollEnvPad];   // this step's observations, [k][16 envs]
    extern __shared__ __align__(128) uint32_t dyn_smem[];
    // Where the env step keeps its item slots (unused during the policy phase): first-layer activations and the two
    // 64 x 64 hidden-layer matrices, re-read from the agent's tensors (L2) and transposed every step.
    float *h1 = reinterpret_cast<float *>(dyn_smem);                  // [2 nets][64 units][16 envs]
    float *w1T = h1 + 2 * kPolHidden * kRollEnvPad;                   // [2 nets][64 k][65]
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int env0 = blockIdx.x * kB;
    const int nb = min(kB, n_envs - env0);

    step_prologue<kB, false>(M);
    for (int i = tid; i < nb * RT_OBS_SIZE; i += nthreads) M.s_obs[i] = A.next_obs[(size_t)env0 * RT_OBS_SIZE + i];
    for (int i = nb * RT_OBS_SIZE + tid; i < kB * RT_OBS_SIZE; i += nthreads) M.s_obs[i] = 0.0f;
    if (tid < kB) s_done[tid] = tid < nb ? A.next_done[env0 + tid] : 0.0f;
    if (tid < kPolOut) {
        const float ls = tid < RT_ACTION_SIZE ? __ldg(A.logstd + tid) : 0.0f;
        s_sigma[tid] = expf(ls);
        s_sigma[kPolOut + tid] = ls;
        s_b2[tid] = tid < RT_ACTION_SIZE ? __ldg(A.actor.b2 + tid) : (tid == 6 ? __ldg(A.critic.b2) : 0.0f);
    }
    for (int i = tid; i < 2 * kPolHidden * RT_OBS_SIZE; i += nthreads) {
        const int net = i / (kPolHidden * RT_OBS_SIZE), r = i - net * (kPolHidden * RT_OBS_SIZE);
        s_w0[i] = __ldg((net ? A.actor.w0 : A.critic.w0) + r);
    }
    for (int i = tid; i < 2 * kPolHidden; i += nthreads) {
        s_b0[i] = __ldg((i >> 6 ? A.actor.b0 : A.critic.b0) + (i & 63));
        s_b1[i] = __ldg((i >> 6 ? A.actor.b1 : A.critic.b1) + (i & 63));
    }
    for (int i = tid; i < kPolHidden * kPolOut; i += nthreads) {
        const int k = i >> 3, o = i & 7;
        s_w2T[i] = o < RT_ACTION_SIZE ? __ldg(A.actor.w2 + o * kPolHidden + k) : (o == 6 ? __ldg(A.critic.w2 + k) : 0.0f);
    }
    for (int i = tid; i < 2 * kPolHidden * kRollEnvPad; i += nthreads) { h1[i] = 0.0f; h2[i] = 0.0f; }   // the padding envs stay finite
    for (int i = tid; i < RT_OBS_SIZE * kRollEnvPad; i += nthreads) s_xT[i] = 0.0f;
    __syncthreads();

    for (int t = 0; t < A.n_steps; t++) {
        const size_t row = (size_t)(A.row0 + t);
        // ---- train.py:139-141: this step's observation and done flag into row t
        for (int i = tid; i < kB * RT_OBS_SIZE; i += nthreads) {
            const int e = i / RT_OBS_SIZE, k = i - e * RT_OBS_SIZE;
            const float v = M.s_obs[i];
            if (e < nb) A.obs_buf[(row * A.n + env0) * RT_OBS_SIZE + i] = v;
            s_xT[k * kRollEnvPad + e] = v;
        }
        if (tid < nb) A.dones_buf[row * A.n + env0 + tid] = s_done[tid];
        __syncthreads();
        // ---- networks.py:132-147 under no_grad.  The hidden-layer matrices come in first (coalesced 16-byte loads of
        // the rows [unit][k], transposing stores), the first layer runs while they are in flight.
        constexpr int kThreads = step_block_threads<kB, false>();
        constexpr int kWQuads = 2 * kPolHidden * kPolHidden / 4;               // 2048 float4 in both matrices
        constexpr int kWPer = (kWQuads + kThreads - 1) / kThreads;
        float4 wreg[kWPer];
#pragma unroll
        for (int j = 0; j < kWPer; j++) {
            const int i = tid + j * kThreads;
            if (i < kWQuads) wreg[j] = __ldg(reinterpret_cast<const float4 *>(i >> 10 ? A.actor.w1 : A.critic.w1) + (i & 1023));
        }
        // first layer: one (net, unit) and four envs per thread (the observations transposed to [k][16 envs])
        for (int idx = tid; idx < 2 * kPolHidden * (kRollEnvPad / 4); idx += nthreads) {
            const int g = idx & 3, nu = idx >> 2;
            if (4 * g >= kB) continue;
            const float b = s_b0[nu];
            float acc[4] = {b, b, b, b};
#pragma unroll
            for (int k = 0; k < RT_OBS_SIZE; k++) {
                const float w = s_w0[nu * RT_OBS_SIZE + k];
                const float4 v = *reinterpret_cast<const float4 *>(s_xT + k * kRollEnvPad + 4 * g);
                acc[0] = fmaf(w, v.x, acc[0]);
                acc[1] = fmaf(w, v.y, acc[1]);
                acc[2] = fmaf(w, v.z, acc[2]);
                acc[3] = fmaf(w, v.w, acc[3]);
            }
            *reinterpret_cast<float4 *>(h1 + nu * kRollEnvPad + 4 * g) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
        }
#pragma unroll
        for (int j = 0; j < kWPer; j++) {
            const int i = tid + j * kThreads;
            if (i < kWQuads) {
                const int net = i >> 10, r = i & 1023, u = r >> 4, k0 = (r & 15) * 4;
                float *dst = w1T + net * kPolHidden * kRollWStride + u;
                dst[(k0 + 0) * kRollWStride] = wreg[j].x;
                dst[(k0 + 1) * kRollWStride] = wreg[j].y;
                dst[(k0 + 2) * kRollWStride] = wreg[j].z;
                dst[(k0 + 3) * kRollWStride] = wreg[j].w;
            }
        }
        __syncthreads();
        // second layer: one (net, unit) and four envs per thread
        for (int idx = tid; idx < 2 * kPolHidden * (kRollEnvPad / 4); idx += nthreads) {
            const int g = idx & 3, nu = idx >> 2, net = nu >> 6, u = nu & 63;
            if (4 * g >= kB) continue;
            const float b = s_b1[nu];
            float acc[4] = {b, b, b, b};
            const float *wcol = w1T + net * kPolHidden * kRollWStride + u;
            const float *x = h1 + net * kPolHidden * kRollEnvPad + 4 * g;
#pragma unroll 8
            for (int k = 0; k < kPolHidden; k++) {
                const float w = wcol[k * kRollWStride];
                const float4 v = *reinterpret_cast<const float4 *>(x + k * kRollEnvPad);
                acc[0] = fmaf(w, v.x, acc[0]);
                acc[1] = fmaf(w, v.y, acc[1]);
                acc[2] = fmaf(w, v.z, acc[2]);
                acc[3] = fmaf(w, v.w, acc[3]);
            }
            *reinterpret_cast<float4 *>(h2 + nu * kRollEnvPad + 4 * g) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
        }
        __syncthreads();
        // output layer: actor means 0..5 and the critic value (column 6), one (env, column) per thread
        for (int idx = tid; idx < 7 * kB; idx += nthreads) {
            const int e = idx / 7, o = idx - e * 7;
            const float *hf = h2 + (o < RT_ACTION_SIZE ? kPolHidden * kRollEnvPad : 0) + e;
            float a = s_b2[o];
#pragma unroll 8
            for (int k = 0; k < kPolHidden; k++) a = fmaf(s_w2T[k * kPolOut + o], hf[k * kRollEnvPad], a);
            s_head[e * kPolOut + o] = a;
        }
        __syncthreads();
        // networks.py:141-147: action = mean + std * N(0,1); log_prob = -(a-mean)^2/(2 var) - log std - log sqrt(2 pi)
        for (int idx = tid; idx < 4 * kB; idx += nthreads) {
            const int q = idx / kB, e = idx - q * kB;
            if (e >= nb) continue;
            if (q == 3) {
                A.values_buf[row * A.n + env0 + e] = s_head[e * kPolOut + 6];            // train.py:147
                continue;
            }
            curandStatePhilox4_32_10_t st;
            curand_init(A.seed, (unsigned long long)(env0 + e) * 4ull + (unsigned long long)q, (unsigned long long)(A.rng_step0 + t) * 4ull, &st);
            const float2 z = curand_normal2(&st);
            const float zz[2] = {z.x, z.y};
            const float mm[2] = {s_head[e * kPolOut + 2 * q], s_head[e * kPolOut + 2 * q + 1]};
            float part = 0.0f;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int o = 2 * q + c;
                const float sd = s_sigma[o];
                const float act = fmaf(sd, zz[c], mm[c]);
                const float d = act - mm[c];
                part += -(d * d) / (2.0f * sd * sd) - s_sigma[kPolOut + o] - 0.918938533204672742f;
                s_act[e * RT_ACTION_SIZE + o] = act;
                A.actions_buf[(row * A.n + env0 + e) * RT_ACTION_SIZE + o] = act;
            }
            s_lp[q * kB + e] = part;
        }
        __syncthreads();
        if (tid < nb) A.logprobs_buf[row * A.n + env0 + tid] = (s_lp[tid] + s_lp[kB + tid]) + s_lp[2 * kB + tid];   // train.py:149
        // ---- envs.step(action) (train.py:151): reward -> row t, done -> M.s_term, next observation -> M.s_obs
        const StepOut none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        step_block<kB, false, false, true>(M, dyn_smem, T, S, rec, cells, beams, n_envs, s_act, none, nullptr, env0, (uint32_t)t,
                                           RollStep{A.rewards_buf + row * A.n, A.episode_stats});
        __syncthreads();
        if (tid < nb) s_done[tid] = M.s_term[tid] ? 1.0f : 0.0f;                          // train.py:153-158
        __syncthreads();
    }
    for (int i = tid; i < nb * RT_OBS_SIZE; i += nthreads) A.next_obs[(size_t)env0 * RT_OBS_SIZE + i] = M.s_obs[i];
    if (tid < nb) A.next_done[env0 + tid] = s_done[tid];
}

}  // namespace

extern "C" {

int rt_rollout(rt_env *e, const rt_mlp_params *p, int n_steps, int64_t row0, int rows, uint64_t seed, int64_t rng_step0,
               float *obs_buf_dev, float *dones_buf_dev, float *values_buf_dev, float *actions_buf_dev, float *logprobs_buf_dev,
               float *rewards_buf_dev, float *next_obs_dev, float *next_done_dev, double *episode_stats_dev, void *stream)
{
    if (!e || !p) return fail(RT_ERR_INVALID, "rt_rollout: NULL handle or parameters");
    if (e->dense) return fail(RT_ERR_STATE, "rt_rollout: not available for dense-mode handles");
    if (!obs_buf_dev || !dones_buf_dev || !values_buf_dev || !actions_buf_dev || !logprobs_buf_dev || !rewards_buf_dev ||
        !next_obs_dev || !next_done_dev)
        return fail(RT_ERR_INVALID, "rt_rollout: NULL buffer");
    if (p->hidden != kPolHidden || p->n_obs != RT_OBS_SIZE || p->n_act != RT_ACTION_SIZE)
        return fail(RT_ERR_INVALID, "rt_rollout: the agent must be the reference's MLP for this env (n_obs 9, hidden 64, n_act 6)");
    if (!p->critic_w0 || !p->critic_b0 || !p->critic_w1 || !p->critic_b1 || !p->critic_w2 || !p->critic_b2 || !p->actor_w0 ||
        !p->actor_b0 || !p->actor_w1 || !p->actor_b1 || !p->actor_w2 || !p->actor_b2 || !p->actor_logstd)
        return fail(RT_ERR_INVALID, "rt_rollout: NULL parameter tensor");
    if (n_steps < 0 || row0 < 0 || rng_step0 < 0 || row0 + n_steps > rows)
        return fail(RT_ERR_INVALID, "rt_rollout: rows [row0, row0 + n_steps) must lie inside the rollout buffers");
    if (n_steps == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(e->device));
    if ((cudaStream_t)stream != e->hstream) e->dev_pending = true;
    RolloutArgs A;
    A.critic = PolicyNet{p->critic_w0, p->critic_b0, p->critic_w1, p->critic_b1, p->critic_w2, p->critic_b2};
    A.actor = PolicyNet{p->actor_w0, p->actor_b0, p->actor_w1, p->actor_b1, p->actor_w2, p->actor_b2};
    A.logstd = p->actor_logstd;
    A.n = e->n; A.n_steps = n_steps; A.row0 = row0; A.rng_step0 = rng_step0; A.seed = seed;
    A.obs_buf = obs_buf_dev; A.dones_buf = dones_buf_dev; A.values_buf = values_buf_dev; A.actions_buf = actions_buf_dev;
    A.logprobs_buf = logprobs_buf_dev; A.rewards_buf = rewards_buf_dev;
    A.next_obs = next_obs_dev; A.next_done = next_done_dev; A.episode_stats = episode_stats_dev;
    const int kb = e->step_kb;
    const size_t roll_smem = e->step_smem > (size_t)kRollPolicyBytes ? e->step_smem : (size_t)kRollPolicyBytes;
    {
        // once per device (the first call of a process is never inside a graph capture: captures are preceded by warm-up)
        static bool attr_set[64][2] = {};
        if (e->device < 0 || e->device >= 64 || !attr_set[e->device][kb == 7]) {
            if (kb == 7) RT_CUDA(cudaFuncSetAttribute(rt_rollout_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roll_smem));
            else RT_CUDA(cudaFuncSetAttribute(rt_rollout_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roll_smem));
            if (e->device >= 0 && e->device < 64) attr_set[e->device][kb == 7] = true;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((e->n + kb - 1) / kb);
    cfg.blockDim = dim3((kb + 1 + (kb >= 14 ? 1 : 0)) * kWarp);
    cfg.dynamicSmemBytes = roll_smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (kb == 7)
        RT_CUDA(cudaLaunchKernelEx(&cfg, rt_rollout_kernel<7>, e->T, e->S, e->rec, e->cells, e->beams, e->n, A));
    else
        RT_CUDA(cudaLaunchKernelEx(&cfg, rt_rollout_kernel<14>, e->T, e->S, e->rec, e->cells, e->beams, e->n, A));
    RT_LAUNCH_CHECK("rt_rollout_kernel");
    return RT_OK;
}

}  // extern "C"
