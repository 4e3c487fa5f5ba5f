// rt_policy.cuh — the rollout side of the PPO loop around the environment step, included by
// rt_env.cu.  Two kernels replace the ~45 small library launches train.py:139-161 makes per rollout step:
//
//   rt_ppo_act_kernel     train.py:139-149 + networks.py:132-147 under no_grad: store obs / done in the rollout buffers,
//        evaluate the reference's MLP agent (critic and actor_mean: Linear(n_obs,64)-Tanh-Linear(64,64)-Tanh-Linear,
//        float32, FFMA), sample the diagonal Gaussian (action = mean + exp(logstd) * N(0,1), Philox4x32-10 keyed by
//        (seed, env, step)), log-probability, store values / actions / logprobs, hand the action to the env step.
//        A block takes 64 envs: the 10,061 parameters sit transposed in shared memory, a thread owns four hidden units
//        for four envs (16 accumulators; per input one 128-bit load of four weights + one 128-bit broadcast load of four
//        activations for 16 FFMA; activation rows are XOR-swizzled in groups of four envs so that both the unit-major
//        stores and the env-major loads are conflict-free).
//   rt_ppo_record_kernel  train.py:154-161: reward and next_done into the rollout buffers, episode statistics of the
//        envs that finished (infos["episode"], infos["reward_components"]) accumulated on the device.
//
// The rollout row and the RNG step come from a device-side counter pair, so one captured CUDA graph serves every step.
// Training (autograd, Adam) stays in PyTorch on the same parameter tensors.
#pragma once
#include <curand_kernel.h>

namespace {

constexpr int kPolHidden = 64;       // feature_dim of the reference's configs (configs/default.yaml.template)
constexpr int kPolTile = 64;         // envs per block
constexpr int kPolThreads = 256;
constexpr int kPolStride = kPolTile;       // activation rows [unit][env], the groups of four envs XOR-swizzled by the row (pol_col)
constexpr int kPolMaxObs = 16;
constexpr int kPolWStride = kPolHidden + 4;   // transposed weight rows [k][unit]: 16-byte aligned rows, four units per 128-bit load
constexpr int kPolOut = 8;           // output columns: actor means 0..n_act-1 (n_act <= 6), critic value 6, unused 7

struct PolicyNet {
    const float *w0, *b0, *w1, *b1, *w2, *b2;      // torch Linear layout: weight [out][in], bias [out]
};
struct PolicyArgs {
    PolicyNet critic, actor;
    const float *logstd;                            // [n_act]
    int n_obs, n_act;
    const float *obs, *next_done;                   // [N][n_obs], [N]
    int n;
    unsigned long long seed;
    const long long *counters;                      // [0] rollout row t, [1] RNG step
    int rows;                                       // rows of the rollout buffers: a row >= rows is not stored
    float *obs_buf, *dones_buf, *values_buf, *actions_buf, *logprobs_buf;   // rollout buffers, row t is written
    float *action_out;                              // [N][n_act]: this step's actions
};

// Column of env e in activation row r: rows are [64 envs] floats, the sixteen groups of four envs permuted by the row's
// (r / 4) mod 8.  A thread stores a 128-bit group per row for rows 4 ug .. 4 ug + 3: the eight lanes of a quarter-warp
// (ug = 0..7, same env group) then hit eight different bank quads; a layer loads row k for one env group per
// half-warp: a broadcast.
__device__ __forceinline__ int pol_col(int r, int e) { return ((((e >> 2) ^ ((r >> 2) & 7)) << 2) | (e & 3)); }

// one hidden layer for this thread's units 4 ug .. 4 ug + 3 and envs 4 eg .. 4 eg + 3:
// out[u][e] = tanh(b[u] + sum_k wT[k][u] * in[k][e]), every sum the fmaf chain k = 0 .. K-1 from the bias
__device__ __forceinline__ void policy_layer(const float *wT, const float *bias, const float *in, int K, float *out, int ug, int eg)
{
    float acc[4][4];
    const float4 b = *reinterpret_cast<const float4 *>(bias + 4 * ug);
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = bb[i];
    // rows in groups of four: the swizzle of the activation row is the same for k = 4 kk .. 4 kk + 3, so one address per group
    auto step = [&](const float *wrow, const float *xrow) {
        const float4 w = *reinterpret_cast<const float4 *>(wrow);
        const float4 x = *reinterpret_cast<const float4 *>(xrow);
        const float ww[4] = {w.x, w.y, w.z, w.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(ww[i], xx[j], acc[i][j]);
    };
    const float *wp = wT + 4 * ug;
    int k = 0;
#pragma unroll 2
    for (; k + 4 <= K; k += 4) {
        const float *xp = in + k * kPolStride + ((eg ^ ((k >> 2) & 7)) << 2);
#pragma unroll
        for (int q = 0; q < 4; q++) step(wp + (k + q) * kPolWStride, xp + q * kPolStride);
    }
    for (; k < K; k++) step(wp + k * kPolWStride, in + k * kPolStride + ((eg ^ ((k >> 2) & 7)) << 2));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = 4 * ug + i;
        *reinterpret_cast<float4 *>(out + r * kPolStride + ((eg ^ (ug & 7)) << 2)) =
            make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
    }
}

__global__ void __launch_bounds__(kPolThreads) rt_ppo_act_kernel(PolicyArgs A)
{
    extern __shared__ __align__(16) float psm[];
    const int n_obs = A.n_obs, n_act = A.n_act;
    // shared-memory plan (floats)
    float *w0T = psm;                                             // [2][kPolMaxObs][65]
    float *b0 = w0T + 2 * kPolMaxObs * kPolWStride;               // [2][64]
    float *w1T = b0 + 2 * kPolHidden;                             // [2][64][65]
    float *b1 = w1T + 2 * kPolHidden * kPolWStride;               // [2][64]
    float *w2T = b1 + 2 * kPolHidden;                             // [64][8]: column o of the output layer
    float *b2 = w2T + kPolHidden * kPolOut;                       // [8]
    float *sigma = b2 + kPolOut;                                  // [8] exp(logstd), then [8] logstd
    float *xin = sigma + 2 * kPolOut;                             // [n_obs][kPolStride]
    float *h1 = xin + kPolMaxObs * kPolStride;                    // [64][kPolStride], one net at a time
    float *h2 = h1 + kPolHidden * kPolStride;                     // [2][64][kPolStride]
    float *lp = h2 + 2 * kPolHidden * kPolStride;                 // [4][kPolTile] partial log-probabilities
    const int t = threadIdx.x;

    // parameters -> shared memory, transposed (net 0 = critic, net 1 = actor)
    for (int net = 0; net < 2; net++) {
        const PolicyNet &P = net ? A.actor : A.critic;
        for (int i = t; i < kPolHidden * n_obs; i += kPolThreads) {           // w0 [64][n_obs]
            const int u = i / n_obs, k = i - u * n_obs;
            w0T[(net * kPolMaxObs + k) * kPolWStride + u] = __ldg(P.w0 + i);
        }
        for (int i = t; i < kPolHidden * kPolHidden; i += kPolThreads) {      // w1 [64][64]
            const int u = i >> 6, k = i & 63;
            w1T[(net * kPolHidden + k) * kPolWStride + u] = __ldg(P.w1 + i);
        }
        if (t < kPolHidden) {
            b0[net * kPolHidden + t] = __ldg(P.b0 + t);
            b1[net * kPolHidden + t] = __ldg(P.b1 + t);
        }
    }
    for (int i = t; i < kPolHidden * kPolOut; i += kPolThreads) {
        const int k = i >> 3, o = i & 7;
        float w = 0.0f;
        if (o < n_act) w = __ldg(A.actor.w2 + o * kPolHidden + k);
        else if (o == 6) w = __ldg(A.critic.w2 + k);
        w2T[i] = w;
    }
    if (t < kPolOut) {
        b2[t] = t < n_act ? __ldg(A.actor.b2 + t) : (t == 6 ? __ldg(A.critic.b2) : 0.0f);
        const float ls = t < n_act ? __ldg(A.logstd + t) : 0.0f;
        sigma[t] = expf(ls);
        sigma[kPolOut + t] = ls;
    }
    const long long row = A.counters[0], rng_step = A.counters[1];
    if (row < 0 || row >= A.rows) {                               // past the end of the rollout buffers: act, but store no row
        A.obs_buf = A.dones_buf = A.values_buf = A.actions_buf = A.logprobs_buf = nullptr;
    }
    const int ug = t & 15, eg = t >> 4;                           // this thread's four units and four envs of a layer

    for (int tile = blockIdx.x; tile * kPolTile < A.n; tile += gridDim.x) {
        const int env0 = tile * kPolTile;
        const int nb = min(kPolTile, A.n - env0);
        __syncthreads();                                           // parameters staged / previous tile done
        // observations: [env][k] in global -> [k][env] in shared memory, and row t of the rollout buffer (train.py:140)
        for (int i = t; i < kPolTile * n_obs; i += kPolThreads) {
            const int e = i / n_obs, k = i - e * n_obs;
            float v = 0.0f;
            if (e < nb) {
                v = A.obs[(size_t)env0 * n_obs + i];
                if (A.obs_buf) A.obs_buf[((size_t)row * A.n + env0) * n_obs + i] = v;
            }
            xin[k * kPolStride + pol_col(k, e)] = v;
        }
        if (t < nb && A.dones_buf) A.dones_buf[(size_t)row * A.n + env0 + t] = A.next_done[env0 + t];   // train.py:141
        __syncthreads();
        for (int net = 0; net < 2; net++) {
            policy_layer(w0T + net * kPolMaxObs * kPolWStride, b0 + net * kPolHidden, xin, n_obs, h1, ug, eg);
            __syncthreads();
            policy_layer(w1T + net * kPolHidden * kPolWStride, b1 + net * kPolHidden, h1, kPolHidden,
                         h2 + net * kPolHidden * kPolStride, ug, eg);
            __syncthreads();
        }
        // output layer: thread = (env, pair of output columns); columns 0..5 read the actor's features, 6 the critic's
        {
            const int e = t & 63, q = t >> 6;                      // q 0..2: means 2q, 2q+1; q 3: value (column 6)
            const float *hf = h2 + (q < 3 ? kPolHidden * kPolStride : 0);
            const int oa = q < 3 ? 2 * q : 6, ob = q < 3 ? 2 * q + 1 : 7;
            float a = b2[oa], b = b2[ob];
#pragma unroll 2
            for (int k = 0; k < kPolHidden; k += 4) {
                const float *hp = hf + k * kPolStride + pol_col(k, e);          // one swizzle per four rows
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float h = hp[q * kPolStride];
                    a = fmaf(w2T[(k + q) * kPolOut + oa], h, a);
                    b = fmaf(w2T[(k + q) * kPolOut + ob], h, b);
                }
            }
            // networks.py:141-147: action = mean + std * N(0,1); log_prob = -(a-mean)^2/(2 var) - log std - log sqrt(2 pi)
            float part = 0.0f;
            if (q < 3 && e < nb) {
                curandStatePhilox4_32_10_t st;
                curand_init(A.seed, (unsigned long long)(env0 + e) * 4ull + (unsigned long long)q, (unsigned long long)rng_step * 4ull, &st);   // one Philox block per step
                const float2 z = curand_normal2(&st);
                const float zz[2] = {z.x, z.y};
                const float mm[2] = {a, b};
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int o = 2 * q + c;
                    if (o < n_act) {
                        const float sd = sigma[o];
                        const float act = fmaf(sd, zz[c], mm[c]);
                        const float d = act - mm[c];
                        part += -(d * d) / (2.0f * sd * sd) - sigma[kPolOut + o] - 0.918938533204672742f;
                        A.action_out[(size_t)(env0 + e) * n_act + o] = act;
                        if (A.actions_buf) A.actions_buf[((size_t)row * A.n + env0 + e) * n_act + o] = act;
                    }
                }
            } else if (q == 3 && e < nb) {
                if (A.values_buf) A.values_buf[(size_t)row * A.n + env0 + e] = a;          // train.py:147
            }
            lp[q * kPolTile + e] = part;
        }
        __syncthreads();
        if (t < nb && A.logprobs_buf)
            A.logprobs_buf[(size_t)row * A.n + env0 + t] = (lp[t] + lp[kPolTile + t]) + lp[2 * kPolTile + t];   // train.py:149
    }
}

struct RecordArgs {
    const float *reward_f32;
    const uint8_t *terminated, *truncated;
    const double *info;                  // [N][RT_INFO_SIZE] or nullptr
    int n;
    const long long *counters;
    int rows;                            // rows of the rewards buffer: a row >= rows is not stored
    float *rewards_buf;                  // [T][N]
    float *next_done;                    // [N]
    double *episode_stats;               // [7]: finished, sum return, sum length, sum last-step tumour / lung / distance / total reward
};

__global__ void __launch_bounds__(256) rt_ppo_record_kernel(RecordArgs A)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n) return;
    const long long row = A.counters[0];
    if (A.rewards_buf && row >= 0 && row < A.rows) A.rewards_buf[(size_t)row * A.n + e] = A.reward_f32[e];   // train.py:154
    const bool term = A.terminated[e] != 0;
    const bool done = term || (A.truncated && A.truncated[e] != 0);                        // train.py:153
    A.next_done[e] = done ? 1.0f : 0.0f;                                                   // train.py:155-158
    if (term && A.episode_stats && A.info) {                                               // train.py:42-66, 160-161
        const double *ip = A.info + (size_t)e * RT_INFO_SIZE;
        atomicAdd(A.episode_stats + 0, 1.0);
        atomicAdd(A.episode_stats + 1, ip[RT_INFO_EPISODE_RETURN]);
        atomicAdd(A.episode_stats + 2, ip[RT_INFO_EPISODE_LENGTH]);
        atomicAdd(A.episode_stats + 3, ip[RT_INFO_REWARD_TUMOUR]);
        atomicAdd(A.episode_stats + 4, ip[RT_INFO_REWARD_LUNG]);
        atomicAdd(A.episode_stats + 5, ip[RT_INFO_REWARD_DISTANCE]);
        atomicAdd(A.episode_stats + 6, ip[RT_INFO_REWARD_TOTAL]);
    }
}

constexpr size_t kPolSmemFloats = 2 * kPolMaxObs * kPolWStride + 2 * kPolHidden + 2 * kPolHidden * kPolWStride + 2 * kPolHidden +
                                  kPolHidden * kPolOut + kPolOut + 2 * kPolOut + kPolMaxObs * kPolStride +
                                  kPolHidden * kPolStride + 2 * kPolHidden * kPolStride + 4 * kPolTile;

}  // namespace

extern "C" {

int rt_ppo_act(const rt_mlp_params *p, const float *obs_dev, const float *next_done_dev, int n, uint64_t seed,
               const int64_t *counters_dev, int rows, float *obs_buf_dev, float *dones_buf_dev, float *values_buf_dev,
               float *actions_buf_dev, float *logprobs_buf_dev, float *action_out_dev, void *stream)
{
    if (!p || !obs_dev || !next_done_dev || !counters_dev || !action_out_dev) return fail(RT_ERR_INVALID, "rt_ppo_act: NULL argument");
    if (n < 0) return fail(RT_ERR_INVALID, "rt_ppo_act: n < 0");
    if (p->hidden != kPolHidden || p->n_obs < 1 || p->n_obs > kPolMaxObs || p->n_act < 1 || p->n_act > 6)
        return fail(RT_ERR_INVALID, "rt_ppo_act: supported shapes are n_obs <= 16, hidden == 64, n_act <= 6");
    if (!p->critic_w0 || !p->critic_b0 || !p->critic_w1 || !p->critic_b1 || !p->critic_w2 || !p->critic_b2 || !p->actor_w0 ||
        !p->actor_b0 || !p->actor_w1 || !p->actor_b1 || !p->actor_w2 || !p->actor_b2 || !p->actor_logstd)
        return fail(RT_ERR_INVALID, "rt_ppo_act: NULL parameter tensor");
    if (n == 0) return RT_OK;
    const size_t smem = kPolSmemFloats * sizeof(float);
    {
        // once per device (the first call of a process comes from a warm-up step, never from inside a graph capture)
        static bool attr_set[64] = {};
        int dev = 0;
        RT_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            RT_CUDA(cudaFuncSetAttribute(rt_ppo_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    PolicyArgs A;
    A.critic = PolicyNet{p->critic_w0, p->critic_b0, p->critic_w1, p->critic_b1, p->critic_w2, p->critic_b2};
    A.actor = PolicyNet{p->actor_w0, p->actor_b0, p->actor_w1, p->actor_b1, p->actor_w2, p->actor_b2};
    A.logstd = p->actor_logstd;
    A.n_obs = p->n_obs; A.n_act = p->n_act;
    A.obs = obs_dev; A.next_done = next_done_dev; A.n = n; A.seed = seed;
    A.counters = reinterpret_cast<const long long *>(counters_dev);
    A.rows = rows;
    A.obs_buf = obs_buf_dev; A.dones_buf = dones_buf_dev; A.values_buf = values_buf_dev;
    A.actions_buf = actions_buf_dev; A.logprobs_buf = logprobs_buf_dev; A.action_out = action_out_dev;
    const int tiles = (n + kPolTile - 1) / kPolTile;
    // two blocks fit an SM (97 KB of shared memory each): beyond that a block takes several tiles and stages the
    // parameters once
    int sms = 148, dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    const int max_blocks = 2 * sms;
    rt_ppo_act_kernel<<<tiles < max_blocks ? tiles : max_blocks, kPolThreads, smem, (cudaStream_t)stream>>>(A);
    RT_LAUNCH_CHECK("rt_ppo_act_kernel");
    return RT_OK;
}

int rt_ppo_record(const float *reward_f32_dev, const uint8_t *terminated_dev, const uint8_t *truncated_dev,
                  const double *info_dev, int n, const int64_t *counters_dev, int rows, float *rewards_buf_dev, float *next_done_dev,
                  double *episode_stats_dev, void *stream)
{
    if (!reward_f32_dev || !terminated_dev || !counters_dev || !next_done_dev) return fail(RT_ERR_INVALID, "rt_ppo_record: NULL argument");
    if (n < 0) return fail(RT_ERR_INVALID, "rt_ppo_record: n < 0");
    if (n == 0) return RT_OK;
    RecordArgs A{reward_f32_dev, terminated_dev, truncated_dev, info_dev, n, reinterpret_cast<const long long *>(counters_dev),
                 rows, rewards_buf_dev, next_done_dev, episode_stats_dev};
    rt_ppo_record_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A);
    RT_LAUNCH_CHECK("rt_ppo_record_kernel");
    return RT_OK;
}

}  // extern "C"
