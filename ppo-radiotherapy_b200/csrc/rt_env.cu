// rt_env.cu — kernels and C ABI (include/rt_env.h) of the batched environment step.
//
// HBM layout per handle (N envs, V voxels):
//   rec     [N]            128-byte env record: pose f64, accumulators, counters, dose generation
//   cells   [N][cstride]   sparse mode: one 8-byte cell per voxel {float32 dose, uint32 generation} in bricks of
//                          2x2x4 voxels = one 128-byte line (cell_index, rt_device.cuh).  A cell of another generation than its env's reads as zero, so
//                          reset never touches the 1.6 MB volume (it bumps EnvRec::gen) and a beam needs no
//                          validity bitmap: every voxel it hits is one 8-byte load and one 8-byte store.
//   dose    [N][vstride]   dense mode (RT_FLAG_DENSE): plain float32 volumes, streamed whole every step
//   lungs bitmask, tumour table, tumour bbox bitmasks, packed voxel lists: < 1 MB, replicated.
//
// The step is rt_step_kernel (rt_step.cuh): per block one scalar warp (a thread per env: pose update in float64,
// beam clip and the serial float32 slab walk, rewards, termination, observation, NEXT_STEP autoreset) and one env
// warp per env (splat targets, sparse dose read-modify-write, tumour / lung deltas).  This file holds the record /
// table definitions, reset, dense-mode, beam, pose, voxel-observation, observation-record and GAE kernels, and the
// C ABI.
#include "../../include/rt_env.h"
#include "rt_device.cuh"

#include <cuda_bf16.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace rt;

// ---------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<unsigned> g_alias_epoch{0};      // bumped whenever this library frees pinned memory (see pinned_alias)

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define RT_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(RT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

#define RT_LAUNCH_CHECK(name)                                                                 \
    do {                                                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return fail(RT_ERR_CUDA, std::string(name) + " launch: " + cudaGetErrorString(_e)); \
    } while (0)

struct __align__(128) EnvRec {
    double pos[3];          // beam_position  (environment.py:101)
    double dir[3];          // beam_direction (environment.py:102)
    double tumour_dose;     // sum(dose * tumours), maintained incrementally
    double lung_dose;       // sum(dose * lungs)
    double ep_return;       // RecordEpisodeStatistics episode_returns
    int32_t t;              // environment.py:47
    int32_t tumour_id;
    int32_t lung_count;     // #voxels of lungs\tumour with dose > 0.2
    int32_t episode;        // episodes started since rt_reset (indexes the tumour schedule)
    int32_t needs_reset;    // terminated on the previous call (NEXT_STEP autoreset)
    int32_t n_beams;
    uint32_t gen;           // dose generation: cells stamped with another value read as zero; bumped by every reset
    int32_t pad;
    double dn[3];           // dir / |dir| as the NEXT step's transforms.py:23 computes it (off that step's critical path)
};
static_assert(sizeof(EnvRec) == 128, "EnvRec must be one 128-byte line");

struct __align__(16) Tumour {
    int32_t lo[3];          // bbox origin
    int32_t dim[3];         // bbox extent
    int32_t n_vox;
    int32_t vox_off;        // into vox_xyz
    float tumour_sum;       // np.sum(tumours)            (environment.py:167)
    float lung_mask_sum;    // np.sum(lungs*(1-tumours))  (environment.py:178)
    float obs_c[3];         // float32(tumour_position()/G*2-1), environment.py:145-148,261
    int32_t pad_[3];
};
static_assert(sizeof(Tumour) == 64, "Tumour table entry layout");
constexpr int kTumourWords = sizeof(Tumour) / 4;
constexpr int kMaxTumourWords = 64;   // bbox occupancy words staged per env (bundled tumours need 38)
constexpr int kMaxPTumourWords = 96;  // same for the bitmask padded by one voxel on axes 1 and 2 (+1 word of slack)

struct Tables {
    Grid G;
    double gnorm;                  // np.linalg.norm(LUNG_SHAPE), environment.py:161
    const uint32_t *lungs_bits;    // bit v = lungs.flat[v]  (= lungs_pad + kLungPadBits / 32)
    const uint32_t *lungs_pad;     // the same mask with kLungPadBits zero bits in front (bit_pair lookups, bulk copies)
    const Tumour *tumours;
    const uint32_t *tumour_bits;   // [n_tumours][bits_words] bbox-local occupancy
    const uint32_t *vox_xyz;       // packed i | j<<8 | k<<16 per tumour voxel
    const uint32_t *tumour_pbits;  // [n_tumours][pbits_words] occupancy of the bbox grown by one voxel on axes 1 and 2
    int n_tumours;
    int bits_words;
    int pbits_words;
    int lung_words16;              // words of lungs_pad (padding included) rounded up to a multiple of 4 (16-byte bulk copies)
    long long *stage_clock;        // optional [N][12] clock64() stamps of the step kernel's stages (rt_set_stage_clock)
};

struct Schedule {
    const int32_t *ids;     // [n_episodes][N] or nullptr
    int n_episodes;
    uint64_t seed;
};

struct StepOut {
    float *obs;
    double *reward;
    float *reward_f32;
    uint8_t *terminated;
    uint8_t *truncated;
    double *info;
};

// Dense mode (RT_FLAG_DENSE): hand-over from the step kernel to rt_dense_kernel, one per env in HBM.
struct __align__(16) DenseWork {
    int mode;                 // 0 accumulate this beam, 1 zero the volume (reset), 2 nothing to do
    int n_hits;
    int tid, t, n_beams, pad_;
    double best;              // min squared distance to the tumour
    double os_t[3], os_r;
    double ep_return;
    int lin[RT_BEAM_CAP];     // distinct voxels of the beam and their summed weights
    float w[RT_BEAM_CAP];
};

// ---------------------------------------------------------------------------------
__device__ __forceinline__ int pick_tumour(const Tables &T, const Schedule &S, int env, int n_envs, int episode)
{
    if (S.ids) {
        int e = episode < S.n_episodes ? episode : S.n_episodes - 1;
        return S.ids[(size_t)e * n_envs + env];
    }
    uint64_t h = splitmix64(S.seed ^ splitmix64(((uint64_t)(uint32_t)env << 32) | (uint32_t)episode));
    return (int)(h % (uint64_t)T.n_tumours);
}

// environment.py:259-268 get_vector_observation -> float32 (SyncVectorEnv copies into a float32 buffer):
// lanes 0-2 position/G*2-1, lanes 3-5 direction, lanes 6-8 tumour centroid/G*2-1 (per-tumour constant).
__device__ __forceinline__ void write_obs(const Tables &T, const Tumour &tm, const double p[3], const double d[3],
                                          float *obs, int lane)
{
    if (lane < 9) {
        const int a = lane % 3;
        float v;
        if (lane < 3) {
            const double g = a == 0 ? (double)T.G.g0 : (a == 1 ? (double)T.G.g1 : (double)T.G.g2);
            v = (float)__dsub_rn(__dmul_rn(__ddiv_rn(a == 0 ? p[0] : (a == 1 ? p[1] : p[2]), g), 2.0), 1.0);
        } else if (lane < 6) {
            v = (float)(a == 0 ? d[0] : (a == 1 ? d[1] : d[2]));
        } else {
            v = a == 0 ? tm.obs_c[0] : (a == 1 ? tm.obs_c[1] : tm.obs_c[2]);
        }
        obs[lane] = v;
    }
}

// environment.py:77-105 for the warp's env: new tumour, centred pose, empty dose (= a new generation: every cell
// written before reads as zero from now on), zero counters.
__device__ __forceinline__ int reset_env(const Tables &T, const Schedule &S, EnvRec *rec, int env, int n_envs, int episode,
                                         int lane, float *obs)
{
    const int tid = pick_tumour(T, S, env, n_envs, episode);
    const double p[3] = {(double)T.G.g0 / 2.0, (double)T.G.g1 / 2.0, (double)T.G.g2 / 2.0};
    const double d[3] = {0.0, 1.0, 0.0};
    if (lane == 0) {
        EnvRec r;
        r.pos[0] = p[0]; r.pos[1] = p[1]; r.pos[2] = p[2];
        r.dir[0] = d[0]; r.dir[1] = d[1]; r.dir[2] = d[2];
        r.dn[0] = d[0]; r.dn[1] = d[1]; r.dn[2] = d[2];             // (0, 1, 0) / 1
        r.tumour_dose = 0.0; r.lung_dose = 0.0; r.ep_return = 0.0;
        r.t = 0; r.tumour_id = tid; r.lung_count = 0; r.episode = episode; r.needs_reset = 0; r.n_beams = 0;
        r.gen = rec[env].gen + 1u;
        r.pad = 0;
        rec[env] = r;
    }
    const Tumour tm = T.tumours[tid];
    if (obs) write_obs(T, tm, p, d, obs + (size_t)env * RT_OBS_SIZE, lane);
    return tid;
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rt_reset_kernel(Tables T, Schedule S, EnvRec *rec, int n_envs,
                                                       const uint8_t *mask, float *obs, DenseWork *dense)
{
    const int env = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    const int lane = threadIdx.x & (kWarp - 1);
    if (env >= n_envs) return;
    const bool mine = mask == nullptr || mask[env];
    if (dense && lane == 0) dense[env].mode = mine ? 1 : 2;
    if (mine) {
        reset_env(T, S, rec, env, n_envs, 0, lane, obs);
    } else if (obs) {
        const EnvRec r = rec[env];
        const Tumour tm = T.tumours[r.tumour_id];
        write_obs(T, tm, r.pos, r.dir, obs + (size_t)env * RT_OBS_SIZE, lane);
    }
}

}  // namespace
#include "rt_step.cuh"
namespace {

// ---------------------------------------------------------------------------------
// Shared-memory hash table voxel -> summed beam weight for the streaming kernels (a beam hits <= 284
// voxels of 201,670, so a per-voxel lookup must be O(1)).  1024 slots, linear probing.
constexpr int kHashSlots = 1024;

__device__ __forceinline__ int hash_slot(int lin) { return (int)(((uint32_t)lin * 2654435761u) >> 22); }

__device__ __forceinline__ void hash_add(int *keys, float *vals, int lin, float w)
{
    int slot = hash_slot(lin);
    while (true) {
        const int prev = atomicCAS(keys + slot, -1, lin);
        if (prev == -1 || prev == lin) {
            atomicAdd(vals + slot, w);      // at most two addends per voxel: (0 + a) + b is order-independent
            return;
        }
        slot = (slot + 1) & (kHashSlots - 1);
    }
}

__device__ __forceinline__ float hash_get(const int *keys, const float *vals, int lin)
{
    int slot = hash_slot(lin);
    while (keys[slot] != lin) slot = (slot + 1) & (kHashSlots - 1);
    return vals[slot];
}

// ---------------------------------------------------------------------------------
// Dense mode, second kernel (BASELINE configs[4], the reference's own dataflow): one block per env streams
// the whole float32 volume  dose = clip(dose + beam*0.1, 0, 1)  (environment.py:107-110) and recomputes
// sum(dose*tumours), sum(dose*lungs) and count(dose*mask > 0.2) from scratch (environment.py:164-191,
// 234-235), then finishes the step.  1.61 MB of HBM traffic per env-step, everything else on chip.
constexpr int kDenseThreads = 512;

__global__ void __launch_bounds__(kDenseThreads) rt_dense_kernel(Tables T, EnvRec *rec, float *dose, DenseWork *dense,
                                                                 StepOut out)
{
    extern __shared__ uint32_t smem[];
    const Grid &G = T.G;
    const int env = blockIdx.x;
    DenseWork &dw = dense[env];
    const int mode = dw.mode;
    if (mode == 2) return;
    float4 *vol4 = reinterpret_cast<float4 *>(dose + (size_t)env * G.vstride);
    const int ngroups = G.vstride / 4;
    if (mode == 1) {                                    // reset: environment.py:104-105
        for (int q = threadIdx.x; q < ngroups; q += blockDim.x) __stcs(vol4 + q, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const int nwords = G.vstride / 32;
    uint32_t *tum_bits = smem;                          // [nwords] voxel belongs to the tumour
    uint32_t *hit_bits = smem + nwords;                 // [nwords] voxel is hit by this step's beam
    int *hkeys = reinterpret_cast<int *>(smem + 2 * nwords);
    float *hvals = reinterpret_cast<float *>(hkeys + kHashSlots);
    __shared__ double red_t[kDenseThreads / kWarp], red_l[kDenseThreads / kWarp];
    __shared__ int red_c[kDenseThreads / kWarp];
    const int tid = dw.tid;
    const Tumour tm = T.tumours[tid];
    const int nhit = dw.n_hits;
    for (int i = threadIdx.x; i < 2 * nwords; i += blockDim.x) smem[i] = 0u;
    for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) { hkeys[i] = -1; hvals[i] = 0.0f; }
    __syncthreads();
    for (int k = threadIdx.x; k < tm.n_vox; k += blockDim.x) {
        const uint32_t pk = __ldg(T.vox_xyz + tm.vox_off + k);
        const int lin = ((int)(pk & 255u) * G.g1 + (int)((pk >> 8) & 255u)) * G.g2 + (int)(pk >> 16);
        atomicOr(tum_bits + (lin >> 5), 1u << (lin & 31));
    }
    for (int k = threadIdx.x; k < nhit; k += blockDim.x) {
        const int lin = dw.lin[k];
        hash_add(hkeys, hvals, lin, dw.w[k]);
        atomicOr(hit_bits + (lin >> 5), 1u << (lin & 31));
    }
    __syncthreads();

    double s_t = 0.0, s_l = 0.0;
    int cnt = 0;
    constexpr int kUnroll = 4;                          // independent 16-byte loads in flight per thread
    for (int q0 = threadIdx.x; q0 < ngroups; q0 += kUnroll * kDenseThreads) {
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int q = q0 + u * kDenseThreads;
            v[u] = q < ngroups ? __ldcs(vol4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int q = q0 + u * kDenseThreads;
            if (q >= ngroups) break;
            const int v0 = q * 4, sh = v0 & 31;
            const uint32_t hb = (hit_bits[v0 >> 5] >> sh) & 15u;
            const uint32_t tb = (tum_bits[v0 >> 5] >> sh) & 15u;
            const uint32_t lb = (__ldg(T.lungs_bits + (v0 >> 5)) >> sh) & 15u;
            float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            if (hb) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if ((hb >> i) & 1u) {
                        const float wsum = hash_get(hkeys, hvals, v0 + i);
                        const float nd = __fadd_rn(e[i], __fmul_rn(wsum, 0.100000001490116119f));
                        e[i] = fminf(fmaxf(nd, 0.0f), 1.0f);
                    }
            }
            // voxels the beam misses: dose + 0*0.1 == dose and the clip is the identity on [0, 1] — written back anyway
            __stcs(vol4 + q, make_float4(e[0], e[1], e[2], e[3]));
            if (tb | lb) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const bool in_t = (tb >> i) & 1u, in_l = (lb >> i) & 1u;
                    if (in_t) s_t += (double)e[i];
                    if (in_l) s_l += (double)e[i];
                    if (in_l && !in_t && e[i] > 0.200000002980232239f) cnt++;
                }
            }
        }
    }
    s_t = warp_sum(s_t);
    s_l = warp_sum(s_l);
    cnt = __reduce_add_sync(kFull, cnt);
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x & (kWarp - 1);
    if (lane == 0) { red_t[warp] = s_t; red_l[warp] = s_l; red_c[warp] = cnt; }
    __syncthreads();
    if (warp != 0) return;
    s_t = lane < kDenseThreads / kWarp ? red_t[lane] : 0.0;
    s_l = lane < kDenseThreads / kWarp ? red_l[lane] : 0.0;
    cnt = lane < kDenseThreads / kWarp ? red_c[lane] : 0;
    const double tumour_dose = warp_sum(s_t), lung_dose = warp_sum(s_l);
    const int lung_count = __reduce_add_sync(kFull, cnt);

    EnvRec *my = rec + env;
    const int t = dw.t;
    float tsum_f32 = (float)tumour_dose;
    float ratio = fdiv_rn_zero_num(tsum_f32, tm.tumour_sum);         // no dose on the tumour yet: the common case
    if (fabsf(ratio - kDoneRatio) < kDoneWindow) {                   // on the termination threshold: NumPy's summation order decides
        __threadfence_block();
        const float *vol = dose + (size_t)env * G.vstride;
        if (lane == 0)
            tsum_f32 = np_pairwise_sparse(G.nvox, T.vox_xyz + tm.vox_off, tm.n_vox, G.g1, G.g2, [=](int lin) { return __ldcg(vol + lin); });
        tsum_f32 = __shfl_sync(kFull, tsum_f32, 0);
        ratio = __fdiv_rn(tsum_f32, tm.tumour_sum);
    }
    const float r_tumour = __fmul_rn(ratio, 10.0f);
    const double r_lung = __dmul_rn(__ddiv_rn((double)lung_count, (double)tm.lung_mask_sum), -1.0);
    const double r_dist = __dmul_rn(__ddiv_rn(sqrt(dw.best), T.gnorm), -1.0);
    const double reward = __dadd_rn(__dadd_rn((double)r_tumour, r_lung), r_dist);
    const bool done = (ratio >= kDoneRatio) || (t >= RT_MAX_TIME_STEPS);
    const double ep_return = dw.ep_return + reward;
    if (lane == 0) {
        my->tumour_dose = tumour_dose; my->lung_dose = lung_dose; my->ep_return = ep_return;
        my->lung_count = lung_count;
        my->needs_reset = done ? 1 : 0;
        if (out.reward) out.reward[env] = reward;
        if (out.reward_f32) out.reward_f32[env] = (float)reward;
        if (out.terminated) out.terminated[env] = done ? 1 : 0;
        if (out.truncated) out.truncated[env] = 0;
    }
    if (out.info && lane < RT_INFO_SIZE) {
        double v;
        switch (lane) {
        case RT_INFO_REWARD_TOTAL: v = reward; break;
        case RT_INFO_REWARD_TUMOUR: v = (double)r_tumour; break;
        case RT_INFO_REWARD_LUNG: v = r_lung; break;
        case RT_INFO_REWARD_DISTANCE: v = r_dist; break;
        case RT_INFO_DOSE_TUMOUR: v = (double)tsum_f32; break;
        case RT_INFO_DOSE_LUNG: v = (double)(float)lung_dose; break;
        case RT_INFO_OVERSHOOT_T0: v = dw.os_t[0]; break;
        case RT_INFO_OVERSHOOT_T0 + 1: v = dw.os_t[1]; break;
        case RT_INFO_OVERSHOOT_T0 + 2: v = dw.os_t[2]; break;
        case RT_INFO_OVERSHOOT_R: v = dw.os_r; break;
        case RT_INFO_EPISODE_RETURN: v = ep_return; break;
        case RT_INFO_EPISODE_LENGTH: v = (double)t; break;
        case RT_INFO_LUNG_COUNT: v = (double)lung_count; break;
        case RT_INFO_STEPPED: v = 1.0; break;
        case RT_INFO_TUMOUR_ID: v = (double)tid; break;
        default: v = (double)t; break;
        }
        out.info[(size_t)env * RT_INFO_SIZE + lane] = v;
    }
}

// ---------------------------------------------------------------------------------
// Stateless geometry kernels (parity surface): one warp per ray / one thread per pose.
struct RayWork {
    Beam beam;
    float ys[kMaxSlabs], zs[kMaxSlabs];
};

// lane 0 of the warp clips the ray and replays the serial walk; returns the beam to every lane
__device__ __forceinline__ Beam ray_prepare(const Grid &G, const double *pos3, const double *dir3, int lane,
                                            RayWork &rw)
{
    if (lane == 0) {
        const double p[3] = {pos3[0], pos3[1], pos3[2]};
        const double d[3] = {dir3[0], dir3[1], dir3[2]};
        const Beam b = beam_setup(G, p, d);
        beam_walk(b, rw.ys, rw.zs);
        rw.beam = b;
    }
    __syncwarp();
    return rw.beam;
}

__global__ void __launch_bounds__(256) rt_beam_kernel(Grid G, const double *__restrict__ pos,
                                                      const double *__restrict__ dir, int m, int cap,
                                                      int32_t *idx, float *wout, int32_t *count)
{
    __shared__ RayWork work[8];
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    const int lane = threadIdx.x & (kWarp - 1);
    if (ray >= m) return;
    RayWork &rw = work[threadIdx.x / kWarp];
    const Beam b = ray_prepare(G, pos + 3 * (size_t)ray, dir + 3 * (size_t)ray, lane, rw);
    if (b.nslab < 0) {
        if (lane == 0) count[ray] = -1;
        return;
    }
    int base = 0;
    for (int kbase = 0; kbase < b.nslab; kbase += kWarp) {
        int lin[4], c0, c1, c2;
        float w[4];
        slab_targets(G, b, rw.ys, rw.zs, kbase + lane, lin, w, c0, c1, c2);
        int mine = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) mine += lin[j] >= 0;
        int incl = mine;                                   // warp inclusive scan
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        int at = base + incl - mine;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (lin[j] >= 0) {
                if (at < cap) {
                    idx[(size_t)ray * cap + at] = lin[j];
                    wout[(size_t)ray * cap + at] = w[j];
                }
                at++;
            }
        base += __shfl_sync(kFull, incl, kWarp - 1);
    }
    if (lane == 0) count[ray] = base < cap ? base : cap;
}

__global__ void __launch_bounds__(256) rt_beam_dense_kernel(Grid G, const double *__restrict__ pos,
                                                            const double *__restrict__ dir, int m, float *out,
                                                            int32_t *status)
{
    __shared__ RayWork work[8];
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    const int lane = threadIdx.x & (kWarp - 1);
    if (ray >= m) return;
    RayWork &rw = work[threadIdx.x / kWarp];
    const Beam b = ray_prepare(G, pos + 3 * (size_t)ray, dir + 3 * (size_t)ray, lane, rw);
    if (lane == 0 && status) status[ray] = b.nslab < 0 ? -1 : 0;
    float *vol = out + (size_t)ray * G.nvox;
    for (int kbase = 0; kbase < b.nslab; kbase += kWarp) {
        int lin[4], c0, c1, c2;
        float w[4];
        slab_targets(G, b, rw.ys, rw.zs, kbase + lane, lin, w, c0, c1, c2);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (lin[j] >= 0) vol[lin[j]] = w[j];
    }
}

__global__ void __launch_bounds__(128) rt_pose_kernel(Grid G, const double *__restrict__ pos,
                                                      const double *__restrict__ dir,
                                                      const float *__restrict__ actions, int m, double *pos_out,
                                                      double *dir_out, double *os_t_out, double *os_r_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    Pose s;
    float a[6];
#pragma unroll
    for (int i = 0; i < 3; i++) { s.p[i] = pos[3 * k + i]; s.d[i] = dir[3 * k + i]; }
#pragma unroll
    for (int i = 0; i < 6; i++) a[i] = actions[6 * k + i];
    double os_t[3], os_r;
    pose_update(G, a, s, os_t, os_r);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        pos_out[3 * k + i] = s.p[i];
        dir_out[3 * k + i] = s.d[i];
        if (os_t_out) os_t_out[3 * k + i] = os_t[i];
    }
    if (os_r_out) os_r_out[k] = os_r;
}

__global__ void __launch_bounds__(128) rt_rotation_kernel(const double *__restrict__ dir,
                                                          const double *__restrict__ rotvec, int m, double min_angle,
                                                          double cos_min, double xy_mag, double *dir_out,
                                                          double *os_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    double d[3] = {dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]};
    const double rv[3] = {rotvec[3 * k], rotvec[3 * k + 1], rotvec[3 * k + 2]};
    double os;
    apply_rotation(d, rv, min_angle, cos_min, xy_mag, os);
    dir_out[3 * k] = d[0]; dir_out[3 * k + 1] = d[1]; dir_out[3 * k + 2] = d[2];
    if (os_out) os_out[k] = os;
}

__global__ void __launch_bounds__(128) rt_translation_kernel(const double *__restrict__ pos,
                                                             const double *__restrict__ tr, int m, double b0,
                                                             double b1, double b2, double *pos_out, double *os_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * m) return;
    const int a = k % 3;
    double os;
    pos_out[k] = translate_axis(pos[k], tr[k], a == 0 ? b0 : (a == 1 ? b1 : b2), os);
    if (os_out) os_out[k] = os;
}

// ---------------------------------------------------------------------------------
// State access kernels.
__global__ void rt_get_pose_kernel(const EnvRec *rec, int n, double *pose)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * 6) return;
    const int e = k / 6, c = k % 6;
    pose[k] = c < 3 ? rec[e].pos[c] : rec[e].dir[c - 3];
}

__global__ void rt_set_pose_kernel(EnvRec *rec, int n, const double *pose)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * 6) return;
    const int e = k / 6, c = k % 6;
    if (c < 3) rec[e].pos[c] = pose[k];
    else rec[e].dir[c - 3] = pose[k];
    if (c == 0) {                                   // dir / |dir| for the next step (transforms.py:23)
        double a = pose[6 * e + 3], b = pose[6 * e + 4], d = pose[6 * e + 5];
        normalize3(a, b, d);
        rec[e].dn[0] = a; rec[e].dn[1] = b; rec[e].dn[2] = d;
    }
}

__global__ void rt_get_counters_kernel(const EnvRec *rec, int n, int32_t *out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const EnvRec r = rec[e];
    int32_t *o = out + (size_t)e * 6;
    o[0] = r.t; o[1] = r.tumour_id; o[2] = r.lung_count; o[3] = r.episode; o[4] = r.needs_reset; o[5] = r.n_beams;
}

// Source of an env's dose volume: sparse-mode cells {dose, generation} or a dense-mode float32 volume.
struct DoseSrc {
    const uint2 *cells;      // [N][cstride] (bricked) or nullptr
    const float *dense;      // [N][vstride] or nullptr
};

// dense float32 dose of one env: a cell of another generation reads as zero
__global__ void rt_get_dose_kernel(Grid G, const EnvRec *rec, DoseSrc D, int env, float *out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= G.nvox) return;
    if (D.cells) {
        const uint2 c = D.cells[(size_t)env * G.cstride + cell_index_lin(G, v)];
        out[v] = c.y == rec[env].gen ? __uint_as_float(c.x) : 0.0f;
    } else {
        out[v] = D.dense[(size_t)env * G.vstride + v];
    }
}

__global__ void rt_get_beams_kernel(const EnvRec *rec, const double *beams, int env, double *out, int32_t *n_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0 && n_out) *n_out = rec[env].n_beams;
    if (k < RT_MAX_TIME_STEPS * 6) out[k] = beams[(size_t)env * RT_MAX_TIME_STEPS * 6 + k];
}

// ---------------------------------------------------------------------------------
// Voxel observation (environment.py:245-257): block per env.  The two view beams are traced by warps 0
// and 1 into a shared hit table, the tumour is expanded into a shared bitset, then every thread streams
// voxel pairs: out[c][v] = clip({lungs, tumours, dose, view}[v], 0, 1) with 8-byte loads and stores
// (V is even, so every channel plane stays 8-byte aligned).  4.03 MB of HBM traffic per env.
constexpr int kVolThreads = 512;

// kPacked = false: live envs [first, first + gridDim.x) (record, dose cells or dense float32 volume).
// kPacked = true: compressed observation records (rt_pack_observations): bfloat16 dose volume, pose and tumour id of
// record index[blockIdx.x] (or blockIdx.x); the other three planes are regenerated, the dose plane is the stored
// bfloat16 value — exactly what a bf16 convolution reads from the float32 observation.
struct PackedObs {
    const __nv_bfloat16 *dose;   // [records][vstride]
    const double *pose;          // [records][6]
    const int32_t *tid;          // [records]
    const int64_t *index;        // [count] or nullptr
};

template <bool kPacked>
__global__ void __launch_bounds__(kVolThreads) rt_volumes_kernel(Tables T, const EnvRec *rec, DoseSrc D, int first,
                                                                 float *out, PackedObs P)
{
    extern __shared__ uint32_t smem[];
    __shared__ RayWork view[2];
    const Grid &G = T.G;
    const int nwords = G.vstride / 32;
    uint32_t *hit_bits = smem;                                  // [nwords] voxel hit by a view beam
    uint32_t *tum_bits = smem + nwords;                         // [nwords] voxel belongs to the tumour
    int *hkeys = reinterpret_cast<int *>(smem + 2 * nwords);    // voxel -> current_beam + horizontal_beam_center
    float *hvals = reinterpret_cast<float *>(hkeys + kHashSlots);
    const int lane = threadIdx.x & (kWarp - 1);
    const int warp = threadIdx.x / kWarp;
    long long env;
    EnvRec r;
    if (kPacked) {
        env = P.index ? P.index[blockIdx.x] : blockIdx.x;
#pragma unroll
        for (int i = 0; i < 3; i++) { r.pos[i] = P.pose[env * 6 + i]; r.dir[i] = P.pose[env * 6 + 3 + i]; }
        r.tumour_id = P.tid[env];
    } else {
        env = first + blockIdx.x;
        r = rec[env];
    }
    const Tumour tm = T.tumours[r.tumour_id];

    for (int i = threadIdx.x; i < 2 * nwords; i += blockDim.x) smem[i] = 0u;
    for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) { hkeys[i] = -1; hvals[i] = 0.0f; }
    __syncthreads();
    if (warp < 2) {
        // environment.py:246-249: beam along the current direction, and along (1,0,0)
        const double horiz[3] = {1.0, 0.0, 0.0};
        const Beam b = ray_prepare(G, r.pos, warp == 0 ? r.dir : horiz, lane, view[warp]);
        for (int kbase = 0; kbase < b.nslab; kbase += kWarp) {
            int lin[4], c0, c1, c2;
            float w[4];
            slab_targets(G, b, view[warp].ys, view[warp].zs, kbase + lane, lin, w, c0, c1, c2);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (lin[j] >= 0) {
                    hash_add(hkeys, hvals, lin[j], w[j]);          // :250 current_beam + horizontal_beam_center
                    atomicOr(hit_bits + (lin[j] >> 5), 1u << (lin[j] & 31));
                }
        }
    } else {
        for (int k = threadIdx.x - 2 * kWarp; k < tm.n_vox; k += blockDim.x - 2 * kWarp) {
            const uint32_t pk = __ldg(T.vox_xyz + tm.vox_off + k);
            const int lin = ((int)(pk & 255u) * G.g1 + (int)((pk >> 8) & 255u)) * G.g2 + (int)(pk >> 16);
            atomicOr(tum_bits + (lin >> 5), 1u << (lin & 31));
        }
    }
    __syncthreads();
    const float2 *vol2 = kPacked || !D.dense ? nullptr : reinterpret_cast<const float2 *>(D.dense + (size_t)env * G.vstride);
    const uint2 *cellv = kPacked || !D.cells ? nullptr : D.cells + (size_t)env * G.cstride;
    const uint32_t gen = kPacked ? 0u : r.gen;
    const __nv_bfloat162 *pk2 = kPacked ? reinterpret_cast<const __nv_bfloat162 *>(P.dose + (size_t)env * G.vstride) : nullptr;
    float *o = out + (size_t)blockIdx.x * 4 * G.nvox;
    float2 *o0 = reinterpret_cast<float2 *>(o), *o1 = reinterpret_cast<float2 *>(o + (size_t)G.nvox);
    float2 *o2 = reinterpret_cast<float2 *>(o + (size_t)2 * G.nvox), *o3 = reinterpret_cast<float2 *>(o + (size_t)3 * G.nvox);
    const int npairs = G.nvox / 2;                               // rt_assemble_volumes requires an even V
    constexpr int kUnroll = 4;
    for (int q0 = threadIdx.x; q0 < npairs; q0 += kUnroll * kVolThreads) {
        float2 d[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int q = q0 + u * kVolThreads;
            d[u] = make_float2(0.f, 0.f);
            if (q < npairs) {
                if (kPacked) {
                    d[u] = __bfloat1622float2(pk2[q]);
                } else if (cellv) {
                    // voxels (2q, 2q + 1) share a row (g2 is even here: nvox even and the pair does not straddle) and a brick
                    // (the other rows of the brick are read a moment later: keep the line in the caches)
                    const uint4 c = __ldg(reinterpret_cast<const uint4 *>(cellv + cell_index_lin(G, 2 * q)));   // another generation reads as zero
                    d[u] = make_float2(c.y == gen ? __uint_as_float(c.x) : 0.0f, c.w == gen ? __uint_as_float(c.z) : 0.0f);
                } else {
                    d[u] = __ldcs(vol2 + q);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int q = q0 + u * kVolThreads;
            if (q >= npairs) break;
            const int v = 2 * q, sh = v & 31;
            const uint32_t hb = (hit_bits[v >> 5] >> sh) & 3u;
            const uint32_t tb = (tum_bits[v >> 5] >> sh) & 3u;
            const uint32_t lb = (__ldg(T.lungs_bits + (v >> 5)) >> sh) & 3u;
            float view2[2] = {0.0f, 0.0f};
            if (hb) {
#pragma unroll
                for (int i = 0; i < 2; i++)
                    if ((hb >> i) & 1u) view2[i] = hash_get(hkeys, hvals, v + i);
            }
            o0[q] = (make_float2((lb & 1u) ? 1.0f : 0.0f, (lb & 2u) ? 1.0f : 0.0f));
            o1[q] = (make_float2((tb & 1u) ? 1.0f : 0.0f, (tb & 2u) ? 1.0f : 0.0f));
            o2[q] = (make_float2(fminf(fmaxf(d[u].x, 0.0f), 1.0f), fminf(fmaxf(d[u].y, 0.0f), 1.0f)));
            o3[q] = (make_float2(fminf(fmaxf(view2[0], 0.0f), 1.0f), fminf(fmaxf(view2[1], 0.0f), 1.0f)));
        }
    }
}

// ---------------------------------------------------------------------------------
// Compressed observation record of live envs [first, first + count): the dose volume as bfloat16 (cells of an earlier
// generation are zero), the pose and the tumour id: 403,392 B; with the three regenerated planes a stored voxel
// observation costs 1/8 of its float32 [4][V] form.
__global__ void __launch_bounds__(256) rt_pack_kernel(Grid G, const EnvRec *rec, const uint2 *cells, int first, long long slot0,
                                                      __nv_bfloat16 *out_dose, double *out_pose, int32_t *out_tid)
{
    const int env = first + blockIdx.y;
    const long long slot = slot0 + blockIdx.y;
    const uint2 *vol = cells + (size_t)env * G.cstride;
    const uint32_t gen = rec[env].gen;
    uint32_t *o = reinterpret_cast<uint32_t *>(out_dose + (size_t)slot * G.vstride);
    const int npairs = G.vstride / 2, vpairs = G.nvox / 2;               // g2 is even: a pair never straddles a row
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += gridDim.x * blockDim.x) {
        float v0 = 0.0f, v1 = 0.0f;
        if (q < vpairs) {
            const uint4 c = __ldg(reinterpret_cast<const uint4 *>(vol + cell_index_lin(G, 2 * q)));
            v0 = c.y == gen ? __uint_as_float(c.x) : 0.0f;
            v1 = c.w == gen ? __uint_as_float(c.z) : 0.0f;
        }
        const __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1);
        o[q] = *reinterpret_cast<const uint32_t *>(&a);
    }
    if (blockIdx.x == 0 && threadIdx.x < 6)
        out_pose[slot * 6 + threadIdx.x] = threadIdx.x < 3 ? rec[env].pos[threadIdx.x] : rec[env].dir[threadIdx.x - 3];
    if (blockIdx.x == 0 && threadIdx.x == 6) out_tid[slot] = rec[env].tumour_id;
}

// ---------------------------------------------------------------------------------
// GAE (train.py:164-181): one thread per env walks t = T-1 .. 0.  float32, one rounding per operation, in
// the order torch evaluates the reference expression.  The recurrence is serial in t, the loads are not: each
// thread keeps kGaeDepth timesteps (3 loads each) in flight, which is what bounds the achieved bandwidth.
constexpr int kGaeDepth = 16;

__global__ void __launch_bounds__(128) rt_gae_kernel(const float *__restrict__ rewards,
                                                     const float *__restrict__ values,
                                                     const float *__restrict__ dones,
                                                     const float *__restrict__ next_value,
                                                     const float *__restrict__ next_done, int T, int N, float g,
                                                     float gl, float *__restrict__ adv, float *__restrict__ ret)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float last = 0.0f;
    float nnt = __fsub_rn(1.0f, next_done[i]);
    float nv = next_value[i];
    for (int t0 = T - 1; t0 >= 0; t0 -= kGaeDepth) {
        float r[kGaeDepth], v[kGaeDepth], d[kGaeDepth];
#pragma unroll
        for (int u = 0; u < kGaeDepth; u++) {
            const int t = t0 - u;
            if (t >= 0) {
                const size_t k = (size_t)t * N + i;
                r[u] = __ldcs(rewards + k);
                v[u] = __ldcs(values + k);
                d[u] = __ldcs(dones + k);
            }
        }
#pragma unroll
        for (int u = 0; u < kGaeDepth; u++) {
            const int t = t0 - u;
            if (t < 0) break;
            const size_t k = (size_t)t * N + i;
            const float delta = __fsub_rn(__fadd_rn(r[u], __fmul_rn(__fmul_rn(g, nv), nnt)), v[u]);
            last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), last));
            __stcs(adv + k, last);
            __stcs(ret + k, __fadd_rn(last, v[u]));
            nnt = __fsub_rn(1.0f, d[u]);      // dones[t] gates step t-1
            nv = v[u];
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------
// Host side.
struct rt_env {
    int device = 0;
    int n = 0;
    uint32_t flags = 0;
    Tables T{};
    Schedule S{};
    EnvRec *rec = nullptr;
    uint2 *cells = nullptr;   // sparse mode: {dose, generation} per voxel
    float *dose = nullptr;    // dense mode: float32 volumes
    double *beams = nullptr;
    DenseWork *dense = nullptr;
    size_t dense_smem = 0;
    size_t step_smem = 0;
    int step_kb = 14;         // envs per block of rt_step_kernel: 7 while one block per SM covers the envs, else 14
    bool use_pdl = true;      // programmatic dependent launch of consecutive steps (rt_set_pdl)
    uint32_t *d_lungs = nullptr;
    Tumour *d_tumours = nullptr;
    uint32_t *d_tbits = nullptr;
    uint32_t *d_ptbits = nullptr;
    uint32_t *d_vox = nullptr;
    int32_t *d_sched = nullptr;
    int64_t bytes = 0;
    // staging for the *_host calls
    cudaStream_t hstream = nullptr;
    float *h_actions = nullptr, *h_obs = nullptr;
    double *h_reward = nullptr, *h_info = nullptr;
    uint8_t *h_term = nullptr, *h_trunc = nullptr, *h_mask = nullptr;
    // Device-pointer entry points run on the caller's stream, the *_host entry points on hstream: a host call that
    // follows device-side work waits for it first (the reverse order needs nothing: host calls return drained).
    bool dev_pending = false;
};

namespace {

Grid make_grid(const int32_t g[3])
{
    Grid G;
    G.g0 = g[0]; G.g1 = g[1]; G.g2 = g[2];
    G.nvox = g[0] * g[1] * g[2];
    G.vstride = (G.nvox + 31) / 32 * 32;
    G.nb1 = (G.g1 + 1) / 2;
    G.nb2 = (G.g2 + 3) / 4;
    G.cstride = ((G.g0 + 1) / 2) * G.nb1 * G.nb2 * 16;
    G.mg1 = (uint32_t)(((1ull << 32) + (uint64_t)G.g1 - 1) / (uint64_t)G.g1);
    G.mg2 = (uint32_t)(((1ull << 32) + (uint64_t)G.g2 - 1) / (uint64_t)G.g2);
    return G;
}

int check_grid(const int32_t g[3])
{
    if (!g) return fail(RT_ERR_INVALID, "grid is NULL");
    for (int i = 0; i < 3; i++)
        if (g[i] < 2 || g[i] > 255) return fail(RT_ERR_INVALID, "grid extents must be in [2, 255]");
    int mx = g[0] > g[1] ? g[0] : g[1];
    mx = mx > g[2] ? mx : g[2];
    if (mx + 1 > kMaxSlabs) return fail(RT_ERR_INVALID, "grid extent too large for the slab walk");
    return RT_OK;
}

template <typename T>
int dev_alloc(T **p, size_t count, int64_t *bytes)
{
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T));
    if (e != cudaSuccess) return fail(RT_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *bytes += (int64_t)(count * sizeof(T));
    return RT_OK;
}

inline int warps_grid(int n_warps, int threads) { return (int)(((int64_t)n_warps * kWarp + threads - 1) / threads); }

}  // namespace

extern "C" {

int rt_abi_version(void) { return RT_ABI_VERSION; }
const char *rt_last_error(void) { return g_err.c_str(); }
int64_t rt_launch_count(void) { return g_launches.load(); }

int rt_create(rt_env **out, int device, int n_envs, uint32_t flags, const rt_phantom_desc *ph)
{
    if (!out || !ph) return fail(RT_ERR_INVALID, "rt_create: NULL argument");
    *out = nullptr;
    if (n_envs < 1) return fail(RT_ERR_INVALID, "rt_create: n_envs must be >= 1");
    if (int rc = check_grid(ph->grid)) return rc;
    if (ph->n_tumours < 1 || !ph->lungs_bits || !ph->vox_offsets || !ph->vox || !ph->centroid ||
        !ph->tumour_sum || !ph->lung_mask_sum)
        return fail(RT_ERR_INVALID, "rt_create: incomplete phantom description");
    RT_CUDA(cudaSetDevice(device));

    rt_env *e = new rt_env();
    e->device = device;
    e->n = n_envs;
    e->flags = flags;
    const Grid G = make_grid(ph->grid);
    e->T.G = G;
    e->T.gnorm = sqrt((double)(G.g0 * G.g0 + G.g1 * G.g1 + G.g2 * G.g2));
    e->T.n_tumours = ph->n_tumours;
    e->T.stage_clock = nullptr;

    // tumour table: bbox, bbox-local bitmask, packed voxel coordinates
    std::vector<Tumour> tum(ph->n_tumours);
    const int total_vox = ph->vox_offsets[ph->n_tumours];
    std::vector<uint32_t> vox_xyz((size_t)(total_vox > 0 ? total_vox : 1));
    int max_bits = 1;
    for (int t = 0; t < ph->n_tumours; t++) {
        Tumour &tm = tum[t];
        const int lo = ph->vox_offsets[t], hi = ph->vox_offsets[t + 1];
        if (hi <= lo) { delete e; return fail(RT_ERR_INVALID, "rt_create: empty tumour"); }
        int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-1, -1, -1};
        for (int k = lo; k < hi; k++) {
            const int v = ph->vox[k];
            if (v < 0 || v >= G.nvox || (k > lo && v <= ph->vox[k - 1])) {
                delete e;
                return fail(RT_ERR_INVALID, "rt_create: tumour voxels must be ascending and inside the grid");
            }
            const int c[3] = {v / (G.g1 * G.g2), (v / G.g2) % G.g1, v % G.g2};
            vox_xyz[k] = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
            for (int a = 0; a < 3; a++) { mn[a] = c[a] < mn[a] ? c[a] : mn[a]; mx[a] = c[a] > mx[a] ? c[a] : mx[a]; }
        }
        for (int a = 0; a < 3; a++) { tm.lo[a] = mn[a]; tm.dim[a] = mx[a] - mn[a] + 1; }
        tm.n_vox = hi - lo;
        tm.vox_off = lo;
        tm.tumour_sum = ph->tumour_sum[t];
        tm.lung_mask_sum = ph->lung_mask_sum[t];
        for (int a = 0; a < 3; a++) {
            // environment.py:261: tumour_position() / LUNG_SHAPE * 2 - 1, one rounding per operation, then float32
            volatile double q = ph->centroid[3 * t + a] / (double)ph->grid[a];
            volatile double m2 = q * 2.0;
            volatile double v = m2 - 1.0;
            tm.obs_c[a] = (float)v;
            tm.pad_[a] = 0;
        }
        const int bits = tm.dim[0] * tm.dim[1] * tm.dim[2];
        max_bits = bits > max_bits ? bits : max_bits;
    }
    e->T.bits_words = (max_bits + 31) / 32;
    if (e->T.bits_words > kMaxTumourWords) {
        delete e;
        return fail(RT_ERR_INVALID, "rt_create: a tumour's bounding box exceeds 2048 voxels");
    }
    std::vector<uint32_t> tbits((size_t)ph->n_tumours * e->T.bits_words, 0u);
    for (int t = 0; t < ph->n_tumours; t++) {
        const Tumour &tm = tum[t];
        for (int k = ph->vox_offsets[t]; k < ph->vox_offsets[t + 1]; k++) {
            const uint32_t pk = vox_xyz[k];
            const int li = (int)(pk & 255u) - tm.lo[0], lj = (int)((pk >> 8) & 255u) - tm.lo[1],
                      lk = (int)(pk >> 16) - tm.lo[2];
            const int b = (li * tm.dim[1] + lj) * tm.dim[2] + lk;
            tbits[(size_t)t * e->T.bits_words + (b >> 5)] |= 1u << (b & 31);
        }
    }

    // padded variant for rt_step3_kernel: bit ((li*(d1+2)) + lj+1)*(d2+2) + lk+1, one spare zero word at the end
    int max_pbits = 1;
    for (int t = 0; t < ph->n_tumours; t++) {
        const Tumour &tm = tum[t];
        const int pb = tm.dim[0] * (tm.dim[1] + 2) * (tm.dim[2] + 2);
        max_pbits = pb > max_pbits ? pb : max_pbits;
    }
    e->T.pbits_words = ((max_pbits + 31) / 32 + 1 + 3) & ~3;        // rows are bulk-copied: multiples of 16 bytes
    if (e->T.pbits_words > kMaxPTumourWords) {
        delete e;
        return fail(RT_ERR_INVALID, "rt_create: a tumour's padded bounding box exceeds 3040 voxels");
    }
    std::vector<uint32_t> ptbits((size_t)ph->n_tumours * e->T.pbits_words, 0u);
    for (int t = 0; t < ph->n_tumours; t++) {
        const Tumour &tm = tum[t];
        for (int k = ph->vox_offsets[t]; k < ph->vox_offsets[t + 1]; k++) {
            const uint32_t pk = vox_xyz[k];
            const int li = (int)(pk & 255u) - tm.lo[0], lj = (int)((pk >> 8) & 255u) - tm.lo[1],
                      lk = (int)(pk >> 16) - tm.lo[2];
            const int b = (li * (tm.dim[1] + 2) + lj + 1) * (tm.dim[2] + 2) + lk + 1;
            ptbits[(size_t)t * e->T.pbits_words + (b >> 5)] |= 1u << (b & 31);
        }
    }

    // device copy of the voxel lists: every list starts at a multiple of four entries and is padded to one by
    // repeating its last voxel (16-byte bulk copies; a repeated voxel does not change the minimum distance)
    std::vector<uint32_t> vox_dev;
    vox_dev.reserve(vox_xyz.size() + 4 * (size_t)ph->n_tumours);
    for (int t = 0; t < ph->n_tumours; t++) {
        tum[t].vox_off = (int)vox_dev.size();
        for (int k = ph->vox_offsets[t]; k < ph->vox_offsets[t + 1]; k++) vox_dev.push_back(vox_xyz[k]);
        while (vox_dev.size() % 4) vox_dev.push_back(vox_dev.back());
    }
    if (vox_dev.size() / 4 >= (1u << 24)) { delete e; return fail(RT_ERR_INVALID, "rt_create: voxel lists too long"); }

    int rc = RT_OK;
    const size_t lung_words = (size_t)(G.nvox + 31) / 32;
    // device copy of the lungs bitmask: kLungPadBits zero bits in front, at least two zero words behind (bit_pair)
    const size_t lung_pad_words = kLungPadBits / 32;
    e->T.lung_words16 = (int)((lung_pad_words + lung_words + 2 + 3) / 4 * 4);
    const bool dense_mode = (flags & RT_FLAG_DENSE) != 0;
    if ((rc = dev_alloc(&e->d_lungs, (size_t)e->T.lung_words16, &e->bytes)) || (rc = dev_alloc(&e->d_tumours, tum.size(), &e->bytes)) ||
        (rc = dev_alloc(&e->d_tbits, tbits.size(), &e->bytes)) || (rc = dev_alloc(&e->d_ptbits, ptbits.size(), &e->bytes)) || (rc = dev_alloc(&e->d_vox, vox_dev.size(), &e->bytes)) ||
        (rc = dev_alloc(&e->rec, (size_t)n_envs, &e->bytes)) ||
        (rc = dense_mode ? dev_alloc(&e->dose, (size_t)n_envs * G.vstride, &e->bytes)
                         : dev_alloc(&e->cells, (size_t)n_envs * G.cstride, &e->bytes))) {
        rt_destroy(e);
        return rc;
    }
    if (flags & RT_FLAG_RECORD_BEAMS)
        if ((rc = dev_alloc(&e->beams, (size_t)n_envs * RT_MAX_TIME_STEPS * 6, &e->bytes))) { rt_destroy(e); return rc; }
    {
        // Envs per block of the step kernel: 7 while that covers the envs with one block per SM (four blocks fit an
        // SM, lungs bitmask through L1), else 14 (two blocks per SM, lungs bitmask staged in shared memory).
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const int per_sm = (n_envs + sms - 1) / sms;
        e->step_kb = per_sm <= 7 ? 7 : 14;
        // dynamic shared memory of a block: the item slots of its envs (kMaxPass x 32 slabs x 4 targets x 8 bytes each) and,
        // for 14-env blocks, the padded lungs bitmask
        e->step_smem = dense_mode ? 0 : (size_t)e->step_kb * kMaxPass * 4 * kWarp * sizeof(uint2) +
                                        (e->step_kb >= 14 ? (size_t)e->T.lung_words16 * sizeof(uint32_t) : 0);
        cudaError_t ae = cudaSuccess;
        if (e->step_smem) {
            if (e->step_kb == 7)
                ae = cudaFuncSetAttribute(rt_step_kernel<7, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->step_smem);
            else {
                ae = cudaFuncSetAttribute(rt_step_kernel<14, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->step_smem);
                if (ae == cudaSuccess)
                    ae = cudaFuncSetAttribute(rt_step_kernel<14, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->step_smem);
            }
        }
        if (ae != cudaSuccess) { rt_destroy(e); return fail(RT_ERR_CUDA, std::string("rt_step_kernel smem: ") + cudaGetErrorString(ae)); }
    }
    if (dense_mode) {
        int gmax = G.g0 > G.g1 ? G.g0 : G.g1;
        gmax = gmax > G.g2 ? gmax : G.g2;
        if (4 * (gmax + 1) > RT_BEAM_CAP) { rt_destroy(e); return fail(RT_ERR_INVALID, "rt_create: dense mode needs 4*(max(grid)+1) <= RT_BEAM_CAP"); }
        if ((rc = dev_alloc(&e->dense, (size_t)n_envs, &e->bytes))) { rt_destroy(e); return rc; }
        e->dense_smem = (size_t)(2 * (G.vstride / 32)) * sizeof(uint32_t) + (size_t)kHashSlots * (sizeof(int) + sizeof(float));
        cudaError_t ae = cudaFuncSetAttribute(rt_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->dense_smem);
        if (ae != cudaSuccess) { rt_destroy(e); return fail(RT_ERR_CUDA, std::string("rt_dense_kernel smem: ") + cudaGetErrorString(ae)); }
    }
    cudaError_t ce = cudaSuccess;
    auto chk = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
    chk(cudaMemset(e->d_lungs, 0, (size_t)e->T.lung_words16 * sizeof(uint32_t)));
    chk(cudaMemcpy(e->d_lungs + lung_pad_words, ph->lungs_bits, lung_words * sizeof(uint32_t), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(e->d_tumours, tum.data(), tum.size() * sizeof(Tumour), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(e->d_tbits, tbits.data(), tbits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(e->d_ptbits, ptbits.data(), ptbits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(e->d_vox, vox_dev.data(), vox_dev.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    chk(cudaMemset(e->rec, 0, (size_t)n_envs * sizeof(EnvRec)));
    // every cell starts in generation 0xffffffff, which no env ever reaches (EnvRec::gen counts resets from 0): the
    // volumes read as zero without being written; dense-mode volumes are zeroed by the first rt_reset
    if (e->cells) chk(cudaMemset(e->cells, 0xff, (size_t)n_envs * G.cstride * sizeof(uint2)));
    chk(cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking));
    chk(cudaMallocHost(&e->h_actions, (size_t)n_envs * RT_ACTION_SIZE * sizeof(float)));
    chk(cudaMallocHost(&e->h_obs, (size_t)n_envs * RT_OBS_SIZE * sizeof(float)));
    chk(cudaMallocHost(&e->h_reward, (size_t)n_envs * sizeof(double)));
    chk(cudaMallocHost(&e->h_info, (size_t)n_envs * RT_INFO_SIZE * sizeof(double)));
    chk(cudaMallocHost(&e->h_term, (size_t)n_envs));
    chk(cudaMallocHost(&e->h_trunc, (size_t)n_envs));
    chk(cudaMallocHost(&e->h_mask, (size_t)n_envs));
    if (ce != cudaSuccess) {
        rt_destroy(e);
        return fail(RT_ERR_CUDA, std::string("rt_create: ") + cudaGetErrorString(ce));
    }
    e->T.lungs_pad = e->d_lungs;
    e->T.lungs_bits = e->d_lungs + lung_pad_words;
    e->T.tumours = e->d_tumours;
    e->T.tumour_bits = e->d_tbits;
    e->T.tumour_pbits = e->d_ptbits;
    e->T.vox_xyz = e->d_vox;
    e->S.ids = nullptr;
    e->S.n_episodes = 0;
    e->S.seed = 0;
    *out = e;
    return RT_OK;
}

int rt_destroy(rt_env *e)
{
    if (!e) return RT_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    cudaFree(e->d_lungs); cudaFree(e->d_tumours); cudaFree(e->d_tbits); cudaFree(e->d_ptbits); cudaFree(e->d_vox);
    cudaFree(e->rec); cudaFree(e->cells); cudaFree(e->dose); cudaFree(e->beams); cudaFree(e->dense); cudaFree(e->d_sched);
    // the *_host staging buffers are device-visible pinned allocations of the same sizes
    cudaFreeHost(e->h_actions); cudaFreeHost(e->h_obs); cudaFreeHost(e->h_reward); cudaFreeHost(e->h_info);
    cudaFreeHost(e->h_term); cudaFreeHost(e->h_trunc); cudaFreeHost(e->h_mask);
    if (e->hstream) cudaStreamDestroy(e->hstream);
    g_alias_epoch.fetch_add(1, std::memory_order_relaxed);      // addresses of freed pinned buffers may be reused: drop cached aliases
    delete e;
    return RT_OK;
}

int rt_num_envs(const rt_env *e) { return e ? e->n : 0; }
int64_t rt_device_bytes(const rt_env *e) { return e ? e->bytes : 0; }

int rt_seed(rt_env *e, uint64_t seed)
{
    if (!e) return fail(RT_ERR_INVALID, "rt_seed: NULL handle");
    e->S.seed = seed;
    return RT_OK;
}

int rt_set_tumour_schedule(rt_env *e, const int32_t *ids_host, int n_episodes)
{
    if (!e) return fail(RT_ERR_INVALID, "rt_set_tumour_schedule: NULL handle");
    RT_CUDA(cudaSetDevice(e->device));
    RT_CUDA(cudaDeviceSynchronize());
    if (e->d_sched) { cudaFree(e->d_sched); e->d_sched = nullptr; }
    e->S.ids = nullptr;
    e->S.n_episodes = 0;
    if (!ids_host) return RT_OK;
    if (n_episodes < 1) return fail(RT_ERR_INVALID, "rt_set_tumour_schedule: n_episodes must be >= 1");
    const size_t cnt = (size_t)n_episodes * e->n;
    for (size_t k = 0; k < cnt; k++)
        if (ids_host[k] < 0 || ids_host[k] >= e->T.n_tumours)
            return fail(RT_ERR_INVALID, "rt_set_tumour_schedule: tumour id out of range");
    RT_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->d_sched), cnt * sizeof(int32_t)));
    RT_CUDA(cudaMemcpy(e->d_sched, ids_host, cnt * sizeof(int32_t), cudaMemcpyHostToDevice));
    e->S.ids = e->d_sched;
    e->S.n_episodes = n_episodes;
    return RT_OK;
}

// A *_host call runs on the handle's own stream; if device-pointer calls were issued since the last one (on
// whatever stream the caller used), wait for them first.
static int sync_before_host_call(rt_env *e)
{
    if (e->dev_pending) {
        RT_CUDA(cudaDeviceSynchronize());
        e->dev_pending = false;
    }
    return RT_OK;
}

int rt_reset(rt_env *e, const uint8_t *mask_dev, float *obs_dev, void *stream)
{
    if (!e) return fail(RT_ERR_INVALID, "rt_reset: NULL handle");
    RT_CUDA(cudaSetDevice(e->device));
    if ((cudaStream_t)stream != e->hstream) e->dev_pending = true;
    rt_reset_kernel<<<warps_grid(e->n, 256), 256, 0, (cudaStream_t)stream>>>(e->T, e->S, e->rec, e->n, mask_dev, obs_dev, e->dense);
    RT_LAUNCH_CHECK("rt_reset_kernel");
    if (e->dense) {
        StepOut none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        rt_dense_kernel<<<e->n, kDenseThreads, e->dense_smem, (cudaStream_t)stream>>>(e->T, e->rec, e->dose, e->dense, none);
        RT_LAUNCH_CHECK("rt_dense_kernel");
    }
    return RT_OK;
}

int rt_step(rt_env *e, const float *actions_dev, float *obs_dev, double *reward_dev, float *reward_f32_dev,
            uint8_t *terminated_dev, uint8_t *truncated_dev, double *info_dev, void *stream)
{
    if (!e || !actions_dev || !obs_dev) return fail(RT_ERR_INVALID, "rt_step: NULL handle, actions or obs");
    RT_CUDA(cudaSetDevice(e->device));
    if ((cudaStream_t)stream != e->hstream) e->dev_pending = true;
    StepOut o{obs_dev, reward_dev, reward_f32_dev, terminated_dev, truncated_dev, info_dev};
    const int kb = e->step_kb;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((e->n + kb - 1) / kb);
    cfg.blockDim = dim3((kb + 1 + (kb >= 14 && !e->dense ? 1 : 0)) * kWarp);     // scalar warp, env warps, predictor warp
    cfg.dynamicSmemBytes = e->step_smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = e->use_pdl && !e->dense ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (e->dense) {
        if (kb == 7)
            RT_CUDA(cudaLaunchKernelEx(&cfg, rt_step_kernel<7, false, true>, e->T, e->S, e->rec, e->cells, e->beams, e->n, actions_dev, o, e->dense));
        else
            RT_CUDA(cudaLaunchKernelEx(&cfg, rt_step_kernel<14, false, true>, e->T, e->S, e->rec, e->cells, e->beams, e->n, actions_dev, o, e->dense));
        RT_LAUNCH_CHECK("rt_step_kernel<dense>");
        rt_dense_kernel<<<e->n, kDenseThreads, e->dense_smem, (cudaStream_t)stream>>>(e->T, e->rec, e->dose, e->dense, o);
        RT_LAUNCH_CHECK("rt_dense_kernel");
        return RT_OK;
    }
    if (kb == 7)
        RT_CUDA(cudaLaunchKernelEx(&cfg, rt_step_kernel<7, false, false>, e->T, e->S, e->rec, e->cells, e->beams, e->n, actions_dev, o, e->dense));
    else if (e->T.stage_clock)
        RT_CUDA(cudaLaunchKernelEx(&cfg, rt_step_kernel<14, true, false>, e->T, e->S, e->rec, e->cells, e->beams, e->n, actions_dev, o, e->dense));
    else
        RT_CUDA(cudaLaunchKernelEx(&cfg, rt_step_kernel<14, false, false>, e->T, e->S, e->rec, e->cells, e->beams, e->n, actions_dev, o, e->dense));
    RT_LAUNCH_CHECK("rt_step_kernel");
    return RT_OK;
}

int rt_reset_host(rt_env *e, const uint8_t *mask_host, float *obs_host)
{
    if (!e || !obs_host) return fail(RT_ERR_INVALID, "rt_reset_host: NULL argument");
    RT_CUDA(cudaSetDevice(e->device));
    if (int rc = sync_before_host_call(e)) return rc;
    uint8_t *dmask = nullptr;
    uint8_t *d_mask_dev = nullptr;
    if (mask_host) {
        memcpy(e->h_mask, mask_host, (size_t)e->n);
        RT_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_mask_dev), e->h_mask, 0));
        dmask = d_mask_dev;
    }
    float *d_obs = nullptr;
    RT_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_obs), e->h_obs, 0));
    if (int rc = rt_reset(e, dmask, d_obs, e->hstream)) return rc;
    RT_CUDA(cudaStreamSynchronize(e->hstream));
    memcpy(obs_host, e->h_obs, (size_t)e->n * RT_OBS_SIZE * sizeof(float));
    return RT_OK;
}

int rt_set_pdl(rt_env *e, int enabled)
{
    if (!e) return fail(RT_ERR_INVALID, "rt_set_pdl: NULL handle");
    e->use_pdl = enabled != 0;
    return RT_OK;
}

int rt_set_stage_clock(rt_env *e, long long *stamps_dev)
{
    if (!e) return fail(RT_ERR_INVALID, "rt_set_stage_clock: NULL handle");
    if (stamps_dev && (e->step_kb != 14 || e->dense))
        return fail(RT_ERR_STATE, "rt_set_stage_clock: the instrumented step kernel exists for sparse-mode handles with 14 envs per block (more than 7 envs per SM)");
    e->T.stage_clock = stamps_dev;
    return RT_OK;
}

int rt_get_pose(rt_env *e, double *pose_dev, void *stream)
{
    if (!e || !pose_dev) return fail(RT_ERR_INVALID, "rt_get_pose: NULL argument");
    RT_CUDA(cudaSetDevice(e->device));
    rt_get_pose_kernel<<<(e->n * 6 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->rec, e->n, pose_dev);
    RT_LAUNCH_CHECK("rt_get_pose_kernel");
    return RT_OK;
}

int rt_set_pose(rt_env *e, const double *pose_dev, void *stream)
{
    if (!e || !pose_dev) return fail(RT_ERR_INVALID, "rt_set_pose: NULL argument");
    RT_CUDA(cudaSetDevice(e->device));
    e->dev_pending = true;
    rt_set_pose_kernel<<<(e->n * 6 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->rec, e->n, pose_dev);
    RT_LAUNCH_CHECK("rt_set_pose_kernel");
    return RT_OK;
}

int rt_get_counters(rt_env *e, int32_t *counters_dev, void *stream)
{
    if (!e || !counters_dev) return fail(RT_ERR_INVALID, "rt_get_counters: NULL argument");
    RT_CUDA(cudaSetDevice(e->device));
    rt_get_counters_kernel<<<(e->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->rec, e->n, counters_dev);
    RT_LAUNCH_CHECK("rt_get_counters_kernel");
    return RT_OK;
}

int rt_get_dose(rt_env *e, int env_index, float *dose_dev, void *stream)
{
    if (!e || !dose_dev) return fail(RT_ERR_INVALID, "rt_get_dose: NULL argument");
    if (env_index < 0 || env_index >= e->n) return fail(RT_ERR_INVALID, "rt_get_dose: env index out of range");
    RT_CUDA(cudaSetDevice(e->device));
    rt_get_dose_kernel<<<(e->T.G.nvox + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->T.G, e->rec, DoseSrc{e->cells, e->dose},
                                                                                  env_index, dose_dev);
    RT_LAUNCH_CHECK("rt_get_dose_kernel");
    return RT_OK;
}

int rt_get_beams(rt_env *e, int env_index, double *beams_dev, int32_t *n_dev, void *stream)
{
    if (!e || !beams_dev) return fail(RT_ERR_INVALID, "rt_get_beams: NULL argument");
    if (!e->beams) return fail(RT_ERR_STATE, "rt_get_beams: handle was created without RT_FLAG_RECORD_BEAMS");
    if (env_index < 0 || env_index >= e->n) return fail(RT_ERR_INVALID, "rt_get_beams: env index out of range");
    RT_CUDA(cudaSetDevice(e->device));
    rt_get_beams_kernel<<<(RT_MAX_TIME_STEPS * 6 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        e->rec, e->beams, env_index, beams_dev, n_dev);
    RT_LAUNCH_CHECK("rt_get_beams_kernel");
    return RT_OK;
}

int rt_assemble_volumes(rt_env *e, int first, int count, float *obs_dev, void *stream)
{
    if (!e || !obs_dev) return fail(RT_ERR_INVALID, "rt_assemble_volumes: NULL argument");
    if (first < 0 || count < 1 || first + count > e->n)
        return fail(RT_ERR_INVALID, "rt_assemble_volumes: env range out of bounds");
    RT_CUDA(cudaSetDevice(e->device));
    if (e->T.G.nvox % 2 || (e->cells && e->T.G.g2 % 2))
        return fail(RT_ERR_INVALID, "rt_assemble_volumes: the voxel count (sparse mode: the last grid extent) must be even");
    const size_t smem = (size_t)(2 * (e->T.G.vstride / 32)) * sizeof(uint32_t) + (size_t)kHashSlots * (sizeof(int) + sizeof(float));
    static bool attr_set = false;
    if (!attr_set) {
        RT_CUDA(cudaFuncSetAttribute(rt_volumes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RT_CUDA(cudaFuncSetAttribute(rt_volumes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    rt_volumes_kernel<false><<<count, kVolThreads, smem, (cudaStream_t)stream>>>(e->T, e->rec, DoseSrc{e->cells, e->dose}, first,
                                                                                obs_dev, PackedObs{});
    RT_LAUNCH_CHECK("rt_volumes_kernel");
    return RT_OK;
}

int rt_observation_record_stride(const rt_env *e) { return e ? e->T.G.vstride : 0; }

int rt_pack_observations(rt_env *e, int first, int count, int64_t slot0, void *dose_bf16_dev, double *pose_dev,
                         int32_t *tumour_id_dev, void *stream)
{
    if (!e || !dose_bf16_dev || !pose_dev || !tumour_id_dev) return fail(RT_ERR_INVALID, "rt_pack_observations: NULL argument");
    if (count == 0) return RT_OK;
    if (first < 0 || count < 0 || first + count > e->n || slot0 < 0)
        return fail(RT_ERR_INVALID, "rt_pack_observations: env range out of bounds");
    if (e->dense) return fail(RT_ERR_STATE, "rt_pack_observations: not available for dense-mode handles");
    if (e->T.G.g2 % 2) return fail(RT_ERR_INVALID, "rt_pack_observations: the last grid extent must be even");
    RT_CUDA(cudaSetDevice(e->device));
    rt_pack_kernel<<<dim3(8, count), 256, 0, (cudaStream_t)stream>>>(e->T.G, e->rec, e->cells, first, (long long)slot0,
                                                                    reinterpret_cast<__nv_bfloat16 *>(dose_bf16_dev), pose_dev,
                                                                    tumour_id_dev);
    RT_LAUNCH_CHECK("rt_pack_kernel");
    return RT_OK;
}

int rt_render_observations(rt_env *e, const void *dose_bf16_dev, const double *pose_dev, const int32_t *tumour_id_dev,
                           const int64_t *index_dev, int count, float *obs_dev, void *stream)
{
    if (!e || !dose_bf16_dev || !pose_dev || !tumour_id_dev || !obs_dev)
        return fail(RT_ERR_INVALID, "rt_render_observations: NULL argument");
    if (count == 0) return RT_OK;
    if (count < 0) return fail(RT_ERR_INVALID, "rt_render_observations: negative count");
    RT_CUDA(cudaSetDevice(e->device));
    if (e->T.G.nvox % 2) return fail(RT_ERR_INVALID, "rt_render_observations: the voxel count must be even");
    const size_t smem = (size_t)(2 * (e->T.G.vstride / 32)) * sizeof(uint32_t) + (size_t)kHashSlots * (sizeof(int) + sizeof(float));
    static bool attr_set = false;
    if (!attr_set) {
        RT_CUDA(cudaFuncSetAttribute(rt_volumes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    PackedObs P{reinterpret_cast<const __nv_bfloat16 *>(dose_bf16_dev), pose_dev, tumour_id_dev, index_dev};
    rt_volumes_kernel<true><<<count, kVolThreads, smem, (cudaStream_t)stream>>>(e->T, nullptr, DoseSrc{nullptr, nullptr}, 0, obs_dev, P);
    RT_LAUNCH_CHECK("rt_volumes_kernel<packed>");
    return RT_OK;
}

// Device alias of a host pointer if (and only if) it is page-locked, else NULL.  The answer is cached per address
// and thread (the attribute query costs about a microsecond and a step passes six buffers); the cache is dropped
// whenever a handle is destroyed, because its pinned staging buffers are freed and their addresses can come back as
// pageable memory.  Callers that free their own pinned buffers while a handle lives must not reuse the addresses for
// pageable arrays passed to the *_host calls (documented in rt_env.h).
static void *pinned_alias(const void *host)
{
    struct Entry { const void *host; void *dev; };
    static thread_local Entry cache[16];
    static thread_local int n_cached = 0;
    static thread_local unsigned epoch = 0;
    if (!host) return nullptr;
    const unsigned now = g_alias_epoch.load(std::memory_order_relaxed);
    if (now != epoch) { n_cached = 0; epoch = now; }
    for (int i = 0; i < n_cached; i++)
        if (cache[i].host == host) return cache[i].dev;
    cudaPointerAttributes at;
    void *dev = nullptr;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess) dev = at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
    else cudaGetLastError();
    if (dev && n_cached < 16) cache[n_cached++] = Entry{host, dev};     // only pinned buffers are worth remembering
    return dev;
}

int rt_step_host(rt_env *e, const float *actions_host, float *obs_host, double *reward_host,
                 uint8_t *terminated_host, uint8_t *truncated_host, double *info_host)
{
    if (!e || !actions_host || !obs_host) return fail(RT_ERR_INVALID, "rt_step_host: NULL argument");
    RT_CUDA(cudaSetDevice(e->device));
    if (int rc = sync_before_host_call(e)) return rc;
    // Host buffers are reached zero-copy: the step kernel reads the actions and writes its
    // results straight over PCIe into page-locked memory (one launch, no copy-engine round
    // trips).  Caller buffers that are already pinned (e.g. torch pin_memory) are used in
    // place; pageable ones are staged through the handle's own pinned buffers.
    const size_t n = (size_t)e->n;
    float *d_act = (float *)pinned_alias(actions_host);
    float *d_obs = (float *)pinned_alias(obs_host);
    double *d_rew = (double *)pinned_alias(reward_host);
    uint8_t *d_term = (uint8_t *)pinned_alias(terminated_host);
    uint8_t *d_trunc = (uint8_t *)pinned_alias(truncated_host);
    double *d_info = (double *)pinned_alias(info_host);
    const bool s_act = !d_act, s_obs = !d_obs, s_rew = reward_host && !d_rew, s_term = terminated_host && !d_term,
               s_trunc = truncated_host && !d_trunc, s_info = info_host && !d_info;
    if (s_act) { memcpy(e->h_actions, actions_host, n * RT_ACTION_SIZE * sizeof(float)); d_act = (float *)pinned_alias(e->h_actions); }
    if (s_obs) d_obs = (float *)pinned_alias(e->h_obs);
    if (s_rew) d_rew = (double *)pinned_alias(e->h_reward);
    if (s_term) d_term = (uint8_t *)pinned_alias(e->h_term);
    if (s_trunc) d_trunc = (uint8_t *)pinned_alias(e->h_trunc);
    if (s_info) d_info = (double *)pinned_alias(e->h_info);
    if (!d_act || !d_obs) return fail(RT_ERR_CUDA, "rt_step_host: pinned staging buffers are not device-mapped");
    if (int rc = rt_step(e, d_act, d_obs, d_rew, nullptr, d_term, d_trunc, d_info, e->hstream)) return rc;
    RT_CUDA(cudaStreamSynchronize(e->hstream));
    if (s_obs) memcpy(obs_host, e->h_obs, n * RT_OBS_SIZE * sizeof(float));
    if (s_rew) memcpy(reward_host, e->h_reward, n * sizeof(double));
    if (s_term) memcpy(terminated_host, e->h_term, n);
    if (s_trunc) memcpy(truncated_host, e->h_trunc, n);
    if (s_info) memcpy(info_host, e->h_info, n * RT_INFO_SIZE * sizeof(double));
    return RT_OK;
}

int rt_beam_voxels(const int32_t grid[3], const double *pos_dev, const double *dir_dev, int m, int cap,
                   int32_t *idx_dev, float *w_dev, int32_t *count_dev, void *stream)
{
    if (int rc = check_grid(grid)) return rc;
    if (m == 0) return RT_OK;
    if (!pos_dev || !dir_dev || !idx_dev || !w_dev || !count_dev) return fail(RT_ERR_INVALID, "rt_beam_voxels: NULL argument");
    int gmax = grid[0] > grid[1] ? grid[0] : grid[1];
    gmax = gmax > grid[2] ? gmax : grid[2];
    if (m < 0 || cap < 4 * (gmax + 1)) return fail(RT_ERR_INVALID, "rt_beam_voxels: cap must be >= 4 * (max(grid) + 1)");
    if (m == 0) return RT_OK;
    rt_beam_kernel<<<warps_grid(m, 256), 256, 0, (cudaStream_t)stream>>>(make_grid(grid), pos_dev, dir_dev, m, cap,
                                                                        idx_dev, w_dev, count_dev);
    RT_LAUNCH_CHECK("rt_beam_kernel");
    return RT_OK;
}

int rt_beam_voxels_dense(const int32_t grid[3], const double *pos_dev, const double *dir_dev, int m, float *out_dev,
                         int32_t *status_dev, void *stream)
{
    if (int rc = check_grid(grid)) return rc;
    if (m == 0) return RT_OK;
    if (!pos_dev || !dir_dev || !out_dev) return fail(RT_ERR_INVALID, "rt_beam_voxels_dense: NULL argument");
    if (m < 0) return fail(RT_ERR_INVALID, "rt_beam_voxels_dense: m < 0");
    if (m == 0) return RT_OK;
    const Grid G = make_grid(grid);
    RT_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)m * G.nvox * sizeof(float), (cudaStream_t)stream));
    rt_beam_dense_kernel<<<warps_grid(m, 256), 256, 0, (cudaStream_t)stream>>>(G, pos_dev, dir_dev, m, out_dev,
                                                                              status_dev);
    RT_LAUNCH_CHECK("rt_beam_dense_kernel");
    return RT_OK;
}

int rt_pose_update(const int32_t grid[3], const double *pos_dev, const double *dir_dev, const float *actions_dev,
                   int m, double *pos_out_dev, double *dir_out_dev, double *overshoot_t_dev,
                   double *overshoot_r_dev, void *stream)
{
    if (int rc = check_grid(grid)) return rc;
    if (m == 0) return RT_OK;
    if (!pos_dev || !dir_dev || !actions_dev || !pos_out_dev || !dir_out_dev)
        return fail(RT_ERR_INVALID, "rt_pose_update: NULL argument");
    if (m < 0) return fail(RT_ERR_INVALID, "rt_pose_update: m < 0");
    if (m == 0) return RT_OK;
    rt_pose_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_grid(grid), pos_dev, dir_dev, actions_dev, m,
                                                                     pos_out_dev, dir_out_dev, overshoot_t_dev,
                                                                     overshoot_r_dev);
    RT_LAUNCH_CHECK("rt_pose_kernel");
    return RT_OK;
}

int rt_apply_rotation(const double *dir_dev, const double *rotvec_dev, int m, double min_angle, double *dir_out_dev,
                      double *overshoot_dev, void *stream)
{
    if (m == 0) return RT_OK;
    if (!dir_dev || !rotvec_dev || !dir_out_dev) return fail(RT_ERR_INVALID, "rt_apply_rotation: NULL argument");
    if (m < 0) return fail(RT_ERR_INVALID, "rt_apply_rotation: m < 0");
    if (m == 0) return RT_OK;
    // transforms.py:36-37 evaluate np.cos(min_angle) and np.sqrt(1 - cos**2) on the host (glibc);
    // so does this entry point, so the clamp target is the reference's to the last bit.
    const double cos_min = cos(min_angle);
    const double xy_mag = sqrt(1.0 - cos_min * cos_min);
    rt_rotation_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dir_dev, rotvec_dev, m, min_angle, cos_min,
                                                                         xy_mag, dir_out_dev, overshoot_dev);
    RT_LAUNCH_CHECK("rt_rotation_kernel");
    return RT_OK;
}

int rt_apply_translation(const double *pos_dev, const double *translation_dev, int m, const double bounds[3],
                         double *pos_out_dev, double *overshoot_dev, void *stream)
{
    if (m == 0) return RT_OK;
    if (!pos_dev || !translation_dev || !bounds || !pos_out_dev)
        return fail(RT_ERR_INVALID, "rt_apply_translation: NULL argument");
    if (m < 0) return fail(RT_ERR_INVALID, "rt_apply_translation: m < 0");
    if (m == 0) return RT_OK;
    rt_translation_kernel<<<(3 * m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        pos_dev, translation_dev, m, bounds[0], bounds[1], bounds[2], pos_out_dev, overshoot_dev);
    RT_LAUNCH_CHECK("rt_translation_kernel");
    return RT_OK;
}

int rt_gae(const float *rewards_dev, const float *values_dev, const float *dones_dev, const float *next_value_dev,
           const float *next_done_dev, int T, int N, double gamma, double gae_lambda, float *advantages_dev,
           float *returns_dev, void *stream)
{
    if (T < 0 || N < 0) return fail(RT_ERR_INVALID, "rt_gae: negative extent");
    if (T == 0 || N == 0) return RT_OK;
    if (!rewards_dev || !values_dev || !dones_dev || !next_value_dev || !next_done_dev || !advantages_dev || !returns_dev)
        return fail(RT_ERR_INVALID, "rt_gae: NULL argument");
    rt_gae_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards_dev, values_dev, dones_dev, next_value_dev,
                                                                    next_done_dev, T, N, (float)gamma,
                                                                    (float)(gamma * gae_lambda), advantages_dev,
                                                                    returns_dev);
    RT_LAUNCH_CHECK("rt_gae_kernel");
    return RT_OK;
}

}  // extern "C"

// FeaturesExtractor3D kernels (tcgen05 conv blocks, tail) and their C ABI
#include "rt_conv.cuh"
#include "rt_policy.cuh"
#include "rt_rollout.cuh"
