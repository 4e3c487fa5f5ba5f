// rt_step_split.cuh — the sparse environment step (RadiotherapyEnv.step, environment.py:193-243) as two
// kernels, an experiment for env counts of many waves per GPU (selected with RT_STEP_KB=-2 only: it ties with the
// fused kernel at 65,536 envs and loses below), included by rt_env.cu after rt_step.cuh.
//
// rt_step3_kernel couples a scalar warp (thread per env: pose, beam set-up, walk, rewards) and kB env warps (warp
// per env: deposition) with three block barriers.  With one wave of blocks that is the shortest dependent chain;
// with many waves it wastes the SM: the env warps of a block idle during its scalar phase, the scalar warp and
// the finished env warps wait for the slowest beam of the block (a third of all stall samples are barriers), and
// 2 x 15 resident warps cannot hide the rest.  Here the two kinds of work are separate launches:
//
//   rt_split_pose_kernel    one THREAD per env, 64 per block: record + action load, float64 translation and
//        rotation (transforms.py:7-69), beam clip / set-up and the serial float32 slab walk (draw_line.py:19-66,
//        98-99), observation (environment.py:259-268), the NEXT_STEP autoreset of the record.  Leaves a 192-byte
//        BeamWork record, an 8-byte brief (sizes of the deposit kernel's bulk copies) and the (intery, interz) of
//        every slab in global memory (one 32-byte store per four slabs; they stay in L2).
//   rt_split_deposit_kernel one WARP per env at a time, PERSISTENT: a block owns a run of consecutive envs and its
//        warps take them one by one from a shared-memory counter; no block barrier.  While a warp works on an env
//        the copy engine (cp.async.bulk on mbarriers) brings the next env's hand-over record, walk values, padded
//        tumour bitmask and sector-valid bitmap to shared memory, so the only load a warp waits for is the dose of
//        re-touched sectors.  Per env: dose deposition with the arithmetic of rt_step3_kernel (environment.py:107-110),
//        one 32-slab pass at a time (the rolled loop has to fit the instruction cache: the SM's warps are all at
//        different points of it), distance-to-tumour minimum (environment.py:150-162) in the shadow of the dose
//        loads, warp reductions.  Rewards, termination, episode statistics and the record update
//        (environment.py:158-191, 214-243) are done one THREAD per env for up to 16 finished envs of the warp at once.
//
// Both are launched with programmatic stream serialization; the pose kernel triggers only after its own
// cudaGridDependencySynchronize, so nothing of call t+1 starts before the deposit kernel of call t has completed.
// Same arithmetic as rt_step3_kernel, same results (tests/test_gpu_parity.py runs every variant on the same episodes).
#pragma once

namespace {

// Hand-over record of one env, 192 bytes: what the deposit kernel needs of the beam (draw_line.py:50-60), of
// the pose and of the tumour table entry, so that its warp has everything after ONE load.
struct __align__(16) BeamWork {
    int nslab, dom, step, x0;   // Beam: slabs to visit, dominant axis, walk direction, first slab coordinate
    int stepping;               // 1 = this call steps the env, 0 = it resets it (NEXT_STEP autoreset)
    int tid;                    // tumour of the running episode (after a reset: the new one)
    double p[3];                // translated beam position
    double os_t[3];             // translation overshoot (info only)
    double os_r;                // rotation overshoot (info only)
    int lo[3], dim[3];          // Tumour: bbox origin and extent
    int n_vox, vox_off;
    float tumour_sum, lung_mask_sum;
    int t, lung_count;          // EnvRec before this step (zero after a reset): what the tail of the deposit kernel updates
    double tumour_dose, lung_dose, ep_return;
    int n_beams;
    int pad_[7];
};
static_assert(sizeof(BeamWork) == 192, "BeamWork layout");

constexpr int kPoseThreads = 64;
constexpr int kVoxStage = 128;           // entries of a tumour's voxel list the deposit kernel stages in shared memory
constexpr int kSplitWarpsPerSM = 28;    // resident warps the deposit kernel is compiled for (one block of 28 or two of 14: 72 registers)

__global__ void __launch_bounds__(kPoseThreads)
rt_split_pose_kernel(Tables T, Schedule S, EnvRec *rec, double *beams, int n_envs, const float *__restrict__ actions,
                     BeamWork *work, uint2 *brief, float2 *yzg, int yz_stride, float *obs_out, int want_info)
{
    extern __shared__ __align__(16) uint32_t dyn_smem[];
    float *s_obs = reinterpret_cast<float *>(dyn_smem);                     // [kPoseThreads][9]
    const Grid &G = T.G;
    const int env0 = blockIdx.x * kPoseThreads;
    const int e = env0 + threadIdx.x;
    const bool mine = e < n_envs;
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();

    EnvRec *my = rec + (mine ? e : 0);
    const double gs[3] = {(double)G.g0, (double)G.g1, (double)G.g2};
    Beam b;
    b.nslab = 0; b.dom = 0; b.o0 = 1; b.o1 = 2; b.step = 1; b.x0 = 0;
    b.y0 = b.z0 = b.sgy = b.sgz = 0.0f;
    float *obs = s_obs + threadIdx.x * RT_OBS_SIZE;
    if (mine) {
        const int needs_reset = my->needs_reset;
        int tid = my->tumour_id;
        Pose s;
        double p0[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { p0[i] = my->pos[i]; s.d[i] = my->dir[i]; }
        const int n_beams = my->n_beams;
        const int episode = my->episode;
        const float2 *ap = reinterpret_cast<const float2 *>(actions + (size_t)e * RT_ACTION_SIZE);
        const float2 a01 = __ldg(ap), a23 = __ldg(ap + 1), a45 = __ldg(ap + 2);
        BeamWork w;
#pragma unroll
        for (int i = 0; i < 7; i++) w.pad_[i] = 0;
        w.t = my->t; w.lung_count = my->lung_count; w.n_beams = n_beams;
        w.tumour_dose = my->tumour_dose; w.lung_dose = my->lung_dose; w.ep_return = my->ep_return;
        if (needs_reset == 0) {
            const float at[3] = {a01.x, a01.y, a23.x};
            const float ar[3] = {a23.y, a45.x, a45.y};
#pragma unroll
            for (int i = 0; i < 3; i++)                                    // environment.py:122-125, transforms.py:65-67
                s.p[i] = translate_axis(p0[i], __dmul_rn(__dmul_rn((double)clip1(at[i]), gs[i]), 0.2), gs[i], w.os_t[i]);
            double rv[3];
#pragma unroll
            for (int i = 0; i < 3; i++)                                    // environment.py:139-141
                rv[i] = (double)__fmul_rn(__fmul_rn(clip1(ar[i]), 3.14159274101257324f), 0.5f);
            const double zc = rotate_env(s.d, rv);                         // transforms.py:7-55
            b = beam_setup(G, s.p, s.d);                                   // draw_line.py:19-66
            if (b.nslab < 0) b.nslab = 0;
            w.os_r = want_info ? overshoot_from_z(zc) : 0.0;               // transforms.py:29-33, 57 (info only)
            const Tumour *tg = T.tumours + tid;
#pragma unroll
            for (int i = 0; i < 3; i++) {                                  // environment.py:259-268
                obs[i] = (float)__dsub_rn(__dmul_rn(__ddiv_rn(s.p[i], gs[i]), 2.0), 1.0);
                obs[3 + i] = (float)s.d[i];
                obs[6 + i] = __ldg(&tg->obs_c[i]);
                my->pos[i] = s.p[i];
                my->dir[i] = s.d[i];
                w.p[i] = s.p[i];
            }
            if (beams && n_beams < RT_MAX_TIME_STEPS) {                    // environment.py:110
                double *bp = beams + ((size_t)e * RT_MAX_TIME_STEPS + n_beams) * 6;
#pragma unroll
                for (int i = 0; i < 3; i++) { bp[i] = s.p[i]; bp[3 + i] = s.d[i]; }
            }
            w.stepping = 1;
        } else {
            // gymnasium 1.0.0 NEXT_STEP: the call after a terminal step resets and reports reward 0
            // (environment.py:77-105; the deposit kernel's warp clears the sector-valid bitmap).
            tid = pick_tumour(T, S, e, n_envs, episode + 1);
            const Tumour *tg = T.tumours + tid;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double p = gs[i] / 2.0, d = i == 1 ? 1.0 : 0.0;
                my->pos[i] = p;
                my->dir[i] = d;
                obs[i] = (float)__dsub_rn(__dmul_rn(__ddiv_rn(p, gs[i]), 2.0), 1.0);
                obs[3 + i] = (float)d;
                obs[6 + i] = __ldg(&tg->obs_c[i]);
                w.p[i] = p;
                w.os_t[i] = 0.0;
            }
            w.os_r = 0.0;
            my->tumour_dose = 0.0; my->lung_dose = 0.0; my->ep_return = 0.0;
            my->t = 0; my->tumour_id = tid; my->lung_count = 0; my->episode = episode + 1; my->needs_reset = 0; my->n_beams = 0;
            w.t = 0; w.lung_count = 0; w.n_beams = 0;
            w.tumour_dose = 0.0; w.lung_dose = 0.0; w.ep_return = 0.0;
            w.stepping = 0;
        }
        w.nslab = b.nslab; w.dom = b.dom; w.step = b.step; w.x0 = b.x0;
        w.tid = tid;
        {
            const Tumour tg = T.tumours[tid];
#pragma unroll
            for (int i = 0; i < 3; i++) { w.lo[i] = tg.lo[i]; w.dim[i] = tg.dim[i]; }
            w.n_vox = tg.n_vox; w.vox_off = tg.vox_off;
            w.tumour_sum = tg.tumour_sum; w.lung_mask_sum = tg.lung_mask_sum;
        }
        // what the deposit kernel's warp must know before the record is in its shared memory (sizes and sources of the
        // bulk copies): x = nslab | stepping << 8 | tid << 16, y = vox_off / 4 | (16-byte chunks of the voxel list) << 24
        brief[e] = make_uint2((uint32_t)b.nslab | ((uint32_t)w.stepping << 8) | ((uint32_t)tid << 16),
                              (uint32_t)(w.vox_off >> 2) | ((uint32_t)((min(w.n_vox, kVoxStage) + 3) >> 2) << 24));
        uint4 *dst = reinterpret_cast<uint4 *>(work + e);
        const uint4 *src = reinterpret_cast<const uint4 *>(&w);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(BeamWork) / 16); i++) dst[i] = src[i];
    }
    // draw_line.py:98-99: the serial walk, straight to the env's row in global memory, four slabs = one 32-byte
    // sector per store (rows are 32-byte aligned and hold a multiple of four entries)
    if (mine) {
        float *row = reinterpret_cast<float *>(yzg + (size_t)e * yz_stride);
        float y = b.y0, z = b.z0;
        for (int k = 0; k < b.nslab; k += 4) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                v[2 * u] = y; v[2 * u + 1] = z;
                y = __fadd_rn(y, b.sgy);
                z = __fadd_rn(z, b.sgz);
            }
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(row + 2 * k), "f"(v[0]), "f"(v[1]),
                         "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        }
    }
    __syncthreads();
    const int nb = min(kPoseThreads, n_envs - env0);
    for (int i = threadIdx.x; i < nb * RT_OBS_SIZE; i += kPoseThreads) obs_out[(size_t)env0 * RT_OBS_SIZE + i] = s_obs[i];
}

struct __align__(8) SplitResult {
    double d_tum, d_lung;    // dose deltas of this beam
    double best;             // min squared distance to the tumour
    int d_cnt;
    int env;
};

constexpr int kTailCap = 16;   // envs a warp finishes (thread per env) in one go

// per-warp shared memory of the deposit kernel (all bulk-copy destinations 16-byte aligned)
struct __align__(16) DepositWarp {
    BeamWork wk[2];                          // hand-over records: this env, next env
    uint32_t tb[2][kMaxPTumourWords];        // padded tumour bitmasks: this env, next env
    uint32_t vox[2][kVoxStage];              // first entries of the tumours' voxel lists: this env, next env
    SplitResult res[kTailCap];               // results waiting for the tail
    unsigned long long mbar_a[2];            // everything staged for the env in slot 0 / 1 (the bitmap included)
};

__device__ __forceinline__ void mbar_expect(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// Persistent: a block owns `envs_per_block` consecutive envs, its kW warps take them one at a time from a
// shared-memory counter.  While a warp works on an env, the copy engine brings the next env's hand-over record,
// walk values, tumour bitmask (slot ^ 1) and - as soon as the target phase has read the current one - its
// sector-valid bitmap: a warp never waits for a first-level load, only for the dose values of re-touched sectors.
template <int kW>
__global__ void __launch_bounds__(kW * kWarp, kSplitWarpsPerSM / kW)
rt_split_deposit_kernel(Tables T, EnvRec *rec, float *dose, uint32_t *valid, int n_envs, int envs_per_block,
                        const BeamWork *work, const uint2 *brief, const float2 *yzg, int yz_stride, StepOut out)
{
    __shared__ int s_next;
    __shared__ __align__(8) unsigned long long mbar_lungs;
    extern __shared__ __align__(128) uint32_t dyn_smem[];   // [lung_words16] lungs | [kW] DepositWarp | [kW][vwords] bitmaps | [kW][2][yz_stride] walks
    const Grid &G = T.G;
    const int le = threadIdx.x / kWarp;
    const int lane = threadIdx.x & (kWarp - 1);
    const uint32_t *slungs = dyn_smem;
    DepositWarp &dw = reinterpret_cast<DepositWarp *>(dyn_smem + T.lung_words16)[le];
    uint32_t *vsm = dyn_smem + T.lung_words16 + kW * (sizeof(DepositWarp) / 4) + (size_t)le * G.vwords;
    float2 *yzs = reinterpret_cast<float2 *>(dyn_smem + T.lung_words16 + kW * (sizeof(DepositWarp) / 4) + (size_t)kW * G.vwords) +
                  (size_t)le * 2 * yz_stride;
    const uint32_t mb_a[2] = {smem_u32(&dw.mbar_a[0]), smem_u32(&dw.mbar_a[1])};
    const uint32_t mb_l = smem_u32(&mbar_lungs);
    if (lane == 0) { mbar_init(mb_a[0], 1); mbar_init(mb_a[1], 1); }
    if (threadIdx.x == 0) {
        mbar_init(mb_l, 1);
        s_next = 0;
        bulk_load(smem_u32(dyn_smem), T.lungs_bits, (uint32_t)(T.lung_words16 * sizeof(uint32_t)), mb_l);   // constant data
    }
    __syncthreads();                                                       // start-up only
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    const int blk0 = blockIdx.x * envs_per_block;
    const int blk_n = min(envs_per_block, n_envs - blk0);
    const uint32_t pbits_bytes = (uint32_t)(T.pbits_words * sizeof(uint32_t));
    const uint32_t bitmap_bytes = (uint32_t)(G.vwords * sizeof(uint32_t));

    const uint32_t next_addr = smem_u32(&s_next);
    auto grab = [&]() -> int {                                             // next env of the block, or -1
        int v = 0;
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(next_addr) : "memory");
        v = __shfl_sync(kFull, v, 0);
        return v < blk_n ? blk0 + v : -1;
    };
    // Everything of an env arrives on ONE mbarrier (slot = parity of the env's position in the warp's sequence): the
    // transaction count announced here covers the sector-valid bitmap as well, whose copy is issued later (issue_b),
    // when the previous env's target phase has released the single bitmap buffer.
    auto issue_a = [&](int env, uint2 br2, int slot) {                     // record + walk + tumour bits + voxel list -> slot
        if (lane == 0) {
            const uint32_t br = br2.x;
            const uint32_t yzb = (((br & 255u) + 1u) >> 1) * 16u;
            const bool st = (br >> 8) & 1u;
            const uint32_t vxb = st ? (br2.y >> 24) * 16u : 0u;
            mbar_expect(mb_a[slot], (uint32_t)sizeof(BeamWork) + yzb + (st ? pbits_bytes + bitmap_bytes : 0u) + vxb);
            bulk_copy(smem_u32(&dw.wk[slot]), work + env, (uint32_t)sizeof(BeamWork), mb_a[slot]);
            if (yzb) bulk_copy(smem_u32(yzs + (size_t)slot * yz_stride), yzg + (size_t)env * yz_stride, yzb, mb_a[slot]);
            if (st) bulk_copy(smem_u32(dw.tb[slot]), T.tumour_pbits + (size_t)(br >> 16) * T.pbits_words, pbits_bytes, mb_a[slot]);
            if (vxb) bulk_copy(smem_u32(dw.vox[slot]), T.vox_xyz + (size_t)(br2.y & 0xffffffu) * 4, vxb, mb_a[slot]);
            // bring the bitmap from HBM to L2 meanwhile
            if (st) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(valid + (size_t)env * G.vwords), "r"(bitmap_bytes) : "memory");
        }
    };
    auto issue_b = [&](int env, int slot) {                                // sector-valid bitmap, on the env's barrier
        if (lane == 0) bulk_copy(smem_u32(vsm), valid + (size_t)env * G.vwords, bitmap_bytes, mb_a[slot]);
    };
    // thread per env: rewards, termination, episode statistics, record update (environment.py:158-191, 214-243)
    auto tail = [&](int cnt) {
        __syncwarp();
        if (lane < cnt) {
            const SplitResult &rs = dw.res[lane];
            const int e = rs.env;
            const BeamWork *we = work + e;
            EnvRec *my = rec + e;
            const int tid = we->tid;
            double reward = 0.0;
            int done = 0;
            double *ip = out.info ? out.info + (size_t)e * RT_INFO_SIZE : nullptr;
            if (we->stepping) {
                const double tumour_dose = we->tumour_dose + rs.d_tum;
                const double lung_dose = we->lung_dose + rs.d_lung;
                double ep_return = we->ep_return;
                const int t = we->t + 1;                                   // environment.py:194
                const int lung_count = we->lung_count + rs.d_cnt;
                const int n_beams = we->n_beams;
                const double r_dist = __dmul_rn(__ddiv_rn(sqrt(rs.best), T.gnorm), -1.0);   // environment.py:158-162
                const float tsum_f32 = (float)tumour_dose;                 // np.sum(dose*tumours) float32
                const float ratio = fdiv_rn_zero_num(tsum_f32, we->tumour_sum);
                const float r_tumour = __fmul_rn(ratio, 10.0f);
                const double r_lung = __dmul_rn(__ddiv_rn((double)lung_count, (double)we->lung_mask_sum), -1.0);
                reward = __dadd_rn(__dadd_rn((double)r_tumour, r_lung), r_dist);
                done = (ratio >= 0.899999976158142090f) || (t >= RT_MAX_TIME_STEPS);
                ep_return += reward;
                my->tumour_dose = tumour_dose; my->lung_dose = lung_dose; my->ep_return = ep_return;
                my->t = t; my->lung_count = lung_count; my->needs_reset = done; my->n_beams = n_beams + 1;
                if (ip) {
                    ip[RT_INFO_REWARD_TOTAL] = reward;
                    ip[RT_INFO_REWARD_TUMOUR] = (double)r_tumour;
                    ip[RT_INFO_REWARD_LUNG] = r_lung;
                    ip[RT_INFO_REWARD_DISTANCE] = r_dist;
                    ip[RT_INFO_DOSE_TUMOUR] = (double)tsum_f32;
                    ip[RT_INFO_DOSE_LUNG] = (double)(float)lung_dose;
                    ip[RT_INFO_OVERSHOOT_T0] = we->os_t[0];
                    ip[RT_INFO_OVERSHOOT_T0 + 1] = we->os_t[1];
                    ip[RT_INFO_OVERSHOOT_T0 + 2] = we->os_t[2];
                    ip[RT_INFO_OVERSHOOT_R] = we->os_r;
                    ip[RT_INFO_EPISODE_RETURN] = ep_return;
                    ip[RT_INFO_EPISODE_LENGTH] = (double)t;
                    ip[RT_INFO_LUNG_COUNT] = (double)lung_count;
                    ip[RT_INFO_STEPPED] = 1.0;
                    ip[RT_INFO_TUMOUR_ID] = (double)tid;
                    ip[RT_INFO_T] = (double)t;
                }
            } else if (ip) {
#pragma unroll
                for (int i = 0; i < RT_INFO_SIZE; i++) ip[i] = i == RT_INFO_TUMOUR_ID ? (double)tid : 0.0;
            }
            if (out.reward) out.reward[e] = reward;
            if (out.reward_f32) out.reward_f32[e] = (float)reward;
            if (out.terminated) out.terminated[e] = (uint8_t)done;
            if (out.truncated) out.truncated[e] = 0;
        }
        __syncwarp();
    };

    int cur_env = blk_n > 0 ? grab() : -1;
    if (cur_env < 0) {
        if (le == 0) mbar_wait(mb_l, 0);                                   // the lungs copy must land before the block retires
        return;
    }
    uint2 cur_br = brief[cur_env];
    issue_a(cur_env, cur_br, 0);
    if ((cur_br.x >> 8) & 1u) issue_b(cur_env, 0);
    int nxt_env = grab();
    uint2 nxt_br = nxt_env >= 0 ? brief[nxt_env] : make_uint2(0u, 0u);
    mbar_wait(mb_l, 0);                                                    // the lungs bitmask has landed
    int cnt = 0;
    const int g2 = G.g2;
    const int dbg = T.debug;                                               // RT_STEP_DEBUG what-if bits (timing experiments: results are wrong by construction)
    for (int it = 0; cur_env >= 0; it++) {
        const int slot = it & 1;
        if (nxt_env >= 0) issue_a(nxt_env, nxt_br, slot ^ 1);              // that slot's env finished in the previous iteration
        const int nn_env = nxt_env >= 0 ? grab() : -1;
        uint2 nn_br = make_uint2(0u, 0u);                                  // consumed in the next iteration: issued NOW
        if (nn_env >= 0) asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(nn_br.x), "=r"(nn_br.y) : "l"(brief + nn_env));
        const bool nxt_steps = nxt_env >= 0 && ((nxt_br.x >> 8) & 1u);
        const int env = cur_env;
        mbar_wait(mb_a[slot], (uint32_t)(it >> 1) & 1u);
        const BeamWork &w = dw.wk[slot];
        if ((cur_br.x >> 8) & 1u) {
            const uint32_t *tb = dw.tb[slot];
            const float2 *myz = yzs + (size_t)slot * yz_stride;
            // ---- dose deposition (environment.py:107-110): targets + loads | zero fill | stores + accumulation, all
            // passes of the beam together (rt_step3_kernel has the commentary)
            Beam b;
            b.nslab = w.nslab; b.dom = w.dom; b.step = w.step; b.x0 = w.x0;
            const int nslab = b.nslab;
            float *vol = dose + (size_t)env * G.vstride;
            uint32_t *vbits = valid + (size_t)env * G.vwords;
            const int li0 = w.lo[0], li1 = w.lo[1] - 1, li2 = w.lo[2] - 1; // origin of the padded bbox
            const int td0 = w.dim[0], td1 = w.dim[1], td2 = w.dim[2];
            const int pd1 = td1 + 2, pd2 = td2 + 2;
            const int variant = b.dom == 0 ? 0 : (b.dom * 2 - 1 + (b.step > 0 ? 0 : 1));   // warp-uniform
            const int nv = w.n_vox;
            const uint32_t *vx = T.vox_xyz + w.vox_off;
            const uint32_t *vxs = dw.vox[slot];                            // its first kVoxStage entries, staged
            // One 32-slab pass at a time (not unrolled: with every warp of the SM at a different point of the loop
            // the code has to fit the instruction cache).  A later pass must see what an earlier one did to a sector
            // they share: fresh sectors are also marked in the staged bitmap, and a warp barrier orders the stores
            // of pass c before the loads of pass c + 1.
            float d_tum = 0.0f, d_lung = 0.0f;
            int d_cnt = 0;
            for (int c0s = 0; c0s < nslab; c0s += kWarp) {
                const bool last = c0s + kWarp >= nslab;
                PassState q;
                const int k = c0s + lane;
                const int kk = k < nslab ? k : 0;
                const float2 cur = myz[kk], prv = myz[kk > 0 ? kk - 1 : 0], nxt = myz[kk + 1 < nslab ? kk + 1 : kk];
                int c0, c1, c2;
                uint32_t inb;
                const SlabCoord sc = b.dom == 0 ? slab_coords<0>(G, b, k, cur, q.base, inb, c0, c1, c2)
                                   : b.dom == 1 ? slab_coords<1>(G, b, k, cur, q.base, inb, c0, c1, c2)
                                                : slab_coords<2>(G, b, k, cur, q.base, inb, c0, c1, c2);
                // the dose loads go out first (a voxel the previous slab owns is loaded for nothing: harmless)
                uint32_t freshm = 0u, lungm = 0u;
                int sec[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool inj = (inb >> j) & 1u;
                    const int l = inj ? q.base + (j >> 1) * g2 + (j & 1) : 0;
                    sec[j] = l >> 3;
                    lungm |= ((slungs[l >> 5] >> (l & 31)) & 1u) << j;
                    const bool fresh = inj && !((vsm[sec[j] >> 5] >> (sec[j] & 31)) & 1u);   // never written this episode: reads as zero
                    q.old[j] = 0.0f;
                    if (inj && !fresh && !(dbg & 4)) q.old[j] = vol[l];                  // re-touched sector: read from HBM / L2
                    freshm |= fresh ? 1u << j : 0u;
                }
                uint32_t drop;
                switch (variant) {
                case 0: slab_weights<0, 0>(b, k, sc, prv, nxt, drop, q.w); break;
                case 1: slab_weights<1, 0>(b, k, sc, prv, nxt, drop, q.w); break;
                case 2: slab_weights<1, 1>(b, k, sc, prv, nxt, drop, q.w); break;
                case 3: slab_weights<2, 0>(b, k, sc, prv, nxt, drop, q.w); break;
                default: slab_weights<2, 1>(b, k, sc, prv, nxt, drop, q.w); break;
                }
                const uint32_t ok = inb & ~drop;
                freshm &= ok;
                uint32_t tmask = 0u;
                const int ti = c0 - li0, tj = c1 - li1, tk = c2 - li2;
                if ((unsigned)ti < (unsigned)td0 && (unsigned)tj <= (unsigned)td1 && (unsigned)tk <= (unsigned)td2) {
                    const int b0 = (ti * pd1 + tj) * pd2 + tk, b1 = b0 + pd2;
                    const uint32_t r0 = __funnelshift_r(tb[b0 >> 5], tb[(b0 >> 5) + 1], b0 & 31) & 3u;
                    const uint32_t r1 = __funnelshift_r(tb[b1 >> 5], tb[(b1 >> 5) + 1], b1 & 31) & 3u;
                    tmask = r0 | (r1 << 2);
                }
                uint32_t fill = freshm;
                if ((freshm & 3u) == 3u && sec[0] == sec[1]) fill &= ~2u;
                if ((freshm & 12u) == 12u && sec[2] == sec[3]) fill &= ~8u;
                q.flags = ok | (fill << 4) | ((tmask & ok) << 8) | ((lungm & ok) << 12);
                __syncwarp();                                              // every lane has done its bitmap lookups
                if (last) {
                    // the copy engine may overwrite the staged bitmap with the next env's
                    if (nxt_steps) issue_b(nxt_env, slot ^ 1);
                }
                if (c0s == 0) {
                    // distance_to_tumour_reward (environment.py:150-162): min over the tumour's voxel list (the staged
                    // part first).  The nearest voxel is searched in float32 and its squared distance evaluated in
                    // float64: a float32 near-tie picks a voxel whose distance differs by < 2e-7 relative, inside the
                    // stated tolerance of the reward (rtol 1e-6).
                    const float q0 = (float)w.p[0], q1 = (float)w.p[1], q2 = (float)w.p[2];
                    float bestf = CUDART_INF_F;
                    uint32_t bestv = 0u;
                    for (int k0 = 0; k0 < nv && k0 < kVoxStage; k0 += kWarp) {
                        const int kq = k0 + lane;
                        const uint32_t pk = vxs[kq < nv ? kq : 0];         // a repeated voxel does not change the minimum
                        const float dx = (float)(pk & 255u) - q0, dy = (float)((pk >> 8) & 255u) - q1, dz = (float)(pk >> 16) - q2;
                        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (d2 < bestf) { bestf = d2; bestv = pk; }
                    }
                    for (int k0 = kVoxStage; k0 < nv; k0 += 4 * kWarp) {   // the rest of a large tumour: four loads in flight
                        uint32_t pk[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int kq = k0 + u * kWarp + lane;
                            pk[u] = __ldg(vx + (kq < nv ? kq : 0));
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float dx = (float)(pk[u] & 255u) - q0, dy = (float)((pk[u] >> 8) & 255u) - q1, dz = (float)(pk[u] >> 16) - q2;
                            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (d2 < bestf) { bestf = d2; bestv = pk[u]; }
                        }
                    }
                    const float mf = __reduce_min_sync(kFull, __float_as_uint(bestf)) == __float_as_uint(bestf) ? 0.0f : 1.0f;   // d2 >= 0: uint order = float order
                    const uint32_t win = __ballot_sync(kFull, mf == 0.0f);
                    const uint32_t pkw = __shfl_sync(kFull, bestv, __ffs(win) - 1);
                    if (lane == 0) {
                        const double dx = (double)(pkw & 255u) - w.p[0], dy = (double)((pkw >> 8) & 255u) - w.p[1], dz = (double)(pkw >> 16) - w.p[2];
                        dw.res[cnt].best = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
                    }
                }
                // first write to a sector this episode: materialise it as zeros and mark it valid for the next step
                // (and, when another pass follows, for that pass)
                {
                    int wrd[2] = {-1, -1};
                    uint32_t msk[2] = {0u, 0u};
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if ((q.flags & (16u << j)) && !(dbg & 1)) {
                            const int sc2 = (q.base + (j >> 1) * g2 + (j & 1)) >> 3;
                            if (!(dbg & 32)) zero_sector(vol + (sc2 << 3));
                            const int wd = sc2 >> 5;
                            const uint32_t bit = 1u << (sc2 & 31);
                            if (wrd[0] < 0 || wrd[0] == wd) { wrd[0] = wd; msk[0] |= bit; }
                            else if (wrd[1] < 0 || wrd[1] == wd) { wrd[1] = wd; msk[1] |= bit; }
                            else { if (!(dbg & 16)) red_or(vbits + wd, bit); if (!last) atomicOr(vsm + wd, bit); }
                        }
                    if (wrd[0] >= 0) { if (!(dbg & 16)) red_or(vbits + wrd[0], msk[0]); if (!last) atomicOr(vsm + wrd[0], msk[0]); }
                    if (wrd[1] >= 0) { if (!(dbg & 16)) red_or(vbits + wrd[1], msk[1]); if (!last) atomicOr(vsm + wrd[1], msk[1]); }
                }
                __syncwarp();   // zero fill (any lane) is ordered before the voxel stores below
                float nd[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    nd[j] = fminf(__fadd_rn(q.old[j], __fmul_rn(q.w[j], 0.100000001490116119f)), 1.0f);   // clip(dose + beam*0.1, 0, 1)
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int l = q.base + r * g2;
                    const uint32_t both = (q.flags >> (2 * r)) & 3u;
                    if (dbg & 2) continue;
                    if (both == 3u && !(l & 1)) {
                        *reinterpret_cast<float2 *>(vol + l) = make_float2(nd[2 * r], nd[2 * r + 1]);
                    } else {
                        if (both & 1u) vol[l] = nd[2 * r];
                        if (both & 2u) vol[l + 1] = nd[2 * r + 1];
                    }
                }
                const uint32_t lmask = (q.flags >> 12) & 15u;
                const uint32_t tmk = (q.flags >> 8) & 15u;
                const uint32_t cmask = lmask & ~tmk;                       // lungs_mask = lungs*(1-tumours) (environment.py:174)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float o = q.old[j];
                    const float delta = (q.flags >> j) & 1u ? nd[j] - o : 0.0f;
                    d_tum += (tmk >> j) & 1u ? delta : 0.0f;
                    d_lung += (lmask >> j) & 1u ? delta : 0.0f;
                    d_cnt += (int)((cmask >> j) & 1u) & (int)(!(o > 0.200000002980232239f) && nd[j] > 0.200000002980232239f);
                }
                if (!last) __syncwarp();                                   // this pass's stores before the next pass's loads
            }
            if (nslab <= 0) {
                // an empty beam (it missed the volume): only the distance term, and the bitmap buffer is free
                if (nxt_steps) issue_b(nxt_env, slot ^ 1);
                const double p0 = w.p[0], p1 = w.p[1], p2 = w.p[2];
                double best = CUDART_INF;
                for (int kq = lane; kq < nv; kq += kWarp) {
                    const uint32_t pk = __ldg(vx + kq);
                    const double dx = (double)(pk & 255u) - p0;
                    const double dy = (double)((pk >> 8) & 255u) - p1;
                    const double dz = (double)(pk >> 16) - p2;
                    const double d2 = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
                    best = d2 < best ? d2 : best;
                }
                best = warp_min(best);
                if (lane == 0) dw.res[cnt].best = best;
            }
            {
                const bool upper = lane >= 16;
                const float keep = upper ? d_lung : d_tum, give = upper ? d_tum : d_lung;
                double v = (double)keep + (double)__shfl_xor_sync(kFull, give, 16);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
                d_cnt = __reduce_add_sync(kFull, d_cnt);
                if (lane == 0) { dw.res[cnt].d_tum = v; dw.res[cnt].d_cnt = d_cnt; dw.res[cnt].env = env; }
                if (lane == 16) dw.res[cnt].d_lung = v;
            }
        } else {
            // autoreset (environment.py:104-105): no sector of the new episode's dose volume is valid
            uint4 *vw = reinterpret_cast<uint4 *>(valid + (size_t)env * G.vwords);
            for (int i = lane; i < G.vwords / 4; i += kWarp) vw[i] = make_uint4(0u, 0u, 0u, 0u);
            if (lane == 0) dw.res[cnt].env = env;
            if (nxt_steps) issue_b(nxt_env, slot ^ 1);                               // nobody is reading the staged bitmap
        }
        if (++cnt == kTailCap) { tail(cnt); cnt = 0; }
        __syncwarp();                                                      // slot's buffers are free for the copy engine
        cur_env = nxt_env; cur_br = nxt_br;
        nxt_env = nn_env; nxt_br = nn_br;
    }
    tail(cnt);
}

}  // namespace
