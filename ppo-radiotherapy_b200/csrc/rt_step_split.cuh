// rt_step_split.cuh — the sparse environment step (RadiotherapyEnv.step, environment.py:193-243) as two
// kernels, for env counts of several waves per GPU (selected in rt_create; RT_STEP_KB=-2 forces it), included by
// rt_env.cu after rt_step.cuh.
//
// rt_step3_kernel couples a scalar warp (thread per env: pose, beam set-up, walk, rewards) and kB env warps (warp
// per env: deposition) with three block barriers.  With one wave of blocks that is the shortest dependent chain;
// with many waves it wastes the SM: the env warps of a block idle during its scalar phase, the scalar warp and
// the finished env warps wait for the slowest beam of the block (a third of all stall samples are barriers), and
// 2 x 15 resident warps cannot hide the rest.  Here the two kinds of work are separate launches:
//
//   rt_split_pose_kernel    one THREAD per env, 64 per block: record + action load, float64 translation and
//        rotation (transforms.py:7-69), beam clip / set-up and the serial float32 slab walk (draw_line.py:19-66,
//        98-99), observation (environment.py:259-268), the NEXT_STEP autoreset of the record.  Leaves a 128-byte
//        BeamWork record and the (intery, interz) of every slab in global memory (written row-contiguously from a
//        shared-memory stage; they stay in L2).
//   rt_split_deposit_kernel one WARP per env, kW independent warps per block, no block barrier: sector-valid
//        bitmap by cp.async.bulk, tumour entry, distance-to-tumour minimum (environment.py:150-162), the dose
//        deposition exactly as in rt_step3_kernel (environment.py:107-110: all passes of a beam through one HBM
//        round trip), warp reductions.  The warp that finishes LAST in its block (shared-memory ticket) computes
//        rewards, termination, episode statistics and the record update (environment.py:158-191, 214-243) for all
//        kW envs of the block with one thread per env and writes the outputs row-contiguously; nobody waits.
//
// Both are launched with programmatic stream serialization; the pose kernel triggers only after its own
// cudaGridDependencySynchronize, so nothing of call t+1 starts before the deposit kernel of call t has completed.
// Same arithmetic as rt_step3_kernel, same results (tests/test_gpu_parity.py runs every variant against the oracle).
#pragma once

namespace {

// Hand-over record of one env, one 128-byte line: what the deposit kernel needs of the beam (draw_line.py:50-60), of
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
    int pad_[2];
};
static_assert(sizeof(BeamWork) == 128, "BeamWork must be one 128-byte line");

constexpr int kPoseThreads = 64;
constexpr int kSplitWarpsPerSM = 32;    // resident warps the deposit kernel is compiled for (64 registers)

__global__ void __launch_bounds__(kPoseThreads)
rt_split_pose_kernel(Tables T, Schedule S, EnvRec *rec, double *beams, int n_envs, const float *__restrict__ actions,
                     BeamWork *work, float2 *yzg, int yz_stride, float *obs_out, int want_info)
{
    extern __shared__ __align__(16) uint32_t dyn_smem[];
    const int ys = yz_stride + 2;                                           // rows 16-byte aligned for the bulk store, 2-way bank conflicts at most
    float2 *s_yz = reinterpret_cast<float2 *>(dyn_smem);                    // [kPoseThreads][ys]
    float *s_obs = reinterpret_cast<float *>(s_yz + (size_t)kPoseThreads * ys);   // [kPoseThreads][9]
    const Grid &G = T.G;
    const int lane = threadIdx.x & (kWarp - 1);
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
        w.pad_[0] = w.pad_[1] = 0;
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
        uint4 *dst = reinterpret_cast<uint4 *>(work + e);
        const uint4 *src = reinterpret_cast<const uint4 *>(&w);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(BeamWork) / 16); i++) dst[i] = src[i];
    }
    // draw_line.py:98-99: the serial walk, every lane of the warp for as many slabs as the longest beam of the warp
    // (stores past a lane's own beam stay inside its row and are never read)
    const int nmax = __reduce_max_sync(kFull, b.nslab);
    {
        float2 *row = s_yz + (size_t)threadIdx.x * ys;
        float y = b.y0, z = b.z0;
#pragma unroll 4
        for (int k = 0; k < nmax; k++) {
            row[k] = make_float2(y, z);
            y = __fadd_rn(y, b.sgy);
            z = __fadd_rn(z, b.sgz);
        }
    }
    __syncthreads();
    const int nb = min(kPoseThreads, n_envs - env0);
    for (int i = threadIdx.x; i < nb * RT_OBS_SIZE; i += kPoseThreads) obs_out[(size_t)env0 * RT_OBS_SIZE + i] = s_obs[i];
    // every thread hands its own row to the copy engine: one bulk store (TMA, no tensor map) of ceil(n / 2) * 16 bytes
    if (mine && b.nslab > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // my generic-proxy stores -> async proxy
        const uint32_t bytes = (uint32_t)((b.nslab + 1) / 2) * 16u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(yzg + (size_t)e * yz_stride), "r"(smem_u32(s_yz + (size_t)threadIdx.x * ys)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // shared memory stays allocated until the copy is complete
    }
}

struct __align__(8) SplitResult {
    double d_tum, d_lung;    // dose deltas of this beam
    double best;             // min squared distance to the tumour
    int d_cnt;
    int pad_;
};

template <int kW>
__global__ void __launch_bounds__(kW * kWarp, kSplitWarpsPerSM / kW)
rt_split_deposit_kernel(Tables T, EnvRec *rec, float *dose, uint32_t *valid, int n_envs, const BeamWork *work,
                        const float2 *yzg, int yz_stride, StepOut out)
{
    __shared__ __align__(16) BeamWork wk[kW];
    __shared__ uint32_t tbits[kW][kMaxPTumourWords];
    __shared__ SplitResult res[kW];
    __shared__ __align__(8) unsigned long long mbars[kW + 1];
    __shared__ int s_done;
    __shared__ double s_rew[kW];
    __shared__ uint8_t s_term[kW];
    __shared__ double s_info[kW * RT_INFO_SIZE];
    extern __shared__ __align__(128) uint32_t dyn_smem[];                  // [lung_words16] lungs bitmask, [kW][vwords] sector-valid bitmaps
    const Grid &G = T.G;
    const int le = threadIdx.x / kWarp;
    const int lane = threadIdx.x & (kWarp - 1);
    const int env0 = blockIdx.x * kW;
    const int env = env0 + le;
    const bool active = env < n_envs;
    const int nb = min(kW, n_envs - env0);
    const uint32_t *slungs = dyn_smem;
    uint32_t *vsm = dyn_smem + T.lung_words16 + (size_t)le * G.vwords;
    const uint32_t mbar = smem_u32(&mbars[le]), mbar_l = smem_u32(&mbars[kW]);
    if (lane == 0) mbar_init(mbar, 1);
    if (threadIdx.x == 0) { mbar_init(mbar_l, 1); s_done = 0; }
    // The bitmap was last written by the deposit kernel of the previous call, which had completed before the pose
    // kernel of this call let us start, and the lungs bitmask is constant: both can be fetched before the grid
    // dependency resolves.
    if (threadIdx.x == 0) bulk_load(smem_u32(dyn_smem), T.lungs_bits, (uint32_t)(T.lung_words16 * sizeof(uint32_t)), mbar_l);
    if (active && lane == 0)
        bulk_load(smem_u32(vsm), valid + (size_t)env * G.vwords, (uint32_t)(G.vwords * sizeof(uint32_t)), mbar);
    __syncthreads();                                                       // s_done, the lungs mbarrier (start-up only)
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (!active) {
        return;
    }

    // one round trip: the hand-over record and the walk values of this lane's slabs (all passes, speculatively)
    const uint32_t wword = reinterpret_cast<const uint32_t *>(work + env)[lane];
    const float2 *myz = yzg + (size_t)env * yz_stride;
    float2 cur[kMaxPass];
#pragma unroll
    for (int c = 0; c < kMaxPass; c++) cur[c] = myz[min(c * kWarp + lane, yz_stride - 1)];
    reinterpret_cast<uint32_t *>(&wk[le])[lane] = wword;
    __syncwarp();
    const BeamWork &w = wk[le];
    const bool stepping = w.stepping != 0;
    uint32_t *tb = tbits[le];
    if (stepping) {
        const int tid = w.tid;
        const int nv = w.n_vox;
        const uint32_t *vx = T.vox_xyz + w.vox_off;
        // second round trip: padded tumour bitmask and the tumour's voxel list
        for (int i = lane; i < T.pbits_words; i += kWarp)                  // pbits_words <= kMaxPTumourWords (rt_create)
            tb[i] = __ldg(T.tumour_pbits + (size_t)tid * T.pbits_words + i);
        // distance_to_tumour_reward (environment.py:150-162): min over the tumour's voxel list
        {
            const double p0 = w.p[0], p1 = w.p[1], p2 = w.p[2];
            double best = CUDART_INF;
            for (int k = lane; k < nv; k += kWarp) {
                const uint32_t pk = __ldg(vx + k);
                const double dx = (double)(pk & 255u) - p0;
                const double dy = (double)((pk >> 8) & 255u) - p1;
                const double dz = (double)(pk >> 16) - p2;
                const double d2 = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
                best = d2 < best ? d2 : best;
            }
            // min of non-negative doubles = min of their bit patterns: two integer warp reductions
            const uint32_t hi = (uint32_t)__double2hiint(best);
            const uint32_t mhi = __reduce_min_sync(kFull, hi);
            const uint32_t mlo = __reduce_min_sync(kFull, hi == mhi ? (uint32_t)__double2loint(best) : 0xffffffffu);
            if (lane == 0) res[le].best = __hiloint2double((int)mhi, (int)mlo);
        }
        __syncwarp();
        // ---- dose deposition (environment.py:107-110): targets + loads | zero fill | stores + accumulation, all
        // passes of the beam together (rt_step3_kernel has the commentary)
        Beam b;
        b.nslab = w.nslab; b.dom = w.dom; b.step = w.step; b.x0 = w.x0;
        const int nslab = b.nslab;
        float *vol = dose + (size_t)env * G.vstride;
        uint32_t *vbits = valid + (size_t)env * G.vwords;
        const int g2 = G.g2;
        const int li0 = w.lo[0], li1 = w.lo[1] - 1, li2 = w.lo[2] - 1;     // origin of the padded bbox
        const int td0 = w.dim[0], td1 = w.dim[1], td2 = w.dim[2];
        const int pd1 = td1 + 2, pd2 = td2 + 2;
        const int variant = b.dom == 0 ? 0 : (b.dom * 2 - 1 + (b.step > 0 ? 0 : 1));   // warp-uniform
        mbar_wait(mbar_l, 0);                                              // the lungs bitmask has landed
        mbar_wait(mbar, 0);                                                // the staged bitmap has landed
        PassState ps[kMaxPass];
#pragma unroll
        for (int c = 0; c < kMaxPass; c++) {
            if (c * kWarp >= nslab) break;                                 // warp-uniform
            PassState &q = ps[c];
            const int k = c * kWarp + lane;
            // the neighbours' walk values come from the neighbouring lanes (slab_weights uses prv only for k > 0 and
            // nxt only for k + 1 < nslab)
            float2 prv, nxt;
            prv.x = __shfl_up_sync(kFull, cur[c].x, 1); prv.y = __shfl_up_sync(kFull, cur[c].y, 1);
            nxt.x = __shfl_down_sync(kFull, cur[c].x, 1); nxt.y = __shfl_down_sync(kFull, cur[c].y, 1);
            if (c > 0) {
                const float tx = __shfl_sync(kFull, cur[c > 0 ? c - 1 : 0].x, kWarp - 1), ty = __shfl_sync(kFull, cur[c > 0 ? c - 1 : 0].y, kWarp - 1);
                if (lane == 0) { prv.x = tx; prv.y = ty; }
            }
            if (c + 1 < kMaxPass) {
                const float tx = __shfl_sync(kFull, cur[c + 1 < kMaxPass ? c + 1 : c].x, 0), ty = __shfl_sync(kFull, cur[c + 1 < kMaxPass ? c + 1 : c].y, 0);
                if (lane == kWarp - 1) { nxt.x = tx; nxt.y = ty; }
            }
            int c0, c1, c2;
            uint32_t inb;
            const SlabCoord sc = b.dom == 0 ? slab_coords<0>(G, b, k, cur[c], q.base, inb, c0, c1, c2)
                               : b.dom == 1 ? slab_coords<1>(G, b, k, cur[c], q.base, inb, c0, c1, c2)
                                            : slab_coords<2>(G, b, k, cur[c], q.base, inb, c0, c1, c2);
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
                if (inj && !fresh) q.old[j] = vol[l];                      // re-touched sector: read from HBM / L2
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
        }
        // first write to a sector this episode: materialise it as zeros and mark it valid for the next step
#pragma unroll
        for (int c = 0; c < kMaxPass; c++) {
            if (c * kWarp >= nslab) break;
            const PassState &q = ps[c];
            int wrd[2] = {-1, -1};
            uint32_t msk[2] = {0u, 0u};
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (q.flags & (16u << j)) {
                    const int sec = (q.base + (j >> 1) * g2 + (j & 1)) >> 3;
                    zero_sector(vol + (sec << 3));
                    const int wd = sec >> 5;
                    const uint32_t bit = 1u << (sec & 31);
                    if (wrd[0] < 0 || wrd[0] == wd) { wrd[0] = wd; msk[0] |= bit; }
                    else if (wrd[1] < 0 || wrd[1] == wd) { wrd[1] = wd; msk[1] |= bit; }
                    else red_or(vbits + wd, bit);
                }
            if (wrd[0] >= 0) red_or(vbits + wrd[0], msk[0]);
            if (wrd[1] >= 0) red_or(vbits + wrd[1], msk[1]);
        }
        __syncwarp();   // zero fill (any lane) is ordered before the voxel stores below
        float d_tum = 0.0f, d_lung = 0.0f;
        int d_cnt = 0;
#pragma unroll
        for (int c = 0; c < kMaxPass; c++) {
            if (c * kWarp >= nslab) break;
            const PassState &q = ps[c];
            float nd[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                nd[j] = fminf(__fadd_rn(q.old[j], __fmul_rn(q.w[j], 0.100000001490116119f)), 1.0f);   // clip(dose + beam*0.1, 0, 1)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int l = q.base + r * g2;
                const uint32_t both = (q.flags >> (2 * r)) & 3u;
                if (both == 3u && !(l & 1)) {
                    *reinterpret_cast<float2 *>(vol + l) = make_float2(nd[2 * r], nd[2 * r + 1]);
                } else {
                    if (both & 1u) vol[l] = nd[2 * r];
                    if (both & 2u) vol[l + 1] = nd[2 * r + 1];
                }
            }
            const uint32_t tmask = (q.flags >> 8) & 15u, lmask = (q.flags >> 12) & 15u;
            const uint32_t cmask = lmask & ~tmask;                         // lungs_mask = lungs*(1-tumours) (environment.py:174)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float o = q.old[j];
                const float delta = (q.flags >> j) & 1u ? nd[j] - o : 0.0f;
                d_tum += (tmask >> j) & 1u ? delta : 0.0f;
                d_lung += (lmask >> j) & 1u ? delta : 0.0f;
                d_cnt += (int)((cmask >> j) & 1u) & (int)(!(o > 0.200000002980232239f) && nd[j] > 0.200000002980232239f);
            }
        }
        {
            const bool upper = lane >= 16;
            const float keep = upper ? d_lung : d_tum, give = upper ? d_tum : d_lung;
            double v = (double)keep + (double)__shfl_xor_sync(kFull, give, 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            d_cnt = __reduce_add_sync(kFull, d_cnt);
            if (lane == 0) { res[le].d_tum = v; res[le].d_cnt = d_cnt; }
            if (lane == 16) res[le].d_lung = v;
        }
    } else {
        // autoreset (environment.py:104-105): no sector of the new episode's dose volume is valid
        mbar_wait(mbar, 0);                                                // the bulk read of the old bitmap is over
        uint4 *vw = reinterpret_cast<uint4 *>(valid + (size_t)env * G.vwords);
        for (int i = lane; i < G.vwords / 4; i += kWarp) vw[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    // ---- the last warp of the block to get here finishes the step for all its envs, one thread per env
    __syncwarp();
    int ticket = 0;
    if (lane == 0) {
        __threadfence_block();
        ticket = atomicAdd(&s_done, 1);
    }
    ticket = __shfl_sync(kFull, ticket, 0);
    if (ticket != nb - 1) return;
    __threadfence_block();
    mbar_wait(mbar_l, 0);                                                  // the lungs copy must land before the block retires
    if (lane < nb) {
        const int e = env0 + lane;
        const BeamWork &we = wk[lane];
        EnvRec *my = rec + e;
        if (we.stepping) {
            const SplitResult &rs = res[lane];
            const double tumour_dose = my->tumour_dose + rs.d_tum;
            const double lung_dose = my->lung_dose + rs.d_lung;
            double ep_return = my->ep_return;
            const int t = my->t + 1;                                       // environment.py:194
            const int lung_count = my->lung_count + rs.d_cnt;
            const int n_beams = my->n_beams;
            const double r_dist = __dmul_rn(__ddiv_rn(sqrt(rs.best), T.gnorm), -1.0);   // environment.py:158-162
            // rewards, termination (environment.py:158-191, 214-220)
            const float tsum_f32 = (float)tumour_dose;                     // np.sum(dose*tumours) float32
            const float ratio = __fdiv_rn(tsum_f32, we.tumour_sum);
            const float r_tumour = __fmul_rn(ratio, 10.0f);
            const double r_lung = __dmul_rn(__ddiv_rn((double)lung_count, (double)we.lung_mask_sum), -1.0);
            const double reward = __dadd_rn(__dadd_rn((double)r_tumour, r_lung), r_dist);
            const bool done = (ratio >= 0.899999976158142090f) || (t >= RT_MAX_TIME_STEPS);
            ep_return += reward;
            my->tumour_dose = tumour_dose; my->lung_dose = lung_dose; my->ep_return = ep_return;
            my->t = t; my->lung_count = lung_count; my->needs_reset = done ? 1 : 0; my->n_beams = n_beams + 1;
            s_rew[lane] = reward;
            s_term[lane] = done ? 1 : 0;
            if (out.info) {
                double *ip = s_info + lane * RT_INFO_SIZE;
                ip[RT_INFO_REWARD_TOTAL] = reward;
                ip[RT_INFO_REWARD_TUMOUR] = (double)r_tumour;
                ip[RT_INFO_REWARD_LUNG] = r_lung;
                ip[RT_INFO_REWARD_DISTANCE] = r_dist;
                ip[RT_INFO_DOSE_TUMOUR] = (double)tsum_f32;
                ip[RT_INFO_DOSE_LUNG] = (double)(float)lung_dose;
                ip[RT_INFO_OVERSHOOT_T0] = we.os_t[0];
                ip[RT_INFO_OVERSHOOT_T0 + 1] = we.os_t[1];
                ip[RT_INFO_OVERSHOOT_T0 + 2] = we.os_t[2];
                ip[RT_INFO_OVERSHOOT_R] = we.os_r;
                ip[RT_INFO_EPISODE_RETURN] = ep_return;
                ip[RT_INFO_EPISODE_LENGTH] = (double)t;
                ip[RT_INFO_LUNG_COUNT] = (double)lung_count;
                ip[RT_INFO_STEPPED] = 1.0;
                ip[RT_INFO_TUMOUR_ID] = (double)we.tid;
                ip[RT_INFO_T] = (double)t;
            }
        } else {
            s_rew[lane] = 0.0;
            s_term[lane] = 0;
            if (out.info) {
                double *ip = s_info + lane * RT_INFO_SIZE;
#pragma unroll
                for (int i = 0; i < RT_INFO_SIZE; i++) ip[i] = i == RT_INFO_TUMOUR_ID ? (double)we.tid : 0.0;
            }
        }
    }
    __syncwarp();
    if (lane < nb) {
        if (out.reward) out.reward[env0 + lane] = s_rew[lane];
        if (out.reward_f32) out.reward_f32[env0 + lane] = (float)s_rew[lane];
        if (out.terminated) out.terminated[env0 + lane] = s_term[lane];
        if (out.truncated) out.truncated[env0 + lane] = 0;
    }
    if (out.info)
        for (int i = lane; i < nb * RT_INFO_SIZE; i += kWarp) out.info[(size_t)env0 * RT_INFO_SIZE + i] = s_info[i];
}

}  // namespace
