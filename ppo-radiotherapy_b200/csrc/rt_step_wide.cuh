// rt_step_wide.cuh — the sparse environment step with one THREAD per env, for env counts large enough to fill
// the GPU with independent threads (tens of thousands of envs per GPU), included by rt_env.cu.
//
// rt_step3_kernel gives an env a warp for the dose deposition: the right choice when there are 28 envs per SM
// and the step is a dependent chain, but SIMT-inefficient — a 35-slab beam uses 35 of 64 lane-slots and every
// scalar decision costs a warp instruction.  Here a warp instruction serves 32 envs at every stage: pose
// update, beam set-up, the slab walk (which is sequential anyway: intery += gradient*step, draw_line.py:98-99),
// the 2x2 splat with the structural duplicate merge, the sparse dose read-modify-write and the rewards, all in
// one thread.  Nothing is staged in shared memory: the env's sector-valid bitmap is private to its thread, so
// bitmap words are read and updated with plain loads and stores (no atomics), the lungs and tumour bitmasks come
// through L1.  Every access of a warp goes to 32 different dose volumes, exactly as scattered as in the
// warp-per-env kernel; what changes is the instruction count per env-step (about 2.5x lower) and that latency is
// hidden by thread-level parallelism across envs instead of within one env.
//
// STATUS: experimental, selected with RT_STEP_KB=-1 only.  It passes the whole parity suite, but measured on B200
// it is slower than rt_step3_kernel at every env count tried (65,536 envs: 308 us against 206; 131,072: 613
// against 413; one thread needs about 210 us for its step).  The slab loop is a long dependent chain per thread
// (three to four memory round trips and about 400 instructions per slab, 35-70 slabs), the lanes of a warp
// diverge on the dominant axis and walk direction of their beams, and 96 registers cap the SM at 20 warps.  It
// would need the next slab's bitmap words and dose values prefetched one iteration ahead, and envs bucketed by
// beam variant, to pay off.
//
// Same arithmetic, same results: the float32 ray walk and splat use the functions of rt_device.cuh, a voxel's two
// contributions are merged before the non-linear update, the dose deltas are accumulated in float64.
#pragma once

namespace {

constexpr int kWideThreads = 128;

__global__ void __launch_bounds__(kWideThreads) rt_step_wide_kernel(Tables T, Schedule S, EnvRec *rec, float *dose,
                                                                    uint32_t *valid, double *beams, int n_envs,
                                                                    const float *__restrict__ actions, StepOut out)
{
    const Grid &G = T.G;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    EnvRec *my = rec + e;
    const double gs[3] = {(double)G.g0, (double)G.g1, (double)G.g2};
    uint32_t *vbits = valid + (size_t)e * G.vwords;

    if (my->needs_reset) {
        // gymnasium 1.0.0 NEXT_STEP: the call after a terminal step resets and reports reward 0 (environment.py:77-105)
        const int episode = my->episode + 1;
        const int tid = pick_tumour(T, S, e, n_envs, episode);
        uint4 *vw = reinterpret_cast<uint4 *>(vbits);
        for (int i = 0; i < G.vwords / 4; i++) vw[i] = make_uint4(0u, 0u, 0u, 0u);
        float *obs = out.obs + (size_t)e * RT_OBS_SIZE;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double p = gs[i] / 2.0, d = i == 1 ? 1.0 : 0.0;
            my->pos[i] = p;
            my->dir[i] = d;
            obs[i] = (float)__dsub_rn(__dmul_rn(__ddiv_rn(p, gs[i]), 2.0), 1.0);
            obs[3 + i] = (float)d;
            obs[6 + i] = __ldg(&T.tumours[tid].obs_c[i]);
        }
        my->tumour_dose = 0.0; my->lung_dose = 0.0; my->ep_return = 0.0;
        my->t = 0; my->tumour_id = tid; my->lung_count = 0; my->episode = episode; my->needs_reset = 0; my->n_beams = 0;
        if (out.reward) out.reward[e] = 0.0;
        if (out.reward_f32) out.reward_f32[e] = 0.0f;
        if (out.terminated) out.terminated[e] = 0;
        if (out.truncated) out.truncated[e] = 0;
        if (out.info) {
            double *ip = out.info + (size_t)e * RT_INFO_SIZE;
#pragma unroll
            for (int i = 0; i < RT_INFO_SIZE; i++) ip[i] = i == RT_INFO_TUMOUR_ID ? (double)tid : 0.0;
        }
        return;
    }

    // ---- pose (environment.py:112-143, transforms.py:7-69)
    Pose s;
    float a[6];
    {
        const float2 *ap = reinterpret_cast<const float2 *>(actions + (size_t)e * RT_ACTION_SIZE);
        const float2 a01 = __ldg(ap), a23 = __ldg(ap + 1), a45 = __ldg(ap + 2);
        a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y; a[4] = a45.x; a[5] = a45.y;
    }
    const int tid = my->tumour_id;
    double os_t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s.d[i] = my->dir[i];
        s.p[i] = translate_axis(my->pos[i], __dmul_rn(__dmul_rn((double)clip1(a[i]), gs[i]), 0.2), gs[i], os_t[i]);
    }
    double rv[3];
    map_rotation(a, rv);
    const double zc = rotate_env(s.d, rv);
    const Beam b = beam_setup(G, s.p, s.d);                                 // draw_line.py:19-66

    // ---- distance_to_tumour_reward (environment.py:150-162)
    const Tumour *tg = T.tumours + tid;
    const int nv = __ldg(&tg->n_vox), voff = __ldg(&tg->vox_off);
    double best = CUDART_INF;
    for (int k = 0; k < nv; k++) {
        const uint32_t pk = __ldg(T.vox_xyz + voff + k);
        const double dx = (double)(pk & 255u) - s.p[0];
        const double dy = (double)((pk >> 8) & 255u) - s.p[1];
        const double dz = (double)(pk >> 16) - s.p[2];
        const double d2 = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
        best = d2 < best ? d2 : best;
    }

    // ---- dose deposition (environment.py:107-110) along the slab walk (draw_line.py:68-100)
    float *vol = dose + (size_t)e * G.vstride;
    const int li0 = __ldg(&tg->lo[0]), li1 = __ldg(&tg->lo[1]) - 1, li2 = __ldg(&tg->lo[2]) - 1;   // padded bbox origin
    const int td0 = __ldg(&tg->dim[0]), td1 = __ldg(&tg->dim[1]), td2 = __ldg(&tg->dim[2]);
    const int pd1 = td1 + 2, pd2 = td2 + 2;
    const uint32_t *tb = T.tumour_pbits + (size_t)tid * T.pbits_words;
    double d_tum = 0.0, d_lung = 0.0;
    int d_cnt = 0;
    float2 prv = make_float2(b.y0, b.z0), cur = prv;
    float2 nxt = make_float2(__fadd_rn(cur.x, b.sgy), __fadd_rn(cur.y, b.sgz));
    for (int k = 0; k < b.nslab; k++) {
        int lin[4], c0, c1, c2;
        float w[4];
        slab_targets_yz(G, b, k, cur, prv, nxt, lin, w, c0, c1, c2);
        // freshness of the (at most four) sectors, then all dose loads of the slab together
        float old[4];
        uint32_t freshm = 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            old[j] = 0.0f;
            if (lin[j] >= 0) {
                const int sec = lin[j] >> 3;
                const bool fresh = !((vbits[sec >> 5] >> (sec & 31)) & 1u);    // never written this episode: reads as zero
                if (fresh) freshm |= 1u << j;
                else old[j] = vol[lin[j]];
            }
        }
        // tumour membership of the 2x2 block (padded bounding-box bitmask, see rt_step3_kernel)
        uint32_t tmask = 0u;
        {
            const int ti = c0 - li0, tj = c1 - li1, tk = c2 - li2;
            if ((unsigned)ti < (unsigned)td0 && (unsigned)tj <= (unsigned)td1 && (unsigned)tk <= (unsigned)td2) {
                const int b0 = (ti * pd1 + tj) * pd2 + tk, b1 = b0 + pd2;
                const uint32_t r0 = __funnelshift_r(__ldg(tb + (b0 >> 5)), __ldg(tb + (b0 >> 5) + 1), b0 & 31) & 3u;
                const uint32_t r1 = __funnelshift_r(__ldg(tb + (b1 >> 5)), __ldg(tb + (b1 >> 5) + 1), b1 & 31) & 3u;
                tmask = r0 | (r1 << 2);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (lin[j] >= 0) {
                const int l = lin[j], sec = l >> 3;
                if ((freshm >> j) & 1u) {
                    // first write to the sector this episode (it may have become valid through an earlier target of
                    // this slab): materialise it as zeros, mark it valid.  The bitmap is private to this thread.
                    const uint32_t wv = vbits[sec >> 5];
                    if (!((wv >> (sec & 31)) & 1u)) {
                        zero_sector(vol + (sec << 3));
                        vbits[sec >> 5] = wv | (1u << (sec & 31));
                    }
                }
                const float o = old[j];
                const float nd = fminf(__fadd_rn(o, __fmul_rn(w[j], 0.100000001490116119f)), 1.0f);   // clip(dose + beam*0.1, 0, 1)
                vol[l] = nd;
                const bool in_t = (tmask >> j) & 1u;
                const bool in_l = (__ldg(T.lungs_bits + (l >> 5)) >> (l & 31)) & 1u;
                if (in_t || in_l) {
                    const double delta = (double)nd - (double)o;
                    if (in_t) d_tum += delta;
                    if (in_l) d_lung += delta;
                    if (!in_t && !(o > 0.200000002980232239f) && nd > 0.200000002980232239f) d_cnt++;
                }
            }
        prv = cur;
        cur = nxt;
        nxt = make_float2(__fadd_rn(nxt.x, b.sgy), __fadd_rn(nxt.y, b.sgz));
    }

    // ---- rewards, termination (environment.py:158-191, 214-220), record, outputs
    const double tumour_dose = my->tumour_dose + d_tum;
    const double lung_dose = my->lung_dose + d_lung;
    const int lung_count = my->lung_count + d_cnt;
    const int t = my->t + 1, n_beams = my->n_beams;
    const float tsum_f32 = (float)tumour_dose;
    const float ratio = fdiv_rn_zero_num(tsum_f32, __ldg(&tg->tumour_sum));
    const float r_tumour = __fmul_rn(ratio, 10.0f);
    const double r_lung = __dmul_rn(__ddiv_rn((double)lung_count, (double)__ldg(&tg->lung_mask_sum)), -1.0);
    const double r_dist = __dmul_rn(__ddiv_rn(sqrt(best), T.gnorm), -1.0);
    const double reward = __dadd_rn(__dadd_rn((double)r_tumour, r_lung), r_dist);
    const bool done = (ratio >= 0.899999976158142090f) || (t >= RT_MAX_TIME_STEPS);
    const double ep_return = my->ep_return + reward;
    float *obs = out.obs + (size_t)e * RT_OBS_SIZE;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        my->pos[i] = s.p[i];
        my->dir[i] = s.d[i];
        obs[i] = (float)__dsub_rn(__dmul_rn(__ddiv_rn(s.p[i], gs[i]), 2.0), 1.0);
        obs[3 + i] = (float)s.d[i];
        obs[6 + i] = __ldg(&tg->obs_c[i]);
    }
    my->tumour_dose = tumour_dose; my->lung_dose = lung_dose; my->ep_return = ep_return;
    my->t = t; my->lung_count = lung_count; my->needs_reset = done ? 1 : 0; my->n_beams = n_beams + 1;
    if (beams && n_beams < RT_MAX_TIME_STEPS) {                            // environment.py:110
        double *bp = beams + ((size_t)e * RT_MAX_TIME_STEPS + n_beams) * 6;
#pragma unroll
        for (int i = 0; i < 3; i++) { bp[i] = s.p[i]; bp[3 + i] = s.d[i]; }
    }
    if (out.reward) out.reward[e] = reward;
    if (out.reward_f32) out.reward_f32[e] = (float)reward;
    if (out.terminated) out.terminated[e] = done ? 1 : 0;
    if (out.truncated) out.truncated[e] = 0;
    if (out.info) {
        double *ip = out.info + (size_t)e * RT_INFO_SIZE;
        ip[RT_INFO_REWARD_TOTAL] = reward;
        ip[RT_INFO_REWARD_TUMOUR] = (double)r_tumour;
        ip[RT_INFO_REWARD_LUNG] = r_lung;
        ip[RT_INFO_REWARD_DISTANCE] = r_dist;
        ip[RT_INFO_DOSE_TUMOUR] = (double)tsum_f32;
        ip[RT_INFO_DOSE_LUNG] = (double)(float)lung_dose;
        ip[RT_INFO_OVERSHOOT_T0] = os_t[0];
        ip[RT_INFO_OVERSHOOT_T0 + 1] = os_t[1];
        ip[RT_INFO_OVERSHOOT_T0 + 2] = os_t[2];
        ip[RT_INFO_OVERSHOOT_R] = overshoot_from_z(zc);
        ip[RT_INFO_EPISODE_RETURN] = ep_return;
        ip[RT_INFO_EPISODE_LENGTH] = (double)t;
        ip[RT_INFO_LUNG_COUNT] = (double)lung_count;
        ip[RT_INFO_STEPPED] = 1.0;
        ip[RT_INFO_TUMOUR_ID] = (double)tid;
        ip[RT_INFO_T] = (double)t;
    }
}

}  // namespace
