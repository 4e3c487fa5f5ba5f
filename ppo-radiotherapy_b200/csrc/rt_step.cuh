// rt_step.cuh — the environment step (RadiotherapyEnv.step, environment.py:193-243), included by rt_env.cu after
// the record / table definitions.
//
// Dose state of a sparse-mode handle: one 8-byte CELL per voxel, {float32 dose, uint32 generation}, C order of the
// reference's volume.  A cell whose generation differs from its env's current one (EnvRec::gen, bumped by every
// reset) reads as zero: reset touches nothing, a beam needs no validity bitmap, no zero fill and no atomics, and
// every voxel a beam hits is one 8-byte load and one 8-byte store by the one lane that owns it.
//
// A block advances kB envs with kB + 1 warps and two kinds of work:
//
//   scalar warp (warp 0), one THREAD per env: everything that is one value per env — record and action load,
//      float64 translation and rotation (transforms.py:7-69), beam clip/set-up and the serial float32 slab walk
//      (draw_line.py:19-66, 98-99), rewards, termination, observation, episode statistics, the record update and
//      the NEXT_STEP autoreset.  One warp instruction serves up to 16 envs.
//   env warps (1..kB), one WARP per env: the only part that is wide — one slab per lane, its 2x2 splat targets,
//      the sparse dose read-modify-write, tumour / lung deltas (environment.py:107-110,164-182) — plus the
//      distance-to-tumour minimum (environment.py:150-162) while the scalar warp is busy.
//
// Timeline of a block (sparse mode; dense mode keeps a block barrier where the mbarrier is):
//      scalar: load, translate | rotate, beam set-up, walk   | obs, pose, next dn                  | distance reward, reward, outputs
//      env:    tumour entry    | voxel list, min distance    | dose deposition                     |
//                          barrier A       mbarrier "beams published" (per env warp)          barrier 2
// An env warp reads the tumour id from the record itself and starts its tumour loads before barrier A (it needs the
// scalar warp only for the translated position); it deposits as soon as the beams are published and its own tumour work
// is done — it never waits for the other env warps before barrier 2.  (Tried: the env warps redoing the translation too,
// no barrier A at all — 12.6 us against 12.2 at 4,096 envs: fourteen more warps on the record's line at kernel start.)
#pragma once

namespace {

struct __align__(16) EnvShared {
    Beam beam;               // scalar warp -> env warp (published through mbars[2]; dense mode: barrier 1)
    int needs_reset;         // scalar warp -> env warp (barrier A)
    int tid;
    uint32_t gen;            // the env's current dose generation
    int d_cnt;               // env warp -> scalar warp (barrier 2)
    double p[3];             // translated beam position (barrier A)
    double best;             // env warp -> scalar warp: min squared distance to the tumour (barrier 2; dense mode: barrier 1)
    double d_tum, d_lung;    // env warp -> scalar warp: dose deltas of this beam (barrier 2)
    double os_t[3];          // translation overshoot (info only), parked by the scalar warp for itself
};

constexpr int kYZStride = kMaxSlabs + 1;   // odd float2 stride: the scalar warp's lanes store to distinct banks
constexpr int kMaxPass = (kMaxSlabs + kWarp - 1) / kWarp;   // 32-slab passes of one beam (3)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// global -> shared bulk copy (TMA, no tensor map); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "RT_MBAR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra RT_MBAR_DONE;\n"
                 "bra RT_MBAR_WAIT;\n"
                 "RT_MBAR_DONE:\n"
                 "}" ::"r"(mbar), "r"(parity) : "memory");
}

// 8-byte asynchronous copy global -> shared (cp.async: the data never passes through registers)
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` (0..2, warp-uniform) of this thread's copy groups are still in flight
__device__ __forceinline__ void cp_async_wait(int pending)
{
    if (pending >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// block barrier of the scalar warp and the env warps only (the predictor warp takes no part in barriers A, 1 and 2)
template <int kThreads>
__device__ __forceinline__ void work_barrier()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
}

#define RT_STAMP3(env_, slot)                                                                    \
    do {                                                                                         \
        if (kClock) T.stage_clock[(size_t)(env_) * 12 + (slot)] = clock64();                     \
    } while (0)

// The four splat targets of slab k (draw_line.py:84-96) in two steps, so that the dose loads can be issued
// between them.  slab_coords: floor / fraction of the slab's (intery, interz), voxel coordinates, bounds;
// target j = 2*dy + dz is voxel  base + dy*g2 + dz,  bit j of `inb` says it is inside the grid.
struct SlabCoord {
    int yf, zf;
    float fy, fz;
};

__device__ __forceinline__ SlabCoord slab_coords(const Grid &G, const Beam &b, int k, float2 cur, int &base, uint32_t &inb,
                                                 int &c0, int &c1, int &c2)
{
    SlabCoord s;
    const bool have = k < b.nslab;
    const float yfl = floorf(cur.x), zfl = floorf(cur.y);                    // :76, :80
    s.yf = (int)yfl; s.zf = (int)zfl;
    s.fy = __fsub_rn(cur.x, yfl); s.fz = __fsub_rn(cur.y, zfl);              // :77, :81
    const int x = b.x0 + k * b.step;
    // idx[dom] = x, idx[o0] = yf, idx[o1] = zf (b.dom is warp-uniform)
    c0 = b.dom == 0 ? x : s.yf;
    c1 = b.dom == 0 ? s.yf : (b.dom == 1 ? x : s.zf);
    c2 = b.dom == 2 ? x : s.zf;
    const bool a0 = have && (unsigned)c0 < (unsigned)G.g0;
    const bool a10 = a0 && (unsigned)c1 < (unsigned)G.g1, a11 = a0 && (unsigned)(c1 + 1) < (unsigned)G.g1;
    const bool a20 = (unsigned)c2 < (unsigned)G.g2, a21 = (unsigned)(c2 + 1) < (unsigned)G.g2;
    base = (c0 * G.g1 + c1) * G.g2 + c2;
    inb = (a10 && a20 ? 1u : 0u) | (a10 && a21 ? 2u : 0u) | (a11 && a20 ? 4u : 0u) | (a11 && a21 ? 8u : 0u);
    return s;
}

// slab_weights: the splat weights, specialised on the warp-uniform dominant axis DOM and on QP = (step > 0 ? 0 : 1).
// Same rule as slab_targets_yz (rt_device.cuh): a voxel that two neighbouring slabs write is owned by the earlier
// slab with weight  (0 + w_k) + w_{k+1};  `drop` gets the bits of the targets the previous slab owns.
template <int DOM, int QP>
__device__ __forceinline__ void slab_weights(const Beam &b, int k, const SlabCoord &s, float2 prev, float2 next,
                                             uint32_t &drop, float (&w)[4])
{
    const float gy = __fsub_rn(1.0f, s.fy), gz = __fsub_rn(1.0f, s.fz);     // :86-87
    w[0] = __fmul_rn(gy, gz); w[1] = __fmul_rn(gy, s.fz); w[2] = __fmul_rn(s.fy, gz); w[3] = __fmul_rn(s.fy, s.fz);
    drop = 0u;
    if (DOM == 0) return;
    // q = offset along the dominant axis, o = the other one: DOM 1 -> j = 2q + o, DOM 2 -> j = 2o + q
    if (k > 0 && (int)floorf(prev.x) == s.yf) {                              // the previous slab owns what both write
        const int D = s.zf - (int)floorf(prev.y);
        constexpr uint32_t m0 = DOM == 1 ? 1u << (2 * QP) : 1u << QP;                // (q, o) = (QP, 0)
        constexpr uint32_t m1 = DOM == 1 ? 1u << (2 * QP + 1) : 1u << (2 + QP);      // (q, o) = (QP, 1)
        if ((unsigned)D <= 1u) drop |= m0;
        if ((unsigned)(D + 1) <= 1u) drop |= m1;
    }
    if (k + 1 < b.nslab) {                                                   // add the next slab's weight for shared voxels
        const float nyfl = floorf(next.x), nzfl = floorf(next.y);
        if ((int)nyfl == s.yf) {
            const int D = s.zf - (int)nzfl;
            const float nfy = __fsub_rn(next.x, nyfl), nfz = __fsub_rn(next.y, nzfl);
            const float ngy = __fsub_rn(1.0f, nfy), ngz = __fsub_rn(1.0f, nfz);
            // my target (q, o) = (1-QP, o) is the neighbour's (QP, o2 = D + o) when o2 is 0 or 1
#pragma unroll
            for (int o = 0; o < 2; o++) {
                const int o2 = D + o;
                if ((unsigned)o2 <= 1u) {
                    const float pw = DOM == 1 ? __fmul_rn(QP ? nfy : ngy, o2 ? nfz : ngz)
                                              : __fmul_rn(o2 ? nfy : ngy, QP ? nfz : ngz);
                    constexpr int jq = DOM == 1 ? 2 * (1 - QP) : (1 - QP);
                    const int j = DOM == 1 ? jq + o : jq + 2 * o;
                    if (j == 0) w[0] = __fadd_rn(w[0], pw);
                    else if (j == 1) w[1] = __fadd_rn(w[1], pw);
                    else if (j == 2) w[2] = __fadd_rn(w[2], pw);
                    else w[3] = __fadd_rn(w[3], pw);
                }
            }
        }
    }
}

// Two neighbouring bits (l, l + 1) of a bitmask padded by kLungPadBits leading zero bits and a trailing zero word.
__device__ __forceinline__ uint32_t bit_pair(const uint32_t *bits, int l)
{
    const int p = l + kLungPadBits;
    return __funnelshift_r(bits[p >> 5], bits[(p >> 5) + 1], p & 31) & 3u;
}
__device__ __forceinline__ uint32_t bit_pair_ldg(const uint32_t *bits, int l)
{
    const int p = l + kLungPadBits;
    return __funnelshift_r(__ldg(bits + (p >> 5)), __ldg(bits + (p >> 5) + 1), p & 31) & 3u;
}

// The dose cells a beam touches are known only after the float64 pose update, the beam set-up and the serial walk —
// 6,000 cycles during which the env warps have nothing to do — and then every env of the GPU asks HBM for them in the
// same microsecond (a beam rarely revisits a brick, so they all miss the L2): 20 MB at 4096 envs, a 3 us burst.  So a
// PREDICTOR warp (one thread per env) redoes the pose update in float32 with fast intrinsics as soon as the record and
// the action are loaded — same formulas as transforms.py, then the exact beam_setup on the predicted direction — and
// the env warps prefetch the bricks of the predicted tube into the L2 (closed-form walk) while the scalar warp is
// still busy with the exact chain.  Nothing depends on the prediction being exact: a brick it misses costs that lane
// one HBM round trip later, a brick it fetches in vain 128 bytes of bandwidth.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// draw_line.py:19-66 with fast divisions: the same clip and set-up as beam_setup (rt_device.cuh), good to a few ulp, which
// is all a prefetch needs
__device__ __forceinline__ Beam beam_setup_fast(const Grid &G, const float p[3], float v0, float v1, float v2)
{
    const float eps = 9.99999997475242708e-07f;
    Beam b;
    b.nslab = 0; b.dom = 0; b.o0 = 1; b.o1 = 2; b.step = 1; b.x0 = 0;
    b.y0 = b.z0 = b.sgy = b.sgz = 0.0f;
    const float v[3] = {v0, v1, v2};
    const int gsz[3] = {G.g0, G.g1, G.g2};
    float t_entry = -CUDART_INF_F, t_exit = CUDART_INF_F;
    bool empty = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float gm1 = (float)(gsz[i] - 1);
        float te = -CUDART_INF_F, tx = CUDART_INF_F;
        if (fabsf(v[i]) > eps) {
            const float r = __frcp_rn(v[i]);
            const float t1 = -p[i] * r, t2 = (gm1 - p[i]) * r;
            te = fminf(t1, t2); tx = fmaxf(t1, t2);
        } else if (p[i] < 0.0f || p[i] > gm1) empty = true;
        t_entry = fmaxf(t_entry, te);
        t_exit = fminf(t_exit, tx);
    }
    if (empty || t_entry > t_exit) return b;
    const float a0 = fabsf(v[0]), a1 = fabsf(v[1]), a2 = fabsf(v[2]);
    int dom = 0;
    float best = a0;
    if (a1 > best) { best = a1; dom = 1; }
    if (a2 > best) { best = a2; dom = 2; }
    const int o0 = dom == 0 ? 1 : 0, o1 = dom == 2 ? 1 : 2;
    const float vd = dom == 0 ? v[0] : (dom == 1 ? v[1] : v[2]), pdm = dom == 0 ? p[0] : (dom == 1 ? p[1] : p[2]);
    const float vo0 = o0 == 0 ? v[0] : v[1], po0 = o0 == 0 ? p[0] : p[1];
    const float vo1 = o1 == 1 ? v[1] : v[2], po1 = o1 == 1 ? p[1] : p[2];
    const int step = vd > 0.0f ? 1 : -1;
    const int x0 = (int)floorf(pdm + t_entry * vd), x1 = (int)floorf(pdm + t_exit * vd);
    const float rden = __frcp_rn(vd + eps);
    b.y0 = po0 + t_entry * vo0;
    b.z0 = po1 + t_entry * vo1;
    b.sgy = step > 0 ? vo0 * rden : -(vo0 * rden);
    b.sgz = step > 0 ? vo1 * rden : -(vo1 * rden);
    b.dom = dom; b.o0 = o0; b.o1 = o1; b.step = step; b.x0 = x0;
    const int n = (x1 - x0) * step + 1;
    b.nslab = n < 0 ? 0 : n;
    return b;
}

__device__ __forceinline__ Beam predict_beam(const Grid &G, const double pd[3], const double dn[3], const float ar[3])
{
    const float rx = clip1(ar[0]) * 1.57079633f, ry = clip1(ar[1]) * 1.57079633f, rz = clip1(ar[2]) * 1.57079633f;
    const float d0 = (float)dn[0], d1 = (float)dn[1], d2 = (float)dn[2];
    const float th2 = rx * rx + ry * ry + rz * rz, th = sqrtf(th2);
    float sc, qw;
    if (th <= 1e-3f) { sc = 0.5f - th2 * (1.0f / 48.0f); qw = 1.0f - th2 * 0.125f; }
    else { float sh, ch; __sincosf(0.5f * th, &sh, &ch); sc = __fdividef(sh, th); qw = ch; }
    const float x = sc * rx, y = sc * ry, z = sc * rz, w = qw;
    const float x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w, xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    float r0 = (x2 - y2 - z2 + w2) * d0 + 2.0f * (xy - zw) * d1 + 2.0f * (xz + yw) * d2;
    float r1 = 2.0f * (xy + zw) * d0 + (-x2 + y2 - z2 + w2) * d1 + 2.0f * (yz - xw) * d2;
    float r2 = 2.0f * (xz - yw) * d0 + 2.0f * (yz + xw) * d1 + (-x2 - y2 + z2 + w2) * d2;
    const float inv = rsqrtf(r0 * r0 + r1 * r1 + r2 * r2);
    r0 *= inv; r1 *= inv; r2 *= inv;
    if (fabsf(r0) >= 0.70710678f) {                                        // transforms.py:35-51
        const float pn = sqrtf(r1 * r1 + r2 * r2);
        const float px = pn < 1e-8f ? 1.0f : __fdividef(r1, pn), py = pn < 1e-8f ? 0.0f : __fdividef(r2, pn);
        r0 = copysignf(0.70710678f, r0); r1 = px * 0.70710678f; r2 = py * 0.70710678f;
    }
    const float inv2 = rsqrtf(r0 * r0 + r1 * r1 + r2 * r2);
    const float pf[3] = {(float)pd[0], (float)pd[1], (float)pd[2]};
    return beam_setup_fast(G, pf, r0 * inv2, r1 * inv2, r2 * inv2);         // draw_line.py:19-66 on the predicted direction
}

__device__ __forceinline__ void prefetch_beam(const Grid &G, const Beam &b, const uint2 *vol, int lane)
{
    for (int k = lane; k < b.nslab; k += kWarp) {
        const float fk = (float)k;
        const float yy = fmaf(fk, b.sgy, b.y0), zz = fmaf(fk, b.sgz, b.z0);  // closed form of the walk: off by an ulp or two
        const int yf = (int)floorf(yy), zf = (int)floorf(zz), xx = b.x0 + k * b.step;
        const int c0 = b.dom == 0 ? xx : yf, c1 = b.dom == 0 ? yf : (b.dom == 1 ? xx : zf), c2 = b.dom == 2 ? xx : zf;
        if ((unsigned)c0 >= (unsigned)G.g0) continue;
        // the 2x2 footprint (c1..c1+1, c2..c2+1) lies in one brick unless c1 is odd or c2 is the last voxel of a brick row
        const int j0 = min(max(c1, 0), G.g1 - 1), j1 = min(max(c1 + 1, 0), G.g1 - 1);
        const int k0 = min(max(c2, 0), G.g2 - 1), k1 = min(max(c2 + 1, 0), G.g2 - 1);
        const bool dj = (j1 >> 1) != (j0 >> 1), dk = (k1 >> 2) != (k0 >> 2);
        // a brick is one 128-byte line of 16 cells: its first cell's address will do
        const int br0 = ((c0 >> 1) * G.nb1 + (j0 >> 1)) * G.nb2, br1 = ((c0 >> 1) * G.nb1 + (j1 >> 1)) * G.nb2;
        prefetch_l2(vol + (br0 + (k0 >> 2)) * 16);
        if (dk) prefetch_l2(vol + (br0 + (k1 >> 2)) * 16);
        if (dj) prefetch_l2(vol + (br1 + (k0 >> 2)) * 16);
        if (dj && dk) prefetch_l2(vol + (br1 + (k1 >> 2)) * 16);
    }
}

// Env warp, before the beam is known — part 1, needs only the env's tumour id: the tumour entry and its padded bitmask
// into shared memory, then the first four voxels per lane of the tumour's voxel list requested (pk4).
template <bool kDense>
__device__ __forceinline__ void env_tumour_fetch(const Tables &T, int tid, Tumour &tm, uint32_t *tb, int lane, uint32_t (&pk4)[4])
{
    if (lane < kTumourWords)
        reinterpret_cast<uint32_t *>(&tm)[lane] = __ldg(reinterpret_cast<const uint32_t *>(T.tumours + tid) + lane);
    if (!kDense)
        for (int i = lane; i < T.pbits_words; i += kWarp)              // pbits_words <= kMaxPTumourWords (rt_create)
            tb[i] = __ldg(T.tumour_pbits + (size_t)tid * T.pbits_words + i);
    __syncwarp();
    const int nv = tm.n_vox;
    const uint32_t *vx = T.vox_xyz + tm.vox_off;
#pragma unroll
    for (int i = 0; i < 4; i++) pk4[i] = lane + i * kWarp < nv ? __ldg(vx + lane + i * kWarp) : 0xffffffffu;
}

// Part 2, needs the translated beam position se.p: the distance-to-tumour reward's minimum (environment.py:150-162) over
// the tumour's voxel list -> se.best.  `meanwhile` runs before the first voxels are used (the per-call kernel prefetches
// the predicted beam there).
template <typename Meanwhile>
__device__ __forceinline__ void env_distance(const Tables &T, EnvShared &se, const Tumour &tm, int lane, const uint32_t (&pk4)[4],
                                             Meanwhile meanwhile)
{
    const double p0 = se.p[0], p1 = se.p[1], p2 = se.p[2];
    double best = CUDART_INF;
    const int nv = tm.n_vox;
    const uint32_t *vx = T.vox_xyz + tm.vox_off;
    meanwhile();
    auto take = [&](uint32_t pk) {
        const double dx = (double)(pk & 255u) - p0;
        const double dy = (double)((pk >> 8) & 255u) - p1;
        const double dz = (double)(pk >> 16) - p2;
        const double d2 = __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
        best = d2 < best ? d2 : best;
    };
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (lane + i * kWarp < nv) take(pk4[i]);
#pragma unroll 4
    for (int k = lane + 4 * kWarp; k < nv; k += kWarp) take(__ldg(vx + k));
    // min of non-negative doubles = min of their bit patterns: two integer warp reductions
    const uint32_t hi = (uint32_t)__double2hiint(best);
    const uint32_t mhi = __reduce_min_sync(kFull, hi);
    const uint32_t mlo = __reduce_min_sync(kFull, hi == mhi ? (uint32_t)__double2loint(best) : 0xffffffffu);
    if (lane == 0) se.best = __hiloint2double((int)mhi, (int)mlo);
}

// Env warp, once the beam and its walk are published: the dose deposition of one env (environment.py:107-110) and the
// tumour / lung deltas it causes (environment.py:164-182), left in se.d_tum / se.d_lung / se.d_cnt.  `vol` = the env's
// cells, `cbuf` = this warp's [kMaxPass][32 slabs][4 targets] item slots, `slungs` = the padded lungs bitmask (shared
// memory when kStageLungs, then `lungs_mbar` is the mbarrier its bulk copy completes on).
template <bool kStageLungs, bool kClock>
__device__ __forceinline__ void deposit_beam(const Tables &T, EnvShared &se, const Tumour &tm, const uint32_t *tb, const float2 *myz,
                                             uint2 *vol, uint2 *cbuf, const uint32_t *slungs, uint32_t lungs_mbar, int env, int lane)
{
    if (lane == 0) RT_STAMP3(env, 3);
    // ---- dose deposition (environment.py:107-110): dose' = clip(dose + beam*0.1, 0, 1) on the voxels hit.
    // A voxel has exactly one owner lane per beam (slab_weights merges the two slabs that can meet), so the
    // 32-slab passes of a beam are independent.
    //   Phase 1 issues the loads of every pass before it uses any (the beam pays one HBM round trip however long it
    //   is); phase 2 takes the passes in order: weights, masks, new dose, deltas, stores.
    //   The MEMORY instructions use another lane mapping than the arithmetic: a lane computes one slab and its
    //   four targets, but load / store instruction i of a pass moves item 32 i + lane of the pass's item list
    //   [slab][target] — the four targets of eight consecutive slabs.  The load/store pipeline pays per distinct
    //   128-byte line of an instruction: with one target of 32 slabs per instruction that is ~15 lines, with all
    //   targets of 8 slabs ~4 (the tube of a beam crosses a brick every 2 to 4 slabs).  Cell indices and cells
    //   change lanes through one shared-memory slot per item (item-major, conflict-free both ways).
    const Grid &G = T.G;
    const Beam b = se.beam;
    const uint32_t gen = se.gen;
    const int g2 = G.g2;
    const int npass = (b.nslab + kWarp - 1) / kWarp;
#pragma unroll
    for (int c = 0; c < kMaxPass; c++) {
        if (c < npass) {                                                       // warp-uniform
            const int k = c * kWarp + lane;
            const float2 cur = myz[k < b.nslab ? k : 0];
            int base, c0, c1, c2;
            uint32_t inb;
            slab_coords(G, b, k, cur, base, inb, c0, c1, c2);
            const int r0 = cell_row_term(G, c0, c1), r1 = cell_row_term(G, c0, c1 + 1);
            const int q0 = cell_col_term(c2), q1 = cell_col_term(c2 + 1);
            // the cell index of every target goes into its own slot (-1: outside the grid) ...
            uint4 *mine4 = reinterpret_cast<uint4 *>(cbuf + c * 4 * kWarp) + 2 * lane;
            mine4[0] = make_uint4(0u, (uint32_t)(inb & 1u ? r0 + q0 : -1), 0u, (uint32_t)(inb & 2u ? r0 + q1 : -1));
            mine4[1] = make_uint4(0u, (uint32_t)(inb & 4u ? r1 + q0 : -1), 0u, (uint32_t)(inb & 8u ? r1 + q1 : -1));
        }
    }
    __syncwarp();
    // ... every load of the beam goes out (item-major lane mapping) before the first one is used; then the dose as this
    // generation sees it joins the index: slot = {dose, cell index}
    uint2 got[kMaxPass][4];
#pragma unroll
    for (int c = 0; c < kMaxPass; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            got[c][i] = make_uint2(0u, ~gen);                                  // outside the grid: reads as "another generation"
            if (c < npass) {
                const int a = (int)cbuf[c * 4 * kWarp + i * kWarp + lane].y;
                if (a >= 0) got[c][i] = __ldcg(vol + a);
            }
        }
#pragma unroll
    for (int c = 0; c < kMaxPass; c++)
        if (c < npass) {
#pragma unroll
            for (int i = 0; i < 4; i++)                                        // another generation reads as zero
                cbuf[c * 4 * kWarp + i * kWarp + lane].x = got[c][i].y == gen ? got[c][i].x : 0u;
        }
    __syncwarp();
    if (lane == 0) RT_STAMP3(env, 4);
    if (kStageLungs) mbar_wait(lungs_mbar, 0);                         // the lungs bitmask has landed
    const int li0 = tm.lo[0], li1 = tm.lo[1] - 1, li2 = tm.lo[2] - 1;  // origin of the padded bbox
    const int td0 = tm.dim[0], td1 = tm.dim[1], td2 = tm.dim[2];
    const int pd1 = td1 + 2, pd2 = td2 + 2;
    const int variant = b.dom == 0 ? 0 : (b.dom * 2 - 1 + (b.step > 0 ? 0 : 1));   // warp-uniform
    // per-lane partial sums of at most 12 float32 deltas; the running totals are float64 (scalar warp)
    float d_tum = 0.0f, d_lung = 0.0f;
    int d_cnt = 0;
#pragma unroll 1
    for (int c = 0; c < npass; c++) {
        const int k = c * kWarp + lane;
        const int kk = k < b.nslab ? k : 0;
        const float2 cur = myz[kk], prv = myz[kk > 0 ? kk - 1 : 0], nxt = myz[kk + 1 < b.nslab ? kk + 1 : kk];
        int base, c0, c1, c2;
        uint32_t inb;
        const SlabCoord sc = slab_coords(G, b, k, cur, base, inb, c0, c1, c2);
        uint32_t drop;
        float w[4];
        switch (variant) {
        case 0: slab_weights<0, 0>(b, k, sc, prv, nxt, drop, w); break;
        case 1: slab_weights<1, 0>(b, k, sc, prv, nxt, drop, w); break;
        case 2: slab_weights<1, 1>(b, k, sc, prv, nxt, drop, w); break;
        case 3: slab_weights<2, 0>(b, k, sc, prv, nxt, drop, w); break;
        default: slab_weights<2, 1>(b, k, sc, prv, nxt, drop, w); break;
        }
        const uint32_t ok = inb & ~drop;
        // lungs membership: the two targets of a row are neighbouring bits of the (padded) bitmask
        uint32_t lmask = 0u;
        if (inb & 3u) lmask = kStageLungs ? bit_pair(slungs, base) : bit_pair_ldg(slungs, base);
        if (inb & 12u) lmask |= (kStageLungs ? bit_pair(slungs, base + g2) : bit_pair_ldg(slungs, base + g2)) << 2;
        // tumour membership of the 2x2 block: one range test against the bounding box grown by one voxel on
        // axes 1 and 2 (the padded bitmask has empty border cells, so the four bits are always addressable)
        uint32_t tmask = 0u;
        const int ti = c0 - li0, tj = c1 - li1, tk = c2 - li2;
        if ((unsigned)ti < (unsigned)td0 && (unsigned)tj <= (unsigned)td1 && (unsigned)tk <= (unsigned)td2) {
            const int b0 = (ti * pd1 + tj) * pd2 + tk, b1 = b0 + pd2;
            const uint32_t r0 = __funnelshift_r(tb[b0 >> 5], tb[(b0 >> 5) + 1], b0 & 31) & 3u;
            const uint32_t r1 = __funnelshift_r(tb[b1 >> 5], tb[(b1 >> 5) + 1], b1 & 31) & 3u;
            tmask = r0 | (r1 << 2);
        }
        tmask &= ok;
        lmask &= ok;
        const uint32_t cmask = lmask & ~tmask;                 // lungs_mask = lungs*(1-tumours) (environment.py:174)
        uint4 *mine4 = reinterpret_cast<uint4 *>(cbuf + c * 4 * kWarp) + 2 * lane;
        const uint4 ca = mine4[0], cb = mine4[1];
        const uint32_t cv[4] = {ca.x, ca.z, cb.x, cb.z}, ci[4] = {ca.y, ca.w, cb.y, cb.w};
        float nd[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float o = __uint_as_float(cv[j]);
            nd[j] = fminf(__fadd_rn(o, __fmul_rn(w[j], 0.100000001490116119f)), 1.0f);   // clip(dose + beam*0.1, 0, 1), dose >= 0
            const float delta = (ok >> j) & 1u ? nd[j] - o : 0.0f;                    // 0 for targets this lane does not write
            d_tum += (tmask >> j) & 1u ? delta : 0.0f;
            d_lung += (lmask >> j) & 1u ? delta : 0.0f;
            // dose is monotone, so the count only grows (environment.py:175-177)
            d_cnt += (int)((cmask >> j) & 1u) & (int)(!(o > 0.200000002980232239f) && nd[j] > 0.200000002980232239f);
        }
        // {new dose, cell index or -1 for targets this lane does not write} back into the item list, then store
        // instruction i writes items 32 i + lane
        mine4[0] = make_uint4(__float_as_uint(nd[0]), ok & 1u ? ci[0] : 0xffffffffu, __float_as_uint(nd[1]), ok & 2u ? ci[1] : 0xffffffffu);
        mine4[1] = make_uint4(__float_as_uint(nd[2]), ok & 4u ? ci[2] : 0xffffffffu, __float_as_uint(nd[3]), ok & 8u ? ci[3] : 0xffffffffu);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint2 it = cbuf[c * 4 * kWarp + i * kWarp + lane];
            if ((int)it.y >= 0) __stcg(vol + (int)it.y, make_uint2(it.x, gen));
        }
    }
    if (lane == 0) RT_STAMP3(env, 6);
    // both sums in one butterfly: after the first exchange the lower half-warp carries the tumour sum, the
    // upper one the lung sum
    {
        const bool upper = lane >= 16;
        const float keep = upper ? d_lung : d_tum, give = upper ? d_tum : d_lung;
        double v = (double)keep + (double)__shfl_xor_sync(kFull, give, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        d_cnt = __reduce_add_sync(kFull, d_cnt);
        if (lane == 0) { se.d_tum = v; se.d_cnt = d_cnt; }
        if (lane == 16) se.d_lung = v;
    }
}

// warps of a block: scalar warp, kB env warps and, for 14-env blocks of sparse-mode handles, the predictor warp
template <int kB, bool kDense>
__host__ __device__ constexpr int step_block_threads() { return (kB + 1 + (kB >= 14 && !kDense ? 1 : 0)) * kWarp; }

// Shared memory of a block (static part).
template <int kB>
struct StepShared {
    EnvShared sh[kB];
    Tumour tum[kB];
    uint32_t tbits[kB][kMaxPTumourWords];
    float2 yz[kB][kYZStride];
    __align__(8) unsigned long long mbars[3];          // [0] lungs bitmask landed, [1] predicted beams published, [2] beams and walks published
    Beam pred[kB];
    // outputs are staged here by the scalar warp's lanes and copied out row-contiguously (full-line stores:
    // the host-buffer entry points map these arrays over PCIe)
    float s_obs[kB * RT_OBS_SIZE];
    double s_rew[kB];
    uint8_t s_term[kB];
    double s_info[kB * RT_INFO_SIZE];
};

// What the fused rollout kernel (rt_rollout.cuh) needs from a step besides the env state: row t of the rewards buffer
// and the episode statistics of train.py:42-66.  Unused (all NULL) by rt_step_kernel.
struct RollStep {
    float *rewards_row;          // [N] row t of rewards [T][N] (train.py:154)
    double *episode_stats;       // [7] finished, sum return, sum length, sum last-step tumour / lung / distance / total reward
};

// One environment step of the block's kB envs (see the top of the file).  `it` counts the steps this block has done
// in this launch (0 for rt_step_kernel; the rollout kernel loops): the lungs bitmask is staged by step 0 only and
// the predictor's mbarrier alternates its phase.  kRoll: `actions` is a shared-memory array [kB][6] indexed by the
// env's position in the block, the observation stays in M.s_obs, reward / done go to `R` and M.s_term.
template <int kB, bool kClock, bool kDense, bool kRoll>
__device__ __forceinline__ void step_block(StepShared<kB> &M, uint32_t *dyn_smem, const Tables &T, const Schedule &S, EnvRec *rec,
                                           uint2 *cells, double *beams, int n_envs, const float *actions, const StepOut &out,
                                           DenseWork *dense, int env0, uint32_t it, const RollStep &R)
{
    EnvShared (&sh)[kB] = M.sh;
    Tumour (&tum)[kB] = M.tum;
    uint32_t (&tbits)[kB][kMaxPTumourWords] = M.tbits;
    float2 (&yz)[kB][kYZStride] = M.yz;
    unsigned long long (&mbars)[3] = M.mbars;
    Beam (&pred)[kB] = M.pred;
    float (&s_obs)[kB * RT_OBS_SIZE] = M.s_obs;
    double (&s_rew)[kB] = M.s_rew;
    uint8_t (&s_term)[kB] = M.s_term;
    double (&s_info)[kB * RT_INFO_SIZE] = M.s_info;
    // dynamic: [kB][kMaxPass][32 slabs][4 targets] item slots of 8 bytes (cell index, then the prefetched cell, then the new
    // dose + cell index), then [lung_words16] padded lungs bitmask (kB >= 14)
    uint2 *cellbuf_all = reinterpret_cast<uint2 *>(dyn_smem);
    uint32_t *lungs_sm = dyn_smem + (kDense ? 0 : kB * kMaxPass * 4 * kWarp * 2);
    const Grid &G = T.G;
    // Warp roles by warp index: the scheduler of an SM sub-partition prefers the warp with the highest index, and the
    // scalar warp carries the block's critical path, so it is the LAST warp; the predictor (if any) is warp 0 and the env
    // warps sit between them.
    constexpr int kEnvWarp0 = (kB >= 14 && !kDense) ? 1 : 0;      // first env warp
    constexpr int kScalarWarp = kEnvWarp0 + kB;
    const int warp = threadIdx.x / kWarp;
    const int lane = threadIdx.x & (kWarp - 1);
    // blocks of 7 envs run four to an SM and read the lungs bitmask through L1 instead of staging it four times
    constexpr bool kStageLungs = kB >= 14 && !kDense;
    constexpr bool kPredict = kB >= 14 && !kDense;
    const uint32_t pred_parity = it & 1u;

    // =====================================================================================================
    if (warp == kScalarWarp) {
        const int e = env0 + lane;
        const bool mine = lane < kB && e < n_envs;
        if (kStageLungs && lane == 0 && it == 0)
            bulk_load(smem_u32(lungs_sm), T.lungs_pad, (uint32_t)(T.lung_words16 * sizeof(uint32_t)), smem_u32(&mbars[0]));
        if (mine) RT_STAMP3(e, 0);
        EnvRec *my = rec + (mine ? e : 0);
        EnvShared &se = sh[lane < kB ? lane : 0];
        const double gs[3] = {(double)G.g0, (double)G.g1, (double)G.g2};
        bool stepping = false;
        Pose s;
        double dn[3] = {0.0, 1.0, 0.0};                                    // direction / |direction| (transforms.py:23)
        float ar[3] = {0.f, 0.f, 0.f};                                     // rotation part of the action
        int tid = 0;
        uint32_t gen = 0u;
        if (mine) {
            // every load is issued before the first use: one round trip to L2
            const int needs_reset = my->needs_reset;
            tid = my->tumour_id;
            gen = my->gen;
            double p0[3];
#pragma unroll
            for (int i = 0; i < 3; i++) { p0[i] = my->pos[i]; dn[i] = my->dn[i]; }
            float2 a01, a23, a45;
            if (kRoll) {
                const float2 *ap = reinterpret_cast<const float2 *>(actions + lane * RT_ACTION_SIZE);
                a01 = ap[0]; a23 = ap[1]; a45 = ap[2];
            } else {
                const float2 *ap = reinterpret_cast<const float2 *>(actions + (size_t)e * RT_ACTION_SIZE);
                a01 = __ldg(ap); a23 = __ldg(ap + 1); a45 = __ldg(ap + 2);
            }
            const float at[3] = {a01.x, a01.y, a23.x};
            ar[0] = a23.y; ar[1] = a45.x; ar[2] = a45.y;
            stepping = needs_reset == 0;
            se.needs_reset = needs_reset;
            se.tid = tid;
            se.gen = gen;
            se.beam.nslab = 0;
            if (stepping) {
#pragma unroll
                for (int i = 0; i < 3; i++) {                              // environment.py:122-125, transforms.py:65-67
                    double os;
                    s.p[i] = translate_axis(p0[i], __dmul_rn(__dmul_rn((double)clip1(at[i]), gs[i]), 0.2), gs[i], os);
                    se.p[i] = s.p[i];
                    if (out.info) se.os_t[i] = os;
                }
                if (kClock) T.stage_clock[(size_t)e * 12 + 8] = clock64() + (long long)(s.p[0] * 0.0);
            }
        }
        work_barrier<(kB + 1) * kWarp>();                                  // ---- barrier A
        double zc = 0.0;
        if (stepping) {
            double rv[3];
#pragma unroll
            for (int i = 0; i < 3; i++)                                    // environment.py:139-141
                rv[i] = (double)__fmul_rn(__fmul_rn(clip1(ar[i]), 3.14159274101257324f), 0.5f);
            zc = rotate_normalized(dn, rv, s.d);                           // transforms.py:25-55 (:23 was done by the previous step)
            if (kClock) T.stage_clock[(size_t)e * 12 + 9] = clock64() + (long long)(s.d[0] * 0.0);
            const Beam b = beam_setup(G, s.p, s.d);                        // draw_line.py:19-66
            if (kClock) T.stage_clock[(size_t)e * 12 + 10] = clock64() + (b.nslab < -5);
            beam_walk2_uniform(b, yz[lane]);                               // draw_line.py:98-99 (rows hold kMaxSlabs + 1 entries)
            se.beam = b;
            RT_STAMP3(e, 1);
        }
        if (kDense) {
            work_barrier<(kB + 1) * kWarp>();                              // ---- barrier 1
        } else {
            // beams and walks published: every env warp goes on as soon as ITS tumour work is done, not the block's slowest
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&mbars[2])) : "memory");
        }

        // While the env warps deposit the dose: everything that does not depend on it.
        const Tumour &tm = tum[lane < kB ? lane : 0];                       // sparse mode: staged by the env warp, complete at barrier 2
        const Tumour *tg = T.tumours + tid;
        int t = 0, n_beams = 0, lung_count = 0;
        double r_dist = 0.0, os_r = 0.0, rcp_mask = 0.0, mask_sum = 1.0;
        double tumour_dose = 0.0, lung_dose = 0.0, ep_return = 0.0;
        if (stepping) {
            tumour_dose = my->tumour_dose; lung_dose = my->lung_dose; ep_return = my->ep_return;   // consumed after barrier 2
            t = my->t + 1;                                                 // environment.py:194
            lung_count = my->lung_count;
            n_beams = my->n_beams;
            float *obs = s_obs + lane * RT_OBS_SIZE;                       // environment.py:259-268
            double nd0 = s.d[0], nd1 = s.d[1], nd2 = s.d[2];
            normalize3(nd0, nd1, nd2);                                     // the next step's transforms.py:23
            my->dn[0] = nd0; my->dn[1] = nd1; my->dn[2] = nd2;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                obs[i] = (float)__dsub_rn(__dmul_rn(__ddiv_rn(s.p[i], gs[i]), 2.0), 1.0);
                obs[3 + i] = (float)s.d[i];
                obs[6 + i] = __ldg(&tg->obs_c[i]);
                my->pos[i] = s.p[i];
                my->dir[i] = s.d[i];
            }
            if (kDense) r_dist = __dmul_rn(__ddiv_rn(sqrt(se.best), T.gnorm), -1.0);   // environment.py:158-162
            mask_sum = (double)__ldg(&tg->lung_mask_sum);
            rcp_mask = __drcp_rn(mask_sum);
            if (out.info || kDense) os_r = overshoot_from_z(zc);           // transforms.py:29-33, 57 (info only)
            if (beams && n_beams < RT_MAX_TIME_STEPS) {                    // environment.py:110
                double *bp = beams + ((size_t)e * RT_MAX_TIME_STEPS + n_beams) * 6;
#pragma unroll
                for (int i = 0; i < 3; i++) { bp[i] = s.p[i]; bp[3 + i] = s.d[i]; }
            }
            if (kDense) {
                // dense mode: rt_dense_kernel streams the volume, recomputes the reductions and finishes the step
                DenseWork &dw = dense[e];
                dw.mode = 0;
                dw.tid = tid; dw.t = t; dw.n_beams = n_beams;
                dw.best = se.best;
                dw.os_t[0] = se.os_t[0]; dw.os_t[1] = se.os_t[1]; dw.os_t[2] = se.os_t[2];
                dw.os_r = os_r;
                dw.ep_return = ep_return;
                my->t = t;
                my->n_beams = n_beams + 1;
            }
        } else if (mine) {
            // gymnasium 1.0.0 NEXT_STEP: the call after a terminal step resets and reports reward 0
            // (environment.py:77-105).  A new generation empties the dose volume.
            const int episode = my->episode + 1;
            tid = pick_tumour(T, S, e, n_envs, episode);
            const Tumour *tg = T.tumours + tid;
            float *obs = s_obs + lane * RT_OBS_SIZE;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double p = gs[i] / 2.0, d = i == 1 ? 1.0 : 0.0;
                my->pos[i] = p;
                my->dir[i] = d;
                my->dn[i] = d;
                obs[i] = (float)__dsub_rn(__dmul_rn(__ddiv_rn(p, gs[i]), 2.0), 1.0);
                obs[3 + i] = (float)d;
                obs[6 + i] = __ldg(&tg->obs_c[i]);
            }
            my->tumour_dose = 0.0; my->lung_dose = 0.0; my->ep_return = 0.0;
            my->t = 0; my->tumour_id = tid; my->lung_count = 0; my->episode = episode; my->needs_reset = 0; my->n_beams = 0;
            my->gen = gen + 1u;
            if (kDense) dense[e].mode = 1;                                 // rt_dense_kernel zeroes the volume
            s_rew[lane] = 0.0;
            s_term[lane] = 0;
            if (out.info) {
                double *ip = s_info + lane * RT_INFO_SIZE;
#pragma unroll
                for (int i = 0; i < RT_INFO_SIZE; i++) ip[i] = i == RT_INFO_TUMOUR_ID ? (double)tid : 0.0;
            }
        }
        const int nb = min(kB, n_envs - env0);                             // envs of this block
        __syncwarp();
        if (out.obs)
            for (int i = lane; i < nb * RT_OBS_SIZE; i += kWarp) out.obs[(size_t)env0 * RT_OBS_SIZE + i] = s_obs[i];
        if (kStageLungs) mbar_wait(smem_u32(&mbars[0]), 0);                // the lungs copy must land before the block retires
        work_barrier<(kB + 1) * kWarp>();                                  // ---- barrier 2
        if (kDense) {
            // the dense kernel writes reward / terminated / info of the envs that stepped
            if (mine && !stepping) {
                if (out.reward) out.reward[e] = 0.0;
                if (out.reward_f32) out.reward_f32[e] = 0.0f;
                if (out.terminated) out.terminated[e] = 0;
                if (out.truncated) out.truncated[e] = 0;
                if (out.info)
                    for (int i = 0; i < RT_INFO_SIZE; i++) out.info[(size_t)e * RT_INFO_SIZE + i] = s_info[lane * RT_INFO_SIZE + i];
            }
        }
        if (!kDense && stepping) {
            RT_STAMP3(e, 11);
            r_dist = __dmul_rn(__ddiv_rn(sqrt(se.best), T.gnorm), -1.0);   // environment.py:158-162 (se.best: env warp, before its deposition)
            tumour_dose += se.d_tum;
            lung_dose += se.d_lung;
            lung_count += se.d_cnt;
            // rewards, termination (environment.py:158-191, 214-220)
            float tsum_f32 = (float)tumour_dose;                           // np.sum(dose*tumours) float32
            float ratio = fdiv_rn_zero_num(tsum_f32, tm.tumour_sum);       // no dose on the tumour yet: the common case
            if (fabsf(ratio - kDoneRatio) < kDoneWindow) {
                // Within 2e-5 of the termination threshold the summation ORDER of np.sum decides `done`
                // (environment.py:186-190): redo the sum exactly as NumPy's pairwise float32 reduction does.
                const uint2 *vol = cells + (size_t)e * G.cstride;
                const int g1 = G.g1, g2 = G.g2, nb1 = G.nb1, nb2 = G.nb2;       // by value: nothing of the kernel's parameters is addressed
                tsum_f32 = np_pairwise_sparse(G.nvox, T.vox_xyz + tm.vox_off, tm.n_vox, g1, g2, [=](int lin) {
                    const int k = lin % g2, ij = lin / g2, i = ij / g1, j = ij - i * g1;
                    const uint2 c = __ldcg(vol + ((((i >> 1) * nb1 + (j >> 1)) * nb2) * 16 + (i & 1) * 8 + (j & 1) * 4 + (k >> 2) * 16 + (k & 3)));
                    return c.y == gen ? __uint_as_float(c.x) : 0.0f;
                });
                ratio = __fdiv_rn(tsum_f32, tm.tumour_sum);
            }
            const float r_tumour = __fmul_rn(ratio, 10.0f);
            const double r_lung = __dmul_rn(div_shared((double)lung_count, mask_sum, rcp_mask), -1.0);
            const double reward = __dadd_rn(__dadd_rn((double)r_tumour, r_lung), r_dist);
            const bool done = (ratio >= kDoneRatio) || (t >= RT_MAX_TIME_STEPS);
            ep_return += reward;
            my->tumour_dose = tumour_dose; my->lung_dose = lung_dose; my->ep_return = ep_return;
            my->t = t; my->lung_count = lung_count; my->needs_reset = done ? 1 : 0; my->n_beams = n_beams + 1;
            s_rew[lane] = reward;
            s_term[lane] = done ? 1 : 0;
            if (kRoll && done && R.episode_stats) {                        // train.py:42-66, 160-161
                atomicAdd(R.episode_stats + 0, 1.0);
                atomicAdd(R.episode_stats + 1, ep_return);
                atomicAdd(R.episode_stats + 2, (double)t);
                atomicAdd(R.episode_stats + 3, (double)r_tumour);
                atomicAdd(R.episode_stats + 4, r_lung);
                atomicAdd(R.episode_stats + 5, r_dist);
                atomicAdd(R.episode_stats + 6, reward);
            }
            if (out.info) {
                double *ip = s_info + lane * RT_INFO_SIZE;
                ip[RT_INFO_REWARD_TOTAL] = reward;
                ip[RT_INFO_REWARD_TUMOUR] = (double)r_tumour;
                ip[RT_INFO_REWARD_LUNG] = r_lung;
                ip[RT_INFO_REWARD_DISTANCE] = r_dist;
                ip[RT_INFO_DOSE_TUMOUR] = (double)tsum_f32;
                ip[RT_INFO_DOSE_LUNG] = (double)(float)lung_dose;
                ip[RT_INFO_OVERSHOOT_T0] = se.os_t[0];
                ip[RT_INFO_OVERSHOOT_T0 + 1] = se.os_t[1];
                ip[RT_INFO_OVERSHOOT_T0 + 2] = se.os_t[2];
                ip[RT_INFO_OVERSHOOT_R] = os_r;
                ip[RT_INFO_EPISODE_RETURN] = ep_return;
                ip[RT_INFO_EPISODE_LENGTH] = (double)t;
                ip[RT_INFO_LUNG_COUNT] = (double)lung_count;
                ip[RT_INFO_STEPPED] = 1.0;
                ip[RT_INFO_TUMOUR_ID] = (double)tid;
                ip[RT_INFO_T] = (double)t;
            }
            RT_STAMP3(e, 7);
        }
        __syncwarp();
        if (!kDense && lane < nb) {
            if (out.reward) out.reward[env0 + lane] = s_rew[lane];
            if (out.reward_f32) out.reward_f32[env0 + lane] = (float)s_rew[lane];
            if (out.terminated) out.terminated[env0 + lane] = s_term[lane];
            if (out.truncated) out.truncated[env0 + lane] = 0;
            if (kRoll) R.rewards_row[env0 + lane] = (float)s_rew[lane];    // train.py:154
        }
        if (!kDense && out.info)
            for (int i = lane; i < nb * RT_INFO_SIZE; i += kWarp) out.info[(size_t)env0 * RT_INFO_SIZE + i] = s_info[i];
        return;
    }

    // =====================================================================================================
    if (kPredict && warp == 0) {
        // predictor warp, one thread per env: float32 pose update -> predicted beam (see predict_beam)
        const int e = env0 + lane;
        Beam pb;
        pb.nslab = 0;
        if (lane < kB && e < n_envs) {
            const EnvRec *my = rec + e;
            const int needs_reset = my->needs_reset;
            double p0[3], dn[3];
#pragma unroll
            for (int i = 0; i < 3; i++) { p0[i] = my->pos[i]; dn[i] = my->dn[i]; }
            float2 a01, a23, a45;
            if (kRoll) {
                const float2 *ap = reinterpret_cast<const float2 *>(actions + lane * RT_ACTION_SIZE);
                a01 = ap[0]; a23 = ap[1]; a45 = ap[2];
            } else {
                const float2 *ap = reinterpret_cast<const float2 *>(actions + (size_t)e * RT_ACTION_SIZE);
                a01 = __ldg(ap); a23 = __ldg(ap + 1); a45 = __ldg(ap + 2);
            }
            if (needs_reset == 0) {
                const double gs[3] = {(double)G.g0, (double)G.g1, (double)G.g2};
                const float at[3] = {a01.x, a01.y, a23.x}, ar[3] = {a23.y, a45.x, a45.y};
                double p[3], os;
#pragma unroll
                for (int i = 0; i < 3; i++) p[i] = translate_axis(p0[i], __dmul_rn(__dmul_rn((double)clip1(at[i]), gs[i]), 0.2), gs[i], os);
                pb = predict_beam(G, p, dn, ar);
            }
        }
        if (lane < kB) pred[lane] = pb;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&mbars[1])) : "memory");
        return;
    }

    // =====================================================================================================
    // env warps
    const int le = warp - kEnvWarp0;
    const int env = env0 + le;
    const bool active = env < n_envs;
    EnvShared &se = sh[le];
    Tumour &tm = tum[le];
    uint32_t *tb = tbits[le];
    // what the tumour work needs from the record, read here so that it starts before the scalar warp has published anything
    const bool stepping = active && __ldcg(&rec[active ? env : 0].needs_reset) == 0;
    uint32_t pk4[4];
    if (stepping) env_tumour_fetch<kDense>(T, __ldcg(&rec[env].tumour_id), tm, tb, lane, pk4);
    work_barrier<(kB + 1) * kWarp>();                                      // ---- barrier A
    if (stepping) {
        env_distance(T, se, tm, lane, pk4, [&]() {
            if (kPredict) {
                mbar_wait(smem_u32(&mbars[1]), pred_parity);               // the predictor warp has published its beams
                prefetch_beam(G, pred[le], cells + (size_t)env * G.cstride, lane);
            }
        });
    }
    if (active && lane == 0) RT_STAMP3(env, 2);
    if (kDense) work_barrier<(kB + 1) * kWarp>();                          // ---- barrier 1
    else mbar_wait(smem_u32(&mbars[2]), pred_parity);                      // beams and walks published
    if (stepping && kDense) {
        // Dense mode: publish the beam (distinct voxels + summed weights); rt_dense_kernel streams the whole volume.
        const Beam b = se.beam;
        const float2 *myz = yz[le];
        DenseWork &dw = dense[env];
        int nhit = 0;
        for (int kbase = 0; kbase < b.nslab; kbase += kWarp) {
            const int k = kbase + lane;
            const int kk = k < b.nslab ? k : 0;
            int lin[4], c0, c1, c2;
            float w[4];
            slab_targets_yz(G, b, k, myz[kk], myz[kk > 0 ? kk - 1 : 0], myz[kk + 1 < b.nslab ? kk + 1 : kk], lin, w, c0, c1, c2);
            int mine = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) mine += lin[j] >= 0;
            int incl = mine;
#pragma unroll
            for (int o = 1; o < kWarp; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            int at = nhit + incl - mine;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (lin[j] >= 0 && at < RT_BEAM_CAP) {
                    dw.lin[at] = lin[j];
                    dw.w[at] = w[j];
                    at++;
                }
            nhit += __shfl_sync(kFull, incl, kWarp - 1);
        }
        if (lane == 0) dw.n_hits = nhit < RT_BEAM_CAP ? nhit : RT_BEAM_CAP;
    }
    if (stepping && !kDense)
        deposit_beam<kStageLungs, kClock>(T, se, tm, tb, yz[le], cells + (size_t)env * G.cstride,
                                          cellbuf_all + (size_t)le * (kMaxPass * 4 * kWarp), kStageLungs ? lungs_sm : T.lungs_pad,
                                          smem_u32(&mbars[0]), env, lane);
    work_barrier<(kB + 1) * kWarp>();                                      // ---- barrier 2
}

// Prologue shared by the kernels built on step_block: mbarriers, and the programmatic-dependent-launch hand-over.
template <int kB, bool kDense>
__device__ __forceinline__ void step_prologue(StepShared<kB> &M)
{
    constexpr int kEnvWarp0 = (kB >= 14 && !kDense) ? 1 : 0;
    constexpr int kScalarWarp = kEnvWarp0 + kB;
    constexpr bool kStageLungs = kB >= 14 && !kDense;
    if (!kDense && threadIdx.x == kScalarWarp * kWarp) {
        mbar_init(smem_u32(&M.mbars[0]), 1);
        mbar_init(smem_u32(&M.mbars[1]), 1);
        mbar_init(smem_u32(&M.mbars[2]), 1);
    }
    if (!kDense) __syncthreads();                                 // the env warps wait on the mbarriers before any other barrier
    // Programmatic dependent launch: nothing the previous launch wrote is read before this point; the trigger
    // lets the next launch's blocks be scheduled as soon as ours retire.
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
}

template <int kB, bool kClock, bool kDense>
__global__ void __launch_bounds__((step_block_threads<kB, kDense>()), 28 / kB)
rt_step_kernel(Tables T, Schedule S, EnvRec *rec, uint2 *cells, double *beams, int n_envs,
               const float *__restrict__ actions, StepOut out, DenseWork *dense)
{
    __shared__ StepShared<kB> M;
    extern __shared__ __align__(128) uint32_t dyn_smem[];
    step_prologue<kB, kDense>(M);
    step_block<kB, kClock, kDense, false>(M, dyn_smem, T, S, rec, cells, beams, n_envs, actions, out, dense, blockIdx.x * kB, 0u,
                                          RollStep{nullptr, nullptr});
}

}  // namespace
