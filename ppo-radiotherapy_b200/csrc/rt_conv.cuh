// rt_conv.cuh — first block of FeaturesExtractor3D (networks.py:22-24) as ONE kernel on the tensor cores:
//     Conv3d(4 -> 16, k = 3)  +  bias  +  ReLU  +  MaxPool3d(2, 2, padding = (pd, ph, 0))
// for the voxel observation [n][4][D][H][W] float32 (environment.py:252-257) -> [n][16][Pd][Ph][Pw] bfloat16.
//
// Why a hand-written kernel: cuDNN runs this 4-channel convolution as an sm80 implicit GEMM plus separate
// layout-conversion, bias, ReLU and pooling kernels that stream the 16-channel activation (5.8 MB/sample in
// bf16) through HBM five times: 61 us/sample on B200 (tools/c3d_layers.py, profiles/README.md).  Fused, the
// block reads the observation once (3.2 MB) and writes the pooled result (0.75 MB).
//
// Implicit GEMM without im2col.  Per sample and output depth d the GEMM is
//     Y[f][co] = sum_{kd,kh,kw,c} X[d+kd][h+kh][w+kw][c] * Wt[co][c][kd][kh][kw],      f = h*W + w
// Flattening (h, w) into f makes positions with w >= W-2 or h >= H-2 wrap around; they are computed and discarded
// (7 %).  ReLU and max commute, so pooling max(0, .) values equals the reference's ReLU -> MaxPool; the implicit
// -inf padding of MaxPool3d is "ignore".  The kernel is described where it is defined (rt_conv1_tc_kernel).
// (included at the end of rt_env.cu: the from-env variant of the first block reads the env records, dose volumes
// and tables defined there)
#pragma once
#include <cuda_bf16.h>
#include <math_constants.h>

namespace {

constexpr int kCin = 4, kCout = 16;

struct ConvShape {
    int D, H, W;          // input
    int Do, Ho, Wo;       // conv output = input - 2
    int pd, ph;           // pool padding on depth / height (width padding must be 0)
    int Pd, Ph, Pw;       // pooled output
    int plane_vox;        // voxels per shared-memory plane buffer (H*W rounded up + halo reach)
    int tiles;            // 32-position warp tiles per plane
    uint32_t mW, mPw, mPhPw;   // ceil(2^32 / divisor), 0 for divisor 1 (2^32 does not fit)
    int r_elems;          // bf16 elements of the R buffer (Ho * Pw * 16, rounded up to 64)
};

// n / divisor for the small dividends used here (n < 2^16, divisor < 2^12: n * (magic * d - 2^32) < 2^32)
__device__ __forceinline__ int fastdiv(int n, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)n, magic) : n; }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

// ---------------------------------------------------------------------------------------------------------
// The block on the 5th-generation tensor cores (tcgen05): accumulators in tensor memory, operands read
// from shared memory through matrix descriptors, one thread issues the MMAs.
//
// Two conv planes at a time.  A pool window is the conv-plane pair (d0, d0+1), which reads the input planes
// d0 .. d0+3.  Shared memory holds input planes in PAIRS, channels-last: a voxel is 16 bytes = [4 channels of
// plane z | 4 channels of plane z+1] with z = d0, d0+2, ... (three pair buffers: two in use, one being loaded),
// and the GEMM is  Y[f][o*16 + co],  N = 32 = two output planes x 16 channels, so every operand byte read from
// shared memory serves both planes:
//     output plane d0   takes  kd = 0, 1 from pair (d0, d0+1)   and  kd = 2 from the lower half of pair (d0+2, d0+3)
//     output plane d0+1 takes  kd = 0 from the upper half of the first pair  and  kd = 1, 2 from the second pair
// (the other weight blocks are zero).
//
// A operand = Toeplitz view of a pair buffer, expressed directly in the shared-memory descriptor: K-major, no
// swizzle, where a core matrix is 8 rows x 16 bytes with rows 16 bytes apart.  Row r of a 128-position tile is
// voxel f0 + r, so the stride between 8-row groups (SBO) is 128 bytes and the second 8-wide K chunk of a row is
// simply ANOTHER VOXEL at a fixed distance: the leading-dimension offset (LBO) is W*16 bytes to pair the rows
// kh = 0 and 1 of the stencil, or 16 bytes to pair kw and kw+1.  The 3x3 in-plane stencil is 5 MMAs
// (M = 128, N = 32, K = 16) per pair buffer — (kh 0|1) x kw 0,1,2, then kh 2 x (kw 0|1) and (kw 2|pad) —
// 10 per tile for two output planes, against 2 x 18 with one K chunk per (kd, kh, kw pair).  The core matrices
// overlap in memory; nothing is gathered or copied.
//
// Warp roles (14 warps), coupled by mbarriers only, so each runs ahead as far as its buffers allow:
//   warps 8-11 load input-plane pairs from HBM into the ring of three pair buffers (pair_free -> pair_ready);
//   warps 12-13 issue the MMAs (one thread each, alternate tiles) into a ring of 16 tensor-memory slots of 32 columns
//      (tcgen05.commit -> full[slot]; empty[slot] comes back from the four warps that drained it) — while the
//      drain warps pool one window it is already 16 tiles into the next;
//   warps 0-7 drain: tcgen05.ld gives every thread one position x 32 channels, then bias, ReLU, max over the
//      two planes and over the w pair (neighbouring lane), bf16 row to R; then max over the h pair out of R
//      (8 channels per 16-byte read, bf16x2 max) and the pooled plane goes to HBM.
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "RT_TC_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra RT_TC_DONE;\n"
                 "bra RT_TC_WAIT;\n"
                 "RT_TC_DONE:\n"
                 "}" ::"r"(mbar), "r"(parity) : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading
// and stride byte offsets in 16-byte units, descriptor version 1 (Blackwell) at bit 46.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "setp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
                 "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void commit(uint32_t mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc

constexpr int kTcTileRows = 128;
constexpr int kTcN = 2 * kCout;              // two output planes x 16 channels
constexpr int kTcSlots = 16;                 // tensor-memory ring: 16 x 32 columns
constexpr int kTmemCols = kTcSlots * kTcN;   // 512
constexpr int kTcMmas = 10;                  // per tile: 5 per pair buffer
constexpr int kBMmaBytes = 2 * kTcN * 16;    // one MMA's B operand: 2 K-chunks x 32 rows x 16 bytes
constexpr int kTcDrainWarps = 8, kTcLoadWarps = 8, kTcMmaWarps = 2;
constexpr int kTcThreads = (kTcDrainWarps + kTcLoadWarps + kTcMmaWarps) * 32;

// Stencil position of K chunk h of MMA m (within a pair buffer): (kh, kw); kw == 3 is padding.
__host__ __device__ __forceinline__ void tc_tap(int m, int h, int &kh, int &kw)
{
    if (m < 3) { kh = h; kw = m; }
    else if (m == 3) { kh = 2; kw = h; }
    else { kh = 2; kw = 2 + h; }
}

// B operand of the tcgen05 path as it lies in shared memory: [mma = b*5 + m][K chunk h][n = o*16 + co][8 slots],
// slot = u*4 + c: channel c of the lower (u = 0) / upper (u = 1) plane of pair buffer b, i.e. input plane
// d0 + 2b + u, which output plane d0 + o sees as kd = 2b + u - o.
__global__ void rt_conv_prepare_tc_kernel(const float *__restrict__ weight, __nv_bfloat16 *__restrict__ bop)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kTcMmas * 2 * kTcN * 8) return;
    const int slot = idx & 7, n = (idx >> 3) & 31, h = (idx >> 8) & 1, mma = idx >> 9;
    const int b = mma / 5, m = mma % 5;
    const int u = slot >> 2, c = slot & 3, o = n >> 4, co = n & 15;
    const int kd = 2 * b + u - o;
    int kh, kw;
    tc_tap(m, h, kh, kw);
    const float v = (kd >= 0 && kd <= 2 && kw < 3) ? weight[(((co * kCin + c) * 3 + kd) * 3 + kh) * 3 + kw] : 0.0f;
    bop[idx] = __float2bfloat16(v);
}

// Source of the first block's input when it is not a materialised observation tensor: the live env state.  The
// loader warps then generate the four observation planes of environment.py:245-257 voxel by voxel — lungs bit,
// tumour bit, dose (cells of another generation read as zero), clip(current beam + horizontal beam, 0, 1) — exactly as
// rt_volumes_kernel would have written them, and the 3.2 MB float32 observation never exists.
struct EnvSource {
    Tables T;
    const EnvRec *rec;
    const uint2 *cells;      // sparse-mode dose cells {dose, generation}
    int first;
};

template <bool kGroupedOut, bool kFromEnv>
__global__ void __launch_bounds__(kTcThreads, 1) rt_conv1_tc_kernel(ConvShape S, int n_samples, int chunks,
                                                                    int pooled_per_chunk, const float *__restrict__ x,
                                                                    const uint4 *__restrict__ bop,
                                                                    const float *__restrict__ bias,
                                                                    __nv_bfloat16 *__restrict__ out, EnvSource E)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4 *pairs = reinterpret_cast<uint4 *>(smem_raw);                               // 3 x [plane_vox] x 16 B
    __nv_bfloat16 *R = reinterpret_cast<__nv_bfloat16 *>(pairs + 3 * (size_t)S.plane_vox);    // [Ho][Pw][16]
    uint4 *bsm = reinterpret_cast<uint4 *>(R + (size_t)S.r_elems);                    // [10][1 KB]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(bsm + kTcMmas * kBMmaBytes / 16);
    unsigned long long *empty = full + kTcSlots;
    unsigned long long *pair_ready = empty + kTcSlots;             // [3] loader warps -> MMA warp
    unsigned long long *pair_free = pair_ready + 3;                // [3] MMA warp (tcgen05.commit) -> loader warps
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pair_free + 3);
    int *hkeys = reinterpret_cast<int *>(tmem_slot + 4);           // kFromEnv: voxel -> beam-view weight (kHashSlots)
    float *hvals = reinterpret_cast<float *>(hkeys + kHashSlots);
    RayWork *view = reinterpret_cast<RayWork *>(hvals + kHashSlots);   // [2]
    uint32_t *hit_sec = reinterpret_cast<uint32_t *>(view + 2);    // kFromEnv: 1 bit per dose sector touched by a view beam
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sample = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const bool is_mma = warp >= kTcDrainWarps + kTcLoadWarps;
    const bool is_loader = warp >= kTcDrainWarps && !is_mma;
    const int mma_rank = warp - (kTcDrainWarps + kTcLoadWarps);   // MMA warp r issues the tiles t = r (mod kTcMmaWarps)

    // pooled planes [p_lo, p_hi) of this block; pooled plane p is the max over conv planes 2p - pd and 2p - pd + 1
    const int p_lo = chunk * pooled_per_chunk;
    const int p_hi = min(S.Pd, p_lo + pooled_per_chunk);
    const bool has_work = sample < n_samples && p_lo < p_hi;                          // block-uniform
    const int iters = has_work ? p_hi - p_lo : 0;
    const int z0 = 2 * p_lo - S.pd;                    // first input plane of pair 0 (may be -1)

    // one-time set-up: tensor memory, mbarriers, weights, zeroed pair buffers
    if (is_mma && mma_rank == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < kTcSlots; i++) {
            tc::mbar_init(tc::smem_u32(&full[i]), 1);
            tc::mbar_init(tc::smem_u32(&empty[i]), 4);             // the four warps (lane quarters) that drain a slot
        }
        for (int i = 0; i < 3; i++) {
            tc::mbar_init(tc::smem_u32(&pair_ready[i]), kTcLoadWarps);
            tc::mbar_init(tc::smem_u32(&pair_free[i]), kTcMmaWarps);
        }
        tc::fence_mbar_init();
    }
    for (int i = tid; i < kTcMmas * kBMmaBytes / 16; i += kTcThreads) bsm[i] = __ldg(bop + i);
    // the halo tail of a buffer is read (times a zero weight, or for discarded rows): it must hold finite values
    for (int i = tid; i < 3 * S.plane_vox; i += kTcThreads) pairs[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int HW = S.H * S.W;
    const uint32_t pairs_addr = tc::smem_u32(pairs);

    if (is_loader && kFromEnv) {
        // ---- the same pair buffers, generated from the env state (see EnvSource)
        const int t0 = tid - kTcDrainWarps * 32, nt = kTcLoadWarps * 32, lw = warp - kTcDrainWarps;
        const Grid &G = E.T.G;
        const int env = E.first + sample;
        if (has_work) {
            const EnvRec *r = E.rec + env;
            const int tumour_id = r->tumour_id;
            const uint32_t gen = r->gen;
            // view planes (environment.py:246-250): beam along the current direction + beam along (1, 0, 0)
            for (int i = t0; i < kHashSlots; i += nt) { hkeys[i] = -1; hvals[i] = 0.0f; }
            for (int i = t0; i < (G.vstride / 8 + 31) / 32; i += nt) hit_sec[i] = 0u;     // one bit per 8 voxels
            asm volatile("bar.sync 2, %0;" ::"n"(kTcLoadWarps * 32) : "memory");
            if (lw < 2) {
                const double pos[3] = {r->pos[0], r->pos[1], r->pos[2]};
                const double dir[3] = {lw == 0 ? r->dir[0] : 1.0, lw == 0 ? r->dir[1] : 0.0, lw == 0 ? r->dir[2] : 0.0};
                const Beam b = ray_prepare(G, pos, dir, lane, view[lw]);
                for (int kbase = 0; kbase < b.nslab; kbase += kWarp) {
                    int lin[4], c0, c1, c2;
                    float w[4];
                    slab_targets(G, b, view[lw].ys, view[lw].zs, kbase + lane, lin, w, c0, c1, c2);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (lin[q] >= 0) {
                            hash_add(hkeys, hvals, lin[q], w[q]);
                            atomicOr(hit_sec + (lin[q] >> 8), 1u << ((lin[q] >> 3) & 31));
                        }
                }
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kTcLoadWarps * 32) : "memory");
            const Tumour tm = E.T.tumours[tumour_id];
            const uint32_t *tb = E.T.tumour_pbits + (size_t)tumour_id * E.T.pbits_words;
            const int pd1 = tm.dim[1] + 2, pd2 = tm.dim[2] + 2;
            const uint2 *vol = E.cells + (size_t)env * G.cstride;
            // Work item = 8 consecutive voxels of the volume (64 bytes of cells) of one plane of the pair: four 16-byte
            // cell loads (a cell of another generation reads as zero), one lungs word, one view-hit bit; the tumour and
            // view tests only run for the few groups that can contain such voxels.
            for (int j = 0; j <= iters; j++) {
                if (j >= 3) tc::mbar_wait(tc::smem_u32(&pair_free[j % 3]), (uint32_t)((j / 3 - 1) & 1));   // MMAs of iteration j-3 done
                uint2 *dst = reinterpret_cast<uint2 *>(pairs + (size_t)(j % 3) * S.plane_vox);       // [voxel][lower | upper plane]
#pragma unroll 1
                for (int u = 0; u < 2; u++) {
                    const int z = z0 + 2 * j + u;
                    if (z < 0 || z >= S.D) {               // plane outside the volume: its half of every voxel is zero
                        for (int v = t0; v < HW; v += nt) dst[2 * v + u] = make_uint2(0u, 0u);
                        continue;
                    }
                    const int lin_lo = z * HW, lin_hi = lin_lo + HW;
                    const bool z_in_tumour = (unsigned)(z - tm.lo[0]) < (unsigned)tm.dim[0];
                    for (int sec = (lin_lo >> 3) + t0; sec <= (lin_hi - 1) >> 3; sec += nt) {
                        const int l0 = sec << 3;
                        float d[8];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            // voxels (l0 + 2i, l0 + 2i + 1) share a row (W is even) and a brick of the cell layout
                            const int lp = min(l0 + 2 * i, G.nvox - 2);
                            const uint4 c = __ldg(reinterpret_cast<const uint4 *>(vol + cell_index_lin(G, lp)));
                            d[2 * i] = c.y == gen ? __uint_as_float(c.x) : 0.0f;
                            d[2 * i + 1] = c.w == gen ? __uint_as_float(c.z) : 0.0f;
                        }
                        const uint32_t lung8 = (__ldg(E.T.lungs_bits + (l0 >> 5)) >> (l0 & 31)) & 255u;
                        const bool view_hit = (hit_sec[sec >> 5] >> (sec & 31)) & 1u;
                        int v = l0 - lin_lo;                                   // may be negative for the plane's first sector
                        int h = v >= 0 ? fastdiv(v, S.mW) : 0, w = v - h * S.W;
#pragma unroll
                        for (int i = 0; i < 8; i++, v++, w++) {
                            if (w == S.W) { w = 0; h++; }
                            if (v < 0 || v >= HW) continue;                    // voxel of the neighbouring plane
                            float tum = 0.0f, vw = 0.0f;
                            if (z_in_tumour) {
                                const int tj = h - tm.lo[1], tk = w - tm.lo[2];
                                if ((unsigned)tj < (unsigned)tm.dim[1] && (unsigned)tk < (unsigned)tm.dim[2]) {
                                    const int bit = ((z - tm.lo[0]) * pd1 + tj + 1) * pd2 + tk + 1;
                                    tum = (__ldg(tb + (bit >> 5)) >> (bit & 31)) & 1u ? 1.0f : 0.0f;
                                }
                            }
                            if (view_hit) {
                                const int lin = l0 + i;
                                for (int slot = hash_slot(lin);; slot = (slot + 1) & (kHashSlots - 1)) {
                                    const int key = hkeys[slot];
                                    if (key == lin) { vw = hvals[slot]; break; }
                                    if (key == -1) break;
                                }
                            }
                            dst[2 * v + u] = make_uint2(pack_bf16((lung8 >> i) & 1u ? 1.0f : 0.0f, tum),
                                                        pack_bf16(fminf(fmaxf(d[i], 0.0f), 1.0f), fminf(fmaxf(vw, 0.0f), 1.0f)));
                        }
                    }
                }
                tc::fence_proxy_async();               // this lane's writes are visible to the tensor core
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&pair_ready[j % 3]));
            }
        }
    } else if (is_loader) {
        // ---- input-plane pairs: pair j = planes (z0 + 2j, z0 + 2j + 1) -> buffer j % 3, channels-last bf16;
        //      planes outside [0, D) are zero.  Pairs 0 .. iters are needed (iteration i reads pairs i and i+1).
        const float *xs = x + (size_t)sample * kCin * S.D * HW;
        const size_t cstride = (size_t)S.D * HW;
        const int t0 = tid - kTcDrainWarps * 32, nt = kTcLoadWarps * 32;
        for (int j = 0; j <= iters && has_work; j++) {
            if (j >= 3) tc::mbar_wait(tc::smem_u32(&pair_free[j % 3]), (uint32_t)((j / 3 - 1) & 1));   // MMAs of iteration j-3 done
            const int za = z0 + 2 * j, zb = za + 1;
            const bool ina = za >= 0 && za < S.D, inb = zb >= 0 && zb < S.D;
            uint4 *dst = pairs + (size_t)(j % 3) * S.plane_vox;
            const float *sa = xs + (size_t)(ina ? za : 0) * HW;
            const float *sb = xs + (size_t)(inb ? zb : 0) * HW;
            if ((HW & 1) == 0) {
                // two voxels per thread and iteration: 8-byte loads (HW even keeps every plane 8-byte aligned),
                // sixteen of them in flight per thread
                const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int v = 2 * t0; v < HW; v += 2 * nt) {
                    float2 a0 = z2, a1 = z2, a2 = z2, a3 = z2, b0 = z2, b1 = z2, b2 = z2, b3 = z2;
                    if (ina) {
                        a0 = __ldg(reinterpret_cast<const float2 *>(sa + v));
                        a1 = __ldg(reinterpret_cast<const float2 *>(sa + cstride + v));
                        a2 = __ldg(reinterpret_cast<const float2 *>(sa + 2 * cstride + v));
                        a3 = __ldg(reinterpret_cast<const float2 *>(sa + 3 * cstride + v));
                    }
                    if (inb) {
                        b0 = __ldg(reinterpret_cast<const float2 *>(sb + v));
                        b1 = __ldg(reinterpret_cast<const float2 *>(sb + cstride + v));
                        b2 = __ldg(reinterpret_cast<const float2 *>(sb + 2 * cstride + v));
                        b3 = __ldg(reinterpret_cast<const float2 *>(sb + 3 * cstride + v));
                    }
                    dst[v] = make_uint4(pack_bf16(a0.x, a1.x), pack_bf16(a2.x, a3.x), pack_bf16(b0.x, b1.x), pack_bf16(b2.x, b3.x));
                    dst[v + 1] = make_uint4(pack_bf16(a0.y, a1.y), pack_bf16(a2.y, a3.y), pack_bf16(b0.y, b1.y), pack_bf16(b2.y, b3.y));
                }
            } else {
#pragma unroll 2
                for (int v = t0; v < HW; v += nt) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
                    if (ina) { a0 = __ldg(sa + v); a1 = __ldg(sa + cstride + v); a2 = __ldg(sa + 2 * cstride + v); a3 = __ldg(sa + 3 * cstride + v); }
                    if (inb) { b0 = __ldg(sb + v); b1 = __ldg(sb + cstride + v); b2 = __ldg(sb + 2 * cstride + v); b3 = __ldg(sb + 3 * cstride + v); }
                    dst[v] = make_uint4(pack_bf16(a0, a1), pack_bf16(a2, a3), pack_bf16(b0, b1), pack_bf16(b2, b3));
                }
            }
            tc::fence_proxy_async();                   // this lane's writes are visible to the tensor core
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&pair_ready[j % 3]));
        }
    } else if (is_mma) {
        // ---- MMA issue
        const uint32_t bsm_addr = tc::smem_u32(bsm);
        for (int i = 0; i < iters; i++) {
            tc::mbar_wait(tc::smem_u32(&pair_ready[i % 3]), (uint32_t)((i / 3) & 1));
            tc::mbar_wait(tc::smem_u32(&pair_ready[(i + 1) % 3]), (uint32_t)(((i + 1) / 3) & 1));
            tc::fence_after_sync();
            const uint32_t buf0 = pairs_addr + (uint32_t)((i % 3) * S.plane_vox * 16);
            const uint32_t buf1 = pairs_addr + (uint32_t)(((i + 1) % 3) * S.plane_vox * 16);
            for (int t = mma_rank; t < S.tiles; t += kTcMmaWarps) {
                const int g = i * S.tiles + t, slot = g % kTcSlots, use = g / kTcSlots;
                tc::mbar_wait(tc::smem_u32(&empty[slot]), (uint32_t)((use & 1) ^ 1));
                tc::fence_after_sync();
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < kTcMmas; q++) {
                        const int b = q / 5, m = q % 5;
                        int kh, kw;
                        tc_tap(m, 0, kh, kw);
                        const uint32_t a_addr = (b ? buf1 : buf0) + (uint32_t)((t * kTcTileRows + kh * S.W + kw) * 16);
                        const uint32_t lbo = m < 3 ? (uint32_t)(S.W * 16) : 16u;
                        tc::mma_f16(tmem_base + (uint32_t)(slot * kTcN), tc::smem_desc(a_addr, lbo, 128),
                                    tc::smem_desc(bsm_addr + q * kBMmaBytes, kTcN * 16, 128), q > 0 ? 1u : 0u);
                    }
                    tc::commit(tc::smem_u32(&full[slot]));
                }
                __syncwarp();
            }
            if (lane == 0) tc::commit(tc::smem_u32(&pair_free[i % 3]));     // buffer i % 3 may be reloaded once these MMAs are done
            __syncwarp();
        }
    } else {
        // ---- drain + pool
        float bv[kCout];
#pragma unroll
        for (int c = 0; c < kCout; c++) bv[c] = __ldg(bias + c);
        const int grp = warp >> 2, quarter = warp & 3;
        const int PhPw = S.Ph * S.Pw;
        for (int i = 0; i < iters; i++) {
            const int d0 = z0 + 2 * i;                 // conv planes d0 (may be -1) and d0 + 1 (may be Do)
            const bool ok_a = d0 >= 0, ok_b = d0 + 1 < S.Do;
            for (int t = grp; t < S.tiles; t += 2) {
                const int g = i * S.tiles + t, slot = g % kTcSlots, use = g / kTcSlots;
                tc::mbar_wait(tc::smem_u32(&full[slot]), (uint32_t)(use & 1));
                tc::fence_after_sync();
                uint32_t acc[kTcN];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * kTcN);
                tc::tmem_ld16(taddr, acc);
                tc::tmem_ld16(taddr + 16, acc + 16);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[slot]));
                const int f = t * kTcTileRows + quarter * 32 + lane;
                const int h = fastdiv(f, S.mW), w = f - h * S.W;
                // max over the depth pair first (ReLU and + bias are monotone), then bias, ReLU, bf16, and the w pair
                // (W and the tile base are even: the odd position is the next lane) on packed bf16x2
                uint32_t pk[kCout / 2];
#pragma unroll
                for (int c = 0; c < kCout; c += 2) {
                    float v0 = ok_a ? __uint_as_float(acc[c]) : -CUDART_INF_F, v1 = ok_a ? __uint_as_float(acc[c + 1]) : -CUDART_INF_F;
                    if (ok_b) { v0 = fmaxf(v0, __uint_as_float(acc[kCout + c])); v1 = fmaxf(v1, __uint_as_float(acc[kCout + c + 1])); }
                    uint32_t mine = pack_bf16(fmaxf(v0 + bv[c], 0.0f), fmaxf(v1 + bv[c + 1], 0.0f));
                    const uint32_t other = __shfl_down_sync(0xffffffffu, mine, 1);
                    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&mine), *reinterpret_cast<const __nv_bfloat162 *>(&other));
                    pk[c / 2] = *reinterpret_cast<const uint32_t *>(&r);
                }
                if ((w & 1) == 0 && w < S.Wo && h < S.Ho) {
                    uint4 *dst = reinterpret_cast<uint4 *>(R + ((size_t)h * S.Pw + (w >> 1)) * kCout);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTcDrainWarps * 32) : "memory");     // R holds the window, pooled along d and w
            // max over the h pair -> pooled plane p; an item is (8 channels, one pooled position)
            const int p = p_lo + i;
            __nv_bfloat16 *op = out + ((size_t)sample * kCout * S.Pd + p) * PhPw;
            for (int item = tid; item < 2 * PhPw; item += kTcDrainWarps * 32) {
                const int half = item >= PhPw ? 1 : 0, pos = item - half * PhPw;
                const int py = fastdiv(pos, S.mPw), px = pos - py * S.Pw;
                const int h0 = 2 * py - S.ph;
                uint4 m = make_uint4(0u, 0u, 0u, 0u);  // every candidate is >= 0 after ReLU
                if (h0 >= 0 && h0 < S.Ho) m = *reinterpret_cast<const uint4 *>(R + ((size_t)h0 * S.Pw + px) * kCout + half * 8);
                if (h0 + 1 >= 0 && h0 + 1 < S.Ho) {
                    const uint4 o = *reinterpret_cast<const uint4 *>(R + ((size_t)(h0 + 1) * S.Pw + px) * kCout + half * 8);
                    auto mx = [](uint32_t a, uint32_t b) {
                        const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
                        return *reinterpret_cast<const uint32_t *>(&r);
                    };
                    m = make_uint4(mx(m.x, o.x), mx(m.y, o.y), mx(m.z, o.z), mx(m.w, o.w));
                }
                if (kGroupedOut) {
                    // [n][group][Pd][Ph*Pw][8 channels]: the layout rt_conv2_tc_kernel bulk-copies into shared memory
                    uint4 *og = reinterpret_cast<uint4 *>(out) + (((size_t)sample * 2 + half) * S.Pd + p) * PhPw + pos;
                    *og = m;
                } else {
                    const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
                    unsigned short *o16 = reinterpret_cast<unsigned short *>(op) + (size_t)(half * 8) * S.Pd * PhPw + pos;
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        o16[(size_t)c * S.Pd * PhPw] = (unsigned short)(c & 1 ? mw[c >> 1] >> 16 : mw[c >> 1] & 0xffffu);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTcDrainWarps * 32) : "memory");     // R is free again
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (is_mma && mma_rank == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}


// ---------------------------------------------------------------------------------------------------------
// Second block of FeaturesExtractor3D (networks.py:25-27): Conv3d(16 -> 16, k = 3, groups = 2) + bias + ReLU +
// MaxPool3d(2, 2), same scheme as rt_conv1_tc_kernel.  The input is the first block's output in the grouped
// channels-last layout [n][group][D][H*W][8 channels] bf16: a voxel of one group is exactly one 16-byte K chunk
// (all 8 slots are real channels), a (plane, group) sub-plane is one contiguous run of H*W*16 bytes, and a
// loader thread brings it to shared memory with a single cp.async.bulk — no conversion, no loader warps.
// A pool window (conv planes 2p, 2p+1) reads input planes 2p .. 2p+3 = two pair buffers x 2 planes x 2 groups =
// 8 sub-planes x 5 MMAs (M 128, N 32, K 16) per tile; N = 32 = [plane 2p | plane 2p+1] x [group 0 | group 1] x 8
// channels, the weight blocks of the other group and of kd outside 0..2 are zero.
constexpr int kC2Mmas = 40;
constexpr int kC2Threads = (kTcDrainWarps + 1 + kTcMmaWarps) * 32;     // 8 drain warps, 1 loader warp, 2 MMA warps

struct Conv2Shape {
    int D, H, W;          // input planes
    int Do, Ho, Wo;       // conv output = input - 2
    int Pd, Ph, Pw;       // pooled output = floor(conv / 2)
    int plane_vox;        // voxels per shared-memory sub-plane (tiles * 128 + halo reach)
    int tiles;            // 128-position tiles per plane
    int r_elems;
    uint32_t mW, mPw;
};

// B operand: [mma = ((b*2 + u)*2 + g)*5 + m][K chunk h][n = o*16 + g'*8 + co][8 input channels of group g]
__global__ void rt_conv2_prepare_kernel(const float *__restrict__ weight, __nv_bfloat16 *__restrict__ bop)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kC2Mmas * 2 * kTcN * 8) return;
    const int c = idx & 7, n = (idx >> 3) & 31, h = (idx >> 8) & 1, mma = idx >> 9;
    const int m = mma % 5, src = mma / 5, g = src & 1, zi = src >> 1;     // zi = 2b + u: input plane d0 + zi
    const int o = n >> 4, g2 = (n >> 3) & 1, co = n & 7;
    const int kd = zi - o;
    int kh, kw;
    tc_tap(m, h, kh, kw);
    const float v = (g2 == g && kd >= 0 && kd <= 2 && kw < 3) ? weight[((((g * 8 + co) * 8 + c) * 3 + kd) * 3 + kh) * 3 + kw] : 0.0f;
    bop[idx] = __float2bfloat16(v);
}

__global__ void __launch_bounds__(kC2Threads, 1) rt_conv2_tc_kernel(Conv2Shape S, int n_samples, int chunks,
                                                                    int pooled_per_chunk, const uint4 *__restrict__ x,
                                                                    const uint4 *__restrict__ bop,
                                                                    const float *__restrict__ bias,
                                                                    __nv_bfloat16 *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4 *pairs = reinterpret_cast<uint4 *>(smem_raw);                               // 3 x [2 planes][2 groups][plane_vox] x 16 B
    const int sub = S.plane_vox;                       // voxels per sub-plane
    __nv_bfloat16 *R = reinterpret_cast<__nv_bfloat16 *>(pairs + 3 * 4 * (size_t)sub);        // [Ho][Pw][16]
    uint4 *bsm = reinterpret_cast<uint4 *>(R + (size_t)S.r_elems);                    // [40][1 KB]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(bsm + kC2Mmas * kBMmaBytes / 16);
    unsigned long long *empty = full + kTcSlots;
    unsigned long long *pair_ready = empty + kTcSlots;
    unsigned long long *pair_free = pair_ready + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pair_free + 3);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sample = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const bool is_mma = warp >= kTcDrainWarps + 1;
    const bool is_loader = warp == kTcDrainWarps;
    const int mma_rank = warp - (kTcDrainWarps + 1);

    const int p_lo = chunk * pooled_per_chunk;
    const int p_hi = min(S.Pd, p_lo + pooled_per_chunk);
    const bool has_work = sample < n_samples && p_lo < p_hi;                          // block-uniform
    const int iters = has_work ? p_hi - p_lo : 0;
    const int z0 = 2 * p_lo;                           // first input plane of pair 0

    if (is_mma && mma_rank == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < kTcSlots; i++) {
            tc::mbar_init(tc::smem_u32(&full[i]), 1);
            tc::mbar_init(tc::smem_u32(&empty[i]), 4);
        }
        for (int i = 0; i < 3; i++) {
            tc::mbar_init(tc::smem_u32(&pair_ready[i]), 1);                // one arrive.expect_tx, then the bulk copies' bytes
            tc::mbar_init(tc::smem_u32(&pair_free[i]), kTcMmaWarps);
        }
        tc::fence_mbar_init();
    }
    for (int i = tid; i < kC2Mmas * kBMmaBytes / 16; i += kC2Threads) bsm[i] = __ldg(bop + i);
    // the halo tail of a sub-plane is read (times a zero weight, or for discarded rows): it must hold finite values
    for (int i = tid; i < 3 * 4 * sub; i += kC2Threads) pairs[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int HW = S.H * S.W;
    const uint32_t pairs_addr = tc::smem_u32(pairs);

    if (is_loader) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(HW * 16);
            for (int j = 0; j <= iters && has_work; j++) {
                if (j >= 3) tc::mbar_wait(tc::smem_u32(&pair_free[j % 3]), (uint32_t)((j / 3 - 1) & 1));
                const uint32_t bar = tc::smem_u32(&pair_ready[j % 3]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * bytes) : "memory");
#pragma unroll
                for (int s4 = 0; s4 < 4; s4++) {
                    const int u = s4 >> 1, g = s4 & 1, z = z0 + 2 * j + u;
                    const uint4 *src = x + (((size_t)sample * 2 + g) * S.D + z) * HW;
                    const uint32_t dst = pairs_addr + (uint32_t)((((j % 3) * 4 + s4) * sub) * 16);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
                }
            }
        }
    } else if (is_mma) {
        const uint32_t bsm_addr = tc::smem_u32(bsm);
        for (int i = 0; i < iters; i++) {
            tc::mbar_wait(tc::smem_u32(&pair_ready[i % 3]), (uint32_t)((i / 3) & 1));
            tc::mbar_wait(tc::smem_u32(&pair_ready[(i + 1) % 3]), (uint32_t)(((i + 1) / 3) & 1));
            tc::fence_after_sync();
            const uint32_t buf0 = pairs_addr + (uint32_t)(((i % 3) * 4 * sub) * 16);
            const uint32_t buf1 = pairs_addr + (uint32_t)((((i + 1) % 3) * 4 * sub) * 16);
            for (int t = mma_rank; t < S.tiles; t += kTcMmaWarps) {
                const int g = i * S.tiles + t, slot = g % kTcSlots, use = g / kTcSlots;
                tc::mbar_wait(tc::smem_u32(&empty[slot]), (uint32_t)((use & 1) ^ 1));
                tc::fence_after_sync();
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < kC2Mmas; q++) {
                        const int m = q % 5, src = q / 5, s4 = src & 3, b = src >> 2;     // src = (b*2 + u)*2 + g
                        int kh, kw;
                        tc_tap(m, 0, kh, kw);
                        const uint32_t a_addr = (b ? buf1 : buf0) + (uint32_t)((s4 * sub + t * kTcTileRows + kh * S.W + kw) * 16);
                        const uint32_t lbo = m < 3 ? (uint32_t)(S.W * 16) : 16u;
                        tc::mma_f16(tmem_base + (uint32_t)(slot * kTcN), tc::smem_desc(a_addr, lbo, 128),
                                    tc::smem_desc(bsm_addr + q * kBMmaBytes, kTcN * 16, 128), q > 0 ? 1u : 0u);
                    }
                    tc::commit(tc::smem_u32(&full[slot]));
                }
                __syncwarp();
            }
            if (lane == 0) tc::commit(tc::smem_u32(&pair_free[i % 3]));
            __syncwarp();
        }
    } else {
        float bv[kCout];
#pragma unroll
        for (int c = 0; c < kCout; c++) bv[c] = __ldg(bias + c);
        const int grp = warp >> 2, quarter = warp & 3;
        const int PhPw = S.Ph * S.Pw;
        for (int i = 0; i < iters; i++) {
            for (int t = grp; t < S.tiles; t += 2) {
                const int g = i * S.tiles + t, slot = g % kTcSlots, use = g / kTcSlots;
                tc::mbar_wait(tc::smem_u32(&full[slot]), (uint32_t)(use & 1));
                tc::fence_after_sync();
                uint32_t acc[kTcN];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * kTcN);
                tc::tmem_ld16(taddr, acc);
                tc::tmem_ld16(taddr + 16, acc + 16);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[slot]));
                const int f = t * kTcTileRows + quarter * 32 + lane;
                const int h = fastdiv(f, S.mW), w = f - h * S.W;
                uint32_t pk[kCout / 2];
#pragma unroll
                for (int c = 0; c < kCout; c += 2) {
                    const float v0 = fmaxf(__uint_as_float(acc[c]), __uint_as_float(acc[kCout + c]));          // depth pair
                    const float v1 = fmaxf(__uint_as_float(acc[c + 1]), __uint_as_float(acc[kCout + c + 1]));
                    uint32_t mine = pack_bf16(fmaxf(v0 + bv[c], 0.0f), fmaxf(v1 + bv[c + 1], 0.0f));
                    const uint32_t other = __shfl_down_sync(0xffffffffu, mine, 1);                             // w pair
                    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&mine), *reinterpret_cast<const __nv_bfloat162 *>(&other));
                    pk[c / 2] = *reinterpret_cast<const uint32_t *>(&r);
                }
                if ((w & 1) == 0 && w < S.Wo && h < S.Ho) {
                    uint4 *dst = reinterpret_cast<uint4 *>(R + ((size_t)h * S.Pw + (w >> 1)) * kCout);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTcDrainWarps * 32) : "memory");
            const int p = p_lo + i;
            unsigned short *op = reinterpret_cast<unsigned short *>(out) + ((size_t)sample * kCout * S.Pd + p) * PhPw;
            for (int item = tid; item < 2 * PhPw; item += kTcDrainWarps * 32) {
                const int half = item >= PhPw ? 1 : 0, pos = item - half * PhPw;
                const int py = fastdiv(pos, S.mPw), px = pos - py * S.Pw;
                const uint4 a = *reinterpret_cast<const uint4 *>(R + ((size_t)(2 * py) * S.Pw + px) * kCout + half * 8);
                const uint4 o = *reinterpret_cast<const uint4 *>(R + ((size_t)(2 * py + 1) * S.Pw + px) * kCout + half * 8);
                auto mx = [](uint32_t u, uint32_t v) {
                    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&u), *reinterpret_cast<const __nv_bfloat162 *>(&v));
                    return *reinterpret_cast<const uint32_t *>(&r);
                };
                const uint32_t mw[4] = {mx(a.x, o.x), mx(a.y, o.y), mx(a.z, o.z), mx(a.w, o.w)};
                unsigned short *o16 = op + (size_t)(half * 8) * S.Pd * PhPw + pos;
#pragma unroll
                for (int c = 0; c < 8; c++)
                    o16[(size_t)c * S.Pd * PhPw] = (unsigned short)(c & 1 ? mw[c >> 1] >> 16 : mw[c >> 1] & 0xffffu);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTcDrainWarps * 32) : "memory");
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (is_mma && mma_rank == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}


// ---------------------------------------------------------------------------------------------------------
// Tail of FeaturesExtractor3D (networks.py:28-45): Conv3d(16 -> 16, k = 3, groups = 4) + bias + ReLU +
// MaxPool3d(2, 2) + Flatten + Linear(n_flat -> F) + bias + ReLU, one block per sample.  2.3 MMAC per sample: too
// small for the tensor pipe to matter; what matters is that the sample's activation (69 KB bf16) is read from HBM
// once and nothing in between goes back.  A thread owns one pooled cell: the 2x2x2 conv outputs under it share a
// 4x4x4 input window per input channel, so every window value and every weight is read from shared memory once
// per cell (364 reads for 864 FMAs); then the 2016 features stay in shared memory for the linear layer.
constexpr int kTailThreads = 256;
constexpr int kTailMaxF = 256;

__global__ void __launch_bounds__(kTailThreads) rt_c3d_tail_kernel(const __nv_bfloat16 *__restrict__ x, int D, int H, int W,
                                                                   const float *__restrict__ w3, const float *__restrict__ b3,
                                                                   const float *__restrict__ wl, const float *__restrict__ bl,
                                                                   int F, float *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int DHW = D * H * W;
    __nv_bfloat16 *xs = reinterpret_cast<__nv_bfloat16 *>(smem_raw);                   // [16][D][H][W]
    float *ws = reinterpret_cast<float *>(xs + ((size_t)16 * DHW + 7) / 8 * 8);         // [16][4][27]
    float *feat = ws + 16 * 4 * 27;                                                     // [16][Pd][Ph][Pw]
    const int Pd = (D - 2) / 2, Ph = (H - 2) / 2, Pw = (W - 2) / 2;
    const int cells = Pd * Ph * Pw, n_flat = 16 * cells;
    const int tid = threadIdx.x;
    const __nv_bfloat16 *xg = x + (size_t)blockIdx.x * 16 * DHW;
    if ((16 * DHW) % 8 == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(xg);
        uint4 *dst = reinterpret_cast<uint4 *>(xs);
        for (int i = tid; i < 16 * DHW / 8; i += kTailThreads) dst[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < 16 * DHW; i += kTailThreads) xs[i] = xg[i];
    }
    for (int i = tid; i < 16 * 4 * 27; i += kTailThreads) ws[i] = __ldg(w3 + i);
    __syncthreads();

    for (int cell = tid; cell < n_flat; cell += kTailThreads) {
        const int c = cell / cells, r = cell - c * cells;
        const int pd = r / (Ph * Pw), r2 = r - pd * (Ph * Pw), ph = r2 / Pw, pw = r2 - ph * Pw;
        const int g = c >> 2;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = 0.0f;
        for (int ci = 0; ci < 4; ci++) {
            const __nv_bfloat16 *xc = xs + (size_t)(4 * g + ci) * DHW + ((2 * pd) * H + 2 * ph) * W + 2 * pw;
            const float *wc = ws + (c * 4 + ci) * 27;
            float win[4][4][4];                          // input window under the 2x2x2 conv outputs of this cell
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++)
#pragma unroll
                    for (int e = 0; e < 4; e++) win[a][b][e] = __bfloat162float(xc[(a * H + b) * W + e]);
#pragma unroll
            for (int kd = 0; kd < 3; kd++)
#pragma unroll
                for (int kh = 0; kh < 3; kh++)
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) {
                        const float wv = wc[(kd * 3 + kh) * 3 + kw];
#pragma unroll
                        for (int o = 0; o < 8; o++)
                            acc[o] = fmaf(win[(o >> 2) + kd][((o >> 1) & 1) + kh][(o & 1) + kw], wv, acc[o]);
                    }
        }
        float m = acc[0];
#pragma unroll
        for (int o = 1; o < 8; o++) m = fmaxf(m, acc[o]);
        feat[cell] = fmaxf(m + __ldg(b3 + c), 0.0f);     // max, + bias and ReLU commute
    }
    __syncthreads();

    // Linear + ReLU: warp w computes outputs w, w + 8, ...; lanes stride the features (coalesced weight rows)
    const int lane = tid & 31, warp = tid >> 5;
    for (int f = warp; f < F; f += kTailThreads / 32) {
        const float *wr = wl + (size_t)f * n_flat;
        float sum = 0.0f;
        for (int i = lane; i < n_flat; i += 32) sum = fmaf(feat[i], __ldg(wr + i), sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) out[(size_t)blockIdx.x * F + f] = fmaxf(sum + __ldg(bl + f), 0.0f);
    }
}

}  // namespace

extern "C" {

// Fused Conv3d(4->16, k=3) + bias + ReLU + MaxPool3d(2, 2, padding=((D-2)%2, (H-2)%2, (W-2)%2)) — networks.py:15-24.
// x_dev float32 [n][4][D][H][W]; weight_dev float32 [16][4][3][3][3]; bias_dev float32 [16];
// out_dev bfloat16 [n][16][Pd][Ph][Pw] with P = (conv_out + pad - 2)/2 + 1; scratch_dev >= 16384 bytes.
// Requires W even (pool padding 0 on the last axis) and a plane that fits shared memory.
static int conv1_launch(const float *x_dev, const float *weight_dev, const float *bias_dev, int n, int D, int H, int W,
                        void *out_dev, void *scratch_dev, void *stream, bool grouped_out, const EnvSource *env_src = nullptr)
{
    if (n == 0) return RT_OK;
    if ((!x_dev && !env_src) || !weight_dev || !bias_dev || !out_dev || !scratch_dev || n < 0 || D < 3 || H < 3 || W < 4) return RT_ERR_INVALID;
    ConvShape S;
    S.D = D; S.H = H; S.W = W;
    S.Do = D - 2; S.Ho = H - 2; S.Wo = W - 2;
    S.pd = S.Do % 2; S.ph = S.Ho % 2;
    if (S.Wo % 2 != 0) return RT_ERR_INVALID;                       // width pool padding must be 0
    S.Pd = (S.Do + 2 * S.pd - 2) / 2 + 1;
    S.Ph = (S.Ho + 2 * S.ph - 2) / 2 + 1;
    S.Pw = (S.Wo - 2) / 2 + 1;
    S.tiles = (H * W + 31) / 32;
    S.plane_vox = S.tiles * 32 + 2 * W + 8;                         // rows read up to f + 2W + 3 (+1 for the voxel pair)
    auto magic = [](int d) { return d == 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
    S.mW = magic(W);
    S.mPw = magic(S.Pw);
    S.mPhPw = magic(S.Ph * S.Pw);
    if (S.tiles * 32 + 64 >= 65536 || S.Ph * S.Pw * kCout >= 65536) return RT_ERR_INVALID;
    S.r_elems = (S.Ho * S.Pw * kCout + 63) / 64 * 64;
    int dev = 0, max_smem = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RT_ERR_CUDA;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // split the depth of a sample over several blocks when there are few samples
    int chunks = 1;
    while (n * chunks < 2 * sms && chunks < S.Pd) chunks++;
    const int per = (S.Pd + chunks - 1) / chunks;
    chunks = (S.Pd + per - 1) / per;

    // 128-row tiles through a ring of tensor-memory slots; a plane too large for three pair buffers in shared memory
    // is not covered (RT_ERR_INVALID: the caller decides what to do, there is no second kernel)
    const int tiles128 = (H * W + kTcTileRows - 1) / kTcTileRows;
    ConvShape T = S;
    T.tiles = tiles128;
    T.plane_vox = (tiles128 * kTcTileRows + 2 * W + 8 + 7) / 8 * 8;
    const size_t smem_tc = (size_t)3 * T.plane_vox * 16 + (size_t)T.r_elems * sizeof(__nv_bfloat16) +
                           (size_t)kTcMmas * kBMmaBytes + (2 * kTcSlots + 6) * 8 + 16 +
                           (size_t)kHashSlots * 8 + 2 * sizeof(RayWork) +       // + from-env: view hash, rays, hit bits (1 per 8 voxels)
                           (size_t)((((size_t)D * H * W + 31) / 32 * 32 / 8 + 31) / 32 * 4 + 128);
    if (smem_tc > (size_t)max_smem || (size_t)3 * T.plane_vox * 16 >= (1u << 18) || W * 16 >= (1 << 18)) return RT_ERR_INVALID;
    {
        if (cudaFuncSetAttribute(rt_conv1_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc) != cudaSuccess ||
            cudaFuncSetAttribute(rt_conv1_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc) != cudaSuccess ||
            cudaFuncSetAttribute(rt_conv1_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc) != cudaSuccess) return RT_ERR_CUDA;
        __nv_bfloat16 *bop = reinterpret_cast<__nv_bfloat16 *>(scratch_dev);
        rt_conv_prepare_tc_kernel<<<(kTcMmas * 2 * kTcN * 8 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(weight_dev, bop);
        if (env_src)
            rt_conv1_tc_kernel<true, true><<<n * chunks, kTcThreads, smem_tc, (cudaStream_t)stream>>>(
                T, n, chunks, per, nullptr, reinterpret_cast<const uint4 *>(bop), bias_dev, reinterpret_cast<__nv_bfloat16 *>(out_dev), *env_src);
        else if (grouped_out)
            rt_conv1_tc_kernel<true, false><<<n * chunks, kTcThreads, smem_tc, (cudaStream_t)stream>>>(
                T, n, chunks, per, x_dev, reinterpret_cast<const uint4 *>(bop), bias_dev, reinterpret_cast<__nv_bfloat16 *>(out_dev), EnvSource{});
        else
            rt_conv1_tc_kernel<false, false><<<n * chunks, kTcThreads, smem_tc, (cudaStream_t)stream>>>(
                T, n, chunks, per, x_dev, reinterpret_cast<const uint4 *>(bop), bias_dev, reinterpret_cast<__nv_bfloat16 *>(out_dev), EnvSource{});
        return cudaGetLastError() == cudaSuccess ? RT_OK : RT_ERR_CUDA;
    }
}

int rt_conv1_relu_pool(const float *x_dev, const float *weight_dev, const float *bias_dev, int n, int D, int H, int W,
                       void *out_dev, void *scratch_dev, void *stream)
{
    return conv1_launch(x_dev, weight_dev, bias_dev, n, D, H, W, out_dev, scratch_dev, stream, false);
}

// Same block, output in the grouped channels-last layout [n][2][Pd][Ph*Pw][8] bf16 that rt_conv2_relu_pool reads.
int rt_conv1_relu_pool_grouped(const float *x_dev, const float *weight_dev, const float *bias_dev, int n, int D, int H,
                               int W, void *out_dev, void *scratch_dev, void *stream)
{
    return conv1_launch(x_dev, weight_dev, bias_dev, n, D, H, W, out_dev, scratch_dev, stream, true);
}

// The first block computed straight from the state of envs [first, first+count) of a (sparse-mode) handle: the
// observation of environment.py:245-257 is generated inside the kernel's loader warps instead of being assembled in
// HBM first (rt_assemble_volumes).  Output in the grouped layout, bit-identical to
// rt_assemble_volumes + rt_conv1_relu_pool_grouped.
int rt_conv1_from_env(rt_env *e, int first, int count, const float *weight_dev, const float *bias_dev, void *out_dev,
                      void *scratch_dev, void *stream)
{
    if (!e) return fail(RT_ERR_INVALID, "rt_conv1_from_env: NULL handle");
    if (first < 0 || count < 0 || first + count > e->n) return fail(RT_ERR_INVALID, "rt_conv1_from_env: env range out of bounds");
    if (e->dense) return fail(RT_ERR_STATE, "rt_conv1_from_env: not available for dense-mode handles");
    if (cudaSetDevice(e->device) != cudaSuccess) return RT_ERR_CUDA;
    EnvSource src{e->T, e->rec, e->cells, first};
    const Grid &G = e->T.G;
    const int rc = conv1_launch(nullptr, weight_dev, bias_dev, count, G.g0, G.g1, G.g2, out_dev, scratch_dev, stream, true, &src);
    if (rc == RT_OK && count > 0) g_launches.fetch_add(2, std::memory_order_relaxed);
    return rc;
}

// Fused Conv3d(16->16, k=3, groups=2) + bias + ReLU + MaxPool3d(2, 2) — networks.py:25-27.
// x_dev bfloat16 [n][2][D][H*W][8] (rt_conv1_relu_pool_grouped); weight_dev float32 [16][8][3][3][3]; bias_dev
// float32 [16]; out_dev bfloat16 [n][16][(D-2)/2][(H-2)/2][(W-2)/2] (NCDHW); scratch_dev >= 65536 bytes.
int rt_conv2_relu_pool(const void *x_dev, const float *weight_dev, const float *bias_dev, int n, int D, int H, int W,
                       void *out_dev, void *scratch_dev, void *stream)
{
    if (n == 0) return RT_OK;
    if (!x_dev || !weight_dev || !bias_dev || !out_dev || !scratch_dev || n < 0 || D < 4 || H < 4 || W < 4) return RT_ERR_INVALID;
    if (W % 2) return RT_ERR_INVALID;                               // the w pair of the pool must not straddle a row
    Conv2Shape S;
    S.D = D; S.H = H; S.W = W;
    S.Do = D - 2; S.Ho = H - 2; S.Wo = W - 2;
    S.Pd = S.Do / 2; S.Ph = S.Ho / 2; S.Pw = S.Wo / 2;
    if (S.Pd < 1 || S.Ph < 1 || S.Pw < 1) return RT_ERR_INVALID;
    S.tiles = (H * W + kTcTileRows - 1) / kTcTileRows;
    S.plane_vox = (S.tiles * kTcTileRows + 2 * W + 8 + 7) / 8 * 8;
    S.r_elems = (S.Ho * S.Pw * kCout + 63) / 64 * 64;
    auto magic = [](int d) { return d == 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
    S.mW = magic(W);
    S.mPw = magic(S.Pw);
    if (S.tiles * kTcTileRows + 64 >= 65536 || S.Ph * S.Pw * kCout >= 65536) return RT_ERR_INVALID;
    int dev = 0, max_smem = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RT_ERR_CUDA;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)12 * S.plane_vox * 16 + (size_t)S.r_elems * sizeof(__nv_bfloat16) +
                        (size_t)kC2Mmas * kBMmaBytes + (2 * kTcSlots + 6) * 8 + 16;
    if (smem > (size_t)max_smem || (size_t)12 * S.plane_vox * 16 >= (1u << 18)) return RT_ERR_INVALID;
    int chunks = 1;
    while (n * chunks < 2 * sms && chunks < S.Pd) chunks++;
    const int per = (S.Pd + chunks - 1) / chunks;
    chunks = (S.Pd + per - 1) / per;
    if (cudaFuncSetAttribute(rt_conv2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return RT_ERR_CUDA;
    __nv_bfloat16 *bop = reinterpret_cast<__nv_bfloat16 *>(scratch_dev);
    rt_conv2_prepare_kernel<<<(kC2Mmas * 2 * kTcN * 8 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(weight_dev, bop);
    rt_conv2_tc_kernel<<<n * chunks, kC2Threads, smem, (cudaStream_t)stream>>>(
        S, n, chunks, per, reinterpret_cast<const uint4 *>(x_dev), reinterpret_cast<const uint4 *>(bop), bias_dev,
        reinterpret_cast<__nv_bfloat16 *>(out_dev));
    return cudaGetLastError() == cudaSuccess ? RT_OK : RT_ERR_CUDA;
}

// Tail of FeaturesExtractor3D (networks.py:28-45): Conv3d(16->16, k=3, groups=4) + ReLU + MaxPool3d(2, 2) + Flatten +
// Linear(16*Pd*Ph*Pw -> F) + ReLU.  x_dev bfloat16 [n][16][D][H][W] (rt_conv2_relu_pool), conv_w float32 [16][4][3][3][3],
// conv_b [16], lin_w float32 [F][16*Pd*Ph*Pw] (flatten order c, d, h, w), lin_b [F] -> out_dev float32 [n][F].
int rt_c3d_tail(const void *x_dev, const float *conv_w_dev, const float *conv_b_dev, const float *lin_w_dev,
                const float *lin_b_dev, int n, int D, int H, int W, int F, float *out_dev, void *stream)
{
    if (n == 0) return RT_OK;
    if (!x_dev || !conv_w_dev || !conv_b_dev || !lin_w_dev || !lin_b_dev || !out_dev || n < 0 || D < 4 || H < 4 || W < 4 ||
        F < 1 || F > kTailMaxF)
        return RT_ERR_INVALID;
    const int Pd = (D - 2) / 2, Ph = (H - 2) / 2, Pw = (W - 2) / 2;
    const size_t smem = ((size_t)16 * D * H * W + 7) / 8 * 8 * sizeof(__nv_bfloat16) + (size_t)16 * 4 * 27 * sizeof(float) +
                        (size_t)16 * Pd * Ph * Pw * sizeof(float);
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RT_ERR_CUDA;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_smem) return RT_ERR_INVALID;
    if (cudaFuncSetAttribute(rt_c3d_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return RT_ERR_CUDA;
    rt_c3d_tail_kernel<<<n, kTailThreads, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16 *>(x_dev), D, H, W,
                                                                     conv_w_dev, conv_b_dev, lin_w_dev, lin_b_dev, F, out_dev);
    return cudaGetLastError() == cudaSuccess ? RT_OK : RT_ERR_CUDA;
}

}  // extern "C"
