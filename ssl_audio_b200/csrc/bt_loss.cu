// Barlow Twins objective, forward + backward, for sm_100a.
//
// Replaces utils/loss.py:15-30 (BarlowTwinsLoss.forward_loss) and its autograd backward
// (reference: /root/reference).  Closed form, N rows, D columns:
//   zh = (z - mu) * r,  r = 1/sqrt(var_biased + eps)           (nn.BatchNorm1d, affine=False)
//   C  = zh1^T zh2 / N
//   L  = alpha * sum_i (C_ii - 1)^2 + lambda * sum_{i!=j} (C_ij + h)^2     (h = 1 iff HSIC)
//   G  = dL/dC,  dL/dzh1 = zh2 G^T / N,  dL/dzh2 = zh1 G / N,  then batch-norm backward.
//
// The D x D matrix is never materialised in fp32.  Four launches:
//   1. bt_colstat_kernel   partial column sums of both views (shifted data), grid = column blocks x row chunks
//   2. bt_normalize_kernel column statistics, fp32 diagonal C_ii and on-diagonal loss, BatchNorm running-stat
//                          update, and the fp16 standardised embeddings zh (operand B of the gradient GEMMs)
//   3. bt_umma_kernel CORR S = z1^T z2 on the tensor cores (tcgen05, RAW bf16 operands straight from the
//                          row-major embeddings as MN-major TMA tiles, fp32 TMEM accumulator); the epilogue
//                          applies batch-norm as a rank-1 correction, reduces the off-diagonal loss in fp32,
//                          emits C (|C_ij| <= 1) in fp16 with the diagonal zeroed, and accumulates the row and
//                          column sums of C o C (warp transpose-reduce for the columns)
//   4. bt_umma_kernel GRAD g1^T = C zh2^T (K-major A) and g2^T = C^T zh1^T (MN-major A over the same C);
//                          the epilogue adds the fp32 diagonal term and applies batch-norm backward straight
//                          out of TMEM, writing dz in the caller's dtype: 6 N D^2 executed FLOP = the
//                          algorithmic count, no fp32 gradient round trip through HBM.
//
// Batch-norm backward, dz = r (g - mean_n(g) - zh mean_n(g o zh)), needs two column means.  Both have closed
// forms in C: mean_n(g) = 0 (columns of zh have zero mean), and
//   mean_n(g1 o zh1)_i = (2 lambda / N) sum_{j != i} C_ij (C_ij + h) + (G_ii / N) C_ii
// (g2: the same with column sums), which is why the CORR epilogue keeps the row / column sums of C o C (and of
// C when HSIC) and no pass over the gradients is needed before they are written.
//
// Row-block mode (multi-GPU, abt_bt_loss_rows_fwd_bwd): the inputs are the rank-ordered gathered
// embeddings (N_g x D); this rank owns dimensions [row_begin, row_begin + row_count) and computes
// that row block of C AND of C^T (second CORR pass with the views swapped), so that both gradient
// GEMMs are complete for its dimensions and batch-norm backward (column-local) needs no reduction.
#include "abt_internal.h"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

namespace abt {

// ------------------------------------------------------------------------------------------
// column-statistics layout inside the workspace (float arrays of length D each)
// ------------------------------------------------------------------------------------------
enum StatSlot {
    S_MU1 = 0, S_R1, S_MU2, S_R2, S_CDIAG,
    S_NMU1, S_RHO1,   // -N * mu1, r1 / N : row constants of the rank-1 batch-norm correction (rows = view-1 dims)
    S_NMU2, S_RHO2,   // same with the views swapped (row-block mode, C^T pass)
    S_ZERO, S_ONE, S_INVN,   // constants 0, 1, 1/N: the CORR epilogue on already standardised operands (multi-GPU) is c = S / N
    S_COUNT
};
// accumulators zeroed at the start of every call (float arrays of length D each, after the 256-byte misc block)
enum AccSlot {
    A_SQ1 = 0,   // sum_{j != i} C_ij^2 over row i of C          (-> batch-norm backward of dz1)
    A_SQ2,       // sum_{i != j} C_ij^2 over column j of C       (-> dz2)
    A_SUM1,      // HSIC: sum_{j != i} C_ij
    A_SUM2,
    A_COUNT
};

template <typename T> struct Ld2;
template <> struct Ld2<__nv_bfloat16> {
    static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
        return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float a, float b) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
    }
};
template <> struct Ld2<__half> {
    static __device__ __forceinline__ float2 ld(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
    static __device__ __forceinline__ void st(__half* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }
};
template <> struct Ld2<float> {
    static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
    static __device__ __forceinline__ void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

constexpr int kColsPerBlock = 64;   // multi-GPU normalise kernel: 32 lanes x 2 columns
constexpr int kRowGroups = 8;       // 256 threads per block; the grid's second dimension splits the rows
constexpr int kColThreads = kRowGroups * 32;
constexpr int kMaxRowSplits = 32;
constexpr int kVecCols = 8;                       // columns per thread: one 16-byte load of 16-bit embeddings
constexpr int kWideCols = 32 * kVecCols;          // 256 columns per block

// eight consecutive embeddings of type T as floats (16-byte loads)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k])); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> struct NeedsRound { static constexpr bool value = true; };
template <> struct NeedsRound<__nv_bfloat16> { static constexpr bool value = false; };     // already bf16: rounding is the identity
template <> struct NeedsRound<__half> { static constexpr bool value = false; };            // fp16 embeddings feed the tensor cores as fp16 (kind::f16, no re-rounding)
// the value the tensor cores see for an input element: fp32 inputs are rounded to bf16, 16-bit inputs are used as they are
template <typename T> __device__ __forceinline__ float in_round(float x) { return NeedsRound<T>::value ? bf16_round(x) : x; }

constexpr int kStatVec = 4;                       // statistics kernel: 4 columns per thread (8-byte loads keep it at ~50 registers, 5 blocks / SM)
constexpr int kStatCols = 32 * kStatVec;          // 128 columns per block

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// ------------------------------------------------------------------------------------------
// 1. partial column sums: block (cb, rs) handles 128 columns x rows [rs * chunk, (rs + 1) * chunk)
//    partials[(rs * 5 + k) * D + col], k = sum da, sum da^2, sum db, sum db^2, sum da db  (da = z1 - z1[0], shifted data)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kColThreads) bt_colstat_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int N, int D, int chunk,
                                                                 float* __restrict__ partials, __nv_bfloat16* __restrict__ zb1,
                                                                 __nv_bfloat16* __restrict__ zb2, unsigned int* __restrict__ counters, float eps,
                                                                 float momentum, float* __restrict__ stats, float* __restrict__ running_mean,
                                                                 float* __restrict__ running_var, double* __restrict__ loss_acc) {
    __shared__ float red[kRowGroups][5][kStatCols];
    __shared__ float on_red[8];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * kStatCols + lane * kStatVec;
    const int n0 = blockIdx.y * chunk, n1 = min(N, n0 + chunk);
    float s1[4], q1[4], s2[4], q2[4], x12[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { s1[c] = q1[c] = s2[c] = q2[c] = x12[c] = 0.f; }
    if (col < D) {
        // shifted-data sums: subtract row 0 so that |mu| >> sigma does not cancel in fp32
        float k1[4], k2[4];
        load4<T>(z1 + col, k1); load4<T>(z2 + col, k2);
        if (NeedsRound<T>::value) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { k1[c] = bf16_round(k1[c]); k2[c] = bf16_round(k2[c]); }
        }
#pragma unroll 4
        for (int n = n0 + rg; n < n1; n += kRowGroups) {
            float av[4], bv[4];
            load4<T>(z1 + (size_t)n * D + col, av); load4<T>(z2 + (size_t)n * D + col, bv);
            if (NeedsRound<T>::value) {
                // the tensor cores consume bf16: statistics are those of the bf16-rounded embeddings
                uint32_t pa[2], pb[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    pa[c] = pack_bf16x2(av[2 * c], av[2 * c + 1]); pb[c] = pack_bf16x2(bv[2 * c], bv[2 * c + 1]);
                    av[2 * c] = __uint_as_float(pa[c] << 16); av[2 * c + 1] = __uint_as_float(pa[c] & 0xffff0000u);
                    bv[2 * c] = __uint_as_float(pb[c] << 16); bv[2 * c + 1] = __uint_as_float(pb[c] & 0xffff0000u);
                }
                if (zb1 != nullptr) {
                    *reinterpret_cast<uint2*>(zb1 + (size_t)n * D + col) = make_uint2(pa[0], pa[1]);
                    *reinterpret_cast<uint2*>(zb2 + (size_t)n * D + col) = make_uint2(pb[0], pb[1]);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float da = av[c] - k1[c], db = bv[c] - k2[c];
                s1[c] += da; q1[c] = fmaf(da, da, q1[c]);
                s2[c] += db; q2[c] = fmaf(db, db, q2[c]);
                x12[c] = fmaf(da, db, x12[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {      // [c][lane] order: conflict-free (column lane * 4 + c lives at c * 32 + lane)
        red[rg][0][c * 32 + lane] = s1[c]; red[rg][1][c * 32 + lane] = q1[c];
        red[rg][2][c * 32 + lane] = s2[c]; red[rg][3][c * 32 + lane] = q2[c];
        red[rg][4][c * 32 + lane] = x12[c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * kStatCols; i += kColThreads) {
        const int k = i / kStatCols, sidx = i % kStatCols, gc = blockIdx.x * kStatCols + (sidx & 31) * kStatVec + (sidx >> 5);
        if (gc < D) {
            float t = 0.f;
#pragma unroll
            for (int g = 0; g < kRowGroups; ++g) t += red[g][k][sidx];
            partials[((size_t)blockIdx.y * 5 + k) * D + gc] = t;
        }
    }
    if (counters == nullptr) return;      // multi-GPU: the partial sums are packed and exchanged instead
    // the LAST row chunk of this column block to finish folds the partial sums in a fixed order (deterministic) into the statistics
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counters + blockIdx.x, 1u) == gridDim.y - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float on = 0.f;
    const int c = threadIdx.x, gc = blockIdx.x * kStatCols + c;
    if (c < kStatCols && gc < D) {
        float t[5] = {0, 0, 0, 0, 0};
        for (int sp = 0; sp < (int)gridDim.y; ++sp)
#pragma unroll
            for (int k = 0; k < 5; ++k) t[k] += __ldcg(partials + ((size_t)sp * 5 + k) * D + gc);
        const float invN = 1.0f / (float)N;
        const float2 f1 = Ld2<T>::ld(z1 + (gc & ~1)), f2 = Ld2<T>::ld(z2 + (gc & ~1));
        const float sh1 = in_round<T>((gc & 1) ? f1.y : f1.x), sh2 = in_round<T>((gc & 1) ? f2.y : f2.x);
        const float m1 = t[0] * invN, m2 = t[2] * invN;
        const float var1 = fmaxf(t[1] * invN - m1 * m1, 0.f), var2 = fmaxf(t[3] * invN - m2 * m2, 0.f);
        const float cov = t[4] * invN - m1 * m2;
        const float mu1 = sh1 + m1, mu2 = sh2 + m2;
        const float r1 = rsqrtf(var1 + eps), r2 = rsqrtf(var2 + eps);
        // one Newton step: rsqrtf is ~2 ulp, BatchNorm uses a correctly rounded 1/sqrt
        const float r1n = r1 * (1.5f - 0.5f * (var1 + eps) * r1 * r1), r2n = r2 * (1.5f - 0.5f * (var2 + eps) * r2 * r2);
        const float cd = cov * r1n * r2n;
        stats[S_MU1 * D + gc] = mu1; stats[S_R1 * D + gc] = r1n;
        stats[S_MU2 * D + gc] = mu2; stats[S_R2 * D + gc] = r2n;
        stats[S_CDIAG * D + gc] = cd;
        stats[S_NMU1 * D + gc] = -(float)N * mu1; stats[S_RHO1 * D + gc] = r1n * invN;
        stats[S_NMU2 * D + gc] = -(float)N * mu2; stats[S_RHO2 * D + gc] = r2n * invN;
        on = (cd - 1.0f) * (cd - 1.0f);
        if (running_mean != nullptr) {
            // BatchNorm1d training-mode side effect, view 1 then view 2 (utils/loss.py:17)
            const float unb = (N > 1) ? (float)N / (float)(N - 1) : 1.0f;
            float rm = running_mean[gc], rv = running_var[gc];
            rm = (1.f - momentum) * rm + momentum * mu1; rv = (1.f - momentum) * rv + momentum * var1 * unb;
            rm = (1.f - momentum) * rm + momentum * mu2; rv = (1.f - momentum) * rv + momentum * var2 * unb;
            running_mean[gc] = rm; running_var[gc] = rv;
        }
    }
    // on-diagonal loss sum_i (C_ii - 1)^2: one double atomic per column block
    on = warp_sum(on);
    if (lane == 0) on_red[rg] = on;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tsum += on_red[w];
        atomicAdd(loss_acc + 2, (double)tsum);
    }
}

// ------------------------------------------------------------------------------------------
// 2. the standardised fp16 embeddings (operand B of the gradient GEMMs; |zh| <= sqrt(N)): elementwise, 16-byte accesses
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kColThreads) bt_normalize_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int N, int D, int chunk,
                                                                   const float* __restrict__ stats, __half* __restrict__ zh1,
                                                                   __half* __restrict__ zh2) {
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * kWideCols + lane * kVecCols;
    if (col >= D) return;
    float m1[8], q1r[8], m2[8], q2r[8];
#pragma unroll
    for (int c4 = 0; c4 < 2; ++c4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(stats + S_MU1 * D + col) + c4), b = __ldg(reinterpret_cast<const float4*>(stats + S_R1 * D + col) + c4);
        const float4 e = __ldg(reinterpret_cast<const float4*>(stats + S_MU2 * D + col) + c4), f = __ldg(reinterpret_cast<const float4*>(stats + S_R2 * D + col) + c4);
        m1[4 * c4] = a.x; m1[4 * c4 + 1] = a.y; m1[4 * c4 + 2] = a.z; m1[4 * c4 + 3] = a.w;
        q1r[4 * c4] = b.x; q1r[4 * c4 + 1] = b.y; q1r[4 * c4 + 2] = b.z; q1r[4 * c4 + 3] = b.w;
        m2[4 * c4] = e.x; m2[4 * c4 + 1] = e.y; m2[4 * c4 + 2] = e.z; m2[4 * c4 + 3] = e.w;
        q2r[4 * c4] = f.x; q2r[4 * c4 + 1] = f.y; q2r[4 * c4 + 2] = f.z; q2r[4 * c4 + 3] = f.w;
    }
    const int n0 = blockIdx.y * chunk, n1 = min(N, n0 + chunk);
#pragma unroll 2
    for (int n = n0 + rg; n < n1; n += kRowGroups) {
        const size_t o = (size_t)n * D + col;
        float av[8], bv[8];
        load8<T>(z1 + o, av); load8<T>(z2 + o, bv);
        uint32_t ha[4], hb[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float a0 = av[2 * c], a1 = av[2 * c + 1], b0 = bv[2 * c], b1 = bv[2 * c + 1];
            if (NeedsRound<T>::value) { a0 = bf16_round(a0); a1 = bf16_round(a1); b0 = bf16_round(b0); b1 = bf16_round(b1); }
            ha[c] = pack_f16x2((a0 - m1[2 * c]) * q1r[2 * c], (a1 - m1[2 * c + 1]) * q1r[2 * c + 1]);
            hb[c] = pack_f16x2((b0 - m2[2 * c]) * q2r[2 * c], (b1 - m2[2 * c + 1]) * q2r[2 * c + 1]);
        }
        *reinterpret_cast<uint4*>(zh1 + o) = make_uint4(ha[0], ha[1], ha[2], ha[3]);
        *reinterpret_cast<uint4*>(zh2 + o) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
    }
}

// ------------------------------------------------------------------------------------------
// 1s. small batches (N <= 128, the one-launch objective): statistics AND standardisation in ONE pass.  A block owns 32 columns
//     (half an operand tile; 256 blocks at D = 8192, two resident per SM) and keeps its N x 32 slice of both views in registers:
//     a thread holds one 16-byte piece (8 columns) of up to 2 rows of each view, so z is read with 16-byte loads and the
//     standardised rows leave as the 16-byte pieces of the tile image.  The shifted sums are reduced through shared memory in two
//     steps; z is read once, no partial-sum round trip, no arrival counters.  The on-diagonal loss leaves as one float per block
//     (summed by the consumer), and block 0 clears the accumulators of the tensor-core kernel, so the call needs no memset either.
// ------------------------------------------------------------------------------------------
constexpr int kSmallCols = 64;                     // columns of one tile image (one operand tile of the tensor-core kernel)
constexpr int kSmallBlockCols = 32;                // columns per block
constexpr int kSmallRowLanes = 64;                 // 8 warps x 8 row lanes; row n is held by row lane n % 64, slot n / 64
template <typename T>
__global__ void __launch_bounds__(kColThreads, 2) bt_stat_norm_small_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int N, int D, float eps,
                                                                            float momentum, float* __restrict__ stats, float* __restrict__ running_mean,
                                                                            float* __restrict__ running_var, __half* __restrict__ zh1,
                                                                            __half* __restrict__ zh2, float* __restrict__ ondiag_part,
                                                                            double* __restrict__ loss_acc, unsigned int* __restrict__ done_counter,
                                                                            int tile_img_rows) {
    // red[row lane][sum k][position]: position h * 16 + piece * 4 + j holds local column piece * 8 + h * 4 + j (16-byte stores)
    __shared__ __align__(16) float red[kSmallRowLanes][5][kSmallBlockCols];
    __shared__ float red2[8][5][kSmallBlockCols];
    __shared__ __align__(16) float colstat[4][kSmallBlockCols];
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int piece = lane & 3, rl = rg * 8 + (lane >> 2);        // 16-byte piece of this block's half tile row; row lane 0..63
    const int col = blockIdx.x * kSmallBlockCols + piece * 8;     // D is a multiple of 64: always in range
    griddep_launch_dependents();                                  // the tensor-core kernel may set itself up while this one runs
    if (blockIdx.x == 0 && threadIdx.x == 0) { loss_acc[0] = 0.0; loss_acc[1] = 0.0; *done_counter = 0u; }
    float sh1[8], sh2[8];
    load8<T>(z1 + col, sh1); load8<T>(z2 + col, sh2);
#pragma unroll
    for (int c = 0; c < 8; ++c) { sh1[c] = in_round<T>(sh1[c]); sh2[c] = in_round<T>(sh2[c]); }
    float va[2][8], vb[2][8];
    float acc[5][8];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[k][c] = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int n = rl + kSmallRowLanes * k;
        if (n < N) { load8<T>(z1 + (size_t)n * D + col, va[k]); load8<T>(z2 + (size_t)n * D + col, vb[k]); }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int n = rl + kSmallRowLanes * k;
        if (n < N) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {                          // shifted data
                va[k][c] = in_round<T>(va[k][c]) - sh1[c]; vb[k][c] = in_round<T>(vb[k][c]) - sh2[c];
                acc[0][c] += va[k][c]; acc[1][c] = fmaf(va[k][c], va[k][c], acc[1][c]);
                acc[2][c] += vb[k][c]; acc[3][c] = fmaf(vb[k][c], vb[k][c], acc[3][c]);
                acc[4][c] = fmaf(va[k][c], vb[k][c], acc[4][c]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        *reinterpret_cast<float4*>(&red[rl][k][piece * 4]) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
        *reinterpret_cast<float4*>(&red[rl][k][16 + piece * 4]) = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
    }
    __syncthreads();
    {   // step 1: thread (position, eighth) folds 8 row lanes of all five sums
        const int pos = threadIdx.x & 31, part = threadIdx.x >> 5;
        float t[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < 8; ++g)
#pragma unroll
            for (int k = 0; k < 5; ++k) t[k] += red[part * 8 + g][k][pos];
#pragma unroll
        for (int k = 0; k < 5; ++k) red2[part][k][pos] = t[k];
    }
    __syncthreads();
    if (threadIdx.x < kSmallBlockCols) {
        const int lc = threadIdx.x, gc = blockIdx.x * kSmallBlockCols + lc;                   // local / global column
        const int pos = ((lc >> 2) & 1) * 16 + (lc >> 3) * 4 + (lc & 3);
        float t[5];
#pragma unroll
        for (int k = 0; k < 5; ++k)
            t[k] = ((red2[0][k][pos] + red2[1][k][pos]) + (red2[2][k][pos] + red2[3][k][pos])) + ((red2[4][k][pos] + red2[5][k][pos]) + (red2[6][k][pos] + red2[7][k][pos]));
        const float invN = 1.0f / (float)N;
        const float2 f1 = Ld2<T>::ld(z1 + (gc & ~1)), f2 = Ld2<T>::ld(z2 + (gc & ~1));
        const float s1 = in_round<T>((lc & 1) ? f1.y : f1.x), s2 = in_round<T>((lc & 1) ? f2.y : f2.x);
        const float m1 = t[0] * invN, m2 = t[2] * invN;
        const float var1 = fmaxf(t[1] * invN - m1 * m1, 0.f), var2 = fmaxf(t[3] * invN - m2 * m2, 0.f);
        const float cov = t[4] * invN - m1 * m2;
        const float mu1 = s1 + m1, mu2 = s2 + m2;
        const float r1 = rsqrtf(var1 + eps), r2 = rsqrtf(var2 + eps);
        const float r1n = r1 * (1.5f - 0.5f * (var1 + eps) * r1 * r1), r2n = r2 * (1.5f - 0.5f * (var2 + eps) * r2 * r2);
        const float cd = cov * r1n * r2n;
        stats[S_MU1 * D + gc] = mu1; stats[S_R1 * D + gc] = r1n;
        stats[S_MU2 * D + gc] = mu2; stats[S_R2 * D + gc] = r2n;
        stats[S_CDIAG * D + gc] = cd;
        colstat[0][lc] = m1; colstat[1][lc] = r1n; colstat[2][lc] = m2; colstat[3][lc] = r2n;
        if (running_mean != nullptr) {
            const float unb = (N > 1) ? (float)N / (float)(N - 1) : 1.0f;
            float rm = running_mean[gc], rv = running_var[gc];
            rm = (1.f - momentum) * rm + momentum * mu1; rv = (1.f - momentum) * rv + momentum * var1 * unb;
            rm = (1.f - momentum) * rm + momentum * mu2; rv = (1.f - momentum) * rv + momentum * var2 * unb;
            running_mean[gc] = rm; running_var[gc] = rv;
        }
        const float on = warp_sum((cd - 1.0f) * (cd - 1.0f));      // kSmallBlockCols == 32: exactly warp 0
        if (lane == 0) ondiag_part[blockIdx.x] = on;
    }
    __syncthreads();
    // standardise from the registers (the values are already shifted: subtract the shifted mean)
    float m1c[8], r1c[8], m2c[8], r2c[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(&colstat[0][piece * 8 + 4 * h]), b = *reinterpret_cast<const float4*>(&colstat[1][piece * 8 + 4 * h]);
        const float4 e = *reinterpret_cast<const float4*>(&colstat[2][piece * 8 + 4 * h]), f = *reinterpret_cast<const float4*>(&colstat[3][piece * 8 + 4 * h]);
        m1c[4 * h] = a.x; m1c[4 * h + 1] = a.y; m1c[4 * h + 2] = a.z; m1c[4 * h + 3] = a.w;
        r1c[4 * h] = b.x; r1c[4 * h + 1] = b.y; r1c[4 * h + 2] = b.z; r1c[4 * h + 3] = b.w;
        m2c[4 * h] = e.x; m2c[4 * h + 1] = e.y; m2c[4 * h + 2] = e.z; m2c[4 * h + 3] = e.w;
        r2c[4 * h] = f.x; r2c[4 * h + 1] = f.y; r2c[4 * h + 2] = f.z; r2c[4 * h + 3] = f.w;
    }
    // Tile-image layout: 64 columns (two blocks) are ONE operand tile of the tensor-core kernel.  It is written exactly as that
    // kernel wants it in shared memory -- tile_img_rows rows of 128 bytes (one per sample, zero rows beyond N), 16-byte pieces
    // XOR-swizzled with the row index (the 128-byte swizzle of a TMA tile) -- so that the kernel fetches it with ONE contiguous
    // bulk copy instead of a 128-row tensor-map box (which a single SM ingests at only ~24 B / clk).
    const int tile = blockIdx.x >> 1, tpiece = (blockIdx.x & 1) * 4 + piece;
    __half* t1 = zh1 + (size_t)tile * tile_img_rows * kSmallCols;
    __half* t2 = zh2 + (size_t)tile * tile_img_rows * kSmallCols;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int n = rl + kSmallRowLanes * k;
        if (n < tile_img_rows) {
            const size_t o = (size_t)n * kSmallCols + (size_t)((tpiece ^ (n & 7)) * 8);
            uint32_t ha[4] = {0u, 0u, 0u, 0u}, hb[4] = {0u, 0u, 0u, 0u};
            if (n < N) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    ha[c] = pack_f16x2((va[k][2 * c] - m1c[2 * c]) * r1c[2 * c], (va[k][2 * c + 1] - m1c[2 * c + 1]) * r1c[2 * c + 1]);
                    hb[c] = pack_f16x2((vb[k][2 * c] - m2c[2 * c]) * r2c[2 * c], (vb[k][2 * c + 1] - m2c[2 * c + 1]) * r2c[2 * c + 1]);
                }
            }
            *reinterpret_cast<uint4*>(t1 + o) = make_uint4(ha[0], ha[1], ha[2], ha[3]);
            *reinterpret_cast<uint4*>(t2 + o) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
        }
    }
    // D % 128 == 64: the walk of the tensor-core kernel visits one block past the last tile; it must read zeros
    const int n_tiles = gridDim.x >> 1;
    if (blockIdx.x == 0 && (n_tiles & 1)) {
        uint4* e1 = reinterpret_cast<uint4*>(zh1 + (size_t)n_tiles * tile_img_rows * kSmallCols);
        uint4* e2 = reinterpret_cast<uint4*>(zh2 + (size_t)n_tiles * tile_img_rows * kSmallCols);
        for (int i = threadIdx.x; i < tile_img_rows * kSmallCols / 8; i += kColThreads) { e1[i] = make_uint4(0u, 0u, 0u, 0u); e2[i] = make_uint4(0u, 0u, 0u, 0u); }
    }
}

// row sums of standardised embeddings stored as tile images (HSIC only): the swizzle permutes columns inside a tile row, the sum does not care
__global__ void __launch_bounds__(256) bt_rowsum_img_kernel(const __half* __restrict__ zimg, int n_tiles, int img_rows, float* __restrict__ out) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int e = threadIdx.x; e < n_tiles * kSmallCols; e += blockDim.x)
        acc += __half2float(zimg[((size_t)(e / kSmallCols) * img_rows + n) * kSmallCols + (e % kSmallCols)]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[n] = t;
    }
}

// row sums of the standardised embeddings (HSIC only): R[n] = sum_j zh[n, j]
template <typename T16>
__global__ void __launch_bounds__(256) bt_rowsum_kernel(const T16* __restrict__ z, int N, int D, const float* __restrict__ mu,
                                                        const float* __restrict__ r, float* __restrict__ out) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) acc += ((float)z[(size_t)n * D + c] - mu[c]) * r[c];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[n] = t;
    }
}

// in-place gradient scaling for autograd's backward: dz *= *scale (both views in one launch, 16-byte accesses)
template <typename T>
__global__ void __launch_bounds__(256) bt_scale_kernel(T* __restrict__ a, T* __restrict__ b, size_t n_vec, const float* __restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.0f) return;        // loss.backward() on the bare loss: grad_output is 1, the stored gradients are already final
    constexpr int E = 16 / sizeof(T);
    T* base = blockIdx.y == 0 ? a : b;
    if (base == nullptr) return;
    uint4* p4 = reinterpret_cast<uint4*>(base);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = p4[i];
        T* e = reinterpret_cast<T*>(&v);
#pragma unroll
        for (int k = 0; k < E; k += 2) {
            const float2 f = Ld2<T>::ld(e + k);
            Ld2<T>::st(e + k, f.x * s, f.y * s);
        }
        p4[i] = v;
    }
}

// same for already standardised fp16 embeddings
__global__ void __launch_bounds__(256) bt_rowsum_zh_kernel(const __half* __restrict__ zh, int N, int D, float* __restrict__ out) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) acc += __half2float(zh[(size_t)n * D + c]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[n] = t;
    }
}

// ------------------------------------------------------------------------------------------
// multi-GPU statistics: every rank reduces its LOCAL rows to 7 numbers per column (5 shifted sums + the two shifts),
// the ranks all-gather those (7 D floats each), and every rank combines them into the GLOBAL batch statistics
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bt_stat_pack_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int D, int n_splits,
                                                           const float* __restrict__ partials, float* __restrict__ pack) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float t = 0.f;
        for (int sp = 0; sp < n_splits; ++sp) t += partials[((size_t)sp * 5 + k) * D + c];
        pack[(size_t)k * D + c] = t;
    }
    const float2 a = Ld2<T>::ld(z1 + (c & ~1)), b = Ld2<T>::ld(z2 + (c & ~1));
    pack[(size_t)5 * D + c] = in_round<T>((c & 1) ? a.y : a.x);
    pack[(size_t)6 * D + c] = in_round<T>((c & 1) ? b.y : b.x);
}

// packs: [world][7][D].  Combines the per-rank shifted sums in double (exact re-centring), writes the statistics arrays, the
// BatchNorm running-stat update, the on-diagonal loss, and the fp16 standardised LOCAL rows (into this rank's slot of the
// all-gather buffers).
template <typename T>
__global__ void __launch_bounds__(kColThreads) bt_normalize_global_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int n_local, int world,
                                                                          int D, int chunk, float eps, float momentum,
                                                                          const float* __restrict__ packs, float* __restrict__ stats,
                                                                          __half* __restrict__ zh1, __half* __restrict__ zh2,
                                                                          __half* __restrict__ zh1_blk, int dr,
                                                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                          double* __restrict__ ondiag) {
    __shared__ float colstat[4][kColsPerBlock];
    __shared__ float on_red[2];
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * kColsPerBlock + lane * 2;
    float on = 0.f;
    if (threadIdx.x < kColsPerBlock) {
        const int c = threadIdx.x, gc = blockIdx.x * kColsPerBlock + c;
        if (gc < D) {
            const double n = (double)n_local, Ng = (double)n_local * world;
            double s1 = 0, s2 = 0;
            for (int r = 0; r < world; ++r) {
                const float* pk = packs + (size_t)r * 7 * D;
                s1 += (double)pk[0 * D + gc] + n * (double)pk[5 * D + gc];
                s2 += (double)pk[2 * D + gc] + n * (double)pk[6 * D + gc];
            }
            const double m1 = s1 / Ng, m2 = s2 / Ng;
            double v1 = 0, v2 = 0, cv = 0;
            for (int r = 0; r < world; ++r) {
                const float* pk = packs + (size_t)r * 7 * D;
                const double a1 = pk[0 * D + gc], q1 = pk[1 * D + gc], a2 = pk[2 * D + gc], q2 = pk[3 * D + gc], x12 = pk[4 * D + gc];
                const double d1 = (double)pk[5 * D + gc] - m1, d2 = (double)pk[6 * D + gc] - m2;
                v1 += q1 + 2.0 * d1 * a1 + n * d1 * d1;
                v2 += q2 + 2.0 * d2 * a2 + n * d2 * d2;
                cv += x12 + d2 * a1 + d1 * a2 + n * d1 * d2;
            }
            const float var1 = (float)fmax(v1 / Ng, 0.0), var2 = (float)fmax(v2 / Ng, 0.0);
            const float mu1 = (float)m1, mu2 = (float)m2;
            const float r1n = (float)(1.0 / sqrt((double)var1 + (double)eps)), r2n = (float)(1.0 / sqrt((double)var2 + (double)eps));
            const float cd = (float)(cv / Ng) * r1n * r2n;
            colstat[0][c] = mu1; colstat[1][c] = r1n; colstat[2][c] = mu2; colstat[3][c] = r2n;
            if (blockIdx.y == 0) {
                const float invN = (float)(1.0 / Ng);
                stats[S_MU1 * D + gc] = mu1; stats[S_R1 * D + gc] = r1n;
                stats[S_MU2 * D + gc] = mu2; stats[S_R2 * D + gc] = r2n;
                stats[S_CDIAG * D + gc] = cd;
                stats[S_ZERO * D + gc] = 0.f; stats[S_ONE * D + gc] = 1.f; stats[S_INVN * D + gc] = invN;
                on = (cd - 1.0f) * (cd - 1.0f);
                if (running_mean != nullptr) {
                    const float unb = (Ng > 1.0) ? (float)(Ng / (Ng - 1.0)) : 1.0f;
                    float rm = running_mean[gc], rv = running_var[gc];
                    rm = (1.f - momentum) * rm + momentum * mu1; rv = (1.f - momentum) * rv + momentum * var1 * unb;
                    rm = (1.f - momentum) * rm + momentum * mu2; rv = (1.f - momentum) * rv + momentum * var2 * unb;
                    running_mean[gc] = rm; running_var[gc] = rv;
                }
            }
        }
        on = warp_sum(on);
        if (lane == 0) on_red[threadIdx.x >> 5] = on;
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.y == 0) atomicAdd(ondiag, (double)(on_red[0] + on_red[1]));
    if (col < D) {
        const float m1[2] = {colstat[0][lane * 2], colstat[0][lane * 2 + 1]}, q1r[2] = {colstat[1][lane * 2], colstat[1][lane * 2 + 1]};
        const float m2[2] = {colstat[2][lane * 2], colstat[2][lane * 2 + 1]}, q2r[2] = {colstat[3][lane * 2], colstat[3][lane * 2 + 1]};
        const int n0 = blockIdx.y * chunk, n1 = min(n_local, n0 + chunk);
#pragma unroll 4
        for (int n = n0 + rg; n < n1; n += kRowGroups) {
            const size_t o = (size_t)n * D + col;
            const float2 a2 = Ld2<T>::ld(z1 + o), b2 = Ld2<T>::ld(z2 + o);
            const float h0 = (in_round<T>(a2.x) - m1[0]) * q1r[0], h1 = (in_round<T>(a2.y) - m1[1]) * q1r[1];
            Ld2<__half>::st(zh1 + o, h0, h1);
            Ld2<__half>::st(zh2 + o, (in_round<T>(b2.x) - m2[0]) * q2r[0], (in_round<T>(b2.y) - m2[1]) * q2r[1]);
            // [dimension owner q][local row][Dr]: the slice every other rank needs as the A operand of its CORR tiles, contiguous
            if (zh1_blk != nullptr) Ld2<__half>::st(zh1_blk + ((size_t)(col / dr) * n_local + n) * dr + (col % dr), h0, h1);
        }
    }
}

// column sums of C o C (and of C) over all rows of a (rows x cols) fp16 block: the batch-norm backward constants of the
// view-2 gradients when the transposed block arrives by all-to-all (multi-GPU)
__global__ void __launch_bounds__(256) bt_colsq_kernel(const __half* __restrict__ c, int rows, int cols, int rows_per_block,
                                                       float* __restrict__ col_sq, float* __restrict__ col_sum, int hsic) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (j >= cols) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float q0 = 0.f, q1 = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(c + (size_t)r * cols + j));
        q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
        s0 += f.x; s1 += f.y;
    }
    atomicAdd(col_sq + j, q0); atomicAdd(col_sq + j + 1, q1);
    if (hsic) { atomicAdd(col_sum + j, s0); atomicAdd(col_sum + j + 1, s1); }
}

// ------------------------------------------------------------------------------------------
// 3./4. tcgen05 GEMM kernel (persistent, warp specialised)
// ------------------------------------------------------------------------------------------
constexpr int BM = 128;            // UMMA M (TMEM lanes)
constexpr int BK = 64;             // K elements per pipeline stage (one 128-byte swizzle row)
constexpr int kStages = 3;
constexpr int kABytes = BM * BK * 2;       // 16 KiB
constexpr int kBBytesMax = 256 * BK * 2;   // 32 KiB
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kStages2 = 5;                       // CTA-pair kernel: half of B per CTA
constexpr int kStageBytes2 = kABytes + kBBytesMax / 2;
constexpr int kAccCols = 256;      // TMEM columns per accumulator stage
constexpr int kNumThreads = 384;   // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 loss scalar, warps 4-11 epilogue
constexpr int kEpiWarps = 8;
constexpr int kColStatBytes = 2 * 2 * 256 * 4;   // CORR epilogue: column mean / scale of the tile's 256 columns, per accumulator stage
constexpr int kStoreBytes = 8 * 4096;            // CORR epilogue: one 32-row x 64-column fp16 box per epilogue warp, staged for the TMA store
constexpr int kPipeBytes = 5 * (kABytes + kBBytesMax / 2);   // = kStages2 * kStageBytes2 >= kStages * kStageBytes
constexpr int kSmemBytes = kPipeBytes + kStoreBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kColStatBytes;
static_assert(kPipeBytes >= kStages * kStageBytes, "pipeline area too small");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// Shared-memory descriptor constants (bytes).  Overridable through abt_debug_set() so that a wrong
// guess about the descriptor encoding can be diagnosed in one GPU session.
struct DescCfg {
    int mn_lbo, mn_sbo, mn_kstep;   // MN-major tiles: 64-element chunk stride, 8-row (K) group stride, bytes per UMMA_K
    int k_lbo, k_sbo, k_kstep;      // K-major tiles
};
static DescCfg g_desc = {8192, 1024, 2048, 16, 1024, 32};

// One GEMM "pass" of a launch (a launch runs 1 or 2 passes over the same tile grid).
struct PassCfg {
    int a_mn;                 // A operand: 1 = MN-major tiles (CORR: raw z columns; GRAD: C^T), 0 = K-major (GRAD: C rows)
    int row0, row_end;        // global dimension index of A-row 0 of the tile grid, and exclusive end of the valid rows
    // ---- CORR epilogue: c = (S + row_nmu[i] * col_mu[j]) * row_rho[i] * col_r[j]
    const float* row_nmu; const float* row_rho; const float* col_mu; const float* col_r;
    int accumulate_loss;
    float* row_sq;            // += sum_j c_ij^2 (j != i), indexed by the global row; may be null
    float* row_sum;           // HSIC only: += sum_j c_ij
    float* col_sq;            // += sum_i c_ij^2 (i != j), indexed by the global column; null when a second pass provides it
    float* col_sum;           // HSIC only
    int a_col0;               // MN-major A: first column of the A matrix that belongs to tile row 0 (row0 for full matrices, 0 for compact ones)
    int blocked_dr;           // > 0: C is stored column-blocked, element (i, j) at ((j / Dr) * Dr + i) * Dr + j % Dr, so that the block
                              // C[rows, q Dr : (q + 1) Dr] is contiguous (multi-GPU: the blocks are exchanged by an all-to-all)
    // ---- GRAD epilogue (batch-norm backward out of TMEM): rows are dimensions of view `side`
    int side;                 // 0: this pass produces dz1, 1: dz2
    const void* z_self;       // 16-bit embeddings of this view / the other view restricted to the pass's dimensions: element (n, local row)
    const void* z_other;      // at z[n * ld + local row]; raw bf16 (zfmt 0) or standardised fp16 (zfmt 1)
    int ld_self, ld_other;
    void* dz;                 // dz[n * ld_dz + (row - row0)], dtype UmmaParams::io_dtype
    int ld_dz;
    const float* sq;          // row_sq / col_sq accumulated by CORR for these rows (global row index)
    const float* sm;          // HSIC: the matching sums of C
};

struct UmmaParams {
    DescCfg dc;
    int mode;          // 0 = CORR, 1 = GRAD, 2 = LINEAR (CTA pairs only)
    int D, N;
    int tiles_m, tiles_n, kblocks;   // per pass
    int pass_count;
    int bn;            // UMMA N of this launch (multiple of 16, <= 256)
    int split_from;    // GRAD: tiles >= split_from (the last, partial wave) are processed as two half-width work items each, so that
                       // the tail of the persistent schedule costs half a tile time; = tile count when nothing is split
    int ab_format;     // operand type of this launch: 1 = bf16, 0 = fp16
    int hsic;
    int write_c;
    double* loss_acc;
    // GRAD
    int io_dtype;                      // abt_dtype of dz
    int zfmt;                          // format of PassCfg::z_self / z_other: 0 raw bf16, 1 standardised fp16, 2 raw fp16
    const float* stats;                // StatSlot arrays
    const float* rs1; const float* rs2;   // HSIC: row sums of zh1 / zh2 over all dimensions
    float alpha, lambda, grad_scale;
    float* loss_out;                   // written by the GRAD launch (single-GPU); may be null
    // LINEAR (mode 2, the projector tail z = h W^T for both views with the column statistics of the rounded outputs)
    void* lin_z1; void* lin_z2;        // (N, D) bf16 outputs
    float* lin_partials;               // [2 * tiles_n][5][D]: sum z1, sum z1^2, sum z2, sum z2^2, sum z1 z2 over 64 samples each
    PassCfg pass[2];
};

// work item -> (pass, row tile, column tile, half): half = -1 for a full-width tile, 0 / 1 for the halves of a split tile
__device__ __forceinline__ void decode_work(const UmmaParams& p, int w, int& pass, int& tm, int& tn, int& half) {
    int tile = w;
    half = -1;
    if (w >= p.split_from) { tile = p.split_from + ((w - p.split_from) >> 1); half = (w - p.split_from) & 1; }
    const int per_pass = p.tiles_m * p.tiles_n;
    pass = tile / per_pass;
    const int r = tile % per_pass;
    tn = r % p.tiles_n; tm = r / p.tiles_n;
}

__device__ __forceinline__ float loss_from_parts(const double* acc, float alpha, float lambda, int hsic, int D) {
    double off = acc[0];
    if (hsic) off = acc[0] + 2.0 * acc[1] + (double)D * (double)(D - 1);
    return (float)((double)alpha * acc[2] + (double)lambda * off);
}

__global__ void bt_loss_scalar_kernel(const double* __restrict__ acc, float alpha, float lambda, int hsic, int D, float* __restrict__ loss_out) {
    if (threadIdx.x == 0) *loss_out = loss_from_parts(acc, alpha, lambda, hsic, D);
}

template <typename T> __device__ __forceinline__ void store_out(T* p, float v);
template <> __device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void store_out<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }

// GRAD epilogue for one 32-sample chunk of one dimension (row): batch-norm backward and the fp32 diagonal term.
//   d loss / d zh_self[n] = hs * acc[n] + gd * zh_other[n]  (+ HSIC: hs * (R_other[n] - zh_other[n]))
//   dz_self[n] = r_self * (that - zh_self[n] * b) * grad_scale
template <typename T>
__device__ __forceinline__ void grad_chunk(const uint32_t (&acc)[32], const UmmaParams& p, const PassCfg& pc, int row, int lrow, int n_base, int n_valid,
                                           float mu_s, float r_s, float mu_o, float r_o, float hs, float gd, float b) {
    const unsigned short* zs = static_cast<const unsigned short*>(pc.z_self) + lrow;
    const unsigned short* zo = static_cast<const unsigned short*>(pc.z_other) + lrow;
    const float* rso = pc.side == 0 ? p.rs2 : p.rs1;
    T* dz = static_cast<T*>(pc.dz) + lrow;
    const float rg = r_s * p.grad_scale;
    // 16 samples at a time, all loads first (the stores may alias as far as the compiler knows; keeping them
    // behind the loads leaves 32+ requests in flight per thread instead of one)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        unsigned short zs_raw[16], zo_raw[16];
        float rsv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int t = h * 16 + u;
            const size_t n = (size_t)(n_base + (t < n_valid ? t : 0));
            zs_raw[u] = __ldg(zs + n * pc.ld_self);
            zo_raw[u] = __ldg(zo + n * pc.ld_other);
            rsv[u] = p.hsic ? __ldg(rso + n) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int t = h * 16 + u;
            if (t < n_valid) {
                const size_t n = (size_t)(n_base + t);
                float zhs, zho;
                if (p.zfmt == 1) { zhs = __half2float(__ushort_as_half(zs_raw[u])); zho = __half2float(__ushort_as_half(zo_raw[u])); }
                else if (p.zfmt == 2) { zhs = (__half2float(__ushort_as_half(zs_raw[u])) - mu_s) * r_s; zho = (__half2float(__ushort_as_half(zo_raw[u])) - mu_o) * r_o; }
                else { zhs = (__uint_as_float((uint32_t)zs_raw[u] << 16) - mu_s) * r_s; zho = (__uint_as_float((uint32_t)zo_raw[u] << 16) - mu_o) * r_o; }
                float g = fmaf(hs, __uint_as_float(acc[t]), gd * zho);
                if (p.hsic) g = fmaf(hs, rsv[u] - zho, g);
                store_out<T>(dz + n * pc.ld_dz, (g - zhs * b) * rg);
            }
        }
    }
}

// CORR epilogue arithmetic for 32 columns of one row: c_t = (S_t + nmu * mu_t) * rho * r_t, diagonal zeroed, packed to fp16;
// cs_mu / cs_r: shared-window addresses of the 32 column means / scales (broadcast reads)
__device__ __forceinline__ void corr_chunk(const uint32_t (&r)[32], uint32_t cs_mu, uint32_t cs_r, float nmu, float rho, int diag_t, bool diag_here,
                                           bool hsic, uint32_t (&packed)[16], float& l2, float& l1) {
#pragma unroll
    for (int t4 = 0; t4 < 8; ++t4) {
        const float4 m4 = lds128(cs_mu + t4 * 16), b4 = lds128(cs_r + t4 * 16);
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
        float cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cc[u] = (fmaf(nmu, mm[u], __uint_as_float(r[t4 * 4 + u])) * rho) * bb[u];
        if (diag_here) {       // warp-uniform: the diagonal is handled in fp32 (cdiag from the statistics pass)
#pragma unroll
            for (int u = 0; u < 4; ++u) cc[u] = (t4 * 4 + u == diag_t) ? 0.f : cc[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) l2 = fmaf(cc[u], cc[u], l2);
        if (hsic) l1 += (cc[0] + cc[1]) + (cc[2] + cc[3]);
        packed[t4 * 2] = pack_f16x2(cc[0], cc[1]);
        packed[t4 * 2 + 1] = pack_f16x2(cc[2], cc[3]);
    }
}

// CG = 1: one CTA per 128 x bn tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x bn tile -- each CTA stages
// its own 128 rows of A and HALF of the B tile, so the operand bytes pulled through L2 per FLOP drop by a third (the single-CTA
// kernel is L2-bandwidth-bound: 48 KB per 4.2 MFLOP k-block = 87 FLOP/B against ~12 TB/s of L2).  The leader CTA (rank 0) issues
// the M = 256 MMAs; TMA completions of both CTAs land on the leader's `full` barriers, MMA commits are multicast to both CTAs.
template <int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
bt_umma_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapC0, const __grid_constant__ CUtensorMap mapC1, const UmmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int kStg = CG == 1 ? kStages : kStages2;
    constexpr int kStgBytes = CG == 1 ? kStageBytes : kStageBytes2;
    static_assert(kStg * kStgBytes <= kPipeBytes, "pipeline stages exceed their area");
    const uint32_t store_s = smem_u32(smem + kPipeBytes);                                // 1024-byte aligned staging boxes
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kPipeBytes + kStoreBytes);
    uint64_t* empty_bar = full_bar + kStg;
    uint64_t* tfull_bar = empty_bar + kStg;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const uint32_t colstat_s = smem_u32(smem + kPipeBytes + kStoreBytes + 256);          // [acc][mu | r][256] floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_m * p.tiles_n * p.pass_count;     // tiles_m counts (128 * CG)-row tiles
    const int total_work = total_tiles + (total_tiles - p.split_from);  // split tiles count twice
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;          // 0 = leader of the pair
    const int work0 = blockIdx.x / CG, work_stride = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapB0);
        if (p.pass_count > 1) { tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapB1); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStg; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kEpiWarps * CG); }
        mbar_fence_init();
    }
    if (warp == 2) {
        if (CG == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Operand "major-ness": CORR reads both raw embeddings as MN-major tiles (B too); GRAD reads C K-major (rows of C) or
    // MN-major (the same C as C^T) and its B operand (standardised fp16 embeddings, one row per sample) K-major.
    const bool b_mn = (p.mode == 0);
    if (warp == 0) {
        // ================= TMA producer (whole warp runs the loop so that the code stays warp-uniform; one elected lane issues) ==========
        int stage = 0; uint32_t phase = 0;
        for (int w = work0; w < total_work; w += work_stride) {
            int pass, tm, tn, half;
            decode_work(p, w, pass, tm, tn, half);
            tm = tm * CG + (int)rank;                 // this CTA's 128-row tile
            const PassCfg& pc = p.pass[pass];
            const CUtensorMap* mA = pass == 1 ? &mapA1 : &mapA0;
            // half-width items of the GRAD tail use the B tensor maps with half the box rows (passed in the C-map slots)
            const CUtensorMap* mB = half < 0 ? (pass == 1 ? &mapB1 : &mapB0) : (pass == 1 ? &mapC1 : &mapC0);
            const int bn_cta = (half < 0 ? p.bn : p.bn / 2) / CG;               // B columns (or rows) staged by this CTA
            const uint32_t tx = (kABytes + (uint32_t)bn_cta * BK * 2) * CG;     // both CTAs' bytes land on the leader's barrier
            int bcol0 = tn * p.bn + (half > 0 ? p.bn / 2 : 0) + (int)rank * bn_cta;
            if (p.mode == 2) {      // LINEAR: the pair's B tile is [128 samples of view 1 | the same 128 samples of view 2]
                mB = rank ? &mapB1 : &mapB0;
                bcol0 = tn * bn_cta;
            }
            const int a_mn = pc.a_mn;
            const int arow = a_mn ? pc.a_col0 + tm * BM : tm * BM;
            const int bdr = pc.blocked_dr;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sA = smem + stage * kStgBytes;
                    uint8_t* sB = sA + kABytes;
                    if (CG == 1) {
                        mbar_expect_tx(&full_bar[stage], tx);
                        if (a_mn) {
                            // the tensor map spans global dimension indices (columns of z or of the full C)
                            tma_load_2d(sA, mA, &full_bar[stage], arow, kb * BK);
                            tma_load_2d(sA + 8192, mA, &full_bar[stage], arow + 64, kb * BK);
                        } else if (bdr > 0) {
                            const int qb = (kb * BK) / bdr;
                            tma_load_2d(sA, mA, &full_bar[stage], kb * BK - qb * bdr, qb * bdr + arow);
                        } else {
                            // the tensor map spans the rows of the (possibly row-block compact) C matrix
                            tma_load_2d(sA, mA, &full_bar[stage], kb * BK, arow);
                        }
                        if (b_mn) {
                            for (int c = 0; c < bn_cta / 64; ++c) tma_load_2d(sB + c * 8192, mB, &full_bar[stage], bcol0 + c * 64, kb * BK);
                        } else {
                            tma_load_2d(sB, mB, &full_bar[stage], kb * BK, bcol0);
                        }
                    } else {
                        const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);       // the leader's barrier
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
                        if (a_mn) {
                            tma_load_2d_2sm(sA, mA, lbar, arow, kb * BK);
                            tma_load_2d_2sm(sA + 8192, mA, lbar, arow + 64, kb * BK);
                        } else if (bdr > 0) {
                            const int qb = (kb * BK) / bdr;
                            tma_load_2d_2sm(sA, mA, lbar, kb * BK - qb * bdr, qb * bdr + arow);
                        } else {
                            tma_load_2d_2sm(sA, mA, lbar, kb * BK, arow);
                        }
                        if (b_mn) {
                            for (int c = 0; c < bn_cta / 64; ++c) tma_load_2d_2sm(sB + c * 8192, mB, lbar, bcol0 + c * 64, kb * BK);
                        } else {
                            tma_load_2d_2sm(sB, mB, lbar, kb * BK, bcol0);
                        }
                    }
                }
                __syncwarp();
                if (++stage == kStg) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =================
        // descriptor = constant high word (SBO, version, swizzle) + low word (address >> 4 | LBO << 16) that advances per UMMA_K
        const uint32_t hi_mn = ((uint32_t)(p.dc.mn_sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        const uint32_t hi_k = ((uint32_t)(p.dc.k_sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        const uint32_t lo_mn = (((uint32_t)p.dc.mn_lbo >> 4) & 0x3FFF) << 16, lo_k = (((uint32_t)p.dc.k_lbo >> 4) & 0x3FFF) << 16;
        const uint32_t b_hi = b_mn ? hi_mn : hi_k, b_lo = b_mn ? lo_mn : lo_k, b_step = (uint32_t)(b_mn ? p.dc.mn_kstep : p.dc.k_kstep) >> 4;
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = work0; w < total_work; w += work_stride) {
            int pass, tm, tn, half;
            decode_work(p, w, pass, tm, tn, half);
            const bool a_mn = p.pass[pass].a_mn != 0;
            const uint32_t idesc = make_idesc_f16(BM * CG, half < 0 ? p.bn : p.bn / 2, a_mn ? 1 : 0, b_mn ? 1 : 0, p.ab_format);
            const uint32_t a_hi = a_mn ? hi_mn : hi_k, a_lo = a_mn ? lo_mn : lo_k, a_step = (uint32_t)(a_mn ? p.dc.mn_kstep : p.dc.k_kstep) >> 4;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccCols;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sA = smem_u32(smem + stage * kStgBytes);
                    const uint32_t sB = sA + kABytes;
                    uint32_t da = a_lo | ((sA & 0x3FFFF) >> 4), db = b_lo | ((sB & 0x3FFFF) >> 4);
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t da64 = ((uint64_t)a_hi << 32) | da, db64 = ((uint64_t)b_hi << 32) | db;
                        if (CG == 2) umma_bf16_ss_2sm(d_tmem, da64, db64, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        else umma_bf16_ss(d_tmem, da64, db64, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        da += a_step; db += b_step;
                    }
                    // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                    if (CG == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    // accumulator complete -> epilogue (of both CTAs)
                    if (kb == p.kblocks - 1) {
                        if (CG == 2) umma_commit_2sm(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
                    }
                }
                __syncwarp();
                if (++stage == kStg) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 3 && lane == 0) {
        // the loss scalar is complete once CORR has run: the GRAD launch publishes it (single-GPU)
        if (p.mode == 1 && blockIdx.x == 0 && p.loss_out != nullptr) *p.loss_out = loss_from_parts(p.loss_acc, p.alpha, p.lambda, p.hsic, p.D);
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> registers -> global =================
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int hf = (warp - 4) >> 2;      // which half of the 32-column chunks
        int acc = 0; uint32_t acc_phase = 0;
        const int D = p.D;
        for (int w = work0; w < total_work; w += work_stride) {
            int pass, tm, tn, half;
            decode_work(p, w, pass, tm, tn, half);
            tm = tm * CG + (int)rank;
            const PassCfg& pc = p.pass[pass];
            const uint32_t cs_mu = colstat_s + acc * 2048, cs_r = cs_mu + 1024;
            if (p.mode == 0) {
                // stage the column statistics of this tile while its MMAs are still running (the previous use of this buffer, two
                // tiles ago, ended before the named barrier of the previous tile; the barrier orders these writes before the reads)
                const int e = threadIdx.x - 128, j = tn * p.bn + e;
                sts32f(cs_mu + e * 4, j < D ? __ldg(pc.col_mu + j) : 0.f);
                sts32f(cs_r + e * 4, j < D ? __ldg(pc.col_r + j) : 0.f);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * kAccCols + (static_cast<uint32_t>(q * 32) << 16);
            const int lrow = tm * BM + q * 32 + lane;    // row inside this pass's tile grid
            const int row = pc.row0 + lrow;              // global dimension index owned by this thread
            const bool row_ok = row < pc.row_end;
            if (p.mode == 0) {
                // ---- CORR: v = S - N mu_i mu'_j;  c = v (r_i / N) r'_j   (batch-norm as a rank-1 correction)
                // Warp (q, hf) owns rows 32q..32q+31 and columns 128 hf..128 hf+127 of the tile: 4 chunks of 32 columns = 2 boxes
                // of 64 columns.  Each box is packed to fp16, staged in shared memory (128-byte swizzle) and written with ONE
                // TMA store (row-major C with lane = row would otherwise need 32 scattered 16-byte stores per instruction);
                // the column sums of C o C are taken from the staged box.  TMEM loads run one chunk ahead of the arithmetic.
                const float nmu = row_ok ? pc.row_nmu[row] : 0.f;
                const float rho = row_ok ? pc.row_rho[row] : 0.f;      // rows beyond the block give c = 0
                float l2 = 0.f, l1 = 0.f;
                const int row_base = pc.row0 + tm * BM + q * 32;       // global row of lane 0
                const uint32_t stg = store_s + (uint32_t)(warp - 4) * 4096u, stg_row = stg + (uint32_t)lane * 128u;
                const uint32_t sw = (uint32_t)(lane & 7);
                const CUtensorMap* mC = pass == 1 ? &mapC1 : &mapC0;
                const int c_first = hf * 4, j_tile = tn * p.bn;
                int nvalid = (D - j_tile - c_first * 32) / 32;          // valid 32-column chunks of this warp (D % 64 == 0: 0, 2 or 4)
                nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
                uint32_t ra[32], rb[32];
                if (nvalid > 0) tmem_ld_32x32(t_addr + c_first * 32, ra);
                for (int pp = 0; pp * 2 < nvalid; ++pp) {
                    const int chA = c_first + pp * 2, jA = j_tile + chA * 32;
                    uint32_t packed[16];
                    tmem_ld_wait();
                    tmem_ld_32x32(t_addr + (chA + 1) * 32, rb);
                    if (p.write_c) {                                   // the staging box is free once the previous store has read it
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                    }
                    corr_chunk(ra, cs_mu + chA * 128, cs_r + chA * 128, nmu, rho, row - jA, jA < row_base + 32 && jA + 32 > row_base, p.hsic != 0, packed, l2, l1);
                    if (p.write_c) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) sts128(stg_row + (((uint32_t)k ^ sw) << 4), packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                    }
                    tmem_ld_wait();
                    if (pp == 0 && nvalid > 2) tmem_ld_32x32(t_addr + (chA + 2) * 32, ra);
                    corr_chunk(rb, cs_mu + (chA + 1) * 128, cs_r + (chA + 1) * 128, nmu, rho, row - jA - 32, jA + 32 < row_base + 32 && jA + 64 > row_base, p.hsic != 0, packed, l2, l1);
                    if (p.write_c) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) sts128(stg_row + (((uint32_t)(k + 4) ^ sw) << 4), packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0 && tm * BM + q * 32 < pc.row_end - pc.row0) {
                            int sc0 = jA, sc1 = tm * BM + q * 32;
                            if (pc.blocked_dr > 0) { const int qb = jA / pc.blocked_dr; sc0 = jA - qb * pc.blocked_dr; sc1 += qb * pc.blocked_dr; }
                            tma_store_2d(mC, stg, sc0, sc1);       // rows / columns beyond the matrix are clipped by the tensor map
                            tma_store_commit();
                        }
                        if (pc.col_sq != nullptr) {
                            // column sums over this warp's 32 rows from the staged (fp16) box: lane l owns columns jA + 2l, 2l + 1
                            float sq0 = 0.f, sq1 = 0.f;
                            const uint32_t woff = (uint32_t)(lane & 3) * 4u, kc = (uint32_t)(lane >> 2);
                            if (!p.hsic) {
#pragma unroll
                                for (int r = 0; r < 32; ++r) {
                                    const uint32_t w = lds32(stg + (uint32_t)r * 128u + ((kc ^ (uint32_t)(r & 7)) << 4) + woff);
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
                                    sq0 = fmaf(f.x, f.x, sq0); sq1 = fmaf(f.y, f.y, sq1);
                                }
                            } else {
                                float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
                                for (int r = 0; r < 32; ++r) {
                                    const uint32_t w = lds32(stg + (uint32_t)r * 128u + ((kc ^ (uint32_t)(r & 7)) << 4) + woff);
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
                                    sq0 = fmaf(f.x, f.x, sq0); sq1 = fmaf(f.y, f.y, sq1);
                                    sm0 += f.x; sm1 += f.y;
                                }
                                atomicAdd(pc.col_sum + jA + 2 * lane, sm0);
                                atomicAdd(pc.col_sum + jA + 2 * lane + 1, sm1);
                            }
                            atomicAdd(pc.col_sq + jA + 2 * lane, sq0);
                            atomicAdd(pc.col_sq + jA + 2 * lane + 1, sq1);
                        }
                    }
                }
                if (row_ok && pc.row_sq != nullptr) {
                    atomicAdd(pc.row_sq + row, l2);
                    if (p.hsic) atomicAdd(pc.row_sum + row, l1);
                }
                if (pc.accumulate_loss) {
                    l2 = warp_sum(l2);
                    if (p.hsic) l1 = warp_sum(l1);
                    if (lane == 0) {
                        atomicAdd(p.loss_acc + 0, (double)l2);
                        if (p.hsic) atomicAdd(p.loss_acc + 1, (double)l1);
                    }
                }
            } else if (p.mode == 2) {
                // ---- LINEAR: accumulator (output dimension row, sample) -> bf16 z1 / z2 and the column statistics of the ROUNDED
                // outputs.  Columns 0..127 are 128 samples of view 1, columns 128..255 the same samples of view 2, and chunk ch of
                // one view and chunk ch of the other belong to the same warp, so the five sums of a dimension -- the cross term
                // sum z1 z2 included -- are thread-local.  One partial per (sample tile, warp half) = 64 samples, plain stores.
                const int n_tile0 = tn * 128;
                float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f, xx = 0.f;
                __nv_bfloat16* zo1 = static_cast<__nv_bfloat16*>(p.lin_z1) + (row_ok ? row : 0);
                __nv_bfloat16* zo2 = static_cast<__nv_bfloat16*>(p.lin_z2) + (row_ok ? row : 0);
                for (int c = 0; c < 2; ++c) {
                    const int ch = hf + 2 * c, n_base = n_tile0 + ch * 32;
                    if (n_base >= p.N) break;                                  // warp-uniform
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32(t_addr + ch * 32, ra);
                    tmem_ld_32x32(t_addr + 128 + ch * 32, rb);
                    tmem_ld_wait();
                    const int n_valid = row_ok ? min(32, p.N - n_base) : 0;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t < n_valid) {
                            const __nv_bfloat16 a16 = __float2bfloat16_rn(__uint_as_float(ra[t])), b16 = __float2bfloat16_rn(__uint_as_float(rb[t]));
                            const float fa = __bfloat162float(a16), fb = __bfloat162float(b16);
                            s1 += fa; q1 = fmaf(fa, fa, q1); s2 += fb; q2 = fmaf(fb, fb, q2); xx = fmaf(fa, fb, xx);
                            zo1[(size_t)(n_base + t) * D] = a16;
                            zo2[(size_t)(n_base + t) * D] = b16;
                        }
                    }
                }
                if (row_ok) {
                    float* pp = p.lin_partials + (size_t)(tn * 2 + hf) * 5 * D + row;
                    pp[0] = s1; pp[D] = q1; pp[2 * (size_t)D] = s2; pp[3 * (size_t)D] = q2; pp[4 * (size_t)D] = xx;
                }
            } else {
                // ---- GRAD: fp32 accumulator (dimension row, sample n) -> batch-norm backward -> dz[n][row - row0]
                const int rr = row_ok ? row : pc.row0;
                const float invN = 1.0f / (float)p.N;
                const float mu1 = p.stats[S_MU1 * D + rr], r1 = p.stats[S_R1 * D + rr];
                const float mu2 = p.stats[S_MU2 * D + rr], r2 = p.stats[S_R2 * D + rr];
                const float cd = p.stats[S_CDIAG * D + rr];
                const float hs = 2.0f * p.lambda * invN;
                const float gd = 2.0f * p.alpha * (cd - 1.0f) * invN;        // G_ii / N
                float b = fmaf(hs, pc.sq[rr], gd * cd);                        // mean_n(g o zh_self)
                if (p.hsic) b = fmaf(hs, pc.sm[rr], b);
                const float mu_s = pc.side == 0 ? mu1 : mu2, r_s = pc.side == 0 ? r1 : r2;
                const float mu_o = pc.side == 0 ? mu2 : mu1, r_o = pc.side == 0 ? r2 : r1;
                const int bn_i = half < 0 ? p.bn : p.bn / 2, n_off = tn * p.bn + (half > 0 ? p.bn / 2 : 0);
                const int nchunks = (bn_i + 31) / 32;
                for (int ch = hf; ch < nchunks; ch += 2) {
                    const int n_base = n_off + ch * 32;
                    if (n_base >= p.N) break;                                  // warp-uniform
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + ch * 32, r);
                    tmem_ld_wait();
                    const int n_valid = row_ok ? min(32, min(bn_i - ch * 32, p.N - n_base)) : 0;
                    if (p.io_dtype == ABT_DTYPE_BF16) grad_chunk<__nv_bfloat16>(r, p, pc, rr, lrow, n_base, n_valid, mu_s, r_s, mu_o, r_o, hs, gd, b);
                    else if (p.io_dtype == ABT_DTYPE_F16) grad_chunk<__half>(r, p, pc, rr, lrow, n_base, n_valid, mu_s, r_s, mu_o, r_o, hs, gd, b);
                    else grad_chunk<float>(r, p, pc, rr, lrow, n_base, n_valid, mu_s, r_s, mu_o, r_o, hs, gd, b);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // the MMA issuer (leader CTA) may overwrite this accumulator stage once every epilogue warp of the pair is done
                if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    if (warp >= 4 && lane == 0) tma_store_wait_all();     // bulk stores of the CORR epilogue
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        if (CG == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace abt
#include "bt_fused.cuh"
namespace abt {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || sym == nullptr) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 16-bit row-major matrix (rows x cols), box = (box_cols x box_rows) with 128-byte swizzle
static int make_map_16(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                       uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return set_error(ABT_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ABT_ERR_CUDA, "cuTensorMapEncodeTiled failed (code %d)", (int)r);
    return 0;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Optional per-call device timing of the launches (bench.py roofline): events are recorded on the launching
// stream at the start of the call, before CORR, after CORR and after GRAD of every loss call while enabled.
constexpr int kTimingRing = 512;
struct TimingState {
    bool enabled = false;
    int count = 0;
    cudaEvent_t ev[kTimingRing][4];
    bool created = false;
};
static TimingState g_timing;

// Workspace layout.  `rows` = number of C rows this call materialises (D single-GPU, row_count in row-block mode);
// `two_c` = a second C block for the transposed pass (row-block mode).
struct WsLayout {
    size_t misc, acc, counters, stats, partials, keep, pack_local, pack_all, rs1, rs2, zb1, zb2, zh1, zh2, zs1, zh1_blk, c1, c2, total;
    size_t zero_bytes;   // misc + acc: cleared at the start of every call
};

static int g_fused = 1;         // single-GPU, N <= 128: the one-launch kernel of bt_fused.cuh (abt_debug_set key 9: 0 = use CORR + GRAD instead)
static int g_fused_pdl = 1;     // programmatic dependent launch of that kernel behind the statistics kernel (abt_debug_set key 10)
static bool fused_applies(int N, bool rows_mode, int world) { return g_fused != 0 && world == 0 && !rows_mode && N <= FB; }

static WsLayout ws_layout(int N, int D, int rows, int dtype, bool two_c, int world = 0) {
    WsLayout L{};
    size_t off = 0;
    L.misc = off; off += 256;
    L.acc = off; off = align_up(off + sizeof(float) * A_COUNT * (size_t)D, 256);
    L.counters = off; off = align_up(off + sizeof(unsigned int) * (((size_t)D + kStatCols - 1) / kStatCols), 256);     // statistics: arrival counters per column block
    L.zero_bytes = off;
    L.stats = off; off = align_up(off + sizeof(float) * S_COUNT * (size_t)D, 256);
    L.partials = off; off = align_up(off + sizeof(float) * 5 * kMaxRowSplits * (size_t)D, 256);
    L.keep = off; off += 256;                                   // multi-GPU: on-diagonal loss sum kept between the calls of one step
    L.pack_local = off; if (world > 0) off = align_up(off + sizeof(float) * 7 * (size_t)D, 256);
    L.pack_all = off; if (world > 0) off = align_up(off + sizeof(float) * 7 * (size_t)D * world, 256);
    L.rs1 = off; off = align_up(off + sizeof(float) * (size_t)N, 256);
    L.rs2 = off; off = align_up(off + sizeof(float) * (size_t)N, 256);
    L.zb1 = off; if (dtype == ABT_DTYPE_F32) off = align_up(off + 2 * (size_t)N * D, 256);       // bf16 copies of fp32 embeddings
    L.zb2 = off; if (dtype == ABT_DTYPE_F32) off = align_up(off + 2 * (size_t)N * D, 256);
    // one-launch path: tile images of (N rounded up to 32) rows x 64 columns, plus one spare (zero) tile
    const size_t zh_rows = fused_applies(N, two_c, world) ? (size_t)((N + 31) / 32 * 32) : (size_t)N;
    const size_t zh_extra = fused_applies(N, two_c, world) ? zh_rows * 128 : 0;
    L.zh1 = off; off = align_up(off + 2 * zh_rows * D + zh_extra, 256);
    L.zh2 = off; off = align_up(off + 2 * zh_rows * D + zh_extra, 256);
    L.zs1 = off; if (world > 0) off = align_up(off + 2 * (size_t)N * rows, 256);          // (N_g, Dr): view-1 columns of this rank's dimensions, all samples
    L.zh1_blk = off; if (world > 0) off = align_up(off + 2 * (size_t)(N / world) * D, 256);  // (world, n_local, Dr): send buffer of that exchange
    L.c1 = off; if (!fused_applies(N, two_c, world)) off = align_up(off + 2 * (size_t)rows * D, 256);     // the fused kernel keeps C on chip
    L.c2 = off; if (two_c) off = align_up(off + 2 * (size_t)rows * D, 256);
    L.total = off;
    return L;
}

static int num_sms() {
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
        if (cached[dev] <= 0) cached[dev] = 148;
    }
    return cached[dev];
}

// ABT_DEBUG_SYNC=1: synchronise after every stage of a loss evaluation and report which one failed
static int debug_sync(cudaStream_t st, const char* stage) {
    static const bool on = std::getenv("ABT_DEBUG_SYNC") != nullptr;
    if (!on) return 0;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "stage '%s' failed: %s", stage, cudaGetErrorString(e));
    return 0;
}

static int g_cta_group = 2;     // 2 = CTA-pair kernel (default), 1 = single-CTA kernel (abt_debug_set key 6)
static int g_reserve_sms = 0;   // SMs the persistent tensor-core kernels leave free (multi-GPU: for NCCL's kernels); abt_debug_set key 12
static int g_dist_reserve_sms = 12;   // ... the value abt_bt_dist_step uses for its own launches (key 14): 8 + 4 CTAs of the two communicators
static int g_comm_max_ctas = 0; // CTA cap of the private NCCL communicators (0 = NCCL's default); abt_debug_set key 13, read at abt_comm_create.
                                // Measured at 2 ranks: a cap of 8 / 16 CTAs makes the step 35 % / 27 % SLOWER (the all-to-alls starve), so the default is no cap
static bool g_tail_split = true; // GRAD: split the last partial wave into half-width items (abt_debug_set key 8)
static int g_dist_xchg = -1;    // multi-GPU exchange schedule: -1 = auto (4 ranks and more), 0 = never, 1 = whenever possible (abt_debug_set key 7)

template <typename T, bool H, int NP> static cudaError_t fused_attr_one() {
    return cudaFuncSetAttribute(bt_fused_kernel<T, H, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmemBytes);
}
template <typename T> static cudaError_t fused_attr_type() {
    cudaError_t e = fused_attr_one<T, false, 128>();
    if (e == cudaSuccess) e = fused_attr_one<T, true, 128>();
    if (e == cudaSuccess) e = fused_attr_one<T, false, 0>();
    if (e == cudaSuccess) e = fused_attr_one<T, true, 0>();
    return e;
}
static cudaError_t fused_set_attr() {
    cudaError_t e = fused_attr_type<__nv_bfloat16>();
    if (e == cudaSuccess) e = fused_attr_type<__half>();
    if (e == cudaSuccess) e = fused_attr_type<float>();
    return e;
}
template <typename T, bool H, int NP> static cudaError_t fused_launch_one(const cudaLaunchConfig_t& cfg, const FusedParams& p) {
    return cudaLaunchKernelEx(&cfg, bt_fused_kernel<T, H, NP>, p);
}
template <typename T> static cudaError_t fused_launch(const cudaLaunchConfig_t& cfg, const FusedParams& p) {
    if (p.n_pad == 128) return p.hsic ? fused_launch_one<T, true, 128>(cfg, p) : fused_launch_one<T, false, 128>(cfg, p);
    return p.hsic ? fused_launch_one<T, true, 0>(cfg, p) : fused_launch_one<T, false, 0>(cfg, p);
}

// cudaFuncSetAttribute is per device: keep one flag per device ordinal
static int ensure_umma_attr() {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(bt_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bt_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = fused_set_attr();
        if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    return 0;
}

// p.tiles_m counts (128 * cg)-row tiles
static int launch_umma(int cg, const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1, const CUtensorMap& c0,
                       const CUtensorMap& c1, const UmmaParams& p, cudaStream_t stream) {
    const int tiles = p.tiles_m * p.tiles_n * p.pass_count;
    const int total = tiles + (p.split_from < tiles ? tiles - p.split_from : 0);
    int avail = num_sms() - g_reserve_sms;
    if (avail < 2 * cg) avail = 2 * cg;
    const int slots = avail / cg;
    const int grid = (total < slots ? total : slots) * cg;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cg == 2 ? cudaLaunchKernelEx(&cfg, bt_umma_kernel<2>, a0, b0, a1, b1, c0, c1, p) : cudaLaunchKernelEx(&cfg, bt_umma_kernel<1>, a0, b0, a1, b1, c0, c1, p);
    count_launch();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "bt_umma_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// row chunks of the statistics kernels: about 64 rows per block, but at least two blocks per SM over the whole grid
static int stat_splits(int N, int D) {
    const int col_blocks = (D + kStatCols - 1) / kStatCols;
    int splits = (N + 127) / 128;
    const int want = (2 * num_sms() + col_blocks - 1) / col_blocks;
    if (splits < want) splits = want;
    if (splits > (N + 7) / 8) splits = (N + 7) / 8;
    if (splits > kMaxRowSplits) splits = kMaxRowSplits;
    return splits < 1 ? 1 : splits;
}

// Everything one loss evaluation needs, for both the single-GPU and the row-block entry points.
struct LossCall {
    const void* z1; const void* z2; int dtype;
    int N, D;
    int row_begin, row_count;   // dimensions owned by this call (0, D single-GPU)
    bool rows_mode;             // compute the C^T row block with a second CORR pass instead of reading C transposed
    bool zh_mode;               // z1 / z2 are standardised fp16 embeddings and the statistics are already in the workspace (multi-GPU)
    int phase;                  // 0 = whole evaluation; 1 = everything except the dz2 GRAD pass; 2 = only the dz2 GRAD pass (after a phase-1 call);
                                // exchange mode: bit mask 8 | F_FRONT (1) | F_DZ1 (2) | F_DZ2 (4)
    bool xchg;                  // multi-GPU exchange mode: A of CORR = the (N, rows) view-1 slice at ws.zs1, C stored column-blocked, no second CORR
                                // pass: the transposed block C[:, rows] (ws.c2) arrives by all-to-all before the dz2 pass
    float alpha, lambda; int hsic; float eps, momentum, grad_scale; int need;
    float* loss_out;            // single-GPU only
    double* loss_parts_out;     // row-block mode: 3 doubles copied out (off-diag sum c^2, off-diag sum c, on-diag sum)
    void* dz1; void* dz2; int ld_dz;
    float* running_mean; float* running_var;
    void* workspace;
};

template <typename T>
static int run_loss(const LossCall& a, const WsLayout& L, cudaStream_t stream) {
    const int N = a.N, D = a.D, R0 = a.row_begin, RC = a.row_count;
    uint8_t* ws = static_cast<uint8_t*>(a.workspace);
    float* stats = reinterpret_cast<float*>(ws + L.stats);
    float* accs = reinterpret_cast<float*>(ws + L.acc);
    float* partials = reinterpret_cast<float*>(ws + L.partials);
    double* loss_acc = reinterpret_cast<double*>(ws + L.misc);
    float* rs1 = reinterpret_cast<float*>(ws + L.rs1);
    float* rs2 = reinterpret_cast<float*>(ws + L.rs2);
    __half* C1 = reinterpret_cast<__half*>(ws + L.c1);
    __half* C2 = reinterpret_cast<__half*>(ws + L.c2);
    __half* zh1 = reinterpret_cast<__half*>(ws + L.zh1);
    __half* zh2 = reinterpret_cast<__half*>(ws + L.zh2);
    // 16-bit embeddings feed the tensor cores as they are (bf16 or fp16, kind::f16 takes both); fp32 ones get a bf16 copy
    const bool is_16 = (a.dtype != ABT_DTYPE_F32);
    const bool raw_f16 = (a.dtype == ABT_DTYPE_F16) && !a.zh_mode;
    __nv_bfloat16* zb1 = is_16 ? nullptr : reinterpret_cast<__nv_bfloat16*>(ws + L.zb1);
    __nv_bfloat16* zb2 = is_16 ? nullptr : reinterpret_cast<__nv_bfloat16*>(ws + L.zb2);
    const void* zq1 = is_16 ? a.z1 : static_cast<const void*>(zb1);
    const void* zq2 = is_16 ? a.z2 : static_cast<const void*>(zb2);
    const int need = a.need & 3;
    const int col_blocks = (D + kColsPerBlock - 1) / kColsPerBlock;
    if (int rc = ensure_umma_attr()) return rc;

    const bool timed = g_timing.enabled && g_timing.count < kTimingRing;
    cudaEvent_t* tev = timed ? g_timing.ev[g_timing.count] : nullptr;
    if (timed) cudaEventRecord(tev[0], stream);
    // which parts of the evaluation this call runs: statistics hand-over + CORR (1), dz1 GEMM (2), dz2 GEMM (4)
    const int pmask = a.xchg ? (a.phase & 7) : (a.phase == 0 ? 7 : (a.phase == 1 ? 3 : 4));
    const bool front = (pmask & 1) != 0;
    __half* ZS1 = reinterpret_cast<__half*>(ws + L.zs1);
    if (fused_applies(N, a.rows_mode, a.zh_mode ? 1 : 0) && need != 0) {
        // ---- small batch: statistics + standardisation in one pass, then ONE tensor-core launch
        //      (S tiles -> loss + fp16 P on chip -> gradient accumulators in TMEM -> batch-norm backward); no memset, no D x D matrix
        float* ondiag_part = partials;
        const int n_pad = (N + 31) / 32 * 32;
        bt_stat_norm_small_kernel<T><<<D / kSmallBlockCols, kColThreads, 0, stream>>>(static_cast<const T*>(a.z1), static_cast<const T*>(a.z2), N, D, a.eps,
                                                                                 a.momentum, stats, a.running_mean, a.running_var, zh1, zh2, ondiag_part,
                                                                                 loss_acc, reinterpret_cast<unsigned int*>(ws + L.misc + 64), n_pad);
        count_launch();
        if (int rc = debug_sync(stream, "statistics")) return rc;
        if (a.hsic) {
            bt_rowsum_img_kernel<<<N, 256, 0, stream>>>(zh1, D / kSmallCols, n_pad, rs1);
            bt_rowsum_img_kernel<<<N, 256, 0, stream>>>(zh2, D / kSmallCols, n_pad, rs2);
            count_launch(2);
        }
        if (timed) { cudaEventRecord(tev[1], stream); cudaEventRecord(tev[2], stream); }
        FusedParams p{};
        p.D = D; p.N = N; p.n_pad = n_pad;
        p.n_blocks = (D + FB - 1) / FB;
        p.pass_count = need == 3 ? 2 : 1;
        p.pass_side[0] = (need & 1) ? 0 : 1; p.pass_side[1] = 1;
        p.hsic = a.hsic;
        p.alpha = a.alpha; p.lambda = a.lambda; p.grad_scale = a.grad_scale;
        p.stats = stats; p.rs1 = rs1; p.rs2 = rs2;
        p.dz1 = a.dz1; p.dz2 = a.dz2;
        p.loss_acc = loss_acc; p.done_counter = reinterpret_cast<unsigned int*>(ws + L.misc + 64); p.loss_out = a.loss_out;
        p.ondiag_part = ondiag_part; p.n_parts = D / kSmallBlockCols;
        p.zimg1 = zh1; p.zimg2 = zh2;
        const int units = p.n_blocks * p.pass_count;
        // programmatic dependent launch: set-up (barriers, TMEM allocation) overlaps the statistics kernel.  Not with HSIC (two more
        // kernels in between) and not while per-launch events are being recorded.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(units < num_sms() ? units : num_sms()); cfg.blockDim = dim3(kTThreads); cfg.dynamicSmemBytes = kXSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        // (not inside a stream capture either: a graph replay has no launch gap to hide, and the captured graph stays free of programmatic edges)
        cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap_status) != cudaSuccess) cap_status = cudaStreamCaptureStatusNone;
        cfg.numAttrs = (!a.hsic && !timed && g_fused_pdl && cap_status == cudaStreamCaptureStatusNone) ? 1 : 0;
        const cudaError_t le = fused_launch<T>(cfg, p);
        if (le != cudaSuccess) return set_error(ABT_ERR_CUDA, "bt_fused_kernel launch: %s", cudaGetErrorString(le));
        count_launch();
        if (int rc = debug_sync(stream, "FUSED")) return rc;
        if (timed) { cudaEventRecord(tev[3], stream); ++g_timing.count; }
        cudaError_t fe = cudaGetLastError();
        if (fe != cudaSuccess) return set_error(ABT_ERR_CUDA, "bt_fused_kernel launch: %s", cudaGetErrorString(fe));
        return 0;
    }
    if (front) cudaMemsetAsync(ws + L.misc, 0, L.zero_bytes, stream);     // loss partial sums + row / column accumulators
    // ---- statistics
    if (!front) {
    } else if (!a.zh_mode) {
        int splits = stat_splits(N, D);
        const int chunk = (N + splits - 1) / splits;
        splits = (N + chunk - 1) / chunk;
        const dim3 cgrid((D + kStatCols - 1) / kStatCols, splits);
        bt_colstat_kernel<T><<<cgrid, kColThreads, 0, stream>>>(static_cast<const T*>(a.z1), static_cast<const T*>(a.z2), N, D, chunk, partials, zb1, zb2,
                                                                reinterpret_cast<unsigned int*>(ws + L.counters), a.eps, a.momentum, stats,
                                                                a.running_mean, a.running_var, loss_acc);
        if (need != 0) {
            int nsplits = (N + 63) / 64;                           // elementwise: 64 rows x 256 columns per block
            if (nsplits > 64) nsplits = 64;
            const int nchunk = (N + nsplits - 1) / nsplits;
            const dim3 ngrid((D + kWideCols - 1) / kWideCols, (N + nchunk - 1) / nchunk);
            bt_normalize_kernel<T><<<ngrid, kColThreads, 0, stream>>>(static_cast<const T*>(a.z1), static_cast<const T*>(a.z2), N, D, nchunk, stats, zh1, zh2);
        }
        count_launch(2);
        if (int rc = debug_sync(stream, "statistics")) return rc;
        if (a.hsic && need != 0) {
            if (raw_f16) {
                bt_rowsum_kernel<__half><<<N, 256, 0, stream>>>(static_cast<const __half*>(zq1), N, D, stats + S_MU1 * D, stats + S_R1 * D, rs1);
                bt_rowsum_kernel<__half><<<N, 256, 0, stream>>>(static_cast<const __half*>(zq2), N, D, stats + S_MU2 * D, stats + S_R2 * D, rs2);
            } else {
                bt_rowsum_kernel<__nv_bfloat16><<<N, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(zq1), N, D, stats + S_MU1 * D, stats + S_R1 * D, rs1);
                bt_rowsum_kernel<__nv_bfloat16><<<N, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(zq2), N, D, stats + S_MU2 * D, stats + S_R2 * D, rs2);
            }
            count_launch(2);
        }
    } else {
        // the statistics (and the on-diagonal loss sum) were produced by abt_bt_dist_normalize into this workspace
        cudaMemcpyAsync(loss_acc + 2, ws + L.keep, sizeof(double), cudaMemcpyDeviceToDevice, stream);
        if (a.hsic && need != 0) {
            if (!a.xchg) bt_rowsum_zh_kernel<<<N, 256, 0, stream>>>(static_cast<const __half*>(a.z1), N, D, rs1);
            bt_rowsum_zh_kernel<<<N, 256, 0, stream>>>(static_cast<const __half*>(a.z2), N, D, rs2);
            count_launch(2);
        }
    }
    if (a.xchg && a.hsic && (pmask & 4) && (need & 2)) {       // view 1 is complete only now
        bt_rowsum_zh_kernel<<<N, 256, 0, stream>>>(static_cast<const __half*>(a.z1), N, D, rs1);
        count_launch();
    }

    const int cg = g_cta_group == 1 ? 1 : 2;
    const int row_tiles = (RC + BM * cg - 1) / (BM * cg);
    // ---- CORR: C[rows, :] (and, in row-block mode, C^T[rows, :] with the views swapped)
    if (int rc = debug_sync(stream, "row sums (HSIC)")) return rc;
    if (timed) cudaEventRecord(tev[1], stream);
    if (front) {
        CUtensorMap m1, m2;
        const CUtensorMapDataType odt = (a.zh_mode || raw_f16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        if (a.xchg) { if (int rc = make_map_16(&m1, odt, ZS1, N, RC, 64, 64)) return rc; }
        else if (int rc = make_map_16(&m1, odt, a.zh_mode ? a.z1 : zq1, N, D, 64, 64)) return rc;
        if (int rc = make_map_16(&m2, odt, a.zh_mode ? a.z2 : zq2, N, D, 64, 64)) return rc;
        UmmaParams p{};
        p.dc = g_desc;
        p.mode = 0; p.D = D; p.N = N;
        p.bn = 256; p.ab_format = (a.zh_mode || raw_f16) ? 0 : 1;
        p.tiles_m = row_tiles; p.tiles_n = (D + p.bn - 1) / p.bn;
        p.kblocks = (N + BK - 1) / BK;
        p.split_from = 1 << 30;                                 // set below once the pass count is known
        p.hsic = a.hsic; p.write_c = need != 0;
        p.loss_acc = loss_acc;
        const bool second = a.rows_mode && !a.xchg && (need & 2);       // the transposed block is only needed for dz2
        PassCfg c0{}, c1{};
        c0.a_mn = 1; c0.row0 = R0; c0.row_end = R0 + RC; c0.a_col0 = a.xchg ? 0 : R0; c0.blocked_dr = a.xchg ? RC : 0;
        c0.row_nmu = stats + S_NMU1 * D; c0.row_rho = stats + S_RHO1 * D; c0.col_mu = stats + S_MU2 * D; c0.col_r = stats + S_R2 * D;
        if (a.zh_mode) { c0.row_nmu = stats + S_ZERO * D; c0.row_rho = stats + S_INVN * D; c0.col_mu = stats + S_ZERO * D; c0.col_r = stats + S_ONE * D; }
        c0.accumulate_loss = 1;                       // C1 / C2 are written through the TMA store maps (mc1 / mc2)
        c0.row_sq = (need & 1) ? accs + A_SQ1 * D : nullptr; c0.row_sum = accs + A_SUM1 * D;
        // (exchange mode: the column sums come from bt_colsq_kernel on the received block)
        c0.col_sq = (!a.rows_mode && (need & 2)) ? accs + A_SQ2 * D : nullptr; c0.col_sum = accs + A_SUM2 * D;
        c1 = c0;
        c1.row_nmu = stats + S_NMU2 * D; c1.row_rho = stats + S_RHO2 * D; c1.col_mu = stats + S_MU1 * D; c1.col_r = stats + S_R1 * D;
        if (a.zh_mode) { c1.row_nmu = c0.row_nmu; c1.row_rho = c0.row_rho; c1.col_mu = c0.col_mu; c1.col_r = c0.col_r; }
        c1.accumulate_loss = 0;
        c1.row_sq = accs + A_SQ2 * D; c1.row_sum = accs + A_SUM2 * D; c1.col_sq = nullptr; c1.col_sum = nullptr;
        p.pass[0] = c0; p.pass[1] = c1;
        p.pass_count = second ? 2 : 1;
        p.split_from = p.tiles_m * p.tiles_n * p.pass_count;    // CORR: no split (its tiles are short)
        // C is written by TMA stores of 32-row x 64-column boxes (row-block compact: RC rows)
        CUtensorMap mc1, mc2;
        if (a.xchg) { if (int rc = make_map_16(&mc1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, D, RC, 64, 32)) return rc; }      // column-blocked: (D, RC)
        else if (int rc = make_map_16(&mc1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, RC, D, 64, 32)) return rc;
        if (int rc = make_map_16(&mc2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, second ? C2 : C1, RC, D, 64, 32)) return rc;
        if (int rc = launch_umma(cg, m1, m2, m2, m1, mc1, mc2, p, stream)) return rc;
    }
    if (int rc = debug_sync(stream, "CORR")) return rc;
    if (timed) cudaEventRecord(tev[2], stream);
    // ---- GRAD (+ batch-norm backward epilogue)
    const int gneed = ((pmask & 2) ? (need & 1) : 0) | ((pmask & 4) ? (need & 2) : 0);     // gradient passes of THIS call
    if (gneed != 0) {
        const int passes = (gneed == 3) ? 2 : 1;
        const int q = 16 * cg;                                  // UMMA N granularity (M = 128: 16, M = 256: 32 so that each CTA stages a multiple of 16)
        int bn = N >= 256 ? 256 : ((N + q - 1) / q) * q;
        // small problems: narrower sample tiles instead of split-K, so that the epilogue always sees complete sums
        while (bn > 32 && row_tiles * ((N + bn - 1) / bn) * passes < (num_sms() - g_reserve_sms) / cg) bn = ((bn / 2 + q - 1) / q) * q;
        CUtensorMap mCk, mCt, mZ2, mZ1;
        if (a.xchg) {
            if (int rc = make_map_16(&mCk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, D, RC, 64, 128)) return rc;         // column-blocked C[rows, :]
            if (int rc = make_map_16(&mCt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C2, D, RC, 64, 64)) return rc;          // received C[:, rows], read MN-major
            if (gneed & 2) {
                const dim3 cgrid((RC / 2 + 255) / 256, 32);
                bt_colsq_kernel<<<cgrid, 256, 0, stream>>>(C2, D, RC, (D + 31) / 32, accs + A_SQ2 * D + R0, accs + A_SUM2 * D + R0, a.hsic);
                count_launch();
            }
        } else if (int rc = make_map_16(&mCk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, RC, D, 64, 128)) return rc;
        if (a.xchg) {
        } else if (a.rows_mode) {
            if (int rc = make_map_16(&mCt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C2, RC, D, 64, 128)) return rc;    // K-major rows of C^T
        } else {
            if (int rc = make_map_16(&mCt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, RC, D, 64, 64)) return rc;     // the same C read MN-major
        }
        const void* zB1 = a.zh_mode ? a.z1 : static_cast<const void*>(zh1);
        const void* zB2 = a.zh_mode ? a.z2 : static_cast<const void*>(zh2);
        if (int rc = make_map_16(&mZ2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, zB2, N, D, 64, bn / cg)) return rc;     // each CTA of a pair stages bn / 2 samples
        if (int rc = make_map_16(&mZ1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, zB1, N, D, 64, bn / cg)) return rc;
        UmmaParams p{};
        p.dc = g_desc;
        p.mode = 1; p.D = D; p.N = N; p.bn = bn; p.ab_format = 0;
        p.tiles_m = row_tiles; p.tiles_n = (N + bn - 1) / bn;
        p.kblocks = (D + BK - 1) / BK;
        p.hsic = a.hsic; p.write_c = 0;
        p.loss_acc = loss_acc;
        p.io_dtype = a.dtype; p.stats = stats; p.rs1 = rs1; p.rs2 = rs2;
        p.zfmt = a.zh_mode ? 1 : (raw_f16 ? 2 : 0);
        const uint16_t* Z1 = static_cast<const uint16_t*>(a.zh_mode ? a.z1 : zq1);      // epilogue reads (16-bit elements)
        const uint16_t* Z2 = static_cast<const uint16_t*>(a.zh_mode ? a.z2 : zq2);
        p.alpha = a.alpha; p.lambda = a.lambda; p.grad_scale = a.grad_scale; p.loss_out = a.loss_out;
        PassCfg d1{}, d2{};
        d1.a_mn = 0; d1.row0 = R0; d1.row_end = R0 + RC; d1.side = 0; d1.dz = a.dz1; d1.ld_dz = a.ld_dz;
        d1.sq = accs + A_SQ1 * D; d1.sm = accs + A_SUM1 * D;
        d1.a_col0 = R0; d1.blocked_dr = a.xchg ? RC : 0;
        d1.z_self = Z1 + R0; d1.ld_self = D; d1.z_other = Z2 + R0; d1.ld_other = D;
        if (a.xchg) { d1.z_self = ZS1; d1.ld_self = RC; }          // view 1 may still be in flight: use the exchanged slice
        d2 = d1;
        d2.a_mn = (a.rows_mode && !a.xchg) ? 0 : 1; d2.side = 1; d2.dz = a.dz2; d2.blocked_dr = 0;
        if (a.xchg) d2.a_col0 = 0;                                   // C[:, rows] is compact: (D, RC)
        d2.sq = accs + A_SQ2 * D; d2.sm = accs + A_SUM2 * D;
        d2.z_self = Z2 + R0; d2.ld_self = D; d2.z_other = Z1 + R0; d2.ld_other = D;
        p.pass_count = passes;
        // tail of the persistent schedule: if the last wave fills at most half of the CTA slots, its tiles are split in two
        // half-width items (one more half wave instead of one more full wave)
        const int n_tiles = p.tiles_m * p.tiles_n * passes, slots = (num_sms() - g_reserve_sms) / cg, rem = n_tiles % slots;
        const bool split = g_tail_split && bn == 256 && n_tiles > slots && rem > 0 && 2 * rem <= slots;
        p.split_from = split ? n_tiles - rem : n_tiles;
        CUtensorMap mZ2h, mZ1h;
        if (int rc = make_map_16(&mZ2h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, zB2, N, D, 64, bn / 2 / cg)) return rc;
        if (int rc = make_map_16(&mZ1h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, zB1, N, D, 64, bn / 2 / cg)) return rc;
        const CUtensorMap *a0, *b0, *a1, *b1;
        if (gneed & 1) { p.pass[0] = d1; a0 = &mCk; b0 = &mZ2; p.pass[1] = d2; a1 = &mCt; b1 = &mZ1; }
        else { p.pass[0] = d2; a0 = &mCt; b0 = &mZ1; p.pass[1] = d2; a1 = &mCt; b1 = &mZ1; }
        const CUtensorMap* h0 = (b0 == &mZ2) ? &mZ2h : &mZ1h;
        const CUtensorMap* h1 = (b1 == &mZ2) ? &mZ2h : &mZ1h;
        if (int rc = launch_umma(cg, *a0, *b0, *a1, *b1, *h0, *h1, p, stream)) return rc;
    } else if (a.loss_out != nullptr && need == 0) {
        bt_loss_scalar_kernel<<<1, 32, 0, stream>>>(loss_acc, a.alpha, a.lambda, a.hsic, D, a.loss_out);
        count_launch();
    }
    if (int rc = debug_sync(stream, "GRAD")) return rc;
    if (timed) { cudaEventRecord(tev[3], stream); ++g_timing.count; }
    if (a.loss_parts_out != nullptr && front) cudaMemcpyAsync(a.loss_parts_out, loss_acc, 3 * sizeof(double), cudaMemcpyDeviceToDevice, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "bt loss launch: %s", cudaGetErrorString(e));
    return 0;
}

static int dispatch(const LossCall& c, const WsLayout& L, cudaStream_t s) {
    switch (c.dtype) {
        case ABT_DTYPE_BF16: return run_loss<__nv_bfloat16>(c, L, s);
        case ABT_DTYPE_F16: return run_loss<__half>(c, L, s);
        case ABT_DTYPE_F32: return run_loss<float>(c, L, s);
        default: return set_error(ABT_ERR_ARG, "unknown dtype %d", c.dtype);
    }
}

static int check_shape(int n_rows, int n_dims, int dtype) {
    if (n_rows < 2 || n_dims < 64 || (n_dims % 64) != 0)
        return set_error(ABT_ERR_ARG, "need n_rows >= 2 and n_dims a multiple of 64 (got %d x %d)", n_rows, n_dims);
    if (dtype < 0 || dtype > 2) return set_error(ABT_ERR_ARG, "unknown dtype %d", dtype);
    return 0;
}

}  // namespace abt

using namespace abt;

extern "C" int abt_debug_timing(int enable) {
    if (enable && !g_timing.created) {
        for (int i = 0; i < kTimingRing; ++i)
            for (int k = 0; k < 4; ++k)
                if (cudaEventCreate(&g_timing.ev[i][k]) != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaEventCreate failed");
        g_timing.created = true;
    }
    g_timing.enabled = enable != 0;
    g_timing.count = 0;
    return 0;
}

// Average milliseconds of the statistics launches, the CORR launch and the GRAD launch recorded since
// abt_debug_timing(1); synchronises on the events.  Any output pointer may be null.
extern "C" int abt_debug_timing_read(float* stats_ms, float* corr_ms, float* grad_ms, int* n_calls) {
    double s = 0, c = 0, g = 0;
    const int n = g_timing.count;
    for (int i = 0; i < n; ++i) {
        float x = 0, y = 0, z = 0;
        if (cudaEventSynchronize(g_timing.ev[i][3]) != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaEventSynchronize failed");
        cudaEventElapsedTime(&x, g_timing.ev[i][0], g_timing.ev[i][1]);
        cudaEventElapsedTime(&y, g_timing.ev[i][1], g_timing.ev[i][2]);
        cudaEventElapsedTime(&z, g_timing.ev[i][2], g_timing.ev[i][3]);
        s += x; c += y; g += z;
    }
    if (stats_ms) *stats_ms = n ? (float)(s / n) : 0.f;
    if (corr_ms) *corr_ms = n ? (float)(c / n) : 0.f;
    if (grad_ms) *grad_ms = n ? (float)(g / n) : 0.f;
    if (n_calls) *n_calls = n;
    return 0;
}

extern "C" int abt_set_reserved_sms(int n_sms) {
    const int prev = g_reserve_sms;
    g_reserve_sms = n_sms < 0 ? 0 : (n_sms > 64 ? 64 : n_sms);
    return prev;
}

extern "C" int abt_debug_set(int key, int value) {
    int* f[6] = {&g_desc.mn_lbo, &g_desc.mn_sbo, &g_desc.mn_kstep, &g_desc.k_lbo, &g_desc.k_sbo, &g_desc.k_kstep};
    if (key == 6) { g_cta_group = value == 1 ? 1 : 2; return 0; }
    if (key == 8) { g_tail_split = value != 0; return 0; }
    if (key == 7) { g_dist_xchg = value < 0 ? -1 : (value != 0 ? 1 : 0); return 0; }
    if (key == 12) { g_reserve_sms = value < 0 ? 0 : (value > 64 ? 64 : value); return 0; }
    if (key == 13) { g_comm_max_ctas = value < 0 ? 0 : value; return 0; }
    if (key == 14) { g_dist_reserve_sms = value < 0 ? 0 : (value > 64 ? 64 : value); return 0; }
    if (key == 9) { g_fused = value != 0; return 0; }
    if (key == 15) { g_gather_blocks = value < 1 ? 1 : (value > 148 ? 148 : value); return 0; }
    if (key == 10) { g_fused_pdl = value != 0; return 0; }
    if (key < 0 || key >= 6) return set_error(ABT_ERR_ARG, "unknown debug key %d", key);
    *f[key] = value;
    return 0;
}

// Debug view of the single-GPU workspace layout (byte offsets), used by tools/gpu_diag.py only.
extern "C" int abt_debug_ws_offsets(int n_rows, int n_dims, int dtype, size_t* out8) {
    const WsLayout L = ws_layout(n_rows, n_dims, n_dims, dtype, false);
    out8[0] = L.stats; out8[1] = L.c1; out8[2] = L.acc; out8[3] = L.zh1; out8[4] = L.zb1; out8[5] = L.zb2; out8[6] = L.misc; out8[7] = L.total;
    return 0;
}

extern "C" int abt_bt_workspace_bytes(int n_rows, int n_dims, int dtype, size_t* bytes) {
    if (bytes == nullptr) return set_error(ABT_ERR_ARG, "bytes is null");
    if (int rc = check_shape(n_rows, n_dims, dtype)) return rc;
    *bytes = ws_layout(n_rows, n_dims, n_dims, dtype, false).total;
    return 0;
}

extern "C" int abt_bt_loss_fwd_bwd(const abt_bt_args* a, abt_stream_t stream) {
    if (a == nullptr) return set_error(ABT_ERR_ARG, "args is null");
    if (int rc = check_shape(a->n_rows, a->n_dims, a->dtype)) return rc;
    if (a->z1 == nullptr || a->z2 == nullptr || a->loss_out == nullptr || a->workspace == nullptr) return set_error(ABT_ERR_ARG, "null pointer argument");
    if ((a->need_grad_mask & 1) && a->dz1 == nullptr) return set_error(ABT_ERR_ARG, "dz1 is null but requested");
    if ((a->need_grad_mask & 2) && a->dz2 == nullptr) return set_error(ABT_ERR_ARG, "dz2 is null but requested");
    if (((reinterpret_cast<uintptr_t>(a->z1) | reinterpret_cast<uintptr_t>(a->z2)) & 15) != 0) return set_error(ABT_ERR_ARG, "z1 / z2 must be 16-byte aligned");
    const WsLayout L = ws_layout(a->n_rows, a->n_dims, a->n_dims, a->dtype, false);
    if (a->workspace_bytes < L.total) return set_error(ABT_ERR_ARG, "workspace too small: %zu < %zu", a->workspace_bytes, L.total);
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0) return set_error(ABT_ERR_ARG, "workspace must be 256-byte aligned");
    if (int rc = check_device_sm100()) return rc;
    LossCall c{};
    c.z1 = a->z1; c.z2 = a->z2; c.dtype = a->dtype; c.N = a->n_rows; c.D = a->n_dims;
    c.row_begin = 0; c.row_count = a->n_dims; c.rows_mode = false;
    c.alpha = a->alpha; c.lambda = a->lambda; c.hsic = a->hsic; c.eps = a->eps; c.momentum = a->momentum; c.grad_scale = a->grad_scale;
    c.need = a->need_grad_mask; c.loss_out = a->loss_out; c.loss_parts_out = nullptr;
    c.dz1 = a->dz1; c.dz2 = a->dz2; c.ld_dz = a->n_dims;
    c.running_mean = a->running_mean; c.running_var = a->running_var; c.workspace = a->workspace;
    return dispatch(c, L, reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// Projector tail: z = h W^T for both views on the tensor cores, with the column statistics the objective needs
// taken from the ROUNDED bf16 outputs in the epilogue.  The result is the hand-over abt_bt_dist_stats_local produces (5 sums + 2
// shifts per column), so the objective continues with abt_bt_dist_normalize / abt_bt_dist_rows_fwd_bwd (world = 1) and never
// reads z for its statistics.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bt_pack_fold_kernel(const float* __restrict__ partials, int n_splits, int D, float* __restrict__ pack) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= D) return;
    for (int k = 0; k < 5; ++k) {
        double t = 0.0;
        for (int r = 0; r < n_splits; ++r) t += (double)partials[((size_t)r * 5 + k) * D + c];      // fixed order: deterministic
        pack[(size_t)k * D + c] = (float)t;
    }
    pack[(size_t)5 * D + c] = 0.f;          // the sums are not shifted (a bias-free Linear after BatchNorm + ReLU has |mean| ~ std)
    pack[(size_t)6 * D + c] = 0.f;
}

static int proj_tail_splits(int n_rows) { return 2 * ((n_rows + 127) / 128); }

extern "C" int abt_proj_tail_workspace_bytes(int n_rows, int n_dims, size_t* bytes) {
    if (bytes == nullptr) return set_error(ABT_ERR_ARG, "bytes is null");
    if (n_rows < 1 || n_dims < 1) return set_error(ABT_ERR_ARG, "bad shape");
    *bytes = sizeof(float) * 5 * (size_t)proj_tail_splits(n_rows) * (size_t)n_dims;
    return 0;
}

extern "C" int abt_proj_tail_fwd(const void* h1, const void* h2, const void* w, int n_rows, int k_dims, int n_dims, void* z1, void* z2,
                                 float* pack, void* workspace, size_t workspace_bytes, abt_stream_t stream_) {
    if (h1 == nullptr || h2 == nullptr || w == nullptr || z1 == nullptr || z2 == nullptr || pack == nullptr || workspace == nullptr)
        return set_error(ABT_ERR_ARG, "null pointer argument");
    if (n_rows < 2 || k_dims < 64 || (k_dims % 8) != 0 || n_dims < 64 || (n_dims % 8) != 0)
        return set_error(ABT_ERR_ARG, "need n_rows >= 2, k_dims and n_dims >= 64 and multiples of 8 (got %d, %d, %d)", n_rows, k_dims, n_dims);
    if (((reinterpret_cast<uintptr_t>(h1) | reinterpret_cast<uintptr_t>(h2) | reinterpret_cast<uintptr_t>(w)) & 15) != 0)
        return set_error(ABT_ERR_ARG, "h1 / h2 / w must be 16-byte aligned");
    size_t need = 0;
    abt_proj_tail_workspace_bytes(n_rows, n_dims, &need);
    if (workspace_bytes < need) return set_error(ABT_ERR_ARG, "workspace too small: %zu < %zu", workspace_bytes, need);
    if (int rc = check_device_sm100()) return rc;
    if (int rc = ensure_umma_attr()) return rc;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CUtensorMap mW, mH1, mH2;
    if (int rc = make_map_16(&mW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w, n_dims, k_dims, 64, 128)) return rc;         // A: 128 output dimensions per CTA, K-major
    if (int rc = make_map_16(&mH1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, h1, n_rows, k_dims, 64, 128)) return rc;       // B: 128 samples of one view per CTA
    if (int rc = make_map_16(&mH2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, h2, n_rows, k_dims, 64, 128)) return rc;
    UmmaParams p{};
    p.dc = g_desc;
    p.mode = 2; p.D = n_dims; p.N = n_rows; p.bn = 256; p.ab_format = 1;
    p.tiles_m = (n_dims + 2 * BM - 1) / (2 * BM); p.tiles_n = (n_rows + 127) / 128;
    p.kblocks = (k_dims + BK - 1) / BK;
    p.pass_count = 1;
    p.split_from = p.tiles_m * p.tiles_n;
    p.lin_z1 = z1; p.lin_z2 = z2; p.lin_partials = static_cast<float*>(workspace);
    PassCfg pc{};
    pc.a_mn = 0; pc.row0 = 0; pc.row_end = n_dims;
    p.pass[0] = pc; p.pass[1] = pc;
    if (int rc = launch_umma(2, mW, mH1, mW, mH2, mW, mW, p, stream)) return rc;
    bt_pack_fold_kernel<<<(n_dims + 255) / 256, 256, 0, stream>>>(static_cast<const float*>(workspace), proj_tail_splits(n_rows), n_dims, pack);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "projector tail launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int abt_bt_rows_workspace_bytes(int n_rows, int n_dims, int row_count, int dtype, size_t* bytes) {
    if (bytes == nullptr) return set_error(ABT_ERR_ARG, "bytes is null");
    if (int rc = check_shape(n_rows, n_dims, dtype)) return rc;
    if (row_count < 8 || row_count > n_dims || (row_count % 8) != 0) return set_error(ABT_ERR_ARG, "row_count must be a multiple of 8 in [8, n_dims]");
    *bytes = ws_layout(n_rows, n_dims, row_count, dtype, true).total;
    return 0;
}

extern "C" int abt_bt_loss_rows_fwd_bwd(const abt_bt_rows_args* a, abt_stream_t stream) {
    if (a == nullptr) return set_error(ABT_ERR_ARG, "args is null");
    if (int rc = check_shape(a->n_rows, a->n_dims, a->dtype)) return rc;
    if (a->row_count < 8 || (a->row_count % 8) != 0 || a->row_begin < 0 || (a->row_begin % 8) != 0 || a->row_begin + a->row_count > a->n_dims)
        return set_error(ABT_ERR_ARG, "row block [%d, %d) must be 8-aligned and inside [0, %d)", a->row_begin, a->row_begin + a->row_count, a->n_dims);
    if (a->zg1 == nullptr || a->zg2 == nullptr || a->loss_parts == nullptr || a->workspace == nullptr) return set_error(ABT_ERR_ARG, "null pointer argument");
    if ((a->need_grad_mask & 1) && a->dzr1 == nullptr) return set_error(ABT_ERR_ARG, "dzr1 is null but requested");
    if ((a->need_grad_mask & 2) && a->dzr2 == nullptr) return set_error(ABT_ERR_ARG, "dzr2 is null but requested");
    const WsLayout L = ws_layout(a->n_rows, a->n_dims, a->row_count, a->dtype, true);
    if (a->workspace_bytes < L.total) return set_error(ABT_ERR_ARG, "workspace too small: %zu < %zu", a->workspace_bytes, L.total);
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0) return set_error(ABT_ERR_ARG, "workspace must be 256-byte aligned");
    if (int rc = check_device_sm100()) return rc;
    LossCall c{};
    c.z1 = a->zg1; c.z2 = a->zg2; c.dtype = a->dtype; c.N = a->n_rows; c.D = a->n_dims;
    c.row_begin = a->row_begin; c.row_count = a->row_count; c.rows_mode = true;
    c.alpha = a->alpha; c.lambda = a->lambda; c.hsic = a->hsic; c.eps = a->eps; c.momentum = a->momentum; c.grad_scale = a->grad_scale;
    c.need = a->need_grad_mask; c.loss_out = nullptr; c.loss_parts_out = a->loss_parts;
    c.dz1 = a->dzr1; c.dz2 = a->dzr2; c.ld_dz = a->row_count;
    c.running_mean = a->running_mean; c.running_var = a->running_var; c.workspace = a->workspace;
    return dispatch(c, L, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int abt_scale_inplace(void* a, void* b, size_t n_elems, int dtype, const float* scale_dev, abt_stream_t stream) {
    if (n_elems == 0 || (a == nullptr && b == nullptr)) return 0;
    if (scale_dev == nullptr) return set_error(ABT_ERR_ARG, "scale is null");
    if (dtype < 0 || dtype > 2) return set_error(ABT_ERR_ARG, "unknown dtype %d", dtype);
    const size_t esz = dtype == ABT_DTYPE_F32 ? 4 : 2;
    if ((n_elems * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(a) & 15) != 0 || (reinterpret_cast<uintptr_t>(b) & 15) != 0)
        return set_error(ABT_ERR_ARG, "buffers must be 16-byte aligned and a multiple of 16 bytes long");
    if (int rc = check_device_sm100()) return rc;
    const size_t n_vec = n_elems * esz / 16;
    const unsigned bx = (unsigned)((n_vec + 255) / 256 < 148u * 8u ? (n_vec + 255) / 256 : 148u * 8u);
    const dim3 grid(bx, 2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ABT_DTYPE_BF16) bt_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(b), n_vec, scale_dev);
    else if (dtype == ABT_DTYPE_F16) bt_scale_kernel<__half><<<grid, 256, 0, st>>>(static_cast<__half*>(a), static_cast<__half*>(b), n_vec, scale_dev);
    else bt_scale_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(a), static_cast<float*>(b), n_vec, scale_dev);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "bt_scale_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------
// multi-GPU objective (one process per GPU): statistics exchange + standardised-embedding gather + row block
// ------------------------------------------------------------------------------------------
static int check_dist(int n_local, int world, int n_dims, int row_count) {
    if (world < 1 || n_local < 1) return set_error(ABT_ERR_ARG, "bad world size / local batch");
    if (int rc = check_shape(n_local * world, n_dims, ABT_DTYPE_BF16)) return rc;
    if (row_count < 8 || row_count > n_dims || (row_count % 8) != 0) return set_error(ABT_ERR_ARG, "row_count must be a multiple of 8 in [8, n_dims]");
    return 0;
}

extern "C" int abt_bt_dist_layout_query(int n_local, int world, int n_dims, int row_count, abt_bt_dist_layout* out) {
    if (out == nullptr) return set_error(ABT_ERR_ARG, "layout is null");
    if (int rc = check_dist(n_local, world, n_dims, row_count)) return rc;
    const WsLayout L = ws_layout(n_local * world, n_dims, row_count, ABT_DTYPE_BF16, true, world);
    out->total_bytes = L.total; out->zh1 = L.zh1; out->zh2 = L.zh2; out->pack_local = L.pack_local; out->pack_all = L.pack_all;
    out->pack_floats = 7 * (size_t)n_dims;
    return 0;
}

template <typename T>
static int dist_stats_local(const void* z1, const void* z2, int N, int D, uint8_t* ws, const WsLayout& L, cudaStream_t stream) {
    float* partials = reinterpret_cast<float*>(ws + L.partials);
    int splits = stat_splits(N, D);
    const int chunk = (N + splits - 1) / splits;
    splits = (N + chunk - 1) / chunk;
    const dim3 sgrid((D + kStatCols - 1) / kStatCols, splits);
    bt_colstat_kernel<T><<<sgrid, kColThreads, 0, stream>>>(static_cast<const T*>(z1), static_cast<const T*>(z2), N, D, chunk, partials, nullptr, nullptr,
                                                            nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr);
    bt_stat_pack_kernel<T><<<(D + 255) / 256, 256, 0, stream>>>(static_cast<const T*>(z1), static_cast<const T*>(z2), D, splits, partials,
                                                                reinterpret_cast<float*>(ws + L.pack_local));
    count_launch(2);
    return 0;
}

extern "C" int abt_bt_dist_stats_local(const void* z1, const void* z2, int dtype, int n_local, int world, int n_dims, int row_count, void* workspace,
                                       abt_stream_t stream) {
    if (z1 == nullptr || z2 == nullptr || workspace == nullptr) return set_error(ABT_ERR_ARG, "null pointer argument");
    if (dtype < 0 || dtype > 2) return set_error(ABT_ERR_ARG, "unknown dtype %d", dtype);
    if (int rc = check_dist(n_local, world, n_dims, row_count)) return rc;
    if (int rc = check_device_sm100()) return rc;
    const WsLayout L = ws_layout(n_local * world, n_dims, row_count, ABT_DTYPE_BF16, true, world);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ABT_DTYPE_BF16) dist_stats_local<__nv_bfloat16>(z1, z2, n_local, n_dims, ws, L, st);
    else if (dtype == ABT_DTYPE_F16) dist_stats_local<__half>(z1, z2, n_local, n_dims, ws, L, st);
    else dist_stats_local<float>(z1, z2, n_local, n_dims, ws, L, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "dist statistics launch: %s", cudaGetErrorString(e));
    return 0;
}

template <typename T>
static void dist_normalize(const void* z1, const void* z2, int N, int world, int rank, int D, int row_count, float eps, float momentum, float* rm,
                           float* rv, uint8_t* ws, const WsLayout& L, cudaStream_t stream, __half* zh1_base = nullptr, __half* zh2_base = nullptr) {
    int splits = (N + 127) / 128;
    if (splits > 8) splits = 8;
    const int chunk = (N + splits - 1) / splits;
    splits = (N + chunk - 1) / chunk;
    const dim3 sgrid((D + kColsPerBlock - 1) / kColsPerBlock, splits);
    __half* zh1 = (zh1_base ? zh1_base : reinterpret_cast<__half*>(ws + L.zh1)) + (size_t)rank * N * D;     // this rank's slot of the gather buffers
    __half* zh2 = (zh2_base ? zh2_base : reinterpret_cast<__half*>(ws + L.zh2)) + (size_t)rank * N * D;
    cudaMemsetAsync(ws + L.keep, 0, 256, stream);      // the on-diagonal loss sum this kernel accumulates (cleared here, so that the statistics
                                                       // hand-over may come from anywhere: abt_bt_dist_stats_local or the fused projector tail)
    bt_normalize_global_kernel<T><<<sgrid, kColThreads, 0, stream>>>(static_cast<const T*>(z1), static_cast<const T*>(z2), N, world, D, chunk, eps, momentum,
                                                                     reinterpret_cast<const float*>(ws + L.pack_all), reinterpret_cast<float*>(ws + L.stats),
                                                                     zh1, zh2, reinterpret_cast<__half*>(ws + L.zh1_blk), row_count, rm, rv,
                                                                     reinterpret_cast<double*>(ws + L.keep));
    count_launch();
}

extern "C" int abt_bt_dist_normalize(const void* z1, const void* z2, int dtype, int n_local, int world, int rank, int n_dims, int row_count, float eps,
                                     float momentum, float* running_mean, float* running_var, void* workspace, abt_stream_t stream) {
    if (z1 == nullptr || z2 == nullptr || workspace == nullptr) return set_error(ABT_ERR_ARG, "null pointer argument");
    if (dtype < 0 || dtype > 2) return set_error(ABT_ERR_ARG, "unknown dtype %d", dtype);
    if (rank < 0 || rank >= world) return set_error(ABT_ERR_ARG, "rank %d outside [0, %d)", rank, world);
    if (int rc = check_dist(n_local, world, n_dims, row_count)) return rc;
    if (int rc = check_device_sm100()) return rc;
    const WsLayout L = ws_layout(n_local * world, n_dims, row_count, ABT_DTYPE_BF16, true, world);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ABT_DTYPE_BF16) dist_normalize<__nv_bfloat16>(z1, z2, n_local, world, rank, n_dims, row_count, eps, momentum, running_mean, running_var, ws, L, st);
    else if (dtype == ABT_DTYPE_F16) dist_normalize<__half>(z1, z2, n_local, world, rank, n_dims, row_count, eps, momentum, running_mean, running_var, ws, L, st);
    else dist_normalize<float>(z1, z2, n_local, world, rank, n_dims, row_count, eps, momentum, running_mean, running_var, ws, L, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "dist normalize launch: %s", cudaGetErrorString(e));
    return 0;
}

static int dist_rows_call(int dtype, int n_local, int world, int n_dims, int row_begin, int row_count, float alpha, float lambda, int hsic,
                          float grad_scale, int need, int phase, double* loss_parts, void* dzr1, void* dzr2, void* workspace, cudaStream_t stream,
                          bool xchg = false, const void* zh1_base = nullptr, const void* zh2_base = nullptr) {
    const int ng = n_local * world;
    const WsLayout L = ws_layout(ng, n_dims, row_count, ABT_DTYPE_BF16, true, world);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    LossCall c{};
    c.z1 = zh1_base ? zh1_base : ws + L.zh1; c.z2 = zh2_base ? zh2_base : ws + L.zh2; c.dtype = dtype; c.N = ng; c.D = n_dims;
    // one rank owning every dimension (the fused projector tail continues here): C^T is the same matrix read transposed, no second CORR pass
    const bool whole = world == 1 && row_begin == 0 && row_count == n_dims && !xchg;
    c.row_begin = row_begin; c.row_count = row_count; c.rows_mode = !whole; c.zh_mode = true; c.phase = phase; c.xchg = xchg;
    c.alpha = alpha; c.lambda = lambda; c.hsic = hsic; c.eps = 0.f; c.momentum = 0.f; c.grad_scale = grad_scale;
    c.need = need; c.loss_out = nullptr; c.loss_parts_out = loss_parts;
    c.dz1 = dzr1; c.dz2 = dzr2; c.ld_dz = row_count;
    c.running_mean = nullptr; c.running_var = nullptr; c.workspace = workspace;
    return run_loss<__nv_bfloat16>(c, L, stream);
}

extern "C" int abt_bt_dist_rows_fwd_bwd(const abt_bt_dist_args* a, abt_stream_t stream) {
    if (a == nullptr) return set_error(ABT_ERR_ARG, "args is null");
    if (int rc = check_dist(a->n_local, a->world, a->n_dims, a->row_count)) return rc;
    if (a->dtype < 0 || a->dtype > 2) return set_error(ABT_ERR_ARG, "unknown dtype %d", a->dtype);
    if (a->row_begin < 0 || (a->row_begin % 8) != 0 || a->row_begin + a->row_count > a->n_dims)
        return set_error(ABT_ERR_ARG, "row block [%d, %d) must be 8-aligned and inside [0, %d)", a->row_begin, a->row_begin + a->row_count, a->n_dims);
    if (a->loss_parts == nullptr || a->workspace == nullptr) return set_error(ABT_ERR_ARG, "null pointer argument");
    if (a->phase < 0 || a->phase > 2) return set_error(ABT_ERR_ARG, "phase must be 0, 1 or 2");
    if ((a->need_grad_mask & 1) && a->phase != 2 && a->dzr1 == nullptr) return set_error(ABT_ERR_ARG, "dzr1 is null but requested");
    if ((a->need_grad_mask & 2) && a->phase != 1 && a->dzr2 == nullptr) return set_error(ABT_ERR_ARG, "dzr2 is null but requested");
    const WsLayout L = ws_layout(a->n_local * a->world, a->n_dims, a->row_count, ABT_DTYPE_BF16, true, a->world);
    if (a->workspace_bytes < L.total) return set_error(ABT_ERR_ARG, "workspace too small: %zu < %zu", a->workspace_bytes, L.total);
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0) return set_error(ABT_ERR_ARG, "workspace must be 256-byte aligned");
    if (int rc = check_device_sm100()) return rc;
    return dist_rows_call(a->dtype, a->n_local, a->world, a->n_dims, a->row_begin, a->row_count, a->alpha, a->lambda, a->hsic, a->grad_scale,
                          a->need_grad_mask, a->phase, a->loss_parts, a->dzr1, a->dzr2, a->workspace, reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// native multi-GPU step: the whole choreography (statistics exchange, embedding all-gather, row block, gradient
// all-to-all, loss all-reduce) issued from ONE call, with the collectives on NCCL (the library PyTorch already loaded,
// resolved at run time) and a private communication stream, so that a training step costs one host call instead of
// a dozen framework-level collectives
// ------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <nccl.h>

namespace abt {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*);     // optional
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*AlltoAll)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);   // NCCL >= 2.28, optional
    bool ok = false;
};

static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.ok) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);     // the copy PyTorch loaded, if any: same soname
    if (h == nullptr) return set_error(ABT_ERR_STATE, "libnccl.so.2 not found: %s", dlerror());
#define ABT_NCCL_SYM(field, name)                                                        \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                            \
    if (g_nccl.field == nullptr) return set_error(ABT_ERR_STATE, "libnccl.so.2 lacks %s", name)
    ABT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    ABT_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    ABT_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    ABT_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    ABT_NCCL_SYM(AllGather, "ncclAllGather");
    ABT_NCCL_SYM(AllReduce, "ncclAllReduce");
    ABT_NCCL_SYM(Send, "ncclSend");
    ABT_NCCL_SYM(Recv, "ncclRecv");
    ABT_NCCL_SYM(GroupStart, "ncclGroupStart");
    ABT_NCCL_SYM(GroupEnd, "ncclGroupEnd");
#undef ABT_NCCL_SYM
    *reinterpret_cast<void**>(&g_nccl.AlltoAll) = dlsym(h, "ncclAlltoAll");
    *reinterpret_cast<void**>(&g_nccl.CommInitRankConfig) = dlsym(h, "ncclCommInitRankConfig");
    g_nccl.ok = true;
    return 0;
}

#define ABT_NCCL_OK(expr)                                                                                              \
    do {                                                                                                               \
        ncclResult_t r__ = (expr);                                                                                     \
        if (r__ != ncclSuccess) return set_error(ABT_ERR_CUDA, "%s: %s", #expr, g_nccl.GetErrorString(r__));           \
    } while (0)

// recv [src rank][n][count] -> dz [n][src * count + j]: the slices a rank received from every dimension owner, in sample-major order
template <int VEC_BYTES>
__global__ void __launch_bounds__(256) bt_unpermute_kernel(const uint4* __restrict__ recv, uint4* __restrict__ dz, int world, int n, int vec_per_slice) {
    const size_t total = (size_t)world * n * vec_per_slice;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec_per_slice);
        const size_t t = i / vec_per_slice;
        const int row = (int)(t % n), src = (int)(t / n);
        dz[((size_t)row * world + src) * vec_per_slice + v] = recv[i];
    }
}

__global__ void bt_dist_loss_kernel(const double* __restrict__ parts, float alpha, float lambda, int hsic, int D, int world, float* __restrict__ loss_out) {
    if (threadIdx.x == 0) {
        double off = parts[0];
        if (hsic) off = parts[0] + 2.0 * parts[1] + (double)D * (double)(D - 1);
        *loss_out = (float)((double)alpha * parts[2] / (double)world + (double)lambda * off);   // the on-diagonal sum is global on every rank
    }
}

// Copy-engine exchange buffer (one per rank, peer-mapped): [parity 0: zh1 (N_g x D fp16) | zh2] [parity 1: ...] [flags]
constexpr int kMaxPeers = 16;
struct ExchLayout { size_t view_bytes, flags, total; };
static ExchLayout exch_layout(int n_local, int world, int D) {
    ExchLayout e{};
    e.view_bytes = align_up((size_t)n_local * world * D * 2, 256);
    e.flags = 4 * e.view_bytes;
    e.total = e.flags + 4096;
    return e;
}
struct PeerFlags { uint32_t* f[kMaxPeers]; };
// One flag barrier per step: thread q tells rank q "rank `rank` has finished standardising its rows for step `epoch`" (a release store
// into q's flag array over NVLink) and waits for the same word from q in this rank's own array.  Every rank runs one such CTA on its
// own GPU; the spin is bounded (a lost peer traps instead of hanging the device).
__global__ void ce_barrier_kernel(PeerFlags pf, int world, int rank, uint32_t epoch) {
    const int q = threadIdx.x;
    if (q >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.f[q] + rank * 16), "r"(epoch) : "memory");
    const uint32_t* mine = pf.f[rank] + q * 16;
    uint32_t v = 0;
    for (unsigned long long spin = 0;; ++spin) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (spin > (1ull << 31)) __trap();
    }
}

struct DistLayout {
    WsLayout L;
    size_t dzr1, dzr2, recv1, recv2, parts, total;
};

static DistLayout dist_layout(int n_local, int world, int D, int row_count) {
    DistLayout d{};
    d.L = ws_layout(n_local * world, D, row_count, ABT_DTYPE_BF16, true, world);
    size_t off = d.L.total;
    const size_t slab = align_up((size_t)n_local * world * row_count * 4, 256);     // sized for fp32 gradients
    d.dzr1 = off; off += slab;
    d.dzr2 = off; off += slab;
    d.recv1 = off; off += slab;
    d.recv2 = off; off += slab;
    d.parts = off; off += 256;
    d.total = off;
    return d;
}

}  // namespace abt

struct abt_comm {
    ncclComm_t comm;          // embedding all-gathers (large)
    ncclComm_t comm2;         // statistics packs, slice / block / gradient all-to-alls, loss all-reduce (small, latency-bound)
    int world, rank;
    cudaStream_t cs, cs2;     // one communication stream per communicator
    cudaEvent_t ev[12];
    cudaStream_t ce[4];       // copy-engine exchange: peer copies are spread over four streams
    cudaEvent_t ce_ev[2][4];  // [view][stream]
    cudaEvent_t ce_ready;
};

extern "C" int abt_comm_unique_id(void* id256) {
    if (id256 == nullptr) return set_error(ABT_ERR_ARG, "id buffer is null");
    if (int rc = load_nccl()) return rc;
    ncclUniqueId id[2];
    ABT_NCCL_OK(g_nccl.GetUniqueId(&id[0]));
    ABT_NCCL_OK(g_nccl.GetUniqueId(&id[1]));
    std::memcpy(id256, id, sizeof(id));
    return 0;
}

extern "C" int abt_comm_create(int world, int rank, const void* id256, abt_comm** out) {
    if (id256 == nullptr || out == nullptr || world < 1 || rank < 0 || rank >= world) return set_error(ABT_ERR_ARG, "bad communicator arguments");
    if (int rc = load_nccl()) return rc;
    if (int rc = check_device_sm100()) return rc;
    abt_comm* c = new abt_comm();
    c->world = world; c->rank = rank;
    ncclUniqueId id[2];
    std::memcpy(id, id256, sizeof(id));
    // The collectives run beside the persistent tensor-core kernels, which leave g_dist_reserve_sms SMs free for them (a CTA of
    // bt_umma_kernel owns a whole SM's registers, so a collective launched behind a GEMM otherwise waits for it to end).  An optional
    // CTA cap for NCCL exists (key 13) but is off: it slows the all-to-alls far more than it helps.
    ncclResult_t r;
    if (g_nccl.CommInitRankConfig != nullptr && g_comm_max_ctas > 0) {
        ncclConfig_t cfg1 = NCCL_CONFIG_INITIALIZER, cfg2 = NCCL_CONFIG_INITIALIZER;
        cfg1.maxCTAs = g_comm_max_ctas;
        cfg2.maxCTAs = g_comm_max_ctas > 4 ? 4 : g_comm_max_ctas;
        r = g_nccl.CommInitRankConfig(&c->comm, world, id[0], rank, &cfg1);
        if (r == ncclSuccess) r = g_nccl.CommInitRankConfig(&c->comm2, world, id[1], rank, &cfg2);
    } else {
        r = g_nccl.CommInitRank(&c->comm, world, id[0], rank);
        if (r == ncclSuccess) r = g_nccl.CommInitRank(&c->comm2, world, id[1], rank);
    }
    if (r != ncclSuccess) { delete c; return set_error(ABT_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&c->cs, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->cs2, cudaStreamNonBlocking, hi) != cudaSuccess) {
        g_nccl.CommDestroy(c->comm); g_nccl.CommDestroy(c->comm2); delete c;
        return set_error(ABT_ERR_CUDA, "cudaStreamCreate failed");
    }
    for (auto& e : c->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& s : c->ce) cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi);
    for (auto& v : c->ce_ev) for (auto& e : v) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ce_ready, cudaEventDisableTiming);
    *out = c;
    return 0;
}

extern "C" int abt_comm_destroy(abt_comm* c) {
    if (c == nullptr) return 0;
    cudaStreamSynchronize(c->cs);
    cudaStreamSynchronize(c->cs2);
    for (auto& e : c->ev) cudaEventDestroy(e);
    for (auto& s : c->ce) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (auto& v : c->ce_ev) for (auto& e : v) cudaEventDestroy(e);
    cudaEventDestroy(c->ce_ready);
    cudaStreamDestroy(c->cs);
    cudaStreamDestroy(c->cs2);
    if (g_nccl.ok) { g_nccl.CommDestroy(c->comm); g_nccl.CommDestroy(c->comm2); }
    delete c;
    return 0;
}

extern "C" int abt_bt_dist_exchange_bytes(int n_local, int world, int n_dims, size_t* bytes) {
    if (bytes == nullptr) return set_error(ABT_ERR_ARG, "bytes is null");
    if (world < 1 || world > kMaxPeers || n_dims % world != 0) return set_error(ABT_ERR_ARG, "world must be in [1, 16] and divide n_dims");
    if (int rc = check_dist(n_local, world, n_dims, n_dims / world)) return rc;
    *bytes = exch_layout(n_local, world, n_dims).total;
    return 0;
}

extern "C" int abt_bt_dist_step_workspace_bytes(int n_local, int world, int n_dims, size_t* bytes) {
    if (bytes == nullptr) return set_error(ABT_ERR_ARG, "bytes is null");
    if (world < 1 || n_dims % world != 0) return set_error(ABT_ERR_ARG, "n_dims must be divisible by the world size");
    if (int rc = check_dist(n_local, world, n_dims, n_dims / world)) return rc;
    *bytes = dist_layout(n_local, world, n_dims, n_dims / world).total;
    return 0;
}

static ncclDataType_t nccl_dtype(int dtype) { return dtype == ABT_DTYPE_BF16 ? ncclBfloat16 : (dtype == ABT_DTYPE_F16 ? ncclFloat16 : ncclFloat32); }

// all-to-all of (world) equal slices: slice q of `send` goes to rank q, slice q of `recv` comes from rank q
static int all_to_all(abt_comm* c, ncclComm_t comm, const void* send, void* recv, size_t slice_elems, int dtype, cudaStream_t st) {
    const size_t esz = dtype == ABT_DTYPE_F32 ? 4 : 2;
    if (g_nccl.AlltoAll != nullptr) {          // one call instead of 2 R point-to-point operations
        ABT_NCCL_OK(g_nccl.AlltoAll(send, recv, slice_elems, nccl_dtype(dtype), comm, st));
        return 0;
    }
    ABT_NCCL_OK(g_nccl.GroupStart());
    ncclResult_t first = ncclSuccess;          // an error inside the group must not leave it open: remember it, close the group, then report
    for (int q = 0; q < c->world && first == ncclSuccess; ++q) {
        first = g_nccl.Send(static_cast<const uint8_t*>(send) + (size_t)q * slice_elems * esz, slice_elems, nccl_dtype(dtype), q, comm, st);
        if (first == ncclSuccess) first = g_nccl.Recv(static_cast<uint8_t*>(recv) + (size_t)q * slice_elems * esz, slice_elems, nccl_dtype(dtype), q, comm, st);
    }
    const ncclResult_t end = g_nccl.GroupEnd();
    ABT_NCCL_OK(first);
    ABT_NCCL_OK(end);
    return 0;
}

extern "C" int abt_bt_dist_step(const abt_bt_dist_step_args* a, abt_comm* c, abt_stream_t stream_) {
    if (a == nullptr || c == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    const int world = c->world, rank = c->rank, N = a->n_local, D = a->n_dims;
    if (world < 1 || D % world != 0) return set_error(ABT_ERR_ARG, "n_dims must be divisible by the world size");
    const int Dr = D / world, R0 = rank * Dr;
    if (int rc = check_dist(N, world, D, Dr)) return rc;
    if (a->dtype < 0 || a->dtype > 2) return set_error(ABT_ERR_ARG, "unknown dtype %d", a->dtype);
    if (a->z1 == nullptr || a->z2 == nullptr || a->loss_out == nullptr || a->workspace == nullptr) return set_error(ABT_ERR_ARG, "null pointer argument");
    const int need = a->need_grad_mask & 3;
    if ((need & 1) && a->dz1 == nullptr) return set_error(ABT_ERR_ARG, "dz1 is null but requested");
    if ((need & 2) && a->dz2 == nullptr) return set_error(ABT_ERR_ARG, "dz2 is null but requested");
    const DistLayout dl = dist_layout(N, world, D, Dr);
    if (a->workspace_bytes < dl.total) return set_error(ABT_ERR_ARG, "workspace too small: %zu < %zu", a->workspace_bytes, dl.total);
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0) return set_error(ABT_ERR_ARG, "workspace must be 256-byte aligned");
    if (int rc = check_device_sm100()) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    const WsLayout& L = dl.L;
    const size_t esz = a->dtype == ABT_DTYPE_F32 ? 4 : 2;
    const size_t slice = (size_t)N * Dr;                       // elements of one (source rank) slice of a gradient slab

    enum { E_P0, E_P1, E_N, E_A, E_B, E_S, E_C3, E_C, E_G1, E_D, E_G2, E_E };
    auto chain = [](cudaStream_t from, cudaEvent_t ev, cudaStream_t to) { cudaEventRecord(ev, from); cudaStreamWaitEvent(to, ev, 0); };
    // copy-engine exchange: the gather buffers live in this rank's peer-mapped exchange buffer (double-buffered by the epoch's parity)
    const bool ce = a->exchange_peers != nullptr && world > 1;
    const ExchLayout xl = exch_layout(N, world, D);
    if (ce) {
        if (world > kMaxPeers) return set_error(ABT_ERR_ARG, "copy-engine exchange supports at most %d ranks", kMaxPeers);
        if (a->exchange_bytes < xl.total) return set_error(ABT_ERR_ARG, "exchange buffers too small: %zu < %zu", a->exchange_bytes, xl.total);
        if (a->exchange_epoch == 0) return set_error(ABT_ERR_ARG, "exchange_epoch starts at 1");
        for (int q = 0; q < world; ++q)
            if (a->exchange_peers[q] == nullptr || (reinterpret_cast<uintptr_t>(a->exchange_peers[q]) & 255) != 0)
                return set_error(ABT_ERR_ARG, "exchange buffer of rank %d is null or not 256-byte aligned", q);
    }
    const size_t par_off = ce ? (size_t)(a->exchange_epoch & 1u) * 2 * xl.view_bytes : 0;
    uint8_t* xbase = ce ? static_cast<uint8_t*>(a->exchange_peers[rank]) : nullptr;
    __half* zh1 = ce ? reinterpret_cast<__half*>(xbase + par_off) : reinterpret_cast<__half*>(ws + L.zh1);
    __half* zh2 = ce ? reinterpret_cast<__half*>(xbase + par_off + xl.view_bytes) : reinterpret_cast<__half*>(ws + L.zh2);
    const void* zo1 = ce ? zh1 : nullptr;      // gather-buffer overrides of the row-block calls
    const void* zo2 = ce ? zh2 : nullptr;
    // all-gather of one view: NCCL, or R - 1 peer copies on the copy engines (view: 0 = zh1, 1 = zh2); completion -> `done` events
    auto gather_view = [&](int view) -> int {
        __half* mine = view == 0 ? zh1 : zh2;
        if (!ce) {
            ABT_NCCL_OK(g_nccl.AllGather(mine + (size_t)rank * N * D, mine, (size_t)N * D, ncclFloat16, c->comm, c->cs));
            return 0;
        }
        const size_t slot_bytes = (size_t)N * D * 2, voff = par_off + (view == 0 ? 0 : xl.view_bytes);
        for (int i = 1; i < world; ++i) {
            const int q = (rank + i) % world, k = (i - 1) & 3;
            cudaMemcpyAsync(reinterpret_cast<uint8_t*>(mine) + (size_t)q * slot_bytes,
                            static_cast<const uint8_t*>(a->exchange_peers[q]) + voff + (size_t)q * slot_bytes, slot_bytes, cudaMemcpyDeviceToDevice, c->ce[k]);
        }
        for (int k = 0; k < 4; ++k) cudaEventRecord(c->ce_ev[view][k], c->ce[k]);
        return 0;
    };
    auto wait_view = [&](int view, cudaEvent_t nccl_ev) {
        if (!ce) { cudaStreamWaitEvent(st, nccl_ev, 0); return; }
        for (int k = 0; k < 4; ++k) cudaStreamWaitEvent(st, c->ce_ev[view][k], 0);
    };
    double* parts = reinterpret_cast<double*>(ws + dl.parts);
    const int vec_per_slice = (int)((size_t)Dr * esz / 16);
    const size_t nvec = (size_t)world * N * vec_per_slice;
    const unsigned ublocks = (unsigned)((nvec + 255) / 256 < 148u * 8u ? (nvec + 255) / 256 : 148u * 8u);

    // while this step runs, the tensor-core kernels leave room for NCCL's CTAs (restored before returning on every path below via the guard)
    struct ReserveGuard { int prev; ReserveGuard(int v) : prev(g_reserve_sms) { g_reserve_sms = v; } ~ReserveGuard() { g_reserve_sms = prev; } };
    ReserveGuard reserve_guard(world > 1 ? g_dist_reserve_sms : g_reserve_sms);
    // 1. local statistics -> all-gather of the 7 D-float packs
    if (int rc = abt_bt_dist_stats_local(a->z1, a->z2, a->dtype, N, world, D, Dr, a->workspace, stream_)) return rc;
    chain(st, c->ev[E_P0], c->cs2);
    ABT_NCCL_OK(g_nccl.AllGather(ws + L.pack_local, ws + L.pack_all, 7 * (size_t)D, ncclFloat32, c->comm2, c->cs2));
    chain(c->cs2, c->ev[E_P1], st);
    // 2. global statistics + standardised local rows (+ the column-blocked copy of view 1)
    if (a->dtype == ABT_DTYPE_BF16) dist_normalize<__nv_bfloat16>(a->z1, a->z2, N, world, rank, D, Dr, a->eps, a->momentum, a->running_mean, a->running_var, ws, L, st, ce ? zh1 : nullptr, ce ? zh2 : nullptr);
    else if (a->dtype == ABT_DTYPE_F16) dist_normalize<__half>(a->z1, a->z2, N, world, rank, D, Dr, a->eps, a->momentum, a->running_mean, a->running_var, ws, L, st, ce ? zh1 : nullptr, ce ? zh2 : nullptr);
    else dist_normalize<float>(a->z1, a->z2, N, world, rank, D, Dr, a->eps, a->momentum, a->running_mean, a->running_var, ws, L, st, ce ? zh1 : nullptr, ce ? zh2 : nullptr);
    if (ce) {
        // every rank's rows are in its own exchange buffer once all ranks have passed this barrier; by the parity argument (see the
        // header) it also guarantees that nobody still reads the buffers this step is about to overwrite two steps from now
        PeerFlags pf{};
        for (int q = 0; q < world; ++q) pf.f[q] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(a->exchange_peers[q]) + xl.flags);
        ce_barrier_kernel<<<1, 32, 0, st>>>(pf, world, rank, a->exchange_epoch);
        count_launch();
        cudaEventRecord(c->ce_ready, st);
        for (int k = 0; k < 4; ++k) cudaStreamWaitEvent(c->ce[k], c->ce_ready, 0);
    }
    cudaEventRecord(c->ev[E_N], st);
    cudaStreamWaitEvent(c->cs, c->ev[E_N], 0);
    cudaStreamWaitEvent(c->cs2, c->ev[E_N], 0);
    // measured on B200 / NVSwitch at N = 1024 per rank, D = 8192: at 2 ranks the C-block exchange (D^2 / 2 bytes) costs what the second
    // CORR pass costs (0.96 vs 0.93 ms per step), at 8 ranks the exchange schedule wins (1.06 vs 1.14 ms)
    const bool xchg = need == 3 && (Dr % 64) == 0 && (g_dist_xchg == 1 || (g_dist_xchg < 0 && world >= 4));
    if (xchg) {
        // Exchange schedule.  View 2 is gathered first; CORR (A = the view-1 columns of this rank's dimensions, which arrive by a small
        // all-to-all) and the dz1 GEMM need nothing else and run while view 1 is still crossing NVLink.  The transposed block of C comes
        // from an all-to-all of C blocks instead of a second CORR pass (6 N D^2 executed FLOP per rank, as on one GPU).
        if (int rc = gather_view(1)) return rc;
        cudaEventRecord(c->ev[E_A], c->cs);
        if (int rc = gather_view(0)) return rc;
        cudaEventRecord(c->ev[E_B], c->cs);
        if (int rc = all_to_all(c, c->comm2, ws + L.zh1_blk, ws + L.zs1, (size_t)N * Dr, ABT_DTYPE_F16, c->cs2)) return rc;
        cudaEventRecord(c->ev[E_S], c->cs2);
        if (a->overlap_cb != nullptr) a->overlap_cb(a->overlap_user);
        wait_view(1, c->ev[E_A]);
        cudaStreamWaitEvent(st, c->ev[E_S], 0);
        if (int rc = dist_rows_call(a->dtype, N, world, D, R0, Dr, a->alpha, a->lambda, a->hsic, a->grad_scale, need, 8 | 1, parts, ws + dl.dzr1,
                                    ws + dl.dzr2, a->workspace, st, true, zo1, zo2)) return rc;
        chain(st, c->ev[E_C3], c->cs2);
        if (int rc = all_to_all(c, c->comm2, ws + L.c1, ws + L.c2, (size_t)Dr * Dr, ABT_DTYPE_F16, c->cs2)) return rc;      // C[rows_r, cols_q] -> rank q
        cudaEventRecord(c->ev[E_C], c->cs2);
        if (int rc = dist_rows_call(a->dtype, N, world, D, R0, Dr, a->alpha, a->lambda, a->hsic, a->grad_scale, need, 8 | 2, parts, ws + dl.dzr1,
                                    ws + dl.dzr2, a->workspace, st, true, zo1, zo2)) return rc;
        chain(st, c->ev[E_G1], c->cs2);
        if (int rc = all_to_all(c, c->comm2, ws + dl.dzr1, ws + dl.recv1, slice, a->dtype, c->cs2)) return rc;
        ABT_NCCL_OK(g_nccl.AllReduce(parts, parts, 3, ncclFloat64, ncclSum, c->comm2, c->cs2));
        cudaEventRecord(c->ev[E_D], c->cs2);
        wait_view(0, c->ev[E_B]);
        cudaStreamWaitEvent(st, c->ev[E_C], 0);
        if (int rc = dist_rows_call(a->dtype, N, world, D, R0, Dr, a->alpha, a->lambda, a->hsic, a->grad_scale, need, 8 | 4, parts, ws + dl.dzr1,
                                    ws + dl.dzr2, a->workspace, st, true, zo1, zo2)) return rc;
        chain(st, c->ev[E_G2], c->cs2);
        if (int rc = all_to_all(c, c->comm2, ws + dl.dzr2, ws + dl.recv2, slice, a->dtype, c->cs2)) return rc;
        cudaEventRecord(c->ev[E_E], c->cs2);
        cudaStreamWaitEvent(st, c->ev[E_D], 0);
        bt_unpermute_kernel<16><<<ublocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws + dl.recv1), static_cast<uint4*>(a->dz1), world, N, vec_per_slice);
        bt_dist_loss_kernel<<<1, 32, 0, st>>>(parts, a->alpha, a->lambda, a->hsic, D, world, a->loss_out);
        cudaStreamWaitEvent(st, c->ev[E_E], 0);
        bt_unpermute_kernel<16><<<ublocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws + dl.recv2), static_cast<uint4*>(a->dz2), world, N, vec_per_slice);
    } else {
        // Plain schedule: both views gathered, then the row block of C and of C^T (second CORR pass); the all-to-all of dz1 overlaps
        // the dz2 GEMM
        if (ce) {
            if (int rc = gather_view(0)) return rc;
            if (int rc = gather_view(1)) return rc;
        } else {
            ABT_NCCL_OK(g_nccl.GroupStart());
            ncclResult_t first = g_nccl.AllGather(zh1 + (size_t)rank * N * D, zh1, (size_t)N * D, ncclFloat16, c->comm, c->cs);
            if (first == ncclSuccess) first = g_nccl.AllGather(zh2 + (size_t)rank * N * D, zh2, (size_t)N * D, ncclFloat16, c->comm, c->cs);
            const ncclResult_t end = g_nccl.GroupEnd();      // never leave the group open on an error
            ABT_NCCL_OK(first);
            ABT_NCCL_OK(end);
        }
        cudaEventRecord(c->ev[E_A], c->cs);
        if (a->overlap_cb != nullptr) a->overlap_cb(a->overlap_user);
        wait_view(0, c->ev[E_A]);
        wait_view(1, c->ev[E_A]);
        const int phase_a = need == 3 ? 1 : 0;
        if (int rc = dist_rows_call(a->dtype, N, world, D, R0, Dr, a->alpha, a->lambda, a->hsic, a->grad_scale, need, phase_a, parts, ws + dl.dzr1,
                                    ws + dl.dzr2, a->workspace, st, false, zo1, zo2)) return rc;
        chain(st, c->ev[E_G1], c->cs2);
        if (need & 1) { if (int rc = all_to_all(c, c->comm2, ws + dl.dzr1, ws + dl.recv1, slice, a->dtype, c->cs2)) return rc; }
        if (need == 2) { if (int rc = all_to_all(c, c->comm2, ws + dl.dzr2, ws + dl.recv2, slice, a->dtype, c->cs2)) return rc; }
        ABT_NCCL_OK(g_nccl.AllReduce(parts, parts, 3, ncclFloat64, ncclSum, c->comm2, c->cs2));
        cudaEventRecord(c->ev[E_D], c->cs2);
        if (need == 3) {
            if (int rc = dist_rows_call(a->dtype, N, world, D, R0, Dr, a->alpha, a->lambda, a->hsic, a->grad_scale, need, 2, parts, ws + dl.dzr1,
                                        ws + dl.dzr2, a->workspace, st, false, zo1, zo2)) return rc;
            chain(st, c->ev[E_G2], c->cs2);
            if (int rc = all_to_all(c, c->comm2, ws + dl.dzr2, ws + dl.recv2, slice, a->dtype, c->cs2)) return rc;
            cudaEventRecord(c->ev[E_E], c->cs2);
        }
        cudaStreamWaitEvent(st, c->ev[E_D], 0);
        if (need & 1) bt_unpermute_kernel<16><<<ublocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws + dl.recv1), static_cast<uint4*>(a->dz1), world, N, vec_per_slice);
        if (need == 2) bt_unpermute_kernel<16><<<ublocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws + dl.recv2), static_cast<uint4*>(a->dz2), world, N, vec_per_slice);
        bt_dist_loss_kernel<<<1, 32, 0, st>>>(parts, a->alpha, a->lambda, a->hsic, D, world, a->loss_out);
        if (need == 3) {
            cudaStreamWaitEvent(st, c->ev[E_E], 0);
            bt_unpermute_kernel<16><<<ublocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws + dl.recv2), static_cast<uint4*>(a->dz2), world, N, vec_per_slice);
        }
    }
    count_launch(need == 3 ? 3 : (need ? 2 : 1));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "dist step launch: %s", cudaGetErrorString(e));
    return 0;
}
