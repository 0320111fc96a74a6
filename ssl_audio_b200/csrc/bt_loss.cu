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
//   1. bt_stats_kernel     column statistics of both views, fp32 diagonal C_ii and on-diagonal loss,
//                          BatchNorm running-stat update, bf16 copies of non-bf16 inputs and the
//                          fp16 standardised embeddings zh used by the gradient GEMMs
//   2. bt_umma_kernel CORR S = z1^T z2 on the tensor cores (tcgen05, RAW bf16 operands straight
//                          from the row-major embeddings as MN-major TMA tiles, fp32 TMEM
//                          accumulator); epilogue applies batch-norm as a rank-1 correction,
//                          reduces the off-diagonal loss in fp32 and emits C (|C_ij| <= 1) in fp16
//                          with the diagonal zeroed
//   3. bt_umma_kernel GRAD g1^T = C zh2^T (K-major A) and g2^T = C^T zh1^T (MN-major A over the
//                          same C), fp32 out: 6 N D^2 executed FLOP = the algorithmic count
//   4. bt_finalize_kernel  adds the fp32 diagonal term, batch-norm backward, output cast, loss.
//
// Row-block mode (multi-GPU, abt_bt_loss_rows_fwd_bwd): the inputs are the rank-ordered gathered
// embeddings (N_g x D); this rank owns dimensions [row_begin, row_begin + row_count) and computes
// that row block of C AND of C^T (second CORR pass with the views swapped), so that both gradient
// GEMMs are complete for its dimensions and batch-norm backward (column-local) needs no reduction.
#include "abt_internal.h"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace abt {

// ------------------------------------------------------------------------------------------
// column-statistics layout inside the workspace (float arrays of length D each)
// ------------------------------------------------------------------------------------------
enum StatSlot {
    S_MU1 = 0, S_R1, S_MU2, S_R2, S_CDIAG,
    S_NMU1, S_RHO1,   // -N * mu1, r1 / N : row constants of the rank-1 batch-norm correction (rows = view-1 dims)
    S_NMU2, S_RHO2,   // same with the views swapped (row-block mode, C^T pass)
    S_COUNT
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

constexpr int kColsPerBlock = 64;   // 32 lanes x 2 columns
constexpr int kRowGroups = 32;      // 1024 threads: enough loads in flight to stream N x 64-column slabs at HBM/L2 speed
constexpr int kColThreads = kRowGroups * 32;

// ------------------------------------------------------------------------------------------
// 1. statistics
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kColThreads) bt_stats_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int N, int D, float eps,
                                                               float momentum, float* __restrict__ stats, __nv_bfloat16* __restrict__ zb1,
                                                               __nv_bfloat16* __restrict__ zb2, __half* __restrict__ zh1,
                                                               __half* __restrict__ zh2, float* __restrict__ running_mean,
                                                               float* __restrict__ running_var, double* __restrict__ loss_acc) {
    __shared__ float red[kRowGroups][5][kColsPerBlock];
    __shared__ float shift[2][kColsPerBlock];
    __shared__ float colstat[4][kColsPerBlock];   // mu1, r1, mu2, r2 of this block's columns
    __shared__ float on_red[2];
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * kColsPerBlock + lane * 2;
    float s1[2] = {0, 0}, q1[2] = {0, 0}, s2[2] = {0, 0}, q2[2] = {0, 0}, x12[2] = {0, 0};
    float k1[2] = {0, 0}, k2[2] = {0, 0};
    const bool ok = col < D;
    if (ok) {
        // shifted-data sums: subtract row 0 so that |mu| >> sigma does not cancel in fp32
        float2 a = Ld2<T>::ld(z1 + col), b = Ld2<T>::ld(z2 + col);
        k1[0] = bf16_round(a.x); k1[1] = bf16_round(a.y);
        k2[0] = bf16_round(b.x); k2[1] = bf16_round(b.y);
        if (rg == 0) {
            shift[0][lane * 2] = k1[0]; shift[0][lane * 2 + 1] = k1[1];
            shift[1][lane * 2] = k2[0]; shift[1][lane * 2 + 1] = k2[1];
        }
#pragma unroll 4
        for (int n = rg; n < N; n += kRowGroups) {
            float2 a2 = Ld2<T>::ld(z1 + (size_t)n * D + col), b2 = Ld2<T>::ld(z2 + (size_t)n * D + col);
            // the tensor cores consume bf16: statistics are those of the bf16-rounded embeddings
            float av[2] = {bf16_round(a2.x), bf16_round(a2.y)}, bv[2] = {bf16_round(b2.x), bf16_round(b2.y)};
            if (zb1 != nullptr) {
                Ld2<__nv_bfloat16>::st(zb1 + (size_t)n * D + col, av[0], av[1]);
                Ld2<__nv_bfloat16>::st(zb2 + (size_t)n * D + col, bv[0], bv[1]);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float da = av[c] - k1[c], db = bv[c] - k2[c];
                s1[c] += da; q1[c] = fmaf(da, da, q1[c]);
                s2[c] += db; q2[c] = fmaf(db, db, q2[c]);
                x12[c] = fmaf(da, db, x12[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        red[rg][0][lane * 2 + c] = s1[c]; red[rg][1][lane * 2 + c] = q1[c];
        red[rg][2][lane * 2 + c] = s2[c]; red[rg][3][lane * 2 + c] = q2[c];
        red[rg][4][lane * 2 + c] = x12[c];
    }
    __syncthreads();
    float on = 0.f;
    if (threadIdx.x < kColsPerBlock) {
        const int c = threadIdx.x, gc = blockIdx.x * kColsPerBlock + c;
        if (gc < D) {
            float t[5] = {0, 0, 0, 0, 0};
#pragma unroll
            for (int g = 0; g < kRowGroups; ++g)
#pragma unroll
                for (int k = 0; k < 5; ++k) t[k] += red[g][k][c];
            const float invN = 1.0f / (float)N;
            const float kk1 = shift[0][c], kk2 = shift[1][c];
            const float m1 = t[0] * invN, m2 = t[2] * invN;
            const float var1 = fmaxf(t[1] * invN - m1 * m1, 0.f), var2 = fmaxf(t[3] * invN - m2 * m2, 0.f);
            const float cov = t[4] * invN - m1 * m2;
            const float mu1 = kk1 + m1, mu2 = kk2 + m2;
            const float r1 = rsqrtf(var1 + eps), r2 = rsqrtf(var2 + eps);
            // one Newton step: rsqrtf is ~2 ulp, BatchNorm uses a correctly rounded 1/sqrt
            const float r1n = r1 * (1.5f - 0.5f * (var1 + eps) * r1 * r1), r2n = r2 * (1.5f - 0.5f * (var2 + eps) * r2 * r2);
            const float cd = cov * r1n * r2n;
            stats[S_MU1 * D + gc] = mu1; stats[S_R1 * D + gc] = r1n;
            stats[S_MU2 * D + gc] = mu2; stats[S_R2 * D + gc] = r2n;
            stats[S_CDIAG * D + gc] = cd;
            stats[S_NMU1 * D + gc] = -(float)N * mu1; stats[S_RHO1 * D + gc] = r1n * invN;
            stats[S_NMU2 * D + gc] = -(float)N * mu2; stats[S_RHO2 * D + gc] = r2n * invN;
            colstat[0][c] = mu1; colstat[1][c] = r1n; colstat[2][c] = mu2; colstat[3][c] = r2n;
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
        // on-diagonal loss sum_i (C_ii - 1)^2: one double atomic per block
        on = warp_sum(on);
        if (lane == 0) on_red[threadIdx.x >> 5] = on;
    }
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(loss_acc + 2, (double)(on_red[0] + on_red[1]));
    // second pass: standardised embeddings in fp16 (operand B of the gradient GEMMs; |zh| <= sqrt(N))
    if (ok && zh1 != nullptr) {
        const float m1[2] = {colstat[0][lane * 2], colstat[0][lane * 2 + 1]}, q1r[2] = {colstat[1][lane * 2], colstat[1][lane * 2 + 1]};
        const float m2[2] = {colstat[2][lane * 2], colstat[2][lane * 2 + 1]}, q2r[2] = {colstat[3][lane * 2], colstat[3][lane * 2 + 1]};
#pragma unroll 4
        for (int n = rg; n < N; n += kRowGroups) {
            const size_t o = (size_t)n * D + col;
            const float2 a2 = Ld2<T>::ld(z1 + o), b2 = Ld2<T>::ld(z2 + o);
            Ld2<__half>::st(zh1 + o, (bf16_round(a2.x) - m1[0]) * q1r[0], (bf16_round(a2.y) - m1[1]) * q1r[1]);
            Ld2<__half>::st(zh2 + o, (bf16_round(b2.x) - m2[0]) * q2r[0], (bf16_round(b2.y) - m2[1]) * q2r[1]);
        }
    }
}

// row sums of the standardised embeddings (HSIC only): R[n] = sum_j zh[n, j]
__global__ void __launch_bounds__(256) bt_rowsum_kernel(const __nv_bfloat16* __restrict__ z, int N, int D, const float* __restrict__ mu,
                                                        const float* __restrict__ r, float* __restrict__ out) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) acc += (__bfloat162float(z[(size_t)n * D + c]) - mu[c]) * r[c];
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
// 2./3. tcgen05 GEMM kernel (persistent, warp specialised)
// ------------------------------------------------------------------------------------------
constexpr int BM = 128;            // UMMA M (TMEM lanes)
constexpr int BK = 64;             // K elements per pipeline stage (one 128-byte swizzle row)
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;       // 16 KiB
constexpr int kBBytesMax = 256 * BK * 2;   // 32 KiB
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kAccCols = 256;      // TMEM columns per accumulator stage
constexpr int kNumThreads = 384;   // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-11 epilogue
constexpr int kEpiWarps = 8;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

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
    // CORR epilogue: c = (S + row_nmu[i] * col_mu[j]) * row_rho[i] * col_r[j]
    const float* row_nmu; const float* row_rho; const float* col_mu; const float* col_r;
    int accumulate_loss;
    __half* c_out;            // CORR: row-major (rows local to row0) x D, fp16, diagonal zeroed
    float* g_out;             // GRAD: g_out[n * ldg + (row - row0)]
    int ldg;
    // Column-blocked C (multi-GPU, Dr = D / world): element (local row i, column j) lives at ((j / Dr) * Dr + i) * Dr + j % Dr, so
    // that block q = C[rows, q Dr : (q+1) Dr] is contiguous for the all-to-all.  CORR writes it, the K-major GRAD pass reads it
    // through a (world * Dr) x Dr tensor map.  0 = plain row-major.
    int blocked_dr;
    float inv_n;              // CORR on standardised fp16 operands (row_nmu == nullptr): c = S * inv_n
};

struct UmmaParams {
    DescCfg dc;
    int mode;          // 0 = CORR, 1 = GRAD
    int D, N;
    int tiles_m, tiles_n, splits, kblocks;   // per pass
    int pass_count;
    int bn;            // UMMA N of this launch (multiple of 16, <= 256)
    int ab_format;     // operand type of this launch: 1 = bf16, 0 = fp16
    int hsic;
    int write_c;
    double* loss_acc;
    PassCfg pass[2];
};

__device__ __forceinline__ void decode_work(const UmmaParams& p, int w, int& pass, int& tm, int& tn, int& kb0, int& kb1) {
    const int per_pass = p.tiles_m * p.tiles_n * p.splits;
    pass = w / per_pass;
    int r = w % per_pass;
    const int split = r % p.splits; r /= p.splits;
    tn = r % p.tiles_n; tm = r / p.tiles_n;
    const int per = (p.kblocks + p.splits - 1) / p.splits;
    kb0 = split * per;
    kb1 = min(p.kblocks, kb0 + per);
}

__global__ void __launch_bounds__(kNumThreads, 1)
bt_umma_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1, const UmmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_work = p.tiles_m * p.tiles_n * p.splits * p.pass_count;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapB0);
        if (p.pass_count > 1) { tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapB1); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Operand "major-ness": CORR reads both raw embeddings as MN-major tiles (B too); GRAD reads C K-major (rows of C) or
    // MN-major (the same C as C^T) and its B operand (standardised fp16 embeddings, one row per sample) K-major.
    const bool b_mn = (p.mode == 0);
    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            int pass, tm, tn, kb0, kb1;
            decode_work(p, w, pass, tm, tn, kb0, kb1);
            const PassCfg& pc = p.pass[pass];
            const CUtensorMap* mA = pass == 1 ? &mapA1 : &mapA0;
            const CUtensorMap* mB = pass == 1 ? &mapB1 : &mapB0;
            const uint32_t tx = kABytes + (uint32_t)p.bn * BK * 2;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * kStageBytes;
                uint8_t* sB = sA + kABytes;
                mbar_expect_tx(&full_bar[stage], tx);
                if (pc.a_mn) {
                    // the tensor map spans global dimension indices (columns of z or of the full C)
                    tma_load_2d(sA, mA, &full_bar[stage], pc.row0 + tm * BM, kb * BK);
                    tma_load_2d(sA + 8192, mA, &full_bar[stage], pc.row0 + tm * BM + 64, kb * BK);
                } else {
                    // the tensor map spans the rows of the (possibly row-block compact / column-blocked) C matrix
                    if (pc.blocked_dr > 0) {
                        const int qb = (kb * BK) / pc.blocked_dr;
                        tma_load_2d(sA, mA, &full_bar[stage], kb * BK - qb * pc.blocked_dr, qb * pc.blocked_dr + tm * BM);
                    } else {
                        tma_load_2d(sA, mA, &full_bar[stage], kb * BK, tm * BM);
                    }
                }
                if (b_mn) {
                    for (int c = 0; c < p.bn / 64; ++c) tma_load_2d(sB + c * 8192, mB, &full_bar[stage], tn * p.bn + c * 64, kb * BK);
                } else {
                    tma_load_2d(sB, mB, &full_bar[stage], kb * BK, tn * p.bn);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer (one thread) =================
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            int pass, tm, tn, kb0, kb1;
            decode_work(p, w, pass, tm, tn, kb0, kb1);
            const bool a_mn = p.pass[pass].a_mn != 0;
            const uint32_t idesc = make_idesc_f16(BM, p.bn, a_mn ? 1 : 0, b_mn ? 1 : 0, p.ab_format);
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccCols;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + stage * kStageBytes);
                const uint32_t sB = sA + kABytes;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                    const uint64_t da = a_mn ? make_smem_desc_sw128(sA + ks * p.dc.mn_kstep, p.dc.mn_lbo, p.dc.mn_sbo)
                                             : make_smem_desc_sw128(sA + ks * p.dc.k_kstep, p.dc.k_lbo, p.dc.k_sbo);
                    const uint64_t db = b_mn ? make_smem_desc_sw128(sB + ks * p.dc.mn_kstep, p.dc.mn_lbo, p.dc.mn_sbo)
                                             : make_smem_desc_sw128(sB + ks * p.dc.k_kstep, p.dc.k_lbo, p.dc.k_sbo);
                    umma_bf16_ss(d_tmem, da, db, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull_bar[acc]);         // accumulator complete -> epilogue
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> registers -> global =================
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int hf = (warp - 4) >> 2;      // which half of the 32-column chunks
        int acc = 0; uint32_t acc_phase = 0;
        const int D = p.D;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            int pass, tm, tn, kb0, kb1;
            decode_work(p, w, pass, tm, tn, kb0, kb1);
            const PassCfg& pc = p.pass[pass];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * kAccCols + (static_cast<uint32_t>(q * 32) << 16);
            const int lrow = tm * BM + q * 32 + lane;    // row inside this pass's tile grid
            const int row = pc.row0 + lrow;              // global dimension index owned by this thread
            const bool row_ok = row < pc.row_end;
            if (p.mode == 0) {
                // ---- CORR: v = S - N mu_i mu'_j;  c = v (r_i / N) r'_j   (batch-norm as a rank-1 correction)
                const bool raw = pc.row_nmu != nullptr;      // raw bf16 operands: apply batch-norm here; else operands are standardised
                const float nmu = (raw && row_ok) ? pc.row_nmu[row] : 0.f;
                const float rho = (raw && row_ok) ? pc.row_rho[row] : pc.inv_n;
                float l2 = 0.f, l1 = 0.f;
                const int nchunks = p.bn / 32;
                for (int ch = hf; ch < nchunks; ch += 2) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + ch * 32, r);
                    tmem_ld_wait();
                    const int j0 = tn * p.bn + ch * 32;
                    if (j0 < D && row_ok) {
                        uint32_t packed[16];
#pragma unroll
                        for (int t4 = 0; t4 < 8; ++t4) {
                            float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(1.f, 1.f, 1.f, 1.f);
                            if (raw) {
                                m4 = __ldg(reinterpret_cast<const float4*>(pc.col_mu + j0) + t4);
                                b4 = __ldg(reinterpret_cast<const float4*>(pc.col_r + j0) + t4);
                            }
                            const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
                            float cc[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int t = t4 * 4 + u;
                                const float v = fmaf(nmu, mm[u], __uint_as_float(r[t]));
                                float c = (v * rho) * bb[u];
                                c = (j0 + t == row) ? 0.f : c;      // diagonal handled in fp32 by the finalize kernel
                                l2 = fmaf(c, c, l2);
                                l1 += c;
                                cc[u] = c;
                            }
                            packed[t4 * 2] = pack_f16x2(cc[0], cc[1]);
                            packed[t4 * 2 + 1] = pack_f16x2(cc[2], cc[3]);
                        }
                        if (p.write_c) {
                            size_t coff = (size_t)lrow * D + j0;
                            if (pc.blocked_dr > 0) {
                                const int qb = j0 / pc.blocked_dr;
                                coff = ((size_t)qb * pc.blocked_dr + lrow) * pc.blocked_dr + (j0 - qb * pc.blocked_dr);
                            }
                            uint4* dst = reinterpret_cast<uint4*>(pc.c_out + coff);
#pragma unroll
                            for (int k = 0; k < 4; ++k) dst[k] = make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                        }
                    }
                }
                if (pc.accumulate_loss) {
                    l2 = warp_sum(l2);
                    if (p.hsic) l1 = warp_sum(l1);
                    if (lane == 0) {
                        atomicAdd(p.loss_acc + 0, (double)l2);
                        if (p.hsic) atomicAdd(p.loss_acc + 1, (double)l1);
                    }
                }
            } else {
                // ---- GRAD: fp32 accumulator (dimension row, sample n) -> g[n][row - row0]
                float* g = pc.g_out;
                const int nchunks = (p.bn + 31) / 32;
                for (int ch = hf; ch < nchunks; ch += 2) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + ch * 32, r);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const int n = tn * p.bn + ch * 32 + t;
                            if (ch * 32 + t < p.bn && n < p.N) {
                                float* dst = g + (size_t)n * pc.ldg + lrow;
                                if (p.splits > 1) atomicAdd(dst, __uint_as_float(r[t]));
                                else *dst = __uint_as_float(r[t]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// 4. finalize: diagonal term (fp32), batch-norm backward, cast, loss scalar
//    handles columns [col_begin, col_begin + col_count); g and dz are indexed with the local column and stride ld
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kColThreads) bt_finalize_kernel(const T* __restrict__ z1, const T* __restrict__ z2, int N, int D, int col_begin,
                                                                  int col_count, int ld, float alpha, float lambda, int hsic, float grad_scale,
                                                                  int need_mask, const float* __restrict__ stats, const float* __restrict__ g1,
                                                                  const float* __restrict__ g2, const float* __restrict__ rowsum1,
                                                                  const float* __restrict__ rowsum2, T* __restrict__ dz1, T* __restrict__ dz2,
                                                                  double* __restrict__ loss_acc, unsigned int* __restrict__ counters,
                                                                  float* __restrict__ loss_out) {
    __shared__ float red[kRowGroups][4][kColsPerBlock];
    __shared__ float mean_s[4][kColsPerBlock];
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int lcol = blockIdx.x * kColsPerBlock + lane * 2;     // local column (index into g / dz)
    const int col = col_begin + lcol;                            // global dimension index (index into z / stats)
    const bool ok = lcol < col_count;
    const float invN = 1.0f / (float)N;
    float mu1[2], r1[2], mu2[2], r2[2], gd[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int gc = ok ? col + c : 0;
        mu1[c] = stats[S_MU1 * D + gc]; r1[c] = stats[S_R1 * D + gc];
        mu2[c] = stats[S_MU2 * D + gc]; r2[c] = stats[S_R2 * D + gc];
        gd[c] = 2.0f * alpha * (stats[S_CDIAG * D + gc] - 1.0f) * invN;   // G_ii / N
    }
    const float hs = 2.0f * lambda * invN;

    // d loss / d zh1[n,i] = (2 lambda/N) sum_{j != i} C_ij zh2[n,j]  +  (G_ii/N) zh2[n,i]
    //                       (+ HSIC: (2 lambda/N)(R2[n] - zh2[n,i]), the "+1" of every off-diagonal G_ij)
    auto side_grad = [&](int c, float zh_other, float graw, float rs_other) -> float {
        float g = hs * graw + gd[c] * zh_other;
        if (hsic) g += hs * (rs_other - zh_other);
        return g;
    };

    if (need_mask != 0) {
        float a1[2] = {0, 0}, b1[2] = {0, 0}, a2[2] = {0, 0}, b2[2] = {0, 0};
        if (ok) {
#pragma unroll 2
            for (int n = rg; n < N; n += kRowGroups) {
                const size_t oz = (size_t)n * D + col, og = (size_t)n * ld + lcol;
                const float2 za = Ld2<T>::ld(z1 + oz), zb = Ld2<T>::ld(z2 + oz);
                const float zav[2] = {bf16_round(za.x), bf16_round(za.y)}, zbv[2] = {bf16_round(zb.x), bf16_round(zb.y)};
                float2 ga = make_float2(0, 0), gb = make_float2(0, 0);
                if (need_mask & 1) ga = *reinterpret_cast<const float2*>(g1 + og);
                if (need_mask & 2) gb = *reinterpret_cast<const float2*>(g2 + og);
                const float gav[2] = {ga.x, ga.y}, gbv[2] = {gb.x, gb.y};
                const float rs1 = hsic ? rowsum1[n] : 0.f, rs2 = hsic ? rowsum2[n] : 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float zh1 = (zav[c] - mu1[c]) * r1[c], zh2 = (zbv[c] - mu2[c]) * r2[c];
                    const float gf1 = side_grad(c, zh2, gav[c], rs2);
                    const float gf2 = side_grad(c, zh1, gbv[c], rs1);
                    a1[c] += gf1; b1[c] = fmaf(gf1, zh1, b1[c]);
                    a2[c] += gf2; b2[c] = fmaf(gf2, zh2, b2[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            red[rg][0][lane * 2 + c] = a1[c]; red[rg][1][lane * 2 + c] = b1[c];
            red[rg][2][lane * 2 + c] = a2[c]; red[rg][3][lane * 2 + c] = b2[c];
        }
        __syncthreads();
        if (threadIdx.x < kColsPerBlock) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = 0.f;
#pragma unroll
                for (int g = 0; g < kRowGroups; ++g) t += red[g][k][threadIdx.x];
                mean_s[k][threadIdx.x] = t * invN;
            }
        }
        __syncthreads();
        if (ok) {
#pragma unroll 2
            for (int n = rg; n < N; n += kRowGroups) {
                const size_t oz = (size_t)n * D + col, og = (size_t)n * ld + lcol;
                const float2 za = Ld2<T>::ld(z1 + oz), zb = Ld2<T>::ld(z2 + oz);
                const float zav[2] = {bf16_round(za.x), bf16_round(za.y)}, zbv[2] = {bf16_round(zb.x), bf16_round(zb.y)};
                float2 ga = make_float2(0, 0), gb = make_float2(0, 0);
                if (need_mask & 1) ga = *reinterpret_cast<const float2*>(g1 + og);
                if (need_mask & 2) gb = *reinterpret_cast<const float2*>(g2 + og);
                const float gav[2] = {ga.x, ga.y}, gbv[2] = {gb.x, gb.y};
                const float rs1 = hsic ? rowsum1[n] : 0.f, rs2 = hsic ? rowsum2[n] : 0.f;
                float o1[2], o2[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float zh1 = (zav[c] - mu1[c]) * r1[c], zh2 = (zbv[c] - mu2[c]) * r2[c];
                    const float gf1 = side_grad(c, zh2, gav[c], rs2);
                    const float gf2 = side_grad(c, zh1, gbv[c], rs1);
                    const int cc = lane * 2 + c;
                    o1[c] = r1[c] * (gf1 - mean_s[0][cc] - zh1 * mean_s[1][cc]) * grad_scale;
                    o2[c] = r2[c] * (gf2 - mean_s[2][cc] - zh2 * mean_s[3][cc]) * grad_scale;
                }
                if (need_mask & 1) Ld2<T>::st(dz1 + og, o1[0], o1[1]);
                if (need_mask & 2) Ld2<T>::st(dz2 + og, o2[0], o2[1]);
            }
        }
    }
    // single-GPU: the last block folds the three partial sums into the loss scalar
    if (loss_out != nullptr && threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(counters, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            const double off2 = *((volatile double*)(loss_acc + 0));
            const double off1 = *((volatile double*)(loss_acc + 1));
            const double ond = *((volatile double*)(loss_acc + 2));
            double off = off2;
            if (hsic) off = off2 + 2.0 * off1 + (double)D * (double)(D - 1);
            *loss_out = (float)((double)alpha * ond + (double)lambda * off);
        }
    }
}

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

// Optional per-call device timing of the two tensor-core launches (bench.py roofline): events are recorded on
// the launching stream around CORR and GRAD of every loss call while enabled.
constexpr int kTimingRing = 512;
struct TimingState {
    bool enabled = false;
    int count = 0;
    cudaEvent_t ev[kTimingRing][3];
    bool created = false;
};
static TimingState g_timing;

// Workspace layout.  `rows` = number of C rows this call materialises (D single-GPU, row_count in row-block mode);
// `two_c` = a second C block for the transposed pass (row-block mode).
struct WsLayout {
    size_t misc, stats, rs1, rs2, g1, g2, zb1, zb2, zh1, zh2, c1, c2, total;
};

static WsLayout ws_layout(int N, int D, int rows, int dtype, bool two_c) {
    WsLayout L{};
    size_t off = 0;
    L.misc = off; off += 256;
    L.stats = off; off = align_up(off + sizeof(float) * S_COUNT * (size_t)D, 256);
    L.rs1 = off; off = align_up(off + sizeof(float) * (size_t)N, 256);
    L.rs2 = off; off = align_up(off + sizeof(float) * (size_t)N, 256);
    L.g1 = off; off = align_up(off + sizeof(float) * (size_t)N * rows, 256);
    L.g2 = off; off = align_up(off + sizeof(float) * (size_t)N * rows, 256);
    L.zb1 = off; if (dtype != ABT_DTYPE_BF16) off = align_up(off + 2 * (size_t)N * D, 256);
    L.zb2 = off; if (dtype != ABT_DTYPE_BF16) off = align_up(off + 2 * (size_t)N * D, 256);
    L.zh1 = off; off = align_up(off + 2 * (size_t)N * D, 256);
    L.zh2 = off; off = align_up(off + 2 * (size_t)N * D, 256);
    L.c1 = off; off = align_up(off + 2 * (size_t)rows * D, 256);
    L.c2 = off; if (two_c) off = align_up(off + 2 * (size_t)rows * D, 256);
    L.total = off;
    return L;
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

static int ensure_umma_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(bt_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    return 0;
}

static void launch_umma(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1, const UmmaParams& p,
                        cudaStream_t stream) {
    const int total = p.tiles_m * p.tiles_n * p.splits * p.pass_count;
    const int grid = total < num_sms() ? total : num_sms();
    bt_umma_kernel<<<grid, kNumThreads, kSmemBytes, stream>>>(a0, b0, a1, b1, p);
    count_launch();
}

// Everything one loss evaluation needs, for both the single-GPU and the row-block entry points.
struct LossCall {
    const void* z1; const void* z2; int dtype;
    int N, D;
    int row_begin, row_count;   // dimensions owned by this call (0, D single-GPU)
    bool rows_mode;             // compute the C^T row block with a second CORR pass instead of reading C transposed
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
    double* loss_acc = reinterpret_cast<double*>(ws + L.misc);
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws + L.misc + 64);
    float* g1 = reinterpret_cast<float*>(ws + L.g1);
    float* g2 = reinterpret_cast<float*>(ws + L.g2);
    float* rs1 = reinterpret_cast<float*>(ws + L.rs1);
    float* rs2 = reinterpret_cast<float*>(ws + L.rs2);
    __half* C1 = reinterpret_cast<__half*>(ws + L.c1);
    __half* C2 = reinterpret_cast<__half*>(ws + L.c2);
    __half* zh1 = reinterpret_cast<__half*>(ws + L.zh1);
    __half* zh2 = reinterpret_cast<__half*>(ws + L.zh2);
    const bool is_bf16 = (a.dtype == ABT_DTYPE_BF16);
    __nv_bfloat16* zb1 = is_bf16 ? nullptr : reinterpret_cast<__nv_bfloat16*>(ws + L.zb1);
    __nv_bfloat16* zb2 = is_bf16 ? nullptr : reinterpret_cast<__nv_bfloat16*>(ws + L.zb2);
    const __nv_bfloat16* zq1 = is_bf16 ? static_cast<const __nv_bfloat16*>(a.z1) : zb1;
    const __nv_bfloat16* zq2 = is_bf16 ? static_cast<const __nv_bfloat16*>(a.z2) : zb2;
    const int need = a.need & 3;
    const int col_blocks = (D + kColsPerBlock - 1) / kColsPerBlock;
    if (int rc = ensure_umma_attr()) return rc;

    cudaMemsetAsync(ws + L.misc, 0, 128, stream);     // loss partial sums + block counter
    bt_stats_kernel<T><<<col_blocks, kColThreads, 0, stream>>>(static_cast<const T*>(a.z1), static_cast<const T*>(a.z2), N, D, a.eps, a.momentum,
                                                                stats, zb1, zb2, need != 0 ? zh1 : nullptr, need != 0 ? zh2 : nullptr,
                                                                a.running_mean, a.running_var, loss_acc);
    count_launch();
    if (a.hsic) {
        bt_rowsum_kernel<<<N, 256, 0, stream>>>(zq1, N, D, stats + S_MU1 * D, stats + S_R1 * D, rs1);
        bt_rowsum_kernel<<<N, 256, 0, stream>>>(zq2, N, D, stats + S_MU2 * D, stats + S_R2 * D, rs2);
        count_launch(2);
    }

    const bool timed = g_timing.enabled && g_timing.count < kTimingRing;
    cudaEvent_t* tev = timed ? g_timing.ev[g_timing.count] : nullptr;
    const int row_tiles = (RC + BM - 1) / BM;
    // ---- CORR: C[rows, :] (and, in row-block mode, C^T[rows, :] with the views swapped)
    if (timed) cudaEventRecord(tev[0], stream);
    {
        CUtensorMap m1, m2;
        if (int rc = make_map_16(&m1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, zq1, N, D, 64, 64)) return rc;
        if (int rc = make_map_16(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, zq2, N, D, 64, 64)) return rc;
        UmmaParams p{};
        p.dc = g_desc;
        p.mode = 0; p.D = D; p.N = N;
        p.bn = 256; p.ab_format = 1;
        p.tiles_m = row_tiles; p.tiles_n = (D + p.bn - 1) / p.bn; p.splits = 1;
        p.kblocks = (N + BK - 1) / BK;
        p.hsic = a.hsic; p.write_c = need != 0;
        p.loss_acc = loss_acc;
        p.pass[0] = PassCfg{1, R0, R0 + RC, stats + S_NMU1 * D, stats + S_RHO1 * D, stats + S_MU2 * D, stats + S_R2 * D, 1, C1, nullptr, 0, 0, 0.f};
        p.pass[1] = PassCfg{1, R0, R0 + RC, stats + S_NMU2 * D, stats + S_RHO2 * D, stats + S_MU1 * D, stats + S_R1 * D, 0, C2, nullptr, 0, 0, 0.f};
        // the transposed block is only needed for dz2
        p.pass_count = (a.rows_mode && (need & 2)) ? 2 : 1;
        launch_umma(m1, m2, m2, m1, p, stream);
    }
    if (timed) cudaEventRecord(tev[1], stream);
    // ---- GRAD
    if (need != 0) {
        const int bn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
        CUtensorMap mCk, mCt, mZ2, mZ1;
        if (int rc = make_map_16(&mCk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, RC, D, 64, 128)) return rc;
        if (a.rows_mode) {
            if (int rc = make_map_16(&mCt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C2, RC, D, 64, 128)) return rc;    // K-major rows of C^T
        } else {
            if (int rc = make_map_16(&mCt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C1, RC, D, 64, 64)) return rc;     // the same C read MN-major
        }
        if (int rc = make_map_16(&mZ2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, zh2, N, D, 64, bn)) return rc;
        if (int rc = make_map_16(&mZ1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, zh1, N, D, 64, bn)) return rc;
        UmmaParams p{};
        p.dc = g_desc;
        p.mode = 1; p.D = D; p.N = N; p.bn = bn; p.ab_format = 0;
        p.tiles_m = row_tiles; p.tiles_n = (N + bn - 1) / bn;
        p.kblocks = (D + BK - 1) / BK;
        const PassCfg pass_dz1{0, R0, R0 + RC, nullptr, nullptr, nullptr, nullptr, 0, nullptr, g1, RC, 0, 0.f};
        const PassCfg pass_dz2{a.rows_mode ? 0 : 1, R0, R0 + RC, nullptr, nullptr, nullptr, nullptr, 0, nullptr, g2, RC, 0, 0.f};
        p.pass_count = (need == 3) ? 2 : 1;
        const CUtensorMap *a0, *b0, *a1, *b1;
        if (need & 1) { p.pass[0] = pass_dz1; a0 = &mCk; b0 = &mZ2; p.pass[1] = pass_dz2; a1 = &mCt; b1 = &mZ1; }
        else { p.pass[0] = pass_dz2; a0 = &mCt; b0 = &mZ1; p.pass[1] = pass_dz2; a1 = &mCt; b1 = &mZ1; }
        const int tiles = p.tiles_m * p.tiles_n * p.pass_count;
        int splits = num_sms() / tiles;
        if (splits < 1) splits = 1;
        while (splits > 1 && p.kblocks / splits < 8) --splits;      // keep >= 8 k-blocks per split
        {   // no empty split: every work item must issue at least one MMA
            const int per = (p.kblocks + splits - 1) / splits;
            splits = (p.kblocks + per - 1) / per;
        }
        p.splits = splits;
        p.hsic = a.hsic; p.write_c = 0;
        p.loss_acc = loss_acc;
        if (splits > 1) {
            if (need & 1) cudaMemsetAsync(g1, 0, sizeof(float) * (size_t)N * RC, stream);
            if (need & 2) cudaMemsetAsync(g2, 0, sizeof(float) * (size_t)N * RC, stream);
        }
        launch_umma(*a0, *b0, *a1, *b1, p, stream);
    }
    if (timed) { cudaEventRecord(tev[2], stream); ++g_timing.count; }
    // ---- finalize (columns of this call's dimension block)
    if (need != 0 || a.loss_out != nullptr) {
        const int fblocks = (RC + kColsPerBlock - 1) / kColsPerBlock;
        bt_finalize_kernel<T><<<fblocks, kColThreads, 0, stream>>>(static_cast<const T*>(a.z1), static_cast<const T*>(a.z2), N, D, R0, RC, a.ld_dz,
                                                                   a.alpha, a.lambda, a.hsic, a.grad_scale, need, stats, g1, g2, rs1, rs2,
                                                                   static_cast<T*>(a.dz1), static_cast<T*>(a.dz2), loss_acc, counters, a.loss_out);
        count_launch();
    }
    if (a.loss_parts_out != nullptr) cudaMemcpyAsync(a.loss_parts_out, loss_acc, 3 * sizeof(double), cudaMemcpyDeviceToDevice, stream);
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
            for (int k = 0; k < 3; ++k)
                if (cudaEventCreate(&g_timing.ev[i][k]) != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaEventCreate failed");
        g_timing.created = true;
    }
    g_timing.enabled = enable != 0;
    g_timing.count = 0;
    return 0;
}

// Average milliseconds of the CORR and GRAD launches recorded since abt_debug_timing(1); synchronises on the events.
extern "C" int abt_debug_timing_read(float* corr_ms, float* grad_ms, int* n_calls) {
    double c = 0, g = 0;
    const int n = g_timing.count;
    for (int i = 0; i < n; ++i) {
        float a = 0, b = 0;
        if (cudaEventSynchronize(g_timing.ev[i][2]) != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaEventSynchronize failed");
        cudaEventElapsedTime(&a, g_timing.ev[i][0], g_timing.ev[i][1]);
        cudaEventElapsedTime(&b, g_timing.ev[i][1], g_timing.ev[i][2]);
        c += a; g += b;
    }
    if (corr_ms) *corr_ms = n ? (float)(c / n) : 0.f;
    if (grad_ms) *grad_ms = n ? (float)(g / n) : 0.f;
    if (n_calls) *n_calls = n;
    return 0;
}

extern "C" int abt_debug_set(int key, int value) {
    int* f[6] = {&g_desc.mn_lbo, &g_desc.mn_sbo, &g_desc.mn_kstep, &g_desc.k_lbo, &g_desc.k_sbo, &g_desc.k_kstep};
    if (key < 0 || key >= 6) return set_error(ABT_ERR_ARG, "unknown debug key %d", key);
    *f[key] = value;
    return 0;
}

// Debug view of the single-GPU workspace layout (byte offsets), used by tools/gpu_diag.py only.
extern "C" int abt_debug_ws_offsets(int n_rows, int n_dims, int dtype, size_t* out8) {
    const WsLayout L = ws_layout(n_rows, n_dims, n_dims, dtype, false);
    out8[0] = L.stats; out8[1] = L.c1; out8[2] = L.g1; out8[3] = L.g2; out8[4] = L.zb1; out8[5] = L.zb2; out8[6] = L.misc; out8[7] = L.total;
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
