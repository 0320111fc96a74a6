// Barlow Twins objective for small batches (N <= 128 rows): ONE tensor-core launch, the D x D matrix never leaves the SM.
//
// Replaces utils/loss.py:17-30 and its autograd backward for the reference's own batch sizes (main.py:115-119 with
// batch 32 ... 128 per GPU; BASELINE config 3: N = 128, D = 2048 / 4096 / 8192).  The two-launch CORR + GRAD form
// (bt_umma_kernel) writes C once and reads it twice; at N = 128 that traffic (3 x 128 MiB at D = 8192) IS the run time.
// Here a CTA owns a 128-row block I of C (and, in the second pass, of C^T) and walks the column blocks J:
//
//     S    = zh_a[:, I]^T zh_b[:, J]            MMA1: K = N, fp16 standardised operands, fp32 accumulator in TMEM
//     loss += sum S^2 (off-diagonal)            epilogue warps, straight out of TMEM
//     P    = fp16(S), diagonal zeroed           -> shared memory (128-byte swizzle, K-major operand A of MMA2)
//     O   += P zh_b[:, J]^T                     MMA2: K = 128 columns of J; O (128 x N fp32) stays in TMEM for the whole walk
//
// and finishes with batch-norm backward on O: dz_a[:, I] = r_a (g - mean_n g - zh_a mean_n(g o zh_a)),
// g = (2 lambda / N^2) O + (G_ii / N) zh_b[:, I]  (+ the HSIC row-sum term), the diagonal in fp32 from the statistics pass.
// This is the shape of an attention forward pass without the softmax ("sequence" = D, "head dimension" = N); the same
// zh_b tile in shared memory serves MMA1 as an MN-major B operand and MMA2 as a K-major one.  Executed FLOP: 8 N D^2 for
// both gradients (S is recomputed by the second pass) against 6 N D^2 algorithmic.
//
// Warp roles (384 threads, 1 CTA / SM): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue.
// The walk over J starts at block I + 1 and ends AT block I: the last zh_b tile is zh_b[:, I], which the final epilogue
// needs (diagonal term), and it is still resident in its ring stage.
#pragma once

namespace abt {

constexpr int FB = 128;                       // rows of C per CTA (UMMA M) and columns per step
constexpr int kFStages = 4;                   // ring of zh_b tiles
constexpr int kFTileBytes = FB * 128 * 2;     // 32 KiB: up to 128 samples x 128 columns, fp16
constexpr int kFOffQ = 0;
constexpr int kFOffK = kFTileBytes;
constexpr int kFOffP = kFOffK + kFStages * kFTileBytes;
constexpr int kFOffBar = kFOffP + 2 * kFTileBytes;
constexpr int kFSmemBytes = kFOffBar + 256 + 1024 /*align slack*/;
static_assert(kFSmemBytes <= 227 * 1024, "fused kernel: shared memory budget");

struct FusedParams {
    int D, N, n_pad;             // n_pad = N rounded up to 16 (UMMA K of MMA1, UMMA N of MMA2); rows >= N are zero-filled by TMA
    int n_blocks;                // ceil(D / 128): row blocks = column blocks
    int pass_count;              // 1 or 2
    int pass_side[2];            // 0: the pass produces dz1 (rows = view-1 dimensions), 1: dz2
    int hsic, io_dtype;
    float alpha, lambda, grad_scale;
    const float* stats;          // StatSlot arrays
    const float* rs1; const float* rs2;      // HSIC: row sums of zh1 / zh2
    void* dz1; void* dz2;
    double* loss_acc;            // [0] += sum_{i != j} c_ij^2, [1] += sum_{i != j} c_ij (HSIC); [2] = on-diagonal sum (statistics pass)
    unsigned int* done_counter;  // zeroed per call: the last CTA to finish publishes the loss
    float* loss_out;             // may be null
    const float* ondiag_part; int n_parts;   // on-diagonal loss: one partial sum per block of the statistics kernel
    const __half* zimg1; const __half* zimg2;    // version 3: standardised embeddings as tile images (null: row-major, read through the tensor maps)
    int n_stages;                // version 3: depth of the zh_b tile ring (<= kXStages)
    int debug;                   // timing experiments only (results are wrong when non-zero): 1 = issuer ignores p_full, 2 = epilogue skips the
                                 // S -> P conversion, 4 = no MMA2, 8 = no MMA1, 64 = no unit epilogue, 128 = no unit prologue, 256 = no tile loads
};

__device__ __forceinline__ float lds_half(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return __half2float(__ushort_as_half(v));
}

// address of element (sample n, local column il) of a TMA-written zh tile: two 64-column chunks of n_pad rows x 128 bytes, 128-byte swizzle
__device__ __forceinline__ uint32_t tile_elem_addr(uint32_t base, uint32_t chunk_bytes, int n, int il) {
    return base + (uint32_t)(il >> 6) * chunk_bytes + (uint32_t)n * 128u + ((((uint32_t)(il & 63) >> 3) ^ ((uint32_t)n & 7u)) << 4) + (uint32_t)(il & 7) * 2u;
}

template <typename T>
__device__ __forceinline__ void fused_store_chunk(const uint32_t (&g)[32], T* dz, int ld, int i, int n0, int N, uint32_t q_s, uint32_t chunk_bytes, int il,
                                                  float mg, float b, float rg) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        const int n = n0 + t;
        if (n < N) {
            const float zs = lds_half(tile_elem_addr(q_s, chunk_bytes, n, il));
            store_out<T>(dz + (size_t)n * ld + i, (__uint_as_float(g[t]) - mg - zs * b) * rg);
        }
    }
}

__global__ void __launch_bounds__(kNumThreads, 1)
bt_fused_kernel(const __grid_constant__ CUtensorMap mapZ1, const __grid_constant__ CUtensorMap mapZ2, const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* k_full = reinterpret_cast<uint64_t*>(smem + kFOffBar);
    uint64_t* k_empty = k_full + kFStages;
    uint64_t* s_full = k_empty + kFStages;      // [2] MMA1 done -> epilogue
    uint64_t* p_full = s_full + 2;              // [2] P written (and S read) -> MMA issuer
    uint64_t* p_empty = p_full + 2;             // [2] MMA2 done reading P
    uint64_t* o_full = p_empty + 2;             // [2] all MMAs of the unit done
    uint64_t* o_empty = o_full + 2;             // [2] final epilogue done with the O buffer
    uint64_t* q_full = o_empty + 2;
    uint64_t* q_empty = q_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);
    const uint32_t q_s = smem_u32(smem + kFOffQ), k_s = smem_u32(smem + kFOffK), p_s = smem_u32(smem + kFOffP);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.n_blocks * p.pass_count;
    const uint32_t chunk_bytes = (uint32_t)p.n_pad * 128u;        // one 64-column chunk of a zh tile
    const uint32_t tile_tx = 2u * chunk_bytes;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapZ1); tma_prefetch_desc(&mapZ2); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kFStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], kEpiWarps); mbar_init(&p_empty[s], 1);
            mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 1);
        }
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        mbar_fence_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: S buffers at 0 / 128, O buffers at 256 / 384

    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0, uiter = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit / p.n_blocks, ib = unit % p.n_blocks;
            const bool side1 = (pass == 0 ? p.pass_side[0] : p.pass_side[1]) != 0;
            const CUtensorMap* mQ = side1 ? &mapZ2 : &mapZ1;
            const CUtensorMap* mK = side1 ? &mapZ1 : &mapZ2;
            mbar_wait(q_empty, (uiter & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(q_full, tile_tx);
                tma_load_2d(smem + kFOffQ, mQ, q_full, ib * FB, 0);
                tma_load_2d(smem + kFOffQ + chunk_bytes, mQ, q_full, ib * FB + 64, 0);
            }
            __syncwarp();
            for (int s = 0; s < p.n_blocks; ++s) {
                int jb = ib + 1 + s;
                if (jb >= p.n_blocks) jb -= p.n_blocks;
                mbar_wait(&k_empty[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* dst = smem + kFOffK + stage * kFTileBytes;
                    mbar_expect_tx(&k_full[stage], tile_tx);
                    tma_load_2d(dst, mK, &k_full[stage], jb * FB, 0);
                    tma_load_2d(dst + chunk_bytes, mK, &k_full[stage], jb * FB + 64, 0);
                }
                __syncwarp();
                if (++stage == kFStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // descriptors: high word = SBO (1024 B between 8-row groups) | version 1 | 128-byte swizzle; low word = address >> 4 | LBO << 16
        const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lo_mn = ((chunk_bytes >> 4) & 0x3FFF) << 16;      // MN-major: LBO = distance between the two 64-column chunks
        const uint32_t lo_k = 1u << 16;                                  // K-major: LBO unused with 128-byte swizzle
        const uint32_t idesc1 = make_idesc_f16(FB, FB, 1, 1, 0);         // S (128 x 128) = Q^T (MN-major) x K (MN-major), fp16
        const uint32_t idesc2 = make_idesc_f16(FB, p.n_pad, 0, 0, 0);    // O (128 x n_pad) += P (K-major) x K^T (K-major)
        const int k1 = p.n_pad / 16;
        int stage = 0; uint32_t phase = 0, uiter = 0, sc = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const uint32_t ob = uiter & 1;
            const uint32_t o_tmem = tmem_base + 256 + ob * 128;
            mbar_wait(q_full, uiter & 1);
            mbar_wait(&o_empty[ob], ((uiter >> 1) & 1) ^ 1);
            tc_fence_after();
            int prev_stage = 0; uint32_t prev_sc = 0;
            for (int s = 0; s <= p.n_blocks; ++s) {
                if (s < p.n_blocks) {
                    const uint32_t b = sc & 1;
                    mbar_wait(&k_full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        uint32_t da = lo_mn | ((q_s & 0x3FFFF) >> 4), db = lo_mn | (((k_s + stage * kFTileBytes) & 0x3FFFF) >> 4);
                        const uint32_t s_tmem = tmem_base + b * 128;
                        for (int ks = 0; ks < k1; ++ks) {
                            umma_bf16_ss(s_tmem, ((uint64_t)d_hi << 32) | da, ((uint64_t)d_hi << 32) | db, idesc1, ks > 0 ? 1u : 0u);
                            da += 2048 >> 4; db += 2048 >> 4;                    // 16 samples = 16 rows of 128 bytes
                        }
                        umma_commit(&s_full[b]);
                    }
                    __syncwarp();
                }
                if (s > 0) {
                    // MMA2 of the previous step (its P tile is written while MMA1 of this step runs)
                    const uint32_t b = prev_sc & 1;
                    mbar_wait(&p_full[b], (prev_sc >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t pa = p_s + b * kFTileBytes, kb = k_s + prev_stage * kFTileBytes;
#pragma unroll
                        for (int ks = 0; ks < FB / 16; ++ks) {
                            const uint32_t da = lo_k | (((pa + (ks >> 2) * 16384u + (ks & 3) * 32u) & 0x3FFFF) >> 4);
                            const uint32_t db = lo_k | (((kb + (ks >> 2) * chunk_bytes + (ks & 3) * 32u) & 0x3FFFF) >> 4);
                            umma_bf16_ss(o_tmem, ((uint64_t)d_hi << 32) | da, ((uint64_t)d_hi << 32) | db, idesc2, (s > 1 || ks > 0) ? 1u : 0u);
                        }
                        umma_commit(&p_empty[b]);
                        if (s < p.n_blocks) umma_commit(&k_empty[prev_stage]);     // the last tile stays for the final epilogue
                        else umma_commit(&o_full[ob]);
                    }
                    __syncwarp();
                }
                if (s < p.n_blocks) {
                    prev_stage = stage; prev_sc = sc;
                    ++sc;
                    if (++stage == kFStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warps =================
        const int q = warp & 3;               // TMEM lane quarter of this warp
        const int h = (warp - 4) >> 2;        // column half (64 columns) this warp handles
        const int il = q * 32 + lane;         // row of the block owned by this thread
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t sw = (uint32_t)(il & 7);
        float* red = reinterpret_cast<float*>(smem + kFOffP);       // [2][128][2], aliases P buffer 0 (free once o_full fired)
        uint32_t uiter = 0, sc = 0;
        int stage = 0;
        const float invN = 1.0f / (float)p.N;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit / p.n_blocks, ib = unit % p.n_blocks;
            const int side = pass == 0 ? p.pass_side[0] : p.pass_side[1];
            const uint32_t ob = uiter & 1;
            float l2 = 0.f, l1 = 0.f;
            int last_stage = 0;
            for (int s = 0; s < p.n_blocks; ++s) {
                const uint32_t b = sc & 1, par = (sc >> 1) & 1;
                mbar_wait(&s_full[b], par);
                tc_fence_after();
                uint32_t ra[32], rb[32];
                const uint32_t t_addr = tmem_base + b * 128 + h * 64 + lane_sel;
                tmem_ld_32x32(t_addr, ra);
                tmem_ld_32x32(t_addr + 32, rb);
                tmem_ld_wait();
                if (s == p.n_blocks - 1 && (q >> 1) == h) {
                    // diagonal block (the walk ends at J = I): the diagonal is handled in fp32 outside the tensor cores
                    const int dt = il - 64 * h;
#pragma unroll
                    for (int t = 0; t < 32; ++t) { ra[t] = (t == dt) ? 0u : ra[t]; rb[t] = (t + 32 == dt) ? 0u : rb[t]; }
                }
                uint32_t pk[32];
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float a0 = __uint_as_float(ra[2 * t]), a1 = __uint_as_float(ra[2 * t + 1]);
                    const float b0 = __uint_as_float(rb[2 * t]), b1 = __uint_as_float(rb[2 * t + 1]);
                    l2 = fmaf(a0, a0, l2); l2 = fmaf(a1, a1, l2); l2 = fmaf(b0, b0, l2); l2 = fmaf(b1, b1, l2);
                    if (p.hsic) l1 += (a0 + a1) + (b0 + b1);
                    pk[t] = pack_f16x2(a0, a1); pk[16 + t] = pack_f16x2(b0, b1);
                }
                mbar_wait(&p_empty[b], par ^ 1);           // MMA2 of two steps ago has finished reading this P buffer
                const uint32_t prow = p_s + b * kFTileBytes + (uint32_t)h * 16384u + (uint32_t)il * 128u;
#pragma unroll
                for (int k = 0; k < 8; ++k) sts128(prow + (((uint32_t)k ^ sw) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
                last_stage = stage;
                ++sc;
                if (++stage == kFStages) stage = 0;
            }
            if (pass == 0) {
                l2 = warp_sum(l2);
                if (p.hsic) l1 = warp_sum(l1);
                if (lane == 0) {
                    atomicAdd(p.loss_acc + 0, (double)l2 * (double)invN * (double)invN);
                    if (p.hsic) atomicAdd(p.loss_acc + 1, (double)l1 * (double)invN);
                }
            }
            // ---- final epilogue: O (dimension il, sample n) -> batch-norm backward -> dz[n][i]
            mbar_wait(&o_full[ob], (uiter >> 1) & 1);
            tc_fence_after();
            const int i = ib * FB + il;
            const bool row_ok = i < p.D;
            const int ii = row_ok ? i : 0;
            const float cd = p.stats[S_CDIAG * p.D + ii];
            const float r_s = p.stats[(side == 0 ? S_R1 : S_R2) * p.D + ii];
            const float hs = 2.0f * p.lambda * invN * invN;            // O is the sum over S = N c
            const float hsn = 2.0f * p.lambda * invN;
            const float gd = 2.0f * p.alpha * (cd - 1.0f) * invN;      // G_ii / N
            const float* rso = side == 0 ? p.rs2 : p.rs1;
            const uint32_t x_s = k_s + last_stage * kFTileBytes;       // zh_other[:, I]: the tile of the last step
            const int n0 = h * 64;
            uint32_t ra[32], rb[32];
            float sg = 0.f, sgz = 0.f;
            const uint32_t o_addr = tmem_base + 256 + ob * 128 + lane_sel;
            if (n0 < p.n_pad) tmem_ld_32x32(o_addr + n0, ra);
            if (n0 + 32 < p.n_pad) tmem_ld_32x32(o_addr + n0 + 32, rb);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int n = n0 + t;
                if (n < p.N) {
                    const float zo = lds_half(tile_elem_addr(x_s, chunk_bytes, n, il)), zs = lds_half(tile_elem_addr(q_s, chunk_bytes, n, il));
                    float g = fmaf(hs, __uint_as_float(ra[t]), gd * zo);
                    if (p.hsic) g = fmaf(hsn, __ldg(rso + n) - zo, g);
                    ra[t] = __float_as_uint(g);
                    sg += g; sgz = fmaf(g, zs, sgz);
                }
            }
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int n = n0 + 32 + t;
                if (n < p.N) {
                    const float zo = lds_half(tile_elem_addr(x_s, chunk_bytes, n, il)), zs = lds_half(tile_elem_addr(q_s, chunk_bytes, n, il));
                    float g = fmaf(hs, __uint_as_float(rb[t]), gd * zo);
                    if (p.hsic) g = fmaf(hsn, __ldg(rso + n) - zo, g);
                    rb[t] = __float_as_uint(g);
                    sg += g; sgz = fmaf(g, zs, sgz);
                }
            }
            // the two column halves of a row live in different warps: exchange the partial sums through shared memory
            red[(h * FB + il) * 2] = sg; red[(h * FB + il) * 2 + 1] = sgz;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            sg += red[((h ^ 1) * FB + il) * 2]; sgz += red[((h ^ 1) * FB + il) * 2 + 1];
            const float mg = sg * invN, bb = sgz * invN, rg = r_s * p.grad_scale;
            if (row_ok) {
                void* dzv = side == 0 ? p.dz1 : p.dz2;
                if (p.io_dtype == ABT_DTYPE_BF16) {
                    fused_store_chunk<__nv_bfloat16>(ra, static_cast<__nv_bfloat16*>(dzv), p.D, i, n0, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                    fused_store_chunk<__nv_bfloat16>(rb, static_cast<__nv_bfloat16*>(dzv), p.D, i, n0 + 32, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                } else if (p.io_dtype == ABT_DTYPE_F16) {
                    fused_store_chunk<__half>(ra, static_cast<__half*>(dzv), p.D, i, n0, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                    fused_store_chunk<__half>(rb, static_cast<__half*>(dzv), p.D, i, n0 + 32, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                } else {
                    fused_store_chunk<float>(ra, static_cast<float*>(dzv), p.D, i, n0, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                    fused_store_chunk<float>(rb, static_cast<float*>(dzv), p.D, i, n0 + 32, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                }
            }
            // release O, Q and the last ring stage (all epilogue warps are done with them after this barrier)
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 128) {
                mbar_arrive(&o_empty[ob]);
                mbar_arrive(q_empty);
                mbar_arrive(&k_empty[last_stage]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    // the last CTA to finish publishes the loss (every pass-0 unit has added its off-diagonal part by then)
    if (threadIdx.x == 0 && p.loss_out != nullptr) {
        __threadfence();
        if (atomicAdd(p.done_counter, 1u) == gridDim.x - 1) {
            __threadfence();
            const double a0 = *reinterpret_cast<volatile double*>(p.loss_acc), a1 = *reinterpret_cast<volatile double*>(p.loss_acc + 1);
            double a2 = 0.0;
            for (int k = 0; k < p.n_parts; ++k) a2 += (double)__ldcg(p.ondiag_part + k);
            double off = a0;
            if (p.hsic) off = a0 + 2.0 * a1 + (double)p.D * (double)(p.D - 1);
            *p.loss_out = (float)((double)p.alpha * a2 + (double)p.lambda * off);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Version 2: both A operands in tensor memory.  With shared-memory operands an M = N = 128 MMA reads 8 KB per 64 clocks --
// exactly the 128 B / clk shared-memory port -- so MMA1 + MMA2 + the P stores + the TMA fills (192 KB per step) made the step
// shared-memory bound (measured: 2270 clk per step against 1024 clk of MMA issue).  Here
//   * Q = zh_a[:, I]^T (constant over the walk) is transposed into TMEM once per unit (128 lanes x n_pad/2 packed columns),
//   * P overwrites the first 64 columns of the S buffer it was computed from (tcgen05.st, fp16 pairs),
// and only the zh_b tile (operand B of both MMAs) and its TMA fill touch shared memory: 96 KB per step.
// Epilogue warps 4-7 take the even steps, warps 8-11 the odd ones: a warp owns all 128 columns of its 32 rows, so P can
// alias S without cross-warp hazards and each group has two step times for its S -> P conversion.
// TMEM columns: S/P buffers at 0 / 128, O at 256, Q at 384.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTStages = 5;
constexpr int kTThreads = 320;      // warp 0 TMA producer (+ TMEM allocation), warp 1 MMA issuer, warps 2-9 epilogue (3 warps on a scheduler: 168 registers)
constexpr int kTOffQ = 0;
constexpr int kTOffK = kFTileBytes;
constexpr int kTOffBar = kTOffK + kTStages * kFTileBytes;
constexpr int kTOffRed = kTOffBar + 256;
constexpr int kTSmemBytes = kTOffRed + 2048 + 1024 /*align slack*/;
static_assert(kTSmemBytes <= 227 * 1024, "fused TS kernel: shared memory budget");

// 32 fp32 accumulator columns of one row -> loss partial sums + 16 packed fp16 pairs
__device__ __forceinline__ void fused_pack_chunk(uint32_t (&r)[32], int diag_t, bool hsic, float& l2, float& l1, uint32_t (&pk)[16]) {
    if (diag_t >= 0 && diag_t < 32) {        // the diagonal element of this row lies in this chunk (diagonal block only)
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = (t == diag_t) ? 0u : r[t];
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const float a0 = __uint_as_float(r[2 * t]), a1 = __uint_as_float(r[2 * t + 1]);
        l2 = fmaf(a0, a0, l2); l2 = fmaf(a1, a1, l2);
        if (hsic) l1 += a0 + a1;
        pk[t] = pack_f16x2(a0, a1);
    }
}

__global__ void __launch_bounds__(kTThreads, 1)
bt_fused_ts_kernel(const __grid_constant__ CUtensorMap mapZ1, const __grid_constant__ CUtensorMap mapZ2, const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* k_full = reinterpret_cast<uint64_t*>(smem + kTOffBar);
    uint64_t* k_empty = k_full + kTStages;
    uint64_t* s_full = k_empty + kTStages;      // [2] MMA1 done -> epilogue group
    uint64_t* p_full = s_full + 2;              // [2] P stored over S -> MMA issuer
    uint64_t* o_full = p_full + 2;              // all MMAs of the unit done
    uint64_t* o_empty = o_full + 1;             // final epilogue done: O, Q (TMEM and smem) and the last ring stage are free
    uint64_t* q_full = o_empty + 1;             // Q tile landed in shared memory
    uint64_t* qt_full = q_full + 1;             // Q transposed into TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qt_full + 1);
    float* red = reinterpret_cast<float*>(smem + kTOffRed);          // [2][128][2]
    const uint32_t q_s = smem_u32(smem + kTOffQ), k_s = smem_u32(smem + kTOffK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.n_blocks * p.pass_count;
    const uint32_t chunk_bytes = (uint32_t)p.n_pad * 128u;
    const uint32_t tile_tx = 2u * chunk_bytes;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapZ1); tma_prefetch_desc(&mapZ2); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); }
        mbar_init(o_full, 1); mbar_init(o_empty, 1); mbar_init(q_full, 1); mbar_init(qt_full, kEpiWarps);
        mbar_fence_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t o_tmem = tmem_base + 256, qt_tmem = tmem_base + 384;

    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0, uiter = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit / p.n_blocks, ib = unit % p.n_blocks;
            const bool side1 = (pass == 0 ? p.pass_side[0] : p.pass_side[1]) != 0;
            const CUtensorMap* mQ = side1 ? &mapZ2 : &mapZ1;
            const CUtensorMap* mK = side1 ? &mapZ1 : &mapZ2;
            mbar_wait(o_empty, (uiter & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(q_full, tile_tx);
                tma_load_2d(smem + kTOffQ, mQ, q_full, ib * FB, 0);
                tma_load_2d(smem + kTOffQ + chunk_bytes, mQ, q_full, ib * FB + 64, 0);
            }
            __syncwarp();
            for (int s = 0; s < p.n_blocks; ++s) {
                int jb = ib + 1 + s;
                if (jb >= p.n_blocks) jb -= p.n_blocks;
                mbar_wait(&k_empty[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* dst = smem + kTOffK + stage * kFTileBytes;
                    mbar_expect_tx(&k_full[stage], tile_tx);
                    tma_load_2d(dst, mK, &k_full[stage], jb * FB, 0);
                    tma_load_2d(dst + chunk_bytes, mK, &k_full[stage], jb * FB + 64, 0);
                }
                __syncwarp();
                if (++stage == kTStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lo_mn = ((chunk_bytes >> 4) & 0x3FFF) << 16;
        const uint32_t lo_k = 1u << 16;
        const uint32_t idesc1 = make_idesc_f16(FB, FB, 0, 1, 0);         // S = Q^T (TMEM, K-major) x K (smem, MN-major)
        const uint32_t idesc2 = make_idesc_f16(FB, p.n_pad, 0, 0, 0);    // O += P (TMEM) x K^T (smem, K-major)
        const int k1 = p.n_pad / 16;
        int stage = 0; uint32_t phase = 0, uiter = 0, sc = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            mbar_wait(qt_full, uiter & 1);
            tc_fence_after();
            int prev_stage = 0; uint32_t prev_sc = 0;
            for (int s = 0; s <= p.n_blocks; ++s) {
                if (s < p.n_blocks) {
                    const uint32_t b = sc & 1;
                    mbar_wait(&k_full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        uint32_t db = lo_mn | (((k_s + stage * kFTileBytes) & 0x3FFFF) >> 4);
                        const uint32_t s_tmem = tmem_base + b * 128;
                        for (int ks = 0; ks < k1; ++ks) {
                            umma_f16_ts(s_tmem, qt_tmem + ks * 8, ((uint64_t)d_hi << 32) | db, idesc1, ks > 0 ? 1u : 0u);
                            db += 2048 >> 4;
                        }
                        umma_commit(&s_full[b]);
                    }
                    __syncwarp();
                }
                if (s > 0) {
                    const uint32_t b = prev_sc & 1;
                    mbar_wait(&p_full[b], (prev_sc >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t kb = k_s + prev_stage * kFTileBytes, p_tmem = tmem_base + b * 128;
#pragma unroll
                        for (int ks = 0; ks < FB / 16; ++ks) {
                            const uint32_t db = lo_k | (((kb + (ks >> 2) * chunk_bytes + (ks & 3) * 32u) & 0x3FFFF) >> 4);
                            umma_f16_ts(o_tmem, p_tmem + ks * 8, ((uint64_t)d_hi << 32) | db, idesc2, (s > 1 || ks > 0) ? 1u : 0u);
                        }
                        if (s < p.n_blocks) umma_commit(&k_empty[prev_stage]);     // the last tile stays for the final epilogue
                        else umma_commit(o_full);
                    }
                    __syncwarp();
                }
                if (s < p.n_blocks) {
                    prev_stage = stage; prev_sc = sc;
                    ++sc;
                    if (++stage == kTStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 2) {
        // ================= epilogue warps =================
        const int q = warp & 3;               // TMEM lane quarter of this warp (warps 2-5 and 6-9 each cover the four quarters)
        const int h = (warp - 2) >> 2;        // group: steps with (step & 1) == h; column half in the unit prologue / epilogue
        const int il = q * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        uint32_t uiter = 0, sc = 0;
        const float invN = 1.0f / (float)p.N;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit / p.n_blocks, ib = unit % p.n_blocks;
            const int side = pass == 0 ? p.pass_side[0] : p.pass_side[1];
            // ---- unit prologue: Q^T into TMEM (this warp: samples 64 h .. 64 h + 63 of its 32 dimensions)
            mbar_wait(q_full, uiter & 1);
            if (h * 64 < p.n_pad) {
                uint32_t qa[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = h * 64 + 2 * c;
                    unsigned short lo = 0, hi = 0;
                    if (n < p.n_pad) {
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(lo) : "r"(tile_elem_addr(q_s, chunk_bytes, n, il)));
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hi) : "r"(tile_elem_addr(q_s, chunk_bytes, n + 1, il)));
                    }
                    qa[c] = (uint32_t)lo | ((uint32_t)hi << 16);
                }
                tmem_st_32x32_x32(qt_tmem + h * 32 + lane_sel, qa);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qt_full);
            // ---- the walk: this group converts S -> P for every other step
            float l2 = 0.f, l1 = 0.f;
            for (int s = 0; s < p.n_blocks; ++s, ++sc) {
                if ((int)(sc & 1) != h) continue;
                const uint32_t b = sc & 1, par = (sc >> 1) & 1;
                mbar_wait(&s_full[b], par);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + b * 128 + lane_sel;
                const int dt = (s == p.n_blocks - 1) ? il : -1;           // diagonal block: column il of row il
                uint32_t ra[32], rb[32], pk[16];
                tmem_ld_32x32(t_addr, ra);
                tmem_ld_32x32(t_addr + 32, rb);
                tmem_ld_wait();
                fused_pack_chunk(ra, dt, p.hsic != 0, l2, l1, pk);
                tmem_st_32x32_x16(t_addr, pk);
                tmem_ld_32x32(t_addr + 64, ra);
                fused_pack_chunk(rb, dt - 32, p.hsic != 0, l2, l1, pk);
                tmem_st_32x32_x16(t_addr + 16, pk);
                tmem_ld_wait();
                tmem_ld_32x32(t_addr + 96, rb);
                fused_pack_chunk(ra, dt - 64, p.hsic != 0, l2, l1, pk);
                tmem_st_32x32_x16(t_addr + 32, pk);
                tmem_ld_wait();
                fused_pack_chunk(rb, dt - 96, p.hsic != 0, l2, l1, pk);
                tmem_st_32x32_x16(t_addr + 48, pk);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
            }
            if (pass == 0) {
                l2 = warp_sum(l2);
                if (p.hsic) l1 = warp_sum(l1);
                if (lane == 0) {
                    atomicAdd(p.loss_acc + 0, (double)l2 * (double)invN * (double)invN);
                    if (p.hsic) atomicAdd(p.loss_acc + 1, (double)l1 * (double)invN);
                }
            }
            // ---- final epilogue: O (dimension il, sample n) -> batch-norm backward -> dz[n][i]; this warp: samples 64 h .. 64 h + 63
            const int last_stage = (int)((sc - 1) % kTStages);
            mbar_wait(o_full, uiter & 1);
            tc_fence_after();
            const int i = ib * FB + il;
            const bool row_ok = i < p.D;
            const int ii = row_ok ? i : 0;
            const float cd = p.stats[S_CDIAG * p.D + ii];
            const float r_s = p.stats[(side == 0 ? S_R1 : S_R2) * p.D + ii];
            const float hs = 2.0f * p.lambda * invN * invN;
            const float hsn = 2.0f * p.lambda * invN;
            const float gd = 2.0f * p.alpha * (cd - 1.0f) * invN;
            const float* rso = side == 0 ? p.rs2 : p.rs1;
            const uint32_t x_s = k_s + last_stage * kFTileBytes;
            const int n0 = h * 64;
            uint32_t ra[32], rb[32];
            float sg = 0.f, sgz = 0.f;
            const uint32_t o_addr = o_tmem + lane_sel;
            if (n0 < p.n_pad) tmem_ld_32x32(o_addr + n0, ra);
            if (n0 + 32 < p.n_pad) tmem_ld_32x32(o_addr + n0 + 32, rb);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int n = n0 + t;
                if (n < p.N) {
                    const float zo = lds_half(tile_elem_addr(x_s, chunk_bytes, n, il)), zs = lds_half(tile_elem_addr(q_s, chunk_bytes, n, il));
                    float g = fmaf(hs, __uint_as_float(ra[t]), gd * zo);
                    if (p.hsic) g = fmaf(hsn, __ldg(rso + n) - zo, g);
                    ra[t] = __float_as_uint(g);
                    sg += g; sgz = fmaf(g, zs, sgz);
                }
            }
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int n = n0 + 32 + t;
                if (n < p.N) {
                    const float zo = lds_half(tile_elem_addr(x_s, chunk_bytes, n, il)), zs = lds_half(tile_elem_addr(q_s, chunk_bytes, n, il));
                    float g = fmaf(hs, __uint_as_float(rb[t]), gd * zo);
                    if (p.hsic) g = fmaf(hsn, __ldg(rso + n) - zo, g);
                    rb[t] = __float_as_uint(g);
                    sg += g; sgz = fmaf(g, zs, sgz);
                }
            }
            red[(h * FB + il) * 2] = sg; red[(h * FB + il) * 2 + 1] = sgz;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            sg += red[((h ^ 1) * FB + il) * 2]; sgz += red[((h ^ 1) * FB + il) * 2 + 1];
            const float mg = sg * invN, bb = sgz * invN, rg = r_s * p.grad_scale;
            if (row_ok) {
                void* dzv = side == 0 ? p.dz1 : p.dz2;
                if (p.io_dtype == ABT_DTYPE_BF16) {
                    fused_store_chunk<__nv_bfloat16>(ra, static_cast<__nv_bfloat16*>(dzv), p.D, i, n0, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                    fused_store_chunk<__nv_bfloat16>(rb, static_cast<__nv_bfloat16*>(dzv), p.D, i, n0 + 32, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                } else if (p.io_dtype == ABT_DTYPE_F16) {
                    fused_store_chunk<__half>(ra, static_cast<__half*>(dzv), p.D, i, n0, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                    fused_store_chunk<__half>(rb, static_cast<__half*>(dzv), p.D, i, n0 + 32, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                } else {
                    fused_store_chunk<float>(ra, static_cast<float*>(dzv), p.D, i, n0, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                    fused_store_chunk<float>(rb, static_cast<float*>(dzv), p.D, i, n0 + 32, p.N, q_s, chunk_bytes, il, mg, bb, rg);
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                mbar_arrive(o_empty);
                mbar_arrive(&k_empty[last_stage]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0 && p.loss_out != nullptr) {
        __threadfence();
        if (atomicAdd(p.done_counter, 1u) == gridDim.x - 1) {
            __threadfence();
            const double a0 = *reinterpret_cast<volatile double*>(p.loss_acc), a1 = *reinterpret_cast<volatile double*>(p.loss_acc + 1);
            double a2 = 0.0;
            for (int k = 0; k < p.n_parts; ++k) a2 += (double)__ldcg(p.ondiag_part + k);
            double off = a0;
            if (p.hsic) off = a0 + 2.0 * a1 + (double)p.D * (double)(p.D - 1);
            *p.loss_out = (float)((double)p.alpha * a2 + (double)p.lambda * off);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Version 3: 64-column steps, four S buffers, MMA issue three steps ahead.
// Measured on version 2 (ncu, D = 8192): the TMA ring is always full and shared memory is idle, but the tensor pipe is busy
// only 53 % of the time -- a step is a latency chain  MMA1 -> commit -> epilogue wake-up -> S -> P conversion -> arrive ->
// issuer wake-up -> MMA2,  and with two S buffers only two such chains overlap (1850 clk per step against 1024 clk of MMAs).
// Here the same 256 TMEM columns hold FOUR 64-column S buffers and the issuer runs MMA1 three steps ahead of MMA2, so four
// chains overlap; the conversion itself is cut to ~100 instructions per thread (packed fp32x2 FMAs for the loss, templates
// instead of run-time branches), and the unit epilogue reads its operands with precomputed swizzle offsets.
// ------------------------------------------------------------------------------------------------------------------
constexpr int JW = 64;                        // columns of C per step
constexpr int kXStages = 11;                  // ring of 16 KiB zh_b tiles (64 columns x up to 128 samples); FusedParams::n_stages <= this are used
constexpr int kXTileBytes = JW * 128 * 2;
constexpr int kXLook = 3;                     // MMA1 runs this many steps ahead of MMA2
constexpr int kXOffQ = 0;
constexpr int kXOffK = kFTileBytes;
constexpr int kXOffBar = kXOffK + kXStages * kXTileBytes;
constexpr int kXOffRed = kXOffBar + 512;
constexpr int kXSmemBytes = kXOffRed + 2048 + 1024 /*align slack*/;
static_assert(kXSmemBytes <= 227 * 1024, "fused kernel v3: shared memory budget");

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fma_f32x2(uint64_t& acc, uint64_t x) { asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(x)); }
__device__ __forceinline__ void add_f32x2(uint64_t& acc, uint64_t x) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(x)); }
__device__ __forceinline__ float sum_f32x2(uint64_t v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (uint32_t)v;
}
__device__ __forceinline__ float half_bits_to_float(uint32_t bits) { return __half2float(__ushort_as_half((unsigned short)bits)); }

template <typename T, bool HSIC>
__global__ void __launch_bounds__(kTThreads, 1)
bt_fused3_kernel(const __grid_constant__ CUtensorMap mapZ1, const __grid_constant__ CUtensorMap mapZ2, const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* k_full = reinterpret_cast<uint64_t*>(smem + kXOffBar);
    uint64_t* k_empty = k_full + kXStages;
    const uint32_t ns = (uint32_t)p.n_stages;
    uint64_t* s_full = k_empty + kXStages;      // [4] MMA1 done -> epilogue group
    uint64_t* p_full = s_full + 4;              // [4] P stored over S -> MMA issuer
    uint64_t* o_full = p_full + 4;              // all MMAs of the unit done
    uint64_t* o_empty = o_full + 1;             // unit epilogue done: Q (TMEM and smem) may be overwritten
    uint64_t* q_full = o_empty + 1;             // Q tile landed in shared memory
    uint64_t* qt_full = q_full + 1;             // Q transposed into TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qt_full + 1);
    float* red = reinterpret_cast<float*>(smem + kXOffRed);          // [2][128][2]
    const uint32_t q_s = smem_u32(smem + kXOffQ), k_s = smem_u32(smem + kXOffK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.n_blocks * p.pass_count;
    const int nsteps = 2 * p.n_blocks;                               // 64-column blocks (one all-zero block when D % 128 == 64)
    const uint32_t chunk_bytes = (uint32_t)p.n_pad * 128u;          // one 64-column tile: n_pad samples x 128 bytes

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapZ1); tma_prefetch_desc(&mapZ2); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kXStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); }
        mbar_init(o_full, 1); mbar_init(o_empty, 1); mbar_init(q_full, 1); mbar_init(qt_full, kEpiWarps);
        mbar_fence_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t o_tmem = tmem_base + 256, qt_tmem = tmem_base + 384;       // S / P buffers: columns 64 b, b < 4
    // launched with programmatic stream serialisation: everything above overlapped the tail of the statistics kernel; its outputs
    // (tile images, statistics, cleared accumulators) are read only after this point
    griddep_wait();

    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0, uiter = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit / p.n_blocks, ib = unit % p.n_blocks;
            const bool side1 = (pass == 0 ? p.pass_side[0] : p.pass_side[1]) != 0;
            const CUtensorMap* mQ = side1 ? &mapZ2 : &mapZ1;
            const CUtensorMap* mK = side1 ? &mapZ1 : &mapZ2;
            const uint8_t* imgQ = reinterpret_cast<const uint8_t*>(side1 ? p.zimg2 : p.zimg1);
            const uint8_t* imgK = reinterpret_cast<const uint8_t*>(side1 ? p.zimg1 : p.zimg2);
            mbar_wait(o_empty, (uiter & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(q_full, 2u * chunk_bytes);
                if (imgQ != nullptr) {
                    bulk_load(smem + kXOffQ, imgQ + (size_t)(2 * ib) * chunk_bytes, 2u * chunk_bytes, q_full);     // tiles 2 ib and 2 ib + 1 are adjacent
                } else {
                    tma_load_2d(smem + kXOffQ, mQ, q_full, ib * FB, 0);
                    tma_load_2d(smem + kXOffQ + chunk_bytes, mQ, q_full, ib * FB + 64, 0);
                }
            }
            __syncwarp();
            int jb = 2 * ib + 2;                                     // the walk ends with blocks 2 ib, 2 ib + 1 (= the columns of I)
            for (int s = 0; s < nsteps; ++s, ++jb) {
                if (jb >= nsteps) jb -= nsteps;
                mbar_wait(&k_empty[stage], phase ^ 1);
                if (elect_one()) {
                    if (p.debug & 256) mbar_arrive(&k_full[stage]);            // timing experiment: handshakes only, no data
                    else {
                        mbar_expect_tx(&k_full[stage], chunk_bytes);
                        if (imgK != nullptr) bulk_load(smem + kXOffK + stage * kXTileBytes, imgK + (size_t)jb * chunk_bytes, chunk_bytes, &k_full[stage]);
                        else tma_load_2d(smem + kXOffK + stage * kXTileBytes, mK, &k_full[stage], jb * JW, 0);
                    }
                }
                __syncwarp();
                if (++stage == (int)ns) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lo_mn = ((chunk_bytes >> 4) & 0x3FFF) << 16;
        const uint32_t lo_k = 1u << 16;
        const uint32_t idesc1 = make_idesc_f16(FB, JW, 0, 1, 0);         // S (128 x 64) = Q^T (TMEM) x K tile (smem, MN-major)
        const uint32_t idesc2 = make_idesc_f16(FB, p.n_pad, 0, 0, 0);    // O (128 x n_pad) += P (TMEM) x K tile^T (smem, K-major)
        const int k1 = p.n_pad / 16;
        uint32_t c1 = 0, c2 = 0, uiter = 0;                               // steps whose MMA1 / MMA2 have been issued (over all units)
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            mbar_wait(qt_full, uiter & 1);
            tc_fence_after();
            for (int s = 0; s < nsteps + kXLook; ++s) {
                if (s < nsteps) {
                    const uint32_t st = c1 % ns, b = c1 & 3;
                    mbar_wait(&k_full[st], (c1 / ns) & 1);
                    if (!(p.debug & 512)) tc_fence_after();
                    if (elect_one()) {
                        uint32_t db = lo_mn | (((k_s + st * kXTileBytes) & 0x3FFFF) >> 4);
                        const uint32_t s_tmem = tmem_base + b * JW;
                        for (int ks = 0; ks < ((p.debug & 8) ? 1 : k1); ++ks) {
                            umma_f16_ts(s_tmem, qt_tmem + ks * 8, ((uint64_t)d_hi << 32) | db, idesc1, ks > 0 ? 1u : 0u);
                            db += 2048 >> 4;
                        }
                        if (p.debug & 1024) mbar_arrive(&s_full[b]); else umma_commit(&s_full[b]);
                    }
                    __syncwarp();
                    ++c1;
                }
                if (s >= kXLook) {
                    const int t = s - kXLook;                              // step whose P tile is consumed now
                    const uint32_t st = c2 % ns, b = c2 & 3;
                    if (!(p.debug & 1)) mbar_wait(&p_full[b], (c2 >> 2) & 1);
                    if (!(p.debug & 512)) tc_fence_after();
                    if (elect_one()) {
                        const uint32_t kb = k_s + st * kXTileBytes, p_tmem = tmem_base + b * JW;
#pragma unroll
                        for (int ks = 0; ks < JW / 16; ++ks) {
                            if ((p.debug & 4) && ks > 0) break;
                            const uint32_t db = lo_k | (((kb + ks * 32u) & 0x3FFFF) >> 4);
                            umma_f16_ts(o_tmem, p_tmem + ks * 8, ((uint64_t)d_hi << 32) | db, idesc2, (t > 0 || ks > 0) ? 1u : 0u);
                        }
                        if (t < nsteps - 2) { if (p.debug & 2048) mbar_arrive(&k_empty[st]); else umma_commit(&k_empty[st]); }       // the last two tiles (= zh_b[:, I]) stay for the unit epilogue
                        else if (t == nsteps - 1) umma_commit(o_full);
                    }
                    __syncwarp();
                    ++c2;
                }
            }
        }
    } else {
        // ================= epilogue warps (2-9) =================
        const int q = warp & 3;               // TMEM lane quarter of this warp
        const int h = (warp - 2) >> 2;        // group: steps with (step & 1) == h; sample half in the unit prologue / epilogue
        const int il = q * 32 + lane;         // row of the block (dimension) owned by this thread
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        // swizzled byte offset of column (il & 63) inside a 128-byte tile row whose index is k modulo 8 (k is a compile-time constant at
        // every use; the eight values are not kept live across the walk)
        const uint32_t xc = (uint32_t)(il & 63) >> 3, xw = (uint32_t)(il & 7) * 2u;
        auto xoff = [&](int k) -> uint32_t { return ((xc ^ (uint32_t)k) << 4) + xw; };
        const uint32_t q_col = q_s + (uint32_t)(il >> 6) * chunk_bytes + (uint32_t)h * 64u * 128u;      // Q tile, this thread's column, sample 64 h
        uint32_t uiter = 0, c = 0;
        const float invN = 1.0f / (float)p.N;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit / p.n_blocks, ib = unit % p.n_blocks;
            const int side = pass == 0 ? p.pass_side[0] : p.pass_side[1];
            // ---- unit prologue: Q^T into TMEM (this warp: samples 64 h .. 64 h + 63 of its 32 dimensions, packed in pairs)
            mbar_wait(q_full, uiter & 1);
            if (h * 64 < p.n_pad && !(p.debug & 128)) {
                uint32_t qa[32];
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const uint32_t lo = lds_u16(q_col + (2 * cc) * 128 + xoff((2 * cc) & 7)), hi = lds_u16(q_col + (2 * cc + 1) * 128 + xoff((2 * cc + 1) & 7));
                    qa[cc] = lo | (hi << 16);
                }
                tmem_st_32x32_x32(qt_tmem + h * 32 + lane_sel, qa);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qt_full);
            // ---- the walk: this group converts S -> P for every other step
            uint64_t a2[4] = {0ull, 0ull, 0ull, 0ull}, a1[2] = {0ull, 0ull};
            const bool count_loss = pass == 0;
            for (int s = 0; s < nsteps; ++s, ++c) {
                if ((int)(c & 1) != h) continue;
                const uint32_t b = c & 3;
                mbar_wait(&s_full[b], (c >> 2) & 1);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + b * JW + lane_sel;
                if (p.debug & 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[b]);
                    continue;
                }
                uint32_t ra[32], rb[32];
                tmem_ld_32x32(t_addr, ra);
                tmem_ld_32x32(t_addr + 32, rb);
                tmem_ld_wait();
                if (s >= nsteps - 2) {
                    // the two diagonal blocks: the diagonal is handled in fp32 outside the tensor cores
                    const int dt = il - (s - (nsteps - 2)) * 64;
                    if (dt >= 0 && dt < 64) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) { ra[t] = (t == dt) ? 0u : ra[t]; rb[t] = (t + 32 == dt) ? 0u : rb[t]; }
                    }
                }
                if (count_loss) {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const uint64_t xa = pack_f32x2(__uint_as_float(ra[2 * t]), __uint_as_float(ra[2 * t + 1]));
                        const uint64_t xb = pack_f32x2(__uint_as_float(rb[2 * t]), __uint_as_float(rb[2 * t + 1]));
                        fma_f32x2(a2[t & 1], xa); fma_f32x2(a2[2 + (t & 1)], xb);
                        if (HSIC) { add_f32x2(a1[0], xa); add_f32x2(a1[1], xb); }
                    }
                }
                uint32_t pk[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2(__uint_as_float(ra[2 * t]), __uint_as_float(ra[2 * t + 1]));
                tmem_st_32x32_x16(t_addr, pk);
#pragma unroll
                for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2(__uint_as_float(rb[2 * t]), __uint_as_float(rb[2 * t + 1]));
                tmem_st_32x32_x16(t_addr + 16, pk);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
            }
            if (count_loss) {
                float l2 = (sum_f32x2(a2[0]) + sum_f32x2(a2[1])) + (sum_f32x2(a2[2]) + sum_f32x2(a2[3]));
                l2 = warp_sum(l2);
                float l1 = 0.f;
                if (HSIC) l1 = warp_sum(sum_f32x2(a1[0]) + sum_f32x2(a1[1]));
                if (lane == 0) {
                    atomicAdd(p.loss_acc + 0, (double)l2 * (double)invN * (double)invN);
                    if (HSIC) atomicAdd(p.loss_acc + 1, (double)l1 * (double)invN);
                }
            }
            // ---- unit epilogue: O (dimension il, sample n) -> batch-norm backward -> dz[n][i]; this warp: samples 64 h .. 64 h + 63
            const uint32_t st_a = (c - 2) % ns, st_b = (c - 1) % ns;      // tiles of the last two steps: zh_b[:, I]
            mbar_wait(o_full, uiter & 1);
            tc_fence_after();
            const int i = ib * FB + il;
            const bool row_ok = i < p.D;
            const int ii = row_ok ? i : 0;
            const float cd = p.stats[S_CDIAG * p.D + ii];
            const float r_s = p.stats[(side == 0 ? S_R1 : S_R2) * p.D + ii];
            const float hs = 2.0f * p.lambda * invN * invN;            // O is the sum over S = N c
            const float hsn = 2.0f * p.lambda * invN;
            const float gd = 2.0f * p.alpha * (cd - 1.0f) * invN;      // G_ii / N
            const float* rso = side == 0 ? p.rs2 : p.rs1;
            const uint32_t x_col = k_s + (il < 64 ? st_a : st_b) * kXTileBytes + (uint32_t)h * 64u * 128u;
            const int n0 = h * 64;
            uint32_t ra[32], rb[32];
            float sg = 0.f, sgz = 0.f;
            const uint32_t o_addr = o_tmem + lane_sel, qz_addr = qt_tmem + h * 32 + lane_sel;
            const bool have_a = n0 < p.n_pad && !(p.debug & 64), have_b = n0 + 32 < p.n_pad && !(p.debug & 64);          // n_pad is a multiple of 32
            // zh_self[n, i] for this thread's samples comes back from the Q^T operand in TMEM (the packed pairs this warp stored in the prologue)
            if (have_a) {
                uint32_t qz[16];
                tmem_ld_32x32(o_addr + n0, ra);
                tmem_ld_32x32_x16(qz_addr, qz);
                tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const float2 zs2 = __half22float2(*reinterpret_cast<const __half2*>(&qz[t >> 1]));
                    const float zs = (t & 1) ? zs2.y : zs2.x;
                    const float zo = half_bits_to_float(lds_u16(x_col + t * 128 + xoff(t & 7)));
                    float g = fmaf(hs, __uint_as_float(ra[t]), gd * zo);
                    if (HSIC && n0 + t < p.N) g = fmaf(hsn, __ldg(rso + n0 + t) - zo, g);            // padded samples (n >= N) contribute nothing
                    ra[t] = __float_as_uint(g);
                    sg += g; sgz = fmaf(g, zs, sgz);
                }
            }
            if (have_b) {
                uint32_t qz[16];
                tmem_ld_32x32(o_addr + n0 + 32, rb);
                tmem_ld_32x32_x16(qz_addr + 16, qz);
                tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const float2 zs2 = __half22float2(*reinterpret_cast<const __half2*>(&qz[t >> 1]));
                    const float zs = (t & 1) ? zs2.y : zs2.x;
                    const float zo = half_bits_to_float(lds_u16(x_col + (32 + t) * 128 + xoff(t & 7)));
                    float g = fmaf(hs, __uint_as_float(rb[t]), gd * zo);
                    if (HSIC && n0 + 32 + t < p.N) g = fmaf(hsn, __ldg(rso + n0 + 32 + t) - zo, g);
                    rb[t] = __float_as_uint(g);
                    sg += g; sgz = fmaf(g, zs, sgz);
                }
            }
            // the two sample halves of a dimension live in different warps: exchange the partial sums through shared memory
            red[(h * FB + il) * 2] = sg; red[(h * FB + il) * 2 + 1] = sgz;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            sg += red[((h ^ 1) * FB + il) * 2]; sgz += red[((h ^ 1) * FB + il) * 2 + 1];
            const float mg = sg * invN, bb = sgz * invN, rg = r_s * p.grad_scale;
            const float mgr = mg * rg, bbr = bb * rg;                              // dz = g rg - mg rg - zs (bb rg)
            if (row_ok) {
                T* dz = static_cast<T*>(side == 0 ? p.dz1 : p.dz2) + (size_t)n0 * p.D + i;
                const int nv = p.N - n0;                                           // valid samples of this half (may be <= 0)
                const size_t ld = (size_t)p.D;
                if (have_a) {
                    uint32_t qz[16];
                    tmem_ld_32x32_x16(qz_addr, qz);
                    tmem_ld_wait();
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const float2 zs2 = __half22float2(*reinterpret_cast<const __half2*>(&qz[t >> 1]));
                        const float zs = (t & 1) ? zs2.y : zs2.x;
                        if (t < nv) store_out<T>(dz, fmaf(__uint_as_float(ra[t]), rg, -fmaf(zs, bbr, mgr)));
                        dz += ld;
                    }
                }
                if (have_b) {
                    uint32_t qz[16];
                    tmem_ld_32x32_x16(qz_addr + 16, qz);
                    tmem_ld_wait();
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const float2 zs2 = __half22float2(*reinterpret_cast<const __half2*>(&qz[t >> 1]));
                        const float zs = (t & 1) ? zs2.y : zs2.x;
                        if (32 + t < nv) store_out<T>(dz, fmaf(__uint_as_float(rb[t]), rg, -fmaf(zs, bbr, mgr)));
                        dz += ld;
                    }
                }
            }
            // release Q and the last two ring stages (every epilogue warp is done with them after this barrier)
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                mbar_arrive(o_empty);
                mbar_arrive(&k_empty[st_a]);
                mbar_arrive(&k_empty[st_b]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    // the last CTA to finish publishes the loss (every pass-0 unit has added its off-diagonal part by then)
    if (threadIdx.x == 0 && p.loss_out != nullptr) {
        __threadfence();
        if (atomicAdd(p.done_counter, 1u) == gridDim.x - 1) {
            __threadfence();
            const double a0 = *reinterpret_cast<volatile double*>(p.loss_acc), a1 = *reinterpret_cast<volatile double*>(p.loss_acc + 1);
            double a2 = 0.0;
            for (int k = 0; k < p.n_parts; ++k) a2 += (double)__ldcg(p.ondiag_part + k);
            double off = a0;
            if (p.hsic) off = a0 + 2.0 * a1 + (double)p.D * (double)(p.D - 1);
            *p.loss_out = (float)((double)p.alpha * a2 + (double)p.lambda * off);
        }
    }
}

}  // namespace abt
