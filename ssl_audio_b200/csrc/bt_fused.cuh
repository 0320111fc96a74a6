// Barlow Twins objective for small batches (N <= 128 rows): ONE tensor-core launch, the D x D matrix never leaves the SM.
//
// Replaces utils/loss.py:17-30 and its autograd backward for the reference's own batch sizes (main.py:115-119 with
// batch 32 ... 128 per GPU; BASELINE config 3: N = 128, D = 2048 / 4096 / 8192).  The two-launch CORR + GRAD form
// (bt_umma_kernel) writes C once and reads it twice; at N = 128 that traffic (3 x 128 MiB at D = 8192) IS the run time.
// Here a CTA owns a 128-row block I of C (and, in the second pass, of C^T) and walks the 64-column blocks J:
//
//     S    = zh_a[:, I]^T zh_b[:, J]            MMA1: K = N, fp16 standardised operands, fp32 accumulator in TMEM
//     loss += sum S^2 (off-diagonal)            epilogue warps, straight out of TMEM (packed fp32x2 FMAs)
//     P    = fp16(S), diagonal zeroed           written back over S in TMEM (tcgen05.st): operand A of MMA2
//     O   += P zh_b[:, J]^T                     MMA2: K = 64 columns of J; O (128 x N fp32) stays in TMEM for the whole walk
//
// and finishes with batch-norm backward on O: dz_a[:, I] = r_a (g - mean_n g - zh_a mean_n(g o zh_a)),
// g = (2 lambda / N^2) O + (G_ii / N) zh_b[:, I]  (+ the HSIC row-sum term), the diagonal in fp32 from the statistics pass.
// This is the shape of an attention forward pass without the softmax ("sequence" = D, "head dimension" = N).  Executed FLOP:
// 8 N D^2 for both gradients (S is recomputed by the second pass) against 6 N D^2 algorithmic.
//
// Both A operands live in tensor memory: Q = zh_a[:, I]^T is transposed into TMEM once per unit, P overwrites the S buffer it
// was computed from; only the zh_b tile (operand B of both MMAs: MN-major for MMA1, K-major for MMA2, the same bytes) is in
// shared memory.  The tiles are stored by the statistics kernel as ready-made shared-memory images (128-byte swizzle), so a
// tile is ONE contiguous 16 KiB bulk copy.  TMEM columns: four S / P buffers at 0 / 64 / 128 / 192, O at 256, Q at 384.
//
// Warp roles (352 threads, 1 CTA / SM): warp 0 TMA producer (+ TMEM allocation), warp 1 issues MMA1, warp 2 issues MMA2,
// warps 3-10 epilogue -- warps 3-6 convert the even steps, warps 7-10 the odd ones (a warp owns all 64 columns of its 32 rows,
// so P can alias S).  A step is a latency chain (MMA1 -> commit -> epilogue wake-up -> S -> P -> arrive -> issuer wake-up ->
// MMA2 -> commit -> S buffer free) and four of them are in flight.  What bounds this kernel (measured with every stage switched
// off in turn, profiles/r2_fused_notes.md): not the tile loads, not the MMAs, not the commits -- the issuing threads.  Every
// step costs a barrier wait or two, descriptor arithmetic and twelve tcgen05.mma issues from ONE thread each, hence two
// issuing warps, and loops that carry no divisions, no constant-bank reloads and no run-time switches.
//
// The walk over J starts at block 2 I + 2 and ends AT blocks 2 I, 2 I + 1: the last two zh_b tiles are zh_b[:, I], which the
// unit epilogue needs (diagonal term), and they are still resident in their ring stages.
#pragma once

namespace abt {

constexpr int FB = 128;                       // rows of C per CTA (UMMA M)
constexpr int JW = 64;                        // columns of C per step
constexpr int kXStages = 11;                  // ring of 16 KiB zh_b tiles (64 columns x up to 128 samples)
constexpr int kXTileBytes = JW * 128 * 2;
constexpr int kTThreads = 352;                // 11 warps: 3 on a scheduler at most -> 168 registers per thread
constexpr int kXOffQ = 0;                     // zh_a[:, I]: two tiles
constexpr int kXOffK = 2 * kXTileBytes;
constexpr int kXOffBar = kXOffK + kXStages * kXTileBytes;
constexpr int kXOffRed = kXOffBar + 512;
constexpr int kXSmemBytes = kXOffRed + 2048 + 1024 /*align slack*/;
static_assert(kXSmemBytes <= 227 * 1024, "fused kernel: shared memory budget");

struct FusedParams {
    int D, N, n_pad;             // n_pad = N rounded up to 32 (UMMA K of MMA1, UMMA N of MMA2); the tile images hold zero rows beyond N
    int n_blocks;                // ceil(D / 128) row blocks
    int pass_count;              // 1 or 2
    int pass_side[2];            // 0: the pass produces dz1 (rows = view-1 dimensions), 1: dz2
    int hsic;
    float alpha, lambda, grad_scale;
    const float* stats;          // StatSlot arrays
    const float* rs1; const float* rs2;      // HSIC: row sums of zh1 / zh2
    void* dz1; void* dz2;
    double* loss_acc;            // [0] += sum_{i != j} c_ij^2, [1] += sum_{i != j} c_ij (HSIC)
    unsigned int* done_counter;  // zeroed per call: the last CTA to finish publishes the loss
    float* loss_out;             // may be null
    const float* ondiag_part; int n_parts;   // on-diagonal loss: one partial sum per block of the statistics kernel
    const __half* zimg1; const __half* zimg2;    // standardised embeddings as tile images: tile t = columns 64 t .. 64 t + 63, n_pad rows of 128 bytes
};

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

// NP: n_pad as a compile-time constant (128: the MMA loops are fully unrolled) or 0 (read from the parameters)
template <typename T, bool HSIC, int NP>
__global__ void __launch_bounds__(kTThreads, 1) bt_fused_kernel(const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* k_full = reinterpret_cast<uint64_t*>(smem + kXOffBar);
    uint64_t* k_empty = k_full + kXStages;
    uint64_t* s_full = k_empty + kXStages;      // [4] MMA1 done -> epilogue group
    uint64_t* p_full = s_full + 4;              // [4] P stored over S -> MMA issuer
    uint64_t* s_free = p_full + 4;              // [4] MMA2 done reading P: the S buffer may be overwritten by MMA1
    uint64_t* o_full = s_free + 4;              // all MMAs of the unit done
    uint64_t* o_empty = o_full + 1;             // unit epilogue done: Q (TMEM and smem) may be overwritten
    uint64_t* q_full = o_empty + 1;             // Q tiles landed in shared memory
    uint64_t* qt_full = q_full + 1;             // Q transposed into TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qt_full + 1);
    float* red = reinterpret_cast<float*>(smem + kXOffRed);          // [2][128][2]
    const uint32_t q_s = smem_u32(smem + kXOffQ), k_s = smem_u32(smem + kXOffK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pad = NP ? NP : p.n_pad;
    const int n_blocks = p.n_blocks;
    const int units = n_blocks * p.pass_count;
    const int nsteps = 2 * n_blocks;                                 // 64-column blocks (one all-zero spare tile when D % 128 == 64)
    const uint32_t chunk_bytes = (uint32_t)n_pad * 128u;            // bytes of one tile image

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kXStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&s_free[s], 1); }
        mbar_init(o_full, 1); mbar_init(o_empty, 1); mbar_init(q_full, 1); mbar_init(qt_full, kEpiWarps);
        mbar_fence_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t o_tmem = tmem_base + 256, qt_tmem = tmem_base + 384;
    // launched with programmatic stream serialisation: everything above overlapped the tail of the statistics kernel; its outputs
    // (tile images, statistics, cleared accumulators) are read only after this point
    griddep_wait();

    if (warp == 0) {
        // ================= TMA producer: one contiguous bulk copy per tile =================
        uint32_t stage = 0, phase = 0, uiter = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit >= n_blocks ? 1 : 0, ib = unit - pass * n_blocks;
            const bool side1 = (pass == 0 ? p.pass_side[0] : p.pass_side[1]) != 0;
            const uint8_t* imgQ = reinterpret_cast<const uint8_t*>(side1 ? p.zimg2 : p.zimg1);
            const uint8_t* imgK = reinterpret_cast<const uint8_t*>(side1 ? p.zimg1 : p.zimg2);
            mbar_wait(o_empty, (uiter & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(q_full, 2u * chunk_bytes);
                bulk_load(smem + kXOffQ, imgQ + (size_t)(2 * ib) * chunk_bytes, 2u * chunk_bytes, q_full);     // tiles 2 ib and 2 ib + 1 are adjacent
            }
            __syncwarp();
            int jb = 2 * ib + 2;                                     // the walk ends with blocks 2 ib, 2 ib + 1 (= the columns of I)
            if (jb >= nsteps) jb -= nsteps;
            const uint8_t* src = imgK + (size_t)jb * chunk_bytes;
            const uint8_t* const src_end = imgK + (size_t)nsteps * chunk_bytes;
            for (int s = 0; s < nsteps; ++s) {
                mbar_wait(&k_empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&k_full[stage], chunk_bytes);
                    bulk_load(smem + kXOffK + stage * kXTileBytes, src, chunk_bytes, &k_full[stage]);
                }
                __syncwarp();
                src += chunk_bytes;
                if (src == src_end) src = imgK;
                if (++stage == kXStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA1 issuer: S[b] = Q^T x tile =================
        // descriptors: high word = SBO (1024 B between 8-row groups) | version 1 | 128-byte swizzle; low word = address >> 4 | LBO << 16
        const uint64_t d_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const uint32_t lo_mn = (((chunk_bytes >> 4) & 0x3FFF) << 16) | ((k_s & 0x3FFFF) >> 4);      // zh_b tile read MN-major (LBO unused: one 64-column chunk)
        const uint32_t idesc1 = make_idesc_f16(FB, JW, 0, 1, 0);         // S (128 x 64) = Q^T (TMEM) x tile (smem, MN-major)
        const int k1 = n_pad / 16;
        uint32_t st1 = 0, ph1 = 0, c1 = 0, uiter = 0;                    // ring position, step counter over all units
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            mbar_wait(qt_full, uiter & 1);
            for (int s = 0; s < nsteps; ++s) {
                const uint32_t b = c1 & 3;
                mbar_wait(&s_free[b], ((c1 >> 2) & 1) ^ 1);               // MMA2 of four steps ago has consumed the P tile in this buffer
                mbar_wait(&k_full[st1], ph1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t s_tmem = tmem_base + b * JW;
                    const uint32_t db = lo_mn + st1 * (kXTileBytes >> 4);
                    if (NP) {
#pragma unroll
                        for (int ks = 0; ks < NP / 16; ++ks) umma_f16_ts(s_tmem, qt_tmem + ks * 8, d_hi | (db + ks * (2048 >> 4)), idesc1, ks > 0 ? 1u : 0u);
                    } else {
                        for (int ks = 0; ks < k1; ++ks) umma_f16_ts(s_tmem, qt_tmem + ks * 8, d_hi | (db + ks * (2048 >> 4)), idesc1, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(&s_full[b]);
                }
                __syncwarp();
                ++c1;
                if (++st1 == kXStages) { st1 = 0; ph1 ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ================= MMA2 issuer: O += P[b] x tile^T =================
        const uint64_t d_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const uint32_t lo_k = (1u << 16) | ((k_s & 0x3FFFF) >> 4);                                    // the same tile read K-major
        const uint32_t idesc2 = make_idesc_f16(FB, n_pad, 0, 0, 0);      // O (128 x n_pad) += P (TMEM) x tile^T (smem, K-major)
        uint32_t st2 = 0, c2 = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            for (int t = 0; t < nsteps; ++t) {
                const uint32_t b = c2 & 3;
                mbar_wait(&p_full[b], (c2 >> 2) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t p_tmem = tmem_base + b * JW;
                    const uint32_t db = lo_k + st2 * (kXTileBytes >> 4);
                    umma_f16_ts(o_tmem, p_tmem, d_hi | db, idesc2, t > 0 ? 1u : 0u);
#pragma unroll
                    for (int ks = 1; ks < JW / 16; ++ks) umma_f16_ts(o_tmem, p_tmem + ks * 8, d_hi | (db + ks * 2), idesc2, 1u);
                    umma_commit(&s_free[b]);
                    if (t < nsteps - 2) umma_commit(&k_empty[st2]);        // the last two tiles (= zh_b[:, I]) stay for the unit epilogue
                    else if (t == nsteps - 1) umma_commit(o_full);
                }
                __syncwarp();
                ++c2;
                if (++st2 == kXStages) st2 = 0;
            }
        }
    } else {
        // ================= epilogue warps (3-10) =================
        const int q = warp & 3;               // TMEM lane quarter of this warp (warps 3-6 and 7-10 each cover the four quarters)
        const int h = (warp - 3) >> 2;        // group: steps with (step & 1) == h; sample half in the unit prologue / epilogue
        const int il = q * 32 + lane;         // row of the block (dimension) owned by this thread
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        // swizzled byte offset of column (il & 63) inside a 128-byte tile row whose index is k modulo 8 (k is a compile-time constant at
        // every use; the eight values are not kept live across the walk)
        const uint32_t xc = (uint32_t)(il & 63) >> 3, xw = (uint32_t)(il & 7) * 2u;
        auto xoff = [&](int k) -> uint32_t { return ((xc ^ (uint32_t)k) << 4) + xw; };
        const uint32_t q_col = q_s + (uint32_t)(il >> 6) * chunk_bytes + (uint32_t)h * 64u * 128u;      // Q tiles, this thread's column, sample 64 h
        uint32_t uiter = 0, c = 0;
        const float invN = 1.0f / (float)p.N;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++uiter) {
            const int pass = unit >= n_blocks ? 1 : 0, ib = unit - pass * n_blocks;
            const int side = pass == 0 ? p.pass_side[0] : p.pass_side[1];
            // ---- unit prologue: Q^T into TMEM (this warp: samples 64 h .. 64 h + 63 of its 32 dimensions, packed in pairs)
            mbar_wait(q_full, uiter & 1);
            if (h * 64 < n_pad) {
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
            const uint32_t st_a = (c - 2) % kXStages, st_b = (c - 1) % kXStages;      // tiles of the last two steps: zh_b[:, I]
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
            const bool have_a = n0 < n_pad, have_b = n0 + 32 < n_pad;          // n_pad is a multiple of 32
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
            if (threadIdx.x == 96) {
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
