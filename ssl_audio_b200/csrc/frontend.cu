// Audio frontend for sm_100a: wav -> power STFT -> HTK mel -> log -> z-score, fused in one kernel.
//
// Replaces torchaudio.transforms.MelSpectrogram as the reference constructs it at
// datasets.py:39-48 (n_fft 1024, hop 160, periodic Hann, center/reflect, power 2, 64 HTK mel
// bands without norm), `(mel + eps).log()` (datasets.py:115) and `(lms - mean)/std`
// (datasets.py:118-119).  Reference tree: /root/reference.
//
// Kernel shape (HBM/FP32-bound, no tensor cores: a 1024-point DFT in split bf16/tf32 fails the
// 1e-3 log-mel tolerance):
//   * one CTA = one clip x 32 consecutive frames; the (31*hop + 1024)-sample span is staged once
//     in shared memory with coalesced loads (reflect padding resolved by index mirroring);
//   * one warp = two frames at a time, packed as re/im of ONE 1024-point complex FFT
//     (two-for-one real FFT).  1024 = 32 x 32: each lane runs an in-register 32-point FFT, the
//     warp transposes through a padded (bank-conflict-free) shared tile, each lane runs a second
//     32-point FFT;  X[k] and X[1024-k] are recombined into the two power spectra;
//   * the mel projection uses the filterbank's sparsity (<= 2 filters per bin, 970 non-zeros of
//     32832), as a block-wide pass over the 16 frames of a round: thread = (frame, band group), CSR weights;
//   * log, z-score, and a staged, time-contiguous (coalesced) store.
#include "abt_internal.h"
#include "fft32_gen.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

struct abt_logmel_plan {
    abt_mel_config cfg;
    float* d_window;   // n_fft, win_length window centre-padded
    float2* d_twiddle; // [k2][n1] -> exp(-2 pi i n1 k2 / 1024)
    int* d_mel_start;  // n_mels
    int* d_mel_len;    // n_mels
    int* d_mel_off;    // n_mels, offset into d_mel_w
    float* d_mel_w;    // nnz
    int nnz;
    int max_len;
};

namespace abt {

constexpr int kNfft = 1024;
constexpr int kMaxMels = 256;
constexpr int kTileFrames = 32;
constexpr int kWarps = 8;
constexpr int kScratchFloats = 2 * 32 * 33 + 2;   // per warp; the +2 staggers the warps' power spectra over the banks for the mel pass
constexpr float kF32Eps = 1.1920928955078125e-07f;

struct LogmelArgs {
    const float* wav;
    long long row_stride;
    const int* wav_offset;
    const int* wav_origin;    // span input: row b holds the samples [wav_origin[b], wav_origin[b] + row_stride) of clip b; nullptr = whole clips
    int n_samples;
    int n_frames_total;    // 1 + n_samples / hop
    int hop;
    const int* frame_start;   // mode C: first frame per clip; nullptr in mode F
    int n_frames_out;         // frames written per clip (mode F: n_frames_total)
    float* out_base;
    const int* out_slot;
    long long out_slot_stride;
    int apply_norm;
    float norm_mean, inv_std;
    float pad_value;          // value of frames beyond the clip end (already normalised)
    const float* window;
    const float2* twiddle;
    const int* mel_start;
    const int* mel_len;
    const int* mel_off;
    const float* mel_w;
    int nnz;
    int n_mels;               // bands (64 in every configuration of the reference; any 1..256 works)
};

__global__ void __launch_bounds__(kWarps * 32, 2) logmel_kernel(const LogmelArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int span = (kTileFrames - 1) * a.hop + kNfft;
    float* s_x = smem;                                  // span samples (padded coordinates)
    float* s_scratch = s_x + ((span + 31) & ~31);       // kWarps * kScratchFloats
    float* s_out = s_scratch + kWarps * kScratchFloats; // n_mels * (kTileFrames + 1)
    float* s_melw = s_out + a.n_mels * (kTileFrames + 1);  // nnz
    int* s_mel_start = reinterpret_cast<int*>(s_melw + a.nnz);   // n_mels start, n_mels len, n_mels off
    __shared__ __align__(8) unsigned long long s_bar;

    const int clip = blockIdx.y;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // -1 in frame_start / wav_offset means "no crop was drawn" (clip not longer than the crop)
    const int f_first = (a.frame_start ? max(a.frame_start[clip], 0) : 0) + tile * kTileFrames;   // absolute frame index
    const float* wav = a.wav + (long long)clip * a.row_stride + (a.wav_offset ? max(a.wav_offset[clip], 0) : 0) - (a.wav_origin ? a.wav_origin[clip] : 0);

    // ---- stage the sample span (reflect padding: padded index s -> original s - n_fft/2, mirrored)
    const long long o_first = (long long)f_first * a.hop - kNfft / 2;      // original index of padded sample 0 of this tile
    const bool interior = o_first >= 0 && o_first + span <= a.n_samples && (span & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(wav + o_first) & 15) == 0;
    if (interior) {
        // one bulk async copy (TMA, 1-D) of the whole span, issued by a single thread
        const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&s_bar));
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(span * 4) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             static_cast<unsigned>(__cvta_generic_to_shared(s_x))),
                         "l"(wav + o_first), "r"(span * 4), "r"(bar)
                         : "memory");
        }
    } else {
        // clip edges / unaligned rows: mirrored scalar loads, eight independent requests in flight per thread
        for (int base = 0; base < span; base += 8 * kWarps * 32) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int sidx = base + u * kWarps * 32 + tid;
                long long o = o_first + sidx;
                if (o < 0) o = -o;
                if (o >= a.n_samples) o = 2LL * (a.n_samples - 1) - o;
                v[u] = (sidx < span && o >= 0 && o < a.n_samples) ? __ldg(wav + o) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int sidx = base + u * kWarps * 32 + tid;
                if (sidx < span) s_x[sidx] = v[u];
            }
        }
    }
    for (int i = tid; i < a.nnz; i += blockDim.x) s_melw[i] = 0.25f * a.mel_w[i];      // x 1/4: see the power spectrum below (exact scaling)
    for (int m = tid; m < a.n_mels; m += blockDim.x) {
        s_mel_start[m] = a.mel_start[m];
        s_mel_start[a.n_mels + m] = a.mel_len[m];
        s_mel_start[2 * a.n_mels + m] = a.mel_off[m];
    }
    __syncthreads();      // also orders the mbarrier init before the waits below
    if (interior) {
        const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&s_bar));
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.b32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }

    float2* sc = reinterpret_cast<float2*>(s_scratch + warp * kScratchFloats);     // 32 x 33 complex, later 513 x (Pa, Pb)

    for (int round = 0; round < kTileFrames / (2 * kWarps); ++round) {
        const int fa = round * 2 * kWarps + 2 * warp;     // local frame indices of this warp's pair
        const int fb = fa + 1;
        const bool any_valid = (f_first + fa) < a.n_frames_total && (tile * kTileFrames + fa) < a.n_frames_out;
        if (any_valid) {   // warp-uniform
            // A complex value is ONE 64-bit register pair (re, im) and all arithmetic runs on Blackwell's packed fp32x2 pipe (FADD2 /
            // FMUL2 / FFMA2; half swaps and signs are operand modifiers): about half the FP instructions of the scalar form.
            unsigned long long x[32];
            // frame a -> real part, frame b -> imaginary part; lane = n1, register = n2, n = n1 + 32 n2
            const float* xa = s_x + fa * a.hop;
            const float* xb = s_x + fb * a.hop;
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) {
                const int n = lane + 32 * n2;
                const float w = __ldg(a.window + n);
                x[n2] = c2_mul(c2_pack(xa[n], xb[n]), c2_pack(w, w));
            }
            fft32p(x);   // register p: Y[n1 = lane][k2 = bitrev(p)]
            // twiddle W1024^(n1 k2) and transpose through shared memory (64-bit accesses, conflict-free with the 33-stride)
            __syncwarp();
            unsigned long long* sc64 = reinterpret_cast<unsigned long long*>(sc);
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const int k2 = bitrev5(p);
                const float2 w = __ldg(a.twiddle + k2 * 32 + lane);
                sc64[lane * 33 + k2] = c2_cmul(x[p], c2_pack(w.x, w.y));
            }
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) x[n1] = sc64[n1 * 33 + lane];     // lane = k2, register = n1
            fft32p(x);   // register p: Z[32 k1 + k2], k1 = bitrev(p), k2 = lane
            __syncwarp();    // all lanes have read the transposed tile: the scratch can hold the power spectra now
            // Z = A + i B with A, B the (Hermitian) spectra of the two real frames: the partner bin Z[1024 - k] lives in lane
            // (32 - lane) & 31, register k1' = 31 - k1 (lane 0: its own register (32 - k1) & 31) -> one 64-bit shuffle per value.
            //   2 A[k] = (re + qr, im - qi),  2 B[k] = (im + qi, qr - re);  with U = Z + Q = (2 Re A, 2 Re B) and
            //   V = (im - qi, qr - re) = swap(Z) (+, -) + swap(Q) (-, +) = (2 Im A, 2 Im B):  (4 |A|^2, 4 |B|^2) = U o U + V o V
            const int lp = (32 - lane) & 31;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const int p = bitrev5(k1), pq = bitrev5(31 - k1), pq0 = bitrev5((32 - k1) & 31);
                unsigned long long qv = __shfl_sync(0xffffffffu, x[pq], lp);
                if (lane == 0) qv = x[pq0];
                float zr, zi, qr, qi;
                c2_unpack(x[p], zr, zi);
                c2_unpack(qv, qr, qi);
                const unsigned long long u = c2_add(x[p], qv);
                const unsigned long long v = c2_add(c2_pack(zi, -zr), c2_pack(-qi, qr));
                sc64[32 * k1 + lane] = c2_fma(u, u, c2_mul(v, v));      // 4 |A|^2, 4 |B|^2: the 1/4 lives in the mel weights
            }
            if (lane == 0) {   // k = 512 pairs with itself: Z[512] = A[512] + i B[512] with both real
                const int p16 = bitrev5(16);
                const unsigned long long z = x[p16];
                sc64[512] = c2_mul(c2_mul(z, z), c2_pack(4.0f, 4.0f));
            }
        }
        // ---- sparse mel projection of the round's 16 frames, across the warps: thread = (frame f, band group g).  The 16 lanes of a
        // half-warp read the same weight (broadcast) and 16 different frames' power -- frame f = 2 * warp + {0, 1} sits at float offset
        // warp * kScratchFloats + 2 k + {0, 1}, i.e. bank (f + 2 k) mod 32: conflict-free.  Band lengths grow with the band index, so a
        // thread takes bands g, 31 - g, 32 + g, 63 - g, ...: about 61 of the 970 non-zero weights each at 64 bands.
        __syncthreads();
        if ((f_first + round * 2 * kWarps) < a.n_frames_total && (tile * kTileFrames + round * 2 * kWarps) < a.n_frames_out) {   // block-uniform
            const int f = tid & 15, g = tid >> 4;
            const float* pf = s_scratch + (f >> 1) * kScratchFloats + (f & 1);
            for (int mb = 0; mb < a.n_mels; mb += 64)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = mb + (q >> 1) * 32 + ((q & 1) ? 31 - g : g);
                if (m >= a.n_mels) continue;
                const int ks = s_mel_start[m], kl = s_mel_start[a.n_mels + m];
                const float* w = s_melw + s_mel_start[2 * a.n_mels + m];
                const float* pk = pf + 2 * ks;
                float m0 = 0.f, m1 = 0.f;
                int k = 0;
#pragma unroll 2
                for (; k + 1 < kl; k += 2) {
                    m0 = fmaf(w[k], pk[2 * k], m0);
                    m1 = fmaf(w[k + 1], pk[2 * k + 2], m1);
                }
                if (k < kl) m0 = fmaf(w[k], pk[2 * k], m0);
                float l = __logf(m0 + m1 + kF32Eps);      // lg2.approx * ln 2: |err| ~ 1e-6, far inside the 1e-3 log-mel tolerance
                if (a.apply_norm) l = (l - a.norm_mean) * a.inv_std;
                s_out[m * (kTileFrames + 1) + round * 2 * kWarps + f] = l;
            }
        }
        if (round + 1 < kTileFrames / (2 * kWarps)) __syncthreads();      // the next round's transposes reuse the scratch
    }
    __syncthreads();
    // ---- coalesced store: rows of up to 32 consecutive frames per mel band
    const long long slot = a.out_slot ? a.out_slot[clip] : clip;
    float* out = a.out_base + slot * a.out_slot_stride;
    const int t_base = tile * kTileFrames;       // frame offset inside the output row
    for (int idx = tid; idx < a.n_mels * kTileFrames; idx += blockDim.x) {
        const int m = idx / kTileFrames, f = idx % kTileFrames;
        const int t_out = t_base + f;
        if (t_out < a.n_frames_out) {
            const bool real = (f_first + f) < a.n_frames_total;
            out[(size_t)m * a.n_frames_out + t_out] = real ? s_out[m * (kTileFrames + 1) + f] : a.pad_value;
        }
    }
}

// ------------------------------------------------------------------------------------------
// precomputed log-mel: time crop / right pad + z-score  (datasets.py:342-354)
// ------------------------------------------------------------------------------------------
__global__ void lms_crop_norm_kernel(const float* __restrict__ lms, int n_mels, int t_full, const int* __restrict__ frame_start, int n_frames,
                                     int apply_norm, float mean, float inv_std, float* __restrict__ out_base,
                                     const int* __restrict__ out_slot, long long out_slot_stride) {
    const int clip = blockIdx.y;
    const int start = frame_start ? frame_start[clip] : 0;
    const long long slot = out_slot ? out_slot[clip] : clip;
    float* out = out_base + slot * out_slot_stride;
    const float* src = lms + (size_t)clip * n_mels * t_full;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_mels * n_frames; idx += gridDim.x * blockDim.x) {
        const int m = idx / n_frames, t = idx % n_frames;
        const int ts = (start < 0 ? 0 : start) + t;
        float v = (ts < t_full) ? src[(size_t)m * t_full + ts] : 0.0f;   // right zero-pad BEFORE normalisation
        if (apply_norm) v = (v - mean) * inv_std;
        out[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------
// views: Mixup -> RandomResizeCrop (bicubic, align_corners) -> RandomLinearFader
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x) {   // |x| <= 1, A = -0.75
    return ((-0.75f + 2.0f) * x - (-0.75f + 3.0f)) * x * x + 1.0f;
}
__device__ __forceinline__ float cubic2(float x) {   // 1 < |x| < 2
    return ((-0.75f * x - 5.0f * -0.75f) * x + 8.0f * -0.75f) * x - 4.0f * -0.75f;
}

constexpr int kViewThreads = 384;

// exp / log on the special-function unit (ex2 / lg2, relative error ~1e-6, far inside the 1e-3 tolerance of the views).  The .ftz forms
// skip the denormal range checks nvcc wraps around __expf / __logf (4 extra instructions per call); flushing is harmless here: every
// sum below carries + eps, so a flushed exp() term is below half an ulp of the result either way.
__device__ __forceinline__ float exp_sfu(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ float lds_f32(unsigned addr) {     // volatile asm: never moved across the __syncthreads between the passes
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float log_sfu(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.6931471805599453f;
}

// One CTA per (clip, view).  Three phases over shared memory:
//   1. the crop rows of the virtual canvas: log-mixup-exp of x with its partner inside the pasted region, zeros outside
//      (every canvas element the taps can touch is written exactly once; float4 global loads);
//   2. horizontal cubic pass  H[y][ox] = sum_c canvas[y][tx_c(ox)] cx_c(ox)   for the crop rows only;
//   3. vertical cubic pass + fader, coalesced stores.
// ATen's upsample_bicubic2d interpolates along x inside each of the 4 tap rows and then along y, so the separable form
// performs the same arithmetic in the same order with 8 instead of 16 taps per output.  A thread keeps its output column:
// the x taps / coefficients live in registers.
__global__ void __launch_bounds__(kViewThreads) views_kernel(const abt_views_args a) {
    extern __shared__ __align__(16) float smem[];
    float* canvas = smem;                                               // canvas_h * canvas_w
    float* hbuf = canvas + a.canvas_h * a.canvas_w;                     // canvas_h * out_w (rows of the crop)
    float4* coef = reinterpret_cast<float4*>(hbuf + a.canvas_h * a.out_w);   // out_w + out_h
    int4* tap = reinterpret_cast<int4*>(coef + (a.out_w + a.out_h));    // out_w + out_h (clamped taps, canvas coordinates)
    const int clip = blockIdx.x, view = blockIdx.y;
    const int tid = threadIdx.x;
    const int pstride = a.param_stride > 0 ? a.param_stride : a.n_views;
    const abt_view_params p = a.params[(size_t)clip * pstride + a.view_offset + view];
    const long long xs = a.x_slot ? a.x_slot[clip] : clip;
    const float* x = a.x + xs * a.x_slot_stride;
    const float* z = nullptr;
    if ((p.flags & 1) && p.z_kind == 1) z = a.bank + (long long)p.z_index * a.bank_slot_stride;
    if ((p.flags & 1) && p.z_kind == 2) z = a.x + (long long)(a.x_slot ? a.x_slot[p.z_index] : p.z_index) * a.x_slot_stride;
    // MixGaussianNoise (augmentations.py:133-140): log((1 - lambd) exp(m) + exp(lambd n) + eps) on the mixed-up clip m, n ~ N(0, 1)
    const float* gn = ((p.flags & 8) && a.noise != nullptr) ? a.noise + ((size_t)clip * a.noise_views + a.view_offset + view) * (size_t)(a.in_h * a.in_w) : nullptr;

    const int y0 = (a.canvas_h - a.in_h) / 2, x0 = (a.canvas_w - a.in_w) / 2;
    const bool rrc = (p.flags & 2) != 0;
    const int ci = rrc ? p.i : y0, cj = rrc ? p.j : x0, chh = rrc ? p.h : a.in_h, cww = rrc ? p.w : a.in_w;

    // ---- phase 1: canvas rows [ci, ci + chh)   (augmentations.py:43-48, :81-85)
    if (((a.canvas_w | x0 | a.in_w) & 3) == 0) {
        // only the crop box is ever read by the taps: rows [ci, ci + chh), float4 groups covering columns [cj, cj + cww);
        // a thread keeps its column group and strides over the rows (one integer division per thread, none per element)
        const int c_lo = cj & ~3, n_groups = ((cj + cww + 3) >> 2) - (cj >> 2);
        const int rows_per_pass = kViewThreads / n_groups, gsub = tid / n_groups, col = c_lo + ((tid % n_groups) << 2);
        for (int r = ci + gsub; gsub < rows_per_pass && r < ci + chh; r += rows_per_pass) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r >= y0 && r < y0 + a.in_h && col >= x0 && col < x0 + a.in_w) {
                const int e = (r - y0) * a.in_w + (col - x0);
                v = __ldg(reinterpret_cast<const float4*>(x + e));
                if (z != nullptr) {
                    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + e));
                    v.x = log_sfu(p.w_x * exp_sfu(v.x) + p.w_z * exp_sfu(zv.x) + kF32Eps);
                    v.y = log_sfu(p.w_x * exp_sfu(v.y) + p.w_z * exp_sfu(zv.y) + kF32Eps);
                    v.z = log_sfu(p.w_x * exp_sfu(v.z) + p.w_z * exp_sfu(zv.z) + kF32Eps);
                    v.w = log_sfu(p.w_x * exp_sfu(v.w) + p.w_z * exp_sfu(zv.w) + kF32Eps);
                }
                if (gn != nullptr) {
                    const float4 nv = __ldg(reinterpret_cast<const float4*>(gn + e));
                    v.x = log_sfu(p.g_keep * exp_sfu(v.x) + exp_sfu(p.g_lambda * nv.x) + kF32Eps);
                    v.y = log_sfu(p.g_keep * exp_sfu(v.y) + exp_sfu(p.g_lambda * nv.y) + kF32Eps);
                    v.z = log_sfu(p.g_keep * exp_sfu(v.z) + exp_sfu(p.g_lambda * nv.z) + kF32Eps);
                    v.w = log_sfu(p.g_keep * exp_sfu(v.w) + exp_sfu(p.g_lambda * nv.w) + kF32Eps);
                }
            }
            *reinterpret_cast<float4*>(canvas + r * a.canvas_w + col) = v;
        }
    } else {
        for (int g = tid; g < chh * a.canvas_w; g += kViewThreads) {
            const int r = ci + g / a.canvas_w, col = g % a.canvas_w;
            float v = 0.f;
            if (r >= y0 && r < y0 + a.in_h && col >= x0 && col < x0 + a.in_w) {
                const int e = (r - y0) * a.in_w + (col - x0);
                v = __ldg(x + e);
                if (z != nullptr) v = log_sfu(p.w_x * exp_sfu(v) + p.w_z * exp_sfu(__ldg(z + e)) + kF32Eps);
                if (gn != nullptr) v = log_sfu(p.g_keep * exp_sfu(v) + exp_sfu(p.g_lambda * __ldg(gn + e)) + kF32Eps);
            }
            canvas[r * a.canvas_w + col] = v;
        }
    }
    // bicubic tables (ATen upsample_bicubic2d, align_corners=True): src = dst * (in-1)/(out-1)
    if (tid < a.out_w + a.out_h) {
        const bool is_x = tid < a.out_w;
        const int o = is_x ? tid : tid - a.out_w;
        const int n_in_ax = is_x ? cww : chh, n_out_ax = is_x ? a.out_w : a.out_h;
        const float scale = n_out_ax > 1 ? (float)(n_in_ax - 1) / (float)(n_out_ax - 1) : 0.f;
        const float pos = scale * (float)o;
        const float fl = floorf(pos);
        const float t = pos - fl;
        const int base = (int)fl;
        int sidx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int sk = base - 1 + k;
            sk = sk < 0 ? 0 : (sk > n_in_ax - 1 ? n_in_ax - 1 : sk);
            sidx[k] = is_x ? sk + cj : sk * a.out_w * 4;   // x taps: canvas columns; y taps: BYTE offsets of the hbuf rows (relative to the crop's first row)
        }
        tap[tid] = make_int4(sidx[0], sidx[1], sidx[2], sidx[3]);
        coef[tid] = make_float4(cubic2(t + 1.0f), cubic1(t), cubic1(1.0f - t), cubic2(2.0f - t));
    }
    __syncthreads();
    // ---- phase 2: horizontal pass over the crop rows; thread = (output column, row group)
    const int n_groups = kViewThreads / a.out_w;
    const int ox = tid % a.out_w, grp = tid / a.out_w;
    const bool active = grp < n_groups;
    if (active) {
        const int4 tx = tap[ox];
        const float4 cx = coef[ox];
        // shared-memory byte addresses: the four taps of this thread's column in row ci + grp, advanced by n_groups rows per iteration
        const unsigned row0 = static_cast<unsigned>(__cvta_generic_to_shared(canvas + (ci + grp) * a.canvas_w));
        unsigned t0 = row0 + tx.x * 4, t1 = row0 + tx.y * 4, t2 = row0 + tx.z * 4, t3 = row0 + tx.w * 4;
        const unsigned row_step = (unsigned)(n_groups * a.canvas_w) * 4u;
        float* hdst = hbuf + grp * a.out_w + ox;
        const int hstep = n_groups * a.out_w;
        for (int yy = grp; yy < chh; yy += n_groups, t0 += row_step, t1 += row_step, t2 += row_step, t3 += row_step, hdst += hstep)
            *hdst = lds_f32(t0) * cx.x + lds_f32(t1) * cx.y + lds_f32(t2) * cx.z + lds_f32(t3) * cx.w;
    }
    __syncthreads();
    // ---- phase 3: vertical pass + fader (torch.linspace(head, tail, T) evaluated from both ends with one rounding)
    if (active) {
        const bool fade = (p.flags & 4) != 0;
        const float step = a.out_w > 1 ? (p.tail - p.head) / (float)(a.out_w - 1) : 0.f;
        const float fv = !fade ? 0.f : ((ox < a.out_w / 2) ? fmaf(step, (float)ox, p.head) : fmaf(-step, (float)(a.out_w - 1 - ox), p.tail));
        // everything that changes per row advances by a constant: the table pointers, the output pointer; the taps are byte offsets
        float* out = a.outs[view] + (size_t)clip * a.out_h * a.out_w + (size_t)grp * a.out_w + ox;
        const size_t out_step = (size_t)n_groups * a.out_w;
        const unsigned hcol = static_cast<unsigned>(__cvta_generic_to_shared(hbuf + ox));
        const int4* tp = tap + a.out_w + grp;
        const float4* cp = coef + a.out_w + grp;
        for (int oy = grp; oy < a.out_h; oy += n_groups, tp += n_groups, cp += n_groups, out += out_step) {
            const int4 ty = *tp;
            const float4 cy = *cp;
            float acc = 0.f;
            acc += lds_f32(hcol + ty.x) * cy.x;
            acc += lds_f32(hcol + ty.y) * cy.y;
            acc += lds_f32(hcol + ty.z) * cy.z;
            acc += lds_f32(hcol + ty.w) * cy.w;
            *out = fade ? acc + fv : acc;
        }
    }
}

__global__ void bank_push_kernel(const float* __restrict__ x, long long x_stride, int clip_elems, float* __restrict__ bank,
                                 long long bank_slot_stride, const int* __restrict__ slot) {
    const int clip = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(x + (long long)clip * x_stride);
    float4* dst = reinterpret_cast<float4*>(bank + (long long)slot[clip] * bank_slot_stride);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < clip_elems / 4; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// crop-first input staging: copy only the samples the cropped frames need, (n_fft + (n_frames-1) hop) per clip,
// from the waveforms -- which may live in MAPPED PINNED HOST memory (zero-copy reads over PCIe) -- into a compact
// device buffer.  For 10 s clips and a 96-frame crop this moves 65 KB instead of 640 KB per clip across PCIe.
// Minimal footprint on purpose -- ONE warp per SM, 32 registers per thread, no shared memory, persistent over the clips: the
// tensor-core kernel of the objective leaves exactly 1024 registers per SM free, so this kernel stays resident next to it and the
// PCIe reads (about 1.3 ms per 1024-clip batch) overlap the compute of the previous step instead of serialising with it.  Four
// independent 16-byte loads per lane keep ~300 KB in flight over PCIe, enough to saturate it.
// ------------------------------------------------------------------------------------------
int g_gather_blocks = 4;
__global__ void __launch_bounds__(1024, 1) wav_span_gather_kernel(const float* __restrict__ wav, long long row_stride, int n_samples, int n_clips,
                                                                                   const int* __restrict__ frame_start, int hop, int half_fft, int span_len,
                                                                                   float* __restrict__ spans, int* __restrict__ origin) {
    // A few fat blocks (one warp per clip, 32 warps per block): PCIe needs only ~100 KB of reads in flight, and confined to a handful
    // of SMs the gather runs BESIDE a persistent tensor-core kernel that leaves those SMs free (abt_debug_set(12, k)) instead of
    // waiting for it -- or making it wait: a GEMM CTA needs a whole SM's registers.
    const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    for (int clip = blockIdx.x * warps + (threadIdx.x >> 5); clip < n_clips; clip += gridDim.x * warps) {
        const int f0 = frame_start ? max(frame_start[clip], 0) : 0;
        int o = f0 * hop - half_fft;                 // first padded-coordinate sample of the crop, in clip coordinates
        o = min(o, n_samples - span_len);
        o = max(o, 0);
        o &= ~3;                                     // keep 16-byte alignment of the source (rows are 16-byte aligned)
        if (lane == 0) origin[clip] = o;
        const int len = min(span_len, n_samples - o);
        const float4* src = reinterpret_cast<const float4*>(wav + (long long)clip * row_stride + o);
        float4* dst = reinterpret_cast<float4*>(spans + (long long)clip * span_len);
        const int n4 = len >> 2;
        int i = lane;
        for (; i + 96 < n4; i += 128) {
            const float4 v0 = __ldcs(src + i), v1 = __ldcs(src + i + 32), v2 = __ldcs(src + i + 64), v3 = __ldcs(src + i + 96);
            dst[i] = v0; dst[i + 32] = v1; dst[i + 64] = v2; dst[i + 96] = v3;
        }
        for (; i < n4; i += 32) dst[i] = __ldcs(src + i);
        if (lane < (len & 3)) reinterpret_cast<float*>(dst)[(n4 << 2) + lane] = reinterpret_cast<const float*>(src)[(n4 << 2) + lane];
    }
}

// ------------------------------------------------------------------------------------------
// NormalizeBatch (augmentations.py:217-232, applied per crop at main.py:62-66): per-channel mean and UNBIASED std over the
// batch, frequency and time axes, out = (x - mean) / max(std, eps).  Two launches: block partial sums (fp32 per thread over a
// short strip, double across threads and blocks), then every block folds the partials and normalises its share.
// ------------------------------------------------------------------------------------------
constexpr int kNbBlocks = 256;     // partial sums per channel

__global__ void __launch_bounds__(256) batch_stats_kernel(const float* __restrict__ x, int n_batch, int n_ch, int hw, double* __restrict__ partials) {
    __shared__ double red[2][8];
    const int ch = blockIdx.y;
    const long long per_ch = (long long)n_batch * hw;
    double s1 = 0.0, s2 = 0.0;
    const float shift = __ldg(x + (size_t)ch * hw);      // shifted data: a large common offset must not cancel in the fp32 squares
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < per_ch; i0 += (long long)gridDim.x * blockDim.x * 8) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const long long i = i0 + k;
            if (i < per_ch) {
                const float v = __ldg(x + ((i / hw) * n_ch + ch) * hw + (i % hw)) - shift;
                a += v; q = fmaf(v, v, q);
            }
        }
        s1 += (double)a; s2 += (double)q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
        partials[((size_t)ch * gridDim.x + blockIdx.x) * 2] = t1;
        partials[((size_t)ch * gridDim.x + blockIdx.x) * 2 + 1] = t2;
    }
}

__global__ void __launch_bounds__(256) batch_norm_apply_kernel(const float* __restrict__ x, float* __restrict__ out, int n_batch, int n_ch, int hw,
                                                               const double* __restrict__ partials, int n_partials) {
    __shared__ double red[2][8];
    __shared__ float ms[2];
    const int ch = blockIdx.y;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) {
        s1 += partials[((size_t)ch * n_partials + i) * 2];
        s2 += partials[((size_t)ch * n_partials + i) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
        const double n = (double)n_batch * hw;
        const double dmean = t1 / n;                                                      // mean of the shifted data
        const double var = n > 1.0 ? fmax((t2 - t1 * dmean) / (n - 1.0), 0.0) : 0.0;     // torch.std: unbiased
        ms[0] = (float)((double)__ldg(x + (size_t)ch * hw) + dmean);
        ms[1] = 1.0f / fmaxf((float)sqrt(var), kF32Eps);                                  // clamp(std, finfo.eps, finfo.max)
    }
    __syncthreads();
    const float mean = ms[0], inv = ms[1];
    const long long per_ch = (long long)n_batch * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_ch; i += (long long)gridDim.x * blockDim.x) {
        const long long o = ((i / hw) * n_ch + ch) * hw + (i % hw);
        out[o] = (__ldg(x + o) - mean) * inv;
    }
}

// ------------------------------------------------------------------------------------------
// RunningNorm (augmentations.py:187-210; --pre_norm, main.py:272-277): online normalisation with running statistics over the
// SAMPLES, in sample order.  Per sample b: m_b = mean(x_b), mu <- m_b (first sample) or mu + (m_b - mu) / n, v_b = mean((x_b - mu)^2),
// s2 likewise, out_b = (x_b - mu) / clamp(sqrt(s2), eps).  (n is the count BEFORE the update, exactly as the reference's RunningMean.)
// Three launches for a batch: per-sample sums, the sequential scalar recurrence (one thread), the elementwise normalisation.
// state: 3 doubles on the device -- count, mu, s2 -- owned by the Python module.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_sums_kernel(const float* __restrict__ x, int elems, double* __restrict__ sums /* [B][2] */) {
    __shared__ double red[2][8];
    const float* xb = x + (size_t)blockIdx.x * elems;
    double s1 = 0.0, s2 = 0.0;
    for (int i0 = threadIdx.x * 8; i0 < elems; i0 += blockDim.x * 8) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k < elems) { const float v = __ldg(xb + i0 + k); a += v; q = fmaf(v, v, q); }
        s1 += (double)a; s2 += (double)q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
        sums[2 * blockIdx.x] = t1 / elems;          // mean(x_b)
        sums[2 * blockIdx.x + 1] = t2 / elems;      // mean(x_b^2)
    }
}

__global__ void running_norm_scan_kernel(const double* __restrict__ sums, int n_batch, long long max_update, double* __restrict__ state,
                                         float* __restrict__ mean_std /* [B][2]: mu_b, 1 / std_b */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double n = state[0], mu = state[1], s2 = state[2];
    for (int b = 0; b < n_batch; ++b) {
        if (n < (double)max_update) {
            const double m = sums[2 * b], q = sums[2 * b + 1];
            mu = (n == 0.0) ? m : mu + (m - mu) / n;
            const double v = fmax(q - 2.0 * mu * m + mu * mu, 0.0);     // mean((x - mu)^2)
            s2 = (n == 0.0) ? v : s2 + (v - s2) / n;
            n += 1.0;
        }
        mean_std[2 * b] = (float)mu;
        mean_std[2 * b + 1] = 1.0f / fmaxf((float)sqrt(s2), kF32Eps);
    }
    state[0] = n; state[1] = mu; state[2] = s2;
}

__global__ void __launch_bounds__(256) sample_norm_apply_kernel(const float* __restrict__ x, float* __restrict__ out, int elems,
                                                                const float* __restrict__ mean_std) {
    const int b = blockIdx.y;
    const float mu = mean_std[2 * b], inv = mean_std[2 * b + 1];
    const size_t base = (size_t)b * elems;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += gridDim.x * blockDim.x) out[base + i] = (__ldg(x + base + i) - mu) * inv;
}

// ------------------------------------------------------------------------------------------
// host: plan tables
// ------------------------------------------------------------------------------------------
static std::vector<float> linspace_f32(double start, double end, int steps) {
    // torch.linspace(dtype=float32): fp32 step, halves evaluated from each end with one rounding
    std::vector<float> v(steps);
    const float s32 = (float)start, e32 = (float)end;
    if (steps == 1) { v[0] = s32; return v; }
    const float step = (e32 - s32) / (float)(steps - 1);
    for (int i = 0; i < steps; ++i)
        v[i] = (i < steps / 2) ? (float)((double)s32 + (double)step * i) : (float)((double)e32 - (double)step * (steps - 1 - i));
    return v;
}

}  // namespace abt

using namespace abt;

#define ABT_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) return set_error(ABT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

extern "C" int abt_logmel_plan_create(const abt_mel_config* cfg, abt_logmel_plan** plan_out) {
    if (cfg == nullptr || plan_out == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (cfg->n_fft != kNfft) return set_error(ABT_ERR_ARG, "n_fft must be 1024 (got %d)", cfg->n_fft);
    if (cfg->n_mels < 1 || cfg->n_mels > kMaxMels) return set_error(ABT_ERR_ARG, "n_mels must be in [1, %d] (got %d)", kMaxMels, cfg->n_mels);
    const int n_mels = cfg->n_mels;
    if (cfg->win_length < 1 || cfg->win_length > kNfft) return set_error(ABT_ERR_ARG, "win_length must be in [1, 1024]");
    if (cfg->hop_length < 1 || cfg->hop_length > 1024) return set_error(ABT_ERR_ARG, "hop_length must be in [1, 1024]");
    if (cfg->apply_norm && !(cfg->norm_std > 0.f)) return set_error(ABT_ERR_ARG, "norm_std must be > 0");
    if (int rc = check_device_sm100()) return rc;

    // window: periodic Hann(win_length), centre-padded to n_fft (torch.stft)
    std::vector<float> win(kNfft, 0.f);
    const int left = (kNfft - cfg->win_length) / 2;
    for (int n = 0; n < cfg->win_length; ++n) win[left + n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / cfg->win_length));
    // twiddles [k2][n1]
    std::vector<float2> tw(1024);
    for (int k2 = 0; k2 < 32; ++k2)
        for (int n1 = 0; n1 < 32; ++n1) {
            const double ang = -2.0 * M_PI * (double)(n1 * k2) / 1024.0;
            tw[k2 * 32 + n1] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
    // HTK mel filterbank, norm=None (torchaudio/functional/functional.py:518-587), as CSR per band
    const int n_freqs = kNfft / 2 + 1;
    std::vector<float> all_freqs = linspace_f32(0.0, (double)(cfg->sample_rate / 2), n_freqs);
    const double m_min = 2595.0 * std::log10(1.0 + (double)cfg->f_min / 700.0);
    const double m_max = 2595.0 * std::log10(1.0 + (double)cfg->f_max / 700.0);
    std::vector<float> m_pts = linspace_f32(m_min, m_max, n_mels + 2);
    std::vector<float> f_pts(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) f_pts[i] = 700.0f * (std::pow(10.0f, m_pts[i] / 2595.0f) - 1.0f);
    std::vector<int> start(n_mels), len(n_mels), off(n_mels);
    std::vector<float> weights;
    int max_len = 0;
    for (int m = 0; m < n_mels; ++m) {
        const float fd0 = f_pts[m + 1] - f_pts[m], fd1 = f_pts[m + 2] - f_pts[m + 1];
        int first = -1, last = -1;
        std::vector<float> col(n_freqs);
        for (int k = 0; k < n_freqs; ++k) {
            const float down = -(f_pts[m] - all_freqs[k]) / fd0;
            const float up = (f_pts[m + 2] - all_freqs[k]) / fd1;
            const float v = std::fmax(0.0f, std::fmin(down, up));
            col[k] = v;
            if (v > 0.f) { if (first < 0) first = k; last = k; }
        }
        start[m] = first < 0 ? 0 : first;
        len[m] = first < 0 ? 0 : last - first + 1;
        // the kernel reads two weights at a time: bin k of band m lives at off[m] + (k - start[m]), and that index must have the parity of k
        if ((((int)weights.size() - start[m]) & 1) != 0) weights.push_back(0.f);
        off[m] = (int)weights.size();
        for (int k = 0; k < len[m]; ++k) weights.push_back(col[start[m] + k]);
        if (len[m] > max_len) max_len = len[m];
    }

    abt_logmel_plan* pl = static_cast<abt_logmel_plan*>(std::calloc(1, sizeof(abt_logmel_plan)));
    if (pl == nullptr) return set_error(ABT_ERR_ARG, "out of host memory");
    pl->cfg = *cfg;
    pl->nnz = (int)weights.size();
    pl->max_len = max_len;
    ABT_CUDA_OK(cudaMalloc(&pl->d_window, sizeof(float) * kNfft));
    ABT_CUDA_OK(cudaMalloc(&pl->d_twiddle, sizeof(float2) * 1024));
    ABT_CUDA_OK(cudaMalloc(&pl->d_mel_start, sizeof(int) * n_mels));
    ABT_CUDA_OK(cudaMalloc(&pl->d_mel_len, sizeof(int) * n_mels));
    ABT_CUDA_OK(cudaMalloc(&pl->d_mel_off, sizeof(int) * n_mels));
    ABT_CUDA_OK(cudaMalloc(&pl->d_mel_w, sizeof(float) * (weights.empty() ? 1 : weights.size())));
    ABT_CUDA_OK(cudaMemcpy(pl->d_window, win.data(), sizeof(float) * kNfft, cudaMemcpyHostToDevice));
    ABT_CUDA_OK(cudaMemcpy(pl->d_twiddle, tw.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice));
    ABT_CUDA_OK(cudaMemcpy(pl->d_mel_start, start.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    ABT_CUDA_OK(cudaMemcpy(pl->d_mel_len, len.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    ABT_CUDA_OK(cudaMemcpy(pl->d_mel_off, off.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    if (!weights.empty()) ABT_CUDA_OK(cudaMemcpy(pl->d_mel_w, weights.data(), sizeof(float) * weights.size(), cudaMemcpyHostToDevice));
    *plan_out = pl;
    return 0;
}

extern "C" int abt_logmel_plan_destroy(abt_logmel_plan* pl) {
    if (pl == nullptr) return 0;
    cudaFree(pl->d_window); cudaFree(pl->d_twiddle); cudaFree(pl->d_mel_start);
    cudaFree(pl->d_mel_len); cudaFree(pl->d_mel_off); cudaFree(pl->d_mel_w);
    std::free(pl);
    return 0;
}

static int launch_logmel(const abt_logmel_plan* pl, const float* wav, int64_t row_stride, const int32_t* wav_offset, const int32_t* wav_origin, int n_clips, int n_samples,
                         const int32_t* frame_start, int n_frames_out,
                         float* out_base, const int32_t* out_slot, int64_t out_slot_stride, cudaStream_t stream) {
    if (n_clips < 0 || n_frames_out < 0) return set_error(ABT_ERR_ARG, "negative size");
    if (n_clips == 0 || n_frames_out == 0) return 0;     // empty batch: nothing to do (pointers may be null)
    if (pl == nullptr || wav == nullptr || out_base == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_samples <= kNfft / 2) return set_error(ABT_ERR_ARG, "n_samples must exceed n_fft/2 = 512 for reflect padding (got %d)", n_samples);
    if (n_clips > 65535) return set_error(ABT_ERR_ARG, "n_clips must be <= 65535 per call");
    if (int rc = check_device_sm100()) return rc;
    LogmelArgs a{};
    a.wav = wav; a.row_stride = row_stride; a.wav_offset = wav_offset; a.wav_origin = wav_origin; a.n_samples = n_samples; a.hop = pl->cfg.hop_length;
    a.n_frames_total = 1 + n_samples / pl->cfg.hop_length;
    a.frame_start = frame_start; a.n_frames_out = n_frames_out;
    a.out_base = out_base; a.out_slot = out_slot; a.out_slot_stride = out_slot_stride;
    a.apply_norm = pl->cfg.apply_norm; a.norm_mean = pl->cfg.norm_mean;
    a.inv_std = pl->cfg.apply_norm ? 1.0f / pl->cfg.norm_std : 1.0f;
    a.pad_value = pl->cfg.apply_norm ? (0.0f - pl->cfg.norm_mean) / pl->cfg.norm_std : 0.0f;
    a.window = pl->d_window; a.twiddle = pl->d_twiddle;
    a.mel_start = pl->d_mel_start; a.mel_len = pl->d_mel_len; a.mel_off = pl->d_mel_off; a.mel_w = pl->d_mel_w; a.nnz = pl->nnz; a.n_mels = pl->cfg.n_mels;
    const int span = (kTileFrames - 1) * a.hop + kNfft;
    const size_t smem = sizeof(float) * (((span + 31) & ~31) + kWarps * kScratchFloats + pl->cfg.n_mels * (kTileFrames + 1) + pl->nnz) + sizeof(int) * 3 * pl->cfg.n_mels;
    if (smem > 227 * 1024) return set_error(ABT_ERR_ARG, "hop_length %d with %d mel bands needs %zu bytes of shared memory (> 227 KiB)", a.hop, a.n_mels, smem);
    static size_t smem_set[64] = {};                  // cudaFuncSetAttribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > smem_set[dev]) {
        ABT_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) smem_set[dev] = smem;
    }
    dim3 grid((n_frames_out + kTileFrames - 1) / kTileFrames, n_clips);
    logmel_kernel<<<grid, kWarps * 32, smem, stream>>>(a);
    count_launch();
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int abt_logmel_fwd(const abt_logmel_plan* plan, const float* wav, int n_clips, int n_samples, float* out, abt_stream_t stream) {
    if (plan == nullptr) return set_error(ABT_ERR_ARG, "plan is null");
    const int n_frames = 1 + n_samples / plan->cfg.hop_length;
    return launch_logmel(plan, wav, n_samples, nullptr, nullptr, n_clips, n_samples, nullptr, n_frames, out, nullptr, (int64_t)plan->cfg.n_mels * n_frames,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int abt_logmel_crop_fwd(const abt_logmel_plan* plan, const float* wav, int64_t wav_row_stride, const int32_t* wav_offset, int n_clips,
                                   int n_samples, const int32_t* frame_start, int n_frames, float* out_base, const int32_t* out_slot,
                                   int64_t out_slot_stride, abt_stream_t stream) {
    if (plan == nullptr) return set_error(ABT_ERR_ARG, "plan is null");
    if (out_slot_stride < (int64_t)plan->cfg.n_mels * n_frames) return set_error(ABT_ERR_ARG, "out_slot_stride smaller than one clip");
    if (wav_row_stride < n_samples) return set_error(ABT_ERR_ARG, "wav_row_stride smaller than n_samples");
    return launch_logmel(plan, wav, wav_row_stride, wav_offset, nullptr, n_clips, n_samples, frame_start, n_frames, out_base, out_slot, out_slot_stride,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int abt_wav_span_len(const abt_logmel_plan* plan, int n_frames, int* span_len) {
    if (plan == nullptr || span_len == nullptr || n_frames < 1) return set_error(ABT_ERR_ARG, "bad argument");
    *span_len = ((n_frames - 1) * plan->cfg.hop_length + kNfft + 4 + 3) & ~3;     // +4: the source start is rounded down to 16 bytes
    return 0;
}

extern "C" int abt_wav_span_gather(const abt_logmel_plan* plan, const float* wav, int wav_on_host, int64_t wav_row_stride, int n_clips,
                                   int n_samples, const int32_t* frame_start, int n_frames, float* spans, int32_t* span_origin,
                                   abt_stream_t stream) {
    if (n_clips == 0) return 0;
    if (plan == nullptr || wav == nullptr || spans == nullptr || span_origin == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_clips < 0 || n_clips > 65535 || n_frames < 1) return set_error(ABT_ERR_ARG, "bad shape");
    if ((wav_row_stride & 3) != 0 || (reinterpret_cast<uintptr_t>(wav) & 15) != 0) return set_error(ABT_ERR_ARG, "waveform rows must be 16-byte aligned");
    if (int rc = check_device_sm100()) return rc;
    int span_len = 0;
    abt_wav_span_len(plan, n_frames, &span_len);
    if (n_samples < span_len) return set_error(ABT_ERR_ARG, "clips of %d samples are shorter than the %d-sample span; use abt_logmel_crop_fwd", n_samples, span_len);
    const float* src = wav;
    if (wav_on_host) {
        void* dptr = nullptr;
        cudaError_t e = cudaHostGetDevicePointer(&dptr, const_cast<float*>(wav), 0);
        if (e != cudaSuccess) return set_error(ABT_ERR_ARG, "host waveforms must be in mapped pinned memory (cudaHostAlloc / cudaHostRegister): %s", cudaGetErrorString(e));
        src = static_cast<const float*>(dptr);
    }
    const int blocks = (n_clips + 31) / 32 < g_gather_blocks ? (n_clips + 31) / 32 : g_gather_blocks;
    wav_span_gather_kernel<<<blocks, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, wav_row_stride, n_samples, n_clips, frame_start,
                                                                                       plan->cfg.hop_length, kNfft / 2, span_len, spans, span_origin);
    count_launch();
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int abt_logmel_span_fwd(const abt_logmel_plan* plan, const float* spans, const int32_t* span_origin, int n_clips, int n_samples,
                                   const int32_t* frame_start, int n_frames, float* out_base, const int32_t* out_slot, int64_t out_slot_stride,
                                   abt_stream_t stream) {
    if (plan == nullptr || span_origin == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (out_slot_stride < (int64_t)plan->cfg.n_mels * n_frames) return set_error(ABT_ERR_ARG, "out_slot_stride smaller than one clip");
    int span_len = 0;
    if (int rc = abt_wav_span_len(plan, n_frames, &span_len)) return rc;
    return launch_logmel(plan, spans, span_len, nullptr, span_origin, n_clips, n_samples, frame_start, n_frames, out_base, out_slot, out_slot_stride,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int abt_lms_crop_norm(const float* lms, int n_clips, int n_mels, int t_full, const int32_t* frame_start, int n_frames, int apply_norm,
                                 float norm_mean, float norm_std, float* out_base, const int32_t* out_slot, int64_t out_slot_stride,
                                 abt_stream_t stream) {
    if (n_clips == 0) return 0;
    if (lms == nullptr || out_base == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_clips < 0 || n_clips > 65535 || n_mels < 1 || t_full < 1 || n_frames < 1) return set_error(ABT_ERR_ARG, "bad shape");
    if (apply_norm && !(norm_std > 0.f)) return set_error(ABT_ERR_ARG, "norm_std must be > 0");
    if (int rc = check_device_sm100()) return rc;
    dim3 grid((n_mels * n_frames + 255) / 256, n_clips);
    lms_crop_norm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(lms, n_mels, t_full, frame_start, n_frames, apply_norm, norm_mean,
                                                                                    apply_norm ? 1.0f / norm_std : 1.0f, out_base, out_slot,
                                                                                    out_slot_stride);
    count_launch();
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int abt_views_fwd(const abt_views_args* a, abt_stream_t stream) {
    if (a == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (a->n_clips == 0 || a->n_views == 0) return 0;
    if (a->x == nullptr || a->params == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (a->n_clips < 0 || a->n_views < 0 || a->n_views > 8) return set_error(ABT_ERR_ARG, "bad n_clips / n_views (at most 8 views per launch)");
    if (a->param_stride < 0 || a->view_offset < 0) return set_error(ABT_ERR_ARG, "bad param_stride / view_offset");
    if (a->noise != nullptr && a->noise_views < a->view_offset + a->n_views)
        return set_error(ABT_ERR_ARG, "noise has %d view planes, this launch reads planes up to %d", a->noise_views, a->view_offset + a->n_views);
    if (a->in_h < 1 || a->in_w < 1 || (a->in_h * a->in_w) % 4 != 0) return set_error(ABT_ERR_ARG, "in_h * in_w must be a positive multiple of 4");
    if (a->canvas_h < a->in_h || a->canvas_w < a->in_w) return set_error(ABT_ERR_ARG, "canvas smaller than input");
    if (a->out_h < 1 || a->out_w < 1 || a->out_h + a->out_w > 256) return set_error(ABT_ERR_ARG, "out_h + out_w must be in [2, 256]");
    if ((a->x_slot_stride % 4) != 0 || (a->bank_slot_stride % 4) != 0) return set_error(ABT_ERR_ARG, "slot strides must be multiples of 4 floats");
    if (int rc = check_device_sm100()) return rc;
    if (a->out_w > kViewThreads) return set_error(ABT_ERR_ARG, "out_w must be <= %d", kViewThreads);
    const size_t smem = sizeof(float) * ((size_t)a->canvas_h * a->canvas_w + (size_t)a->canvas_h * a->out_w) + 32 * (size_t)(a->out_w + a->out_h);
    if (smem > 200 * 1024) return set_error(ABT_ERR_ARG, "canvas too large for shared memory");
    static size_t smem_set[64] = {};                  // per device; 0 = only the default 48 KiB is available
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || smem > smem_set[dev])) {
        ABT_CUDA_OK(cudaFuncSetAttribute(views_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) smem_set[dev] = smem;
    }
    dim3 grid(a->n_clips, a->n_views);
    views_kernel<<<grid, kViewThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
    count_launch();
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int abt_bank_push(const float* x, int64_t x_stride, int n_clips, int clip_elems, float* bank, int64_t bank_slot_stride,
                             const int32_t* slot, abt_stream_t stream) {
    if (n_clips == 0) return 0;
    if (x == nullptr || bank == nullptr || slot == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_clips < 0 || n_clips > 65535 || clip_elems < 4 || clip_elems % 4 != 0 || x_stride % 4 != 0 || bank_slot_stride % 4 != 0)
        return set_error(ABT_ERR_ARG, "bad shape (clip_elems and strides must be multiples of 4)");
    if (int rc = check_device_sm100()) return rc;
    dim3 grid((clip_elems / 4 + 255) / 256, n_clips);
    bank_push_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_stride, clip_elems, bank, bank_slot_stride, slot);
    count_launch();
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int abt_normalize_batch_workspace_bytes(int n_channels, size_t* bytes) {
    if (bytes == nullptr || n_channels < 1) return set_error(ABT_ERR_ARG, "bad argument");
    *bytes = sizeof(double) * 2 * (size_t)kNbBlocks * n_channels;
    return 0;
}

extern "C" int abt_normalize_batch(const float* x, int n_batch, int n_channels, int hw, float* out, void* workspace, abt_stream_t stream) {
    if (n_batch == 0) return 0;
    if (x == nullptr || out == nullptr || workspace == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_batch < 0 || n_channels < 1 || n_channels > 65535 || hw < 1) return set_error(ABT_ERR_ARG, "bad shape");
    if (int rc = check_device_sm100()) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(kNbBlocks, n_channels);
    batch_stats_kernel<<<grid, 256, 0, st>>>(x, n_batch, n_channels, hw, static_cast<double*>(workspace));
    const long long per_ch = (long long)n_batch * hw;
    const unsigned blocks = (unsigned)((per_ch + 1023) / 1024 < 148 * 8 ? (per_ch + 1023) / 1024 : 148 * 8);
    batch_norm_apply_kernel<<<dim3(blocks, n_channels), 256, 0, st>>>(x, out, n_batch, n_channels, hw, static_cast<const double*>(workspace), kNbBlocks);
    count_launch(2);
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

// mean and UNBIASED standard deviation of a whole tensor (dataset statistics, datasets.py:362-376: `lms_vectors.mean()`,
// `lms_vectors.std()`): the per-channel partial sums above with one channel, folded by one block into two doubles
__global__ void __launch_bounds__(256) mean_std_finalize_kernel(const float* __restrict__ x, const double* __restrict__ partials, int n_partials, double n,
                                                                double* __restrict__ out2) {
    __shared__ double red[2][8];
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) { s1 += partials[(size_t)i * 2]; s2 += partials[(size_t)i * 2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
        const double dmean = t1 / n;
        const double var = n > 1.0 ? fmax((t2 - t1 * dmean) / (n - 1.0), 0.0) : 0.0;
        out2[0] = (double)__ldg(x) + dmean;
        out2[1] = sqrt(var);
    }
}

extern "C" int abt_mean_std(const float* x, int n_batch, int elems, double* out2, void* workspace, abt_stream_t stream) {
    if (x == nullptr || out2 == nullptr || workspace == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_batch < 1 || elems < 1) return set_error(ABT_ERR_ARG, "bad shape");
    if (int rc = check_device_sm100()) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    batch_stats_kernel<<<dim3(kNbBlocks, 1), 256, 0, st>>>(x, n_batch, 1, elems, static_cast<double*>(workspace));
    mean_std_finalize_kernel<<<1, 256, 0, st>>>(x, static_cast<const double*>(workspace), kNbBlocks, (double)n_batch * (double)elems, out2);
    count_launch(2);
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int abt_running_norm_workspace_bytes(int n_batch, size_t* bytes) {
    if (bytes == nullptr || n_batch < 0) return set_error(ABT_ERR_ARG, "bad argument");
    *bytes = (size_t)n_batch * (2 * sizeof(double) + 2 * sizeof(float)) + 256;
    return 0;
}

extern "C" int abt_running_norm(const float* x, int n_batch, int elems, long long max_update, double* state3, float* out, void* workspace,
                                abt_stream_t stream) {
    if (n_batch == 0) return 0;
    if (x == nullptr || out == nullptr || state3 == nullptr || workspace == nullptr) return set_error(ABT_ERR_ARG, "null argument");
    if (n_batch < 0 || n_batch > 65535 || elems < 1) return set_error(ABT_ERR_ARG, "bad shape");
    if (int rc = check_device_sm100()) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double* sums = static_cast<double*>(workspace);
    float* mean_std = reinterpret_cast<float*>(sums + 2 * (size_t)n_batch);
    sample_sums_kernel<<<n_batch, 256, 0, st>>>(x, elems, sums);
    running_norm_scan_kernel<<<1, 32, 0, st>>>(sums, n_batch, max_update, state3, mean_std);
    const unsigned bx = (unsigned)((elems + 1023) / 1024 < 64 ? (elems + 1023) / 1024 : 64);
    sample_norm_apply_kernel<<<dim3(bx, n_batch), 256, 0, st>>>(x, out, elems, mean_std);
    count_launch(3);
    ABT_CUDA_OK(cudaGetLastError());
    return 0;
}
