// Step-adjacent optimiser math as multi-tensor kernels.
//
// Replaces the per-parameter Python loops of the reference:
//   * LARS.step            utils/utils.py:162-189  -- per tensor: dp = g + wd p;  q = eta |p| / |dp| (1 when either norm is 0);
//                                                     mu = momentum mu + q dp;  p -= lr mu          (2 torch.norm + ~8 launches per tensor)
//   * update_moving_average utils/utils.py:328-331 -- ma = beta ma + (1 - beta) p                    (BYOL target network, per tensor)
// All tensors of a step are described by one device table (pointer, element count, flags) and a chunk map (chunk -> tensor, offset),
// so a step costs three launches (chunk norms, per-tensor trust ratios, apply) regardless of the number of parameters.  fp32 parameters only (the reference keeps
// fp32 master weights under autocast).  Norm partials are reduced in a fixed order: results are deterministic.
#include "abt_internal.h"

#include <stdint.h>

namespace abt {

constexpr int kOptChunk = 8192;      // elements per block
constexpr int kOptThreads = 256;

struct TensorRec {                   // mirrors abt_opt_tensor (include/abt_b200.h)
    float* p; const float* g; float* aux; long long n; int flags; int chunk0;      // chunk0: first chunk of this tensor in the chunk map
};

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < kOptThreads / 32) t = red[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;          // valid in thread 0
}

// partial[chunk] = (sum p^2, sum (g + wd p)^2) over the chunk's elements
__global__ void __launch_bounds__(kOptThreads) lars_norm_kernel(const TensorRec* __restrict__ tensors, const int* __restrict__ chunk_tensor, int n_chunks,
                                                                float weight_decay, float2* __restrict__ partial) {
    __shared__ float red[kOptThreads / 32];
    const int c = blockIdx.x;
    const TensorRec t = tensors[chunk_tensor[c]];
    const long long off = (long long)(c - t.chunk0) * kOptChunk;
    const int n = (int)((t.n - off) < kOptChunk ? (t.n - off) : kOptChunk);
    const float wd = (t.flags & 1) ? weight_decay : 0.f;
    const float* p = t.p + off;
    const float* g = t.g + off;
    float sp = 0.f, sd = 0.f;
    if ((t.flags & 2) != 0) {
        const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
        if (vec) {
            for (int i = threadIdx.x * 4; i + 3 < n; i += kOptThreads * 4) {
                const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(g + i);
                const float d0 = fmaf(wd, a.x, b.x), d1 = fmaf(wd, a.y, b.y), d2 = fmaf(wd, a.z, b.z), d3 = fmaf(wd, a.w, b.w);
                sp += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
                sd += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
            for (int i = (n & ~3) + threadIdx.x; i < n; i += kOptThreads) { const float a = p[i], d = fmaf(wd, a, g[i]); sp += a * a; sd += d * d; }
        } else {
            for (int i = threadIdx.x; i < n; i += kOptThreads) { const float a = p[i], d = fmaf(wd, a, g[i]); sp += a * a; sd += d * d; }
        }
    }
    const float tp = block_sum(sp, red), td = block_sum(sd, red);
    if (threadIdx.x == 0) partial[c] = make_float2(tp, td);
}

// q[t] = eta |p_t| / |dp_t| (1 when tensor t is excluded from the adaptation or either norm is zero): one block per tensor sums the
// tensor's chunk partials -- fixed assignment, fixed tree, double: deterministic.  (Summing them in every block of the apply kernel
// made an 8192 x 8192 weight cost 8192 serial additions in each of its 8192 blocks: 7 ms per step for a 156 M-parameter model.)
__global__ void __launch_bounds__(kOptThreads) lars_ratio_kernel(const TensorRec* __restrict__ tensors, float eta, const float2* __restrict__ partial,
                                                                 float* __restrict__ q_out) {
    __shared__ double sp_s[kOptThreads], sd_s[kOptThreads];
    const TensorRec t = tensors[blockIdx.x];
    double sp = 0.0, sd = 0.0;
    if (t.flags & 2) {
        const int nc = (int)((t.n + kOptChunk - 1) / kOptChunk);
        for (int k = threadIdx.x; k < nc; k += kOptThreads) { const float2 v = partial[t.chunk0 + k]; sp += (double)v.x; sd += (double)v.y; }
    }
    sp_s[threadIdx.x] = sp; sd_s[threadIdx.x] = sd;
    __syncthreads();
    for (int o = kOptThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sp_s[threadIdx.x] += sp_s[threadIdx.x + o]; sd_s[threadIdx.x] += sd_s[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float q = 1.f;
        if (t.flags & 2) {
            const float pn = (float)sqrt(sp_s[0]), un = (float)sqrt(sd_s[0]);
            q = (pn > 0.f && un > 0.f) ? eta * pn / un : 1.f;
        }
        q_out[blockIdx.x] = q;
    }
}

__global__ void __launch_bounds__(kOptThreads) lars_apply_kernel(const TensorRec* __restrict__ tensors, const int* __restrict__ chunk_tensor, int n_chunks,
                                                                 float lr, float weight_decay, float momentum, const float* __restrict__ q_arr) {
    const int c = blockIdx.x;
    const int ti = chunk_tensor[c];
    const TensorRec t = tensors[ti];
    const long long off = (long long)(c - t.chunk0) * kOptChunk;
    const int n = (int)((t.n - off) < kOptChunk ? (t.n - off) : kOptChunk);
    const float wd = (t.flags & 1) ? weight_decay : 0.f;
    const float q = q_arr[ti];
    float* p = t.p + off;
    const float* g = t.g + off;
    float* mu = t.aux + off;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mu)) & 15) == 0;
    int done = 0;
    if (vec) {
        for (int i = threadIdx.x * 4; i + 3 < n; i += kOptThreads * 4) {
            const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(g + i), m0 = *reinterpret_cast<const float4*>(mu + i);
            float4 m, o;
            m.x = fmaf(momentum, m0.x, fmaf(wd, a.x, b.x) * q); m.y = fmaf(momentum, m0.y, fmaf(wd, a.y, b.y) * q);
            m.z = fmaf(momentum, m0.z, fmaf(wd, a.z, b.z) * q); m.w = fmaf(momentum, m0.w, fmaf(wd, a.w, b.w) * q);
            o.x = fmaf(-lr, m.x, a.x); o.y = fmaf(-lr, m.y, a.y); o.z = fmaf(-lr, m.z, a.z); o.w = fmaf(-lr, m.w, a.w);
            *reinterpret_cast<float4*>(mu + i) = m;
            *reinterpret_cast<float4*>(p + i) = o;
        }
        done = n & ~3;
    }
    for (int i = done + threadIdx.x; i < n; i += kOptThreads) {
        const float a = p[i];
        const float dp = fmaf(wd, a, g[i]) * q;
        const float m = fmaf(momentum, mu[i], dp);
        mu[i] = m;
        p[i] = fmaf(-lr, m, a);
    }
}

// ma = beta ma + (1 - beta) cur      (rec.p = ma, rec.g = cur)
__global__ void __launch_bounds__(kOptThreads) ema_kernel(const TensorRec* __restrict__ tensors, const int* __restrict__ chunk_tensor, int n_chunks,
                                                          float beta) {
    const int c = blockIdx.x;
    const TensorRec t = tensors[chunk_tensor[c]];
    const long long off = (long long)(c - t.chunk0) * kOptChunk;
    const int n = (int)((t.n - off) < kOptChunk ? (t.n - off) : kOptChunk);
    float* ma = t.p + off;
    const float* cur = t.g + off;
    const float ob = 1.f - beta;
    for (int i = threadIdx.x; i < n; i += kOptThreads) ma[i] = ma[i] * beta + ob * cur[i];     // the reference's evaluation order: old * beta + (1 - beta) * new
}

static int check_table(const void* tensors_dev, const int* chunk_tensor_dev, int n_tensors, int n_chunks) {
    if (n_tensors <= 0 || n_chunks <= 0) return set_error(ABT_ERR_ARG, "empty tensor table");
    if (tensors_dev == nullptr || chunk_tensor_dev == nullptr) return set_error(ABT_ERR_ARG, "null table pointer");
    return check_device_sm100();
}

}  // namespace abt

using namespace abt;

static_assert(sizeof(abt_opt_tensor) == sizeof(TensorRec), "abt_opt_tensor layout");

extern "C" int abt_opt_chunk_elems(void) { return kOptChunk; }

extern "C" int abt_lars_step(const abt_opt_tensor* tensors_dev, const int* chunk_tensor_dev, int n_tensors, int n_chunks, float lr, float weight_decay,
                             float momentum, float eta, void* partial_dev, abt_stream_t stream) {
    if (int rc = check_table(tensors_dev, chunk_tensor_dev, n_tensors, n_chunks)) return rc;
    if (partial_dev == nullptr) return set_error(ABT_ERR_ARG, "partial buffer is null (8 bytes per chunk + 4 bytes per tensor)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const TensorRec* T = reinterpret_cast<const TensorRec*>(tensors_dev);
    float* q_arr = static_cast<float*>(partial_dev) + 2 * (size_t)n_chunks;       // per-tensor trust ratios, behind the chunk partials
    lars_norm_kernel<<<n_chunks, kOptThreads, 0, st>>>(T, chunk_tensor_dev, n_chunks, weight_decay, static_cast<float2*>(partial_dev));
    lars_ratio_kernel<<<n_tensors, kOptThreads, 0, st>>>(T, eta, static_cast<const float2*>(partial_dev), q_arr);
    lars_apply_kernel<<<n_chunks, kOptThreads, 0, st>>>(T, chunk_tensor_dev, n_chunks, lr, weight_decay, momentum, q_arr);
    count_launch(3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "LARS launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int abt_ema_update(const abt_opt_tensor* tensors_dev, const int* chunk_tensor_dev, int n_tensors, int n_chunks, float beta, abt_stream_t stream) {
    if (int rc = check_table(tensors_dev, chunk_tensor_dev, n_tensors, n_chunks)) return rc;
    ema_kernel<<<n_chunks, kOptThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const TensorRec*>(tensors_dev), chunk_tensor_dev, n_chunks, beta);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "EMA launch: %s", cudaGetErrorString(e));
    return 0;
}
