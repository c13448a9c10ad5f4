// Small fp32 kernels behind the per-op (autograd) surface of the drop-in modules: LangPuller.forward /
// get_projected_weight / loss1, ResNet.regloss / reglossnovel, LinearMap and the classifier nn.Linear
// (reference models/resnet_language.py:12-18, 74-97, 187, 229-240).  The fused session driver does not use these;
// they exist so the UNMODIFIED reference loop can run on top of the B200 modules one launch per op.
#include <algorithm>
#include "common.h"

namespace {
using namespace srb;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[i, :] = softmax_j(<En[i], Eb[j]> / temp) @ W0          one CTA per novel label
__global__ void __launch_bounds__(256) semantic_pullers_kernel(const float* __restrict__ En, const float* __restrict__ Eb,
                                                               const float* __restrict__ W0, int n_base, int e, int d,
                                                               float temp, int mask_diag, float* __restrict__ out) {
    extern __shared__ float sc[];  // [n_base]
    const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = warp; j < n_base; j += nw) {
        float s = 0.f;
        for (int k = lane; k < e; k += 32) s = fmaf(En[(int64_t)i * e + k], Eb[(int64_t)j * e + k], s);
        s = warp_sum_f(s);
        if (lane == 0) sc[j] = (mask_diag && j == i) ? -9999.f / temp : s / temp;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = 0; j < n_base; ++j) mx = fmaxf(mx, sc[j]);
    float den = 0.f;
    for (int j = 0; j < n_base; ++j) den += expf(sc[j] - mx);
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < n_base; ++j) acc = fmaf(expf(sc[j] - mx) / den, W0[(int64_t)j * d + k], acc);
        out[(int64_t)i * d + k] = acc;
    }
}

// y[i, o] = <x[i], w[o]> + b[o]            one warp per output element
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                  int n, int k, int m, float* __restrict__ y) {
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (int64_t)n * m) return;
    const int i = (int)(gw / m), o = (int)(gw % m);
    float s = 0.f;
    for (int t = lane; t < k; t += 32) s = fmaf(x[(int64_t)i * k + t], w[(int64_t)o * k + t], s);
    s = warp_sum_f(s);
    if (lane == 0) y[gw] = s + (b ? b[o] : 0.f);
}

// dw[o, t] = sum_i dy[i, o] * x[i, t];  db[o] = sum_i dy[i, o]         one thread per dw element
__global__ void linear_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int n, int k, int m,
                                  float* __restrict__ dw, float* __restrict__ db) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)m * k) {
        const int o = (int)(idx / k), t = (int)(idx % k);
        float s = 0.f;
        for (int i = 0; i < n; ++i) s = fmaf(dy[(int64_t)i * m + o], x[(int64_t)i * k + t], s);
        dw[idx] = s;
    }
    if (db != nullptr && idx < m) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += dy[(int64_t)i * m + (int)idx];
        db[idx] = s;
    }
}

// out[0] = sum (a - b)^2  (fp64 accumulation, one CTA)
__global__ void __launch_bounds__(1024) sqdist_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                      float* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float dlt = a[i] - b[i];
        s += (double)dlt * (double)dlt;
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        out[0] = (float)t;
    }
}

// out = (a - b) * coef, coef = scale * gout[0] * (sq ? 1 / sqrt(sq[0]) (0 when sq[0] == 0) : 1)
__global__ void diff_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float scale,
                                  const float* __restrict__ gout, const float* __restrict__ sq, float* __restrict__ out) {
    float coef = scale * (gout ? gout[0] : 1.f);
    if (sq != nullptr) {
        const float nrm = sqrtf(sq[0]);
        coef = nrm > 0.f ? coef / nrm : 0.f;
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (a[i] - b[i]) * coef;
}

// out[i, :] = (x[i] Q^T) Q       one CTA per row; Q = qt [q, d] with orthonormal rows
__global__ void __launch_bounds__(256) project_rows_kernel(const float* __restrict__ x, const float* __restrict__ Q, int q,
                                                           int d, float* __restrict__ out) {
    extern __shared__ float sm[];  // x row [d] then u [q]
    float* sx = sm;
    float* su = sm + d;
    const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < d; k += blockDim.x) sx[k] = x[(int64_t)i * d + k];
    __syncthreads();
    for (int j = warp; j < q; j += nw) {
        float s = 0.f;
        for (int k = lane; k < d; k += 32) s = fmaf(Q[(int64_t)j * d + k], sx[k], s);
        s = warp_sum_f(s);
        if (lane == 0) su[j] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < q; ++j) s = fmaf(su[j], Q[(int64_t)j * d + k], s);
        out[(int64_t)i * d + k] = s;
    }
}
// MSELoss(mean) forward + gradient: loss[0] += sum((y-t)^2)/n (caller zeroes it); dy = 2 (y - t) / n
__global__ void mse_grad_kernel(const float* __restrict__ y, const float* __restrict__ t, int64_t n, float* __restrict__ dy,
                                float* __restrict__ loss) {
    __shared__ double red[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (i < n) {
        const float dlt = y[i] - t[i];
        dy[i] = 2.f * dlt / (float)n;
        s = (double)dlt * (double)dlt;
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += red[k];
        atomicAdd(loss, (float)(tot / (double)n));
    }
}

// torch.optim.SGD without momentum: p -= lr * (g + wd * p)
__global__ void sgd_update_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float lr, float wd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] - lr * fmaf(wd, p[i], g[i]);
}
}  // namespace

extern "C" int32_t sr_mse_grad(const float* y, const float* target, int64_t n, float* dy, float* loss, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!y || !target || !dy || !loss || n < 1) return fail(SR_E_ARG, "sr_mse_grad: bad arguments");
    mse_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(y, target, n, dy, loss);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_sgd_update(float* param, const float* grad, int64_t n, float lr, float weight_decay, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!param || !grad || n < 1) return fail(SR_E_ARG, "sr_sgd_update: bad arguments");
    sgd_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(param, grad, n, lr, weight_decay);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_semantic_pullers(const float* novel_embeds, const float* base_embeds, const float* base_weight,
                                       int32_t n_novel, int32_t n_base, int32_t embed_dim, int32_t dim, float temperature,
                                       int32_t mask_diagonal, float* pullers, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!novel_embeds || !base_embeds || !base_weight || !pullers || n_novel < 1 || n_base < 1 || embed_dim < 1 || dim < 1)
        return fail(SR_E_ARG, "sr_semantic_pullers: bad arguments");
    if (n_base * sizeof(float) > 48 * 1024) return fail(SR_E_ARG, "sr_semantic_pullers: n_base too large");
    semantic_pullers_kernel<<<n_novel, 256, n_base * sizeof(float), stream>>>(novel_embeds, base_embeds, base_weight, n_base,
                                                                              embed_dim, dim, temperature, mask_diagonal,
                                                                              pullers);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_linear_fwd(const float* x, const float* w, const float* bias, int32_t n, int32_t k, int32_t m, float* y,
                                 void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x || !w || !y || n < 1 || k < 1 || m < 1) return fail(SR_E_ARG, "sr_linear_fwd: bad arguments");
    const int64_t threads = (int64_t)n * m * 32;
    linear_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(x, w, bias, n, k, m, y);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_linear_bwd(const float* dy, const float* x, int32_t n, int32_t k, int32_t m, float* dw, float* dbias,
                                 void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!dy || !x || !dw || n < 1 || k < 1 || m < 1) return fail(SR_E_ARG, "sr_linear_bwd: bad arguments");
    const int64_t total = std::max<int64_t>((int64_t)m * k, m);
    linear_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dy, x, n, k, m, dw, dbias);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_sqdist(const float* a, const float* b, int64_t n, float* out, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!a || !b || !out || n < 1) return fail(SR_E_ARG, "sr_sqdist: bad arguments");
    sqdist_kernel<<<1, 1024, 0, stream>>>(a, b, n, out);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_diff_scale(const float* a, const float* b, int64_t n, float scale, const float* gout, const float* sq,
                                 float* out, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!a || !b || !out || n < 1) return fail(SR_E_ARG, "sr_diff_scale: bad arguments");
    diff_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, b, n, scale, gout, sq, out);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}

extern "C" int32_t sr_project_rows(const float* x, const float* qt, int32_t n, int32_t q_rows, int32_t dim, float* out,
                                   void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x || !qt || !out || n < 1 || q_rows < 1 || dim < 1) return fail(SR_E_ARG, "sr_project_rows: bad arguments");
    const size_t smem = (size_t)(dim + q_rows) * sizeof(float);
    if (smem > 48 * 1024) return fail(SR_E_ARG, "sr_project_rows: dim + q_rows too large");
    project_rows_kernel<<<n, 256, smem, stream>>>(x, qt, q_rows, dim, out);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}
