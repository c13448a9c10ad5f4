// learn_mapping.py:41-67 as ONE launch: LinearMap(e, d) fitted to the base classifier rows by `epochs` full-batch SGD steps
// on nn.MSELoss (lr, weight_decay, no momentum).
//
// Output dimension m of a linear map depends on nothing but its own weight row: y[:, m] = X w_m + b_m, dw_m = dy[:, m]^T X.
// So the fit splits over the output dimensions with NO exchange between CTAs: a CTA keeps X (n x e, <= 184 KB), its rows of
// W, its columns of the targets and its dy in shared memory and runs every step locally - no grid barrier, no global traffic
// inside the loop except one partial loss per step.  (The per-op route, sr_linear_fwd / sr_mse_grad / sr_linear_bwd /
// sr_sgd_update, is five launches per step sequenced from Python: ~25 us per step, host-bound.)
#include <algorithm>
#include <cstdint>
#include "common.h"
#include "head_common.cuh"

namespace {
using namespace srb;
constexpr int kT = 256;

struct FitParams {
    const float* x;        // [n][e]
    const float* target;   // [n][d]
    float* w;              // [d][e]
    float* b;              // [d]
    int n, e, d, epochs, M;   // M: output dimensions per CTA
    float lr, wd;
    double* loss_part;     // [epochs][gridDim.x]: sum of squared errors of this CTA's columns
};

__global__ void __launch_bounds__(kT, 1) fit_linear_map_kernel(const FitParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ double red[32];
    const int n = p.n, e = p.e, M = p.M;
    const int ep = e + 1;                 // row pitch of X: walking down a column is then bank-conflict free
    float* Xs = sm;                       // [n][ep]
    float* Ws = Xs + (size_t)n * ep;      // [M][e]
    float* Ts = Ws + (size_t)M * e;       // [n][M]
    float* Dy = Ts + (size_t)n * M;       // [n][M]
    float* bs = Dy + (size_t)n * M;       // [M]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * M;
    const int Mc = min(M, p.d - m0);
    for (int i = tid; i < n * e; i += kT) Xs[(i / e) * ep + i % e] = p.x[i];
    for (int i = tid; i < Mc * e; i += kT) Ws[i] = p.w[(size_t)m0 * e + i];
    for (int i = tid; i < n * Mc; i += kT) Ts[(i / Mc) * M + i % Mc] = p.target[(size_t)(i / Mc) * p.d + m0 + i % Mc];
    if (tid < Mc) bs[tid] = p.b[m0 + tid];
    __syncthreads();
    const float count = (float)((int64_t)n * p.d);     // nn.MSELoss(reduction='mean') over n * d elements
    for (int epoch = 0; epoch < p.epochs; ++epoch) {
        // ---- forward + loss gradient: one warp per (sample, output dimension) ----
        double ls = 0.0;
        for (int pr = warp; pr < n * Mc; pr += kT / 32) {
            const int i = pr / Mc, m = pr - i * Mc;
            const float* xr = Xs + i * ep;
            const float* wr = Ws + m * e;
            float acc = 0.f;
            for (int k = lane; k < e; k += 32) acc = fmaf(xr[k], wr[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const float dlt = (acc + bs[m]) - Ts[i * M + m];
                Dy[i * M + m] = 2.f * dlt / count;
                ls += (double)dlt * (double)dlt;
            }
        }
        ls = block_sum(ls, red);          // (its barriers also publish Dy)
        if (tid == 0) p.loss_part[(size_t)epoch * gridDim.x + blockIdx.x] = ls;
        // ---- weight gradient + SGD step: one thread per (output dimension, input dimension) ----
        for (int o = tid; o < Mc * e; o += kT) {
            const int m = o / e, k = o - m * e;
            float g = 0.f;
            for (int i = 0; i < n; ++i) g = fmaf(Dy[i * M + m], Xs[i * ep + k], g);
            const float w = Ws[o];
            Ws[o] = w - p.lr * fmaf(p.wd, w, g);
        }
        __syncthreads();                  // Dy is read by the bias update below and rewritten by the next step
        if (tid < Mc) {
            float g = 0.f;
            for (int i = 0; i < n; ++i) g += Dy[i * M + tid];
            const float bv = bs[tid];
            bs[tid] = bv - p.lr * fmaf(p.wd, bv, g);
        }
        __syncthreads();
    }
    for (int i = tid; i < Mc * e; i += kT) p.w[(size_t)m0 * e + i] = Ws[i];
    if (tid < Mc) p.b[m0 + tid] = bs[tid];
}

// loss[epoch] = (sum over CTAs, in a fixed order) / (n d)
__global__ void fit_loss_kernel(const double* __restrict__ part, int epochs, int ctas, double count, float* __restrict__ loss) {
    const int ep = blockIdx.x * blockDim.x + threadIdx.x;
    if (ep >= epochs) return;
    double s = 0.0;
    for (int c = 0; c < ctas; ++c) s += part[(size_t)ep * ctas + c];
    loss[ep] = (float)(s / count);
}

int dims_per_cta(int d) { return (d + current_device_sms() - 1) / current_device_sms(); }
size_t fit_smem(int n, int e, int M) { return ((size_t)n * (e + 1) + (size_t)M * e + 2 * (size_t)n * M + M) * sizeof(float); }
}  // namespace

extern "C" int64_t sr_fit_linear_map_workspace_bytes(int32_t n, int32_t e, int32_t d, int32_t epochs) {
    if (n < 1 || e < 1 || d < 1 || epochs < 0) return -1;
    const int M = dims_per_cta(d);
    if (fit_smem(n, e, M) > 220 * 1024) return 0;      // does not fit: the caller uses the per-op route
    const int ctas = (d + M - 1) / M;
    return srb::align_up((int64_t)std::max(epochs, 1) * ctas * 8, 256);
}

extern "C" int32_t sr_fit_linear_map(const float* x, const float* target, float* weight, float* bias, int32_t n, int32_t e,
                                     int32_t d, int32_t epochs, float lr, float weight_decay, float* loss_trace,
                                     void* workspace, int64_t workspace_bytes, void* stream_v) {
    using namespace srb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (!x || !target || !weight || !bias || n < 1 || e < 1 || d < 1 || epochs < 0)
        return fail(SR_E_ARG, "sr_fit_linear_map: bad arguments");
    if (epochs == 0) return SR_OK;
    const int64_t need = sr_fit_linear_map_workspace_bytes(n, e, d, epochs);
    if (need == 0) return fail(SR_E_ARG, "sr_fit_linear_map: %d x %d inputs do not fit in shared memory", n, e);
    if (!workspace || workspace_bytes < need)
        return fail(SR_E_SMALLWS, "sr_fit_linear_map: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
    FitParams p;
    p.x = x; p.target = target; p.w = weight; p.b = bias;
    p.n = n; p.e = e; p.d = d; p.epochs = epochs; p.M = dims_per_cta(d);
    p.lr = lr; p.wd = weight_decay;
    p.loss_part = static_cast<double*>(workspace);
    const int ctas = (d + p.M - 1) / p.M;
    static PerDeviceOnce once;
    once_per_device(once, [] {
        cudaFuncSetAttribute(fit_linear_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    fit_linear_map_kernel<<<ctas, kT, fit_smem(n, e, p.M), stream>>>(p);
    if (loss_trace)
        fit_loss_kernel<<<(epochs + 127) / 128, 128, 0, stream>>>(p.loss_part, epochs, ctas, (double)n * (double)d, loss_trace);
    SR_CUDA_OK(cudaGetLastError());
    return SR_OK;
}
