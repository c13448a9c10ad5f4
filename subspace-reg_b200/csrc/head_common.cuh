// Device helpers shared by the two persistent head kernels (head.cu: tiled, any size; head_small.cu: resident operands).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Sum over the block (result valid in every thread). `red` = 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// Three block sums for the price of one barrier pair (same summation order as three block_sum calls; blocks of <= 10 warps).
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red) {
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) { red[w] = a; red[nw + w] = b; red[2 * nw + w] = c; }
    __syncthreads();
    double ta = 0.0, tb = 0.0, tc = 0.0;
    for (int i = 0; i < nw; ++i) { ta += red[i]; tb += red[nw + i]; tc += red[2 * nw + i]; }
    a = ta; b = tb; c = tc;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct HeadCtrl {          // lives at the start of the workspace (zeroed by the host before launch)
    unsigned int barrier;  // monotonically increasing arrival counter
    int stop;
    int epochs_done;
    int stable_count;
    float prev_loss;
    int error;
    int pad[2];
    unsigned long long t_ns[12];  // profiling aid: CTA 0 ns in phase 1 / barrier 1 / phase 2 / barrier 2; loss CTA, pull CTA phase 1;
                                  // [6..8] phase 1: W stream, logits to smem, softmax + dlogits; [9..11] phase 2: prologue, DLt stream, update
    double norm_base_sq;   // ||W[:nb] - W0||_F^2
    double norm_prev_sq;   // ||W[nb:nb+np] - Wres||_F^2
};

// Where this launch starts: the host's values, or - chained launches - the previous launch's status block on the device.
struct HeadStart {
    int epoch0, step0, stable_count0;
    float prev_loss;
    bool already_stopped;
};
} // namespace srb
#include "../../include/srb200.h"
namespace srb {
__device__ __forceinline__ HeadStart head_start(const sr_head_args& a) {
    HeadStart s;
    s.epoch0 = a.epoch0; s.step0 = a.step0; s.stable_count0 = a.stable_count0; s.prev_loss = a.prev_loss;
    s.already_stopped = false;
    if (a.resume_status != nullptr) {
        s.already_stopped = a.resume_status[1] != 0;
        s.stable_count0 = a.resume_status[2];
        s.epoch0 = a.resume_status[4];
        s.step0 = a.resume_status[4];
        s.prev_loss = __int_as_float(a.resume_status[5]);
    }
    return s;
}
__device__ __forceinline__ void head_write_status(const sr_head_args& a, const HeadStart& st, int epochs_done, int stop,
                                                  int stable_count, float last_loss, int error) {
    a.status[0] = epochs_done;
    a.status[1] = stop;
    a.status[2] = stable_count;
    a.status[3] = error;
    a.status[4] = st.epoch0 + epochs_done;
    a.status[5] = __float_as_int(last_loss);
    a.status[6] = 0;
    a.status[7] = 0;
}

__device__ __forceinline__ void grid_barrier(HeadCtrl* ctrl, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(&ctrl->barrier, 1u);
        const long long t0 = clock64();
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(&ctrl->barrier) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 8000000000LL) {  // never hang the device
                ctrl->error = 1;
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}


}  // namespace srb

#include "../../include/srb200.h"
namespace srb {
bool head_small_applicable(const sr_head_args* a);
int64_t head_small_workspace_bytes(const sr_head_args* a);
int32_t head_small_run(const sr_head_args* a, cudaStream_t stream);
void head_small_shape(const sr_head_args* a, int* rows_per_cta, int* cols_per_cta, int* ctas);   // host only
bool head_cluster_applicable(const sr_head_args* a);
int64_t head_cluster_workspace_bytes(const sr_head_args* a);
int32_t head_cluster_run(const sr_head_args* a, cudaStream_t stream);
bool head_tc_applicable(const sr_head_args* a);
int64_t head_tc_workspace_bytes(const sr_head_args* a);
int32_t head_tc_run(const sr_head_args* a, cudaStream_t stream);
int64_t eval_tc_workspace_bytes(int n, int dim, int n_classes);
int32_t eval_tc_logits(const sr_eval_args* a, cudaStream_t stream, const float** z, int* pitch);
}  // namespace srb
