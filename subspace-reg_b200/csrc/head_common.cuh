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
}  // namespace srb
