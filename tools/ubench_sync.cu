// Micro-benchmarks of the synchronisation / issue primitives the convolution kernel is built from (sm_100a).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I subspace-reg_b200/csrc tools/ubench_sync.cu -o gpurun_out/ubench -lcuda
// Each line reports cycles per pipeline step for a producer warp / consumer warp(s) ring of `depth` slots.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "ptx.cuh"

using namespace srb;

struct UB {
    CUtensorMap tm;
    int steps, depth, n_cons, use_commit, mma_per_step, mma_n, tma_rows, n_tma, elect, slot_bytes, row_bytes;
    long long* out;
};

__global__ void __launch_bounds__(96, 1) k_ring(const __grid_constant__ UB p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16], empty[16];
    __shared__ uint32_t tmem_slot;
    __shared__ long long t_end[3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.depth; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], p.n_cons);
        }
        mbar_fence_init();
        tma_prefetch_desc(&p.tm);
    }
    if (warp == 1) {
        tmem_alloc_dyn(&tmem_slot, 512);
        tmem_relinquish();
    }
    // zero the ring so the MMAs read finite data
    for (int i = threadIdx.x; i < p.depth * p.slot_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t a_full = smem_u32(&full[0]), a_empty = smem_u32(&empty[0]), base = smem_u32(smem);
    const long long t0 = clock64();
    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int row = (blockIdx.x * 977) & 32767;
            for (int it = 0; it < p.steps; ++it) {
                mbar_wait_a(a_empty + 8u * s, ph ^ 1u);
                if (p.n_tma) {
                    mbar_expect_tx_a(a_full + 8u * s, (uint32_t)(p.n_tma * p.tma_rows * p.row_bytes));
                    for (int j = 0; j < p.n_tma; ++j) {
                        tma_load_2d_a(base + (uint32_t)(s * p.slot_bytes + j * p.tma_rows * p.row_bytes), &p.tm, a_full + 8u * s, 0, row);
                        row = (row + p.tma_rows) & 32767;
                    }
                } else {
                    mbar_arrive_a(a_full + 8u * s);
                }
                if (++s == p.depth) { s = 0; ph ^= 1u; }
            }
            t_end[0] = clock64();
        }
    } else if (warp - 1 < p.n_cons) {
        const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.mma_n);
        const uint32_t dhi = umma_desc_hi(128);
        const uint32_t acc = tmem_base + (uint32_t)((warp - 1) * 256);
        if (!p.elect) {
            if (lane == 0) {
                int s = 0;
                uint32_t ph = 0;
                for (int it = 0; it < p.steps; ++it) {
                    mbar_wait_a(a_full + 8u * s, ph);
                    tc_fence_after();
                    const uint32_t lo = umma_desc_lo(base + (uint32_t)(s * p.slot_bytes));
                    for (int k = 0; k < p.mma_per_step; ++k) umma_f16_split(acc, lo + 2 * (k & 3), lo + 2 * (k & 3), dhi, idesc, 1u);
                    if (p.use_commit) umma_commit_a(a_empty + 8u * s);
                    else mbar_arrive_a(a_empty + 8u * s);
                    if (++s == p.depth) { s = 0; ph ^= 1u; }
                }
                t_end[warp] = clock64();
            }
        } else {
            // warp-converged variant: every lane polls, one elected lane issues
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < p.steps; ++it) {
                mbar_wait_a(a_full + 8u * s, ph);
                tc_fence_after();
                const uint32_t lo = umma_desc_lo(base + (uint32_t)(s * p.slot_bytes));
                if (elect_one()) {
                    for (int k = 0; k < p.mma_per_step; ++k) umma_f16_split(acc, lo + 2 * (k & 3), lo + 2 * (k & 3), dhi, idesc, 1u);
                    if (p.use_commit) umma_commit_a(a_empty + 8u * s);
                    else mbar_arrive_a(a_empty + 8u * s);
                }
                __syncwarp();
                if (++s == p.depth) { s = 0; ph ^= 1u; }
            }
            if (lane == 0) t_end[warp] = clock64();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // drain: make sure every MMA retired before TMEM is released
        long long m = t_end[0];
        for (int i = 1; i <= p.n_cons; ++i) m = t_end[i] > m ? t_end[i] : m;
        p.out[blockIdx.x] = m - t0;
    }
    if (warp == 1) {
        if (lane == 0) {
            umma_commit_a(a_full);   // reuse as a drain barrier: its phase is irrelevant now
        }
        __syncwarp();
        // crude drain: wait long enough for outstanding MMAs
        const long long t1 = clock64();
        while (clock64() - t1 < 200000) {}
        tmem_dealloc_dyn(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);     \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

int main() {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeTiledFn encode = (EncodeTiledFn)fnp;
    void* gbuf;
    const size_t rows = 65536;
    CK(cudaMalloc(&gbuf, rows * 128));
    CK(cudaMemset(gbuf, 0, rows * 128));
    long long* dout;
    CK(cudaMalloc(&dout, 148 * sizeof(long long)));
    CK(cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    auto run = [&](const char* name, int grid, int depth, int n_cons, int use_commit, int mma_per_step, int mma_n, int tma_rows,
                   int n_tma, int elect, int row_bytes = 128, int slot_bytes = 32768) {
        UB p;
        memset(&p, 0, sizeof(p));
        p.steps = 4096;
        p.depth = depth;
        p.n_cons = n_cons;
        p.use_commit = use_commit;
        p.mma_per_step = mma_per_step;
        p.mma_n = mma_n;
        p.tma_rows = tma_rows;
        p.n_tma = n_tma;
        p.elect = elect;
        p.slot_bytes = slot_bytes;
        p.row_bytes = row_bytes;
        p.out = dout;
        if (n_tma) {
            cuuint64_t gdim[2] = {(cuuint64_t)row_bytes / 2, rows};
            cuuint64_t gstr[1] = {(cuuint64_t)row_bytes};
            cuuint32_t box[2] = {(cuuint32_t)row_bytes / 2, (cuuint32_t)tma_rows};
            cuuint32_t est[2] = {1, 1};
            CUresult r = encode(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gbuf, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
        }
        if (depth * p.slot_bytes > 196 * 1024) { printf("%s: ring too large\n", name); return; }
        for (int rep = 0; rep < 2; ++rep) {
            k_ring<<<grid, 96, 200 * 1024>>>(p);
            CK(cudaDeviceSynchronize());
        }
        long long h[148];
        CK(cudaMemcpy(h, dout, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double cps = (double)mx / p.steps;
        printf("%-44s grid %3d depth %d cons %d commit %d mma %2d N %3d tma %dx%3d rows elect %d : %8.1f cyc/step", name, grid, depth,
               n_cons, use_commit, mma_per_step, mma_n, n_tma, tma_rows, elect, cps);
        if (n_tma) printf("  (%.1f B/cyc/SM, %.2f rows/cyc, %d B rows)", (double)n_tma * tma_rows * row_bytes / cps, (double)n_tma * tma_rows / cps, row_bytes);
        if (mma_per_step) printf("  (%.1f cyc/MMA/warp, pipe floor %d)", cps / mma_per_step, n_cons * 128 * mma_n / 256);
        printf("\n");
    };

    if (getenv("UB_TMA")) {
        for (int g : {1, 148}) {
            for (int rb : {128, 64, 32}) {
                for (int nt : {1, 2, 3, 4}) {
                    if (nt * 256 * rb > 65536) continue;
                    run("tma rate", g, 3, 1, 0, 0, 64, 256, nt, 0, rb, 65536);
                }
            }
        }
        return 0;
    }
    // 1. pure handshake
    run("handshake arrive", 1, 4, 1, 0, 0, 64, 0, 0, 0);
    run("handshake commit", 1, 4, 1, 1, 0, 64, 0, 0, 0);
    run("handshake commit depth2", 1, 2, 1, 1, 0, 64, 0, 0, 0);
    run("handshake commit depth6", 1, 6, 1, 1, 0, 64, 0, 0, 0);
    run("handshake commit 2 consumers", 1, 4, 2, 1, 0, 64, 0, 0, 0);
    run("handshake arrive 2 consumers", 1, 4, 2, 0, 0, 64, 0, 0, 0);
    run("handshake commit 2 consumers, 148 CTAs", 148, 4, 2, 1, 0, 64, 0, 0, 0);
    // 2. MMA issue rate (no TMA)
    for (int n : {64, 160, 256}) {
        for (int k : {4, 8, 16}) {
            run("mma lane0", 1, 4, 1, 1, k, n, 0, 0, 0);
            run("mma elect", 1, 4, 1, 1, k, n, 0, 0, 1);
            run("mma lane0 2 warps", 1, 4, 2, 1, k, n, 0, 0, 0);
            run("mma elect 2 warps", 1, 4, 2, 1, k, n, 0, 0, 1);
        }
    }
    // 3. TMA only (L2-resident 8 MB source), one CTA and all SMs
    for (int g : {1, 148}) {
        for (int r : {32, 64, 128, 256}) {
            run("tma ring", g, 4, 1, 0, 0, 64, r, 1, 0);
            run("tma ring x2", g, 4, 1, 0, 0, 64, r / 2, 2, 0);
        }
        run("tma ring depth6 256 rows", g, 6, 1, 0, 0, 64, 256, 1, 0);
    }
    // 4. everything: TMA + MMAs, 148 CTAs
    for (int n : {64, 160}) {
        run("full lane0", 148, 4, 2, 1, 4, n, 128, 2, 0);
        run("full elect", 148, 4, 2, 1, 4, n, 128, 2, 1);
        run("full lane0 12 mma", 148, 4, 2, 1, 12, n, 128, 2, 0);
        run("full elect 12 mma", 148, 4, 2, 1, 12, n, 128, 2, 1);
    }
    return 0;
}
