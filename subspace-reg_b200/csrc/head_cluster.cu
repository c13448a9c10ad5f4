// Paper-size persistent head kernel, second generation: thread-block clusters split BOTH reductions of an epoch.
//
// head_small_kernel (head_small.cu) gives every row CTA all of W^T to stream (256 KB per CTA and epoch) and every column
// CTA all of the dlogits (144 KB): each phase is a serial chain of ten ~0.47 us chunk hand-offs.  Here a CLUSTER of eight
// CTAs owns a block of rows (phase 1) / a block of feature columns (phase 2) and splits the reduction dimension between
// its CTAs, so a CTA brings in ONE 19-32 KB chunk per phase with one bulk TMA copy:
//
//   phase 1  cluster c = rows [c RB, (c+1) RB); CTA j = features [j d/8, (j+1) d/8): partial logits P_j[RB][C] from its
//            resident X block and its chunk of W^T -> own shared memory; cluster barrier; CTA j adds the eight partials
//            of RB/8 rows through distributed shared memory (9.6 KB of DSMEM reads), then softmax, CE, top-k, dlogits.
//            Cluster 0 also forms the projection coefficients u = W_new Q the same way (partial over the feature block).
//   phase 2  cluster c = columns [c d/NCL, (c+1) d/NCL); CTA j = samples [j SB, (j+1) SB): partial dW_j[C][d/NCL] from its
//            resident X block and its chunk of the dlogits; cluster barrier; CTA j adds the eight partials of C/8 classes
//            (DSMEM), applies every regulariser gradient, weight decay and the optimiser to that (class, column) slice
//            whose master copy / momentum / anchors stay resident, publishes W, W^T and the norm partials.
//
// Two grid barriers per epoch remain (logits need all of W, dW needs all dlogits); everything else of head_small's
// protocol is kept: loss of epoch e-1 assembled during phase 1 of epoch e, the reference's stopping rule
// (language_eval.py:298-318) read after barrier 1, closed-form projection gradient, chained launches (resume_status).
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include "common.h"
#include "head_common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace {
using namespace srb;

constexpr int kT = 256;
constexpr int CS = 8;       // CTAs per cluster
constexpr int MAXR = 8;     // rows per thread in phase 1 (RB <= 64)

struct ClParams {
    sr_head_args a;
    int NT, NCL, G;         // rows, work clusters, grid = NCL * CS
    int RB, RS, KB, DCB, SB, CB, cw;
    int MAXNEW;             // newest classes one CTA can own: ceil(n_new / CS) + 1
    int TC;                 // phase 2: columns per thread tile (4 or 8)
    int RPT, RBP;           // phase 1: rows per thread, padded rows per cluster (= RPT * row groups)
    HeadCtrl* ctrl;
    float* DL;              // [CS * SB][cw]  dlogits, sample-major (rows >= NT stay zero)
    float* Wt;              // [d][cw]        W^T, kept in step with `weight`
    float* rowloss;         // [2][NT]
    int* rowhit;            // [2][NT]
    double* nb_part;        // [2][G]
    double* nn_part;        // [2][G]
    double* pull_part;      // [G]
    float* u;               // [n_new][q]
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ int64_t feat_row(const sr_head_args& a, int r) {
    return r < a.n_support ? (int64_t)a.support_row0 + r : (int64_t)a.memory_row0 + (r - a.n_support);
}

// Loss of epoch `e` from the partial results in the workspace + the reference's stopping rule.  One CTA.
__device__ void assemble_loss(const ClParams& p, const HeadStart& st, int e) {
    __shared__ double s7[kT / 32][7];
    const sr_head_args& a = p.a;
    const int par = e & 1;
    const int tid = threadIdx.x;
    double v[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // ce support, ce memory, top-1, top-5, ||dW0||^2, ||dWres||^2, ||Pw-w||^2
    for (int r = tid; r < p.NT; r += kT) {
        const float l = p.rowloss[par * p.NT + r];
        const int h = p.rowhit[par * p.NT + r];
        if (r < a.n_support) { v[0] += (double)l; v[2] += (double)(h & 1); v[3] += (double)((h >> 1) & 1); }
        else v[1] += (double)l;
    }
    for (int i = tid; i < p.G; i += kT) {
        v[4] += p.nb_part[par * p.G + i];
        v[5] += p.nn_part[par * p.G + i];
        v[6] += p.pull_part[i];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = warp_sum(v[k]);
    if ((tid & 31) == 0)
#pragma unroll
        for (int k = 0; k < 7; ++k) s7[tid >> 5][k] = v[k];
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            double t = 0.0;
            for (int w = 0; w < kT / 32; ++w) t += s7[w][k];
            v[k] = t;
        }
        const bool has_base = a.base_weight != nullptr;
        const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
        const bool has_pull = a.pull_mode != SR_PULL_NONE;
        const float ce_s = (float)(v[0] / (double)a.n_support);
        const float ce_m = a.n_memory > 0 ? (float)(v[1] / (double)a.n_memory) : 0.f;
        const float reg_b = has_base ? a.lmbd_base * (float)sqrt(v[4]) : 0.f;
        const float reg_n = has_prev ? a.lmbd_novel * (float)sqrt(v[5]) : 0.f;
        const float pull = has_pull ? a.gamma * (float)v[6] : 0.f;
        float loss = ce_s;
        if (a.n_memory > 0) loss += ce_m;
        if (has_base) loss += reg_b;
        if (has_prev) loss += reg_n;
        if (has_pull) loss += pull;
        float* tr = a.loss_trace + (int64_t)e * SR_TRACE_COLS;
        tr[0] = loss; tr[1] = ce_s; tr[2] = ce_m; tr[3] = reg_b; tr[4] = reg_n; tr[5] = pull; tr[6] = (float)v[2]; tr[7] = (float)v[3];
        int stop = 0;
        int sc = e == 0 ? st.stable_count0 : p.ctrl->stable_count;
        const float prev = e == 0 ? st.prev_loss : p.ctrl->prev_loss;
        if (a.stable) {
            if (fabs((double)loss - (double)prev) < a.convergence_epsilon) sc += 1; else sc = 0;
            if (sc == a.stable_epochs) stop = 1;
        }
        const int epoch = st.epoch0 + e + 1;
        if (epoch >= a.max_novel_epochs || ((double)loss <= a.target_train_loss && epoch >= a.min_novel_epochs + 1)) stop = 1;
        p.ctrl->stable_count = sc;
        p.ctrl->prev_loss = loss;
        p.ctrl->epochs_done = e + 1;
        p.ctrl->stop = stop;
    }
    __syncthreads();
}

// Partial logits of this thread: RPT consecutive rows x 4 classes over the CTA's KB features.  The k loop is unrolled by
// four with every shared-memory load of a group issued before its FMAs: with two warps per scheduler the loop is otherwise
// bound by the load latency (measured 9.5 us for 960 FMAs per thread; 1 us of FMA issue).
template <int RPT>
__device__ __forceinline__ void logits_partial(const float* __restrict__ Ws, const float* __restrict__ Xk, float* __restrict__ P,
                                               int KB, int cw, int RBP, int RB, int cg4, int rg) {
    float acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    const float* wsrc = Ws + cg4 * 4;
    const float* xsrc = Xk + rg * RPT;
    for (int k = 0; k < KB; k += 4) {     // KB is a multiple of 8
        float4 w[4];
        float x[4][RPT];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            w[u] = *reinterpret_cast<const float4*>(wsrc + (k + u) * cw);
#pragma unroll
            for (int i = 0; i < RPT; ++i) x[u][i] = xsrc[(k + u) * RBP + i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                acc[i][0] = fmaf(x[u][i], w[u].x, acc[i][0]);
                acc[i][1] = fmaf(x[u][i], w[u].y, acc[i][1]);
                acc[i][2] = fmaf(x[u][i], w[u].z, acc[i][2]);
                acc[i][3] = fmaf(x[u][i], w[u].w, acc[i][3]);
            }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i)
        if (rg * RPT + i < RB)
            *reinterpret_cast<float4*>(P + (rg * RPT + i) * cw + cg4 * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// Partial dW tile of this thread: 4 classes x TC columns over the CTA's SB samples (same load-first unrolling).
template <int TC>
__device__ __forceinline__ void dw_partial(const float* __restrict__ DLs, const float* __restrict__ Xn, float* __restrict__ D,
                                           int SB, int cw, int DCB, int C, int cg_, int colg) {
    float acc[4][TC];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int x = 0; x < TC; ++x) acc[c][x] = 0.f;
    const float* dsrc = DLs + cg_ * 4;
    const float* xsrc = Xn + colg * TC;
    for (int s = 0; s < SB; s += 4) {     // SB is a multiple of 4
        float4 dv[4];
        float xv[4][TC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dv[u] = *reinterpret_cast<const float4*>(dsrc + (s + u) * cw);
#pragma unroll
            for (int v = 0; v < TC / 4; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(xsrc + (s + u) * DCB + v * 4);
                xv[u][4 * v] = t.x; xv[u][4 * v + 1] = t.y; xv[u][4 * v + 2] = t.z; xv[u][4 * v + 3] = t.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int x = 0; x < TC; ++x) acc[c][x] = fmaf(ds[c], xv[u][x], acc[c][x]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (cg_ * 4 + c < C)
#pragma unroll
            for (int v = 0; v < TC / 4; ++v)
                *reinterpret_cast<float4*>(D + (cg_ * 4 + c) * DCB + colg * TC + v * 4) =
                    make_float4(acc[c][4 * v], acc[c][4 * v + 1], acc[c][4 * v + 2], acc[c][4 * v + 3]);
}

// profiling aid: CTA 0 accumulates the nanoseconds between consecutive stamps of an epoch into ctrl->t_ns[0..11]
#define CL_STAMP(i)                                                        \
    do {                                                                   \
        if (cta == t_cta && tid == 0) {                                    \
            const unsigned long long t_ = global_ns();                     \
            p.ctrl->t_ns[i] += t_ - t_prev;                                \
            t_prev = t_;                                                   \
        }                                                                  \
    } while (0)

template <int CP>
__global__ void __launch_bounds__(kT, 1) head_cluster_kernel(const ClParams p) {
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ double red[32];
    __shared__ int s_stop;
    __shared__ uint64_t s_bar[2];   // [0] W^T chunk landed, [1] dlogits chunk landed
    cg::cluster_group cluster = cg::this_cluster();
    const sr_head_args& a = p.a;
    const int C = a.n_classes, d = a.dim, NT = p.NT, q = a.q_rows, cw = p.cw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const int cl = cta / CS, j = cta % CS;           // cluster index, rank in cluster
    const bool is_loss = cta == p.G - 1;
    const bool has_base = a.base_weight != nullptr;
    const bool has_prev = a.reserve_weight != nullptr && a.n_prev_novel > 0;
    const bool proj = a.pull_mode == SR_PULL_PROJECT && q < d;   // q >= d: P = I, the term vanishes
    const bool fixed = a.pull_mode == SR_PULL_FIXED;
    const bool do_u = proj && cl == 0;               // cluster 0 forms the projection coefficients
    const int new0 = C - a.n_new;
    const int n_opt = a.optimizer == SR_OPT_ADAM ? 2 : 1;
    const int RB = p.RB, RS = p.RS, KB = p.KB, DCB = p.DCB, SB = p.SB, CB = p.CB;
    // my classes in phase 2: c = lc * CS + j (interleaved, so that the newest classes - the ones with the projection work -
    // spread over the CTAs of a cluster instead of landing on the last one)
    const int n_mine = C > j ? (C - j + CS - 1) / CS : 0;
    const bool own_new = (proj || fixed) && n_mine > 0 && (n_mine - 1) * CS + j >= new0;
    const HeadStart st = head_start(a);
    if (st.already_stopped) {   // chained launch after the stopping rule fired (every CTA takes this exit)
        if (is_loss && tid == 0) head_write_status(a, st, 0, 1, st.stable_count0, st.prev_loss, 0);
        return;
    }

    // ---- shared memory carve-up (floats; every region a multiple of 4 floats so that 16-byte accesses stay aligned) ----
    auto up4 = [](int n) { return (n + 3) & ~3; };
    float* sp = reinterpret_cast<float*>(dyn);
    float* Ws = sp;   sp += up4(KB * cw);                   // [KB][cw]   chunk of W^T (bulk copy target)
    float* DLs = sp;  sp += up4(SB * cw);                   // [SB][cw]   chunk of the dlogits (bulk copy target)
    float* Xk = sp;   sp += up4(KB * p.RBP);                // [KB][RBP]  my rows x my features, k-major (rows >= RB are zero)
    float* P = sp;    sp += up4(RB * cw);                   // [RB][cw]   partial logits (read by the cluster)
    float* Zs = sp;   sp += up4(RS * cw);                   // [RS][cw]   full logits of my RS rows
    float* Xn = sp;   sp += up4(SB * DCB);                  // [SB][DCB]  my samples x my columns
    float* D = sp;    sp += up4(C * DCB);                   // [C][DCB]   partial dW (read by the cluster)
    // (regions read by cluster siblings - P, D, Up - sit at offsets that are identical in every CTA of the cluster:
    // everything above and the two below depend on launch-wide or cluster-wide conditions only)
    float* Up = sp;   sp += do_u ? up4(a.n_new * q) : 0;    // partial u (read by the cluster)
    float* Qk = sp;   sp += do_u ? up4(q * KB) : 0;         // [KB][q]  Q restricted to my feature block, k-major
    float* Wc = sp;   sp += up4(CB * DCB);                  // master copy of my (class slice, column block)
    float* Vc = sp;   sp += up4(n_opt * CB * DCB);
    float* W0c = sp;  sp += has_base ? up4(CB * DCB) : 0;
    float* Rc = sp;   sp += has_prev ? up4(CB * DCB) : 0;
    float* Pc = sp;   sp += own_new ? up4((proj ? q : a.n_new) * DCB) : 0;   // Q^T columns or puller columns
    float* Us = sp;   sp += (own_new && proj) ? up4(a.n_new * q) : 0;
    float* Pw = sp;   sp += (own_new && proj) ? up4(p.MAXNEW * DCB) : 0;   // projection of my newest classes, my columns

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    const uint32_t bar_w = smem_u32(&s_bar[0]), bar_d = smem_u32(&s_bar[1]);

    // ---- one-time loads ----
    const int r0 = cl * RB, k0 = j * KB;            // phase-1 block
    for (int i = tid; i < KB * p.RBP; i += kT) {
        const int k = i / p.RBP, r = i % p.RBP;
        const int n = r0 + r;
        Xk[i] = (r < RB && n < NT) ? a.feat[feat_row(a, n) * d + k0 + k] : 0.f;
    }
    const int col0 = cl * DCB, s0 = j * SB;         // phase-2 block
    for (int i = tid; i < SB * DCB; i += kT) {
        const int s = i / DCB, c = i % DCB;
        const int n = s0 + s;
        Xn[i] = n < NT ? a.feat[feat_row(a, n) * d + col0 + c] : 0.f;
    }
    for (int i = tid; i < n_mine * DCB; i += kT) {
        const int c = (i / DCB) * CS + j, col = col0 + i % DCB;
        Wc[i] = a.weight[(int64_t)c * d + col];
        p.Wt[(int64_t)col * cw + c] = Wc[i];
        for (int s = 0; s < n_opt; ++s) Vc[s * CB * DCB + i] = a.opt_state[(int64_t)s * C * d + (int64_t)c * d + col];
        if (has_base && c < a.n_base) W0c[i] = a.base_weight[(int64_t)c * d + col];
        if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) Rc[i] = a.reserve_weight[(int64_t)(c - a.n_base) * d + col];
    }
    if (own_new) {
        const int rows = proj ? q : a.n_new;
        for (int i = tid; i < rows * DCB; i += kT) Pc[i] = a.pull[(int64_t)(i / DCB) * d + col0 + i % DCB];
    }
    if (do_u)
        for (int i = tid; i < q * KB; i += kT) Qk[(i % KB) * q + i / KB] = a.pull[(int64_t)(i / KB) * d + k0 + i % KB];
    __syncthreads();
    {   // norm partials of the initial W (buffer parity 0)
        double nb = 0.0, nn = 0.0, unused = 0.0;
        for (int i = tid; i < n_mine * DCB; i += kT) {
            const int c = (i / DCB) * CS + j;
            if (has_base && c < a.n_base) { const float dl = Wc[i] - W0c[i]; nb += (double)dl * dl; }
            if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) { const float dl = Wc[i] - Rc[i]; nn += (double)dl * dl; }
        }
        block_sum3(nb, nn, unused, red);
        if (tid == 0) { p.nb_part[cta] = nb; p.nn_part[cta] = nn; p.pull_part[cta] = 0.0; }
    }
    unsigned int bar_target = 0;
    grid_barrier(p.ctrl, bar_target);

    constexpr int CG = CP / 4;          // 4-class groups (16 or 32)
    const int cg4 = tid % CG, rg = tid / CG;   // 4-class group, row group (kT / CG of them)
    const uint32_t w_bytes = (uint32_t)(KB * cw * 4), d_bytes = (uint32_t)(SB * cw * 4);
    int e = 0;
    bool stopped = false;
    unsigned long long t_prev = 0;
    const int t_cta = p.G > CS ? CS : 0;   // the CTA whose phase times are recorded: rank 0 of cluster 1 (no u work)
    for (; e < a.max_epochs; ++e) {
        const int par = e & 1;
        if (cta == t_cta && tid == 0) t_prev = global_ns();
        // ======================= phase 1: logits, softmax, dlogits =======================
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx_a(bar_w, w_bytes);
            bulk_load_1d(smem_u32(Ws), p.Wt + (int64_t)k0 * cw, w_bytes, bar_w);
        }
        if (is_loss && e > 0) assemble_loss(p, st, e - 1);   // while the chunk is in flight
        mbar_wait_a(bar_w, (uint32_t)(e & 1));
        CL_STAMP(0);   // W^T chunk landed
        if (cg4 * 4 < cw) {
            switch (p.RPT) {
                case 1: logits_partial<1>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
                case 2: logits_partial<2>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
                case 3: logits_partial<3>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
                case 4: logits_partial<4>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
                case 5: logits_partial<5>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
                case 6: logits_partial<6>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
                default: logits_partial<8>(Ws, Xk, P, KB, cw, p.RBP, RB, cg4, rg); break;
            }
        }
        if (do_u) {   // partial u_j[i][jq] = sum over my feature block of W_new[i][k] Q[jq][k]
            for (int o = tid; o < a.n_new * q; o += kT) {
                const int i = o / q, jq = o % q;
                const float* wk = Ws + (new0 + i);
                const float* qk = Qk + jq;            // consecutive lanes = consecutive jq: conflict-free rows of [KB][q]
                float s0 = 0.f, s1 = 0.f;
                int k = 0;
                for (; k + 2 <= KB; k += 2) {
                    s0 = fmaf(qk[k * q], wk[k * cw], s0);
                    s1 = fmaf(qk[(k + 1) * q], wk[(k + 1) * cw], s1);
                }
                for (; k < KB; ++k) s0 = fmaf(qk[k * q], wk[k * cw], s0);
                Up[o] = s0 + s1;
            }
        }
        CL_STAMP(1);   // partial logits (+ partial u)
        cluster.sync();
        CL_STAMP(2);   // cluster barrier
        // my RS rows: add the eight partials through distributed shared memory
        for (int o = tid; o < RS * (cw / 4); o += kT) {
            const int r = o / (cw / 4), c4 = o % (cw / 4);
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int rk = 0; rk < CS; ++rk) {
                const float* rp = cluster.map_shared_rank(P, rk);
                const float4 v = *reinterpret_cast<const float4*>(rp + (j * RS + r) * cw + c4 * 4);
                z.x += v.x; z.y += v.y; z.z += v.z; z.w += v.w;
            }
            *reinterpret_cast<float4*>(Zs + r * cw + c4 * 4) = z;
        }
        if (do_u) {   // my slice of u (n_new * q values split over the eight CTAs)
            const int total = a.n_new * q, per = (total + CS - 1) / CS;
            for (int o = j * per + tid; o < min(total, (j + 1) * per); o += kT) {
                float s = 0.f;
#pragma unroll
                for (int rk = 0; rk < CS; ++rk) s += cluster.map_shared_rank(Up, rk)[o];
                p.u[o] = s;
            }
        }
        __syncthreads();
        CL_STAMP(3);   // DSMEM reduction
        for (int r = warp; r < RS; r += kT / 32) {
            const int n = r0 + j * RS + r;
            if (n >= NT) continue;
            const float* z = Zs + r * cw;
            const bool is_sup = n < a.n_support;
            const int y = (int)(is_sup ? a.labels_support[n] : a.labels_memory[n - a.n_support]);
            const float inv_n = 1.f / (float)(is_sup ? a.n_support : a.n_memory);
            float mx = -INFINITY;
            for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
            mx = warp_max(mx);
            float se = 0.f;
            for (int c = lane; c < C; c += 32) se += expf(z[c] - mx);
            se = warp_sum(se);
            const float lse = mx + logf(se);
            const float zy = z[y];
            int greater = 0, tie_before = 0;
            for (int c = lane; c < C; c += 32) {
                const float zc = z[c];
                greater += zc > zy ? 1 : 0;
                tie_before += (zc == zy && c < y) ? 1 : 0;
            }
            greater = __reduce_add_sync(0xffffffffu, greater);
            tie_before = __reduce_add_sync(0xffffffffu, tie_before);
            for (int c = lane; c < C; c += 32) {
                const float pr = expf(z[c] - lse);
                p.DL[(int64_t)n * cw + c] = (pr - (c == y ? 1.f : 0.f)) * inv_n;
            }
            if (lane == 0) {
                p.rowloss[par * NT + n] = lse - zy;
                const int rank = greater + tie_before;
                p.rowhit[par * NT + n] = (rank == 0 ? 1 : 0) | (rank < 5 ? 2 : 0);
            }
        }
        CL_STAMP(4);   // softmax, dlogits
        grid_barrier(p.ctrl, bar_target);
        CL_STAMP(5);   // grid barrier 1
        if (tid == 0) s_stop = p.ctrl->stop;
        __syncthreads();
        if (e > 0 && s_stop) { stopped = true; break; }

        // ======================= phase 2: dW, regularisers, optimiser =======================
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx_a(bar_d, d_bytes);
            bulk_load_1d(smem_u32(DLs), p.DL + (int64_t)s0 * cw, d_bytes, bar_d);
        }
        // norms of W_e - anchors (partials from the previous epoch / init) and the projection coefficients of this epoch
        double nbs = 0.0, nns = 0.0;
        for (int i = tid; i < p.G; i += kT) { nbs += p.nb_part[par * p.G + i]; nns += p.nn_part[par * p.G + i]; }
        if (own_new && proj)
            for (int i = tid; i < a.n_new * q; i += kT) Us[i] = __ldcg(p.u + i);
        mbar_wait_a(bar_d, (uint32_t)(e & 1));
        CL_STAMP(6);   // dlogits chunk landed (+ norm partials, u)
        {
            // partial dW: thread = 4 classes x TC columns over my SB samples (TC = 8 when 4-column tiles would need a
            // second pass over the threads)
            const int ncg = cw / 4;
            if (p.TC == 8) {
                const int ncolg = DCB / 8;
                for (int t = tid; t < ncg * ncolg; t += kT) dw_partial<8>(DLs, Xn, D, SB, cw, DCB, C, t / ncolg, t % ncolg);
            } else {
                const int ncolg = DCB / 4;
                for (int t = tid; t < ncg * ncolg; t += kT) dw_partial<4>(DLs, Xn, D, SB, cw, DCB, C, t / ncolg, t % ncolg);
            }
        }
        {
            double unused = 0.0;
            block_sum3(nbs, nns, unused, red);
        }
        if (own_new && proj) {
            // P w for my newest classes and my columns, all threads: (class, column) = one 60-term dot product each
            const int lc_first = new0 > j ? (new0 - j + CS - 1) / CS : 0;      // first local class index that is a new class
            const int n_new_mine = n_mine - lc_first;
            for (int t = tid; t < n_new_mine * DCB; t += kT) {
                const int ln = t / DCB, lcol = t % DCB;
                const int c = (lc_first + ln) * CS + j;
                const float* ui = Us + (c - new0) * q;
                float pw0 = 0.f, pw1 = 0.f, pw2 = 0.f, pw3 = 0.f;
                int jq = 0;
                for (; jq + 4 <= q; jq += 4) {
                    pw0 = fmaf(ui[jq], Pc[jq * DCB + lcol], pw0);
                    pw1 = fmaf(ui[jq + 1], Pc[(jq + 1) * DCB + lcol], pw1);
                    pw2 = fmaf(ui[jq + 2], Pc[(jq + 2) * DCB + lcol], pw2);
                    pw3 = fmaf(ui[jq + 3], Pc[(jq + 3) * DCB + lcol], pw3);
                }
                for (; jq < q; ++jq) pw0 = fmaf(ui[jq], Pc[jq * DCB + lcol], pw0);
                Pw[ln * DCB + lcol] = (pw0 + pw1) + (pw2 + pw3);
            }
        }
        CL_STAMP(7);   // partial dW (+ projection of the newest classes)
        cluster.sync();
        CL_STAMP(8);   // cluster barrier
        const float nb = has_base ? (float)sqrt(nbs) : 0.f, nn = has_prev ? (float)sqrt(nns) : 0.f;
        const float sb = nb > 0.f ? a.lmbd_base / nb : 0.f;
        const float sn = nn > 0.f ? a.lmbd_novel / nn : 0.f;
        const int step = st.step0 + e;
        float bc1 = 1.f, bc2s = 1.f;
        if (a.optimizer == SR_OPT_ADAM) {
            bc1 = (float)(1.0 - pow((double)a.beta1, (double)(step + 1)));
            bc2s = (float)sqrt(1.0 - pow((double)a.beta2, (double)(step + 1)));
        }
        double nbp = 0.0, nnp = 0.0, pp = 0.0;
        for (int t = tid; t < n_mine * (DCB / 4); t += kT) {
            const int lc = t / (DCB / 4), colg = t % (DCB / 4);
            const int c = lc * CS + j;
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int rk = 0; rk < CS; ++rk) {
                const float* rp = cluster.map_shared_rank(D, rk);
                const float4 v = *reinterpret_cast<const float4*>(rp + c * DCB + colg * 4);
                g4.x += v.x; g4.y += v.y; g4.z += v.z; g4.w += v.w;
            }
            float gs[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int lcol = colg * 4 + x;
                const int idx = lc * DCB + lcol;
                const float w = Wc[idx];
                float g = gs[x];
                if (has_base && c < a.n_base) g += sb * (w - W0c[idx]);
                if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) g += sn * (w - Rc[idx]);
                if (c >= new0 && (proj || fixed)) {
                    float r;
                    if (proj) {
                        r = Pw[(lc - (new0 > j ? (new0 - j + CS - 1) / CS : 0)) * DCB + lcol] - w;
                    } else {
                        r = Pc[(c - new0) * DCB + lcol] - w;
                    }
                    pp += (double)r * (double)r;
                    g += -2.f * a.gamma * r;
                }
                g = fmaf(a.weight_decay, w, g);
                float wnew;
                if (a.optimizer == SR_OPT_SGD) {
                    float v = Vc[idx];
                    v = step == 0 ? g : fmaf(a.momentum, v, g);
                    Vc[idx] = v;
                    wnew = w - a.lr * v;
                } else {
                    float m1 = Vc[idx], m2 = Vc[CB * DCB + idx];
                    m1 = m1 + (1.f - a.beta1) * (g - m1);
                    m2 = a.beta2 * m2 + (1.f - a.beta2) * g * g;
                    Vc[idx] = m1;
                    Vc[CB * DCB + idx] = m2;
                    wnew = w - (a.lr / bc1) * (m1 / (sqrtf(m2) / bc2s + a.adam_eps));
                }
                Wc[idx] = wnew;
                a.weight[(int64_t)c * d + col0 + lcol] = wnew;
                p.Wt[(int64_t)(col0 + lcol) * cw + c] = wnew;
                if (has_base && c < a.n_base) { const float dl = wnew - W0c[idx]; nbp += (double)dl * dl; }
                if (has_prev && c >= a.n_base && c < a.n_base + a.n_prev_novel) { const float dl = wnew - Rc[idx]; nnp += (double)dl * dl; }
            }
        }
        block_sum3(nbp, nnp, pp, red);
        if (tid == 0) {
            p.nb_part[(par ^ 1) * p.G + cta] = nbp;
            p.nn_part[(par ^ 1) * p.G + cta] = nnp;
            p.pull_part[cta] = pp;
        }
        CL_STAMP(9);   // DSMEM reduction + update
        grid_barrier(p.ctrl, bar_target);
        CL_STAMP(10);  // grid barrier 2
    }
    // ---- tail: loss of the last applied epoch when the loop ran out of epochs; write back optimiser state ----
    if (!stopped && is_loss && e > 0) assemble_loss(p, st, e - 1);
    for (int i = tid; i < n_mine * DCB; i += kT) {
        const int c = (i / DCB) * CS + j, col = col0 + i % DCB;
        for (int s = 0; s < n_opt; ++s) a.opt_state[(int64_t)s * C * d + (int64_t)c * d + col] = Vc[s * CB * DCB + i];
    }
    if (is_loss) {
        __syncthreads();
        if (tid == 0) {
            const int done = p.ctrl->epochs_done;
            head_write_status(a, st, done, p.ctrl->stop, done > 0 ? p.ctrl->stable_count : st.stable_count0,
                              done > 0 ? p.ctrl->prev_loss : st.prev_loss, p.ctrl->error);
        }
    }
    cluster.sync();   // nobody leaves while a sibling may still read its shared memory
}

struct ClLayout {
    int64_t ctrl, DL, Wt, rowloss, rowhit, nb, nn, pull, u, total;
};

bool cl_geometry(const sr_head_args* a, int ncl, ClParams* p) {
    const int nt = a->n_support + a->n_memory;
    const int d = a->dim, C = a->n_classes;
    if (ncl < 1 || d % (ncl * 4) != 0 || d % (CS * 4) != 0) return false;
    p->NT = nt;
    p->NCL = ncl;
    p->G = ncl * CS;
    p->RB = (int)align_up((nt + ncl - 1) / ncl, CS);
    p->RS = p->RB / CS;
    p->KB = d / CS;
    p->DCB = d / ncl;
    p->SB = (int)align_up((nt + CS - 1) / CS, 4);
    p->CB = (C + CS - 1) / CS;
    p->cw = (C + 3) & ~3;
    const int CP = C <= 64 ? 64 : 128;
    const int RGN = kT / (CP / 4);
    p->RPT = (p->RB + RGN - 1) / RGN;
    if (p->RPT == 7) p->RPT = 8;
    if (p->RPT > MAXR) return false;
    p->RBP = p->RPT * RGN;
    const int ncg = p->cw / 4;
    p->TC = (ncg * (p->DCB / 4) > kT && p->DCB % 8 == 0) ? 8 : 4;
    p->MAXNEW = (a->n_new + CS - 1) / CS + 1;
    return true;
}

size_t cl_smem_bytes(const sr_head_args* a, const ClParams& p) {
    auto up4 = [](int64_t n) { return (n + 3) & ~(int64_t)3; };
    const int C = a->n_classes, d = a->dim, q = a->q_rows;
    const bool proj = a->pull_mode == SR_PULL_PROJECT && q < d;
    const bool fixed = a->pull_mode == SR_PULL_FIXED;
    const bool has_base = a->base_weight != nullptr;
    const bool has_prev = a->reserve_weight != nullptr && a->n_prev_novel > 0;
    const int n_opt = a->optimizer == SR_OPT_ADAM ? 2 : 1;
    int64_t f = up4((int64_t)p.KB * p.cw) + up4((int64_t)p.SB * p.cw) + up4((int64_t)p.KB * p.RBP) + up4((int64_t)p.RB * p.cw) +
                up4((int64_t)p.RS * p.cw) + up4((int64_t)p.SB * p.DCB) + up4((int64_t)C * p.DCB) + up4((int64_t)p.CB * p.DCB) +
                up4((int64_t)n_opt * p.CB * p.DCB);
    if (has_base) f += up4((int64_t)p.CB * p.DCB);
    if (has_prev) f += up4((int64_t)p.CB * p.DCB);
    if (proj || fixed) f += up4((int64_t)(proj ? q : a->n_new) * p.DCB);
    if (proj) f += 2 * up4((int64_t)a->n_new * q) + up4((int64_t)q * p.KB) + up4((int64_t)p.MAXNEW * p.DCB);
    return (size_t)f * sizeof(float) + 128;
    (void)d;
}

ClLayout cl_layout(const sr_head_args* a, const ClParams& p) {
    ClLayout L;
    int64_t off = 0;
    L.ctrl = off;    off += align_up(sizeof(HeadCtrl), 256);
    L.DL = off;      off += align_up((int64_t)CS * p.SB * p.cw * 4, 256);
    L.Wt = off;      off += align_up((int64_t)a->dim * p.cw * 4, 256);
    L.rowloss = off; off += align_up(2ll * p.NT * 4, 256);
    L.rowhit = off;  off += align_up(2ll * p.NT * 4, 256);
    L.nb = off;      off += align_up(2ll * p.G * 8, 256);
    L.nn = off;      off += align_up(2ll * p.G * 8, 256);
    L.pull = off;    off += align_up((int64_t)p.G * 8, 256);
    L.u = off;       off += align_up((int64_t)std::max(a->n_new, 1) * std::max(a->q_rows, 1) * 4, 256);
    L.total = off;
    return L;
}

int g_max_clusters[kMaxDevices][2];     // per device, per template instance: 0 = not queried yet, -1 = unusable
constexpr int kClSmemMax = 200 * 1024;

template <int CP>
int max_clusters(int dev) {
    int& slot = g_max_clusters[dev][CP == 64 ? 0 : 1];
    if (slot != 0) return slot;
    slot = -1;
    if (cudaFuncSetAttribute(head_cluster_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClSmemMax) != cudaSuccess)
        return slot;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16 * CS);
    cfg.blockDim = dim3(kT);
    cfg.dynamicSmemBytes = kClSmemMax;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, head_cluster_kernel<CP>, &cfg) == cudaSuccess && n > 0) slot = n;
    else cudaGetLastError();
    return slot;
}

// Work clusters for this problem: as many as the device keeps co-resident (cooperative launch), dividing d.
bool cl_plan(const sr_head_args* a, ClParams* p) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
    const int mc = a->n_classes <= 64 ? max_clusters<64>(dev) : max_clusters<128>(dev);
    if (mc < 2) return false;
    static const int cand[] = {16, 10, 8, 5, 4, 2};
    for (int ncl : cand) {
        if (ncl > mc) continue;
        if (!cl_geometry(a, ncl, p)) continue;
        if (cl_smem_bytes(a, *p) > (size_t)kClSmemMax) continue;
        return true;
    }
    return false;
}

}  // namespace

namespace srb {

bool head_cluster_applicable(const sr_head_args* a) {
    // Opt-in (SRB_HEAD_CLUSTER=1): measured on B200 this kernel is on par with head_small, not ahead of it (16.8 / 17.9 /
    // 19.9 us per epoch against 15.9 / 17.1 / 18.8 at the session 1 / 4 / 8 shapes; DESIGN.md section 4.2 has the breakdown).
    const char* e = getenv("SRB_HEAD_CLUSTER");
    if (e == nullptr || atoi(e) == 0) return false;
    const int nt = a->n_support + a->n_memory;
    if (a->logits_support != nullptr) return false;
    if (a->n_classes > 128 || a->n_classes < 8 || a->dim % 64 != 0 || a->dim > 1024 || nt > 1024 || nt < 16) return false;
    if (a->pull_mode == SR_PULL_PROJECT && a->q_rows < a->dim && (a->q_rows > 256 || a->n_new > 16)) return false;
    ClParams p;
    return cl_plan(a, &p);
}

int64_t head_cluster_workspace_bytes(const sr_head_args* a) {
    // (sized for the coarsest plan so that the query does not need a device: SB / cw / G do not depend on the cluster count
    // except G <= 16 * CS)
    ClParams p;
    p.NT = a->n_support + a->n_memory;
    p.SB = (int)align_up((p.NT + CS - 1) / CS, 4);
    p.cw = (a->n_classes + 3) & ~3;
    p.G = 16 * CS;
    return cl_layout(a, p).total;
}

int32_t head_cluster_run(const sr_head_args* a, cudaStream_t stream) {
    ClParams p;
    p.a = *a;
    if (!cl_plan(a, &p)) return fail(SR_E_ARG, "sr_head_run: no cluster plan for this shape");
    p.a = *a;
    const ClLayout L = cl_layout(a, p);
    if (a->workspace_bytes < L.total)
        return fail(SR_E_SMALLWS, "sr_head_run: workspace %lld < %lld", (long long)a->workspace_bytes, (long long)L.total);
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    p.ctrl = reinterpret_cast<HeadCtrl*>(ws + L.ctrl);
    p.DL = reinterpret_cast<float*>(ws + L.DL);
    p.Wt = reinterpret_cast<float*>(ws + L.Wt);
    p.rowloss = reinterpret_cast<float*>(ws + L.rowloss);
    p.rowhit = reinterpret_cast<int*>(ws + L.rowhit);
    p.nb_part = reinterpret_cast<double*>(ws + L.nb);
    p.nn_part = reinterpret_cast<double*>(ws + L.nn);
    p.pull_part = reinterpret_cast<double*>(ws + L.pull);
    p.u = reinterpret_cast<float*>(ws + L.u);
    SR_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)L.total, stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.G);
    cfg.blockDim = dim3(kT);
    cfg.dynamicSmemBytes = cl_smem_bytes(a, p);
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    if (getenv("SRB_HEAD_DBG")) {
        static int once = 0;
        if (!once++)
            fprintf(stderr, "head_cluster plan: NCL %d (grid %d) RB %d RS %d KB %d DCB %d SB %d CB %d cw %d smem %zu B\n", p.NCL,
                    p.G, p.RB, p.RS, p.KB, p.DCB, p.SB, p.CB, p.cw, (size_t)cfg.dynamicSmemBytes);
    }
    if (a->n_classes <= 64) SR_CUDA_OK(cudaLaunchKernelEx(&cfg, head_cluster_kernel<64>, p));
    else SR_CUDA_OK(cudaLaunchKernelEx(&cfg, head_cluster_kernel<128>, p));
    return SR_OK;
}

}  // namespace srb
