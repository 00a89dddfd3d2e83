// Sparse convolution = pair-GEMM (gather rows -> per-offset dense contraction -> T) followed by a
// deterministic output-stationary reduce with the BatchNorm / residual / ReLU epilogue fused in.
// Replaces torchsparse's sparseconv_forward host loop reached from models/basic_blocks.py:14-21,
// 32-44 (spnn.Conv3d + spnn.BatchNorm + spnn.ReLU) and the `+` at :55.
//
// This file: the SIMT fp32 pair-GEMM (exact fp32 FMA chain; the reference kernel the tcgen05
// version in spconv_tc.cu is validated against), the reduce/epilogue, segmented max-pool.
//
//   T[kofs[k] + pos, :] = F[in_idx[k][pos], :] @ W[k]            (pair-GEMM, weight-stationary)
//   out[o, :] = act( scale * sum_{k asc} T[kofs[k] + slot[k][o], :] + shift (+ resid[o, :]) )
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

extern int g_tune_reduce_ctas;      // spconv_tc.cu (ir_tune_set)
#define PG_TP 32
#define PG_THREADS 128
#define PG_MAXC 128

template <int COUT>
__global__ void __launch_bounds__(PG_THREADS)
k_pairgemm_simt(IrConvBatch b, int cin, int K) {
    const IrConvProblem& P = b.p[blockIdx.y];
    const float* __restrict__ F = P.fin;
    const int* __restrict__ in_idx = P.in_idx;
    const long long seg_cap = P.seg_cap;
    const int* __restrict__ count = P.count;
    const float* __restrict__ W = P.weight;
    float* __restrict__ T = P.T;
    constexpr int G = PG_THREADS / COUT;      // pair groups per block
    constexpr int PPT = PG_TP / G;            // pairs per thread
    __shared__ __align__(16) float As[PG_TP][PG_MAXC];
    __shared__ int s_kofs[33], s_tofs[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {                     // pair / tile prefixes, one lane per kernel offset
        const int c = (lane < K) ? count[lane] : 0;
        const int t = (c + PG_TP - 1) / PG_TP;
        int cinc = c, tinc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int c2 = __shfl_up_sync(0xffffffffu, cinc, o), t2 = __shfl_up_sync(0xffffffffu, tinc, o);
            if (lane >= o) { cinc += c2; tinc += t2; }
        }
        s_kofs[lane] = cinc - c; s_tofs[lane] = tinc - t;
        if (lane == 31) { s_kofs[32] = cinc; s_tofs[32] = tinc; }
    }
    __syncthreads();
    ir_pdl_trigger();
    ir_pdl_wait();                        // the input features come from the previous kernel
    const int ntiles = s_tofs[32];
    const int cp = (cin + 3) & ~3;
    const int co = tid % COUT, g = tid / COUT;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int k = 0;
        while (k + 1 < K && tile >= s_tofs[k + 1]) ++k;
        const int p0 = (tile - s_tofs[k]) * PG_TP;
        const int np = min(PG_TP, ((k + 1 < 32 ? s_kofs[k + 1] : s_kofs[32]) - s_kofs[k]) - p0);
        for (int r = warp; r < PG_TP; r += PG_THREADS / 32) {
            const int j = (r < np) ? in_idx[(long long)k * seg_cap + p0 + r] : -1;
            for (int c = lane; c < cp; c += 32)
                As[r][c] = (j >= 0 && c < cin) ? F[(long long)j * cin + c] : 0.f;
        }
        __syncthreads();
        float acc[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) acc[p] = 0.f;
        const float* Wk = W + (long long)k * cin * COUT + co;
        for (int c4 = 0; c4 < cp; c4 += 4) {
            const float w0 = Wk[(long long)(c4 + 0) * COUT];
            const float w1 = (c4 + 1 < cin) ? Wk[(long long)(c4 + 1) * COUT] : 0.f;
            const float w2 = (c4 + 2 < cin) ? Wk[(long long)(c4 + 2) * COUT] : 0.f;
            const float w3 = (c4 + 3 < cin) ? Wk[(long long)(c4 + 3) * COUT] : 0.f;
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const float4 a = *reinterpret_cast<const float4*>(&As[g * PPT + p][c4]);
                acc[p] = fmaf(a.x, w0, acc[p]);
                acc[p] = fmaf(a.y, w1, acc[p]);
                acc[p] = fmaf(a.z, w2, acc[p]);
                acc[p] = fmaf(a.w, w3, acc[p]);
            }
        }
        float* Trow = T + (long long)(s_kofs[k] + p0 + g * PPT) * COUT + co;
#pragma unroll
        for (int p = 0; p < PPT; ++p)
            if (g * PPT + p < np) Trow[(long long)p * COUT] = acc[p];
        __syncthreads();
    }
}

int irk_pairgemm_simt(const IrConvBatch& b, int cin, int cout, int K, cudaStream_t st) {
    IR_CHECK_ARG(cin >= 1 && cin <= PG_MAXC && K <= 32 && b.G >= 1 && b.G <= IR_MAX_GROUPS);
    long long tiles = 0;
    for (int g = 0; g < b.G; ++g) tiles = tiles > ((long long)K * b.p[g].n_max / PG_TP + K) ? tiles : ((long long)K * b.p[g].n_max / PG_TP + K);
    const dim3 grid(ir_min_i(tiles > 0 ? tiles : 1, IR_NUM_SMS * 16), b.G);
    switch (cout) {
        case 32:  IR_CHECK_CUDA(ir_launch_pdl(k_pairgemm_simt<32>, grid, dim3(PG_THREADS), 0, st, b, cin, K)); break;
        case 64:  IR_CHECK_CUDA(ir_launch_pdl(k_pairgemm_simt<64>, grid, dim3(PG_THREADS), 0, st, b, cin, K)); break;
        case 128: IR_CHECK_CUDA(ir_launch_pdl(k_pairgemm_simt<128>, grid, dim3(PG_THREADS), 0, st, b, cin, K)); break;
        default: ir_set_error("pairgemm_simt: unsupported cout %d", cout); return IR_ERR_UNSUPPORTED;
    }
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ reduce + epilogue
template <int COUT>
__global__ void __launch_bounds__(256)
k_reduce_epilogue(IrConvBatch b, int K) {
    const IrConvProblem& P = b.p[blockIdx.y];
    const float* __restrict__ T = P.T;
    const int* __restrict__ slot = P.slot;
    const long long seg_cap = P.seg_cap;
    const int* __restrict__ count = P.count;
    const int* __restrict__ n_dev = P.n_out_dev;
    const float* __restrict__ scale = P.scale;
    const float* __restrict__ shift = P.shift;
    const float* __restrict__ resid = P.resid;
    const int relu = P.relu;
    float* __restrict__ out = P.out;
    constexpr int V = COUT / 32;
    __shared__ int s_kofs[32];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) {
        const int v = (lane < K) ? count[lane] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        s_kofs[lane] = inc - v;
    }
    __syncthreads();
    ir_pdl_trigger();                     // the next layer's pair-GEMM may start staging its weights
    ir_pdl_wait();                        // T comes from the pair-GEMM launched just before
    ir_stamp_begin(b.stamp);
    const int n = *n_dev;
    float sc[V], sh[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        sc[v] = scale ? scale[lane * V + v] : 1.f;
        sh[v] = shift ? shift[lane * V + v] : 0.f;
    }
    const int wpb = blockDim.x >> 5;
    float amax = 0.f;
    // (prefetching the next row's pair positions was measured slower inside the step: 0.494 vs 0.477 ms)
    for (long long o = (long long)blockIdx.x * wpb + (tid >> 5); o < n; o += (long long)gridDim.x * wpb) {
        const int my = (lane < K) ? slot[(long long)lane * seg_cap + o] : -1;
        // present offsets of this row, ascending k, compacted into lanes 0..cnt-1 (lane L holds the T row of the L-th
        // present offset): only those rows are loaded, U at a time; missing pairs cost nothing and the order of the
        // additions stays fixed (bitwise deterministic)
        const unsigned present = __ballot_sync(0xffffffffu, my >= 0);
        const int cnt = __popc(present);
        const int my_row = ((lane < K) ? s_kofs[lane] : 0) + (my >= 0 ? my : 0);
        const int crow = __shfl_sync(0xffffffffu, my_row, (lane < cnt) ? (int)__fns(present, 0, lane + 1) : 0);
        float acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0.f;
        constexpr int U = 9;
        for (int j0 = 0; j0 < cnt; j0 += U) {
            float t[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = j0 + u < cnt;
                const int trow = __shfl_sync(0xffffffffu, crow, (j0 + u) & 31);
                const float* row = T + (long long)trow * COUT + lane * V;
                if (V == 4) {
                    const float4 q = ok ? *reinterpret_cast<const float4*>(row) : make_float4(0.f, 0.f, 0.f, 0.f);
                    t[u][0] = q.x; t[u][1 % V] = q.y; t[u][2 % V] = q.z; t[u][3 % V] = q.w;
                } else if (V == 2) {
                    const float2 q = ok ? *reinterpret_cast<const float2*>(row) : make_float2(0.f, 0.f);
                    t[u][0] = q.x; t[u][1 % V] = q.y;
                } else {
                    t[u][0] = ok ? row[0] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += t[u][v];
        }
        float* orow = out + o * COUT + lane * V;
        const float* rrow = resid ? resid + o * COUT + lane * V : nullptr;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float y = fmaf(acc[v], sc[v], sh[v]);
            if (rrow) y += rrow[v];
            if (relu) y = fmaxf(y, 0.f);
            acc[v] = y;
            amax = fmaxf(amax, fabsf(y));
        }
        if (V == 4) *reinterpret_cast<float4*>(orow) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
        else if (V == 2) *reinterpret_cast<float2*>(orow) = make_float2(acc[0], acc[1 % V]);
        else orow[0] = acc[0];
    }
    if (P.out_absmax) {                   // range guard of the next layer's split-fp16 gather (non-negative floats order as uints)
        amax = warp_max(amax);
        if (lane == 0 && amax > 0.f) atomicMax(reinterpret_cast<unsigned*>(P.out_absmax), __float_as_uint(amax));
    }
    __syncthreads();
    ir_stamp_end(b.stamp);
}

int irk_reduce_epilogue(const IrConvBatch& b, int cout, int K, cudaStream_t st) {
    IR_CHECK_ARG(K <= 32 && b.G >= 1 && b.G <= IR_MAX_GROUPS);
    long long rows = 1;
    for (int g = 0; g < b.G; ++g) rows = rows > b.p[g].n_max ? rows : b.p[g].n_max;
    const dim3 grid(ir_min_i(ir_div_up(rows, 8), rows <= 8192 ? (g_tune_reduce_ctas + 1) / 2 : g_tune_reduce_ctas), b.G);
    switch (cout) {
        case 32: IR_CHECK_CUDA(ir_launch_pdl(k_reduce_epilogue<32>, grid, dim3(256), 0, st, b, K)); break;
        case 64: IR_CHECK_CUDA(ir_launch_pdl(k_reduce_epilogue<64>, grid, dim3(256), 0, st, b, K)); break;
        case 128: IR_CHECK_CUDA(ir_launch_pdl(k_reduce_epilogue<128>, grid, dim3(256), 0, st, b, K)); break;
        default: ir_set_error("reduce: unsupported cout %d", cout); return IR_ERR_UNSUPPORTED;
    }
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ fused stem (Cin <= 8 -> 32, k3)
// The stem has 7 input channels: a pair-GEMM + reduce round trip through T costs more than the
// arithmetic.  Direct output-stationary form instead: one warp per output row, lane = output channel,
// the 27 x Cin x 32 weights in shared memory, neighbours resolved through the same rulebook
// (slot -> in_idx), folded BN + ReLU in the epilogue.  Replaces stem.0 of models/basic_blocks.py:64-66.
#define STEM_COUT 32
#define STEM_MAXCIN 8
#define STEM_WARPS 8
// Per output row (one warp): (1) lanes < 27 resolve the neighbour rows through the rulebook (slot -> in_idx), the next
// row's lookups are issued before this row's arithmetic; (2) the present neighbours are compacted (ascending k) and ALL
// their feature values (<= 27 x Cin <= 216) are fetched by ONE round of independent loads, 7 per lane, into the warp's
// shared-memory slab; (3) lane = output channel accumulates over present neighbours only, in ascending k (deterministic),
// reading the features as shared-memory broadcasts.  Three dependent memory latencies per row instead of one per neighbour.
__global__ void __launch_bounds__(32 * STEM_WARPS)
k_stem_direct(IrConvBatch b, int cin) {
    const IrConvProblem& P = b.p[blockIdx.y];
    __shared__ float Ws[27 * STEM_MAXCIN * STEM_COUT];
    __shared__ float Xs[STEM_WARPS][27 * STEM_MAXCIN + 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 27 * cin * STEM_COUT; i += 32 * STEM_WARPS) Ws[i] = P.weight[i];     // (27, cin, 32), constant
    __syncthreads();
    ir_pdl_trigger();
    ir_pdl_wait();
    ir_stamp_begin(b.stamp);
    const int n = *P.n_out_dev;
    const float sc = P.scale ? P.scale[lane] : 1.f, sh = P.shift ? P.shift[lane] : 0.f;
    const int* __restrict__ slot = P.slot + (long long)lane * P.seg_cap;
    const int* __restrict__ in_idx = P.in_idx + (long long)lane * P.seg_cap;
    const float* __restrict__ F = P.fin;
    float* xs = Xs[warp];
    const long long stride = (long long)gridDim.x * STEM_WARPS;
    float amax = 0.f;
    long long o = (long long)blockIdx.x * STEM_WARPS + warp;
    int my_j = -1;
    if (o < n && lane < 27) {
        const int pos = slot[o];
        if (pos >= 0) my_j = in_idx[pos];
    }
    for (; o < n; o += stride) {
        const int cur_j = my_j;
        // next row's rulebook lookups (two dependent loads) fly while this row is computed
        int pos_n = -1;
        if (o + stride < n && lane < 27) pos_n = slot[o + stride];
        const unsigned present = __ballot_sync(0xffffffffu, cur_j >= 0);
        const int cnt = __popc(present);
        // lane L takes over (k, row) of the L-th present neighbour
        const int src = (lane < cnt) ? (int)__fns(present, 0, lane + 1) : 0;
        const int cj = __shfl_sync(0xffffffffu, cur_j, src);
        const int ck = src;
        // one round of independent loads: element e = (neighbour e / cin, channel e % cin)
        const int total = cnt * cin;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int e = lane + 32 * r;
            const int nb = e / cin, ci = e - nb * cin;
            const int jr = __shfl_sync(0xffffffffu, cj, nb & 31);
            if (e < total) xs[e] = __ldg(F + (long long)jr * cin + ci);
        }
        my_j = (pos_n >= 0) ? in_idx[pos_n] : -1;
        __syncwarp();
        float acc = 0.f;
        for (int nb = 0; nb < cnt; ++nb) {
            const int k = __shfl_sync(0xffffffffu, ck, nb);
            const float* w = Ws + (k * cin) * STEM_COUT + lane;
            const float* x = xs + nb * cin;
            for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], w[ci * STEM_COUT], acc);
        }
        __syncwarp();
        float y = fmaf(acc, sc, sh);
        if (P.resid) y += P.resid[o * STEM_COUT + lane];
        if (P.relu) y = fmaxf(y, 0.f);
        P.out[o * STEM_COUT + lane] = y;
        amax = fmaxf(amax, fabsf(y));
    }
    if (P.out_absmax) {
        amax = warp_max(amax);
        if (lane == 0 && amax > 0.f) atomicMax(reinterpret_cast<unsigned*>(P.out_absmax), __float_as_uint(amax));
    }
    __syncthreads();
    ir_stamp_end(b.stamp);
}

int irk_stem_direct(const IrConvBatch& b, int cin, cudaStream_t st) {
    IR_CHECK_ARG(cin >= 1 && cin <= STEM_MAXCIN && b.G >= 1 && b.G <= IR_MAX_GROUPS);
    long long rows = 1;
    for (int g = 0; g < b.G; ++g) rows = rows > b.p[g].n_max ? rows : b.p[g].n_max;
    const dim3 grid(ir_min_i(ir_div_up(rows, 8), rows <= 8192 ? (g_tune_reduce_ctas + 1) / 2 : g_tune_reduce_ctas), b.G);
    IR_CHECK_CUDA(ir_launch_pdl(k_stem_direct, grid, dim3(256), 0, st, b, cin));
    IR_CHECK_LAUNCH();
    return IR_OK;
}
