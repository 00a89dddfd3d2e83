// Matching heads: fused Linear -> {BN-affine | LayerNorm} -> ReLU -> Linear -> {L2-normalise, dot,
// cosine}.  A thread-block CLUSTER of 8 CTAs serves 8 rows: each CTA computes 1/8 of a layer's output
// columns and broadcasts them through distributed shared memory; warp-shuffle reductions for the
// norms / dots; plus the per-scene softmax/argmax over candidates.
// Reference: models/attribute_module.py:88-90,108-126; models/relation_module.py:82,101-103;
// models/scene_module.py:44-57,84-104; lib/eval_helper.py:61-67 (host argmax of summed scores).
#include "../../include/instancerefer_b200.h"
#include <cooperative_groups.h>

#include "common.cuh"

#define MH_TM 8          // rows per cluster
#define MH_MAXD 256
#define MH_CL 8          // CTAs per cluster: each owns 1/8 of a layer's output columns
#define MH_THREADS 256
#define MH_MAXSL 32      // max K slices per CTA

namespace cg = cooperative_groups;

// One layer for the 8 rows of the tile, columns split over the 8 CTAs of the cluster:
// dst[r][n] = b[n] + sum_k WT[k][n] * src[r][k].  Inside a CTA: thread = (K slice, local column),
// coalesced weight reads (WT is (in,out)), activations broadcast from shared memory, slices combined
// in a fixed order; the finished column block is written into the `dst` buffer of EVERY CTA of the
// cluster through distributed shared memory, so after cluster.sync() each CTA holds the full rows.
__device__ __forceinline__ void cluster_linear(cg::cluster_group& cluster, int rank,
                                               const float (*src)[MH_MAXD], int K,
                                               const float* __restrict__ WT, const float* __restrict__ b, int N,
                                               float (*dst)[MH_MAXD], float (*part)[MH_TM][32]) {
    const bool split = N >= 64;                       // tiny layers (e.g. 9 logits) stay on rank 0
    const int nc = split ? N / MH_CL : N;             // columns owned by this CTA (<= 32)
    const int c0 = split ? rank * nc : 0;
    const bool active = split || rank == 0;
    const int nsl = min(MH_THREADS / nc, MH_MAXSL);
    const int col = threadIdx.x % nc, slice = threadIdx.x / nc;
    if (active && slice < nsl) {
        const int kper = (K + nsl - 1) / nsl;
        const int k0 = slice * kper, k1 = min(K, k0 + kper);
        float acc[MH_TM];
#pragma unroll
        for (int r = 0; r < MH_TM; ++r) acc[r] = 0.f;
        const float* w = WT + c0 + col;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) {
            const float wv = __ldg(w + (long long)k * N);
#pragma unroll
            for (int r = 0; r < MH_TM; ++r) acc[r] = fmaf(wv, src[r][k], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < MH_TM; ++r) part[slice][r][col] = acc[r];
    }
    __syncthreads();
    if (active) {
        for (int i = threadIdx.x; i < MH_TM * nc; i += MH_THREADS) {
            const int r = i / nc, c = i - r * nc;
            float v = b ? b[c0 + c] : 0.f;
            for (int sl = 0; sl < nsl; ++sl) v += part[sl][r][c];
            float* cell = &dst[r][c0 + c];
#pragma unroll
            for (int q = 0; q < MH_CL; ++q) *cluster.map_shared_rank(cell, q) = v;
        }
    }
    cluster.sync();
}

__global__ void __cluster_dims__(MH_CL, 1, 1) __launch_bounds__(MH_THREADS)
k_mlp_head(const float* __restrict__ x, int M, int K, const float* __restrict__ W1T,
           const float* __restrict__ b1, int N1, int norm, const float* __restrict__ g,
           const float* __restrict__ beta, const float* __restrict__ W2T, const float* __restrict__ b2,
           int N2, int mode, const float* __restrict__ partner, const int* __restrict__ seg,
           float* __restrict__ y, float* __restrict__ score) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ float xs[MH_TM][MH_MAXD];
    __shared__ float hs[MH_TM][MH_MAXD];
    __shared__ float part[MH_MAXSL][MH_TM][32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = (blockIdx.x / MH_CL) * MH_TM;
    const int rows = min(MH_TM, M - r0);
    for (int i = tid; i < MH_TM * K; i += MH_THREADS) {
        const int r = i / K, k = i - r * K;
        xs[r][k] = (r < rows) ? x[(long long)(r0 + r) * K + k] : 0.f;
    }
    __syncthreads();
    cluster_linear(cluster, rank, xs, K, W1T, b1, N1, hs, part);
    {   // normalisation + ReLU on the full rows (every CTA, redundantly): warp w owns row w
        const int r = w;
        if (norm == 2) {
            float s = 0.f;
            for (int n = lane; n < N1; n += 32) s += hs[r][n];
            const float mean = warp_sum(s) / (float)N1;
            float v = 0.f;
            for (int n = lane; n < N1; n += 32) { const float d = hs[r][n] - mean; v = fmaf(d, d, v); }
            const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)N1 + 1e-5f);
            for (int n = lane; n < N1; n += 32)
                hs[r][n] = fmaxf(fmaf((hs[r][n] - mean) * rstd, g[n], beta[n]), 0.f);
        } else if (norm == 1) {
            for (int n = lane; n < N1; n += 32) hs[r][n] = fmaxf(fmaf(hs[r][n], g[n], beta[n]), 0.f);
        } else {
            for (int n = lane; n < N1; n += 32) hs[r][n] = fmaxf(hs[r][n], 0.f);
        }
    }
    __syncthreads();
    cluster_linear(cluster, rank, hs, N1, W2T, b2, N2, xs, part);      // xs now holds y (all CTAs)
    if (rank != 0) return;
    const int r = w;
    if (r >= rows) return;
    const long long row = r0 + r;
    if (mode == 0) {
        for (int n = lane; n < N2; n += 32) y[row * N2 + n] = xs[r][n];
        return;
    }
    float ss = 0.f;
    for (int n = lane; n < N2; n += 32) ss = fmaf(xs[r][n], xs[r][n], ss);
    const float nrm = sqrtf(warp_sum(ss));
    if (mode == 1) {
        const float inv = 1.0f / fmaxf(nrm, 1e-12f);
        for (int n = lane; n < N2; n += 32) y[row * N2 + n] = xs[r][n] * inv;
        return;
    }
    const float* p = partner + (long long)seg[row] * N2;
    if (y) for (int n = lane; n < N2; n += 32) y[row * N2 + n] = xs[r][n];
    if (mode == 2) {
        const float inv = 1.0f / fmaxf(nrm, 1e-12f);
        float d = 0.f;
        for (int n = lane; n < N2; n += 32) d = fmaf(xs[r][n] * inv, p[n], d);
        d = warp_sum(d);
        if (lane == 0) score[row] = d;
    } else {
        float d = 0.f, pp = 0.f;
        for (int n = lane; n < N2; n += 32) { d = fmaf(xs[r][n], p[n], d); pp = fmaf(p[n], p[n], pp); }
        d = warp_sum(d);
        pp = sqrtf(warp_sum(pp));
        if (lane == 0) score[row] = d / (fmaxf(nrm, 1e-8f) * fmaxf(pp, 1e-8f));
    }
}

extern "C" int ir_mlp_head(const float* x, int32_t M, int32_t K, const float* W1T, const float* b1,
                           int32_t N1, int32_t norm, const float* g, const float* beta,
                           const float* W2T, const float* b2, int32_t N2, int32_t mode,
                           const float* partner, const int32_t* seg, float* y, float* score,
                           ir_stream_t stream) {
    IR_CHECK_ARG(x && W1T && W2T && M > 0 && K > 0 && K <= MH_MAXD && N1 > 0 && N1 <= MH_MAXD && N2 > 0 && N2 <= MH_MAXD);
    IR_CHECK_ARG((N1 < 64 ? N1 <= 32 : N1 % MH_CL == 0) && (N2 < 64 ? N2 <= 32 : N2 % MH_CL == 0));
    IR_CHECK_ARG(norm >= 0 && norm <= 2 && mode >= 0 && mode <= 3);
    IR_CHECK_ARG(norm == 0 || (g && beta));
    IR_CHECK_ARG(mode >= 2 ? (partner && seg && score) : (y != nullptr));
    k_mlp_head<<<ir_div_up(M, MH_TM) * MH_CL, MH_THREADS, 0, (cudaStream_t)stream>>>(
        x, M, K, W1T, b1, N1, norm, g, beta, W2T, b2, N2, mode, partner, seg, y, score);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ per-scene candidate softmax
__global__ void k_candidate_softmax(const float* __restrict__ sa, const float* __restrict__ sr,
                                    const float* __restrict__ ss, const int* __restrict__ seg_ofs,
                                    float* __restrict__ prob, int* __restrict__ argmax) {
    const int s = blockIdx.x, lane = threadIdx.x;
    const int a = seg_ofs[s], b = seg_ofs[s + 1];
    float m = -INFINITY;
    int mi = 0x7FFFFFFF;
    for (int i = a + lane; i < b; i += 32) {
        const float v = sa[i] + sr[i] + ss[i];
        if (v > m) { m = v; mi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    float z = 0.f;
    for (int i = a + lane; i < b; i += 32) z += expf((sa[i] + sr[i] + ss[i]) - m);
    z = warp_sum(z);
    for (int i = a + lane; i < b; i += 32) prob[i] = expf((sa[i] + sr[i] + ss[i]) - m) / z;
    if (lane == 0) argmax[s] = (b > a) ? (mi - a) : -1;
}

extern "C" int ir_candidate_softmax(const float* s_attr, const float* s_rel, const float* s_scene,
                                    const int32_t* seg_ofs, int32_t n_seg, float* prob,
                                    int32_t* argmax, ir_stream_t stream) {
    IR_CHECK_ARG(s_attr && s_rel && s_scene && seg_ofs && prob && argmax && n_seg > 0);
    k_candidate_softmax<<<n_seg, 32, 0, (cudaStream_t)stream>>>(s_attr, s_rel, s_scene, seg_ofs, prob, argmax);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
