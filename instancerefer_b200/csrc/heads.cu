// Matching heads: fused Linear -> {BN-affine | LayerNorm} -> ReLU -> Linear -> {L2-normalise, dot,
// cosine} with warp-shuffle reductions, plus the per-scene softmax/argmax over candidates.
// Reference: models/attribute_module.py:88-90,108-126; models/relation_module.py:82,101-103;
// models/scene_module.py:44-57,84-104; lib/eval_helper.py:61-67 (host argmax of summed scores).
#include "../../include/instancerefer_b200.h"
#include "common.cuh"

#define MH_TM 8
#define MH_MAXD 256

// rows [r0, r0+8) of x (M,K): dst[r][n] = b[n] + sum_k W[n][k] * src[r][k]   (8 warps over n)
__device__ __forceinline__ void tile_linear(const float (*src)[MH_MAXD], int K, const float* __restrict__ W,
                                            const float* __restrict__ b, int N, float (*dst)[MH_MAXD]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int n = w; n < N; n += 8) {
        float acc[MH_TM];
#pragma unroll
        for (int r = 0; r < MH_TM; ++r) acc[r] = 0.f;
        const float* wr = W + (long long)n * K;
        for (int k = lane; k < K; k += 32) {
            const float wv = wr[k];
#pragma unroll
            for (int r = 0; r < MH_TM; ++r) acc[r] = fmaf(wv, src[r][k], acc[r]);
        }
        const float bv = b ? b[n] : 0.f;
#pragma unroll
        for (int r = 0; r < MH_TM; ++r) {
            const float v = warp_sum(acc[r]);
            if (lane == r) dst[r][n] = v + bv;
        }
    }
}

__global__ void __launch_bounds__(256)
k_mlp_head(const float* __restrict__ x, int M, int K, const float* __restrict__ W1,
           const float* __restrict__ b1, int N1, int norm, const float* __restrict__ g,
           const float* __restrict__ beta, const float* __restrict__ W2, const float* __restrict__ b2,
           int N2, int mode, const float* __restrict__ partner, const int* __restrict__ seg,
           float* __restrict__ y, float* __restrict__ score) {
    __shared__ float xs[MH_TM][MH_MAXD];
    __shared__ float hs[MH_TM][MH_MAXD];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = blockIdx.x * MH_TM;
    const int rows = min(MH_TM, M - r0);
    for (int i = tid; i < MH_TM * K; i += 256) {
        const int r = i / K, k = i - r * K;
        xs[r][k] = (r < rows) ? x[(long long)(r0 + r) * K + k] : 0.f;
    }
    __syncthreads();
    tile_linear(xs, K, W1, b1, N1, hs);
    __syncthreads();
    {   // normalisation + ReLU: warp w owns row w
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
    tile_linear(hs, N1, W2, b2, N2, xs);      // xs now holds y
    __syncthreads();
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

extern "C" int ir_mlp_head(const float* x, int32_t M, int32_t K, const float* W1, const float* b1,
                           int32_t N1, int32_t norm, const float* g, const float* beta,
                           const float* W2, const float* b2, int32_t N2, int32_t mode,
                           const float* partner, const int32_t* seg, float* y, float* score,
                           ir_stream_t stream) {
    IR_CHECK_ARG(x && W1 && W2 && M > 0 && K > 0 && K <= MH_MAXD && N1 > 0 && N1 <= MH_MAXD && N2 > 0 && N2 <= MH_MAXD);
    IR_CHECK_ARG(norm >= 0 && norm <= 2 && mode >= 0 && mode <= 3);
    IR_CHECK_ARG(norm == 0 || (g && beta));
    IR_CHECK_ARG(mode >= 2 ? (partner && seg && score) : (y != nullptr));
    k_mlp_head<<<ir_div_up(M, MH_TM), 256, 0, (cudaStream_t)stream>>>(x, M, K, W1, b1, N1, norm, g, beta, W2, b2,
                                                                      N2, mode, partner, seg, y, score);
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
