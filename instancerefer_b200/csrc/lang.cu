// Language encoder kernels: small-M linear (warp-shuffle reductions), packed bidirectional GRU layer
// with the recurrent matrix resident in shared memory, masked token-attention pooling.
// Reference: models/lang_module.py:22-37,51-108.
#include "../../include/instancerefer_b200.h"
#include "common.cuh"

// ------------------------------------------------------------------ y = act(x W^T + b), small M
#define LIN_TM 8
#define LIN_MAXK 512
__global__ void __launch_bounds__(256)
k_linear(const float* __restrict__ x, int M, int K, const float* __restrict__ W,
         const float* __restrict__ b, int N, int relu, float* __restrict__ y) {
    __shared__ float xs[LIN_TM][LIN_MAXK];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = blockIdx.y * LIN_TM;
    const int rows = min(LIN_TM, M - r0);
    for (int i = tid; i < LIN_TM * K; i += 256) {
        const int r = i / K, k = i - r * K;
        xs[r][k] = (r < rows) ? x[(long long)(r0 + r) * K + k] : 0.f;
    }
    __syncthreads();
    const int n = blockIdx.x * 8 + w;
    if (n >= N) return;
    float acc[LIN_TM];
#pragma unroll
    for (int r = 0; r < LIN_TM; ++r) acc[r] = 0.f;
    const float* wr = W + (long long)n * K;
    for (int k = lane; k < K; k += 32) {
        const float wv = wr[k];
#pragma unroll
        for (int r = 0; r < LIN_TM; ++r) acc[r] = fmaf(wv, xs[r][k], acc[r]);
    }
    const float bv = b ? b[n] : 0.f;
#pragma unroll
    for (int r = 0; r < LIN_TM; ++r) {
        float v = warp_sum(acc[r]) + bv;
        if (relu) v = fmaxf(v, 0.f);
        if (lane == 0 && r < rows) y[(long long)(r0 + r) * N + n] = v;
    }
}

extern "C" int ir_linear(const float* x, int32_t M, int32_t K, const float* W, const float* b,
                         int32_t N, int32_t relu, float* y, ir_stream_t stream) {
    IR_CHECK_ARG(x && W && y && M > 0 && N > 0 && K > 0 && K <= LIN_MAXK);
    dim3 grid(ir_div_up(N, 8), ir_div_up(M, LIN_TM));
    k_linear<<<grid, 256, 0, (cudaStream_t)stream>>>(x, M, K, W, b, N, relu, y);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ GRU layer (both directions)
// grid (B, 2 directions); 2*3H threads.  Thread (gate row j, half) keeps its 64 recurrent weights in
// REGISTERS for the whole sequence; the hidden state lives in shared memory (padded so the two
// halves read different banks); partial dot products are combined with one warp shuffle.
#define GRU_H 128
#define GRU_HALF 64
#define GRU_PAD 4
__global__ void __launch_bounds__(6 * GRU_H)
k_gru_layer(const float* __restrict__ xproj, const float* __restrict__ whh, const float* __restrict__ bhh,
            const long long* __restrict__ lengths, int L, float* __restrict__ out) {
    constexpr int H = GRU_H, G = 3 * GRU_H;
    __shared__ __align__(16) float s_h[2 * (GRU_HALF + GRU_PAD)];
    __shared__ float s_hp[G];
    const int b = blockIdx.x, dir = blockIdx.y, t = threadIdx.x;
    const int j = t >> 1, half = t & 1;
    float w[GRU_HALF];
    {
        const float4* src = reinterpret_cast<const float4*>(whh + ((size_t)dir * G + j) * H + half * GRU_HALF);
#pragma unroll
        for (int i = 0; i < GRU_HALF / 4; ++i) {
            const float4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
    const float bj = bhh[dir * G + j];
    if (t < H) s_h[(t / GRU_HALF) * (GRU_HALF + GRU_PAD) + (t % GRU_HALF)] = 0.f;
    int len = (int)lengths[b];
    len = max(0, min(len, L));
    __syncthreads();
    const float4* hv = reinterpret_cast<const float4*>(s_h + half * (GRU_HALF + GRU_PAD));
    // input projections of the current step live in registers; the next step's are prefetched while the
    // recurrent matvec runs (they do not depend on h)
    float xr = 0.f, xz = 0.f, xn = 0.f;
    if (t < H && len > 0) {
        const float* xp = xproj + (((size_t)b * L + (dir ? len - 1 : 0)) * 2 + dir) * G;
        xr = xp[t]; xz = xp[H + t]; xn = xp[2 * H + t];
    }
    for (int s = 0; s < len; ++s) {
        const int tt = dir ? (len - 1 - s) : s;
        float nr = 0.f, nz = 0.f, nn = 0.f;
        if (t < H && s + 1 < len) {
            const float* xp = xproj + (((size_t)b * L + (dir ? tt - 1 : tt + 1)) * 2 + dir) * G;
            nr = xp[t]; nz = xp[H + t]; nn = xp[2 * H + t];
        }
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < GRU_HALF / 4; ++i) {
            const float4 h4 = hv[i];
            a0 = fmaf(w[4 * i], h4.x, a0);
            a1 = fmaf(w[4 * i + 1], h4.y, a1);
            a0 = fmaf(w[4 * i + 2], h4.z, a0);
            a1 = fmaf(w[4 * i + 3], h4.w, a1);
        }
        float acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0) s_hp[j] = acc + bj;
        __syncthreads();
        if (t < H) {
            const int hi = (t / GRU_HALF) * (GRU_HALF + GRU_PAD) + (t % GRU_HALF);
            const float r = 1.f / (1.f + expf(-(xr + s_hp[t])));
            const float z = 1.f / (1.f + expf(-(xz + s_hp[H + t])));
            const float n = tanhf(xn + r * s_hp[2 * H + t]);
            const float h = (1.f - z) * n + z * s_h[hi];
            out[((size_t)b * L + tt) * (2 * H) + dir * H + t] = h;
            s_h[hi] = h;
            xr = nr; xz = nz; xn = nn;
        }
        __syncthreads();
    }
    if (t < H)
        for (int tt = len; tt < L; ++tt) out[((size_t)b * L + tt) * (2 * H) + dir * H + t] = 0.f;
}

extern "C" int ir_gru_layer(const float* xproj, const float* whh, const float* bhh,
                            const int64_t* lengths, int32_t B, int32_t L, int32_t H, float* out,
                            ir_stream_t stream) {
    IR_CHECK_ARG(xproj && whh && bhh && lengths && out && B > 0 && L > 0 && H == GRU_H);
    k_gru_layer<<<dim3(B, 2), 6 * GRU_H, 0, (cudaStream_t)stream>>>(xproj, whh, bhh, (const long long*)lengths, L, out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ token attention pooling (x4)
#define TA_MAXL 128
__global__ void __launch_bounds__(256)
k_token_attention(const float* __restrict__ feats, const float* __restrict__ embed,
                  long long embed_stride, const long long* __restrict__ lengths,
                  const float* __restrict__ fcw, const float* __restrict__ fcb, int B, int L, int D,
                  int E, float* __restrict__ atten, float* __restrict__ pooled) {
    __shared__ float s_a[4][TA_MAXL];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int len = (int)lengths[b];
    for (int p = w; p < 4 * L; p += 8) {
        const int h = p / L, t = p - h * L;
        const float* f = feats + ((long long)b * L + t) * D;
        float a = 0.f;
        for (int d = lane; d < D; d += 32) a = fmaf(f[d], fcw[h * D + d], a);
        a = warp_sum(a);
        if (lane == 0) s_a[h][t] = a + fcb[h];
    }
    __syncthreads();
    if (w < 4) {
        const int h = w;
        float m = -INFINITY;
        for (int t = lane; t < L; t += 32) m = fmaxf(m, s_a[h][t]);
        m = warp_max(m);
        float s = 0.f;
        for (int t = lane; t < L; t += 32) {
            const float e = expf(s_a[h][t] - m);
            s_a[h][t] = e;
            s += e;
        }
        s = warp_sum(s);
        float s2 = 0.f;
        for (int t = lane; t < L; t += 32) {
            const float a = (s_a[h][t] / s) * ((t < len) ? 1.f : 0.f);
            s_a[h][t] = a;
            s2 += a;
        }
        s2 = warp_sum(s2);
        for (int t = lane; t < L; t += 32) {
            const float a = s_a[h][t] / s2;
            s_a[h][t] = a;
            atten[((long long)h * B + b) * L + t] = a;
        }
    }
    __syncthreads();
    for (int e = tid; e < E; e += 256) {
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int t = 0; t < L; ++t) {
            const float v = embed[(long long)b * embed_stride + (long long)t * E + e];
            acc0 = fmaf(s_a[0][t], v, acc0);
            acc1 = fmaf(s_a[1][t], v, acc1);
            acc2 = fmaf(s_a[2][t], v, acc2);
            acc3 = fmaf(s_a[3][t], v, acc3);
        }
        pooled[((long long)0 * B + b) * E + e] = acc0;
        pooled[((long long)1 * B + b) * E + e] = acc1;
        pooled[((long long)2 * B + b) * E + e] = acc2;
        pooled[((long long)3 * B + b) * E + e] = acc3;
    }
}

extern "C" int ir_token_attention(const float* feats, const float* embed, int64_t embed_stride,
                                  const int64_t* lengths, const float* fcw, const float* fcb,
                                  int32_t B, int32_t L, int32_t D, int32_t E, float* atten,
                                  float* pooled, ir_stream_t stream) {
    IR_CHECK_ARG(feats && embed && lengths && fcw && fcb && atten && pooled);
    IR_CHECK_ARG(B > 0 && L > 0 && L <= TA_MAXL && D > 0 && E > 0);
    k_token_attention<<<B, 256, 0, (cudaStream_t)stream>>>(feats, embed, embed_stride, (const long long*)lengths,
                                                          fcw, fcb, B, L, D, E, atten, pooled);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
