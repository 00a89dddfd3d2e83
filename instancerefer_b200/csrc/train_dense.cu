// Dense training-step kernels: everything the reference's train-mode forward
// and torch autograd do outside the sparse encoders — Linear / LayerNorm / Dropout / normalise-and-match
// heads (models/attribute_module.py:88-126, relation_module.py:82-103, scene_module.py:44-104), the
// Conv2d pair as im2col + GEMM (scene_module.py:33-38), BEV densification backward
// (basic_blocks.py:195-243), language-guided attention backward (scene_module.py:73-83), the packed
// biGRU and token-attention backward (lang_module.py:51-93) and the EdgeConv edge gather / max
// (basic_blocks.py:98-133).  All fp32 SIMT: these operators are a few MFLOP each and latency-bound.
#include <math.h>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"

// ------------------------------------------------------------------ GEMM  C = op(A) op(B) (+bias)(relu)(+C)
// A(m,k) = ta ? A[k*lda+m] : A[m*lda+k];   B(k,n) = tb ? B[n*ldb+k] : B[k*ldb+n].
// 64x64x16 tiles, 256 threads, thread (ty,tx) owns rows ty+16i, cols tx+16j (conflict-free smem reads,
// 64-byte coalesced stores).
#define GM_BM 64
#define GM_BN 64
#define GM_BK 32
#define GM_LD (GM_BK * GM_BM / 256)     // operand elements per thread and tile (8)
__global__ void __launch_bounds__(256)
k_gemm(int M, int N, int K, const float* __restrict__ A, int lda, int ta, const float* __restrict__ B, int ldb,
       int tb, float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu, int accumulate) {
    __shared__ float As[GM_BK][GM_BM + 1];
    __shared__ float Bs[GM_BK][GM_BN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // split-K: slice z of gridDim.z handles a GM_BK-aligned K range and adds into C with atomics
    const int kper = ((K + gridDim.z - 1) / gridDim.z + GM_BK - 1) / GM_BK * GM_BK;
    const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
    const bool split = gridDim.z > 1;
    // These GEMMs are small and latency-bound: the next tile's global loads are issued into registers before
    // the current tile is multiplied (software pipeline), 8 + 8 independent loads per thread in flight.
    float ra[GM_LD], rb[GM_LD];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < GM_LD; ++q) {
            const int idx = tid + q * 256;
            int m, k;
            if (ta) { m = idx & 63; k = idx >> 6; } else { k = idx & 31; m = idx >> 5; }
            const int gm = m0 + m, gk = k0 + k;
            ra[q] = (gm < M && gk < kend) ? (ta ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk]) : 0.f;
            int n, kb;
            if (tb) { kb = idx & 31; n = idx >> 5; } else { n = idx & 63; kb = idx >> 6; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[q] = (gn < N && gkb < kend) ? (tb ? B[(long long)gn * ldb + gkb] : B[(long long)gkb * ldb + gn]) : 0.f;
        }
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += GM_BK) {
#pragma unroll
        for (int q = 0; q < GM_LD; ++q) {
            const int idx = tid + q * 256;
            if (ta) As[idx >> 6][idx & 63] = ra[q]; else As[idx & 31][idx >> 5] = ra[q];
            if (tb) Bs[idx & 31][idx >> 5] = rb[q]; else Bs[idx >> 6][idx & 63] = rb[q];
        }
        __syncthreads();
        if (k0 + GM_BK < kend) fetch(k0 + GM_BK);
#pragma unroll
        for (int kk = 0; kk < GM_BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty + 16 * i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx + 16 * j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (bias && blockIdx.z == 0) v += bias[gn];
            if (split) { atomicAdd(&C[(long long)gm * ldc + gn], v); continue; }
            if (accumulate) v += C[(long long)gm * ldc + gn];
            if (relu) v = fmaxf(v, 0.f);
            C[(long long)gm * ldc + gn] = v;
        }
    }
}

// Skinny GEMMs of the heads (M <= 16 rows: one row per scene or per candidate, A not transposed): the 64x64-tile kernel
// above spends its time on one CTA's serial K loop (or on a memset + split-K atomics); here every output column gets its
// own warp (B given as (N,K): lanes stride over K, coalesced) or the K range is cut over the eight warps of a CTA (B given
// as (K,N): lanes over columns, coalesced), all loads of a thread in flight at once, sums in a fixed order.
#define GS_MAXM 16
__global__ void __launch_bounds__(256)
k_gemm_skinny_nk(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu, int accumulate) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float acc[GS_MAXM];
#pragma unroll
    for (int m = 0; m < GS_MAXM; ++m) acc[m] = 0.f;
    const float* brow = B + (long long)n * ldb;
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
        const float b = brow[k];
#pragma unroll
        for (int m = 0; m < GS_MAXM; ++m)
            if (m < M) acc[m] = fmaf(A[(long long)m * lda + k], b, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < GS_MAXM; ++m) {
        if (m >= M) break;
        float v = warp_sum(acc[m]);
        if (lane == 0) {
            if (bias) v += bias[n];
            if (accumulate) v += C[(long long)m * ldc + n];
            if (relu) v = fmaxf(v, 0.f);
            C[(long long)m * ldc + n] = v;
        }
    }
}
__global__ void __launch_bounds__(256)
k_gemm_skinny_kn(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu, int accumulate) {
    __shared__ float sh[8][GS_MAXM][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5, n = blockIdx.x * 32 + lane;
    float acc[GS_MAXM];
#pragma unroll
    for (int m = 0; m < GS_MAXM; ++m) acc[m] = 0.f;
    if (n < N) {
#pragma unroll 4
        for (int k = g; k < K; k += 8) {
            const float b = B[(long long)k * ldb + n];
#pragma unroll
            for (int m = 0; m < GS_MAXM; ++m)
                if (m < M) acc[m] = fmaf(A[(long long)m * lda + k], b, acc[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < GS_MAXM; ++m) sh[g][m][lane] = acc[m];
    __syncthreads();
    // thread (g, lane) finishes rows g and g + 8 of column n
    for (int m = g; m < M; m += 8) {
        if (n >= N) break;
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += sh[q][m][lane];
        if (bias) v += bias[n];
        if (accumulate) v += C[(long long)m * ldc + n];
        if (relu) v = fmaxf(v, 0.f);
        C[(long long)m * ldc + n] = v;
    }
}

extern "C" int ir_gemm(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t trans_a,
                       const float* B, int32_t ldb, int32_t trans_b, float* C, int32_t ldc, const float* bias,
                       int32_t relu, int32_t accumulate, ir_stream_t stream) {
    IR_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0);
    if (!trans_a && M <= GS_MAXM && K >= 32) {
        if (trans_b) k_gemm_skinny_nk<<<ir_div_up(N, 8), 256, 0, (cudaStream_t)stream>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, accumulate);
        else k_gemm_skinny_kn<<<ir_div_up(N, 32), 256, 0, (cudaStream_t)stream>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, accumulate);
        IR_CHECK_LAUNCH();
        return IR_OK;
    }
    dim3 grid(ir_div_up(N, GM_BN), ir_div_up(M, GM_BM));
    // few output tiles and a long K (im2col conv GEMMs, weight gradients): split K over gridDim.z so the
    // launch fills the GPU; partial tiles are added with atomics into a zeroed (or accumulated-into) C
    const int tiles = grid.x * grid.y;
    if (!relu && tiles < IR_NUM_SMS && K >= 128 && ldc == N) {
        int z = ir_min_i(ir_div_up(2 * IR_NUM_SMS, tiles), K / 64);
        if (z > 1) {
            grid.z = z;
            if (!accumulate) IR_CHECK_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * 4, (cudaStream_t)stream));
        }
    }
    k_gemm<<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ column sums (bias gradients), deterministic
__global__ void __launch_bounds__(256)
k_colsum(const float* __restrict__ x, int M, int N, float* __restrict__ out) {
    __shared__ float sh[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (c < N)
        for (int r = w; r < M; r += 8) s += x[(long long)r * N + c];
    sh[w][lane] = s;
    __syncthreads();
    if (w == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += sh[q][lane];
        out[c] = t;
    }
}
extern "C" int ir_colsum(const float* x, int32_t M, int32_t N, float* out, ir_stream_t stream) {
    IR_CHECK_ARG(x && out && M > 0 && N > 0);
    k_colsum<<<ir_div_up(N, 32), 256, 0, (cudaStream_t)stream>>>(x, M, N, out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ elementwise: ReLU mask, dropout
__global__ void k_relu_bwd(const float* __restrict__ dy, const float* __restrict__ y, long long n, float* __restrict__ dx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
extern "C" int ir_relu_bwd(const float* dy, const float* y, int64_t n, float* dx, ir_stream_t stream) {
    IR_CHECK_ARG(dy && y && dx && n > 0);
    k_relu_bwd<<<ir_min_i(ir_div_up(n, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(dy, y, n, dx);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// counter-based generator (murmur-style finaliser of seed ^ index): keep with prob 1-p, scale 1/(1-p)
__device__ __forceinline__ float u01(unsigned long long seed, unsigned long long i) {
    unsigned long long k = seed + i * 0x9E3779B97F4A7C15ull;
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (float)(k >> 40) * (1.0f / 16777216.0f);
}
// seed_step (optional, device): a per-step counter folded into the seed on the device, so that a launch replayed
// from a CUDA graph (host seed baked in at capture) still draws a fresh mask every step
__global__ void k_dropout(const float* __restrict__ x, long long n, float p, unsigned long long seed,
                          const unsigned long long* __restrict__ seed_step, float* __restrict__ y,
                          unsigned char* __restrict__ mask) {
    const float scale = 1.f / (1.f - p);
    if (seed_step) seed += *seed_step * 0xD1B54A32D192ED03ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool keep = u01(seed, (unsigned long long)i) >= p;
        mask[i] = keep;
        y[i] = keep ? x[i] * scale : 0.f;
    }
}
__global__ void k_dropout_bwd(const float* __restrict__ dy, const unsigned char* __restrict__ mask, long long n,
                              float scale, float* __restrict__ dx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = mask[i] ? dy[i] * scale : 0.f;
}
static const unsigned long long* g_dropout_seed_step = nullptr;
extern "C" int ir_dropout_seed_step(const uint64_t* step_dev) {
    g_dropout_seed_step = (const unsigned long long*)step_dev;
    return IR_OK;
}
extern "C" int ir_dropout_fwd(const float* x, int64_t n, float p, uint64_t seed, float* y, uint8_t* mask,
                              ir_stream_t stream) {
    IR_CHECK_ARG(x && y && mask && n > 0 && p >= 0.f && p < 1.f);
    k_dropout<<<ir_min_i(ir_div_up(n, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(x, n, p, seed, g_dropout_seed_step,
                                                                                           y, mask);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
extern "C" int ir_dropout_bwd(const float* dy, const uint8_t* mask, int64_t n, float p, float* dx, ir_stream_t stream) {
    IR_CHECK_ARG(dy && mask && dx && n > 0 && p >= 0.f && p < 1.f);
    k_dropout_bwd<<<ir_min_i(ir_div_up(n, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(dy, mask, n, 1.f / (1.f - p), dx);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ LayerNorm (+ReLU), warp per row
__global__ void __launch_bounds__(256)
k_layernorm_fwd(const float* __restrict__ x, int M, int N, const float* __restrict__ g, const float* __restrict__ b,
                float eps, int relu, float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* xr = x + (long long)r * N;
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += xr[c];
    const float mu = warp_sum(s) / N;
    float v = 0.f;
    for (int c = lane; c < N; c += 32) { const float d = xr[c] - mu; v = fmaf(d, d, v); }
    const float rs = rsqrtf(warp_sum(v) / N + eps);
    for (int c = lane; c < N; c += 32) {
        float o = (xr[c] - mu) * rs * g[c] + b[c];
        if (relu) o = fmaxf(o, 0.f);
        y[(long long)r * N + c] = o;
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
}
// dgamma / dbeta are accumulated with atomics into zeroed arrays (<= a few thousand rows)
__global__ void __launch_bounds__(256)
k_layernorm_bwd(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, int M, int N,
                const float* __restrict__ g, const float* __restrict__ mean, const float* __restrict__ rstd, int relu,
                float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= M) return;
    const long long o = (long long)r * N;
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < N; c += 32) {
        float gv = dy[o + c];
        if (relu && !(y[o + c] > 0.f)) gv = 0.f;
        const float xh = (x[o + c] - mu) * rs;
        const float t = gv * g[c];
        s1 += t;
        s2 = fmaf(t, xh, s2);
        atomicAdd(&dgamma[c], gv * xh);
        atomicAdd(&dbeta[c], gv);
    }
    s1 = warp_sum(s1) / N;
    s2 = warp_sum(s2) / N;
    for (int c = lane; c < N; c += 32) {
        float gv = dy[o + c];
        if (relu && !(y[o + c] > 0.f)) gv = 0.f;
        const float xh = (x[o + c] - mu) * rs;
        dx[o + c] = rs * (gv * g[c] - s1 - xh * s2);
    }
}
extern "C" int ir_layernorm_fwd(const float* x, int32_t M, int32_t N, const float* gamma, const float* beta, float eps,
                                int32_t relu, float* y, float* mean, float* rstd, ir_stream_t stream) {
    IR_CHECK_ARG(x && gamma && beta && y && mean && rstd && M > 0 && N > 0);
    k_layernorm_fwd<<<ir_div_up(M, 8), 256, 0, (cudaStream_t)stream>>>(x, M, N, gamma, beta, eps, relu, y, mean, rstd);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
extern "C" int ir_layernorm_bwd(const float* dy, const float* y, const float* x, int32_t M, int32_t N, const float* gamma,
                                const float* mean, const float* rstd, int32_t relu, float* dx, float* dgamma,
                                float* dbeta, ir_stream_t stream) {
    IR_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && M > 0 && N > 0 && (!relu || y));
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)N * 4, st));
    IR_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)N * 4, st));
    k_layernorm_bwd<<<ir_div_up(M, 8), 256, 0, st>>>(dy, y, x, M, N, gamma, mean, rstd, relu, dx, dgamma, dbeta);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ match scores
// mode 0: score[r] = <l2n(a_r), p_r>, l2n(a) = a / max(|a|, 1e-12)          (attribute: F.normalize + dot)
// mode 1: score[r] = <a_r, p_r> / max(|a_r| |p_r|, 1e-8)                     (relation / scene: cosine_similarity)
// with p_r = partner[seg[r]].  Backward gives da and (deterministically, one CTA per partner row) dpartner.
__global__ void __launch_bounds__(256)
k_match_fwd(const float* __restrict__ a, const float* __restrict__ partner, const int* __restrict__ seg, int M, int N,
            int mode, float* __restrict__ score) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* ar = a + (long long)r * N;
    const float* pr = partner + (long long)seg[r] * N;
    float ab = 0.f, aa = 0.f, pp = 0.f;
    for (int c = lane; c < N; c += 32) { ab = fmaf(ar[c], pr[c], ab); aa = fmaf(ar[c], ar[c], aa); pp = fmaf(pr[c], pr[c], pp); }
    ab = warp_sum(ab); aa = warp_sum(aa); pp = warp_sum(pp);
    if (lane == 0) score[r] = mode == 0 ? ab / fmaxf(sqrtf(aa), 1e-12f) : ab / fmaxf(sqrtf(aa) * sqrtf(pp), 1e-8f);
}
__global__ void __launch_bounds__(256)
k_match_bwd_a(const float* __restrict__ ds, const float* __restrict__ a, const float* __restrict__ partner,
              const int* __restrict__ seg, int M, int N, int mode, float* __restrict__ da) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* ar = a + (long long)r * N;
    const float* pr = partner + (long long)seg[r] * N;
    float ab = 0.f, aa = 0.f, pp = 0.f;
    for (int c = lane; c < N; c += 32) { ab = fmaf(ar[c], pr[c], ab); aa = fmaf(ar[c], ar[c], aa); pp = fmaf(pr[c], pr[c], pp); }
    ab = warp_sum(ab); aa = warp_sum(aa); pp = warp_sum(pp);
    const float na = sqrtf(aa), np_ = sqrtf(pp), g = ds[r];
    for (int c = lane; c < N; c += 32) {
        float d;
        if (mode == 0) {
            const float den = fmaxf(na, 1e-12f);            // s = ab/na : ds/da = p/na - ab a / na^3
            d = pr[c] / den - (na > 1e-12f ? ab * ar[c] / (den * den * den) : 0.f);
        } else {
            const float den = fmaxf(na * np_, 1e-8f);       // s = ab/(na np): ds/da = p/den - s a/na^2
            d = pr[c] / den - (na * np_ > 1e-8f ? (ab / den) * ar[c] / aa : 0.f);
        }
        da[(long long)r * N + c] = g * d;
    }
}
// one CTA per partner row b: sums over the rows r in [ofs[b], ofs[b+1]) (candidates of a scene are contiguous)
__global__ void __launch_bounds__(256)
k_match_bwd_p(const float* __restrict__ ds, const float* __restrict__ a, const float* __restrict__ partner,
              const int* __restrict__ row_ofs, int N, int mode, float* __restrict__ dpartner) {
    __shared__ float red[8];
    __shared__ float s_coef[2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = row_ofs[b], r1 = row_ofs[b + 1];
    const float* pr = partner + (long long)b * N;
    float pp = 0.f;
    for (int c = tid; c < N; c += 256) pp = fmaf(pr[c], pr[c], pp);
    pp = warp_sum(pp);
    if (lane == 0) red[w] = pp;
    __syncthreads();
    pp = 0.f;
    for (int q = 0; q < 8; ++q) pp += red[q];
    const float np_ = sqrtf(pp);
    float acc[2] = {0.f, 0.f};                                // N <= 512
    for (int r = r0; r < r1; ++r) {
        const float* ar = a + (long long)r * N;
        __syncthreads();
        float ab = 0.f, aa = 0.f;
        for (int c = tid; c < N; c += 256) { ab = fmaf(ar[c], pr[c], ab); aa = fmaf(ar[c], ar[c], aa); }
        ab = warp_sum(ab); aa = warp_sum(aa);
        __shared__ float r2[2][8];
        if (lane == 0) { r2[0][w] = ab; r2[1][w] = aa; }
        __syncthreads();
        if (tid == 0) {
            float sab = 0.f, saa = 0.f;
            for (int q = 0; q < 8; ++q) { sab += r2[0][q]; saa += r2[1][q]; }
            s_coef[0] = sab; s_coef[1] = sqrtf(saa);
        }
        __syncthreads();
        ab = s_coef[0];
        const float na = s_coef[1], g = ds[r];
        int i = 0;
        for (int c = tid; c < N; c += 256, ++i) {
            float d;
            if (mode == 0) d = ar[c] / fmaxf(na, 1e-12f);
            else {
                const float den = fmaxf(na * np_, 1e-8f);
                d = ar[c] / den - (na * np_ > 1e-8f ? (ab / den) * pr[c] / pp : 0.f);
            }
            acc[i] = fmaf(g, d, acc[i]);
        }
    }
    int i = 0;
    for (int c = tid; c < N; c += 256, ++i) dpartner[(long long)b * N + c] = acc[i];
}
extern "C" int ir_match_fwd(const float* a, const float* partner, const int32_t* seg, int32_t M, int32_t N, int32_t mode,
                            float* score, ir_stream_t stream) {
    IR_CHECK_ARG(a && partner && seg && score && M > 0 && N > 0 && (mode == 0 || mode == 1));
    k_match_fwd<<<ir_div_up(M, 8), 256, 0, (cudaStream_t)stream>>>(a, partner, seg, M, N, mode, score);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
extern "C" int ir_match_bwd(const float* dscore, const float* a, const float* partner, const int32_t* seg,
                            const int32_t* row_ofs, int32_t M, int32_t N, int32_t n_partner, int32_t mode, float* da,
                            float* dpartner, ir_stream_t stream) {
    IR_CHECK_ARG(dscore && a && partner && seg && row_ofs && da && dpartner && M > 0 && N > 0 && N <= 512 && n_partner > 0);
    cudaStream_t st = (cudaStream_t)stream;
    k_match_bwd_a<<<ir_div_up(M, 8), 256, 0, st>>>(dscore, a, partner, seg, M, N, mode, da);
    IR_CHECK_LAUNCH();
    k_match_bwd_p<<<n_partner, 256, 0, st>>>(dscore, a, partner, row_ofs, N, mode, dpartner);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// L2 row normalisation y = x / max(|x|, 1e-12) (F.normalize) and its backward
__global__ void __launch_bounds__(256)
k_l2norm(const float* __restrict__ x, const float* __restrict__ dy, int M, int N, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* xr = x + (long long)r * N;
    float xx = 0.f, xd = 0.f;
    for (int c = lane; c < N; c += 32) { xx = fmaf(xr[c], xr[c], xx); if (dy) xd = fmaf(xr[c], dy[(long long)r * N + c], xd); }
    xx = warp_sum(xx); xd = warp_sum(xd);
    const float n = sqrtf(xx), den = fmaxf(n, 1e-12f);
    for (int c = lane; c < N; c += 32) {
        if (!dy) out[(long long)r * N + c] = xr[c] / den;
        else out[(long long)r * N + c] = dy[(long long)r * N + c] / den - (n > 1e-12f ? xr[c] * xd / (den * den * den) : 0.f);
    }
}
extern "C" int ir_l2norm_fwd(const float* x, int32_t M, int32_t N, float* y, ir_stream_t stream) {
    IR_CHECK_ARG(x && y && M > 0 && N > 0);
    k_l2norm<<<ir_div_up(M, 8), 256, 0, (cudaStream_t)stream>>>(x, nullptr, M, N, y);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
extern "C" int ir_l2norm_bwd(const float* dy, const float* x, int32_t M, int32_t N, float* dx, ir_stream_t stream) {
    IR_CHECK_ARG(dy && x && dx && M > 0 && N > 0);
    k_l2norm<<<ir_div_up(M, 8), 256, 0, (cudaStream_t)stream>>>(x, dy, M, N, dx);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ Conv2d 3x3 (valid) as im2col + GEMM, NHWC
// col[(b,y,x), (ky,kx,c)] = in[b, y+ky, x+kx, c];   col2im is the gather-form adjoint.
__global__ void k_im2col3(const float* __restrict__ in, int B, int H, int W, int C, float* __restrict__ col) {
    const int Ho = H - 2, Wo = W - 2;
    const long long total = (long long)B * Ho * Wo * 9 * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int tap = (int)(t % 9); t /= 9;
        const int x = (int)(t % Wo); t /= Wo;
        const int y = (int)(t % Ho);
        const int b = (int)(t / Ho);
        col[i] = in[(((long long)b * H + y + tap / 3) * W + x + tap % 3) * C + c];
    }
}
__global__ void k_col2im3(const float* __restrict__ dcol, int B, int H, int W, int C, float* __restrict__ din) {
    const int Ho = H - 2, Wo = W - 2;
    const long long total = (long long)B * H * W * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H);
        const int b = (int)(t / H);
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int oy = y - ky, ox = x - kx;
                if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo)
                    s += dcol[((((long long)b * Ho + oy) * Wo + ox) * 9 + ky * 3 + kx) * C + c];
            }
        din[i] = s;
    }
}
extern "C" int ir_im2col_3x3(const float* in, int32_t B, int32_t H, int32_t W, int32_t C, float* col, ir_stream_t stream) {
    IR_CHECK_ARG(in && col && B > 0 && H > 2 && W > 2 && C > 0);
    const long long total = (long long)B * (H - 2) * (W - 2) * 9 * C;
    k_im2col3<<<ir_min_i(ir_div_up(total, 256), IR_NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(in, B, H, W, C, col);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
extern "C" int ir_col2im_3x3(const float* dcol, int32_t B, int32_t H, int32_t W, int32_t C, float* din, ir_stream_t stream) {
    IR_CHECK_ARG(dcol && din && B > 0 && H > 2 && W > 2 && C > 0);
    const long long total = (long long)B * H * W * C;
    k_col2im3<<<ir_min_i(ir_div_up(total, 256), IR_NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(dcol, B, H, W, C, din);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ BEV densification backward
// forward (ir_bev, raw mode): tmp[r] = F[r] @ kern[z_r]; dense[cell_r] += tmp[r] for kept rows.
// dF[r] = d dense[cell_r] @ kern[z_r]^T ;  dkern[z] = sum_{r: z_r = z} F[r]^T d dense[cell_r].
#define BV_C 128
__global__ void __launch_bounds__(BV_C)
k_bev_bwd_feats(const float* __restrict__ ddense, const int4* __restrict__ coords, const int* __restrict__ cell,
                const int* __restrict__ n_dev, int stride, const float* __restrict__ kern, float* __restrict__ dF) {
    __shared__ float g[BV_C];
    const int n = *n_dev, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int cl = cell[r];
        if (cl < 0) { dF[(long long)r * BV_C + threadIdx.x] = 0.f; continue; }
        __syncthreads();
        g[threadIdx.x] = ddense[(long long)cl * BV_C + threadIdx.x];
        __syncthreads();
        const float* kz = kern + (long long)(coords[r].z / stride) * BV_C * BV_C;
        for (int ci = w; ci < BV_C; ci += BV_C / 32) {           // warp per input channel: coalesced kernel row
            const float* kr = kz + (long long)ci * BV_C;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < BV_C / 32; ++q) s = fmaf(g[lane + 32 * q], kr[lane + 32 * q], s);
            s = warp_sum(s);
            if (lane == 0) dF[(long long)r * BV_C + ci] = s;
        }
    }
}
// grid (BV_C/8, n_z): CTA owns dkern[z][ci0..ci0+7][:]; 8 row groups x 128 output channels: group g scans rows g, g+8, ...
// in order, the eight partial tiles are added in group order (fixed summation order, no atomics)
#define BVK_GROUPS 8
__global__ void __launch_bounds__(BV_C * BVK_GROUPS)
k_bev_bwd_kernel(const float* __restrict__ ddense, const float* __restrict__ F, const int4* __restrict__ coords,
                 const int* __restrict__ cell, const int* __restrict__ n_dev, int stride, float* __restrict__ dkern) {
    __shared__ float sh[BVK_GROUPS][8][BV_C];
    const int n = *n_dev, z = blockIdx.y, ci0 = blockIdx.x * 8, co = threadIdx.x % BV_C, grp = threadIdx.x / BV_C;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int r = grp; r < n; r += BVK_GROUPS) {
        const int cl = cell[r];
        if (cl < 0 || coords[r].z / stride != z) continue;         // uniform over the group's four warps
        const float gv = ddense[(long long)cl * BV_C + co];
        const float4 f0 = *reinterpret_cast<const float4*>(F + (long long)r * BV_C + ci0);
        const float4 f1 = *reinterpret_cast<const float4*>(F + (long long)r * BV_C + ci0 + 4);
        acc[0] = fmaf(f0.x, gv, acc[0]); acc[1] = fmaf(f0.y, gv, acc[1]);
        acc[2] = fmaf(f0.z, gv, acc[2]); acc[3] = fmaf(f0.w, gv, acc[3]);
        acc[4] = fmaf(f1.x, gv, acc[4]); acc[5] = fmaf(f1.y, gv, acc[5]);
        acc[6] = fmaf(f1.z, gv, acc[6]); acc[7] = fmaf(f1.w, gv, acc[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sh[grp][i][co] = acc[i];
    __syncthreads();
    float s = 0.f;                                                  // thread (grp, co) finishes output row i = grp
#pragma unroll
    for (int q = 0; q < BVK_GROUPS; ++q) s += sh[q][grp][co];
    dkern[((long long)z * BV_C + ci0 + grp) * BV_C + co] = s;
}
extern "C" int ir_bev_bwd(const float* ddense, const float* feats, const int32_t* coords, const int32_t* cell,
                          const int32_t* n_dev, int64_t n_max, int32_t stride, const float* kernel, int32_t n_z,
                          float* dfeats, float* dkernel, ir_stream_t stream) {
    IR_CHECK_ARG(ddense && feats && coords && cell && n_dev && kernel && dfeats && dkernel && n_max > 0 && n_z > 0 && stride > 0);
    cudaStream_t st = (cudaStream_t)stream;
    k_bev_bwd_feats<<<ir_min_i(n_max, IR_NUM_SMS * 16), BV_C, 0, st>>>(ddense, (const int4*)coords, cell, n_dev, stride, kernel, dfeats);
    IR_CHECK_LAUNCH();
    k_bev_bwd_kernel<<<dim3(BV_C / 8, n_z), BV_C * BVK_GROUPS, 0, st>>>(ddense, feats, (const int4*)coords, cell, n_dev, stride, dkernel);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ scene attention backward
// forward: l_i = <f_i, q>/sqrt(C); a = softmax(l); s = sum_i a_i f_i.   One CTA per scene, thread = channel.
__global__ void __launch_bounds__(128)
k_scene_attention_bwd(const float* __restrict__ feats, const float* __restrict__ q, const float* __restrict__ atten,
                      const float* __restrict__ ds, const float* __restrict__ datten_in, int ncell, int C,
                      float* __restrict__ dfeats, float* __restrict__ dq) {
    extern __shared__ float sm[];
    float* da = sm;                 // [ncell]
    float* dl = sm + ncell;         // [ncell]
    __shared__ float s_dot;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    const float* f = feats + (long long)b * ncell * C;
    const float* a = atten + (long long)b * ncell;
    const float inv = rsqrtf((float)C);
    for (int i = w; i < ncell; i += nw) {                      // da_i = <ds, f_i> (+ upstream d atten)
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(ds[(long long)b * C + c], f[(long long)i * C + c], s);
        s = warp_sum(s);
        if (lane == 0) da[i] = s + (datten_in ? datten_in[(long long)b * ncell + i] : 0.f);
    }
    __syncthreads();
    if (w == 0) {
        float s = 0.f;
        for (int i = lane; i < ncell; i += 32) s = fmaf(a[i], da[i], s);
        s = warp_sum(s);
        if (lane == 0) s_dot = s;
    }
    __syncthreads();
    for (int i = tid; i < ncell; i += blockDim.x) dl[i] = a[i] * (da[i] - s_dot);
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        const float qc = q[(long long)b * C + c] * inv, dsc = ds[(long long)b * C + c];
        float dqc = 0.f;
        for (int i = 0; i < ncell; ++i) {
            const float fv = f[(long long)i * C + c];
            dfeats[((long long)b * ncell + i) * C + c] = a[i] * dsc + dl[i] * qc;
            dqc = fmaf(dl[i], fv, dqc);
        }
        dq[(long long)b * C + c] = dqc * inv;
    }
}
extern "C" int ir_scene_attention_bwd(const float* feats, const float* q, const float* atten, const float* dscene,
                                      const float* datten, int32_t B, int32_t ncell, int32_t C, float* dfeats, float* dq,
                                      ir_stream_t stream) {
    IR_CHECK_ARG(feats && q && atten && dscene && dfeats && dq && B > 0 && ncell > 0 && ncell <= 4096 && C > 0);
    k_scene_attention_bwd<<<B, 128, (size_t)2 * ncell * 4, (cudaStream_t)stream>>>(feats, q, atten, dscene, datten, ncell, C, dfeats, dq);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ token attention backward
// forward (ir_token_attention): l = feats . fcw_h + fcb_h; s = softmax_L(l); u = s*mask; a = u / sum u;
// pooled_h = a @ embed.   One CTA per sample; the four heads are accumulated in registers / smem, so
// dfeats and dembed need no atomics; dfcw / dfcb are written per sample (B,4,D)/(B,4) and summed by ir_colsum.
#define TAB_MAXL 128
__global__ void __launch_bounds__(256)
k_token_attention_bwd(const float* __restrict__ feats, const float* __restrict__ embed, long long embed_stride,
                      const long long* __restrict__ lengths, const float* __restrict__ fcw, const float* __restrict__ fcb,
                      const float* __restrict__ atten, const float* __restrict__ dpooled, int B, int L, int D, int E,
                      float* __restrict__ dfeats, float* __restrict__ dembed, float* __restrict__ dfcw_part,
                      float* __restrict__ dfcb_part) {
    __shared__ float s_a[4][TAB_MAXL], s_dl[4][TAB_MAXL], s_s[4][TAB_MAXL];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int len = max(0, min((int)lengths[b], L));
    // recompute the un-normalised softmax s (needed for d softmax) and load a
    for (int p = w; p < 4 * L; p += 8) {
        const int h = p / L, t = p - h * L;
        const float* f = feats + ((long long)b * L + t) * D;
        float v = 0.f;
        for (int d = lane; d < D; d += 32) v = fmaf(f[d], fcw[h * D + d], v);
        v = warp_sum(v);
        if (lane == 0) { s_s[h][t] = v + fcb[h]; s_a[h][t] = atten[((long long)h * B + b) * L + t]; }
    }
    __syncthreads();
    if (w < 4) {
        const int h = w;
        float m = -INFINITY;
        for (int t = lane; t < L; t += 32) m = fmaxf(m, s_s[h][t]);
        m = warp_max(m);
        float z = 0.f;
        for (int t = lane; t < L; t += 32) { const float e = expf(s_s[h][t] - m); s_s[h][t] = e; z += e; }
        z = warp_sum(z);
        float su = 0.f;
        for (int t = lane; t < L; t += 32) { s_s[h][t] /= z; if (t < len) su += s_s[h][t]; }
        su = warp_sum(su);                                     // sum of masked softmax
        // da_t = <dpooled_h, embed_t>
        float dot_a = 0.f;
        for (int t = 0; t < L; ++t) {
            float v = 0.f;
            for (int e = lane; e < E; e += 32)
                v = fmaf(dpooled[((long long)h * B + b) * E + e], embed[(long long)b * embed_stride + (long long)t * E + e], v);
            v = warp_sum(v);
            if (lane == 0) s_dl[h][t] = v;
            dot_a = fmaf(v, s_a[h][t], dot_a);                 // a_t is 0 on pads
        }
        __syncwarp();
        // a = u/su: du_t = (da_t - sum_j da_j a_j)/su; ds_t = du_t * mask_t; dl = s*(ds - sum_j s_j ds_j)
        float sd = 0.f;
        for (int t = lane; t < L; t += 32) {
            const float dsv = (t < len) ? (s_dl[h][t] - dot_a) / su : 0.f;
            s_dl[h][t] = dsv;
            sd = fmaf(s_s[h][t], dsv, sd);
        }
        sd = warp_sum(sd);
        float db = 0.f;
        for (int t = lane; t < L; t += 32) { const float v = s_s[h][t] * (s_dl[h][t] - sd); s_dl[h][t] = v; db += v; }
        db = warp_sum(db);
        if (lane == 0) dfcb_part[b * 4 + h] = db;
    }
    __syncthreads();
    for (int d = tid; d < D; d += 256) {                        // dfeats += dl (x) fcw ; dfcw_part = sum_t dl_t f_t
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        const float w0 = fcw[d], w1 = fcw[D + d], w2 = fcw[2 * D + d], w3 = fcw[3 * D + d];
        for (int t = 0; t < L; ++t) {
            const long long o = ((long long)b * L + t) * D + d;
            const float fv = feats[o];
            dfeats[o] = s_dl[0][t] * w0 + s_dl[1][t] * w1 + s_dl[2][t] * w2 + s_dl[3][t] * w3;
            g0 = fmaf(s_dl[0][t], fv, g0); g1 = fmaf(s_dl[1][t], fv, g1);
            g2 = fmaf(s_dl[2][t], fv, g2); g3 = fmaf(s_dl[3][t], fv, g3);
        }
        dfcw_part[((long long)b * 4 + 0) * D + d] = g0; dfcw_part[((long long)b * 4 + 1) * D + d] = g1;
        dfcw_part[((long long)b * 4 + 2) * D + d] = g2; dfcw_part[((long long)b * 4 + 3) * D + d] = g3;
    }
    for (int e = tid; e < E; e += 256) {                        // dembed_t = sum_h a_h,t dpooled_h
        const float p0 = dpooled[((long long)0 * B + b) * E + e], p1 = dpooled[((long long)1 * B + b) * E + e];
        const float p2 = dpooled[((long long)2 * B + b) * E + e], p3 = dpooled[((long long)3 * B + b) * E + e];
        for (int t = 0; t < L; ++t)
            dembed[((long long)b * L + t) * E + e] = s_a[0][t] * p0 + s_a[1][t] * p1 + s_a[2][t] * p2 + s_a[3][t] * p3;
    }
}
extern "C" int ir_token_attention_bwd(const float* feats, const float* embed, int64_t embed_stride, const int64_t* lengths,
                                      const float* fcw, const float* fcb, const float* atten, const float* dpooled,
                                      int32_t B, int32_t L, int32_t D, int32_t E, float* dfeats, float* dembed,
                                      float* dfcw_part, float* dfcb_part, ir_stream_t stream) {
    IR_CHECK_ARG(feats && embed && lengths && fcw && fcb && atten && dpooled && dfeats && dembed && dfcw_part && dfcb_part);
    IR_CHECK_ARG(B > 0 && L > 0 && L <= TAB_MAXL && D > 0 && E > 0);
    k_token_attention_bwd<<<B, 256, 0, (cudaStream_t)stream>>>(feats, embed, embed_stride, (const long long*)lengths, fcw, fcb,
                                                              atten, dpooled, B, L, D, E, dfeats, dembed, dfcw_part, dfcb_part);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ GRU layer backward (BPTT)
// grid (B, 2 directions), 384 threads (one per gate row), W_hh of the direction resident in shared
// memory (384 x 129 floats, padded rows: row-wise AND column-wise reads are conflict-free).
// Per step, from dh (upstream d out[t] + carry):
//   recompute r,z,n from xproj[t] and h_prev (= out of the previous step of this direction);
//   dn = dh (1-z) (1-n^2); dz = dh (h_prev - n) z (1-z); dr = dn (W_hn h_prev + b_hn) r (1-r);
//   dxproj[t] = [dr, dz, dn]; dhp[t] = [dr, dz, dn*r]; dh_prev = dh z + W_hh^T dhp.
// dhp and h_prev are written out per step: dW_hh = dhp^T @ h_prev and db_hh = colsum(dhp) are then ONE
// GEMM / column sum per direction over all (sample, step) rows instead of a serial rank-1 update chain.
#define GB_H 128
#define GB_LD 129
__global__ void __launch_bounds__(3 * GB_H)
k_gru_layer_bwd(const float* __restrict__ xproj, const float* __restrict__ whh, const float* __restrict__ bhh,
                const long long* __restrict__ lengths, const float* __restrict__ out, const float* __restrict__ dout,
                int L, float* __restrict__ dxproj, float* __restrict__ dhp_out, float* __restrict__ hprev_out) {
    constexpr int H = GB_H, G = 3 * GB_H;
    extern __shared__ float Wsm[];                                        // [G][GB_LD]
    __shared__ float s_hprev[H], s_hp[G], s_dhp[G], s_dh[H], s_part[3][H];
    const int b = blockIdx.x, dir = blockIdx.y, j = threadIdx.x;
    const int len = max(0, min((int)lengths[b], L));
    {
        const float* W = whh + (size_t)dir * G * H;
        for (int i = j; i < G * H; i += G) Wsm[(i >> 7) * GB_LD + (i & (H - 1))] = __ldg(W + i);
    }
    const float bj = bhh[dir * G + j];
    if (j < H) s_dh[j] = 0.f;
    for (int t = len; t < L; ++t) {                                       // pads carry no gradient
        const size_t o = (((size_t)b * L + t) * 2 + dir);
        dxproj[o * G + j] = 0.f;
        dhp_out[o * G + j] = 0.f;
        if (j < H) hprev_out[o * H + j] = 0.f;
    }
    __syncthreads();
    const int pc = j & (H - 1), part = j >> 7;
    for (int s = len - 1; s >= 0; --s) {
        const int tt = dir ? (len - 1 - s) : s;                           // time index of step s
        const int tp = dir ? tt + 1 : tt - 1;                             // previous step's time index
        const size_t o = (((size_t)b * L + tt) * 2 + dir);
        if (j < H) {
            const float hv = (s > 0) ? out[((size_t)b * L + tp) * (2 * H) + dir * H + j] : 0.f;
            s_hprev[j] = hv;
            hprev_out[o * H + j] = hv;
            s_dh[j] += dout[((size_t)b * L + tt) * (2 * H) + dir * H + j];
        }
        __syncthreads();
        {   // hp_j = W_hh[j,:] . h_prev + b_j
            const float* wr = Wsm + j * GB_LD;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
            for (int c = 0; c < H; c += 2) { a0 = fmaf(wr[c], s_hprev[c], a0); a1 = fmaf(wr[c + 1], s_hprev[c + 1], a1); }
            s_hp[j] = a0 + a1 + bj;
        }
        __syncthreads();
        float dhz = 0.f;
        if (j < H) {
            const float* xp = xproj + o * G;
            const float r = 1.f / (1.f + expf(-(xp[j] + s_hp[j])));
            const float z = 1.f / (1.f + expf(-(xp[H + j] + s_hp[H + j])));
            const float n = tanhf(xp[2 * H + j] + r * s_hp[2 * H + j]);
            const float dh = s_dh[j];
            const float dn = dh * (1.f - z) * (1.f - n * n);
            const float dz = dh * (s_hprev[j] - n) * z * (1.f - z);
            const float dr = dn * s_hp[2 * H + j] * r * (1.f - r);
            float* dx = dxproj + o * G;
            dx[j] = dr; dx[H + j] = dz; dx[2 * H + j] = dn;
            s_dhp[j] = dr; s_dhp[H + j] = dz; s_dhp[2 * H + j] = dn * r;
            dhz = dh * z;                                                 // carry through the z gate
        }
        __syncthreads();
        dhp_out[o * G + j] = s_dhp[j];
        {   // W_hh^T dhp, split over the three gate blocks: thread (part, pc) sums rows part*H .. +H of column pc
            const float* wc = Wsm + (size_t)part * H * GB_LD + pc;
            const float* dp = s_dhp + part * H;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
            for (int i = 0; i < H; i += 2) { a0 = fmaf(wc[i * GB_LD], dp[i], a0); a1 = fmaf(wc[(i + 1) * GB_LD], dp[i + 1], a1); }
            s_part[part][pc] = a0 + a1;
        }
        __syncthreads();
        if (j < H) s_dh[j] = dhz + s_part[0][j] + s_part[1][j] + s_part[2][j];
        __syncthreads();
    }
}
extern "C" int ir_gru_layer_bwd(const float* xproj, const float* whh, const float* bhh, const int64_t* lengths,
                                const float* out, const float* dout, int32_t B, int32_t L, int32_t H, float* dxproj,
                                float* dhp, float* hprev, ir_stream_t stream) {
    IR_CHECK_ARG(xproj && whh && bhh && lengths && out && dout && dxproj && dhp && hprev && B > 0 && L > 0 && H == GB_H);
    const size_t smem = (size_t)3 * GB_H * GB_LD * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_gru_layer_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    k_gru_layer_bwd<<<dim3(B, 2), 3 * GB_H, smem, (cudaStream_t)stream>>>(xproj, whh, bhh, (const long long*)lengths, out, dout, L,
                                                                         dxproj, dhp, hprev);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ EdgeConv pieces (train mode)
// edge e = (query m, slot s) with neighbour j = nbr[m][s] (or -1).  Inputs of the two edge MLPs:
//   w_in[e] = [xyz_j - xyz_i, onehot_i, onehot_j]          (3 + 2*ncls)
//   e_in[e] = [x_i, w[e], x_j]                              (3*F)      (models/basic_blocks.py:131-132)
// x carries no gradient (instance statistics); only w does, so the backward of the concat is a slice.
__global__ void k_edge_inputs(const float* __restrict__ x, const float* __restrict__ xyz, const int* __restrict__ qidx,
                              const int* __restrict__ nbr, int nq, int k, int F, int ncls, const float* __restrict__ w,
                              float* __restrict__ w_in, float* __restrict__ e_in) {
    const int E = nq * k, Dw = 3 + 2 * ncls, De = 3 * F;
    const int D = w ? De : Dw;
    const long long total = (long long)E * D;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / D), c = (int)(t - (long long)e * D);
        const int i = qidx[e / k], j = nbr[e];
        float v = 0.f;
        if (j >= 0) {
            if (!w) {
                if (c < 3) v = xyz[(long long)j * 3 + c] - xyz[(long long)i * 3 + c];
                else if (c < 3 + ncls) v = x[(long long)i * F + (F - ncls) + (c - 3)];
                else v = x[(long long)j * F + (F - ncls) + (c - 3 - ncls)];
            } else {
                if (c < F) v = x[(long long)i * F + c];
                else if (c < 2 * F) v = w[(long long)e * F + (c - F)];
                else v = x[(long long)j * F + (c - 2 * F)];
            }
        }
        if (w) e_in[t] = v; else w_in[t] = v;
    }
}
extern "C" int ir_edge_inputs(const float* x, const float* xyz, const int32_t* qidx, const int32_t* nbr, int32_t nq,
                              int32_t k, int32_t F, int32_t ncls, const float* w, float* w_in, float* e_in,
                              ir_stream_t stream) {
    IR_CHECK_ARG(x && xyz && qidx && nbr && nq > 0 && k > 0 && F > ncls && ncls > 0 && (w ? e_in != nullptr : w_in != nullptr));
    const long long total = (long long)nq * k * (w ? 3 * F : 3 + 2 * ncls);
    k_edge_inputs<<<ir_min_i(ir_div_up(total, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(x, xyz, qidx, nbr, nq, k, F, ncls, w, w_in, e_in);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
// out[m,c] = max over valid slots of msg[m,s,c] (0 if none), arg = winning slot (first on ties); backward scatters.
__global__ void k_edge_max(const float* __restrict__ msg, const int* __restrict__ nbr, int nq, int k, int C,
                           float* __restrict__ out, int* __restrict__ arg) {
    const long long total = (long long)nq * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(t / C), c = (int)(t - (long long)m * C);
        float best = -INFINITY;
        int bs = -1;
        for (int s = 0; s < k; ++s) {
            if (nbr[m * k + s] < 0) continue;
            const float v = msg[((long long)m * k + s) * C + c];
            if (v > best) { best = v; bs = s; }
        }
        out[t] = bs >= 0 ? best : 0.f;
        arg[t] = bs;
    }
}
__global__ void k_edge_max_bwd(const float* __restrict__ dout, const int* __restrict__ arg, int nq, int k, int C,
                               float* __restrict__ dmsg) {
    const long long total = (long long)nq * k * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        const long long e = t / C;
        const int m = (int)(e / k), s = (int)(e - (long long)m * k);
        dmsg[t] = (arg[(long long)m * C + c] == s) ? dout[(long long)m * C + c] : 0.f;
    }
}
extern "C" int ir_edge_max_fwd(const float* msg, const int32_t* nbr, int32_t nq, int32_t k, int32_t C, float* out,
                               int32_t* arg, ir_stream_t stream) {
    IR_CHECK_ARG(msg && nbr && out && arg && nq > 0 && k > 0 && C > 0);
    k_edge_max<<<ir_min_i(ir_div_up((long long)nq * C, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(msg, nbr, nq, k, C, out, arg);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
extern "C" int ir_edge_max_bwd(const float* dout, const int32_t* arg, int32_t nq, int32_t k, int32_t C, float* dmsg,
                               ir_stream_t stream) {
    IR_CHECK_ARG(dout && arg && dmsg && nq > 0 && k > 0 && C > 0);
    k_edge_max_bwd<<<ir_min_i(ir_div_up((long long)nq * k * C, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(dout, arg, nq, k, C, dmsg);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ fused two-layer head (one call per direction)
// Linear -> {BatchNorm1d (batch statistics) | LayerNorm} -> ReLU [-> Dropout] -> Linear: the nn.Sequential heads of
// models/{attribute,relation,scene}_module.py as ONE library call forward and ONE backward (the host issues 2 calls
// per head and step instead of ~14).  Arena (caller-owned, ir_mlp_head_arena_bytes):
//   h1 (M,N1) pre-norm | h2 (M,N1) normalised + ReLU | h3 (M,N1) after dropout | stats 2*max(M,N1) | mask (M,N1) u8 |
//   BN partial-sum scratch | backward temporaries 2 x (M,N1).
extern "C" int ir_bn_train_fwd(const float*, const int32_t*, int32_t, int32_t, const float*, const float*, const float*, int32_t,
                               float, float, float*, float*, float*, float*, float*, float*, ir_stream_t);
extern "C" int ir_bn_train_bwd(const float*, const float*, const float*, const int32_t*, int32_t, int32_t, const float*,
                               const float*, const float*, int32_t, float*, float*, float*, float*, float*, float*, ir_stream_t);
extern "C" int64_t ir_bn_scratch_floats(int32_t);

struct HeadArena {
    float *h1, *h2, *h3, *stat_a, *stat_b, *bn_scratch, *t1, *t2;
    uint8_t* mask;
    int64_t bytes;
};
static HeadArena head_arena(void* base, int64_t M, int64_t N1) {
    HeadArena a;
    char* p = (char*)base;
    auto take = [&](int64_t bytes) { char* o = p; p += (bytes + 255) / 256 * 256; return o; };
    const int64_t mn = M * N1 * 4, st = (M > N1 ? M : N1) * 4;
    a.h1 = (float*)take(mn); a.h2 = (float*)take(mn); a.h3 = (float*)take(mn);
    a.stat_a = (float*)take(st); a.stat_b = (float*)take(st);
    a.mask = (uint8_t*)take(M * N1);
    a.bn_scratch = (float*)take(ir_bn_scratch_floats(256) * 4);
    a.t1 = (float*)take(mn); a.t2 = (float*)take(mn);
    a.bytes = p - (char*)base;
    return a;
}
extern "C" int64_t ir_mlp_head_arena_bytes(int32_t M, int32_t N1) { return head_arena(nullptr, M, N1).bytes; }

extern "C" int ir_mlp_head_train_fwd(const ir_mlp_head_t* h, const float* x, void* arena, float* y, ir_stream_t stream) {
    IR_CHECK_ARG(h && x && arena && y && h->M > 0 && h->K > 0 && h->N1 > 0 && h->N2 > 0 && (h->norm == 1 || h->norm == 2));
    const HeadArena a = head_arena(arena, h->M, h->N1);
    int r;
    if ((r = ir_gemm(h->M, h->N1, h->K, x, h->K, 0, h->w1, h->K, 1, a.h1, h->N1, h->b1, 0, 0, stream)) != IR_OK) return r;
    if (h->norm == 1) r = ir_bn_train_fwd(a.h1, nullptr, h->M, h->N1, h->gamma, h->beta, nullptr, 1, h->eps, h->momentum,
                                          h->running_mean, h->running_var, a.bn_scratch, a.stat_a, a.stat_b, a.h2, stream);
    else r = ir_layernorm_fwd(a.h1, h->M, h->N1, h->gamma, h->beta, h->eps, 1, a.h2, a.stat_a, a.stat_b, stream);
    if (r != IR_OK) return r;
    const float* h3 = a.h2;
    if (h->drop_p > 0.f) {
        if ((r = ir_dropout_fwd(a.h2, (int64_t)h->M * h->N1, h->drop_p, h->seed, a.h3, a.mask, stream)) != IR_OK) return r;
        h3 = a.h3;
    }
    return ir_gemm(h->M, h->N2, h->N1, h3, h->N1, 0, h->w2, h->N1, 1, y, h->N2, h->b2, 0, 0, stream);
}

extern "C" int ir_mlp_head_train_bwd(const ir_mlp_head_t* h, const float* x, void* arena, const float* dy, float* dx,
                                     float* dw1, float* db1, float* dgamma, float* dbeta, float* dw2, float* db2,
                                     ir_stream_t stream) {
    IR_CHECK_ARG(h && x && arena && dy && dw1 && db1 && dgamma && dbeta && dw2 && db2);
    const HeadArena a = head_arena(arena, h->M, h->N1);
    const float* h3 = h->drop_p > 0.f ? a.h3 : a.h2;
    int r;
    // second Linear: dW2 = dy^T h3, db2 = colsum(dy), dh3 = dy W2
    if ((r = ir_gemm(h->N2, h->N1, h->M, dy, h->N2, 1, h3, h->N1, 0, dw2, h->N1, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_colsum(dy, h->M, h->N2, db2, stream)) != IR_OK) return r;
    if ((r = ir_gemm(h->M, h->N1, h->N2, dy, h->N2, 0, h->w2, h->N1, 0, a.t1, h->N1, nullptr, 0, 0, stream)) != IR_OK) return r;
    const float* dh2 = a.t1;
    if (h->drop_p > 0.f) {
        if ((r = ir_dropout_bwd(a.t1, a.mask, (int64_t)h->M * h->N1, h->drop_p, a.t2, stream)) != IR_OK) return r;
        dh2 = a.t2;
    }
    // norm + ReLU backward -> dh1 (into the buffer dh2 does not occupy)
    float* dh1 = (dh2 == a.t1) ? a.t2 : a.t1;
    if (h->norm == 1) r = ir_bn_train_bwd(dh2, a.h2, a.h1, nullptr, h->M, h->N1, a.stat_a, a.stat_b, h->gamma, 1, a.bn_scratch,
                                          dh1, nullptr, dgamma, dbeta, nullptr, stream);
    else r = ir_layernorm_bwd(dh2, a.h2, a.h1, h->M, h->N1, h->gamma, a.stat_a, a.stat_b, 1, dh1, dgamma, dbeta, stream);
    if (r != IR_OK) return r;
    // first Linear
    if ((r = ir_gemm(h->N1, h->K, h->M, dh1, h->N1, 1, x, h->K, 0, dw1, h->K, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_colsum(dh1, h->M, h->N1, db1, stream)) != IR_OK) return r;
    if (dx) return ir_gemm(h->M, h->K, h->N1, dh1, h->N1, 0, h->w1, h->K, 0, dx, h->K, nullptr, 0, 0, stream);
    return IR_OK;
}

// ------------------------------------------------------------------ language encoder, train mode, one call per direction
// models/lang_module.py:51-108: word MLP (Linear-ReLU-Dropout-Linear-ReLU) -> 2-layer packed biGRU -> four masked
// attention pools over the projected embeddings -> classifier.  The chain of primitives above behind ONE library call
// forward and ONE backward; intermediates in a caller-owned arena (ir_lang_train_arena_bytes).
struct LangArena {
    float *h1, *h1d, *e, *xp[2], *out[2], *whh[2], *bhh[2], *fcw, *fcb, *atten;
    float *dpooled, *dfeats, *dembed, *dxp, *dhp, *hprev, *dout0, *t1, *t2, *dfcw_part, *dfcb_part;
    uint8_t* mask;
    int64_t bytes;
};
static LangArena lang_arena(void* base, const ir_lang_t* p) {
    LangArena a;
    char* q = (char*)base;
    auto take = [&](int64_t bytes) { char* o = q; q += (bytes + 255) / 256 * 256; return o; };
    const int64_t BL = (int64_t)p->B * p->L, D = p->D, H = p->H;
    a.h1 = (float*)take(BL * D * 4); a.h1d = (float*)take(BL * D * 4); a.e = (float*)take(BL * D * 4);
    for (int l = 0; l < 2; ++l) {
        a.xp[l] = (float*)take(BL * 6 * H * 4); a.out[l] = (float*)take(BL * 2 * H * 4);
        a.whh[l] = (float*)take(2 * 3 * H * H * 4); a.bhh[l] = (float*)take(2 * 3 * H * 4);
    }
    a.fcw = (float*)take(4 * 2 * H * 4); a.fcb = (float*)take(256);
    a.atten = (float*)take(4 * (int64_t)p->B * p->L * 4);
    a.mask = (uint8_t*)take(BL * D);
    a.dpooled = (float*)take(4 * (int64_t)p->B * D * 4);
    a.dfeats = (float*)take(BL * 2 * H * 4); a.dembed = (float*)take(BL * D * 4);
    a.dxp = (float*)take(BL * 6 * H * 4); a.dhp = (float*)take(BL * 6 * H * 4); a.hprev = (float*)take(BL * 2 * H * 4);
    a.dout0 = (float*)take(BL * 2 * H * 4); a.t1 = (float*)take(BL * D * 4); a.t2 = (float*)take(BL * D * 4);
    a.dfcw_part = (float*)take((int64_t)p->B * 4 * 2 * H * 4); a.dfcb_part = (float*)take((int64_t)p->B * 4 * 4);
    a.bytes = q - (char*)base;
    return a;
}
extern "C" int64_t ir_lang_train_arena_bytes(const ir_lang_t* p) { return p ? lang_arena(nullptr, p).bytes : 0; }

#define LANG_CHECK(p) IR_CHECK_ARG((p) && (p)->B > 0 && (p)->L > 0 && (p)->H == 128 && (p)->D == 2 * (p)->H && (p)->E_in > 0 && (p)->n_cls > 0)

extern "C" int ir_lang_train_fwd(const ir_lang_t* p, const float* x, const int64_t* lengths, void* arena, float* pooled,
                                 float* scores, ir_stream_t stream) {
    LANG_CHECK(p);
    IR_CHECK_ARG(x && lengths && arena && pooled);
    const LangArena a = lang_arena(arena, p);
    cudaStream_t st = (cudaStream_t)stream;
    const int BL = p->B * p->L, D = p->D, H = p->H, G3 = 3 * p->H;
    int r;
    if ((r = ir_gemm(BL, D, p->E_in, x, p->E_in, 0, p->w0, p->E_in, 1, a.h1, D, p->b0, 1, 0, stream)) != IR_OK) return r;
    const float* h1d = a.h1;
    if (p->drop_p > 0.f) {
        if ((r = ir_dropout_fwd(a.h1, (int64_t)BL * D, p->drop_p, p->seed, a.h1d, a.mask, stream)) != IR_OK) return r;
        h1d = a.h1d;
    }
    if ((r = ir_gemm(BL, D, D, h1d, D, 0, p->w3, D, 1, a.e, D, p->b3, 1, 0, stream)) != IR_OK) return r;
    for (int l = 0; l < 2; ++l) {
        const float* in = l == 0 ? a.e : a.out[0];
        for (int d = 0; d < 2; ++d) {
            if ((r = ir_gemm(BL, G3, D, in, D, 0, p->wih[l][d], D, 1, a.xp[l] + d * G3, 2 * G3, p->bih[l][d], 0, 0, stream)) != IR_OK) return r;
            IR_CHECK_CUDA(cudaMemcpyAsync(a.whh[l] + (size_t)d * G3 * H, p->whh[l][d], (size_t)G3 * H * 4, cudaMemcpyDeviceToDevice, st));
            IR_CHECK_CUDA(cudaMemcpyAsync(a.bhh[l] + d * G3, p->bhh[l][d], (size_t)G3 * 4, cudaMemcpyDeviceToDevice, st));
        }
        if ((r = ir_gru_layer(a.xp[l], a.whh[l], a.bhh[l], lengths, p->B, p->L, H, a.out[l], stream)) != IR_OK) return r;
    }
    for (int h = 0; h < 4; ++h) {
        IR_CHECK_CUDA(cudaMemcpyAsync(a.fcw + h * D, p->fcw[h], (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
        IR_CHECK_CUDA(cudaMemcpyAsync(a.fcb + h, p->fcb[h], 4, cudaMemcpyDeviceToDevice, st));
    }
    if ((r = ir_token_attention(a.out[1], a.e, (int64_t)p->L * D, lengths, a.fcw, a.fcb, p->B, p->L, D, D, a.atten, pooled, stream)) != IR_OK) return r;
    if (scores && p->wc)
        return ir_gemm(p->B, p->n_cls, D, pooled + (size_t)p->B * D, D, 0, p->wc, D, 1, scores, p->n_cls, p->bc, 0, 0, stream);
    return IR_OK;
}

extern "C" int ir_lang_train_bwd(const ir_lang_t* p, const float* x, const int64_t* lengths, void* arena, const float* pooled,
                                 const float* dpooled, const float* dscores, const ir_lang_grads_t* g, ir_stream_t stream) {
    LANG_CHECK(p);
    IR_CHECK_ARG(x && lengths && arena && pooled && dpooled && g);
    const LangArena a = lang_arena(arena, p);
    cudaStream_t st = (cudaStream_t)stream;
    const int B = p->B, BL = p->B * p->L, D = p->D, H = p->H, G3 = 3 * p->H;
    int r;
    IR_CHECK_CUDA(cudaMemcpyAsync(a.dpooled, dpooled, (size_t)4 * B * D * 4, cudaMemcpyDeviceToDevice, st));
    if (dscores && p->wc) {                                    // classifier on pooled[1]
        if ((r = ir_gemm(p->n_cls, D, B, dscores, p->n_cls, 1, pooled + (size_t)B * D, D, 0, g->dwc, D, nullptr, 0, 0, stream)) != IR_OK) return r;
        if ((r = ir_colsum(dscores, B, p->n_cls, g->dbc, stream)) != IR_OK) return r;
        if ((r = ir_gemm(B, D, p->n_cls, dscores, p->n_cls, 0, p->wc, D, 0, a.dpooled + (size_t)B * D, D, nullptr, 0, 1, stream)) != IR_OK) return r;
    }
    if ((r = ir_token_attention_bwd(a.out[1], a.e, (int64_t)p->L * D, lengths, a.fcw, a.fcb, a.atten, a.dpooled, B, p->L, D, D,
                                    a.dfeats, a.dembed, a.dfcw_part, a.dfcb_part, stream)) != IR_OK) return r;
    if ((r = ir_colsum(a.dfcw_part, B, 4 * D, g->dfcw, stream)) != IR_OK) return r;
    if ((r = ir_colsum(a.dfcb_part, B, 4, g->dfcb, stream)) != IR_OK) return r;
    const float* dout = a.dfeats;
    for (int l = 1; l >= 0; --l) {
        const float* in = l == 0 ? a.e : a.out[0];
        if ((r = ir_gru_layer_bwd(a.xp[l], a.whh[l], a.bhh[l], lengths, a.out[l], dout, B, p->L, H, a.dxp, a.dhp, a.hprev, stream)) != IR_OK) return r;
        if ((r = ir_colsum(a.dhp, BL, 2 * G3, g->dbhh[l], stream)) != IR_OK) return r;          // (2,3H) contiguous
        if ((r = ir_colsum(a.dxp, BL, 2 * G3, g->dbih[l], stream)) != IR_OK) return r;
        float* din = l == 0 ? a.dembed : a.dout0;              // layer 0 adds into the pools' gradient of e
        for (int d = 0; d < 2; ++d) {
            if ((r = ir_gemm(G3, H, BL, a.dhp + d * G3, 2 * G3, 1, a.hprev + d * H, 2 * H, 0, g->dwhh[l][d], H, nullptr, 0, 0, stream)) != IR_OK) return r;
            if ((r = ir_gemm(G3, D, BL, a.dxp + d * G3, 2 * G3, 1, in, D, 0, g->dwih[l][d], D, nullptr, 0, 0, stream)) != IR_OK) return r;
            if ((r = ir_gemm(BL, D, G3, a.dxp + d * G3, 2 * G3, 0, p->wih[l][d], D, 0, din, D, nullptr, 0, (l == 0 || d > 0) ? 1 : 0, stream)) != IR_OK) return r;
        }
        dout = a.dout0;
    }
    // word MLP: e = relu(h1d W3^T + b3), h1 = relu(x W0^T + b0)
    const float* h1d = p->drop_p > 0.f ? a.h1d : a.h1;
    if ((r = ir_relu_bwd(a.dembed, a.e, (int64_t)BL * D, a.t1, stream)) != IR_OK) return r;
    if ((r = ir_gemm(D, D, BL, a.t1, D, 1, h1d, D, 0, g->dw3, D, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_colsum(a.t1, BL, D, g->db3, stream)) != IR_OK) return r;
    if ((r = ir_gemm(BL, D, D, a.t1, D, 0, p->w3, D, 0, a.t2, D, nullptr, 0, 0, stream)) != IR_OK) return r;
    const float* dh1 = a.t2;
    if (p->drop_p > 0.f) {
        if ((r = ir_dropout_bwd(a.t2, a.mask, (int64_t)BL * D, p->drop_p, a.t1, stream)) != IR_OK) return r;
        dh1 = a.t1;
    }
    float* gbuf = (dh1 == a.t1) ? a.t2 : a.t1;
    if ((r = ir_relu_bwd(dh1, a.h1, (int64_t)BL * D, gbuf, stream)) != IR_OK) return r;
    if ((r = ir_gemm(D, p->E_in, BL, gbuf, D, 1, x, p->E_in, 0, g->dw0, p->E_in, nullptr, 0, 0, stream)) != IR_OK) return r;
    return ir_colsum(gbuf, BL, D, g->db0, stream);
}

extern "C" int ir_lang_train_view(const ir_lang_t* p, int64_t* off_feats, int64_t* off_atten) {
    LANG_CHECK(p);
    IR_CHECK_ARG(off_feats && off_atten);
    const LangArena a = lang_arena(nullptr, p);
    *off_feats = (char*)a.out[1] - (char*)nullptr;
    *off_atten = (char*)a.atten - (char*)nullptr;
    return IR_OK;
}

// ------------------------------------------------------------------ DynamicEdgeConv, train mode, one call per direction
// models/basic_blocks.py:98-133 given the kNN lists: per edge w = W2w relu(W1w [p_j-p_i, oh_i, oh_j]),
// msg = W2m relu(W1m [x_i, w, x_j]); out_i = max_j msg.  x carries no gradient (instance statistics), so the backward
// ends at the two edge MLPs' parameters.  Arena: ir_edgeconv_train_arena_bytes.
struct EdgeArena {
    float *w_in, *a1, *w, *e_in, *m1, *msg, *t1, *t2, *t3, *t4;
    int* arg;
    int64_t bytes;
};
static EdgeArena edge_arena(void* base, const ir_edgeconv_t* p) {
    EdgeArena a;
    char* q = (char*)base;
    auto take = [&](int64_t bytes) { char* o = q; q += (bytes + 255) / 256 * 256; return o; };
    const int64_t E = (int64_t)p->nq * p->k;
    a.w_in = (float*)take(E * (3 + 2 * p->ncls) * 4); a.a1 = (float*)take(E * p->H1 * 4); a.w = (float*)take(E * p->F * 4);
    a.e_in = (float*)take(E * 3 * p->F * 4); a.m1 = (float*)take(E * p->Fout * 4); a.msg = (float*)take(E * p->Fout * 4);
    a.arg = (int*)take((int64_t)p->nq * p->Fout * 4);
    a.t1 = (float*)take(E * p->Fout * 4); a.t2 = (float*)take(E * p->Fout * 4);
    a.t3 = (float*)take(E * p->H1 * 4); a.t4 = (float*)take(E * p->H1 * 4);
    a.bytes = q - (char*)base;
    return a;
}
extern "C" int64_t ir_edgeconv_train_arena_bytes(const ir_edgeconv_t* p) { return p ? edge_arena(nullptr, p).bytes : 0; }
#define EDGE_CHECK(p) IR_CHECK_ARG((p) && (p)->nq > 0 && (p)->k > 0 && (p)->F > (p)->ncls && (p)->ncls > 0 && (p)->H1 > 0 && (p)->Fout > 0 && (p)->F <= (p)->H1)

extern "C" int ir_edgeconv_train_fwd(const ir_edgeconv_t* p, const float* x, const float* xyz, const int32_t* qidx,
                                     const int32_t* nbr, void* arena, float* out, ir_stream_t stream) {
    EDGE_CHECK(p);
    IR_CHECK_ARG(x && xyz && qidx && nbr && arena && out);
    const EdgeArena a = edge_arena(arena, p);
    const int E = p->nq * p->k, Dw = 3 + 2 * p->ncls, De = 3 * p->F;
    int r;
    if ((r = ir_edge_inputs(x, xyz, qidx, nbr, p->nq, p->k, p->F, p->ncls, nullptr, a.w_in, nullptr, stream)) != IR_OK) return r;
    if ((r = ir_gemm(E, p->H1, Dw, a.w_in, Dw, 0, p->ww1, Dw, 1, a.a1, p->H1, p->bw1, 1, 0, stream)) != IR_OK) return r;
    if ((r = ir_gemm(E, p->F, p->H1, a.a1, p->H1, 0, p->ww2, p->H1, 1, a.w, p->F, p->bw2, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_edge_inputs(x, xyz, qidx, nbr, p->nq, p->k, p->F, p->ncls, a.w, nullptr, a.e_in, stream)) != IR_OK) return r;
    if ((r = ir_gemm(E, p->Fout, De, a.e_in, De, 0, p->wm1, De, 1, a.m1, p->Fout, p->bm1, 1, 0, stream)) != IR_OK) return r;
    if ((r = ir_gemm(E, p->Fout, p->Fout, a.m1, p->Fout, 0, p->wm2, p->Fout, 1, a.msg, p->Fout, p->bm2, 0, 0, stream)) != IR_OK) return r;
    return ir_edge_max_fwd(a.msg, nbr, p->nq, p->k, p->Fout, out, a.arg, stream);
}

extern "C" int ir_edgeconv_train_bwd(const ir_edgeconv_t* p, void* arena, const float* dout, const ir_edgeconv_grads_t* g,
                                     ir_stream_t stream) {
    EDGE_CHECK(p);
    IR_CHECK_ARG(arena && dout && g);
    const EdgeArena a = edge_arena(arena, p);
    const int E = p->nq * p->k, Dw = 3 + 2 * p->ncls, De = 3 * p->F, Fo = p->Fout;
    int r;
    if ((r = ir_edge_max_bwd(dout, a.arg, p->nq, p->k, Fo, a.t1, stream)) != IR_OK) return r;                 // dmsg (0 on unused edges)
    if ((r = ir_gemm(Fo, Fo, E, a.t1, Fo, 1, a.m1, Fo, 0, g->dwm2, Fo, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_colsum(a.t1, E, Fo, g->dbm2, stream)) != IR_OK) return r;
    if ((r = ir_gemm(E, Fo, Fo, a.t1, Fo, 0, p->wm2, Fo, 0, a.t2, Fo, nullptr, 0, 0, stream)) != IR_OK) return r;   // dm1
    if ((r = ir_relu_bwd(a.t2, a.m1, (int64_t)E * Fo, a.t1, stream)) != IR_OK) return r;                       // g1
    if ((r = ir_gemm(Fo, De, E, a.t1, Fo, 1, a.e_in, De, 0, g->dwm1, De, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_colsum(a.t1, E, Fo, g->dbm1, stream)) != IR_OK) return r;
    // dw = g1 @ W1m[:, F:2F]  (only the w slice of [x_i, w, x_j] carries a gradient)
    if ((r = ir_gemm(E, p->F, Fo, a.t1, Fo, 0, p->wm1 + p->F, De, 0, a.t3, p->F, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_gemm(p->F, p->H1, E, a.t3, p->F, 1, a.a1, p->H1, 0, g->dww2, p->H1, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_colsum(a.t3, E, p->F, g->dbw2, stream)) != IR_OK) return r;
    if ((r = ir_gemm(E, p->H1, p->F, a.t3, p->F, 0, p->ww2, p->H1, 0, a.t4, p->H1, nullptr, 0, 0, stream)) != IR_OK) return r;   // da1
    if ((r = ir_relu_bwd(a.t4, a.a1, (int64_t)E * p->H1, a.t2, stream)) != IR_OK) return r;                    // g2 (t2 is large enough)
    if ((r = ir_gemm(p->H1, Dw, E, a.t2, p->H1, 1, a.w_in, Dw, 0, g->dww1, Dw, nullptr, 0, 0, stream)) != IR_OK) return r;
    return ir_colsum(a.t2, E, p->H1, g->dbw1, stream);
}

// ------------------------------------------------------------------ scene tail, train mode, one call per direction
// models/scene_module.py:25-38,70-71: SparseCrop + ToDenseBEVConvolution -> BatchNorm2d (batch statistics) -> ReLU ->
// Conv2d 3x3 -> BatchNorm2d -> ReLU -> Dropout -> Conv2d 3x3, NHWC, as a chain of the primitives above behind one
// library call per direction.  Arena: ir_scene_tail_arena_bytes.
__global__ void k_pack_conv_w(const float* __restrict__ w, int cout, int cin, float* __restrict__ wp, int unpack) {
    // pack: wp[(tap*cin + ci)*cout + co] = w[((co*cin + ci)*9) + tap];  unpack: the inverse (w <- wp)
    const int total = cout * cin * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout, t = i / cout, ci = t % cin, tap = t / cin;
        const int j = (co * cin + ci) * 9 + tap;
        if (unpack) const_cast<float*>(w)[j] = wp[i]; else wp[i] = w[j];
    }
}
extern "C" int ir_bev(const float*, const int32_t*, const int32_t*, int64_t, int32_t, const float*, const float*, const float*,
                      int32_t, float*, int32_t*, float*, float*, ir_stream_t);

struct SceneArena {
    float *tmp, *dense, *a0, *st0a, *st0b, *wp1, *col1, *y1, *a1, *st1a, *st1b, *a1d, *wp2, *col2, *bn_scratch;
    float *dcol, *t1, *t2, *dwp;
    int* cell;
    uint8_t* mask;
    int64_t bytes;
};
static SceneArena scene_arena(void* base, int64_t n_rows, int64_t B) {
    SceneArena a;
    char* q = (char*)base;
    auto take = [&](int64_t bytes) { char* o = q; q += (bytes + 255) / 256 * 256; return o; };
    const int64_t C = 128, r0 = B * 375, r1 = B * 299, r2 = B * 231;
    a.tmp = (float*)take((n_rows > 0 ? n_rows : 1) * C * 4); a.cell = (int*)take((n_rows > 0 ? n_rows : 1) * 4);
    a.dense = (float*)take(r0 * C * 4); a.a0 = (float*)take(r0 * C * 4); a.st0a = (float*)take(C * 4); a.st0b = (float*)take(C * 4);
    a.wp1 = (float*)take(9 * C * C * 4); a.col1 = (float*)take(r1 * 9 * C * 4); a.y1 = (float*)take(r1 * C * 4); a.a1 = (float*)take(r1 * C * 4);
    a.st1a = (float*)take(C * 4); a.st1b = (float*)take(C * 4); a.a1d = (float*)take(r1 * C * 4); a.mask = (uint8_t*)take(r1 * C);
    a.wp2 = (float*)take(9 * C * C * 4); a.col2 = (float*)take(r2 * 9 * C * 4);
    a.bn_scratch = (float*)take(ir_bn_scratch_floats(256) * 4);
    a.dcol = (float*)take(r1 * 9 * C * 4); a.t1 = (float*)take(r0 * C * 4); a.t2 = (float*)take(r0 * C * 4); a.dwp = (float*)take(9 * C * C * 4);
    a.bytes = q - (char*)base;
    return a;
}
extern "C" int64_t ir_scene_tail_arena_bytes(int64_t n_rows, int32_t B) { return scene_arena(nullptr, n_rows, B).bytes; }
#define SCENE_CHECK(p) IR_CHECK_ARG((p) && (p)->B > 0 && (p)->n_rows > 0 && (p)->kernel && (p)->w1 && (p)->w2)

extern "C" int ir_scene_tail_train_fwd(const ir_scene_tail_t* p, const float* f4, const int32_t* coords, const int32_t* n_dev,
                                       void* arena, float* out, ir_stream_t stream) {
    SCENE_CHECK(p);
    IR_CHECK_ARG(f4 && coords && n_dev && arena && out);
    const SceneArena a = scene_arena(arena, p->n_rows, p->B);
    cudaStream_t st = (cudaStream_t)stream;
    const int B = p->B, C = 128, r0 = B * 375, r1 = B * 299, r2 = B * 231;
    int r;
    if ((r = ir_bev(f4, coords, n_dev, p->n_rows, 16, p->kernel, nullptr, nullptr, B, a.tmp, a.cell, a.dense, nullptr, stream)) != IR_OK) return r;
    if ((r = ir_bn_train_fwd(a.dense, nullptr, r0, C, p->g0, p->be0, nullptr, 1, p->eps, p->mom0, p->rm0, p->rv0, a.bn_scratch,
                             a.st0a, a.st0b, a.a0, stream)) != IR_OK) return r;
    k_pack_conv_w<<<IR_NUM_SMS * 2, 256, 0, st>>>(p->w1, C, C, a.wp1, 0);
    IR_CHECK_LAUNCH();
    if ((r = ir_im2col_3x3(a.a0, B, 15, 25, C, a.col1, stream)) != IR_OK) return r;
    if ((r = ir_gemm(r1, C, 9 * C, a.col1, 9 * C, 0, a.wp1, C, 0, a.y1, C, p->b1, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_bn_train_fwd(a.y1, nullptr, r1, C, p->g1, p->be1, nullptr, 1, p->eps, p->mom1, p->rm1, p->rv1, a.bn_scratch,
                             a.st1a, a.st1b, a.a1, stream)) != IR_OK) return r;
    const float* a1d = a.a1;
    if (p->drop_p > 0.f) {
        if ((r = ir_dropout_fwd(a.a1, (int64_t)r1 * C, p->drop_p, p->seed, a.a1d, a.mask, stream)) != IR_OK) return r;
        a1d = a.a1d;
    }
    k_pack_conv_w<<<IR_NUM_SMS * 2, 256, 0, st>>>(p->w2, C, C, a.wp2, 0);
    IR_CHECK_LAUNCH();
    if ((r = ir_im2col_3x3(a1d, B, 13, 23, C, a.col2, stream)) != IR_OK) return r;
    return ir_gemm(r2, C, 9 * C, a.col2, 9 * C, 0, a.wp2, C, 0, out, C, p->b2, 0, 0, stream);
}

extern "C" int ir_scene_tail_train_bwd(const ir_scene_tail_t* p, const float* f4, const int32_t* coords, const int32_t* n_dev,
                                       void* arena, const float* dout, float* df4, const ir_scene_tail_grads_t* g,
                                       ir_stream_t stream) {
    SCENE_CHECK(p);
    IR_CHECK_ARG(f4 && coords && n_dev && arena && dout && df4 && g);
    const SceneArena a = scene_arena(arena, p->n_rows, p->B);
    cudaStream_t st = (cudaStream_t)stream;
    const int B = p->B, C = 128, r0 = B * 375, r1 = B * 299, r2 = B * 231;
    int r;
    // conv 2
    if ((r = ir_gemm(9 * C, C, r2, a.col2, 9 * C, 1, dout, C, 0, a.dwp, C, nullptr, 0, 0, stream)) != IR_OK) return r;
    k_pack_conv_w<<<IR_NUM_SMS * 2, 256, 0, st>>>(g->dw2, C, C, a.dwp, 1);
    IR_CHECK_LAUNCH();
    if ((r = ir_colsum(dout, r2, C, g->db2, stream)) != IR_OK) return r;
    if ((r = ir_gemm(r2, 9 * C, C, dout, C, 0, a.wp2, C, 1, a.dcol, 9 * C, nullptr, 0, 0, stream)) != IR_OK) return r;
    if ((r = ir_col2im_3x3(a.dcol, B, 13, 23, C, a.t1, stream)) != IR_OK) return r;                        // d a1d
    const float* da1 = a.t1;
    if (p->drop_p > 0.f) {
        if ((r = ir_dropout_bwd(a.t1, a.mask, (int64_t)r1 * C, p->drop_p, a.t2, stream)) != IR_OK) return r;
        da1 = a.t2;
    }
    float* dy1 = (da1 == a.t1) ? a.t2 : a.t1;
    if ((r = ir_bn_train_bwd(da1, a.a1, a.y1, nullptr, r1, C, a.st1a, a.st1b, p->g1, 1, a.bn_scratch, dy1, nullptr, g->dg1, g->dbe1,
                             nullptr, stream)) != IR_OK) return r;
    // conv 1
    if ((r = ir_gemm(9 * C, C, r1, a.col1, 9 * C, 1, dy1, C, 0, a.dwp, C, nullptr, 0, 0, stream)) != IR_OK) return r;
    k_pack_conv_w<<<IR_NUM_SMS * 2, 256, 0, st>>>(g->dw1, C, C, a.dwp, 1);
    IR_CHECK_LAUNCH();
    if ((r = ir_colsum(dy1, r1, C, g->db1, stream)) != IR_OK) return r;
    if ((r = ir_gemm(r1, 9 * C, C, dy1, C, 0, a.wp1, C, 1, a.dcol, 9 * C, nullptr, 0, 0, stream)) != IR_OK) return r;
    float* da0 = (dy1 == a.t1) ? a.t2 : a.t1;
    if ((r = ir_col2im_3x3(a.dcol, B, 15, 25, C, da0, stream)) != IR_OK) return r;
    float* dd = (da0 == a.t1) ? a.t2 : a.t1;
    if ((r = ir_bn_train_bwd(da0, a.a0, a.dense, nullptr, r0, C, a.st0a, a.st0b, p->g0, 1, a.bn_scratch, dd, nullptr, g->dg0, g->dbe0,
                             nullptr, stream)) != IR_OK) return r;
    return ir_bev_bwd(dd, f4, coords, a.cell, n_dev, p->n_rows, 16, p->kernel, 5, df4, g->dkernel, stream);
}
