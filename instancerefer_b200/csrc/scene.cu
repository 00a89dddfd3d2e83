// Scene head kernels: SparseCrop + ToDenseBEVConvolution (+BN2d+ReLU), Conv2d 3x3, language-guided
// attention over BEV cells.  Reference: models/basic_blocks.py:174-243, models/scene_module.py:25-38,
// 69-83.  Activations are NHWC; all reductions run in a fixed order (deterministic).
#include "../../include/instancerefer_b200.h"
#include <string.h>

#include "common.cuh"

#define BEV_X 240
#define BEV_Y 400
#define BEV_Z 80
#define BEV_H 15
#define BEV_W 25
#define BEV_C 128

// one CTA (128 threads = output channels) per voxel row: crop test, z-slice matvec
__global__ void __launch_bounds__(BEV_C)
k_bev_matvec(const float* __restrict__ F, const int4* __restrict__ coords, const int* __restrict__ n_dev,
             int stride, const float* __restrict__ kern, int B, float* __restrict__ tmp,
             int* __restrict__ cell) {
    __shared__ float f[BEV_C];
    const int n = *n_dev;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int4 c = coords[r];
        const bool keep = c.x >= 0 && c.y >= 0 && c.z >= 0 && c.x < BEV_X && c.y < BEV_Y && c.z < BEV_Z &&
                          c.w >= 0 && c.w < B;
        if (!keep) {
            if (threadIdx.x == 0) cell[r] = -1;
            continue;
        }
        __syncthreads();
        f[threadIdx.x] = F[(long long)r * BEV_C + threadIdx.x];
        __syncthreads();
        const float* kz = kern + (long long)(c.z / stride) * BEV_C * BEV_C + threadIdx.x;
        // 128 taps in 4 batches of 32 independent (coalesced, L2-resident) weight loads: 4 memory latencies per row
        // instead of 16; the additions keep their order (ci ascending)
        float acc = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BEV_C; c0 += 32) {
            float wv[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) wv[u] = __ldg(kz + (long long)(c0 + u) * BEV_C);
#pragma unroll
            for (int u = 0; u < 32; ++u) acc = fmaf(f[c0 + u], wv[u], acc);
        }
        tmp[(long long)r * BEV_C + threadIdx.x] = acc;
        if (threadIdx.x == 0) cell[r] = c.w * (BEV_H * BEV_W) + (c.x / stride) * BEV_W + (c.y / stride);
    }
}

// one CTA per dense cell: sum matching rows in ascending row order, BN2d affine, ReLU
__global__ void __launch_bounds__(BEV_C)
k_bev_cell(const float* __restrict__ tmp, const int* __restrict__ cell, const int* __restrict__ n_dev,
           const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out,
           float* __restrict__ out_absmax) {
    __shared__ int s_list[16];
    __shared__ int s_wcnt[BEV_C / 32];
    __shared__ int s_total;
    const int n = *n_dev;
    const int me = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int r0 = 0; r0 < n; r0 += BEV_C) {
        const int r = r0 + tid;
        const bool m = (r < n) && (cell[r] == me);
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) s_wcnt[w] = __popc(bal);
        __syncthreads();
        int base = s_total;
        for (int i = 0; i < w; ++i) base += s_wcnt[i];
        if (m) {
            const int p = base + __popc(bal & ((1u << lane) - 1u));
            if (p < 16) s_list[p] = r;
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_total;
            for (int i = 0; i < BEV_C / 32; ++i) t += s_wcnt[i];
            s_total = t;
        }
        __syncthreads();
    }
    const int cnt = min(s_total, 16);
    float acc = 0.f;
    for (int i = 0; i < cnt; ++i) acc += tmp[(long long)s_list[i] * BEV_C + tid];
    // scale == NULL: raw sums (train mode: batch-statistics BN2d + ReLU follow as their own kernel)
    const float y = scale ? fmaxf(fmaf(acc, scale[tid], shift[tid]), 0.f) : acc;
    out[(long long)me * BEV_C + tid] = y;
    if (out_absmax) {                      // range guard of the tcgen05 Conv2d that consumes the BEV map
        const float m = warp_max(fabsf(y));
        if (lane == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned*>(out_absmax), __float_as_uint(m));
    }
}

extern "C" int ir_bev(const float* feats, const int32_t* coords, const int32_t* n_dev, int64_t n_max,
                      int32_t stride, const float* kernel, const float* bn_scale,
                      const float* bn_shift, int32_t B, float* tmp, int32_t* cell, float* out,
                      float* out_absmax, ir_stream_t stream) {
    IR_CHECK_ARG(feats && coords && n_dev && kernel && tmp && cell && out && ((bn_scale == nullptr) == (bn_shift == nullptr)));
    IR_CHECK_ARG(stride == 16 && B > 0 && n_max > 0);
    cudaStream_t st = (cudaStream_t)stream;
    k_bev_matvec<<<ir_min_i(n_max, IR_NUM_SMS * 16), BEV_C, 0, st>>>(feats, (const int4*)coords, n_dev, stride,
                                                                     kernel, B, tmp, cell);
    IR_CHECK_LAUNCH();
    k_bev_cell<<<B * BEV_H * BEV_W, BEV_C, 0, st>>>(tmp, cell, n_dev, bn_scale, bn_shift, out, out_absmax);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ conv2d 3x3 (valid), NHWC, C=128
// CTA = 4 output pixels x 128 output channels x 8 K-slices (1024 threads): every thread walks 144 of
// the 1152 (ky,kx,cin) taps with 8 independent coalesced weight loads in flight; slices are combined
// through shared memory in a fixed order (deterministic).
#define C2_TPX 4
#define C2_KS 8   // K slices
__global__ void __launch_bounds__(BEV_C * C2_KS)
k_conv2d_3x3(const float* __restrict__ in, int H, int W, const float* __restrict__ wpack,
             const float* __restrict__ bias, const float* __restrict__ scale,
             const float* __restrict__ shift, int relu, float* __restrict__ out) {
    constexpr int C = BEV_C;
    constexpr int TAPS = 9 * C;                 // 1152 taps, tap t = (ky*3+kx)*C + cin
    constexpr int PER = TAPS / C2_KS;           // 144 taps per slice
    __shared__ float patch[3][C2_TPX + 2][C];
    __shared__ float red[C2_KS][C2_TPX][C];
    const int Ho = H - 2, Wo = W - 2;
    const int x0 = blockIdx.x * C2_TPX, y = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, co = tid & (C - 1), ks = tid / C;
    for (int i = tid; i < 3 * (C2_TPX + 2) * C; i += C * C2_KS) {
        const int c = i % C, px = (i / C) % (C2_TPX + 2), py = i / (C * (C2_TPX + 2));
        const int xx = x0 + px;
        patch[py][px][c] = (xx < W) ? in[(((long long)b * H + (y + py)) * W + xx) * C + c] : 0.f;
    }
    __syncthreads();
    float acc[C2_TPX];
#pragma unroll
    for (int p = 0; p < C2_TPX; ++p) acc[p] = 0.f;
    const int t0 = ks * PER;
    const float* wk = wpack + (long long)t0 * C + co;
    // 144 taps in 6 batches of 24 independent coalesced weight loads (6 memory latencies per thread instead of 18);
    // the order of the additions is unchanged
    constexpr int UB = 24;
    static_assert(PER % UB == 0, "tap batches");
#pragma unroll 1
    for (int i0 = 0; i0 < PER; i0 += UB) {
        float wv[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) wv[u] = __ldg(wk + (long long)(i0 + u) * C);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int t = t0 + i0 + u;
            const int kk = t / C, ci = t - kk * C;
            const int ky = kk / 3, kx = kk - ky * 3;
#pragma unroll
            for (int p = 0; p < C2_TPX; ++p) acc[p] = fmaf(wv[u], patch[ky][p + kx][ci], acc[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < C2_TPX; ++p) red[ks][p][co] = acc[p];
    __syncthreads();
    if (ks < C2_TPX) {                          // slice index reused as the pixel this thread finalises
        const int p = ks;
        if (x0 + p < Wo) {
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < C2_KS; ++q) v += red[q][p][co];
            v += bias ? bias[co] : 0.f;
            if (scale) v = fmaf(v, scale[co], shift[co]);
            if (relu) v = fmaxf(v, 0.f);
            out[(((long long)b * Ho + y) * Wo + (x0 + p)) * C + co] = v;
        }
    }
}

extern "C" int ir_conv2d_3x3(const float* in, int32_t B, int32_t H, int32_t W, int32_t C,
                             const float* wpack, const float* bias, const float* scale,
                             const float* shift, int32_t relu, float* out, ir_stream_t stream) {
    IR_CHECK_ARG(in && wpack && out && C == BEV_C && H >= 3 && W >= 3 && B > 0);
    dim3 grid(ir_div_up(W - 2, C2_TPX), H - 2, B);
    k_conv2d_3x3<<<grid, BEV_C * C2_KS, 0, (cudaStream_t)stream>>>(in, H, W, wpack, bias, scale, shift, relu, out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ conv2d 3x3 on the tcgen05 rule GEMM
// A 3x3 valid convolution over a dense (B,H,W) grid IS a sparse convolution whose rulebook is known in closed form
// (tap k = ky*3+kx of output pixel o reads input pixel o + ky*W + kx; every output has all nine pairs), with the
// conv weight repacked to (9, Cin, Cout).  So the two Conv2d of the scene head run through the same weight-stationary
// tcgen05 pair-GEMM + ordered reduce as the sparse layers (range-scaled split-fp16, bias / folded BN / ReLU in the
// reduce epilogue) instead of a SIMT direct convolution: 9 x 299 pairs = 43 tiles for the first conv.
// Replaces nn.Conv2d x2 of models/scene_module.py:33-38.  The caller builds the rulebook once per grid shape.
#include "kernels.cuh"
extern "C" int ir_conv2d_3x3_tc(const float* in, const int32_t* in_idx, const int32_t* slot, const int32_t* count,
                                const int32_t* n_out_dev, int64_t n_out, const float* wprep, const float* scale,
                                const float* shift, int32_t relu, const float* in_absmax, float* out_absmax, float* T,
                                float* out, ir_stream_t stream) {
    IR_CHECK_ARG(in && in_idx && slot && count && n_out_dev && wprep && T && out && n_out > 0);
    IR_CHECK_ARG((reinterpret_cast<uintptr_t>(wprep) & 15) == 0);
    IrConvBatch b;
    memset(&b, 0, sizeof(b));
    b.G = 1;
    b.p[0] = IrConvProblem{in, in_idx, slot, count, n_out_dev, wprep, scale, shift, nullptr, T, out,
                           (long long)n_out, (long long)n_out, relu, in_absmax, out_absmax};
    int r = irk_pairgemm_tc(b, BEV_C, BEV_C, 9, (cudaStream_t)stream);
    if (r != IR_OK) return r;
    return irk_reduce_epilogue(b, BEV_C, 9, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ attention over BEV cells
__global__ void __launch_bounds__(256)
k_scene_attention(const float* __restrict__ feats, const float* __restrict__ q, int ncell, int C,
                  float* __restrict__ atten, float* __restrict__ scene_feat) {
    extern __shared__ float s_att[];             // ncell logits
    __shared__ float s_red[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* fb = feats + (long long)b * ncell * C;
    const float* qb = q + (long long)b * C;
    const float inv = 1.0f / sqrtf((float)C);
    // logits: a warp takes 4 cells at a time (their loads and shuffle reductions are independent: one memory latency per
    // 4 cells); per-cell arithmetic order is unchanged
    for (int cell0 = w * 4; cell0 < ncell; cell0 += 32) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (cell0 + u < ncell)
                for (int c = lane; c < C; c += 32) a[u] = fmaf(fb[(long long)(cell0 + u) * C + c], qb[c], a[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = warp_sum(a[u]);
        if (lane < 4 && cell0 + lane < ncell) s_att[cell0 + lane] = (lane == 0 ? a[0] : lane == 1 ? a[1] : lane == 2 ? a[2] : a[3]) * inv;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int i = tid; i < ncell; i += 256) m = fmaxf(m, s_att[i]);
    m = warp_max(m);
    if (lane == 0) s_red[w] = m;
    __syncthreads();
    m = s_red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = tid; i < ncell; i += 256) {
        const float e = expf(s_att[i] - m);
        s_att[i] = e;
        s += e;
    }
    s = warp_sum(s);
    if (lane == 0) s_red[w] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_red[i];
    const float rs = 1.0f / s;
    for (int i = tid; i < ncell; i += 256) {
        const float a = s_att[i] * rs;
        s_att[i] = a;
        atten[(long long)b * ncell + i] = a;
    }
    __syncthreads();
    // weighted sum: warp w takes cells w, w+8, ... for all channels (lane = 4 consecutive channels, one 512-byte row per
    // load instruction, 8 rows in flight), the 8 per-warp partials are added in warp order (deterministic)
    __shared__ float s_part[8][BEV_C];
    if (C == BEV_C) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = w; c0 < ncell; c0 += 64) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int cell = c0 + 8 * u;
                v[u] = (cell < ncell) ? *reinterpret_cast<const float4*>(fb + (long long)cell * C + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int cell = c0 + 8 * u;
                const float a = (cell < ncell) ? s_att[cell] : 0.f;
                acc.x = fmaf(v[u].x, a, acc.x); acc.y = fmaf(v[u].y, a, acc.y);
                acc.z = fmaf(v[u].z, a, acc.z); acc.w = fmaf(v[u].w, a, acc.w);
            }
        }
        *reinterpret_cast<float4*>(&s_part[w][lane * 4]) = acc;
        __syncthreads();
        if (tid < BEV_C) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += s_part[q][tid];
            scene_feat[(long long)b * C + tid] = t;
        }
    } else {
        for (int c = tid; c < C; c += 256) {
            float acc = 0.f;
            for (int cell = 0; cell < ncell; ++cell) acc = fmaf(fb[(long long)cell * C + c], s_att[cell], acc);
            scene_feat[(long long)b * C + c] = acc;
        }
    }
}

extern "C" int ir_scene_attention(const float* feats, const float* q, int32_t B, int32_t ncell,
                                  int32_t C, float* atten, float* scene_feat, ir_stream_t stream) {
    IR_CHECK_ARG(feats && q && atten && scene_feat && B > 0 && ncell > 0 && ncell <= 8192 && C > 0);
    k_scene_attention<<<B, 256, ncell * sizeof(float), (cudaStream_t)stream>>>(feats, q, ncell, C, atten, scene_feat);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
