// Training-step kernels of the sparse-voxel encoders: what torchsparse's
// sparseconv_backward host loop (two GEMMs per kernel offset), spnn.BatchNorm in train mode and the
// backward of spnn.GlobalMaxPooling do for models/basic_blocks.py:10-56 and
// models/attribute_module.py:105.
//
//   dgrad : dX[i] = sum_k dY[o] @ W[k]^T over pairs (i,o,k)  == the forward pair-GEMM + reduce run on
//           the TRANSPOSED rulebook (k_rulebook_transpose) with W^T — no new GEMM kernel.
//   wgrad : dW[k] = sum_pairs X[i]^T dY[o]                   (k_wgrad: per-offset gathered outer products)
//   BN    : batch statistics over the live rows, normalise (+ residual, ReLU) and the matching backward.
#include "../../include/instancerefer_b200.h"
#include "common.cuh"
#include "kernels.cuh"

// ------------------------------------------------------------------ rulebook transpose
// forward rulebook of a map: in_idx[k][pos] = input row of pair pos, slot[k][o] = pair of output row o.
// transposed:                out_idx[k][pos] = output row of pair pos, slot_in[k][i] = pair of input row i
// (inside one offset every input row occurs at most once, exactly like every output row).
__global__ void k_rulebook_transpose(const int* __restrict__ in_idx, const int* __restrict__ slot,
                                     long long seg_cap, const int* __restrict__ n_out_dev,
                                     int* __restrict__ out_idx, int* __restrict__ slot_in) {
    const int n = *n_out_dev;
    const long long base = (long long)blockIdx.y * seg_cap;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        const int pos = slot[base + o];
        if (pos >= 0) {
            out_idx[base + pos] = o;
            slot_in[base + in_idx[base + pos]] = pos;
        }
    }
}

extern "C" int ir_rulebook_transpose(const int32_t* in_idx, const int32_t* slot, int32_t K,
                                     int64_t seg_cap, const int32_t* n_out_dev, int64_t n_max,
                                     int32_t* out_idx, int32_t* slot_in, ir_stream_t stream) {
    IR_CHECK_ARG(in_idx && slot && n_out_dev && out_idx && slot_in && K > 0 && K <= 32 && seg_cap > 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(slot_in, 0xFF, (size_t)K * seg_cap * 4, st));
    const dim3 grid(ir_min_i(ir_div_up(n_max > 0 ? n_max : 1, 256), IR_NUM_SMS * 4), K);
    k_rulebook_transpose<<<grid, 256, 0, st>>>(in_idx, slot, seg_cap, n_out_dev, out_idx, slot_in);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ wgrad
// grid (nsplit, K), 256 threads as a 16 x 16 grid over (ci, co); thread (ty,tx) owns the MI x MJ
// entries ci = i*16+ty, co = j*16+tx of dW[k] (interleaved so shared-memory reads are conflict-free
// and the final atomics touch consecutive floats).  Pairs of the CTA's chunk are staged 16 at a time.
#define WG_PB 16
template <int V> struct WgVec;
template <> struct WgVec<4> { typedef float4 T; };
template <> struct WgVec<2> { typedef float2 T; };
template <int V> __device__ __forceinline__ void wg_load(const float* p, float* dst) {
    const typename WgVec<V>::T v = *reinterpret_cast<const typename WgVec<V>::T*>(p);
    if (V == 4) { const float4 q = *reinterpret_cast<const float4*>(&v); dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w; }
    else { const float2 q = *reinterpret_cast<const float2*>(&v); dst[0] = q.x; dst[1] = q.y; }
}
// Thread (ty,tx) owns ci = (i/VI)*16*VI + ty*VI + i%VI and co = (j/VJ)*16*VJ + tx*VJ + j%VJ: its operands are
// VI/VJ-wide contiguous groups, so one 16-byte shared-memory load feeds 4 rows/columns of the outer product
// (the warp's 16 tx lanes read 256 contiguous bytes: conflict-free; the two ty values broadcast).
// Gathered rows are staged by a 3-deep cp.async pipeline (16 pairs per stage): the L2 round trip of the
// next two batches overlaps the outer products of the current one.
#define WG_STAGES 3
__device__ __forceinline__ void wg_cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;                                        // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
k_wgrad(const float* __restrict__ X, const float* __restrict__ dY, const int* __restrict__ in_idx,
        const int* __restrict__ out_idx, const int* __restrict__ count, long long seg_cap,
        float* __restrict__ dW) {
    constexpr int MI = CIN / 16, MJ = COUT / 16;
    constexpr int VI = MI < 4 ? MI : 4, VJ = MJ < 4 ? MJ : 4;
    extern __shared__ __align__(16) float wg_smem[];
    float (*xs)[WG_PB][CIN] = reinterpret_cast<float (*)[WG_PB][CIN]>(wg_smem);
    float (*ds)[WG_PB][COUT] = reinterpret_cast<float (*)[WG_PB][COUT]>(wg_smem + WG_STAGES * WG_PB * CIN);
    const int k = blockIdx.y;
    const int cnt = count[k];
    const int chunk = (cnt + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * chunk, p_end = min(cnt, p_begin + chunk);
    if (p_begin >= p_end) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[MI][MJ];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < MJ; ++j) acc[i][j] = 0.f;
    const long long base = (long long)k * seg_cap;
    const int nb = (p_end - p_begin + WG_PB - 1) / WG_PB;
    // rulebook indices are fetched one batch AHEAD of the cp.async that consumes them, so the issue step never
    // waits on an L2 round trip: a thread stages (CIN+COUT)/64 row chunks per batch
    constexpr int NX = WG_PB * (CIN / 4) / 256, ND = WG_PB * (COUT / 4) / 256;     // chunks per thread (1..2 each)
    constexpr int NXr = NX > 0 ? NX : 1, NDr = ND > 0 ? ND : 1;
    int ix[NXr], id[NDr];
    auto load_idx = [&](int bi) {
        const int p0 = p_begin + bi * WG_PB;
#pragma unroll
        for (int q = 0; q < NXr; ++q) {
            const int i = tid + q * 256, r = i / (CIN / 4);
            ix[q] = (bi < nb && i < WG_PB * (CIN / 4) && p0 + r < p_end) ? in_idx[base + p0 + r] : -1;
        }
#pragma unroll
        for (int q = 0; q < NDr; ++q) {
            const int i = tid + q * 256, r = i / (COUT / 4);
            id[q] = (bi < nb && i < WG_PB * (COUT / 4) && p0 + r < p_end) ? out_idx[base + p0 + r] : -1;
        }
    };
    auto issue = [&](int bi) {                       // uses ix/id loaded for batch bi
        if (bi < nb) {
            const int st = bi % WG_STAGES;
#pragma unroll
            for (int q = 0; q < NXr; ++q) {
                const int i = tid + q * 256;
                if (i < WG_PB * (CIN / 4)) {
                    const int r = i / (CIN / 4), c4 = i - r * (CIN / 4);
                    const bool v = ix[q] >= 0;
                    wg_cp_async16(&xs[st][r][c4 * 4], v ? X + (long long)ix[q] * CIN + c4 * 4 : X, v);
                }
            }
#pragma unroll
            for (int q = 0; q < NDr; ++q) {
                const int i = tid + q * 256;
                if (i < WG_PB * (COUT / 4)) {
                    const int r = i / (COUT / 4), c4 = i - r * (COUT / 4);
                    const bool v = id[q] >= 0;
                    wg_cp_async16(&ds[st][r][c4 * 4], v ? dY + (long long)id[q] * COUT + c4 * 4 : dY, v);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");               // one group per batch slot, empty or not
    };
    for (int bi = 0; bi < WG_STAGES - 1; ++bi) { load_idx(bi); issue(bi); }
    load_idx(WG_STAGES - 1);
    for (int bi = 0; bi < nb; ++bi) {
        issue(bi + WG_STAGES - 1);
        load_idx(bi + WG_STAGES);                                            // consumed by the NEXT iteration's issue
        asm volatile("cp.async.wait_group %0;" ::"n"(WG_STAGES - 1) : "memory");   // batch bi has landed
        __syncthreads();
        const int st = bi % WG_STAGES;
#pragma unroll 4
        for (int p = 0; p < WG_PB; ++p) {
            float a[MI], b[MJ];
#pragma unroll
            for (int i = 0; i < MI; i += VI) wg_load<VI>(&xs[st][p][(i / VI) * 16 * VI + ty * VI], &a[i]);
#pragma unroll
            for (int j = 0; j < MJ; j += VJ) wg_load<VJ>(&ds[st][p][(j / VJ) * 16 * VJ + tx * VJ], &b[j]);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < MJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();                                                     // stage st is refilled by the next issue
    }
    float* Wk = dW + (long long)k * CIN * COUT;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < MJ; ++j)
            atomicAdd(&Wk[((i / VI) * 16 * VI + ty * VI + i % VI) * COUT + (j / VJ) * 16 * VJ + tx * VJ + j % VJ], acc[i][j]);
}

template <int CIN, int COUT>
static int wgrad_launch(dim3 grid, cudaStream_t st, const float* x, const float* dy, const int* in_idx, const int* out_idx,
                        const int* count, long long seg_cap, float* dW) {
    constexpr int smem = WG_STAGES * WG_PB * (CIN + COUT) * 4;
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad<CIN, COUT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    k_wgrad<CIN, COUT><<<grid, 256, smem, st>>>(x, dy, in_idx, out_idx, count, seg_cap, dW);
    return IR_OK;
}

// small-Cin form (the stem, Cin = 7): one thread per (ci, co), pairs streamed from L2
__global__ void __launch_bounds__(256)
k_wgrad_small(const float* __restrict__ X, const float* __restrict__ dY, const int* __restrict__ in_idx,
              const int* __restrict__ out_idx, const int* __restrict__ count, long long seg_cap,
              int cin, int cout, float* __restrict__ dW) {
    const int k = blockIdx.y;
    const int cnt = count[k];
    const int chunk = (cnt + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * chunk, p_end = min(cnt, p_begin + chunk);
    const int tid = threadIdx.x;
    if (p_begin >= p_end || tid >= cin * cout) return;
    const int ci = tid / cout, co = tid - ci * cout;
    const long long base = (long long)k * seg_cap;
    float acc = 0.f;
    int p = p_begin;
    for (; p + 4 <= p_end; p += 4) {                 // four independent gather chains in flight
        int i[4], o[4];
        float xv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { i[u] = in_idx[base + p + u]; o[u] = out_idx[base + p + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { xv[u] = __ldg(X + (long long)i[u] * cin + ci); dv[u] = __ldg(dY + (long long)o[u] * cout + co); }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = fmaf(xv[u], dv[u], acc);
    }
    for (; p < p_end; ++p)
        acc = fmaf(__ldg(X + (long long)in_idx[base + p] * cin + ci), __ldg(dY + (long long)out_idx[base + p] * cout + co), acc);
    atomicAdd(&dW[((long long)k * cin + ci) * cout + co], acc);
}

extern "C" int ir_spconv_wgrad(const float* x, int32_t cin, const float* dy, int32_t cout, int32_t K,
                               const int32_t* in_idx, const int32_t* out_idx, const int32_t* count,
                               int64_t seg_cap, float* dW, ir_stream_t stream) {
    IR_CHECK_ARG(x && dy && in_idx && out_idx && count && dW && K > 0 && K <= 32 && seg_cap > 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)K * cin * cout * 4, st));
    // <= 4 CTAs per SM in total and never one more than a whole number of 2-CTA/SM waves (the 128x128 tile is
    // register-limited to two CTAs per SM: 594 CTAs would leave a 2-CTA third wave)
    const int nsplit = (4 * IR_NUM_SMS) / K > 0 ? (4 * IR_NUM_SMS) / K : 1;
    const dim3 grid(nsplit, K);
    int r = IR_OK;
    if (cin == 128 && cout == 128) r = wgrad_launch<128, 128>(grid, st, x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin == 64 && cout == 128) r = wgrad_launch<64, 128>(grid, st, x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin == 64 && cout == 64) r = wgrad_launch<64, 64>(grid, st, x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin == 32 && cout == 64) r = wgrad_launch<32, 64>(grid, st, x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin * cout <= 256) k_wgrad_small<<<dim3(ir_div_up(16 * IR_NUM_SMS, K), K), 256, 0, st>>>(x, dy, in_idx, out_idx, count, seg_cap, cin, cout, dW);
    else { ir_set_error("spconv_wgrad: unsupported channels %d -> %d", cin, cout); return IR_ERR_UNSUPPORTED; }
    if (r != IR_OK) return r;
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// wgrad with the tcgen05 kernel where the shape allows (Cin, Cout in {64,128}) and a range hint for dY is
// available; otherwise the SIMT kernel above.
extern "C" int ir_spconv_wgrad_scaled(const float* x, int32_t cin, const float* dy, const float* dy_absmax, int32_t cout,
                                      int32_t K, const int32_t* in_idx, const int32_t* out_idx, const int32_t* count,
                                      int64_t seg_cap, int32_t use_tc, float* dW, ir_stream_t stream) {
    const bool tc = use_tc && dy_absmax && (cin == 64 || cin == 128) && (cout == 64 || cout == 128) && !(cin == 128 && cout == 64);
    if (!tc) return ir_spconv_wgrad(x, cin, dy, cout, K, in_idx, out_idx, count, seg_cap, dW, stream);
    IR_CHECK_ARG(x && dy && in_idx && out_idx && count && dW && K > 0 && K <= 27 && seg_cap > 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)K * cin * cout * 4, st));
    return irk_wgrad_tc(x, cin, dy, cout, K, in_idx, out_idx, count, seg_cap, dy_absmax, dW, st);
}

// ------------------------------------------------------------------ BatchNorm, train mode, (rows, C)
// Used for spnn.BatchNorm over voxels, nn.BatchNorm1d over samples and nn.BatchNorm2d over NHWC cells
// (rows = B*H*W).  Two-level deterministic reductions: every CTA writes fp32 partial sums of its row
// stripe to scratch[part][2*C]; the consuming kernel adds the <= BN_MAX_PARTS partials in fp64 in its prologue.
// scratch: float[BN_MAX_PARTS * 2 * C] (ir_bn_scratch_floats).
#define BN_MAX_PARTS 128
#define BN_ROWS_PER_CTA 32
// few, fat partial CTAs: every CTA of the consuming kernel re-adds all partials of all channels in its prologue
static inline int bn_parts(long long n, int C) { return ir_min_i(ir_div_up(n > 0 ? n : 1, BN_ROWS_PER_CTA), C >= 128 ? 64 : BN_MAX_PARTS); }
extern "C" int64_t ir_bn_scratch_floats(int32_t C) { return (int64_t)BN_MAX_PARTS * 2 * C; }

// mode 0: (sum x, sum x^2);  mode 1 (backward): g = dy*[y>0]; (sum g, sum g*xhat).  A thread owns four adjacent
// channels (16-byte loads); C/4 threads cover a row, the 256/(C/4) thread groups of a CTA stride over its rows.
template <int MODE>
__global__ void __launch_bounds__(256)
k_bn_partials(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
              const int* __restrict__ n_dev, int n_host, int C, const float* __restrict__ mean,
              const float* __restrict__ rstd, int relu, float* __restrict__ scratch, float* __restrict__ absmax_out) {
    if (absmax_out && blockIdx.x == 0 && threadIdx.x == 0) *absmax_out = 0.f;    // k_bn_bwd_apply (next launch) maxes into it
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const int tid = threadIdx.x, C4 = C >> 2;
    const int c4 = tid % C4, g = tid / C4, G = blockDim.x / C4;
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {1.f, 1.f, 1.f, 1.f};
    if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { mu[u] = mean[c4 * 4 + u]; rs[u] = rstd[c4 * 4 + u]; }
    }
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    const long long step = (long long)gridDim.x * G;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * G + g; r < n; r += step) {
        const float4 q = reinterpret_cast<const float4*>(x)[r * C4 + c4];
        const float xv[4] = {q.x, q.y, q.z, q.w};
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { s0[u] += xv[u]; s1[u] = fmaf(xv[u], xv[u], s1[u]); }
        } else {
            const float4 d = reinterpret_cast<const float4*>(dy)[r * C4 + c4];
            float gv[4] = {d.x, d.y, d.z, d.w};
            if (relu) {
                const float4 yy = reinterpret_cast<const float4*>(y)[r * C4 + c4];
                if (!(yy.x > 0.f)) gv[0] = 0.f;
                if (!(yy.y > 0.f)) gv[1] = 0.f;
                if (!(yy.z > 0.f)) gv[2] = 0.f;
                if (!(yy.w > 0.f)) gv[3] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s0[u] += gv[u]; s1[u] = fmaf(gv[u], (xv[u] - mu[u]) * rs[u], s1[u]); }
        }
    }
    __shared__ float sh[2][4][256];
#pragma unroll
    for (int u = 0; u < 4; ++u) { sh[0][u][tid] = s0[u]; sh[1][u][tid] = s1[u]; }
    __syncthreads();
    if (g == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float a = s0[u], b = s1[u];
            for (int q = 1; q < G; ++q) { a += sh[0][u][q * C4 + c4]; b += sh[1][u][q * C4 + c4]; }
            scratch[(long long)blockIdx.x * 2 * C + c4 * 4 + u] = a;
            scratch[(long long)blockIdx.x * 2 * C + C + c4 * 4 + u] = b;
        }
    }
}

// Second level of the reduction, done by EVERY CTA of the consuming (apply) kernel in its prologue instead of by a
// launch of its own: the <= 64 partials of each channel are added in fp64 in one fixed order (thread group h takes
// partials h, h+H, ...; the groups are combined in order), so all CTAs hold bit-identical sums.  256 threads, C | 256.
__device__ __forceinline__ void bn_sum_parts(const float* __restrict__ scratch, int parts, int C, double& s, double& ss) {
    __shared__ double sh[2][4][256];
    const int tid = threadIdx.x, C4 = C >> 2, c4 = tid % C4, h = tid / C4, H = 256 / C4;      // 16-byte loads, H >= 4 groups
    double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int p = h; p < parts; p += H) {
        const float4 qa = reinterpret_cast<const float4*>(scratch + (long long)p * 2 * C)[c4];
        const float4 qb = reinterpret_cast<const float4*>(scratch + (long long)p * 2 * C + C)[c4];
        a[0] += (double)qa.x; a[1] += (double)qa.y; a[2] += (double)qa.z; a[3] += (double)qa.w;
        b[0] += (double)qb.x; b[1] += (double)qb.y; b[2] += (double)qb.z; b[3] += (double)qb.w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { sh[0][u][tid] = a[u]; sh[1][u][tid] = b[u]; }
    __syncthreads();
    s = 0.0; ss = 0.0;
    if (tid < C) {
        const int u = tid & 3, cc = tid >> 2;                                               // channel tid = 4*cc + u
        for (int q = 0; q < H; ++q) { s += sh[0][u][q * C4 + cc]; ss += sh[1][u][q * C4 + cc]; }
    }
}

// y = relu?((x - mean) * rstd * gamma + beta + resid); prologue: batch statistics from the partials (CTA 0 also
// publishes mean / rstd for the backward and updates the running statistics)
__global__ void __launch_bounds__(256)
k_bn_apply(const float* __restrict__ x, const float* __restrict__ scratch, int parts, const int* __restrict__ n_dev, int n_host,
           int C, float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
           float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ gamma,
           const float* __restrict__ beta, const float* __restrict__ resid, int relu, float* __restrict__ y) {
    __shared__ float s_mean[256], s_rstd[256];
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    double s, ss;
    bn_sum_parts(scratch, parts, C, s, ss);
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        const double inv = n > 0 ? 1.0 / n : 0.0;
        const double m = s * inv;
        double var = ss * inv - m * m;
        if (var < 0) var = 0;
        const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)eps));
        s_mean[c] = mf; s_rstd[c] = rf;
        if (blockIdx.x == 0) {
            mean_out[c] = mf;
            rstd_out[c] = rf;
            if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mf;
            if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(n > 1 ? var * n / (n - 1) : var);
        }
    }
    __syncthreads();
    const long long total4 = (long long)n * C / 4;                 // C is a multiple of 4
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* r4 = reinterpret_cast<const float4*>(resid);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % C);
        const float4 xv = x4[i];
        float v[4] = {xv.x, xv.y, xv.z, xv.w};
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (resid) { const float4 q = r4[i]; r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float o = (v[u] - s_mean[c + u]) * s_rstd[c + u] * gamma[c + u] + beta[c + u] + r[u];
            v[u] = relu ? fmaxf(o, 0.f) : o;
        }
        y4[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

static inline int bn_apply_grid(long long n, int C) { return ir_min_i(ir_div_up(n * C, 1024), IR_NUM_SMS * 2); }

extern "C" int ir_bn_train_fwd(const float* x, const int32_t* n_dev, int32_t n, int32_t C, const float* gamma,
                               const float* beta, const float* resid, int32_t relu, float eps, float momentum,
                               float* running_mean, float* running_var, float* scratch, float* mean,
                               float* rstd, float* y, ir_stream_t stream) {
    IR_CHECK_ARG(x && gamma && beta && scratch && mean && rstd && y && n > 0 && C >= 4 && C <= 256 && 256 % C == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int parts = bn_parts(n, C);
    k_bn_partials<0><<<parts, 256, 0, st>>>(x, nullptr, nullptr, n_dev, n, C, nullptr, nullptr, 0, scratch, nullptr);
    IR_CHECK_LAUNCH();
    k_bn_apply<<<bn_apply_grid(n, C), 256, 0, st>>>(x, scratch, parts, n_dev, n, C, eps, momentum, running_mean, running_var, mean, rstd,
                                                    gamma, beta, resid, relu, y);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// backward: g = dy * [y > 0] (if relu); dbeta = sum g; dgamma = sum g*xhat;
//           dx = gamma*rstd*(g - dbeta/n - xhat*dgamma/n); dresid = g
// prologue: dgamma / dbeta from the partials (published by CTA 0)
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
               const float* __restrict__ scratch, int parts, const int* __restrict__ n_dev, int n_host, int C,
               const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma, int relu,
               float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx,
               float* __restrict__ dresid, float* __restrict__ absmax_out) {
    __shared__ float s_dg[256], s_db[256];
    double s, sx;
    bn_sum_parts(scratch, parts, C, s, sx);
    if (threadIdx.x < C) {
        s_db[threadIdx.x] = (float)s; s_dg[threadIdx.x] = (float)sx;
        if (blockIdx.x == 0) { dbeta[threadIdx.x] = (float)s; dgamma[threadIdx.x] = (float)sx; }
    }
    __syncthreads();
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const float inv = n > 0 ? 1.f / n : 0.f;
    const long long total4 = (long long)n * C / 4;
    float amax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % C);
        const float4 d4 = reinterpret_cast<const float4*>(dy)[i];
        const float4 x4 = reinterpret_cast<const float4*>(x)[i];
        float g[4] = {d4.x, d4.y, d4.z, d4.w};
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        if (relu) {
            const float4 y4 = reinterpret_cast<const float4*>(y)[i];
            if (!(y4.x > 0.f)) g[0] = 0.f;
            if (!(y4.y > 0.f)) g[1] = 0.f;
            if (!(y4.z > 0.f)) g[2] = 0.f;
            if (!(y4.w > 0.f)) g[3] = 0.f;
        }
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xv[u] - mean[c + u]) * rstd[c + u];
            o[u] = gamma[c + u] * rstd[c + u] * (g[u] - s_db[c + u] * inv - xh * s_dg[c + u] * inv);
            amax = fmaxf(amax, fabsf(o[u]));
        }
        reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (dresid) reinterpret_cast<float4*>(dresid)[i] = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (absmax_out) {                      // max |dx| (non-negative floats order like their bit patterns)
        amax = warp_max(amax);
        if ((threadIdx.x & 31) == 0 && amax > 0.f && amax < 3.0e38f) atomicMax(reinterpret_cast<unsigned*>(absmax_out), __float_as_uint(amax));
    }
}

extern "C" int ir_bn_train_bwd(const float* dy, const float* y, const float* x, const int32_t* n_dev, int32_t n,
                               int32_t C, const float* mean, const float* rstd, const float* gamma, int32_t relu,
                               float* scratch, float* dx, float* dresid, float* dgamma, float* dbeta,
                               float* absmax_out, ir_stream_t stream) {
    IR_CHECK_ARG(dy && x && mean && rstd && gamma && scratch && dx && dgamma && dbeta && n > 0 && C >= 4 && C <= 256 && 256 % C == 0);
    IR_CHECK_ARG(!relu || y);
    cudaStream_t st = (cudaStream_t)stream;
    const int parts = bn_parts(n, C);
    k_bn_partials<1><<<parts, 256, 0, st>>>(x, dy, y, n_dev, n, C, mean, rstd, relu, scratch, absmax_out);
    IR_CHECK_LAUNCH();
    k_bn_bwd_apply<<<bn_apply_grid(n, C), 256, 0, st>>>(dy, y, x, scratch, parts, n_dev, n, C, mean, rstd, gamma, relu, dgamma, dbeta,
                                                        dx, dresid, absmax_out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ GlobalMaxPooling backward
// The gradient of out[b,c] = max_{rows of b} F[row,c] goes to the FIRST row attaining the maximum.
__global__ void k_segmax_argmin(const float* __restrict__ F, const int4* __restrict__ coords, const int* __restrict__ n_dev,
                                int C, int n_seg, const float* __restrict__ pooled, int* __restrict__ arg) {
    const long long total = (long long)(*n_dev) * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / C), c = (int)(i - (long long)row * C);
        const int b = coords[row].w;
        if (b >= 0 && b < n_seg && F[i] == pooled[(long long)b * C + c]) atomicMin(&arg[(long long)b * C + c], row);
    }
}
__global__ void k_segmax_bwd(const int4* __restrict__ coords, const int* __restrict__ n_dev, int C, int n_seg,
                             const float* __restrict__ dpooled, const int* __restrict__ arg, float* __restrict__ dF) {
    const long long total = (long long)(*n_dev) * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / C), c = (int)(i - (long long)row * C);
        const int b = coords[row].w;
        dF[i] = (b >= 0 && b < n_seg && arg[(long long)b * C + c] == row) ? dpooled[(long long)b * C + c] : 0.f;
    }
}

extern "C" int ir_segmax_bwd(const float* feats, const int32_t* coords, const int32_t* n_dev, int64_t n_max,
                             int32_t C, int32_t n_seg, const float* pooled, const float* dpooled,
                             int32_t* arg_scratch, float* dfeats, ir_stream_t stream) {
    IR_CHECK_ARG(feats && coords && n_dev && pooled && dpooled && arg_scratch && dfeats && C > 0 && n_seg > 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(arg_scratch, 0x7F, (size_t)n_seg * C * 4, st));
    const int grid = ir_min_i(ir_div_up(n_max * C > 0 ? n_max * C : 1, 256), IR_NUM_SMS * 8);
    k_segmax_argmin<<<grid, 256, 0, st>>>(feats, (const int4*)coords, n_dev, C, n_seg, pooled, arg_scratch);
    IR_CHECK_LAUNCH();
    k_segmax_bwd<<<grid, 256, 0, st>>>((const int4*)coords, n_dev, C, n_seg, dpooled, arg_scratch, dfeats);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
