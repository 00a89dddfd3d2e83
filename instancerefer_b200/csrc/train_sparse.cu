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
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
k_wgrad(const float* __restrict__ X, const float* __restrict__ dY, const int* __restrict__ in_idx,
        const int* __restrict__ out_idx, const int* __restrict__ count, long long seg_cap,
        float* __restrict__ dW) {
    constexpr int MI = CIN / 16, MJ = COUT / 16;
    __shared__ __align__(16) float xs[WG_PB][CIN];
    __shared__ __align__(16) float ds[WG_PB][COUT];
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
    for (int p0 = p_begin; p0 < p_end; p0 += WG_PB) {
        const int np = min(WG_PB, p_end - p0);
        // stage gathered rows with 16-byte loads (CIN, COUT are multiples of 16)
        for (int i = tid; i < WG_PB * (CIN / 4); i += 256) {
            const int r = i / (CIN / 4), c4 = i - r * (CIN / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < np) v = __ldg(reinterpret_cast<const float4*>(X + (long long)in_idx[base + p0 + r] * CIN) + c4);
            reinterpret_cast<float4*>(&xs[r][0])[c4] = v;
        }
        for (int i = tid; i < WG_PB * (COUT / 4); i += 256) {
            const int r = i / (COUT / 4), c4 = i - r * (COUT / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < np) v = __ldg(reinterpret_cast<const float4*>(dY + (long long)out_idx[base + p0 + r] * COUT) + c4);
            reinterpret_cast<float4*>(&ds[r][0])[c4] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < WG_PB; ++p) {
            float a[MI], b[MJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = xs[p][i * 16 + ty];
#pragma unroll
            for (int j = 0; j < MJ; ++j) b[j] = ds[p][j * 16 + tx];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < MJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Wk = dW + (long long)k * CIN * COUT;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < MJ; ++j) atomicAdd(&Wk[(i * 16 + ty) * COUT + j * 16 + tx], acc[i][j]);
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
    for (int p = p_begin; p < p_end; ++p)
        acc = fmaf(__ldg(X + (long long)in_idx[base + p] * cin + ci), __ldg(dY + (long long)out_idx[base + p] * cout + co), acc);
    atomicAdd(&dW[((long long)k * cin + ci) * cout + co], acc);
}

extern "C" int ir_spconv_wgrad(const float* x, int32_t cin, const float* dy, int32_t cout, int32_t K,
                               const int32_t* in_idx, const int32_t* out_idx, const int32_t* count,
                               int64_t seg_cap, float* dW, ir_stream_t stream) {
    IR_CHECK_ARG(x && dy && in_idx && out_idx && count && dW && K > 0 && K <= 32 && seg_cap > 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)K * cin * cout * 4, st));
    const int nsplit = ir_div_up(4 * IR_NUM_SMS, K);
    const dim3 grid(nsplit, K);
    if (cin == 128 && cout == 128) k_wgrad<128, 128><<<grid, 256, 0, st>>>(x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin == 64 && cout == 128) k_wgrad<64, 128><<<grid, 256, 0, st>>>(x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin == 64 && cout == 64) k_wgrad<64, 64><<<grid, 256, 0, st>>>(x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin == 32 && cout == 64) k_wgrad<32, 64><<<grid, 256, 0, st>>>(x, dy, in_idx, out_idx, count, seg_cap, dW);
    else if (cin * cout <= 256) k_wgrad_small<<<grid, 256, 0, st>>>(x, dy, in_idx, out_idx, count, seg_cap, cin, cout, dW);
    else { ir_set_error("spconv_wgrad: unsupported channels %d -> %d", cin, cout); return IR_ERR_UNSUPPORTED; }
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ BatchNorm, train mode, (rows, C)
// Used for spnn.BatchNorm over voxels, nn.BatchNorm1d over samples and nn.BatchNorm2d over NHWC cells
// (rows = B*H*W).  Statistics are accumulated per CTA in fp32 over <= ~100 rows, then in fp64 atomics.
// scratch: double[2*C].
__global__ void __launch_bounds__(256)
k_bn_stats(const float* __restrict__ x, const int* __restrict__ n_dev, int n_host, int C,
           double* __restrict__ scratch) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const int tid = threadIdx.x;
    const int c = tid % C, g = tid / C, G = blockDim.x / C;
    float s = 0.f, ss = 0.f;
    for (long long r = (long long)blockIdx.x * G + g; r < n; r += (long long)gridDim.x * G) {
        const float v = x[r * C + c];
        s += v;
        ss = fmaf(v, v, ss);
    }
    __shared__ float sh[2][256];
    sh[0][tid] = s; sh[1][tid] = ss;
    __syncthreads();
    if (g == 0) {
        for (int q = 1; q < G; ++q) { s += sh[0][q * C + c]; ss += sh[1][q * C + c]; }
        atomicAdd(&scratch[c], (double)s);
        atomicAdd(&scratch[C + c], (double)ss);
    }
}

__global__ void k_bn_finalize(const double* __restrict__ scratch, const int* __restrict__ n_dev, int n_host,
                              int C, float eps, float momentum, float* __restrict__ running_mean,
                              float* __restrict__ running_var, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const double inv = n > 0 ? 1.0 / n : 0.0;
    const double m = scratch[c] * inv;
    double var = scratch[C + c] * inv - m * m;
    if (var < 0) var = 0;
    mean_out[c] = (float)m;
    rstd_out[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(n > 1 ? var * n / (n - 1) : var);
}

__global__ void __launch_bounds__(256)
k_bn_apply(const float* __restrict__ x, const int* __restrict__ n_dev, int n_host, int C,
           const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
           const float* __restrict__ beta, const float* __restrict__ resid, int relu, float* __restrict__ y) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const long long total = (long long)n * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = (x[i] - mean[c]) * rstd[c] * gamma[c] + beta[c];
        if (resid) v += resid[i];
        if (relu) v = fmaxf(v, 0.f);
        y[i] = v;
    }
}

extern "C" int ir_bn_train_fwd(const float* x, const int32_t* n_dev, int32_t n, int32_t C, const float* gamma,
                               const float* beta, const float* resid, int32_t relu, float eps, float momentum,
                               float* running_mean, float* running_var, double* scratch, float* mean,
                               float* rstd, float* y, ir_stream_t stream) {
    IR_CHECK_ARG(x && gamma && beta && scratch && mean && rstd && y && n > 0 && C > 0 && C <= 256 && 256 % C == 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)2 * C * sizeof(double), st));
    const int G = 256 / C;
    const int grid = ir_min_i(ir_div_up(n, (long long)G * 64), IR_NUM_SMS * 4);
    k_bn_stats<<<grid, 256, 0, st>>>(x, n_dev, n, C, scratch);
    IR_CHECK_LAUNCH();
    k_bn_finalize<<<ir_div_up(C, 128), 128, 0, st>>>(scratch, n_dev, n, C, eps, momentum, running_mean, running_var, mean, rstd);
    IR_CHECK_LAUNCH();
    k_bn_apply<<<ir_min_i(ir_div_up((long long)n * C, 1024), IR_NUM_SMS * 8), 256, 0, st>>>(x, n_dev, n, C, mean, rstd, gamma, beta, resid, relu, y);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// backward: g = dy * [y > 0] (if relu); dbeta = sum g; dgamma = sum g*xhat;
//           dx = gamma*rstd*(g - dbeta/n - xhat*dgamma/n); dresid = g
__global__ void __launch_bounds__(256)
k_bn_bwd_reduce(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                const int* __restrict__ n_dev, int n_host, int C, const float* __restrict__ mean,
                const float* __restrict__ rstd, int relu, double* __restrict__ scratch) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const int tid = threadIdx.x;
    const int c = tid % C, g = tid / C, G = blockDim.x / C;
    const float mu = mean[c], rs = rstd[c];
    float s = 0.f, sx = 0.f;
    for (long long r = (long long)blockIdx.x * G + g; r < n; r += (long long)gridDim.x * G) {
        float gv = dy[r * C + c];
        if (relu && !(y[r * C + c] > 0.f)) gv = 0.f;
        s += gv;
        sx = fmaf(gv, (x[r * C + c] - mu) * rs, sx);
    }
    __shared__ float sh[2][256];
    sh[0][tid] = s; sh[1][tid] = sx;
    __syncthreads();
    if (g == 0) {
        for (int q = 1; q < G; ++q) { s += sh[0][q * C + c]; sx += sh[1][q * C + c]; }
        atomicAdd(&scratch[c], (double)s);
        atomicAdd(&scratch[C + c], (double)sx);
    }
}

__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
               const int* __restrict__ n_dev, int n_host, int C, const float* __restrict__ mean,
               const float* __restrict__ rstd, const float* __restrict__ gamma, int relu,
               const double* __restrict__ scratch, float* __restrict__ dx, float* __restrict__ dresid,
               float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const float inv = n > 0 ? 1.f / n : 0.f;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] = (float)scratch[c];
            dgamma[c] = (float)scratch[C + c];
        }
    const long long total = (long long)n * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float gv = dy[i];
        if (relu && !(y[i] > 0.f)) gv = 0.f;
        const float xh = (x[i] - mean[c]) * rstd[c];
        dx[i] = gamma[c] * rstd[c] * (gv - (float)scratch[c] * inv - xh * (float)scratch[C + c] * inv);
        if (dresid) dresid[i] = gv;
    }
}

extern "C" int ir_bn_train_bwd(const float* dy, const float* y, const float* x, const int32_t* n_dev, int32_t n,
                               int32_t C, const float* mean, const float* rstd, const float* gamma, int32_t relu,
                               double* scratch, float* dx, float* dresid, float* dgamma, float* dbeta,
                               ir_stream_t stream) {
    IR_CHECK_ARG(dy && x && mean && rstd && gamma && scratch && dx && dgamma && dbeta && n > 0 && C > 0 && C <= 256 && 256 % C == 0);
    IR_CHECK_ARG(!relu || y);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)2 * C * sizeof(double), st));
    const int G = 256 / C;
    const int grid = ir_min_i(ir_div_up(n, (long long)G * 64), IR_NUM_SMS * 4);
    k_bn_bwd_reduce<<<grid, 256, 0, st>>>(dy, y, x, n_dev, n, C, mean, rstd, relu, scratch);
    IR_CHECK_LAUNCH();
    k_bn_bwd_apply<<<ir_min_i(ir_div_up((long long)n * C, 1024), IR_NUM_SMS * 8), 256, 0, st>>>(
        dy, y, x, n_dev, n, C, mean, rstd, gamma, relu, scratch, dx, dresid, dgamma, dbeta);
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
