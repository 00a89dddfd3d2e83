// Loss and optimiser kernels of the training step:
//   get_loss (lib/loss_helper.py:131-161,189-269): language / region cross-entropy, axis-aligned box
//   IoU -> one-hot cluster label -> ContrastiveLoss, each fused with its own gradient;
//   torch.optim.Adam over ONE flat parameter buffer (scripts/train.py:93, lib/solver.py:200-205).
#include <math.h>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"

// ------------------------------------------------------------------ cross entropy (mean) + gradient
// single CTA: warp per row; the per-row losses are summed in row order, tile by tile of CE_TILE rows, by one thread
// (deterministic for any B: the reference's CrossEntropyLoss has no batch limit either)
#define CE_TILE 1024
__global__ void __launch_bounds__(256)
k_cross_entropy(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int N,
                float* __restrict__ loss, float* __restrict__ dlogits) {
    __shared__ float s_loss[CE_TILE];
    __shared__ float s_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float invB = 1.f / B;
    if (threadIdx.x == 0) s_total = 0.f;
    for (int r0 = 0; r0 < B; r0 += CE_TILE) {
        const int nr = min(CE_TILE, B - r0);
        for (int rr = w; rr < nr; rr += 8) {
            const int r = r0 + rr;
            const float* z = logits + (long long)r * N;
            float m = -INFINITY;
            for (int j = lane; j < N; j += 32) m = fmaxf(m, z[j]);
            m = warp_max(m);
            float s = 0.f;
            for (int j = lane; j < N; j += 32) s += expf(z[j] - m);
            s = warp_sum(s);
            const float lse = m + logf(s);
            const int y = (int)labels[r];
            for (int j = lane; j < N; j += 32)
                dlogits[(long long)r * N + j] = (expf(z[j] - lse) - (j == y ? 1.f : 0.f)) * invB;
            if (lane == 0) s_loss[rr] = lse - z[y];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = s_total;
            for (int rr = 0; rr < nr; ++rr) t += s_loss[rr];
            s_total = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = s_total * invB;
}

extern "C" int ir_cross_entropy(const float* logits, const int64_t* labels, int32_t B, int32_t N,
                                float* loss, float* dlogits, ir_stream_t stream) {
    IR_CHECK_ARG(logits && labels && loss && dlogits && B > 0 && N > 0);
    k_cross_entropy<<<1, 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)labels, B, N, loss, dlogits);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ 9-way region label
// compute_scene_mask_loss (lib/loss_helper.py:131-153): thirds of the scene's xy extent.
__global__ void k_region_label(const double* __restrict__ center, const double* __restrict__ pmin,
                               const double* __restrict__ pmax, int B, int f32, long long* __restrict__ label) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    bool f[2], s[2];
    for (int a = 0; a < 2; ++a) {
        const double lo = pmin[b * 3 + a], hi = pmax[b * 3 + a], c = center[b * 3 + a];
        if (f32) {           // all operands were fp32 in the caller: keep fp32 rounding of the thirds
            const float lof = (float)lo, hif = (float)hi;
            f[a] = (float)c <= lof + (hif - lof) / 3.f;
            s[a] = (float)c <= lof + (hif - lof) / 3.f * 2.f;
        } else {
            f[a] = c <= lo + (hi - lo) / 3;
            s[a] = c <= lo + (hi - lo) / 3 * 2;
        }
    }
    int l = (f[0] && f[1]) ? 0 : 4;
    if (!f[0] && s[0] && f[1]) l = 1;
    if (!s[0] && f[1]) l = 2;
    if (f[0] && !f[1] && s[0]) l = 3;       // (the reference's condition, lib/loss_helper.py:148)
    if (!s[0] && !f[1] && s[1]) l = 5;
    if (f[0] && !s[1]) l = 6;
    if (!f[0] && s[0] && !s[1]) l = 7;
    if (!s[0] && !s[1]) l = 8;
    label[b] = l;
}

extern "C" int ir_region_label(const double* ref_center, const double* point_min, const double* point_max,
                               int32_t B, int32_t inputs_were_f32, int64_t* label, ir_stream_t stream) {
    IR_CHECK_ARG(ref_center && point_min && point_max && label && B > 0);
    k_region_label<<<ir_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(ref_center, point_min, point_max, B,
                                                                        inputs_were_f32, (long long*)label);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ reference loss
// One warp per scene.  Boxes: obb = [cx,cy,cz, l,w,h, heading]; corners (+-l/2,+-w/2,+-h/2) rotated by
// roty(heading)^T and translated (utils/box_util.py:310-333), min/max per axis, axis-aligned IoU
// (:154-179) in fp64 like the reference's numpy.  label = one-hot of the FIRST maximal IoU.
// For scenes with >= 2 candidates and max IoU >= 0.2: ContrastiveLoss(margin .2, gamma 5) on
// score = attr + rel + scene, and its gradient w.r.t. every score entry.
__device__ void box_min_max(const double* o, double mn[3], double mx[3]) {
    const double c = cos(o[6]), s = sin(o[6]);
    mn[0] = mn[1] = mn[2] = 1e300;
    mx[0] = mx[1] = mx[2] = -1e300;
    for (int q = 0; q < 8; ++q) {
        const double x = ((q & 3) < 2 ? 0.5 : -0.5) * o[3];                 // l/2, l/2,-l/2,-l/2, ...
        const double y = ((q & 3) == 0 || (q & 3) == 3 ? 0.5 : -0.5) * o[4]; // w/2,-w/2,-w/2, w/2, ...
        const double z = (q < 4 ? 0.5 : -0.5) * o[5];
        const double p[3] = {c * x + s * z + o[0], y + o[1], -s * x + c * z + o[2]};
        for (int a = 0; a < 3; ++a) { mn[a] = fmin(mn[a], p[a]); mx[a] = fmax(mx[a], p[a]); }
    }
}

__global__ void __launch_bounds__(32)
k_ref_loss(const double* __restrict__ pred_obb, const int* __restrict__ obb_ofs, const double* __restrict__ gt_obb,
           const int* __restrict__ score_ofs, const float* __restrict__ sa, const float* __restrict__ sr,
           const float* __restrict__ ss, float margin, float gamma, double iou_thresh,
           float* __restrict__ label, float* __restrict__ loss_scene, float* __restrict__ dscore,
           float* __restrict__ iou_max_out) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int o0 = obb_ofs[b], n = obb_ofs[b + 1] - o0;
    if (lane == 0) { loss_scene[b] = 0.f; iou_max_out[b] = 0.f; }
    if (n == 0) return;
    double gmn[3], gmx[3];
    box_min_max(gt_obb + (long long)b * 7, gmn, gmx);
    const double gvol = (gmx[0] - gmn[0]) * (gmx[1] - gmn[1]) * (gmx[2] - gmn[2]);
    double best = -1.0;
    int best_j = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
        double mn[3], mx[3];
        box_min_max(pred_obb + (long long)(o0 + j) * 7, mn, mx);
        double inter = 1.0;
        for (int a = 0; a < 3; ++a) inter *= fmax(fmin(mx[a], gmx[a]) - fmax(mn[a], gmn[a]), 0.0);
        const double vol = (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]);
        const double iou = inter / (vol + gvol - inter + 1e-8);
        if (iou > best) { best = iou; best_j = j; }          // strided scan keeps the lowest j per lane
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
    }
    for (int j = lane; j < n; j += 32) label[o0 + j] = (j == best_j) ? 1.f : 0.f;
    if (lane == 0) iou_max_out[b] = (float)best;
    const int s0 = score_ofs[b];
    if (s0 < 0 || n < 2) return;
    if (best < iou_thresh) {                 // fp64 threshold like the reference's `max_iou < 0.2` (lib/loss_helper.py:246)
        for (int j = lane; j < n; j += 32) dscore[s0 + j] = 0.f;
        return;
    }
    // t_j = gamma*score_j for negatives, 0 at the positive (score*label.logical_not()); lse over all j
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) {
        const float t = (j == best_j) ? 0.f : gamma * (sa[s0 + j] + sr[s0 + j] + ss[s0 + j]);
        m = fmaxf(m, t);
    }
    m = warp_max(m);
    float e = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float t = (j == best_j) ? 0.f : gamma * (sa[s0 + j] + sr[s0 + j] + ss[s0 + j]);
        e += expf(t - m);
    }
    e = warp_sum(e);
    const float lse = m + logf(e);
    const float sim = gamma * (sa[s0 + best_j] + sr[s0 + best_j] + ss[s0 + best_j]);
    const float l = lse - sim + margin;
    const bool active = l > 0.f;
    if (lane == 0) loss_scene[b] = active ? l : 0.f;
    for (int j = lane; j < n; j += 32) {
        float g = 0.f;
        if (active) {
            if (j == best_j) g = -gamma;
            else g = gamma * expf(gamma * (sa[s0 + j] + sr[s0 + j] + ss[s0 + j]) - lse);
        }
        dscore[s0 + j] = g;
    }
}

extern "C" int ir_ref_loss(const double* pred_obb, const int32_t* obb_ofs, const double* gt_obb,
                           const int32_t* score_ofs, int32_t B, const float* s_attr, const float* s_rel,
                           const float* s_scene, float margin, float gamma, double iou_thresh, float* label,
                           float* loss_scene, float* dscore, float* iou_max, ir_stream_t stream) {
    IR_CHECK_ARG(pred_obb && obb_ofs && gt_obb && score_ofs && s_attr && s_rel && s_scene && label &&
                 loss_scene && dscore && iou_max && B > 0);
    k_ref_loss<<<B, 32, 0, (cudaStream_t)stream>>>(pred_obb, obb_ofs, gt_obb, score_ofs, s_attr, s_rel, s_scene,
                                                   margin, gamma, iou_thresh, label, loss_scene, dscore, iou_max);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ Adam over a flat buffer
// g <- grad_scale*g + wd*p;  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;
// p <- p - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)          (torch.optim.Adam, amsgrad off)
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
       long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
       const unsigned char* __restrict__ block_skip, const float* __restrict__ hyper) {
    if (hyper) { lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2]; }     // per-step scalars of a graph-replayed step
    const float step = lr / bc1;
    // 16-byte accesses (a skip byte covers 16 of them), scalar tail for n % 4 elements
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        if (block_skip && block_skip[i >> 4]) continue;      // parameter without a gradient this step: torch.optim.Adam skips it
        const float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
        float pv[4] = {P.x, P.y, P.z, P.w}, mv[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
        const float gg[4] = {G.x, G.y, G.z, G.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float gv = fmaf(wd, pv[u], grad_scale * gg[u]);
            mv[u] = b1 * mv[u] + (1.f - b1) * gv;
            vv[u] = b2 * vv[u] + (1.f - b2) * gv * gv;
            pv[u] = pv[u] - step * mv[u] / (sqrtf(vv[u]) / bc2_sqrt + eps);
        }
        m4[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        p4[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    }
    if (blockIdx.x == 0) {
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            if (block_skip && block_skip[i >> 6]) continue;
            const float pv = p[i];
            const float gv = fmaf(wd, pv, grad_scale * g[i]);
            const float mv = b1 * m[i] + (1.f - b1) * gv;
            const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
            m[i] = mv;
            v[i] = vv;
            p[i] = pv - step * mv / (sqrtf(vv) / bc2_sqrt + eps);
        }
    }
}

extern "C" int ir_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                            float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                            float grad_scale, const uint8_t* block_skip, ir_stream_t stream) {
    IR_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1);
    IR_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                   reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0);
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    const int grid = ir_min_i(ir_div_up(n, 256 * 4), IR_NUM_SMS * 8);
    k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                   weight_decay, bc1, sqrtf(bc2), grad_scale, block_skip, nullptr);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// The three scalars that change from step to step (lr: a scheduler may move it; the two bias corrections) as they
// are needed by a launch replayed from a CUDA graph: ir_adam_hyper computes them on the host exactly like
// ir_adam_step, the caller copies them to the device before the replay, ir_adam_step_dev reads them there.
extern "C" int ir_adam_hyper(float lr, float beta1, float beta2, int32_t step, float* hyper3) {
    IR_CHECK_ARG(hyper3 && step >= 1);
    hyper3[0] = lr;
    hyper3[1] = (float)(1.0 - pow((double)beta1, (double)step));
    hyper3[2] = sqrtf((float)(1.0 - pow((double)beta2, (double)step)));
    return IR_OK;
}
extern "C" int ir_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const float* hyper_dev, float beta1, float beta2, float eps, float weight_decay,
                                float grad_scale, const uint8_t* block_skip, ir_stream_t stream) {
    IR_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && hyper_dev);
    IR_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                   reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0);
    const int grid = ir_min_i(ir_div_up(n, 256 * 4), IR_NUM_SMS * 8);
    k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, 0.f, beta1, beta2, eps,
                                                   weight_decay, 1.f, 1.f, grad_scale, block_skip, hyper_dev);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ get_eval (lib/eval_helper.py:11-114)
// One warp per scene: candidate with the highest summed score (first maximum), its box against the
// ground-truth box (axis-aligned IoU of the rotated corner boxes, utils/box_util.py:95-133,290-308),
// ref_acc (= argmax matches the IoU label when the scene has >= 2 candidates, IoU > 0.25 otherwise) and the
// un-rotated corner boxes of construct_bbox_corners (utils/util.py:21-32) for the visualisation lists.
__global__ void __launch_bounds__(32)
k_ref_eval(const double* __restrict__ pred_obb, const int* __restrict__ obb_ofs, const double* __restrict__ gt_obb,
           const int* __restrict__ score_ofs, const float* __restrict__ sa, const float* __restrict__ sr,
           const float* __restrict__ ss, const float* __restrict__ label, int* __restrict__ pred_idx,
           float* __restrict__ ref_acc, double* __restrict__ iou_out, double* __restrict__ pred_corners,
           double* __restrict__ gt_corners) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int o0 = obb_ofs[b], n = obb_ofs[b + 1] - o0;
    const int s0 = score_ofs[b];
    int best_j = 0, target = 0;
    if (n >= 2 && s0 >= 0) {
        float best = -INFINITY, bl = -INFINITY;
        best_j = 0x7fffffff; target = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            const float v = sa[s0 + j] + sr[s0 + j] + ss[s0 + j];
            if (v > best) { best = v; best_j = j; }
            const float l = label[o0 + j];
            if (l > bl) { bl = l; target = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
            if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
            const float ol = __shfl_xor_sync(0xffffffffu, bl, o);
            const int ot = __shfl_xor_sync(0xffffffffu, target, o);
            if (ol > bl || (ol == bl && ot < target)) { bl = ol; target = ot; }
        }
    }
    if (lane != 0) return;
    double po[7] = {0, 0, 0, 0, 0, 0, 0};                            // no candidate: a zero box (:55-57)
    if (n >= 1) for (int q = 0; q < 7; ++q) po[q] = pred_obb[(long long)(o0 + best_j) * 7 + q];
    const double* go = gt_obb + (long long)b * 7;
    double mn1[3], mx1[3], mn2[3], mx2[3];
    box_min_max(po, mn1, mx1);
    box_min_max(go, mn2, mx2);
    double inter = 1.0;
    for (int a = 0; a < 3; ++a) inter *= fmax(fmin(mx1[a], mx2[a]) - fmax(mn1[a], mn2[a]), 0.0);
    const double v1 = (mx1[0] - mn1[0]) * (mx1[1] - mn1[1]) * (mx1[2] - mn1[2]);
    const double v2 = (mx2[0] - mn2[0]) * (mx2[1] - mn2[1]) * (mx2[2] - mn2[2]);
    const double iou = inter / (v1 + v2 - inter + 1e-8);
    iou_out[b] = iou;
    pred_idx[b] = n >= 1 ? best_j : -1;
    ref_acc[b] = (n >= 2) ? (target == best_j ? 1.f : 0.f) : (iou > 0.25 ? 1.f : 0.f);
    for (int q = 0; q < 8; ++q) {
        const double sx = ((q & 3) < 2 ? 0.5 : -0.5), sy = ((q & 3) == 0 || (q & 3) == 3 ? 0.5 : -0.5), sz = (q < 4 ? 0.5 : -0.5);
        double* pc = pred_corners + ((long long)b * 8 + q) * 3;
        double* gc = gt_corners + ((long long)b * 8 + q) * 3;
        pc[0] = sx * po[3] + po[0]; pc[1] = sy * po[4] + po[1]; pc[2] = sz * po[5] + po[2];
        gc[0] = sx * go[3] + go[0]; gc[1] = sy * go[4] + go[1]; gc[2] = sz * go[5] + go[2];
    }
}

extern "C" int ir_ref_eval(const double* pred_obb, const int32_t* obb_ofs, const double* gt_obb,
                           const int32_t* score_ofs, int32_t B, const float* s_attr, const float* s_rel,
                           const float* s_scene, const float* label, int32_t* pred_idx, float* ref_acc,
                           double* iou, double* pred_corners, double* gt_corners, ir_stream_t stream) {
    IR_CHECK_ARG(pred_obb && obb_ofs && gt_obb && score_ofs && s_attr && s_rel && s_scene && label && pred_idx &&
                 ref_acc && iou && pred_corners && gt_corners && B > 0);
    k_ref_eval<<<B, 32, 0, (cudaStream_t)stream>>>(pred_obb, obb_ofs, gt_obb, score_ofs, s_attr, s_rel, s_scene, label,
                                                   pred_idx, ref_acc, iou, pred_corners, gt_corners);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
