// One-call training passes of a sparse encoder (SURVEY.md §8 rows a7 + a14): all 13
// [conv -> train-mode BN (-> + skip) -> ReLU] layers of SparseConvEncoder / BEVEncoder
// (models/basic_blocks.py:59-95,136-171) forward, and their backward (BN backward, wgrad, dgrad on the
// transposed rulebooks, skip-gradient folded into the dgrad epilogue), as ONE chain of asynchronous
// launches per direction: the host issues two library calls per encoder and training step instead of
// ~200 per-operator calls.  Activations live in a caller-owned arena (ir_encoder_train_layout).
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"

// A helper stream per calling stream, owned by the library: work that does not sit on the layer-to-layer dependency chain
// (transposed rulebooks and W^T during the forward, each layer's wgrad during the backward) is forked onto it and joined
// back with events, so it runs beside the chain — also inside a CUDA-graph capture, where fork/join become graph edges.
struct SideStream { cudaStream_t s; cudaEvent_t fork, join, ring[3]; };
static SideStream* side_stream_for(cudaStream_t main) {
    static std::unordered_map<cudaStream_t, SideStream> pool;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = pool.find(main);
    if (it != pool.end()) return &it->second;
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;        // creation may happen while `main` is capturing
    cudaThreadExchangeStreamCaptureMode(&mode);
    SideStream ss;
    bool ok = cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3; ++i) ok = ok && cudaEventCreateWithFlags(&ss.ring[i], cudaEventDisableTiming) == cudaSuccess;
    cudaThreadExchangeStreamCaptureMode(&mode);
    if (!ok) return nullptr;
    return &(pool[main] = ss);
}
static int side_fork(SideStream* ss, cudaStream_t main) {
    IR_CHECK_CUDA(cudaEventRecord(ss->fork, main));
    IR_CHECK_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
    return IR_OK;
}
static int side_join(SideStream* ss, cudaStream_t main) {
    IR_CHECK_CUDA(cudaEventRecord(ss->join, ss->s));
    IR_CHECK_CUDA(cudaStreamWaitEvent(main, ss->join, 0));
    return IR_OK;
}

static const int kCh[5] = {32, 64, 128, 128, 128};
// layer -> (input level, output level, K, map id in kcount: 0-4 = k3 at level l, 5-8 = k2 l->l+1)
struct LayerInfo { int lin, lout, K, map, cin, cout, resid_from; };
static LayerInfo layer_info(int idx, int cin0) {
    LayerInfo li;
    if (idx == 0) { li.lin = 0; li.lout = 0; li.K = 27; li.map = 0; li.cin = cin0; li.cout = kCh[0]; li.resid_from = -1; return li; }
    const int s = (idx - 1) / 3 + 1, r = (idx - 1) % 3;
    if (r == 0) { li.lin = s - 1; li.lout = s; li.K = 8; li.map = 5 + s - 1; li.cin = kCh[s - 1]; li.cout = kCh[s]; li.resid_from = -1; }
    else { li.lin = s; li.lout = s; li.K = 27; li.map = s; li.cin = kCh[s]; li.cout = kCh[s]; li.resid_from = (r == 2) ? idx - 2 : -1; }
    return li;
}

static inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }

extern "C" int ir_encoder_train_layout(int64_t n_max, const int32_t* n_lvl_in, int32_t cin, ir_encoder_train_layout_t* L) {
    IR_CHECK_ARG(n_max > 0 && L && cin >= 1 && cin <= 128);
    memset(L, 0, sizeof(*L));
    int32_t bound[5];
    for (int l = 0; l < 5; ++l) bound[l] = n_lvl_in ? n_lvl_in[l] : (int32_t)n_max;     // NULL: shape-independent (capacity) layout
    const int32_t* n_lvl = bound;
    int64_t off = 0, rows_max = 1;
    auto take = [&](int64_t bytes) { int64_t o = off; off = al(off + bytes); return o; };
    for (int l = 0; l < 5; ++l) { IR_CHECK_ARG(n_lvl[l] >= 0 && n_lvl[l] <= n_max); if (n_lvl[l] > rows_max) rows_max = n_lvl[l]; }
    for (int i = 0; i < IR_ENC_LAYERS; ++i) {
        const LayerInfo li = layer_info(i, cin);
        const int64_t rows = n_lvl[li.lout] > 0 ? n_lvl[li.lout] : 1;
        L->off_y[i] = take(rows * li.cout * 4);
        L->off_out[i] = take(rows * li.cout * 4);
        L->off_mean[i] = take(li.cout * 4);
        L->off_rstd[i] = take(li.cout * 4);
    }
    L->off_bn_scratch = take(ir_bn_scratch_floats(256) * 4);
    for (int m = 0; m < 9; ++m) {
        const int K = m < 5 ? 27 : 8;
        L->off_tr_out[m] = take((int64_t)K * n_max * 4);
        L->off_tr_slot[m] = take((int64_t)K * n_max * 4);
    }
    for (int i = 0; i < 6; ++i) L->off_grad[i] = take(rows_max * 128 * 4);
    {   // transposed weights (K,Cout,Cin) of layers 1..12, back to back (256-byte aligned each)
        int64_t tot = 0;
        for (int i = 1; i < IR_ENC_LAYERS; ++i) { const LayerInfo li = layer_info(i, cin); tot += al((int64_t)li.K * li.cin * li.cout * 4); }
        L->off_wt = take(tot);
    }
    L->off_absmax = take(256);
    L->total_bytes = off;
    return IR_OK;
}

__global__ void k_transpose_w(const float* __restrict__ w, int K, int cin, int cout, float* __restrict__ wt) {
    const long long total = (long long)K * cin * cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        const long long t = i / cin;
        const int co = (int)(t % cout), k = (int)(t / cout);
        wt[i] = w[((long long)k * cin + ci) * cout + co];                  // wt[k][co][ci]
    }
}

struct Tr {
    ir_encoder_layout_t W;          // encoder workspace (rulebooks)
    ir_encoder_train_layout_t A;    // arena
    char *ws, *ar;
    int64_t n_max;
    int32_t n[5];                   // host row counts, or the capacity n_max when the caller gave none
    bool bounded;                   // true: row counts are read on the device (launch sequence independent of the input)
    const int* nlvl_dev() const { return (const int*)(ws + W.off_nlvl); }
    const int* ndev(int l) const { return bounded ? nlvl_dev() + l : nullptr; }
    const int* kcount(int map) const { return (const int*)(ws + W.off_kcount) + map * 32; }
    const int* in_idx(int map) const { return (const int*)(ws + (map < 5 ? W.off_k3_in[map] : W.off_k2_in[map - 5])); }
    const int* slot(int map) const { return (const int*)(ws + (map < 5 ? W.off_k3_slot[map] : W.off_k2_slot[map - 5])); }
    float* T() const { return (float*)(ws + W.off_T); }
    float* y(int i) const { return (float*)(ar + A.off_y[i]); }
    float* out(int i) const { return (float*)(ar + A.off_out[i]); }
    float* mean(int i) const { return (float*)(ar + A.off_mean[i]); }
    float* rstd(int i) const { return (float*)(ar + A.off_rstd[i]); }
    float* bn_scratch() const { return (float*)(ar + A.off_bn_scratch); }
    int* tr_out(int map) const { return (int*)(ar + A.off_tr_out[map]); }
    int* tr_slot(int map) const { return (int*)(ar + A.off_tr_slot[map]); }
    float* grad(int i) const { return (float*)(ar + A.off_grad[i]); }
    float* wt(int layer, int cin0) const {
        int64_t o = A.off_wt;
        for (int i = 1; i < layer; ++i) { const LayerInfo li = layer_info(i, cin0); o += al((int64_t)li.K * li.cin * li.cout * 4); }
        return (float*)(ar + o);
    }
    float* absmax() const { return (float*)(ar + A.off_absmax); }
};

static int tr_open(const ir_encoder_train_params* p, void* ws, int64_t n_max, const int32_t* n_lvl, void* arena, Tr* t) {
    IR_CHECK_ARG(p && ws && arena);
    int r;
    if ((r = ir_encoder_layout(n_max, &t->W)) != IR_OK) return r;
    if ((r = ir_encoder_train_layout(n_max, n_lvl, p->cin, &t->A)) != IR_OK) return r;
    t->ws = (char*)ws; t->ar = (char*)arena; t->n_max = n_max; t->bounded = (n_lvl == nullptr);
    for (int l = 0; l < 5; ++l) {
        t->n[l] = n_lvl ? n_lvl[l] : (int32_t)n_max;
        IR_CHECK_ARG(t->n[l] > 0);
    }
    return IR_OK;
}

extern "C" int ir_encoder_train_forward(const ir_encoder_train_params* p, const float* feats0, void* ws, int64_t n_max,
                                        const int32_t* n_lvl, void* arena, ir_stream_t stream) {
    Tr t;
    int r;
    if ((r = tr_open(p, ws, n_max, n_lvl, arena, &t)) != IR_OK) return r;
    const float* f0 = feats0 ? feats0 : (const float*)(t.ws + t.W.off_feat0);
    cudaStream_t st = (cudaStream_t)stream;
    SideStream* ss = side_stream_for(st);
    IR_CHECK_ARG(ss);
    // What the backward needs beside the activations — the transposed rulebooks of all nine maps (out_idx for wgrad,
    // slot_in for the dgrad reduce) and W^T of every layer with an input gradient — depends only on the maps and on
    // the weights of this step: built here, on the helper stream beside the 13 layers, instead of in front of the
    // backward chain.
    if ((r = side_fork(ss, st)) != IR_OK) return r;
    for (int m = 0; m < 9; ++m) {
        const int K = m < 5 ? 27 : 8, lout = m < 5 ? m : m - 5 + 1;
        if ((r = ir_rulebook_transpose(t.in_idx(m), t.slot(m), K, n_max, t.nlvl_dev() + lout, t.n[lout], t.tr_out(m),
                                       t.tr_slot(m), (ir_stream_t)ss->s)) != IR_OK) return r;
    }
    for (int i = 1; i < IR_ENC_LAYERS; ++i) {
        const LayerInfo li = layer_info(i, p->cin);
        const long long tot = (long long)li.K * li.cin * li.cout;
        k_transpose_w<<<ir_min_i(ir_div_up(tot, 256), IR_NUM_SMS * 4), 256, 0, ss->s>>>(p->weight[i], li.K, li.cin, li.cout, t.wt(i, p->cin));
        IR_CHECK_LAUNCH();
    }
    for (int i = 0; i < IR_ENC_LAYERS; ++i) {
        const LayerInfo li = layer_info(i, p->cin);
        const float* fin = (i == 0) ? f0 : t.out(i - 1);
        const int tc = (p->use_tc & 1) && li.cin >= 32 && ((reinterpret_cast<uintptr_t>(p->weight[i]) & 15) == 0);
        if ((r = ir_spconv_layer(fin, li.cin, li.cout, li.K, t.in_idx(li.map), n_max, t.slot(li.map), t.kcount(li.map),
                                 t.nlvl_dev() + li.lout, t.n[li.lout], p->weight[i], tc ? p->weight[i] : nullptr, tc,
                                 nullptr, nullptr, nullptr, 0, t.T(), t.y(i), stream)) != IR_OK) return r;
        if ((r = ir_bn_train_fwd(t.y(i), t.ndev(li.lout), t.n[li.lout], li.cout, p->gamma[i], p->beta[i],
                                 li.resid_from >= 0 ? t.out(li.resid_from) : nullptr, 1, p->eps, p->momentum[i],
                                 p->running_mean[i], p->running_var[i], t.bn_scratch(), t.mean(i), t.rstd(i), t.out(i),
                                 stream)) != IR_OK) return r;
    }
    return side_join(ss, st);
}

extern "C" int ir_encoder_train_backward(const ir_encoder_train_params* p, const float* feats0, void* ws, int64_t n_max,
                                         const int32_t* n_lvl, void* arena, const float* dout,
                                         const ir_encoder_train_grads* g, ir_stream_t stream) {
    return ir_encoder_train_backward_range(p, feats0, ws, n_max, n_lvl, arena, dout, g, 4, 0, stream);
}

// Stages stage_hi .. stage_lo of the backward (4 .. 1 = the residual stages from the output down, 0 = the stem).  A
// caller that wants the gradients of the deep stages early — they hold 2/3 of an encoder's parameters and are final long
// before the large shallow levels are done — splits the pass in two calls on the same arena (the gradient between them
// stays in the arena) and starts their all-reduce in between.
extern "C" int ir_encoder_train_backward_range(const ir_encoder_train_params* p, const float* feats0, void* ws, int64_t n_max,
                                               const int32_t* n_lvl, void* arena, const float* dout,
                                               const ir_encoder_train_grads* g, int32_t stage_hi, int32_t stage_lo,
                                               ir_stream_t stream) {
    Tr t;
    int r;
    if ((r = tr_open(p, ws, n_max, n_lvl, arena, &t)) != IR_OK) return r;
    IR_CHECK_ARG(g && stage_hi <= 4 && stage_lo >= 0 && stage_lo <= stage_hi && (dout || stage_hi < 4));
    cudaStream_t st = (cudaStream_t)stream;
    const float* f0 = feats0 ? feats0 : (const float*)(t.ws + t.W.off_feat0);
    // (transposed rulebooks and W^T were built by ir_encoder_train_forward on the same arena)
    SideStream* ss = side_stream_for(st);
    IR_CHECK_ARG(ss);
    float *S0 = t.grad(0), *S2 = t.grad(2), *S3 = t.grad(3);
    const float* up = stage_hi == 4 ? dout : S0;   // gradient w.r.t. the output of the layer being processed
    float* dyb[3] = {t.grad(1), t.grad(4), t.grad(5)};         // ring of dY buffers (+ one range scalar each)
    bool ring_used[3] = {false, false, false};
    // one layer: BN backward (-> DY, skip gradient in `dres`), wgrad, optional dgrad (+ `add`) into `dx`.
    // wgrad and dgrad both read DY and nothing of each other, and at the large levels the wgrad is the longer of the two:
    // it goes to the helper stream and is NOT joined per layer — DY lives in a ring of three buffers, so the chain
    // (BN backward -> dgrad -> next layer) runs up to two layers ahead of the weight gradients, and only a layer that
    // wants a ring slot back waits for the wgrad that still reads it.
    auto layer_bwd = [&](int i, const float* gin, float* dres, float* dx, const float* add) -> int {
        const LayerInfo li = layer_info(i, p->cin);
        const float* xin = (i == 0) ? f0 : t.out(i - 1);
        const int slot = i % 3;
        float* DY = dyb[slot];
        float* amax = t.absmax() + slot;
        int rr;
        if (ring_used[slot]) IR_CHECK_CUDA(cudaStreamWaitEvent(st, ss->ring[slot], 0));
        if ((rr = ir_bn_train_bwd(gin, t.out(i), t.y(i), t.ndev(li.lout), t.n[li.lout], li.cout, t.mean(i), t.rstd(i), p->gamma[i], 1,
                                  t.bn_scratch(), DY, dres, g->dgamma[i], g->dbeta[i], amax, stream)) != IR_OK) return rr;
        if ((rr = side_fork(ss, st)) != IR_OK) return rr;
        if ((rr = ir_spconv_wgrad_scaled(xin, li.cin, DY, amax, li.cout, li.K, t.in_idx(li.map), t.tr_out(li.map),
                                         t.kcount(li.map), n_max, (p->use_tc >> 2) & 1, g->dweight[i], (ir_stream_t)ss->s)) != IR_OK) return rr;
        IR_CHECK_CUDA(cudaEventRecord(ss->ring[slot], ss->s));
        ring_used[slot] = true;
        if (dx) {
            // dgrad = forward pipeline on the transposed rulebook with W^T; `add` (the skip gradient) rides in the
            // reduce epilogue's residual slot.  use_tc bit 1: tcgen05 pair-GEMM with the gathered gradient rows
            // range-scaled by max|dY| (written by the BN backward above); otherwise the exact fp32 SIMT pair-GEMM.
            if ((rr = ir_spconv_layer_scaled(DY, amax, li.cout, li.cin, li.K, t.tr_out(li.map), n_max, t.tr_slot(li.map),
                                             t.kcount(li.map), t.nlvl_dev() + li.lin, t.n[li.lin], t.wt(i, p->cin), (p->use_tc >> 1) & 1,
                                             add, t.T(), dx, stream)) != IR_OK) return rr;
        }
        return IR_OK;
    };
    for (int s = stage_hi; s >= 1 && s >= stage_lo; --s) {
        const int a = 1 + 3 * (s - 1), b = a + 1, c = a + 2;
        if ((r = layer_bwd(c, up, S2, S3, nullptr)) != IR_OK) return r;      // skip gradient -> S2, d out[b] -> S3
        if ((r = layer_bwd(b, S3, nullptr, S3, S2)) != IR_OK) return r;      // d out[a] = dgrad + skip -> S3 (S3 was consumed by BN bwd)
        if ((r = layer_bwd(a, S3, nullptr, S0, nullptr)) != IR_OK) return r; // d out[prev] -> S0
        up = S0;
    }
    if (stage_lo == 0 && (r = layer_bwd(0, up, nullptr, nullptr, nullptr)) != IR_OK) return r;  // stem: no input gradient
    return side_join(ss, st);                                                 // all weight gradients are final
}
