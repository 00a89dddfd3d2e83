// Sparse-voxel encoder orchestration + C ABI: workspace layout, reset, voxelise, map build,
// 13-layer feature pass, segmented max-pool.  One call = one chain of async launches on `stream`;
// no allocation, no host sync, device-side counts.
// Reference: SparseConvEncoder / BEVEncoder (models/basic_blocks.py:59-95,136-171),
// GlobalMaxPooling (models/attribute_module.py:105).
#include <stdarg.h>
#include <string.h>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"
#include "kernels.cuh"

// ------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";
void ir_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* ir_last_error(void) { return g_err; }
extern "C" int ir_version(void) { return 100; }
extern "C" int ir_check_device(int device) {
    cudaDeviceProp prop;
    IR_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        ir_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                     prop.major, prop.minor);
        return IR_ERR_UNSUPPORTED;
    }
    return IR_OK;
}

// ------------------------------------------------------------------ launch counter + profiling hooks
static long long g_launches = 0;
void ir_count_launch(int n) { g_launches += n; }
extern "C" int64_t ir_launch_count(void) { return g_launches; }

#define IR_PROF_MAX 512
static int g_prof_on = 0, g_prof_n = 0;
static int g_prof_force_gemm = 0;     // tests: run the stem through the generic pair-GEMM + reduce path
static cudaEvent_t g_prof_ev[IR_PROF_MAX][3];
static int g_prof_made = 0;
static int g_prof_meta[IR_PROF_MAX][4];   // cin, cout, K, use_tc

static int g_prof_rep = 1;
extern "C" int ir_profile_enable(int on) {
    g_prof_on = on > 0;
    g_prof_rep = on > 1 ? on : 1;        // on = N > 1: launch each timed kernel N times back to back
    g_prof_n = 0;
    return IR_OK;
}
// per conv layer since ir_profile_enable(1): pair-GEMM ms, reduce ms, (cin,cout,K,tc). Synchronises.
extern "C" int ir_profile_read(float* gemm_ms, float* reduce_ms, int32_t* meta, int32_t cap, int32_t* n_out) {
    IR_CHECK_ARG(gemm_ms && reduce_ms && meta && n_out);
    const int n = g_prof_n < cap ? g_prof_n : cap;
    for (int i = 0; i < n; ++i) {
        IR_CHECK_CUDA(cudaEventSynchronize(g_prof_ev[i][2]));
        IR_CHECK_CUDA(cudaEventElapsedTime(&gemm_ms[i], g_prof_ev[i][0], g_prof_ev[i][1]));
        IR_CHECK_CUDA(cudaEventElapsedTime(&reduce_ms[i], g_prof_ev[i][1], g_prof_ev[i][2]));
        gemm_ms[i] /= (float)g_prof_rep;           // per-launch average
        reduce_ms[i] /= (float)g_prof_rep;
        for (int j = 0; j < 4; ++j) meta[4 * i + j] = g_prof_meta[i][j];
    }
    *n_out = n;
    g_prof_n = 0;
    return IR_OK;
}
static cudaEvent_t prof_event(int i, int j) {
    while (g_prof_made <= i && g_prof_made < IR_PROF_MAX) {
        for (int q = 0; q < 3; ++q) cudaEventCreate(&g_prof_ev[g_prof_made][q]);
        ++g_prof_made;
    }
    return g_prof_ev[i][j];
}

// 0 (default): one pair-GEMM + one reduce launch per layer (k_pairgemm_tc / k_reduce_epilogue), PDL-chained;
// 1: the 13 layers of ir_encoder_features[_pair] run as ONE persistent launch (encoder_persist.cu) — measured slower at
//    the bench size, faster for small scenes where per-launch floors dominate.
static int g_encoder_mode = 0;
extern "C" int ir_encoder_mode_set(int mode) {
    IR_CHECK_ARG(mode == 0 || mode == 1);
    g_encoder_mode = mode;
    return IR_OK;
}

// timeline aid (tools/timeline.py): one thread writes %globaltimer into buf[idx]; usable inside stream capture
__global__ void k_stamp(unsigned long long* buf, int idx) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    buf[idx] = t;
}
extern "C" int ir_debug_stamp(uint64_t* buf, int32_t idx, ir_stream_t stream) {
    IR_CHECK_ARG(buf != nullptr && idx >= 0);
    k_stamp<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)buf, idx);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ layout
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

extern "C" int ir_encoder_layout(int64_t n_max, ir_encoder_layout_t* L) {
    IR_CHECK_ARG(n_max > 0 && n_max < (1ll << 26) && L != nullptr);
    memset(L, 0, sizeof(*L));
    L->n_max = n_max;
    int64_t cap = 1024;
    while (cap < 2 * n_max) cap <<= 1;
    L->cap = cap;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 1024); return o; };
    L->off_nlvl = take(8 * 4);
    L->off_kcount = take(9 * 32 * 4);
    L->scan_stride = 1 + (n_max + 511) / 512 + 1;          // ticket + one word per 512-row scan tile (coords.cu SCAN_TILE)
    L->off_scan = take(5 * L->scan_stride * 8);
    L->off_sync = take(512);               // ticket / phase counters of the persistent encoder kernel
    L->zero_bytes = off - L->off_nlvl;
    L->off_keys = take(5 * cap * 8);
    L->off_vals = take(5 * cap * 8);
    for (int l = 0; l < IR_ENC_LEVELS; ++l) L->off_coords[l] = take(n_max * 16);
    L->off_pslot = take(5 * n_max * 4);
    for (int l = 0; l < IR_ENC_LEVELS; ++l) {
        L->off_k3_in[l] = take(27 * n_max * 4);
        L->off_k3_slot[l] = take(27 * n_max * 4);
    }
    for (int l = 0; l < 4; ++l) {
        L->off_k2_in[l] = take(8 * n_max * 4);
        L->off_k2_slot[l] = take(8 * n_max * 4);
    }
    L->off_feat0 = take(n_max * 8 * 4);
    for (int i = 0; i < 3; ++i) L->off_feat[i] = take(n_max * 128 * 4);
    L->off_T = take(27 * n_max * 128 * 4);
    L->total_bytes = off;
    return IR_OK;
}

extern "C" size_t ir_encoder_workspace_bytes(int64_t n_max) {
    ir_encoder_layout_t L;
    if (ir_encoder_layout(n_max, &L) != IR_OK) return 0;
    return (size_t)L.total_bytes;
}

struct Ws {
    ir_encoder_layout_t L;
    char* base;
    int* nlvl() const { return (int*)(base + L.off_nlvl); }
    int* kcount(int map) const { return (int*)(base + L.off_kcount) + map * 32; }
    void* sync() const { return (void*)(base + L.off_sync); }
    unsigned long long* scan(int i) const { return (unsigned long long*)(base + L.off_scan) + i * L.scan_stride; }
    IrTable table(int l) const {
        return ir_table_view2(base + L.off_keys + (int64_t)l * L.cap * 8,
                              base + L.off_vals + (int64_t)l * L.cap * 8, L.cap);
    }
    int32_t* coords(int l) const { return (int32_t*)(base + L.off_coords[l]); }
    int* vslot() const { return (int*)(base + L.off_pslot); }                       // voxelize insert slots
    int* pslot() const { return (int*)(base + L.off_pslot) + L.n_max; }             // ancestor slots [4][n_max]
    IrLevels levels() const { IrLevels lt; for (int l = 0; l < 5; ++l) lt.t[l] = table(l); return lt; }
    int* k3_in(int l) const { return (int*)(base + L.off_k3_in[l]); }
    int* k3_slot(int l) const { return (int*)(base + L.off_k3_slot[l]); }
    int* k2_in(int l) const { return (int*)(base + L.off_k2_in[l]); }
    int* k2_slot(int l) const { return (int*)(base + L.off_k2_slot[l]); }
    float* feat0() const { return (float*)(base + L.off_feat0); }
    float* feat(int i) const { return (float*)(base + L.off_feat[i]); }
    float* T() const { return (float*)(base + L.off_T); }
};

static int ws_open(void* ws, int64_t n_max, Ws* w) {
    IR_CHECK_ARG(ws != nullptr);
    int r = ir_encoder_layout(n_max, &w->L);
    if (r != IR_OK) return r;
    w->base = (char*)ws;
    return IR_OK;
}

extern "C" int ir_encoder_reset(void* ws, int64_t n_max, ir_stream_t stream) {
    Ws w;
    int r = ws_open(ws, n_max, &w);
    if (r != IR_OK) return r;
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(w.base + w.L.off_nlvl, 0, (size_t)w.L.zero_bytes, st));
    // keys (EMPTY) and {minrow,row} (NOROW) share the 0x7F byte pattern and are contiguous
    IR_CHECK_CUDA(cudaMemsetAsync(w.base + w.L.off_keys, 0x7F, (size_t)(w.L.off_vals - w.L.off_keys + 5 * w.L.cap * 8), st));
    return IR_OK;
}

extern "C" int ir_voxelize(const float* pts, const int32_t* cand, int32_t n_cand, int32_t ppi,
                           int32_t fdim, double voxel, void* ws, int64_t n_max,
                           ir_stream_t stream) {
    Ws w;
    int r = ws_open(ws, n_max, &w);
    if (r != IR_OK) return r;
    IR_CHECK_ARG(pts && cand && n_cand > 0 && ppi > 0 && fdim >= 3 && fdim <= 8 && voxel > 0);
    IR_CHECK_ARG((int64_t)n_cand * ppi <= n_max && n_cand < IR_MAX_BATCH);
    return irk_voxelize(pts, cand, n_cand, ppi, fdim, voxel, w.levels(), w.vslot(), w.coords(0),
                        w.feat0(), w.nlvl() + 0, w.scan(0), w.pslot(), n_max, (cudaStream_t)stream);
}

extern "C" int ir_encoder_build_maps(const int32_t* coords0, int32_t n0, const int32_t* n0_dev,
                                     void* ws, int64_t n_max, ir_stream_t stream) {
    Ws w;
    int r = ws_open(ws, n_max, &w);
    if (r != IR_OK) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t* c0 = w.coords(0);
    long long rows0 = n_max;
    if (coords0 != nullptr) {
        IR_CHECK_ARG(n0 >= 0 && n0 <= n_max);
        if ((r = ir_encoder_reset(ws, n_max, stream)) != IR_OK) return r;
        if ((r = irk_levels_from_coords(coords0, n0, n0_dev, w.levels(), w.pslot(), n_max, w.nlvl(), st)) != IR_OK) return r;
        c0 = coords0;
        rows0 = n0;
    }
    // levels 1..4 (stride 2,4,8,16): one ordered compaction per level, all from the level-0 rows
    if ((r = irk_levels_compact(c0, rows0, w.nlvl(), w.levels(), w.pslot(), n_max, w.coords(1), w.coords(2),
                                w.coords(3), w.coords(4), w.scan(0), w.L.scan_stride, st)) != IR_OK) return r;
    // kernel maps: k3 at every level, k2s2 between levels — one launch
    IrKmapArgs ka;
    ka.nlvl = w.nlvl();
    ka.lt = w.levels();
    ka.kcount = w.kcount(0);
    ka.n_max = n_max;
    for (int l = 0; l < IR_ENC_LEVELS; ++l) {
        ka.coords[l] = (const int4*)((l == 0) ? c0 : w.coords(l));
        ka.k3_in[l] = w.k3_in(l);
        ka.k3_slot[l] = w.k3_slot(l);
    }
    for (int l = 0; l < 4; ++l) { ka.k2_in[l] = w.k2_in(l); ka.k2_slot[l] = w.k2_slot(l); }
    return irk_kmap_all(ka, rows0, st);
}

// Live per-launch spans (bench.py roofline): while a stamp buffer is set, every conv kernel launched through conv_layer
// (also under stream capture, so the replayed graph keeps writing them) records {begin, end} GPU-timer ns into slot
// 2*i / 2*i+1 (pair-GEMM / reduce-or-stem of the i-th layer call); meta = (cin, cout, K, tcgen05) per layer call.
#define IR_STAMP_MAX 256
static unsigned long long* g_stamp_buf = nullptr;
static int g_stamp_n = 0;
static int g_stamp_meta[IR_STAMP_MAX][4];
extern "C" int ir_conv_stamps_set(uint64_t* buf) {
    g_stamp_buf = (unsigned long long*)buf;
    if (buf) g_stamp_n = 0;               // slots restart with every new buffer; the meta of the last one stays readable
    return IR_OK;
}
extern "C" int ir_conv_stamps_meta(int32_t* meta, int32_t cap, int32_t* n_out) {
    IR_CHECK_ARG(meta && n_out);
    const int n = g_stamp_n < cap ? g_stamp_n : cap;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 4; ++j) meta[4 * i + j] = g_stamp_meta[i][j];
    *n_out = n;
    return IR_OK;
}

static int conv_layer(const IrConvBatch& b_in, int cin, int cout, int K, const float* const* wprep, int use_tc,
                      cudaStream_t st) {
    int r;
    IrConvBatch b = b_in;
    bool tc = use_tc && cin >= 32;
    unsigned long long* stamp = nullptr;
    if (g_stamp_buf && g_stamp_n < IR_STAMP_MAX) {
        bool tcq = tc;
        for (int g = 0; g < b.G; ++g) tcq = tcq && wprep[g] != nullptr;
        g_stamp_meta[g_stamp_n][0] = cin; g_stamp_meta[g_stamp_n][1] = cout; g_stamp_meta[g_stamp_n][2] = K; g_stamp_meta[g_stamp_n][3] = tcq;
        stamp = g_stamp_buf + 4 * g_stamp_n++;           // {gemm begin, gemm end, reduce begin, reduce end}
    }
    for (int g = 0; g < b.G; ++g) tc = tc && wprep[g] != nullptr;
    const int pi = (g_prof_on && g_prof_n < IR_PROF_MAX) ? g_prof_n++ : -1;
    if (pi >= 0) {
        g_prof_meta[pi][0] = cin; g_prof_meta[pi][1] = cout; g_prof_meta[pi][2] = K; g_prof_meta[pi][3] = tc;
        cudaEventRecord(prof_event(pi, 0), st);
    }
    if (!tc && K == 27 && cout == 32 && cin <= 8 && !g_prof_force_gemm) {
        // stem: direct fused conv (no T round trip); recorded as the "reduce" span of the profile
        if (pi >= 0) cudaEventRecord(prof_event(pi, 1), st);
        b.stamp = stamp ? stamp + 2 : nullptr;
        r = irk_stem_direct(b, cin, st);
        if (pi >= 0) cudaEventRecord(prof_event(pi, 2), st);
        return r;
    }
    const int rep = (pi >= 0) ? g_prof_rep : 1;      // profiling: idempotent repeats amortise the host launch gap
    for (int it = 0; it < rep; ++it) {
        b.stamp = stamp;
        if (tc) {
            IrConvBatch bt = b;
            for (int g = 0; g < b.G; ++g) bt.p[g].weight = wprep[g];  // 16-byte aligned copy for the TMA bulk copy
            r = irk_pairgemm_tc(bt, cin, cout, K, st);
        } else {
            r = irk_pairgemm_simt(b, cin, cout, K, st);
        }
        if (r != IR_OK) return r;
    }
    if (pi >= 0) cudaEventRecord(prof_event(pi, 1), st);
    b.stamp = stamp ? stamp + 2 : nullptr;
    for (int it = 0; it < rep; ++it) {
        r = irk_reduce_epilogue(b, cout, K, st);
        if (r != IR_OK) return r;
    }
    if (pi >= 0) cudaEventRecord(prof_event(pi, 2), st);
    return r;
}

extern "C" int ir_spconv_layer(const float* feat_in, int32_t cin, int32_t cout, int32_t K,
                               const int32_t* in_idx, int64_t seg_cap,
                               const int32_t* slot, const int32_t* count,
                               const int32_t* n_out_dev, int64_t n_max, const float* weight,
                               const float* wprep, int32_t use_tc, const float* scale,
                               const float* shift, const float* resid, int32_t relu, float* T,
                               float* out, ir_stream_t stream) {
    IR_CHECK_ARG(feat_in && in_idx && slot && count && n_out_dev && weight && T && out);
    IrConvBatch b;
    memset(&b, 0, sizeof(b));
    b.G = 1;
    b.p[0] = IrConvProblem{feat_in, in_idx, slot, count, n_out_dev, weight, scale, shift, resid, T, out,
                           (long long)seg_cap, (long long)n_max, relu};
    return conv_layer(b, cin, cout, K, &wprep, use_tc, (cudaStream_t)stream);
}

// ir_spconv_layer with the tcgen05 input-range scale (kernels.cuh: IrConvProblem::in_absmax); used by the
// dgrad of the training step, whose inputs are gradients of arbitrary magnitude.
extern "C" int ir_spconv_layer_scaled(const float* feat_in, const float* in_absmax, int32_t cin, int32_t cout, int32_t K,
                                      const int32_t* in_idx, int64_t seg_cap, const int32_t* slot, const int32_t* count,
                                      const int32_t* n_out_dev, int64_t n_max, const float* weight, int32_t use_tc,
                                      const float* resid, float* T, float* out, ir_stream_t stream) {
    IR_CHECK_ARG(feat_in && in_idx && slot && count && n_out_dev && weight && T && out);
    IrConvBatch b;
    memset(&b, 0, sizeof(b));
    b.G = 1;
    b.p[0] = IrConvProblem{feat_in, in_idx, slot, count, n_out_dev, weight, nullptr, nullptr, resid, T, out,
                           (long long)seg_cap, (long long)n_max, 0, in_absmax};
    const bool tc = use_tc && cin >= 32 && cout >= 32 && in_absmax != nullptr && (reinterpret_cast<uintptr_t>(weight) & 15) == 0;
    const float* wp = tc ? weight : nullptr;
    return conv_layer(b, cin, cout, K, &wp, tc, (cudaStream_t)stream);
}

// The 13-layer feature pass for one or two encoders (same topology) with shared launches.
static int encoder_features_multi(int G, const ir_encoder_params* const* ps, const float* const* feats0,
                                  void* const* wss, const int64_t* n_maxs, float* const* outs, cudaStream_t st) {
    Ws w[IR_MAX_GROUPS];
    int r;
    for (int g = 0; g < G; ++g) {
        if ((r = ws_open(wss[g], n_maxs[g], &w[g])) != IR_OK) return r;
        IR_CHECK_ARG(ps[g] != nullptr && outs[g] != nullptr && ps[g]->cin >= 1 && ps[g]->cin <= 128);
        IR_CHECK_ARG(ps[g]->cin == ps[0]->cin && ps[g]->use_tc == ps[0]->use_tc);
    }
    static const int ch[5] = {32, 64, 128, 128, 128};
    // the 13 layers as (shape, per-problem buffers); executed either by ONE persistent launch (encoder_persist.cu)
    // or layer by layer (pair-GEMM + reduce launches; the profiling hooks and the SIMT path live there)
    IrConvProblem layers[IR_ENC_LAYERS][IR_MAX_GROUPS];
    int cins[IR_ENC_LAYERS], couts[IR_ENC_LAYERS], Ks[IR_ENC_LAYERS];
    memset(layers, 0, sizeof(layers));
    auto add = [&](int idx, int cin, int cout, int K, auto&& fill) {
        cins[idx] = cin; couts[idx] = cout; Ks[idx] = K;
        for (int g = 0; g < G; ++g) {
            IrConvProblem& P = layers[idx][g];
            fill(g, P);
            P.weight = ps[g]->weight[idx];
            P.scale = ps[g]->bn_scale[idx];
            P.shift = ps[g]->bn_shift[idx];
            P.T = w[g].T();
            P.seg_cap = n_maxs[g];
            P.n_max = n_maxs[g];
            P.relu = 1;
        }
    };
    // stem: k3 at level 0
    add(0, ps[0]->cin, ch[0], 27, [&](int g, IrConvProblem& P) {
        P.fin = feats0[g] ? feats0[g] : w[g].feat0();
        P.in_idx = w[g].k3_in(0); P.slot = w[g].k3_slot(0); P.count = w[g].kcount(0);
        P.n_out_dev = w[g].nlvl() + 0; P.resid = nullptr; P.out = w[g].feat(0);
    });
    for (int s = 1; s <= 4; ++s) {
        const int l = s - 1, li = 1 + 3 * (s - 1);
        // down: k2 s2, level l -> l+1
        add(li + 0, ch[l], ch[s], 8, [&](int g, IrConvProblem& P) {
            P.fin = w[g].feat(0);
            P.in_idx = w[g].k2_in(l); P.slot = w[g].k2_slot(l); P.count = w[g].kcount(5 + l);
            P.n_out_dev = w[g].nlvl() + s; P.resid = nullptr; P.out = w[g].feat(1);
        });
        // residual block at level s: relu(bn(conv(relu(bn(conv(X))))) + X)
        add(li + 1, ch[s], ch[s], 27, [&](int g, IrConvProblem& P) {
            P.fin = w[g].feat(1);
            P.in_idx = w[g].k3_in(s); P.slot = w[g].k3_slot(s); P.count = w[g].kcount(s);
            P.n_out_dev = w[g].nlvl() + s; P.resid = nullptr; P.out = w[g].feat(2);
        });
        add(li + 2, ch[s], ch[s], 27, [&](int g, IrConvProblem& P) {
            P.fin = w[g].feat(2);
            P.in_idx = w[g].k3_in(s); P.slot = w[g].k3_slot(s); P.count = w[g].kcount(s);
            P.n_out_dev = w[g].nlvl() + s; P.resid = w[g].feat(1);
            P.out = (s == 4) ? outs[g] : w[g].feat(0);
        });
    }
    bool persist = g_encoder_mode == 1 && ps[0]->use_tc && !g_prof_on && ps[0]->cin <= 8;
    for (int g = 0; g < G && persist; ++g)
        for (int i = 1; i < IR_ENC_LAYERS; ++i) persist = persist && ps[g]->wprep[i] != nullptr;
    if (persist) {
        IrConvProblem lt[IR_ENC_LAYERS][IR_MAX_GROUPS];
        memcpy(lt, layers, sizeof(lt));
        for (int i = 1; i < IR_ENC_LAYERS; ++i)
            for (int g = 0; g < G; ++g) lt[i][g].weight = ps[g]->wprep[i];     // 16-byte aligned copies (TMA bulk source)
        return irk_encoder_persist(G, lt, cins, couts, Ks, IR_ENC_LAYERS, w[0].sync(), st);
    }
    // range guard of the split-fp16 GEMM: every reduce / stem epilogue records max|out| of its layer (one float per layer
    // and problem in the workspace's sync area, cleared here); the next layer's pair-GEMM scales its gathered rows by the
    // power of two that brings this maximum to 2^13 and un-scales in its epilogue (exact).
    float* amax[IR_MAX_GROUPS] = {nullptr, nullptr};
    if (ps[0]->use_tc) {
        for (int g = 0; g < G; ++g) {
            amax[g] = (float*)((char*)w[g].sync() + 4 * 34);
            IR_CHECK_CUDA(cudaMemsetAsync(amax[g], 0, 4 * 2 * IR_ENC_LAYERS, st));
        }
    }
    for (int i = 0; i < IR_ENC_LAYERS; ++i) {
        IrConvBatch b;
        memset(&b, 0, sizeof(b));
        b.G = G;
        const float* wp[IR_MAX_GROUPS] = {nullptr, nullptr};
        for (int g = 0; g < G; ++g) {
            b.p[g] = layers[i][g];
            wp[g] = ps[g]->wprep[i];
            if (amax[g]) {
                b.p[g].out_absmax = amax[g] + 2 * i;
                b.p[g].in_absmax = (i >= 1) ? amax[g] + 2 * (i - 1) : nullptr;
            }
        }
        if ((r = conv_layer(b, cins[i], couts[i], Ks[i], wp, ps[0]->use_tc, st)) != IR_OK) return r;
    }
    return IR_OK;
}

extern "C" int ir_encoder_features(const ir_encoder_params* p, const float* feats0, void* ws,
                                   int64_t n_max, float* feats_out, ir_stream_t stream) {
    return encoder_features_multi(1, &p, &feats0, &ws, &n_max, &feats_out, (cudaStream_t)stream);
}

extern "C" int ir_encoder_features_pair(const ir_encoder_params* pa, const float* feats0a, void* wsa,
                                        int64_t n_max_a, float* out_a, const ir_encoder_params* pb,
                                        const float* feats0b, void* wsb, int64_t n_max_b, float* out_b,
                                        ir_stream_t stream) {
    const ir_encoder_params* ps[2] = {pa, pb};
    const float* f0[2] = {feats0a, feats0b};
    void* wss[2] = {wsa, wsb};
    const int64_t nm[2] = {n_max_a, n_max_b};
    float* outs[2] = {out_a, out_b};
    return encoder_features_multi(2, ps, f0, wss, nm, outs, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ segmented max (GlobalMaxPooling)
__device__ __forceinline__ unsigned enc_f32(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

__global__ void k_segmax_scatter(const float* __restrict__ F, const int4* __restrict__ coords,
                                 const int* __restrict__ n_dev, int C, int n_seg,
                                 unsigned* __restrict__ enc) {
    const int n = *n_dev;
    const long long total = (long long)n * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / C), c = (int)(i - (long long)row * C);
        const int b = coords[row].w;
        if (b >= 0 && b < n_seg) atomicMax(&enc[(long long)b * C + c], enc_f32(F[i]));
    }
}
__global__ void k_segmax_decode(const unsigned* __restrict__ enc, int total, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        const unsigned e = enc[i];
        out[i] = e ? dec_f32(e) : 0.f;
    }
}

extern "C" int ir_segmax(const float* feats, const int32_t* coords, const int32_t* n_dev,
                         int64_t n_max, int32_t C, int32_t n_seg, uint32_t* enc_scratch,
                         float* out, ir_stream_t stream) {
    IR_CHECK_ARG(feats && coords && n_dev && enc_scratch && out && C > 0 && n_seg > 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(enc_scratch, 0, (size_t)n_seg * C * 4, st));
    const int grid = ir_min_i(ir_div_up(n_max * C > 0 ? n_max * C : 1, 256), IR_NUM_SMS * 8);
    k_segmax_scatter<<<grid, 256, 0, st>>>(feats, (const int4*)coords, n_dev, C, n_seg, enc_scratch);
    IR_CHECK_LAUNCH();
    k_segmax_decode<<<ir_div_up((long long)n_seg * C, 256), 256, 0, st>>>(enc_scratch, n_seg * C, out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
