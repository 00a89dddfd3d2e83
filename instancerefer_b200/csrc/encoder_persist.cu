// Persistent sparse-encoder kernel: all 13 conv layers of one or two encoders (SparseConvEncoder / BEVEncoder,
// models/basic_blocks.py:59-95,136-171) in ONE launch.
//
// Why this shape (numbers for the bench workload): a conv layer here is 3-120 k pairs over <= 26 k rows
// while its weights are 27 x Cin x Cout — per SM the weights (1.8 MB as split fp16) are 4x the gathered rows, so an
// output-stationary kernel that re-streams W[k] per row tile is bound by L2 -> SM weight traffic (148 x 1.8 MB per layer
// against the ~6.3 kB/clk L2 cap), not by the rows.  The weight-stationary pair-GEMM + ordered reduce stays; what this
// kernel removes is everything around it: 25 launches + prologues per encoder (TMEM allocation, barrier set-up, schedule
// scans), the grid-sized-by-capacity launches of the small levels, and the launch gaps between dependent layers.
//
//   * work = tickets handed out by one global atomic counter, in phase order:
//       phase 0            stem (direct fused conv, rows)
//       phase 2l-1, 2l     layer l >= 1:  pair-GEMM items (one kernel offset + a tile range, weights parked in TMEM)
//                                         then reduce items (row ranges: ordered sum over k, BN, residual, ReLU)
//     an item of phase p starts when done[p-1] == items[p-1]; a CTA only ever waits for tickets that RUNNING CTAs hold
//     (tickets are taken in order), so the kernel needs no co-residency guarantee and cannot deadlock against other
//     streams' kernels; the next ticket is prefetched while the current item runs, and the weights of a GEMM item are
//     staged (TMA bulk copy -> fp16 hi/lo -> TMEM) BEFORE the phase wait, i.e. behind the previous layer's reduce;
//   * item counts come from the device-side rulebook counts: the small levels use as many CTAs as they have tiles;
//   * TMEM, mbarriers and the schedule tables are set up once per CTA for all 25 phases;
//   * every reduce item also records max|out| of its layer; the next layer's GEMM scales its gathered rows by the power
//     of two that brings this maximum to 2^13 before the fp16 hi/lo split (exact, un-scaled in the epilogue): the
//     split-fp16 contraction is safe for any activation magnitude (range guard of the inference path);
//   * in-kernel data (activations, T) is read with ld.global.cg: other CTAs wrote it during this launch.
// The arithmetic of an item is the one of k_pairgemm_tc / k_reduce_epilogue / k_stem_direct (spconv_tc.cu, spconv.cu).
#include <stdlib.h>
#include <string.h>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace ep {
using namespace tc;

constexpr int MAX_LAYERS = 13;
constexpr int MAX_PHASES = 2 * MAX_LAYERS - 1;
constexpr int VSLOTS = 64;                    // virtual offsets (2 problems x 27) padded to two warp rounds
constexpr int RMIN = 17;                      // minimum rows of a reduce item: one per warp
constexpr int STEM_FLOATS = 27 * 8 * 32;      // stem weights of one problem in shared memory (Cin <= 8 -> 32)

// sync words (uint32 index) inside the caller's zero-initialised 512-byte area
constexpr int SY_TICKET = 0, SY_EXIT = 1, SY_DONE = 2, SY_ABSMAX = 34, SY_STAMP64 = 32;   // stamps: u64 index (byte 256)

struct LayerIO {
    const float* fin; const int* in_idx; const int* slot; const int* count; const int* n_out_dev;
    const float* weight; const float* scale; const float* shift; const float* resid; float* T; float* out;
};
struct Program {
    LayerIO io[MAX_LAYERS][2];
    long long seg_cap[2];
    int cin[MAX_LAYERS], cout[MAX_LAYERS], K[MAX_LAYERS];
    int n_layers, G, nominal, relu_last;
    unsigned* sync;
    unsigned long long* dbg;       // optional per-item time stamps [ticket][8] (tools/bench_encoder.py), or NULL
};

struct Sched {                                 // shared memory, built once per CTA
    int ginc[MAX_LAYERS][VSLOTS];              // inclusive prefix of GEMM items over virtual offsets
    int tiles[MAX_LAYERS][VSLOTS];
    int cnt[MAX_LAYERS][VSLOTS];
    int kofs[MAX_LAYERS][VSLOTS];              // first T row of the offset inside its problem
    int rows[MAX_LAYERS][2], rpi[MAX_LAYERS][2], ri0[MAX_LAYERS];   // reduce: rows, rows per item, items of problem 0
    int items[32], base[32], tick[32];       // real items, first ticket, tickets (>= items) per phase
};

struct Smem {
    static constexpr int OFF_STAGE = 0;
    static constexpr int STAGE_AREA = 65536;                 // max(NS*STAGE_BYTES, W_RAW 128x128x4, 2 stems)
    static constexpr int OFF_BAR = OFF_STAGE + STAGE_AREA;
    static constexpr int N_BAR = 2 * NS + 6;
    static constexpr int OFF_MISC = OFF_BAR + N_BAR * 8;     // s_tmem, ticket, item[16]
    static constexpr int OFF_SCHED = OFF_MISC + 256;
    static constexpr int BYTES = OFF_SCHED + (int)sizeof(Sched) + 1024;
};
static_assert(2 * STEM_FLOATS * 4 <= Smem::STAGE_AREA, "stem weights fit the stage area");
static_assert(NS * STAGE_BYTES <= Smem::STAGE_AREA, "stages fit");

__device__ __forceinline__ int ld_acquire(const unsigned* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float scale_from_absmax(float m) {
    // power of two bringing the maximum to [2^13, 2^14); 1 when the layer is all-zero or not finite
    return (m > 0.f && m < 3.0e38f) ? exp2f(13.f - floorf(log2f(m))) : 1.f;
}

struct Ctx {
    uint8_t* sm;
    uint32_t base;          // shared-space address of sm
    uint32_t tmem_base;
    int tid, lane, warp;
    unsigned long long* dbg;     // this item's stamp row or NULL
};

// ------------------------------------------------------------------------------------------ GEMM item
template <int CIN, int COUT>
struct Cfg {
    static constexpr int CINP = (CIN < PANEL) ? PANEL : CIN;
    static constexpr int KP = CINP / PANEL;
    static constexpr int W_RAW_BYTES = CIN * COUT * 4;
    static constexpr int COL_W_HI = 2 * TILE_M;
    static constexpr int COL_W_LO = COL_W_HI + CINP / 2;
    static_assert(COL_W_LO + CINP / 2 <= 256, "TMEM column budget (256 per CTA, two CTAs per SM)");
    static_assert(W_RAW_BYTES <= Smem::STAGE_AREA, "raw weight tile fits the stage area");
};

// All tickets of the previous phase finished: their writes are visible (release = the finisher's fence + atomic,
// acquire here, CTA barrier for the other threads).  need < 0: already known, only the barrier.
__device__ __forceinline__ void phase_wait(const unsigned* done, int need, int tid) {
    if (need >= 0 && tid == 0) {
        const long long t0 = clock64();
        while (ld_acquire(done) < need) {
            __nanosleep(40);
            if (clock64() - t0 > 4000000000ll) __trap();        // a protocol bug traps instead of hanging the GPU
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(N_THREADS) : "memory");     // named barrier: every role joins where it can
}

// One pair-GEMM item = offset kk of one problem, tiles [t_begin, t_end) of TILE_M pairs.  Its arguments live in shared
// memory (filled by the decoding warp); every warp role is its own function so that each gets its own register
// allocation (the 32-register TMEM load/store tuples of the epilogue and of the weight staging fragment a shared one).
struct GemmArgs {
    const float* fin; const int* idx_k; float* T; const uint8_t* w_src; const unsigned* in_absmax;
    uint32_t base, tmem_base;
    int t_begin, t_end, kofs, kcount;
};

struct Bars {
    uint32_t s_bar;
    __device__ __forceinline__ uint32_t full(int s) const { return s_bar + 8u * s; }
    __device__ __forceinline__ uint32_t empty(int s) const { return s_bar + 8u * (NS + s); }
    __device__ __forceinline__ uint32_t tfull(int b) const { return s_bar + 8u * (2 * NS + b); }
    __device__ __forceinline__ uint32_t tempty(int b) const { return s_bar + 8u * (2 * NS + 2 + b); }
    __device__ __forceinline__ uint32_t wfull() const { return s_bar + 8u * (2 * NS + 4); }
    __device__ __forceinline__ uint32_t wdone() const { return s_bar + 8u * (2 * NS + 5); }
};

// W[k] (Cin,Cout) fp32 staged in shared memory by TMA bulk copies -> x 2^8 -> fp16 hi/lo -> TMEM, where it stays for
// the item: the MMA reads its weight operand from tensor memory.  16 warps, 4 per TMEM sub-partition (warp % 4);
// `quarter` selects the warp's share of the packed K columns.
template <int CIN, int COUT>
__device__ __noinline__ void gemm_weights(const GemmArgs* __restrict__ A, const uint8_t* sm, int quarter, int warp, int lane) {
    using C = Cfg<CIN, COUT>;
    constexpr int NCOL = C::CINP / 2 / 4;
    const Bars B{A->base + Smem::OFF_BAR};
    mbar_wait(B.wfull(), 0u);
    const int sp = warp & 3;
    if (sp * 32 < COUT) {
        const float* ws = reinterpret_cast<const float*>(sm + Smem::OFF_STAGE) + sp * 32 + lane;
        const uint32_t tw = A->tmem_base + ((uint32_t)(sp * 32) << 16) + quarter * NCOL;
        uint32_t hi[NCOL], lo[NCOL];
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            const int c = 2 * (quarter * NCOL + q);
            const float w0 = (c < CIN) ? ws[c * COUT] * W_SCALE : 0.f;
            const float w1 = (c + 1 < CIN) ? ws[(c + 1) * COUT] * W_SCALE : 0.f;
            split2(w0, w1, hi[q], lo[q]);
        }
        if constexpr (NCOL == 16) { tmem_st16(tw + C::COL_W_HI, hi); tmem_st16(tw + C::COL_W_LO, lo); }
        else                      { tmem_st8(tw + C::COL_W_HI, hi);  tmem_st8(tw + C::COL_W_LO, lo); }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(B.wdone());
}

// Gather producers: NS groups of 128 threads, one shared-memory stage each.  Work items = (tile, 64-channel panel);
// group g owns the items g, g+NS, ...  A thread owns TILE_M/16 (row, 8-channel chunk) tasks per item: two 16-byte loads,
// range scale, fp32 -> fp16 hi/lo split, two 16-byte swizzled stores; the loads of the group's next item are issued
// right after the current one was handed to the tensor core.
template <int CIN, int COUT>
__device__ __noinline__ void gemm_producer(const GemmArgs* __restrict__ A, uint8_t* sm, int pw, int lane,
                                           const unsigned* done, int need) {
    using C = Cfg<CIN, COUT>;
    const Bars B{A->base + Smem::OFF_BAR};
    const float* F = A->fin;
    const int* __restrict__ idx_k = A->idx_k;
    const int t_begin = A->t_begin, kcount = A->kcount;
    const int n_items = (A->t_end - t_begin) * C::KP;
    const int grp = pw / WARPS_PER_GROUP;
    const int gt = (pw % WARPS_PER_GROUP) * 32 + lane;
    const int j = gt & 7;
    const int rbase = gt >> 3;
    int idx[TASKS];
    float4 va[TASKS], vb[TASKS];
    auto load_idx = [&](int it) {
        const int p0 = (t_begin + it / C::KP) * TILE_M;
#pragma unroll
        for (int i = 0; i < TASKS; ++i) {
            const int r = p0 + rbase + 16 * i;
            idx[i] = (it < n_items && r < kcount) ? __ldg(idx_k + r) : -1;
        }
    };
    auto load_rows = [&](int it) {
        const int ch = (it % C::KP) * PANEL + j * 8;
#pragma unroll
        for (int i = 0; i < TASKS; ++i) {
            va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx[i] >= 0 && ch < CIN) {
                const float4* src = reinterpret_cast<const float4*>(F + (long long)idx[i] * CIN + ch);
                va[i] = __ldcg(src);
                vb[i] = __ldcg(src + 1);
            }
        }
    };
    load_idx(grp);                                          // the rulebook is final long before this launch
    phase_wait(done, need, 1);                              // previous phase complete (tid 0 polls; this is the barrier)
    if (grp >= n_items) return;
    const unsigned amax_bits = __ldcg(A->in_absmax);        // in flight with the row loads; used at the first split
    load_rows(grp);
    load_idx(grp + NS);
    const float in_s = scale_from_absmax(__uint_as_float(amax_bits));
    uint8_t* st_hi = sm + Smem::OFF_STAGE + grp * STAGE_BYTES;
    uint8_t* st_lo = st_hi + PANEL_BYTES;
    uint32_t round = 0;
    mbar_wait(B.wdone(), 0u);                               // the raw weight tile has left the stage area
#pragma unroll 1
    for (int it = grp; it < n_items; it += NS, ++round) {
        mbar_wait(B.empty(grp), (round & 1u) ^ 1u);
#pragma unroll
        for (int i = 0; i < TASKS; ++i) {
            const int r = rbase + 16 * i;
            const int off = r * 128 + ((j ^ (r & 7)) << 4);
            uint4 h, l;
            split2(va[i].x * in_s, va[i].y * in_s, h.x, l.x);
            split2(va[i].z * in_s, va[i].w * in_s, h.y, l.y);
            split2(vb[i].x * in_s, vb[i].y * in_s, h.z, l.z);
            split2(vb[i].z * in_s, vb[i].w * in_s, h.w, l.w);
            *reinterpret_cast<uint4*>(st_hi + off) = h;
            *reinterpret_cast<uint4*>(st_lo + off) = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.full(grp));
        if (it + NS < n_items) load_rows(it + NS);            // idx[] holds the rows of item it+NS
        load_idx(it + 2 * NS);
    }
}

// MMA issuer (one thread): A = W^T from TMEM (lanes = output channels, 8 columns = 16 fp16), B = gathered pairs
// (128B-swizzled K-major stage), fp32 accumulators double-buffered in TMEM; hi*hi + hi*lo + lo*hi per K = 16.
template <int CIN, int COUT>
__device__ __noinline__ void gemm_mma(const GemmArgs* __restrict__ A) {
    using C = Cfg<CIN, COUT>;
    const Bars B{A->base + Smem::OFF_BAR};
    const uint32_t s_stage = A->base + Smem::OFF_STAGE, tmem_base = A->tmem_base;
    const int t_begin = A->t_begin, t_end = A->t_end;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TILE_M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t it = 0, acc_it = 0;
    mbar_wait(B.wdone(), 0u);
    tc_fence_after();
    for (int tile = t_begin; tile < t_end; ++tile) {
        const uint32_t b = acc_it & 1u;
        mbar_wait(B.tempty(b), ((acc_it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * TILE_M;
#pragma unroll 1
        for (int panel = 0; panel < C::KP; ++panel, ++it) {
            const int stage = it % NS;
            mbar_wait(B.full(stage), (it / NS) & 1u);
            tc_fence_after();
            const uint32_t g_hi = s_stage + stage * STAGE_BYTES;
            const uint32_t g_lo = g_hi + PANEL_BYTES;
            const uint32_t w_hi = tmem_base + C::COL_W_HI + panel * (PANEL / 2);
            const uint32_t w_lo = tmem_base + C::COL_W_LO + panel * (PANEL / 2);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t dg_hi = make_desc(g_hi + ks * 32), dg_lo = make_desc(g_lo + ks * 32);
                mma_f16_ts(d_tmem, w_hi + ks * 8, dg_hi, idesc, (panel | ks) ? 1u : 0u);
                mma_f16_ts(d_tmem, w_hi + ks * 8, dg_lo, idesc, 1u);
                mma_f16_ts(d_tmem, w_lo + ks * 8, dg_hi, idesc, 1u);
            }
            tc_commit(B.empty(stage));
        }
        tc_commit(B.tfull(b));
        ++acc_it;
    }
}

// Epilogue (warps 0-3): TMEM lane = output channel, column = pair, so every warp store instruction writes 32
// consecutive channels of one T row (one 128-byte line).
template <int CIN, int COUT>
__device__ __noinline__ void gemm_epilogue(const GemmArgs* __restrict__ A, int warp, int lane) {
    const Bars B{A->base + Smem::OFF_BAR};
    const uint32_t tmem_base = A->tmem_base;
    const int t_begin = A->t_begin, t_end = A->t_end, kofs = A->kofs, kcount = A->kcount;
    float* T = A->T;
    const float out_s = W_UNSCALE / scale_from_absmax(__uint_as_float(__ldcg(A->in_absmax)));
    uint32_t acc_it = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int p0 = tile * TILE_M;
        const int np = min(TILE_M, kcount - p0);
        const uint32_t b = acc_it & 1u;
        mbar_wait(B.tfull(b), (acc_it >> 1) & 1u);
        tc_fence_after();
        const int ch = warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + b * TILE_M;
        if (warp * 32 < COUT) {
            float* tcol = T + (long long)(kofs + p0) * COUT + ch;
#pragma unroll 1
            for (int c0 = 0; c0 < TILE_M; c0 += 32) {
                if (c0 >= np) break;
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (c0 + q < np) tcol[(long long)(c0 + q) * COUT] = __uint_as_float(v[q]) * out_s;
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.tempty(b));
        ++acc_it;
    }
}

template <int CIN, int COUT>
__device__ __forceinline__ void gemm_item(const Ctx& cx, const GemmArgs* A, const unsigned* done, int need) {
    using C = Cfg<CIN, COUT>;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    // ---- before the phase wait: this offset's weights (they do not depend on the previous layer)
    if (warp != 4) {
        if (tid == 0) {
            const Bars B{A->base + Smem::OFF_BAR};
            mbar_expect_tx(B.wfull(), (uint32_t)C::W_RAW_BYTES);
            const uint8_t* src = A->w_src;
            for (int o = 0; o < C::W_RAW_BYTES; o += 16384)
                bulk_g2s(A->base + Smem::OFF_STAGE + o, src + o, (uint32_t)min(16384, C::W_RAW_BYTES - o), B.wfull());
        }
        gemm_weights<CIN, COUT>(A, cx.sm, warp < 4 ? 0 : 1 + (warp - 5) / 4, warp, lane);
    }
    if (cx.dbg && tid == 0) cx.dbg[1] = globaltimer();
    if (warp >= 5) {
        gemm_producer<CIN, COUT>(A, cx.sm, warp - 5, lane, done, need);     // joins the phase barrier after its index loads
    } else {
        phase_wait(done, need, tid);                            // previous phase complete (CTA barrier inside)
        if (cx.dbg && tid == 0) cx.dbg[2] = globaltimer();
        if (warp == 4) { if (lane == 0) gemm_mma<CIN, COUT>(A); }
        else gemm_epilogue<CIN, COUT>(A, warp, lane);
    }
}

// ------------------------------------------------------------------------------------------ reduce item
// out[o] = act(scale * sum_{k asc} T[kofs[k] + slot[k][o]] + shift (+ resid[o])) for rows [r0, r1): warp per row, fixed
// order, no atomics (bitwise deterministic); returns the warp's max |out| in every lane.
template <int COUT>
__device__ __noinline__ float reduce_rows(const LayerIO& io, long long seg_cap, int K, const int* s_kofs, int r0, int r1,
                                          int relu, int warp, int lane, int nwarps) {
    constexpr int V = COUT / 32;
    constexpr int U = (V == 4) ? 8 : 9;          // independent T-row loads in flight per lane
    const float* T = io.T;
    const int* __restrict__ slot = io.slot + (long long)lane * seg_cap;
    const float* resid = io.resid;
    const int my_kofs = (lane < K) ? s_kofs[lane] : 0;
    float amax = 0.f;
    int o = r0 + warp;
    int my_next = (lane < K && o < r1) ? __ldg(slot + o) : -1;
    for (; o < r1; o += nwarps) {
        const int my = my_next;
        my_next = (lane < K && o + nwarps < r1) ? __ldg(slot + o + nwarps) : -1;      // next row's pair positions
        const int my_row = my_kofs + (my >= 0 ? my : 0);
        // present offsets of this row, ascending k: only those T rows are loaded, U at a time (missing pairs cost nothing;
        // the order of the additions is fixed, so the result is bitwise deterministic)
        const unsigned present = __ballot_sync(0xffffffffu, my >= 0);
        const int cnt = __popc(present);
        // lane L takes over the T row of the L-th present offset: the present rows sit compacted in lanes 0..cnt-1
        const int crow = __shfl_sync(0xffffffffu, my_row, (lane < cnt) ? (int)__fns(present, 0, lane + 1) : 0);
        float acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0.f;
        for (int j0 = 0; j0 < cnt; j0 += U) {
            float t[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = j0 + u < cnt;
                const int trow = __shfl_sync(0xffffffffu, crow, (j0 + u) & 31);
                const float* row = T + (long long)trow * COUT + lane * V;
                if constexpr (V == 4) {
                    const float4 q = ok ? __ldcg(reinterpret_cast<const float4*>(row)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    t[u][0] = q.x; t[u][1 % V] = q.y; t[u][2 % V] = q.z; t[u][3 % V] = q.w;
                } else if constexpr (V == 2) {
                    const float2 q = ok ? __ldcg(reinterpret_cast<const float2*>(row)) : make_float2(0.f, 0.f);
                    t[u][0] = q.x; t[u][1 % V] = q.y;
                } else {
                    t[u][0] = ok ? __ldcg(row) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += t[u][v];
        }
        // folded BN, residual, ReLU (scale / shift re-read per row: L1 hits; keeps the T batch in registers)
        float* orow = io.out + (long long)o * COUT + lane * V;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float sc = io.scale ? __ldg(io.scale + lane * V + v) : 1.f;
            const float sh = io.shift ? __ldg(io.shift + lane * V + v) : 0.f;
            float y = fmaf(acc[v], sc, sh);
            if (resid) y += __ldcg(resid + (long long)o * COUT + lane * V + v);
            if (relu) y = fmaxf(y, 0.f);
            acc[v] = y;
            amax = fmaxf(amax, fabsf(y));
        }
        if constexpr (V == 4) *reinterpret_cast<float4*>(orow) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
        else if constexpr (V == 2) *reinterpret_cast<float2*>(orow) = make_float2(acc[0], acc[1 % V]);
        else orow[0] = acc[0];
    }
    return warp_max(amax);
}

// ------------------------------------------------------------------------------------------ stem item
// Direct fused k3 conv Cin <= 8 -> 32 (k_stem_direct): warp per output row, lane = output channel, the 27 x Cin x 32
// weights of the problem in shared memory (Ws); the level-0 features come from the kernels before this launch.
__device__ __noinline__ float stem_rows(const LayerIO& io, long long seg_cap, int cin, const float* Ws, int r0, int r1,
                                           int warp, int lane, int nwarps) {
    const float sc = io.scale ? __ldg(io.scale + lane) : 1.f, sh = io.shift ? __ldg(io.shift + lane) : 0.f;
    float amax = 0.f;
    for (int o = r0 + warp; o < r1; o += nwarps) {
        int my_j = -1;
        if (lane < 27) {
            const int pos = __ldg(io.slot + (long long)lane * seg_cap + o);
            if (pos >= 0) my_j = __ldg(io.in_idx + (long long)lane * seg_cap + pos);
        }
        float acc = 0.f;
#pragma unroll 1
        for (int k = 0; k < 27; ++k) {
            const int jrow = __shfl_sync(0xffffffffu, my_j, k);
            if (jrow >= 0) {
                const float* f = io.fin + (long long)jrow * cin;
                const float* w = Ws + (k * cin) * 32 + lane;
                for (int ci = 0; ci < cin; ++ci) acc = fmaf(__ldg(f + ci), w[ci * 32], acc);
            }
        }
        float y = fmaf(acc, sc, sh);
        y = fmaxf(y, 0.f);
        io.out[(long long)o * 32 + lane] = y;
        amax = fmaxf(amax, fabsf(y));
    }
    return warp_max(amax);
}

// ------------------------------------------------------------------------------------------ the kernel
__device__ __forceinline__ int phase_layer(int p) { return (p + 1) >> 1; }         // 0 | 1 1 | 2 2 | ...
__device__ __forceinline__ bool phase_is_gemm(int p) { return p > 0 && (p & 1); }

__global__ void __launch_bounds__(N_THREADS, 2)
k_encoder_persist(const __grid_constant__ Program P) {
    extern __shared__ uint8_t smem_raw[];
    Ctx cx;
    const uint32_t raw = smem_u32(smem_raw);
    cx.base = (raw + 1023u) & ~1023u;
    cx.sm = smem_raw + (cx.base - raw);
    cx.tid = threadIdx.x; cx.lane = cx.tid & 31; cx.warp = cx.tid >> 5;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    uint8_t* sm = cx.sm;
    const uint32_t s_bar = cx.base + Smem::OFF_BAR;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + Smem::OFF_MISC);
    int* s_ticket = reinterpret_cast<int*>(sm + Smem::OFF_MISC + 8);
    int* s_item = reinterpret_cast<int*>(sm + Smem::OFF_MISC + 16);      // [3] problem, [4] r0, [5] r1 (row items)
    GemmArgs* s_gemm = reinterpret_cast<GemmArgs*>(sm + Smem::OFF_MISC + 48);
    Sched& S = *reinterpret_cast<Sched*>(sm + Smem::OFF_SCHED);
    unsigned* sync = P.sync;
    const int n_phases = 2 * P.n_layers - 1;

    auto init_barriers = [&]() {
        for (int s = 0; s < NS; ++s) { mbar_init(s_bar + 8u * s, WARPS_PER_GROUP); mbar_init(s_bar + 8u * (NS + s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(s_bar + 8u * (2 * NS + b), 1); mbar_init(s_bar + 8u * (2 * NS + 2 + b), 4); }
        mbar_init(s_bar + 8u * (2 * NS + 4), 1);
        mbar_init(s_bar + 8u * (2 * NS + 5), 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    };
    if (tid == 32) init_barriers();
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ir_pdl_wait();                         // rulebooks, counts and level-0 features come from the kernels before us
    ir_pdl_trigger();

    // ---- schedule tables: warp l computes layer l (counts are final)
    if (warp < P.n_layers) {
        const int l = warp, K = P.K[l], V = P.G * K;
        int total_items = 0;
        if (l >= 1) {
            int c[2], t[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int v = lane + 32 * h;
                c[h] = 0;
                if (v < V) c[h] = __ldg(P.io[l][v >= K ? 1 : 0].count + (v >= K ? v - K : v));
                t[h] = (c[h] + TILE_M - 1) / TILE_M;
            }
            int cinc[2] = {c[0], c[1]}, tinc[2] = {t[0], t[1]};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int c2 = __shfl_up_sync(0xffffffffu, cinc[h], o), t2 = __shfl_up_sync(0xffffffffu, tinc[h], o);
                    if (lane >= o) { cinc[h] += c2; tinc[h] += t2; }
                }
            }
            const int csum0 = __shfl_sync(0xffffffffu, cinc[0], 31), tsum0 = __shfl_sync(0xffffffffu, tinc[0], 31);
            cinc[1] += csum0; tinc[1] += tsum0;
            const int T_all = __shfl_sync(0xffffffffu, tinc[1], 31);
            const int cbase1 = __shfl_sync(0xffffffffu, cinc[0] - c[0], K & 31);       // pairs of problem 0 (K < 32)
            const int nonempty = __popc(__ballot_sync(0xffffffffu, t[0] > 0)) + __popc(__ballot_sync(0xffffffffu, t[1] > 0));
            const int spare = max(0, P.nominal - nonempty);
            int g[2], ginc[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                g[h] = 0;
                if (t[h] > 0) g[h] = min(t[h], 1 + (int)(((long long)spare * t[h]) / max(T_all, 1)));
                ginc[h] = g[h];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int g2 = __shfl_up_sync(0xffffffffu, ginc[h], o);
                    if (lane >= o) ginc[h] += g2;
                }
            }
            ginc[1] += __shfl_sync(0xffffffffu, ginc[0], 31);
            total_items = __shfl_sync(0xffffffffu, ginc[1], 31);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int v = lane + 32 * h;
                S.ginc[l][v] = ginc[h];
                S.tiles[l][v] = t[h];
                S.cnt[l][v] = c[h];
                S.kofs[l][v] = (cinc[h] - c[h]) - ((v >= K) ? cbase1 : 0);
            }
            if (lane == 0) S.items[2 * l - 1] = total_items;
        }
        // row items: one size for both problems, chosen so that their item counts together fit one ticket round
        int n = 0, rpi = RMIN, it = 0;
        if (lane < P.G) n = __ldg(P.io[l][lane].n_out_dev);
        const int n_all = n + __shfl_xor_sync(0xffffffffu, n, 1);              // lanes 0/1 hold the two row counts
        if (lane < P.G) {
            rpi = max(RMIN, (n_all + P.nominal - 3) / max(P.nominal - 2, 1));
            it = (n + rpi - 1) / rpi;
        }
        const int it0 = __shfl_sync(0xffffffffu, it, 0), it1 = __shfl_sync(0xffffffffu, it, 1);
        if (lane < 2) { S.rows[l][lane] = n; S.rpi[l][lane] = rpi; }
        if (lane == 0) { S.ri0[l] = it0; S.items[2 * l] = it0 + it1; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cx.tmem_base = *s_tmem;
    // Every phase hands out exactly `nominal` (= grid size) tickets; the ones beyond its real item count are null
    // items.  With one ticket taken per CTA per item this keeps ticket rounds and phases aligned: each CTA gets exactly
    // one (possibly null) item per phase, nobody runs two items of a phase while others idle at its barrier.
    if (warp == 0) {
        const int v = (lane < n_phases) ? max(P.nominal, S.items[lane]) : 0;
        S.tick[lane] = v;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        S.base[lane] = inc - v;                       // lanes >= n_phases hold the total
        if (lane == 0) *s_ticket = (int)atomicAdd(sync + SY_TICKET, 1u);
    }
    __syncthreads();
    const int total = S.base[n_phases];

    int phase = 0, known_done = -1, stem_loaded = 0;
    while (true) {
        const int ticket = *s_ticket;
        if (ticket >= total) break;
        while (ticket >= S.base[phase + 1]) ++phase;
        const int l = phase_layer(phase);
        const bool is_gemm = phase_is_gemm(phase);
        const int idx = ticket - S.base[phase];
        const bool null_item = idx >= S.items[phase];
        if (ticket == 0 && tid == 0) reinterpret_cast<unsigned long long*>(sync)[SY_STAMP64 + 31] = globaltimer();
        if (warp == 0 && !null_item) {
            if (is_gemm) {
                const int K = P.K[l];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int v = lane + 32 * h;
                    const int gi = S.ginc[l][v], gp = (v > 0) ? S.ginc[l][v - 1] : 0;
                    if (idx >= gp && idx < gi) {
                        const int g = gi - gp, r = idx - gp, t = S.tiles[l][v];
                        const int grp_id = (v >= K) ? 1 : 0;
                        const int kk = v - grp_id * K;
                        const LayerIO& o = P.io[l][grp_id];
                        s_item[3] = grp_id;
                        s_gemm->fin = o.fin;
                        s_gemm->idx_k = o.in_idx + (long long)kk * P.seg_cap[grp_id];
                        s_gemm->T = o.T;
                        s_gemm->w_src = reinterpret_cast<const uint8_t*>(o.weight) + (size_t)kk * P.cin[l] * P.cout[l] * 4;
                        s_gemm->in_absmax = sync + SY_ABSMAX + 2 * (l - 1) + grp_id;
                        s_gemm->base = cx.base;
                        s_gemm->tmem_base = cx.tmem_base;
                        s_gemm->t_begin = (int)(((long long)r * t) / g);
                        s_gemm->t_end = (int)(((long long)(r + 1) * t) / g);
                        s_gemm->kofs = S.kofs[l][v];
                        s_gemm->kcount = S.cnt[l][v];
                    }
                }
            } else if (lane == 0) {
                const int grp_id = (idx >= S.ri0[l]) ? 1 : 0;
                const int jj = idx - (grp_id ? S.ri0[l] : 0);
                const int r0 = jj * S.rpi[l][grp_id];
                s_item[3] = grp_id;
                s_item[4] = r0;
                s_item[5] = min(S.rows[l][grp_id], r0 + S.rpi[l][grp_id]);
            }
        }
        __syncthreads();
        cx.dbg = P.dbg ? P.dbg + (long long)ticket * 8 : nullptr;
        if (cx.dbg && tid == 0) { cx.dbg[0] = globaltimer(); cx.dbg[7] = ((unsigned long long)phase << 32) | blockIdx.x; }
        const int g = null_item ? 0 : s_item[3];
        const LayerIO& io = P.io[l][g];
        const long long seg_cap = P.seg_cap[g];
        // dependency of this item: every ticket of the previous phase (skipped when this CTA already saw it complete)
        const unsigned* dep = sync + SY_DONE + (phase > 0 ? phase - 1 : 0);
        const int need = (phase > 0 && known_done < phase - 1) ? S.tick[phase - 1] : -1;
        if (phase > 0 && !null_item) known_done = phase - 1;
        unsigned* amax_out = sync + SY_ABSMAX + 2 * l + g;
        if (null_item) {
            // nothing to do: the item only keeps the ticket rounds aligned with the phases
        } else if (is_gemm) {
            const int cin = P.cin[l], cout = P.cout[l];
            if (cin == 128)                   gemm_item<128, 128>(cx, s_gemm, dep, need);
            else if (cin == 64 && cout == 64) gemm_item<64, 64>(cx, s_gemm, dep, need);
            else if (cin == 64)               gemm_item<64, 128>(cx, s_gemm, dep, need);
            else                              gemm_item<32, 64>(cx, s_gemm, dep, need);
        } else if (phase == 0) {
            float* Ws = reinterpret_cast<float*>(sm + Smem::OFF_STAGE);
            if (!stem_loaded) {
                const int nw = 27 * P.cin[0] * 32;
                for (int gg = 0; gg < P.G; ++gg)
                    for (int i = tid; i < nw; i += N_THREADS) Ws[gg * STEM_FLOATS + i] = __ldg(P.io[0][gg].weight + i);
                stem_loaded = 1;
                __syncthreads();
            }
            const float m = stem_rows(io, seg_cap, P.cin[0], Ws + g * STEM_FLOATS, s_item[4], s_item[5], warp, lane, N_THREADS / 32);
            if (lane == 0 && m > 0.f) atomicMax(amax_out, __float_as_uint(m));
        } else {
            phase_wait(dep, need, tid);
            if (cx.dbg && tid == 0) cx.dbg[2] = globaltimer();
            const int K = P.K[l], cout = P.cout[l];
            const int* s_kofs = &S.kofs[l][g * K];
            const int relu = (l == P.n_layers - 1) ? P.relu_last : 1;
            float m;
            if (cout == 128)     m = reduce_rows<128>(io, seg_cap, K, s_kofs, s_item[4], s_item[5], relu, warp, lane, N_THREADS / 32);
            else if (cout == 64) m = reduce_rows<64>(io, seg_cap, K, s_kofs, s_item[4], s_item[5], relu, warp, lane, N_THREADS / 32);
            else                 m = reduce_rows<32>(io, seg_cap, K, s_kofs, s_item[4], s_item[5], relu, warp, lane, N_THREADS / 32);
            if (lane == 0 && m > 0.f) atomicMax(amax_out, __float_as_uint(m));
        }
        tc_fence_before();
        if (cx.dbg && tid == 0) cx.dbg[3] = globaltimer();
        // The next ticket is taken when this CTA is (almost) done with the current one — by the bookkeeping thread (the
        // MMA warp's lane 0, which finishes a GEMM item before the epilogue drains) — never earlier: tickets claimed
        // ahead of time by busy CTAs would leave late or slow CTAs without their share of a phase.
        if (warp == 4 && lane == 0) *s_ticket = (int)atomicAdd(sync + SY_TICKET, 1u);
        __syncthreads();                         // every global write of this item has been issued; next ticket published
        tc_fence_after();
        if (warp == 4 && lane == 0) {
            if (cx.dbg) cx.dbg[4] = globaltimer();
            // completion: release the item's writes, count it.  Only this thread waits for the fence / atomic round
            // trips; the other warps are already decoding the next item (the decode barrier orders the re-init below)
            __threadfence();
            const unsigned old = atomicAdd(sync + SY_DONE + phase, 1u);
            if ((int)old + 1 == S.tick[phase]) reinterpret_cast<unsigned long long*>(sync)[SY_STAMP64 + phase] = globaltimer();
            if (cx.dbg) cx.dbg[5] = globaltimer();
        }
        if (is_gemm && tid == 32) {              // fresh parities for the next GEMM item
            for (int b = 0; b < Smem::N_BAR; ++b)
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(s_bar + 8u * b) : "memory");
            init_barriers();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem_base), "r"(256u) : "memory");
    }
    if (tid == 0) {
        // the last CTA to leave clears the ticket / phase counters / range maxima for the next launch
        // (the per-phase time stamps stay readable until then)
        __threadfence();
        if (atomicAdd(sync + SY_EXIT, 1u) == gridDim.x - 1) {
            for (int i = 0; i < 2 * SY_STAMP64; ++i) sync[i] = 0u;
            __threadfence();
        }
    }
}

}  // namespace ep

extern int g_tune_pairgemm_ctas;
static unsigned long long* g_persist_dbg = nullptr;
// profiling aid: per-item time stamps of the next persistent launches into buf (u64 [n_tickets][8]) or NULL = off
// resident CTAs per SM of the persistent kernel as the runtime computes it (expected: 2)
extern "C" int ir_encoder_persist_occupancy(void) {
    int n = 0;
    cudaFuncSetAttribute(ep::k_encoder_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, ep::Smem::BYTES);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ep::k_encoder_persist, tc::N_THREADS, (size_t)ep::Smem::BYTES) != cudaSuccess) return -1;
    if (getenv("IR_VERBOSE")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, ep::k_encoder_persist);
        fprintf(stderr, "k_encoder_persist: regs %d static smem %zu max dyn smem %d local %zu maxThreads %d carveout %d -> %d CTAs/SM\n",
                fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock,
                fa.preferredShmemCarveout, n);
        for (int smem = 60000; smem <= 110000; smem += 10000) {
            int m = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, ep::k_encoder_persist, tc::N_THREADS, (size_t)smem);
            fprintf(stderr, "  dyn smem %d -> %d\n", smem, m);
        }
        for (int thr = 416; thr <= 544; thr += 32) {
            int m = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, ep::k_encoder_persist, thr, (size_t)60000);
            fprintf(stderr, "  threads %d (smem 60000) -> %d\n", thr, m);
        }
    }
    return n;
}
extern "C" int ir_encoder_persist_debug(uint64_t* buf) { g_persist_dbg = (unsigned long long*)buf; return IR_OK; }

// Builds the 13-layer program of encoder_features_multi (encoder.cu) and launches the persistent kernel.
int irk_encoder_persist(int G, const IrConvProblem (*layers)[IR_MAX_GROUPS], const int* cin, const int* cout, const int* K,
                        int n_layers, void* sync, cudaStream_t st) {
    IR_CHECK_ARG(G >= 1 && G <= 2 && n_layers >= 1 && n_layers <= ep::MAX_LAYERS && sync != nullptr);
    IR_CHECK_ARG(cin[0] >= 1 && cin[0] <= 8 && cout[0] == 32 && K[0] == 27);
    ep::Program P;
    memset(&P, 0, sizeof(P));
    for (int l = 0; l < n_layers; ++l) {
        P.cin[l] = cin[l]; P.cout[l] = cout[l]; P.K[l] = K[l];
        if (l >= 1) {
            const bool ok = (cin[l] == 32 && cout[l] == 64) || (cin[l] == 64 && cout[l] == 64) ||
                            (cin[l] == 64 && cout[l] == 128) || (cin[l] == 128 && cout[l] == 128);
            IR_CHECK_ARG(ok && K[l] <= 27);
        }
        for (int g = 0; g < G; ++g) {
            const IrConvProblem& p = layers[l][g];
            IR_CHECK_ARG(p.fin && p.in_idx && p.slot && p.count && p.n_out_dev && p.weight && p.T && p.out);
            IR_CHECK_ARG(l == 0 || (reinterpret_cast<uintptr_t>(p.weight) & 15) == 0);
            P.io[l][g] = ep::LayerIO{p.fin, p.in_idx, p.slot, p.count, p.n_out_dev, p.weight, p.scale, p.shift, p.resid, p.T, p.out};
            P.seg_cap[g] = p.seg_cap;
        }
    }
    P.n_layers = n_layers;
    P.G = G;
    P.relu_last = 1;
    P.sync = (unsigned*)sync;
    P.dbg = g_persist_dbg;
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(ep::k_encoder_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, ep::Smem::BYTES));
        attr_done = true;
    }
    const int grid = g_tune_pairgemm_ctas > 0 ? g_tune_pairgemm_ctas : 2 * IR_NUM_SMS;
    P.nominal = grid;
    IR_CHECK_CUDA(ir_launch_pdl(ep::k_encoder_persist, dim3(grid), dim3(tc::N_THREADS), (size_t)ep::Smem::BYTES, st, P));
    IR_CHECK_LAUNCH();
    return IR_OK;
}
