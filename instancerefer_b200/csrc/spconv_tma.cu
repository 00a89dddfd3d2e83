// tcgen05 pair-GEMM with a TMA gather: the feature rows of a tile are fetched by the tensor memory accelerator
// (cp.async.bulk.tensor.2d tile::gather4 -> UTMALDG: four rulebook rows x 128 bytes per instruction, row index -1 =
// out of bounds = zero fill) into a raw fp32 stage in shared memory, already laid out in the 128-byte swizzle the
// converter reads back; the 12 converter warps only do shared -> fp16 hi/lo split -> shared, no global loads and no
// registers held across a memory latency.  Everything else (weight-stationary schedule, W^T parked in TMEM, TS-mode
// tcgen05.mma, double-buffered TMEM accumulators, coalesced T epilogue) is the kernel of spconv_tc.cu.
//
//   warp 17 (TMA warp): per (tile, 64-channel panel) item: 64 rulebook indices (2 per lane), one expect_tx, then every
//     lane issues ONE gather4 (row quad lane/2, 32-float half lane%2) -> a 16 KB raw stage in one warp instruction;
//   raw stage g belongs to converter group g (items g, g+3, ...): raw_full[g] (transaction barrier) / raw_empty[g].
// Replaces the same reference site as k_pairgemm_tc (torchsparse sparseconv_forward gather + GEMM, reached from
// models/basic_blocks.py:14,32,39).
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

extern int g_tune_pairgemm_ctas;      // spconv_tc.cu (ir_tune_set)

namespace tma {
using namespace tc;

constexpr int N_THREADS_TMA = N_THREADS + 32;          // + the TMA warp
constexpr int RAW_BYTES = TILE_M * PANEL * 4;          // one raw fp32 stage: 64 pairs x 64 channels
constexpr int RAW_HALF = TILE_M * 128;                 // 32-float (128-byte) column half of a raw stage

struct Maps { CUtensorMap m[IR_MAX_GROUPS]; };

template <int CIN, int COUT>
struct Cfg {
    static constexpr int CINP = (CIN < PANEL) ? PANEL : CIN;
    static constexpr int KP = CINP / PANEL;
    static constexpr int W_RAW_BYTES = CIN * COUT * 4;
    static constexpr int OFF_STAGE = 0;                            // hi/lo stages [0, 48 KB); W staging aliases [0, W_RAW)
    static constexpr int OFF_RAW = NS * STAGE_BYTES;               // raw ring [48 KB, 96 KB): raw[0] overlaps the W staging
    static constexpr int AREA = (OFF_RAW + NS * RAW_BYTES > W_RAW_BYTES) ? OFF_RAW + NS * RAW_BYTES : W_RAW_BYTES;
    static constexpr int OFF_BAR = AREA;
    static constexpr int N_BAR = 4 * NS + 6;
    static constexpr int OFF_MISC = OFF_BAR + N_BAR * 8;
    static constexpr int SMEM_BYTES = OFF_MISC + 16 + 16 * 4 + 1024;
    static constexpr int COL_W_HI = 2 * TILE_M;
    static constexpr int COL_W_LO = COL_W_HI + CINP / 2;
    static constexpr int TMEM_COLS = (COL_W_LO + CINP / 2 <= 256) ? 256 : 512;
    static constexpr bool RAW0_ALIASES_W = W_RAW_BYTES > OFF_RAW;  // the first use of raw[0] waits for the weights to leave
    static_assert(AREA % 1024 == 0 && OFF_RAW % 1024 == 0, "swizzle atoms are 1024-byte aligned");
    static_assert(W_RAW_BYTES <= OFF_RAW + RAW_BYTES, "W staging may only overlap raw[0]");
};

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

template <int CIN, int COUT, bool SCALED>
__global__ void __launch_bounds__(N_THREADS_TMA, 2)
k_pairgemm_tma(IrConvBatch batch, int K, const __grid_constant__ Maps maps) {
    using C = Cfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t s_stage = base + C::OFF_STAGE;
    const uint32_t s_rawst = base + C::OFF_RAW;
    const uint32_t s_bar = base + C::OFF_BAR;
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (NS + s); };
    auto bar_tfull = [&](int b) { return s_bar + 8u * (2 * NS + b); };
    auto bar_tempty = [&](int b) { return s_bar + 8u * (2 * NS + 2 + b); };
    const uint32_t bar_wfull = s_bar + 8u * (2 * NS + 4);
    const uint32_t bar_wdone = s_bar + 8u * (2 * NS + 5);
    auto bar_rfull = [&](int s) { return s_bar + 8u * (2 * NS + 6 + s); };
    auto bar_rempty = [&](int s) { return s_bar + 8u * (3 * NS + 6 + s); };
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + C::OFF_MISC);
    int* s_sched = reinterpret_cast<int*>(sm + C::OFF_MISC + 16);      // k, tile_begin, tile_end, kofs, count

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        // Pair / tile prefixes and the CTA <-> offset assignment.  Virtual offset v = g*K + k runs over
        // the (up to two) problems of this launch; lane handles v = lane and v = lane + 32.  Every CTA
        // serves ONE virtual offset (its weights are staged once); CTAs are dealt to offsets in
        // proportion to their tile counts, each offset with work gets at least one.
        const int V = batch.G * K;
        int c[2], t[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int v = lane + 32 * h;
            c[h] = 0;
            if (v < V) c[h] = __ldg(((v < K) ? batch.p[0].count : batch.p[1].count) + ((v < K) ? v : v - K));
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
        // exclusive pair prefix at the start of problem 1 (v == K): pairs of problem 0
        const int cbase1 = __shfl_sync(0xffffffffu, cinc[0] - c[0], K & 31);      // K < 32 always
        const int nonempty = __popc(__ballot_sync(0xffffffffu, t[0] > 0)) + __popc(__ballot_sync(0xffffffffu, t[1] > 0));
        const int spare = max(0, (int)gridDim.x - nonempty);
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
        if (lane == 0) { s_sched[0] = -1; s_sched[1] = 0; s_sched[2] = 0; s_sched[3] = 0; s_sched[4] = 0; s_sched[5] = 0; }
        __syncwarp();
        const int bx = (int)blockIdx.x;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cta_lo = ginc[h] - g[h];
            if (g[h] > 0 && bx >= cta_lo && bx < ginc[h]) {
                const int v = lane + 32 * h;
                const int grp_id = (v < K) ? 0 : 1;
                const int r = bx - cta_lo;
                s_sched[0] = v - grp_id * K;                                    // offset k served by this CTA
                s_sched[1] = (int)(((long long)r * t[h]) / g[h]);               // first tile (within k)
                s_sched[2] = (int)(((long long)(r + 1) * t[h]) / g[h]);         // end tile (within k)
                s_sched[3] = (cinc[h] - c[h]) - (grp_id ? cbase1 : 0);          // kofs[k] inside this problem's T
                s_sched[4] = c[h];                                              // pairs of this offset
                s_sched[5] = grp_id;                                            // which problem
            }
        }
    }
    if (tid == 32) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full(s), WARPS_PER_GROUP); mbar_init(bar_empty(s), 1);
            mbar_init(bar_rfull(s), 1); mbar_init(bar_rempty(s), WARPS_PER_GROUP);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 4); }
        mbar_init(bar_wfull, 1);
        mbar_init(bar_wdone, 16);          // weights resident in TMEM (one arrival per staging warp)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(s_tmem)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    ir_pdl_trigger();
    const int kk = s_sched[0];
    const int t_begin = s_sched[1], t_end = (kk >= 0) ? s_sched[2] : 0;
    const int kofs = s_sched[3], kcount = s_sched[4];
    const IrConvProblem& P = batch.p[s_sched[5]];
    const int* __restrict__ in_idx = P.in_idx;
    const long long seg_cap = P.seg_cap;
    const float* __restrict__ weight = P.weight;
    float* __restrict__ T = P.T;
    const float* __restrict__ in_absmax = P.in_absmax;
    const int n_items = (t_end - t_begin) * C::KP;
    auto input_scale = [&]() -> float {
        if (!SCALED || in_absmax == nullptr) return 1.f;
        const float m = *in_absmax;
        return (m > 0.f && m < 3.0e38f) ? exp2f(13.f - floorf(log2f(m))) : 1.f;
    };
    auto weights_to_tmem = [&](int quarter) {
        constexpr int NCOL = C::CINP / 2 / 4;
        mbar_wait(bar_wfull, 0u);
        const int sp = warp & 3;
        if (sp * 32 < COUT) {
            const float* ws = reinterpret_cast<const float*>(sm + C::OFF_STAGE) + sp * 32 + lane;
            const uint32_t tw = tmem_base + ((uint32_t)(sp * 32) << 16) + quarter * NCOL;
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
        if (lane == 0) mbar_arrive(bar_wdone);
    };

    if (warp == 17) {
        // ===================== TMA gather warp =====================
        const CUtensorMap* map = &maps.m[s_sched[5]];
        const int* idx_k = in_idx + (long long)kk * seg_cap;
        const int q = lane >> 1, h = lane & 1;                  // row quad, 32-float column half of this lane's gather4
        constexpr uint32_t HALVES = (CIN >= 64) ? 2u : 1u;      // Cin = 32: only the first half exists (the rest is zero-filled by the converters)
        // rulebook indices of an item: 2 per lane; prefetched two items ahead (the rulebook is final long before this launch)
        auto load_idx = [&](int it, int& ia, int& ib) {
            const int p0 = (t_begin + it / C::KP) * TILE_M;
            const int ra = p0 + lane, rb = p0 + 32 + lane;
            ia = (it < n_items && ra < kcount) ? __ldg(idx_k + ra) : -1;      // -1 = out of bounds = zero row
            ib = (it < n_items && rb < kcount) ? __ldg(idx_k + rb) : -1;
        };
        int ia0, ib0, ia1, ib1;
        load_idx(0, ia0, ib0);
        load_idx(1, ia1, ib1);
        ir_pdl_wait();                                          // the feature rows come from the previous kernel
#pragma unroll 1
        for (int it = 0; it < n_items; ++it) {
            const int g = it % NS;
            const int ia = ia0, ib = ib0;
            ia0 = ia1; ib0 = ib1;
            load_idx(it + 2, ia1, ib1);
            // raw stage of group g: slot (g+1)%NS, so the slot that overlaps the raw weight tile is the one used last
            const int slot = (g + 1) % NS;
            if (C::RAW0_ALIASES_W && it == NS - 1) mbar_wait(bar_wdone, 0u);
            mbar_wait(bar_rempty(g), (((uint32_t)(it / NS)) & 1u) ^ 1u);
            int r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = 4 * q + i;                                 // 0..63
                const int va = __shfl_sync(0xffffffffu, ia, row & 31), vb = __shfl_sync(0xffffffffu, ib, row & 31);
                r[i] = (row < 32) ? va : vb;
            }
            if (lane == 0) mbar_expect_tx(bar_rfull(g), HALVES * (uint32_t)RAW_HALF);
            __syncwarp();
            if ((uint32_t)h < HALVES)
                tma_gather4(s_rawst + slot * RAW_BYTES + h * RAW_HALF + q * 512, map, bar_rfull(g), (it % C::KP) * PANEL + h * 32,
                            r[0], r[1], r[2], r[3]);
        }
    } else if (warp >= 5) {
        // ===================== converters: NS groups of 128 threads, raw fp32 stage -> fp16 hi/lo stage ==========
        const int pw = warp - 5;
        const int grp = pw / WARPS_PER_GROUP;
        const int gt = (pw % WARPS_PER_GROUP) * 32 + lane;
        const int j = gt & 7;                  // 16-byte (8 x fp16) chunk of the 128-byte hi/lo row = channels 8j..8j+7
        const int rbase = gt >> 3;             // rows rbase + 16*i
        if (t_end > t_begin) weights_to_tmem(1 + pw / 4);       // this warp's share of the weight columns
        float in_s = 1.f;
        if (SCALED) { ir_pdl_wait(); in_s = input_scale(); }
        uint8_t* st_hi = sm + C::OFF_STAGE + grp * STAGE_BYTES;
        uint8_t* st_lo = st_hi + PANEL_BYTES;
        const uint8_t* rw = sm + C::OFF_RAW + ((grp + 1) % NS) * RAW_BYTES + (j >> 2) * RAW_HALF;     // this thread's column half
        const int c0 = 2 * (j & 3);            // first of its two 16-byte fp32 chunks inside the 128-byte half row
        uint32_t round = 0;
        if (grp < n_items) mbar_wait(bar_wdone, 0u);            // raw weight tile has left the stage area
#pragma unroll 1
        for (int it = grp; it < n_items; it += NS, ++round) {
            mbar_wait(bar_rfull(grp), round & 1u);
            float4 va[TASKS], vb[TASKS];
#pragma unroll
            for (int i = 0; i < TASKS; ++i) {
                const int r = rbase + 16 * i;
                if (CIN >= 64 || j < 4) {
                    va[i] = *reinterpret_cast<const float4*>(rw + r * 128 + (((c0) ^ (r & 7)) << 4));
                    vb[i] = *reinterpret_cast<const float4*>(rw + r * 128 + (((c0 + 1) ^ (r & 7)) << 4));
                } else {
                    va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rempty(grp));          // raw stage read: the TMA warp may refill it
            mbar_wait(bar_empty(grp), (round & 1u) ^ 1u);
#pragma unroll
            for (int i = 0; i < TASKS; ++i) {
                const int r = rbase + 16 * i;
                const int off = r * 128 + ((j ^ (r & 7)) << 4);
                uint4 hh, ll;
                if (SCALED) {
                    va[i].x *= in_s; va[i].y *= in_s; va[i].z *= in_s; va[i].w *= in_s;
                    vb[i].x *= in_s; vb[i].y *= in_s; vb[i].z *= in_s; vb[i].w *= in_s;
                }
                split2(va[i].x, va[i].y, hh.x, ll.x);
                split2(va[i].z, va[i].w, hh.y, ll.y);
                split2(vb[i].x, vb[i].y, hh.z, ll.z);
                split2(vb[i].z, vb[i].w, hh.w, ll.w);
                *reinterpret_cast<uint4*>(st_hi + off) = hh;
                *reinterpret_cast<uint4*>(st_lo + off) = ll;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(grp));
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TILE_M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0, acc_it = 0;
            if (t_end > t_begin) {
                mbar_wait(bar_wdone, 0u);
                tc_fence_after();
            }
            for (int tile = t_begin; tile < t_end; ++tile) {
                const uint32_t b = acc_it & 1u;
                mbar_wait(bar_tempty(b), ((acc_it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + b * TILE_M;
#pragma unroll 1
                for (int panel = 0; panel < C::KP; ++panel, ++it) {
                    const int stage = it % NS;
                    mbar_wait(bar_full(stage), (it / NS) & 1u);
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
                    tc_commit(bar_empty(stage));
                }
                tc_commit(bar_tfull(b));
                ++acc_it;
            }
        }
    } else {
        // ===================== epilogue (warps 0-3: TMEM lanes 32*warp ..) =====================
        if (t_end > t_begin) {
            if (tid == 0) {
                mbar_expect_tx(bar_wfull, (uint32_t)C::W_RAW_BYTES);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(weight) + (size_t)kk * C::W_RAW_BYTES;
                for (int o = 0; o < C::W_RAW_BYTES; o += 16384)
                    bulk_g2s(s_stage + o, src + o, (uint32_t)min(16384, C::W_RAW_BYTES - o), bar_wfull);
            }
            weights_to_tmem(0);
        }
        ir_pdl_wait();                                          // T is still being read by the previous reduce
        ir_stamp_begin(batch.stamp);
        const float out_s = SCALED ? W_UNSCALE / input_scale() : W_UNSCALE;
        uint32_t acc_it = 0;
        for (int tile = t_begin; tile < t_end; ++tile) {
            const int p0 = tile * TILE_M;
            const int np = min(TILE_M, kcount - p0);
            const uint32_t b = acc_it & 1u;
            mbar_wait(bar_tfull(b), (acc_it >> 1) & 1u);
            tc_fence_after();
            const int ch = warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + b * TILE_M;
            if (warp * 32 < COUT) {
                float* tcol = T + (long long)(kofs + p0) * COUT + ch;
#pragma unroll 1
                for (int c1 = 0; c1 < TILE_M; c1 += 32) {
                    if (c1 >= np) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + c1, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int qq = 0; qq < 32; ++qq)
                        if (c1 + qq < np) tcol[(long long)(c1 + qq) * COUT] = __uint_as_float(v[qq]) * out_s;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(b));
            ++acc_it;
        }
    }
    tc_fence_before();
    __syncthreads();
    ir_stamp_end(batch.stamp);
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ---- host: tensor maps (fp32 (rows, Cin), box = 32 floats x 1 row, 128-byte swizzle, zero fill out of bounds)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}
static int make_map(CUtensorMap* m, const float* base, long long rows, int cin) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { ir_set_error("cuTensorMapEncodeTiled unavailable"); return IR_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cin * 4};
    cuuint32_t box[2] = {32, 1};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ir_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return IR_ERR_CUDA; }
    return IR_OK;
}

template <int CIN, int COUT, bool SCALED>
static int launch(const IrConvBatch& b, int K, const Maps& maps, cudaStream_t st) {
    using C = Cfg<CIN, COUT>;
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_pairgemm_tma<CIN, COUT, SCALED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    long long tiles_max = 0;
    for (int g = 0; g < b.G; ++g) tiles_max += (long long)K * b.p[g].n_max / TILE_M + K;
    const int grid = ir_min_i(tiles_max > 0 ? tiles_max : 1, g_tune_pairgemm_ctas);
    IR_CHECK_CUDA(ir_launch_pdl(k_pairgemm_tma<CIN, COUT, SCALED>, dim3(grid), dim3(N_THREADS_TMA), (size_t)C::SMEM_BYTES, st, b, K, maps));
    IR_CHECK_LAUNCH();
    return IR_OK;
}

}  // namespace tma

// rows_in[g]: number of rows of p[g].fin the tensor map may address (capacity of the buffer)
int irk_pairgemm_tma(const IrConvBatch& b, int cin, int cout, int K, cudaStream_t st) {
    IR_CHECK_ARG(K <= 27 && b.G >= 1 && b.G <= IR_MAX_GROUPS);
    tma::Maps maps;
    memset(&maps, 0, sizeof(maps));
    bool scaled = true;
    for (int g = 0; g < b.G; ++g) {
        IR_CHECK_ARG(b.p[g].weight != nullptr && (reinterpret_cast<uintptr_t>(b.p[g].weight) & 15) == 0);
        IR_CHECK_ARG((reinterpret_cast<uintptr_t>(b.p[g].fin) & 15) == 0);
        int r = tma::make_map(&maps.m[g], b.p[g].fin, b.p[g].n_max, cin);
        if (r != IR_OK) return r;
        scaled = scaled && b.p[g].in_absmax != nullptr;
    }
    if (b.G == 1) maps.m[1] = maps.m[0];
    if (scaled) {
        if (cin == 32 && cout == 64) return tma::launch<32, 64, true>(b, K, maps, st);
        if (cin == 64 && cout == 64) return tma::launch<64, 64, true>(b, K, maps, st);
        if (cin == 64 && cout == 128) return tma::launch<64, 128, true>(b, K, maps, st);
        if (cin == 128 && cout == 128) return tma::launch<128, 128, true>(b, K, maps, st);
    } else {
        if (cin == 32 && cout == 64) return tma::launch<32, 64, false>(b, K, maps, st);
        if (cin == 64 && cout == 64) return tma::launch<64, 64, false>(b, K, maps, st);
        if (cin == 64 && cout == 128) return tma::launch<64, 128, false>(b, K, maps, st);
        if (cin == 128 && cout == 128) return tma::launch<128, 128, false>(b, K, maps, st);
    }
    ir_set_error("pairgemm_tma: unsupported channels %d -> %d", cin, cout);
    return IR_ERR_UNSUPPORTED;
}
