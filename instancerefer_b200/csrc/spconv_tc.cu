// tcgen05 pair-GEMM for the sparse convolution: the per-rule dense contraction
//     T[kofs[k] + pos, :] = F[in_idx[k][pos], :] @ W[k]
// on the 5th-gen tensor cores with fp32-class accuracy from a split-fp16 scheme:
//     x = x_hi + x_lo (two fp16, ~22 mantissa bits),  D += W_hi*x_hi + W_hi*x_lo + W_lo*x_hi  (fp32 accumulate)
// (weights pre-scaled by 2^8 so their lo parts stay normal; the epilogue multiplies by 2^-8).
//
//   * one persistent CTA per SM; every CTA serves ONE kernel offset k and a contiguous range of its
//     128-pair tiles (weight-stationary);
//   * W[k] is staged once by TMA bulk copies (cp.async.bulk -> UBLKCP) and parked in TENSOR MEMORY as
//     packed fp16 hi/lo (tcgen05.st): the MMA reads its weight operand from TMEM, so shared memory
//     only carries the gathered rows;
//   * 3 groups of 4 producer warps (one group per shared-memory stage) gather feature rows with
//     coalesced 16-byte loads, split them into fp16 hi/lo and write 128B-swizzled K-major stages
//     (64 channels per stage); a group's next loads are issued right after it hands a stage over;
//   * one thread issues tcgen05.mma kind::f16 (M = 128 output channels, zero-padded when Cout = 64;
//     N = 128 pairs; K = 16), accumulators double-buffered in TMEM;
//   * 4 epilogue warps drain TMEM with tcgen05.ld: TMEM lane = output channel, column = pair, so
//     every warp store instruction writes 32 consecutive channels of one T row (one 128-byte line).
// Validated against the SIMT fp32 kernel in spconv.cu and the CPU oracle.
#include <cuda_fp16.h>

#include "../../include/instancerefer_b200.h"
#include "common.cuh"
#include "kernels.cuh"

// tuning aid (tools/bench_spconv.py): bit0 skip gather loads, bit1 skip T stores, bit2 skip MMA issue
__device__ int g_tc_debug = 0;
extern "C" int ir_debug_set(int flags) {
    return cudaMemcpyToSymbol(g_tc_debug, &flags, sizeof(int)) == cudaSuccess ? IR_OK : IR_ERR_CUDA;
}

int g_tune_pairgemm_ctas = 2 * IR_NUM_SMS;      // tuning knobs (ir_tune_set): CTAs of a pair-GEMM launch ...
int g_tune_reduce_ctas = 5 * IR_NUM_SMS;        // ... and of a reduce / stem launch (swept in the step: 4-5 per SM beats 8)
extern "C" int ir_tune_set(int pairgemm_ctas, int reduce_ctas) {
    if (pairgemm_ctas > 0) g_tune_pairgemm_ctas = pairgemm_ctas;
    if (reduce_ctas > 0) g_tune_reduce_ctas = reduce_ctas;
    return IR_OK;
}

#include "tc_common.cuh"

namespace tc {

template <int CIN, int COUT>
struct Cfg {
    static constexpr int CINP = (CIN < PANEL) ? PANEL : CIN;     // channels padded to whole panels
    static constexpr int KP = CINP / PANEL;                      // K panels (= pipeline items) per tile
    static constexpr int W_RAW_BYTES = CIN * COUT * 4;           // W[k] (Cin,Cout) fp32, staged once by TMA
    static constexpr int OFF_STAGE = 0;
    static constexpr int STAGE_AREA = (NS * STAGE_BYTES > W_RAW_BYTES) ? NS * STAGE_BYTES : W_RAW_BYTES;
    static constexpr int OFF_BAR = OFF_STAGE + STAGE_AREA;
    static constexpr int N_BAR = 2 * NS + 6;
    static constexpr int OFF_MISC = OFF_BAR + N_BAR * 8;
    static constexpr int SMEM_BYTES = OFF_MISC + 16 + 16 * 4 + 1024;   // + alignment slack
    // TMEM columns: two fp32 accumulators (128 pair-columns each), then W^T hi and lo as packed fp16
    // pairs (Cin/2 columns each)
    static constexpr int COL_W_HI = 2 * TILE_M;
    static constexpr int COL_W_LO = COL_W_HI + CINP / 2;
    static constexpr int TMEM_COLS = (COL_W_LO + CINP / 2 <= 256) ? 256 : 512;
    static_assert(COL_W_LO + CINP / 2 <= 512, "TMEM column budget");
    static_assert(STAGE_AREA % 1024 == 0, "barriers follow the stage area");
};

// SCALED: the gathered rows are multiplied by a power of two derived from *in_absmax (IrConvProblem) before
// the fp16 split and the result is scaled back (training dgrad); the inference kernel carries no such code.
template <int CIN, int COUT, bool SCALED>
__global__ void __launch_bounds__(N_THREADS, 2)
k_pairgemm_tc(IrConvBatch batch, int K) {
    using C = Cfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t s_stage = base + C::OFF_STAGE;
    const uint32_t s_bar = base + C::OFF_BAR;
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (NS + s); };
    auto bar_tfull = [&](int b) { return s_bar + 8u * (2 * NS + b); };
    auto bar_tempty = [&](int b) { return s_bar + 8u * (2 * NS + 2 + b); };
    const uint32_t bar_wfull = s_bar + 8u * (2 * NS + 4);
    const uint32_t bar_wdone = s_bar + 8u * (2 * NS + 5);
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
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full(s), WARPS_PER_GROUP); mbar_init(bar_empty(s), 1); }
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

    ir_pdl_trigger();                      // the reduce kernel behind us may be scheduled as SMs free up
    const int dbg = g_tc_debug;
    const int kk = s_sched[0];
    const int t_begin = s_sched[1], t_end = (kk >= 0) ? s_sched[2] : 0;
    const int kofs = s_sched[3], kcount = s_sched[4];
    const IrConvProblem& P = batch.p[s_sched[5]];
    const float* __restrict__ F = P.fin;
    const int* __restrict__ in_idx = P.in_idx;
    const long long seg_cap = P.seg_cap;
    const float* __restrict__ weight = P.weight;
    float* __restrict__ T = P.T;
    const float* __restrict__ in_absmax = P.in_absmax;
    // input scale (power of two, exact): 2^13 / 2^floor(log2(max|F|)); 1 when no maximum is supplied
    auto input_scale = [&]() -> float {
        if (!SCALED || in_absmax == nullptr) return 1.f;
        const float m = *in_absmax;
        return (m > 0.f && m < 3.0e38f) ? exp2f(13.f - floorf(log2f(m))) : 1.f;
    };

    // W[k] (Cin,Cout) fp32 is staged in shared memory by one TMA bulk-copy chain (issued by thread 0);
    // 16 warps (4 per TMEM sub-partition) then scale it by 2^8, split it into fp16 hi/lo and park it in
    // TMEM as packed pairs, where it stays for the whole kernel: the MMA reads its weight operand from
    // TMEM, not from shared memory.  A warp's TMEM lanes are fixed by warp%4 (lanes = output channels);
    // `quarter` selects its share of the packed K columns.
    auto weights_to_tmem = [&](int quarter) {
        constexpr int NCOL = C::CINP / 2 / 4;                      // packed columns per warp (8 or 16)
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

    if (warp >= 5) {
        // ===================== gather producers: NS independent groups of 128 threads ==========
        // Work items = (tile, 64-channel panel) in order; group g owns stage g and the items
        // it = g, g+NS, ...  A thread owns TILE_M/16 (row, 8-channel chunk) tasks per item: two 16-byte loads,
        // fp32 -> fp16 hi/lo split, two 16-byte swizzled shared-memory stores.  The loads of the group's
        // next item are issued right after the current item was handed to the tensor core.
        const int pw = warp - 5;
        const int grp = pw / WARPS_PER_GROUP;
        const int gt = (pw % WARPS_PER_GROUP) * 32 + lane;     // thread inside the group (0..127)
        const int j = gt & 7;                  // 16-byte (8 x fp16) chunk inside the 128-byte panel row
        const int rbase = gt >> 3;             // rows rbase + 16*i, i < TASKS
        const int* idx_k = in_idx + (long long)kk * seg_cap;
        const int n_items = (t_end - t_begin) * C::KP;
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
            const int ch = (it % C::KP) * PANEL + j * 8;           // first channel of this thread's chunk
#pragma unroll
            for (int i = 0; i < TASKS; ++i) {
                va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx[i] >= 0 && ch < CIN && !(dbg & 1)) {
                    const float4* src = reinterpret_cast<const float4*>(F + (long long)idx[i] * CIN + ch);
                    va[i] = __ldg(src);
                    vb[i] = __ldg(src + 1);
                }
            }
        };
        if (grp < n_items) load_idx(grp);                       // the rulebook is final long before this launch
        if (t_end > t_begin) weights_to_tmem(1 + pw / 4);       // this warp's share of the weight columns
        ir_pdl_wait();                                          // feature rows come from the previous kernel
        const float in_s = input_scale();
        if (grp < n_items) {
            load_rows(grp);
            load_idx(grp + NS);
        }
        uint8_t* st_hi = sm + C::OFF_STAGE + grp * STAGE_BYTES;
        uint8_t* st_lo = st_hi + PANEL_BYTES;
        uint32_t round = 0;
        if (grp < n_items) mbar_wait(bar_wdone, 0u);            // raw weight tile has left the stage area
#pragma unroll 1
        for (int it = grp; it < n_items; it += NS, ++round) {
            mbar_wait(bar_empty(grp), (round & 1u) ^ 1u);
#pragma unroll
            for (int i = 0; i < TASKS; ++i) {
                const int r = rbase + 16 * i;
                const int off = r * 128 + ((j ^ (r & 7)) << 4);
                uint4 h, l;
                if (SCALED) {
                    va[i].x *= in_s; va[i].y *= in_s; va[i].z *= in_s; va[i].w *= in_s;
                    vb[i].x *= in_s; vb[i].y *= in_s; vb[i].z *= in_s; vb[i].w *= in_s;
                }
                split2(va[i].x, va[i].y, h.x, l.x);
                split2(va[i].z, va[i].w, h.y, l.y);
                split2(vb[i].x, vb[i].y, h.z, l.z);
                split2(vb[i].z, vb[i].w, h.w, l.w);
                *reinterpret_cast<uint4*>(st_hi + off) = h;
                *reinterpret_cast<uint4*>(st_lo + off) = l;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(grp));
            if (it + NS < n_items) load_rows(it + NS);            // idx[] holds the rows of item it+NS
            load_idx(it + 2 * NS);
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // fp32 accumulate, fp16 x fp16, both K-major, N = 128 pairs, M = 128 (padded) channels
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TILE_M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0, acc_it = 0;
            if (t_end > t_begin) {              // weights (hi, lo) are placed in TMEM by the epilogue warps
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
                    for (int ks = 0; ks < 4; ++ks) {            // 4 x K=16 per 64-channel panel
                        const uint64_t dg_hi = make_desc(g_hi + ks * 32), dg_lo = make_desc(g_lo + ks * 32);
                        if (dbg & 4) continue;
                        // A = W^T from TMEM (lanes = channels, 8 columns = 16 fp16), B = gathered pairs
                        mma_f16_ts(d_tmem, w_hi + ks * 8, dg_hi, idesc, (panel | ks) ? 1u : 0u);
                        mma_f16_ts(d_tmem, w_hi + ks * 8, dg_lo, idesc, 1u);
                        mma_f16_ts(d_tmem, w_lo + ks * 8, dg_hi, idesc, 1u);
                    }
                    tc_commit(bar_empty(stage));          // smem stage reusable once these MMAs retire
                }
                tc_commit(bar_tfull(b));                  // accumulator ready for the epilogue
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
            const int ch = warp * 32 + lane;                         // TMEM lane = output channel
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + b * TILE_M;
            if (warp * 32 < COUT) {
                float* tcol = T + (long long)(kofs + p0) * COUT + ch;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_M; c0 += 32) {
                    if (c0 >= np) break;                             // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_wait_ld();
                    if (!(dbg & 2)) {
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (c0 + q < np) tcol[(long long)(c0 + q) * COUT] = __uint_as_float(v[q]) * out_s;
                    }
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

template <int CIN, int COUT, bool SCALED = false>
int launch(const IrConvBatch& b, int K, cudaStream_t st) {
    using C = Cfg<CIN, COUT>;
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_pairgemm_tc<CIN, COUT, SCALED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    long long tiles_max = 0, rows_cap = 0;
    for (int g = 0; g < b.G; ++g) {
        tiles_max += (long long)K * b.p[g].n_max / TILE_M + K;
        rows_cap += b.p[g].n_max;
    }
    // default: two CTAs per SM; one per SM when the whole problem is capped at 8 k rows (small scenes: measured in the
    // 10 k-point x 8-instance sweep point, the narrower launches of the two encoders overlap instead of queueing)
    const int width = rows_cap <= 8192 ? (g_tune_pairgemm_ctas + 1) / 2 : g_tune_pairgemm_ctas;
    const int grid = ir_min_i(tiles_max > 0 ? tiles_max : 1, width);
    IR_CHECK_CUDA(ir_launch_pdl(k_pairgemm_tc<CIN, COUT, SCALED>, dim3(grid), dim3(N_THREADS), (size_t)C::SMEM_BYTES, st, b, K));
    IR_CHECK_LAUNCH();
    return IR_OK;
}


// =====================================================================================================
// wgrad on tcgen05:  dW[k] (Cin x Cout) = sum over the pairs of offset k of  X[in]^T (x) dY[out]
// (the second GEMM of the sparse-conv backward).  Both operands are GATHERED rows, i.e. matrices whose
// contraction index (the pair) is the slow one: A = X^T is (M = Cin) x (K = pairs) with M contiguous,
// B = dY is (K = pairs) x (N = Cout) with N contiguous — MN-major operands for UMMA.  The 128B-swizzled
// row image the forward producers write (one 128-byte row = 64 fp16 channels of one pair) IS the canonical
// MN-major SWIZZLE_128B layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) [fp16 elements]: a swizzle atom is
// 64 channels x 8 pairs, LBO = distance between 64-channel panels, SBO = 1024 B between 8-pair groups.
//   * grid (splits, K): a CTA owns a contiguous chunk of the pairs of ONE offset and ONE fp32 accumulator
//     (M = 128 lanes x N = Cout columns) in tensor memory for its whole life;
//   * 3 producer groups gather 64 pairs per stage: X rows -> fp16 hi/lo panels, dY rows (range-scaled by the
//     power of two from max|dY|, as in the scaled dgrad) -> fp16 hi/lo panels;
//   * one thread issues, per 16 pairs, hi*hi + hi*lo + lo*hi (SS-mode tcgen05.mma kind::f16, both MN-major);
//   * the epilogue drains TMEM once and adds the tile into dW[k] with fp32 atomics (splits share it).
template <int CIN, int COUT>
struct WgCfg {
    static constexpr int PA = 2;                       // A panels: M is always 128 (Cin = 64 -> second panel is zero)
    static constexpr int PA_REAL = CIN / 64;
    static constexpr int PB = COUT / 64;
    static constexpr int PANEL_B = 64 * 128;           // 64 pairs x 128 B
    static constexpr int OFF_XHI = 0, OFF_XLO = PA * PANEL_B;
    static constexpr int OFF_DHI = 2 * PA * PANEL_B, OFF_DLO = (2 * PA + PB) * PANEL_B;
    static constexpr int STAGE_B = 2 * (PA + PB) * PANEL_B;
    static constexpr int OFF_BAR = NS * STAGE_B;
    static constexpr int N_BAR = 2 * NS + 1;
    static constexpr int OFF_MISC = OFF_BAR + N_BAR * 8;
    static constexpr int SMEM_BYTES = OFF_MISC + 16 + 1024;
    static constexpr int TMEM_COLS = 128;
    static_assert(CIN % 64 == 0 && COUT % 64 == 0 && CIN <= 128 && COUT <= 128, "wgrad_tc shapes");
};

// MN-major, 128B swizzle: LBO = bytes between 64-element MN atoms, SBO = 1024 B between 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(N_THREADS, 1)
k_wgrad_tc(const float* __restrict__ X, const float* __restrict__ dY, const int* __restrict__ in_idx,
           const int* __restrict__ out_idx, const int* __restrict__ count, long long seg_cap,
           const float* __restrict__ dy_absmax, float* __restrict__ dW) {
    using C = WgCfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t s_bar = base + C::OFF_BAR;
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (NS + s); };
    const uint32_t bar_done = s_bar + 8u * (2 * NS);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + C::OFF_MISC);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int k = blockIdx.y;
    const int cnt = count[k];
    int chunk = (cnt + (int)gridDim.x - 1) / (int)gridDim.x;
    chunk = max(chunk, 8 * TILE_M);                                     // small offsets: fewer CTAs, fewer redundant tile adds
    chunk = (chunk + TILE_M - 1) / TILE_M * TILE_M;                     // whole stages per CTA
    const int p_begin = blockIdx.x * chunk, p_end = min(cnt, p_begin + chunk);
    const int n_items = p_end > p_begin ? (p_end - p_begin + TILE_M - 1) / TILE_M : 0;

    if (tid == 32) {
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full(s), WARPS_PER_GROUP); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(s_tmem)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (C::PA_REAL < C::PA) {                                           // Cin = 64: the upper half of M reads zeros
        for (int s = 0; s < NS; ++s)
            for (int i = tid; i < C::PANEL_B / 16; i += N_THREADS) {
                reinterpret_cast<uint4*>(sm + s * C::STAGE_B + C::OFF_XHI + C::PANEL_B)[i] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(sm + s * C::STAGE_B + C::OFF_XLO + C::PANEL_B)[i] = make_uint4(0, 0, 0, 0);
            }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const long long kbase = (long long)k * seg_cap;
    float dy_scale = 1.f;
    {
        const float m = dy_absmax ? *dy_absmax : 0.f;
        if (m > 0.f && m < 3.0e38f) dy_scale = exp2f(13.f - floorf(log2f(m)));
    }

    if (warp >= 5) {
        // ===================== gather producers (one group of 128 threads per stage) =====================
        const int pw = warp - 5;
        const int grp = pw / WARPS_PER_GROUP;
        const int gt = (pw % WARPS_PER_GROUP) * 32 + lane;
        const int j = gt & 7;                 // 16-byte chunk (8 channels) inside a 128-byte panel row
        const int rbase = gt >> 3;            // rows rbase + 16*i
        uint32_t round = 0;
        int nx[TASKS], no[TASKS];                                          // rulebook rows of the group's NEXT item
        auto load_idx = [&](int it) {
            const int p0 = p_begin + it * TILE_M;
#pragma unroll
            for (int i = 0; i < TASKS; ++i) {
                const int p = p0 + rbase + 16 * i;
                const bool ok = it < n_items && p < p_end;
                nx[i] = ok ? __ldg(in_idx + kbase + p) : -1;
                no[i] = ok ? __ldg(out_idx + kbase + p) : -1;
            }
        };
        load_idx(grp);
#pragma unroll 1
        for (int it = grp; it < n_items; it += NS, ++round) {
            int ix[TASKS], io[TASKS];
#pragma unroll
            for (int i = 0; i < TASKS; ++i) { ix[i] = nx[i]; io[i] = no[i]; }
            load_idx(it + NS);                                             // in flight while this item is gathered
            mbar_wait(bar_empty(grp), (round & 1u) ^ 1u);
            uint8_t* stg = sm + grp * C::STAGE_B;
#pragma unroll 1
            for (int pn = 0; pn < C::PA_REAL + C::PB; ++pn) {
                const bool isx = pn < C::PA_REAL;
                const int pp = isx ? pn : pn - C::PA_REAL;
                const float* src = isx ? X : dY;
                const int width = isx ? CIN : COUT;
                const float sc = isx ? 1.f : dy_scale;
                uint8_t* hi = stg + (isx ? C::OFF_XHI : C::OFF_DHI) + pp * C::PANEL_B;
                uint8_t* lo = stg + (isx ? C::OFF_XLO : C::OFF_DLO) + pp * C::PANEL_B;
                float4 va[TASKS], vb[TASKS];
#pragma unroll
                for (int i = 0; i < TASKS; ++i) {
                    const int row = isx ? ix[i] : io[i];
                    va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row >= 0) {
                        const float4* g = reinterpret_cast<const float4*>(src + (long long)row * width + pp * 64 + j * 8);
                        va[i] = __ldg(g);
                        vb[i] = __ldg(g + 1);
                    }
                }
#pragma unroll
                for (int i = 0; i < TASKS; ++i) {
                    const int r = rbase + 16 * i;
                    const int off = r * 128 + ((j ^ (r & 7)) << 4);
                    uint4 h, l;
                    split2(va[i].x * sc, va[i].y * sc, h.x, l.x);
                    split2(va[i].z * sc, va[i].w * sc, h.y, l.y);
                    split2(vb[i].x * sc, vb[i].y * sc, h.z, l.z);
                    split2(vb[i].z * sc, vb[i].w * sc, h.w, l.w);
                    *reinterpret_cast<uint4*>(hi + off) = h;
                    *reinterpret_cast<uint4*>(lo + off) = l;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(grp));
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (lane == 0 && n_items > 0) {
            // fp32 accumulate, fp16 x fp16, A and B MN-major, N = Cout, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int it = 0; it < n_items; ++it) {
                const int stage = it % NS;
                mbar_wait(bar_full(stage), (it / NS) & 1u);
                tc_fence_after();
                const uint32_t sb = base + stage * C::STAGE_B;
#pragma unroll
                for (int ks = 0; ks < TILE_M / 16; ++ks) {                 // 16 pairs (= 16 rows = 2048 B) per MMA
                    const uint64_t a_hi = make_desc_mn(sb + C::OFF_XHI + ks * 2048, C::PANEL_B);
                    const uint64_t a_lo = make_desc_mn(sb + C::OFF_XLO + ks * 2048, C::PANEL_B);
                    const uint64_t b_hi = make_desc_mn(sb + C::OFF_DHI + ks * 2048, C::PANEL_B);
                    const uint64_t b_lo = make_desc_mn(sb + C::OFF_DLO + ks * 2048, C::PANEL_B);
                    mma_f16_ss(tmem_base, a_hi, b_hi, idesc, (it | ks) ? 1u : 0u);
                    mma_f16_ss(tmem_base, a_hi, b_lo, idesc, 1u);
                    mma_f16_ss(tmem_base, a_lo, b_hi, idesc, 1u);
                }
                tc_commit(bar_empty(stage));
            }
            tc_commit(bar_done);
        }
    } else if (n_items > 0) {
        // ===================== epilogue: TMEM lane = ci, column = co =====================
        mbar_wait(bar_done, 0u);
        tc_fence_after();
        const int ci = warp * 32 + lane;
        if (warp * 32 < CIN) {
            const float unscale = 1.f / dy_scale;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            float* wrow = dW + ((long long)k * CIN + ci) * COUT;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 32; ++q) atomicAdd(wrow + c0 + q, __uint_as_float(v[q]) * unscale);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

template <int CIN, int COUT>
int launch_wgrad(const float* x, const float* dy, const int* in_idx, const int* out_idx, const int* count, long long seg_cap,
                 const float* dy_absmax, float* dW, int K, cudaStream_t st) {
    using C = WgCfg<CIN, COUT>;
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad_tc<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    const int nsplit = IR_NUM_SMS / K > 0 ? IR_NUM_SMS / K : 1;            // one CTA per SM, one wave
    k_wgrad_tc<CIN, COUT><<<dim3(nsplit, K), N_THREADS, C::SMEM_BYTES, st>>>(x, dy, in_idx, out_idx, count, seg_cap, dy_absmax, dW);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

}  // namespace tc

// 0: gathered rows by 16-byte global loads of the producer warps (k_pairgemm_tc); 1: by TMA tile::gather4 into a raw
// shared-memory stage (k_pairgemm_tma, spconv_tma.cu).  Forward shapes only; the dgrad shapes keep the LDG kernel.
int g_gather_mode = 0;
extern "C" int ir_gather_mode_set(int mode) {
    IR_CHECK_ARG(mode == 0 || mode == 1);
    g_gather_mode = mode;
    return IR_OK;
}

int irk_pairgemm_tc(const IrConvBatch& b, int cin, int cout, int K, cudaStream_t st) {
    IR_CHECK_ARG(K <= 27 && b.G >= 1 && b.G <= IR_MAX_GROUPS);
    if (g_gather_mode == 1 && ((cin == 32 && cout == 64) || (cin == 64 && cout == 64) || (cin == 64 && cout == 128) ||
                               (cin == 128 && cout == 128))) {
        bool aligned = true;
        for (int g = 0; g < b.G; ++g) aligned = aligned && (reinterpret_cast<uintptr_t>(b.p[g].fin) & 15) == 0;
        if (aligned) return irk_pairgemm_tma(b, cin, cout, K, st);
    }
    for (int g = 0; g < b.G; ++g) IR_CHECK_ARG(b.p[g].weight != nullptr && (reinterpret_cast<uintptr_t>(b.p[g].weight) & 15) == 0);
    bool scaled = true;
    for (int g = 0; g < b.G; ++g) scaled = scaled && b.p[g].in_absmax != nullptr;
    if (scaled) {          // range-scaled variant: the guarded inference path and the dgrad shapes of the training step
        if (cin == 32 && cout == 64) return tc::launch<32, 64, true>(b, K, st);
        if (cin == 64 && cout == 128) return tc::launch<64, 128, true>(b, K, st);
        if (cin == 64 && cout == 32) return tc::launch<64, 32, true>(b, K, st);
        if (cin == 64 && cout == 64) return tc::launch<64, 64, true>(b, K, st);
        if (cin == 128 && cout == 64) return tc::launch<128, 64, true>(b, K, st);
        if (cin == 128 && cout == 128) return tc::launch<128, 128, true>(b, K, st);
        ir_set_error("pairgemm_tc (scaled): unsupported channels %d -> %d", cin, cout);
        return IR_ERR_UNSUPPORTED;
    }
    if (cin == 32 && cout == 64) return tc::launch<32, 64>(b, K, st);
    if (cin == 64 && cout == 64) return tc::launch<64, 64>(b, K, st);
    if (cin == 64 && cout == 128) return tc::launch<64, 128>(b, K, st);
    if (cin == 128 && cout == 128) return tc::launch<128, 128>(b, K, st);
    ir_set_error("pairgemm_tc: unsupported channels %d -> %d", cin, cout);
    return IR_ERR_UNSUPPORTED;
}

// The tcgen05 kernel consumes the reference weight layout (K,Cin,Cout) directly (TMA bulk copy of
// W[k], tf32 hi/lo split on the way into TMEM): the "prepared" image is a 16-byte-aligned copy.
extern "C" int64_t ir_spconv_wprep_floats(int32_t K, int32_t cin, int32_t cout) {
    return (int64_t)K * cin * cout;
}

extern "C" int ir_spconv_prepare_weights(const float* weight, int32_t K, int32_t cin, int32_t cout, float* out,
                                         ir_stream_t stream) {
    IR_CHECK_ARG(weight && out && K > 0 && cin % 32 == 0 && cout % 32 == 0 && cout <= 128);
    IR_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    IR_CHECK_CUDA(cudaMemcpyAsync(out, weight, (size_t)K * cin * cout * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return IR_OK;
}

// wgrad on tcgen05 (see k_wgrad_tc); dW must be zeroed by the caller.  Shapes: Cin, Cout in {64, 128}.
int irk_wgrad_tc(const float* x, int cin, const float* dy, int cout, int K, const int* in_idx, const int* out_idx,
                 const int* count, long long seg_cap, const float* dy_absmax, float* dW, cudaStream_t st) {
    IR_CHECK_ARG(K > 0 && K <= 27);
    if (cin == 64 && cout == 64) return tc::launch_wgrad<64, 64>(x, dy, in_idx, out_idx, count, seg_cap, dy_absmax, dW, K, st);
    if (cin == 64 && cout == 128) return tc::launch_wgrad<64, 128>(x, dy, in_idx, out_idx, count, seg_cap, dy_absmax, dW, K, st);
    if (cin == 128 && cout == 128) return tc::launch_wgrad<128, 128>(x, dy, in_idx, out_idx, count, seg_cap, dy_absmax, dW, K, st);
    ir_set_error("wgrad_tc: unsupported channels %d -> %d", cin, cout);
    return IR_ERR_UNSUPPORTED;
}
