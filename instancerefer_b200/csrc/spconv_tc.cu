// tcgen05 pair-GEMM for the sparse convolution: the per-rule dense contraction
//     T[kofs[k] + pos, :] = F[in_idx[k][pos], :] @ W[k]
// on the 5th-gen tensor cores with fp32-accurate 3xTF32 split accumulation.
//
//   * one persistent CTA per SM, a contiguous chunk of 128-pair tiles (pairs are grouped by kernel
//     offset k, so a CTA keeps W[k] resident: weight-stationary);
//   * W[k]^T (hi and lo tf32 parts) is staged once per k change by TMA bulk copies
//     (cp.async.bulk -> UBLKCP) from a pre-swizzled image (ir_spconv_prepare_weights);
//   * 8 producer warps gather feature rows with coalesced 16-byte loads, split them into tf32
//     hi/lo parts and write them into 128B-swizzled K-major stages (32 channels per stage);
//   * one thread issues tcgen05.mma kind::tf32 (M=128 pairs, N=Cout, K=8): hi*hi + lo*hi + hi*lo,
//     accumulators double-buffered in TMEM;
//   * 4 epilogue warps drain TMEM with tcgen05.ld and store T rows with 16-byte vector stores.
// D lane = pair, D column = output channel.  Validated against the SIMT kernel in spconv.cu.
#include "../../include/instancerefer_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace tc {

constexpr int TILE_M = 128;          // pairs per tile (UMMA M)
constexpr int PANEL = 32;            // fp32 channels per 128-byte swizzle row
constexpr int PANEL_BYTES = TILE_M * 128;          // one operand panel of the gathered tile
constexpr int STAGE_BYTES = 2 * PANEL_BYTES;       // hi + lo
constexpr int N_PRODUCER_WARPS = 8;
constexpr int N_THREADS = (4 + 1 + N_PRODUCER_WARPS) * 32;   // epilogue x4, mma x1, producers x8

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, 128B swizzle, 8-row groups 1024 B apart (SBO), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int CIN, int COUT, int NS>
struct Cfg {
    static constexpr int KP = CIN / PANEL;                       // K panels per tile
    static constexpr int W_PANEL_BYTES = COUT * 128;             // one weight panel (hi or lo)
    static constexpr int W_HALF_BYTES = KP * W_PANEL_BYTES;      // hi (or lo) image of one offset
    static constexpr int W_BYTES = 2 * W_HALF_BYTES;
    static constexpr int OFF_STAGE = W_BYTES;
    static constexpr int OFF_BAR = OFF_STAGE + NS * STAGE_BYTES;
    static constexpr int N_BAR = 2 * NS + 6;
    static constexpr int OFF_MISC = OFF_BAR + N_BAR * 8;
    static constexpr int SMEM_BYTES = OFF_MISC + 16 + 2 * 33 * 4 + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = (2 * COUT <= 32) ? 32 : (2 * COUT <= 64) ? 64 : (2 * COUT <= 128) ? 128 : 256;
};

template <int CIN, int COUT, int NS>
__global__ void __launch_bounds__(N_THREADS, 1)
k_pairgemm_tc(const float* __restrict__ F, int K, const int* __restrict__ in_idx, long long seg_cap,
              const int* __restrict__ count, const float* __restrict__ wprep, float* __restrict__ T) {
    using C = Cfg<CIN, COUT, NS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t s_w = base;
    const uint32_t s_stage = base + C::OFF_STAGE;
    const uint32_t s_bar = base + C::OFF_BAR;
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (NS + s); };
    auto bar_tfull = [&](int b) { return s_bar + 8u * (2 * NS + b); };
    auto bar_tempty = [&](int b) { return s_bar + 8u * (2 * NS + 2 + b); };
    const uint32_t bar_wfull = s_bar + 8u * (2 * NS + 4);
    const uint32_t bar_wdone = s_bar + 8u * (2 * NS + 5);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + C::OFF_MISC);
    int* s_kofs = reinterpret_cast<int*>(sm + C::OFF_MISC + 16);
    int* s_tofs = s_kofs + 33;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        int a = 0, t = 0;
        for (int k = 0; k < K; ++k) {
            s_kofs[k] = a; s_tofs[k] = t;
            const int c = count[k];
            a += c; t += (c + TILE_M - 1) / TILE_M;
        }
        s_kofs[K] = a; s_tofs[K] = t;
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full(s), N_PRODUCER_WARPS); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 4); }
        mbar_init(bar_wfull, 1);
        mbar_init(bar_wdone, 1);
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

    const int ntiles = s_tofs[K];
    const int t_begin = (int)(((long long)blockIdx.x * ntiles) / gridDim.x);
    const int t_end = (int)(((long long)(blockIdx.x + 1) * ntiles) / gridDim.x);

    if (warp >= 5) {
        // ===================== gather producers (256 threads) =====================
        const int pt = tid - 5 * 32;
        const int j = pt & 7;                 // 16-byte chunk inside the 128-byte panel row
        const int rbase = pt >> 3;            // rows rbase + 32*i
        uint32_t it = 0;
        int kk = 0;
        for (int tile = t_begin; tile < t_end; ++tile) {
            while (tile >= s_tofs[kk + 1]) ++kk;
            const int p0 = (tile - s_tofs[kk]) * TILE_M;
            const int np = min(TILE_M, (s_kofs[kk + 1] - s_kofs[kk]) - p0);
            int idx[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rbase + 32 * i;
                idx[i] = (r < np) ? __ldg(in_idx + (long long)kk * seg_cap + p0 + r) : -1;
            }
#pragma unroll 1
            for (int panel = 0; panel < C::KP; ++panel, ++it) {
                const int stage = it % NS;
                mbar_wait(bar_empty(stage), ((it / NS) & 1u) ^ 1u);
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx[i] >= 0)
                        v[i] = __ldg(reinterpret_cast<const float4*>(F + (long long)idx[i] * CIN + panel * PANEL + j * 4));
                }
                uint8_t* st_hi = sm + C::OFF_STAGE + stage * STAGE_BYTES;
                uint8_t* st_lo = st_hi + PANEL_BYTES;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = rbase + 32 * i;
                    const int off = r * 128 + ((j ^ (r & 7)) << 4);
                    float4 h, l;
                    h.x = tf32_hi(v[i].x); l.x = tf32_hi(v[i].x - h.x);
                    h.y = tf32_hi(v[i].y); l.y = tf32_hi(v[i].y - h.y);
                    h.z = tf32_hi(v[i].z); l.z = tf32_hi(v[i].z - h.z);
                    h.w = tf32_hi(v[i].w); l.w = tf32_hi(v[i].w - h.w);
                    *reinterpret_cast<float4*>(st_hi + off) = h;
                    *reinterpret_cast<float4*>(st_lo + off) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full(stage));
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                       ((uint32_t)(TILE_M >> 4) << 24);
            uint32_t it = 0, acc_it = 0, wphase = 0, dphase = 0;
            int kk = 0, cur_k = -1;
            for (int tile = t_begin; tile < t_end; ++tile) {
                while (tile >= s_tofs[kk + 1]) ++kk;
                if (kk != cur_k) {
                    if (cur_k >= 0) {            // all MMAs reading the old weights must be done
                        tc_commit(bar_wdone);
                        mbar_wait(bar_wdone, dphase);
                        dphase ^= 1u;
                    }
                    mbar_expect_tx(bar_wfull, (uint32_t)C::W_BYTES);
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(wprep) + (size_t)kk * C::W_BYTES;
                    for (int o = 0; o < C::W_BYTES; o += 16384)
                        bulk_g2s(s_w + o, src + o, (uint32_t)min(16384, C::W_BYTES - o), bar_wfull);
                    mbar_wait(bar_wfull, wphase);
                    wphase ^= 1u;
                    cur_k = kk;
                }
                const uint32_t b = acc_it & 1u;
                mbar_wait(bar_tempty(b), ((acc_it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + b * COUT;
#pragma unroll 1
                for (int panel = 0; panel < C::KP; ++panel, ++it) {
                    const int stage = it % NS;
                    mbar_wait(bar_full(stage), (it / NS) & 1u);
                    tc_fence_after();
                    const uint32_t a_hi = s_stage + stage * STAGE_BYTES;
                    const uint32_t a_lo = a_hi + PANEL_BYTES;
                    const uint32_t w_hi = s_w + panel * C::W_PANEL_BYTES;
                    const uint32_t w_lo = w_hi + C::W_HALF_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t da_hi = make_desc(a_hi + ks * 32), da_lo = make_desc(a_lo + ks * 32);
                        const uint64_t db_hi = make_desc(w_hi + ks * 32), db_lo = make_desc(w_lo + ks * 32);
                        mma_tf32(d_tmem, da_hi, db_hi, idesc, (panel | ks) ? 1u : 0u);
                        mma_tf32(d_tmem, da_lo, db_hi, idesc, 1u);
                        mma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                    tc_commit(bar_empty(stage));          // smem stage reusable once these MMAs retire
                }
                tc_commit(bar_tfull(b));                  // accumulator ready for the epilogue
                ++acc_it;
            }
        }
    } else {
        // ===================== epilogue (warps 0-3: TMEM lanes 32*warp ..) =====================
        uint32_t acc_it = 0;
        int kk = 0;
        for (int tile = t_begin; tile < t_end; ++tile) {
            while (tile >= s_tofs[kk + 1]) ++kk;
            const int p0 = (tile - s_tofs[kk]) * TILE_M;
            const int np = min(TILE_M, (s_kofs[kk + 1] - s_kofs[kk]) - p0);
            const uint32_t b = acc_it & 1u;
            mbar_wait(bar_tfull(b), (acc_it >> 1) & 1u);
            tc_fence_after();
            const int row = warp * 32 + lane;
            float* trow = T + (long long)(s_kofs[kk] + p0 + row) * COUT;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + b * COUT;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_wait_ld();
                if (row < np) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<uint4*>(trow + c0 + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
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
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// weight image: per offset k: hi[KP][COUT][32 floats, 16B chunks XOR-swizzled by row&7], then lo
__global__ void k_prepare_weights(const float* __restrict__ W, int K, int cin, int cout, float* __restrict__ out) {
    const long long total = (long long)K * cin * cout;
    const int kp = cin / PANEL;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % cout);
        const int c = (int)((i / cout) % cin);
        const int k = (int)(i / ((long long)cout * cin));
        const float w = W[i];                      // W[k][c][n]
        const float hi = tf32_hi(w);
        const float lo = tf32_hi(w - hi);
        const int p = c / PANEL, jj = (c % PANEL) / 4, e = c % 4;
        const long long half = (long long)kp * cout * PANEL;
        const long long o = (long long)k * 2 * half + (long long)p * cout * PANEL + (long long)n * PANEL + ((jj ^ (n & 7)) * 4) + e;
        out[o] = hi;
        out[o + half] = lo;
    }
}

template <int CIN, int COUT, int NS>
int launch(const float* F, int K, const int* in_idx, long long seg_cap, const int* count, const float* wprep,
           float* T, long long pairs_max, cudaStream_t st) {
    using C = Cfg<CIN, COUT, NS>;
    static bool attr_done = false;
    if (!attr_done) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_pairgemm_tc<CIN, COUT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    const long long tiles_max = pairs_max / TILE_M + K;
    const int grid = ir_min_i(tiles_max > 0 ? tiles_max : 1, IR_NUM_SMS);
    k_pairgemm_tc<CIN, COUT, NS><<<grid, N_THREADS, C::SMEM_BYTES, st>>>(F, K, in_idx, seg_cap, count, wprep, T);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

}  // namespace tc

int irk_pairgemm_tc(const float* feat_in, int cin, int cout, int K, const int* in_idx, long long seg_cap,
                    const int* count, const float* wprep, float* T, long long pairs_max, cudaStream_t st) {
    IR_CHECK_ARG(K <= 32 && wprep != nullptr);
    if (cin == 32 && cout == 64) return tc::launch<32, 64, 4>(feat_in, K, in_idx, seg_cap, count, wprep, T, pairs_max, st);
    if (cin == 64 && cout == 64) return tc::launch<64, 64, 4>(feat_in, K, in_idx, seg_cap, count, wprep, T, pairs_max, st);
    if (cin == 64 && cout == 128) return tc::launch<64, 128, 4>(feat_in, K, in_idx, seg_cap, count, wprep, T, pairs_max, st);
    if (cin == 128 && cout == 128) return tc::launch<128, 128, 3>(feat_in, K, in_idx, seg_cap, count, wprep, T, pairs_max, st);
    ir_set_error("pairgemm_tc: unsupported channels %d -> %d", cin, cout);
    return IR_ERR_UNSUPPORTED;
}

extern "C" int64_t ir_spconv_wprep_floats(int32_t K, int32_t cin, int32_t cout) {
    return (int64_t)K * 2 * cin * cout;
}

extern "C" int ir_spconv_prepare_weights(const float* weight, int32_t K, int32_t cin, int32_t cout, float* out,
                                         ir_stream_t stream) {
    IR_CHECK_ARG(weight && out && K > 0 && cin % 32 == 0 && cout % 8 == 0);
    const long long total = (long long)K * cin * cout;
    tc::k_prepare_weights<<<ir_min_i(ir_div_up(total, 256), IR_NUM_SMS * 8), 256, 0, (cudaStream_t)stream>>>(
        weight, K, cin, cout, out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
