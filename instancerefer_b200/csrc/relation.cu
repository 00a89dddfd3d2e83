// Relation module kernels: per-instance feature rows, brute-force segmented kNN, fused EdgeConv
// (edge-weight MLP + message MLP + max over neighbours).
// Reference: models/relation_module.py:38-78,94-100; models/basic_blocks.py:98-133
// (torch_cluster.knn + PyG MessagePassing(aggr='max')).
#include "../../include/instancerefer_b200.h"
#include "common.cuh"

// ------------------------------------------------------------------ per-instance column means
__global__ void __launch_bounds__(256)
k_instance_mean(const float* __restrict__ pts, int ppi, int fdim, float* __restrict__ mean) {
    __shared__ float red[8][8];
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    const float* base = pts + (long long)inst * ppi * fdim;
    for (int p = tid; p < ppi; p += 256)
        for (int c = 0; c < fdim; ++c) acc[c] += base[(long long)p * fdim + c];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float v = warp_sum(acc[c]);
        if (lane == 0) red[w][c] = v;
    }
    __syncthreads();
    if (tid < fdim) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i][tid];
        mean[(long long)inst * fdim + tid] = s / (float)ppi;
    }
}

extern "C" int ir_instance_mean(const float* pts, int32_t n_inst, int32_t ppi, int32_t fdim,
                                float* mean, ir_stream_t stream) {
    IR_CHECK_ARG(pts && mean && n_inst > 0 && ppi > 0 && fdim > 0 && fdim <= 8);
    k_instance_mean<<<n_inst, 256, 0, (cudaStream_t)stream>>>(pts, ppi, fdim, mean);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ kNN inside scene segments
// one warp per query; k selection rounds, each picks the smallest (d, idx) strictly greater than the
// previous pick => ascending distance, ties -> lower index (torch_cluster's strict '>' insertion).
__global__ void __launch_bounds__(128)
k_knn(const float* __restrict__ xyz, const int* __restrict__ seg_ofs, const int* __restrict__ qidx,
      const int* __restrict__ qseg, int nq, int k, int* __restrict__ nbr) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int s0 = seg_ofs[qseg[q]], s1 = seg_ofs[qseg[q] + 1];
    const int qi = qidx[q];
    const float qx = xyz[3 * qi], qy = xyz[3 * qi + 1], qz = xyz[3 * qi + 2];
    float dprev = -1.f;
    int iprev = -1;
    for (int r = 0; r < k; ++r) {
        float bd = INFINITY;
        int bi = 0x7FFFFFFF;
        for (int i = s0 + lane; i < s1; i += 32) {
            const float dx = xyz[3 * i] - qx, dy = xyz[3 * i + 1] - qy, dz = xyz[3 * i + 2] - qz;
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const bool after = (d > dprev) || (d == dprev && i > iprev);
            if (after && (d < bd || (d == bd && i < bi))) { bd = d; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        const bool found = bi != 0x7FFFFFFF;
        if (lane == 0) nbr[(long long)q * k + r] = found ? bi : -1;
        if (!found) {
            for (int rr = r + 1; rr < k; ++rr)
                if (lane == 0) nbr[(long long)q * k + rr] = -1;
            break;
        }
        dprev = bd;
        iprev = bi;
    }
}

extern "C" int ir_knn(const float* xyz, const int32_t* seg_ofs, const int32_t* qidx,
                      const int32_t* qseg, int32_t nq, int32_t k, int32_t* nbr, ir_stream_t stream) {
    IR_CHECK_ARG(xyz && seg_ofs && qidx && qseg && nbr && nq > 0 && k > 0 && k <= 100);
    k_knn<<<ir_div_up(nq, 4), 128, 0, (cudaStream_t)stream>>>(xyz, seg_ofs, qidx, qseg, nq, k, nbr);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ fused EdgeConv
// Persistent CTAs (128 threads) loop over queries; the four weight matrices ((in,out) layout, 120 KB
// for F=25 / 18 classes / 128 outputs) are staged ONCE per CTA into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier) and every matvec then reads conflict-free shared memory.
#define EC_MAXK 16
#define EC_MAXF 32
#define EC_HW 64
#define EC_HM 128
__device__ __forceinline__ uint32_t ec_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128)
k_edgeconv(const float* __restrict__ x, const float* __restrict__ xyz, const int* __restrict__ qidx,
           const int* __restrict__ nbr, int nq, int k, int F, int ncls, const float* __restrict__ Ww1,
           const float* __restrict__ bw1, const float* __restrict__ Ww2, const float* __restrict__ bw2,
           const float* __restrict__ Wm1, const float* __restrict__ bm1, const float* __restrict__ Wm2,
           const float* __restrict__ bm2, int Fout, float* __restrict__ out) {
    extern __shared__ __align__(16) float ec_w[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_j[EC_MAXK];
    __shared__ float s_win[EC_MAXK][3 + 2 * EC_MAXF];
    __shared__ float s_hw[EC_MAXK][EC_HW];
    __shared__ float s_ein[EC_MAXK][3 * EC_MAXF];
    __shared__ float s_hm[EC_MAXK][EC_HM];
    const int tid = threadIdx.x;
    const int nw = 3 + 2 * ncls;     // edge-weight MLP input width
    const int ne = 3 * F;            // message MLP input width
    const int n1 = nw * EC_HW, n2 = EC_HW * F, n3 = ne * EC_HM, n4 = EC_HM * Fout;
    const int o2 = (n1 + 3) & ~3, o3 = (o2 + n2 + 3) & ~3, o4 = (o3 + n3 + 3) & ~3;     // 16-byte aligned sections
    float* sWw1 = ec_w;
    float* sWw2 = ec_w + o2;
    float* sWm1 = ec_w + o3;
    float* sWm2 = ec_w + o4;
    const uint32_t bar = ec_smem_u32(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)((n1 + n2 + n3 + n4) * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const float* srcs[4] = {Ww1, Ww2, Wm1, Wm2};
        float* dsts[4] = {sWw1, sWw2, sWm1, sWm2};
        const int ns[4] = {n1, n2, n3, n4};
        for (int m = 0; m < 4; ++m) {
            for (int o = 0; o < ns[m]; o += 4096) {                 // <= 16 KB per bulk copy
                const int c = min(4096, ns[m] - o);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ec_smem_u32(dsts[m] + o)), "l"(srcs[m] + o), "r"((uint32_t)(c * 4)), "r"(bar) : "memory");
            }
        }
    }
    __syncthreads();
    {   // wait for the weights (phase 0)
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar) : "memory");
        }
    }
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
        const int qi = qidx[q];
        __syncthreads();
        if (tid < k) s_j[tid] = nbr[(long long)q * k + tid];
        __syncthreads();
        // stage inputs
        for (int i = tid; i < k * nw; i += 128) {
            const int e = i / nw, c = i - e * nw;
            const int j = s_j[e];
            float v = 0.f;
            if (j >= 0) {
                if (c < 3) v = xyz[3 * j + c] - xyz[3 * qi + c];
                else if (c < 3 + ncls) v = x[(long long)qi * F + (F - ncls) + (c - 3)];
                else v = x[(long long)j * F + (F - ncls) + (c - 3 - ncls)];
            }
            s_win[e][c] = v;
        }
        for (int i = tid; i < k * F; i += 128) {
            const int e = i / F, c = i - e * F;
            const int j = s_j[e];
            s_ein[e][c] = x[(long long)qi * F + c];
            s_ein[e][2 * F + c] = (j >= 0) ? x[(long long)j * F + c] : 0.f;
        }
        __syncthreads();
        // edge-weight MLP layer 1: (nw -> 64), ReLU
        for (int i = tid; i < k * EC_HW; i += 128) {
            const int e = i / EC_HW, u = i - e * EC_HW;
            float a = bw1[u];
            for (int c = 0; c < nw; ++c) a = fmaf(sWw1[c * EC_HW + u], s_win[e][c], a);
            s_hw[e][u] = fmaxf(a, 0.f);
        }
        __syncthreads();
        // edge-weight MLP layer 2: (64 -> F)
        for (int i = tid; i < k * F; i += 128) {
            const int e = i / F, c = i - e * F;
            float a = bw2[c];
            for (int u = 0; u < EC_HW; ++u) a = fmaf(sWw2[u * F + c], s_hw[e][u], a);
            s_ein[e][F + c] = a;
        }
        __syncthreads();
        // message MLP layer 1: (3F -> 128), ReLU ; thread = hidden unit, all edges
        {
            const int u = tid;
            float a[EC_MAXK];
#pragma unroll
            for (int e = 0; e < EC_MAXK; ++e) a[e] = bm1[u];
            for (int c = 0; c < ne; ++c) {
                const float wv = sWm1[c * EC_HM + u];
#pragma unroll
                for (int e = 0; e < EC_MAXK; ++e)
                    if (e < k) a[e] = fmaf(wv, s_ein[e][c], a[e]);
            }
#pragma unroll
            for (int e = 0; e < EC_MAXK; ++e)
                if (e < k) s_hm[e][u] = fmaxf(a[e], 0.f);
        }
        __syncthreads();
        // message MLP layer 2: (128 -> Fout) + max over valid edges
        for (int u = tid; u < Fout; u += 128) {
            float a[EC_MAXK];
#pragma unroll
            for (int e = 0; e < EC_MAXK; ++e) a[e] = bm2[u];
            for (int v = 0; v < EC_HM; ++v) {
                const float wv = sWm2[v * Fout + u];
#pragma unroll
                for (int e = 0; e < EC_MAXK; ++e)
                    if (e < k) a[e] = fmaf(wv, s_hm[e][v], a[e]);
            }
            float m = -INFINITY;
            bool any = false;
#pragma unroll
            for (int e = 0; e < EC_MAXK; ++e)
                if (e < k && s_j[e] >= 0) { m = fmaxf(m, a[e]); any = true; }
            out[(long long)q * Fout + u] = any ? m : 0.f;
        }
    }
}

extern "C" int ir_edgeconv(const float* x, const float* xyz, const int32_t* qidx, const int32_t* nbr,
                           int32_t nq, int32_t k, int32_t F, int32_t ncls, const float* Ww1,
                           const float* bw1, const float* Ww2, const float* bw2, const float* Wm1,
                           const float* bm1, const float* Wm2, const float* bm2, int32_t Fout,
                           float* out, ir_stream_t stream) {
    IR_CHECK_ARG(x && xyz && qidx && nbr && Ww1 && bw1 && Ww2 && bw2 && Wm1 && bm1 && Wm2 && bm2 && out);
    IR_CHECK_ARG(nq > 0 && k > 0 && k <= EC_MAXK && F > 0 && F <= EC_MAXF && ncls > 0 && ncls <= F && Fout > 0);
    IR_CHECK_ARG(((uintptr_t)Ww1 & 15) == 0 && ((uintptr_t)Ww2 & 15) == 0 && ((uintptr_t)Wm1 & 15) == 0 && ((uintptr_t)Wm2 & 15) == 0);
    const int nw = 3 + 2 * ncls, ne = 3 * F;
    const int n1 = nw * EC_HW, n2 = EC_HW * F, n3 = ne * EC_HM, n4 = EC_HM * Fout;
    IR_CHECK_ARG(n1 % 4 == 0 && n2 % 4 == 0 && n3 % 4 == 0 && n4 % 4 == 0);        // 16-byte bulk copies
    const int o2 = (n1 + 3) & ~3, o3 = (o2 + n2 + 3) & ~3, o4 = (o3 + n3 + 3) & ~3;
    const size_t smem = (size_t)(o4 + n4) * 4;
    IR_CHECK_ARG(smem <= 180 * 1024);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        IR_CHECK_CUDA(cudaFuncSetAttribute(k_edgeconv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    k_edgeconv<<<ir_min_i(nq, IR_NUM_SMS), 128, smem, (cudaStream_t)stream>>>(x, xyz, qidx, nbr, nq, k, F, ncls, Ww1, bw1, Ww2, bw2,
                                                                            Wm1, bm1, Wm2, bm2, Fout, out);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
