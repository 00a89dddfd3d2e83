// Internal launcher prototypes shared between the .cu files (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

// coords.cu
struct IrLevels { IrTable t[5]; };      // hash tables of levels 0..4 (stride 1,2,4,8,16)
struct IrKmapArgs {
    const int4* coords[5];
    const int* nlvl;
    IrLevels lt;
    int* k3_in[5];
    int* k3_slot[5];
    int* k2_in[4];
    int* k2_slot[4];
    int* kcount;      // [9][32]
    long long n_max;
};
int irk_levels_from_coords(const int32_t* coords0, int n0, const int* n0_dev, IrLevels lt, int* pslot,
                           long long n_max, int* nlvl, cudaStream_t st);
int irk_levels_compact(const int32_t* coords0, long long n0_max, int* nlvl, IrLevels lt, const int* pslot,
                       long long n_max, int32_t* c1, int32_t* c2, int32_t* c3, int32_t* c4,
                       unsigned long long* scan_state, long long scan_stride, cudaStream_t st);
int irk_voxelize(const float* pts, const int* cand, int n_cand, int ppi, int fdim, double voxel,
                 IrLevels lt, int* vslot, int32_t* coords_out, float* feats_out, int* nlvl,
                 unsigned long long* scan_state, int* pslot, long long n_max, cudaStream_t st);
int irk_kmap_all(const IrKmapArgs& a, long long rows_max, cudaStream_t st);

// A sparse-conv launch serves up to two independent problems with identical layer shapes (the
// instance encoder and the scene encoder run the same 13 layers): one launch, CTAs dealt to both.
#define IR_MAX_GROUPS 2
struct IrConvProblem {
    const float* fin;        // (rows_in, Cin)
    const int* in_idx;       // [K][seg_cap]
    const int* slot;         // [K][seg_cap]
    const int* count;        // [K]
    const int* n_out_dev;    // device row count of the output level
    const float* weight;     // (K,Cin,Cout) fp32 (16-byte aligned)
    const float* scale;      // folded BN (or NULL)
    const float* shift;
    const float* resid;      // (rows_out, Cout) or NULL
    float* T;                // pair products
    float* out;              // (rows_out, Cout)
    long long seg_cap;
    long long n_max;
    int relu;
    const float* in_absmax;  // tcgen05 path only: device scalar max|fin| (or NULL).  The gathered rows are scaled by
                             // the power of two that brings this maximum to 2^13 before the fp16 hi/lo split and the
                             // result is scaled back exactly: keeps small-magnitude inputs (gradients) out of fp16's
                             // subnormal range.
    float* out_absmax;       // reduce / stem epilogue: device scalar receiving max|out| of this launch (atomicMax on the
                             // bit pattern; the caller zeroes it), or NULL.  Feeds the next layer's in_absmax.
};
struct IrConvBatch {
    IrConvProblem p[IR_MAX_GROUPS];
    int G;
    unsigned long long* stamp;   // optional {first CTA past its dependency wait, last CTA done} in GPU-timer ns (atomicMin /
                                 // atomicMax; the caller presets {~0, 0}), or NULL: live per-launch spans inside a replayed graph
};
__device__ __forceinline__ void ir_stamp_begin(unsigned long long* stamp) {
    if (stamp && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMin(stamp, t);
    }
}
__device__ __forceinline__ void ir_stamp_end(unsigned long long* stamp) {
    if (stamp && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMax(stamp + 1, t);
    }
}

// spconv.cu  (SIMT fp32 pair-GEMM + deterministic reduce/epilogue)
int irk_pairgemm_simt(const IrConvBatch& b, int cin, int cout, int K, cudaStream_t st);
int irk_reduce_epilogue(const IrConvBatch& b, int cout, int K, cudaStream_t st);
int irk_stem_direct(const IrConvBatch& b, int cin, cudaStream_t st);   // fused k3 conv for Cin <= 8 -> 32

// spconv_tc.cu  (tcgen05 / TMEM / TMA pair-GEMM, split-fp16)
int irk_pairgemm_tc(const IrConvBatch& b, int cin, int cout, int K, cudaStream_t st);
int irk_wgrad_tc(const float* x, int cin, const float* dy, int cout, int K, const int* in_idx, const int* out_idx,
                 const int* count, long long seg_cap, const float* dy_absmax, float* dW, cudaStream_t st);

// spconv_tma.cu  (the same pair-GEMM with the gather done by TMA tile::gather4; fin must be 16-byte aligned, n_max rows)
int irk_pairgemm_tma(const IrConvBatch& b, int cin, int cout, int K, cudaStream_t st);

// encoder_persist.cu  (all conv layers of one or two encoders in one persistent launch; `sync` = 512 zeroed bytes)
int irk_encoder_persist(int G, const IrConvProblem (*layers)[IR_MAX_GROUPS], const int* cin, const int* cout, const int* K,
                        int n_layers, void* sync, cudaStream_t st);
