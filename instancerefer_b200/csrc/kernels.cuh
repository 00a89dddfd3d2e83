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

// spconv.cu  (SIMT fp32 pair-GEMM + deterministic reduce/epilogue)
int irk_pairgemm_simt(const float* feat_in, int cin, int cout, int K, const int* in_idx,
                      long long seg_cap, const int* count, const float* weight, float* T,
                      long long pairs_max, cudaStream_t st);
int irk_reduce_epilogue(const float* T, int cout, int K, const int* slot, long long seg_cap,
                        const int* count, const int* n_out_dev, long long n_max, const float* scale,
                        const float* shift, const float* resid, int relu, float* out,
                        cudaStream_t st);

// spconv_tc.cu  (tcgen05 / TMEM / TMA pair-GEMM, 3xTF32)
int irk_pairgemm_tc(const float* feat_in, int cin, int cout, int K, const int* in_idx,
                    long long seg_cap, const int* count, const float* wprep, float* T,
                    long long pairs_max, cudaStream_t st);
