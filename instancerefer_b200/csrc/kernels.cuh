// Internal launcher prototypes shared between the .cu files (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

// coords.cu
int irk_set_int(int* p, int v, const int* src, cudaStream_t st);
int irk_hash_build(const int32_t* coords, const int* n_dev, long long n_max, IrTable t,
                   cudaStream_t st);
int irk_downsample(const int32_t* coords, const int* n_dev, long long n_max, int new_stride,
                   IrTable t, int* pslot, int32_t* coords_out, int* n_out_dev,
                   unsigned long long* scan_state, cudaStream_t st);
int irk_voxelize(const float* pts, const int* cand, int n_cand, int ppi, int fdim, double voxel,
                 IrTable t, int* pslot, int32_t* coords_out, float* feats_out,
                 int* n_out_dev, unsigned long long* scan_state, cudaStream_t st);
int irk_kmap(int ks, const int32_t* coords_out, const int* n_out_dev, long long n_max,
             IrTable t, int stride, int* in_idx, long long seg_cap, int* slot,
             int* count, cudaStream_t st);

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
