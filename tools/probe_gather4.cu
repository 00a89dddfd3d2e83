// Probe (run on the GPU box): semantics of cp.async.bulk.tensor.2d tile::gather4 on sm_100a —
// tensor-map box shape it accepts, smem layout it produces under SWIZZLE_128B, out-of-bounds rows.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/probe_gather4 tools/probe_gather4.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap tm, int col0, int r0, int r1, int r2, int r3, int dst_off,
                        uint32_t tx_bytes, uint16_t* out, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* tile = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
    for (int i = threadIdx.x; i < 4096 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(tile)[i] = 0xDEAD;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(tx_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(smem_u32(tile + dst_off)), "l"(&tm), "r"(smem_u32(&bar)), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
            : "memory");
        long long t0 = clock64();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            if (clock64() - t0 > 200000000ll) break;
        }
        *status = ok ? 1 : -1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(tile)[i];
}

static uint16_t val(int row, int col) { return (uint16_t)((row * 131 + col) & 0x7FFF); }

int main() {
    const int R = 1000, Cc = 128;
    uint16_t* h = new uint16_t[R * Cc];
    for (int r = 0; r < R; ++r) for (int c = 0; c < Cc; ++c) h[r * Cc + c] = val(r, c);
    uint16_t* d; cudaMalloc(&d, R * Cc * 2); cudaMemcpy(d, h, R * Cc * 2, cudaMemcpyHostToDevice);
    uint16_t* dout; cudaMalloc(&dout, 4096); int* dstat; cudaMalloc(&dstat, 4);
    EncodeTiled enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)Cc, (cuuint64_t)R};
        cuuint64_t strides[1] = {(cuuint64_t)Cc * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
        cuuint32_t es[2] = {1, 1};
        CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== box {64,%d}: encode rc=%d\n", boxrows, (int)cr);
        if (cr != CUDA_SUCCESS) continue;
        const int tests[3][6] = {{64, 5, 900, 17, 3, 0}, {0, 7, -1, 2000, 8, 512}, {64, 1, 2, 3, 4, 1024 + 512}};
        for (int t = 0; t < 3; ++t) {
            const int* T = tests[t];
            cudaMemset(dout, 0, 4096); cudaMemset(dstat, 0, 4);
            k_probe<<<1, 128, 8192>>>(tm, T[0], T[1], T[2], T[3], T[4], T[5], 512u, dout, dstat);
            cudaError_t e = cudaDeviceSynchronize();
            int st = 0; cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
            uint16_t o[2048]; cudaMemcpy(o, dout, 4096, cudaMemcpyDeviceToHost);
            printf("-- test %d col0=%d rows={%d,%d,%d,%d} dst_off=%d: sync=%s status=%d\n", t, T[0], T[1], T[2], T[3], T[4], T[5],
                   cudaGetErrorString(e), st);
            if (e != cudaSuccess) { printf("kernel failed; stopping\n"); return 2; }
            // describe every 16-byte chunk of the 4 KB tile that is not the 0xDEAD fill
            for (int ch = 0; ch < 256; ++ch) {
                const uint16_t* p = o + ch * 8;
                if (p[0] == 0xDEAD && p[7] == 0xDEAD) continue;
                // find which (row, col) this chunk came from
                int fr = -9, fc = -9;
                bool zero = true; for (int i = 0; i < 8; ++i) zero = zero && p[i] == 0;
                if (!zero)
                    for (int k = 0; k < 4 && fr < 0; ++k) {
                        const int r = T[1 + k];
                        if (r < 0 || r >= R) continue;
                        for (int c = 0; c < Cc; c += 8)
                            if (p[0] == val(r, c) && p[7] == val(r, c + 7)) { fr = r; fc = c; break; }
                    }
                const int smrow = ch / 8, smchunk = ch % 8;
                printf("   smem row %2d chunk %d (byte %4d): %s row=%d col=%d  [expect SW128 logical chunk %d]\n", smrow, smchunk, ch * 16,
                       zero ? "ZERO" : (fr >= 0 ? "data" : "????"), fr, fc, smchunk ^ (smrow & 7));
            }
        }
    }
    printf("done\n");
    return 0;
}
