// Probe: which kernel feature makes cudaOccupancyMaxActiveBlocksPerMultiprocessor drop to 1 for a 544-thread, 56-reg,
// 80 KB-smem kernel?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe_occ tools/probe_occupancy.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
struct Big { char b[2496]; int* out; };
__device__ __noinline__ float callee(const float* p, int n) { float s = 0; for (int i = 0; i < n; ++i) s += p[i] * p[i ^ 1]; return s; }
extern __shared__ uint8_t sm[];
template <int V>
__global__ void __launch_bounds__(544, 2) k(const __grid_constant__ Big P, const float* in, int n) {
    float acc = 0.f;
    if (V == 0) acc = in[threadIdx.x];
    if (V == 1) acc = callee(in, n);
    if (V == 2) {
        uint32_t* s_t = (uint32_t*)sm;
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_t)), "r"(256u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s_t), "r"(256u) : "memory");
        acc = in[threadIdx.x];
    }
    if (V == 3) { asm volatile("griddepcontrol.wait;" ::: "memory"); acc = in[threadIdx.x]; }
    if (V == 4) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); acc = (float)t + in[threadIdx.x]; __nanosleep(40); }
    if (V == 5) {
        uint64_t* b = (uint64_t*)sm;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)));
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)));
        }
        acc = in[threadIdx.x];
    }
    if (V == 6) { if (in[0] > 1e30f) __trap(); acc = in[threadIdx.x] + (float)clock64(); }
    P.out[threadIdx.x] = (int)acc + sm[threadIdx.x];
}
template <int V> void report(const char* name) {
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80868);
    int n = -1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k<V>, 544, 80868);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<V>);
    printf("%-28s regs %3d local %3zu -> %d CTAs/SM\n", name, fa.numRegs, fa.localSizeBytes, n);
}
int main() {
    report<0>("plain"); report<1>("noinline call"); report<2>("tcgen05.alloc"); report<3>("griddepcontrol.wait");
    report<4>("globaltimer+nanosleep"); report<5>("mbarrier.inval"); report<6>("trap+clock64");
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    for (int smem = 40000; smem <= 116000; smem += 4000) {
        int a = -1, b = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k<2>, 544, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k<0>, 544, smem);
        printf("smem %6d: tcgen05 kernel %d, plain kernel %d\n", smem, a, b);
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("regsPerSM %d smemPerSM %zu smemPerBlockOptin %zu reserved %zu maxThreadsPerSM %d\n", p.regsPerMultiprocessor,
           p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.reservedSharedMemPerBlock, p.maxThreadsPerMultiProcessor);
    return 0;
}
