// Shared helpers for the instancerefer_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define IR_OK 0
#define IR_ERR_ARG (-1)
#define IR_ERR_CUDA (-2)
#define IR_ERR_UNSUPPORTED (-3)

#define IR_NUM_SMS 148

void ir_set_error(const char* fmt, ...);
void ir_count_launch(int n);

#define IR_CHECK_ARG(cond)                                                        \
    do {                                                                          \
        if (!(cond)) {                                                            \
            ir_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);   \
            return IR_ERR_ARG;                                                    \
        }                                                                         \
    } while (0)

#define IR_CHECK_LAUNCH()                                                         \
    do {                                                                          \
        ir_count_launch(1);                                                       \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) {                                                  \
            ir_set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return IR_ERR_CUDA;                                                   \
        }                                                                         \
    } while (0)

#define IR_CHECK_CUDA(call)                                                       \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            ir_set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return IR_ERR_CUDA;                                                   \
        }                                                                         \
    } while (0)

// Programmatic dependent launch (PDL): the kernel may begin while its stream predecessor is still
// running; it must execute ir_pdl_wait() before touching anything the predecessor writes.  Kernels
// call ir_pdl_trigger() early to let their successor start its own independent prologue.
__device__ __forceinline__ void ir_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ir_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t ir_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline int ir_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int ir_min_i(long long a, long long b) { return (int)(a < b ? a : b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------ coordinate hash table
// keys u64[cap] | minrow i32[cap] | row i32[cap]; open addressing, linear probe.  EMPTY and NOROW
// share the byte pattern 0x7F so one memset clears a whole table set; batch ids must stay < 0x7F7F.
#define IR_EMPTY_KEY 0x7F7F7F7F7F7F7F7Full
#define IR_MAX_BATCH 0x7F7F
#define IR_NOROW 0x7F7F7F7F

struct IrTable {
    unsigned long long* keys;
    int* minrow;
    int* row;
    unsigned mask;
};

__host__ __device__ inline IrTable ir_table_view2(void* keys, void* vals, long long cap) {
    IrTable t;
    t.keys = (unsigned long long*)keys;
    t.minrow = (int*)vals;
    t.row = t.minrow + cap;
    t.mask = (unsigned)(cap - 1);
    return t;
}

__host__ __device__ inline IrTable ir_table_view(void* base, long long cap) {
    IrTable t;
    t.keys = (unsigned long long*)base;
    t.minrow = (int*)((char*)base + cap * 8);
    t.row = t.minrow + cap;
    t.mask = (unsigned)(cap - 1);
    return t;
}

// collision-free pack of (x,y,z,b): b:16 | x+32768:16 | y+32768:16 | z+32768:16
__device__ __forceinline__ unsigned long long ir_pack_key(int x, int y, int z, int b) {
    return ((unsigned long long)(unsigned)(b & 0xFFFF) << 48) |
           ((unsigned long long)(unsigned)((x + 32768) & 0xFFFF) << 32) |
           ((unsigned long long)(unsigned)((y + 32768) & 0xFFFF) << 16) |
           (unsigned long long)(unsigned)((z + 32768) & 0xFFFF);
}
__device__ __forceinline__ unsigned ir_hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (unsigned)k;
}
__device__ __forceinline__ int ir_ht_insert(const IrTable& t, unsigned long long key) {
    unsigned h = ir_hash64(key) & t.mask;
    while (true) {
        unsigned long long prev = atomicCAS(&t.keys[h], IR_EMPTY_KEY, key);
        if (prev == IR_EMPTY_KEY || prev == key) return (int)h;
        h = (h + 1) & t.mask;
    }
}
__device__ __forceinline__ int ir_ht_find(const IrTable& t, unsigned long long key) {
    unsigned h = ir_hash64(key) & t.mask;
    while (true) {
        unsigned long long cur = t.keys[h];
        if (cur == key) return (int)h;
        if (cur == IR_EMPTY_KEY) return -1;
        h = (h + 1) & t.mask;
    }
}
