// Scan preprocessing: the array work of data/scannet/prepare_data.py:30-216 and
// data/scannet/scannet_utils.py:18-44,97-116 on the device.  File parsing (PLY / JSON / TSV) stays on the host
// (instancerefer_b200/prepare_data.py); everything here is byte/integer/fp32 work with the exact arithmetic of the
// numpy original (separate multiplies and adds, no FMA contraction, correctly rounded sqrt/div), so results are
// bit-identical to the reference's .npy files.
#include "common.cuh"
#include "../../include/instancerefer_b200.h"

namespace {

constexpr int TPB = 256;
constexpr int KEEP_BLOCK = 1024;            // vertices per block in the stable compaction
constexpr int MAX_SMEM_OBJECTS = 1024;      // per-block min/max tables of k_box_minmax

struct Scratch {
    float* face_n;       // (n_faces, 3)
    int32_t* last;       // (3, n_verts): last face that names the vertex as corner c
    uint32_t* mm;        // (n_objects, 6) ordered-uint min xyz | max xyz
    int64_t* blk;        // (n_blocks + 1) per-block keep counts / offsets
    int64_t bytes;
};
Scratch scratch_of(void* base, int64_t n_verts, int64_t n_faces, int64_t n_objects) {
    Scratch s;
    char* q = (char*)base;
    auto take = [&](int64_t b) { char* o = q; q += (b + 255) / 256 * 256; return o; };
    s.face_n = (float*)take((n_faces > 0 ? n_faces : 1) * 3 * 4);
    s.last = (int32_t*)take((n_verts > 0 ? n_verts : 1) * 3 * 4);
    s.mm = (uint32_t*)take((n_objects > 0 ? n_objects : 1) * 6 * 4);
    s.blk = (int64_t*)take(((n_verts + KEEP_BLOCK - 1) / KEEP_BLOCK + 2) * 8);
    s.bytes = q - (char*)base;
    return s;
}

// ---- normals: scannet_utils.py:18-44.  `normals[faces[:,c]] += n` is a buffered fancy-index update: for a vertex
// named by several faces only the LAST face's normal lands, once per corner role c; the three roles add up in order.
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
    const float l2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const float d = __fadd_rn(__fsqrt_rn(l2), 1e-8f);
    x = __fdiv_rn(x, d);
    y = __fdiv_rn(y, d);
    z = __fdiv_rn(z, d);
}
__global__ void k_face_normals(const float* __restrict__ v9, int64_t n_verts, const int32_t* __restrict__ faces,
                               int64_t n_faces, float* __restrict__ face_n, int32_t* __restrict__ last) {
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < n_faces; f += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        const float* p0 = v9 + 9 * (int64_t)i0;
        const float* p1 = v9 + 9 * (int64_t)i1;
        const float* p2 = v9 + 9 * (int64_t)i2;
        const float a0 = __fsub_rn(p1[0], p0[0]), a1 = __fsub_rn(p1[1], p0[1]), a2 = __fsub_rn(p1[2], p0[2]);
        const float b0 = __fsub_rn(p2[0], p0[0]), b1 = __fsub_rn(p2[1], p0[1]), b2 = __fsub_rn(p2[2], p0[2]);
        float nx = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
        float ny = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
        float nz = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
        normalize3(nx, ny, nz);
        face_n[3 * f] = nx;
        face_n[3 * f + 1] = ny;
        face_n[3 * f + 2] = nz;
        atomicMax(&last[i0], (int32_t)f);
        atomicMax(&last[n_verts + i1], (int32_t)f);
        atomicMax(&last[2 * n_verts + i2], (int32_t)f);
    }
}
__global__ void k_vertex_normals(float* __restrict__ v9, int64_t n_verts, const float* __restrict__ face_n,
                                 const int32_t* __restrict__ last) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_verts; v += (int64_t)gridDim.x * blockDim.x) {
        float x = 0.f, y = 0.f, z = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int32_t f = last[c * n_verts + v];
            if (f >= 0) {
                x = __fadd_rn(x, face_n[3 * (int64_t)f]);
                y = __fadd_rn(y, face_n[3 * (int64_t)f + 1]);
                z = __fadd_rn(z, face_n[3 * (int64_t)f + 2]);
            }
        }
        normalize3(x, y, z);
        v9[9 * v + 6] = x;
        v9[9 * v + 7] = y;
        v9[9 * v + 8] = z;
    }
}

// ---- axis alignment: prepare_data.py:60-66 (homogeneous points times the transposed 4x4, in fp64, stored as fp32)
struct Mat4 { double m[16]; };
__global__ void k_align(const float* __restrict__ v9, int64_t n, Mat4 M, float* __restrict__ out) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const float* p = v9 + 9 * v;
        const double x = p[0], y = p[1], z = p[2];
        float* o = out + 9 * v;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double acc = __dmul_rn(x, M.m[4 * r]);
            acc = __fma_rn(y, M.m[4 * r + 1], acc);
            acc = __fma_rn(z, M.m[4 * r + 2], acc);
            acc = __dadd_rn(acc, M.m[4 * r + 3]);
            o[r] = (float)acc;
        }
#pragma unroll
        for (int c = 3; c < 9; ++c) o[c] = p[c];
    }
}

// ---- per-vertex labels from the segment tables: prepare_data.py:73-90
__global__ void k_vertex_labels(const int32_t* __restrict__ seg, int64_t n, const int32_t* __restrict__ seg_label,
                                const int32_t* __restrict__ seg_object, int32_t n_tab, uint32_t* __restrict__ label,
                                uint32_t* __restrict__ inst) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = seg[v];
        const bool ok = s >= 0 && s < n_tab;
        label[v] = ok ? (uint32_t)seg_label[s] : 0u;
        inst[v] = ok ? (uint32_t)seg_object[s] : 0u;
    }
}

// ---- per-object axis-aligned boxes: prepare_data.py:92-131.  min/max are order-independent, so atomics on the
// order-preserving integer image of the floats give the numpy result bit for bit.
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__global__ void k_box_init(uint32_t* mm, int32_t n_obj) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_obj * 6; i += gridDim.x * blockDim.x)
        mm[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}
__global__ void k_box_minmax(const float* __restrict__ v9, const uint32_t* __restrict__ inst, int64_t n, int32_t n_obj,
                             uint32_t* __restrict__ mm) {
    extern __shared__ uint32_t tab[];
    const bool local = n_obj <= MAX_SMEM_OBJECTS;
    if (local) {
        for (int i = threadIdx.x; i < n_obj * 6; i += blockDim.x) tab[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
        __syncthreads();
    }
    uint32_t* t = local ? tab : mm;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t id = inst[v];
        if (id == 0u || id > (uint32_t)n_obj) continue;
        uint32_t* row = t + 6 * (int64_t)(id - 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t o = f2ord(v9[9 * v + c]);
            atomicMin(&row[c], o);
            atomicMax(&row[3 + c], o);
        }
    }
    if (local) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_obj * 6; i += blockDim.x) {
            const uint32_t o = tab[i];
            if ((i % 6) < 3) { if (o != 0xffffffffu) atomicMin(&mm[i], o); }
            else if (o != 0u) atomicMax(&mm[i], o);
        }
    }
}
__global__ void k_box_finish(const uint32_t* __restrict__ mm, const int32_t* __restrict__ obj_label, int32_t n_obj,
                             double* __restrict__ boxes) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_obj) return;
    double* b = boxes + 8 * (int64_t)o;
    if (mm[6 * o] == 0xffffffffu) {            // no vertex carries this id: the row stays zero (prepare_data.py:98)
        for (int c = 0; c < 8; ++c) b[c] = 0.0;
        return;
    }
    for (int c = 0; c < 3; ++c) {
        const float lo = ord2f(mm[6 * o + c]), hi = ord2f(mm[6 * o + 3 + c]);
        b[c] = (double)__fdiv_rn(__fadd_rn(lo, hi), 2.f);
        b[3 + c] = (double)__fsub_rn(hi, lo);
    }
    b[6] = (double)(uint32_t)obj_label[o];
    b[7] = (double)o;
}

// ---- PointGroup proposals -> per-vertex labels: prepare_data.py:141-148 (later proposals overwrite earlier ones).
// A stream over the (n_inst, n) byte masks from the last proposal down: every thread owns 4 consecutive vertices (one
// 32-bit load per proposal row when n % 4 == 0), issues PG_DEPTH row loads before looking at them, and stops as soon
// as its four vertices are decided.
constexpr int PG_DEPTH = 8;
template <bool VEC>
__global__ void k_pointgroup(const uint8_t* __restrict__ masks, const int32_t* __restrict__ cls, int32_t n_inst, int64_t n,
                             uint32_t* __restrict__ label, uint32_t* __restrict__ inst) {
    const int64_t groups = (n + 3) / 4;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v0 = 4 * g;
        const int lanes = (int)(n - v0 < 4 ? n - v0 : 4);
        uint32_t id[4] = {0u, 0u, 0u, 0u};
        uint32_t open_mask = lanes == 4 ? 0xffffffffu : ((1u << (8 * lanes)) - 1u);   // byte j nonzero = vertex j undecided
        for (int hi = n_inst - 1; hi >= 0 && open_mask; hi -= PG_DEPTH) {
            uint32_t w[PG_DEPTH];
#pragma unroll
            for (int d = 0; d < PG_DEPTH; ++d) {
                const int i = hi - d;
                uint32_t x = 0u;
                if (i >= 0) {
                    const uint8_t* row = masks + (int64_t)i * n + v0;
                    if (VEC) x = *reinterpret_cast<const uint32_t*>(row);
                    else
                        for (int j = 0; j < lanes; ++j) x |= (uint32_t)row[j] << (8 * j);
                }
                w[d] = x;
            }
#pragma unroll
            for (int d = 0; d < PG_DEPTH; ++d) {
                const int i = hi - d;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t byte = 0xffu << (8 * j);
                    if ((w[d] & byte) && (open_mask & byte)) {
                        id[j] = (uint32_t)(i + 1);
                        open_mask &= ~byte;
                    }
                }
            }
        }
        for (int j = 0; j < lanes; ++j) {
            inst[v0 + j] = id[j];
            label[v0 + j] = id[j] ? (uint32_t)cls[id[j] - 1] : 0u;
        }
    }
}

// ---- np.logical_not(np.in1d(labels, DONOTCARE)) + boolean indexing: prepare_data.py:185-189 (order-preserving)
__device__ __forceinline__ bool keep_of(uint32_t lb, const int32_t* dc, int n_dc) {
    for (int i = 0; i < n_dc; ++i)
        if ((uint32_t)dc[i] == lb) return false;
    return true;
}
__global__ void k_keep_count(const uint32_t* __restrict__ sem, int64_t n, const int32_t* __restrict__ dc, int n_dc,
                             int64_t* __restrict__ blk) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * KEEP_BLOCK;
    int c = 0;
    for (int i = threadIdx.x; i < KEEP_BLOCK; i += blockDim.x) {
        const int64_t v = base + i;
        if (v < n && keep_of(sem[v], dc, n_dc)) ++c;
    }
    atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = cnt;
}
__global__ void k_keep_scan(int64_t* __restrict__ blk, int64_t n_blocks, int64_t* __restrict__ count_out) {
    // one block; exclusive scan of the per-block counts in place, total -> blk[n_blocks] and count_out
    __shared__ int64_t part[TPB];
    const int64_t per = (n_blocks + TPB - 1) / TPB;
    const int64_t lo = threadIdx.x * per, hi = lo + per < n_blocks ? lo + per : n_blocks;
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += blk[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < TPB; ++i) { const int64_t t = part[i]; part[i] = run; run += t; }
        blk[n_blocks] = run;
        *count_out = run;
    }
    __syncthreads();
    int64_t run = part[threadIdx.x];
    for (int64_t i = lo; i < hi; ++i) { const int64_t t = blk[i]; blk[i] = run; run += t; }
}
__global__ void k_keep_write(const uint32_t* __restrict__ sem, int64_t n, const int32_t* __restrict__ dc, int n_dc,
                             const int64_t* __restrict__ blk, int64_t* __restrict__ idx) {
    // KEEP_BLOCK == 4 * blockDim.x: every thread owns 4 consecutive vertices; warp/block scan of the per-thread counts
    __shared__ int wsum[TPB / 32];
    const int64_t base = (int64_t)blockIdx.x * KEEP_BLOCK + threadIdx.x * 4;
    bool k[4];
    int c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t v = base + j;
        k[j] = v < n && keep_of(sem[v], dc, n_dc);
        c += k[j];
    }
    int inc = c;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < w; ++i) woff += wsum[i];
    int64_t pos = blk[blockIdx.x] + woff + inc - c;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (k[j]) idx[pos++] = base + j;
}
__global__ void k_gather_rows(const uint8_t* __restrict__ src, int64_t row_bytes, const int64_t* __restrict__ idx,
                              const int64_t* __restrict__ count, int64_t m_max, uint8_t* __restrict__ dst) {
    const int64_t m = count ? (*count < m_max ? *count : m_max) : m_max;
    const int64_t words = row_bytes / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m * words; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / words, c = i % words;
        ((uint32_t*)dst)[i] = ((const uint32_t*)src)[idx[r] * words + c];
    }
}

int grid_for(int64_t n) { return ir_min_i(ir_div_up(n > 0 ? n : 1, TPB), IR_NUM_SMS * 8); }

}  // namespace

extern "C" int64_t ir_prepare_scratch_bytes(int64_t n_verts, int64_t n_faces, int32_t n_objects) {
    if (n_verts < 0 || n_faces < 0 || n_objects < 0) return 0;
    return scratch_of(nullptr, n_verts, n_faces, n_objects).bytes;
}

extern "C" int ir_mesh_normals(float* verts9, int64_t n_verts, const int32_t* faces, int64_t n_faces, void* scratch,
                               ir_stream_t stream) {
    IR_CHECK_ARG(n_verts >= 0 && n_faces >= 0 && n_verts < (1ll << 31) && n_faces < (1ll << 31));
    if (n_verts == 0) return IR_OK;
    IR_CHECK_ARG(verts9 && scratch && (faces || n_faces == 0));
    const Scratch s = scratch_of(scratch, n_verts, n_faces, 0);
    cudaStream_t st = (cudaStream_t)stream;
    IR_CHECK_CUDA(cudaMemsetAsync(s.last, 0xff, (size_t)n_verts * 3 * 4, st));
    if (n_faces > 0) {
        k_face_normals<<<grid_for(n_faces), TPB, 0, st>>>(verts9, n_verts, faces, n_faces, s.face_n, s.last);
        IR_CHECK_LAUNCH();
    }
    k_vertex_normals<<<grid_for(n_verts), TPB, 0, st>>>(verts9, n_verts, s.face_n, s.last);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

extern "C" int ir_align_vertices(const float* verts9, int64_t n_verts, const double* matrix16_host, float* aligned9,
                                 ir_stream_t stream) {
    IR_CHECK_ARG(n_verts >= 0 && matrix16_host);
    if (n_verts == 0) return IR_OK;
    IR_CHECK_ARG(verts9 && aligned9);
    Mat4 M;
    for (int i = 0; i < 16; ++i) M.m[i] = matrix16_host[i];
    k_align<<<grid_for(n_verts), TPB, 0, (cudaStream_t)stream>>>(verts9, n_verts, M, aligned9);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

extern "C" int ir_vertex_labels(const int32_t* seg_of_vert, int64_t n_verts, const int32_t* seg_label,
                                const int32_t* seg_object, int32_t n_seg_table, uint32_t* label_ids,
                                uint32_t* instance_ids, ir_stream_t stream) {
    IR_CHECK_ARG(n_verts >= 0 && n_seg_table >= 0);
    if (n_verts == 0) return IR_OK;
    IR_CHECK_ARG(seg_of_vert && label_ids && instance_ids && (n_seg_table == 0 || (seg_label && seg_object)));
    k_vertex_labels<<<grid_for(n_verts), TPB, 0, (cudaStream_t)stream>>>(seg_of_vert, n_verts, seg_label, seg_object,
                                                                          n_seg_table, label_ids, instance_ids);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

extern "C" int ir_instance_boxes(const float* verts9, const uint32_t* instance_ids, int64_t n_verts, int32_t n_objects,
                                 const int32_t* obj_label, void* scratch, double* boxes, ir_stream_t stream) {
    IR_CHECK_ARG(n_verts >= 0 && n_objects >= 0);
    if (n_objects == 0) return IR_OK;
    IR_CHECK_ARG(obj_label && scratch && boxes && (n_verts == 0 || (verts9 && instance_ids)));
    const Scratch s = scratch_of(scratch, 0, 0, n_objects);
    cudaStream_t st = (cudaStream_t)stream;
    k_box_init<<<ir_div_up(n_objects * 6, TPB), TPB, 0, st>>>(s.mm, n_objects);
    IR_CHECK_LAUNCH();
    if (n_verts > 0) {
        const size_t smem = n_objects <= MAX_SMEM_OBJECTS ? (size_t)n_objects * 6 * 4 : 0;
        k_box_minmax<<<ir_min_i(ir_div_up(n_verts, TPB * 4), IR_NUM_SMS * 4), TPB, smem, st>>>(verts9, instance_ids, n_verts,
                                                                                               n_objects, s.mm);
        IR_CHECK_LAUNCH();
    }
    k_box_finish<<<ir_div_up(n_objects, TPB), TPB, 0, st>>>(s.mm, obj_label, n_objects, boxes);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

extern "C" int ir_pointgroup_labels(const uint8_t* masks, const int32_t* cls, int32_t n_inst, int64_t n_verts,
                                    uint32_t* label_ids_pg, uint32_t* instance_ids_pg, ir_stream_t stream) {
    IR_CHECK_ARG(n_verts >= 0 && n_inst >= 0);
    if (n_verts == 0) return IR_OK;
    IR_CHECK_ARG(label_ids_pg && instance_ids_pg && (n_inst == 0 || (masks && cls)));
    const int grid = grid_for((n_verts + 3) / 4);
    if (n_verts % 4 == 0 && (reinterpret_cast<uintptr_t>(masks) & 3) == 0)
        k_pointgroup<true><<<grid, TPB, 0, (cudaStream_t)stream>>>(masks, cls, n_inst, n_verts, label_ids_pg, instance_ids_pg);
    else
        k_pointgroup<false><<<grid, TPB, 0, (cudaStream_t)stream>>>(masks, cls, n_inst, n_verts, label_ids_pg, instance_ids_pg);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

extern "C" int ir_keep_index(const uint32_t* sem_labels, int64_t n_verts, const int32_t* donotcare, int32_t n_donotcare,
                             void* scratch, int64_t* idx, int64_t* count_dev, ir_stream_t stream) {
    IR_CHECK_ARG(n_verts >= 0 && n_donotcare >= 0 && count_dev && scratch);
    IR_CHECK_ARG(n_donotcare == 0 || donotcare);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_verts == 0) {
        IR_CHECK_CUDA(cudaMemsetAsync(count_dev, 0, 8, st));
        return IR_OK;
    }
    IR_CHECK_ARG(sem_labels && idx);
    static_assert(KEEP_BLOCK == 4 * TPB, "k_keep_write owns 4 vertices per thread");
    const Scratch s = scratch_of(scratch, n_verts, 0, 0);
    const int64_t nb = (n_verts + KEEP_BLOCK - 1) / KEEP_BLOCK;
    k_keep_count<<<(int)nb, TPB, 0, st>>>(sem_labels, n_verts, donotcare, n_donotcare, s.blk);
    IR_CHECK_LAUNCH();
    k_keep_scan<<<1, TPB, 0, st>>>(s.blk, nb, count_dev);
    IR_CHECK_LAUNCH();
    k_keep_write<<<(int)nb, TPB, 0, st>>>(sem_labels, n_verts, donotcare, n_donotcare, s.blk, idx);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

extern "C" int ir_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, const int64_t* count_dev,
                              int64_t m_max, void* dst, ir_stream_t stream) {
    IR_CHECK_ARG(row_bytes > 0 && row_bytes % 4 == 0 && m_max >= 0);
    if (m_max == 0) return IR_OK;
    IR_CHECK_ARG(src && idx && dst);
    k_gather_rows<<<grid_for(m_max * (row_bytes / 4)), TPB, 0, (cudaStream_t)stream>>>((const uint8_t*)src, row_bytes, idx,
                                                                                       count_dev, m_max, (uint8_t*)dst);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
