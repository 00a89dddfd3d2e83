// Coordinate pipeline: hash build, first-wins voxelisation, stride-2 downsample, kernel maps.
// Integer work, bit-exact against oracle/sparse_ref.py (first-occurrence row order; rulebook
// compared after canonicalisation inside each kernel offset).  Replaces what the reference reaches
// through torchsparse at models/basic_blocks.py:14,32,39 (kernel-map build inside spnn.Conv3d),
// :68,73,78,83 (stride-2 downsample) and models/attribute_module.py:65-71,101
// (sparse_quantize + sparse_collate_tensors).  No host sync anywhere: all counts stay on device.
#include "common.cuh"
#include "kernels.cuh"

// ------------------------------------------------------------------ single-pass ordered compaction
// Ordered stream compaction with decoupled look-back: tiles are handed out by an atomic ticket; a tile publishes its
// own count at once (status 1 = aggregate), then one warp looks back over up to 32 predecessors per step, adding
// aggregates until it meets a tile whose inclusive prefix is known (status 2), and publishes its own inclusive prefix.
// One 64-bit word per tile (status << 32 | value): all tiles count concurrently and the look-back costs one or two L2
// round trips instead of one per tile (a chain of 16 tiles used to cost ~20 us for 32 k points).  Integer arithmetic:
// the result does not depend on the order in which tiles finish.
#define SCAN_THREADS 256
#define SCAN_ITEMS 2
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <class FlagFn, class EmitFn>
__device__ __forceinline__ void compact_ordered(int n, unsigned long long* state, int* n_out,
                                                FlagFn flag, EmitFn emit) {
    __shared__ int s_tile, s_prefix, s_warp[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n == 0 && blockIdx.x == 0 && tid == 0) *n_out = 0;
    while (true) {
        if (tid == 0) s_tile = (int)atomicAdd((unsigned int*)state, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int base = tile * SCAN_TILE + tid * SCAN_ITEMS;
        int f[SCAN_ITEMS], cnt = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            const int idx = base + j;
            f[j] = (idx < n) ? (flag(idx) ? 1 : 0) : 0;
            cnt += f[j];
        }
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            int ws = (lane < SCAN_THREADS / 32) ? s_warp[lane] : 0;
            int winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            const int tile_sum = __shfl_sync(0xffffffffu, winc, SCAN_THREADS / 32 - 1);
            if (lane < SCAN_THREADS / 32) s_warp[lane] = winc - ws;      // exclusive warp offsets
            volatile unsigned long long* st = state + 1;
            unsigned prefix = 0;
            if (tile > 0) {
                if (lane == 0) st[tile] = (1ull << 32) | (unsigned long long)(unsigned)tile_sum;      // aggregate available
                int look = tile - 1;
                while (true) {
                    const int idx = look - lane;                     // lane 0 = nearest predecessor
                    unsigned long long v = 2ull << 32;               // before tile 0: inclusive prefix 0
                    if (idx >= 0) {
                        unsigned spins = 0;
                        do { v = st[idx]; if (++spins > (1u << 27)) __trap(); } while ((v >> 32) == 0ull);      // bounded: a protocol bug traps
                    }
                    const unsigned status = (unsigned)(v >> 32), val = (unsigned)(v & 0xFFFFFFFFull);
                    const unsigned inc_mask = __ballot_sync(0xffffffffu, status == 2u);
                    const int first_inc = __ffs(inc_mask) - 1;       // nearest tile with a known inclusive prefix
                    unsigned c = (first_inc < 0 || lane <= first_inc) ? val : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    prefix += c;
                    if (first_inc >= 0) break;
                    look -= 32;
                }
            }
            if (lane == 0) {
                st[tile] = (2ull << 32) | (unsigned long long)(prefix + (unsigned)tile_sum);          // inclusive prefix available
                s_prefix = (int)prefix;
                if (tile == ntiles - 1) *n_out = (int)prefix + tile_sum;
            }
        }
        __syncthreads();
        int pos = s_prefix + s_warp[w] + (inc - cnt);
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j)
            if (f[j]) emit(base + j, pos++);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ kernels
typedef IrLevels LevelTables;

__device__ __forceinline__ int4 parent_coord(int4 c, int new_stride) {
    // floor(c / ns) * ns for power-of-two ns, negatives included (two's complement floor)
    const int m = ~(new_stride - 1);
    return make_int4(c.x & m, c.y & m, c.z & m, c.w);
}

// Level registration, one thread per (row, level): level 0 hashes the row itself; level l >= 1 makes
// the ancestor voxel (stride 2^l) of level-0 row i remember the smallest level-0 row below it.  Row
// order of level l = order of that minimum = the first-occurrence order that repeated torchsparse-
// style downsampling produces (first occurrence is transitive).  grid.y = level - first_level.
__global__ void k_insert_levels(const int4* __restrict__ coords, int n_host, const int* __restrict__ n_dev,
                                IrLevels lt, int* __restrict__ pslot, long long n_max, int first_level,
                                int* __restrict__ nlvl0) {
    const int n = n_dev ? *n_dev : n_host;
    const int l = blockIdx.y + first_level;
    if (nlvl0 != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *nlvl0 = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = coords[i];
        if (l == 0) {
            const int s = ir_ht_insert(lt.t[0], ir_pack_key(c.x, c.y, c.z, c.w));
            atomicMin(&lt.t[0].minrow[s], i);
            atomicMin(&lt.t[0].row[s], i);
        } else {
            const int4 p = parent_coord(c, 1 << l);
            const int s = ir_ht_insert(lt.t[l], ir_pack_key(p.x, p.y, p.z, p.w));
            atomicMin(&lt.t[l].minrow[s], i);
            pslot[(long long)(l - 1) * n_max + i] = s;
        }
    }
}

// ordered compaction of levels 1..4 (blockIdx.y = level-1), all from the level-0 rows
__global__ void __launch_bounds__(SCAN_THREADS)
k_levels_compact(const int4* __restrict__ coords0, const int* __restrict__ nlvl, LevelTables lt,
                 const int* __restrict__ pslot, long long n_max, int4* __restrict__ c1, int4* __restrict__ c2,
                 int4* __restrict__ c3, int4* __restrict__ c4, int* __restrict__ nlvl_out,
                 unsigned long long* state, long long state_stride) {
    const int l = blockIdx.y + 1;
    const int n = nlvl[0];
    const IrTable t = lt.t[l];
    const int* ps = pslot + (long long)(l - 1) * n_max;
    int4* cout = (l == 1) ? c1 : (l == 2) ? c2 : (l == 3) ? c3 : c4;
    compact_ordered(
        n, state + (long long)l * state_stride, nlvl_out + l,
        [&](int i) { return t.minrow[ps[i]] == i; },
        [&](int i, int r) {
            cout[r] = parent_coord(coords0[i], 1 << l);
            t.row[ps[i]] = r;
        });
}

// points (n_inst, ppi, fdim) fp32; candidate m uses instance cand[m]; voxel key = (floor(xyz/voxel), m)
__global__ void k_vox_insert(const float* __restrict__ pts, const int* __restrict__ cand, int n_pts,
                             int ppi, int fdim, double voxel, IrTable t, int* __restrict__ pslot) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pts; p += gridDim.x * blockDim.x) {
        const int m = p / ppi, j = p - m * ppi;
        const float* src = pts + ((long long)cand[m] * ppi + j) * fdim;
        const int x = (int)floor((double)src[0] / voxel);
        const int y = (int)floor((double)src[1] / voxel);
        const int z = (int)floor((double)src[2] / voxel);
        const int s = ir_ht_insert(t, ir_pack_key(x, y, z, m));
        atomicMin(&t.minrow[s], p);
        pslot[p] = s;
    }
}

// first-point-wins compaction -> level-0 rows (coords + the winning point's feature row).
// vslot: the insert slots above.
__global__ void __launch_bounds__(SCAN_THREADS)
k_vox_compact(const float* __restrict__ pts, const int* __restrict__ cand, int n_pts, int ppi,
              int fdim, double voxel, IrTable t, const int* __restrict__ vslot,
              int4* __restrict__ coords_out, float* __restrict__ feats_out, int* __restrict__ n_out,
              unsigned long long* state) {
    compact_ordered(
        n_pts, state, n_out,
        [&](int p) { return t.minrow[vslot[p]] == p; },
        [&](int p, int r) {
            const int m = p / ppi, j = p - m * ppi;
            const float* src = pts + ((long long)cand[m] * ppi + j) * fdim;
            const int4 c = make_int4((int)floor((double)src[0] / voxel), (int)floor((double)src[1] / voxel),
                                     (int)floor((double)src[2] / voxel), m);
            coords_out[r] = c;
            for (int q = 0; q < fdim; ++q) feats_out[(long long)r * fdim + q] = src[q];
            t.row[vslot[p]] = r;
        });
}

// All nine kernel maps of an encoder in one launch (grid.y walks map x offset group).  Appendix-A
// offset enumeration:
//   k3: k = (dz+1)*9 + (dy+1)*3 + (dx+1), offsets {-1,0,1}*stride
//   k2: k = 4*bx + 2*by + bz,            offsets {0,1}*stride   (stride = INPUT level stride)
// One hash probe per thread, one block-aggregated append per CTA.  Emits per offset k:
// in_idx[k*n_max + pos] (input row of pair pos), count[k], slot[k*n_max + o] = pos or -1.
typedef IrKmapArgs KmapArgs;

#define KM_K3_GROUPS 9          // 27 offsets = 9 groups of 3 per thread
#define KM_NY (5 * KM_K3_GROUPS + 4 * 2)   // + 4 k2 maps x 2 groups of 4
__global__ void __launch_bounds__(256)
k_kmap_all(KmapArgs a) {
    // blockIdx.y < 45: k3 map of level y/9, offsets 3*(y%9) .. +2;  else k2s2 map l=(y-45)/2, offsets
    // 4*((y-45)%2) .. +3.  A thread probes its 3-4 offsets back to back (independent L2 latencies).
    const int y = blockIdx.y;
    const bool is3 = y < 5 * KM_K3_GROUPS;
    const int l = is3 ? y / KM_K3_GROUPS : (y - 5 * KM_K3_GROUPS) / 2;
    const int k0 = is3 ? 3 * (y % KM_K3_GROUPS) : 4 * ((y - 5 * KM_K3_GROUPS) % 2);
    const int NK = is3 ? 3 : 4;
    const int lo = is3 ? l : l + 1;                       // output level
    const int stride = 1 << l;
    const int n = a.nlvl[lo];
    const int4* __restrict__ coords_out = a.coords[lo];
    const IrTable tin = a.lt.t[l];
    int* __restrict__ in_base = is3 ? a.k3_in[l] : a.k2_in[l];
    int* __restrict__ slot_base = is3 ? a.k3_slot[l] : a.k2_slot[l];
    int* __restrict__ cnt = a.kcount + (is3 ? l : 5 + l) * 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ int s_wcnt[4][8];
    __shared__ int s_base[4];
    const int n_round = (n + 255) & ~255;                  // whole CTAs iterate together (block-level scan)
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_round; o += gridDim.x * blockDim.x) {
        const bool live = o < n;
        int4 c = make_int4(0, 0, 0, 0);
        if (live) c = coords_out[o];
        int sl[4], j[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            sl[u] = -1;
            j[u] = -1;
            const int k = k0 + u;
            if (live && u < NK && !(is3 && k == 13)) {
                int dx, dy, dz;
                if (is3) { dx = (k % 3 - 1) * stride; dy = ((k / 3) % 3 - 1) * stride; dz = (k / 9 - 1) * stride; }
                else     { dx = (k >> 2) * stride; dy = ((k >> 1) & 1) * stride; dz = (k & 1) * stride; }
                sl[u] = ir_ht_find(tin, ir_pack_key(c.x + dx, c.y + dy, c.z + dz, c.w));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (sl[u] >= 0) j[u] = tin.row[sl[u]];
            if (live && u < NK && is3 && k0 + u == 13) j[u] = o;       // centre offset: the row itself
        }
        // one atomicAdd per CTA and offset (the 27 / 8 counters are hot): ballot inside the warp,
        // 8-entry scan across warps, threads 0..3 claim the block's ranges
        unsigned m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            m[u] = __ballot_sync(0xffffffffu, j[u] >= 0);
            if (lane == 0) s_wcnt[u][wid] = __popc(m[u]);
        }
        __syncthreads();
        if (threadIdx.x < NK) {
            const int u = threadIdx.x;
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int cc = s_wcnt[u][w]; s_wcnt[u][w] = tot; tot += cc; }
            s_base[u] = tot ? atomicAdd(cnt + k0 + u, tot) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < NK) {
                const long long seg = (long long)(k0 + u) * a.n_max;
                int pos = -1;
                if (j[u] >= 0) {
                    pos = s_base[u] + s_wcnt[u][wid] + __popc(m[u] & ((1u << lane) - 1u));
                    in_base[seg + pos] = j[u];
                }
                if (live) slot_base[seg + o] = pos;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host launchers (internal)
static inline int grid_for(long long n, int threads) {
    int g = ir_div_up(n > 0 ? n : 1, threads);
    const int cap = IR_NUM_SMS * 8;
    return g < cap ? g : cap;
}
static inline int scan_grid(long long n) {
    return ir_min_i(ir_div_up(n > 0 ? n : 1, SCAN_TILE), IR_NUM_SMS * 2);
}

int irk_levels_from_coords(const int32_t* coords0, int n0, const int* n0_dev, IrLevels lt, int* pslot,
                           long long n_max, int* nlvl, cudaStream_t st) {
    k_insert_levels<<<dim3(grid_for(n0, 256), 5), 256, 0, st>>>((const int4*)coords0, n0, n0_dev, lt, pslot, n_max, 0, nlvl);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

int irk_levels_compact(const int32_t* coords0, long long n0_max, int* nlvl, IrLevels lt, const int* pslot,
                       long long n_max, int32_t* c1, int32_t* c2, int32_t* c3, int32_t* c4,
                       unsigned long long* scan_state, long long scan_stride, cudaStream_t st) {
    k_levels_compact<<<dim3(scan_grid(n0_max), 4), SCAN_THREADS, 0, st>>>(
        (const int4*)coords0, nlvl, lt, pslot, n_max, (int4*)c1, (int4*)c2, (int4*)c3, (int4*)c4, nlvl,
        scan_state, scan_stride);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

int irk_voxelize(const float* pts, const int* cand, int n_cand, int ppi, int fdim, double voxel,
                 IrLevels lt, int* vslot, int32_t* coords_out, float* feats_out, int* nlvl,
                 unsigned long long* scan_state, int* pslot, long long n_max, cudaStream_t st) {
    const long long n_pts = (long long)n_cand * ppi;
    k_vox_insert<<<grid_for(n_pts, 256), 256, 0, st>>>(pts, cand, (int)n_pts, ppi, fdim, voxel, lt.t[0], vslot);
    IR_CHECK_LAUNCH();
    k_vox_compact<<<scan_grid(n_pts), SCAN_THREADS, 0, st>>>(pts, cand, (int)n_pts, ppi, fdim, voxel, lt.t[0], vslot,
                                                            (int4*)coords_out, feats_out, nlvl, scan_state);
    IR_CHECK_LAUNCH();
    // ancestors of the new level-0 rows (levels 1..4), one thread per (row, level)
    k_insert_levels<<<dim3(grid_for(n_pts, 256), 4), 256, 0, st>>>((const int4*)coords_out, 0, nlvl, lt, pslot, n_max, 1, nullptr);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

int irk_kmap_all(const IrKmapArgs& a, long long rows_max, cudaStream_t st) {
    const int gx = ir_min_i(ir_div_up(rows_max > 0 ? rows_max : 1, 256), IR_NUM_SMS);
    k_kmap_all<<<dim3(gx, KM_NY), 256, 0, st>>>(a);
    IR_CHECK_LAUNCH();
    return IR_OK;
}

// ------------------------------------------------------------------ standalone voxeliser (loader side)
// sparse_quantize of whole scenes in the loader (lib/dataset.py:255-261) + sparse_collate batch index:
// first-point-wins voxelisation of (n_cloud, ppi, fdim) point clouds into caller buffers, with its own
// small scratch (hash table sized by the POINT count, insert slots, scan state) instead of an encoder
// workspace.  Same kernels as ir_voxelize, so rows come out in first-occurrence order, bit-exact.
#include "../../include/instancerefer_b200.h"
static inline long long vp_cap(long long n_pts) { long long c = 1024; while (c < 2 * n_pts) c <<= 1; return c; }
extern "C" size_t ir_voxelize_points_scratch_bytes(int64_t n_pts) {
    const long long cap = vp_cap(n_pts);
    return (size_t)(cap * 16 + ((n_pts * 4 + 1023) / 1024) * 1024 + (3 + n_pts / SCAN_TILE) * 8 + 1024);
}
extern "C" int ir_voxelize_points(const float* pts, const int32_t* cloud, int32_t n_cloud, int32_t ppi, int32_t fdim,
                                  double voxel, void* scratch, int32_t* coords_out, float* feats_out,
                                  int32_t* count_out, ir_stream_t stream) {
    IR_CHECK_ARG(pts && cloud && scratch && coords_out && feats_out && count_out);
    IR_CHECK_ARG(n_cloud > 0 && n_cloud < IR_MAX_BATCH && ppi > 0 && fdim >= 3 && fdim <= 8 && voxel > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_pts = (long long)n_cloud * ppi;
    IR_CHECK_ARG(n_pts < (1ll << 30));
    const long long cap = vp_cap(n_pts);
    char* base = (char*)scratch;
    const IrTable t = ir_table_view(base, cap);                          // keys u64[cap] | minrow | row
    int* vslot = (int*)(base + cap * 16);
    unsigned long long* state = (unsigned long long*)(base + cap * 16 + ((n_pts * 4 + 1023) / 1024) * 1024);
    IR_CHECK_CUDA(cudaMemsetAsync(base, 0x7F, (size_t)cap * 16, st));
    IR_CHECK_CUDA(cudaMemsetAsync(state, 0, (size_t)(3 + n_pts / SCAN_TILE) * 8, st));
    k_vox_insert<<<grid_for(n_pts, 256), 256, 0, st>>>(pts, cloud, (int)n_pts, ppi, fdim, voxel, t, vslot);
    IR_CHECK_LAUNCH();
    k_vox_compact<<<scan_grid(n_pts), SCAN_THREADS, 0, st>>>(pts, cloud, (int)n_pts, ppi, fdim, voxel, t, vslot,
                                                            (int4*)coords_out, feats_out, count_out, state);
    IR_CHECK_LAUNCH();
    return IR_OK;
}
