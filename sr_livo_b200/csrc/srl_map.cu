// srl_map.cu — HBM-resident voxel map: bulk mirror (K4) and order-preserving insertion (K3).
//
// K3 replaces lioOptimization::addPointsToMap / addPointToMap (src/lioOptimization.cpp:400-446,520-554):
// the reference inserts the registered frame point by point, and a point's acceptance depends on the points
// already accepted into its voxel (including earlier points of the same sweep).  Voxels are independent of each
// other, so the GPU version is: float-round + key per point -> stable radix sort by key (keeps sweep order inside
// a voxel) -> one warp per touched voxel replays the reference's sequential rule over that voxel's points.
// Resulting block contents (points and their order) are identical to the reference's voxelBlock::points.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <tr1/unordered_map>

#include "srl_internal.h"

namespace srl {

constexpr unsigned long long kInvalidKey = 1ull << 49;

__device__ __forceinline__ int slot_find_rw(const Slot* slots, unsigned int mask, unsigned long long key, int x, int y, int z) {
    unsigned int idx = hash_key(x, y, z) & mask;
    for (;;) {
        const unsigned long long k = slots[idx].key;
        if (k == key) return (int)idx;
        if (k == 0ull) return -1;
        idx = (idx + 1) & mask;
    }
}

__device__ __forceinline__ int slot_claim(Slot* slots, unsigned int mask, unsigned long long key, int x, int y, int z,
                                          unsigned int block, unsigned int count) {
    unsigned int idx = hash_key(x, y, z) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(&slots[idx].key, 0ull, key);
        if (old == 0ull) { slots[idx].block = block; slots[idx].count = count; return (int)idx; }
        if (old == key) return -1;   // duplicate key (caller guarantees uniqueness)
        idx = (idx + 1) & mask;
    }
}

// ---- K4: bulk mirror of a host voxelHashMap ---------------------------------------------------------------
__global__ void k_upload_slots(Slot* slots, unsigned int mask, float* blocks, const short* keys, const int* counts,
                               long long n_voxels, int* dup_flag) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_voxels) return;
    const int x = keys[3 * v], y = keys[3 * v + 1], z = keys[3 * v + 2];
    const unsigned long long key = pack_key(x, y, z);
    if (slot_claim(slots, mask, key, x, y, z, (unsigned)v, (unsigned)counts[v]) < 0) *dup_flag = 1;
    unsigned int* meta = reinterpret_cast<unsigned int*>(blocks + (size_t)v * kBlockFloats);
    meta[kMetaKeyLo] = (unsigned int)(key & 0xffffffffu); meta[kMetaKeyHi] = (unsigned int)(key >> 32); meta[kMetaCount] = (unsigned)counts[v];
}
__global__ void k_upload_points(float* blocks, const float* xyz, long long n_voxels, int cap) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_voxels * cap) return;
    const long long v = e / cap;
    const int i = (int)(e % cap);
    float* b = blocks + (size_t)v * kBlockFloats + 4 * i;
    b[0] = xyz[3 * e]; b[1] = xyz[3 * e + 1]; b[2] = xyz[3 * e + 2];
}
__global__ void k_download(const float* blocks, long long n_voxels, int cap, short* keys, int* counts, float* xyz) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_voxels * cap) return;
    const long long v = e / cap;
    const int i = (int)(e % cap);
    const float* b = blocks + (size_t)v * kBlockFloats;
    const unsigned int* meta = reinterpret_cast<const unsigned int*>(b);
    const int cnt = (int)meta[kMetaCount];
    if (i == 0) {
        short x, y, z;
        unpack_key((unsigned long long)meta[kMetaKeyLo] | ((unsigned long long)meta[kMetaKeyHi] << 32), x, y, z);
        keys[3 * v] = x; keys[3 * v + 1] = y; keys[3 * v + 2] = z;
        counts[v] = cnt;
    }
    xyz[3 * e] = i < cnt ? b[4 * i] : 0.f;
    xyz[3 * e + 1] = i < cnt ? b[4 * i + 1] : 0.f;
    xyz[3 * e + 2] = i < cnt ? b[4 * i + 2] : 0.f;
}

// ---- K3: insertion ------------------------------------------------------------------------------------------
// rgbPoint ctor: position = position_.cast<float>() (src/cloudMap.cpp:5-9); key from the float-rounded position
// (src/lioOptimization.cpp:403-405)
__global__ void k_insert_keys(const double* __restrict__ xyz, long long n, double size, unsigned long long* keys,
                              unsigned int* idx, float* fxyz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fx = __double2float_rn(xyz[3 * i]), fy = __double2float_rn(xyz[3 * i + 1]), fz = __double2float_rn(xyz[3 * i + 2]);
    fxyz[3 * i] = fx; fxyz[3 * i + 1] = fy; fxyz[3 * i + 2] = fz;
    const double qx = __ddiv_rn((double)fx, size), qy = __ddiv_rn((double)fy, size), qz = __ddiv_rn((double)fz, size);
    unsigned long long key = kInvalidKey;
    if (fabs(qx) < 32765.0 && fabs(qy) < 32765.0 && fabs(qz) < 32765.0) key = pack_key((int)qx, (int)qy, (int)qz);
    keys[i] = key;
    idx[i] = (unsigned int)i;
}

__global__ void k_seg_flags(const unsigned long long* __restrict__ keys, long long n, unsigned char* flags) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long k = keys[j];
    flags[j] = (k != kInvalidKey && (j == 0 || keys[j - 1] != k)) ? 1 : 0;
}

__global__ void k_seg_lookup(const Slot* slots, unsigned int mask, const unsigned long long* __restrict__ keys,
                             const unsigned int* __restrict__ seg_start, const int* n_seg_p, int allow_new,
                             int* seg_slot, unsigned int* is_new) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= *n_seg_p) return;
    const unsigned long long key = keys[seg_start[s]];
    short x, y, z;
    unpack_key(key, x, y, z);
    const int slot = slot_find_rw(slots, mask, key, x, y, z);
    seg_slot[s] = slot;
    is_new[s] = (slot < 0 && allow_new) ? 1u : 0u;
}

__global__ void k_seg_claim(Slot* slots, unsigned int mask, float* blocks, const unsigned long long* __restrict__ keys,
                            const unsigned int* __restrict__ seg_start, const int* n_seg_p,
                            const unsigned int* __restrict__ is_new, const unsigned int* __restrict__ new_rank,
                            long long block_base, int* seg_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= *n_seg_p || !is_new[s]) return;
    const unsigned long long key = keys[seg_start[s]];
    short x, y, z;
    unpack_key(key, x, y, z);
    const unsigned int blk = (unsigned int)(block_base + new_rank[s]);
    seg_slot[s] = slot_claim(slots, mask, key, x, y, z, blk, 0u);
    unsigned int* meta = reinterpret_cast<unsigned int*>(blocks + (size_t)blk * kBlockFloats);
    meta[kMetaKeyLo] = (unsigned int)(key & 0xffffffffu); meta[kMetaKeyHi] = (unsigned int)(key >> 32); meta[kMetaCount] = 0;
}

// one warp per touched voxel: the reference's per-point rule, replayed in sweep order
__global__ void __launch_bounds__(256) k_seg_process(Slot* slots, float* blocks, const unsigned long long* __restrict__ keys,
                                                      const unsigned int* __restrict__ idx, const float* __restrict__ fxyz,
                                                      const unsigned int* __restrict__ seg_start, const int* n_seg_p,
                                                      const int* __restrict__ seg_slot, long long n, double size, int cap,
                                                      double min_dist, int min_num_points, long long* n_points) {
    const int lane = threadIdx.x & 31;
    const int s = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (s >= *n_seg_p) return;
    const int slot = seg_slot[s];
    if (slot < 0) return;   // voxel absent and min_num_points > 0: nothing is ever created (:437)
    const unsigned int blk = slots[slot].block;
    int count = (int)slots[slot].count;
    float* bp = blocks + (size_t)blk * kBlockFloats;
    float ex = 0.f, ey = 0.f, ez = 0.f;
    if (lane < count) { ex = bp[4 * lane]; ey = bp[4 * lane + 1]; ez = bp[4 * lane + 2]; }
    const long long start = seg_start[s];
    const unsigned long long key = keys[start];
    const double sq_init = 10 * size * size;            // :413
    const double min_sq = min_dist * min_dist;          // :427
    int added = 0;
    for (long long j = start; j < n; ++j) {
        if (keys[j] != key) break;
        if (count >= cap) break;                        // IsFull(): nothing more is ever added (:411)
        const unsigned int i = idx[j];
        const float fx = fxyz[3 * (size_t)i], fy = fxyz[3 * (size_t)i + 1], fz = fxyz[3 * (size_t)i + 2];
        bool add;
        if (count == 0) {
            add = (min_num_points <= 0);                // absent voxel: created with the point (:437-445)
        } else {
            double sq = CUDART_INF;
            if (lane < count) {
                const double dx = __dsub_rn((double)ex, (double)fx), dy = __dsub_rn((double)ey, (double)fy), dz = __dsub_rn((double)ez, (double)fz);
                sq = __dadd_rn(__dmul_rn(dx, dx), __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz)));
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sq = fmin(sq, __shfl_xor_sync(0xffffffffu, sq, o));
            const double sq_min = fmin(sq_init, sq);
            add = (sq_min > min_sq) && (min_num_points <= 0 || count >= min_num_points);   // :427-433
        }
        if (add) {
            if (lane == count) { ex = fx; ey = fy; ez = fz; bp[4 * count] = fx; bp[4 * count + 1] = fy; bp[4 * count + 2] = fz; }
            ++count; ++added;
        }
    }
    if (lane == 0 && added) {
        slots[slot].count = (unsigned int)count;
        reinterpret_cast<unsigned int*>(bp)[kMetaCount] = (unsigned int)count;
        atomicAdd(reinterpret_cast<unsigned long long*>(n_points), (unsigned long long)added);
    }
}

// ---- N2: gridSampling / subSampleFrame (src/utility.cpp:167-201) ----------------------------------------------------
// cell key from the DOUBLE coordinate (src/utility.cpp:171-173: static_cast<short>(frame[i].point[k] / size_voxel))
__global__ void k_cell_keys(const double* __restrict__ xyz, long long n, double size, unsigned long long* keys, unsigned int* idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double qx = __ddiv_rn(xyz[3 * i], size), qy = __ddiv_rn(xyz[3 * i + 1], size), qz = __ddiv_rn(xyz[3 * i + 2], size);
    unsigned long long key = kInvalidKey;
    if (fabs(qx) < 32765.0 && fabs(qy) < 32765.0 && fabs(qz) < 32765.0) key = pack_key((int)qx, (int)qy, (int)qz);
    keys[i] = key;
    idx[i] = (unsigned int)i;
}
// after the stable sort by cell: the head of each run is the cell's first point in frame order; flag it at its own index
__global__ void k_first_in_cell(const unsigned long long* __restrict__ keys_sorted, const unsigned int* __restrict__ idx_sorted,
                                long long n, unsigned char* is_first) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long k = keys_sorted[j];
    is_first[idx_sorted[j]] = (k != kInvalidKey && (j == 0 || keys_sorted[j - 1] != k)) ? 1 : 0;
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace srl

using namespace srl;

extern "C" {

int srl_map_create(srl_ctx* ctx, double voxel_size, int32_t max_num_points_in_voxel, size_t max_voxels, srl_map** out) {
    if (!ctx || !out) return SRL_BAD_ARG;
    if (!(voxel_size > 0) || max_num_points_in_voxel < 1 || max_num_points_in_voxel > kBlockCap || max_voxels == 0)
        return set_err(ctx, SRL_BAD_ARG, "srl_map_create: voxel_size>0, 1<=max_num_points_in_voxel<=20, max_voxels>0 required");
    // the pass kernels address points as 32-bit float indices into the block pool (block * 80 + 4 * i)
    if (max_voxels > (size_t(1) << 25)) return set_err(ctx, SRL_BAD_ARG, "srl_map_create: max_voxels is limited to 2^25 (33.5 M voxels, 10.7 GB of blocks)");
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    srl_map* m = new srl_map();
    m->ctx = ctx; m->voxel_size = voxel_size; m->cap = max_num_points_in_voxel; m->max_voxels = max_voxels;
    size_t capacity = 1024;
    while (capacity < 2 * max_voxels) capacity <<= 1;
    m->capacity = capacity;
    cudaError_t e;
    if ((e = cudaMalloc(&m->d_slots, capacity * sizeof(Slot))) != cudaSuccess ||
        (e = cudaMalloc(&m->d_blocks, max_voxels * kBlockFloats * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc(&m->d_counters, 4 * sizeof(long long))) != cudaSuccess) {
        srl_map_destroy(m);
        return cuda_fail(ctx, e, "srl_map_create/cudaMalloc");
    }
    *out = m;
    return srl_map_clear(m);
}

void srl_map_destroy(srl_map* m) {
    if (!m) return;
    cudaFree(m->d_slots); cudaFree(m->d_blocks); cudaFree(m->d_counters);
    delete m;
}

int srl_map_clear(srl_map* m) {
    if (!m) return SRL_BAD_ARG;
    srl_ctx* ctx = m->ctx;
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    SRL_CUDA(ctx, cudaMemsetAsync(m->d_slots, 0, m->capacity * sizeof(Slot), ctx->stream));
    SRL_CUDA(ctx, cudaMemsetAsync(m->d_counters, 0, 4 * sizeof(long long), ctx->stream));
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    m->n_voxels = 0;
    return SRL_OK;
}

// ---- removePointsFarFromLocation (src/lioOptimization.cpp:556-572; row N4): a voxel goes when its FIRST point is farther
// than `distance` from `location`.  Open addressing has no cheap erase, and the block pool must stay dense (block index
// = slot payload), so eviction is mark -> compact the pool (stable: surviving blocks keep their relative order) ->
// rebuild the slot table from the keys kept in the blocks.
__global__ void k_far_flags(const float* __restrict__ blocks, long long n_voxels, double lx, double ly, double lz, double dist2,
                            unsigned* __restrict__ keep) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_voxels) return;
    const float* b = blocks + (size_t)v * kBlockFloats;
    const unsigned cnt = reinterpret_cast<const unsigned*>(b)[kMetaCount];
    // rgbPoint::getPosition() widens the stored floats; (pt - location).squaredNorm() reduces as x^2 + (y^2 + z^2)
    const double dx = __dsub_rn((double)b[0], lx), dy = __dsub_rn((double)b[1], ly), dz = __dsub_rn((double)b[2], lz);
    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz)));
    keep[v] = (cnt > 0 && !(d2 > dist2)) ? 1u : 0u;
}
__global__ void k_compact_blocks(const float* __restrict__ blocks, long long n_voxels, const unsigned* __restrict__ keep,
                                 const unsigned* __restrict__ new_index, float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 per thread
    const long long v = e / (kBlockFloats / 4);
    if (v >= n_voxels || !keep[v]) return;
    const int q = (int)(e % (kBlockFloats / 4));
    reinterpret_cast<float4*>(out + (size_t)new_index[v] * kBlockFloats)[q] = reinterpret_cast<const float4*>(blocks + (size_t)v * kBlockFloats)[q];
}
__global__ void k_rebuild_slots(Slot* slots, unsigned int mask, const float* __restrict__ blocks, long long n_voxels, long long* n_points) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_voxels) return;
    const unsigned int* meta = reinterpret_cast<const unsigned int*>(blocks + (size_t)v * kBlockFloats);
    const unsigned long long key = (unsigned long long)meta[kMetaKeyLo] | ((unsigned long long)meta[kMetaKeyHi] << 32);
    short x, y, z;
    unpack_key(key, x, y, z);
    slot_claim(slots, mask, key, x, y, z, (unsigned)v, meta[kMetaCount]);
    atomicAdd(reinterpret_cast<unsigned long long*>(n_points), (unsigned long long)meta[kMetaCount]);
}

int srl_map_remove_far(srl_map* m, const double location[3], double distance, int64_t* n_removed) {
    if (!m || !location) return SRL_BAD_ARG;
    srl_ctx* ctx = m->ctx;
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_removed) *n_removed = 0;
    const long long nv = (long long)m->n_voxels;
    if (nv == 0) return SRL_OK;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (unsigned*)nullptr, (unsigned*)nullptr, (int)nv, ctx->stream);
    // worst case every block survives: the compacted copy needs a pool-sized scratch
    const size_t need = 2 * al((size_t)nv * 4) + al(tmp) + al((size_t)nv * kBlockFloats * sizeof(float)) + 256;
    int rc = ensure_scratch(ctx, need);
    if (rc != SRL_OK) return rc;
    char* p = static_cast<char*>(ctx->d_scratch);
    unsigned* keep = reinterpret_cast<unsigned*>(p); p += al((size_t)nv * 4);
    unsigned* nidx = reinterpret_cast<unsigned*>(p); p += al((size_t)nv * 4);
    void* cub_tmp = p; p += al(tmp);
    float* pool = reinterpret_cast<float*>(p);
    const int T = 256;
    k_far_flags<<<(unsigned)((nv + T - 1) / T), T, 0, ctx->stream>>>(m->d_blocks, nv, location[0], location[1], location[2], distance * distance, keep);
    SRL_CUDA(ctx, cudaGetLastError());
    SRL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, tmp, keep, nidx, (int)nv, ctx->stream));
    unsigned last_keep = 0, last_idx = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&last_keep, keep + (nv - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(&last_idx, nidx + (nv - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const long long n_keep = (long long)last_idx + last_keep;
    ctx->launches += 2;
    if (n_keep == nv) return SRL_OK;   // nothing to evict
    const long long n_quads = nv * (kBlockFloats / 4);
    k_compact_blocks<<<(unsigned)((n_quads + T - 1) / T), T, 0, ctx->stream>>>(m->d_blocks, nv, keep, nidx, pool);
    SRL_CUDA(ctx, cudaGetLastError());
    if (n_keep > 0) SRL_CUDA(ctx, cudaMemcpyAsync(m->d_blocks, pool, (size_t)n_keep * kBlockFloats * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    SRL_CUDA(ctx, cudaMemsetAsync(m->d_slots, 0, m->capacity * sizeof(Slot), ctx->stream));
    SRL_CUDA(ctx, cudaMemsetAsync(m->d_counters, 0, sizeof(long long), ctx->stream));
    if (n_keep > 0) {
        k_rebuild_slots<<<(unsigned)((n_keep + T - 1) / T), T, 0, ctx->stream>>>(m->d_slots, (unsigned)(m->capacity - 1), m->d_blocks, n_keep, m->d_counters);
        SRL_CUDA(ctx, cudaGetLastError());
    }
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += 2;
    m->n_voxels = n_keep;
    if (n_removed) *n_removed = nv - n_keep;
    return SRL_OK;
}

int srl_map_stats(srl_map* m, int64_t* n_voxels, int64_t* n_points) {
    if (!m) return SRL_BAD_ARG;
    srl_ctx* ctx = m->ctx;
    long long np = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&np, m->d_counters, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_voxels) *n_voxels = m->n_voxels;
    if (n_points) *n_points = np;
    return SRL_OK;
}

int srl_map_upload(srl_map* m, const int16_t* keys, const int32_t* counts, const float* xyz, size_t n_voxels) {
    if (!m || (n_voxels && (!keys || !counts || !xyz))) return SRL_BAD_ARG;
    srl_ctx* ctx = m->ctx;
    if (n_voxels > m->max_voxels) return set_err(ctx, SRL_MAP_FULL, "srl_map_upload: more voxels than max_voxels");
    int rc = srl_map_clear(m);
    if (rc != SRL_OK || n_voxels == 0) return rc;
    const int cap = m->cap;
    long long total_pts = 0;
    for (size_t v = 0; v < n_voxels; ++v) {
        if (counts[v] < 0 || counts[v] > cap) return set_err(ctx, SRL_BAD_ARG, "srl_map_upload: count outside [0, cap]");
        total_pts += counts[v];
    }
    const size_t b_keys = align_up(n_voxels * 3 * sizeof(short)), b_cnt = align_up(n_voxels * sizeof(int)),
                 b_xyz = align_up(n_voxels * cap * 3 * sizeof(float));
    if ((rc = ensure_scratch(ctx, b_keys + b_cnt + b_xyz + 256)) != SRL_OK) return rc;
    char* base = static_cast<char*>(ctx->d_scratch);
    short* d_keys = reinterpret_cast<short*>(base);
    int* d_cnt = reinterpret_cast<int*>(base + b_keys);
    float* d_xyz = reinterpret_cast<float*>(base + b_keys + b_cnt);
    int* d_dup = reinterpret_cast<int*>(base + b_keys + b_cnt + b_xyz);
    SRL_CUDA(ctx, cudaMemcpyAsync(d_keys, keys, n_voxels * 3 * sizeof(short), cudaMemcpyHostToDevice, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(d_cnt, counts, n_voxels * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(d_xyz, xyz, n_voxels * cap * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SRL_CUDA(ctx, cudaMemsetAsync(d_dup, 0, sizeof(int), ctx->stream));
    const int T = 256;
    k_upload_slots<<<(unsigned)((n_voxels + T - 1) / T), T, 0, ctx->stream>>>(m->d_slots, (unsigned)(m->capacity - 1), m->d_blocks,
                                                                             d_keys, d_cnt, (long long)n_voxels, d_dup);
    k_upload_points<<<(unsigned)((n_voxels * cap + T - 1) / T), T, 0, ctx->stream>>>(m->d_blocks, d_xyz, (long long)n_voxels, cap);
    ctx->launches += 2;
    SRL_CUDA(ctx, cudaGetLastError());
    int dup = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&dup, d_dup, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(m->d_counters, &total_pts, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (dup) return set_err(ctx, SRL_BAD_ARG, "srl_map_upload: duplicate voxel keys");
    m->n_voxels = (int64_t)n_voxels;
    return SRL_OK;
}

int srl_map_download(srl_map* m, int16_t* keys, int32_t* counts, float* xyz, size_t max_voxels, int64_t* n_voxels) {
    if (!m) return SRL_BAD_ARG;
    srl_ctx* ctx = m->ctx;
    const size_t nv = (size_t)m->n_voxels;
    if (n_voxels) *n_voxels = (int64_t)nv;
    if (nv == 0) return SRL_OK;
    if (nv > max_voxels || !keys || !counts || !xyz) return set_err(ctx, SRL_BAD_ARG, "srl_map_download: output too small");
    const int cap = m->cap;
    const size_t b_keys = align_up(nv * 3 * sizeof(short)), b_cnt = align_up(nv * sizeof(int)), b_xyz = align_up(nv * cap * 3 * sizeof(float));
    int rc;
    if ((rc = ensure_scratch(ctx, b_keys + b_cnt + b_xyz)) != SRL_OK) return rc;
    char* base = static_cast<char*>(ctx->d_scratch);
    short* d_keys = reinterpret_cast<short*>(base);
    int* d_cnt = reinterpret_cast<int*>(base + b_keys);
    float* d_xyz = reinterpret_cast<float*>(base + b_keys + b_cnt);
    const int T = 256;
    k_download<<<(unsigned)((nv * cap + T - 1) / T), T, 0, ctx->stream>>>(m->d_blocks, (long long)nv, cap, d_keys, d_cnt, d_xyz);
    ctx->launches += 1;
    SRL_CUDA(ctx, cudaGetLastError());
    SRL_CUDA(ctx, cudaMemcpyAsync(keys, d_keys, nv * 3 * sizeof(short), cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(counts, d_cnt, nv * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(xyz, d_xyz, nv * cap * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SRL_OK;
}

static int map_insert_impl(srl_map* m, const double* d_xyz, size_t n, double min_distance_points, int32_t min_num_points,
                           int64_t* n_added, char* scratch_after_points) {
    srl_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    // scratch carve-up
    char* p = scratch_after_points;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
    unsigned long long* keys_a = reinterpret_cast<unsigned long long*>(take(n * 8));
    unsigned long long* keys_b = reinterpret_cast<unsigned long long*>(take(n * 8));
    unsigned int* idx_a = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned int* idx_b = reinterpret_cast<unsigned int*>(take(n * 4));
    float* fxyz = reinterpret_cast<float*>(take(n * 12));
    unsigned char* flags = reinterpret_cast<unsigned char*>(take(n));
    unsigned int* seg_start = reinterpret_cast<unsigned int*>(take(n * 4));
    int* seg_slot = reinterpret_cast<int*>(take(n * 4));
    unsigned int* is_new = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned int* new_rank = reinterpret_cast<unsigned int*>(take(n * 4));
    int* d_nseg = reinterpret_cast<int*>(take(256));
    unsigned int* d_total_new = reinterpret_cast<unsigned int*>(take(256));
    size_t tmp_sort = 0, tmp_sel = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 50, st);
    cub::DeviceSelect::Flagged(nullptr, tmp_sel, thrust::counting_iterator<unsigned int>(0), flags, seg_start, d_nseg, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, is_new, new_rank, (int)n, st);
    const size_t tmp_bytes = std::max(tmp_sort, std::max(tmp_sel, tmp_scan));
    void* d_tmp = take(tmp_bytes);

    const int T = 256;
    const unsigned gb = (unsigned)((n + T - 1) / T);
    const unsigned mask = (unsigned)(m->capacity - 1);
    k_insert_keys<<<gb, T, 0, st>>>(d_xyz, (long long)n, m->voxel_size, keys_a, idx_a, fxyz);
    size_t tb = tmp_bytes;
    cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 50, st);
    k_seg_flags<<<gb, T, 0, st>>>(keys_b, (long long)n, flags);
    tb = tmp_bytes;
    cub::DeviceSelect::Flagged(d_tmp, tb, thrust::counting_iterator<unsigned int>(0), flags, seg_start, d_nseg, (int)n, st);
    int n_seg = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&n_seg, d_nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRL_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 4;
    if (n_seg == 0) { if (n_added) *n_added = 0; return SRL_OK; }
    const unsigned gs = (unsigned)((n_seg + T - 1) / T);
    k_seg_lookup<<<gs, T, 0, st>>>(m->d_slots, mask, keys_b, seg_start, d_nseg, min_num_points <= 0 ? 1 : 0, seg_slot, is_new);
    tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(d_tmp, tb, is_new, new_rank, n_seg, st);
    // total new voxels = new_rank[last] + is_new[last]
    unsigned int last_rank = 0, last_new = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&last_rank, new_rank + (n_seg - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SRL_CUDA(ctx, cudaMemcpyAsync(&last_new, is_new + (n_seg - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SRL_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 2;
    const long long total_new = (long long)last_rank + last_new;
    (void)d_total_new;
    if ((size_t)(m->n_voxels + total_new) > m->max_voxels)
        return set_err(ctx, SRL_MAP_FULL, "srl_map_insert: voxel pool exhausted (raise max_voxels)");
    long long before = 0, after = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&before, m->d_counters, sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (total_new > 0) {
        k_seg_claim<<<gs, T, 0, st>>>(m->d_slots, mask, m->d_blocks, keys_b, seg_start, d_nseg, is_new, new_rank,
                                      (long long)m->n_voxels, seg_slot);
        ctx->launches += 1;
    }
    const unsigned gw = (unsigned)(((long long)n_seg * 32 + T - 1) / T);
    k_seg_process<<<gw, T, 0, st>>>(m->d_slots, m->d_blocks, keys_b, idx_b, fxyz, seg_start, d_nseg, seg_slot, (long long)n,
                                    m->voxel_size, m->cap, min_distance_points, min_num_points, m->d_counters);
    ctx->launches += 1;
    SRL_CUDA(ctx, cudaGetLastError());
    SRL_CUDA(ctx, cudaMemcpyAsync(&after, m->d_counters, sizeof(long long), cudaMemcpyDeviceToHost, st));
    SRL_CUDA(ctx, cudaStreamSynchronize(st));
    m->n_voxels += total_new;
    if (n_added) *n_added = after - before;
    return SRL_OK;
}

static size_t insert_scratch_bytes(size_t n) {
    // generous upper bound: arrays + CUB temp (radix sort of 64-bit keys needs ~ (8+4)*n + small)
    return n * (8 + 8 + 4 + 4 + 12 + 1 + 4 + 4 + 4 + 4) + n * 16 + (1u << 20) + 16 * 256;
}

namespace {
struct CellKey { short x, y, z; bool operator==(const CellKey& o) const { return x == o.x && y == o.y && z == o.z; } };
struct CellHash {   // std::hash<voxel> of the reference (include/cloudMap.h:173-184)
    std::size_t operator()(const CellKey& v) const {
        const std::size_t kP1 = 73856093, kP2 = 19349669, kP3 = 83492791;
        return v.x * kP1 + v.y * kP2 + v.z * kP3;
    }
};
}  // namespace

int srl_grid_sampling(srl_ctx* ctx, const double* xyz_world, size_t n, double size, uint32_t* out, size_t* n_out) {
    if (!ctx || (n && (!xyz_world || !out)) || !n_out || !(size > 0)) return SRL_BAD_ARG;
    *n_out = 0;
    if (n == 0) return SRL_OK;
    if (n > 0x7fffffffULL) return set_err(ctx, SRL_BAD_ARG, "srl_grid_sampling: n must fit in int32");
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t tmp_sort = 0, tmp_sel = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr,
                                    (unsigned int*)nullptr, (int)n, 0, 50, st);
    cub::DeviceSelect::Flagged(nullptr, tmp_sel, thrust::counting_iterator<unsigned int>(0), (unsigned char*)nullptr,
                               (unsigned int*)nullptr, (int*)nullptr, (int)n, st);
    const size_t tmp_bytes = std::max(tmp_sort, tmp_sel);
    const size_t need = align_up(n * 24) + 2 * align_up(n * 8) + 3 * align_up(n * 4) + align_up(n) + 256 + align_up(tmp_bytes);
    int rc = ensure_scratch(ctx, need);
    if (rc != SRL_OK) return rc;
    char* p = static_cast<char*>(ctx->d_scratch);
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
    double* d_xyz = reinterpret_cast<double*>(take(n * 24));
    unsigned long long* ka = reinterpret_cast<unsigned long long*>(take(n * 8));
    unsigned long long* kb = reinterpret_cast<unsigned long long*>(take(n * 8));
    unsigned int* ia = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned int* ib = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned int* sel = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned char* is_first = reinterpret_cast<unsigned char*>(take(n));
    int* d_count = reinterpret_cast<int*>(take(256));
    void* d_tmp = take(tmp_bytes);
    SRL_CUDA(ctx, cudaMemcpyAsync(d_xyz, xyz_world, n * 24, cudaMemcpyHostToDevice, st));
    const int T = 256;
    const unsigned gb = (unsigned)((n + T - 1) / T);
    k_cell_keys<<<gb, T, 0, st>>>(d_xyz, (long long)n, size, ka, ia);
    size_t tb = tmp_bytes;
    cub::DeviceRadixSort::SortPairs(d_tmp, tb, ka, kb, ia, ib, (int)n, 0, 50, st);
    k_first_in_cell<<<gb, T, 0, st>>>(kb, ib, (long long)n, is_first);
    tb = tmp_bytes;
    cub::DeviceSelect::Flagged(d_tmp, tb, thrust::counting_iterator<unsigned int>(0), is_first, sel, d_count, (int)n, st);
    ctx->launches += 2;
    SRL_CUDA(ctx, cudaGetLastError());
    int m = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&m, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    SRL_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<unsigned int> first((size_t)m);
    if (m) SRL_CUDA(ctx, cudaMemcpy(first.data(), sel, (size_t)m * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    // `first` = the frame indices that open a new cell, in frame order: exactly the sequence of node insertions the
    // reference's grid sees (later points of a cell only push_back into an existing node).  Replaying it through the same
    // libstdc++ container gives the reference's iteration order (src/utility.cpp:180-187).
    std::tr1::unordered_map<CellKey, unsigned int, CellHash> grid;
    for (int j = 0; j < m; ++j) {
        const unsigned int i = first[(size_t)j];
        CellKey k;
        k.x = static_cast<short>(xyz_world[3 * (size_t)i] / size);
        k.y = static_cast<short>(xyz_world[3 * (size_t)i + 1] / size);
        k.z = static_cast<short>(xyz_world[3 * (size_t)i + 2] / size);
        grid[k] = i;
    }
    size_t w = 0;
    for (std::tr1::unordered_map<CellKey, unsigned int, CellHash>::const_iterator it = grid.begin(); it != grid.end(); ++it) out[w++] = it->second;
    *n_out = w;
    return SRL_OK;
}

int srl_map_insert_device(srl_map* m, const double* d_xyz_world, size_t n, double min_distance_points, int32_t min_num_points,
                          int64_t* n_added) {
    if (!m || (n && !d_xyz_world)) return SRL_BAD_ARG;
    if (n_added) *n_added = 0;
    if (n == 0) return SRL_OK;
    srl_ctx* ctx = m->ctx;
    if (n > 0x7fffffffULL) return set_err(ctx, SRL_BAD_ARG, "srl_map_insert: n must fit in int32");
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_scratch(ctx, insert_scratch_bytes(n))) != SRL_OK) return rc;
    return map_insert_impl(m, d_xyz_world, n, min_distance_points, min_num_points, n_added, static_cast<char*>(ctx->d_scratch));
}

int srl_map_insert_sweep(srl_map* m, srl_sweep* sw, const double q[4], const double t[3], const double R_il[9], const double t_il[3],
                         double min_distance_points, int32_t min_num_points, int64_t* n_added) {
    if (!m || !sw || !q || !t || !R_il || !t_il) return SRL_BAD_ARG;
    if (n_added) *n_added = 0;
    srl_ctx* ctx = m->ctx;
    if (sw->ctx != ctx) return set_err(ctx, SRL_BAD_ARG, "map and sweep belong to different contexts");
    const size_t n = sw->n;
    if (n == 0) return SRL_OK;
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pts_bytes = align_up(n * 3 * sizeof(double));
    int rc;
    if ((rc = ensure_scratch(ctx, pts_bytes + insert_scratch_bytes(n))) != SRL_OK) return rc;
    char* base = static_cast<char*>(ctx->d_scratch);
    if ((rc = srl_sweep_transform_device(ctx, sw, q, t, R_il, t_il, reinterpret_cast<double*>(base))) != SRL_OK) return rc;
    return map_insert_impl(m, reinterpret_cast<const double*>(base), n, min_distance_points, min_num_points, n_added, base + pts_bytes);
}

int srl_map_insert(srl_map* m, const double* xyz_world, size_t n, double min_distance_points, int32_t min_num_points,
                   int64_t* n_added) {
    if (!m || (n && !xyz_world)) return SRL_BAD_ARG;
    if (n_added) *n_added = 0;
    if (n == 0) return SRL_OK;
    srl_ctx* ctx = m->ctx;
    if (n > 0x7fffffffULL) return set_err(ctx, SRL_BAD_ARG, "srl_map_insert: n must fit in int32");
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pts_bytes = align_up(n * 3 * sizeof(double));
    int rc;
    if ((rc = ensure_scratch(ctx, pts_bytes + insert_scratch_bytes(n))) != SRL_OK) return rc;
    char* base = static_cast<char*>(ctx->d_scratch);
    SRL_CUDA(ctx, cudaMemcpyAsync(base, xyz_world, n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return map_insert_impl(m, reinterpret_cast<const double*>(base), n, min_distance_points, min_num_points, n_added, base + pts_bytes);
}

}  // extern "C"
