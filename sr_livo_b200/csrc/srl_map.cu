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

// ---- N4: colour map (src/lioOptimization.cpp:448-551 colour branch; src/rgbMapTracker.cpp:181-237; src/cloudMap.cpp:59-101) ----
// The colour map is a second voxel map (same slot table / block pool) whose points carry a colour estimate, plus
//   * a fine occupancy set (cells of min_distance_points, the reference's Hash_map_3d hashmap_3d_points) that decides
//     which stored points also enter rgb_points_vec, and
//   * the list of voxels first visited by the current sweep(s) (voxels_recent_visited_temp), which the renderer walks.
// addPointToColorMap is sequential in the reference; the same (sort by voxel, replay per voxel in sweep order) scheme as K3
// reproduces it: a voxel accepts the first (cap - count) offered points, a fine cell is claimed by the first ACCEPTED
// point of the sweep that falls into it, and both lists are emitted in sweep order.
struct ColorPoint {            // colour state of one stored point (rgbPoint minus position), index = block * cap + i
    short rgb[3]; short n_rgb;
    float cov[3]; float pad;
    double obs_dist, last_obs;
};
constexpr unsigned kNoIndex = 0xffffffffu;

// selected points (every step-th of the frame): voxel key + fine key from the float-rounded position (getPosition())
__global__ void k_color_keys(const double* __restrict__ xyz, long long m, int step, double size, double fine, unsigned long long* vkeys,
                             unsigned long long* fkeys, unsigned int* idx, float* fxyz) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const long long i = j * step;
    const float fx = __double2float_rn(xyz[3 * i]), fy = __double2float_rn(xyz[3 * i + 1]), fz = __double2float_rn(xyz[3 * i + 2]);
    fxyz[3 * j] = fx; fxyz[3 * j + 1] = fy; fxyz[3 * j + 2] = fz;
    const double qx = __ddiv_rn((double)fx, size), qy = __ddiv_rn((double)fy, size), qz = __ddiv_rn((double)fz, size);
    const double gx = __ddiv_rn((double)fx, fine), gy = __ddiv_rn((double)fy, fine), gz = __ddiv_rn((double)fz, fine);
    const bool ok = fabs(qx) < 32765.0 && fabs(qy) < 32765.0 && fabs(qz) < 32765.0 && fabs(gx) < 32765.0 && fabs(gy) < 32765.0 && fabs(gz) < 32765.0;
    vkeys[j] = ok ? pack_key((int)qx, (int)qy, (int)qz) : kInvalidKey;
    fkeys[j] = ok ? pack_key((int)gx, (int)gy, (int)gz) : kInvalidKey;
    idx[j] = (unsigned int)j;
}

// one thread per touched voxel: append the first (cap - count) offered points in sweep order, note the visit
__global__ void k_color_seg_process(Slot* slots, float* blocks, ColorPoint* cpts, double* last_visited,
                                    const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idx,
                                    const float* __restrict__ fxyz, const unsigned int* __restrict__ seg_start, const int* n_seg_p,
                                    const int* __restrict__ seg_slot, const unsigned int* __restrict__ is_new, long long m, int cap,
                                    double t_end, double t_last_process, unsigned int* accept_id, unsigned int* seg_first,
                                    long long* n_points) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= *n_seg_p) return;
    seg_first[s] = kNoIndex;
    const int slot = seg_slot[s];
    if (slot < 0) return;
    const unsigned int blk = slots[slot].block;
    int count = (int)slots[slot].count;
    float* bp = blocks + (size_t)blk * kBlockFloats;
    const long long start = seg_start[s];
    const unsigned long long key = keys[start];
    if (is_new[s]) last_visited[blk] = 0.0;                       // voxelBlock::last_visited_time = 0.0 (include/cloudMap.h:153)
    int added = 0;
    for (long long j = start; j < m && keys[j] == key && count < cap; ++j) {
        const unsigned int i = idx[j];
        bp[4 * count] = fxyz[3 * (size_t)i]; bp[4 * count + 1] = fxyz[3 * (size_t)i + 1]; bp[4 * count + 2] = fxyz[3 * (size_t)i + 2];
        ColorPoint cp;                                            // rgbPoint::reset() (src/cloudMap.cpp:12-19)
        cp.rgb[0] = cp.rgb[1] = cp.rgb[2] = 0; cp.n_rgb = 0; cp.cov[0] = cp.cov[1] = cp.cov[2] = 0.f; cp.pad = 0.f; cp.obs_dist = 0.0; cp.last_obs = 0.0;
        cpts[(size_t)blk * kBlockCap + count] = cp;
        accept_id[i] = blk * (unsigned)kBlockCap + (unsigned)count;
        ++count; ++added;
    }
    if (added) {
        slots[slot].count = (unsigned int)count;
        reinterpret_cast<unsigned int*>(bp)[kMetaCount] = (unsigned int)count;
        atomicAdd(reinterpret_cast<unsigned long long*>(n_points), (unsigned long long)added);
    }
    // :486-490 / :509-513: once per voxel and sweep end time, whether or not a point was stored
    if (fabs(t_end - t_last_process) > 1e-5 && fabs(last_visited[blk] - t_end) > 1e-5) {
        last_visited[blk] = t_end;
        seg_first[s] = idx[start];                                // the sweep position of the voxel's first offered point
    }
}
__global__ void k_color_seg_keys(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ seg_start, const int* n_seg_p,
                                 unsigned long long* seg_key) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < *n_seg_p) seg_key[s] = keys[seg_start[s]];
}
// after sorting the visited voxels by the sweep position of their first point: append them to the recent list
__global__ void k_color_append_recent(const unsigned int* __restrict__ first_sorted, const unsigned long long* __restrict__ key_sorted,
                                      int n_seg, unsigned long long* recent, long long* counters, long long capacity) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_seg || first_sorted[t] == kNoIndex) return;
    const long long base = counters[1];
    if (base + t < capacity) recent[base + t] = key_sorted[t];
    if (t + 1 == n_seg || first_sorted[t + 1] == kNoIndex) counters[3] = t + 1;   // how many were appended (applied by the host)
}
// accepted points only keep their fine key
__global__ void k_color_mask_fine(unsigned long long* fkeys, const unsigned int* __restrict__ accept_id, long long m) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m && accept_id[j] == kNoIndex) fkeys[j] = kInvalidKey;
}
// heads of the fine-cell runs (= the first accepted point of the sweep in that cell) whose cell is still free win
__global__ void k_color_fine_winners(const Slot* fine, unsigned int fmask, const unsigned long long* __restrict__ fk_sorted,
                                     const unsigned int* __restrict__ idx_sorted, long long m, unsigned int* winner) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const unsigned long long k = fk_sorted[j];
    if (k == kInvalidKey || (j > 0 && fk_sorted[j - 1] == k)) return;
    short x, y, z;
    unpack_key(k, x, y, z);
    if (slot_find_rw(fine, fmask, k, x, y, z) < 0) winner[idx_sorted[j]] = 1u;
}
__global__ void k_color_emit_rgb(Slot* fine, unsigned int fmask, const unsigned long long* __restrict__ fkeys_by_idx,
                                 const unsigned int* __restrict__ winner, const unsigned int* __restrict__ rank,
                                 const unsigned int* __restrict__ accept_id, long long m, unsigned int* rgb_points, long long* counters,
                                 long long capacity) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m || !winner[j]) return;
    const long long pos = counters[0] + rank[j];
    if (pos < capacity) rgb_points[pos] = accept_id[j];            // point.point_index = rgb_points_vec.size() (:478, :503)
    const unsigned long long k = fkeys_by_idx[j];
    short x, y, z;
    unpack_key(k, x, y, z);
    slot_claim(fine, fmask, k, x, y, z, (unsigned)pos, 1u);       // hashmap_3d_points.insert (:481, :506)
}

struct CamConst { double R[9], t_cw[3], t_wc[3], fx, fy, cx, cy, fov; int cols, rows; };
__device__ __forceinline__ unsigned char sat_u8(double v) {       // cv::saturate_cast<uchar>(double): cvRound (half to even), clamp
    const long long r = __double2ll_rn(v);
    return (unsigned char)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
__device__ __forceinline__ unsigned char sat_add_u8(unsigned char a, unsigned char b) { const int r = (int)a + (int)b; return (unsigned char)(r > 255 ? 255 : r); }
// one warp per distinct recent voxel, a lane per stored point; a voxel listed `mult` times is rendered `mult` times in a row
__global__ void __launch_bounds__(256) k_color_render(const Slot* slots, unsigned int mask, const float* __restrict__ blocks, ColorPoint* cpts,
                                                       const unsigned long long* __restrict__ uniq_keys, const int* __restrict__ mult,
                                                       const int* n_uniq_p, CamConst c, const unsigned char* __restrict__ img, double obs_time,
                                                       unsigned long long* n_rendered) {
    const int lane = threadIdx.x & 31;
    const int u_i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (u_i >= *n_uniq_p) return;
    const unsigned long long key = uniq_keys[u_i];
    short kx, ky, kz;
    unpack_key(key, kx, ky, kz);
    const int slot = slot_find_rw(slots, mask, key, kx, ky, kz);
    if (slot < 0) return;
    const unsigned blk = slots[slot].block;
    const int count = (int)slots[slot].count;
    if (lane >= count) return;
    const float* bp = blocks + (size_t)blk * kBlockFloats + 4 * lane;
    const double px = (double)bp[0], py = (double)bp[1], pz = (double)bp[2];
    // project3dTo2d (src/lioOptimization.cpp:132-152), scale 1: products and sums rounded one by one like the host code
    double cxm, cym, czm;
    matvec3_exact(c.R, px, py, pz, cxm, cym, czm);
    const double pcx = __dadd_rn(cxm, c.t_cw[0]), pcy = __dadd_rn(cym, c.t_cw[1]), pcz = __dadd_rn(czm, c.t_cw[2]);
    if (pcz < 0.001) return;
    const double u = __dadd_rn(__ddiv_rn(__dmul_rn(pcx, c.fx), pcz), c.cx);
    const double v = __dadd_rn(__ddiv_rn(__dmul_rn(pcy, c.fy), pcz), c.cy);
    // if2dPointsAvailable (:49-60)
    if (!((u >= __dadd_rn(__dmul_rn(c.fov, (double)c.cols), 1.0)) && (ceil(u) < __dmul_rn(__dsub_rn(1.0, c.fov), (double)c.cols)) &&
          (v >= __dadd_rn(__dmul_rn(c.fov, (double)c.rows), 1.0)) && (ceil(v) < __dmul_rn(__dsub_rn(1.0, c.fov), (double)c.rows)))) return;
    const double dx = __dsub_rn(px, c.t_wc[0]), dy = __dsub_rn(py, c.t_wc[1]), dz = __dsub_rn(pz, c.t_wc[2]);
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz))));
    // getSubPixel<cv::Vec3b> (:71-98): four saturated products, three saturated sums per channel
    const int fr = (int)floor(v), fc = (int)floor(u);
    const double frac_r = __dsub_rn(v, (double)fr), frac_c = __dsub_rn(u, (double)fc);
    const double w00 = __dmul_rn(__dsub_rn(1.0, frac_r), __dsub_rn(1.0, frac_c)), w10 = __dmul_rn(frac_r, __dsub_rn(1.0, frac_c));
    const double w01 = __dmul_rn(__dsub_rn(1.0, frac_r), frac_c), w11 = __dmul_rn(frac_r, frac_c);
    double color[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const unsigned char a = sat_u8(__dmul_rn((double)img[((size_t)fr * c.cols + fc) * 3 + ch], w00));
        const unsigned char b = sat_u8(__dmul_rn((double)img[((size_t)(fr + 1) * c.cols + fc) * 3 + ch], w10));
        const unsigned char cc = sat_u8(__dmul_rn((double)img[((size_t)fr * c.cols + fc + 1) * 3 + ch], w01));
        const unsigned char d = sat_u8(__dmul_rn((double)img[((size_t)(fr + 1) * c.cols + fc + 1) * 3 + ch], w11));
        color[ch] = (double)sat_add_u8(sat_add_u8(sat_add_u8(a, b), cc), d);
    }
    // rgbPoint::updateRgb (src/cloudMap.cpp:59-101), mixed float / double arithmetic as written there
    ColorPoint cp = cpts[(size_t)blk * kBlockCap + lane];
    const double sigma = 15.0, process_noise_sigma = 0.1;
    unsigned rendered = 0;
    const int reps = mult[u_i];
    for (int rep = 0; rep < reps; ++rep) {
        if (cp.obs_dist != 0 && (dist > __dmul_rn(cp.obs_dist, 1.2))) continue;
        if (cp.n_rgb == 0) {
            cp.last_obs = obs_time; cp.obs_dist = dist;
#pragma unroll
            for (int i = 0; i < 3; ++i) { cp.rgb[i] = (short)round(color[i]); cp.cov[i] = (float)sigma; }
            cp.n_rgb = 1;
            continue;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            cp.cov[i] = __double2float_rn(__dadd_rn((double)cp.cov[i], __dmul_rn(process_noise_sigma, __dsub_rn(obs_time, cp.last_obs))));
            const double old_sigma = (double)cp.cov[i];
            const double c2 = (double)__fmul_rn(cp.cov[i], cp.cov[i]);
            cp.cov[i] = __double2float_rn(__dsqrt_rn(__ddiv_rn(1.0, __dadd_rn(__ddiv_rn(1.0, c2), __ddiv_rn(1.0, __dmul_rn(sigma, sigma))))));
            const double n2 = (double)__fmul_rn(cp.cov[i], cp.cov[i]);
            const double mix = __dadd_rn(__ddiv_rn((double)cp.rgb[i], __dmul_rn(old_sigma, old_sigma)), __ddiv_rn(color[i], __dmul_rn(sigma, sigma)));
            cp.rgb[i] = (short)(int)__dmul_rn(n2, mix);           // double -> short truncates
        }
        if (dist < cp.obs_dist) cp.obs_dist = dist;
        cp.last_obs = obs_time;
        cp.n_rgb = (short)(cp.n_rgb + 1);
        ++rendered;
    }
    cpts[(size_t)blk * kBlockCap + lane] = cp;
    if (rendered) atomicAdd(n_rendered, (unsigned long long)rendered);
}
// colour state + last visited time in the block order of srl_map_download
__global__ void k_color_download(const ColorPoint* __restrict__ cpts, const float* __restrict__ blocks, const double* __restrict__ last_visited,
                                 long long n_voxels, int cap, short* rgb, short* n_rgb, float* cov, double* obs_dist, double* last_obs,
                                 double* visited) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_voxels * cap) return;
    const long long v = e / cap;
    const int i = (int)(e % cap);
    const int cnt = (int)reinterpret_cast<const unsigned int*>(blocks + (size_t)v * kBlockFloats)[kMetaCount];
    const bool on = i < cnt;
    const ColorPoint cp = cpts[(size_t)v * kBlockCap + i];
    for (int a = 0; a < 3; ++a) { rgb[3 * e + a] = on ? cp.rgb[a] : (short)0; cov[3 * e + a] = (on && cp.n_rgb > 0) ? cp.cov[a] : 0.f; }
    n_rgb[e] = on ? cp.n_rgb : (short)0;
    obs_dist[e] = on ? cp.obs_dist : 0.0;
    last_obs[e] = on ? cp.last_obs : 0.0;
    if (i == 0) visited[v] = last_visited[v];
}
// rgb_points_vec entries as (voxel key, index in block)
__global__ void k_color_rgb_ids(const unsigned int* __restrict__ rgb_points, long long n, const float* __restrict__ blocks, int cap, short* out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned id = rgb_points[t];
    const unsigned blk = id / (unsigned)kBlockCap;
    const unsigned int* meta = reinterpret_cast<const unsigned int*>(blocks + (size_t)blk * kBlockFloats);
    short x, y, z;
    unpack_key((unsigned long long)meta[kMetaKeyLo] | ((unsigned long long)meta[kMetaKeyHi] << 32), x, y, z);
    out[4 * t] = x; out[4 * t + 1] = y; out[4 * t + 2] = z; out[4 * t + 3] = (short)(id - blk * (unsigned)kBlockCap);
}

__global__ void k_gather_points(const double* __restrict__ xyz, const unsigned int* __restrict__ sel, int m, double* out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const size_t i = sel[j];
    out[3 * j] = xyz[3 * i]; out[3 * j + 1] = xyz[3 * i + 1]; out[3 * j + 2] = xyz[3 * i + 2];
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
    // the frame may already be in HBM (e.g. the output of srl_distort_frame_* / srl_sweep_transform_device): no H2D then
    cudaPointerAttributes attr;
    const bool on_device = cudaPointerGetAttributes(&attr, xyz_world) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    if (!on_device) cudaGetLastError();
    double* d_stage = reinterpret_cast<double*>(take(n * 24));
    const double* d_xyz = on_device ? xyz_world : d_stage;
    unsigned long long* ka = reinterpret_cast<unsigned long long*>(take(n * 8));
    unsigned long long* kb = reinterpret_cast<unsigned long long*>(take(n * 8));
    unsigned int* ia = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned int* ib = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned int* sel = reinterpret_cast<unsigned int*>(take(n * 4));
    unsigned char* is_first = reinterpret_cast<unsigned char*>(take(n));
    int* d_count = reinterpret_cast<int*>(take(256));
    void* d_tmp = take(tmp_bytes);
    if (!on_device) SRL_CUDA(ctx, cudaMemcpyAsync(d_stage, xyz_world, n * 24, cudaMemcpyHostToDevice, st));
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
    std::vector<double> first_xyz;                       // device input: the coordinates of those points come back, not the frame
    if (on_device && m) {
        k_gather_points<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(d_xyz, sel, m, d_stage);
        SRL_CUDA(ctx, cudaGetLastError());
        first_xyz.resize((size_t)m * 3);
        SRL_CUDA(ctx, cudaMemcpyAsync(first_xyz.data(), d_stage, (size_t)m * 24, cudaMemcpyDeviceToHost, st));
        SRL_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // `first` = the frame indices that open a new cell, in frame order: exactly the sequence of node insertions the
    // reference's grid sees (later points of a cell only push_back into an existing node).  Replaying it through the same
    // libstdc++ container gives the reference's iteration order (src/utility.cpp:180-187).
    std::tr1::unordered_map<CellKey, unsigned int, CellHash> grid;
    for (int j = 0; j < m; ++j) {
        const unsigned int i = first[(size_t)j];
        const double* pt = on_device ? &first_xyz[3 * (size_t)j] : &xyz_world[3 * (size_t)i];
        CellKey k;
        k.x = static_cast<short>(pt[0] / size);
        k.y = static_cast<short>(pt[1] / size);
        k.z = static_cast<short>(pt[2] / size);
        grid[k] = i;
    }
    size_t w = 0;
    for (std::tr1::unordered_map<CellKey, unsigned int, CellHash>::const_iterator it = grid.begin(); it != grid.end(); ++it) out[w++] = it->second;
    *n_out = w;
    return SRL_OK;
}


// ---- N4: colour map host side ---------------------------------------------------------------------------------------
struct srl_color_map {
    srl_ctx* ctx = nullptr;
    srl_map* vox = nullptr;                 // color_voxel_map (include/lioOptimization.h:275)
    double min_dist = 0.15;
    srl::ColorPoint* d_cpts = nullptr;      // max_voxels * 20
    double* d_last_visited = nullptr;       // max_voxels
    srl::Slot* d_fine = nullptr;            // hashmap_3d_points as an occupancy set: key = fine cell, block = index into rgb_points
    size_t fine_capacity = 0;
    unsigned int* d_rgb_points = nullptr;   // rgb_points_vec: point ids (block * 20 + index)
    size_t max_rgb_points = 0;
    unsigned long long* d_recent_temp = nullptr;   // voxels_recent_visited_temp (packed keys)
    unsigned long long* d_recent = nullptr;        // map_tracker->voxels_recent_visited
    size_t recent_capacity = 0;
    long long* d_counters = nullptr;        // [0] rgb points, [1] recent_temp size, [2] scratch, [3] scratch
    int64_t n_rgb_points = 0, n_recent_temp = 0, n_recent = 0, n_new_recent = 0;
};

int srl_color_map_create(srl_ctx* ctx, double voxel_size, int32_t max_num_points_in_voxel, size_t max_voxels, double min_distance_points,
                         srl_color_map** out) {
    if (!ctx || !out || !(min_distance_points > 0)) return SRL_BAD_ARG;
    *out = nullptr;
    srl_color_map* cm = new srl_color_map();
    cm->ctx = ctx; cm->min_dist = min_distance_points;
    int rc = srl_map_create(ctx, voxel_size, max_num_points_in_voxel, max_voxels, &cm->vox);
    if (rc != SRL_OK) { delete cm; return rc; }
    cm->max_rgb_points = max_voxels * (size_t)kBlockCap;
    cm->recent_capacity = 4 * max_voxels + 1024;
    size_t cap = 1024;
    while (cap < 2 * cm->max_rgb_points) cap <<= 1;
    cm->fine_capacity = cap;
    cudaError_t e;
    if ((e = cudaMalloc(&cm->d_cpts, max_voxels * kBlockCap * sizeof(ColorPoint))) != cudaSuccess ||
        (e = cudaMalloc(&cm->d_last_visited, max_voxels * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&cm->d_fine, cap * sizeof(Slot))) != cudaSuccess ||
        (e = cudaMalloc(&cm->d_rgb_points, cm->max_rgb_points * sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaMalloc(&cm->d_recent_temp, cm->recent_capacity * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&cm->d_recent, cm->recent_capacity * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&cm->d_counters, 8 * sizeof(long long))) != cudaSuccess ||
        (e = cudaMemsetAsync(cm->d_fine, 0, cap * sizeof(Slot), ctx->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(cm->d_cpts, 0, max_voxels * kBlockCap * sizeof(ColorPoint), ctx->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(cm->d_last_visited, 0, max_voxels * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(cm->d_counters, 0, 8 * sizeof(long long), ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        srl_color_map_destroy(cm);
        return cuda_fail(ctx, e, "srl_color_map_create");
    }
    *out = cm;
    return SRL_OK;
}

void srl_color_map_destroy(srl_color_map* cm) {
    if (!cm) return;
    if (cm->vox) srl_map_destroy(cm->vox);
    cudaFree(cm->d_cpts); cudaFree(cm->d_last_visited); cudaFree(cm->d_fine); cudaFree(cm->d_rgb_points);
    cudaFree(cm->d_recent_temp); cudaFree(cm->d_recent); cudaFree(cm->d_counters);
    delete cm;
}

srl_map* srl_color_map_voxels(srl_color_map* cm) { return cm ? cm->vox : nullptr; }

int srl_color_map_stats(srl_color_map* cm, int64_t* n_voxels, int64_t* n_points, int64_t* n_rgb_points, int64_t* n_recent, int64_t* n_new_recent) {
    if (!cm) return SRL_BAD_ARG;
    int rc = srl_map_stats(cm->vox, n_voxels, n_points);
    if (n_rgb_points) *n_rgb_points = cm->n_rgb_points;
    if (n_recent) *n_recent = cm->n_recent;
    if (n_new_recent) *n_new_recent = cm->n_new_recent;
    return rc;
}

int srl_color_map_add_points(srl_color_map* cm, const double* xyz_world, size_t n, int32_t add_point_step, double time_sweep_end,
                             double time_last_process, int32_t to_rendering, int64_t* n_stored) {
    if (!cm || (n && !xyz_world) || add_point_step < 1) return SRL_BAD_ARG;
    srl_ctx* ctx = cm->ctx;
    srl_map* m = cm->vox;
    cudaStream_t st = ctx->stream;
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_stored) *n_stored = 0;
    if (to_rendering) cm->n_recent_temp = 0;                                   // :523-527
    const int64_t recent_before = cm->n_recent_temp;                           // :529
    const size_t msel = (n + (size_t)add_point_step - 1) / (size_t)add_point_step;   // points with idx % step == 0
    if (msel > 0x7fffffffULL) return set_err(ctx, SRL_BAD_ARG, "srl_color_map_add_points: too many points");
    if (msel) {
        cudaPointerAttributes attr;
        const bool on_device = cudaPointerGetAttributes(&attr, xyz_world) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
        if (!on_device) cudaGetLastError();
        size_t tmp_sort = 0, tmp_sort32 = 0, tmp_sel = 0, tmp_scan = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr,
                                        (unsigned int*)nullptr, (int)msel, 0, 50, st);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort32, (unsigned int*)nullptr, (unsigned int*)nullptr, (unsigned long long*)nullptr,
                                        (unsigned long long*)nullptr, (int)msel, 0, 32, st);
        cub::DeviceSelect::Flagged(nullptr, tmp_sel, thrust::counting_iterator<unsigned int>(0), (unsigned char*)nullptr,
                                   (unsigned int*)nullptr, (int*)nullptr, (int)msel, st);
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)msel, st);
        const size_t tmp_bytes = std::max(std::max(tmp_sort, tmp_sort32), std::max(tmp_sel, tmp_scan));
        const size_t need = (on_device ? 0 : align_up(n * 24)) + 5 * align_up(msel * 8) + 12 * align_up(msel * 4) + align_up(msel * 12) +
                            align_up(msel) + 4 * 256 + align_up(tmp_bytes);
        int rc = ensure_scratch(ctx, need);
        if (rc != SRL_OK) return rc;
        char* p = static_cast<char*>(ctx->d_scratch);
        auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
        const double* d_xyz = xyz_world;
        if (!on_device) {
            double* buf = reinterpret_cast<double*>(take(n * 24));
            SRL_CUDA(ctx, cudaMemcpyAsync(buf, xyz_world, n * 24, cudaMemcpyHostToDevice, st));
            d_xyz = buf;
        }
        unsigned long long* vk_a = reinterpret_cast<unsigned long long*>(take(msel * 8));
        unsigned long long* vk_b = reinterpret_cast<unsigned long long*>(take(msel * 8));
        unsigned long long* fk_a = reinterpret_cast<unsigned long long*>(take(msel * 8));
        unsigned long long* fk_b = reinterpret_cast<unsigned long long*>(take(msel * 8));
        unsigned long long* seg_key = reinterpret_cast<unsigned long long*>(take(msel * 8));
        unsigned int* idx_a = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* idx_b = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* idx_c = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* seg_start = reinterpret_cast<unsigned int*>(take(msel * 4));
        int* seg_slot = reinterpret_cast<int*>(take(msel * 4));
        unsigned int* is_new = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* new_rank = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* accept_id = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* seg_first = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* seg_first_sorted = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* winner = reinterpret_cast<unsigned int*>(take(msel * 4));
        unsigned int* wrank = reinterpret_cast<unsigned int*>(take(msel * 4));
        float* fxyz = reinterpret_cast<float*>(take(msel * 12));
        unsigned char* flags = reinterpret_cast<unsigned char*>(take(msel));
        int* d_nseg = reinterpret_cast<int*>(take(256));
        void* d_tmp = take(tmp_bytes);
        const int T = 256;
        const unsigned gb = (unsigned)((msel + T - 1) / T);
        const unsigned vmask = (unsigned)(m->capacity - 1), fmask = (unsigned)(cm->fine_capacity - 1);
        k_color_keys<<<gb, T, 0, st>>>(d_xyz, (long long)msel, add_point_step, m->voxel_size, cm->min_dist, vk_a, fk_a, idx_a, fxyz);
        SRL_CUDA(ctx, cudaMemsetAsync(accept_id, 0xff, msel * 4, st));
        SRL_CUDA(ctx, cudaMemsetAsync(winner, 0, msel * 4, st));
        size_t tb = tmp_bytes;
        cub::DeviceRadixSort::SortPairs(d_tmp, tb, vk_a, vk_b, idx_a, idx_b, (int)msel, 0, 50, st);
        k_seg_flags<<<gb, T, 0, st>>>(vk_b, (long long)msel, flags);
        tb = tmp_bytes;
        cub::DeviceSelect::Flagged(d_tmp, tb, thrust::counting_iterator<unsigned int>(0), flags, seg_start, d_nseg, (int)msel, st);
        int n_seg = 0;
        SRL_CUDA(ctx, cudaMemcpyAsync(&n_seg, d_nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
        SRL_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 4;
        if (n_seg > 0) {
            const unsigned gs = (unsigned)((n_seg + T - 1) / T);
            k_seg_lookup<<<gs, T, 0, st>>>(m->d_slots, vmask, vk_b, seg_start, d_nseg, 1, seg_slot, is_new);   // min_num_points = 0 (:539)
            tb = tmp_bytes;
            cub::DeviceScan::ExclusiveSum(d_tmp, tb, is_new, new_rank, n_seg, st);
            unsigned int last_rank = 0, last_new = 0;
            SRL_CUDA(ctx, cudaMemcpyAsync(&last_rank, new_rank + (n_seg - 1), 4, cudaMemcpyDeviceToHost, st));
            SRL_CUDA(ctx, cudaMemcpyAsync(&last_new, is_new + (n_seg - 1), 4, cudaMemcpyDeviceToHost, st));
            SRL_CUDA(ctx, cudaStreamSynchronize(st));
            const long long total_new = (long long)last_rank + last_new;
            if ((size_t)(m->n_voxels + total_new) > m->max_voxels)
                return set_err(ctx, SRL_MAP_FULL, "srl_color_map_add_points: voxel pool exhausted (raise max_voxels)");
            long long before = 0, after = 0;
            SRL_CUDA(ctx, cudaMemcpyAsync(&before, m->d_counters, sizeof(long long), cudaMemcpyDeviceToHost, st));
            if (total_new > 0)
                k_seg_claim<<<gs, T, 0, st>>>(m->d_slots, vmask, m->d_blocks, vk_b, seg_start, d_nseg, is_new, new_rank, (long long)m->n_voxels, seg_slot);
            k_color_seg_process<<<gs, T, 0, st>>>(m->d_slots, m->d_blocks, cm->d_cpts, cm->d_last_visited, vk_b, idx_b, fxyz, seg_start, d_nseg,
                                                  seg_slot, is_new, (long long)msel, m->cap, time_sweep_end, time_last_process, accept_id,
                                                  seg_first, m->d_counters);
            m->n_voxels += total_new;
            // ---- recent list: the voxels this sweep visited for the first time, in the order of their first point
            k_color_seg_keys<<<gs, T, 0, st>>>(vk_b, seg_start, d_nseg, seg_key);
            tb = tmp_bytes;
            cub::DeviceRadixSort::SortPairs(d_tmp, tb, seg_first, seg_first_sorted, seg_key, fk_b /* reused as sorted keys */, n_seg, 0, 32, st);
            const long long host_cnt[4] = {cm->n_rgb_points, cm->n_recent_temp, 0, 0};
            SRL_CUDA(ctx, cudaMemcpyAsync(cm->d_counters, host_cnt, sizeof(host_cnt), cudaMemcpyHostToDevice, st));
            k_color_append_recent<<<gs, T, 0, st>>>(seg_first_sorted, fk_b, n_seg, cm->d_recent_temp, cm->d_counters, (long long)cm->recent_capacity);
            long long appended = 0;
            SRL_CUDA(ctx, cudaMemcpyAsync(&appended, cm->d_counters + 3, sizeof(long long), cudaMemcpyDeviceToHost, st));
            SRL_CUDA(ctx, cudaMemcpyAsync(&after, m->d_counters, sizeof(long long), cudaMemcpyDeviceToHost, st));
            SRL_CUDA(ctx, cudaStreamSynchronize(st));
            if ((size_t)(cm->n_recent_temp + appended) > cm->recent_capacity)
                return set_err(ctx, SRL_MAP_FULL, "srl_color_map_add_points: recent-voxel list exhausted (render or clear it)");
            cm->n_recent_temp += appended;
            if (n_stored) *n_stored = after - before;
            // ---- rgb_points_vec: the first accepted point of the sweep in every still-free fine cell, in sweep order
            k_color_mask_fine<<<gb, T, 0, st>>>(fk_a, accept_id, (long long)msel);
            tb = tmp_bytes;
            cub::DeviceRadixSort::SortPairs(d_tmp, tb, fk_a, fk_b, idx_a, idx_c, (int)msel, 0, 50, st);
            k_color_fine_winners<<<gb, T, 0, st>>>(cm->d_fine, fmask, fk_b, idx_c, (long long)msel, winner);
            tb = tmp_bytes;
            cub::DeviceScan::ExclusiveSum(d_tmp, tb, winner, wrank, (int)msel, st);
            unsigned int lr = 0, lw = 0;
            SRL_CUDA(ctx, cudaMemcpyAsync(&lr, wrank + (msel - 1), 4, cudaMemcpyDeviceToHost, st));
            SRL_CUDA(ctx, cudaMemcpyAsync(&lw, winner + (msel - 1), 4, cudaMemcpyDeviceToHost, st));
            SRL_CUDA(ctx, cudaStreamSynchronize(st));
            const long long n_win = (long long)lr + lw;
            if ((size_t)(cm->n_rgb_points + n_win) > cm->max_rgb_points)
                return set_err(ctx, SRL_MAP_FULL, "srl_color_map_add_points: rgb point list exhausted");
            if (n_win > 0)
                k_color_emit_rgb<<<gb, T, 0, st>>>(cm->d_fine, fmask, fk_a, winner, wrank, accept_id, (long long)msel, cm->d_rgb_points, cm->d_counters,
                                                   (long long)cm->max_rgb_points);
            SRL_CUDA(ctx, cudaGetLastError());
            cm->n_rgb_points += n_win;
            ctx->launches += 12;
        }
    }
    if (to_rendering) {                                                         // :544-550
        if (cm->n_recent_temp) SRL_CUDA(ctx, cudaMemcpyAsync(cm->d_recent, cm->d_recent_temp, (size_t)cm->n_recent_temp * 8, cudaMemcpyDeviceToDevice, st));
        cm->n_recent = cm->n_recent_temp;
        cm->n_new_recent = cm->n_recent - recent_before;
    }
    SRL_CUDA(ctx, cudaStreamSynchronize(st));
    return SRL_OK;
}

int srl_color_map_render_recent(srl_color_map* cm, const srl_camera* cam, const uint8_t* image_bgr, double obs_time, int64_t* n_rendered) {
    if (!cm || !cam || !image_bgr || cam->cols < 2 || cam->rows < 2) return SRL_BAD_ARG;
    srl_ctx* ctx = cm->ctx;
    srl_map* m = cm->vox;
    cudaStream_t st = ctx->stream;
    SRL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_rendered) *n_rendered = 0;
    const size_t nr = (size_t)cm->n_recent;
    if (nr == 0) return SRL_OK;
    cudaPointerAttributes attr;
    const bool on_device = cudaPointerGetAttributes(&attr, image_bgr) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    if (!on_device) cudaGetLastError();
    const size_t img_bytes = (size_t)cam->rows * cam->cols * 3;
    size_t tmp_sort = 0, tmp_rle = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)nr, 0, 50, st);
    cub::DeviceRunLengthEncode::Encode(nullptr, tmp_rle, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)nr, st);
    const size_t tmp_bytes = std::max(tmp_sort, tmp_rle);
    const size_t need = (on_device ? 0 : align_up(img_bytes)) + 2 * align_up(nr * 8) + align_up(nr * 4) + 2 * 256 + align_up(tmp_bytes);
    int rc = ensure_scratch(ctx, need);
    if (rc != SRL_OK) return rc;
    char* p = static_cast<char*>(ctx->d_scratch);
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
    const unsigned char* d_img = image_bgr;
    if (!on_device) {
        unsigned char* buf = reinterpret_cast<unsigned char*>(take(img_bytes));
        SRL_CUDA(ctx, cudaMemcpyAsync(buf, image_bgr, img_bytes, cudaMemcpyHostToDevice, st));
        d_img = buf;
    }
    unsigned long long* sorted = reinterpret_cast<unsigned long long*>(take(nr * 8));
    unsigned long long* uniq = reinterpret_cast<unsigned long long*>(take(nr * 8));
    int* mult = reinterpret_cast<int*>(take(nr * 4));
    int* d_nuniq = reinterpret_cast<int*>(take(256));
    unsigned long long* d_count = reinterpret_cast<unsigned long long*>(take(256));
    void* d_tmp = take(tmp_bytes);
    SRL_CUDA(ctx, cudaMemsetAsync(d_count, 0, 8, st));
    size_t tb = tmp_bytes;
    cub::DeviceRadixSort::SortKeys(d_tmp, tb, cm->d_recent, sorted, (int)nr, 0, 50, st);
    tb = tmp_bytes;
    cub::DeviceRunLengthEncode::Encode(d_tmp, tb, sorted, uniq, mult, d_nuniq, (int)nr, st);
    CamConst c;
    quat_to_rot(cam->q_camera_world, c.R);
    for (int i = 0; i < 3; ++i) { c.t_cw[i] = cam->t_camera_world[i]; c.t_wc[i] = cam->t_world_camera[i]; }
    c.fx = cam->fx; c.fy = cam->fy; c.cx = cam->cx; c.cy = cam->cy; c.fov = cam->fov_margin; c.cols = cam->cols; c.rows = cam->rows;
    const int T = 256;
    k_color_render<<<(unsigned)((nr * 32 + T - 1) / T), T, 0, st>>>(m->d_slots, (unsigned)(m->capacity - 1), m->d_blocks, cm->d_cpts, uniq, mult, d_nuniq, c, d_img,
                                                               obs_time, d_count);
    SRL_CUDA(ctx, cudaGetLastError());
    unsigned long long cnt = 0;
    SRL_CUDA(ctx, cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, st));
    SRL_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 3;
    if (n_rendered) *n_rendered = (int64_t)cnt;
    return SRL_OK;
}

int srl_color_map_download_state(srl_color_map* cm, size_t max_voxels, int16_t* rgb, int16_t* n_rgb, float* cov, double* obs_dist, double* last_obs,
                                 double* last_visited) {
    if (!cm) return SRL_BAD_ARG;
    srl_ctx* ctx = cm->ctx;
    srl_map* m = cm->vox;
    const size_t nv = (size_t)m->n_voxels;
    if (nv == 0) return SRL_OK;
    if (nv > max_voxels || !rgb || !n_rgb || !cov || !obs_dist || !last_obs || !last_visited) return set_err(ctx, SRL_BAD_ARG, "srl_color_map_download_state: output too small");
    const int cap = m->cap;
    const size_t e = nv * (size_t)cap;
    const size_t need = align_up(e * 6) + align_up(e * 2) + align_up(e * 12) + 2 * align_up(e * 8) + align_up(nv * 8);
    int rc = ensure_scratch(ctx, need);
    if (rc != SRL_OK) return rc;
    char* p = static_cast<char*>(ctx->d_scratch);
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
    short* d_rgb = reinterpret_cast<short*>(take(e * 6));
    short* d_n = reinterpret_cast<short*>(take(e * 2));
    float* d_cov = reinterpret_cast<float*>(take(e * 12));
    double* d_od = reinterpret_cast<double*>(take(e * 8));
    double* d_lo = reinterpret_cast<double*>(take(e * 8));
    double* d_lv = reinterpret_cast<double*>(take(nv * 8));
    const int T = 256;
    k_color_download<<<(unsigned)((e + T - 1) / T), T, 0, ctx->stream>>>(cm->d_cpts, m->d_blocks, cm->d_last_visited, (long long)nv, cap, d_rgb, d_n, d_cov, d_od, d_lo, d_lv);
    SRL_CUDA(ctx, cudaGetLastError());
    SRL_CUDA(ctx, cudaMemcpyAsync(rgb, d_rgb, e * 6, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(n_rgb, d_n, e * 2, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(cov, d_cov, e * 12, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(obs_dist, d_od, e * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(last_obs, d_lo, e * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaMemcpyAsync(last_visited, d_lv, nv * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SRL_OK;
}

int srl_color_map_download_lists(srl_color_map* cm, int16_t* rgb_points /* n_rgb_points * 4 */, int16_t* recent /* n_recent * 3 */) {
    if (!cm) return SRL_BAD_ARG;
    srl_ctx* ctx = cm->ctx;
    srl_map* m = cm->vox;
    if (cm->n_rgb_points && rgb_points) {
        const size_t n = (size_t)cm->n_rgb_points;
        int rc = ensure_scratch(ctx, n * 8);
        if (rc != SRL_OK) return rc;
        short* d_out = static_cast<short*>(ctx->d_scratch);
        const int T = 256;
        k_color_rgb_ids<<<(unsigned)((n + T - 1) / T), T, 0, ctx->stream>>>(cm->d_rgb_points, (long long)n, m->d_blocks, m->cap, d_out);
        SRL_CUDA(ctx, cudaGetLastError());
        SRL_CUDA(ctx, cudaMemcpyAsync(rgb_points, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (cm->n_recent && recent) {
        std::vector<unsigned long long> keys((size_t)cm->n_recent);
        SRL_CUDA(ctx, cudaMemcpyAsync(keys.data(), cm->d_recent, keys.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < keys.size(); ++i) { short x, y, z; unpack_key(keys[i], x, y, z); recent[3 * i] = x; recent[3 * i + 1] = y; recent[3 * i + 2] = z; }
    }
    SRL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
