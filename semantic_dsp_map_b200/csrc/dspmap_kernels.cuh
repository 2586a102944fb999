// dspmap_kernels.cuh — the per-frame kernels of the particle map engine (sm_100a).
//
// One frame = K1 project -> (P2 object moves) -> U2 vertex mask + sweep + pixel binning -> U3 ck / weight update
//             -> B1 birth + resample -> O1 occupancy + label fusion.
// Each kernel header names the reference routine it replaces (paths relative to the reference root).
// Iteration order is the oracle's "canonical" order (ascending particle index inside a pixel list, ascending
// reference rank for births), evaluated sequentially per output element so float sums are bit-identical.
#pragma once
#include "dspmap_device.cuh"

#include <cstddef>

namespace dspk {

// device-resident per-frame counters (zeroed at frame start)
struct Counters {
    uint32_t seq_head;         // = seq_tail = the frame's time stamp once the frame's results are final (host-side completion marker)
    uint32_t n_pairs;          // (pixel, particle) pairs produced by the sweep (may exceed vis_cap -> overflow)
    uint32_t n_valid_points;
    uint32_t n_swept_voxels;
    uint32_t n_swept_nonempty;
    uint32_t n_born;
    uint32_t n_birth_voxels;   // also the length of the birth entry list
    uint32_t n_resampled;
    uint32_t n_occupied;
    uint32_t n_free;
    uint32_t n_nonempty_voxels;
    uint32_t n_moved;
    uint32_t n_removed;
    uint32_t n_selected;       // particles selected by the object-move collection
    uint32_t n_move_voxels;    // also the length of the move entry list
    uint32_t n_movable;        // live particles of movable tracks listed by the occupancy pass
    uint32_t move_cursor0;     // random-table cursor the object moves of this frame start from
    uint32_t hash_overflow;    // insertion hash table full (cannot happen with cap >= 2 * candidates)
    uint32_t scratch;
    uint32_t n_sweep_work;     // swept voxels with live slots (work list of the slot pass)
    uint32_t n_occ_work;       // voxels whose occupancy needs the slot records
    uint32_t comm_errors;      // barriers that timed out (sharded maps)
    uint32_t ck_next;          // k_ck: next pixel tile to hand out (dynamic work distribution)
    uint32_t wu_next;          // k_weight_update: next 32-particle group to hand out
    uint32_t n_mine;           // sharded map: listed particles this rank owns (k_select_mine)
    uint32_t mov_out[8];       // moved particles sent to each destination rank this frame
    uint32_t n_visible;        // listed particles (total of the pixel lists)
    uint32_t n_birth_big;      // touched voxels whose candidate search was handed to whole blocks (k_birth_gather_big)
    uint32_t seq_tail;
};

// survives across frames
struct Persist {
    uint32_t cursor_ring;      // RingBufferOperations::gaussian_random_calculator_ cursor (operations.h:1460)
    uint32_t cursor_map;       // SemanticDSPMap::gaussian_random_ cursor (semantic_dsp_map.h:265)
    uint32_t n_movable;        // length of the movable-particle list left by the last occupancy pass
};

struct Store {
    float* x; float* y; float* z; float* w;
    uint64_t* meta;
    uint32_t* hdr;             // per voxel, see dspmap_device.cuh
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---- sharded maps: peer memory over NVLink ------------------------------------------------------------------------
// Each rank owns a contiguous range of storage-z slabs and an "exchange arena" (one cudaMalloc, exported with CUDA IPC
// and mapped by every peer).  Kernels write what other ranks need straight into the peers' arenas; ranks meet at
// device-side barriers (monotonic flags in peer memory, system-scope fences, a timeout instead of a hang).
constexpr int MAX_RANKS = 8;
constexpr int COMM_COUNTS = 16;              // u32 per (destination rank, source rank)
enum { CC_PAIRS = 0, CC_OCC = 1, CC_FREE = 2, CC_MOVREC = 3, CC_MOVE0 = 4 /* .. CC_MOVE0 + unused */ };

// one listed (visible) particle: everything the pixel lists and the two stencil passes need, 32 bytes
struct __align__(16) VisRec {
    uint32_t pix, idx;
    float x, y;
    float z, w;
    uint32_t tf;      // track | forget << 16
    uint32_t order;   // sort key inside the pixel's list (the particle index, or the reference's visiting order)
};

// one moved particle on its way to the rank that owns its destination voxel, 32 bytes
struct __align__(16) MovRec {
    float x, y, z, w;
    uint64_t meta;
    uint32_t seq, voxel;
};

struct CommView {
    int32_t rank, size;
    uint32_t rec_cap;                  // records per (rank) segment of the visible list
    uint32_t mov_cap;                  // records per (source rank) segment of the incoming moved-particle list
    uint32_t* flags[MAX_RANKS];        // peer p's barrier flags, indexed by source rank
    uint32_t* counts[MAX_RANKS];       // peer p's counts[source rank][COMM_COUNTS]
    VisRec* recs[MAX_RANKS];           // peer p's visible-record arena: [source rank][rec_cap]
    MovRec* movs[MAX_RANKS];           // peer p's incoming moved particles: [source rank][mov_cap]
    float4* obs_a[MAX_RANKS];          // peer p's observation planes (dspmap_stencil.cuh), written by pass 1 of the stencil
    uint2* obs_b[MAX_RANKS];
    uint32_t* move_counts[MAX_RANKS];  // peer p's per-move particle counts: [source rank][MAX_MOVES]
};

// all ranks arrive, then all leave.  One block of 32 threads; thread p handles peer p.
__global__ void k_comm_barrier(CommView c, uint32_t epoch, uint32_t* __restrict__ err) {
    const int p = threadIdx.x;
    __threadfence_system();
    if (p < c.size) {
        *(volatile uint32_t*)(c.flags[p] + c.rank) = epoch;
        const long long t0 = clock64();
        while (*(volatile uint32_t*)(c.flags[c.rank] + p) < epoch) {
            if (clock64() - t0 > 4000000000ll) { atomicAdd(err, 1u); break; }   // ~2 s: report instead of hanging the GPU
        }
    }
    __threadfence_system();
}

// Sharded map, end of a frame (after the last barrier): this rank's counter block and its table of per-rank counts go to the
// host - plain stores into the handle's pinned copies, the block's two stamps last - and the counters are zeroed for the next
// frame.  (An unsharded map does this from k_occ_scan; here the barriers after the scan still count their time-outs.)
__global__ void __launch_bounds__(64)
k_publish_counters(Counters* __restrict__ cnt, const uint32_t* __restrict__ counts, Counters* __restrict__ host_cnt, uint32_t* __restrict__ host_counts,
                   uint32_t n_counts, uint32_t frame_stamp) {
    constexpr int NW = sizeof(Counters) / 4;
    volatile uint32_t* src = reinterpret_cast<volatile uint32_t*>(cnt);
    volatile uint32_t* dst = reinterpret_cast<volatile uint32_t*>(host_cnt);
    const volatile uint32_t* csrc = counts;
    volatile uint32_t* cdst = host_counts;
    for (uint32_t i = threadIdx.x; i < n_counts; i += blockDim.x) cdst[i] = csrc[i];
    for (int i = 1 + (int)threadIdx.x; i < NW - 1; i += (int)blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { dst[0] = frame_stamp; dst[NW - 1] = frame_stamp; }
    for (int i = (int)threadIdx.x; i < NW; i += (int)blockDim.x) src[i] = 0u;
}

// =====================================================================================================
// K1a depth_int       the internal-resolution depth image (manualResize of the depth Mat, pointcloud_tools.h:1104-1133):
//                     everything up to the pixel lists needs only this, so it runs before the masks have arrived.
// =====================================================================================================
__global__ void __launch_bounds__(256)
k_depth_int(const __grid_constant__ FrameParams f, const float* __restrict__ depth, int depth_stride, float* __restrict__ depth_int) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= f.W * f.H) return;
    const int i = q / f.W, j = q - i * f.W;
    int si = i, sj = j;
    if (f.boost && !f.pre_decimated) {
        si = min((int)((float)i * 2.0f), f.in_H - 1);
        sj = min((int)((float)j * 2.0f), f.in_W - 1);
    }
    depth_int[q] = depth[(size_t)si * depth_stride + sj];
}

// =====================================================================================================
// K1  project_label   replaces PointCloudTools::manualResize + generateLabeledPointCloud
//                     (utils/pointcloud_tools.h:1104-1133, 88-310)
// one thread per internal pixel; double-precision back-projection, float narrowing at the end.
// =====================================================================================================
__global__ void __launch_bounds__(256)
k_project(const __grid_constant__ FrameParams f, const MaskDesc* __restrict__ masks,
          const uint16_t* __restrict__ lut_static_track, const uint8_t* __restrict__ lut_static_label,
          const float* __restrict__ depth_int, float4* __restrict__ cloud_pos, uint32_t* __restrict__ cloud_meta,
          Counters* __restrict__ cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    if (q < f.W * f.H) {
        const int i = q / f.W, j = q - i * f.W;
        int si = i, sj = j;
        if (f.boost && !f.pre_decimated) {  // :1104-1133, src = (int)(i * (1/0.5f)) clamped
            si = min((int)((float)i * 2.0f), f.in_H - 1);
            sj = min((int)((float)j * 2.0f), f.in_W - 1);
        }
        const float d = depth_int[q];

        // track id image, :116-155 (first static mask only), :162-215 (instance masks overwrite in list order)
        uint32_t inst = 65535u;
        for (int m = 0; m < f.n_masks; ++m) {
            if (masks[m].is_static) {
                inst = lut_static_track[masks[m].data[(size_t)si * masks[m].row_stride + sj]];
                break;
            }
        }
        int owner = -1;  // last instance mask whose track id equals the pixel's final id (label / bbox source)
        if (f.consider_instance) {
            for (int m = 0; m < f.n_masks; ++m) {
                if (masks[m].is_static) continue;
                if (masks[m].data[(size_t)si * masks[m].row_stride + sj] > 0) inst = (uint32_t)masks[m].track_id & 0xffffu;
            }
            for (int m = 0; m < f.n_masks; ++m)
                if (!masks[m].is_static && ((uint32_t)masks[m].track_id & 0xffffu) == inst) owner = m;
        }

        float4 pos = make_float4(0.f, 0.f, 0.f, f.sigma_invalid);
        uint32_t meta = 0;
        const bool depth_ok = !(d != d) && !((double)d < f.dmin) && !((double)d > f.dmax);           // :228
        if (depth_ok && !(f.zed2_filters && inst == f.sky_track)) {                                 // :234-241
            const double vj = (double)j, vi = (double)i, dd = (double)d;
            double pc[3], pg[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) pc[r] = ((f.Kinv[r * 3 + 0] * vj + f.Kinv[r * 3 + 1] * vi) + f.Kinv[r * 3 + 2] * 1.0) * dd;  // :243-245
#pragma unroll
            for (int r = 0; r < 3; ++r) pg[r] = ((f.R[r * 3 + 0] * pc[0] + f.R[r * 3 + 1] * pc[1]) + f.R[r * 3 + 2] * pc[2]) + f.t[r];  // :247-249
            const float sigma = f.depth_noise ? (f.n0 + f.n1 * d) : 0.1f;                            // :285-289
            uint32_t label, track = inst;
            bool background = false;
            if (f.zed2_filters && f.consider_instance && inst < f.max_movable) {                     // :254-273
                double b0[3] = {0.0, 0.0, 0.0}, b1[3] = {0.0, 0.0, 0.0};  // unordered_map::operator[] default
                if (owner >= 0) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) { b0[a] = masks[owner].bbox_min[a]; b1[a] = masks[owner].bbox_max[a]; }
                }
                background = pg[0] < b0[0] || pg[0] > b1[0] || pg[1] < b0[1] || pg[1] > b1[1] || pg[2] < b0[2] || pg[2] > b1[2];
            }
            if (background) { label = 0; track = 65535u; }       // sigma is never assigned on this branch of the reference (:256-272)
            else if (inst > f.max_movable) label = lut_static_label[inst];                           // :277-279
            else label = owner >= 0 ? ((uint32_t)masks[owner].label_id & 0xffu) : 0u;                 // :281
            pos = make_float4((float)pg[0], (float)pg[1], (float)pg[2], (background && f.sigma_gated_set) ? f.sigma_invalid : sigma);   // :298-300
            meta = track | (label << 16) | (1u << 24);
            valid = true;
        }
        cloud_pos[q] = pos;
        cloud_meta[q] = meta;
    }
    const int n = __syncthreads_count(valid);
    if (threadIdx.x == 0 && n) atomicAdd(&cnt->n_valid_points, (uint32_t)n);
}

// =====================================================================================================
// U2a  vertex_mask    the "is this lattice vertex inside the frustum" predicate of the BFS
//                     (mc_ring/operations.h:1336-1338, isPointInFrustum :1240-1258), one bit per vertex.
// one warp per 32 consecutive x vertices; blockIdx.y = row (z, y) in MAP coordinates.  Vertices outside the
// host-computed bounding box of the frustum (conservative, in vertex indices) are written as 0 without the test.
// =====================================================================================================
constexpr uint32_t BFS_UNSET = 0xffffffffu, BFS_TENTATIVE = 0x80000000u;

__global__ void __launch_bounds__(256)
k_vertex_mask(const __grid_constant__ FrameParams f, uint32_t* __restrict__ vmask, uint32_t words_per_row, int y0, int ny, int z0, uint32_t* __restrict__ vrank) {
    // blockIdx.y enumerates the (y, z) rows of the bounding box grown by one vertex (what k_sweep_select may read)
    const uint32_t xw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (xw >= words_per_row) return;
    const uint32_t z = (uint32_t)z0 + blockIdx.y / (uint32_t)ny, y = (uint32_t)y0 + blockIdx.y % (uint32_t)ny;
    const uint32_t row = z * (f.NY + 1) + y;
    const uint32_t x = xw * 32 + lane_id();
    bool in = false;
    if (x <= f.NX && (int)x >= f.vb_lo[0] && (int)x <= f.vb_hi[0] && (int)y >= f.vb_lo[1] && (int)y <= f.vb_hi[1] && (int)z >= f.vb_lo[2] && (int)z <= f.vb_hi[2])
        in = point_in_frustum(f, (float)(int)x * f.voxel + f.vox_off[0], (float)(int)y * f.voxel + f.vox_off[1], (float)(int)z * f.voxel + f.vox_off[2]);
    const uint32_t bits = __ballot_sync(0xffffffffu, in);
    if (lane_id() == 0) vmask[(size_t)row * words_per_row + xw] = bits;
    if (vrank && x <= f.NX) vrank[(size_t)row * (f.NX + 1) + x] = BFS_UNSET;      // reference order: no vertex of these rows is reached yet
}

__device__ __forceinline__ bool vbit(const uint32_t* __restrict__ vmask, uint32_t wpr, uint32_t NYp1, uint32_t x, uint32_t y, uint32_t z) {
    return (__ldg(vmask + (size_t)(z * NYp1 + y) * wpr + (x >> 5)) >> (x & 31u)) & 1u;
}

// =====================================================================================================
// U2a' bfs_order      the ORDER of the reference's breadth-first search (mc_ring/operations.h:1326-1456), for
//                     dspmap_params.reference_iteration_order: rank[vertex] = position of the in-frustum vertex in the order the
//                     search takes vertices off its queue (vertices outside the frustum are taken off and dropped: they order nothing).
// A vertex enters the queue when the first of its neighbours (in queue order; +x -x +y -y +z -z within one) pushes it, so level
// L+1 in queue order = the (parent position, direction) pairs that reach a not-yet-ranked in-frustum vertex FIRST, ascending.
// One cooperative kernel, level-synchronous: every (frontier vertex, direction) item bids for its target with atomicMin of
// its item number; the item that won owns the target; an ordered compaction of the owners is the next frontier.
// Ranks live in one word per lattice vertex: UNSET, TENTATIVE | winning bid, or the final rank.
// =====================================================================================================
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t val, uint32_t* s_warp, uint32_t& total);

struct BfsScratch {
    uint32_t* rank;           // (NX+1) x (NY+1) x (NZ+1), indexed like the vertex mask rows: (z * (NY+1) + y) * (NX+1) + x
    uint32_t* frontier[2];    // vertex ids, in queue order
    uint32_t* block_count;    // owners per block of the current level
    uint32_t frontier_cap;
    uint32_t* overflow;
};

__global__ void __launch_bounds__(1024, 1)
k_bfs_order(const __grid_constant__ FrameParams f, const uint32_t* __restrict__ vmask, uint32_t wpr, BfsScratch sc) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_level[2];
    const uint32_t NXp = f.NX + 1, NYp = f.NY + 1, NZp = f.NZ + 1;
    const uint32_t n_threads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    // grid-wide barrier: monotonic counter in block_count[gridDim.x] (the kernel is launched cooperatively: all blocks are resident)
    uint32_t* bar = sc.block_count + gridDim.x;
    uint32_t bar_target = 0;
    auto grid_sync = [&]() {
        __syncthreads();
        bar_target += gridDim.x;
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            while (*(volatile uint32_t*)bar < bar_target) {}
            __threadfence();
        }
        __syncthreads();
    };
    auto lin = [&](uint32_t x, uint32_t y, uint32_t z) { return (z * NYp + y) * NXp + x; };
    if (gtid == 0) {
        const uint32_t s0 = lin((uint32_t)f.start_v[0], (uint32_t)f.start_v[1], (uint32_t)f.start_v[2]);
        sc.frontier[0][0] = s0;
        __stcg(sc.rank + s0, 0u);
    }
    uint32_t nf = 1, next_rank = 1;
    int cur = 0;
    grid_sync();
    // target vertex of item t = (frontier position t / 6, direction t % 6), or BFS_UNSET
    auto target = [&](const uint32_t* __restrict__ F, uint32_t t) -> uint32_t {
        const uint32_t u = __ldcg(F + t / 6u), d = t % 6u;
        uint32_t x = u % NXp, y = (u / NXp) % NYp, z = u / (NXp * NYp);
        // dir[6] = {+x, -x, +y, -y, +z, -z}, operations.h:1437
        if (d == 0) { if (++x >= NXp) return BFS_UNSET; }
        else if (d == 1) { if (x-- == 0) return BFS_UNSET; }
        else if (d == 2) { if (++y >= NYp) return BFS_UNSET; }
        else if (d == 3) { if (y-- == 0) return BFS_UNSET; }
        else if (d == 4) { if (++z >= NZp) return BFS_UNSET; }
        else { if (z-- == 0) return BFS_UNSET; }
        if (!vbit(vmask, wpr, NYp, x, y, z)) return BFS_UNSET;
        return lin(x, y, z);
    };
    while (nf) {
        const uint32_t* F = sc.frontier[cur];
        uint32_t* Fn = sc.frontier[cur ^ 1];
        const uint32_t total = 6u * nf, J = (total + n_threads - 1) / n_threads;     // J consecutive items per thread: thread order = item order
        const uint32_t t0 = gtid * J;
        // bids
        for (uint32_t j = 0; j < J && t0 + j < total; ++j) {
            const uint32_t n = target(F, t0 + j);
            if (n != BFS_UNSET && __ldcg(sc.rank + n) >= BFS_TENTATIVE) atomicMin(sc.rank + n, BFS_TENTATIVE | (t0 + j));
        }
        grid_sync();
        // owners of this thread's items, block-wide exclusive scan of the counts
        uint32_t mine = 0;
        for (uint32_t j = 0; j < J && t0 + j < total; ++j) {
            const uint32_t n = target(F, t0 + j);
            mine += n != BFS_UNSET && __ldcg(sc.rank + n) == (BFS_TENTATIVE | (t0 + j));
        }
        uint32_t block_total;
        const uint32_t off = block_exclusive_scan(mine, s_warp, block_total);
        if (threadIdx.x == 0) __stcg(sc.block_count + blockIdx.x, block_total);
        grid_sync();
        if (threadIdx.x < 32) {          // one warp: owners in the blocks before this one, owners of the whole level
            uint32_t bsum = 0, tsum = 0;
            for (uint32_t b = threadIdx.x; b < gridDim.x; b += 32) {
                const uint32_t c = __ldcg(sc.block_count + b);
                bsum += b < blockIdx.x ? c : 0u;
                tsum += c;
            }
            bsum = __reduce_add_sync(0xffffffffu, bsum); tsum = __reduce_add_sync(0xffffffffu, tsum);
            if (threadIdx.x == 0) { s_level[0] = bsum; s_level[1] = tsum; }
        }
        __syncthreads();
        const uint32_t base = s_level[0], level_total = s_level[1];
        uint32_t pos = base + off;
        for (uint32_t j = 0; j < J && t0 + j < total; ++j) {
            const uint32_t n = target(F, t0 + j);
            if (n != BFS_UNSET && __ldcg(sc.rank + n) == (BFS_TENTATIVE | (t0 + j))) {
                if (pos < sc.frontier_cap) __stcg(Fn + pos, n); else atomicAdd(sc.overflow, 1u);
                __stcg(sc.rank + n, next_rank + pos);
                ++pos;
            }
        }
        next_rank += level_total;
        nf = min(level_total, sc.frontier_cap);
        cur ^= 1;
        grid_sync();
    }
}

// ---- warp work queue: P-lane groups take the voxels that need slot-level work ---------------------------------
// `need` marks the lanes whose voxel needs it.  Each step hands the next G = 32/P such voxels to the G groups of the
// warp: lane = g*P + s works on slot s of the voxel owned by lane `src` (0xffffffff: group idle this step).
template <int P>
struct GroupQueue {
    static constexpr int G = 32 / P;
    uint32_t pending;
    __device__ __forceinline__ explicit GroupQueue(bool need) : pending(__ballot_sync(0xffffffffu, need)) {}
    __device__ __forceinline__ bool next(uint32_t& src) {
        if (!pending) return false;
        src = __fns(pending, 0, (int)(lane_id() / P) + 1);
#pragma unroll
        for (int i = 0; i < G; ++i) pending &= pending - 1;
        return true;
    }
};
template <int P> __device__ __forceinline__ uint32_t group_mask() { return (P == 32 ? 0xffffffffu : ((1u << P) - 1u)) << ((lane_id() / P) * P); }

// =====================================================================================================
// U2b  sweep          replaces getIdxOfVisibleParitlces minus its queue (mc_ring/operations.h:1344-1435)
// Two launches.
//   select : one thread per storage voxel of this shard decides "swept" (one of the 8 corner vertices in the frustum)
//            and reads the 4-byte header.  Empty swept voxels are finished on the spot (corner visibility test,
//            :1418-1431); swept voxels with live slots are appended to a work list (order is irrelevant: the pixel
//            lists are sorted afterwards).
//   slots  : one P-lane group per listed voxel, lane s = slot s (coalesced 8 B tail + 3 x 4 B position per lane):
//            invalidate when stale, reset to the newborn weight under a too-far depth pixel, skip when occluded,
//            otherwise append a (pixel, particle) pair (one atomic per warp) and count it per pixel.
// =====================================================================================================
__device__ __forceinline__ bool corner_visible(const FrameParams& f, const float* __restrict__ depth_int, uint32_t cmx, uint32_t cmy, uint32_t cmz) {
    const float gx = (float)(int)cmx * f.voxel + f.pmin[0] + f.map_center[0];   // mapXYZIdxToGlobalPose, operations.h:986-991
    const float gy = (float)(int)cmy * f.voxel + f.pmin[1] + f.map_center[1];
    const float gz = (float)(int)cmz * f.voxel + f.pmin[2] + f.map_center[2];
    int row, col;
    float cz;
    if (project_to_image(f, gx, gy, gz, row, col, cz)) return cz <= __ldg(depth_int + row * f.W + col);
    return false;
}

__global__ void __launch_bounds__(256)
k_sweep_select(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ vmask, uint32_t words_per_row,
               const float* __restrict__ depth_int, uint32_t* __restrict__ work, uint32_t* __restrict__ work_hdr, Counters* __restrict__ cnt, int3 box_lo, int3 box_n,
               const uint32_t* __restrict__ vrank, uint32_t* __restrict__ work_key) {
    // the grid covers only the voxels of the frustum's bounding box, in MAP coordinates: blockIdx.y = (my, mz) row of the
    // box, x along the threads.  Voxels outside the box touch no in-frustum vertex.
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t mz = (uint32_t)(box_lo.z + (int)(blockIdx.y / (uint32_t)box_n.y)), my = (uint32_t)(box_lo.y + (int)(blockIdx.y % (uint32_t)box_n.y));
    const uint32_t mx = (uint32_t)(box_lo.x + ix);
    bool swept = false;
    uint32_t v = 0, visit_key = BFS_UNSET;
    if (ix < box_n.x) {
        uint32_t bx, by, bz;
        map_to_buffer(f, (int)mx, (int)my, (int)mz, bx, by, bz);
        if (bz >= f.z_lo && bz < f.z_hi) {        // this shard's slabs
            v = buffer_to_voxel(f, bx, by, bz);
            const uint32_t nyp = f.NY + 1;
            if (!vrank) {
                // the eight corner bits: vertices mx, mx + 1 of the four vertex rows (my + b, mz + c) = two adjacent bits of one or two words per row
                const uint32_t w = mx >> 5, sh = mx & 31u;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const uint32_t* row = vmask + (size_t)((mz + c) * nyp + (my + b)) * words_per_row;
                        const uint32_t w0 = __ldg(row + w), w1 = (sh == 31u && w + 1 < words_per_row) ? __ldg(row + w + 1) : 0u;
                        swept |= (__funnelshift_r(w0, w1, sh) & 3u) != 0u;
                    }
            } else
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const bool in = vbit(vmask, words_per_row, nyp, mx + a, my + b, mz + c);
                        if (!vrank) { swept |= in; continue; }
                        // reference order: the voxel is visited from the first of its corner vertices the search reaches, as that
                        // vertex's voxel number (dx, dy, dz) = (-a, -b, -c) of the loops dx, dy, dz = -1..0 (operations.h:1344-1346);
                        // a corner inside the frustum that the search never reaches visits nothing
                        const uint32_t r = in ? vrank[((size_t)(mz + c) * nyp + (my + b)) * (f.NX + 1) + (mx + a)] : BFS_UNSET;
                        if (r < BFS_TENTATIVE) { swept = true; visit_key = min(visit_key, r * 8u + (uint32_t)((1 - a) * 4 + (1 - b) * 2 + (1 - c))); }
                    }
        }
    }
    __shared__ uint32_t s_swept, s_heavy, s_base;
    if (threadIdx.x == 0) { s_swept = 0; s_heavy = 0; }
    const uint32_t hdr = swept ? st.hdr[v] : 0u;
    const bool heavy = swept && hdr_live(hdr) != 0;
    if (swept && !heavy) {
        if (corner_visible(f, depth_int, mx, my, mz)) st.hdr[v] = (f.T & 0xffffu) | (hdr & HDR_MOVABLE);
    }
    __syncthreads();
    // one pair of global atomics per block: position of this block's heavy voxels in the work list
    const uint32_t b_swept = __ballot_sync(0xffffffffu, swept), b_heavy = __ballot_sync(0xffffffffu, heavy);
    uint32_t woff = 0;
    if (lane_id() == 0) {
        if (b_swept) atomicAdd(&s_swept, (uint32_t)__popc(b_swept));
        if (b_heavy) woff = atomicAdd(&s_heavy, (uint32_t)__popc(b_heavy));
    }
    woff = __shfl_sync(0xffffffffu, woff, 0);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_swept) atomicAdd(&cnt->n_swept_voxels, s_swept);
        s_base = s_heavy ? atomicAdd(&cnt->n_sweep_work, s_heavy) : 0u;
    }
    __syncthreads();
    if (heavy) {
        const uint32_t pos = s_base + woff + __popc(b_heavy & ((1u << lane_id()) - 1u));
        work[pos] = v;
        work_hdr[pos] = hdr;              // nothing changes a listed voxel's header before the slot pass reads it: one gather less there
        if (work_key) work_key[pos] = visit_key;
    }
}

template <int PB>
__global__ void __launch_bounds__(256)
k_sweep_slots(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
              const uint32_t* __restrict__ stamp_z, const uint32_t* __restrict__ work, const uint32_t* __restrict__ work_hdr, const float* __restrict__ depth_int,
              uint32_t* __restrict__ pix_count, CommView cv, Counters* __restrict__ cnt, const uint32_t* __restrict__ work_key) {
    constexpr int P = 1 << PB, G = 32 / P;
    const uint32_t n_work = cnt->n_sweep_work;
    const uint32_t s = lane_id() & (P - 1), g = lane_id() / P, gbase = lane_id() & ~(uint32_t)(P - 1);
    const uint32_t gmask = group_mask<P>();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t n_nonempty = 0;
    // the pair list is reserved PAIR_CHUNK entries at a time per warp; unused tails are filled with a sentinel pixel
    constexpr uint32_t PAIR_CHUNK = 32;
    uint32_t chunk_base = 0, chunk_used = PAIR_CHUNK;
    auto put_sentinel = [&](uint32_t pos) {
        for (int pr = 0; pr < cv.size; ++pr) (cv.recs[pr] + (size_t)cv.rank * cv.rec_cap + pos)->pix = 0xffffffffu;
    };
    // the voxel and its header word come from the work list (coalesced); the next iteration's pair is requested before this
    // iteration's slot records, so the list is never on the chain of dependent loads (list -> slots -> depth pixel)
    uint32_t nx_v = 0, nx_hdr = 0;
    if (warp * G + g < n_work) { nx_v = work[warp * G + g]; nx_hdr = work_hdr[warp * G + g]; }
    for (uint32_t w0 = warp * G; w0 < n_work; w0 += n_warps * G) {     // warp-uniform trip count
        const uint32_t wi = w0 + g;
        const bool active = wi < n_work;
        const uint32_t g_v = nx_v, g_hdr = nx_hdr;
        {
            const uint32_t wn = wi + n_warps * G;
            nx_v = 0; nx_hdr = 0;
            if (wn < n_work) { nx_v = work[wn]; nx_hdr = work_hdr[wn]; }
        }
        const size_t base = (size_t)g_v << PB;
        // position of the particle inside its pixel's list: ascending particle index, or (reference order) the order the search
        // visits the voxels in, slot order inside a voxel
        const uint32_t order_key = (work_key && active) ? ((work_key[wi] << PB) | s) : (uint32_t)(base + s);
        const bool has = active && s != 0 && ((g_hdr >> (16 + s)) & 1u);
        uint64_t m = 0;
        float x = 0.f, y = 0.f, z = 0.f, wgt = 0.f;
        if (has) { m = st.meta[base + s]; x = st.x[base + s]; y = st.y[base + s]; z = st.z[base + s]; wgt = st.w[base + s]; }
        uint32_t bx, by, bz;
        voxel_to_buffer(f, g_v, bx, by, bz);
        const uint32_t stamp = slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz);
        const bool stale = has && meta_ts(m) < stamp;
        const bool counted = has && !stale;
        bool observed = false, emit = false;
        uint32_t pix = 0;
        if (stale) st.meta[base + s] = meta_with_status(m, ST_INVALID);  // :1374-1378 outdated by the ring shift
        int row, col;
        float cz;
        if (counted && project_to_image(f, x, y, z, row, col, cz)) {
            const float d = __ldg(depth_int + row * f.W + col);
            if (d > f.depth_max) {                 // :1389-1395 free space seen through this pixel
                st.w[base + s] = OCC_INIT_WEIGHT;
                observed = true;
            } else if (cz > d * ONE_SIGMA_COEFF) {  // :1397-1400 occluded
            } else {
                observed = true;
                emit = true;
                pix = (uint32_t)(row * f.W + col);
            }
        }
        const uint32_t b_stale = __ballot_sync(0xffffffffu, stale), b_cnt = __ballot_sync(0xffffffffu, counted);
        const uint32_t b_obs = __ballot_sync(0xffffffffu, observed), b_emit = __ballot_sync(0xffffffffu, emit);
        if (b_emit) {
            const uint32_t ne = __popc(b_emit);
            if (chunk_used + ne > PAIR_CHUNK) {                      // warp-uniform: retire the chunk, reserve a new one
                for (uint32_t k = chunk_used + lane_id(); k < PAIR_CHUNK; k += 32)
                    if (chunk_base + k < f.vis_cap) put_sentinel(chunk_base + k);
                uint32_t nb = 0;
                if (lane_id() == 0) nb = atomicAdd(&cnt->n_pairs, PAIR_CHUNK);
                chunk_base = __shfl_sync(0xffffffffu, nb, 0);
                chunk_used = 0;
            }
            if (emit) {
                const uint32_t pos = chunk_base + chunk_used + __popc(b_emit & ((1u << lane_id()) - 1u));
                if (pos < f.vis_cap) {
                    const uint4 lo = make_uint4(pix, (uint32_t)(base + s), __float_as_uint(x), __float_as_uint(y));
                    const uint4 hi = make_uint4(__float_as_uint(z), __float_as_uint(wgt), meta_track(m) | (meta_forget(m) << 16), order_key);
                    for (int pr = 0; pr < cv.size; ++pr) {          // every rank builds the full pixel lists
                        uint4* dst = reinterpret_cast<uint4*>(cv.recs[pr] + (size_t)cv.rank * cv.rec_cap + pos);
                        dst[0] = lo; dst[1] = hi;
                    }
                    atomicAdd(pix_count + pix, 1u);
                }
            }
            chunk_used += ne;
        }
        if (active && s == 0) {                                     // group leader: the voxel's header
            const bool g_observed = (b_obs & gmask) != 0;
            const uint32_t g_nvalid = __popc(b_cnt & gmask);
            const uint32_t new_live = hdr_live(g_hdr) & ~((b_stale & gmask) >> gbase);
            bool stamp0 = g_observed;                                // :1415-1417
            if (!g_observed && g_nvalid == 0) {                      // every live slot was stale: the empty-voxel rule applies
                uint32_t gmx, gmy, gmz;
                buffer_to_map(f, bx, by, bz, gmx, gmy, gmz);
                stamp0 = corner_visible(f, depth_int, gmx, gmy, gmz);
            }
            const uint32_t nh = (stamp0 ? (f.T & 0xffffu) : hdr_ts(g_hdr)) | (g_hdr & HDR_MOVABLE) | (new_live << 16);
            if (nh != g_hdr) st.hdr[g_v] = nh;
            n_nonempty += g_nvalid > 0;
        }
    }
    for (uint32_t k = chunk_used + lane_id(); k < PAIR_CHUNK; k += 32)          // retire the last chunk
        if (chunk_base + k < f.vis_cap) put_sentinel(chunk_base + k);
    if (cv.size > 1) __threadfence_system();      // sharded map: peers read what this kernel wrote after the next barrier
    n_nonempty = __reduce_add_sync(0xffffffffu, n_nonempty);
    if (lane_id() == 0 && n_nonempty) atomicAdd(&cnt->n_swept_nonempty, n_nonempty);
}

// =====================================================================================================
// exclusive scan of the per-pixel counts -> CSR offsets (replaces particle_to_pixel_num_array /
// particle_to_pixel_index_map, mc_ring/buffer.h:90-93).  Two launches: block sums, then block-local scans.
// =====================================================================================================
constexpr int SCAN_BLOCK = 1024;

// exclusive scan over the block (blockDim.x a multiple of 32, <= 1024); total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t val, uint32_t* s_warp /*[33]*/, uint32_t& total) {
    uint32_t inc = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane_id() >= (uint32_t)o) inc += n;
    }
    const uint32_t wid = threadIdx.x >> 5;
    if (lane_id() == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = (lane_id() < (blockDim.x >> 5)) ? s_warp[lane_id()] : 0u;
        uint32_t winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane_id() >= (uint32_t)o) winc += n;
        }
        s_warp[lane_id()] = winc - ws;       // exclusive warp offsets
        if (lane_id() == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    const uint32_t r = s_warp[wid] + inc - val;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_blocksum(const uint32_t* __restrict__ count, uint32_t n, uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t s_warp[33];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    uint32_t total;
    block_exclusive_scan(i < n ? count[i] : 0u, s_warp, total);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_apply(const uint32_t* __restrict__ count, uint32_t n, const uint32_t* __restrict__ block_sum, uint32_t* __restrict__ offset, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_base;
    // prefix of the preceding blocks' sums
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += SCAN_BLOCK) part += block_sum[b];
    uint32_t tot;
    block_exclusive_scan(part, s_warp, tot);
    if (threadIdx.x == 0) s_base = tot;
    __syncthreads();
    const uint32_t base = s_base;
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t c = i < n ? count[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(c, s_warp, total);
    if (i < n) offset[i] = base + ex;
    if (i == n - 1) { offset[n] = base + ex + c; if (total_out) *total_out = base + ex + c; }
}

// this rank's record count goes into every rank's counts table
__global__ void k_publish_pairs(CommView cv, const Counters* __restrict__ cnt) {
    const int p = threadIdx.x;
    if (p < cv.size) cv.counts[p][cv.rank * COMM_COUNTS + CC_PAIRS] = cnt->n_pairs;
    __threadfence_system();
}

// per-pixel counts of the records that arrived from the OTHER ranks (own records were counted by the sweep)
__global__ void __launch_bounds__(256)
k_count_remote(CommView cv, const VisRec* __restrict__ recs, const uint32_t* __restrict__ counts, uint32_t* __restrict__ pix_count) {
    for (int g = 0; g < cv.size; ++g) {
        if (g == cv.rank) continue;
        const uint32_t n = min(counts[g * COMM_COUNTS + CC_PAIRS], cv.rec_cap);
        const VisRec* seg = recs + (size_t)g * cv.rec_cap;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
            const uint32_t pix = seg[k].pix;
            if (pix != 0xffffffffu) atomicAdd(pix_count + pix, 1u);
        }
    }
}

// scatter (particle index, record reference) keys into their pixel segments (arrival order; sorted next)
__global__ void __launch_bounds__(256)
k_fill(CommView cv, const VisRec* __restrict__ recs, const uint32_t* __restrict__ counts, const Counters* __restrict__ cnt,
       const uint32_t* __restrict__ offset, uint32_t* __restrict__ pix_count, uint64_t* __restrict__ csr_key) {
    for (int g = 0; g < cv.size; ++g) {
        const uint32_t n = min(g == cv.rank ? cnt->n_pairs : counts[g * COMM_COUNTS + CC_PAIRS], cv.rec_cap);
        const VisRec* seg = recs + (size_t)g * cv.rec_cap;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
            const uint32_t pix = seg[k].pix;
            if (pix == 0xffffffffu) continue;     // unused tail of a reserved chunk
            const uint32_t pos = offset[pix] + (atomicSub(pix_count + pix, 1u) - 1u);   // leaves the counts at zero
            csr_key[pos] = ((uint64_t)seg[k].order << 32) | (uint64_t)((uint32_t)g * cv.rec_cap + k);
        }
    }
}

// canonical order inside each pixel list: ascending particle index (insertion sort; lists are short)
__global__ void __launch_bounds__(256)
k_sort_lists(const __grid_constant__ FrameParams f, const uint32_t* __restrict__ offset, uint64_t* __restrict__ csr_key) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= f.W * f.H) return;
    const uint32_t b = offset[q], e = offset[q + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
        const uint64_t key = csr_key[i];
        uint32_t j = i;
        while (j > b && csr_key[j - 1] > key) { csr_key[j] = csr_key[j - 1]; --j; }
        csr_key[j] = key;
    }
}

// lay the listed particles out in list order for the two stencil passes (from the records, not from the store)
__global__ void __launch_bounds__(256)
k_gather_visible(const uint32_t* __restrict__ n_visible, const VisRec* __restrict__ recs, const uint64_t* __restrict__ csr_key,
                 uint32_t* __restrict__ csr_idx, float4* __restrict__ vis_pos, uint2* __restrict__ vis_mp) {
    const uint32_t n = *n_visible;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint64_t key = csr_key[k];
        const uint4* r = reinterpret_cast<const uint4*>(recs + (uint32_t)key);
        const uint4 lo = r[0], hi = r[1];
        csr_idx[k] = lo.y;                            // the particle index (the key's upper half is the list order)
        vis_pos[k] = make_float4(__uint_as_float(lo.z), __uint_as_float(lo.w), __uint_as_float(hi.x), __uint_as_float(hi.y));
        vis_mp[k] = make_uint2(hi.z, lo.x);          // {track | forget << 16, pixel}
    }
}

// =====================================================================================================
// U3  the two stencil passes of updateParticles (k_ck, k_weight_update): dspmap_stencil.cuh
// =====================================================================================================
}  // namespace dspk
#include "dspmap_stencil.cuh"
namespace dspk {

// =====================================================================================================
// ordered insertion of candidates into voxels.
// The reference inserts points one at a time into "the first vacant slot" of their voxel
// (mc_ring/operations.h:782-803), so the outcome for a voxel depends only on the ORDER of the candidates that
// fall into it.  A voxel can absorb at most topk = 2P-1 candidates before every later one is a no-op (P-1
// slots, one resample that frees at most P-1 more), so: hash the voxel, keep the topk smallest ranks per
// voxel with a cascade of atomicMin (order-independent result), then replay each touched voxel serially.
// Touched hash entries are appended to a list so the replay never scans the table.
// =====================================================================================================
__device__ __forceinline__ uint32_t hash_slot(uint32_t v, uint32_t shift) { return (v * 2654435761u) >> shift; }

// cap/shift: the power-of-two prefix of the table in use (births: 2x the pixel count; moves: the whole table)
__device__ __forceinline__ void topk_insert(const FrameParams& f, uint32_t cap, uint32_t shift, uint32_t v, uint32_t rank,
                                            uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks, uint32_t* __restrict__ entry_list,
                                            uint32_t* n_entries, uint32_t* overflow) {
    uint32_t h = hash_slot(v, shift);
    for (uint32_t probe = 0;; ++probe) {
        if (probe >= cap) { atomicAdd(overflow, 1u); return; }
        const uint32_t old = atomicCAS(keys + h, EMPTY_KEY, v);
        if (old == EMPTY_KEY) { entry_list[atomicAdd(n_entries, 1u)] = h; break; }
        if (old == v) break;
        h = (h + 1) & (cap - 1);
    }
    uint32_t* r = ranks + (size_t)h * f.topk;
    // early out: the last entry only ever decreases
    if (*(volatile uint32_t*)(r + f.topk - 1) < rank) return;
    uint32_t carry = rank;
    for (uint32_t s = 0; s < f.topk && carry != NO_RANK; ++s) {
        const uint32_t old = atomicMin(r + s, carry);
        carry = max(old, carry);
    }
}

// birth rank = position of the pixel in the reference's 3x3-interleaved raster (semantic_dsp_map.h:776-783)
__device__ __forceinline__ uint32_t birth_rank(const FrameParams& f, int i, int j) {
    const int hc = (f.H + 2) / 3, wc = (f.W + 2) / 3;
    return (uint32_t)((((i % 3) * 3 + (j % 3)) * hc + i / 3) * wc + j / 3);
}
__device__ __forceinline__ int birth_rank_to_pixel(const FrameParams& f, uint32_t rank) {
    const int hc = (f.H + 2) / 3, wc = (f.W + 2) / 3;
    const int b = (int)(rank % (uint32_t)wc);
    const int a = (int)((rank / (uint32_t)wc) % (uint32_t)hc);
    const int ph = (int)(rank / (uint32_t)(wc * hc));
    return ((ph / 3) + 3 * a) * f.W + (ph % 3) + 3 * b;
}

// birth rank -> pixel; -1 for the padding ranks of image sizes that are not multiples of 3
__device__ __forceinline__ int birth_rank_pixel_checked(const FrameParams& f, uint32_t rank) {
    const int hc = (f.H + 2) / 3, wc = (f.W + 2) / 3;
    if (rank >= (uint32_t)(9 * hc * wc)) return -1;
    const int b = (int)(rank % (uint32_t)wc), a = (int)((rank / (uint32_t)wc) % (uint32_t)hc), ph = (int)(rank / (uint32_t)(wc * hc));
    const int i = ph / 3 + 3 * a, j = ph % 3 + 3 * b;
    if (i >= f.H || j >= f.W) return -1;
    return i * f.W + j;
}

// Several newborn particles per point (nb_ptc_num_per_point > 1 with the depth-noise model): the reference perturbs each
// by sigma * three values of its pre-drawn Gaussian table, consumed in birth order (semantic_dsp_map.h:1177-1190;
// utils/basic_algorithms.h:426-440).  Valid pixel number k (in birth order) owns the draws cursor + 3*(k*nb + n) + 1..3.
__device__ __forceinline__ float3 newborn_position(const FrameParams& f, const float4 p, uint32_t k_valid, uint32_t n,
                                                   const float* __restrict__ randoms, uint32_t cursor) {
    if (f.nb_ptc == 1) return make_float3(p.x + 0.f, p.y + 0.f, p.z + 0.f);       // :1183-1184 zero noise
    const uint64_t d0 = (uint64_t)cursor + 3ull * ((uint64_t)k_valid * (uint32_t)f.nb_ptc + n);
    const float n0 = p.w * randoms[(d0 + 1) % f.rand_n], n1 = p.w * randoms[(d0 + 2) % f.rand_n], n2 = p.w * randoms[(d0 + 3) % f.rand_n];
    return make_float3(p.x + n0, p.y + n1, p.z + n2);
}

// valid flag of every birth rank (input of the scan that numbers the valid pixels in birth order; nb > 1 only)
__global__ void __launch_bounds__(256)
k_birth_flags(const __grid_constant__ FrameParams f, const uint32_t* __restrict__ cloud_meta, uint32_t* __restrict__ flags, uint32_t n_ranks) {
    const uint32_t rank = blockIdx.x * blockDim.x + threadIdx.x;
    if (rank >= n_ranks) return;
    const int q = birth_rank_pixel_checked(f, rank);
    flags[rank] = (q >= 0 && pt_valid(cloud_meta[q])) ? 1u : 0u;
}

// One thread per birth candidate = (birth RANK, copy n) — not per pixel: the lowest ranks are binned by the first blocks,
// so by the time the later 3x3 phases arrive most per-voxel lists are full and their cascades end at the first comparison.
__global__ void __launch_bounds__(256)
k_birth_bin(const __grid_constant__ FrameParams f, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta, uint32_t* __restrict__ keys,
            uint32_t* __restrict__ ranks, uint32_t* __restrict__ entry_list, Counters* __restrict__ cnt,
            const uint32_t* __restrict__ valid_prefix, const float* __restrict__ randoms, const Persist* __restrict__ persist) {
    const uint32_t cand = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nb = (uint32_t)f.nb_ptc, rank = cand / nb, n = cand - rank * nb;
    const int q = birth_rank_pixel_checked(f, rank);
    if (q < 0) return;
    if (!pt_valid(cloud_meta[q])) return;
    const float3 p = newborn_position(f, cloud_pos[q], nb > 1 ? valid_prefix[rank] : 0u, n, randoms, persist->cursor_map);
    uint32_t bx, by, bz;
    const uint32_t v = global_pos_to_voxel(f, p.x, p.y, p.z, bx, by, bz);
    if (v == INVALID_INDEX) return;
    if (bz < f.z_lo || bz >= f.z_hi) return;   // another shard owns this voxel
    topk_insert(f, f.birth_hash_cap, f.birth_hash_shift, v, cand, keys, ranks, entry_list, &cnt->n_birth_voxels, &cnt->hash_overflow);
}

// ---- one newborn particle per point (the default): the candidates of a voxel are FOUND instead of being collected ----------
// mark   : one thread per pixel; the pixels of a warp that fall into the same voxel act as one: a single hash insertion and four
//          atomic min / max that grow the voxel's bounding rectangle of candidate pixels (a voxel's candidates are a compact patch
//          of the image).
//          The pixels also count themselves per voxel, and the first topk to arrive leave their birth ranks in the voxel's rank
//          list (arrival order, i.e. no order).
// gather : one 8-lane group per touched voxel.  A voxel with at most topk candidates - nine in ten: everything but near
//          surfaces - has them all in its rank list already and only sorts them.  The others walk the rectangle in birth-rank
//          order - phase by phase of the 3x3 interleave, row by row - and keep the first topk pixels whose point falls into the
//          voxel: a wall one metre away is hundreds of candidates per voxel, and they all stop after the first rows of phase 0.
//          (Walking the rectangle of a voxel with three candidates far apart in the image was the tail of this kernel.)
struct BirthBox { uint32_t i_min, i_max, j_min, j_max, count; };      // idle state: {0xffffffff, 0, 0xffffffff, 0, 0}

__global__ void __launch_bounds__(256)
k_birth_mark(const __grid_constant__ FrameParams f, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta, uint32_t* __restrict__ keys,
             BirthBox* __restrict__ boxes, uint32_t* __restrict__ ranks, uint32_t* __restrict__ entry_list, Counters* __restrict__ cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = INVALID_INDEX, pi = 0, pj = 0;
    if (q < f.W * f.H && pt_valid(cloud_meta[q])) {
        const float4 p = cloud_pos[q];
        uint32_t bx, by, bz;
        v = global_pos_to_voxel(f, p.x + 0.f, p.y + 0.f, p.z + 0.f, bx, by, bz);
        if (v != INVALID_INDEX && (bz < f.z_lo || bz >= f.z_hi)) v = INVALID_INDEX;   // another shard owns this voxel
        pi = (uint32_t)(q / f.W); pj = (uint32_t)(q - (int)pi * f.W);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, v);
    const uint32_t i_lo = __reduce_min_sync(peers, pi), i_hi = __reduce_max_sync(peers, pi), j_lo = __reduce_min_sync(peers, pj), j_hi = __reduce_max_sync(peers, pj);
    const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
    uint32_t h = EMPTY_KEY, base = 0;
    if (v != INVALID_INDEX && leader == lane_id()) {
        h = hash_slot(v, f.birth_hash_shift);
        for (uint32_t probe = 0;; ++probe) {
            if (probe >= f.birth_hash_cap) { atomicAdd(&cnt->hash_overflow, 1u); h = EMPTY_KEY; break; }
            uint32_t old = *(volatile uint32_t*)(keys + h);
            if (old == EMPTY_KEY) old = atomicCAS(keys + h, EMPTY_KEY, v);
            if (old == EMPTY_KEY) { entry_list[atomicAdd(&cnt->n_birth_voxels, 1u)] = h; break; }
            if (old == v) break;
            h = (h + 1) & (f.birth_hash_cap - 1);
        }
        if (h != EMPTY_KEY) {
            BirthBox* b = boxes + h;
            atomicMin(&b->i_min, i_lo); atomicMax(&b->i_max, i_hi); atomicMin(&b->j_min, j_lo); atomicMax(&b->j_max, j_hi);
            base = atomicAdd(&b->count, (uint32_t)__popc(peers));
        }
    }
    __syncwarp();
    h = __shfl_sync(0xffffffffu, h, leader);
    base = __shfl_sync(0xffffffffu, base, leader);
    if (v == INVALID_INDEX || h == EMPTY_KEY) return;
    const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane_id()) - 1u));
    if (pos < f.topk) ranks[(size_t)h * f.topk + pos] = birth_rank(f, (int)pi, (int)pj);
}

__global__ void k_fill_birth_boxes(BirthBox* __restrict__ boxes, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) boxes[i] = BirthBox{0xffffffffu, 0u, 0xffffffffu, 0u, 0u};
}

struct BirthBig { uint32_t e, i_min, i_max, j_min, j_max; };      // a voxel handed to k_birth_gather_big: hash slot + rectangle
constexpr int BG_LANES = 8;      // lanes per touched voxel
constexpr int BG_VALS = 4;       // rank-list entries per lane: topk = 2 * MAX_P - 1 = 31 at most (static_assert below MAX_P)
__global__ void __launch_bounds__(256)
k_birth_gather(const __grid_constant__ FrameParams f, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta, const uint32_t* __restrict__ keys,
               BirthBox* __restrict__ boxes, uint32_t* __restrict__ ranks, const uint32_t* __restrict__ entry_list, Counters* __restrict__ cnt,
               BirthBig* __restrict__ big_list) {
    const uint32_t n_entries = cnt->n_birth_voxels, K = f.topk;
    const uint32_t sub = lane_id() & (BG_LANES - 1), gmask = ((1u << BG_LANES) - 1u) << (lane_id() & ~(uint32_t)(BG_LANES - 1));
    const uint32_t n_groups = (gridDim.x * blockDim.x) / BG_LANES;
    const int hc = (f.H + 2) / 3, wc = (f.W + 2) / 3;
    for (uint32_t li0 = (blockIdx.x * blockDim.x + threadIdx.x) / BG_LANES;; li0 += n_groups) {
        // warp-uniform exit (the groups of a warp stay together for the ballots below)
        if (__all_sync(0xffffffffu, li0 >= n_entries)) break;
        const bool live = li0 < n_entries;
        const uint32_t e = live ? entry_list[li0] : 0u, v = live ? keys[e] : INVALID_INDEX;
        int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
        uint32_t n_cand = 0;
        if (live) {
            const BirthBox bb = boxes[e];
            i0 = (int)bb.i_min; i1 = (int)bb.i_max; j0 = (int)bb.j_min; j1 = (int)bb.j_max; n_cand = bb.count;
        }
        __syncwarp();
        if (live && sub == 0) boxes[e] = BirthBox{0xffffffffu, 0u, 0xffffffffu, 0u, 0u};     // idle again for the next frame
        uint32_t* out = ranks + (size_t)e * K;
        // at most topk candidates: they are all in the list (k_birth_mark), in arrival order - every value finds its place by
        // counting the smaller ones (ranks are distinct); the lists of the warp's four groups are sorted side by side
        const bool listed = live && n_cand <= K;
        {
            uint32_t mine[BG_VALS];
#pragma unroll
            for (int t = 0; t < BG_VALS; ++t) { const uint32_t k = sub + (uint32_t)t * BG_LANES; mine[t] = (listed && k < n_cand) ? out[k] : NO_RANK; }
            uint32_t place[BG_VALS] = {};
            const uint32_t lane0 = lane_id() & ~(uint32_t)(BG_LANES - 1);
#pragma unroll
            for (int t2 = 0; t2 < BG_VALS; ++t2)
#pragma unroll
                for (int src = 0; src < BG_LANES; ++src) {
                    const uint32_t other = __shfl_sync(0xffffffffu, mine[t2], lane0 + src);
#pragma unroll
                    for (int t = 0; t < BG_VALS; ++t) place[t] += other < mine[t];
                }
            __syncwarp();
            if (listed) {
#pragma unroll
                for (int t = 0; t < BG_VALS; ++t) {
                    const uint32_t k = sub + (uint32_t)t * BG_LANES;
                    if (k < n_cand) out[place[t]] = mine[t];
                }
                for (uint32_t t = n_cand + sub; t < K; t += BG_LANES) out[t] = NO_RANK;
            }
        }
        // More candidates than the list holds: walk the rectangle.  The walk ends after about  area * topk / candidates  cells (the
        // hits are spread over the nine phases like the cells) - a few for a near surface, hundreds for a voxel with thirty
        // candidates scattered over a large rectangle; those go to whole blocks (k_birth_gather_big) instead of eight lanes.
        const uint32_t area = live ? (uint32_t)((i1 - i0 + 1) * (j1 - j0 + 1)) : 0u;
        const bool big = live && !listed && (uint64_t)area * K > (uint64_t)n_cand * (8u * BG_LANES);
        if (big && sub == 0) big_list[atomicAdd(&cnt->n_birth_big, 1u)] = BirthBig{e, (uint32_t)i0, (uint32_t)i1, (uint32_t)j0, (uint32_t)j1};
        const bool scan = live && !listed && !big;
        uint32_t found = scan ? 0u : K;
        for (int ph = 0; ph < 9; ++ph) {
            const int pi = ph / 3, pj = ph - 3 * pi;
            // cells (a, b) of this phase inside the rectangle: pixel (3a + pi, 3b + pj)
            const int a0 = (i0 - pi + 2) / 3, a1 = (i1 - pi >= 0) ? (i1 - pi) / 3 : -1, b0 = (j0 - pj + 2) / 3, b1 = (j1 - pj >= 0) ? (j1 - pj) / 3 : -1;
            const int nb = b1 - b0 + 1, ncell = (live && a1 >= a0 && nb > 0 && found < K) ? (a1 - a0 + 1) * nb : 0;
            const int ncell_max = (int)__reduce_max_sync(0xffffffffu, (uint32_t)ncell);
            for (int c0 = 0; c0 < ncell_max; c0 += BG_LANES) {
                const int c = c0 + (int)sub;
                bool hit = false;
                uint32_t rank = 0;
                if (c < ncell && found < K) {
                    const int a = a0 + c / nb, b = b0 + c % nb, q = (3 * a + pi) * f.W + 3 * b + pj;
                    if (pt_valid(cloud_meta[q])) {
                        const float4 p = cloud_pos[q];
                        uint32_t bx, by, bz;
                        hit = global_pos_to_voxel(f, p.x + 0.f, p.y + 0.f, p.z + 0.f, bx, by, bz) == v;
                        rank = (uint32_t)((ph * hc + a) * wc + b);
                    }
                }
                const uint32_t hits = __ballot_sync(0xffffffffu, hit) & gmask;
                if (hit) {
                    const uint32_t pos = found + (uint32_t)__popc(hits & ((1u << lane_id()) - 1u));
                    if (pos < K) out[pos] = rank;
                }
                found += (uint32_t)__popc(hits);
                if (__all_sync(0xffffffffu, c0 + BG_LANES >= ncell || found >= K)) break;     // every group of the warp is done with this phase
            }
        }
        if (scan)
            for (uint32_t t = min(found, K) + sub; t < K; t += BG_LANES) out[t] = NO_RANK;
    }
}

// the rectangle walk of k_birth_gather for one voxel per BLOCK: 256 cells of a phase at a time, ordered compaction of the hits
__global__ void __launch_bounds__(256)
k_birth_gather_big(const __grid_constant__ FrameParams f, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta, const uint32_t* __restrict__ keys,
                   uint32_t* __restrict__ ranks, const BirthBig* __restrict__ big_list, const Counters* __restrict__ cnt) {
    __shared__ uint32_t s_wcnt[8];
    const uint32_t n_big = cnt->n_birth_big, K = f.topk, wid = threadIdx.x >> 5;
    const int hc = (f.H + 2) / 3, wc = (f.W + 2) / 3;
    for (uint32_t bi = blockIdx.x; bi < n_big; bi += gridDim.x) {
        const BirthBig bb = big_list[bi];
        const uint32_t v = keys[bb.e];
        const int i0 = (int)bb.i_min, i1 = (int)bb.i_max, j0 = (int)bb.j_min, j1 = (int)bb.j_max;
        uint32_t* out = ranks + (size_t)bb.e * K;
        uint32_t found = 0;                                      // the same in every thread of the block
        for (int ph = 0; ph < 9 && found < K; ++ph) {
            const int pi = ph / 3, pj = ph - 3 * pi;
            const int a0 = (i0 - pi + 2) / 3, a1 = (i1 - pi >= 0) ? (i1 - pi) / 3 : -1, b0 = (j0 - pj + 2) / 3, b1 = (j1 - pj >= 0) ? (j1 - pj) / 3 : -1;
            const int nb = b1 - b0 + 1, ncell = (a1 >= a0 && nb > 0) ? (a1 - a0 + 1) * nb : 0;
            for (int c0 = 0; c0 < ncell && found < K; c0 += 256) {
                const int c = c0 + (int)threadIdx.x;
                bool hit = false;
                uint32_t rank = 0;
                if (c < ncell) {
                    const int a = a0 + c / nb, b = b0 + c % nb, q = (3 * a + pi) * f.W + 3 * b + pj;
                    if (pt_valid(cloud_meta[q])) {
                        const float4 p = cloud_pos[q];
                        uint32_t bx, by, bz;
                        hit = global_pos_to_voxel(f, p.x + 0.f, p.y + 0.f, p.z + 0.f, bx, by, bz) == v;
                        rank = (uint32_t)((ph * hc + a) * wc + b);
                    }
                }
                const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                if (lane_id() == 0) s_wcnt[wid] = (uint32_t)__popc(hits);
                __syncthreads();
                uint32_t before = 0, total = 0;
#pragma unroll
                for (uint32_t w = 0; w < 8; ++w) { const uint32_t t = s_wcnt[w]; before += w < wid ? t : 0u; total += t; }
                if (hit) {
                    const uint32_t pos = found + before + (uint32_t)__popc(hits & ((1u << lane_id()) - 1u));
                    if (pos < K) out[pos] = rank;
                }
                found += total;
                __syncthreads();
            }
        }
        for (uint32_t t = min(found, K) + threadIdx.x; t < K; t += 256) out[t] = NO_RANK;
    }
}

// the map's random cursor moves on by three draws per candidate of every valid pixel (nb > 1 only)
__global__ void k_birth_advance_cursor(const __grid_constant__ FrameParams f, const Counters* __restrict__ cnt, Persist* __restrict__ persist) {
    persist->cursor_map = (uint32_t)(((uint64_t)persist->cursor_map + 3ull * (uint32_t)f.nb_ptc * cnt->n_valid_points) % f.rand_n);
}

// ---- per-voxel slot state held in registers while candidates are replayed ------------------------------------------
template <int P>
struct VoxelSlots {
    uint64_t m[P];
    float w[P];
    uint32_t live, vac, upd;   // bit s: status != INVALID / vacant (operations.h:810-816) / status == UPDATED

    __device__ __forceinline__ void load(const Store& st, size_t base, uint32_t hdr, uint32_t stamp) {
        live = hdr_live(hdr);
        vac = 0; upd = 0;
        if (live) {
#pragma unroll
            for (int i = 0; i < P / 2; ++i) {
                const uint4 q = reinterpret_cast<const uint4*>(st.meta + base)[i];
                m[2 * i] = (uint64_t)q.x | ((uint64_t)q.y << 32);
                m[2 * i + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
            }
#pragma unroll
            for (int i = 0; i < P / 2; ++i) {
                const float2 q = reinterpret_cast<const float2*>(st.w + base)[i];
                w[2 * i] = q.x; w[2 * i + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) { m[i] = 0; w[i] = 0.f; }
        }
#pragma unroll
        for (int s = 1; s < P; ++s) {
            const bool lv = (live >> s) & 1u;
            if (!lv || meta_ts(m[s]) < stamp) vac |= 1u << s;
            if (lv && meta_status(m[s]) == ST_UPDATED) upd |= 1u << s;
        }
    }
    // first vacant slot 1..P-1 (operations.h:790-796); 0 when full
    __device__ __forceinline__ uint32_t first_vacant() const { return vac ? (uint32_t)(__ffs(vac) - 1) : 0u; }
    __device__ __forceinline__ void occupy(uint32_t s) { vac &= ~(1u << s); upd &= ~(1u << s); live |= 1u << s; }

    // SemanticDSPMap::resampleParticlesInVoxel (semantic_dsp_map.h:1448-1519); returns true when a resample ran
    __device__ __forceinline__ bool resample(const Store& st, size_t base) {
        const uint32_t trigger = P >> 1;
        if ((uint32_t)__popc(upd) <= trigger) return false;
        float wsum = 0.f;
#pragma unroll
        for (int s = 1; s < P; ++s)
            if ((upd >> s) & 1u) wsum += w[s];
        uint32_t kill = 0;
        if (wsum < 0.01f) {
            kill = upd;
        } else {
            float wpp = wsum / (float)trigger;
            if (wpp > 1.f) wpp = 1.f;
            float cum = 0.f, thr = wpp;
#pragma unroll
            for (int s = 1; s < P; ++s) {
                if (!((upd >> s) & 1u)) continue;
                cum += w[s];
                if (cum < thr) {
                    kill |= 1u << s;
                } else {
                    w[s] = wpp;
                    st.w[base + s] = wpp;
                    thr += wpp;
                    while (cum > thr) thr += wpp;
                }
            }
        }
#pragma unroll
        for (int s = 1; s < P; ++s)
            if ((kill >> s) & 1u) st.meta[base + s] = meta_with_status(m[s], ST_INVALID);
        live &= ~kill; upd &= ~kill; vac |= kill;
        return true;
    }
};

// =====================================================================================================
// B1  birth_replay    replaces the birth loop + addNewbornParticle[WithNoise]AndResample
//                     (semantic_dsp_map.h:768-800, 1148-1230) — one thread per touched voxel replays its candidates in
//                     reference order on a register copy of the slot state.  Accepted candidates are left in the
//                     voxel's rank list as (rank | slot << 24) for k_birth_write; the hash entry is released.
// =====================================================================================================
template <int PB>
__global__ void __launch_bounds__(128)
k_birth_replay(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
               const uint32_t* __restrict__ stamp_z, uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks,
               const uint32_t* __restrict__ entry_list, uint32_t* __restrict__ entry_voxel, Counters* __restrict__ cnt) {
    constexpr int P = 1 << PB, K = 2 * P - 1;
    const uint32_t n_entries = cnt->n_birth_voxels;
    for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_entries; li += gridDim.x * blockDim.x) {
        const uint32_t e = entry_list[li];
        const uint32_t v = keys[e];
        keys[e] = EMPTY_KEY;
        entry_voxel[li] = v;
        uint32_t* r = ranks + (size_t)e * K;
        uint32_t rk[K];
#pragma unroll
        for (int t = 0; t < K; ++t) rk[t] = r[t];
        uint32_t bx, by, bz;
        voxel_to_buffer(f, v, bx, by, bz);
        const uint32_t stamp = slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz);
        const size_t base = (size_t)v << PB;
        const uint32_t hdr = st.hdr[v];
        VoxelSlots<P> vs;
        vs.load(st, base, hdr, stamp);
        bool resampled = false, frozen = false;
        uint32_t born = 0, n_res = 0;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const uint32_t rank = rk[t];
            if (rank == NO_RANK) continue;
            if (frozen) { r[t] = NO_RANK; continue; }
            uint32_t s = vs.first_vacant();
            const bool try_resample = f.depth_noise ? ((s == 0) && !resampled)     // :1205-1226 only when the voxel is full
                                                    : !resampled;                  // :1165-1170 after every add
            if (f.depth_noise && try_resample) {
                if (vs.resample(st, base)) { resampled = true; ++n_res; s = vs.first_vacant(); }
                else frozen = true;            // full, and a resample can never trigger this frame
            }
            if (s) {
                vs.occupy(s);
                ++born;
            } else if (f.depth_noise) {
                frozen = true;                 // full after the one allowed resample
            }
            if (!f.depth_noise && try_resample) {
                if (vs.resample(st, base)) { resampled = true; ++n_res; }
                else if (s == 0) frozen = true;  // full and births never create UPDATED particles
            } else if (!f.depth_noise && s == 0) {
                frozen = true;
            }
            r[t] = s ? (rank | (s << 24)) : NO_RANK;
        }
        const uint32_t nh = (hdr & 0x1ffffu) | (vs.live << 16);
        if (nh != hdr) st.hdr[v] = nh;
        if (born) atomicAdd(&cnt->n_born, born);
        if (n_res) atomicAdd(&cnt->n_resampled, n_res);
    }
}

// one thread per (touched voxel, candidate): write the accepted newborn particles (operations.h:171-184)
__global__ void __launch_bounds__(256)
k_birth_write(const __grid_constant__ FrameParams f, Store st, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta,
              uint32_t* __restrict__ ranks, const uint32_t* __restrict__ entry_list, const uint32_t* __restrict__ entry_voxel,
              const Counters* __restrict__ cnt, const uint32_t* __restrict__ valid_prefix, const float* __restrict__ randoms,
              const Persist* __restrict__ persist) {
    const uint32_t K = f.topk, n = cnt->n_birth_voxels * K;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t li = t / K, i = t - li * K;
        uint32_t* r = ranks + (size_t)entry_list[li] * K + i;
        const uint32_t a = *r;
        if (a == NO_RANK) continue;
        *r = NO_RANK;
        const uint32_t s = a >> 24, v = entry_voxel[li];
        const uint32_t cand = a & 0xffffffu, nb = (uint32_t)f.nb_ptc, rank = cand / nb;
        const int q = birth_rank_to_pixel(f, rank);
        const float3 p = newborn_position(f, cloud_pos[q], nb > 1 ? valid_prefix[rank] : 0u, cand - rank * nb, randoms, persist->cursor_map);
        const uint32_t pm = cloud_meta[q];
        const size_t idx = ((size_t)v << f.Pb) + s;
        st.x[idx] = p.x; st.y[idx] = p.y; st.z[idx] = p.z;
        st.w[idx] = OCC_INIT_WEIGHT;
        st.meta[idx] = meta_pack(f.T, pt_track(pm), pt_label(pm), ST_REGULAR_BORN, 0);
        if (pt_track(pm) <= f.max_movable && !(st.hdr[v] & HDR_MOVABLE)) atomicOr(st.hdr + v, HDR_MOVABLE);
    }
}

// =====================================================================================================
// O1  occupancy_labels  replaces getOccupancyResult + determineIfVoxelOccupied + calculateWeightAndSemanticsInVoxel
//                       + getVoxelGlobalPosition (semantic_dsp_map.h:1239-1383; mc_ring/operations.h:390-448, 623-648)
// Four launches, no inter-block dependency inside any of them:
//   headers  : each thread streams OCC_VPT consecutive header words (one 16-byte load).  Unknown voxels and known-but-
//              empty voxels are decided from the header alone; voxels with live slots are appended to a work list.
//   slots    : one P-lane group per listed voxel: lane s holds slot s (coalesced 8 B tail + 4 B weight) and every lane
//              replays the reference's slot-order sums through shuffles (weight sum, guessed weight, summed weight of
//              its own track), so all floats are added in the reference order; the arg-max track is a butterfly over
//              (weight desc, track asc).  Output: one state byte per voxel, sparse per-voxel results for emitted
//              voxels, per-tile counts.
//   scan     : exclusive scan of the per-block counts.
//   emit     : ordered compaction (ascending storage index) of occupied / free voxels and of the live particles of
//              movable tracks — the next frame's object-move step works from that list instead of a map pass.
// =====================================================================================================
constexpr int OCC_BLOCK = 256;
constexpr int OCC_VPT = 4;                       // voxels per thread (one 16-byte header load)
constexpr int OCC_TILE = OCC_BLOCK * OCC_VPT;
constexpr int OCC_SCAN_TPT = 4;                  // tiles per thread of the tile-count scan
constexpr int OCC_SCAN_CHUNK = 1024 * OCC_SCAN_TPT;   // tiles per scan block
constexpr int MAX_P = 16;
static_assert(2 * MAX_P - 1 <= BG_LANES * BG_VALS, "k_birth_gather sorts a voxel's rank list in BG_VALS registers per lane");
// state byte: bits 0..1 = state + 1 (0 unknown, 1 free, 2 occupied, 3 guessed occupied), bit 2 = holds movable particles
constexpr uint32_t VS_MOVABLE = 4u;
constexpr uint32_t VS_ASSIGNS = 8u;     // the voxel's label histogram had a positive entry: it assigned (label, track), res_tl holds them

struct ResultArrays {
    uint32_t* voxel_index; float* xyz; float* weight_sum; uint16_t* track_id; uint8_t* label_id; int8_t* state;
    uint32_t* free_voxel_index; float* free_xyz;
    uint8_t* rgb;           // 3 bytes per occupied entry (written when FrameParams::col_enabled)
};

// ---- colour of one occupied voxel: getOccupancyResult, semantic_dsp_map.h:1274-1351 ------------------------------
// The two conversions restate cv::cvtColor on one CV_8UC3 pixel (H in [0,180); S and H through 12-bit fixed-point
// reciprocals, rounded to nearest like OpenCV's tables; back through float sectors with the two "1 - s*h" terms FUSED
// and the final scaling truncated) — bit-identical to cv2 4.13 on all 2^24 inputs (tests/test_colour_cpu.py pins the
// oracle's copy of the same arithmetic, tests/test_colour_gpu.py this one).
__device__ __forceinline__ int fixed_recip12(int num, int den) {   // lrint((num << 12) / (double)den), den >= 1
    // the quotient is never at an exact .5 (num << 12 is even, ties need an odd numerator over 2*den), so
    // round-half-up in integers equals the reference's round-half-even
    const long long n2 = ((long long)num << 13) + den;
    return (int)(n2 / (2 * den));
}
__device__ __forceinline__ void rgb_to_hsv_u8(int r, int g, int b, int& h, int& s, int& v) {
    v = max(r, max(g, b));
    const int vmin = min(r, min(g, b)), diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    const int sdiv = v ? fixed_recip12(255, v) : 0;
    const int hdiv = diff ? fixed_recip12(180, 6 * diff) : 0;
    s = (diff * sdiv + (1 << 11)) >> 12;
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * hdiv + (1 << 11)) >> 12;
    h += h < 0 ? 180 : 0;
}
__device__ __forceinline__ void hsv_to_rgb_u8(int hh, int ss, int vv, int& r, int& g, int& b) {
    float h = (float)hh;
    const float s = __fmul_rn((float)ss, 1.f / 255.f), v = __fmul_rn((float)vv, 1.f / 255.f);
    float fr, fg, fb;
    if (ss == 0) {
        fr = fg = fb = v;
    } else {
        h = __fmul_rn(h, 6.f / 180.f);
        while (h >= 6.f) h = __fsub_rn(h, 6.f);
        int sector = (int)floorf(h);
        h = __fsub_rn(h, (float)sector);
        if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
        const float t0 = v, t1 = __fmul_rn(v, __fsub_rn(1.f, s)), t2 = __fmul_rn(v, __fmaf_rn(-s, h, 1.f)),
                    t3 = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.f, h), 1.f));
        // sector table of OpenCV: (b, g, r) = tab[{1,3,0}, {1,0,2}, {3,0,1}, {0,2,1}, {0,1,3}, {2,1,0}]
        switch (sector) {
            case 0: fb = t1; fg = t3; fr = t0; break;
            case 1: fb = t1; fg = t0; fr = t2; break;
            case 2: fb = t3; fg = t0; fr = t1; break;
            case 3: fb = t0; fg = t2; fr = t1; break;
            case 4: fb = t0; fg = t1; fr = t3; break;
            default: fb = t2; fg = t1; fr = t0; break;
        }
    }
    r = min(max((int)floorf(__fmul_rn(fr, 255.f)), 0), 255);
    g = min(max((int)floorf(__fmul_rn(fg, 255.f)), 0), 255);
    b = min(max((int)floorf(__fmul_rn(fb, 255.f)), 0), 255);
}
// col_tab: 768 bytes label colours (b, g, r per label) followed by the 256-entry instance map
__device__ __forceinline__ uint32_t voxel_colour(const FrameParams& f, const uint8_t* __restrict__ col_tab, int state, uint32_t label, uint32_t track,
                                                 float px, float py, float pz, float cx, float cy, float cz) {
    int r, g, b;
    if (state == 1) {
        if ((int)label == f.col_background) {                       // :1277-1294 jet by height (color_map_jet_256_, :50-63)
            const float hgt = f.col_scheme == 3 ? (py + 2.f) : (-pz + 2.f);
            const int ci = min(max((int)(hgt * 51.2f), 0), 255);
            if (ci < 64) { r = 0; g = 0; b = ci * 4; }
            else if (ci < 128) { r = 0; g = (ci - 64) * 4; b = 255; }
            else if (ci < 192) { r = (ci - 128) * 4; g = 255; b = 255 - (ci - 128) * 4; }
            else { r = 255; g = 255 - (ci - 192) * 4; b = 0; }
            if (f.col_eval) r = g = b = 0;
        } else if (f.col_scheme == 0 || track > f.max_movable) {    // :1297-1309 label colour, stored (b, g, r)
            b = col_tab[3 * label]; g = col_tab[3 * label + 1]; r = col_tab[3 * label + 2];
        } else if (f.col_eval) {                                    // :1311-1315
            r = (int)label; g = (int)(track >> 8); b = (int)(track & 0xffu);
        } else {                                                    // :1317-1319 (table index taken modulo 256)
            r = 160; g = col_tab[768 + (track & 0xffu)]; b = col_tab[768 + label];
        }
    } else {                                                        // guessed occupied: white, :1325-1331
        r = g = b = 255;
    }
    if (!f.col_eval) {                                              // :1333-1351
        int h, s, v;
        rgb_to_hsv_u8(r, g, b, h, s, v);
        if (!point_in_frustum(f, cx, cy, cz)) v = (int)((float)v * 0.7f) & 0xff;
        hsv_to_rgb_u8(h, s, v, r, g, b);
    }
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

struct OccScratch {
    uint8_t* vstate;        // V bytes
    float* res_w;           // V, written for emitted voxels only
    uint32_t* res_tl;       // V, track | label << 16
    uint16_t* mov_mask;     // V, slots holding live particles of movable tracks (written where VS_MOVABLE)
    uint4* blk_cnt;         // per tile: occupied, free, movable particles
    uint4* blk_off;         // exclusive prefix of blk_cnt inside the tile's scan chunk (OCC_SCAN_CHUNK tiles)
    uint4* chunk_tot;       // per scan chunk: totals (x, y, z) and number of non-empty tiles (w); one extra entry at the end:
                            // x = ticket counter of the "last block finishes" step, y = total number of non-empty tiles
    uint4* chunk_base;      // per scan chunk: exclusive prefix of chunk_tot
    uint32_t* tile_list;    // non-empty tiles, ascending inside each chunk's own segment [chunk * OCC_SCAN_CHUNK, ...)
};

// persistent warps, each streaming 128-voxel pieces (32 lanes x one 16-byte load) of the header plane with the next
// piece's load already in flight; no block-level synchronisation.  blk_cnt must be zero on entry (k_occ_scan leaves it so).
template <int PB>
__global__ void __launch_bounds__(OCC_BLOCK)
k_occ_headers(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
              const uint32_t* __restrict__ stamp_z, OccScratch sc, uint32_t* __restrict__ work, uint32_t n_tiles, Counters* __restrict__ cnt) {
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_pieces = n_tiles * (OCC_TILE / 128);
    uint32_t piece = warp;
    uint4 h4 = make_uint4(0, 0, 0, 0);
    if (piece < n_pieces) h4 = *reinterpret_cast<const uint4*>(st.hdr + v_lo + piece * 128 + lane_id() * 4);
    for (; piece < n_pieces; piece += n_warps) {
        const uint32_t v0 = v_lo + piece * 128 + lane_id() * 4;
        const uint32_t hdrs[OCC_VPT] = {h4.x, h4.y, h4.z, h4.w};
        if (piece + n_warps < n_pieces) h4 = *reinterpret_cast<const uint4*>(st.hdr + v0 + (size_t)n_warps * 128);   // prefetch
        // never-observed voxels (header 0: no time stamp, no live slot) are unknown whatever the slab stamps say
        const bool any_set = (hdrs[0] | hdrs[1] | hdrs[2] | hdrs[3]) != 0;
        if (!__any_sync(0xffffffffu, any_set)) {
            *reinterpret_cast<uint32_t*>(sc.vstate + (v0 - v_lo)) = 0u;
            continue;
        }
        uint32_t sb4 = 0, heavy = 0, n_free = 0;
        if (any_set) {
            uint32_t bx0, by, bz;
            voxel_to_buffer(f, v0, bx0, by, bz);
            const uint32_t stamp_yz = max(__ldg(stamp_y + by), __ldg(stamp_z + bz));
#pragma unroll
            for (int u = 0; u < OCC_VPT; ++u) {
                const uint32_t hdr = hdrs[u];
                const uint32_t stamp = max(stamp_yz, __ldg(stamp_x + bx0 + u));
                const uint32_t ts0 = hdr_ts(hdr);
                const bool known = ts0 != 0 && !(ts0 < stamp);                  // isVoxelValid, operations.h:824-837
                const bool hv = hdr_live(hdr) != 0 && (known || (hdr & HDR_MOVABLE));
                // decided by the header alone: unknown (-1), or known and empty (0: weight sum 0, no guessed weight)
                sb4 |= (known ? 1u : 0u) << (8 * u);
                heavy |= (hv ? 1u : 0u) << u;
                n_free += known && !hv;
            }
        }
        *reinterpret_cast<uint32_t*>(sc.vstate + (v0 - v_lo)) = sb4;
        if (f.want_free) {
            n_free = __reduce_add_sync(0xffffffffu, n_free);
            if (lane_id() == 0 && n_free) atomicAdd(reinterpret_cast<uint32_t*>(sc.blk_cnt + piece / (OCC_TILE / 128)) + 1, n_free);
        }
        if (!__any_sync(0xffffffffu, heavy != 0)) continue;
        // append this warp's heavy voxels to the work list (one atomic per warp)
        const uint32_t nh = __popc(heavy);
        uint32_t inc = nh;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane_id() >= (uint32_t)o) inc += n;
        }
        const uint32_t warp_total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t base = 0;
        if (lane_id() == 31) base = atomicAdd(&cnt->n_occ_work, warp_total);
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t pos = base + inc - nh;
#pragma unroll
        for (int u = 0; u < OCC_VPT; ++u)
            if ((heavy >> u) & 1u) work[pos++] = v0 + u;
    }
}

// ---- the same pass with the header plane staged through shared memory by the bulk-copy (TMA) unit ----------------------
// The plane is the one thing the engine streams whole from HBM (4 B per voxel, 134 MB at 512 x 128 x 512), and a load per
// thread with one prefetch keeps 64 KB per SM in flight - short of what HBM3e needs at full rate.  Here every CTA runs a ring of
// OHB_STAGES chunks of OHB_CHUNK voxels: one elected thread issues cp.async.bulk global -> shared for the chunk OHB_STAGES ahead,
// completion is counted in bytes on the stage's mbarrier, the CTA's threads wait on the barrier's phase and read their header
// words from shared memory (16-byte, conflict-free).  Four CTAs per SM x 32 KB = 128 KB per SM in flight, issued by four threads.
constexpr int OHB_STAGES = 4;
constexpr int OHB_CHUNK = 2048;                  // voxels per stage (8 KB)
constexpr int OHB_BLOCK = 256;
static_assert(OCC_TILE % OHB_CHUNK == 0 || OHB_CHUNK % OCC_TILE == 0, "chunks and tiles nest");

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_global_to_shared(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_BAR:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_BAR;\n"
        "bra WAIT_BAR;\n"
        "DONE_BAR:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

template <int PB>
__global__ void __launch_bounds__(OHB_BLOCK)
k_occ_headers_bulk(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
                   const uint32_t* __restrict__ stamp_z, OccScratch sc, uint32_t* __restrict__ work, uint32_t n_tiles, Counters* __restrict__ cnt) {
    __shared__ __align__(128) uint32_t s_hdr[OHB_STAGES][OHB_CHUNK];
    __shared__ __align__(8) uint64_t s_bar[OHB_STAGES];
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb);
    const uint32_t n_vox = n_tiles * OCC_TILE;                               // the plane is padded to whole tiles
    const uint32_t n_chunks = (n_vox + OHB_CHUNK - 1) / OHB_CHUNK;
    const uint32_t my_chunks = blockIdx.x < n_chunks ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar), buf0 = (uint32_t)__cvta_generic_to_shared(s_hdr);
    auto issue = [&](uint32_t k) {       // chunk number k of this CTA into its stage (one thread)
        const uint32_t chunk = blockIdx.x + k * gridDim.x, stage = k % OHB_STAGES;
        const uint32_t vox0 = chunk * OHB_CHUNK, bytes = min((uint32_t)OHB_CHUNK, n_vox - vox0) * 4u;
        mbar_expect_tx(bar0 + 8 * stage, bytes);
        bulk_global_to_shared(buf0 + stage * (OHB_CHUNK * 4), st.hdr + v_lo + vox0, bytes, bar0 + 8 * stage);
    };
    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < OHB_STAGES; ++s2) mbar_init(bar0 + 8 * s2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (uint32_t k = 0; k < min(my_chunks, (uint32_t)OHB_STAGES); ++k) issue(k);
    for (uint32_t k = 0; k < my_chunks; ++k) {
        const uint32_t chunk = blockIdx.x + k * gridDim.x, stage = k % OHB_STAGES;
        const uint32_t vox0 = chunk * OHB_CHUNK, n_here = min((uint32_t)OHB_CHUNK, n_vox - vox0);
        mbar_wait(bar0 + 8 * stage, (k / OHB_STAGES) & 1u);
#pragma unroll
        for (int j = 0; j < OHB_CHUNK / (OHB_BLOCK * OCC_VPT); ++j) {
            const uint32_t q = (uint32_t)j * OHB_BLOCK + threadIdx.x;        // 16-byte group inside the chunk (whole warps in or out)
            if (q * OCC_VPT >= n_here) break;
            const uint4 h4 = *reinterpret_cast<const uint4*>(&s_hdr[stage][q * OCC_VPT]);
            const uint32_t v0 = v_lo + vox0 + q * OCC_VPT;
            const uint32_t hdrs[OCC_VPT] = {h4.x, h4.y, h4.z, h4.w};
            // never-observed voxels (header 0: no time stamp, no live slot) are unknown whatever the slab stamps say
            const bool any_set = (hdrs[0] | hdrs[1] | hdrs[2] | hdrs[3]) != 0;
            if (!__any_sync(0xffffffffu, any_set)) {
                *reinterpret_cast<uint32_t*>(sc.vstate + (v0 - v_lo)) = 0u;
                continue;
            }
            uint32_t sb4 = 0, heavy = 0, n_free = 0;
            if (any_set) {
                uint32_t bx0, by, bz;
                voxel_to_buffer(f, v0, bx0, by, bz);
                const uint32_t stamp_yz = max(__ldg(stamp_y + by), __ldg(stamp_z + bz));
#pragma unroll
                for (int u = 0; u < OCC_VPT; ++u) {
                    const uint32_t hdr = hdrs[u];
                    const uint32_t stamp = max(stamp_yz, __ldg(stamp_x + bx0 + u));
                    const uint32_t ts0 = hdr_ts(hdr);
                    const bool known = ts0 != 0 && !(ts0 < stamp);                  // isVoxelValid, operations.h:824-837
                    const bool hv = hdr_live(hdr) != 0 && (known || (hdr & HDR_MOVABLE));
                    sb4 |= (known ? 1u : 0u) << (8 * u);
                    heavy |= (hv ? 1u : 0u) << u;
                    n_free += known && !hv;
                }
            }
            *reinterpret_cast<uint32_t*>(sc.vstate + (v0 - v_lo)) = sb4;
            if (f.want_free) {
                n_free = __reduce_add_sync(0xffffffffu, n_free);
                if (lane_id() == 0 && n_free) atomicAdd(reinterpret_cast<uint32_t*>(sc.blk_cnt + (v0 - v_lo) / OCC_TILE) + 1, n_free);
            }
            if (!__any_sync(0xffffffffu, heavy != 0)) continue;
            const uint32_t nh = __popc(heavy);
            uint32_t inc = nh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane_id() >= (uint32_t)o) inc += n;
            }
            const uint32_t warp_total = __shfl_sync(0xffffffffu, inc, 31);
            uint32_t base = 0;
            if (lane_id() == 31) base = atomicAdd(&cnt->n_occ_work, warp_total);
            base = __shfl_sync(0xffffffffu, base, 31);
            uint32_t pos = base + inc - nh;
#pragma unroll
            for (int u = 0; u < OCC_VPT; ++u)
                if ((heavy >> u) & 1u) work[pos++] = v0 + u;
        }
        __syncthreads();                                                     // the stage has been read by everyone
        if (threadIdx.x == 0 && k + OHB_STAGES < my_chunks) issue(k + OHB_STAGES);
    }
}

template <int PB>
__global__ void __launch_bounds__(256)
k_occ_slots(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
            const uint32_t* __restrict__ stamp_z, OccScratch sc, const uint32_t* __restrict__ work, Counters* __restrict__ cnt) {
    constexpr int P = 1 << PB, G = 32 / P;
    const uint32_t n_work = cnt->n_occ_work;
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb);
    const uint32_t sl = lane_id() & (P - 1), g = lane_id() / P, gbase = lane_id() & ~(uint32_t)(P - 1);
    const uint32_t gmask = group_mask<P>();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t n_nonempty = 0;
    for (uint32_t w0 = warp * G; w0 < n_work; w0 += n_warps * G) {     // warp-uniform trip count
        const uint32_t wi = w0 + g;
        uint32_t g_v = wi < n_work ? work[wi] : 0xffffffffu;
        const bool active = g_v != 0xffffffffu;                 // sentinel: unused tail of a reserved chunk
        if (!active) g_v = v_lo;
        const uint32_t g_hdr = active ? st.hdr[g_v] : 0u;
        const size_t base = (size_t)g_v << PB;
        const bool has = active && sl != 0 && ((g_hdr >> (16 + sl)) & 1u);
        uint64_t m = 0;
        float w = 0.f;
        if (has) { m = st.meta[base + sl]; w = st.w[base + sl]; }
        uint32_t bx, by, bz;
        voxel_to_buffer(f, g_v, bx, by, bz);
        const uint32_t g_stamp = slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz);
        const uint32_t ts0 = hdr_ts(g_hdr);
        const bool g_known = ts0 != 0 && !(ts0 < g_stamp);
        const bool nonvac = has && g_known && !(meta_ts(m) < g_stamp);   // isParticleVacant, operations.h:810-816
        float wc = w;
        if (nonvac && w > 1.f) { wc = 1.f; st.w[base + sl] = 1.f; }     // :416-419 clamp the stored weight
        const uint32_t status = meta_status(m);
        const bool guessed = nonvac && status == ST_GUESSED_BORN;
        const bool pruned = nonvac && status == ST_UPDATED && wc < OCC_INIT_WEIGHT;   // :423-427
        if (pruned) st.meta[base + sl] = meta_with_status(m, ST_INVALID);
        const bool included = nonvac && !pruned;
        const uint32_t my_track = meta_track(m);
        const bool movable = has && !pruned && my_track <= f.max_movable;   // still live after this pass
        const uint32_t packed = my_track | ((meta_label(m) & 0xffu) << 16) | ((uint32_t)nonvac << 24) | ((uint32_t)guessed << 25) | ((uint32_t)included << 26);
        // replay the slot loop: every lane accumulates the voxel sums and the sum of its own track, in slot order
        float ws = 0.f, gs = 0.f, tot = 0.f;
        uint32_t last_label = 0;
#pragma unroll
        for (int s2 = 1; s2 < P; ++s2) {
            const float w2 = __shfl_sync(0xffffffffu, w, gbase + s2), wc2 = __shfl_sync(0xffffffffu, wc, gbase + s2);
            const uint32_t p2 = __shfl_sync(0xffffffffu, packed, gbase + s2);
            if ((p2 >> 24) & 1u) ws += w2;                          // the unclamped weight, :414
            if ((p2 >> 25) & 1u) gs += wc2;
            if (((p2 >> 26) & 1u) && (p2 & 0xffffu) == my_track) { tot += wc2; last_label = (p2 >> 16) & 0xffu; }
        }
        // arg-max over the included slots: largest summed weight, lowest track on ties, strictly positive
        // (std::map ascending iteration with '>' from 0, :439-446)
        float best_w = included ? tot : -1.f;
        uint32_t best_tl = my_track | (last_label << 16);
#pragma unroll
        for (int o = 1; o < P; o <<= 1) {
            const float ow = __shfl_xor_sync(0xffffffffu, best_w, o);
            const uint32_t otl = __shfl_xor_sync(0xffffffffu, best_tl, o);
            if (ow > best_w || (ow == best_w && (otl & 0xffffu) < (best_tl & 0xffffu))) { best_w = ow; best_tl = otl; }
        }
        const uint32_t b_pruned = __ballot_sync(0xffffffffu, pruned), b_nonvac = __ballot_sync(0xffffffffu, nonvac);
        const uint32_t b_mov = __ballot_sync(0xffffffffu, movable);
        if (active && sl == 0) {
            uint32_t sb = 0;
            uint32_t* tile_cnt = reinterpret_cast<uint32_t*>(sc.blk_cnt + (g_v - v_lo) / OCC_TILE);
            if (g_known) {
                int state;
                if (ws > f.occ_thr) state = 1;
                else if (gs >= OCC_INIT_WEIGHT) state = 2;
                else state = 0;
                sb = (uint32_t)(state + 1);
                // label / track are out-parameters the reference assigns only when the histogram has a positive entry
                // (operations.h:439-446); otherwise the caller's variables keep what an earlier voxel left (see k_occ_emit)
                if (best_w > 0.f) { sb |= VS_ASSIGNS; sc.res_tl[g_v] = best_tl; }
                if (state > 0) { sc.res_w[g_v] = ws; atomicAdd(tile_cnt + 0, 1u); }
                else if (f.want_free) atomicAdd(tile_cnt + 1, 1u);
                n_nonempty += (b_nonvac & gmask) != 0;
            }
            const uint32_t mv = (b_mov & gmask) >> gbase;
            if (mv) { sc.mov_mask[g_v] = (uint16_t)mv; sb |= VS_MOVABLE; atomicAdd(tile_cnt + 2, (uint32_t)__popc(mv)); }
            sc.vstate[g_v - v_lo] = (uint8_t)sb;
            const uint32_t nh = hdr_ts(g_hdr) | (mv ? HDR_MOVABLE : 0u) | ((hdr_live(g_hdr) & ~((b_pruned & gmask) >> gbase)) << 16);
            if (nh != g_hdr) st.hdr[g_v] = nh;
        }
    }
    n_nonempty = __reduce_add_sync(0xffffffffu, n_nonempty);
    if (lane_id() == 0 && n_nonempty) atomicAdd(&cnt->n_nonempty_voxels, n_nonempty);
}

// exclusive scan of the per-block (occupied, free, movable) counts; single block, any number of tiles
__global__ void __launch_bounds__(1024)
k_occ_scan(OccScratch sc, uint32_t n_blocks, Counters* __restrict__ cnt, Persist* __restrict__ persist, CommView cv, uint32_t frame_stamp,
           Counters* __restrict__ host_cnt) {
    // one block per chunk of OCC_SCAN_CHUNK tiles (a 256 x 64 x 256 map is one chunk, 1024 x 128 x 1024 thirty-two): offsets
    // inside the chunk, the chunk's totals, and the block that finishes last turns the totals into chunk bases
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_last;
    constexpr int TPT = OCC_SCAN_TPT;
    const uint32_t b0 = blockIdx.x * OCC_SCAN_CHUNK + threadIdx.x * TPT;
    uint4 v[TPT];
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
        v[i] = (b0 + i) < n_blocks ? sc.blk_cnt[b0 + i] : make_uint4(0, 0, 0, 0);
        if ((b0 + i) < n_blocks) sc.blk_cnt[b0 + i] = make_uint4(0, 0, 0, 0);      // read once per frame, here: zero again for the next frame's header pass
        s0 += v[i].x; s1 += v[i].y; s2 += v[i].z; s3 += (v[i].x | v[i].y | v[i].z) != 0;
    }
    uint32_t t0, t1, t2, t3;
    uint32_t e0 = block_exclusive_scan(s0, s_warp, t0);
    uint32_t e1 = block_exclusive_scan(s1, s_warp, t1);
    uint32_t e2 = block_exclusive_scan(s2, s_warp, t2);
    const bool multi = gridDim.x > 1;                         // a small map is one chunk: no list, no hand-over between blocks
    uint32_t e3 = 0;
    t3 = 0;
    if (multi) e3 = block_exclusive_scan(s3, s_warp, t3);
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
        if ((b0 + i) < n_blocks) sc.blk_off[b0 + i] = make_uint4(e0, e1, e2, (v[i].x | v[i].y | v[i].z) != 0);   // .w: the tile emits something
        if (multi && (v[i].x | v[i].y | v[i].z) != 0) sc.tile_list[blockIdx.x * OCC_SCAN_CHUNK + e3++] = b0 + i;   // the emit pass visits only these
        e0 += v[i].x; e1 += v[i].y; e2 += v[i].z;
    }
    uint32_t* tail = reinterpret_cast<uint32_t*>(sc.chunk_tot + gridDim.x);
    if (threadIdx.x == 0) {
        sc.chunk_tot[blockIdx.x] = make_uint4(t0, t1, t2, t3);
        s_last = 1u;
        if (multi) {
            __threadfence();
            const uint32_t ticket = atomicAdd(tail, 1u);
            s_last = ticket == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        uint32_t run0 = 0, run1 = 0, run2 = 0, run3 = 0;
        for (uint32_t c = 0; c < gridDim.x; ++c) {
            const volatile uint32_t* tot = reinterpret_cast<const volatile uint32_t*>(sc.chunk_tot + c);
            sc.chunk_base[c] = make_uint4(run0, run1, run2, run3);
            run0 += tot[0]; run1 += tot[1]; run2 += tot[2]; run3 += tot[3];
        }
        tail[0] = 0u;                                             // the ticket counter is ready for the next frame
        tail[1] = run3;                                           // length of the non-empty tile list
        cnt->n_occupied = run0; cnt->n_free = run1; cnt->n_movable = run2; persist->n_movable = run2;
        cnt->seq_head = frame_stamp; cnt->seq_tail = frame_stamp;    // sharded maps: copied to the host after the last kernel of the frame
        for (int p = 0; p < cv.size; ++p) {   // every rank learns every rank's totals (the emit pass needs the lower ranks')
            uint32_t* row = cv.counts[p] + cv.rank * COMM_COUNTS;
            row[CC_OCC] = run0; row[CC_FREE] = run1;
        }
    }
    if (host_cnt) {
        // Unsharded map: nothing after this point changes a counter (the emit pass reads the scan's tables only), so the block goes
        // to the host from here - plain stores into the handle's pinned copy, body first, the two stamps the host polls for
        // last - and is zeroed for the next frame.  No copy-engine operation at either end of a frame.  (One warp, two words
        // per lane: the loads come from L2 - other kernels' atomics live there - and should not queue up behind one thread.)
        __syncthreads();                                          // thread 0 has written its totals (the branch is block-uniform)
        if (threadIdx.x < 32) {
            constexpr int NW = sizeof(Counters) / 4;
            static_assert(offsetof(Counters, seq_head) == 0 && offsetof(Counters, seq_tail) == (NW - 1) * 4 && NW - 2 <= 64, "stamps at both ends, body of at most 64 words");
            volatile uint32_t* src = reinterpret_cast<volatile uint32_t*>(cnt);
            volatile uint32_t* dst = reinterpret_cast<volatile uint32_t*>(host_cnt);
            const int k0 = 1 + (int)threadIdx.x, k1 = 33 + (int)threadIdx.x;
            const uint32_t w0 = k0 < NW - 1 ? src[k0] : 0u, w1 = k1 < NW - 1 ? src[k1] : 0u;
            if (k0 < NW - 1) dst[k0] = w0;
            if (k1 < NW - 1) dst[k1] = w1;
            __threadfence_system();
            __syncwarp();
            if (threadIdx.x == 0) { dst[0] = frame_stamp; dst[NW - 1] = frame_stamp; }
            for (int i = (int)threadIdx.x; i < NW; i += 32) src[i] = 0u;
        }
    }
    if (cv.size > 1) __threadfence_system();      // sharded map: peers read what this kernel wrote after the next barrier
}

// one tile (OCC_TILE voxels) of the ordered emission; every thread of the block takes part in the two scans
template <int PB>
__device__ __forceinline__ void occ_emit_tile(const FrameParams& f, const Store& st, const OccScratch& sc, const ResultArrays& out,
                                              uint32_t* __restrict__ movl_idx, uint16_t* __restrict__ movl_track, uint32_t movl_cap,
                                              const uint8_t* __restrict__ col_tab, uint32_t tile, uint32_t base_occ, uint32_t base_free, uint32_t* s_warp, bool use_list) {
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb);
    const uint32_t v0 = v_lo + tile * OCC_TILE + threadIdx.x * OCC_VPT;
    const uint32_t sb4 = *reinterpret_cast<const uint32_t*>(sc.vstate + (v0 - v_lo));
    uint32_t n_occ = 0, n_free = 0, n_mov = 0;
    uint32_t mmask[OCC_VPT];
#pragma unroll
    for (int u = 0; u < OCC_VPT; ++u) {
        const uint32_t sb = (sb4 >> (8 * u)) & 0xffu, stt = sb & 3u;
        n_occ += stt >= 2;
        n_free += (f.want_free && stt == 1);
        mmask[u] = (sb & VS_MOVABLE) ? (uint32_t)sc.mov_mask[v0 + u] : 0u;
        n_mov += __popc(mmask[u]);
    }
    uint32_t tot_a, tot_b;
    const uint32_t ex_a = block_exclusive_scan(n_occ | (n_free << 16), s_warp, tot_a);   // <= 1024 each
    const uint32_t ex_b = block_exclusive_scan(n_mov, s_warp, tot_b);
    if ((n_occ | n_free | n_mov) == 0) return;                // after the scans: no barrier follows inside this tile
    const uint4 off = sc.blk_off[tile];
    uint4 cb = make_uint4(0u, 0u, 0u, 0u);
    if (use_list) cb = sc.chunk_base[tile / OCC_SCAN_CHUNK];   // a one-chunk map has base 0
    uint32_t pos_occ = base_occ + cb.x + off.x + (ex_a & 0xffffu), pos_free = base_free + cb.y + off.y + (ex_a >> 16), pos_mov = cb.z + off.z + ex_b;
    uint32_t bx0, by, bz;
    voxel_to_buffer(f, v0, bx0, by, bz);
#pragma unroll
    for (int u = 0; u < OCC_VPT; ++u) {
        const uint32_t stt = (sb4 >> (8 * u)) & 3u, v = v0 + u;
        const bool is_occ = stt >= 2, is_free = f.want_free && stt == 1;
        if (is_occ || is_free) {
            // voxelIdxToMapFramePos + map centre (operations.h:940-944, 970-983)
            uint32_t mx, my, mz;
            buffer_to_map(f, bx0 + u, by, bz, mx, my, mz);
            float p0 = (float)mx * f.voxel + f.pmin[0], p1 = (float)my * f.voxel + f.pmin[1], p2 = (float)mz * f.voxel + f.pmin[2];
            p0 += f.map_center[0]; p1 += f.map_center[1]; p2 += f.map_center[2];
            const float c0 = p0, c1 = p1, c2 = p2;                  // the voxel corner in the global frame
            if (f.zero_center) { p0 = p0 - f.cam[0]; p1 = p1 - f.cam[1]; p2 = p2 - f.cam[2]; }
            if (is_occ) {
                const uint32_t pos = pos_occ++;
                // getOccupancyResult declares label / track once per loop iteration WITHOUT initialising them
                // (semantic_dsp_map.h:1245-1246) and determineIfVoxelOccupied leaves them untouched when every particle of the
                // voxel was pruned in the same call: a compiled reference then emits what the most recent voxel in storage
                // order assigned - an unknown voxel assigns (0, 0) (operations.h:402-408), a voxel with a positive histogram
                // entry its arg-max.  Rare (a handful of voxels per frame at low thresholds): walk back over the state bytes.
                // A sharded map does not look across its slab boundary (defined as (0, 0) there, like the very first voxel).
                uint32_t tl = 0;
                if (((sb4 >> (8 * u)) & VS_ASSIGNS) != 0) tl = sc.res_tl[v];
                else {
                    for (uint32_t b = v; b > v_lo;) {
                        --b;
                        const uint32_t sbb = sc.vstate[b - v_lo];
                        if ((sbb & 3u) == 0u) break;                                 // unknown voxel: (0, 0)
                        if (sbb & VS_ASSIGNS) { tl = sc.res_tl[b]; break; }
                    }
                }
                out.voxel_index[pos] = v;
                out.xyz[3 * (size_t)pos + 0] = p0; out.xyz[3 * (size_t)pos + 1] = p1; out.xyz[3 * (size_t)pos + 2] = p2;
                out.weight_sum[pos] = sc.res_w[v];
                out.track_id[pos] = (uint16_t)(tl & 0xffffu);
                out.label_id[pos] = (uint8_t)((tl >> 16) & 0xffu);
                out.state[pos] = (int8_t)((int)stt - 1);
                if (f.col_enabled) {
                    const uint32_t c = voxel_colour(f, col_tab, (int)stt - 1, (tl >> 16) & 0xffu, tl & 0xffffu, p0, p1, p2, c0, c1, c2);
                    out.rgb[3 * (size_t)pos + 0] = (uint8_t)c; out.rgb[3 * (size_t)pos + 1] = (uint8_t)(c >> 8); out.rgb[3 * (size_t)pos + 2] = (uint8_t)(c >> 16);
                }
            } else {
                const uint32_t pos = pos_free++;
                out.free_voxel_index[pos] = v;
                out.free_xyz[3 * (size_t)pos + 0] = p0; out.free_xyz[3 * (size_t)pos + 1] = p1; out.free_xyz[3 * (size_t)pos + 2] = p2;
            }
        }
        uint32_t mm = mmask[u];
        while (mm) {
            const uint32_t s = __ffs(mm) - 1;
            mm &= mm - 1;
            const size_t idx = ((size_t)v << PB) + s;
            if (pos_mov < movl_cap) { movl_idx[pos_mov] = (uint32_t)idx; movl_track[pos_mov] = (uint16_t)meta_track(st.meta[idx]); }
            ++pos_mov;
        }
    }
}

// Small maps (one scan chunk): one block per tile, empty tiles leave after one load - every tile in flight at once.
// Large maps (mostly unknown or empty tiles, where launching a block per tile costs more than the work): a single wave of
// blocks takes the non-empty tiles round-robin from the list the scan left; list entry g lives in the segment of the chunk
// whose base covers it.
template <int PB>
__global__ void __launch_bounds__(OCC_BLOCK)
k_occ_emit(const __grid_constant__ FrameParams f, Store st, OccScratch sc, ResultArrays out, uint32_t* __restrict__ movl_idx,
           uint16_t* __restrict__ movl_track, uint32_t movl_cap, CommView cv, const uint32_t* __restrict__ counts,
           const uint8_t* __restrict__ col_tab, uint32_t n_chunks, int use_list) {
    __shared__ uint32_t s_warp[33];
    // sharded map: `out` is rank 0's result arrays; this rank's entries start after the lower ranks' (ascending storage index)
    uint32_t base_occ = 0, base_free = 0;
    for (int g = 0; g < cv.rank; ++g) { base_occ += counts[g * COMM_COUNTS + CC_OCC]; base_free += counts[g * COMM_COUNTS + CC_FREE]; }
    if (!use_list) {
        if (sc.blk_off[blockIdx.x].w == 0) return;            // nothing to emit from this tile
        occ_emit_tile<PB>(f, st, sc, out, movl_idx, movl_track, movl_cap, col_tab, blockIdx.x, base_occ, base_free, s_warp, false);
        if (cv.size > 1) __threadfence_system();      // sharded map: peers read what this kernel wrote after the next barrier
        return;
    }
    const uint32_t n_list = reinterpret_cast<const uint32_t*>(sc.chunk_tot + n_chunks)[1];
    uint32_t c = 0;
    for (uint32_t g = blockIdx.x; g < n_list; g += gridDim.x) {
        while (c + 1 < n_chunks && sc.chunk_base[c + 1].w <= g) ++c;                  // g only grows
        const uint32_t tile = sc.tile_list[c * OCC_SCAN_CHUNK + (g - sc.chunk_base[c].w)];
        occ_emit_tile<PB>(f, st, sc, out, movl_idx, movl_track, movl_cap, col_tab, tile, base_occ, base_free, s_warp, true);
        __syncthreads();      // s_warp is reused by the next tile
    }
    if (cv.size > 1) __threadfence_system();      // sharded map: peers read what this kernel wrote after the next barrier
}

// =====================================================================================================
// P2/P3  object moves  replaces moveParticlesInSetsByTransformations (mc_ring/operations.h:321-362) and
//                      ObjectSet::removeObjectByTrackID (object_layer.h:414-425).
// set(T) = { p : status(p) != INVALID and track(p) == T } in ascending particle index (canonical order), read from
// the movable-particle list the previous frame's occupancy pass left behind (nothing touches the store in between).
// =====================================================================================================
constexpr int MAX_MOVES = 256;           // moved tracks per frame; removals are batched, MAX_MOVES per batch (no limit per frame)
constexpr uint32_t SEQ_REMOVE = 0xfffffffeu;

struct MoveTable {
    int32_t n_moves, n_remove;
    int32_t move_track[MAX_MOVES];
    int32_t remove_track[MAX_MOVES];
    float t[MAX_MOVES][12];   // rows 0..2 of the 4x4
};

__device__ __forceinline__ int find_track(const int32_t* __restrict__ tab, int n, int tr) {
    for (int k = 0; k < n; ++k)
        if (tab[k] == tr) return k;
    return -1;
}

// pass 1a: this rank's number of listed particles per moved track, published to every rank
__global__ void __launch_bounds__(1024)
k_move_count(const MoveTable* __restrict__ mt, const uint16_t* __restrict__ movl_track, uint32_t movl_cap, const Persist* __restrict__ persist,
             CommView cv) {
    __shared__ uint32_t s_count[MAX_MOVES];
    const uint32_t n = min(persist->n_movable, movl_cap);
    const int n_moves = mt->n_moves;
    if (threadIdx.x < MAX_MOVES) s_count[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const int mv = find_track(mt->move_track, n_moves, (int)movl_track[k]);
        if (mv >= 0) atomicAdd(&s_count[mv], 1u);
    }
    __syncthreads();
    if (threadIdx.x < MAX_MOVES)
        for (int pr = 0; pr < cv.size; ++pr) cv.move_counts[pr][cv.rank * MAX_MOVES + threadIdx.x] = s_count[threadIdx.x];
    __threadfence_system();
}

// pass 1b (single block): sequence number of every listed particle of a moved track in the reference's processing
// order — object by object, ascending particle index inside an object (= ascending rank, then ascending local index) —
// = offset[move] + particles of that move on lower ranks + stable local rank.  Entries of a removed track are tagged
// SEQ_REMOVE, everything else NO_RANK.  Also advances the random cursor by the GLOBAL draw count (replicated state).
__global__ void __launch_bounds__(1024)
k_move_prepare(const MoveTable* __restrict__ mt, const uint16_t* __restrict__ movl_track, uint32_t* __restrict__ movl_seq,
               uint32_t movl_cap, uint32_t rand_n, Counters* __restrict__ cnt, Persist* __restrict__ persist, CommView cv,
               const uint32_t* __restrict__ move_counts) {
    __shared__ uint32_t s_offset[MAX_MOVES], s_running[MAX_MOVES];
    __shared__ uint32_t s_wcount[32][MAX_MOVES];
    __shared__ uint32_t s_local[MAX_MOVES];
    const uint32_t n = min(persist->n_movable, movl_cap);
    const int n_moves = mt->n_moves;
    if (cv.size == 1) {                        // unsharded: count here (sharded maps ran k_move_count and a barrier)
        if (threadIdx.x < MAX_MOVES) s_local[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
            const int mv = find_track(mt->move_track, n_moves, (int)movl_track[k]);
            if (mv >= 0) atomicAdd(&s_local[mv], 1u);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int m = 0; m < n_moves; ++m) {
            uint32_t below = 0, all = 0;
            for (int g = 0; g < cv.size; ++g) { const uint32_t c = cv.size == 1 ? s_local[m] : move_counts[g * MAX_MOVES + m]; if (g < cv.rank) below += c; all += c; }
            s_offset[m] = acc + below;
            s_running[m] = 0;
            acc += all;
        }
        cnt->n_selected = acc;
        cnt->move_cursor0 = persist->cursor_ring;
        persist->cursor_ring = (uint32_t)(((uint64_t)persist->cursor_ring + 3ull * acc) % rand_n);   // basic_algorithms.h:426-433
    }
    __syncthreads();
    const uint32_t wid = threadIdx.x >> 5;
    for (uint32_t c = 0; c < n; c += blockDim.x) {
        for (int i = threadIdx.x; i < 32 * n_moves; i += blockDim.x) s_wcount[i / n_moves][i % n_moves] = 0;
        __syncthreads();
        const uint32_t k = c + threadIdx.x;
        int mv = -1;
        bool rm = false;
        if (k < n) {
            const int tr = (int)movl_track[k];
            mv = find_track(mt->move_track, n_moves, tr);
            if (mv < 0) rm = find_track(mt->remove_track, mt->n_remove, tr) >= 0;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, mv);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane_id()) - 1u));
        if (mv >= 0 && rank_in_warp == 0) s_wcount[wid][mv] = __popc(peers);
        __syncthreads();
        if ((int)threadIdx.x < n_moves) {          // exclusive prefix over the warps of this chunk, per move
            uint32_t acc = s_running[threadIdx.x];
            for (int w = 0; w < 32; ++w) { const uint32_t t = s_wcount[w][threadIdx.x]; s_wcount[w][threadIdx.x] = acc; acc += t; }
            s_running[threadIdx.x] = acc;
        }
        __syncthreads();
        if (k < n) movl_seq[k] = mv >= 0 ? s_offset[mv] + s_wcount[wid][mv] + rank_in_warp : (rm ? SEQ_REMOVE : NO_RANK);
        __syncthreads();
    }
}

// ---- the same numbering for the usual frame (at most four moved objects), over the whole GPU -------------------------------
// The movable list of a large map is long (every live particle of every movable track: 3 x 10^5 entries at 512 x 128 x 512), and
// one block walking it twice was 66 us there.  Two launches of MPB_BLOCKS blocks, block b owning entries [b, b + 1) * MPB_CHUNK
// (thread order = list order, MPB_EPT consecutive entries per thread):
//   k_move_count_blocks : per block, the number of entries of each moved track (two tracks share one word of a block scan);
//   k_move_number       : every block adds up the counts of the blocks before it (its base per track) and all of them (the
//                         rank's totals; sharded maps: this rank's row of the published counts), numbers its own entries, and
//                         block 0 does the bookkeeping of k_move_prepare's thread 0.
constexpr int MPB_EPT = 4, MPB_CHUNK = SCAN_BLOCK * MPB_EPT;

__device__ __forceinline__ void move_classify(const MoveTable* __restrict__ mt, int n_moves, const uint16_t* __restrict__ movl_track, uint32_t n, uint32_t k0,
                                              int (&mvv)[MPB_EPT], bool (&rmv)[MPB_EPT], uint32_t& pk01, uint32_t& pk23) {
    pk01 = 0; pk23 = 0;
#pragma unroll
    for (int i = 0; i < MPB_EPT; ++i) {
        const uint32_t k = k0 + i;
        mvv[i] = -1; rmv[i] = false;
        if (k < n) {
            const int tr = (int)movl_track[k];
            mvv[i] = find_track(mt->move_track, n_moves, tr);
            if (mvv[i] < 0) rmv[i] = find_track(mt->remove_track, mt->n_remove, tr) >= 0;
        }
        pk01 += (mvv[i] == 0) + ((uint32_t)(mvv[i] == 1) << 16);
        pk23 += (mvv[i] == 2) + ((uint32_t)(mvv[i] == 3) << 16);
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_move_count_blocks(const MoveTable* __restrict__ mt, const uint16_t* __restrict__ movl_track, uint32_t movl_cap, const Persist* __restrict__ persist,
                    uint4* __restrict__ blk_counts) {
    __shared__ uint32_t s_scan[33];
    const uint32_t n = min(persist->n_movable, movl_cap);
    const uint32_t k0 = blockIdx.x * MPB_CHUNK;
    if (k0 >= n) { if (threadIdx.x == 0) blk_counts[blockIdx.x] = make_uint4(0, 0, 0, 0); return; }      // (block-uniform)
    int mvv[MPB_EPT];
    bool rmv[MPB_EPT];
    uint32_t pk01, pk23, tot01 = 0, tot23 = 0;
    move_classify(mt, mt->n_moves, movl_track, n, k0 + threadIdx.x * MPB_EPT, mvv, rmv, pk01, pk23);
    block_exclusive_scan(pk01, s_scan, tot01);
    block_exclusive_scan(pk23, s_scan, tot23);
    if (threadIdx.x == 0) blk_counts[blockIdx.x] = make_uint4(tot01 & 0xffffu, tot01 >> 16, tot23 & 0xffffu, tot23 >> 16);
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_move_number(const MoveTable* __restrict__ mt, const uint16_t* __restrict__ movl_track, uint32_t* __restrict__ movl_seq, uint32_t movl_cap, uint32_t rand_n,
              Counters* __restrict__ cnt, Persist* __restrict__ persist, CommView cv, const uint32_t* __restrict__ move_counts, const uint4* __restrict__ blk_counts) {
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_base[4], s_offset[4];
    const uint32_t n = min(persist->n_movable, movl_cap);
    const int n_moves = mt->n_moves;
    // counts of the blocks before this one, and of all blocks (gridDim.x <= SCAN_BLOCK: thread t looks at block t)
    const uint4 c4 = threadIdx.x < gridDim.x ? blk_counts[threadIdx.x] : make_uint4(0, 0, 0, 0);
    const uint32_t cs[4] = {c4.x, c4.y, c4.z, c4.w};
    uint32_t total[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint32_t ex = block_exclusive_scan(cs[m], s_scan, total[m]);
        if (threadIdx.x == blockIdx.x) s_base[m] = ex;
    }
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int m = 0; m < n_moves; ++m) {      // offset of a track = everything of the tracks before it + this track on the lower ranks
            uint32_t below = 0, all = 0;
            for (int g = 0; g < cv.size; ++g) { const uint32_t c = cv.size == 1 ? total[m] : move_counts[g * MAX_MOVES + m]; if (g < cv.rank) below += c; all += c; }
            s_offset[m] = acc + below;
            acc += all;
        }
        if (blockIdx.x == 0) {
            cnt->n_selected = acc;
            cnt->move_cursor0 = persist->cursor_ring;
            persist->cursor_ring = (uint32_t)(((uint64_t)persist->cursor_ring + 3ull * acc) % rand_n);   // basic_algorithms.h:426-433
        }
    }
    __syncthreads();
    const uint32_t k0 = blockIdx.x * MPB_CHUNK + threadIdx.x * MPB_EPT;
    if (blockIdx.x * MPB_CHUNK >= n) return;                                                             // (block-uniform)
    int mvv[MPB_EPT];
    bool rmv[MPB_EPT];
    uint32_t pk01, pk23, tot01 = 0, tot23 = 0;
    move_classify(mt, n_moves, movl_track, n, k0, mvv, rmv, pk01, pk23);
    const uint32_t ex01 = block_exclusive_scan(pk01, s_scan, tot01), ex23 = block_exclusive_scan(pk23, s_scan, tot23);
    uint32_t before[4] = {s_base[0] + (ex01 & 0xffffu), s_base[1] + (ex01 >> 16), s_base[2] + (ex23 & 0xffffu), s_base[3] + (ex23 >> 16)};
#pragma unroll
    for (int i = 0; i < MPB_EPT; ++i) {
        if (k0 + i >= n) continue;
        uint32_t seq = rmv[i] ? SEQ_REMOVE : NO_RANK;
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (mvv[i] == m) seq = s_offset[m] + before[m]++;
        movl_seq[k0 + i] = seq;
    }
}

// pass 2: transform + table noise, delete the original, hand the moved particle to the rank that owns its destination
// voxel (a record in that rank's arena); removed tracks' particles are invalidated here too
__global__ void __launch_bounds__(256)
k_move_apply(const __grid_constant__ FrameParams f, Counters* __restrict__ cnt, Store st, const MoveTable* __restrict__ mt,
             const float* __restrict__ randoms, const Persist* __restrict__ persist, const uint32_t* __restrict__ movl_idx,
             const uint16_t* __restrict__ movl_track, const uint32_t* __restrict__ movl_seq, uint32_t movl_cap, CommView cv) {
    uint32_t* out_count = cnt->mov_out;
    const uint32_t n = min(persist->n_movable, movl_cap);
    const uint32_t cur = cnt->move_cursor0;
    const uint32_t slabs = f.z_hi - f.z_lo;
    uint32_t n_removed = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t seq = movl_seq[k];
        if (seq == NO_RANK) continue;
        const uint32_t idx = movl_idx[k];
        const uint64_t m = st.meta[idx];
        st.meta[idx] = meta_with_status(m, ST_INVALID);                        // operations.h:346 / object_layer.h:418-421
        atomicAnd(st.hdr + (idx >> f.Pb), ~(1u << (16 + (idx & (f.P - 1)))));
        if (seq == SEQ_REMOVE) { ++n_removed; continue; }
        const float* t = mt->t[find_track(mt->move_track, mt->n_moves, (int)movl_track[k])];
        const float ox = st.x[idx], oy = st.y[idx], oz = st.z[idx];
        const float nx = ((t[0] * ox + t[1] * oy) + t[2] * oz) + t[3];       // operations.h:340
        const float ny = ((t[4] * ox + t[5] * oy) + t[6] * oz) + t[7];
        const float nz = ((t[8] * ox + t[9] * oy) + t[10] * oz) + t[11];
        // pre-incremented, wrapping cursor (utils/basic_algorithms.h:426-433): draw d uses (cursor + d) mod n
        const uint64_t d0 = (uint64_t)cur + 3ull * seq;
        const float px = nx + randoms[(d0 + 1) % f.rand_n];
        const float py = ny + randoms[(d0 + 2) % f.rand_n];
        const float pz = nz + randoms[(d0 + 3) % f.rand_n];
        uint32_t bx, by, bz;
        const uint32_t v = global_pos_to_voxel(f, px, py, pz, bx, by, bz);
        if (v == INVALID_INDEX) continue;                                      // left the map
        const uint32_t owner = bz / slabs;
        const uint32_t pos = atomicAdd(out_count + owner, 1u);
        if (pos < cv.mov_cap) {
            uint4* dst = reinterpret_cast<uint4*>(cv.movs[owner] + (size_t)cv.rank * cv.mov_cap + pos);
            dst[0] = make_uint4(__float_as_uint(px), __float_as_uint(py), __float_as_uint(pz), __float_as_uint(st.w[idx]));
            dst[1] = make_uint4((uint32_t)m, (uint32_t)(m >> 32), seq, v);
        }
    }
    if (n_removed) atomicAdd(&cnt->n_removed, n_removed);
    if (cv.size > 1) __threadfence_system();      // sharded map: peers read what this kernel wrote after the next barrier
}

// tell every destination rank how many moved particles this rank sent it
__global__ void k_move_publish(CommView cv, const Counters* __restrict__ cnt) {
    const int p = threadIdx.x;
    if (p < cv.size) cv.counts[p][cv.rank * COMM_COUNTS + CC_MOVREC] = cnt->mov_out[p];
    __threadfence_system();
}

// pass 3: bin the moved particles that arrived here by destination voxel, ranked by sequence number
__global__ void __launch_bounds__(256)
k_move_bin(const __grid_constant__ FrameParams f, CommView cv, const MovRec* __restrict__ movs, const uint32_t* __restrict__ counts,
           uint32_t* __restrict__ seq_ref, uint32_t seq_cap, uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks,
           uint32_t* __restrict__ entry_list, Counters* __restrict__ cnt) {
    for (int g = 0; g < cv.size; ++g) {
        const uint32_t n = g == cv.rank ? cnt->mov_out[cv.rank] : counts[g * COMM_COUNTS + CC_MOVREC];
        if (n > cv.mov_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&cnt->hash_overflow, 1u); }
        const MovRec* seg = movs + (size_t)g * cv.mov_cap;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < min(n, cv.mov_cap); k += gridDim.x * blockDim.x) {
            const uint32_t seq = seg[k].seq, v = seg[k].voxel;
            if (seq >= seq_cap) { atomicAdd(&cnt->hash_overflow, 1u); continue; }      // reported as a capacity error by the host
            seq_ref[seq] = (uint32_t)g * cv.mov_cap + k;
            topk_insert(f, f.hash_cap, f.hash_shift, v, seq, keys, ranks, entry_list, &cnt->n_move_voxels, &cnt->hash_overflow);
        }
    }
}

// fallback removal by a pass over the header plane (used when the movable list does not describe the store any more:
// a track that is moved AND removed in the same frame, which the reference removes after the move)
__global__ void __launch_bounds__(256)
k_remove_tracks_scan(const __grid_constant__ FrameParams f, Store st, const MoveTable* __restrict__ mt, Counters* __restrict__ cnt) {
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb), v_hi = f.z_hi << (f.Xb + f.Yb);
    uint32_t n_removed = 0;
    for (uint32_t v = v_lo + blockIdx.x * blockDim.x + threadIdx.x; v < v_hi; v += gridDim.x * blockDim.x) {
        const uint32_t hdr = st.hdr[v];
        if (!(hdr & HDR_MOVABLE)) continue;
        const size_t base = (size_t)v << f.Pb;
        uint32_t live = hdr_live(hdr), new_live = live;
        while (live) {
            const uint32_t s = __ffs(live) - 1;
            live &= live - 1;
            const uint64_t m = st.meta[base + s];
            if (find_track(mt->remove_track, mt->n_remove, (int)meta_track(m)) >= 0) {
                st.meta[base + s] = meta_with_status(m, ST_INVALID);
                new_live &= ~(1u << s);
                ++n_removed;
            }
        }
        if (new_live != hdr_live(hdr)) st.hdr[v] = (hdr & 0x1ffffu) | (new_live << 16);
    }
    if (n_removed) atomicAdd(&cnt->n_removed, n_removed);
}

// pass 4: re-insert in sequence order (operations.h:351-360); one thread per destination voxel decides the slots
template <int PB>
__global__ void __launch_bounds__(128)
k_move_replay(const __grid_constant__ FrameParams f, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
              const uint32_t* __restrict__ stamp_z, uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks,
              const uint32_t* __restrict__ entry_list, uint32_t* __restrict__ entry_voxel, Counters* __restrict__ cnt) {
    constexpr int P = 1 << PB, K = 2 * P - 1;
    const uint32_t n_entries = cnt->n_move_voxels;
    for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_entries; li += gridDim.x * blockDim.x) {
        const uint32_t e = entry_list[li];
        const uint32_t v = keys[e];
        keys[e] = EMPTY_KEY;
        entry_voxel[li] = v;
        uint32_t* r = ranks + (size_t)e * K;
        uint32_t rk[K];
#pragma unroll
        for (int t = 0; t < K; ++t) rk[t] = r[t];
        uint32_t bx, by, bz;
        voxel_to_buffer(f, v, bx, by, bz);
        const uint32_t stamp = slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz);
        const uint32_t hdr = st.hdr[v];
        VoxelSlots<P> vs;
        vs.load(st, (size_t)v << PB, hdr, stamp);
        uint32_t moved = 0;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const uint32_t seq = rk[t];
            if (seq == NO_RANK) continue;
            const uint32_t s = vs.first_vacant();
            if (s) { vs.occupy(s); ++moved; }
            r[t] = s ? (seq | (s << 24)) : NO_RANK;
        }
        if (moved) {
            st.hdr[v] = hdr_ts(hdr) | HDR_MOVABLE | (vs.live << 16);   // moved particles belong to a movable track by construction
            atomicAdd(&cnt->n_moved, moved);
        }
    }
}

__global__ void __launch_bounds__(256)
k_move_write(const __grid_constant__ FrameParams f, Store st, const MovRec* __restrict__ movs, const uint32_t* __restrict__ seq_ref, uint32_t seq_cap,
             uint32_t* __restrict__ ranks, const uint32_t* __restrict__ entry_list, const uint32_t* __restrict__ entry_voxel,
             const Counters* __restrict__ cnt) {
    const uint32_t K = f.topk, n = cnt->n_move_voxels * K;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t li = t / K, i = t - li * K;
        uint32_t* r = ranks + (size_t)entry_list[li] * K + i;
        const uint32_t a = *r;
        if (a == NO_RANK) continue;
        *r = NO_RANK;
        const uint32_t s = a >> 24, seq = a & 0xffffffu;
        if (seq >= seq_cap) continue;                                                  // cannot happen: k_move_bin drops such records
        const size_t idx = ((size_t)entry_voxel[li] << f.Pb) + s;
        const uint4* rec = reinterpret_cast<const uint4*>(movs + seq_ref[seq]);
        const uint4 lo = rec[0], hi = rec[1];
        st.x[idx] = __uint_as_float(lo.x); st.y[idx] = __uint_as_float(lo.y); st.z[idx] = __uint_as_float(lo.z); st.w[idx] = __uint_as_float(lo.w);
        st.meta[idx] = (uint64_t)hi.x | ((uint64_t)hi.y << 32);
    }
}

// =====================================================================================================
// store initialisation (mc_ring/operations.h:684-723): everything zero, slot 0 of each voxel = TIMEPTC
// =====================================================================================================
__global__ void __launch_bounds__(256)
k_clear_store(Store st, size_t n_slots, uint32_t slot_bits) {
    const uint32_t slot_mask = (1u << slot_bits) - 1u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x) {
        st.x[i] = 0.f; st.y[i] = 0.f; st.z[i] = 0.f; st.w[i] = 0.f;
        if ((i & slot_mask) == 0) { st.meta[i] = meta_pack(0, 0, 0, ST_TIMEPTC, 0); st.hdr[i >> slot_bits] = 0u; }
        else st.meta[i] = 0ull;
    }
}

__global__ void k_fill_u32(uint32_t* p, size_t n, uint32_t value) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = value;
}

// live particles of one track (device-side ObjectParticleHashMap::indices_map[track].size(), object_layer.h:20-52)
__global__ void __launch_bounds__(256)
k_count_track(Store st, uint32_t n_voxels, uint32_t slot_bits, uint32_t track, uint32_t* __restrict__ out) {
    uint32_t c = 0;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_voxels; v += gridDim.x * blockDim.x) {
        uint32_t live = hdr_live(st.hdr[v]);
        while (live) {
            const uint32_t s = __ffs(live) - 1;
            live &= live - 1;
            if (meta_track(st.meta[((size_t)v << slot_bits) + s]) == track) ++c;
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane_id() == 0 && c) atomicAdd(out, c);
}

// DSPMAP_DUMP_PARTICLES_HASH (include/dspmap_c.h): order-independent digest of the reassembled 24-byte records
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256)
k_store_hash(Store st, size_t slot_lo, size_t n_slots, uint32_t slot_bits, unsigned long long* __restrict__ out) {
    const size_t slot_mask = ((size_t)1 << slot_bits) - 1;
    uint64_t local = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_slots; k += (size_t)gridDim.x * blockDim.x) {
        const size_t i = slot_lo + k;
        uint64_t m = st.meta[i];
        if ((i & slot_mask) == 0) m = (m & ~0xffffull) | hdr_ts(st.hdr[i >> slot_bits]);   // slot 0 = the time particle: its stamp lives in the header
        const uint64_t a = (uint64_t)__float_as_uint(st.x[i]) | ((uint64_t)__float_as_uint(st.y[i]) << 32);
        const uint64_t b = (uint64_t)__float_as_uint(st.z[i]) | ((uint64_t)__float_as_uint(st.w[i]) << 32);
        local += mix64(mix64(mix64((uint64_t)i * 0x9E3779B97F4A7C15ull + a) ^ b) + m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane_id() == 0 && local) atomicAdd(out, (unsigned long long)local);
}

// brute-force check of InvDiv against the compiler's IEEE division (test hook)
__global__ void k_selftest_invdiv(uint64_t n, uint64_t seed, float b_lo, float b_hi, float a_max, unsigned long long* __restrict__ mismatches) {
    uint64_t local = 0;
    const bool wide = a_max < 0.f;      // a_max < 0: test div_by_stored (log-uniform divisor in [b_lo, b_hi], dividend 2^-100 .. 2^100)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + i * 0x9E3779B97F4A7C15ull;            // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        const float ub = (float)(uint32_t)(z & 0xffffffu) * (1.0f / 16777216.0f);
        const float b = b_lo + ub * (b_hi - b_lo);
        // dividend: random sign, random mantissa, magnitude spread over many binades up to a_max
        const uint32_t bits = (uint32_t)(z >> 24);
        const float mant = 1.0f + (float)(bits & 0x7fffffu) * (1.0f / 8388608.0f);
        const int ex = (int)((bits >> 23) & 31u) - 24;            // 2^-24 .. 2^7
        float a = ldexpf(mant, ex) * ((bits >> 28) & 1u ? -1.f : 1.f);
        if (fabsf(a) > a_max) a = a_max * 0.73f;
        float q1;
        const float q2 = a / b;
        if (wide) {                                        // div_by_stored over its whole claimed range
            const float lb = log2f(b_lo) + ub * (log2f(b_hi) - log2f(b_lo));
            const float bw = exp2f(lb);
            const int exw = (int)((bits >> 23) & 255u) % 201 - 100;    // 2^-100 .. 2^100
            const float aw = ldexpf(mant, exw) * ((bits >> 31) ? -1.f : 1.f);
            const float2 cy = ck_and_reciprocal(bw);
            bool exact = true;
            q1 = div_by_stored(aw, bw, cy.y, exact);
            const float q2w = aw / bw;
            if (exact && __float_as_uint(q1) != __float_as_uint(q2w)) ++local;
            continue;
        }
        const InvDiv d(b);
        q1 = d.ok ? d.div<true>(a) : d.div<false>(a);
        if (__float_as_uint(q1) != __float_as_uint(q2)) ++local;
    }
    if (local) atomicAdd(mismatches, (unsigned long long)local);
}

// all 2^24 8-bit triples through both colour-space conversions; out = 2 x 2^24 packed (c0 | c1 << 8 | c2 << 16) words
__global__ void k_selftest_colour(uint32_t* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 24); i += gridDim.x * blockDim.x) {
        const int a = (int)(i >> 16), b = (int)((i >> 8) & 255u), c = (int)(i & 255u);
        int h, s, v, r, g, bb;
        rgb_to_hsv_u8(a, b, c, h, s, v);
        out[i] = (uint32_t)(h & 255) | ((uint32_t)s << 8) | ((uint32_t)v << 16);
        hsv_to_rgb_u8(a, b, c, r, g, bb);
        out[(1u << 24) + i] = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)bb << 16);
    }
}

}  // namespace dspk
