// dspmap_kernels.cuh — the per-frame kernels of the particle map engine (sm_100a).
//
// One frame = K1 project -> (P2 object moves) -> U2 vertex mask + sweep + pixel binning -> U3 ck / weight update
//             -> B1 birth + resample -> O1 occupancy + label fusion.
// Each kernel header names the reference routine it replaces (paths relative to the reference root).
// Iteration order is the oracle's "canonical" order (ascending particle index inside a pixel list, ascending
// reference rank for births), evaluated sequentially per output element so float sums are bit-identical.
#pragma once
#include "dspmap_device.cuh"

namespace dspk {

// device-resident per-frame counters (zeroed at frame start)
struct Counters {
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
    uint32_t ticket_occ;       // dynamic block ids for the ordered compactions
    uint32_t ticket_sel;
    uint32_t hash_overflow;    // insertion hash table full (cannot happen with cap >= 2 * candidates)
    uint32_t scratch;
    uint32_t pad[2];
};

// survives across frames
struct Persist {
    uint32_t cursor_ring;      // RingBufferOperations::gaussian_random_calculator_ cursor (operations.h:1460)
    uint32_t cursor_map;       // SemanticDSPMap::gaussian_random_ cursor (semantic_dsp_map.h:265)
};

struct Store {
    float* x; float* y; float* z; float* w;
    uint64_t* meta;
    uint32_t* hdr;             // per voxel, see dspmap_device.cuh
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// =====================================================================================================
// K1  project_label   replaces PointCloudTools::manualResize + generateLabeledPointCloud
//                     (utils/pointcloud_tools.h:1104-1133, 88-310)
// one thread per internal pixel; double-precision back-projection, float narrowing at the end.
// =====================================================================================================
__global__ void __launch_bounds__(256)
k_project(int slot, const float* __restrict__ depth, int depth_stride, const MaskDesc* __restrict__ masks,
          const uint16_t* __restrict__ lut_static_track, const uint8_t* __restrict__ lut_static_label,
          float* __restrict__ depth_int, float4* __restrict__ cloud_pos, uint32_t* __restrict__ cloud_meta,
          Counters* __restrict__ cnt) {
    const FrameParams& f = c_fp[slot];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    if (q < f.W * f.H) {
        const int i = q / f.W, j = q - i * f.W;
        int si = i, sj = j;
        if (f.boost && !f.pre_decimated) {  // :1104-1133, src = (int)(i * (1/0.5f)) clamped
            si = min((int)((float)i * 2.0f), f.in_H - 1);
            sj = min((int)((float)j * 2.0f), f.in_W - 1);
        }
        const float d = depth[(size_t)si * depth_stride + sj];
        depth_int[q] = d;

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
            if (background) { label = 0; track = 65535u; }
            else if (inst > f.max_movable) label = lut_static_label[inst];                           // :277-279
            else label = owner >= 0 ? ((uint32_t)masks[owner].label_id & 0xffu) : 0u;                 // :281
            pos = make_float4((float)pg[0], (float)pg[1], (float)pg[2], sigma);                       // :298-300
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
// one warp per 32 consecutive x vertices; rows are (z, y) in MAP coordinates.
// =====================================================================================================
__global__ void __launch_bounds__(256)
k_vertex_mask(int slot, uint32_t* __restrict__ vmask, uint32_t words_per_row, uint32_t n_words) {
    const FrameParams& f = c_fp[slot];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_words) return;
    const uint32_t row = w / words_per_row, xw = w - row * words_per_row;
    const uint32_t z = row / (f.NY + 1), y = row - z * (f.NY + 1);
    const uint32_t x = xw * 32 + lane_id();
    bool in = false;
    if (x <= f.NX)
        in = point_in_frustum(f, (float)(int)x * f.voxel + f.vox_off[0], (float)(int)y * f.voxel + f.vox_off[1], (float)(int)z * f.voxel + f.vox_off[2]);
    const uint32_t bits = __ballot_sync(0xffffffffu, in);
    if (lane_id() == 0) vmask[w] = bits;
}

__device__ __forceinline__ bool vbit(const uint32_t* __restrict__ vmask, uint32_t wpr, uint32_t NYp1, uint32_t x, uint32_t y, uint32_t z) {
    return (__ldg(vmask + (size_t)(z * NYp1 + y) * wpr + (x >> 5)) >> (x & 31u)) & 1u;
}

// =====================================================================================================
// U2b  sweep          replaces getIdxOfVisibleParitlces minus its queue (mc_ring/operations.h:1344-1435)
// one thread per storage voxel of this shard.  A voxel is swept when one of its 8 corner vertices is in the
// frustum.  The header word says which slots are live, so an empty voxel costs 4 bytes.  Swept particles are:
// invalidated when stale, reset to the newborn weight under a too-far depth pixel, skipped when occluded,
// otherwise appended as a (pixel, particle) pair and counted per pixel.
// =====================================================================================================
__global__ void __launch_bounds__(256)
k_sweep(int slot, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
        const uint32_t* __restrict__ stamp_z, const uint32_t* __restrict__ vmask, uint32_t words_per_row,
        const float* __restrict__ depth_int, uint32_t* __restrict__ pix_count, uint2* __restrict__ pairs,
        Counters* __restrict__ cnt) {
    const FrameParams& f = c_fp[slot];
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb), v_hi = f.z_hi << (f.Xb + f.Yb);
    const uint32_t v = v_lo + blockIdx.x * blockDim.x + threadIdx.x;
    bool swept = false;
    uint32_t bx = 0, by = 0, bz = 0, mx = 0, my = 0, mz = 0;
    if (v < v_hi) {
        voxel_to_buffer(f, v, bx, by, bz);
        buffer_to_map(f, bx, by, bz, mx, my, mz);
        const uint32_t nyp = f.NY + 1;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a) swept |= vbit(vmask, words_per_row, nyp, mx + a, my + b, mz + c);
    }
    if (!__any_sync(0xffffffffu, swept)) return;

    const size_t base = (size_t)v << f.Pb;
    const uint32_t hdr = swept ? st.hdr[v] : 0u;
    uint32_t todo = hdr_live(hdr), live = todo;
    const uint32_t stamp = todo ? slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz) : 0u;
    int nvalid = 0;
    bool observed = false;
    while (__any_sync(0xffffffffu, todo != 0)) {
        bool emit = false;
        uint32_t pix = 0, s = 0;
        if (todo) {
            s = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint64_t m = st.meta[base + s];
            if (meta_ts(m) < stamp) {
                st.meta[base + s] = meta_with_status(m, ST_INVALID);  // :1374-1378 outdated by the ring shift
                live &= ~(1u << s);
            } else {
                ++nvalid;
                int row, col;
                float cz;
                if (project_to_image(f, st.x[base + s], st.y[base + s], st.z[base + s], row, col, cz)) {
                    const float d = __ldg(depth_int + row * f.W + col);
                    if (d > f.depth_max) {             // :1389-1395 free space seen through this pixel
                        st.w[base + s] = OCC_INIT_WEIGHT;
                        observed = true;
                    } else if (cz > d * ONE_SIGMA_COEFF) {  // :1397-1400 occluded
                    } else {
                        observed = true;
                        emit = true;
                        pix = (uint32_t)(row * f.W + col);
                    }
                }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, emit);
        if (bal) {
            uint32_t off = 0;
            const uint32_t leader = __ffs(bal) - 1;
            if (lane_id() == leader) off = atomicAdd(&cnt->n_pairs, (uint32_t)__popc(bal));
            off = __shfl_sync(0xffffffffu, off, leader);
            if (emit) {
                const uint32_t pos = off + __popc(bal & ((1u << lane_id()) - 1u));
                if (pos < f.vis_cap) {
                    pairs[pos] = make_uint2(pix, (uint32_t)(base + s));
                    atomicAdd(pix_count + pix, 1u);
                }
            }
        }
    }
    if (swept) {
        bool stamp0 = observed;                                    // :1415-1417
        if (!observed && nvalid == 0) {                            // :1418-1431 imaginary particle at the voxel's min corner
            const float gx = (float)(int)mx * f.voxel + f.pmin[0] + f.map_center[0];
            const float gy = (float)(int)my * f.voxel + f.pmin[1] + f.map_center[1];
            const float gz = (float)(int)mz * f.voxel + f.pmin[2] + f.map_center[2];
            int row, col;
            float cz;
            if (project_to_image(f, gx, gy, gz, row, col, cz)) stamp0 = cz <= __ldg(depth_int + row * f.W + col);
        }
        const uint32_t nh = (stamp0 ? (f.T & 0xffffu) : hdr_ts(hdr)) | (hdr & HDR_MOVABLE) | (live << 16);
        if (nh != hdr) st.hdr[v] = nh;
    }
    const uint32_t b_swept = __ballot_sync(0xffffffffu, swept), b_nonempty = __ballot_sync(0xffffffffu, swept && nvalid > 0);
    if (lane_id() == 0) {
        atomicAdd(&cnt->n_swept_voxels, (uint32_t)__popc(b_swept));
        if (b_nonempty) atomicAdd(&cnt->n_swept_nonempty, (uint32_t)__popc(b_nonempty));
    }
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
k_scan_apply(const uint32_t* __restrict__ count, uint32_t n, const uint32_t* __restrict__ block_sum, uint32_t* __restrict__ offset) {
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
    if (i == n - 1) offset[n] = base + ex + c;
}

// scatter the pairs into their pixel segments (arrival order; sorted next)
__global__ void __launch_bounds__(256)
k_fill(int slot, const Counters* __restrict__ cnt, const uint2* __restrict__ pairs, const uint32_t* __restrict__ offset,
       uint32_t* __restrict__ pix_count, uint32_t* __restrict__ csr_idx, uint32_t* __restrict__ vis_pix) {
    const uint32_t n = min(cnt->n_pairs, c_fp[slot].vis_cap);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[k];
        const uint32_t pos = offset[pr.x] + (atomicSub(pix_count + pr.x, 1u) - 1u);   // leaves the counts at zero
        csr_idx[pos] = pr.y;
        vis_pix[pos] = pr.x;
    }
}

// canonical order inside each pixel list: ascending particle index (insertion sort; lists are short)
__global__ void __launch_bounds__(256)
k_sort_lists(int slot, const uint32_t* __restrict__ offset, uint32_t* __restrict__ csr_idx) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= c_fp[slot].W * c_fp[slot].H) return;
    const uint32_t b = offset[q], e = offset[q + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
        const uint32_t key = csr_idx[i];
        uint32_t j = i;
        while (j > b && csr_idx[j - 1] > key) { csr_idx[j] = csr_idx[j - 1]; --j; }
        csr_idx[j] = key;
    }
}

// gather the listed particles into compact, list-ordered arrays read by the two stencil passes
__global__ void __launch_bounds__(256)
k_gather_visible(int slot, const Counters* __restrict__ cnt, Store st, const uint32_t* __restrict__ csr_idx,
                 float4* __restrict__ vis_pos, uint32_t* __restrict__ vis_meta) {
    const uint32_t n = min(cnt->n_pairs, c_fp[slot].vis_cap);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t idx = csr_idx[k];
        const uint64_t m = st.meta[idx];
        vis_pos[k] = make_float4(st.x[idx], st.y[idx], st.z[idx], st.w[idx]);
        vis_meta[k] = meta_track(m) | (meta_forget(m) << 16);
    }
}

// =====================================================================================================
// U3a  ck_accumulate  replaces SemanticDSPMap::updateParticles pass 1 (semantic_dsp_map.h:973-1037)
// one thread per observation pixel; window rows are contiguous CSR ranges; Gaussian LUT in shared memory.
// =====================================================================================================
__device__ __forceinline__ void stage_lut(float* s_lut, const float* __restrict__ lut) {
    for (int i = threadIdx.x; i < PDF_NUM / 4; i += blockDim.x) reinterpret_cast<float4*>(s_lut)[i] = reinterpret_cast<const float4*>(lut)[i];
    __syncthreads();
}

__global__ void __launch_bounds__(512)
k_ck(int slot, const float* __restrict__ lut, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta,
     const uint32_t* __restrict__ offset, const float4* __restrict__ vis_pos, const uint32_t* __restrict__ vis_meta,
     float* __restrict__ ck) {
    extern __shared__ float s_lut[];
    stage_lut(s_lut, lut);
    const FrameParams& f = c_fp[slot];
    const int npix = f.W * f.H, r = f.win_r;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < npix; q += gridDim.x * blockDim.x) {
        const uint32_t om = cloud_meta[q];
        if (!pt_valid(om)) { ck[q] = 0.f; continue; }
        const float4 o = cloud_pos[q];
        const uint32_t otrack = pt_track(om);
        const int i = q / f.W, j = q - i * f.W;
        const int j0 = max(j - r, 0), j1 = min(j + r, f.W - 1);
        float acc = 0.f;
        for (int ni = max(i - r, 0); ni <= min(i + r, f.H - 1); ++ni) {
            const uint32_t kb = offset[ni * f.W + j0], ke = offset[ni * f.W + j1 + 1];
            for (uint32_t k = kb; k < ke; ++k) {
                const uint32_t pm = vis_meta[k];
                const bool same = (pm & 0xffffu) == otrack;
                if (f.indep && !same) continue;
                const float4 p = vis_pos[k];
                float gk = query_pdf(s_lut, p.x, o.x, o.w) * query_pdf(s_lut, p.y, o.y, o.w) * query_pdf(s_lut, p.z, o.z, o.w);
                if (!f.indep) {
                    gk *= forgetting_factor(f, pm >> 16);
                    if (!same) gk *= f.tau;
                }
                acc += p.w * gk;
            }
        }
        ck[q] = acc * f.Pd + f.kappa;
    }
}

// =====================================================================================================
// U3b  weight_update  replaces SemanticDSPMap::updateParticles pass 2 (semantic_dsp_map.h:1041-1119)
// one thread per listed particle; sigma is the PARTICLE's pixel's sigma (:1047).
// =====================================================================================================
__global__ void __launch_bounds__(512)
k_weight_update(int slot, const Counters* __restrict__ cnt, const float* __restrict__ lut, const float4* __restrict__ cloud_pos,
                const uint32_t* __restrict__ cloud_meta, const float* __restrict__ ck, const uint32_t* __restrict__ csr_idx,
                const uint32_t* __restrict__ vis_pix, const float4* __restrict__ vis_pos, const uint32_t* __restrict__ vis_meta,
                Store st) {
    extern __shared__ float s_lut[];
    stage_lut(s_lut, lut);
    const FrameParams& f = c_fp[slot];
    const uint32_t n = min(cnt->n_pairs, f.vis_cap);
    const int r = f.win_r;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int q = (int)vis_pix[k];
        const int i = q / f.W, j = q - i * f.W;
        const float sigma = cloud_pos[q].w;
        const float4 p = vis_pos[k];
        const uint32_t pm = vis_meta[k];
        const uint32_t ptrack = pm & 0xffffu, pforget = pm >> 16;
        const float phi = f.indep ? 1.f : forgetting_factor(f, pforget);
        float acc = 0.f;
        bool right_id = false;
        for (int ni = max(i - r, 0); ni <= min(i + r, f.H - 1); ++ni) {
            for (int nj = max(j - r, 0); nj <= min(j + r, f.W - 1); ++nj) {
                const int oq = ni * f.W + nj;
                const uint32_t om = cloud_meta[oq];
                if (!pt_valid(om)) continue;
                const bool same = pt_track(om) == ptrack;
                if (f.indep && !same) continue;
                const float4 o = cloud_pos[oq];
                float gk = query_pdf(s_lut, p.x, o.x, sigma) * query_pdf(s_lut, p.y, o.y, sigma) * query_pdf(s_lut, p.z, o.z, sigma);
                if (!f.indep) {
                    if (!same) gk *= f.tau;
                    else if (gk > MIN_RIGHTLY_UPDATED_PDF) right_id = true;
                    gk *= phi;
                }
                acc += gk / ck[oq];
            }
        }
        const uint32_t idx = csr_idx[k];
        const uint64_t m = st.meta[idx];
        uint32_t forget = meta_forget(m);
        if (!f.indep) {
            if (right_id) forget = 0;
            else if (forget < 5) forget += 1;
        }
        st.w[idx] = p.w * (acc * f.Pd + 1.f - f.Pd);
        st.meta[idx] = meta_pack(f.T, meta_track(m), meta_label(m), ST_UPDATED, forget);
    }
}

// =====================================================================================================
// ordered insertion of candidates into voxels.
// The reference inserts points one at a time into "the first vacant slot" of their voxel
// (mc_ring/operations.h:782-803), so the outcome for a voxel depends only on the ORDER of the candidates that
// fall into it.  A voxel can absorb at most topk = 2P-1 candidates before every later one is a no-op (P-1
// slots, one resample that frees at most P-1 more), so: hash the voxel, keep the topk smallest ranks per
// voxel with a cascade of atomicMin (order-independent result), then replay each touched voxel serially.
// Touched hash entries are appended to a list so the replay never scans the table.
// =====================================================================================================
__device__ __forceinline__ uint32_t hash_slot(const FrameParams& f, uint32_t v) { return (v * 2654435761u) >> f.hash_shift; }

__device__ __forceinline__ void topk_insert(const FrameParams& f, uint32_t v, uint32_t rank, uint32_t* __restrict__ keys,
                                            uint32_t* __restrict__ ranks, uint32_t* __restrict__ entry_list, uint32_t* n_entries,
                                            uint32_t* overflow) {
    uint32_t h = hash_slot(f, v);
    for (uint32_t probe = 0;; ++probe) {
        if (probe >= f.hash_cap) { atomicAdd(overflow, 1u); return; }
        const uint32_t old = atomicCAS(keys + h, EMPTY_KEY, v);
        if (old == EMPTY_KEY) { entry_list[atomicAdd(n_entries, 1u)] = h; break; }
        if (old == v) break;
        h = (h + 1) & (f.hash_cap - 1);
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

__global__ void __launch_bounds__(256)
k_birth_bin(int slot, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta, uint32_t* __restrict__ keys,
            uint32_t* __restrict__ ranks, uint32_t* __restrict__ entry_list, Counters* __restrict__ cnt) {
    const FrameParams& f = c_fp[slot];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= f.W * f.H) return;
    if (!pt_valid(cloud_meta[q])) return;
    const float4 p = cloud_pos[q];
    uint32_t bx, by, bz;
    const uint32_t v = global_pos_to_voxel(f, p.x, p.y, p.z, bx, by, bz);
    if (v == INVALID_INDEX) return;
    if (bz < f.z_lo || bz >= f.z_hi) return;   // another shard owns this voxel
    const int i = q / f.W, j = q - i * f.W;
    topk_insert(f, v, birth_rank(f, i, j), keys, ranks, entry_list, &cnt->n_birth_voxels, &cnt->hash_overflow);
}

// first vacant slot 1..P-1 of a voxel (operations.h:790-796); 0 when full.  live = header mask (bit s)
__device__ __forceinline__ uint32_t first_vacant(const Store& st, size_t base, uint32_t P, uint32_t live, uint32_t stamp) {
    for (uint32_t s = 1; s < P; ++s) {
        if (!((live >> s) & 1u)) return s;
        if (meta_ts(st.meta[base + s]) < stamp) return s;
    }
    return 0;
}

// SemanticDSPMap::resampleParticlesInVoxel (semantic_dsp_map.h:1448-1519); returns true when a resample ran
__device__ __forceinline__ bool resample_voxel(const Store& st, size_t base, uint32_t P, uint32_t& live) {
    float wsum = 0.f;
    uint32_t n_upd = 0, upd = 0;
    for (uint32_t s = 1; s < P; ++s)
        if (((live >> s) & 1u) && meta_status(st.meta[base + s]) == ST_UPDATED) { wsum += st.w[base + s]; ++n_upd; upd |= 1u << s; }
    const uint32_t trigger = P >> 1;
    if (n_upd <= trigger) return false;
    if (wsum < 0.01f) {
        for (uint32_t s = 1; s < P; ++s)
            if ((upd >> s) & 1u) st.meta[base + s] = meta_with_status(st.meta[base + s], ST_INVALID);
        live &= ~upd;
        return true;
    }
    float wpp = wsum / (float)trigger;
    if (wpp > 1.f) wpp = 1.f;
    float cum = 0.f, thr = wpp;
    for (uint32_t s = 1; s < P; ++s) {
        if (!((upd >> s) & 1u)) continue;
        cum += st.w[base + s];
        if (cum < thr) {
            st.meta[base + s] = meta_with_status(st.meta[base + s], ST_INVALID);
            live &= ~(1u << s);
        } else {
            st.w[base + s] = wpp;
            thr += wpp;
            while (cum > thr) thr += wpp;
        }
    }
    return true;
}

// =====================================================================================================
// B1  birth_replay    replaces the birth loop + addNewbornParticle[WithNoise]AndResample
//                     (semantic_dsp_map.h:768-800, 1148-1230) — one thread per touched voxel, candidates in
//                     reference order.  Resets the hash entry it consumed.
// =====================================================================================================
__global__ void __launch_bounds__(128)
k_birth_replay(int slot, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
               const uint32_t* __restrict__ stamp_z, const float4* __restrict__ cloud_pos, const uint32_t* __restrict__ cloud_meta,
               uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks, const uint32_t* __restrict__ entry_list,
               Counters* __restrict__ cnt) {
    const FrameParams& f = c_fp[slot];
    const uint32_t n_entries = cnt->n_birth_voxels;
    for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_entries; li += gridDim.x * blockDim.x) {
        const uint32_t e = entry_list[li];
        const uint32_t v = keys[e];
        keys[e] = EMPTY_KEY;
        uint32_t bx, by, bz;
        voxel_to_buffer(f, v, bx, by, bz);
        const uint32_t stamp = slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz);
        const size_t base = (size_t)v << f.Pb;
        const uint32_t hdr = st.hdr[v];
        uint32_t live = hdr_live(hdr), movable = hdr & HDR_MOVABLE;
        uint32_t* r = ranks + (size_t)e * f.topk;
        bool resampled = false, frozen = false;
        uint32_t born = 0, n_res = 0;
        for (uint32_t t = 0; t < f.topk; ++t) {
            const uint32_t rank = r[t];
            if (rank == NO_RANK) break;
            r[t] = NO_RANK;
            if (frozen) continue;
            const int q = birth_rank_to_pixel(f, rank);
            uint32_t s = first_vacant(st, base, f.P, live, stamp);
            bool try_resample;
            if (f.depth_noise) try_resample = (s == 0) && !resampled;      // :1205-1226 only when the voxel is full
            else try_resample = !resampled;                                 // :1165-1170 after every add
            if (f.depth_noise && try_resample) {
                if (resample_voxel(st, base, f.P, live)) { resampled = true; ++n_res; s = first_vacant(st, base, f.P, live, stamp); }
                else frozen = true;            // full, and a resample can never trigger this frame
            }
            if (s) {
                const float4 p = cloud_pos[q];
                const uint32_t pm = cloud_meta[q];
                st.x[base + s] = p.x; st.y[base + s] = p.y; st.z[base + s] = p.z;
                st.w[base + s] = OCC_INIT_WEIGHT;
                st.meta[base + s] = meta_pack(f.T, pt_track(pm), pt_label(pm), ST_REGULAR_BORN, 0);   // operations.h:171-184
                live |= 1u << s;
                if (pt_track(pm) <= f.max_movable) movable = HDR_MOVABLE;
                ++born;
            } else if (f.depth_noise) {
                frozen = true;                 // full after the one allowed resample
            }
            if (!f.depth_noise && try_resample) {
                if (resample_voxel(st, base, f.P, live)) { resampled = true; ++n_res; }
                else if (s == 0) frozen = true;  // full and births never create UPDATED particles
            } else if (!f.depth_noise && s == 0) {
                frozen = true;
            }
        }
        const uint32_t nh = hdr_ts(hdr) | movable | (live << 16);
        if (nh != hdr) st.hdr[v] = nh;
        if (born) atomicAdd(&cnt->n_born, born);
        if (n_res) atomicAdd(&cnt->n_resampled, n_res);
    }
}

// =====================================================================================================
// ordered stream compaction across blocks: decoupled look-back over (flag | count) words, one warp looking
// at 32 predecessors at a time.  Blocks take a dynamic ticket so a block never waits on one that has not started.
// =====================================================================================================
constexpr uint64_t LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_MASK = (1ull << 62) - 1;

// call from every lane of warp 0; returns this block's exclusive prefix (two packed 31-bit counters) to all lanes
__device__ __forceinline__ uint64_t lookback_prefix_warp(volatile uint64_t* status, uint32_t bid, uint64_t block_total) {
    const uint32_t lane = lane_id();
    if (lane == 0) status[bid] = (bid == 0 ? LB_INC : LB_AGG) | block_total;
    if (bid == 0) return 0;
    uint64_t prefix = 0;
    for (int j = (int)bid - 1;; j -= 32) {
        const int k = j - (int)lane;
        uint64_t s = LB_INC;   // before the first block: an inclusive zero
        if (k >= 0) {
            do { s = status[k]; } while ((s >> 62) == 0);
        }
        const uint32_t inc_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        const uint32_t first_inc = inc_mask ? (uint32_t)(__ffs(inc_mask) - 1) : 32u;
        uint64_t val = (lane <= first_inc) ? (s & LB_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
        prefix += __shfl_sync(0xffffffffu, val, 0);
        if (inc_mask) break;
    }
    if (lane == 0) status[bid] = LB_INC | (prefix + block_total);
    return prefix;
}

// =====================================================================================================
// O1  occupancy_labels  replaces getOccupancyResult + determineIfVoxelOccupied + calculateWeightAndSemanticsInVoxel
//                       + getVoxelGlobalPosition (semantic_dsp_map.h:1239-1383; mc_ring/operations.h:390-448, 623-648)
// one thread per OCC_VPT consecutive storage voxels; header word first, slot records only for live slots of known
// voxels; sequential slot loop (the reference's summation order); ordered compaction of the emitted voxels.
// =====================================================================================================
constexpr int OCC_BLOCK = 256;
constexpr int OCC_VPT = 4;                       // voxels per thread (one 16-byte header load)
constexpr int OCC_TILE = OCC_BLOCK * OCC_VPT;
constexpr int MAX_P = 16;

struct ResultArrays {
    uint32_t* voxel_index; float* xyz; float* weight_sum; uint16_t* track_id; uint8_t* label_id; int8_t* state;
    uint32_t* free_voxel_index; float* free_xyz;
};

__global__ void __launch_bounds__(OCC_BLOCK)
k_occupancy(int slot, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
            const uint32_t* __restrict__ stamp_z, ResultArrays out, int8_t* __restrict__ voxel_state,
            uint64_t* __restrict__ lb_status, Counters* __restrict__ cnt) {
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_warp[33];
    __shared__ uint64_t s_prefix;
    const FrameParams& f = c_fp[slot];
    if (threadIdx.x == 0) s_bid = atomicAdd(&cnt->ticket_occ, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t v_lo = f.z_lo << (f.Xb + f.Yb);
    const uint32_t v0 = v_lo + bid * OCC_TILE + threadIdx.x * OCC_VPT;   // the grid covers the shard exactly (V multiple of 1024)

    const uint4 h4 = *reinterpret_cast<const uint4*>(st.hdr + v0);
    const uint32_t hdrs[OCC_VPT] = {h4.x, h4.y, h4.z, h4.w};
    uint32_t bx0, by, bz;
    voxel_to_buffer(f, v0, bx0, by, bz);
    const uint32_t stamp_yz = max(__ldg(stamp_y + by), __ldg(stamp_z + bz));

    int state[OCC_VPT];
    float wsum[OCC_VPT];
    uint32_t tl[OCC_VPT];     // track | label << 16
    uint32_t n_occ = 0, n_free = 0, n_nonempty = 0;
#pragma unroll
    for (int u = 0; u < OCC_VPT; ++u) {
        const uint32_t hdr = hdrs[u], v = v0 + u;
        const uint32_t stamp = max(stamp_yz, __ldg(stamp_x + bx0 + u));
        const uint32_t ts0 = hdr_ts(hdr);
        state[u] = -1; wsum[u] = -1.f; tl[u] = 0;
        if (ts0 != 0 && !(ts0 < stamp)) {                   // isVoxelValid, operations.h:824-837
            float ws = 0.f, gsum = 0.f;
            uint32_t live = hdr_live(hdr), new_live = live;
            uint32_t track = 0, label = 0;
            if (live) {
                const size_t base = (size_t)v << f.Pb;
                // per-track weight histogram (std::map in the reference), at most P-1 distinct tracks
                uint32_t h_track[MAX_P - 1], h_label[MAX_P - 1];
                float h_w[MAX_P - 1];
                int nh = 0;
                bool any = false;
                while (live) {
                    const uint32_t s = __ffs(live) - 1;
                    live &= live - 1;
                    const uint64_t m = st.meta[base + s];
                    if (meta_ts(m) < stamp) continue;           // vacant: outdated by the ring shift (operations.h:810-816)
                    any = true;
                    float w = st.w[base + s];
                    ws += w;                                    // the unclamped weight, operations.h:414
                    if (w > 1.f) { w = 1.f; st.w[base + s] = 1.f; }
                    const uint32_t status = meta_status(m);
                    if (status == ST_GUESSED_BORN) gsum += w;
                    else if (status == ST_UPDATED && w < OCC_INIT_WEIGHT) {
                        st.meta[base + s] = meta_with_status(m, ST_INVALID);
                        new_live &= ~(1u << s);
                        continue;
                    }
                    const uint32_t tr = meta_track(m);
                    int k = 0;
                    while (k < nh && h_track[k] != tr) ++k;
                    if (k == nh) { h_track[k] = tr; h_w[k] = 0.f; ++nh; }
                    h_w[k] += w;
                    h_label[k] = meta_label(m) & 0xffu;
                }
                // first maximum in ascending track order, strict > from 0 (operations.h:439-446)
                float mw = 0.f;
                uint32_t best = 0xffffffffu;
                for (int k = 0; k < nh; ++k)
                    if (h_w[k] > 0.f && (h_w[k] > mw || (h_w[k] == mw && h_track[k] < best))) { mw = h_w[k]; best = h_track[k]; track = h_track[k]; label = h_label[k]; }
                if (any) ++n_nonempty;
                if (new_live != hdr_live(hdr)) st.hdr[v] = (hdr & 0x1ffffu) | (new_live << 16);
            }
            wsum[u] = ws; tl[u] = track | (label << 16);
            if (ws > f.occ_thr) state[u] = 1;
            else if (gsum >= OCC_INIT_WEIGHT) state[u] = 2;
            else state[u] = 0;
        }
        if (f.want_state) voxel_state[v] = (int8_t)state[u];
        n_occ += state[u] > 0;
        n_free += (f.want_free && state[u] == 0);
    }
    // block-wide exclusive scan of the two counters, packed
    uint32_t total;
    const uint32_t packed = n_occ | (n_free << 16);                 // <= 4 per thread, <= 1024 per block: 16 bits each suffice
    const uint32_t ex = block_exclusive_scan(packed, s_warp, total);
    if (threadIdx.x < 32) {
        const uint64_t tot64 = (uint64_t)(total & 0xffffu) | ((uint64_t)(total >> 16) << 31);
        const uint64_t prefix = lookback_prefix_warp(lb_status, bid, tot64);
        if (threadIdx.x == 0) {
            s_prefix = prefix;
            if (bid == gridDim.x - 1) {
                const uint64_t inc = prefix + tot64;
                cnt->n_occupied = (uint32_t)(inc & 0x7fffffffu);
                cnt->n_free = (uint32_t)((inc >> 31) & 0x7fffffffu);
            }
        }
    }
    n_nonempty = __reduce_add_sync(0xffffffffu, n_nonempty);
    if (lane_id() == 0 && n_nonempty) atomicAdd(&cnt->n_nonempty_voxels, n_nonempty);
    __syncthreads();
    if (n_occ | n_free) {
        uint32_t pos_occ = (uint32_t)(s_prefix & 0x7fffffffu) + (ex & 0xffffu);
        uint32_t pos_free = (uint32_t)((s_prefix >> 31) & 0x7fffffffu) + (ex >> 16);
#pragma unroll
        for (int u = 0; u < OCC_VPT; ++u) {
            const bool is_occ = state[u] > 0, is_free = f.want_free && state[u] == 0;
            if (!is_occ && !is_free) continue;
            // voxelIdxToMapFramePos + map centre (operations.h:940-944, 970-983)
            uint32_t mx, my, mz;
            buffer_to_map(f, bx0 + u, by, bz, mx, my, mz);
            float p0 = (float)mx * f.voxel + f.pmin[0], p1 = (float)my * f.voxel + f.pmin[1], p2 = (float)mz * f.voxel + f.pmin[2];
            p0 += f.map_center[0]; p1 += f.map_center[1]; p2 += f.map_center[2];
            if (f.zero_center) { p0 = p0 - f.cam[0]; p1 = p1 - f.cam[1]; p2 = p2 - f.cam[2]; }
            if (is_occ) {
                const uint32_t pos = pos_occ++;
                out.voxel_index[pos] = v0 + u;
                out.xyz[3 * (size_t)pos + 0] = p0; out.xyz[3 * (size_t)pos + 1] = p1; out.xyz[3 * (size_t)pos + 2] = p2;
                out.weight_sum[pos] = wsum[u];
                out.track_id[pos] = (uint16_t)(tl[u] & 0xffffu);
                out.label_id[pos] = (uint8_t)(tl[u] >> 16);
                out.state[pos] = (int8_t)state[u];
            } else {
                const uint32_t pos = pos_free++;
                out.free_voxel_index[pos] = v0 + u;
                out.free_xyz[3 * (size_t)pos + 0] = p0; out.free_xyz[3 * (size_t)pos + 1] = p1; out.free_xyz[3 * (size_t)pos + 2] = p2;
            }
        }
    }
}

// =====================================================================================================
// P2/P3  object moves  replaces moveParticlesInSetsByTransformations (mc_ring/operations.h:321-362) and
//                      ObjectSet::removeObjectByTrackID (object_layer.h:414-425).
// set(T) = { p : status(p) != INVALID and track(p) == T } in ascending particle index (canonical order).
// =====================================================================================================
constexpr int SEL_BLOCK = 1024;
constexpr int MAX_MOVES = 64;

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

// pass 1: ordered list of (particle index, move index) for every particle of a moved track; particles of a
// removed track are invalidated on the spot.  One thread per voxel; only voxels whose header carries the
// "movable" flag are opened (the flag is cleared when no live movable-track particle is left).
__global__ void __launch_bounds__(SEL_BLOCK)
k_select_tracks(int slot, Store st, const MoveTable* __restrict__ mt, uint32_t* __restrict__ sel_idx, uint32_t* __restrict__ sel_move,
                uint64_t* __restrict__ lb_status, Counters* __restrict__ cnt, uint32_t sel_cap) {
    __shared__ uint32_t s_bid, s_warp[33];
    __shared__ uint64_t s_prefix;
    const FrameParams& f = c_fp[slot];
    if (threadIdx.x == 0) s_bid = atomicAdd(&cnt->ticket_sel, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t v = (f.z_lo << (f.Xb + f.Yb)) + bid * SEL_BLOCK + threadIdx.x;
    const size_t base = (size_t)v << f.Pb;
    const uint32_t hdr = st.hdr[v];
    uint32_t sel = 0, n_removed = 0;
    if (hdr & HDR_MOVABLE) {
        uint32_t live = hdr_live(hdr), new_live = live;
        bool any_movable = false;
        while (live) {
            const uint32_t s = __ffs(live) - 1;
            live &= live - 1;
            const uint64_t m = st.meta[base + s];
            const int tr = (int)meta_track(m);
            if ((uint32_t)tr > f.max_movable) continue;
            if (find_track(mt->move_track, mt->n_moves, tr) >= 0) { sel |= 1u << s; any_movable = true; }
            else if (find_track(mt->remove_track, mt->n_remove, tr) >= 0) {
                st.meta[base + s] = meta_with_status(m, ST_INVALID);
                new_live &= ~(1u << s);
                ++n_removed;
            } else any_movable = true;
        }
        const uint32_t nh = hdr_ts(hdr) | (any_movable ? HDR_MOVABLE : 0u) | (new_live << 16);
        if (nh != hdr) st.hdr[v] = nh;
    }
    uint32_t total;
    const uint32_t ex = block_exclusive_scan((uint32_t)__popc(sel), s_warp, total);
    if (threadIdx.x < 32) {
        const uint64_t prefix = lookback_prefix_warp(lb_status, bid, total);
        if (threadIdx.x == 0) {
            s_prefix = prefix;
            if (bid == gridDim.x - 1) cnt->n_selected = (uint32_t)(prefix + total);
        }
    }
    n_removed = __reduce_add_sync(0xffffffffu, n_removed);
    if (lane_id() == 0 && n_removed) atomicAdd(&cnt->n_removed, n_removed);
    __syncthreads();
    uint32_t pos = (uint32_t)s_prefix + ex;
    while (sel) {
        const uint32_t s = __ffs(sel) - 1;
        sel &= sel - 1;
        if (pos < sel_cap) {
            sel_idx[pos] = (uint32_t)(base + s);
            sel_move[pos] = (uint32_t)find_track(mt->move_track, mt->n_moves, (int)meta_track(st.meta[base + s]));
        }
        ++pos;
    }
}

// pass 2: sequence number of every selected particle in the reference's processing order (object by object,
// ascending index inside an object) = offset[move] + rank among the same move's entries.  Single block.
__global__ void __launch_bounds__(1024)
k_move_sequence(const Counters* __restrict__ cnt, const MoveTable* __restrict__ mt, const uint32_t* __restrict__ sel_move,
                uint32_t* __restrict__ sel_seq, uint32_t sel_cap) {
    __shared__ uint32_t s_count[MAX_MOVES], s_offset[MAX_MOVES], s_running[MAX_MOVES];
    __shared__ uint32_t s_warp[33];
    const uint32_t n = min(cnt->n_selected, sel_cap);
    if (threadIdx.x < MAX_MOVES) { s_count[threadIdx.x] = 0; s_running[threadIdx.x] = 0; }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) atomicAdd(&s_count[sel_move[k]], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int m = 0; m < mt->n_moves; ++m) { s_offset[m] = acc; acc += s_count[m]; }
    }
    __syncthreads();
    // stable rank inside each move: chunks of blockDim entries, one move at a time
    for (uint32_t c = 0; c < n; c += blockDim.x) {
        const uint32_t k = c + threadIdx.x;
        const int my = k < n ? (int)sel_move[k] : -1;
        for (int m = 0; m < mt->n_moves; ++m) {
            if (s_count[m] == 0) continue;
            uint32_t total;
            const uint32_t ex = block_exclusive_scan(my == m ? 1u : 0u, s_warp, total);
            if (my == m) sel_seq[k] = s_offset[m] + s_running[m] + ex;
            __syncthreads();
            if (threadIdx.x == 0) s_running[m] += total;
            __syncthreads();
        }
    }
}

// pass 3: transform + table noise into a staging record, delete the original, bin by destination voxel
__global__ void __launch_bounds__(256)
k_move_apply(int slot, Counters* __restrict__ cnt, Store st, const MoveTable* __restrict__ mt, const float* __restrict__ randoms,
             const Persist* __restrict__ persist, const uint32_t* __restrict__ sel_idx, const uint32_t* __restrict__ sel_move,
             const uint32_t* __restrict__ sel_seq, float4* __restrict__ mov_pos, uint64_t* __restrict__ mov_meta,
             uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks, uint32_t* __restrict__ entry_list, uint32_t sel_cap) {
    const FrameParams& f = c_fp[slot];
    const uint32_t n = min(cnt->n_selected, sel_cap);
    const uint32_t cur = persist->cursor_ring;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t idx = sel_idx[k], seq = sel_seq[k];
        const float* t = mt->t[sel_move[k]];
        const float ox = st.x[idx], oy = st.y[idx], oz = st.z[idx];
        const float nx = ((t[0] * ox + t[1] * oy) + t[2] * oz) + t[3];       // operations.h:340
        const float ny = ((t[4] * ox + t[5] * oy) + t[6] * oz) + t[7];
        const float nz = ((t[8] * ox + t[9] * oy) + t[10] * oz) + t[11];
        // pre-incremented, wrapping cursor (utils/basic_algorithms.h:426-433): draw d uses (cursor + d) mod n
        const uint64_t d0 = (uint64_t)cur + 3ull * seq;
        const float px = nx + randoms[(d0 + 1) % f.rand_n];
        const float py = ny + randoms[(d0 + 2) % f.rand_n];
        const float pz = nz + randoms[(d0 + 3) % f.rand_n];
        const uint64_t m = st.meta[idx];
        mov_pos[seq] = make_float4(px, py, pz, st.w[idx]);
        mov_meta[seq] = m;
        st.meta[idx] = meta_with_status(m, ST_INVALID);                        // operations.h:346
        atomicAnd(st.hdr + (idx >> f.Pb), ~(1u << (16 + (idx & (f.P - 1)))));
        uint32_t bx, by, bz;
        const uint32_t v = global_pos_to_voxel(f, px, py, pz, bx, by, bz);
        if (v != INVALID_INDEX) topk_insert(f, v, seq, keys, ranks, entry_list, &cnt->n_move_voxels, &cnt->hash_overflow);
    }
}

__global__ void k_advance_cursor(const Counters* __restrict__ cnt, Persist* __restrict__ persist, uint32_t rand_n, uint32_t sel_cap) {
    const uint32_t n = min(cnt->n_selected, sel_cap);
    persist->cursor_ring = (uint32_t)(((uint64_t)persist->cursor_ring + 3ull * n) % rand_n);
}

// pass 4: re-insert in sequence order (operations.h:351-360); one thread per destination voxel
__global__ void __launch_bounds__(128)
k_move_replay(int slot, Store st, const uint32_t* __restrict__ stamp_x, const uint32_t* __restrict__ stamp_y,
              const uint32_t* __restrict__ stamp_z, const float4* __restrict__ mov_pos, const uint64_t* __restrict__ mov_meta,
              uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks, const uint32_t* __restrict__ entry_list,
              Counters* __restrict__ cnt) {
    const FrameParams& f = c_fp[slot];
    const uint32_t n_entries = cnt->n_move_voxels;
    for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_entries; li += gridDim.x * blockDim.x) {
        const uint32_t e = entry_list[li];
        const uint32_t v = keys[e];
        keys[e] = EMPTY_KEY;
        uint32_t bx, by, bz;
        voxel_to_buffer(f, v, bx, by, bz);
        const uint32_t stamp = slab_stamp(stamp_x, stamp_y, stamp_z, bx, by, bz);
        const size_t base = (size_t)v << f.Pb;
        const uint32_t hdr = st.hdr[v];
        uint32_t live = hdr_live(hdr);
        uint32_t* r = ranks + (size_t)e * f.topk;
        uint32_t moved = 0;
        bool full = false;
        for (uint32_t t = 0; t < f.topk; ++t) {
            const uint32_t seq = r[t];
            if (seq == NO_RANK) break;
            r[t] = NO_RANK;
            if (full) continue;
            const uint32_t s = first_vacant(st, base, f.P, live, stamp);
            if (!s) { full = true; continue; }
            const float4 p = mov_pos[seq];
            st.x[base + s] = p.x; st.y[base + s] = p.y; st.z[base + s] = p.z; st.w[base + s] = p.w;
            st.meta[base + s] = mov_meta[seq];
            live |= 1u << s;
            ++moved;
        }
        if (moved) {
            st.hdr[v] = hdr_ts(hdr) | HDR_MOVABLE | (live << 16);   // moved particles belong to a movable track by construction
            atomicAdd(&cnt->n_moved, moved);
        }
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

}  // namespace dspk
