// dspmap_engine.cu — host side of the B200 particle-map engine and its C ABI (include/dspmap_c.h).
//
// The host keeps exactly what the reference keeps in scalar globals — ring offsets, slab stamps, the time stamp,
// the camera matrices (mc_ring/buffer.h:97-120, utils/data_base.h:22) — and drives one stream of kernels per
// frame.  Every large array lives in HBM; nothing here computes on particles, pixels or voxels on the CPU and
// there is no CPU fallback: without a CUDA device dspmap_create fails with DSPMAP_ERR_NO_DEVICE.
#include "../../include/dspmap_c.h"
#include "dspmap_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <atomic>
#if defined(__AVX2__)
#include <immintrin.h>
#endif
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace dspk;

namespace {

#define CU_TRY(h, expr)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return (h)->fail(DSPMAP_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- small fixed-size linear algebra, evaluation order pinned (no FMA: host flags -ffp-contract=off) ----------
// Eigen::Quaternion::toRotationMatrix, row-major
template <typename T>
void quat_to_rot(const T q[4], T R[9]) {
    const T w = q[0], x = q[1], y = q[2], z = q[3];
    const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = T(1) - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = T(1) - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = T(1) - (txx + tyy);
}
// fixed 3x3 inverse: cofactors * (1/det)
void inverse3d(const double m[9], double o[9]) {
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
    };
    const double det = (cof(0, 0) * m[0] + cof(1, 0) * m[3]) + cof(2, 0) * m[6];
    const double invdet = 1.0 / det;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o[r * 3 + c] = cof(c, r) * invdet;
}
// fixed 4x4 float inverse: adjugate / determinant
void inverse4f(const float m[16], float o[16]) {
    auto M = [&](int r, int c) { return m[r * 4 + c]; };
    auto d3 = [&](int i1, int i2, int i3, int j1, int j2, int j3) { return M(i1, j1) * (M(i2, j2) * M(i3, j3) - M(i2, j3) * M(i3, j2)); };
    float adj[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4, j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
            const float c = (d3(i1, i2, i3, j1, j2, j3) + d3(i2, i3, i1, j1, j2, j3)) + d3(i3, i1, i2, j1, j2, j3);
            adj[j * 4 + i] = ((i + j) & 1) ? -c : c;
        }
    const float det = ((M(0, 0) * adj[0] + M(1, 0) * adj[1]) + M(2, 0) * adj[2]) + M(3, 0) * adj[3];
    for (int k = 0; k < 16; ++k) o[k] = adj[k] / det;
}

static inline void host_pause() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}

// A few persistent host threads for the pageable -> pinned staging copies of dspmap_update (row blocks of the depth
// image and of every mask are independent).  Workers sleep on a condition variable between frames.
//
// Work items are handed out as TICKETS: one 64-bit counter holding (job generation << 32 | next item).  A ticket names
// the job it was drawn for, the job's (function, item count) live in a two-entry table indexed by generation parity, and
// run() does not return before every thread that may hold a ticket of its job has left work() (the `active_` count) - so
// a helper that is descheduled between drawing a ticket and looking at the job can neither run an item of the NEXT job
// nor touch a function object that has gone out of scope.
class StagePool {
public:
    explicit StagePool(int n_workers, int spin_us = 200) : spin_us_(spin_us) {
        for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~StagePool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_.store(true); }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // runs fn(i) for i in [0, n) on the workers and the calling thread; returns when all are done
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        const uint64_t gen = (ticket_.load(std::memory_order_relaxed) >> 32) + 1;
        {
            std::lock_guard<std::mutex> lk(mu_);
            Job& j = jobs_[gen & 1];
            j.fn = &fn; j.n = n;
            pending_.store(n, std::memory_order_relaxed);
            ticket_.store(gen << 32, std::memory_order_release);          // publishes the job
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        work();
        // the tail of a staging copy is microseconds: spin.  active_ == 0: nobody is left inside work() with a ticket of this job
        while (pending_.load(std::memory_order_acquire) != 0 || active_.load(std::memory_order_acquire) != 0) cpu_relax();
    }

private:
    struct Job { const std::function<void(int)>* fn = nullptr; int n = 0; };
    static void cpu_relax() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    void work() {
        active_.fetch_add(1, std::memory_order_acq_rel);
        for (;;) {
            const uint64_t t = ticket_.fetch_add(1, std::memory_order_acq_rel);
            const Job& j = jobs_[(t >> 32) & 1];                          // the job this ticket was drawn for
            const int i = (int)(uint32_t)t;
            if ((t >> 32) == 0 || i >= j.n) break;
            (*j.fn)(i);
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
        active_.fetch_sub(1, std::memory_order_acq_rel);
    }
    // A frame stages twice within ~0.1 ms (depth, then the masks under the sweep): after a job the workers keep polling for
    // the next one for a short while before they block on the condition variable, so the second job starts without a wake-up.
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const auto t0 = std::chrono::steady_clock::now();
            int spins = 0;
            while (!stop_.load(std::memory_order_acquire) && (ticket_.load(std::memory_order_acquire) >> 32) == seen) {
                cpu_relax();
                if ((++spins & 255) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) {
                    std::unique_lock<std::mutex> lk(mu_);
                    sleepers_.fetch_add(1, std::memory_order_acq_rel);
                    cv_.wait(lk, [&] { return stop_.load() || (ticket_.load(std::memory_order_acquire) >> 32) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                    break;
                }
            }
            if (stop_.load(std::memory_order_acquire)) return;
            seen = ticket_.load(std::memory_order_acquire) >> 32;
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    Job jobs_[2];
    std::atomic<uint64_t> ticket_{0};
    std::atomic<int> pending_{0}, active_{0}, sleepers_{0};
    std::atomic<bool> stop_{false};
    const int spin_us_;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        return count ? cudaMalloc(&p, count * sizeof(T)) : cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = std::max(count, n + n / 2);
        return cudaMallocHost(&p, n * sizeof(T));
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

}  // namespace

struct dspmap {
    dspmap_config cfg{};
    dspmap_params prm{};
    FrameParams fp{};
    std::string err = "";
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, own_stream = nullptr, side_stream = nullptr;   // side: independent kernels run beside the main chain
    cudaEvent_t ev_fork[3]{}, ev_join[3]{};
    uint64_t launches = 0;

    // host ring state (mc_ring/buffer.h:97-120)
    std::vector<uint32_t> stamps[3];
    int moved[3]{}, equiv[3]{};
    float map_center[3]{}, ego_center[3]{}, last_pos[3]{};
    uint32_t T = 0;
    bool forgetting_init = false;
    float forget_tab[5]{};

    // label tables
    uint16_t h_static_track[256];
    std::vector<uint8_t> h_static_label;  // 65536
    uint32_t max_movable = 65523, sky_track = 65533;
    uint32_t rand_n = 1;

    // device: store (allocated for this shard's slabs; pointers are biased so the global particle index applies)
    DevBuf<float> sx, sy, sz, sw;
    DevBuf<uint64_t> smeta;
    DevBuf<uint32_t> shdr, d_entry_list, d_entry_voxel;
    Store st{};
    size_t n_slots_local = 0, slot_lo = 0;
    uint32_t v_local = 0;
    DevBuf<float> d_lut, d_randoms;
    DevBuf<uint16_t> d_static_track;
    DevBuf<uint8_t> d_static_label;
    DevBuf<FrameParams> d_fp;
    DevBuf<Counters> d_cnt;
    DevBuf<Persist> d_persist;
    // images
    DevBuf<float> d_depth_int;
    DevBuf<float4> d_cloud_pos;
    DevBuf<uint32_t> d_cloud_meta, d_pix_count, d_pix_offset, d_scan_block;
    // visibility
    DevBuf<uint32_t> d_vmask;
    uint32_t words_per_row = 0, n_vwords = 0;
    DevBuf<uint64_t> d_csr_key;
    DevBuf<uint32_t> d_csr_idx, d_seq_ref, d_mov_out, d_mine_list;
    DevBuf<uint32_t> d_vrank, d_bfs_front[2], d_bfs_blocks, d_work_key;   // reference iteration order (k_bfs_order), allocated on first use
    int bfs_grid = 0;
    DevBuf<uint32_t> d_row_order;      // image rows by listed particles in their window, descending (k_ck_row_order)
    DevBuf<uint2> d_vis_mp;            // list-ordered {track | forget << 16, pixel}
    DevBuf<float4> d_vis_pos;
    // ordered insertion
    DevBuf<uint32_t> d_keys, d_ranks, d_birth_flags, d_birth_prefix, d_birth_block;
    DevBuf<BirthBig> d_birth_big;     // touched voxels whose candidate search runs block-wide
    DevBuf<BirthBox> d_birth_box;     // per hash slot of the birth table: bounding rectangle of the voxel's candidate pixels
    // object moves
    DevBuf<uint4> d_move_blk;           // per-block counts of the moved tracks' list entries (k_move_count_blocks)
    DevBuf<uint32_t> d_movl_idx, d_movl_seq;
    DevBuf<uint16_t> d_movl_track, d_mov_mask;
    DevBuf<uint8_t> d_vstate;
    DevBuf<float> d_res_w;
    DevBuf<uint32_t> d_res_tl;
    DevBuf<uint4> d_blk_cnt, d_blk_off, d_chunk_tot, d_chunk_base;
    DevBuf<uint32_t> d_tile_list;
    uint32_t n_occ_chunks = 0;
    DevBuf<uint32_t> d_work, d_work_hdr;      // work lists of the per-voxel passes; header words beside the sweep's list
    uint32_t n_occ_blocks = 0, birth_cap1 = 0, birth_shift1 = 0;
    uint32_t sel_cap = 0, seq_cap = 0;
    // occupancy
    // exchange arena (one allocation, exported over CUDA IPC in sharded maps): barrier flags, counts, visible records,
    // incoming moved particles, the ck image and the (global) result arrays
    DevBuf<uint8_t> arena;
    struct ArenaLayout { size_t flags, counts, move_counts, recs, movs, obs_a, obs_b, r_idx, r_xyz, r_w, r_track, r_label, r_state, r_free_idx, r_free_xyz, r_rgb, total, res2; } lay{};
    CommView cv{};
    uint32_t comm_epoch = 0;
    void* peer_base[MAX_RANKS]{};
    bool comm_connected = false;
    float4* d_obs_a = nullptr;        // observation planes of the stencil passes (padded; dspmap_stencil.cuh)
    uint2* d_obs_b = nullptr;
    DevBuf<uint8_t> result2;           // second set of result arrays (unsharded maps; allocated by the first pipelined frame)
    dspmap_colors colors{};            // dspmap_set_colors
    DevBuf<uint8_t> d_col_tab;         // 768 B label colours (b, g, r) + 256 B instance map
    uint32_t* d_counts = nullptr;      // this rank's counts[source rank][COMM_COUNTS]

    // Frame slots.  Everything the host writes for ONE frame and the device reads later (pinned staging of the images,
    // descriptor tables, slab stamps), everything the device writes for one frame and the host reads later (the counter
    // block, the result arrays and their pinned host copies) exists twice, indexed by the parity of the frame's time stamp:
    // frame t+1 can be staged and enqueued while frame t still runs and frame t-1's result is on its way to the host
    // (dspmap_update_async).  The synchronous calls use the same slots, one frame at a time.
    struct FrameSlot {
        PinBuf<float> p_depth;
        PinBuf<uint8_t> p_masks;
        PinBuf<Counters> p_cnt;
        PinBuf<uint32_t> p_stamps;
        PinBuf<MaskDesc> p_maskdesc;
        PinBuf<MoveTable> p_moves;
        PinBuf<uint32_t> hr_idx, hr_free_idx;
        PinBuf<float> hr_xyz, hr_w, hr_free_xyz;
        PinBuf<uint16_t> hr_track;
        PinBuf<uint8_t> hr_label, hr_rgb;
        PinBuf<int8_t> hr_state;
        DevBuf<float> d_depth_in;          // where dspmap_update uploads the frame's images (on the upload stream, so that
        DevBuf<uint8_t> d_mask_in;         // frame t+1's images travel while frame t's kernels run)
        DevBuf<uint32_t> d_stamps;         // x | y | z slab stamps, mask descriptors, move tables of the slot's frame: uploaded on
        DevBuf<MaskDesc> d_masks;          // the upload stream as well - a copy-engine operation IN the kernels' stream costs
        DevBuf<MoveTable> d_moves;         // the stream several microseconds each time, a wait for a finished upload next to nothing
        bool stamps_stale = true;          // the slot's device copy of the stamps is older than the host's
        cudaEvent_t up_depth = nullptr, up_masks = nullptr, up_tab = nullptr, frame_done = nullptr;
        cudaEvent_t tr_begin = nullptr, tr_end = nullptr;   // DSPMAP_TRACE: the frame's first and last operation on the main stream
        ResultArrays dres{};               // this slot's result arrays on the device
        cudaEvent_t copy_done = nullptr;   // the device->host transfer of this slot's result
        bool copy_issued = false;          // ... has been enqueued since the slot's frame finished
        bool host_result_valid = false;    // ... and has completed
        bool inflight = false;             // the slot's frame is enqueued, its counters have not been read yet
        bool want_host = false;            // the frame came through dspmap_update_async: its result goes to the host
        int want_free = 0, col_enabled = 0;
        uint32_t T = 0;                    // time stamp of the slot's frame (the ticket of dspmap_update_async)
        uint32_t result_n = 0, result_nf = 0, n_occ_global = 0;
        int rc = DSPMAP_OK;                // what finishing the frame returned
        double t_issued = 0;
        void release() {
            p_depth.release(); p_masks.release(); p_cnt.release(); p_stamps.release(); p_maskdesc.release(); p_moves.release();
            hr_idx.release(); hr_free_idx.release(); hr_xyz.release(); hr_w.release(); hr_free_xyz.release(); hr_track.release();
            hr_label.release(); hr_rgb.release(); hr_state.release();
            d_depth_in.release(); d_mask_in.release(); d_stamps.release(); d_masks.release(); d_moves.release();
            for (cudaEvent_t* e : {&copy_done, &up_depth, &up_masks, &up_tab, &frame_done, &tr_begin, &tr_end}) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
        }
    } slot[2];
    int cur = 0;                           // slot of the most recent frame
    cudaStream_t copy_stream = nullptr;    // result transfers of pipelined frames
    cudaStream_t upload_stream = nullptr;  // image uploads of dspmap_update
    double t_frame_begin = 0, trace_issue_ms = 0, trace_wait_ms = 0, trace_prev_ms = 0, trace_copy_issue_ms = 0, trace_result_ms = 0, trace_gpu_ms = 0, trace_gap_ms = 0;   // DSPMAP_TRACE
    std::unique_ptr<StagePool> stage_pool;

    Counters last{};                       // of the most recent FINISHED frame
    uint32_t n_visible_last = 0;
    dspmap_frame_stats stats{};
    bool profiling = false;
    cudaEvent_t ev[10]{};

    int fail(int rc, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return rc;
    }

    // ---- ring shift: RingBufferOperations::updateEgoCenterPos / updateRingbufferIndexParams
    //      (mc_ring/operations.h:68-96, 1111-1230) — O(slabs moved) scalar work, stays on the host ----
    uint32_t axis_n(int a) const { return a == 0 ? fp.NX : (a == 1 ? fp.NY : fp.NZ); }
    uint32_t wrap(int idx, int a) const {
        const int n = (int)axis_n(a);
        return idx < 0 ? (uint32_t)(idx + n) : (idx >= n ? (uint32_t)(idx - n) : (uint32_t)idx);
    }
    void ring_params() {
        int center[3];
        for (int a = 0; a < 3; ++a) {
            center[a] = static_cast<int>(ego_center[a] * fp.voxel_recip);
            map_center[a] = static_cast<float>(center[a]) * fp.voxel;
        }
        for (int a = 0; a < 3; ++a) {
            const int n = (int)axis_n(a), d = center[a] - moved[a];
            for (int i = 0; i < d; ++i) { stamps[a][wrap(i + equiv[a], a)] = T; stamps_changed(); }
            for (int i = 0; i < -d; ++i) { stamps[a][wrap(n - 1 - i + equiv[a], a)] = T; stamps_changed(); }
        }
        for (int a = 0; a < 3; ++a) {
            const int n = (int)axis_n(a);
            moved[a] = center[a];
            equiv[a] = moved[a] > 0 ? moved[a] % n : (moved[a] < 0 ? -(-moved[a] % n) : 0);
        }
    }
    void ring_shift(const float pos[3]) {
        const float lim = std::min(std::min((1 << (fp.Xb - 2)) * fp.voxel, (1 << (fp.Yb - 2)) * fp.voxel), (1 << (fp.Zb - 2)) * fp.voxel);
        float mv[3] = {pos[0] - last_pos[0], pos[1] - last_pos[1], pos[2] - last_pos[2]};
        const float sq = (mv[0] * mv[0] + mv[1] * mv[1]) + mv[2] * mv[2];
        float unit[3] = {mv[0], mv[1], mv[2]};
        if (sq > 0.f) { const float nn = sqrtf(sq); for (int a = 0; a < 3; ++a) unit[a] = mv[a] / nn; }
        float dist = sqrtf(sq);
        float np[3] = {last_pos[0], last_pos[1], last_pos[2]};
        while (dist > lim) {
            for (int a = 0; a < 3; ++a) np[a] = np[a] + unit[a] * lim;
            for (int a = 0; a < 3; ++a) ego_center[a] = np[a];
            ring_params();
            for (int a = 0; a < 3; ++a) mv[a] = pos[a] - np[a];
            dist = sqrtf((mv[0] * mv[0] + mv[1] * mv[1]) + mv[2] * mv[2]);
        }
        for (int a = 0; a < 3; ++a) ego_center[a] = pos[a];
        ring_params();
        for (int a = 0; a < 3; ++a) last_pos[a] = pos[a];
    }

    void stamps_changed() { slot[0].stamps_stale = slot[1].stamps_stale = true; }
    void reset_host_state() {
        T = 0;
        for (int a = 0; a < 3; ++a) std::fill(stamps[a].begin(), stamps[a].end(), 0u);
        stamps_changed();
    }
    // host table -> the slot's device copy, on the upload stream; the kernels' stream waits for it (cheap once the copy is done)
    cudaError_t upload_table(FrameSlot& sl, void* dst, const void* src, size_t bytes) {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, upload_stream);
        if (e == cudaSuccess) e = cudaEventRecord(sl.up_tab, upload_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, sl.up_tab, 0);
        return e;
    }

    void release_all() {
        sx.release(); sy.release(); sz.release(); sw.release(); smeta.release(); shdr.release(); d_entry_list.release(); d_entry_voxel.release(); d_lut.release(); d_randoms.release();
        d_static_track.release(); d_static_label.release(); d_fp.release(); d_cnt.release(); d_persist.release();
        d_depth_int.release(); d_cloud_pos.release(); d_cloud_meta.release();
        d_pix_count.release(); d_pix_offset.release(); d_scan_block.release(); d_vmask.release(); d_csr_key.release(); d_seq_ref.release(); d_mov_out.release(); d_csr_idx.release(); d_mine_list.release();
        d_vis_mp.release(); d_row_order.release(); d_vrank.release(); d_bfs_front[0].release(); d_bfs_front[1].release(); d_bfs_blocks.release(); d_work_key.release(); d_vis_pos.release(); d_keys.release(); d_birth_box.release(); d_birth_big.release(); d_ranks.release(); d_birth_flags.release(); d_birth_prefix.release(); d_birth_block.release(); d_move_blk.release(); d_movl_idx.release(); d_movl_seq.release(); d_movl_track.release(); d_mov_mask.release(); d_vstate.release(); d_res_w.release(); d_res_tl.release(); d_blk_cnt.release(); d_blk_off.release(); d_chunk_tot.release(); d_chunk_base.release(); d_tile_list.release(); d_work.release(); d_work_hdr.release();
        for (int p = 0; p < MAX_RANKS; ++p) if (peer_base[p] && p != cfg.shard_rank) { cudaIpcCloseMemHandle(peer_base[p]); peer_base[p] = nullptr; }
        arena.release();
        result2.release();
        for (auto& sl : slot) sl.release();
        if (copy_stream) { cudaStreamDestroy(copy_stream); copy_stream = nullptr; }
        if (upload_stream) { cudaStreamDestroy(upload_stream); upload_stream = nullptr; }
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (own_stream) { cudaStreamDestroy(own_stream); own_stream = nullptr; }
        if (side_stream) { cudaStreamDestroy(side_stream); side_stream = nullptr; }
        for (auto& e : ev_fork) if (e) { cudaEventDestroy(e); e = nullptr; }
        for (auto& e : ev_join) if (e) { cudaEventDestroy(e); e = nullptr; }
        stream = nullptr;
    }

    // (re)builds the comm view from the mapped arenas; result arrays always point at rank 0's arena once connected
    void bind_arena(int have_rank0) {
        const int G = cfg.shard_count, me = cfg.shard_rank;
        cv.rank = me; cv.size = G; cv.rec_cap = fp.vis_cap; cv.mov_cap = sel_cap;
        for (int p = 0; p < G; ++p) {
            uint8_t* b = (uint8_t*)(peer_base[p] ? peer_base[p] : arena.p);
            cv.flags[p] = (uint32_t*)(b + lay.flags); cv.counts[p] = (uint32_t*)(b + lay.counts); cv.move_counts[p] = (uint32_t*)(b + lay.move_counts);
            cv.recs[p] = (VisRec*)(b + lay.recs); cv.movs[p] = (MovRec*)(b + lay.movs); cv.obs_a[p] = (float4*)(b + lay.obs_a); cv.obs_b[p] = (uint2*)(b + lay.obs_b);
        }
        uint8_t* mine = arena.p;
        d_obs_a = (float4*)(mine + lay.obs_a); d_obs_b = (uint2*)(mine + lay.obs_b); d_counts = (uint32_t*)(mine + lay.counts);
        uint8_t* r0 = (uint8_t*)(peer_base[0] && have_rank0 >= 0 ? peer_base[0] : arena.p);
        slot[0].dres = result_arrays_at(r0, 0);
        slot[1].dres = lay.res2 ? result_arrays_at(r0 + lay.res2, lay.r_idx) : (result2.p ? result_arrays_at(result2.p, lay.r_idx) : slot[0].dres);
    }
    // the nine result arrays inside a block laid out like the arena's result part (`bias` = offset of that part's first array)
    ResultArrays result_arrays_at(uint8_t* b, size_t bias) const {
        b -= bias;
        return ResultArrays{(uint32_t*)(b + lay.r_idx), (float*)(b + lay.r_xyz), (float*)(b + lay.r_w), (uint16_t*)(b + lay.r_track), (uint8_t*)(b + lay.r_label),
                            (int8_t*)(b + lay.r_state), (uint32_t*)(b + lay.r_free_idx), (float*)(b + lay.r_free_xyz), b + lay.r_rgb};
    }
    int barrier() {
        if (cfg.shard_count == 1) return DSPMAP_OK;
        ++launches; k_comm_barrier<<<1, 32, 0, stream>>>(cv, ++comm_epoch, &d_cnt.p->comm_errors);
        return DSPMAP_OK;
    }
    int init(const dspmap_config& c);
    int clear_device();
    int apply_params();
    int run_moves(const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove, bool scan_remove, int slot);
    int frame(const float* d_depth, int rows, int cols, int stride, const dspmap_mask* masks, int n_masks, const double pos[3],
              const double quat[4], const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove,
              int want_free, int pre_decimated, const std::function<int(dspmap_mask*)>* late_masks = nullptr,
              const std::function<int()>* late_depth = nullptr, bool pipelined = false, bool want_host = false);
    int frame_enqueue(const float* d_depth, int rows, int cols, int stride, const dspmap_mask* masks, int n_masks, const double pos[3],
                      const double quat[4], const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove,
                      int want_free, int pre_decimated, const std::function<int(dspmap_mask*)>* late_masks, const std::function<int()>* late_depth);
    int frame_finish(int s);
    int issue_result_copy(int s, cudaStream_t on);
    int fetch_result_to_host(int s);
    int drain(bool whole_frame = true);
    int next_slot() const { return (int)((T + 1) & 1u); }
};

// kernels templated on the slot-count bits
#define DISPATCH_PB(pb, CALL)                                   \
    switch (pb) {                                               \
        case 1: { constexpr int PB = 1; CALL; break; }          \
        case 2: { constexpr int PB = 2; CALL; break; }          \
        case 3: { constexpr int PB = 3; CALL; break; }          \
        default: { constexpr int PB = 4; CALL; break; }         \
    }

// ------------------------------------------------------------------------------------------------------------
int dspmap::init(const dspmap_config& c) {
    cfg = c;
    // runSystemChecking, mc_ring/operations.h:54-58
    if (c.voxel_bits_x < 2 || c.voxel_bits_y < 2 || c.voxel_bits_z < 2 || c.slot_bits < 1 ||
        c.voxel_bits_x + c.voxel_bits_y + c.voxel_bits_z + c.slot_bits > 31)
        return fail(DSPMAP_ERR_CONFIG, "grid bits (%d,%d,%d)+%d must each be >= 2 (slots >= 1) and sum to <= 31", c.voxel_bits_x, c.voxel_bits_y, c.voxel_bits_z, c.slot_bits);
    if ((1 << c.slot_bits) > MAX_P) return fail(DSPMAP_ERR_CONFIG, "at most %d slots per voxel", MAX_P);
    if (!(c.voxel_size > 0.f) || c.image_width < 2 || c.image_height < 2 || !(c.fx > 0.f) || !(c.fy > 0.f) || !(c.depth_max > c.depth_min))
        return fail(DSPMAP_ERR_CONFIG, "bad voxel size / image size / intrinsics / depth range");
    if (c.shard_count < 1 || c.shard_rank < 0 || c.shard_rank >= c.shard_count || ((1 << c.voxel_bits_z) % c.shard_count) != 0)
        return fail(DSPMAP_ERR_CONFIG, "shard_count must divide the storage-z extent and shard_rank must be in range");
    if (c.shard_count > MAX_RANKS) return fail(DSPMAP_ERR_CONFIG, "at most %d shards", MAX_RANKS);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(DSPMAP_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
    }
    if (c.device < 0 || c.device >= ndev) return fail(DSPMAP_ERR_INVALID_ARG, "device %d of %d", c.device, ndev);
    device = c.device;
    CU_TRY(this, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(this, cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    CU_TRY(this, cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    CU_TRY(this, cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    for (auto& e : ev_fork) CU_TRY(this, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ev_join) CU_TRY(this, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ev) CU_TRY(this, cudaEventCreate(&e));

    FrameParams& f = fp;
    memset(&f, 0, sizeof f);
    f.Xb = c.voxel_bits_x; f.Yb = c.voxel_bits_y; f.Zb = c.voxel_bits_z; f.Pb = c.slot_bits;
    f.NX = 1u << f.Xb; f.NY = 1u << f.Yb; f.NZ = 1u << f.Zb; f.P = 1u << f.Pb;
    f.V = f.NX * f.NY * f.NZ;
    f.voxel = c.voxel_size;
    f.voxel_recip = 1.f / f.voxel;                               // operations.h:766
    f.pmax[0] = (f.NX >> 1) * f.voxel; f.pmax[1] = (f.NY >> 1) * f.voxel; f.pmax[2] = (f.NZ >> 1) * f.voxel;  // :735-741
    for (int a = 0; a < 3; ++a) f.pmin[a] = -f.pmax[a];
    if (c.boost_mode) {                                          // settings.h:136-143
        const float s = 0.5f;
        f.fx = s * c.fx; f.fy = s * c.fy; f.cx = s * c.cx; f.cy = s * c.cy;
        f.W = (int)(s * c.image_width); f.H = (int)(s * c.image_height);
    } else {
        f.fx = c.fx; f.fy = c.fy; f.cx = c.cx; f.cy = c.cy; f.W = c.image_width; f.H = c.image_height;
    }
    f.in_W = c.image_width; f.in_H = c.image_height; f.boost = c.boost_mode ? 1 : 0;
    f.depth_min = c.depth_min; f.depth_max = c.depth_max;
    f.dmin = (double)c.depth_min; f.dmax = (double)c.depth_max;
    f.tan_hx = (float)tan(atan2(f.W / 2.0, (double)f.fx));      // operations.h:1249-1250
    f.tan_hy = (float)tan(atan2(f.H / 2.0, (double)f.fy));
    f.win_r = c.window_half > 0 ? c.window_half : (c.boost_mode ? 3 : 5);   // semantic_dsp_map.h:964-970
    f.zed2_filters = c.zed2_filters ? 1 : 0;
    f.consider_instance = c.consider_instance ? 1 : 0;
    const uint32_t slabs = f.NZ / (uint32_t)c.shard_count;
    f.z_lo = slabs * (uint32_t)c.shard_rank; f.z_hi = f.z_lo + slabs;
    f.topk = 2 * f.P - 1;
    const double Km[9] = {(double)f.fx, 0, (double)f.cx, 0, (double)f.fy, (double)f.cy, 0, 0, 1};   // pointcloud_tools.h:113
    inverse3d(Km, f.Kinv);

    // defaults of the SemanticDSPMap constructor (semantic_dsp_map.h:25-43)
    prm.detection_probability = 0.95f; prm.noise_number = 0.1f; prm.nb_ptc_num_per_point = 3; prm.occupancy_threshold = 0.2f;
    prm.forgetting_rate = 1.0f; prm.max_forget_count = 5; prm.id_transition_probability = 0.1f; prm.depth_noise_first_order = 0.0f;
    prm.depth_noise_zero_order = 0.1f; prm.consider_depth_noise = 0; prm.use_independent_filter = 0; prm.zero_center = 0;

    const size_t npix = (size_t)f.W * f.H, npix_in = (size_t)f.in_W * f.in_H;
    v_local = slabs << (f.Xb + f.Yb);
    n_slots_local = (size_t)v_local << f.Pb;
    slot_lo = ((size_t)f.z_lo << (f.Xb + f.Yb)) << f.Pb;
    CU_TRY(this, sx.alloc(n_slots_local)); CU_TRY(this, sy.alloc(n_slots_local)); CU_TRY(this, sz.alloc(n_slots_local));
    CU_TRY(this, sw.alloc(n_slots_local)); CU_TRY(this, smeta.alloc(n_slots_local));
    constexpr size_t PAD = OCC_TILE;
    const size_t v_pad = ((size_t)v_local + PAD - 1) / PAD * PAD;   // whole tiles for the per-voxel passes
    CU_TRY(this, shdr.alloc(v_pad));
    CU_TRY(this, cudaMemset(shdr.p, 0, v_pad * sizeof(uint32_t)));
    st.x = sx.p - slot_lo; st.y = sy.p - slot_lo; st.z = sz.p - slot_lo; st.w = sw.p - slot_lo; st.meta = smeta.p - slot_lo;
    st.hdr = shdr.p - (slot_lo >> f.Pb);

    for (int a = 0; a < 3; ++a) stamps[a].assign(axis_n(a), 0u);
    for (auto& sl : slot) CU_TRY(this, sl.d_stamps.alloc(f.NX + f.NY + f.NZ));
    CU_TRY(this, d_lut.alloc(PDF_NUM));
    CU_TRY(this, d_randoms.alloc(1));
    CU_TRY(this, cudaMemset(d_randoms.p, 0, sizeof(float)));
    rand_n = 1;
    CU_TRY(this, d_static_track.alloc(256)); CU_TRY(this, d_static_label.alloc(65536));
    for (int i = 0; i < 256; ++i) h_static_track[i] = 65535;
    h_static_label.assign(65536, 0);
    CU_TRY(this, cudaMemcpy(d_static_track.p, h_static_track, sizeof h_static_track, cudaMemcpyHostToDevice));
    CU_TRY(this, cudaMemset(d_static_label.p, 0, 65536));
    CU_TRY(this, d_fp.alloc(1)); CU_TRY(this, d_cnt.alloc(1)); CU_TRY(this, d_persist.alloc(1));
    for (auto& sl : slot) { CU_TRY(this, sl.d_moves.alloc(1)); CU_TRY(this, sl.d_masks.alloc(64)); }
    CU_TRY(this, cudaMemset(d_persist.p, 0, sizeof(Persist)));
    for (auto& sl : slot) CU_TRY(this, sl.d_depth_in.alloc(npix_in));
    CU_TRY(this, d_depth_int.alloc(npix));
    CU_TRY(this, d_cloud_pos.alloc(npix)); CU_TRY(this, d_cloud_meta.alloc(npix));
    CU_TRY(this, d_pix_count.alloc(npix)); CU_TRY(this, d_pix_offset.alloc(npix + 1));
    CU_TRY(this, d_scan_block.alloc((npix + SCAN_BLOCK - 1) / SCAN_BLOCK));
    words_per_row = (f.NX + 1 + 31) / 32;
    n_vwords = words_per_row * (f.NY + 1) * (f.NZ + 1);
    CU_TRY(this, d_vmask.alloc(n_vwords));

    const size_t usable = n_slots_local - v_local;
    // capacity of ONE rank's visible-record segment; the pixel lists hold the records of all ranks
    f.vis_cap = (uint32_t)std::min<size_t>(usable, c.max_visible ? (size_t)c.max_visible : (size_t)4 << 20);
    const size_t list_cap = (size_t)f.vis_cap * c.shard_count;
    CU_TRY(this, d_csr_key.alloc(list_cap)); CU_TRY(this, d_csr_idx.alloc(list_cap));
    if (c.shard_count > 1) CU_TRY(this, d_mine_list.alloc(list_cap + 32));
    CU_TRY(this, d_vis_mp.alloc(list_cap)); CU_TRY(this, d_vis_pos.alloc(list_cap)); CU_TRY(this, d_row_order.alloc(f.H));

    sel_cap = (uint32_t)std::min<size_t>(usable, (size_t)1 << 20);
    CU_TRY(this, d_move_blk.alloc((sel_cap + MPB_CHUNK - 1) / MPB_CHUNK));
    CU_TRY(this, d_movl_idx.alloc(sel_cap)); CU_TRY(this, d_movl_seq.alloc(sel_cap)); CU_TRY(this, d_movl_track.alloc(sel_cap));
    seq_cap = (size_t)sel_cap * c.shard_count >= ((size_t)1 << 24) ? (1u << 24) : sel_cap * (uint32_t)c.shard_count;   // sequence numbers are global (and 24 bits in the rank lists)
    CU_TRY(this, d_seq_ref.alloc(seq_cap)); CU_TRY(this, d_mov_out.alloc(MAX_RANKS));
    CU_TRY(this, cudaMemset(d_mov_out.p, 0, MAX_RANKS * sizeof(uint32_t)));

    // hash table for ordered insertion: >= 2x the larger candidate population (pixels, moved particles)
    size_t want = 2 * std::max<size_t>(npix, sel_cap);
    uint32_t cap = 1024, bits = 10;
    while (cap < want) { cap <<= 1; ++bits; }
    f.hash_cap = cap; f.hash_shift = 32 - bits;
    {
        uint32_t bcap = 1024, bbits = 10;
        while (bcap < 2 * npix) { bcap <<= 1; ++bbits; }
        f.birth_hash_cap = bcap; f.birth_hash_shift = 32 - bbits;
        birth_cap1 = bcap; birth_shift1 = 32 - bbits;
    }
    CU_TRY(this, d_keys.alloc(cap)); CU_TRY(this, d_ranks.alloc((size_t)cap * f.topk)); CU_TRY(this, d_entry_list.alloc(cap)); CU_TRY(this, d_entry_voxel.alloc(cap));
    CU_TRY(this, d_birth_box.alloc(f.birth_hash_cap)); CU_TRY(this, d_birth_big.alloc(f.birth_hash_cap));

    n_occ_blocks = (uint32_t)(v_pad / OCC_TILE);
    CU_TRY(this, d_col_tab.alloc(1024)); CU_TRY(this, cudaMemset(d_col_tab.p, 0, 1024));
    CU_TRY(this, d_vstate.alloc(v_pad)); CU_TRY(this, d_res_w.alloc(v_pad)); CU_TRY(this, d_res_tl.alloc(v_pad)); CU_TRY(this, d_mov_mask.alloc(v_pad));
    n_occ_chunks = (n_occ_blocks + OCC_SCAN_CHUNK - 1) / OCC_SCAN_CHUNK;
    CU_TRY(this, d_chunk_tot.alloc(n_occ_chunks + 1)); CU_TRY(this, d_chunk_base.alloc(n_occ_chunks));
    CU_TRY(this, cudaMemset(d_chunk_tot.p, 0, (n_occ_chunks + 1) * sizeof(uint4))); CU_TRY(this, cudaMemset(d_chunk_base.p, 0, n_occ_chunks * sizeof(uint4)));
    CU_TRY(this, d_tile_list.alloc((size_t)n_occ_chunks * OCC_SCAN_CHUNK));
    CU_TRY(this, d_blk_cnt.alloc(n_occ_blocks)); CU_TRY(this, d_blk_off.alloc(n_occ_blocks)); CU_TRY(this, d_work.alloc(v_pad + (size_t)sm_count * 8 * 8 * 512 + 512));
    CU_TRY(this, d_work_hdr.alloc(v_pad));
    CU_TRY(this, cudaMemset(d_vstate.p, 0, v_pad));
    {   // exchange arena
        const int G = c.shard_count;
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
        lay.flags = take(64 * 4); lay.counts = take((size_t)MAX_RANKS * COMM_COUNTS * 4); lay.move_counts = take((size_t)MAX_RANKS * MAX_MOVES * 4);
        lay.recs = take((size_t)G * f.vis_cap * sizeof(VisRec)); lay.movs = take((size_t)G * sel_cap * sizeof(MovRec));
        const size_t n_obs = obs_cells(f.W, f.H, f.win_r);
        lay.obs_a = take(n_obs * 16); lay.obs_b = take(n_obs * 8);
        const size_t Vg = f.V;   // rank 0 receives the global result
        lay.r_idx = take(Vg * 4); lay.r_xyz = take(Vg * 12); lay.r_w = take(Vg * 4); lay.r_track = take(Vg * 2); lay.r_label = take(Vg); lay.r_state = take(Vg);
        lay.r_free_idx = take(Vg * 4); lay.r_free_xyz = take(Vg * 12);
        lay.r_rgb = take(Vg * 3);
        lay.total = off;
        // a sharded map's result arrays live in rank 0's arena (the peers write into them): both sets of a pipelined map are part
        // of every rank's arena, so that the layout is the same everywhere; an unsharded map allocates its second set on demand
        lay.res2 = G > 1 ? lay.total : 0;
        CU_TRY(this, arena.alloc(lay.total + (G > 1 ? lay.total - lay.r_idx : 0)));
        CU_TRY(this, cudaMemset(arena.p, 0, lay.flags + 64 * 4 + (size_t)MAX_RANKS * COMM_COUNTS * 4 + (size_t)MAX_RANKS * MAX_MOVES * 4 + 1024));
        CU_TRY(this, cudaMemset(arena.p + lay.obs_a, 0, n_obs * 16)); CU_TRY(this, cudaMemset(arena.p + lay.obs_b, 0, n_obs * 8));
        peer_base[c.shard_rank] = arena.p;
        bind_arena(c.shard_rank == 0 ? 0 : -1);
        comm_connected = G == 1;
    }

    for (auto& sl : slot) {
        CU_TRY(this, sl.p_cnt.reserve(1)); CU_TRY(this, sl.p_stamps.reserve(f.NX + f.NY + f.NZ + 1 + MAX_RANKS * COMM_COUNTS));
        CU_TRY(this, sl.p_maskdesc.reserve(64)); CU_TRY(this, sl.p_moves.reserve(1));
        memset(sl.p_cnt.p, 0, sizeof(Counters));
        CU_TRY(this, cudaEventCreateWithFlags(&sl.copy_done, cudaEventDisableTiming));
        CU_TRY(this, cudaEventCreateWithFlags(&sl.up_depth, cudaEventDisableTiming));
        CU_TRY(this, cudaEventCreateWithFlags(&sl.up_masks, cudaEventDisableTiming));
        CU_TRY(this, cudaEventCreateWithFlags(&sl.up_tab, cudaEventDisableTiming));
        CU_TRY(this, cudaEventCreateWithFlags(&sl.frame_done, cudaEventDisableTiming));
        CU_TRY(this, cudaEventCreate(&sl.tr_begin)); CU_TRY(this, cudaEventCreate(&sl.tr_end));
    }
    CU_TRY(this, cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    CU_TRY(this, cudaStreamCreateWithFlags(&upload_stream, cudaStreamNonBlocking));

    // Gaussian PDF LUT (utils/basic_algorithms.h:403-407, 455-459): (1/sqrt(2*M_PI_2f32)) * exp(-v^2/2), v=(i-10000)*0.001
    {
        std::vector<float> lut(PDF_NUM);
        const float half_pi_f32 = 1.57079632679489661923f;
        for (int i = 0; i < PDF_NUM; ++i) {
            const float value = (float)(i - PDF_NUM / 2) * 0.001f;
            lut[i] = (1.f / (sqrtf(2.f * half_pi_f32))) * expf(-(value * value) / (2));
        }
        CU_TRY(this, cudaMemcpy(d_lut.p, lut.data(), PDF_NUM * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (f.win_r > CK_MAX_R) return fail(DSPMAP_ERR_CONFIG, "window_half %d: at most %d", f.win_r, CK_MAX_R);
#define SET_SMEM(K, BYTES) CU_TRY(this, cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)))
    SET_SMEM((k_ck<0, false>), CK_SMEM_BYTES); SET_SMEM((k_ck<3, false>), CK_SMEM_BYTES); SET_SMEM((k_ck<5, false>), CK_SMEM_BYTES);
    SET_SMEM((k_ck<0, true>), CK_SMEM_BYTES); SET_SMEM((k_ck<3, true>), CK_SMEM_BYTES); SET_SMEM((k_ck<5, true>), CK_SMEM_BYTES);
    SET_SMEM((k_weight_update<0, false>), WU_SMEM_BYTES); SET_SMEM((k_weight_update<3, false>), WU_SMEM_BYTES); SET_SMEM((k_weight_update<5, false>), WU_SMEM_BYTES);
    SET_SMEM((k_weight_update<0, true>), WU_SMEM_BYTES); SET_SMEM((k_weight_update<3, true>), WU_SMEM_BYTES); SET_SMEM((k_weight_update<5, true>), WU_SMEM_BYTES);
#undef SET_SMEM

    for (int a = 0; a < 3; ++a) { moved[a] = 0; equiv[a] = 0; map_center[a] = 0.f; ego_center[a] = 0.f; last_pos[a] = 0.f; }
    return clear_device();
}

// RingBufferOperations::clear (mc_ring/operations.h:684-723) + SemanticDSPMap::clear (semantic_dsp_map.h:74-81)
int dspmap::clear_device() {
    reset_host_state();
    ++launches; k_clear_store<<<sm_count * 8, 256, 0, stream>>>(Store{sx.p, sy.p, sz.p, sw.p, smeta.p, shdr.p}, n_slots_local, (uint32_t)fp.Pb);
    ++launches; k_fill_u32<<<sm_count * 4, 256, 0, stream>>>(d_keys.p, d_keys.n, EMPTY_KEY);
    ++launches; k_fill_u32<<<sm_count * 4, 256, 0, stream>>>(d_ranks.p, d_ranks.n, NO_RANK);
    ++launches; k_fill_birth_boxes<<<sm_count * 4, 256, 0, stream>>>(d_birth_box.p, d_birth_box.n);
    CU_TRY(this, cudaMemsetAsync(d_pix_count.p, 0, d_pix_count.n * sizeof(uint32_t), stream));
    CU_TRY(this, cudaMemsetAsync(d_pix_offset.p, 0, d_pix_offset.n * sizeof(uint32_t), stream));
    CU_TRY(this, cudaMemsetAsync(d_cnt.p, 0, sizeof(Counters), stream));
    // the movable-particle list is empty again; the random-table cursors survive a clear, as in the reference
    CU_TRY(this, cudaMemsetAsync(&d_persist.p->n_movable, 0, sizeof(uint32_t), stream));
    CU_TRY(this, cudaMemsetAsync(d_vstate.p, 0, d_vstate.n, stream));
    CU_TRY(this, cudaMemsetAsync(d_blk_cnt.p, 0, (size_t)n_occ_blocks * sizeof(uint4), stream));   // from here on k_occ_scan re-zeroes it every frame
    CU_TRY(this, cudaGetLastError());
    CU_TRY(this, cudaStreamSynchronize(stream));
    memset(&last, 0, sizeof last);
    memset(&stats, 0, sizeof stats);
    for (auto& sl : slot) { memset(sl.p_cnt.p, 0, sizeof(Counters)); sl.host_result_valid = false; sl.inflight = false; sl.copy_issued = false; sl.result_n = sl.result_nf = sl.n_occ_global = 0; sl.T = 0; sl.rc = DSPMAP_OK; }
    n_visible_last = 0;
    return DSPMAP_OK;
}

int dspmap::apply_params() {
    FrameParams& f = fp;
    f.Pd = prm.detection_probability; f.kappa = prm.noise_number; f.tau = prm.id_transition_probability;
    f.occ_thr = prm.occupancy_threshold; f.max_forget = prm.max_forget_count;
    f.depth_noise = prm.consider_depth_noise ? 1 : 0; f.indep = prm.use_independent_filter ? 1 : 0;
    f.n0 = prm.depth_noise_zero_order; f.n1 = prm.depth_noise_first_order;
    f.sigma_invalid = prm.uninit_sigma_mode ? prm.uninit_sigma_value : (prm.consider_depth_noise ? prm.depth_noise_zero_order : 0.1f);
    f.sigma_gated_set = prm.uninit_sigma_mode ? 1 : 0;
    f.ref_order = prm.reference_iteration_order ? 1 : 0;
    f.zero_center = prm.zero_center ? 1 : 0;
    f.col_enabled = colors.enabled ? 1 : 0; f.col_eval = colors.eval_format ? 1 : 0; f.col_scheme = colors.scheme; f.col_background = colors.background_label;
    f.nb_ptc = prm.consider_depth_noise ? prm.nb_ptc_num_per_point : 1;
    // one newborn per point: a table prefix of 2x the pixel count is enough (and cache friendlier); more: the whole table
    f.birth_hash_cap = f.nb_ptc > 1 ? f.hash_cap : birth_cap1; f.birth_hash_shift = f.nb_ptc > 1 ? f.hash_shift : birth_shift1;
    return DSPMAP_OK;
}

// moveParticlesInSetsByTransformations + removeObjectByTrackID (mc_ring/operations.h:321-362; object_layer.h:414-425)
// scan_remove: the movable list no longer describes the store (second phase of a move-then-remove frame)
int dspmap::run_moves(const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove, bool scan_remove, int slot) {
    // one staging table per call of a frame (slot): nothing waits for an earlier table's copy to finish
    MoveTable& mt = this->slot[cur].p_moves.p[slot];
    mt.n_moves = n_moves; mt.n_remove = n_remove;
    for (int i = 0; i < n_moves; ++i) { mt.move_track[i] = moves[i].track_id; memcpy(mt.t[i], moves[i].t, 12 * sizeof(float)); }
    for (int i = 0; i < n_remove; ++i) mt.remove_track[i] = remove_tracks[i];
    MoveTable* d_mt = this->slot[cur].d_moves.p + slot;
    // only the used part of the table travels (the rigid transforms are its tail)
    const size_t mt_bytes = n_moves > 0 ? offsetof(MoveTable, t) + (size_t)n_moves * sizeof(mt.t[0]) : offsetof(MoveTable, t);
    CU_TRY(this, upload_table(this->slot[cur], d_mt, &mt, mt_bytes));
    const uint32_t* stx = this->slot[cur].d_stamps.p;
    const uint32_t* sty = stx + fp.NX;
    const uint32_t* stz = sty + fp.NY;
    if (scan_remove) {
        ++launches; k_remove_tracks_scan<<<sm_count * 8, 256, 0, stream>>>(fp, st, d_mt, d_cnt.p);
    } else {
        if (cv.size > 1) {
            ++launches; k_move_count<<<1, 1024, 0, stream>>>(d_mt, d_movl_track.p, sel_cap, d_persist.p, cv);
            barrier();   // every rank knows how many particles of each moved track the lower ranks hold
        }
        if (n_moves <= 4) {      // the usual frame: numbering over the whole GPU
            const uint32_t nb = (sel_cap + MPB_CHUNK - 1) / MPB_CHUNK;          // <= SCAN_BLOCK blocks (sel_cap <= 2^20)
            ++launches; k_move_count_blocks<<<nb, SCAN_BLOCK, 0, stream>>>(d_mt, d_movl_track.p, sel_cap, d_persist.p, d_move_blk.p);
            ++launches; k_move_number<<<nb, SCAN_BLOCK, 0, stream>>>(d_mt, d_movl_track.p, d_movl_seq.p, sel_cap, rand_n, d_cnt.p, d_persist.p, cv,
                                                                     cv.move_counts[cv.rank], d_move_blk.p);
        } else {
            ++launches; k_move_prepare<<<1, 1024, 0, stream>>>(d_mt, d_movl_track.p, d_movl_seq.p, sel_cap, rand_n, d_cnt.p, d_persist.p, cv,
                                                             cv.move_counts[cv.rank]);
        }
        ++launches; k_move_apply<<<sm_count * 2, 256, 0, stream>>>(fp, d_cnt.p, st, d_mt, d_randoms.p, d_persist.p, d_movl_idx.p, d_movl_track.p,
                                                             d_movl_seq.p, sel_cap, cv);
        if (n_moves > 0) {
            if (cv.size > 1) {
                ++launches; k_move_publish<<<1, 32, 0, stream>>>(cv, d_cnt.p);
                barrier();   // the moved particles bound for this rank have arrived
            }
            ++launches; k_move_bin<<<sm_count * 2, 256, 0, stream>>>(fp, cv, cv.movs[cv.rank], d_counts, d_seq_ref.p, seq_cap, d_keys.p, d_ranks.p,
                                                               d_entry_list.p, d_cnt.p);
            DISPATCH_PB(fp.Pb, (k_move_replay<PB><<<sm_count * 2, 128, 0, stream>>>(fp, st, stx, sty, stz, d_keys.p, d_ranks.p, d_entry_list.p, d_entry_voxel.p, d_cnt.p)));
            ++launches;
            ++launches; k_move_write<<<sm_count * 2, 256, 0, stream>>>(fp, st, cv.movs[cv.rank], d_seq_ref.p, seq_cap, d_ranks.p, d_entry_list.p, d_entry_voxel.p, d_cnt.p);
        }
    }
    CU_TRY(this, cudaGetLastError());
    return DSPMAP_OK;
}

static double host_now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// SemanticDSPMap::update (semantic_dsp_map.h:170-251) -> subObjectLevelUpdate (:576-955)
int dspmap::frame(const float* d_depth, int rows, int cols, int stride, const dspmap_mask* masks, int n_masks, const double pos[3],
                  const double quat[4], const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove,
                  int want_free, int pre_decimated, const std::function<int(dspmap_mask*)>* late_masks, const std::function<int()>* late_depth,
                  bool pipelined, bool want_host) {
    // Synchronous call: enqueue the frame, wait for its counters.  Pipelined call (dspmap_update_async): enqueue frame t, THEN
    // wait for frame t-1 - which ran while the host staged and enqueued this one - and start the transfer of its result.
    const int prev = cur;
    if (!pipelined || profiling) { const int rc = drain(false); if (rc != DSPMAP_OK) return rc; }
    if (slot[next_slot()].inflight) { const int rc = frame_finish(next_slot()); if (rc != DSPMAP_OK) return rc; }   // (protocol: cannot happen)
    int rc = frame_enqueue(d_depth, rows, cols, stride, masks, n_masks, pos, quat, moves, n_moves, remove_tracks, n_remove, want_free, pre_decimated, late_masks, late_depth);
    if (rc != DSPMAP_OK) return rc;
    slot[cur].want_host = want_host;
    if (!pipelined || profiling) return frame_finish(cur);
    if (prev != cur && slot[prev].inflight) {
        static const bool trace = getenv("DSPMAP_TRACE") != nullptr;
        const double ta = trace ? host_now_ms() : 0.0;
        rc = frame_finish(prev);
        const double tb = trace ? host_now_ms() : 0.0;
        if (rc == DSPMAP_OK && slot[prev].want_host) rc = issue_result_copy(prev, copy_stream);
        if (trace) {
            trace_prev_ms += tb - ta; trace_copy_issue_ms += host_now_ms() - tb;
            if ((T % 16) == 0) {
                fprintf(stderr, "[dspmap] pipelined, mean of 16: enqueue %.3f ms, then waited %.3f ms for the previous frame, result copy issue %.3f ms, wait_result blocked %.3f ms\n",
                        trace_issue_ms / 16, trace_prev_ms / 16, trace_copy_issue_ms / 16, trace_result_ms / 16);
                trace_issue_ms = trace_prev_ms = trace_copy_issue_ms = trace_result_ms = trace_wait_ms = 0;
            }
        }
        if (rc != DSPMAP_OK) { err = "frame " + std::to_string(slot[prev].T) + " (reported by the call that enqueued frame " + std::to_string(T) + "): " + err; return rc; }
    }
    return DSPMAP_OK;
}

// Waits for the frames in flight.  A frame has "finished" for the host when its counters have arrived, which is before its emit
// pass has written the last result entry: whole_frame also waits for that (callers that touch device memory outside the stream).
int dspmap::drain(bool whole_frame) {
    for (int k = 0; k < 2; ++k) {
        const int s = (cur + 1 + k) & 1;      // older frame first
        if (slot[s].inflight) { const int rc = frame_finish(s); if (rc != DSPMAP_OK) return rc; }
    }
    if (whole_frame && T > 0) CU_TRY(this, cudaEventSynchronize(slot[cur].frame_done));
    return DSPMAP_OK;
}

int dspmap::frame_enqueue(const float* d_depth, int rows, int cols, int stride, const dspmap_mask* masks, int n_masks, const double pos[3],
                          const double quat[4], const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove,
                          int want_free, int pre_decimated, const std::function<int(dspmap_mask*)>* late_masks, const std::function<int()>* late_depth) {
    // late_depth: the depth image is not on the device yet; the callback (host staging + H2D enqueue) runs after the vertex
    // mask and the object moves - which need the pose only - have been launched, so the GPU works while the host stages.
    // late_masks: the mask images are not on the device yet; the callback (host staging + H2D enqueue) runs after the
    // mask-independent kernels have been launched, so the copy engine and the CPU work under them
    FrameParams& f = fp;
    f.pre_decimated = pre_decimated;
    // the noise path spawns nb particles per point only under the depth-noise model; addNewbornParticleAndResample (no depth
    // noise) always adds one (semantic_dsp_map.h:1148-1171)
    const int nb_eff = prm.consider_depth_noise ? prm.nb_ptc_num_per_point : 1;
    {
        const long long n_cand = 9ll * ((fp.H + 2) / 3) * ((fp.W + 2) / 3) * nb_eff;
        if (n_cand >= (1ll << 24)) return fail(DSPMAP_ERR_UNSUPPORTED, "nb_ptc_num_per_point = %d gives %lld birth candidates per frame (limit 2^24)", nb_eff, n_cand);
    }
    // everything that can refuse the frame is checked BEFORE the time stamp, the ring or the store are touched
    if (n_moves > MAX_MOVES) return fail(DSPMAP_ERR_CAPACITY, "%d object moves in one frame (limit %d); nothing was changed", n_moves, MAX_MOVES);
    if (!comm_connected) return fail(DSPMAP_ERR_COMM, "sharded map: call dspmap_comm_connect on every rank before the first update");
    t_frame_begin = host_now_ms();
    T += 1;                                                       // :173
    f.T = T;
    cur = (int)(T & 1u);
    FrameSlot& sl = slot[cur];
    sl.T = T; sl.host_result_valid = false; sl.want_host = false; sl.rc = DSPMAP_OK;
    sl.want_free = want_free ? 1 : 0;
    // the slot's result arrays may still be on their way to the host (frame T-2, pipelined): the emit pass waits for that copy
    const bool wait_copy = sl.copy_issued;
    sl.copy_issued = false;
    apply_params();
    if (!forgetting_init) {                                       // utils/basic_algorithms.h:32-48, built once
        for (int i = 0; i < 5; ++i) forget_tab[i] = (float)pow(2.5, -i / prm.forgetting_rate);
        forgetting_init = true;
    }
    memcpy(f.forget_tab, forget_tab, sizeof forget_tab);

    // camera matrices: double for the projection (pointcloud_tools.h:105-110), float for the map (:744-747)
    quat_to_rot<double>(quat, f.R);
    for (int a = 0; a < 3; ++a) f.t[a] = pos[a];
    const float posf[3] = {(float)pos[0], (float)pos[1], (float)pos[2]};
    const float qf[4] = {(float)quat[0], (float)quat[1], (float)quat[2], (float)quat[3]};
    float Rf[9];
    quat_to_rot<float>(qf, Rf);
    const float c2g[16] = {Rf[0], Rf[1], Rf[2], posf[0], Rf[3], Rf[4], Rf[5], posf[1], Rf[6], Rf[7], Rf[8], posf[2], 0.f, 0.f, 0.f, 1.f};
    inverse4f(c2g, f.ext);
    float inv_ext[16];
    inverse4f(f.ext, inv_ext);
    f.cam[0] = inv_ext[3]; f.cam[1] = inv_ext[7]; f.cam[2] = inv_ext[11];   // :1241

    ring_shift(posf);                                             // :584-585
    for (int a = 0; a < 3; ++a) { f.equiv[a] = equiv[a]; f.map_center[a] = map_center[a]; f.vox_off[a] = map_center[a] + f.pmin[a]; }

    // BFS start vertex (operations.h:1312-1338), clamped into the lattice
    {
        const float sg[3] = {(inv_ext[0] * 0.f + inv_ext[1] * 0.f + inv_ext[2] * 1.f) + inv_ext[3],
                             (inv_ext[4] * 0.f + inv_ext[5] * 0.f + inv_ext[6] * 1.f) + inv_ext[7],
                             (inv_ext[8] * 0.f + inv_ext[9] * 0.f + inv_ext[10] * 1.f) + inv_ext[11]};
        int sv[3];
        for (int a = 0; a < 3; ++a) {
            sv[a] = static_cast<int>(((sg[a] - map_center[a]) + f.pmax[a]) * f.voxel_recip);
            sv[a] = std::min(std::max(sv[a], 0), (int)axis_n(a));
            f.start_v[a] = sv[a];
        }
        const float vx = (float)sv[0] * f.voxel + f.vox_off[0], vy = (float)sv[1] * f.voxel + f.vox_off[1], vz = (float)sv[2] * f.voxel + f.vox_off[2];
        const float c0 = ((f.ext[0] * vx + f.ext[1] * vy) + f.ext[2] * vz) + f.ext[3] * 1.0f;
        const float c1 = ((f.ext[4] * vx + f.ext[5] * vy) + f.ext[6] * vz) + f.ext[7] * 1.0f;
        const float c2 = ((f.ext[8] * vx + f.ext[9] * vy) + f.ext[10] * vz) + f.ext[11] * 1.0f;
        f.start_in = !(c2 < f.depth_min || c2 > f.depth_max) && !(std::abs(c0) > c2 * f.tan_hx) && !(std::abs(c1) > c2 * f.tan_hy);
    }
    // conservative bounding box of the frustum in lattice-vertex indices: its 8 corners in the map frame, +-2 vertices
    {
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int k = 0; k < 8; ++k) {
            const float z = (k & 1) ? f.depth_max : f.depth_min;
            const float x = ((k & 2) ? 1.f : -1.f) * z * f.tan_hx, y = ((k & 4) ? 1.f : -1.f) * z * f.tan_hy;
            for (int a = 0; a < 3; ++a) {
                const float g = inv_ext[a * 4 + 0] * x + inv_ext[a * 4 + 1] * y + inv_ext[a * 4 + 2] * z + inv_ext[a * 4 + 3];
                const float vi = (g - f.vox_off[a]) * f.voxel_recip;
                lo[a] = std::min(lo[a], vi); hi[a] = std::max(hi[a], vi);
            }
        }
        for (int a = 0; a < 3; ++a) {
            f.vb_lo[a] = std::max(0, (int)std::floor(lo[a]) - 2);
            f.vb_hi[a] = std::min((int)axis_n(a), (int)std::ceil(hi[a]) + 2);
        }
    }
    f.n_masks = n_masks;
    f.max_movable = max_movable; f.sky_track = sky_track;
    f.want_free = want_free ? 1 : 0; f.want_state = 0;
    f.rand_n = rand_n;

    // ---- uploads -------------------------------------------------------------------------------------------
    if (sl.stamps_stale) {
        uint32_t* s = sl.p_stamps.p;
        for (int a = 0; a < 3; ++a) { memcpy(s, stamps[a].data(), stamps[a].size() * 4); s += stamps[a].size(); }
        CU_TRY(this, upload_table(sl, sl.d_stamps.p, sl.p_stamps.p, (f.NX + f.NY + f.NZ) * 4));
        sl.stamps_stale = false;
    }
    // (the counters are zero: k_occ_scan - sharded maps: k_publish_counters - hands them to the host and clears them)
    static const bool trace_gpu = getenv("DSPMAP_TRACE") != nullptr;
    if (trace_gpu) cudaEventRecord(sl.tr_begin, stream);
    const uint32_t* stx = sl.d_stamps.p;
    const uint32_t* sty = stx + f.NX;
    const uint32_t* stz = sty + f.NY;
    const int npix = f.W * f.H;

    if (profiling) cudaEventRecord(ev[0], stream);
    const bool depth_kernel = !((pre_decimated || !f.boost) && stride == f.W);
    if (late_depth && depth_kernel) { const int rc = (*late_depth)(); if (rc != DSPMAP_OK) return rc; late_depth = nullptr; }
    if (f.ref_order) {
        if (cv.size > 1) return fail(DSPMAP_ERR_UNSUPPORTED, "reference_iteration_order is not available for sharded maps");
        if (!d_vrank.p) {
            const size_t n_vert = (size_t)(f.NX + 1) * (f.NY + 1) * (f.NZ + 1);
            CU_TRY(this, cudaStreamSynchronize(stream));
            CU_TRY(this, d_vrank.alloc(n_vert)); CU_TRY(this, cudaMemset(d_vrank.p, 0xff, n_vert * 4));
            const size_t fcap = std::min<size_t>(n_vert, (size_t)1 << 22);
            CU_TRY(this, d_bfs_front[0].alloc(fcap)); CU_TRY(this, d_bfs_front[1].alloc(fcap));
            CU_TRY(this, d_work_key.alloc(d_work.n));
            int per_sm = 0;
            CU_TRY(this, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bfs_order, 1024, 0));
            if (per_sm < 1) return fail(DSPMAP_ERR_CUDA, "k_bfs_order does not fit an SM");
            bfs_grid = sm_count;
            CU_TRY(this, d_bfs_blocks.alloc(bfs_grid + 2));
        }
    }
    // ---- U2a vertex mask: depends on the pose only, so it runs on the side stream beside the depth copy and the moves ------
    if (f.start_in) {
        const int y0 = std::max(f.vb_lo[1] - 1, 0), y1 = std::min(f.vb_hi[1] + 1, (int)f.NY);
        const int z0 = std::max(f.vb_lo[2] - 1, 0), z1 = std::min(f.vb_hi[2] + 1, (int)f.NZ);
        CU_TRY(this, cudaEventRecord(ev_fork[0], stream));
        CU_TRY(this, cudaStreamWaitEvent(side_stream, ev_fork[0], 0));
        ++launches; k_vertex_mask<<<dim3((words_per_row * 32 + 255) / 256, (y1 - y0 + 1) * (z1 - z0 + 1)), 256, 0, side_stream>>>(fp, d_vmask.p, words_per_row, y0, y1 - y0 + 1, z0, f.ref_order ? d_vrank.p : nullptr);
        CU_TRY(this, cudaEventRecord(ev_join[0], side_stream));
    }
    // ---- K1a internal depth image ----------------------------------------------------------------------------------
    const float* depth_int = d_depth_int.p;
    if ((pre_decimated || !f.boost) && stride == f.W) depth_int = d_depth;      // already the internal image: no copy
    else { ++launches; k_depth_int<<<(npix + 255) / 256, 256, 0, stream>>>(fp, d_depth, stride, d_depth_int.p); }
    if (profiling) cudaEventRecord(ev[1], stream);
    // ---- P2/P3 object predict / prune ---------------------------------------------------------------------------
    if (n_moves > 0 || n_remove > 0) {
        // a track that is both moved and removed is removed AFTER its move (the reference order)
        std::vector<int32_t> rm_now, rm_after;
        for (int i = 0; i < n_remove; ++i) {
            bool also_moved = false;
            for (int k = 0; k < n_moves; ++k) also_moved |= moves[k].track_id == remove_tracks[i];
            (also_moved ? rm_after : rm_now).push_back(remove_tracks[i]);
        }
        // the reference has no limit on removals per frame (every unregistered detection is one): they are independent, so
        // they go MAX_MOVES at a time, the first batch together with the moves
        const int nb_now = std::max(1, ((int)rm_now.size() + MAX_MOVES - 1) / MAX_MOVES), nb_after = ((int)rm_after.size() + MAX_MOVES - 1) / MAX_MOVES;
        CU_TRY(this, sl.p_moves.reserve(nb_now + nb_after));
        if (sl.d_moves.n < (size_t)(nb_now + nb_after)) { CU_TRY(this, cudaStreamSynchronize(stream)); CU_TRY(this, sl.d_moves.alloc(nb_now + nb_after)); }
        int slot = 0;
        for (int b = 0; b < nb_now; ++b, ++slot) {
            const int lo = b * MAX_MOVES, n = std::min((int)rm_now.size() - lo, MAX_MOVES);
            const int rc = run_moves(b == 0 ? moves : nullptr, b == 0 ? n_moves : 0, rm_now.data() + lo, std::max(n, 0), false, slot);
            if (rc != DSPMAP_OK) return rc;
        }
        for (int b = 0; b < nb_after; ++b, ++slot) {
            const int lo = b * MAX_MOVES, n = std::min((int)rm_after.size() - lo, MAX_MOVES);
            const int rc = run_moves(nullptr, 0, rm_after.data() + lo, n, true, slot);
            if (rc != DSPMAP_OK) return rc;
        }
    }
    if (late_depth) { const int rc = (*late_depth)(); if (rc != DSPMAP_OK) return rc; }
    if (profiling) cudaEventRecord(ev[2], stream);
    // ---- U2 visibility ---------------------------------------------------------------------------------------------
    if (f.start_in) {
        CU_TRY(this, cudaStreamWaitEvent(stream, ev_join[0], 0));   // the vertex mask was computed on the side stream
        {   // voxels whose vertices m..m+1 can intersect the vertex box [vb_lo, vb_hi]
            const int3 lo = make_int3(std::max(f.vb_lo[0] - 1, 0), std::max(f.vb_lo[1] - 1, 0), std::max(f.vb_lo[2] - 1, 0));
            const int3 hi = make_int3(std::min(f.vb_hi[0], (int)f.NX - 1), std::min(f.vb_hi[1], (int)f.NY - 1), std::min(f.vb_hi[2], (int)f.NZ - 1));
            const int3 n = make_int3(hi.x - lo.x + 1, hi.y - lo.y + 1, hi.z - lo.z + 1);
            if (n.x > 0 && n.y > 0 && n.z > 0) {
                if (f.ref_order) {      // the order the reference's search reaches the vertices in (a cooperative launch: grid-wide barriers)
                    CU_TRY(this, cudaMemsetAsync(d_bfs_blocks.p, 0, (bfs_grid + 2) * sizeof(uint32_t), stream));
                    BfsScratch bs{d_vrank.p, {d_bfs_front[0].p, d_bfs_front[1].p}, d_bfs_blocks.p, (uint32_t)d_bfs_front[0].n, &d_cnt.p->hash_overflow};
                    const uint32_t* vm = d_vmask.p;
                    uint32_t wpr = words_per_row;
                    void* args[] = {(void*)&fp, (void*)&vm, (void*)&wpr, (void*)&bs};
                    ++launches; CU_TRY(this, cudaLaunchCooperativeKernel((const void*)k_bfs_order, dim3(bfs_grid), dim3(1024), args, 0, stream));
                }
                ++launches; k_sweep_select<<<dim3((n.x + 255) / 256, n.y * n.z), 256, 0, stream>>>(fp, st, d_vmask.p, words_per_row, depth_int, d_work.p, d_work_hdr.p, d_cnt.p, lo, n,
                                                                                       f.ref_order ? d_vrank.p : nullptr, f.ref_order ? d_work_key.p : nullptr);
            }
        }
        DISPATCH_PB(f.Pb, (k_sweep_slots<PB><<<sm_count * 8, 256, 0, stream>>>(fp, st, stx, sty, stz, d_work.p, d_work_hdr.p, depth_int, d_pix_count.p, cv, d_cnt.p, f.ref_order ? d_work_key.p : nullptr)));
        ++launches;
    }
    // every rank tells every rank how many visible records it wrote into their arenas
    if (cv.size > 1) {
        ++launches; k_publish_pairs<<<1, 32, 0, stream>>>(cv, d_cnt.p);
        barrier();
        ++launches; k_count_remote<<<sm_count * 2, 256, 0, stream>>>(cv, cv.recs[cv.rank], d_counts, d_pix_count.p);
    }
    const int nsb = (npix + SCAN_BLOCK - 1) / SCAN_BLOCK;
    ++launches; k_scan_blocksum<<<nsb, SCAN_BLOCK, 0, stream>>>(d_pix_count.p, (uint32_t)npix, d_scan_block.p);
    ++launches; k_scan_apply<<<nsb, SCAN_BLOCK, 0, stream>>>(d_pix_count.p, (uint32_t)npix, d_scan_block.p, d_pix_offset.p, &d_cnt.p->n_visible);
    // the order in which pass 1 of the stencil takes the image rows needs the offsets only: beside the list building, on the side stream
    CU_TRY(this, cudaEventRecord(ev_fork[2], stream));
    CU_TRY(this, cudaStreamWaitEvent(side_stream, ev_fork[2], 0));
    ++launches; k_ck_row_order<<<1, 1024, (size_t)f.H * sizeof(uint32_t), side_stream>>>(fp, d_pix_offset.p, d_row_order.p);
    CU_TRY(this, cudaEventRecord(ev_join[2], side_stream));
    ++launches; k_fill<<<sm_count * 4, 256, 0, stream>>>(cv, cv.recs[cv.rank], d_counts, d_cnt.p, d_pix_offset.p, d_pix_count.p, d_csr_key.p);
    ++launches; k_sort_lists<<<(npix + 255) / 256, 256, 0, stream>>>(fp, d_pix_offset.p, d_csr_key.p);
    ++launches; k_gather_visible<<<sm_count * 4, 256, 0, stream>>>(d_pix_offset.p + npix, cv.recs[cv.rank], d_csr_key.p, d_csr_idx.p, d_vis_pos.p, d_vis_mp.p);
    if (profiling) cudaEventRecord(ev[3], stream);
    // ---- K1 projection + labels (needs the masks) -----------------------------------------------------------------------
    std::vector<dspmap_mask> mask_copy(masks, masks + n_masks);
    if (late_masks) { const int rc = (*late_masks)(mask_copy.data()); if (rc != DSPMAP_OK) return rc; }
    if ((size_t)n_masks > sl.d_masks.n) { CU_TRY(this, cudaStreamSynchronize(stream)); CU_TRY(this, sl.d_masks.alloc(n_masks)); }
    CU_TRY(this, sl.p_maskdesc.reserve(std::max(n_masks, 1)));
    for (int m = 0; m < n_masks; ++m) {
        MaskDesc& md = sl.p_maskdesc.p[m];
        md.data = mask_copy[m].data; md.row_stride = mask_copy[m].row_stride; md.is_static = mask_copy[m].is_static;
        md.track_id = mask_copy[m].track_id; md.label_id = mask_copy[m].label_id;
        for (int a = 0; a < 3; ++a) { md.bbox_min[a] = mask_copy[m].bbox_min[a]; md.bbox_max[a] = mask_copy[m].bbox_max[a]; }
    }
    if (n_masks) CU_TRY(this, upload_table(sl, sl.d_masks.p, sl.p_maskdesc.p, n_masks * sizeof(MaskDesc)));
    ++launches; k_project<<<(npix + 255) / 256, 256, 0, stream>>>(fp, sl.d_masks.p, d_static_track.p, d_static_label.p,
                                                     depth_int, d_cloud_pos.p, d_cloud_meta.p, d_cnt.p);
    if (profiling) cudaEventRecord(ev[4], stream);
    // ---- B1a birth binning: needs only the labelled points, so it runs on the side stream beside the two stencil passes ----
    {
        const int n_ranks = 9 * ((f.H + 2) / 3) * ((f.W + 2) / 3);
        CU_TRY(this, cudaEventRecord(ev_fork[1], stream));
        CU_TRY(this, cudaStreamWaitEvent(side_stream, ev_fork[1], 0));
        if (f.nb_ptc > 1) {     // number the valid pixels in birth order: they own consecutive draws of the random table
            if (d_birth_flags.n < (size_t)n_ranks + 1) { CU_TRY(this, d_birth_flags.alloc(n_ranks + 1)); CU_TRY(this, d_birth_prefix.alloc(n_ranks + 1)); CU_TRY(this, d_birth_block.alloc((n_ranks + SCAN_BLOCK - 1) / SCAN_BLOCK)); }
            const int nb_blocks = (n_ranks + SCAN_BLOCK - 1) / SCAN_BLOCK;
            ++launches; k_birth_flags<<<(n_ranks + 255) / 256, 256, 0, side_stream>>>(fp, d_cloud_meta.p, d_birth_flags.p, (uint32_t)n_ranks);
            ++launches; k_scan_blocksum<<<nb_blocks, SCAN_BLOCK, 0, side_stream>>>(d_birth_flags.p, (uint32_t)n_ranks, d_birth_block.p);
            ++launches; k_scan_apply<<<nb_blocks, SCAN_BLOCK, 0, side_stream>>>(d_birth_flags.p, (uint32_t)n_ranks, d_birth_block.p, d_birth_prefix.p, nullptr);
        }
        if (f.nb_ptc == 1) {    // one newborn per point: mark the touched voxels, then find every voxel's first candidates in the image
            ++launches; k_birth_mark<<<(npix + 255) / 256, 256, 0, side_stream>>>(fp, d_cloud_pos.p, d_cloud_meta.p, d_keys.p, d_birth_box.p, d_ranks.p, d_entry_list.p, d_cnt.p);
            ++launches; k_birth_gather<<<sm_count * 8, 256, 0, side_stream>>>(fp, d_cloud_pos.p, d_cloud_meta.p, d_keys.p, d_birth_box.p, d_ranks.p, d_entry_list.p, d_cnt.p, d_birth_big.p);
            ++launches; k_birth_gather_big<<<sm_count * 4, 256, 0, side_stream>>>(fp, d_cloud_pos.p, d_cloud_meta.p, d_keys.p, d_ranks.p, d_birth_big.p, d_cnt.p);
        } else {
            ++launches; k_birth_bin<<<(n_ranks * f.nb_ptc + 255) / 256, 256, 0, side_stream>>>(fp, d_cloud_pos.p, d_cloud_meta.p, d_keys.p, d_ranks.p, d_entry_list.p, d_cnt.p,
                                                                                        d_birth_prefix.p, d_randoms.p, d_persist.p);
        }
        CU_TRY(this, cudaEventRecord(ev_join[1], side_stream));
    }
    // ---- U3 weight update --------------------------------------------------------------------------------------------
#define DISPATCH_R1(r, CALL)                                    \
    switch (r) {                                                \
        case 3: { constexpr int RT = 3; CALL; break; }          \
        case 5: { constexpr int RT = 5; CALL; break; }          \
        default: { constexpr int RT = 0; CALL; break; }         \
    }
#define DISPATCH_R(r, CALL)                                     \
    if (f.indep) { constexpr bool INDEP = true; DISPATCH_R1(r, CALL) } else { constexpr bool INDEP = false; DISPATCH_R1(r, CALL) }
    CU_TRY(this, cudaStreamWaitEvent(stream, ev_join[2], 0));   // the row order was computed on the side stream
    DISPATCH_R(f.win_r, (k_ck<RT, INDEP><<<sm_count, CK_WARPS * 32, CK_SMEM_BYTES, stream>>>(fp, d_lut.p, d_cloud_pos.p, d_cloud_meta.p, d_pix_offset.p, d_vis_pos.p, d_vis_mp.p, d_row_order.p, cv, d_cnt.p)));
    ++launches;
    if (cv.size > 1) { ++launches; k_select_mine<<<sm_count * 2, 256, 0, stream>>>(fp, d_pix_offset.p + npix, d_csr_idx.p, d_mine_list.p, d_cnt.p); }
    barrier();   // sharded map: every rank's ck image is complete
    DISPATCH_R(f.win_r, (k_weight_update<RT, INDEP><<<sm_count, WU_THREADS, WU_SMEM_BYTES, stream>>>(fp, d_pix_offset.p + npix, d_lut.p, d_cloud_pos.p, d_obs_a, d_obs_b, d_csr_idx.p,
                                                                                  d_vis_pos.p, d_vis_mp.p, st, d_cnt.p, cv.size > 1 ? d_mine_list.p : nullptr)));
    ++launches;
    if (profiling) cudaEventRecord(ev[5], stream);
    // ---- B1 birth + resample -----------------------------------------------------------------------------------------
    {
        CU_TRY(this, cudaStreamWaitEvent(stream, ev_join[1], 0));   // the candidates were binned on the side stream
        DISPATCH_PB(f.Pb, (k_birth_replay<PB><<<sm_count * 4, 128, 0, stream>>>(fp, st, stx, sty, stz, d_keys.p, d_ranks.p, d_entry_list.p, d_entry_voxel.p, d_cnt.p)));
        ++launches;
        ++launches; k_birth_write<<<sm_count * 8, 256, 0, stream>>>(fp, st, d_cloud_pos.p, d_cloud_meta.p, d_ranks.p, d_entry_list.p, d_entry_voxel.p, d_cnt.p,
                                                                d_birth_prefix.p, d_randoms.p, d_persist.p);
        if (f.nb_ptc > 1) { ++launches; k_birth_advance_cursor<<<1, 1, 0, stream>>>(fp, d_cnt.p, d_persist.p); }
    }
    if (profiling) cudaEventRecord(ev[6], stream);
    // ---- O1 occupancy -------------------------------------------------------------------------------------------------
    const ResultArrays out = sl.dres;
    OccScratch osc{d_vstate.p, d_res_w.p - (slot_lo >> f.Pb), d_res_tl.p - (slot_lo >> f.Pb), d_mov_mask.p - (slot_lo >> f.Pb), d_blk_cnt.p, d_blk_off.p, d_chunk_tot.p, d_chunk_base.p, d_tile_list.p};
    if (profiling) cudaEventRecord(ev[8], stream);
    // DSPMAP_OCC_HEADERS=bulk: the header plane staged through shared memory by the bulk-copy unit (k_occ_headers_bulk).  Measured
    // at configs B / D / E: 18.4 / 63.1 / 137.9 us against 17.1 / 57.7 / 134.7 us for a 16-byte load per thread with one prefetch -
    // the pass already runs at the copy peak on the margin (504 MB more from D to E in 77 us = 6.5 TB/s); what is left at the
    // smaller sizes is start-up, not bytes in flight.  Kept for comparisons, not the default.
    static const bool headers_by_loads = !(getenv("DSPMAP_OCC_HEADERS") && !strcmp(getenv("DSPMAP_OCC_HEADERS"), "bulk"));
    if (headers_by_loads) { DISPATCH_PB(f.Pb, (k_occ_headers<PB><<<sm_count * 8, OCC_BLOCK, 0, stream>>>(fp, st, stx, sty, stz, osc, d_work.p, n_occ_blocks, d_cnt.p))); }
    else { DISPATCH_PB(f.Pb, (k_occ_headers_bulk<PB><<<sm_count * 4, OHB_BLOCK, 0, stream>>>(fp, st, stx, sty, stz, osc, d_work.p, n_occ_blocks, d_cnt.p))); }
    ++launches;
    if (profiling) cudaEventRecord(ev[9], stream);
    DISPATCH_PB(f.Pb, (k_occ_slots<PB><<<sm_count * 8, 256, 0, stream>>>(fp, st, stx, sty, stz, osc, d_work.p, d_cnt.p)));
    ++launches;
    ++launches; k_occ_scan<<<n_occ_chunks, 1024, 0, stream>>>(osc, n_occ_blocks, d_cnt.p, d_persist.p, cv, T, cv.size == 1 ? sl.p_cnt.p : nullptr);
    // before the barrier: on a sharded map the PEERS' emit passes write into rank 0's arrays too, and they pass this barrier only
    // when rank 0 has
    if (wait_copy) CU_TRY(this, cudaStreamWaitEvent(stream, sl.copy_done, 0));
    barrier();   // the lower ranks' totals are this rank's base offsets
    {
        const int use_list = n_occ_chunks > 1;
        const uint32_t emit_grid = use_list ? std::min<uint32_t>(n_occ_blocks, (uint32_t)sm_count * 8) : n_occ_blocks;
        DISPATCH_PB(f.Pb, (k_occ_emit<PB><<<emit_grid, OCC_BLOCK, 0, stream>>>(fp, st, osc, out, d_movl_idx.p, d_movl_track.p, sel_cap, cv, d_counts, d_col_tab.p, n_occ_chunks, use_list)));
    }
    ++launches;
    barrier();   // rank 0 holds the whole result
    if (profiling) cudaEventRecord(ev[7], stream);
    CU_TRY(this, cudaEventRecord(sl.frame_done, stream));      // the result arrays are complete
    if (cv.size > 1) {
        ++launches; k_publish_counters<<<1, 64, 0, stream>>>(d_cnt.p, d_counts, sl.p_cnt.p, sl.p_stamps.p + (f.NX + f.NY + f.NZ) + 1, (uint32_t)(MAX_RANKS * COMM_COUNTS), T);
    }
    if (trace_gpu) cudaEventRecord(sl.tr_end, stream);
    CU_TRY(this, cudaGetLastError());
    static const bool trace_frame = getenv("DSPMAP_TRACE") != nullptr;
    sl.t_issued = trace_frame ? host_now_ms() : 0.0;
    sl.col_enabled = f.col_enabled;
    sl.inflight = true;
    if (trace_frame) trace_issue_ms += sl.t_issued - t_frame_begin;
    return DSPMAP_OK;
}

// the end of a frame on the host: wait for its counter block, derive the statistics and the result size, report overflows
int dspmap::frame_finish(int s) {
    FrameSlot& sl = slot[s];
    const FrameParams& f = fp;
    const uint32_t T = sl.T;
    static const bool trace_frame = getenv("DSPMAP_TRACE") != nullptr;
    sl.inflight = false;
    {
        // The counter block arrives in the handle's pinned copy with the frame's stamp at both ends (written by k_occ_scan, on a
        // sharded map by k_publish_counters): the host polls it instead of going through a stream synchronisation - the wake-up
        // is a cache miss, not a driver call.  The stream is looked at now and then so that a failed launch ends the wait.
        volatile Counters* pc = sl.p_cnt.p;
        for (uint32_t spins = 0; pc->seq_head != T || pc->seq_tail != T; ++spins) {
            host_pause();
            if ((spins & 0x3fffu) == 0x3fffu && cudaStreamQuery(stream) != cudaErrorNotReady) break;
        }
        if (pc->seq_head != T || pc->seq_tail != T) CU_TRY(this, cudaStreamSynchronize(stream));
    }
    if (trace_frame) {   // host time to issue the frame's launches vs time left waiting for the device afterwards
        trace_wait_ms += host_now_ms() - sl.t_issued;
        if ((T % 16) == 0 && s == cur) { fprintf(stderr, "[dspmap] frame %u: issue %.3f ms, wait %.3f ms (mean of 16)\n", T, trace_issue_ms / 16, trace_wait_ms / 16); trace_issue_ms = trace_wait_ms = 0; }
    }
    last = *sl.p_cnt.p;
    if (trace_frame && T > 2) {
        float gpu_ms = 0.f, gap_ms = 0.f;
        if (cudaEventSynchronize(sl.tr_end) == cudaSuccess && cudaEventElapsedTime(&gpu_ms, sl.tr_begin, sl.tr_end) == cudaSuccess) trace_gpu_ms += gpu_ms;
        if (cudaEventElapsedTime(&gap_ms, slot[s ^ 1].tr_end, sl.tr_begin) == cudaSuccess) trace_gap_ms += gap_ms;
        if ((T % 16) == 0) { fprintf(stderr, "[dspmap] device, mean of 16: frame %.3f ms first to last operation, %.3f ms between the previous frame's last and this frame's first\n", trace_gpu_ms / 16, trace_gap_ms / 16); trace_gpu_ms = trace_gap_ms = 0; }
    }

    stats.time_stamp = T;
    stats.n_swept_voxels = last.n_swept_voxels; stats.n_swept_nonempty = last.n_swept_nonempty;
    n_visible_last = last.n_visible;
    {   // global result size (rank 0 reads the whole list; the other ranks report their own share only)
        const uint32_t* cc = sl.p_stamps.p + (f.NX + f.NY + f.NZ) + 1;
        uint32_t occ = 0, fre = 0;
        if (cv.size == 1) { occ = last.n_occupied; fre = last.n_free; }
        else for (int g = 0; g < cv.size; ++g) { occ += cc[g * COMM_COUNTS + CC_OCC]; fre += cc[g * COMM_COUNTS + CC_FREE]; }
        sl.result_n = cv.rank == 0 ? occ : 0; sl.result_nf = cv.rank == 0 ? fre : 0;
        sl.n_occ_global = occ;
    }
    if (last.comm_errors) return sl.rc = fail(DSPMAP_ERR_COMM, "a peer did not reach a barrier within the timeout");
    stats.n_visible = n_visible_last; stats.n_valid_points = last.n_valid_points;
    stats.n_born = last.n_born; stats.n_birth_voxels = last.n_birth_voxels; stats.n_resampled = last.n_resampled;
    stats.n_occupied = sl.n_occ_global; stats.n_nonempty_voxels = last.n_nonempty_voxels;
    stats.n_moved = last.n_moved; stats.n_removed = last.n_removed;
    if (profiling) {
        cudaEventSynchronize(ev[7]);      // the counters arrive before the emit pass has ended
        // events: 0 start | 1 depth | 2 moves | 3 visibility | 4 project | 5 update | 6 birth | 7 occupancy
        float d01, d12, d23, d34, d45, d56, d67;
        cudaEventElapsedTime(&d01, ev[0], ev[1]); cudaEventElapsedTime(&d12, ev[1], ev[2]); cudaEventElapsedTime(&d23, ev[2], ev[3]);
        cudaEventElapsedTime(&d34, ev[3], ev[4]); cudaEventElapsedTime(&d45, ev[4], ev[5]); cudaEventElapsedTime(&d56, ev[5], ev[6]);
        cudaEventElapsedTime(&d67, ev[6], ev[7]);
        stats.ms_stage[0] = d01 + d34; cudaEventElapsedTime(&stats.ms_stage[1], ev[8], ev[9]);   // [1]: the header-streaming kernel alone
        stats.ms_stage[2] = d12; stats.ms_stage[3] = d23;
        stats.ms_stage[4] = d45; stats.ms_stage[5] = d56; stats.ms_stage[6] = d67;
        cudaEventElapsedTime(&stats.ms_stage[7], ev[0], ev[7]);
    }
    if (last.n_pairs > f.vis_cap) return sl.rc = fail(DSPMAP_ERR_CAPACITY, "visible-particle list overflow: %u > max_visible %u", last.n_pairs, f.vis_cap);
    if (last.n_movable > sel_cap) return sl.rc = fail(DSPMAP_ERR_CAPACITY, "movable-particle list overflow: %u > %u", last.n_movable, sel_cap);
    if (last.hash_overflow) return sl.rc = fail(DSPMAP_ERR_CAPACITY, "insertion hash table overflow");
    return DSPMAP_OK;
}

// enqueues the device->host transfer of slot s's result (its frame has finished) on `on` and records the slot's copy_done event
int dspmap::issue_result_copy(int s, cudaStream_t on) {
    FrameSlot& sl = slot[s];
    if (sl.copy_issued || sl.host_result_valid) return DSPMAP_OK;
    // pinned result buffers grow in big steps (cudaMallocHost takes milliseconds: not something to do every few frames)
    const size_t grain = (size_t)1 << 18;
    const size_t n = sl.result_n, nf = sl.want_free ? sl.result_nf : 0;
    const size_t n_cap = (std::max<size_t>(n, 1) + grain - 1) / grain * grain;
    CU_TRY(this, sl.hr_idx.reserve(n_cap)); CU_TRY(this, sl.hr_xyz.reserve(3 * n_cap));
    CU_TRY(this, sl.hr_w.reserve(n_cap)); CU_TRY(this, sl.hr_track.reserve(n_cap));
    CU_TRY(this, sl.hr_label.reserve(n_cap)); CU_TRY(this, sl.hr_state.reserve(n_cap));
    CU_TRY(this, sl.hr_rgb.reserve(3 * n_cap));
    const ResultArrays& r = sl.dres;
    // the counters reach the host before the emit pass has written the result arrays: another stream waits for that pass
    if (on != stream) CU_TRY(this, cudaStreamWaitEvent(on, sl.frame_done, 0));
    if (n) {
        CU_TRY(this, cudaMemcpyAsync(sl.hr_idx.p, r.voxel_index, n * 4, cudaMemcpyDeviceToHost, on));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_xyz.p, r.xyz, n * 12, cudaMemcpyDeviceToHost, on));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_w.p, r.weight_sum, n * 4, cudaMemcpyDeviceToHost, on));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_track.p, r.track_id, n * 2, cudaMemcpyDeviceToHost, on));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_label.p, r.label_id, n, cudaMemcpyDeviceToHost, on));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_state.p, r.state, n, cudaMemcpyDeviceToHost, on));
        if (sl.col_enabled) CU_TRY(this, cudaMemcpyAsync(sl.hr_rgb.p, r.rgb, 3 * n, cudaMemcpyDeviceToHost, on));
        else memset(sl.hr_rgb.p, 0, 3 * n);
    }
    if (nf) {
        const size_t nf_cap = (nf + grain - 1) / grain * grain;
        CU_TRY(this, sl.hr_free_idx.reserve(nf_cap)); CU_TRY(this, sl.hr_free_xyz.reserve(3 * nf_cap));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_free_idx.p, r.free_voxel_index, nf * 4, cudaMemcpyDeviceToHost, on));
        CU_TRY(this, cudaMemcpyAsync(sl.hr_free_xyz.p, r.free_xyz, nf * 12, cudaMemcpyDeviceToHost, on));
    }
    CU_TRY(this, cudaEventRecord(sl.copy_done, on));
    sl.copy_issued = true;
    return DSPMAP_OK;
}

int dspmap::fetch_result_to_host(int s) {
    FrameSlot& sl = slot[s];
    if (sl.host_result_valid) return DSPMAP_OK;
    if (!sl.copy_issued) { const int rc = issue_result_copy(s, stream); if (rc != DSPMAP_OK) return rc; }
    CU_TRY(this, cudaEventSynchronize(sl.copy_done));
    sl.host_result_valid = true;
    return DSPMAP_OK;
}

// =================================================== C ABI ===================================================
extern "C" {

int dspmap_version(void) { return 100; }

int dspmap_create(const dspmap_config* cfg, dspmap_t** out) {
    if (!cfg || !out) return DSPMAP_ERR_INVALID_ARG;
    *out = nullptr;
    dspmap* h = new (std::nothrow) dspmap();
    if (!h) return DSPMAP_ERR_INVALID_ARG;
    const int rc = h->init(*cfg);
    if (rc != DSPMAP_OK) {
        fprintf(stderr, "dspmap_create: %s\n", h->err.c_str());
        h->release_all();
        delete h;
        return rc;
    }
    *out = h;
    return DSPMAP_OK;
}

void dspmap_destroy(dspmap_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->drain();
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    h->release_all();
    delete h;
}

int dspmap_clear(dspmap_t* h) {
    if (!h) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    h->drain();          // a frame that failed is cleared like any other
    CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
    return h->clear_device();
}

const char* dspmap_last_error(const dspmap_t* h) { return h ? h->err.c_str() : "null handle"; }

int dspmap_set_params(dspmap_t* h, const dspmap_params* p) {
    if (!h || !p) return DSPMAP_ERR_INVALID_ARG;
    if (p->max_forget_count > 5 || p->max_forget_count < 0 || p->nb_ptc_num_per_point < 1)
        return h->fail(DSPMAP_ERR_INVALID_ARG, "max_forget_count must be in 0..5 and nb_ptc_num_per_point >= 1");
    h->prm = *p;
    return DSPMAP_OK;
}

int dspmap_set_label_luts(dspmap_t* h, const uint16_t static_value_to_track[256], const uint16_t* static_tracks,
                          const uint8_t* static_labels, int32_t n_static, uint16_t max_movable_track, uint16_t sky_track) {
    if (!h || !static_value_to_track || n_static < 0 || (n_static > 0 && (!static_tracks || !static_labels))) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    { const int rc = h->drain(); if (rc != DSPMAP_OK) return rc; }      // a pipelined frame may still read the tables
    memcpy(h->h_static_track, static_value_to_track, 256 * sizeof(uint16_t));
    std::fill(h->h_static_label.begin(), h->h_static_label.end(), (uint8_t)0);
    for (int i = 0; i < n_static; ++i) h->h_static_label[static_tracks[i]] = static_labels[i];
    h->max_movable = max_movable_track;
    h->sky_track = sky_track;
    CU_TRY(h, cudaMemcpy(h->d_static_track.p, h->h_static_track, 256 * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_static_label.p, h->h_static_label.data(), 65536, cudaMemcpyHostToDevice));
    return DSPMAP_OK;
}

int dspmap_set_random_table(dspmap_t* h, const float* table, uint32_t n) {
    if (!h || !table || n < 1) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    { const int rc = h->drain(); if (rc != DSPMAP_OK) return rc; }
    CU_TRY(h, h->d_randoms.alloc(n));
    CU_TRY(h, cudaMemcpy(h->d_randoms.p, table, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    h->rand_n = n;
    return DSPMAP_OK;
}

int dspmap_set_colors(dspmap_t* h, const dspmap_colors* c) {
    if (!h || !c) return DSPMAP_ERR_INVALID_ARG;
    if (c->scheme != 0 && c->scheme != 1 && c->scheme != 3) return h->fail(DSPMAP_ERR_INVALID_ARG, "colour scheme %d (0, 1 or 3)", c->scheme);
    CU_TRY(h, cudaSetDevice(h->device));
    { const int rc = h->drain(); if (rc != DSPMAP_OK) return rc; }
    h->colors = *c;
    uint8_t tab[1024];
    memcpy(tab, c->label_bgr, 768);
    memcpy(tab + 768, c->instance_map, 256);
    CU_TRY(h, cudaMemcpy(h->d_col_tab.p, tab, sizeof(tab), cudaMemcpyHostToDevice));
    return h->apply_params();
}

static int check_frame_args(dspmap_t* h, const float* depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                            int32_t n_masks, const double* position, const double* orientation, const dspmap_object_move* moves,
                            int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove) {
    if (!h) return DSPMAP_ERR_INVALID_ARG;
    if (!depth || !position || !orientation || n_masks < 0 || n_moves < 0 || n_remove < 0 || (n_masks && !masks) || (n_moves && !moves) ||
        (n_remove && !remove_tracks))
        return h->fail(DSPMAP_ERR_INVALID_ARG, "null pointer or negative count");
    // the reference indexes with the compile-time size and reads out of bounds otherwise (semantic_dsp_map.h:212-221 vs :974-979)
    if (rows != h->cfg.image_height || cols != h->cfg.image_width || row_stride < cols)
        return h->fail(DSPMAP_ERR_IMAGE_SIZE, "depth is %dx%d (stride %d), map is configured for %dx%d", rows, cols, row_stride, h->cfg.image_height, h->cfg.image_width);
    for (int m = 0; m < n_masks; ++m) {
        if (!masks[m].data || masks[m].row_stride < cols) return h->fail(DSPMAP_ERR_IMAGE_SIZE, "mask %d: null data or stride < width", m);
        if ((masks[m].rows || masks[m].cols) && (masks[m].rows != rows || masks[m].cols != cols))
            return h->fail(DSPMAP_ERR_IMAGE_SIZE, "mask %d is %dx%d, the depth image %dx%d", m, masks[m].rows, masks[m].cols, rows, cols);
    }
    return DSPMAP_OK;
}

// pipelined frames need a second set of result arrays: frame t's emit pass must not write where frame t-1's transfer reads
static int ensure_pipelined(dspmap_t* h) {
    if (h->cfg.shard_count > 1 || h->result2.p) return DSPMAP_OK;      // sharded maps carry both sets in their arenas
    CU_TRY(h, h->result2.alloc(h->lay.total - h->lay.r_idx));
    h->slot[1].dres = h->result_arrays_at(h->result2.p, h->lay.r_idx);
    return DSPMAP_OK;
}

static int update_device_impl(dspmap_t* h, const float* d_depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                              int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
                              int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace, bool pipelined, uint32_t* ticket) {
    int rc = check_frame_args(h, d_depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove);
    if (rc != DSPMAP_OK) return rc;
    CU_TRY(h, cudaSetDevice(h->device));
    if (pipelined && (rc = ensure_pipelined(h)) != DSPMAP_OK) return rc;
    const uint32_t t_before = h->T;
    rc = h->frame(d_depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove, want_freespace, 0,
                  nullptr, nullptr, pipelined, false);
    if (ticket) *ticket = h->T != t_before ? h->T : 0;
    return rc;
}

int dspmap_update_device(dspmap_t* h, const float* d_depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                         int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
                         int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace) {
    return update_device_impl(h, d_depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove, want_freespace, false, nullptr);
}

int dspmap_update_device_async(dspmap_t* h, const float* d_depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                               int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
                               int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace, uint32_t* ticket) {
    if (!ticket) return DSPMAP_ERR_INVALID_ARG;
    *ticket = 0;
    return update_device_impl(h, d_depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove, want_freespace, true, ticket);
}

static int update_host_impl(dspmap_t* h, const float* depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                            int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
                            int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace, bool pipelined, uint32_t* ticket) {
    static const bool trace = getenv("DSPMAP_TRACE") != nullptr;
    const double t_a = trace ? host_now_ms() : 0.0;
    int rc = check_frame_args(h, depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove);
    if (rc != DSPMAP_OK) return rc;
    CU_TRY(h, cudaSetDevice(h->device));
    if (pipelined && (rc = ensure_pipelined(h)) != DSPMAP_OK) return rc;
    if (!pipelined || h->profiling) { if ((rc = h->drain(false)) != DSPMAP_OK) return rc; }
    // the pinned staging buffers of the frame's slot: free since the frame two back finished
    dspmap::FrameSlot& sl = h->slot[h->next_slot()];
    if (sl.inflight && (rc = h->frame_finish(h->next_slot())) != DSPMAP_OK) return rc;
    const uint32_t t_before = h->T;
    // Stage the caller's (pageable) images through pinned memory.  Under BOOST_MODE the nearest 1/2 resize of
    // manualResize (pointcloud_tools.h:1104-1133: src = (int)(i * 2.0f), clamped) is applied while staging, so only the
    // quarter of the pixels the map will read crosses PCIe; K1 then reads the staged image 1:1 (same source pixels).
    const FrameParams& f = h->fp;
    const int srows = f.boost ? f.H : rows, scols = f.boost ? f.W : cols;
    const size_t npix = (size_t)srows * scols;
    std::vector<int> src_col(scols);
    for (int j = 0; j < scols; ++j) src_col[j] = f.boost ? std::min((int)((float)j * 2.0f), cols - 1) : j;
    CU_TRY(h, sl.p_depth.reserve(npix));
    std::vector<dspmap_mask> dm(masks, masks + n_masks);
    if (n_masks) {
        CU_TRY(h, sl.p_masks.reserve(npix * n_masks));
        if (sl.d_mask_in.n < npix * n_masks) CU_TRY(h, sl.d_mask_in.alloc(npix * n_masks));
    }
    if (!h->stage_pool) {
        const char* env = getenv("DSPMAP_STAGE_THREADS");
        const int want = env ? atoi(env) : 6;   // helper threads for the staging copies; 0 = stage on the calling thread
        const char* spin = getenv("DSPMAP_STAGE_SPIN_US");   // how long an idle helper polls for the next job before it sleeps
        h->stage_pool.reset(new StagePool(std::max(0, std::min(want, (int)std::thread::hardware_concurrency() - 1)), spin ? atoi(spin) : 200));
    }
    // drow[j] = srow[2 j]: the BOOST_MODE column decimation of manualResize (pointcloud_tools.h:1104-1133)
    auto take_even = [](auto* drow, const auto* srow, int n) {
        int j = 0;
#if defined(__AVX2__)
        if constexpr (sizeof(*drow) == 1) {
            const __m256i lo = _mm256_set1_epi16(0x00ff);
            for (; j + 32 <= n; j += 32) {
                const __m256i a = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(srow + 2 * j)), lo);
                const __m256i b = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(srow + 2 * j + 32)), lo);
                const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xd8);
                _mm256_storeu_si256(reinterpret_cast<__m256i*>(drow + j), p);
            }
        } else if constexpr (sizeof(*drow) == 4) {
            const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
            for (; j + 8 <= n; j += 8) {
                const __m256 a = _mm256_permutevar8x32_ps(_mm256_loadu_ps(reinterpret_cast<const float*>(srow + 2 * j)), idx);
                const __m256 b = _mm256_permutevar8x32_ps(_mm256_loadu_ps(reinterpret_cast<const float*>(srow + 2 * j + 8)), idx);
                _mm256_storeu_ps(reinterpret_cast<float*>(drow + j), _mm256_permute2f128_ps(a, b, 0x20));
            }
        }
#endif
        for (; j < n; ++j) drow[j] = srow[2 * j];
    };
    auto stage_rows = [&](auto* dst, const auto* src, int src_stride, int r0, int r1) {
        for (int i = r0; i < r1; ++i) {
            const int si = f.boost ? std::min((int)((float)i * 2.0f), rows - 1) : i;
            const auto* srow = src + (size_t)si * src_stride;
            auto* drow = dst + (size_t)i * scols;
            if (f.boost && 2 * scols <= cols) take_even(drow, srow, scols);                             // src = 2j exactly
            else if (f.boost) for (int j = 0; j < scols; ++j) drow[j] = srow[src_col[j]];
            else memcpy(drow, srow, (size_t)scols * sizeof(*dst));
        }
    };
    // depth first: the sweep and the pixel lists need nothing else, so the GPU starts while the masks are still being staged;
    // and even the depth is staged only after the pose-only work (vertex mask, object moves) has been launched
    const std::function<int()> late_depth = [&]() -> int {
        const int depth_parts = 8;
        h->stage_pool->run(depth_parts, [&](int t) { stage_rows(sl.p_depth.p, depth, row_stride, srows * t / depth_parts, srows * (t + 1) / depth_parts); });
        // the slot's device buffers were last read by the frame two back, which has finished: no wait before the upload
        CU_TRY(h, cudaMemcpyAsync(sl.d_depth_in.p, sl.p_depth.p, npix * sizeof(float), cudaMemcpyHostToDevice, h->upload_stream));
        CU_TRY(h, cudaEventRecord(sl.up_depth, h->upload_stream));
        CU_TRY(h, cudaStreamWaitEvent(h->stream, sl.up_depth, 0));
        return DSPMAP_OK;
    };
    const std::function<int(dspmap_mask*)> late_masks = [&](dspmap_mask* dmv) -> int {
        if (!n_masks) return DSPMAP_OK;
        const int parts = 2;   // row blocks per mask: enough tasks to keep every helper busy
        h->stage_pool->run(n_masks * parts, [&](int t) {
            const int m = t / parts, b = t % parts;
            stage_rows(sl.p_masks.p + (size_t)m * npix, masks[m].data, masks[m].row_stride, srows * b / parts, srows * (b + 1) / parts);
        });
        for (int m = 0; m < n_masks; ++m) { dmv[m].data = sl.d_mask_in.p + (size_t)m * npix; dmv[m].row_stride = scols; }
        CU_TRY(h, cudaMemcpyAsync(sl.d_mask_in.p, sl.p_masks.p, npix * n_masks, cudaMemcpyHostToDevice, h->upload_stream));
        CU_TRY(h, cudaEventRecord(sl.up_masks, h->upload_stream));
        CU_TRY(h, cudaStreamWaitEvent(h->stream, sl.up_masks, 0));
        return DSPMAP_OK;
    };
    const double t_b = trace ? host_now_ms() : 0.0;
    rc = h->frame(sl.d_depth_in.p, rows, cols, scols, dm.data(), n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove,
                  want_freespace, f.boost ? 1 : 0, &late_masks, &late_depth, pipelined, true);
    if (ticket) *ticket = h->T != t_before ? h->T : 0;
    if (rc != DSPMAP_OK) return rc;
    const double t_c = trace ? host_now_ms() : 0.0;
    if (!h->slot[h->cur].inflight) rc = h->fetch_result_to_host(h->cur);      // synchronous frame: the result is on the host on return
    if (trace && (h->T % 16) == 0)
        fprintf(stderr, "[dspmap] frame %u: stage+h2d-enqueue %.3f ms, frame (launch+sync) %.3f ms, result d2h %.3f ms\n", h->T, t_b - t_a, t_c - t_b,
                host_now_ms() - t_c);
    return rc;
}

int dspmap_update(dspmap_t* h, const float* depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                  int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
                  int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace) {
    return update_host_impl(h, depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove, want_freespace, false, nullptr);
}

int dspmap_update_async(dspmap_t* h, const float* depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
                        int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
                        int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace, uint32_t* ticket) {
    if (!ticket) return DSPMAP_ERR_INVALID_ARG;
    *ticket = 0;
    return update_host_impl(h, depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves, n_moves, remove_tracks, n_remove, want_freespace, true, ticket);
}

static void fill_result(const dspmap::FrameSlot& sl, bool host, dspmap_result* out) {
    out->count = sl.result_n;
    out->free_count = sl.want_free ? sl.result_nf : 0;
    if (host) {
        out->voxel_index = sl.hr_idx.p; out->xyz = sl.hr_xyz.p; out->weight_sum = sl.hr_w.p; out->track_id = sl.hr_track.p;
        out->label_id = sl.hr_label.p; out->state = sl.hr_state.p; out->free_voxel_index = sl.hr_free_idx.p; out->free_xyz = sl.hr_free_xyz.p;
        out->rgb = sl.hr_rgb.p;
    } else {
        const ResultArrays& r = sl.dres;
        out->voxel_index = r.voxel_index; out->xyz = r.xyz; out->weight_sum = r.weight_sum; out->track_id = r.track_id;
        out->label_id = r.label_id; out->state = r.state; out->free_voxel_index = r.free_voxel_index; out->free_xyz = r.free_xyz;
        out->rgb = r.rgb;
    }
}

int dspmap_wait_result(dspmap_t* h, uint32_t ticket, dspmap_result* out) {
    if (!h || !out) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    const int s = (int)(ticket & 1u);
    dspmap::FrameSlot& sl = h->slot[s];
    if (ticket == 0 || sl.T != ticket) return h->fail(DSPMAP_ERR_INVALID_ARG, "frame %u is not held any more (a result lives until two more frames have been submitted)", ticket);
    if (sl.inflight) {
        if (s == h->cur && h->slot[s ^ 1].inflight) { const int rc = h->frame_finish(s ^ 1); if (rc != DSPMAP_OK) return rc; }   // older frame first
        const int rc = h->frame_finish(s);
        if (rc != DSPMAP_OK) return rc;
    }
    if (sl.rc != DSPMAP_OK) return h->fail(sl.rc, "frame %u failed (reported when it finished)", ticket);
    if (sl.want_host) {
        if (!sl.copy_issued && !sl.host_result_valid) { const int rc = h->issue_result_copy(s, h->copy_stream); if (rc != DSPMAP_OK) return rc; }
        static const bool trace = getenv("DSPMAP_TRACE") != nullptr;
        const double ta = trace ? host_now_ms() : 0.0;
        const int rc = h->fetch_result_to_host(s);
        if (trace) h->trace_result_ms += host_now_ms() - ta;
        if (rc != DSPMAP_OK) return rc;
    }
    else CU_TRY(h, cudaEventSynchronize(sl.frame_done));      // device view: the emit pass has written the arrays
    fill_result(sl, sl.want_host, out);
    return DSPMAP_OK;
}

int dspmap_get_result(dspmap_t* h, dspmap_result* out) {
    if (!h || !out) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    int rc = h->drain();
    if (rc != DSPMAP_OK) return rc;
    rc = h->fetch_result_to_host(h->cur);
    if (rc != DSPMAP_OK) return rc;
    fill_result(h->slot[h->cur], true, out);
    return DSPMAP_OK;
}

int dspmap_get_result_device(dspmap_t* h, dspmap_result* out) {
    if (!h || !out) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    const int rc = h->drain();
    if (rc != DSPMAP_OK) return rc;
    fill_result(h->slot[h->cur], false, out);
    return DSPMAP_OK;
}

int dspmap_get_stats(dspmap_t* h, dspmap_frame_stats* out) {
    if (!h || !out) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    const int rc = h->drain();
    if (rc != DSPMAP_OK) return rc;
    *out = h->stats;
    return DSPMAP_OK;
}

int dspmap_set_profiling(dspmap_t* h, int32_t on) {
    if (!h) return DSPMAP_ERR_INVALID_ARG;
    h->profiling = on != 0;
    return DSPMAP_OK;
}

int dspmap_set_stream(dspmap_t* h, void* cuda_stream) {
    if (!h) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    { const int rc = h->drain(); if (rc != DSPMAP_OK) return rc; }
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return DSPMAP_OK;
}

int dspmap_get_launch_count(dspmap_t* h, uint64_t* count) {
    if (!h || !count) return DSPMAP_ERR_INVALID_ARG;
    *count = h->launches;
    return DSPMAP_OK;
}

int dspmap_object_particle_count(dspmap_t* h, int32_t track_id, uint32_t* count) {
    if (!h || !count) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    { const int rc = h->drain(); if (rc != DSPMAP_OK) return rc; }
    CU_TRY(h, cudaMemsetAsync(&h->d_cnt.p->scratch, 0, sizeof(uint32_t), h->stream));
    ++h->launches; k_count_track<<<h->sm_count * 8, 256, 0, h->stream>>>(Store{h->sx.p, h->sy.p, h->sz.p, h->sw.p, h->smeta.p, h->shdr.p}, h->v_local, (uint32_t)h->fp.Pb, (uint32_t)track_id & 0xffffu, &h->d_cnt.p->scratch);
    CU_TRY(h, cudaMemcpyAsync(count, &h->d_cnt.p->scratch, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return DSPMAP_OK;
}

int dspmap_comm_export(dspmap_t* h, uint8_t handle[DSPMAP_COMM_HANDLE_BYTES]) {
    if (!h || !handle) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    memset(handle, 0, DSPMAP_COMM_HANDLE_BYTES);
    static_assert(sizeof(cudaIpcMemHandle_t) <= DSPMAP_COMM_HANDLE_BYTES, "handle buffer too small");
    cudaIpcMemHandle_t ipc;
    CU_TRY(h, cudaIpcGetMemHandle(&ipc, h->arena.p));
    memcpy(handle, &ipc, sizeof ipc);
    return DSPMAP_OK;
}

int dspmap_comm_connect(dspmap_t* h, const uint8_t* handles) {
    if (!h || !handles) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    const int G = h->cfg.shard_count, me = h->cfg.shard_rank;
    for (int p = 0; p < G; ++p) {
        if (p == me) { h->peer_base[p] = h->arena.p; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + (size_t)p * DSPMAP_COMM_HANDLE_BYTES, sizeof ipc);
        void* base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return h->fail(DSPMAP_ERR_COMM, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
        h->peer_base[p] = base;
    }
    h->bind_arena(0);
    h->comm_connected = true;
    return DSPMAP_OK;
}

int dspmap_selftest_division(dspmap_t* h, uint64_t n, uint64_t seed, float b_lo, float b_hi, float a_max, uint64_t* mismatches) {
    if (!h || !mismatches) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    unsigned long long* d = nullptr;
    CU_TRY(h, cudaMalloc(&d, sizeof(unsigned long long)));
    CU_TRY(h, cudaMemsetAsync(d, 0, sizeof(unsigned long long), h->stream));
    k_selftest_invdiv<<<h->sm_count * 8, 256, 0, h->stream>>>(n, seed, b_lo, b_hi, a_max, d);
    unsigned long long out = 0;
    CU_TRY(h, cudaMemcpyAsync(&out, d, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(d);
    *mismatches = out;
    return DSPMAP_OK;
}

int dspmap_selftest_colour(dspmap_t* h, uint32_t* rgb_to_hsv, uint32_t* hsv_to_rgb) {
    if (!h || !rgb_to_hsv || !hsv_to_rgb) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    uint32_t* d = nullptr;
    const size_t n = (size_t)1 << 24;
    CU_TRY(h, cudaMalloc(&d, 2 * n * sizeof(uint32_t)));
    k_selftest_colour<<<h->sm_count * 8, 256, 0, h->stream>>>(d);
    cudaError_t e = cudaMemcpyAsync(rgb_to_hsv, d, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hsv_to_rgb, d + n, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    CU_TRY(h, e);
    return DSPMAP_OK;
}

int dspmap_dump(dspmap_t* h, int32_t what, void* dst, size_t cap, size_t* n_bytes) {
    if (!h) return DSPMAP_ERR_INVALID_ARG;
    CU_TRY(h, cudaSetDevice(h->device));
    { const int rc = h->drain(); if (rc != DSPMAP_OK) return rc; }
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const FrameParams& f = h->fp;
    const size_t npix = (size_t)f.W * f.H;
    std::vector<uint8_t> tmp;
    auto pull = [&](void* host, const void* dev, size_t bytes) { return cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost); };
    switch (what) {
        case DSPMAP_DUMP_POINTS: {
            std::vector<float4> pos(npix);
            std::vector<uint32_t> meta(npix);
            CU_TRY(h, pull(pos.data(), h->d_cloud_pos.p, npix * sizeof(float4)));
            CU_TRY(h, pull(meta.data(), h->d_cloud_meta.p, npix * 4));
            tmp.resize(npix * 20);
            for (size_t q = 0; q < npix; ++q) {
                uint8_t* r = tmp.data() + q * 20;
                memcpy(r, &pos[q], 16);
                const uint16_t tr = (uint16_t)pt_track(meta[q]);
                memcpy(r + 16, &tr, 2);
                r[18] = (uint8_t)pt_label(meta[q]);
                r[19] = (uint8_t)pt_valid(meta[q]);
            }
            break;
        }
        case DSPMAP_DUMP_PARTICLES: {
            const size_t n = h->n_slots_local;
            std::vector<float> x(n), y(n), z(n), w(n);
            std::vector<uint64_t> m(n);
            CU_TRY(h, pull(x.data(), h->sx.p, n * 4)); CU_TRY(h, pull(y.data(), h->sy.p, n * 4)); CU_TRY(h, pull(z.data(), h->sz.p, n * 4));
            CU_TRY(h, pull(w.data(), h->sw.p, n * 4)); CU_TRY(h, pull(m.data(), h->smeta.p, n * 8));
            std::vector<uint32_t> hd(h->v_local);
            CU_TRY(h, pull(hd.data(), h->shdr.p, (size_t)h->v_local * 4));
            for (size_t v = 0; v < h->v_local; ++v) m[v << f.Pb] = (m[v << f.Pb] & ~0xffffull) | hdr_ts(hd[v]);   // slot 0 = the time particle
            tmp.resize(n * 24);
            for (size_t i = 0; i < n; ++i) {
                uint8_t* r = tmp.data() + i * 24;
                memcpy(r, &x[i], 4); memcpy(r + 4, &y[i], 4); memcpy(r + 8, &z[i], 4); memcpy(r + 12, &w[i], 4); memcpy(r + 16, &m[i], 8);
            }
            break;
        }
        case DSPMAP_DUMP_PARTICLES_HASH: {
            unsigned long long* d = nullptr;
            CU_TRY(h, cudaMalloc(&d, sizeof(unsigned long long)));
            CU_TRY(h, cudaMemset(d, 0, sizeof(unsigned long long)));
            k_store_hash<<<h->sm_count * 16, 256, 0, h->stream>>>(h->st, h->slot_lo, h->n_slots_local, (uint32_t)f.Pb, d);
            tmp.resize(8);
            const cudaError_t e = cudaMemcpyAsync(tmp.data(), d, 8, cudaMemcpyDeviceToHost, h->stream);
            const cudaError_t e2 = cudaStreamSynchronize(h->stream);
            cudaFree(d);
            CU_TRY(h, e); CU_TRY(h, e2);
            break;
        }
        case DSPMAP_DUMP_PIXEL_OFFSETS:
            tmp.resize((npix + 1) * 4);
            CU_TRY(h, pull(tmp.data(), h->d_pix_offset.p, tmp.size()));
            break;
        case DSPMAP_DUMP_PIXEL_LISTS: {
            const size_t n = h->n_visible_last;
            tmp.resize(n * 4);
            if (n) CU_TRY(h, pull(tmp.data(), h->d_csr_idx.p, tmp.size()));
            break;
        }
        case DSPMAP_DUMP_CK:
        {   // the .w of every pixel's {x, y, z, ck} record of the padded observation plane; 0 where the pixel is invalid
            const int R = f.win_r, Wp = obs_stride(f.W, R);
            std::vector<float> plane(obs_cells(f.W, f.H, R) * 4);
            std::vector<uint32_t> flags(obs_cells(f.W, f.H, R) * 2);
            CU_TRY(h, pull(plane.data(), h->d_obs_a, plane.size() * 4));
            CU_TRY(h, pull(flags.data(), h->d_obs_b, flags.size() * 4));
            tmp.resize(npix * 4);
            float* out = reinterpret_cast<float*>(tmp.data());
            for (int i = 0; i < f.H; ++i)
                for (int j = 0; j < f.W; ++j) {
                    const size_t c = (size_t)(i + R) * Wp + (j + R);
                    out[(size_t)i * f.W + j] = (flags[2 * c + 1] & OB_VALID) ? plane[4 * c + 3] : 0.f;
                }
        }
            break;
        case DSPMAP_DUMP_RING: {
            tmp.resize(4 * (9 + f.NX + f.NY + f.NZ));
            uint8_t* wp = tmp.data();
            memcpy(wp, h->moved, 12); wp += 12; memcpy(wp, h->equiv, 12); wp += 12; memcpy(wp, h->map_center, 12); wp += 12;
            for (int a = 0; a < 3; ++a) { memcpy(wp, h->stamps[a].data(), h->stamps[a].size() * 4); wp += h->stamps[a].size() * 4; }
            break;
        }
        case DSPMAP_DUMP_VOXEL_STATE: {
            tmp.resize(h->v_local);
            CU_TRY(h, pull(tmp.data(), h->d_vstate.p, tmp.size()));
            for (auto& b8 : tmp) b8 = (uint8_t)(int8_t)((int)(b8 & 3u) - 1);   // state byte -> -1 unknown, 0 free, 1 occupied, 2 guessed
            break;
        }
        default:
            return h->fail(DSPMAP_ERR_INVALID_ARG, "unknown dump %d", what);
    }
    if (n_bytes) *n_bytes = tmp.size();
    if (dst && cap) memcpy(dst, tmp.data(), std::min(cap, tmp.size()));
    return DSPMAP_OK;
}

}  // extern "C"
