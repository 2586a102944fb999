// ref_harness.cpp — runs the REFERENCE's own hot-path code (compiled unmodified from /root/reference/include)
// behind a small C interface, so the CPU oracle (dsp_oracle.cpp) can be pinned against it.
//
// TEST INFRASTRUCTURE ONLY.  Built into oracle/_ref/libdsp_ref_s<SETTING>.so by oracle/Makefile when
// /root/reference is present; loaded only by tests/ and tests/golden/make_ref_golden.py.  No reference source
// is copied: this file #includes "semantic_dsp_map.h" from the reference tree where it lies and compiles it
// against the stand-in third-party headers in oracle/ref_stubs/ (Eigen / OpenCV / PCL / Boost are absent from
// this image; what the stand-ins do and do not anchor is stated at the top of ref_stubs/Eigen/Dense).
//
// What runs is the reference's code: PointCloudTools::generateLabeledPointCloud, SemanticDSPMap::
// subObjectLevelUpdate and everything below it (ring shift, object moves, BFS visibility, updateParticles,
// birth + resample, getOccupancyResult incl. colouring), plus ObjectInfoHandler::readObjectInfo for the label
// tables.  What does NOT run is objectLevelUpdate (include/semantic_dsp_map.h:306-566; Kabsch / RANSAC object
// layer, out of scope — SURVEY.md §8): ref_update() restates the dozen glue lines of SemanticDSPMap::update()
// around it (:170-251) and injects the object layer's per-frame OUTPUT (which tracks move by which rigid
// transform, which tracks are dropped) into ObjectSet exactly where objectLevelUpdate would have left it.
//
// The reference keeps its state in namespace-scope globals and function statics (buffer.h:86-120,
// operations.h:70, semantic_dsp_map.h:965-969): one map per process, never re-created.  Run every scenario in a
// fresh process.
//
// Build: -fno-access-control lets this file reach private members (op_mt_, pt_tools_, object_set_, the two
// random-table cursors); the reference itself is not edited.
#include <Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <pcl/point_types.h>

#include <chrono>
#include <cstring>
#include <fstream>
#include <mutex>
#include <random>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "semantic_dsp_map.h"  // the reference's, via -I/root/reference/include

#include "../include/dspmap_c.h"

namespace {

SemanticDSPMap* g_map = nullptr;
pcl::PointCloud<pcl::PointXYZRGB>::Ptr g_occ, g_free;
std::vector<std::vector<LabeledPoint>> g_points;
std::vector<int32_t> g_move_order;
double g_seconds_last = 0.0;
int g_rows = 0, g_cols = 0;

SemanticDSPMap& map() {
    if (!g_map) g_map = new SemanticDSPMap();  // constructor fills the random table from std::random_device; see ref_set_random_table
    return *g_map;
}

}  // namespace

extern "C" {

typedef struct ref_mask {
    const uint8_t* data;   // MONO8, rows x cols
    int32_t rows, cols, row_stride;
    int32_t track_id;
    const char* label;     // MaskKpts.label ("static" or a class name)
    const double* kpts;    // n_kpts x 3, MaskKpts.kpts_current
    int32_t n_kpts;
} ref_mask;

// The compile-time configuration of this build of the reference (settings.h), in the C-ABI's terms.
int ref_describe(dspmap_config* out) {
    std::memset(out, 0, sizeof(*out));
    out->voxel_bits_x = C_VOXEL_NUM_AXIS_X_N;
    out->voxel_bits_y = C_VOXEL_NUM_AXIS_Y_N;
    out->voxel_bits_z = C_VOXEL_NUM_AXIS_Z_N;
    out->slot_bits = C_MAX_PARTICLE_NUM_PER_VOXEL_N;
    out->voxel_size = C_VOXEL_SIZE;
    out->fx = g_camera_fx_set; out->fy = g_camera_fy_set; out->cx = g_camera_cx_set; out->cy = g_camera_cy_set;
    out->image_width = g_image_width_set;
    out->image_height = g_image_height_set;
    out->boost_mode = BOOST_MODE;
    out->depth_min = g_depth_range_min;
    out->depth_max = g_depth_range_max;
    out->consider_instance = g_consider_instance ? 1 : 0;
    out->zed2_filters = (SETTING == 3) ? 1 : 0;
    out->window_half = 0;
    out->shard_count = 1;
    return SETTING;
}

// ObjectInfoHandler::readObjectInfo (utils/object_info_handler.h:28-91), as src/mapping.cpp:89-93 calls it.
int ref_read_object_info(const char* csv_path) {
    ObjectInfoHandler h;
    h.readObjectInfo(csv_path);
    return g_max_movable_object_instance_id;
}

// The five setters src/mapping.cpp:135-140 calls.
int ref_set_params(const dspmap_params* p, int max_observation_lost_time, int eval_format) {
    SemanticDSPMap& m = map();
    m.setMapParameters(p->detection_probability, p->noise_number, p->nb_ptc_num_per_point, p->occupancy_threshold, max_observation_lost_time,
                       p->forgetting_rate, p->max_forget_count, 0.5f, p->id_transition_probability);
    m.setMapOptions(p->consider_depth_noise != 0, p->use_independent_filter != 0);
    m.setVisualizeOptions(p->zero_center != 0, eval_format != 0);
    m.setDepthNoiseModelParameters(p->depth_noise_first_order, p->depth_noise_zero_order);
    return 0;
}

// Determinism: overwrite the std::random_device-seeded table (utils/basic_algorithms.h:394-401, a namespace-scope
// vector shared by both calculators) and rewind both cursors (semantic_dsp_map.h:265, mc_ring/operations.h:1460).
int ref_set_random_table(const float* table, uint32_t n) {
    SemanticDSPMap& m = map();
    if (n != GAUSSIAN_RANDOM_NUM) return -1;
    for (uint32_t i = 0; i < n; ++i) gaussian_randoms[i] = table[i];
    m.gaussian_random_.gaussian_random_seq_ = 0;
    m.op_mt_.gaussian_random_calculator_.gaussian_random_seq_ = 0;
    return 0;
}

// The instance colour permutation is shuffled from std::random_device (semantic_dsp_map.h:45-48): inject it.
int ref_set_color_map(const int32_t* map256) {
    SemanticDSPMap& m = map();
    for (int i = 0; i < 256; ++i) m.color_map_int_256_[i] = map256[i];
    return 0;
}

int ref_update(const float* depth, int32_t rows, int32_t cols, int32_t row_stride, const ref_mask* masks, int32_t n_masks,
               const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves, int32_t n_moves,
               const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace) {
    SemanticDSPMap& m = map();

    // ---- inputs in the reference's own types (src/mapping.cpp:180-251) ----
    cv::Mat depth_value_mat(rows, cols, CV_32FC1);
    for (int i = 0; i < rows; ++i) std::memcpy(depth_value_mat.ptr<float>(i), depth + static_cast<size_t>(i) * row_stride, sizeof(float) * cols);
    std::vector<MaskKpts> ins_seg_result(n_masks);
    for (int k = 0; k < n_masks; ++k) {
        MaskKpts& mk = ins_seg_result[k];
        mk.track_id = masks[k].track_id;
        mk.label = masks[k].label;
        mk.mask = cv::Mat(masks[k].rows, masks[k].cols, CV_8UC1);
        for (int i = 0; i < masks[k].rows; ++i) std::memcpy(mk.mask.ptr<uchar>(i), masks[k].data + static_cast<size_t>(i) * masks[k].row_stride, masks[k].cols);
        for (int j = 0; j < masks[k].n_kpts; ++j) mk.kpts_current.push_back(Eigen::Vector3d(masks[k].kpts[3 * j], masks[k].kpts[3 * j + 1], masks[k].kpts[3 * j + 2]));
        mk.bbox = BBox2D{0, 0, 0, 0};
    }
    Eigen::Vector3d camera_position(position[0], position[1], position[2]);
    Eigen::Quaterniond camera_orientation(orientation_wxyz[0], orientation_wxyz[1], orientation_wxyz[2], orientation_wxyz[3]);
    g_occ.reset(new pcl::PointCloud<pcl::PointXYZRGB>);
    g_free.reset(new pcl::PointCloud<pcl::PointXYZRGB>);

    const auto t0 = std::chrono::steady_clock::now();

    // ---- SemanticDSPMap::update, semantic_dsp_map.h:173-186 ----
    global_time_stamp += 1;
    for (size_t i = 0; i < ins_seg_result.size(); ++i)
        if (ins_seg_result[i].label != "static" && ins_seg_result[i].track_id > g_max_movable_object_instance_id)
            ins_seg_result[i].track_id = ins_seg_result[i].track_id % g_max_movable_object_instance_id;

    // ---- stand-in for objectLevelUpdate (:189-191): leave ObjectSet as that function would ----
    if (g_consider_instance) {
        auto& table = m.object_set_.object_tracking_hash_map;
        for (const MaskKpts& mk : ins_seg_result) {  // Case 1 (:346-364): a newly observed movable object joins the set
            if (mk.label == "static" || mk.track_id > g_max_movable_object_instance_id) continue;
            if (!m.object_set_.checkIfObjectExists(mk.track_id)) {
                MJObject object;
                object.time_stamp = global_time_stamp;
                object.confidence = 1.f;
                object.label = g_label_id_map_default[mk.label];
                m.object_set_.addNewObject(object, mk.track_id);
            }
        }
        for (auto& kv : table) kv.second.object.rigidbody_moved_vec.assign(1, false);
        for (int i = 0; i < n_moves; ++i) {  // updateObject's result (:484-510): "moved", with this rigid transform, observed now
            if (!m.object_set_.checkIfObjectExists(moves[i].track_id)) {
                MJObject object;
                object.time_stamp = global_time_stamp;
                object.confidence = 1.f;
                object.label = 0;
                m.object_set_.addNewObject(object, moves[i].track_id);
            }
            auto& obj = table.at(moves[i].track_id);
            Eigen::Matrix4d T;
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) T(r, c) = static_cast<double>(moves[i].t[4 * r + c]);
            obj.object.rigidbody_tmatrix_vec.assign(1, T);
            obj.object.rigidbody_moved_vec.assign(1, true);
            obj.observation_time_step = global_time_stamp;
        }
        for (int i = 0; i < n_remove; ++i) {  // a moving object unobserved for max_obersevation_lost_time_ frames (:600-606)
            if (!m.object_set_.checkIfObjectExists(remove_tracks[i])) {
                MJObject object;
                object.label = 0;
                m.object_set_.addNewObject(object, remove_tracks[i]);
            }
            auto& obj = table.at(remove_tracks[i]);
            obj.object.rigidbody_tmatrix_vec.assign(1, Eigen::Matrix4d::Identity());
            obj.object.rigidbody_moved_vec.assign(1, true);
            obj.observation_time_step = static_cast<int>(global_time_stamp) - m.max_obersevation_lost_time_;
        }
        // the order in which subObjectLevelUpdate (:593-688) will visit the moving objects: a host-container artefact
        // (std::unordered_map iteration) that fixes the order of random-table draws; reported so the oracle can be fed
        // its `moves` in the same order.
        g_move_order.clear();
        for (auto it = table.begin(); it != table.end(); ++it) {
            if (it->second.object.rigidbody_moved_vec.size() == 0 || !it->second.object.rigidbody_moved_vec[0]) continue;
            if (static_cast<int>(global_time_stamp - it->second.observation_time_step) >= m.max_obersevation_lost_time_) continue;
            int track = it->second.track_id;
            if (m.object_set_.obj_ptc_hash_map.checkIfObjectExists(track)) g_move_order.push_back(track);
        }
    }

    // ---- :209-242 ----
    m.pt_tools_.updateCameraPose(camera_position, camera_orientation);
#if BOOST_MODE == 0
    int cols_i = depth_value_mat.cols;
    int rows_i = depth_value_mat.rows;
#else
    int cols_i = depth_value_mat.cols * g_image_rescale;
    int rows_i = depth_value_mat.rows * g_image_rescale;
#endif
    std::vector<std::vector<LabeledPoint>> labeled_point_cloud(rows_i, std::vector<LabeledPoint>(cols_i));
    std::unordered_map<uint16_t, std::vector<Eigen::Vector3d>> tracked_objects_points;
    std::unordered_map<int, int> track_to_label_id_map;
    m.pt_tools_.generateLabeledPointCloud(depth_value_mat, ins_seg_result, labeled_point_cloud, tracked_objects_points, track_to_label_id_map,
                                          m.depth_noise_model_first_order_, m.depth_noise_model_zero_order_);
    m.subObjectLevelUpdate(labeled_point_cloud, tracked_objects_points, track_to_label_id_map, depth_value_mat, ins_seg_result, camera_position,
                           camera_orientation, g_occ, g_free, want_freespace != 0);

    g_seconds_last = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_points.swap(labeled_point_cloud);
    g_rows = rows_i;
    g_cols = cols_i;
    return 0;
}

// test hook: the stand-in cv::cvtColor the reference's colouring runs on here (checked against the real cv2 module)
void ref_cvt_color(int32_t code, const uint8_t* in, uint8_t* out, uint32_t n) {
    cv::Mat src(1, static_cast<int>(n), CV_8UC3, const_cast<uint8_t*>(in)), dst;
    cv::cvtColor(src, dst, code == 0 ? cv::COLOR_RGB2HSV : cv::COLOR_HSV2RGB);
    std::memcpy(out, dst.data, 3 * static_cast<size_t>(n));
}

double ref_last_seconds(void) { return g_seconds_last; }
uint32_t ref_time_stamp(void) { return global_time_stamp; }

// out: 3 floats + 3 bytes (r,g,b) per point, in emission order (ascending storage voxel index)
uint32_t ref_get_cloud(int32_t which_free, float* xyz, uint8_t* rgb, uint32_t cap) {
    const auto& c = which_free ? g_free : g_occ;
    if (!c) return 0;
    const uint32_t n = static_cast<uint32_t>(c->points.size());
    for (uint32_t i = 0; i < n && i < cap; ++i) {
        const pcl::PointXYZRGB& p = c->points[i];
        if (xyz) { xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z; }
        if (rgb) { rgb[3 * i] = p.r; rgb[3 * i + 1] = p.g; rgb[3 * i + 2] = p.b; }
    }
    return n;
}

// getVoxelGlobalPosition for every storage voxel index (pure index math, operations.h:645-648, 940-983): lets a test
// recover the voxel index of each emitted point without touching the particle store.
uint32_t ref_voxel_positions(float* xyz, uint32_t cap) {
    SemanticDSPMap& m = map();
    for (uint32_t v = 0; v < C_VOXEL_NUM_TOTAL && v < cap; ++v) {
        Eigen::Vector3f p;
        m.op_mt_.getVoxelGlobalPosition(v, p);
        xyz[3 * v] = p.x(); xyz[3 * v + 1] = p.y(); xyz[3 * v + 2] = p.z();
    }
    return C_VOXEL_NUM_TOTAL;
}

int32_t ref_move_order(int32_t* tracks, int32_t cap) {
    for (size_t i = 0; i < g_move_order.size() && static_cast<int32_t>(i) < cap; ++i) tracks[i] = g_move_order[i];
    return static_cast<int32_t>(g_move_order.size());
}

// Same payloads as dspmap_dump (include/dspmap_c.h) so one comparison routine serves oracle, engine and reference.
int ref_dump(int32_t what, void* dst, size_t cap, size_t* n_bytes) {
    std::vector<uint8_t> buf;
    auto put = [&](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); buf.insert(buf.end(), b, b + n); };
    switch (what) {
    case DSPMAP_DUMP_POINTS:
        for (int i = 0; i < g_rows; ++i)
            for (int j = 0; j < g_cols; ++j) {
                const LabeledPoint& p = g_points[i][j];
                float f[4] = {p.position.x(), p.position.y(), p.position.z(), p.sigma};
                uint8_t valid = p.is_valid ? 1 : 0;
                if (!valid) { f[0] = f[1] = f[2] = f[3] = 0.f; }  // uninitialised in the reference (pointcloud_tools.h:224-232)
                uint16_t track = valid ? p.track_id : 0;
                uint8_t label = valid ? p.label_id : 0;
                put(f, 16); put(&track, 2); put(&label, 1); put(&valid, 1);
            }
        break;
    case DSPMAP_DUMP_PARTICLES:
        static_assert(sizeof(Particle) == 24, "reference record");
        put(PARTICLE_ARRAY.data(), sizeof(Particle) * static_cast<size_t>(C_MAX_PARTICLE_NUM));
        break;
    case DSPMAP_DUMP_PIXEL_OFFSETS: {
        uint32_t acc = 0;
        for (int i = 0; i < g_image_height; ++i)
            for (int j = 0; j < g_image_width; ++j) { put(&acc, 4); acc += particle_to_pixel_num_array[i][j]; }
        put(&acc, 4);
        break;
    }
    case DSPMAP_DUMP_PIXEL_LISTS:  // BFS discovery order inside a pixel (operations.h:1404-1407)
        for (int i = 0; i < g_image_height; ++i)
            for (int j = 0; j < g_image_width; ++j) {
                const uint32_t n = particle_to_pixel_num_array[i][j];
                if (!n) continue;
                const auto& v = particle_to_pixel_index_map[i * g_image_width + j];
                put(v.data(), 4 * static_cast<size_t>(n));
            }
        break;
    case DSPMAP_DUMP_RING: {
        const int32_t steps[6] = {buffer_moved_steps_x, buffer_moved_steps_y, buffer_moved_steps_z, buffer_moved_equivalent_steps_x,
                                  buffer_moved_equivalent_steps_y, buffer_moved_equivalent_steps_z};
        const float c[3] = {map_center_pos.x(), map_center_pos.y(), map_center_pos.z()};
        put(steps, sizeof(steps)); put(c, sizeof(c));
        put(voxel_time_stamps_x, sizeof(voxel_time_stamps_x));
        put(voxel_time_stamps_y, sizeof(voxel_time_stamps_y));
        put(voxel_time_stamps_z, sizeof(voxel_time_stamps_z));
        break;
    }
    default:
        return DSPMAP_ERR_INVALID_ARG;
    }
    if (n_bytes) *n_bytes = buf.size();
    if (dst) {
        if (cap < buf.size()) return DSPMAP_ERR_INVALID_ARG;
        std::memcpy(dst, buf.data(), buf.size());
    }
    return DSPMAP_OK;
}

}  // extern "C"
