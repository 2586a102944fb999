// semantic_dsp_map.h — drop-in replacement for the reference header of the same name
// (reference include/semantic_dsp_map.h:23-251): class SemanticDSPMap with the same public methods, implemented over
// the C ABI of the B200 engine (include/dspmap_c.h, libdspmap_b200.so).  src/mapping.cpp includes this header instead
// of the reference one and compiles unchanged: same constructor, the five setters it calls (mapping.cpp:135-140) and
// update() (mapping.cpp:273).  Host work kept here = what the reference does with string keys and tiny per-object
// loops: label look-ups, bbox from key points, the track-id wrap.  Nothing here touches particles, pixels or voxels.
//
// Needs the same third-party headers as the reference (Eigen, OpenCV, PCL).  The object layer
// (objectLevelUpdate, reference :306-566: a handful of objects, key-point Kabsch, moving-or-not filter, coasting) runs on
// the host in dsp_object_tracker.h, fed from update() like the reference does; its per-frame product (one rigid matrix per
// moving object, tracks to drop) goes to the engine with the frame.  An integrator with an object layer of their own calls
// setObjectMoves() before update() instead, which switches the built-in one off.
#pragma once
#include <Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "dspmap_c.h"
#include "dsp_object_tracker.h"

// reference include/utils/data_base.h:24-30 and include/utils/tracking_result_handler.h:15-26 (src/mapping.cpp:216-219 writes bbox.x1 .. y2)
struct BBox2D { int x1 = 0, y1 = 0, x2 = 0, y2 = 0; };
struct MaskKpts {
    int track_id = 0;
    std::string label;
    std::vector<Eigen::Vector3d> kpts_current;
    std::vector<Eigen::Vector3d> kpts_previous;
    cv::Mat mask;
    BBox2D bbox;
};

// label tables: what ObjectInfoHandler::readObjectInfo leaves in the reference's globals
// (include/utils/data_base.h:108-212; include/utils/object_info_handler.h:28-91)
struct DspLabelTables {
    std::unordered_map<std::string, int> label_id;            // g_label_id_map_default
    std::unordered_map<int, std::string> label_id_reversed;   // g_label_id_map_reversed
    std::unordered_map<std::string, int> label_id_static;     // g_label_id_map_static
    std::unordered_map<std::string, int> label_to_instance;   // g_label_to_instance_id_map_default
    std::map<int, std::string> instance_to_label;             // g_instance_id_to_label_map_default
    std::unordered_map<int, cv::Vec3b> label_color;           // g_label_color_map_default, (b, g, r)
    int max_movable_instance_id = 65523;                      // g_max_movable_object_instance_id

    void addRow(int id, const std::string& name, int instance_id, int color_b = 0, int color_g = 0, int color_r = 0) {   // object_info_handler.h:51-86
        label_id[name] = id;
        label_id_reversed[id] = name;
        label_color[id] = cv::Vec3b((unsigned char)color_b, (unsigned char)color_g, (unsigned char)color_r);
        if (instance_id > 0) {
            label_id_static[name] = id;
            label_to_instance[name] = instance_id;
            instance_to_label[instance_id] = name;
            if (instance_id - 1 < max_movable_instance_id) max_movable_instance_id = instance_id - 1;
        }
    }
    template <class M, class K> static int getOr0(const M& m, const K& k) { auto it = m.find(k); return it == m.end() ? 0 : it->second; }

    // the tables the reference starts with when no CSV is read (include/utils/data_base.h:108-232)
    static DspLabelTables builtin() {
        DspLabelTables t;
        struct Row { int id; const char* name; int inst; int b, g, r; };
        static const Row rows[] = {{0, "Background", 65535, 0, 0, 0}, {2, "Terrain", 65534, 200, 0, 210}, {3, "Sky", 65533, 255, 200, 90},
                                   {4, "Tree", 65532, 0, 199, 0}, {5, "Vegetation", 65531, 0, 240, 90}, {6, "Building", 65530, 140, 140, 140},
                                   {7, "Road", 65529, 100, 60, 100}, {8, "GuardRail", 65528, 255, 100, 250}, {9, "TrafficSign", 65527, 0, 255, 255},
                                   {10, "TrafficLight", 65526, 0, 200, 200}, {11, "Pole", 65525, 0, 130, 255}, {12, "Misc", 65524, 80, 80, 80},
                                   {13, "Truck", -1, 60, 60, 160}, {14, "Car", -1, 80, 127, 255}, {15, "Person", -1, 139, 139, 0}};
        for (const Row& r : rows) t.addRow(r.id, r.name, r.inst, r.b, r.g, r.r);
        t.max_movable_instance_id = 65523;
        return t;
    }
};

// The reference keeps its label tables in namespace-scope globals that ObjectInfoHandler::readObjectInfo rewrites
// (include/utils/data_base.h:108-232; include/utils/object_info_handler.h:28-91) and that the map reads at every frame.
// Same arrangement here: one process-wide table set with a version stamp; the map re-sends its look-up tables to the engine
// when the stamp moved (src/mapping.cpp constructs the map BEFORE it reads the CSV, :71 vs :89-93).
struct DspLabelGlobals {
    DspLabelTables tables = DspLabelTables::builtin();
    unsigned version = 1;
};
inline DspLabelGlobals& dspLabelGlobals() { static DspLabelGlobals g; return g; }

// reference include/utils/object_info_handler.h:20-152
class ObjectInfoHandler {
public:
    void readObjectInfo(const std::string& file_path) {
        if (file_path.substr(file_path.find_last_of(".") + 1) != "csv") {                       // :31-35
            std::fprintf(stderr, "The file path is not a csv file. Will use default settings in data_base.h\n");
            return;
        }
        std::ifstream file(file_path);
        std::string line;
        std::getline(file, line);                                                                // header line, :138
        DspLabelTables t;
        t.max_movable_instance_id = 65535;                                                       // :49
        while (std::getline(file, line)) {
            std::stringstream ss(line);
            std::string item[6];
            for (auto& it : item) std::getline(ss, it, ',');
            t.addRow(std::stoi(item[0]), item[1], std::stoi(item[2]), std::stoi(item[3]), std::stoi(item[4]), std::stoi(item[5]));
        }
        dspLabelGlobals().tables = t;
        ++dspLabelGlobals().version;
        std::printf("Object information csv file read successfully. File: %s\n g_max_movable_object_instance_id = %d\n", file_path.c_str(), t.max_movable_instance_id);
    }
};

// reference include/utils/tracking_result_handler.h:28-34
class TrackingResultHandler {
public:
    std::vector<MaskKpts> tracking_result;
};

class SemanticDSPMap {
public:
    // The compile-time constants of the reference (include/settings/settings.h) become a runtime config.  The default
    // constructor - the one src/mapping.cpp:71 uses - takes defaultConfig(), which starts as the configuration the
    // reference ships (SETTING 3, settings.h:100-119, 136-143) and can be overwritten before the map is constructed.
    static dspmap_config& defaultConfig() {
        static dspmap_config c = [] {
            dspmap_config d{};
            d.voxel_bits_x = 7; d.voxel_bits_y = 5; d.voxel_bits_z = 7; d.slot_bits = 2; d.voxel_size = 0.15f;
            d.fx = d.fy = 527.8191528320312f; d.cx = 633.9357299804688f; d.cy = 366.3338623046875f;
            d.image_width = 1280; d.image_height = 720; d.boost_mode = 1; d.depth_min = 0.3f; d.depth_max = 15.f;
            d.consider_instance = 1; d.zed2_filters = 1; d.shard_count = 1;
            return d;
        }();
        return c;
    }
    SemanticDSPMap() : SemanticDSPMap(defaultConfig(), dspLabelGlobals().tables) { follow_globals_ = true; labels_version_ = dspLabelGlobals().version; }
    explicit SemanticDSPMap(const dspmap_config& cfg, const DspLabelTables& labels = DspLabelTables()) : cfg_(cfg), labels_(labels) {
        if (dspmap_create(&cfg_, &h_) != DSPMAP_OK) { std::fprintf(stderr, "dspmap_create failed\n"); std::exit(-1); }   // reference: exit(-1), operations.h:57
        prm_ = dspmap_params{0.95f, 0.1f, 3, 0.2f, 1.0f, 5, 0.1f, 0.0f, 0.1f, 0, 0, 0};   // ctor defaults, reference :25-43
        for (int i = 0; i < 256; ++i) color_map_int_256_[i] = (uint8_t)i;                       // reference :45-48: a random shuffle of 0..255
        // The reference draws its instance colours and its jitter table from std::random_device.  DSPMAP_ADAPTER_SEED in the
        // environment replaces that entropy by a fixed seed (reproducible runs: tests, regression comparisons).
        std::random_device rd_device;
        const char* fixed_seed = std::getenv("DSPMAP_ADAPTER_SEED");
        std::mt19937 seed_source(fixed_seed ? (uint32_t)std::strtoul(fixed_seed, nullptr, 10) : rd_device());
        auto rd = [&seed_source]() { return (uint32_t)seed_source(); };
        std::mt19937 gen(rd());
        for (int i = 255; i > 0; --i) { std::uniform_int_distribution<int> dist(0, i); std::swap(color_map_int_256_[i], color_map_int_256_[dist(gen)]); }
        // DSPMAP_ADAPTER_REFERENCE_ORDER=1: follow the reference's iteration order and its compiled uninitialised-sigma reads
        // (setReferenceIterationOrder, followCompiledReferenceSigma): a scene without moving objects is then published voxel for
        // voxel, colour for colour as the reference node publishes it (regression comparisons against a recorded reference run)
        const char* exact = std::getenv("DSPMAP_ADAPTER_REFERENCE_ORDER");
        follow_reference_ = exact && exact[0] == '1';
        const char* piped = std::getenv("DSPMAP_ADAPTER_PIPELINED");      // = setPipelinedOutput(true), for callers that cannot be edited
        pipelined_ = piped && piped[0] == '1';
        pushLabels();
        pushColors();
        setGaussianTableSeed(rd());                                                              // reference :65 -> basic_algorithms.h:394-402
        configureTracker(cfg_.zed2_filters ? dsp_adapter::ObjectTracker::kBoxKeypoints : dsp_adapter::ObjectTracker::kMatchedKeypoints);
    }
    ~SemanticDSPMap() { dspmap_destroy(h_); }
    SemanticDSPMap(const SemanticDSPMap&) = delete;
    SemanticDSPMap& operator=(const SemanticDSPMap&) = delete;

    void clear() { dspmap_clear(h_); known_objects_.clear(); born_tracks_.clear(); tracker_.clear(); step_ = 0; }                                                          // reference :74-81
    void setTemplatePath(std::string) {}                                                         // template matching: out of scope
    void setMapParameters(float detection_probability, float noise_number, int nb_ptc_num_per_point, float occupancy_threshold,
                          int max_obersevation_lost_time, float forgetting_rate = 1.f, int max_forget_count = 5,
                          float match_score_threshold = 0.5f, float id_transition_probability = 0.1f) {   // reference :101-119
        prm_.detection_probability = detection_probability; prm_.noise_number = noise_number;
        prm_.nb_ptc_num_per_point = nb_ptc_num_per_point; prm_.occupancy_threshold = occupancy_threshold;
        prm_.forgetting_rate = forgetting_rate; prm_.max_forget_count = max_forget_count;
        prm_.id_transition_probability = id_transition_probability;
        max_obersevation_lost_time_ = max_obersevation_lost_time; match_score_threshold_ = match_score_threshold;
        pushParams();
    }
    void setMapOptions(bool if_consider_depth_noise, bool if_use_independent_filter) {          // reference :121-126
        prm_.consider_depth_noise = if_consider_depth_noise; prm_.use_independent_filter = if_use_independent_filter;
        pushParams();
    }
    void setVisualizeOptions(bool visualize_with_zero_center, bool if_out_evaluation_format = false) {   // reference :130-134
        prm_.zero_center = visualize_with_zero_center; if_out_evaluation_format_ = if_out_evaluation_format;
        pushParams();
        pushColors();
    }
    // The table of N(0, 0.05) draws that moved and noisy-newborn particles are jittered with (gaussian_randoms,
    // utils/basic_algorithms.h:377-402: 10^6 draws of std::normal_distribution<> from a std::random_device-seeded mt19937).
    // Random per process in the reference; a fixed seed here makes runs reproducible.
    void setGaussianTableSeed(uint32_t seed) {
        std::mt19937 gen(seed);
        std::normal_distribution<> d(0, 0.05);                                                   // prediction_stddev_, reference :40
        std::vector<float> table(1000000);
        for (float& v : table) v = (float)d(gen);
        checked(dspmap_set_random_table(h_, table.data(), (uint32_t)table.size()), "dspmap_set_random_table");
    }
    // the instance colour permutation (color_map_int_256_); random per process in the reference, settable here for reproducible output
    void setInstanceColorMap(const uint8_t map256[256]) { for (int i = 0; i < 256; ++i) color_map_int_256_[i] = map256[i]; pushColors(); }
    void setBeyesianMovementParameters(double distance_threshold, double probability_threshold, double increment, double decrement) {   // reference :142-150
        tracker_.setMovementFilter(distance_threshold, probability_threshold, increment, decrement);
    }
    // Which "#if SETTING" flavour of the reference's object layer to follow (default: box key points when the configuration
    // is the ZED2 one, matched key points otherwise); kAlwaysMoving is the SETTING 1 rule.
    void setObjectLayerMode(dsp_adapter::ObjectTracker::Mode mode) { configureTracker(mode); }
    void setObjectLayerSeed(uint32_t seed) { tracker_.setSeed(seed); }
    const dsp_adapter::ObjectTracker& objectTracker() const { return tracker_; }
    void setOccupancyThreshold(float t) { prm_.occupancy_threshold = t; pushParams(); }   // reference :153-156
    void setDepthNoiseModelParameters(float first_order, float zero_order) {                     // reference :161-166
        prm_.depth_noise_first_order = first_order; prm_.depth_noise_zero_order = zero_order;
        pushParams();
    }
    // what an integrator's OWN object layer decided for the coming frame (rigid motion per moving track, tracks to drop);
    // the first call switches the built-in object layer off for good
    void setObjectMoves(const std::vector<dspmap_object_move>& moves, const std::vector<int32_t>& remove_tracks) {
        moves_ = moves; remove_ = remove_tracks; external_object_layer_ = true;
    }

    // reference :170-251.  Unlike the reference the caller's depth/mask Mats are NOT resized in place.
    void update(cv::Mat& depth_value_mat, std::vector<MaskKpts>& ins_seg_result, Eigen::Vector3d& camera_position,
                Eigen::Quaterniond& camera_orientation, pcl::PointCloud<pcl::PointXYZRGB>::Ptr& occupied_point_cloud,
                pcl::PointCloud<pcl::PointXYZRGB>::Ptr& freespace_point_cloud, bool if_get_freespace = false, double time_stamp_double = 0.0) {
        ++step_;                                                                                 // global_time_stamp, reference :173
        if (follow_globals_ && labels_version_ != dspLabelGlobals().version) {   // readObjectInfo ran since the tables were sent
            labels_ = dspLabelGlobals().tables;
            labels_version_ = dspLabelGlobals().version;
            pushLabels();
            pushColors();
        }
        std::vector<dspmap_mask> masks(ins_seg_result.size());
        for (size_t i = 0; i < ins_seg_result.size(); ++i) {
            MaskKpts& m = ins_seg_result[i];
            if (m.label != "static" && m.track_id > labels_.max_movable_instance_id) {           // reference :179-186
                std::printf("Reach the maximum movable object instance id. ID reallocated.\n");
                m.track_id = m.track_id % labels_.max_movable_instance_id;
            }
            dspmap_mask& d = masks[i];
            if (m.mask.type() != CV_8UC1) { std::fprintf(stderr, "mask %zu is not CV_8UC1: frame dropped\n", i); moves_.clear(); return; }
            d.data = m.mask.data; d.row_stride = (int32_t)m.mask.step; d.is_static = m.label == "static";
            d.rows = m.mask.rows; d.cols = m.mask.cols;                                           // checked against the depth image by the engine
            d.track_id = m.track_id; d.label_id = DspLabelTables::getOr0(labels_.label_id, m.label);
            double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {DBL_MIN, DBL_MIN, DBL_MIN};     // pointcloud_tools.h:178-199
            for (const auto& k : m.kpts_current)
                for (int a = 0; a < 3; ++a) { if (k[a] < mn[a]) mn[a] = k[a]; if (k[a] > mx[a]) mx[a] = k[a]; }
            for (int a = 0; a < 3; ++a) { d.bbox_min[a] = mn[a] - 1.0; d.bbox_max[a] = mx[a] + 1.0; }
        }
        const double pos[3] = {camera_position.x(), camera_position.y(), camera_position.z()};
        const double quat[4] = {camera_orientation.w(), camera_orientation.x(), camera_orientation.y(), camera_orientation.z()};
        if (cfg_.consider_instance && !external_object_layer_) runObjectLayer(ins_seg_result, pos, quat, time_stamp_double);   // reference :188-190, 588-704
        if (cfg_.consider_instance)   // pointcloud_tools.h:168 reads g_label_id_map_default with operator[]: an unknown label of a
            for (const MaskKpts& m : ins_seg_result)   // non-static mask is known from the next frame on, with id 0
                if (m.label != "static") labels_.label_id.emplace(m.label, 0);
        appendFloatingTracks(ins_seg_result);
        dspmap_result r;
        if (pipelined_) {
            // setPipelinedOutput: this frame is staged and enqueued, the clouds handed back are those of the PREVIOUS call, whose
            // device->host transfer ran under this frame's kernels (dspmap_update_async / dspmap_wait_result, include/dspmap_c.h)
            uint32_t ticket = 0;
            const int rc = dspmap_update_async(h_, reinterpret_cast<const float*>(depth_value_mat.data), depth_value_mat.rows, depth_value_mat.cols,
                                               (int32_t)(depth_value_mat.step / sizeof(float)), masks.data(), (int32_t)masks.size(), pos, quat,
                                               moves_.data(), (int32_t)moves_.size(), remove_.data(), (int32_t)remove_.size(), if_get_freespace, &ticket);
            moves_.clear();
            if (rc != DSPMAP_OK) std::fprintf(stderr, "dspmap_update_async: %s\n", dspmap_last_error(h_));
            if (ticket) remove_.clear();                      // this frame was accepted (an error then belongs to the frame before it)
            const uint32_t due = pending_ticket_;
            pending_ticket_ = ticket;
            if (!due || dspmap_wait_result(h_, due, &r) != DSPMAP_OK) return;
        } else {
        const int rc = dspmap_update(h_, reinterpret_cast<const float*>(depth_value_mat.data), depth_value_mat.rows, depth_value_mat.cols,
                                     (int32_t)(depth_value_mat.step / sizeof(float)), masks.data(), (int32_t)masks.size(), pos, quat,
                                     moves_.data(), (int32_t)moves_.size(), remove_.data(), (int32_t)remove_.size(), if_get_freespace);
        moves_.clear();
        // reference: no error channel.  A refused frame changed nothing in the engine (arguments are validated before any state
        // is touched), so the tracks that were due for removal stay queued and go with the next frame.
        if (rc != DSPMAP_OK) { std::fprintf(stderr, "dspmap_update: %s\n", dspmap_last_error(h_)); return; }
        remove_.clear();
        if (dspmap_get_result(h_, &r) != DSPMAP_OK) return;
        }
        appendClouds(r, occupied_point_cloud, freespace_point_cloud);
    }

    // Not in the reference's API.  on: update() returns as soon as its frame is enqueued and fills the clouds with the result of
    // the PREVIOUS update() (empty on the first call); flush() hands out the last frame's.  Same map, same clouds, one call later -
    // the host->device copy, the kernels and the device->host copy of consecutive frames overlap (a camera stream, not a request).
    void setPipelinedOutput(bool on) { pipelined_ = on; }
    void flush(pcl::PointCloud<pcl::PointXYZRGB>::Ptr& occupied_point_cloud, pcl::PointCloud<pcl::PointXYZRGB>::Ptr& freespace_point_cloud) {
        dspmap_result r;
        const uint32_t due = pending_ticket_;
        pending_ticket_ = 0;
        if (due && dspmap_wait_result(h_, due, &r) == DSPMAP_OK) appendClouds(r, occupied_point_cloud, freespace_point_cloud);
    }

private:
    void appendClouds(const dspmap_result& r, pcl::PointCloud<pcl::PointXYZRGB>::Ptr& occupied_point_cloud,
                      pcl::PointCloud<pcl::PointXYZRGB>::Ptr& freespace_point_cloud) {
        for (uint32_t i = 0; i < r.count; ++i) {                                                  // reference :1274-1351
            pcl::PointXYZRGB pt;
            pt.x = r.xyz[3 * i]; pt.y = r.xyz[3 * i + 1]; pt.z = r.xyz[3 * i + 2];
            pt.r = r.rgb[3 * i]; pt.g = r.rgb[3 * i + 1]; pt.b = r.rgb[3 * i + 2];               // coloured on the device (dspmap_set_colors)
            occupied_point_cloud->push_back(pt);
        }
        for (uint32_t i = 0; i < r.free_count; ++i) {                                             // reference :1354-1376
            pcl::PointXYZRGB pt;
            pt.x = r.free_xyz[3 * i]; pt.y = r.free_xyz[3 * i + 1]; pt.z = r.free_xyz[3 * i + 2];
            pt.r = 0; pt.g = 255; pt.b = 0;
            freespace_point_cloud->push_back(pt);
        }
    }

public:
    dspmap_t* handle() { return h_; }
    // follow a COMPILED reference where it reads an uninitialised sigma (dspmap_params.uninit_sigma_mode); not in the reference's API
    // pixel lists in the order the reference's breadth-first search visits the voxels in (dspmap_params.reference_iteration_order)
    void setReferenceIterationOrder(bool on) { prm_.reference_iteration_order = on; pushParams(); }
    void followCompiledReferenceSigma(bool on, float value = 0.f) { prm_.uninit_sigma_mode = on; prm_.uninit_sigma_value = value; pushParams(); }

private:
    // every setter of the C ABI can refuse its argument; the reference has no error channel, so: clamp what the engine's
    // tables cannot hold (the reference's own forgetting table has five entries, basic_algorithms.h:33), say so, and treat
    // anything else like the constructor treats a failed create (exit(-1), mc_ring/operations.h:57)
    void checked(int rc, const char* what) {
        if (rc == DSPMAP_OK) return;
        std::fprintf(stderr, "%s: %s\n", what, dspmap_last_error(h_));
        std::exit(-1);
    }
    void pushParams() {
        dspmap_params p = prm_;
        if (p.max_forget_count > 5 || p.max_forget_count < 0) {
            std::fprintf(stderr, "max_forget_count %d clamped to 0..5 (size of the forgetting table)\n", p.max_forget_count);
            p.max_forget_count = p.max_forget_count < 0 ? 0 : 5;
        }
        if (follow_reference_) {      // what a compiled reference reads at its uninitialised-sigma sites: 0 with the depth-noise model, else 0.1
            p.reference_iteration_order = 1;
            p.uninit_sigma_mode = 1; p.uninit_sigma_value = p.consider_depth_noise ? 0.f : 0.1f;
        }
        if (p.nb_ptc_num_per_point < 1) { std::fprintf(stderr, "nb_ptc_num_per_point %d raised to 1\n", p.nb_ptc_num_per_point); p.nb_ptc_num_per_point = 1; }
        checked(dspmap_set_params(h_, &p), "dspmap_set_params");
    }
    void pushLabels() {
        uint16_t v2t[256];
        for (int v = 0; v < 256; ++v) {   // pointcloud_tools.h:137-138 with operator[] default-insert semantics
            auto it = labels_.label_id_reversed.find(v + 1);
            v2t[v] = (uint16_t)(it == labels_.label_id_reversed.end() ? 0 : DspLabelTables::getOr0(labels_.label_to_instance, it->second));
        }
        std::vector<uint16_t> tracks;
        std::vector<uint8_t> labs;
        for (const auto& kv : labels_.instance_to_label) {   // pointcloud_tools.h:277-279
            tracks.push_back((uint16_t)kv.first);
            labs.push_back((uint8_t)DspLabelTables::getOr0(labels_.label_id_static, kv.second));
        }
        checked(dspmap_set_label_luts(h_, v2t, tracks.data(), labs.data(), (int32_t)tracks.size(), (uint16_t)labels_.max_movable_instance_id,
                              (uint16_t)DspLabelTables::getOr0(labels_.label_to_instance, std::string("Sky"))), "dspmap_set_label_luts");   // pointcloud_tools.h:236
    }

    // reference :712-737 ("floating objects"): after the moves and removals of a frame, every track that owns particles in
    // obj_ptc_hash_map (particles register at birth when track <= g_max_movable_object_instance_id, :1160-1162) without being
    // a tracked object (object_tracking_hash_map, filled by objectLevelUpdate only under g_consider_instance, :346-364) is
    // removed with all its particles - e.g. pixels of a movable CLASS of the static mask (instance id -1 -> track 0,
    // pointcloud_tools.h:137-138) live for one frame.  Mirrors just enough of ObjectSet's bookkeeping to name those tracks.
    void appendFloatingTracks(const std::vector<MaskKpts>& ins_seg_result) {
        if (cfg_.consider_instance)
            for (const MaskKpts& m : ins_seg_result)
                if (m.label != "static" && m.track_id <= labels_.max_movable_instance_id) {
                    born_tracks_.insert(m.track_id);
                    if (external_object_layer_) known_objects_.insert(m.track_id);
                }
        if (external_object_layer_) {
            for (const auto& mv : moves_) known_objects_.insert(mv.track_id);
            for (int32_t t : remove_) known_objects_.erase(t);
        }
        for (int32_t t : remove_) born_tracks_.erase(t);
        std::set<int> candidates = born_tracks_;
        for (int v = 0; v < 256; ++v) {   // tracks the static-mask look-up can produce (pointcloud_tools.h:137-138)
            auto it = labels_.label_id_reversed.find(v + 1);
            const int t = it == labels_.label_id_reversed.end() ? 0 : DspLabelTables::getOr0(labels_.label_to_instance, it->second);
            if (t <= labels_.max_movable_instance_id) candidates.insert(t);
        }
        for (int t : candidates) {
            if (external_object_layer_ ? known_objects_.count(t) > 0 : tracker_.isTracked(t)) continue;
            bool listed = false;
            for (int32_t r : remove_) listed |= r == t;
            if (!listed) remove_.push_back(t);
            born_tracks_.erase(t);
            tracker_.forget(t);
        }
    }

    void configureTracker(dsp_adapter::ObjectTracker::Mode mode) {
        const double scale = cfg_.boost_mode ? 0.5 : 1.0;                                        // settings.h:128-143 (float products of the reference)
        dsp_adapter::ObjectTracker::Camera cam{(double)(float)(scale * cfg_.fx), (double)(float)(scale * cfg_.fy), (double)(float)(scale * cfg_.cx),
                                               (double)(float)(scale * cfg_.cy), (int)(scale * cfg_.image_width), (int)(scale * cfg_.image_height)};
        const int biggest = std::max(std::max(cfg_.voxel_bits_x, cfg_.voxel_bits_y), cfg_.voxel_bits_z);   // mc_ring/buffer.h:31-32
        tracker_.configure(mode, cam, (double)(cfg_.voxel_size * (float)(1 << (biggest - 1))) * 1.2);    // reference :356
    }

    // objectLevelUpdate (reference :306-566) on this frame's tracking result, then the head of subObjectLevelUpdate
    // (:588-704): the moves and drops of the frame
    void runObjectLayer(const std::vector<MaskKpts>& ins_seg_result, const double pos[3], const double quat[4], double time_stamp) {
        std::vector<dsp_adapter::ObjectTracker::Observation> seen;
        for (const MaskKpts& m : ins_seg_result) {
            if (m.track_id > labels_.max_movable_instance_id || m.label == "static") continue;   // reference :317
            dsp_adapter::ObjectTracker::Observation o;
            o.track_id = m.track_id;
            auto it = labels_.label_id.find(m.label);
            o.label_known = it != labels_.label_id.end();
            o.label_id = o.label_known ? it->second : 0;
            for (const auto& k : m.kpts_current) o.kpts_current.push_back({k[0], k[1], k[2]});
            for (const auto& k : m.kpts_previous) o.kpts_previous.push_back({k[0], k[1], k[2]});
            seen.push_back(std::move(o));
        }
        tracker_.observe(seen, pos, quat, time_stamp, step_);
        std::vector<dsp_adapter::ObjectTracker::Move> mv;
        std::vector<int> dropped;
        tracker_.plan(step_, max_obersevation_lost_time_, mv, dropped);
        moves_.clear(); remove_.clear();
        for (const auto& m : mv) { dspmap_object_move d; d.track_id = m.track_id; for (int k = 0; k < 16; ++k) d.t[k] = m.t[k]; moves_.push_back(d); }
        for (int t : dropped) remove_.push_back(t);
    }

    void pushColors() {   // reference :1274-1351; which "#if SETTING" block applies follows from the configuration
        dspmap_colors c{};
        c.enabled = 1;
        c.eval_format = if_out_evaluation_format_;
        c.scheme = cfg_.zed2_filters ? 3 : (cfg_.consider_instance ? 1 : 0);
        c.background_label = DspLabelTables::getOr0(labels_.label_id, std::string("Background"));   // :1276, operator[] default-insert
        for (const auto& kv : labels_.label_color)
            if (kv.first >= 0 && kv.first < 256) for (int k = 0; k < 3; ++k) c.label_bgr[kv.first][k] = kv.second[k];
        for (int i = 0; i < 256; ++i) c.instance_map[i] = color_map_int_256_[i];
        checked(dspmap_set_colors(h_, &c), "dspmap_set_colors");
    }

    bool follow_reference_ = false;
    bool pipelined_ = false;            // setPipelinedOutput
    uint32_t pending_ticket_ = 0;       // the frame whose clouds the next update() / flush() hands out
    dspmap_config cfg_;
    DspLabelTables labels_;
    uint8_t color_map_int_256_[256];
    std::set<int> known_objects_, born_tracks_;
    bool follow_globals_ = false;
    unsigned labels_version_ = 0;
    dspmap_params prm_{};
    dspmap_t* h_ = nullptr;
    int max_obersevation_lost_time_ = 5;                                                       // reference :30
    float match_score_threshold_ = 0.5f;
    bool if_out_evaluation_format_ = false;
    std::vector<dspmap_object_move> moves_;
    std::vector<int32_t> remove_;
    dsp_adapter::ObjectTracker tracker_;
    bool external_object_layer_ = false;
    uint32_t step_ = 0;
};
