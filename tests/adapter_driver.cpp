// Drives adapter/semantic_dsp_map.h the way src/mapping.cpp does (mapping.cpp:135-140, 273), against the stand-ins for
// Eigen / OpenCV / PCL in oracle/ref_stubs (the same ones the reference itself is compiled against for the parity pin).  Prints one line per frame: "<frame> <n_occupied> <fnv1a of the occupied voxel positions>".
#include "../adapter/semantic_dsp_map.h"

#include <cmath>
#include <cstring>

int main(int argc, char** argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 3;
    const bool pipelined = argc > 2 && std::strcmp(argv[2], "pipelined") == 0;      // clouds come out one call later (setPipelinedOutput)
    dspmap_config cfg{};
    cfg.voxel_bits_x = 6; cfg.voxel_bits_y = 5; cfg.voxel_bits_z = 6; cfg.slot_bits = 3; cfg.voxel_size = 0.15f;
    cfg.fx = cfg.fy = 66.f; cfg.cx = 79.5f; cfg.cy = 47.5f; cfg.image_width = 160; cfg.image_height = 96; cfg.boost_mode = 0;
    cfg.depth_min = 0.3f; cfg.depth_max = 15.f; cfg.consider_instance = 1; cfg.zed2_filters = 0; cfg.shard_count = 1;
    DspLabelTables labels;
    labels.addRow(0, "background", 65535); labels.addRow(1, "road", 65534); labels.addRow(14, "Vehicle", -1);
    SemanticDSPMap dsp_map_(cfg, labels);
    dsp_map_.setMapParameters(0.8f, 0.2f, 1, 0.15f, 5, 1.0f, 5, 0.5f, 0.5f);
    dsp_map_.setMapOptions(true, false);
    dsp_map_.setVisualizeOptions(false, true);
    dsp_map_.setBeyesianMovementParameters(0.1, 0.9, 0.3, 0.1);
    dsp_map_.setDepthNoiseModelParameters(0.02f, 0.3f);
    dsp_map_.setPipelinedOutput(pipelined);
    auto print = [](int t, const pcl::PointCloud<pcl::PointXYZRGB>::Ptr& occ) {
        uint64_t h = 1469598103934665603ull;
        for (const auto& p : occ->points) { uint32_t b[3]; std::memcpy(b, &p.x, 12); for (uint32_t w : b) { h ^= w; h *= 1099511628211ull; } h ^= p.r; h *= 1099511628211ull; }
        std::printf("%d %zu %016llx\n", t, occ->size(), (unsigned long long)h);
    };
    for (int t = 0; t < frames; ++t) {
        cv::Mat depth(96, 160, CV_32FC1);
        float* d = reinterpret_cast<float*>(depth.data);
        for (int i = 0; i < 96; ++i)
            for (int j = 0; j < 160; ++j) d[i * 160 + j] = i > 52 ? 1.5f * 66.f / (float)(i - 47) : 3.5f;   // ground plane + wall
        std::vector<MaskKpts> ins_seg;
        MaskKpts st; st.track_id = 65535; st.label = "static"; st.mask = cv::Mat(96, 160, CV_8UC1);
        std::memset(st.mask.data, 0, 96 * 160);
        ins_seg.push_back(st);
        MaskKpts car; car.track_id = 3; car.label = "Vehicle"; car.mask = cv::Mat(96, 160, CV_8UC1);
        std::memset(car.mask.data, 0, 96 * 160);
        for (int i = 30; i < 50; ++i) for (int j = 60; j < 90; ++j) car.mask.data[i * 160 + j] = 255;
        ins_seg.push_back(car);
        Eigen::Vector3d pos(0.0, 0.0, 0.1 * t);
        Eigen::Quaterniond q(1.0, 0.0, 0.0, 0.0);
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr occ(new pcl::PointCloud<pcl::PointXYZRGB>), fre(new pcl::PointCloud<pcl::PointXYZRGB>);
        dsp_map_.update(depth, ins_seg, pos, q, occ, fre, false, 0.1 * t);
        if (!pipelined) print(t, occ);
        else if (t > 0) print(t - 1, occ);          // the previous frame's clouds
    }
    if (pipelined) {
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr occ(new pcl::PointCloud<pcl::PointXYZRGB>), fre(new pcl::PointCloud<pcl::PointXYZRGB>);
        dsp_map_.flush(occ, fre);
        print(frames - 1, occ);
    }
    return 0;
}
