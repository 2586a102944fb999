// The reference's OWN object layer (objectLevelUpdate, ObjectSet, Kabsch / RANSAC; include/semantic_dsp_map.h:306-566,
// include/object_layer.h, include/utils/basic_algorithms.h:56-195) driven through SemanticDSPMap::update on empty depth
// images, printing after every frame what it decided for every object: tracked, moving, moved probability, the rigid
// matrix.  TEST INFRASTRUCTURE, build container only (tests/golden/make_object_layer_golden.py); third-party headers from
// oracle/ref_stubs.  One process per scenario: the reference's state is global / function-static.
#include <Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <pcl/point_types.h>
#include <array>
#include <fstream>
#include <mutex>
#include <set>
#include <random>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#define private public
#include "semantic_dsp_map.h"
#undef private
#include "scenarios.h"

int main(int argc, char** argv) {
    const std::string want = argc > 1 ? argv[1] : "";
    const std::string csv_path = argc > 2 ? argv[2] : "labels.csv";
    for (const scen::Scenario& s : scen::all()) {
        if (s.name != want) continue;
        if (s.setting != SETTING) { std::fprintf(stderr, "scenario %s is for SETTING %d\n", s.name.c_str(), s.setting); return 2; }
        static SemanticDSPMap map;
        ObjectInfoHandler h;
        h.readObjectInfo(csv_path);
        map.setMapParameters(0.8f, 0.2f, 1, 0.15f, s.max_lost, 1.f, 5, 0.6f, 0.5f);
        map.setMapOptions(true, false);
        map.setVisualizeOptions(false, true);
        map.setBeyesianMovementParameters(0.2, 0.3, 0.2, 0.05);
        map.setDepthNoiseModelParameters(0.02f, 0.3f);
        std::set<int> ids;
        for (const auto& f : s.frames) for (const auto& o : f.objects) ids.insert(o.track_id);
        int t = 0;
        for (const scen::Frame& f : s.frames) {
            cv::Mat depth = cv::Mat::zeros(g_image_height_set, g_image_width_set, CV_32FC1);
            std::vector<MaskKpts> seg(1 + f.objects.size());
            seg[0].track_id = 65535; seg[0].label = "static"; seg[0].mask = cv::Mat::zeros(g_image_height_set, g_image_width_set, CV_8UC1);
            for (size_t i = 0; i < f.objects.size(); ++i) {
                seg[i + 1].track_id = f.objects[i].track_id; seg[i + 1].label = f.objects[i].label;
                seg[i + 1].mask = cv::Mat::zeros(g_image_height_set, g_image_width_set, CV_8UC1);
                for (const auto& k : f.objects[i].kpts) seg[i + 1].kpts_current.push_back(Eigen::Vector3d(k[0], k[1], k[2]));
                for (const auto& k : scen::previous(s, (size_t)t, f.objects[i])) seg[i + 1].kpts_previous.push_back(Eigen::Vector3d(k[0], k[1], k[2]));
            }
            Eigen::Vector3d pos(f.pos[0], f.pos[1], f.pos[2]);
            Eigen::Quaterniond q(f.quat_wxyz[0], f.quat_wxyz[1], f.quat_wxyz[2], f.quat_wxyz[3]);
            pcl::PointCloud<pcl::PointXYZRGB>::Ptr occ(new pcl::PointCloud<pcl::PointXYZRGB>), fre(new pcl::PointCloud<pcl::PointXYZRGB>);
            map.update(depth, seg, pos, q, occ, fre, false, f.time_stamp);
            for (int id : ids) {
                auto it = map.object_set_.object_tracking_hash_map.find(id);
                if (it == map.object_set_.object_tracking_hash_map.end()) { std::fprintf(stderr, "OBJ %d %d absent\n", t, id); continue; }
                const auto& o = it->second.object;
                const bool has = !o.rigidbody_moved_vec.empty();
                std::fprintf(stderr, "OBJ %d %d tracked has_flag=%d moving=%d p=%.17g", t, id, (int)has, (int)(has && o.rigidbody_moved_vec[0]), o.moved_probability);
                if (!o.rigidbody_tmatrix_vec.empty())
                    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) std::fprintf(stderr, " %.17g", o.rigidbody_tmatrix_vec[0](r, c));
                std::fprintf(stderr, "\n");
            }
            ++t;
        }
    }
    return 0;
}
