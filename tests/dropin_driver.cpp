// The SAME caller source compiled twice: against the reference's own include/semantic_dsp_map.h (-DUSE_REFERENCE, on the
// CPU, with the stand-in third-party headers of oracle/ref_stubs) and against adapter/semantic_dsp_map.h (the B200 engine).
// It drives the class the way src/mapping.cpp does (readObjectInfo :89-93, the five setters :135-140 with the values of
// cfg/options_zed2.yaml, update() :273) on procedurally generated 1280x720 frames - the configuration the reference
// ships (SETTING 3) - and writes every frame's occupied cloud (x, y, z, r, g, b) to a file.  The one place the two builds
// differ is the header they include: the adapter's default constructor takes the shipped SETTING 3 configuration.
// Scene "static" (default): objects carry two key points only, so the reference's object layer registers them
// (semantic_dsp_map.h:346-364) but never estimates a motion (:383 needs four).
// Scene "moving": the box slides 0.125 m a frame along +x and reports four 3-D box key points, so the whole object layer
// runs (objectLevelUpdate :306-566: Kabsch on the key points, the moving-or-not filter, the stored key frame); after
// frame 6 the detector loses the box, which then coasts on its estimated velocity (:545-563) until
// max_obersevation_lost_time (3 here) steps without an observation drop it together with its particles (:597-604, 694-704).
#ifdef USE_REFERENCE
#include <Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <pcl/point_types.h>
#include <fstream>
#include <mutex>
#include <random>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include "semantic_dsp_map.h"
#else
#include "../adapter/semantic_dsp_map.h"
#include <fstream>
#include <sstream>
#endif

#include <cmath>
#include <cstdio>
#include <cstring>

static const char* kCsv =
    "label_id,label,instance_id(-1 means not fixed),color_b,color_g,color_r\n"
    "0,background,65535,0,0,0\n1,road,65534,128,64,128\n2,sidewalk,65533,232,35,244\n3,building,65532,70,70,70\n4,wall,65531,156,102,102\n"
    "5,fence,65530,153,153,190\n6,pole,65529,153,153,153\n7,traffic_light,65528,30,170,250\n8,traffic_sign,65527,0,220,220\n"
    "9,vegetation,65526,35,142,107\n10,terrain,65525,152,251,152\n11,sky,65524,180,130,70\n12,person,-1,60,20,220\n13,rider,65522,0,0,255\n"
    "14,Vehicle,-1,142,0,0\n15,truck,65521,70,0,0\n";

int main(int argc, char** argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 4;
    const char* out_path = argc > 2 ? argv[2] : "dropin_out.bin";
    const bool moving = argc > 3 && std::strcmp(argv[3], "moving") == 0;
    const std::string csv_path = std::string(out_path) + ".csv";
    { std::ofstream f(csv_path); f << kCsv; }

    SemanticDSPMap dsp_map_;                                  // src/mapping.cpp:71: a default-constructed member ...
    ObjectInfoHandler object_info_handler;                    // ... and only then the label CSV, :89-93
    object_info_handler.readObjectInfo(csv_path);
    // src/mapping.cpp:135-140 with cfg/options_zed2.yaml
    dsp_map_.setMapParameters(0.8f, 0.2f, 1, 0.15f, moving ? 3 : 20, 1.f, 5, 0.6f, 0.5f);
    dsp_map_.setMapOptions(true, false);
    dsp_map_.setVisualizeOptions(false, true);
    dsp_map_.setBeyesianMovementParameters(0.2, 0.3, 0.2, 0.05);
    dsp_map_.setDepthNoiseModelParameters(0.02f, 0.3f);

    FILE* out = std::fopen(out_path, "wb");
    const int H = 720, W = 1280;
    const float fx = 527.8191528320312f, cx = 633.9357299804688f, cy = 366.3338623046875f;
    for (int t = 0; t < frames; ++t) {
        const double cam_z = moving ? 0.125 * t : 0.3 * t;
        const float box_x = moving ? 0.125f * (float)t : 0.f;                  // exact in float32
        const bool box_seen = !moving || t <= 6;
        cv::Mat depth(H, W, CV_32FC1);
        cv::Mat cls(H, W, CV_8UC1), car(H, W, CV_8UC1);
        // ground plane 1.5 m below the camera, a wall 7 m ahead of the start, a 1.5 x 1.5 m box face at 4 m
        const float wall = (float)(7.0 - cam_z), box = (float)(4.0 - cam_z);
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float d = wall;
                unsigned char c = 3;                                       // wall -> label 4
                const float ry = ((float)i - cy) / fx, rx = ((float)j - cx) / fx;
                if (ry > 1e-6f && 1.5f / ry < d) { d = 1.5f / ry; c = 0; }  // road
                unsigned char m = 0;
                const float bx = rx * box - box_x, by = ry * box;
                if (box > 0.5f && bx > -0.2f && bx < 1.3f && by > 0.0f && by < 1.5f && box < d) {
                    d = box;
                    if (box_seen) { c = 13; m = 255; } else { c = 3; }       // a lost detection leaves the pixels to the static classes
                }
                if ((i * 131 + j * 17 + t * 7) % 97 == 0) d = 16.f;         // a sprinkle of far pixels
                depth.at<float>(i, j) = d; cls.at<uchar>(i, j) = c; car.at<uchar>(i, j) = m;
            }
        std::vector<MaskKpts> ins_seg(box_seen ? 2 : 1);
        ins_seg[0].track_id = 65535; ins_seg[0].label = "static"; ins_seg[0].mask = cls;
        if (box_seen) {
            ins_seg[1].track_id = 3; ins_seg[1].label = "Vehicle"; ins_seg[1].mask = car;
            const double bxd = box_x;
            ins_seg[1].kpts_current.push_back(Eigen::Vector3d(bxd - 0.25, 0.0, 4.0));   // exact in float32: the wire format carries float32 key points
            if (moving) ins_seg[1].kpts_current.push_back(Eigen::Vector3d(bxd - 0.125, 0.0, 4.0));   // "width" 0.125 m: the 0.2 m movement threshold applies (:451-457)
            ins_seg[1].kpts_current.push_back(Eigen::Vector3d(bxd + 1.25, 1.5, 5.0));
            if (moving) ins_seg[1].kpts_current.push_back(Eigen::Vector3d(bxd - 0.25, 0.5, 4.5));    // the last key point decides "out of view" (:419-422)
            ins_seg[1].kpts_previous = ins_seg[1].kpts_current;
        }
        Eigen::Vector3d pos(0.0, 0.0, cam_z);
        Eigen::Quaterniond q(1.0, 0.0, 0.0, 0.0);
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr occ(new pcl::PointCloud<pcl::PointXYZRGB>), fre(new pcl::PointCloud<pcl::PointXYZRGB>);
        dsp_map_.update(depth, ins_seg, pos, q, occ, fre, false, 0.1 * t);
        const uint32_t n = (uint32_t)occ->points.size();
        std::fwrite(&n, 4, 1, out);
        for (const auto& p : occ->points) {
            const float xyz[3] = {p.x, p.y, p.z};
            const unsigned char rgb[3] = {p.r, p.g, p.b};
            std::fwrite(xyz, 4, 3, out);
            std::fwrite(rgb, 1, 3, out);
        }
        std::fprintf(stderr, "frame %d: %u occupied\n", t, n);
    }
    std::fclose(out);
    return 0;
}
