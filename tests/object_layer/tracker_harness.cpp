// The adapter's host object layer on the scenarios of scenarios.h, printing the same lines as reference_harness.cpp.
// Default build: through the C ABI (include/dspmap_c.h: dspmap_objects_*, libdspmap_b200.so - no GPU is needed for these entry
// points).  -DDIRECT_HEADER: adapter/dsp_object_tracker.h used directly, the way adapter/semantic_dsp_map.h uses it.
#include <array>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>
#ifdef DIRECT_HEADER
#include "../../adapter/dsp_object_tracker.h"
#else
#include "../../include/dspmap_c.h"
#endif
#include "scenarios.h"

int main(int argc, char** argv) {
    const std::string want = argc > 1 ? argv[1] : "";
    for (const scen::Scenario& s : scen::all()) {
        if (s.name != want) continue;
        std::set<std::string> known = {"Vehicle", "person"};   // the label CSV of the pin; pointcloud_tools.h:168 default-inserts the labels it meets
        std::set<int> ids;
        for (const auto& f : s.frames) for (const auto& o : f.objects) ids.insert(o.track_id);
#ifdef DIRECT_HEADER
        using dsp_adapter::ObjectTracker;
        ObjectTracker tr;
        if (s.setting == 3)   // SETTING 3 as shipped: boost mode halves the intrinsics (settings.h:100-119, 136-143); 1.2 x half of the 128-voxel axis
            tr.configure(ObjectTracker::kBoxKeypoints, {(double)(0.5f * 527.8191528320312f), (double)(0.5f * 527.8191528320312f), (double)(0.5f * 633.9357299804688f),
                                                        (double)(0.5f * 366.3338623046875f), 640, 360}, (double)(0.15f * 64) * 1.2);
        else if (s.setting == 2)   // SETTING 2 (settings.h:79-98): Virtual KITTI 2 camera, 0.2 m voxels, 256-voxel axes
            tr.configure(ObjectTracker::kMatchedKeypoints, {725.0087, 725.0087, 620.5, 187.0, 1242, 375}, (double)(0.2f * 128) * 1.2);
        else                       // SETTING 1 (settings.h:55-77): CODA camera, 0.15 m voxels, 256-voxel axes
            tr.configure(ObjectTracker::kAlwaysMoving, {569.8286, 565.4818, 439.2660, 360.5810, 960, 540}, (double)(0.15f * 128) * 1.2);
        tr.setMovementFilter(0.2, 0.3, 0.2, 0.05);
        uint32_t step = 0;
#else
        dspmap_config cfg{};
        if (s.setting == 3) {
            cfg.voxel_bits_x = 7; cfg.voxel_bits_y = 5; cfg.voxel_bits_z = 7; cfg.slot_bits = 2; cfg.voxel_size = 0.15f;
            cfg.fx = cfg.fy = 527.8191528320312f; cfg.cx = 633.9357299804688f; cfg.cy = 366.3338623046875f;
            cfg.image_width = 1280; cfg.image_height = 720; cfg.boost_mode = 1;
        } else if (s.setting == 2) {
            cfg.voxel_bits_x = 8; cfg.voxel_bits_y = 7; cfg.voxel_bits_z = 8; cfg.slot_bits = 3; cfg.voxel_size = 0.2f;
            cfg.fx = cfg.fy = 725.0087f; cfg.cx = 620.5f; cfg.cy = 187.f; cfg.image_width = 1242; cfg.image_height = 375; cfg.boost_mode = 0;
        } else {
            cfg.voxel_bits_x = 8; cfg.voxel_bits_y = 8; cfg.voxel_bits_z = 7; cfg.slot_bits = 2; cfg.voxel_size = 0.15f;
            cfg.fx = 569.8286f; cfg.fy = 565.4818f; cfg.cx = 439.2660f; cfg.cy = 360.5810f; cfg.image_width = 960; cfg.image_height = 540; cfg.boost_mode = 0;
        }
        dspmap_objects_t* tr = nullptr;
        if (dspmap_objects_create(&cfg, s.setting, &tr) != DSPMAP_OK) return 3;   // DSPMAP_OBJECTS_* = the reference's SETTING number
        dspmap_objects_set_movement_filter(tr, 0.2, 0.3, 0.2, 0.05);
#endif
        int t = 0;
        for (const scen::Frame& f : s.frames) {
            std::vector<std::vector<double>> cur(f.objects.size()), prev(f.objects.size());
            for (size_t i = 0; i < f.objects.size(); ++i) {
                for (const auto& k : f.objects[i].kpts) cur[i].insert(cur[i].end(), k.begin(), k.end());
                for (const auto& k : scen::previous(s, (size_t)t, f.objects[i])) prev[i].insert(prev[i].end(), k.begin(), k.end());
            }
#ifdef DIRECT_HEADER
            ++step;
            std::vector<ObjectTracker::Observation> seen;
            for (size_t i = 0; i < f.objects.size(); ++i) {
                const auto& o = f.objects[i];
                ObjectTracker::Observation ob;
                ob.track_id = o.track_id; ob.label_known = known.count(o.label) > 0; ob.label_id = o.label == "Vehicle" ? 14 : (o.label == "person" ? 12 : 0);
                for (size_t k = 0; k + 2 < cur[i].size(); k += 3) { ob.kpts_current.push_back({cur[i][k], cur[i][k + 1], cur[i][k + 2]}); ob.kpts_previous.push_back({prev[i][k], prev[i][k + 1], prev[i][k + 2]}); }
                seen.push_back(ob);
            }
            tr.observe(seen, f.pos, f.quat_wxyz, f.time_stamp, step);
            std::vector<ObjectTracker::Move> moves;
            std::vector<int> dropped;
            tr.plan(step, s.max_lost, moves, dropped);
#else
            std::vector<dspmap_object_observation> seen(f.objects.size());
            for (size_t i = 0; i < f.objects.size(); ++i) {
                const auto& o = f.objects[i];
                seen[i] = dspmap_object_observation{o.track_id, (int32_t)known.count(o.label), o.label == "Vehicle" ? 14 : (o.label == "person" ? 12 : 0),
                                                    (int32_t)(cur[i].size() / 3), cur[i].data(), prev[i].data()};
            }
            dspmap_object_move moves[16];
            int32_t dropped[16], n_moves = 0, n_dropped = 0;
            if (dspmap_objects_step(tr, seen.data(), (int32_t)seen.size(), f.pos, f.quat_wxyz, f.time_stamp, s.max_lost, moves, 16, &n_moves, dropped, 16, &n_dropped) != DSPMAP_OK) return 4;
            for (int32_t i = 0; i < n_moves; ++i) std::fprintf(stderr, "MOVE %d %d %.9g %.9g %.9g\n", t, moves[i].track_id, moves[i].t[3], moves[i].t[7], moves[i].t[11]);
            for (int32_t i = 0; i < n_dropped; ++i) std::fprintf(stderr, "DROP %d %d\n", t, dropped[i]);
#endif
            for (const auto& o : f.objects) known.insert(o.label);
            for (int id : ids) {
                double T[16];
#ifdef DIRECT_HEADER
                const bool tracked = tr.isTracked(id), has_flag = tr.hasMovedFlag(id), moving = tr.isMoving(id);
                const double p = tr.movedProbability(id);
                tr.currentMatrix(id, T);
#else
                int32_t tracked = 0, has_flag = 0, moving = 0;
                double p = 0;
                dspmap_objects_query(tr, id, &tracked, &has_flag, &moving, &p, T);
#endif
                if (!tracked) { std::fprintf(stderr, "OBJ %d %d absent\n", t, id); continue; }
                std::fprintf(stderr, "OBJ %d %d tracked has_flag=%d moving=%d p=%.17g", t, id, (int)has_flag, (int)moving, p);
                if (has_flag) for (int k = 0; k < 16; ++k) std::fprintf(stderr, " %.17g", T[k]);
                std::fprintf(stderr, "\n");
            }
            ++t;
        }
#ifndef DIRECT_HEADER
        dspmap_objects_destroy(tr);
#endif
    }
    return 0;
}
