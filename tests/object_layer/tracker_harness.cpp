// The adapter's host object layer (adapter/dsp_object_tracker.h) on the scenarios of scenarios.h, printing the same lines as
// reference_harness.cpp.  No GPU, no third-party headers.
#include <array>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include "../../adapter/dsp_object_tracker.h"
#include "scenarios.h"

int main(int argc, char** argv) {
    const std::string want = argc > 1 ? argv[1] : "";
    using dsp_adapter::ObjectTracker;
    for (const scen::Scenario& s : scen::all()) {
        if (s.name != want) continue;
        ObjectTracker tr;
        if (s.setting == 3)   // SETTING 3 as shipped: boost mode halves the intrinsics (settings.h:100-119, 136-143); 1.2 x half of the 128-voxel axis
            tr.configure(ObjectTracker::kBoxKeypoints, {(double)(0.5f * 527.8191528320312f), (double)(0.5f * 527.8191528320312f), (double)(0.5f * 633.9357299804688f),
                                                        (double)(0.5f * 366.3338623046875f), 640, 360}, (double)(0.15f * 64) * 1.2);
        else                  // SETTING 2 (settings.h:79-98): Virtual KITTI 2 camera, 0.2 m voxels, 256-voxel axes
            tr.configure(ObjectTracker::kMatchedKeypoints, {725.0087, 725.0087, 620.5, 187.0, 1242, 375}, (double)(0.2f * 128) * 1.2);
        tr.setMovementFilter(0.2, 0.3, 0.2, 0.05);
        std::set<std::string> known = {"Vehicle", "person"};   // the label CSV of the pin; pointcloud_tools.h:168 default-inserts the labels it meets
        std::set<int> ids;
        for (const auto& f : s.frames) for (const auto& o : f.objects) ids.insert(o.track_id);
        uint32_t step = 0;
        int t = 0;
        for (const scen::Frame& f : s.frames) {
            ++step;
            std::vector<ObjectTracker::Observation> seen;
            for (const auto& o : f.objects) {
                ObjectTracker::Observation ob;
                ob.track_id = o.track_id; ob.label_known = known.count(o.label) > 0; ob.label_id = o.label == "Vehicle" ? 14 : (o.label == "person" ? 12 : 0);
                for (const auto& k : o.kpts) ob.kpts_current.push_back({k[0], k[1], k[2]});
                for (const auto& k : scen::previous(s, (size_t)t, o)) ob.kpts_previous.push_back({k[0], k[1], k[2]});
                seen.push_back(ob);
            }
            tr.observe(seen, f.pos, f.quat_wxyz, f.time_stamp, step);
            for (const auto& o : f.objects) known.insert(o.label);
            std::vector<ObjectTracker::Move> moves;
            std::vector<int> dropped;
            tr.plan(step, s.max_lost, moves, dropped);
            for (int id : ids) {
                if (!tr.isTracked(id)) { std::fprintf(stderr, "OBJ %d %d absent\n", t, id); continue; }
                double T[16];
                const bool hasT = tr.currentMatrix(id, T);
                std::fprintf(stderr, "OBJ %d %d tracked has_flag=%d moving=%d p=%.17g", t, id, (int)tr.hasMovedFlag(id), (int)tr.isMoving(id), tr.movedProbability(id));
                if (hasT) for (int k = 0; k < 16; ++k) std::fprintf(stderr, " %.17g", T[k]);
                std::fprintf(stderr, "\n");
            }
            ++t;
        }
    }
    return 0;
}
