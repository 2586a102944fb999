// C ABI of the host object layer (include/dspmap_c.h, "object layer"): a thin shell over adapter/dsp_object_tracker.h.
// Host only - these entry points work without a GPU; they decide, they do not touch particles.
#include <new>
#include <vector>

#include "../../adapter/dsp_object_tracker.h"
#include "../../include/dspmap_c.h"

struct dspmap_objects {
    dsp_adapter::ObjectTracker tracker;
    uint32_t step = 0;      // global_time_stamp, semantic_dsp_map.h:173
    std::vector<dsp_adapter::ObjectTracker::Move> last_moves;   // the plan of the last step (dspmap_objects_last_plan)
    std::vector<int> last_dropped;
};

static int copy_plan(const dspmap_objects* o, dspmap_object_move* moves, int32_t moves_cap, int32_t* n_moves, int32_t* dropped, int32_t dropped_cap,
                     int32_t* n_dropped) {
    *n_moves = (int32_t)o->last_moves.size();
    *n_dropped = (int32_t)o->last_dropped.size();
    for (int32_t i = 0; i < *n_moves && i < moves_cap; ++i) { moves[i].track_id = o->last_moves[(size_t)i].track_id; for (int k = 0; k < 16; ++k) moves[i].t[k] = o->last_moves[(size_t)i].t[k]; }
    for (int32_t i = 0; i < *n_dropped && i < dropped_cap; ++i) dropped[i] = o->last_dropped[(size_t)i];
    return (*n_moves > moves_cap || *n_dropped > dropped_cap) ? DSPMAP_ERR_CAPACITY : DSPMAP_OK;
}

extern "C" {

int dspmap_objects_create(const dspmap_config* cfg, int32_t mode, dspmap_objects_t** out) {
    if (!cfg || !out || mode < 0 || mode > 3) return DSPMAP_ERR_INVALID_ARG;
    if (cfg->voxel_bits_x < 1 || cfg->voxel_bits_y < 1 || cfg->voxel_bits_z < 1 || !(cfg->voxel_size > 0.f)) return DSPMAP_ERR_CONFIG;
    dspmap_objects* o = new (std::nothrow) dspmap_objects();
    if (!o) return DSPMAP_ERR_INVALID_ARG;
    const float scale = cfg->boost_mode ? 0.5f : 1.0f;                                          // settings.h:128-143 (float products)
    dsp_adapter::ObjectTracker::Camera cam{(double)(scale * cfg->fx), (double)(scale * cfg->fy), (double)(scale * cfg->cx), (double)(scale * cfg->cy),
                                           (int)(scale * (float)cfg->image_width), (int)(scale * (float)cfg->image_height)};
    int biggest = cfg->voxel_bits_x > cfg->voxel_bits_y ? cfg->voxel_bits_x : cfg->voxel_bits_y;   // mc_ring/buffer.h:31-32
    if (cfg->voxel_bits_z > biggest) biggest = cfg->voxel_bits_z;
    o->tracker.configure((dsp_adapter::ObjectTracker::Mode)mode, cam, (double)(cfg->voxel_size * (float)(1 << (biggest - 1))) * 1.2);   // semantic_dsp_map.h:356
    *out = o;
    return DSPMAP_OK;
}

void dspmap_objects_destroy(dspmap_objects_t* o) { delete o; }

int dspmap_objects_clear(dspmap_objects_t* o) {
    if (!o) return DSPMAP_ERR_INVALID_ARG;
    o->tracker.clear();
    o->last_moves.clear(); o->last_dropped.clear();
    o->step = 0;
    return DSPMAP_OK;
}

int dspmap_objects_set_movement_filter(dspmap_objects_t* o, double distance_threshold, double probability_threshold, double increment, double decrement) {
    if (!o) return DSPMAP_ERR_INVALID_ARG;
    o->tracker.setMovementFilter(distance_threshold, probability_threshold, increment, decrement);
    return DSPMAP_OK;
}

int dspmap_objects_set_seed(dspmap_objects_t* o, uint32_t seed) {
    if (!o) return DSPMAP_ERR_INVALID_ARG;
    o->tracker.setSeed(seed);
    return DSPMAP_OK;
}

int dspmap_objects_step(dspmap_objects_t* o, const dspmap_object_observation* observations, int32_t n_observations, const double position[3],
                        const double orientation_wxyz[4], double time_stamp, int32_t max_lost, dspmap_object_move* moves, int32_t moves_cap, int32_t* n_moves,
                        int32_t* dropped, int32_t dropped_cap, int32_t* n_dropped) {
    if (!o || !position || !orientation_wxyz || !n_moves || !n_dropped || n_observations < 0 || (n_observations > 0 && !observations) || moves_cap < 0 ||
        dropped_cap < 0 || (moves_cap > 0 && !moves) || (dropped_cap > 0 && !dropped))
        return DSPMAP_ERR_INVALID_ARG;
    std::vector<dsp_adapter::ObjectTracker::Observation> seen((size_t)n_observations);
    for (int32_t i = 0; i < n_observations; ++i) {
        const dspmap_object_observation& in = observations[i];
        if (in.n_kpts < 0 || (in.n_kpts > 0 && !in.kpts_current)) return DSPMAP_ERR_INVALID_ARG;
        auto& ob = seen[(size_t)i];
        ob.track_id = in.track_id; ob.label_known = in.label_known != 0; ob.label_id = in.label_id;
        const double* prev = in.kpts_previous ? in.kpts_previous : in.kpts_current;
        for (int32_t k = 0; k < in.n_kpts; ++k) {
            ob.kpts_current.push_back({in.kpts_current[3 * k], in.kpts_current[3 * k + 1], in.kpts_current[3 * k + 2]});
            ob.kpts_previous.push_back({prev[3 * k], prev[3 * k + 1], prev[3 * k + 2]});
        }
    }
    ++o->step;
    o->tracker.observe(seen, position, orientation_wxyz, time_stamp, o->step);
    o->last_moves.clear(); o->last_dropped.clear();
    o->tracker.plan(o->step, max_lost, o->last_moves, o->last_dropped);
    return copy_plan(o, moves, moves_cap, n_moves, dropped, dropped_cap, n_dropped);
}

int dspmap_objects_last_plan(dspmap_objects_t* o, dspmap_object_move* moves, int32_t moves_cap, int32_t* n_moves, int32_t* dropped, int32_t dropped_cap,
                             int32_t* n_dropped) {
    if (!o || !n_moves || !n_dropped || moves_cap < 0 || dropped_cap < 0 || (moves_cap > 0 && !moves) || (dropped_cap > 0 && !dropped)) return DSPMAP_ERR_INVALID_ARG;
    return copy_plan(o, moves, moves_cap, n_moves, dropped, dropped_cap, n_dropped);
}

int dspmap_objects_query(dspmap_objects_t* o, int32_t track_id, int32_t* tracked, int32_t* has_flag, int32_t* moving, double* moved_probability, double* t16) {
    if (!o) return DSPMAP_ERR_INVALID_ARG;
    if (tracked) *tracked = o->tracker.isTracked(track_id);
    if (has_flag) *has_flag = o->tracker.hasMovedFlag(track_id);
    if (moving) *moving = o->tracker.isMoving(track_id);
    if (moved_probability) *moved_probability = o->tracker.movedProbability(track_id);
    if (t16) o->tracker.currentMatrix(track_id, t16);
    return DSPMAP_OK;
}

int dspmap_objects_forget(dspmap_objects_t* o, int32_t track_id) {
    if (!o) return DSPMAP_ERR_INVALID_ARG;
    o->tracker.forget(track_id);
    return DSPMAP_OK;
}

}  // extern "C"
