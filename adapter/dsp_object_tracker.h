// dsp_object_tracker.h — the host-side object layer of the drop-in adapter (host edition).
//
// What the reference does per frame with a handful of tracked objects before it touches a particle:
//   objectLevelUpdate              include/semantic_dsp_map.h:306-566   (register / update / predict objects)
//   estimateTransformation[RANSAC] include/utils/basic_algorithms.h:56-195 (Kabsch on key points, 3-point RANSAC)
//   ObjectSet / ObjectTransformations / MotionEstimation   include/object_layer.h:58-581 (moving-or-not filter,
//                                   constant-velocity prediction for unobserved objects)
//   head of subObjectLevelUpdate   include/semantic_dsp_map.h:588-737   (which objects move by which matrix, which are dropped)
// Its per-frame product is tiny: one rigid 4x4 per moving object and a list of tracks to drop.  Here that product is
// what goes into dspmap_update() as `moves` / `remove_tracks`; the particles themselves never leave the GPU.
//
// Written on plain doubles (no Eigen types inside) with a fixed evaluation order.  Behaviour notes:
//  * The reference seeds its RANSAC from std::random_device (basic_algorithms.h:112-114); for a rigid motion every
//    3-point sample gives the same matrix, so a fixed-seed generator is used (setSeed) and runs are reproducible.
//  * Iteration over the reference's unordered_map of objects is implementation-defined; moves are issued in ascending track id.
//  * Template re-matching (to_match_with_previous with getFlagUseTemplateMatching(), disabled by default) is out of scope (8 f4).
//  * Where the reference indexes stored key points out of bounds (an object first seen with fewer than four key points, or
//    one reported with more than four, semantic_dsp_map.h:442-445) the tracker re-records the key points and skips the update.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <map>
#include <random>
#include <set>
#include <vector>

namespace dsp_adapter {

using Vec3 = std::array<double, 3>;
struct Rigid {                       // x -> R x + t, R row-major
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double t[3] = {0, 0, 0};
    Vec3 apply(const Vec3& p) const {
        return {(R[0] * p[0] + R[1] * p[1]) + R[2] * p[2] + t[0], (R[3] * p[0] + R[4] * p[1]) + R[5] * p[2] + t[1], (R[6] * p[0] + R[7] * p[1]) + R[8] * p[2] + t[2]};
    }
};
inline Vec3 sub(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline double norm(const Vec3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Eigen-decomposition of a symmetric 3x3 (cyclic Jacobi); eigenvectors in the COLUMNS of `vec`, eigenvalues descending.
inline void symmetricEigen3(const double a_in[9], double vec[9], double val[3]) {
    double a[9];
    for (int i = 0; i < 9; ++i) { a[i] = a_in[i]; vec[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = std::fabs(a[1]) + std::fabs(a[2]) + std::fabs(a[5]);
        const double diag = std::fabs(a[0]) + std::fabs(a[4]) + std::fabs(a[8]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = a[3 * p + q];
                if (apq == 0.0) continue;
                const double theta = (a[3 * q + q] - a[3 * p + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {   // A <- A J
                    const double akp = a[3 * k + p], akq = a[3 * k + q];
                    a[3 * k + p] = c * akp - s * akq; a[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {   // A <- J^T A
                    const double apk = a[3 * p + k], aqk = a[3 * q + k];
                    a[3 * p + k] = c * apk - s * aqk; a[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = vec[3 * k + p], vkq = vec[3 * k + q];
                    vec[3 * k + p] = c * vkp - s * vkq; vec[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    int order[3] = {0, 1, 2};
    const double d[3] = {a[0], a[4], a[8]};
    std::sort(order, order + 3, [&](int x, int y) { return d[x] > d[y]; });
    double sorted[9];
    for (int k = 0; k < 3; ++k) { val[k] = d[order[k]]; for (int i = 0; i < 3; ++i) sorted[3 * i + k] = vec[3 * i + order[k]]; }
    for (int i = 0; i < 9; ++i) vec[i] = sorted[i];
}

// Kabsch: the rotation + translation that best maps P onto Q (basic_algorithms.h:56-97: R = V U^T of the SVD of
// H = Pc Qc^T, third column of V negated when det R < 0, t = centroid_Q - R centroid_P).  V = eigenvectors of H^T H,
// U = H V / sigma; a vanishing third singular value (three sample points are always coplanar) leaves u3 = u1 x u2.
inline Rigid kabsch(const std::vector<Vec3>& P, const std::vector<Vec3>& Q) {
    const size_t n = P.size();
    Vec3 cp{0, 0, 0}, cq{0, 0, 0};
    for (size_t j = 0; j < n; ++j) for (int a = 0; a < 3; ++a) { cp[a] += P[j][a]; cq[a] += Q[j][a]; }
    for (int a = 0; a < 3; ++a) { cp[a] /= (double)n; cq[a] /= (double)n; }
    double H[9] = {0};
    for (size_t j = 0; j < n; ++j) {
        const Vec3 p = sub(P[j], cp), q = sub(Q[j], cq);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] += p[r] * q[c];
    }
    double HtH[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += H[3 * k + r] * H[3 * k + c]; HtH[3 * r + c] = s; }
    double V[9], ev[3];
    symmetricEigen3(HtH, V, ev);
    double U[9];
    const double s0 = std::sqrt(std::max(ev[0], 0.0));
    int have = 0;
    for (int k = 0; k < 3; ++k) {
        const double sk = std::sqrt(std::max(ev[k], 0.0));
        if (sk > 1e-9 * s0 && sk > 0) {
            double u[3];
            for (int r = 0; r < 3; ++r) u[r] = ((H[3 * r] * V[k] + H[3 * r + 1] * V[3 + k]) + H[3 * r + 2] * V[6 + k]) / sk;
            for (int c2 = 0; c2 < k; ++c2) {   // one Gram-Schmidt pass: H V is orthogonal up to rounding
                double dpr = 0; for (int r = 0; r < 3; ++r) dpr += u[r] * U[3 * r + c2];
                for (int r = 0; r < 3; ++r) u[r] -= dpr * U[3 * r + c2];
            }
            const double nu = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            for (int r = 0; r < 3; ++r) U[3 * r + k] = u[r] / nu;
            have = k + 1;
        } else break;
    }
    if (have == 2) {   // u3 = u1 x u2
        U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    } else if (have < 2) {   // collinear or coincident sample: no rotation can be recovered; the reference's result is arbitrary here
        Rigid T;
        for (int a = 0; a < 3; ++a) T.t[a] = cq[a] - cp[a];
        return T;
    }
    auto build = [&](double R[9]) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[3 * r + c] = (V[3 * r] * U[3 * c] + V[3 * r + 1] * U[3 * c + 1]) + V[3 * r + 2] * U[3 * c + 2]; };
    Rigid T;
    build(T.R);
    const double det = T.R[0] * (T.R[4] * T.R[8] - T.R[5] * T.R[7]) - T.R[1] * (T.R[3] * T.R[8] - T.R[5] * T.R[6]) + T.R[2] * (T.R[3] * T.R[7] - T.R[4] * T.R[6]);
    if (det < 0) { for (int r = 0; r < 3; ++r) V[3 * r + 2] = -V[3 * r + 2]; build(T.R); }
    for (int a = 0; a < 3; ++a) T.t[a] = cq[a] - ((T.R[3 * a] * cp[0] + T.R[3 * a + 1] * cp[1]) + T.R[3 * a + 2] * cp[2]);
    return T;
}

// basic_algorithms.h:107-190.  Returns the mean squared error over the inliers.
inline double kabschRansac(const std::vector<Vec3>& P, const std::vector<Vec3>& Q, Rigid& result, std::vector<int>& inliers, int max_iterations, double threshold,
                           bool recompute_with_inliers, std::mt19937& rng) {
    const int N = (int)P.size();
    int max_inliers = -1;
    Rigid best;
    std::uniform_int_distribution<int> dist(0, N - 1);
    for (int it = 0; it < max_iterations; ++it) {
        int idx[3], got = 0;
        while (got < 3) {
            const int r = dist(rng);
            bool dup = false;
            for (int k = 0; k < got; ++k) dup |= idx[k] == r;
            if (!dup) idx[got++] = r;
        }
        const Rigid T = kabsch({P[idx[0]], P[idx[1]], P[idx[2]]}, {Q[idx[0]], Q[idx[1]], Q[idx[2]]});
        std::vector<int> in;
        for (int j = 0; j < N; ++j) if (norm(sub(T.apply(P[j]), Q[j])) < threshold) in.push_back(j);
        if ((int)in.size() > max_inliers) { max_inliers = (int)in.size(); best = T; inliers = in; }
        if (max_inliers > 0.9 * N) break;
    }
    if (recompute_with_inliers && inliers.size() >= 3) {
        std::vector<Vec3> Pi, Qi;
        for (int j : inliers) { Pi.push_back(P[j]); Qi.push_back(Q[j]); }
        result = kabsch(Pi, Qi);
    } else {
        result = best;
    }
    double err_in = 0;
    for (int j : inliers) { const Vec3 d = sub(result.apply(P[j]), Q[j]); err_in += d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; }
    return err_in / (double)inliers.size();
}

class ObjectTracker {
public:
    // which "#if SETTING" flavour of the reference's object layer applies (semantic_dsp_map.h:339-343, object_layer.h:479-509)
    enum Mode {
        kStatic = 0,            // SETTING 0: objects are never moving
        kAlwaysMoving = 1,      // SETTING 1: matched key points (>= 5, kpts_previous -> kpts_current), every updated object moves
        kMatchedKeypoints = 2,  // SETTING 2: matched key points, moving-or-not filter on the reference point's displacement
        kBoxKeypoints = 3       // SETTING 3: four 3-D box key points per object, compared with the ones stored from the last frame
    };
    struct Camera { double fx, fy, cx, cy; int width, height; };   // the INTERNAL image (g_camera_fx .., g_image_width ..)
    struct Observation {          // one non-static entry of ins_seg_result with track_id <= g_max_movable_object_instance_id
        int track_id;
        bool label_known;         // label found in g_label_id_map_default
        int label_id;
        std::vector<Vec3> kpts_current, kpts_previous;
    };

    void configure(Mode mode, const Camera& cam, double map_half_size_scaled) { mode_ = mode; cam_ = cam; map_half_size_scaled_ = map_half_size_scaled; }
    void setMovementFilter(double distance_threshold, double probability_threshold, double increment, double decrement) {   // semantic_dsp_map.h:142-148
        dist_thr_ = distance_threshold; prob_thr_ = probability_threshold; inc_ = increment; dec_ = decrement;
    }
    void setSeed(uint32_t seed) { rng_.seed(seed); }
    void clear() { objects_.clear(); last_kpts_.clear(); last_kpts_time_.clear(); key_kpts_.clear(); key_kpts_time_.clear(); last_time_stamp_ = 0.0; }
    bool isTracked(int track_id) const { return objects_.count(track_id) > 0; }
    size_t size() const { return objects_.size(); }
    bool isMoving(int track_id) const { auto it = objects_.find(track_id); return it != objects_.end() && it->second.has_moved_flag && it->second.moved; }
    bool hasMovedFlag(int track_id) const { auto it = objects_.find(track_id); return it != objects_.end() && it->second.has_moved_flag; }
    // rigidbody_tmatrix_vec[0] as a row-major 4x4; false while the object has never been updated (the vector is empty)
    bool currentMatrix(int track_id, double out[16]) const {
        auto it = objects_.find(track_id);
        if (it == objects_.end() || !it->second.has_moved_flag) return false;
        const Rigid& T = it->second.T;
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out[4 * r + c] = T.R[3 * r + c]; out[4 * r + 3] = T.t[r]; }
        out[12] = out[13] = out[14] = 0.0; out[15] = 1.0;
        return true;
    }
    double movedProbability(int track_id) const { auto it = objects_.find(track_id); return it == objects_.end() ? -1.0 : it->second.moved_probability; }

    // objectLevelUpdate, semantic_dsp_map.h:306-566.  `step` = global_time_stamp of this frame.
    void observe(const std::vector<Observation>& seen, const double cam_pos[3], const double cam_quat_wxyz[4], double time_stamp, uint32_t step) {
        std::set<int> observed;
        const size_t min_kpts = (mode_ == kAlwaysMoving || mode_ == kMatchedKeypoints) ? 5 : 4;
        for (const Observation& o : seen) {
            const int id = o.track_id;
            observed.insert(id);
            if (!o.label_known) { std::printf("Warning: Object %d is not in the g_label_id_map_default. Ignore it.\n", id); continue; }
            bool success = false;
            auto it = objects_.find(id);
            if (it == objects_.end()) {                                                    // case 1, :346-381
                double closest = std::numeric_limits<double>::max();
                for (const Vec3& p : o.kpts_current) {
                    const double d = std::max(std::max(std::fabs(p[0] - cam_pos[0]), std::fabs(p[1] - cam_pos[1])), std::fabs(p[2] - cam_pos[2]));
                    if (d < closest) closest = d;
                }
                if (closest > map_half_size_scaled_) continue;
                Tracked t;
                t.label = o.label_id; t.observation_step = step; t.observation_count = 1;
                objects_[id] = t;
                success = true;
                if (mode_ == kBoxKeypoints) rememberKeypoints(id, o.kpts_current, time_stamp, true);
            } else if (o.kpts_current.size() >= min_kpts) {                                 // case 2, :383-513
                Rigid T;
                Vec3 reference_point{0, 0, 0};
                double confidence = 1.0, time_interval = 0.15;
                int moved_observation = -1;
                if (mode_ == kBoxKeypoints) {
                    const bool out_of_fov = outOfFov(cam_pos, cam_quat_wxyz, o.kpts_current.back(), 5);   // :419-422: the LAST key point decides
                    const double time_diff = time_stamp - last_kpts_time_[id];
                    auto last = last_kpts_.find(id);
                    if (out_of_fov) {
                        success = false;
                    } else if (last == last_kpts_.end() || last->second.size() < 4 || o.kpts_current.size() != 4) {
                        if (last == last_kpts_.end()) std::printf("Error: No 3D keypoints for object %d in last_kpts_3d_map_. Ignore updating it.\n", id);
                        rememberKeypoints(id, o.kpts_current, time_stamp, true);
                    } else {
                        const std::vector<Vec3> P(last->second.begin(), last->second.begin() + 4), &Q = o.kpts_current;
                        std::vector<int> inliers;
                        kabschRansac(P, Q, T, inliers, 2, 0.5, false, rng_);
                        double thr = dist_thr_;                                             // :451-457
                        const double width = norm(sub(Q[1], Q[0]));
                        if (thr < width) thr = width;
                        auto key = key_kpts_.find(id);
                        const double moved_dist = key == key_kpts_.end() || key->second.empty() ? 0.0 : norm(sub(Q[0], key->second[0]));
                        moved_observation = moved_dist > thr ? 1 : 0;
                        if (time_stamp - key_kpts_time_[id] > 2.0) { key_kpts_[id] = o.kpts_current; key_kpts_time_[id] = time_stamp; }
                        reference_point = P[0];
                        rememberKeypoints(id, o.kpts_current, time_stamp, false);
                        time_interval = time_diff;
                        success = true;
                    }
                } else {                                                                     // :386-410, 489-504
                    if (o.kpts_previous.size() == o.kpts_current.size()) {
                        std::vector<int> inliers;
                        const double mse = kabschRansac(o.kpts_previous, o.kpts_current, T, inliers, 100, 0.5, true, rng_);
                        if (mse > 0.2f || inliers.size() < 5 || inliers.size() / (double)o.kpts_current.size() < 0.5f) {
                            std::printf("Transformation is not valid.mse = %g, inlier_indices.size = %zu\n", mse, inliers.size());
                        } else {
                            success = true;
                            reference_point = inliers.empty() ? o.kpts_previous[0] : o.kpts_previous[inliers[0]];
                            if (!inliers.empty() && inliers.size() < 3) confidence = 0.0;
                        }
                    }
                }
                if (success) updateObject(it->second, o.label_id, T, reference_point, time_interval, moved_observation, step);
                (void)confidence;   // the reference passes it on but never reads it (object_layer.h:467)
            }
            if ((mode_ == kAlwaysMoving || mode_ == kMatchedKeypoints) && !success) {      // case 3, :519-540
                Tracked& t = objects_.at(id);
                if (t.has_moved_flag && t.moved) {
                    if (t.velocity_valid) predictAndSet(t, 0.2);
                    else { t.observation_step = step; t.match_with_previous = true; }       // setFlagsUpdateByMatching, object_layer.h:526-534
                }
            }
        }
        for (auto& kv : objects_) {                                                         // :545-563: unobserved moving objects coast
            if (observed.count(kv.first)) continue;
            Tracked& t = kv.second;
            if (!t.has_moved_flag || !t.moved) continue;
            double dt = time_stamp - last_time_stamp_;
            if (std::fabs(dt) > 1.0) dt = 1.0;
            predictAndSet(t, dt);
        }
        last_time_stamp_ = time_stamp;
    }

    struct Move { int track_id; float t[16]; };
    // head of subObjectLevelUpdate, semantic_dsp_map.h:588-704: moving objects either move by their current matrix or, when
    // unobserved for max_lost steps, are dropped together with their particles.
    void plan(uint32_t step, int max_lost, std::vector<Move>& moves, std::vector<int>& dropped) {
        for (auto& kv : objects_) {
            const Tracked& t = kv.second;
            if (!t.has_moved_flag || !t.moved) continue;
            if ((int64_t)step - (int64_t)t.observation_step >= (int64_t)max_lost) { dropped.push_back(kv.first); continue; }
            Move m;
            m.track_id = kv.first;
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m.t[4 * r + c] = (float)t.T.R[3 * r + c]; m.t[4 * r + 3] = (float)t.T.t[r]; }
            m.t[12] = m.t[13] = m.t[14] = 0.f; m.t[15] = 1.f;
            moves.push_back(m);
        }
        for (int id : dropped) { objects_.erase(id); last_kpts_.erase(id); last_kpts_time_.erase(id); }   // :694-704
    }
    // a "floating" track (:712-737) loses its stored key points as well
    void forget(int track_id) { last_kpts_.erase(track_id); last_kpts_time_.erase(track_id); key_kpts_.erase(track_id); key_kpts_time_.erase(track_id); }

private:
    struct Tracked {                 // ObjectSet::ObjectInTracking + MJObject + ObjectTransformations, object_layer.h:207-341,353-366
        int label = 0;
        uint32_t observation_step = 0;
        int observation_count = 0;
        bool has_moved_flag = false, moved = false, match_with_previous = false;
        double moved_probability = 0.5;
        Rigid T;                     // rigidbody_tmatrix_vec[0]
        // sliding window of the transformations seen while moving (max 5, none older than 10 steps)
        std::vector<Rigid> win_T;
        std::vector<double> win_dt;
        std::vector<Vec3> win_ref;
        std::vector<uint32_t> win_step;
        bool velocity_valid = false;
        Vec3 velocity{0, 0, 0};
    };

    void rememberKeypoints(int id, const std::vector<Vec3>& k, double ts, bool key_too) {
        last_kpts_[id] = k; last_kpts_time_[id] = ts;
        if (key_too) { key_kpts_[id] = k; key_kpts_time_[id] = ts; }
    }

    // isPointOutOfFOV, semantic_dsp_map.h:1421-1442
    bool outOfFov(const double pos[3], const double q[4], const Vec3& p, int margin) const {
        const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        const double w = q[0] / n2, x = -q[1] / n2, y = -q[2] / n2, z = -q[3] / n2;      // inverse quaternion
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
        const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
        const Vec3 d{p[0] - pos[0], p[1] - pos[1], p[2] - pos[2]};
        const double cx_ = (R[0] * d[0] + R[1] * d[1]) + R[2] * d[2], cy_ = (R[3] * d[0] + R[4] * d[1]) + R[5] * d[2], cz_ = (R[6] * d[0] + R[7] * d[1]) + R[8] * d[2];
        if (cz_ <= 0) return true;
        const double u = cam_.fx * (cx_ / cz_) + cam_.cx, v = cam_.fy * (cy_ / cz_) + cam_.cy;
        return u < margin || u >= cam_.width - margin || v < margin || v >= cam_.height - margin;
    }

    // ObjectSet::updateObject, object_layer.h:467-523
    void updateObject(Tracked& t, int label, const Rigid& T, const Vec3& reference_point, double time_interval, int moved_observation, uint32_t step) {
        bool moving = false;
        if (mode_ == kAlwaysMoving) {
            moving = true;
        } else if (mode_ != kStatic) {
            bool up;
            if (moved_observation == -1) up = norm(sub(T.apply(reference_point), reference_point)) > dist_thr_;
            else up = moved_observation == 1;
            if (up) t.moved_probability += inc_; else t.moved_probability -= dec_;
            moving = t.moved_probability > prob_thr_;
        }
        t.moved_probability = std::min(1.0, std::max(0.0, t.moved_probability));
        t.label = label; t.T = T; t.observation_step = step; ++t.observation_count;
        t.match_with_previous = false; t.has_moved_flag = true; t.moved = moving;
        if (moving) pushTransformation(t, T, time_interval, reference_point, step);
    }

    // ObjectTransformations::update + MotionEstimation::estimateByTransformations / estimate, object_layer.h:94-152, 206-258.
    // Only the translational velocity is kept: the prediction never applies the estimated rotation (object_layer.h:190).
    void pushTransformation(Tracked& t, const Rigid& T, double dt, const Vec3& ref, uint32_t step) {
        t.win_T.push_back(T); t.win_dt.push_back(dt); t.win_ref.push_back(ref); t.win_step.push_back(step);
        auto pop_front = [&] { t.win_T.erase(t.win_T.begin()); t.win_dt.erase(t.win_dt.begin()); t.win_ref.erase(t.win_ref.begin()); t.win_step.erase(t.win_step.begin()); };
        while (!t.win_ref.empty() && step - t.win_step[0] > 10) pop_front();
        if (t.win_T.size() > 5) pop_front();
        if (t.win_T.size() < 2) { t.velocity_valid = false; return; }
        Vec3 sum{0, 0, 0};
        for (size_t i = 0; i < t.win_T.size(); ++i) {
            const Vec3 p0 = t.win_ref[i], p1{p0[0] + 1, p0[1], p0[2]}, p2{p0[0], p0[1] + 1, p0[2]};
            const Vec3 q0 = t.win_T[i].apply(p0), q1 = t.win_T[i].apply(p1), q2 = t.win_T[i].apply(p2);
            for (int a = 0; a < 3; ++a) {
                const double prev = ((p0[a] + p1[a]) + p2[a]) / 3.0, curr = ((q0[a] + q1[a]) + q2[a]) / 3.0;
                sum[a] += (curr - prev) / t.win_dt[i];
            }
        }
        const double denom = (double)(t.win_T.size() - 1);
        for (int a = 0; a < 3; ++a) t.velocity[a] = sum[a] / denom;
        t.velocity_valid = true;
    }

    // ObjectSet::predictAndSetTransformation, object_layer.h:539-566: pure translation by velocity * dt when a velocity
    // exists, else the last matrix stays
    void predictAndSet(Tracked& t, double dt) {
        if (t.velocity_valid) {
            Rigid T;
            for (int a = 0; a < 3; ++a) T.t[a] = t.velocity[a] * dt;
            t.T = T;
        }
        t.match_with_previous = false;
    }

    Mode mode_ = kBoxKeypoints;
    Camera cam_{};
    double map_half_size_scaled_ = 0.0;
    double dist_thr_ = 0.1, prob_thr_ = 0.69, inc_ = 0.1, dec_ = 0.15;   // constructor defaults, semantic_dsp_map.h:36-39
    std::map<int, Tracked> objects_;
    std::map<int, std::vector<Vec3>> last_kpts_, key_kpts_;
    std::map<int, double> last_kpts_time_, key_kpts_time_;
    double last_time_stamp_ = 0.0;
    std::mt19937 rng_{20240911u};
};

}  // namespace dsp_adapter
