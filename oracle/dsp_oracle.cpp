// dsp_oracle.cpp — CPU oracle for the semantic_dsp_map per-frame hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
// and only as the checker or as the timed CPU baseline.
//
// PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path, so
// the oracle is pinned against the reference's OWN CODE instead: oracle/build_ref.sh compiles the reference's
// headers, unmodified and where they lie under /root/reference/include, against stand-in Eigen / OpenCV / PCL /
// Boost headers (oracle/ref_stubs; those libraries are absent from this image) into oracle/_ref/, and
// tests/test_reference_pin_cpu.py demands that this file, run in the reference's iteration order, reproduces
// that build BIT FOR BIT on seeded scenarios: ring state, labelled points, per-pixel particle lists in BFS
// discovery order, every field of every particle record after every frame, the occupied / free voxel lists and
// positions and the RGB bytes of every emitted point (SETTING 3 as shipped and SETTING 2).  Fixtures generated
// from that build are committed under tests/golden/ref_*.npz for hosts without /root/reference.
// This file restates the reference algorithm function by function from its sources; every function cites the
// reference file:line it follows (paths are relative to the reference root).  Third-party arithmetic on the
// path is Eigen's fixed-size matrix code (un-vendored, version unpinned: find_package(Eigen3),
// CMakeLists.txt:20); it is restated here — and identically in the stand-in Eigen the reference build uses, so
// the pin does not cover it — as closed-form cofactor inverses and
// left-to-right dot products WITHOUT fused multiply-add (build with -ffp-contract=off).
//
// Where the reference reads uninitialised memory the oracle DEFINES the value
// (documented at each site, all listed in DESIGN.md §"Defined UB").
//
// Two iteration orders are offered:
//   order 0 "reference": per-pixel particle lists in BFS discovery order, object particle
//                        sets kept in std::unordered_set exactly like the reference host code;
//   order 1 "canonical": per-pixel lists in ascending particle index, voxel sweep = every voxel
//                        touching an in-frustum lattice vertex, object sets = {p : track(p)=T}
//                        in ascending particle index.  The CUDA engine matches this order
//                        bit-exactly; tests also report how many discrete outputs differ
//                        between the two orders.
#include "../include/dspmap_c.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <queue>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---- types -------------------------------------------------------------------------

// include/mc_ring/buffer.h:43-50
enum : uint8_t { ST_INVALID = 0, ST_UPDATED = 1, ST_REGULAR_BORN = 2, ST_GUESSED_BORN = 3, ST_COPIED = 4, ST_TIMEPTC = 5 };

// include/mc_ring/buffer.h:57-79 (24 bytes)
struct Particle {
    float x, y, z, weight;
    uint16_t time_stamp, track_id, label_id;
    uint8_t status, forget_count;
};
static_assert(sizeof(Particle) == 24, "reference AoS record is 24 bytes");

// include/utils/data_base.h:78-92 (20 bytes)
struct LabeledPoint {
    float x, y, z, sigma;
    uint16_t track_id;
    uint8_t label_id;
    uint8_t is_valid;
};
static_assert(sizeof(LabeledPoint) == 20, "reference LabeledPoint is 20 bytes");

constexpr uint32_t INVALID_INDEX = 0xffffffffu;          // mc_ring/operations.h:35
constexpr float OCC_INIT_WEIGHT = 0.05f;                 // settings.h:147
constexpr float MIN_RIGHTLY_UPDATED_PDF = 0.1f;          // settings.h:149
constexpr float DEPTH_ERR_STDDEV_1M = 0.1f;              // settings.h:150
constexpr int PDF_NUM = 20000;                           // basic_algorithms.h:378

// ---- small fixed-size linear algebra with a pinned evaluation order ------------------

// Eigen::Quaternion<T>::toRotationMatrix (closed form), row-major 3x3
template <typename T>
void quat_to_rot(const T q[4] /*w,x,y,z*/, T R[9]) {
    const T w = q[0], x = q[1], y = q[2], z = q[3];
    const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x;
    const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = T(1) - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
    R[3] = txy + twz;          R[4] = T(1) - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = T(1) - (txx + tyy);
}

// 3x3 inverse by cofactors / determinant (the closed form Eigen uses for fixed 3x3)
void inverse3d(const double m[9], double o[9]) {
    auto M = [&](int r, int c) { return m[r * 3 + c]; };
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = (c00 * M(0, 0) + c10 * M(1, 0)) + c20 * M(2, 0);
    const double invdet = 1.0 / det;
    // result(r,c) = cofactor(c,r) * invdet
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o[r * 3 + c] = cof(c, r) * invdet;
}

// 4x4 float inverse by 3x3 cofactor expansion (Eigen generic fixed-4x4 path:
// cofactor_4x4 = sum of three general_det3_helper terms; result /= determinant)
void inverse4f(const float m[16], float o[16]) {
    auto M = [&](int r, int c) { return m[r * 4 + c]; };
    auto det3h = [&](int i1, int i2, int i3, int j1, int j2, int j3) {
        return M(i1, j1) * (M(i2, j2) * M(i3, j3) - M(i2, j3) * M(i3, j2));
    };
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4;
        const int j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
        return (det3h(i1, i2, i3, j1, j2, j3) + det3h(i2, i3, i1, j1, j2, j3)) + det3h(i3, i1, i2, j1, j2, j3);
    };
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const float c = cof(i, j);
            r[j * 4 + i] = ((i + j) & 1) ? -c : c;  // adjugate = transposed signed cofactors
        }
    // determinant = column 0 of m dotted with row 0 of the adjugate
    const float det = ((M(0, 0) * r[0] + M(1, 0) * r[1]) + M(2, 0) * r[2]) + M(3, 0) * r[3];
    for (int k = 0; k < 16; ++k) o[k] = r[k] / det;
}

// row * (x,y,z,1), left to right, no FMA
inline void mat4_apply(const float m[16], float x, float y, float z, float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = ((m[r * 4 + 0] * x + m[r * 4 + 1] * y) + m[r * 4 + 2] * z) + m[r * 4 + 3] * 1.0f;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}


// ---- colouring of the occupied list: semantic_dsp_map.h:1274-1351 -------------------------------------------
// cv::cvtColor(COLOR_RGB2HSV / COLOR_HSV2RGB) on one CV_8UC3 pixel, restated from OpenCV's published 8-bit algorithm
// and checked bit for bit against cv2 4.13 on all 2^24 inputs (tests/test_colour_cpu.py): H in [0,180), S and H through
// 12-bit fixed-point reciprocal tables; back through float sectors with the two "1 - s*h" terms fused and the final
// scaling truncated.
inline void rgb_to_hsv_u8(const uint8_t in[3], uint8_t out[3]) {
    static int sdiv[256], hdiv[256];
    static bool init = false;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; ++i) {
            sdiv[i] = (int)lrint((255 << 12) / (1. * i));
            hdiv[i] = (int)lrint((180 << 12) / (6. * i));
        }
        init = true;
    }
    const int r = in[0], g = in[1], b = in[2];
    const int v = std::max(r, std::max(g, b)), vmin = std::min(r, std::min(g, b));
    const int diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
    int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * hdiv[diff] + (1 << 11)) >> 12;
    h += h < 0 ? 180 : 0;
    out[0] = (uint8_t)h; out[1] = (uint8_t)s; out[2] = (uint8_t)v;
}
inline void hsv_to_rgb_u8(const uint8_t in[3], uint8_t out[3]) {
    float h = (float)in[0];
    const float s = in[1] * (1.f / 255.f), v = in[2] * (1.f / 255.f);
    float r, g, b;
    if (in[1] == 0) {
        r = g = b = v;
    } else {
        static const int sector_data[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
        h *= 6.f / 180.f;
        while (h >= 6.f) h -= 6.f;
        int sector = (int)floorf(h);
        h -= (float)sector;
        if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
        float tab[4];
        tab[0] = v;
        tab[1] = v * (1.f - s);
        tab[2] = v * fmaf(-s, h, 1.f);
        tab[3] = v * fmaf(-s, 1.f - h, 1.f);
        b = tab[sector_data[sector][0]]; g = tab[sector_data[sector][1]]; r = tab[sector_data[sector][2]];
    }
    auto sat = [](float x) { const float y = floorf(x * 255.f); return (uint8_t)(y < 0.f ? 0.f : (y > 255.f ? 255.f : y)); };
    out[0] = sat(r); out[1] = sat(g); out[2] = sat(b);
}
// color_map_jet_256_, semantic_dsp_map.h:50-63
inline void jet_256(int i, uint8_t out[3]) {
    if (i < 64) { out[0] = 0; out[1] = 0; out[2] = (uint8_t)(i * 4); }
    else if (i < 128) { out[0] = 0; out[1] = (uint8_t)((i - 64) * 4); out[2] = 255; }
    else if (i < 192) { out[0] = (uint8_t)((i - 128) * 4); out[1] = 255; out[2] = (uint8_t)(255 - (i - 128) * 4); }
    else { out[0] = 255; out[1] = (uint8_t)(255 - (i - 192) * 4); out[2] = 0; }
}

// ---- the oracle map ---------------------------------------------------------------------

struct Oracle {
    dspmap_config cfg{};
    dspmap_params prm{};
    int order = 1;

    // derived constants (settings.h / buffer.h / operations.h:726-767)
    int Xb = 0, Yb = 0, Zb = 0, Pb = 0;
    uint32_t NX = 0, NY = 0, NZ = 0, P = 0, V = 0;
    size_t N = 0;
    float voxel = 0, voxel_recip = 0;
    float pmin[3]{}, pmax[3]{};
    float fx = 0, fy = 0, cx = 0, cy = 0;
    int W = 0, H = 0;  // internal image size
    int win_r = 0;
    float tan_half_fovx = 0, tan_half_fovy = 0;
    bool fov_init = false;

    // state (buffer.h:83-120, data_base.h:22)
    std::vector<Particle> store;
    std::vector<uint32_t> stamps[3];
    int moved[3]{}, equiv[3]{};
    float map_center[3]{}, ego_center[3]{}, last_pos[3]{};
    uint32_t T = 0;

    // label tables
    uint16_t static_value_to_track[256]{};
    std::map<uint16_t, uint8_t> static_label;
    uint16_t max_movable = 65523, sky_track = 65533;

    // Gaussian tables (basic_algorithms.h:377-461)
    std::vector<float> pdf_lut;
    std::vector<float> randoms;
    int cursor_map = 0;   // SemanticDSPMap::gaussian_random_ (semantic_dsp_map.h:265)
    int cursor_ring = 0;  // RingBufferOperations::gaussian_random_calculator_ (operations.h:1460)
    float forgetting_table[5]{};
    bool forgetting_init = false;

    // per-frame scratch
    std::vector<float> depth_int;
    std::vector<LabeledPoint> cloud;
    std::vector<std::vector<uint32_t>> pix_lists;
    std::vector<float> ck;
    float extrinsic[16]{};
    std::vector<int8_t> voxel_state;

    // reference-order object particle sets (object_layer.h:20-52)
    std::unordered_map<int, std::unordered_set<uint32_t>> obj_sets;

    // outputs
    std::vector<uint32_t> out_idx, free_idx;
    std::vector<float> out_xyz, out_w, free_xyz;
    std::vector<uint16_t> out_track;
    std::vector<uint8_t> out_label;
    std::vector<int8_t> out_state;
    dspmap_frame_stats stats{};
    uint64_t sweep_set_mismatch = 0;  // |BFS voxel set  xor  in-frustum-vertex voxel set|, reference order only
    uint64_t start_vertex_outside = 0;
    dspmap_colors colors{};          // orc_set_colors: colouring of the occupied list (semantic_dsp_map.h:1274-1351)
    std::vector<uint8_t> out_rgb;
    bool carry_undefined_label = true;   // orc_set_undefined_label_carry: see occupancy()
    bool invalid_sigma_set = false;  // orc_set_invalid_sigma: see project()
    float invalid_sigma_value = 0.f;
    bool diagnostics = false;  // the sweep-set check costs a full pass over the map: off when the oracle is being timed
    double stage_s[8]{};
    std::vector<uint32_t> csr_off, csr_idx;

    // ------------------------------------------------------------------------------
    int init(const dspmap_config& c, int order_mode) {
        cfg = c;
        order = order_mode;
        Xb = c.voxel_bits_x; Yb = c.voxel_bits_y; Zb = c.voxel_bits_z; Pb = c.slot_bits;
        // operations.h:54-58 runSystemChecking
        if (Xb < 2 || Yb < 2 || Zb < 2 || Pb < 1 || Xb + Yb + Zb + Pb > 31) return DSPMAP_ERR_CONFIG;
        NX = 1u << Xb; NY = 1u << Yb; NZ = 1u << Zb; P = 1u << Pb;
        V = NX * NY * NZ;
        N = (size_t)V * P;
        voxel = c.voxel_size;
        if (c.boost_mode) {  // settings.h:136-143
            const float s = 0.5f;
            fx = s * c.fx; fy = s * c.fy; cx = s * c.cx; cy = s * c.cy;
            W = (int)(s * c.image_width); H = (int)(s * c.image_height);
        } else {             // settings.h:128-134
            fx = c.fx; fy = c.fy; cx = c.cx; cy = c.cy; W = c.image_width; H = c.image_height;
        }
        win_r = c.window_half > 0 ? c.window_half : (c.boost_mode ? 3 : 5);  // semantic_dsp_map.h:964-970
        store.assign(N, Particle{});
        stamps[0].assign(NX, 0); stamps[1].assign(NY, 0); stamps[2].assign(NZ, 0);
        initialize();
        // defaults of the SemanticDSPMap ctor (semantic_dsp_map.h:25-43)
        prm.detection_probability = 0.95f; prm.noise_number = 0.1f; prm.nb_ptc_num_per_point = 3;
        prm.occupancy_threshold = 0.2f; prm.forgetting_rate = 1.0f; prm.max_forget_count = 5;
        prm.id_transition_probability = 0.1f; prm.depth_noise_first_order = 0.0f;
        prm.depth_noise_zero_order = 0.1f; prm.consider_depth_noise = 0; prm.use_independent_filter = 0;
        prm.zero_center = 0;
        for (int i = 0; i < 256; ++i) static_value_to_track[i] = 65535;
        build_pdf_lut();
        randoms.assign(1, 0.0f);
        pix_lists.assign((size_t)W * H, {});
        ck.assign((size_t)W * H, 0.f);
        cloud.assign((size_t)W * H, LabeledPoint{});
        depth_int.assign((size_t)W * H, 0.f);
        voxel_state.assign(V, -1);
        return DSPMAP_OK;
    }

    // basic_algorithms.h:403-407, 455-459: LUT of (1/sqrt(2*pi/2)) * exp(-v^2/2), v = (i-10000)*0.001
    void build_pdf_lut() {
        pdf_lut.resize(PDF_NUM);
        const float half_pi_f32 = 1.57079632679489661923f;  // M_PI_2f32 (glibc: a _Float32 constant)
        for (int i = 0; i < PDF_NUM; ++i) {
            const float value = (float)(i - PDF_NUM / 2) * 0.001f;
            pdf_lut[i] = (1.f / (sqrtf(2.f * half_pi_f32))) * expf(-(value * value) / (2));
        }
    }

    // basic_algorithms.h:417-422
    inline float query_pdf(float x, float mu, float sigma) const {
        const float corrected_x = (x - mu) / sigma;
        // DEFINED UB: a NaN argument (0/0 when sigma == 0 and x == mu) indexes the table out of range in the
        // reference; it is defined here as the tail value.
        if (corrected_x > 9.9f || corrected_x < -9.9f || corrected_x != corrected_x) return 1e-9f;
        return pdf_lut[static_cast<int>(corrected_x * 1000 + 10000)];
    }

    // basic_algorithms.h:426-433 (pre-increment, wrap)
    inline float next_random(int& cursor) const {
        cursor += 1;
        if (cursor >= (int)randoms.size()) cursor = 0;
        return randoms[cursor];
    }

    // basic_algorithms.h:32-48; the table is built once, from the first call's rate
    inline float forgetting_factor(int forget_count) {
        if (!forgetting_init) {
            for (int i = 0; i < 5; ++i) forgetting_table[i] = (float)pow(2.5, -i / prm.forgetting_rate);
            forgetting_init = true;
        }
        if (forget_count < prm.max_forget_count) return forgetting_table[forget_count];
        return 0.f;
    }

    // operations.h:684-723
    void clear() {
        T = 0;
        for (int a = 0; a < 3; ++a) std::fill(stamps[a].begin(), stamps[a].end(), 0u);
        for (uint32_t i = 0; i < V; ++i) {
            const size_t s = (size_t)i << Pb;
            store[s] = Particle{};
            store[s].status = ST_TIMEPTC;
            for (uint32_t j = 1; j < P; ++j) { store[s + j] = Particle{}; store[s + j].status = ST_INVALID; }
        }
        obj_sets.clear();  // semantic_dsp_map.h:74-81 -> ObjectSet::clear
    }

    // operations.h:726-767
    void initialize() {
        clear();
        for (int a = 0; a < 3; ++a) { moved[a] = 0; equiv[a] = 0; map_center[a] = 0.f; ego_center[a] = 0.f; last_pos[a] = 0.f; }
        pmax[0] = (NX >> 1) * voxel; pmax[1] = (NY >> 1) * voxel; pmax[2] = (NZ >> 1) * voxel;
        for (int a = 0; a < 3; ++a) pmin[a] = -pmax[a];
        voxel_recip = 1.f / voxel;
    }

    inline uint32_t axis_n(int a) const { return a == 0 ? NX : (a == 1 ? NY : NZ); }

    // operations.h:1037-1107
    inline uint32_t axis_correct(int idx, int a) const {
        const int n = (int)axis_n(a);
        if (idx < 0) return (uint32_t)(idx + n);
        if (idx >= n) return (uint32_t)(idx - n);
        return (uint32_t)idx;
    }

    // operations.h:1196-1230
    inline int equivalent_steps(int steps, int a) const {
        const int n = (int)axis_n(a);
        if (steps > 0) return steps % n;
        if (steps < 0) return -(-steps % n);
        return steps;
    }

    // operations.h:1111-1191
    void update_ring_params() {
        int center[3];
        for (int a = 0; a < 3; ++a) {
            center[a] = static_cast<int>(ego_center[a] * voxel_recip);
            map_center[a] = static_cast<float>(center[a]) * voxel;
        }
        for (int a = 0; a < 3; ++a) {
            const int n = (int)axis_n(a);
            const int new_moved = center[a] - moved[a];
            if (new_moved > 0) {
                for (int i = 0; i < new_moved; ++i) stamps[a][axis_correct(i + equiv[a], a)] = T;
            } else if (new_moved < 0) {
                for (int i = 0; i < -new_moved; ++i) stamps[a][axis_correct(n - 1 - i + equiv[a], a)] = T;
            }
        }
        for (int a = 0; a < 3; ++a) { moved[a] = center[a]; equiv[a] = equivalent_steps(moved[a], a); }
    }

    // operations.h:68-96
    void update_ego_center(const float pos[3]) {
        const float mx = (1 << (Xb - 2)) * voxel, my = (1 << (Yb - 2)) * voxel, mz = (1 << (Zb - 2)) * voxel;
        const float max_once = std::min(std::min(mx, my), mz);
        float mv[3] = {pos[0] - last_pos[0], pos[1] - last_pos[1], pos[2] - last_pos[2]};
        const float sq = (mv[0] * mv[0] + mv[1] * mv[1]) + mv[2] * mv[2];
        float unit[3] = {mv[0], mv[1], mv[2]};
        if (sq > 0.f) { const float nn = sqrtf(sq); unit[0] = mv[0] / nn; unit[1] = mv[1] / nn; unit[2] = mv[2] / nn; }
        float dist = sqrtf(sq);
        float np[3] = {last_pos[0], last_pos[1], last_pos[2]};
        while (dist > max_once) {
            for (int a = 0; a < 3; ++a) np[a] = np[a] + unit[a] * max_once;
            for (int a = 0; a < 3; ++a) ego_center[a] = np[a];
            update_ring_params();
            for (int a = 0; a < 3; ++a) mv[a] = pos[a] - np[a];
            dist = sqrtf((mv[0] * mv[0] + mv[1] * mv[1]) + mv[2] * mv[2]);
        }
        for (int a = 0; a < 3; ++a) ego_center[a] = pos[a];
        update_ring_params();
        for (int a = 0; a < 3; ++a) last_pos[a] = pos[a];
    }

    // operations.h:994-1019
    inline void map_to_buffer(int mx, int my, int mz, uint32_t& bx, uint32_t& by, uint32_t& bz) const {
        bx = axis_correct(mx + equiv[0], 0); by = axis_correct(my + equiv[1], 1); bz = axis_correct(mz + equiv[2], 2);
    }
    // operations.h:1022-1033
    inline void buffer_to_map(uint32_t bx, uint32_t by, uint32_t bz, uint32_t& mx, uint32_t& my, uint32_t& mz) const {
        mx = axis_correct((int)bx - equiv[0], 0); my = axis_correct((int)by - equiv[1], 1); mz = axis_correct((int)bz - equiv[2], 2);
    }
    // operations.h:890-903 (row major, STORAGE_TYPE 0)
    inline uint32_t buffer_to_voxel(uint32_t bx, uint32_t by, uint32_t bz) const { return (((bz << Yb) | by) << Xb) | bx; }
    // operations.h:947-967
    inline void voxel_to_buffer(uint32_t v, uint32_t& bx, uint32_t& by, uint32_t& bz) const {
        bx = v & (NX - 1); by = (v >> Xb) & (NY - 1); bz = (v >> (Xb + Yb)) & (NZ - 1);
    }

    // operations.h:849-883 (globalFramePostoVoxelIdx -> mapFramePostoVoxelIdx)
    inline uint32_t global_pos_to_voxel(float x, float y, float z, uint32_t& bx, uint32_t& by, uint32_t& bz) const {
        const float mp[3] = {x - map_center[0], y - map_center[1], z - map_center[2]};
        // static_cast<uint32_t>(float): defined here as truncation of the value when it is in
        // [0, 2^32); negative / NaN / huge inputs (UB in C++) are treated as "outside the map",
        // which is what the x86-64 conversion produces in practice (cvttss2si -> huge value).
        uint32_t mi[3];
        for (int a = 0; a < 3; ++a) {
            const float f = (mp[a] - pmin[a]) * voxel_recip;
            if (!(f > -1.0f) || !(f < 4294967296.0f)) return INVALID_INDEX;
            mi[a] = (f < 0.f) ? 0u : static_cast<uint32_t>(f);
        }
        if (mi[0] < NX && mi[1] < NY && mi[2] < NZ) {
            map_to_buffer((int)mi[0], (int)mi[1], (int)mi[2], bx, by, bz);
            return buffer_to_voxel(bx, by, bz);
        }
        return INVALID_INDEX;
    }

    // operations.h:810-816
    inline bool is_vacant(const Particle& p, uint32_t bx, uint32_t by, uint32_t bz) const {
        return p.status == ST_INVALID || p.time_stamp < stamps[0][bx] || p.time_stamp < stamps[1][by] || p.time_stamp < stamps[2][bz];
    }
    // operations.h:824-837
    inline bool is_voxel_valid(uint32_t v, uint32_t bx, uint32_t by, uint32_t bz) const {
        const Particle& tp = store[(size_t)v << Pb];
        if (tp.time_stamp == 0) return false;
        return !(tp.time_stamp < stamps[0][bx] || tp.time_stamp < stamps[1][by] || tp.time_stamp < stamps[2][bz]);
    }

    // operations.h:782-803
    uint32_t add_particle(const Particle& p, uint32_t& voxel_index) {
        uint32_t bx = 0, by = 0, bz = 0;
        voxel_index = global_pos_to_voxel(p.x, p.y, p.z, bx, by, bz);
        if (voxel_index == INVALID_INDEX) return INVALID_INDEX;
        const size_t s = (size_t)voxel_index << Pb;
        for (uint32_t i = 1; i < P; ++i) {
            if (is_vacant(store[s + i], bx, by, bz)) { store[s + i] = p; return (uint32_t)(s + i); }
        }
        return INVALID_INDEX;
    }

    // operations.h:171-184
    uint32_t add_new_particle(float x, float y, float z, uint8_t label, uint16_t track, uint32_t& voxel_index) {
        Particle p{};
        p.status = ST_REGULAR_BORN; p.time_stamp = (uint16_t)T; p.forget_count = 0;
        p.x = x; p.y = y; p.z = z; p.weight = OCC_INIT_WEIGHT; p.label_id = label; p.track_id = track;
        return add_particle(p, voxel_index);
    }

    // ---- K1: pointcloud_tools.h:1104-1133 + :88-310 --------------------------------------
    template <typename Tp>
    static void half_resize(const Tp* src, int rows, int cols, int stride, Tp* dst, int new_rows, int new_cols) {
        const float scale_inv = 1.f / 0.5f;
        for (int i = 0; i < new_rows; ++i)
            for (int j = 0; j < new_cols; ++j) {
                int si = static_cast<int>(i * scale_inv), sj = static_cast<int>(j * scale_inv);
                si = std::min(si, rows - 1); sj = std::min(sj, cols - 1);
                dst[(size_t)i * new_cols + j] = src[(size_t)si * stride + sj];
            }
    }

    void project(const float* depth, int rows, int cols, int stride, const dspmap_mask* masks, int n_masks,
                 const double pos[3], const double quat[4]) {
        // depth -> internal resolution (in-place resize of the caller's Mat in the reference)
        if (cfg.boost_mode) half_resize<float>(depth, rows, cols, stride, depth_int.data(), H, W);
        else for (int i = 0; i < H; ++i) memcpy(&depth_int[(size_t)i * W], depth + (size_t)i * stride, sizeof(float) * W);

        // camera_to_global (double), pointcloud_tools.h:105-110
        double R[9];
        quat_to_rot<double>(quat, R);
        // inverse intrinsics (double from float entries), :113
        const double K[9] = {(double)fx, 0, (double)cx, 0, (double)fy, (double)cy, 0, 0, 1};
        double Kinv[9];
        inverse3d(K, Kinv);

        // track id image, :116-155
        std::vector<uint16_t> track_mask((size_t)W * H, 0);
        std::vector<uint8_t> mbuf((size_t)W * H);
        bool has_static = false;
        for (int m = 0; m < n_masks; ++m) {
            if (!masks[m].is_static) continue;
            const uint8_t* md; int ms;
            if (cfg.boost_mode) { half_resize<uint8_t>(masks[m].data, rows, cols, masks[m].row_stride, mbuf.data(), H, W); md = mbuf.data(); ms = W; }
            else { md = masks[m].data; ms = masks[m].row_stride; }
            for (int j = 0; j < H; ++j)
                for (int k = 0; k < W; ++k) track_mask[(size_t)j * W + k] = static_value_to_track[md[(size_t)j * ms + k]];
            has_static = true;
            break;  // :142 only the first static mask
        }
        if (!has_static) std::fill(track_mask.begin(), track_mask.end(), (uint16_t)65535);

        // movable objects, :162-215
        std::unordered_map<int, int> track_to_label;
        std::unordered_map<uint16_t, double> bmin[3], bmax[3];
        if (cfg.consider_instance) {
            for (int m = 0; m < n_masks; ++m) {
                if (masks[m].is_static) continue;
                track_to_label[masks[m].track_id] = masks[m].label_id;
                const uint8_t* md; int ms;
                if (cfg.boost_mode) { half_resize<uint8_t>(masks[m].data, rows, cols, masks[m].row_stride, mbuf.data(), H, W); md = mbuf.data(); ms = W; }
                else { md = masks[m].data; ms = masks[m].row_stride; }
                if (cfg.zed2_filters)
                    for (int a = 0; a < 3; ++a) { bmin[a][(uint16_t)masks[m].track_id] = masks[m].bbox_min[a]; bmax[a][(uint16_t)masks[m].track_id] = masks[m].bbox_max[a]; }
                for (int j = 0; j < H; ++j)
                    for (int k = 0; k < W; ++k)
                        if (md[(size_t)j * ms + k] > 0) track_mask[(size_t)j * W + k] = (uint16_t)masks[m].track_id;
            }
        }

        const double dmin = (double)cfg.depth_min, dmax = (double)cfg.depth_max;  // :21-22 (double members)
        uint32_t n_valid = 0;
        for (int i = 0; i < H; ++i) {
            for (int j = 0; j < W; ++j) {
                LabeledPoint pt{};
                const float d = depth_int[(size_t)i * W + j];
                // DEFINED UB: the reference copies an uninitialised LabeledPoint for invalid pixels
                // (:224-232); sigma of such a pixel is read later by updateParticles pass 2
                // (semantic_dsp_map.h:1047).  The oracle defines it as the zero-depth noise term.  What an actual
                // build of the reference reads there is whatever its stack slot holds (0 in oracle/_ref's build);
                // orc_set_invalid_sigma lets the pin test feed that value in so everything else can be compared.
                const bool invalid_sigma_set = this->invalid_sigma_set || prm.uninit_sigma_mode != 0;      // test knob, or the C-ABI flag (dspmap_params)
                const float invalid_sigma_value = this->invalid_sigma_set ? this->invalid_sigma_value : prm.uninit_sigma_value;
                const float sigma_invalid = invalid_sigma_set ? invalid_sigma_value : (prm.consider_depth_noise ? prm.depth_noise_zero_order : 0.1f);
                if (std::isnan(d) || (double)d < dmin || (double)d > dmax) {  // :228
                    pt.sigma = sigma_invalid; pt.is_valid = 0; cloud[(size_t)i * W + j] = pt; continue;
                }
                const uint16_t inst = track_mask[(size_t)i * W + j];
                if (cfg.zed2_filters && inst == sky_track) {  // :234-241
                    pt.sigma = sigma_invalid; pt.is_valid = 0; cloud[(size_t)i * W + j] = pt; continue;
                }
                // :243-249 (double): (Kinv * (j,i,1)) * depth ; R * p + t
                const double vj = (double)j, vi = (double)i;
                double pc[3];
                for (int r = 0; r < 3; ++r) pc[r] = ((Kinv[r * 3 + 0] * vj + Kinv[r * 3 + 1] * vi) + Kinv[r * 3 + 2] * 1.0) * (double)d;
                double pg[3];
                for (int r = 0; r < 3; ++r) pg[r] = ((R[r * 3 + 0] * pc[0] + R[r * 3 + 1] * pc[1]) + R[r * 3 + 2] * pc[2]) + pos[r];

                // sigma, :285-289.  DEFINED UB: the bbox-gated background branch (:256-272) leaves sigma
                // uninitialised in the reference; the oracle assigns the noise model there too (a compiled reference
                // reads the same stack slot as for invalid pixels: orc_set_invalid_sigma covers both sites).
                const float sigma = prm.consider_depth_noise ? (prm.depth_noise_zero_order + prm.depth_noise_first_order * d) : 0.1f;

                if (cfg.zed2_filters && cfg.consider_instance && inst < max_movable) {  // :254-273
                    // std::unordered_map::operator[] default-inserts 0.0 for an unknown id
                    const double x0 = bmin[0][inst], x1 = bmax[0][inst], y0 = bmin[1][inst], y1 = bmax[1][inst], z0 = bmin[2][inst], z1 = bmax[2][inst];
                    if (pg[0] < x0 || pg[0] > x1 || pg[1] < y0 || pg[1] > y1 || pg[2] < z0 || pg[2] > z1) {
                        pt.x = (float)pg[0]; pt.y = (float)pg[1]; pt.z = (float)pg[2];
                        pt.sigma = invalid_sigma_set ? invalid_sigma_value : sigma; pt.label_id = 0; pt.track_id = 65535; pt.is_valid = 1;
                        cloud[(size_t)i * W + j] = pt; ++n_valid; continue;
                    }
                }
                uint8_t label;
                if (inst > max_movable) {  // :277-279, std::map/unordered_map operator[] -> 0 when absent
                    auto it = static_label.find(inst);
                    label = it == static_label.end() ? 0 : it->second;
                } else {
                    label = (uint8_t)track_to_label[inst];  // :281
                }
                pt.x = (float)pg[0]; pt.y = (float)pg[1]; pt.z = (float)pg[2];
                pt.sigma = sigma; pt.label_id = label; pt.track_id = inst; pt.is_valid = 1;
                cloud[(size_t)i * W + j] = pt; ++n_valid;
            }
        }
        stats.n_valid_points = n_valid;
    }

    // ---- P2/P3: operations.h:321-362 + object_layer.h:414-425 -------------------------------
    void collect_track_set(int track, std::vector<uint32_t>& out) const {
        out.clear();
        for (size_t i = 0; i < N; ++i) {
            if ((i & (P - 1)) == 0) continue;
            const Particle& p = store[i];
            if (p.status != ST_INVALID && p.track_id == (uint16_t)track) out.push_back((uint32_t)i);
        }
    }

    void move_objects(const dspmap_object_move* moves, int n_moves, const int32_t* remove_tracks, int n_remove) {
        uint32_t n_moved = 0, n_removed = 0;
        if (n_moves > 0) {
            std::vector<std::vector<uint32_t>> sets(n_moves);
            std::vector<bool> present(n_moves, true);
            for (int i = 0; i < n_moves; ++i) {
                if (order == 0) {  // reference: obj_ptc_hash_map.indices_map.at(track) in unordered_set iteration order
                    auto it = obj_sets.find(moves[i].track_id);  // semantic_dsp_map.h:611 checkIfObjectExists
                    if (it == obj_sets.end()) { present[i] = false; continue; }
                    sets[i].assign(it->second.begin(), it->second.end());
                } else {
                    collect_track_set(moves[i].track_id, sets[i]);
                }
            }
            std::vector<std::vector<Particle>> np(n_moves);
            for (int i = 0; i < n_moves; ++i) {
                if (!present[i]) continue;
                const float* t = moves[i].t;
                np[i].resize(sets[i].size());
                for (size_t j = 0; j < sets[i].size(); ++j) {
                    const uint32_t idx = sets[i][j];
                    Particle q = store[idx];
                    // :340 R*p + t (float, rows left to right)
                    const float ox = q.x, oy = q.y, oz = q.z;
                    const float nx = ((t[0] * ox + t[1] * oy) + t[2] * oz) + t[3];
                    const float ny = ((t[4] * ox + t[5] * oy) + t[6] * oz) + t[7];
                    const float nz = ((t[8] * ox + t[9] * oy) + t[10] * oz) + t[11];
                    q.x = nx + next_random(cursor_ring);  // :342-344 queryNormalRandomPresetSD
                    q.y = ny + next_random(cursor_ring);
                    q.z = nz + next_random(cursor_ring);
                    np[i][j] = q;
                    store[idx].status = ST_INVALID;  // :346
                }
            }
            for (int i = 0; i < n_moves; ++i) {
                if (!present[i]) continue;
                std::unordered_set<uint32_t> fresh;
                for (size_t j = 0; j < np[i].size(); ++j) {
                    uint32_t v;
                    const uint32_t idx = add_particle(np[i][j], v);  // :356
                    if (idx != INVALID_INDEX) { fresh.insert(idx); ++n_moved; }
                }
                if (order == 0) obj_sets[moves[i].track_id] = fresh;  // semantic_dsp_map.h:697-699
            }
        }
        for (int r = 0; r < n_remove; ++r) {  // object_layer.h:414-425
            if (order == 0) {
                for (uint32_t idx : obj_sets[remove_tracks[r]]) { if (store[idx].status != ST_INVALID) ++n_removed; store[idx].status = ST_INVALID; }
                obj_sets.erase(remove_tracks[r]);
            } else {
                std::vector<uint32_t> s;
                collect_track_set(remove_tracks[r], s);
                for (uint32_t idx : s) { store[idx].status = ST_INVALID; ++n_removed; }
            }
        }
        stats.n_moved = n_moved; stats.n_removed = n_removed;
    }

    // ---- U2: operations.h:1240-1290, 1297-1457 ---------------------------------------------
    void init_fov() {
        if (fov_init) return;  // static const in the reference (:1249-1250)
        tan_half_fovx = (float)tan(atan2(W / 2.0, (double)fx));
        tan_half_fovy = (float)tan(atan2(H / 2.0, (double)fy));
        fov_init = true;
    }
    inline bool point_in_frustum(float x, float y, float z) const {
        float c[4];
        mat4_apply(extrinsic, x, y, z, c);
        if (c[2] < cfg.depth_min || c[2] > cfg.depth_max) return false;
        if (std::abs(c[0]) > c[2] * tan_half_fovx) return false;
        if (std::abs(c[1]) > c[2] * tan_half_fovy) return false;
        return true;
    }
    // :1267-1290
    inline bool project_to_image(float x, float y, float z, int& row, int& col, float& cz) const {
        float c[4];
        mat4_apply(extrinsic, x, y, z, c);
        if (c[2] < cfg.depth_min || c[2] > cfg.depth_max) return false;
        // (K * p) / z with K = [fx 0 cx; 0 fy cy; 0 0 1]
        const float u = ((fx * c[0] + 0.f * c[1]) + cx * c[2]) / c[2];
        const float v = ((0.f * c[0] + fy * c[1]) + cy * c[2]) / c[2];
        // DEFINED UB: static_cast<int> of a non-finite / out-of-int-range float; such values are
        // rejected (x86 yields INT_MIN, which fails the range test the same way)
        if (!(u > -2147483648.0f && u < 2147483648.0f) || !(v > -2147483648.0f && v < 2147483648.0f)) return false;
        row = static_cast<int>(v); col = static_cast<int>(u);
        if (row < 0 || row >= H || col < 0 || col >= W) return false;
        cz = c[2];
        return true;
    }

    // body of the per-voxel block, :1358-1435; returns number of non-invalid particles seen
    void sweep_voxel(int mx, int my, int mz) {
        uint32_t bx, by, bz;
        map_to_buffer(mx, my, mz, bx, by, bz);
        const uint32_t v = buffer_to_voxel(bx, by, bz);
        const size_t s = (size_t)v << Pb;
        bool observed = false;
        int nvalid = 0;
        const float one_sigma = DEPTH_ERR_STDDEV_1M + 1.f;  // :1387
        for (uint32_t i = 1; i < P; ++i) {
            Particle* p = &store[s + i];
            if (p->status == ST_INVALID) continue;
            if (p->time_stamp < stamps[0][bx] || p->time_stamp < stamps[1][by] || p->time_stamp < stamps[2][bz]) { p->status = ST_INVALID; continue; }
            ++nvalid;
            int row, col; float cz;
            if (project_to_image(p->x, p->y, p->z, row, col, cz)) {
                const float d = depth_int[(size_t)row * W + col];
                if (d > cfg.depth_max) { p->weight = OCC_INIT_WEIGHT; observed = true; continue; }
                if (cz > d * one_sigma) continue;
                observed = true;
                pix_lists[(size_t)row * W + col].push_back((uint32_t)(s + i));
            }
        }
        ++stats.n_swept_voxels;
        if (nvalid > 0) ++stats.n_swept_nonempty;
        if (observed) {
            store[s].time_stamp = (uint16_t)T;
        } else if (nvalid == 0) {
            // mapXYZIdxToGlobalPose :986-991: the voxel's min corner
            const float gx = mx * voxel + pmin[0] + map_center[0];
            const float gy = my * voxel + pmin[1] + map_center[1];
            const float gz = mz * voxel + pmin[2] + map_center[2];
            int row, col; float cz;
            if (project_to_image(gx, gy, gz, row, col, cz))
                if (cz <= depth_int[(size_t)row * W + col]) store[s].time_stamp = (uint16_t)T;
        }
    }

    void visibility() {
        init_fov();
        for (auto& l : pix_lists) l.clear();  // :653-661
        stats.n_swept_voxels = 0; stats.n_swept_nonempty = 0;
        const int VX = (int)NX + 1, VY = (int)NY + 1, VZ = (int)NZ + 1;
        const float off[3] = {map_center[0] + pmin[0], map_center[1] + pmin[1], map_center[2] + pmin[2]};  // :1304

        // start vertex, :1312-1321
        float inv_ext[16];
        inverse4f(extrinsic, inv_ext);
        const float sg[3] = {(inv_ext[0] * 0.f + inv_ext[1] * 0.f + inv_ext[2] * 1.f) + inv_ext[3],
                             (inv_ext[4] * 0.f + inv_ext[5] * 0.f + inv_ext[6] * 1.f) + inv_ext[7],
                             (inv_ext[8] * 0.f + inv_ext[9] * 0.f + inv_ext[10] * 1.f) + inv_ext[11]};
        int start[3];
        for (int a = 0; a < 3; ++a) start[a] = static_cast<int>(((sg[a] - map_center[a]) + pmax[a]) * voxel_recip);
        // DEFINED UB: the reference indexes visited_points with the unchecked start vertex (:1332);
        // the oracle clamps it into the vertex lattice.
        start[0] = std::min(std::max(start[0], 0), VX - 1);
        start[1] = std::min(std::max(start[1], 0), VY - 1);
        start[2] = std::min(std::max(start[2], 0), VZ - 1);

        auto vertex_in = [&](int x, int y, int z) {
            return point_in_frustum((float)x * voxel + off[0], (float)y * voxel + off[1], (float)z * voxel + off[2]);  // :1338
        };
        const bool start_in = vertex_in(start[0], start[1], start[2]);
        if (!start_in) ++start_vertex_outside;

        auto vidx = [&](int x, int y, int z) { return ((size_t)x * VY + y) * VZ + z; };
        auto cidx = [&](int x, int y, int z) { return ((size_t)x * NY + y) * NZ + z; };

        if (order == 1) {
            // canonical sweep: if the start vertex is in the frustum, every voxel that touches an
            // in-frustum vertex, each exactly once; lists are sorted afterwards.
            if (start_in) {
                std::vector<uint8_t> vin((size_t)VX * VY * VZ);
                for (int x = 0; x < VX; ++x) for (int y = 0; y < VY; ++y) for (int z = 0; z < VZ; ++z) vin[vidx(x, y, z)] = vertex_in(x, y, z);
                for (int z = 0; z < (int)NZ; ++z) for (int y = 0; y < (int)NY; ++y) for (int x = 0; x < (int)NX; ++x) {
                    bool any = false;
                    for (int a = 0; a < 2 && !any; ++a) for (int b = 0; b < 2 && !any; ++b) for (int c = 0; c < 2 && !any; ++c) any = vin[vidx(x + a, y + b, z + c)];
                    if (any) sweep_voxel(x, y, z);
                }
            }
            for (auto& l : pix_lists) std::sort(l.begin(), l.end());
            return;
        }

        // reference BFS, :1299-1456
        std::vector<bool> visited((size_t)VX * VY * VZ, false), added((size_t)NX * NY * NZ, false);
        struct V3 { int x, y, z; };
        std::queue<V3> q;
        q.push({start[0], start[1], start[2]});
        static const int dir[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        while (!q.empty()) {
            const V3 cur = q.front(); q.pop();
            if (visited[vidx(cur.x, cur.y, cur.z)]) continue;
            visited[vidx(cur.x, cur.y, cur.z)] = true;
            if (!vertex_in(cur.x, cur.y, cur.z)) continue;
            for (int dx = -1; dx <= 0; ++dx) for (int dy = -1; dy <= 0; ++dy) for (int dz = -1; dz <= 0; ++dz) {
                const int ax = cur.x + dx, ay = cur.y + dy, az = cur.z + dz;
                if (ax < 0 || ax >= (int)NX || ay < 0 || ay >= (int)NY || az < 0 || az >= (int)NZ) continue;
                if (added[cidx(ax, ay, az)]) continue;
                sweep_voxel(ax, ay, az);
                added[cidx(ax, ay, az)] = true;
            }
            for (int i = 0; i < 6; ++i) {
                const int nx = cur.x + dir[i][0], ny = cur.y + dir[i][1], nz = cur.z + dir[i][2];
                if (nx < 0 || nx > (int)NX || ny < 0 || ny > (int)NY || nz < 0 || nz > (int)NZ || visited[vidx(nx, ny, nz)]) continue;
                q.push({nx, ny, nz});
            }
        }
        if (!diagnostics) return;
        // diagnostic: how far is the BFS-reached voxel set from "touches any in-frustum vertex"?
        uint64_t mismatch = 0;
        for (int x = 0; x < (int)NX; ++x) for (int y = 0; y < (int)NY; ++y) for (int z = 0; z < (int)NZ; ++z) {
            bool any = false;
            if (start_in)
                for (int a = 0; a < 2 && !any; ++a) for (int b = 0; b < 2 && !any; ++b) for (int c = 0; c < 2 && !any; ++c) any = vertex_in(x + a, y + b, z + c);
            if (any != (bool)added[cidx(x, y, z)]) ++mismatch;
        }
        sweep_set_mismatch += mismatch;
    }

    // ---- U3: semantic_dsp_map.h:960-1121 ----------------------------------------------------
    void update_particles() {
        const float Pd = prm.detection_probability, kappa = prm.noise_number, tau = prm.id_transition_probability;
        const bool indep = prm.use_independent_filter != 0;
        uint32_t n_vis = 0;
        for (int i = 0; i < H; ++i) {
            for (int j = 0; j < W; ++j) {
                const LabeledPoint& o = cloud[(size_t)i * W + j];
                if (!o.is_valid) { ck[(size_t)i * W + j] = 0.f; continue; }
                const float sigma = o.sigma;
                float acc = 0.f;
                for (int m = -win_r; m <= win_r; ++m) {
                    for (int n = -win_r; n <= win_r; ++n) {
                        const int ni = i + m, nj = j + n;
                        if (ni < 0 || ni >= H || nj < 0 || nj >= W) continue;
                        const auto& lst = pix_lists[(size_t)ni * W + nj];
                        for (uint32_t pidx : lst) {
                            const Particle* p = &store[pidx];
                            if (indep && p->track_id != o.track_id) continue;
                            float gk = query_pdf(p->x, o.x, sigma) * query_pdf(p->y, o.y, sigma) * query_pdf(p->z, o.z, sigma);
                            if (!indep) {
                                gk *= forgetting_factor(p->forget_count);
                                if (p->track_id != o.track_id) gk *= tau;
                            }
                            acc += p->weight * gk;
                        }
                    }
                }
                ck[(size_t)i * W + j] = acc * Pd + kappa;
            }
        }
        for (int i = 0; i < H; ++i) {
            for (int j = 0; j < W; ++j) {
                const auto& lst = pix_lists[(size_t)i * W + j];
                if (lst.empty()) continue;
                const float sigma = cloud[(size_t)i * W + j].sigma;
                for (uint32_t pidx : lst) {
                    Particle* p = &store[pidx];
                    float acc = 0.f;
                    bool right_id = false;
                    for (int m = -win_r; m <= win_r; ++m) {
                        for (int n = -win_r; n <= win_r; ++n) {
                            const int ni = i + m, nj = j + n;
                            if (ni < 0 || ni >= H || nj < 0 || nj >= W) continue;
                            const LabeledPoint& o = cloud[(size_t)ni * W + nj];
                            if (!o.is_valid) continue;
                            if (indep && o.track_id != p->track_id) continue;
                            float gk = query_pdf(p->x, o.x, sigma) * query_pdf(p->y, o.y, sigma) * query_pdf(p->z, o.z, sigma);
                            if (!indep) {
                                if (p->track_id != o.track_id) gk *= tau;
                                else if (gk > MIN_RIGHTLY_UPDATED_PDF) right_id = true;
                                gk *= forgetting_factor(p->forget_count);
                            }
                            acc += gk / ck[(size_t)ni * W + nj];
                        }
                    }
                    p->weight *= (acc * Pd + 1.f - Pd);
                    p->status = ST_UPDATED;
                    p->time_stamp = (uint16_t)T;
                    if (!indep) {
                        if (right_id) p->forget_count = 0;
                        else if (p->forget_count < 5) p->forget_count += 1;
                    }
                    ++n_vis;
                }
            }
        }
        stats.n_visible = n_vis;
    }

    // ---- B1: semantic_dsp_map.h:768-800, 1148-1230, 1448-1519 -----------------------------
    bool resample_voxel(uint32_t v) {
        const size_t s = (size_t)v << Pb;
        float wsum = 0.f;
        uint32_t n_upd = 0;
        for (uint32_t i = 1; i < P; ++i)
            if (store[s + i].status == ST_UPDATED) { wsum += store[s + i].weight; ++n_upd; }
        const uint32_t trigger = P >> 1;
        if (n_upd <= trigger) return false;
        ++stats.n_resampled;
        if (wsum < 0.01f) {
            for (uint32_t i = 1; i < P; ++i)
                if (store[s + i].status == ST_UPDATED) {
                    store[s + i].status = ST_INVALID;
                    if (order == 0) obj_sets[store[s + i].track_id].erase((uint32_t)(s + i));
                }
            return true;
        }
        float wpp = wsum / trigger;
        if (wpp > 1.f) wpp = 1.f;
        float cum = 0.f, thr = wpp;
        for (uint32_t i = 1; i < P; ++i) {
            Particle& p = store[s + i];
            if (p.status != ST_UPDATED) continue;
            cum += p.weight;
            if (cum < thr) {
                p.status = ST_INVALID;
                if (order == 0) obj_sets[p.track_id].erase((uint32_t)(s + i));
            } else {
                p.weight = wpp;
                thr += wpp;
                while (cum > thr) thr += wpp;
            }
        }
        return true;
    }

    void birth() {
        std::unordered_set<uint32_t> resampled;
        std::unordered_set<uint32_t> touched;
        uint32_t n_born = 0;
        stats.n_resampled = 0;
        const int interval = 3;
        const int nb = prm.nb_ptc_num_per_point;
        for (int rs = 0; rs < interval; ++rs) for (int cs = 0; cs < interval; ++cs)
            for (int i = rs; i < H; i += interval) for (int j = cs; j < W; j += interval) {
                const LabeledPoint& pt = cloud[(size_t)i * W + j];
                if (!pt.is_valid) continue;
                if (!prm.consider_depth_noise) {
                    // :1148-1171
                    uint32_t v;
                    const uint32_t idx = add_new_particle(pt.x, pt.y, pt.z, pt.label_id, pt.track_id, v);
                    if (v != INVALID_INDEX) touched.insert(v);
                    if (idx != INVALID_INDEX) { ++n_born; if (order == 0 && pt.track_id <= max_movable) obj_sets[pt.track_id].insert(idx); }
                    if (v != INVALID_INDEX && resampled.count(v) == 0)
                        if (resample_voxel(v)) resampled.insert(v);
                } else {
                    // :1177-1230
                    for (int n = 0; n < nb; ++n) {
                        float noise[3] = {0.f, 0.f, 0.f};
                        if (nb != 1) {
                            float s = pt.sigma;
                            noise[0] = s * next_random(cursor_map); noise[1] = s * next_random(cursor_map); noise[2] = s * next_random(cursor_map);
                        }
                        const float px = pt.x + noise[0], py = pt.y + noise[1], pz = pt.z + noise[2];
                        uint32_t v;
                        uint32_t idx = add_new_particle(px, py, pz, pt.label_id, pt.track_id, v);
                        if (v != INVALID_INDEX) touched.insert(v);
                        if (idx != INVALID_INDEX) {
                            ++n_born; if (order == 0 && pt.track_id <= max_movable) obj_sets[pt.track_id].insert(idx);
                        } else if (v != INVALID_INDEX && resampled.count(v) == 0) {
                            if (resample_voxel(v)) {
                                resampled.insert(v);
                                idx = add_new_particle(px, py, pz, pt.label_id, pt.track_id, v);
                                if (idx != INVALID_INDEX) { ++n_born; if (order == 0 && pt.track_id <= max_movable) obj_sets[pt.track_id].insert(idx); }
                            }
                        }
                    }
                }
            }
        stats.n_born = n_born;
        stats.n_birth_voxels = (uint32_t)touched.size();
    }

    // ---- semantic_dsp_map.h:1274-1351: the colour of one occupied voxel --------------------------------
    // pt = the emitted position (camera-relative when zero-centred), corner = the voxel corner in the global frame
    void colour_of(int state, uint8_t label, uint16_t track, const float pt[3], const float corner[3], uint8_t rgb[3]) const {
        if (state == 1) {
            if ((int)label == colors.background_label) {  // :1277-1294 jet by height
                const float h = colors.scheme == 3 ? (pt[1] + 2.f) : (-pt[2] + 2.f);
                const int ci = std::min(std::max(static_cast<int>(h * 51.2f), 0), 255);
                jet_256(ci, rgb);
                if (colors.eval_format) rgb[0] = rgb[1] = rgb[2] = 0;
            } else if (colors.scheme == 0 || track > max_movable) {  // :1297-1309 label colour (stored b, g, r)
                rgb[0] = colors.label_bgr[label][2]; rgb[1] = colors.label_bgr[label][1]; rgb[2] = colors.label_bgr[label][0];
            } else if (colors.eval_format) {  // :1311-1315
                rgb[0] = label; rgb[1] = (uint8_t)(track >> 8); rgb[2] = (uint8_t)(track & 0xFF);
            } else {  // :1317-1319; DEFINED UB: the reference indexes the 256-entry table with the 16-bit track id
                rgb[0] = 160; rgb[1] = colors.instance_map[track & 0xFF]; rgb[2] = colors.instance_map[label];
            }
        } else {  // guessed occupied: white, :1325-1331
            rgb[0] = rgb[1] = rgb[2] = 255;
        }
        if (!colors.eval_format) {  // :1333-1351
            uint8_t hsv[3];
            rgb_to_hsv_u8(rgb, hsv);
            if (!point_in_frustum(corner[0], corner[1], corner[2])) hsv[2] = (uint8_t)(hsv[2] * 0.7f);
            hsv_to_rgb_u8(hsv, rgb);
        }
    }

    // ---- O1: semantic_dsp_map.h:1239-1383; operations.h:390-448, 623-648 --------------------
    void occupancy(bool want_free) {
        out_idx.clear(); out_xyz.clear(); out_w.clear(); out_track.clear(); out_label.clear(); out_state.clear(); out_rgb.clear();
        free_idx.clear(); free_xyz.clear();
        uint8_t carried_label = 0; uint16_t carried_track = 0;
        float inv_ext[16];
        inverse4f(extrinsic, inv_ext);
        const float cam[3] = {inv_ext[3], inv_ext[7], inv_ext[11]};  // :1241
        uint32_t n_nonempty = 0;
        for (uint32_t v = 0; v < V; ++v) {
            uint32_t bx, by, bz;
            voxel_to_buffer(v, bx, by, bz);
            const size_t s = (size_t)v << Pb;
            float wsum, gsum = 0.f;
            // DEFINED UB: label/track are uninitialised out-params in the reference when no particle
            // enters the histogram (semantic_dsp_map.h:1245-1246).  A compiled reference keeps the two
            // variables in one place across loop iterations, so what it actually emits is the value the most
            // recent voxel in storage order assigned (an unknown voxel assigns 0, operations.h:402-408; the
            // very first voxel starts from 0).  That is the definition here and in the engine (verified
            // against oracle/_ref); orc_set_undefined_label_carry(0) gives plain zeros instead.
            uint8_t label = carry_undefined_label ? carried_label : 0;
            uint16_t track = carry_undefined_label ? carried_track : 0;
            int state;
            if (!is_voxel_valid(v, bx, by, bz)) {
                wsum = -1.f;
                label = 0; track = 0;
            } else {
                wsum = 0.f;
                std::map<uint16_t, float> tw;
                std::map<uint16_t, uint8_t> tl;
                bool any = false;
                for (uint32_t i = 1; i < P; ++i) {
                    Particle& p = store[s + i];
                    if (is_vacant(p, bx, by, bz)) continue;
                    any = true;
                    wsum += p.weight;
                    if (p.weight > 1.f) p.weight = 1.f;
                    if (p.status == ST_GUESSED_BORN) gsum += p.weight;
                    else if (p.status == ST_UPDATED && p.weight < OCC_INIT_WEIGHT) { p.status = ST_INVALID; continue; }
                    if (tw.find(p.track_id) == tw.end()) tw[p.track_id] = 0.f;
                    tw[p.track_id] += p.weight;
                    tl[p.track_id] = (uint8_t)p.label_id;
                }
                if (any) ++n_nonempty;
                float mw = 0.f;
                for (auto& kv : tw) if (kv.second > mw) { mw = kv.second; track = kv.first; label = tl[kv.first]; }
            }
            if (wsum > prm.occupancy_threshold) state = 1;
            else if (wsum < 0) state = -1;
            else if (gsum >= OCC_INIT_WEIGHT) state = 2;
            else state = 0;
            voxel_state[v] = (int8_t)state;
            carried_label = label; carried_track = track;
            if (state > 0 || (want_free && state == 0)) {
                // voxelIdxToMapFramePos + map_center, operations.h:970-983, 940-944
                uint32_t mx, my, mz;
                buffer_to_map(bx, by, bz, mx, my, mz);
                float p3[3] = {mx * voxel + pmin[0], my * voxel + pmin[1], mz * voxel + pmin[2]};
                for (int a = 0; a < 3; ++a) p3[a] += map_center[a];
                const float corner[3] = {p3[0], p3[1], p3[2]};
                if (prm.zero_center) for (int a = 0; a < 3; ++a) p3[a] = p3[a] - cam[a];
                if (state > 0) {
                    uint8_t rgb[3] = {0, 0, 0};
                    if (colors.enabled) colour_of(state, label, track, p3, corner, rgb);
                    out_rgb.insert(out_rgb.end(), rgb, rgb + 3);
                    out_idx.push_back(v); out_xyz.insert(out_xyz.end(), p3, p3 + 3); out_w.push_back(wsum);
                    out_track.push_back(track); out_label.push_back(label); out_state.push_back((int8_t)state);
                } else {
                    free_idx.push_back(v); free_xyz.insert(free_xyz.end(), p3, p3 + 3);
                }
            }
        }
        stats.n_occupied = (uint32_t)out_idx.size();
        stats.n_nonempty_voxels = n_nonempty;
    }

    // ---- semantic_dsp_map.h:170-251 + 576-955 ------------------------------------------------
    int update(const float* depth, int rows, int cols, int stride, const dspmap_mask* masks, int n_masks,
               const double pos[3], const double quat[4], const dspmap_object_move* moves, int n_moves,
               const int32_t* remove_tracks, int n_remove, int want_free) {
        if (!depth || rows != cfg.image_height || cols != cfg.image_width) return DSPMAP_ERR_IMAGE_SIZE;
        const double t0 = now_s();
        T += 1;  // :173
        project(depth, rows, cols, stride, masks, n_masks, pos, quat);
        const double t1 = now_s();
        const float posf[3] = {(float)pos[0], (float)pos[1], (float)pos[2]};
        update_ego_center(posf);  // :584-585
        const double t2 = now_s();
        move_objects(moves, n_moves, remove_tracks, n_remove);
        const double t3 = now_s();
        // extrinsic, :744-747
        const float qf[4] = {(float)quat[0], (float)quat[1], (float)quat[2], (float)quat[3]};
        float Rf[9];
        quat_to_rot<float>(qf, Rf);
        float c2g[16] = {Rf[0], Rf[1], Rf[2], posf[0], Rf[3], Rf[4], Rf[5], posf[1], Rf[6], Rf[7], Rf[8], posf[2], 0.f, 0.f, 0.f, 1.f};
        inverse4f(c2g, extrinsic);
        visibility();  // :749
        const double t4 = now_s();
        update_particles();  // :759-764
        const double t5 = now_s();
        birth();  // :768-800
        const double t6 = now_s();
        occupancy(want_free != 0);  // :900-910
        const double t7 = now_s();
        stats.time_stamp = T;
        const double st[8] = {t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t7 - t0};
        for (int k = 0; k < 8; ++k) { stage_s[k] += st[k]; stats.ms_stage[k] = (float)(st[k] * 1e3); }
        return DSPMAP_OK;
    }

    void build_csr() {
        const size_t np = (size_t)W * H;
        csr_off.assign(np + 1, 0);
        csr_idx.clear();
        for (size_t q = 0; q < np; ++q) { csr_off[q] = (uint32_t)csr_idx.size(); csr_idx.insert(csr_idx.end(), pix_lists[q].begin(), pix_lists[q].end()); }
        csr_off[np] = (uint32_t)csr_idx.size();
    }
};

}  // namespace

// ---- C API (consumed through ctypes by tests/ and bench.py only) --------------------------
extern "C" {

void* orc_create(const dspmap_config* cfg, int order_mode) {
    if (!cfg) return nullptr;
    Oracle* o = new Oracle();
    if (o->init(*cfg, order_mode) != DSPMAP_OK) { delete o; return nullptr; }
    return o;
}
void orc_destroy(void* h) { delete static_cast<Oracle*>(h); }
int orc_clear(void* h) { auto* o = static_cast<Oracle*>(h); o->clear(); return DSPMAP_OK; }

int orc_set_params(void* h, const dspmap_params* p) {
    auto* o = static_cast<Oracle*>(h);
    if (!p || p->max_forget_count > 5 || p->max_forget_count < 0 || p->nb_ptc_num_per_point < 1) return DSPMAP_ERR_INVALID_ARG;
    o->prm = *p;
    return DSPMAP_OK;
}
int orc_set_label_luts(void* h, const uint16_t static_value_to_track[256], const uint16_t* static_tracks,
                       const uint8_t* static_labels, int32_t n_static, uint16_t max_movable_track, uint16_t sky_track) {
    auto* o = static_cast<Oracle*>(h);
    memcpy(o->static_value_to_track, static_value_to_track, 256 * sizeof(uint16_t));
    o->static_label.clear();
    for (int i = 0; i < n_static; ++i) o->static_label[static_tracks[i]] = static_labels[i];
    o->max_movable = max_movable_track;
    o->sky_track = sky_track;
    return DSPMAP_OK;
}
int orc_set_random_table(void* h, const float* table, uint32_t n) {
    auto* o = static_cast<Oracle*>(h);
    if (!table || n < 1) return DSPMAP_ERR_INVALID_ARG;
    o->randoms.assign(table, table + n);
    return DSPMAP_OK;
}
int orc_update(void* h, const float* depth, int32_t rows, int32_t cols, int32_t row_stride, const dspmap_mask* masks,
               int32_t n_masks, const double position[3], const double orientation_wxyz[4], const dspmap_object_move* moves,
               int32_t n_moves, const int32_t* remove_tracks, int32_t n_remove, int32_t want_freespace) {
    return static_cast<Oracle*>(h)->update(depth, rows, cols, row_stride, masks, n_masks, position, orientation_wxyz, moves,
                                           n_moves, remove_tracks, n_remove, want_freespace);
}
int orc_get_result(void* h, dspmap_result* out) {
    auto* o = static_cast<Oracle*>(h);
    out->count = (uint32_t)o->out_idx.size();
    out->free_count = (uint32_t)o->free_idx.size();
    out->voxel_index = o->out_idx.data(); out->xyz = o->out_xyz.data(); out->weight_sum = o->out_w.data();
    out->track_id = o->out_track.data(); out->label_id = o->out_label.data(); out->state = o->out_state.data();
    out->free_voxel_index = o->free_idx.data(); out->free_xyz = o->free_xyz.data();
    out->rgb = o->out_rgb.data();
    return DSPMAP_OK;
}
int orc_get_stats(void* h, dspmap_frame_stats* out) { *out = static_cast<Oracle*>(h)->stats; return DSPMAP_OK; }
// which: 0 sweep-set mismatch (cumulative), 1 frames whose BFS start vertex was outside the frustum
uint64_t orc_get_extra(void* h, int which) {
    auto* o = static_cast<Oracle*>(h);
    return which == 0 ? o->sweep_set_mismatch : o->start_vertex_outside;
}
int orc_set_colors(void* h, const dspmap_colors* c) {
    if (!h || !c) return DSPMAP_ERR_INVALID_ARG;
    if (c->scheme != 0 && c->scheme != 1 && c->scheme != 3) return DSPMAP_ERR_INVALID_ARG;
    static_cast<Oracle*>(h)->colors = *c;
    return DSPMAP_OK;
}
void orc_set_undefined_label_carry(void* h, int on) { static_cast<Oracle*>(h)->carry_undefined_label = on != 0; }
// test hooks for the two colour-space conversions (tests/test_colour_cpu.py checks them against cv2)
void orc_rgb_to_hsv(const uint8_t* in, uint8_t* out, size_t n) { for (size_t i = 0; i < n; ++i) rgb_to_hsv_u8(in + 3 * i, out + 3 * i); }
void orc_hsv_to_rgb(const uint8_t* in, uint8_t* out, size_t n) { for (size_t i = 0; i < n; ++i) hsv_to_rgb_u8(in + 3 * i, out + 3 * i); }
void orc_set_invalid_sigma(void* h, float v) { auto* o = static_cast<Oracle*>(h); o->invalid_sigma_set = true; o->invalid_sigma_value = v; }
void orc_set_diagnostics(void* h, int on) { static_cast<Oracle*>(h)->diagnostics = on != 0; }
double orc_stage_seconds(void* h, int stage) { return static_cast<Oracle*>(h)->stage_s[stage & 7]; }

int orc_dump(void* h, int32_t what, void* dst, size_t cap, size_t* n_bytes) {
    auto* o = static_cast<Oracle*>(h);
    const void* src = nullptr;
    size_t n = 0;
    std::vector<uint8_t> tmp;
    switch (what) {
        case DSPMAP_DUMP_POINTS: src = o->cloud.data(); n = o->cloud.size() * sizeof(LabeledPoint); break;
        case DSPMAP_DUMP_PARTICLES: src = o->store.data(); n = o->store.size() * sizeof(Particle); break;
        case DSPMAP_DUMP_PIXEL_OFFSETS: o->build_csr(); src = o->csr_off.data(); n = o->csr_off.size() * 4; break;
        case DSPMAP_DUMP_PIXEL_LISTS: o->build_csr(); src = o->csr_idx.data(); n = o->csr_idx.size() * 4; break;
        case DSPMAP_DUMP_CK: src = o->ck.data(); n = o->ck.size() * 4; break;
        case DSPMAP_DUMP_VOXEL_STATE: src = o->voxel_state.data(); n = o->voxel_state.size(); break;
        case DSPMAP_DUMP_PARTICLES_HASH: {   // include/dspmap_c.h: digest of the records above, for maps too large to dump
            auto mix = [](uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
            uint64_t sum = 0;
            for (size_t i = 0; i < o->store.size(); ++i) {
                uint64_t w[3];
                memcpy(w, &o->store[i], 24);
                sum += mix(mix(mix((uint64_t)i * 0x9E3779B97F4A7C15ull + w[0]) ^ w[1]) + w[2]);
            }
            tmp.resize(8);
            memcpy(tmp.data(), &sum, 8);
            src = tmp.data(); n = 8;
            break;
        }
        case DSPMAP_DUMP_RING: {
            tmp.resize(4 * (9 + o->NX + o->NY + o->NZ));
            uint8_t* w = tmp.data();
            memcpy(w, o->moved, 12); w += 12; memcpy(w, o->equiv, 12); w += 12; memcpy(w, o->map_center, 12); w += 12;
            for (int a = 0; a < 3; ++a) { memcpy(w, o->stamps[a].data(), o->stamps[a].size() * 4); w += o->stamps[a].size() * 4; }
            src = tmp.data(); n = tmp.size();
            break;
        }
        default: return DSPMAP_ERR_INVALID_ARG;
    }
    if (n_bytes) *n_bytes = n;
    if (dst && cap) memcpy(dst, src, std::min(cap, n));
    return DSPMAP_OK;
}

}  // extern "C"
