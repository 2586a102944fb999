/*
 * dspmap_c.h — C ABI of the B200-native particle occupancy-map update engine.
 *
 * This is the drop-in boundary for the per-frame hot path of tud-amr/semantic_dsp_map
 * (SemanticDSPMap::update -> generateLabeledPointCloud -> subObjectLevelUpdate).  The
 * reference has no FFI layer of its own: its "operator API" is the public section of
 * class SemanticDSPMap (reference include/semantic_dsp_map.h:23-251).  Every entry
 * point below names the reference interface it replaces.  Plain pointers and sizes
 * only; no C++/torch types cross this boundary; no exceptions; rc 0 = ok, <0 = error.
 *
 * All host pointers are caller-owned and only read during the call, unless stated.
 * One handle = one map (the reference allows exactly one map per process because its
 * store is a namespace-scope global, include/mc_ring/buffer.h:86; handles lift that).
 * A handle is not thread-safe (same as the reference: single ROS spinner thread).
 */
#ifndef DSPMAP_C_H
#define DSPMAP_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dspmap dspmap_t;

/* ---- error codes -------------------------------------------------------- */
enum {
    DSPMAP_OK = 0,
    DSPMAP_ERR_INVALID_ARG = -1, /* null pointer, bad enum, bad size                          */
    DSPMAP_ERR_CONFIG = -2,      /* e.g. index bits > 31 (reference exit(-1), operations.h:55) */
    DSPMAP_ERR_IMAGE_SIZE = -3,  /* depth/mask size != configured size (reference: silent OOB,
                                    semantic_dsp_map.h:212-221 vs :974-979)                    */
    DSPMAP_ERR_CUDA = -4,        /* a CUDA call failed; see dspmap_last_error()               */
    DSPMAP_ERR_CAPACITY = -5,    /* a device work list overflowed its configured capacity      */
    DSPMAP_ERR_NO_DEVICE = -6,   /* no CUDA device: there is NO CPU fallback                   */
    DSPMAP_ERR_UNSUPPORTED = -7, /* option the engine does not implement (stated in DESIGN.md) */
    DSPMAP_ERR_COMM = -8         /* sharded map: peers not connected, or a peer missed a barrier */
};

/* ---- configuration (compile-time constants of the reference made runtime) ---- */
typedef struct dspmap_config {
    /* grid: reference include/settings/settings.h:32-119, include/mc_ring/buffer.h:28-41 */
    int32_t voxel_bits_x;      /* C_VOXEL_NUM_AXIS_X_N                                  */
    int32_t voxel_bits_y;      /* C_VOXEL_NUM_AXIS_Y_N                                  */
    int32_t voxel_bits_z;      /* C_VOXEL_NUM_AXIS_Z_N                                  */
    int32_t slot_bits;         /* C_MAX_PARTICLE_NUM_PER_VOXEL_N (slot 0 = time particle) */
    float   voxel_size;        /* C_VOXEL_SIZE [m]                                      */
    /* camera as delivered by the sensor (the g_camera_*_set constants) */
    float   fx, fy, cx, cy;
    int32_t image_width;       /* g_image_width_set                                     */
    int32_t image_height;      /* g_image_height_set                                    */
    int32_t boost_mode;        /* BOOST_MODE: 1 = internal image is the nearest 1/2 downsample
                                  (settings.h:25-29,136-143; pointcloud_tools.h:1104-1133) */
    float   depth_min;         /* g_depth_range_min                                     */
    float   depth_max;         /* g_depth_range_max                                     */
    int32_t consider_instance; /* g_consider_instance                                   */
    int32_t zed2_filters;      /* the "#if SETTING == 3" blocks of generateLabeledPointCloud:
                                  sky rejection + 3-D bbox gate (pointcloud_tools.h:234-273) */
    int32_t window_half;       /* update window half size; 0 = reference default
                                  (3 when boost_mode else 5, semantic_dsp_map.h:964-970)  */
    /* sharding: the map's storage-z slabs are split over shard_count ranks */
    int32_t shard_rank;        /* 0..shard_count-1                                      */
    int32_t shard_count;       /* 1 = whole map on this GPU                             */
    int32_t device;            /* CUDA device ordinal                                   */
    uint32_t max_visible;      /* capacity of the per-frame visible-particle list, 0 = auto */
    uint32_t reserved0;
} dspmap_config;

/* runtime parameters: reference setMapParameters / setMapOptions /
 * setDepthNoiseModelParameters (semantic_dsp_map.h:101-166) */
typedef struct dspmap_params {
    float   detection_probability;     /* Pd                                   */
    float   noise_number;              /* kappa                                */
    int32_t nb_ptc_num_per_point;      /* newborn particles per point          */
    float   occupancy_threshold;
    float   forgetting_rate;
    int32_t max_forget_count;          /* <= 5 (table size, basic_algorithms.h:33) */
    float   id_transition_probability;
    float   depth_noise_first_order;   /* n1 */
    float   depth_noise_zero_order;    /* n0 */
    int32_t consider_depth_noise;      /* g_if_consider_depth_noise            */
    int32_t use_independent_filter;    /* g_if_use_independent_filter          */
    int32_t zero_center;               /* visualize_with_zero_center_ (output positions minus camera) */
    /* The reference never initialises LabeledPoint::sigma for an invalid pixel and for a pixel the ZED2 bounding-box
     * gate turned into background (utils/pointcloud_tools.h:224-232, 256-272); updateParticles reads it
     * (semantic_dsp_map.h:985, 1047).  0 (default): both sites get the depth-noise model's value (defined behaviour).
     * 1: both sites carry uninit_sigma_value instead - what a COMPILED reference reads there (0.0 in the GCC builds of
     * oracle/build_ref.sh and of the reference node), so that a replay follows the shipped binary voxel for voxel. */
    int32_t uninit_sigma_mode;
    float   uninit_sigma_value;
    /* Order of the particles inside a pixel's list, which is the order pass 1 of updateParticles adds their terms in
     * (semantic_dsp_map.h:995-1031).  0 (default): ascending particle index ("canonical": fastest, independent of the
     * sweep).  1: the REFERENCE's order - the order in which getIdxOfVisibleParitlces' breadth-first search over the
     * lattice vertices reaches the voxels (mc_ring/operations.h:1326-1456: FIFO from the vertex in front of the camera,
     * neighbours pushed +x -x +y -y +z -z, a vertex's eight voxels dx, dy, dz = -1..0 with dz fastest), slot order inside
     * a voxel; the swept set is then exactly what that search reaches.  The search is replayed level by level on the
     * device (one cooperative kernel, ~3 grid-wide barriers per level), so a frame costs a few milliseconds more; every
     * float of a scene without moving objects then equals the reference's bit for bit.  Not available for sharded maps. */
    int32_t reference_iteration_order;
} dspmap_params;

/* one segmentation mask: reference MaskKpts (utils/tracking_result_handler.h:15-26),
 * with the string/keypoint logic pre-evaluated by the host adapter */
typedef struct dspmap_mask {
    const uint8_t* data;   /* MONO8, image_height x image_width (input resolution)      */
    int32_t row_stride;    /* bytes between rows                                        */
    int32_t is_static;     /* MaskKpts.label == "static": pixel value v -> class v+1    */
    int32_t track_id;      /* MaskKpts.track_id after the wrap of semantic_dsp_map.h:179-186 */
    int32_t label_id;      /* g_label_id_map_default[MaskKpts.label]                    */
    double  bbox_min[3];   /* min over kpts_current - 1 m (pointcloud_tools.h:178-199)  */
    double  bbox_max[3];   /* max over kpts_current + 1 m; used only when zed2_filters  */
    int32_t rows, cols;    /* size of the mask image; must equal the depth image's (else DSPMAP_ERR_IMAGE_SIZE: the
                              reference iterates over the mask's own size, pointcloud_tools.h:203-212, the engine
                              indexes masks with depth coordinates).  0, 0 = not stated: the caller vouches for it */
} dspmap_mask;

/* rigid motion of one tracked object for this frame: what objectLevelUpdate (host,
 * semantic_dsp_map.h:306-566) leaves in rigidbody_tmatrix_vec[0] and what
 * subObjectLevelUpdate feeds to moveParticlesInSetsByTransformations
 * (semantic_dsp_map.h:588-699; mc_ring/operations.h:321-362) */
typedef struct dspmap_object_move {
    int32_t track_id;
    float   t[16];         /* row-major 4x4, float (reference casts Matrix4d -> Matrix4f) */
} dspmap_object_move;

/* result view: occupied (state 1), guessed-occupied (state 2) voxels in ascending
 * storage index, as getOccupancyResult emits them (semantic_dsp_map.h:1239-1383) */
typedef struct dspmap_result {
    uint32_t  count;          /* number of entries                                     */
    uint32_t  free_count;     /* entries in the free-space view (0 unless requested)   */
    const uint32_t* voxel_index;   /* storage voxel index                              */
    const float*    xyz;           /* 3 floats per entry: voxel min corner, global frame
                                      (minus camera position when zero_center)         */
    const float*    weight_sum;    /* sum of particle weights of the voxel             */
    const uint16_t* track_id;      /* fused instance id                                */
    const uint8_t*  label_id;      /* fused semantic label                             */
    const int8_t*   state;         /* 1 occupied, 2 guessed occupied                   */
    const uint32_t* free_voxel_index;
    const float*    free_xyz;
    const uint8_t*  rgb;           /* 3 bytes (r, g, b) per entry: the colour getOccupancyResult gives the point
                                      (semantic_dsp_map.h:1274-1351); zeros until dspmap_set_colors enabled it.
                                      Free-space points are always (0, 255, 0) (:1371-1373) and carry no array. */
} dspmap_result;

/* per-frame counters (also the inputs of the roofline accounting) */
typedef struct dspmap_frame_stats {
    uint32_t time_stamp;        /* global_time_stamp after this frame                  */
    uint32_t n_swept_voxels;    /* voxels visited by the frustum sweep                 */
    uint32_t n_swept_nonempty;  /* ... of which held at least one particle             */
    uint32_t n_visible;         /* particles binned to pixels (updated this frame)     */
    uint32_t n_valid_points;    /* valid pixels of the labelled point image            */
    uint32_t n_born;            /* particles born                                       */
    uint32_t n_birth_voxels;    /* distinct voxels touched by births                   */
    uint32_t n_resampled;       /* voxels resampled                                    */
    uint32_t n_occupied;        /* result entries                                      */
    uint32_t n_nonempty_voxels; /* voxels with >=1 non-vacant particle at extraction    */
    uint32_t n_moved;           /* particles moved by object transforms                */
    uint32_t n_removed;         /* particles deleted by track removal                  */
    float    ms_stage[8];       /* device time per stage when profiling is on: project, occ_headers kernel alone
                                   (the HBM-streaming kernel of the roofline), predict, visibility, update, birth,
                                   occupancy, total */
} dspmap_frame_stats;

/* ---- lifecycle ---------------------------------------------------------- */

/* replaces: SemanticDSPMap() + RingBufferOperations() ctor chain
 * (semantic_dsp_map.h:25-67; mc_ring/operations.h:44-64,684-767) */
int  dspmap_create(const dspmap_config* cfg, dspmap_t** out);
void dspmap_destroy(dspmap_t* h);
/* replaces: SemanticDSPMap::clear (semantic_dsp_map.h:74-81) */
int  dspmap_clear(dspmap_t* h);
/* human-readable text of the last error on this handle (never NULL) */
const char* dspmap_last_error(const dspmap_t* h);
/* library/ABI version, e.g. 100 = 1.0.0 */
int  dspmap_version(void);

/* ---- setters ------------------------------------------------------------ */

/* replaces: setMapParameters + setMapOptions + setVisualizeOptions(zero_center) +
 * setDepthNoiseModelParameters + setOccupancyThreshold (semantic_dsp_map.h:101-166) */
int dspmap_set_params(dspmap_t* h, const dspmap_params* p);

/* replaces: the label globals of utils/data_base.h:108-212 as rewritten by
 * ObjectInfoHandler::readObjectInfo (utils/object_info_handler.h:28-91).
 *   static_value_to_track[v] = g_label_to_instance_id_map_default[g_label_id_map_reversed[v+1]]
 *                              evaluated with std::unordered_map::operator[] semantics
 *                              (a missing key yields 0), pointcloud_tools.h:137-138
 *   static_tracks[i] -> static_labels[i]  : g_label_id_map_static[g_instance_id_to_label_map_default[t]]
 *   max_movable_track                     : g_max_movable_object_instance_id
 *   sky_track                             : g_label_to_instance_id_map_default["Sky"] (pointcloud_tools.h:236) */
int dspmap_set_label_luts(dspmap_t* h, const uint16_t static_value_to_track[256],
                          const uint16_t* static_tracks, const uint8_t* static_labels,
                          int32_t n_static, uint16_t max_movable_track, uint16_t sky_track);

/* replaces: the std::random_device-seeded table of
 * GaussianRandomCalculator::calculateGaussianTable (utils/basic_algorithms.h:394-401);
 * injecting it makes runs reproducible.  n must be >= 1. */
int dspmap_set_random_table(dspmap_t* h, const float* table, uint32_t n);

/* Colouring of the occupied list — the rest of getOccupancyResult (include/semantic_dsp_map.h:1274-1351): jet colour by
 * height for background voxels, the label table colour for static classes, instance colours for movable tracks, the
 * evaluation-format packing (label, track >> 8, track & 0xFF), white for guessed voxels, and — unless the evaluation
 * format is on — the 8-bit RGB -> HSV -> RGB round trip of cv::cvtColor with V scaled by 0.7 for voxels outside the
 * camera frustum (checkIfPointInFrustum on the voxel corner).  The HSV conversions restate OpenCV's 8-bit algorithm
 * and are bit-identical to cv2 4.13 on all 2^24 inputs (tests/test_colour_cpu.py). */
typedef struct dspmap_colors {
    int32_t enabled;          /* 0: dspmap_result.rgb stays zero (default)                                  */
    int32_t eval_format;      /* if_out_evaluation_format_ (setVisualizeOptions, semantic_dsp_map.h:130-134) */
    int32_t scheme;           /* which "#if SETTING" block colours a voxel (semantic_dsp_map.h:1279-1322):
                                 0 = SETTING 0: every non-background voxel gets its label colour, jet over -z;
                                 1 = SETTING 1 / 2: static -> label colour, movable -> instance colours, jet over -z;
                                 3 = SETTING 3: as 1, jet over +y                                            */
    int32_t background_label; /* g_label_id_map_default["Background"], default-insert semantics (:1276)      */
    uint8_t label_bgr[256][3];/* g_label_color_map_default[label] as (b, g, r); absent labels read (0,0,0)    */
    uint8_t instance_map[256];/* color_map_int_256_: a std::random_device shuffle of 0..255 in the reference
                                 (:45-48), so it is injected.  The reference indexes it with the 16-bit track id
                                 (out of bounds for ids >= 256); the index is taken modulo 256 here.         */
} dspmap_colors;
int dspmap_set_colors(dspmap_t* h, const dspmap_colors* c);

/* ---- the per-frame hot path ---------------------------------------------- */

/* replaces: SemanticDSPMap::update (semantic_dsp_map.h:170-251) minus the host-side
 * objectLevelUpdate, whose outputs arrive as `moves` / `remove_tracks`.
 * depth: CV_32FC1 metres, rows x cols, row_stride in floats; HOST memory. */
int dspmap_update(dspmap_t* h, const float* depth, int32_t rows, int32_t cols,
                  int32_t row_stride, const dspmap_mask* masks, int32_t n_masks,
                  const double position[3], const double orientation_wxyz[4],
                  const dspmap_object_move* moves, int32_t n_moves,
                  const int32_t* remove_tracks, int32_t n_remove,
                  int32_t want_freespace);

/* same, with depth and mask data already resident in DEVICE memory (bench `value`
 * arm; pointers inside `masks` are device pointers); results stay on the device */
int dspmap_update_device(dspmap_t* h, const float* d_depth, int32_t rows, int32_t cols,
                         int32_t row_stride, const dspmap_mask* masks, int32_t n_masks,
                         const double position[3], const double orientation_wxyz[4],
                         const dspmap_object_move* moves, int32_t n_moves,
                         const int32_t* remove_tracks, int32_t n_remove,
                         int32_t want_freespace);

/* ---- pipelined frames ---------------------------------------------------------------------------------
 * The reference's update() returns with the clouds filled (semantic_dsp_map.h:170-251), so a caller's frame costs
 * staging + host->device copy + kernels + device->host copy, one after the other.  A camera STREAM does not need that:
 * dspmap_update_async stages and enqueues frame t and returns while it runs; the call that enqueues frame t+1 reads
 * frame t's counters once it has finished and starts the transfer of its result on a copy stream, where it overlaps
 * frame t+1's kernels; dspmap_wait_result(t) returns when that transfer is done.  Same arguments, same arithmetic, same
 * results as dspmap_update - only the hand-over is deferred (mapping.cpp would publish frame t's clouds while frame t+1
 * is being computed).  Loop:   update_async(0,&k0); for t = 1..: { update_async(t,&k); wait_result(k_prev,&r); use r; }
 *   *ticket   names the frame (its time stamp; never 0 for an accepted frame).
 *   errors    a frame's capacity / communication errors are found when its counters are read: they are returned by the
 *             NEXT dspmap_update_async (whose own frame is enqueued all the same) and by dspmap_wait_result(ticket).
 *   lifetime  the views of dspmap_wait_result(t) stay valid until dspmap_wait_result(t+2) is called or frame t+3 is
 *             submitted, whichever comes first.
 *   mixing    every other entry point first waits for the frames in flight.  With profiling on, and for sharded maps,
 *             the async calls run the frame synchronously (the ticket works the same).
 * dspmap_update_device_async: the same for images resident in device memory; its result stays on the device
 * (dspmap_wait_result then returns device pointers, as dspmap_get_result_device does). */
int dspmap_update_async(dspmap_t* h, const float* depth, int32_t rows, int32_t cols,
                        int32_t row_stride, const dspmap_mask* masks, int32_t n_masks,
                        const double position[3], const double orientation_wxyz[4],
                        const dspmap_object_move* moves, int32_t n_moves,
                        const int32_t* remove_tracks, int32_t n_remove,
                        int32_t want_freespace, uint32_t* ticket);
int dspmap_update_device_async(dspmap_t* h, const float* d_depth, int32_t rows, int32_t cols,
                               int32_t row_stride, const dspmap_mask* masks, int32_t n_masks,
                               const double position[3], const double orientation_wxyz[4],
                               const dspmap_object_move* moves, int32_t n_moves,
                               const int32_t* remove_tracks, int32_t n_remove,
                               int32_t want_freespace, uint32_t* ticket);
int dspmap_wait_result(dspmap_t* h, uint32_t ticket, dspmap_result* out);

/* result of the last update: host view (pinned buffers owned by the handle, valid until
 * the next update) — what the reference push_backs into the caller's clouds */
int dspmap_get_result(dspmap_t* h, dspmap_result* out);
/* device view of the same (valid after dspmap_update_device too) */
int dspmap_get_result_device(dspmap_t* h, dspmap_result* out);
int dspmap_get_stats(dspmap_t* h, dspmap_frame_stats* out);
/* enable per-stage CUDA-event timing (the reference's VERBOSE_MODE stage timers,
 * semantic_dsp_map.h:912-953) */
int dspmap_set_profiling(dspmap_t* h, int32_t on);

/* run the handle's kernels and copies on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL = the
 * handle's own stream).  Lets the caller bracket frames with its own events / overlap them with other work. */
int dspmap_set_stream(dspmap_t* h, void* cuda_stream);
/* kernels launched by this handle since creation (bench evidence: "gpu_launches") */
int dspmap_get_launch_count(dspmap_t* h, uint64_t* count);

/* number of live particles carrying `track_id` (device-side replacement of
 * ObjectParticleHashMap::indices_map[track].size(), object_layer.h:20-52) */
int dspmap_object_particle_count(dspmap_t* h, int32_t track_id, uint32_t* count);

/* ---- sharded maps (one process per GPU, all GPUs on one NVLink/NVSwitch node) ---------------------- */
/* A map created with shard_count = G > 1 holds the storage-z slabs [rank*NZ/G, (rank+1)*NZ/G) of the particle store.
 * Ranks exchange what the frame couples — visible-particle records, the ck image, moved particles, result offsets —
 * by writing straight into each other's "exchange arena" over NVLink (CUDA IPC peer mappings) and meeting at
 * device-side barriers; there is no host synchronisation and no collective library in the data path.
 *   1. every rank: dspmap_create(cfg with shard_rank/shard_count), dspmap_comm_export(h, handle)
 *   2. the caller's own plumbing all-gathers the G handles (bench.py: torch.distributed)
 *   3. every rank: dspmap_comm_connect(h, handles)   with handles = G * DSPMAP_COMM_HANDLE_BYTES bytes in rank order
 * Every rank then calls dspmap_update* with the SAME frame; rank 0's dspmap_get_result returns the whole map's result
 * (identical, bit for bit, to an unsharded map's), the other ranks return an empty list. */
#define DSPMAP_COMM_HANDLE_BYTES 128
int dspmap_comm_export(dspmap_t* h, uint8_t handle[DSPMAP_COMM_HANDLE_BYTES]);
int dspmap_comm_connect(dspmap_t* h, const uint8_t* handles);

/* ---- parity / debug dumps (tests only) ------------------------------------- */
enum {
    DSPMAP_DUMP_POINTS = 1,   /* Hi*Wi records of 20 B: float x,y,z,sigma; u16 track; u8 label; u8 valid
                                 (LabeledPoint, utils/data_base.h:78-92)                  */
    DSPMAP_DUMP_PARTICLES = 2,/* N records of 24 B in the reference AoS layout (buffer.h:57-79) */
    DSPMAP_DUMP_PIXEL_OFFSETS = 3, /* Hi*Wi+1 u32: CSR offsets of the per-pixel particle lists */
    DSPMAP_DUMP_PIXEL_LISTS = 4,   /* n_visible u32 particle indices, pixel-major, ascending  */
    DSPMAP_DUMP_CK = 5,       /* Hi*Wi f32: ck*Pd+kappa, 0 where the pixel is invalid     */
    DSPMAP_DUMP_RING = 6,     /* ring state: 3 i32 moved steps, 3 i32 equivalent steps, 3 f32 map centre,
                                 then stamps x,y,z as u32                                  */
    DSPMAP_DUMP_VOXEL_STATE = 7, /* V i8: -1 unknown, 0 free, 1 occupied, 2 guessed          */
    DSPMAP_DUMP_PARTICLES_HASH = 8 /* one u64: order-independent digest of the DSPMAP_DUMP_PARTICLES records (for maps whose
                                 store is too large to dump): sum over particle index i of
                                 mix(mix(mix(i * 0x9E3779B97F4A7C15 + (x | y << 32)) ^ (z | w << 32)) + tail), x..w the float
                                 bit patterns, tail the record's last 8 bytes as a little-endian u64, mix = the splitmix64
                                 finaliser.  A sharded map returns the sum over its own slabs (add the ranks' values). */
};
/* test hook: evaluates the engine's hoisted-reciprocal division (csrc/dspmap_device.cuh InvDiv) against the compiler's
 * IEEE division on n pseudo-random operand pairs (divisor in [b_lo, b_hi], |dividend| <= a_max) and counts results that
 * differ in any bit */
int dspmap_selftest_division(dspmap_t* h, uint64_t n, uint64_t seed, float b_lo, float b_hi, float a_max, uint64_t* mismatches);
/* Every 8-bit triple (a, b, c) = (i >> 16, i >> 8 & 255, i & 255), i < 2^24, through the device's RGB->HSV and HSV->RGB
 * conversions (the ones dspmap_set_colors uses); each output word packs the three result bytes, first in bits 0..7. */
int dspmap_selftest_colour(dspmap_t* h, uint32_t* rgb_to_hsv /* 2^24 */, uint32_t* hsv_to_rgb /* 2^24 */);
/* copies at most cap bytes to dst (host); *n_bytes = full size of the dump */
int dspmap_dump(dspmap_t* h, int32_t what, void* dst, size_t cap, size_t* n_bytes);

/* ---- object layer (host only: no GPU, no map handle needed) -----------------------------
 * What the reference decides per frame about its handful of tracked objects before it touches
 * a particle: objectLevelUpdate (semantic_dsp_map.h:306-566: register / update / coast objects;
 * Kabsch + 3-point RANSAC on the key points, utils/basic_algorithms.h:56-195; moving-or-not
 * filter and constant-velocity prediction, object_layer.h:94-258, 467-566) and the head of
 * subObjectLevelUpdate (:588-704: which objects move by which matrix, which are dropped).
 * Its product per frame - `moves` and `dropped` - is what dspmap_update() takes.
 * Implementation: adapter/dsp_object_tracker.h (the C++ adapter uses it directly). */
typedef struct dspmap_objects dspmap_objects_t;

enum {                       /* which "#if SETTING" flavour of the reference's object layer  */
    DSPMAP_OBJECTS_STATIC = 0,            /* SETTING 0: objects never move                   */
    DSPMAP_OBJECTS_ALWAYS_MOVING = 1,     /* SETTING 1: matched key points, every update moves */
    DSPMAP_OBJECTS_MATCHED_KEYPOINTS = 2, /* SETTING 2: matched key points + movement filter  */
    DSPMAP_OBJECTS_BOX_KEYPOINTS = 3      /* SETTING 3: four 3-D box key points per object    */
};

/* one non-static entry of ins_seg_result (utils/tracking_result_handler.h:15-26) with
 * track_id <= g_max_movable_object_instance_id, after the id wrap of semantic_dsp_map.h:179-186 */
typedef struct dspmap_object_observation {
    int32_t track_id;
    int32_t label_known;          /* label found in g_label_id_map_default (:330-334)        */
    int32_t label_id;
    int32_t n_kpts;
    const double* kpts_current;   /* n_kpts x 3, global frame                                */
    const double* kpts_previous;  /* n_kpts x 3 (matched key points) or NULL = same as current */
} dspmap_object_observation;

/* cfg: grid bits, voxel size, intrinsics, image size and boost mode are read (map extent for the
 * "too far to register" rule :356, internal camera for the out-of-view rule :1421-1442) */
int  dspmap_objects_create(const dspmap_config* cfg, int32_t mode, dspmap_objects_t** out);
void dspmap_objects_destroy(dspmap_objects_t* o);
int  dspmap_objects_clear(dspmap_objects_t* o);                       /* object_set_.clear(), :74-81 */
/* setBeyesianMovementParameters, semantic_dsp_map.h:142-148 */
int  dspmap_objects_set_movement_filter(dspmap_objects_t* o, double distance_threshold, double probability_threshold,
                                        double increment, double decrement);
/* the reference seeds its RANSAC from std::random_device; fixed by default here */
int  dspmap_objects_set_seed(dspmap_objects_t* o, uint32_t seed);
/* One frame: objectLevelUpdate on the observations, then the move / drop decisions of the frame.
 * max_lost = max_obersevation_lost_time (setMapParameters :101).  Returns DSPMAP_ERR_CAPACITY when
 * an output array is too small: n_moves / n_dropped then hold the needed sizes, the tracker HAS stepped, and the
 * frame's plan stays available - fetch it with dspmap_objects_last_plan into larger arrays (do not step again). */
int  dspmap_objects_step(dspmap_objects_t* o, const dspmap_object_observation* observations, int32_t n_observations,
                         const double position[3], const double orientation_wxyz[4], double time_stamp, int32_t max_lost,
                         dspmap_object_move* moves, int32_t moves_cap, int32_t* n_moves,
                         int32_t* dropped, int32_t dropped_cap, int32_t* n_dropped);
/* the move / drop decisions of the last dspmap_objects_step again (same outputs, no state change) */
int  dspmap_objects_last_plan(dspmap_objects_t* o, dspmap_object_move* moves, int32_t moves_cap, int32_t* n_moves,
                              int32_t* dropped, int32_t dropped_cap, int32_t* n_dropped);
/* state of one object after the last step: tracked (in object_tracking_hash_map), has_flag
 * (rigidbody_moved_vec not empty), moving, moved_probability, t16 = rigidbody_tmatrix_vec[0]
 * (row-major double 4x4; untouched while has_flag is 0).  Any output may be NULL. */
int  dspmap_objects_query(dspmap_objects_t* o, int32_t track_id, int32_t* tracked, int32_t* has_flag, int32_t* moving,
                          double* moved_probability, double* t16);
/* a "floating" track (:712-737) loses its stored key points */
int  dspmap_objects_forget(dspmap_objects_t* o, int32_t track_id);

#ifdef __cplusplus
}
#endif
#endif /* DSPMAP_C_H */
