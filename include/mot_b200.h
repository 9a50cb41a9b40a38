/* mot_b200.h -- C ABI of the B200-native LiDAR tracker hot path (libmot_b200.so).
 *
 * The reference (MLCS-Yonsei/multiple-object-tracking-lidar) has no plugin/FFI interface; the drop-in
 * boundary is the set of C++ call sites inside ObstacleTrack::clusterPointCloud and callIHGP.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference repository;
 * MOT.cpp = src/multiple_object_tracking_lidar.cpp, MOT.h = include/multiple_object_tracking_lidar/
 * multiple_object_tracking_lidar.h, IHGP.cpp/.hpp = src/ihgp/InfiniteHorizonGP.cpp and its header).
 * include/mot_b200_pcl.hpp is the header-only C++ adapter that keeps those call sites source compatible.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary; nothing throws.
 *   - every function returns an int status: 0 = MOT_OK, negative = error, positive = completed with a warning
 *     (mot_last_error gives the text in both cases).
 *   - points are arrays of 4 floats (x, y, z, pad) == the memory layout of pcl::PointXYZ, so
 *     &cloud.points[0] is passed as-is ("xyz16").
 *   - input buffers are caller owned and only read; output buffers are caller owned, capacities are
 *     passed in elements; all device memory is owned by the handle.
 *   - calls block until their outputs are host visible (the reference consumes results synchronously on
 *     the ROS spinner thread, MOT.cpp:117-121).  A handle is not re-entrant; use one handle per
 *     thread / stream / GPU.
 *   - there is no CPU fallback: without a CUDA device mot_create fails with MOT_ERR_CUDA.
 */
#ifndef MOT_B200_H_
#define MOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mot_handle mot_handle;

enum {
    MOT_OK = 0,
    MOT_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, tolerance <= 0, ...) */
    MOT_ERR_CUDA = -2,      /* CUDA runtime error or no device */
    MOT_ERR_CAPACITY = -3,  /* input larger than the handle's max_points / output buffer too small */
    MOT_ERR_NO_MAP = -4,    /* removeStatic requested before mot_set_map (reference: map_init, MOT.cpp:128) */
    MOT_ERR_STATE = -5,     /* result query without a preceding clustering call / IHGP not configured */
    MOT_ERR_NONFINITE = -6, /* NaN/Inf coordinate in a cloud passed to clustering (reference assumes is_dense) */
    /* positive = the call completed, with a condition the reference handles by carrying on (mot_last_error has the text) */
    MOT_WARN_TRACKS_FULL = 1,    /* mot_tracks_step: max_tracks reached, some centroids were not registered (id -1, zero rows) */
    MOT_WARN_VOXEL_OVERFLOW = 2  /* mot_voxel_grid: leaf too small for the extent; input returned unchanged, as PCL does */
};

/* Per-cluster table row (north_star: centroid / bbox per cluster; 40 bytes). */
typedef struct mot_cluster_stat {
    int32_t count;      /* number of points */
    float mean[3];      /* arithmetic centroid (pcl::compute3DCentroid semantics, MOT.h:62) */
    float bbox_min[3];
    float bbox_max[3];
} mot_cluster_stat;

/* one row per track holding exactly what ObstacleTrack::publishObstacles (MOT.cpp:253-295) writes into a
 * costmap_converter::ObstacleMsg (48 bytes). */
typedef struct mot_obstacle {
    int32_t id;        /* obstacle.id (MOT.cpp:266) */
    float radius;      /* 0.3 (MOT.cpp:267) */
    float x, y;        /* polygon.points[0] = filtered position (MOT.cpp:288-290); z = 0 */
    float vx, vy;      /* velocities.twist.linear (MOT.cpp:272-273); remaining twist components are 0 */
    float vel_cov[6];  /* diagonal of velocities.covariance: .1 .1 1e9 1e9 1e9 .1 (MOT.cpp:279-284) */
} mot_obstacle;

/* Stage timings of the last frame call, milliseconds of GPU time (CUDA events). */
typedef struct mot_timings {
    float remove_static_ms;
    float grid_build_ms;    /* keys + radix sort + reorder + cell tables */
    float union_find_ms;    /* neighbour test + hooking + pointer jumping */
    float cluster_table_ms; /* size filter, ordering, CSR emission */
    float reduce_ms;        /* segmented reduction + circumcentre */
    float total_ms;
} mot_timings;

const char* mot_version(void);

/* Opaque handle owning device workspace, stream and pinned staging for up to max_points points per
 * frame (or per frame batch) and max_tracks IHGP tracks. */
int mot_create(int device, size_t max_points, size_t max_tracks, mot_handle** out);
int mot_destroy(mot_handle* h);
const char* mot_last_error(mot_handle* h);

/* Replaces ObstacleTrack::mapCallback (MOT.cpp:235-251): occ is nav_msgs/OccupancyGrid.data, row-major
 * height x width, -1 unknown, 0..100 occupancy; resolution/origin/orientation are MapMetaData.
 * static_tolerance is the launch parameter `static_tolarance`, clamped to [0,4] (MOT.cpp:95-96).
 * Builds the dilated "blocked" bitmap once so that removeStatic is one bit lookup per point. */
int mot_set_map(mot_handle* h, const int8_t* occ, int width, int height, float resolution, double origin_x,
                double origin_y, const double quat_xyzw[4], int static_tolerance);

/* Launch parameters cluster_tolerance / min_cluster_size / max_cluster_size
 * (MOT.cpp:90-92 -> ec.setClusterTolerance / setMinClusterSize / setMaxClusterSize, MOT.cpp:481-483). */
int mot_set_cluster_params(mot_handle* h, float cluster_tolerance, int min_cluster_size, int max_cluster_size);

/* Replaces ObstacleTrack::removeStatic (MOT.cpp:664-706; decl MOT.h:182).  Kept points are written in
 * input order.  out_xyz16 may alias nothing in xyz16.  *m receives the number kept. */
int mot_remove_static(mot_handle* h, const float* xyz16, size_t n, float* out_xyz16, size_t out_capacity, size_t* m);

/* SURVEY 8f-3 ("next" row): replaces pcl::fromROSMsg (MOT.cpp:448-449) for the x / y / z FLOAT32 fields of a
 * sensor_msgs/PointCloud2: data holds n_points records of point_step bytes, off_* are the byte offsets of the three
 * fields (any alignment), is_bigendian as in the message.  Output is the pcl::PointXYZ layout (pad = 1).  With
 * drop_nonfinite != 0 points with a NaN/Inf coordinate are removed, order preserved (pcl::removeNaNFromPointCloud);
 * with 0 they are kept, as fromROSMsg does.  data / out_xyz16 may be host or device pointers. */
int mot_unpack_pointcloud2(mot_handle* h, const uint8_t* data, size_t n_points, uint32_t point_step, uint32_t off_x,
                           uint32_t off_y, uint32_t off_z, int is_bigendian, int drop_nonfinite, float* out_xyz16,
                           size_t out_capacity, size_t* m);

/* SURVEY 8f-1 ("next" row): replaces pcl::VoxelGrid::setLeafSize + filter (MOT.cpp:452-456; the tracker uses the
 * leaf (L, L, 20 L), L = voxel_leaf_size).  One centroid per occupied voxel, emitted in ascending voxel index
 * (ijk = floor(p * inv_leaf) - min_b in fp32, index = i + j*dx + k*dx*dy), exactly PCL's arithmetic; the mean is
 * accumulated in fp64 (PCL: fp32 in an unspecified order), so coordinates agree to ~1 ulp.  xyz16 / out_xyz16 may be
 * host or device pointers. */
int mot_voxel_grid(mot_handle* h, const float* xyz16, size_t n, float leaf_x, float leaf_y, float leaf_z, float* out_xyz16,
                   size_t out_capacity, size_t* m);

/* Replaces pcl::search::KdTree::setInputCloud + pcl::EuclideanClusterExtraction::extract
 * (MOT.cpp:472-488).  Output is the CSR form of std::vector<pcl::PointIndices>: cluster c owns
 * point_indices[cluster_offsets[c] .. cluster_offsets[c+1]), indices ascending inside a cluster (PCL sorts
 * them), clusters ordered by size descending (PCL's extract()), ties by smallest index ascending.
 * Components with fewer than min or more than max points are dropped whole. */
int mot_cluster(mot_handle* h, const float* xyz16, size_t m, int32_t* cluster_offsets, size_t offsets_capacity,
                int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters);

/* Per-cluster count / mean / bbox of the clusters found by the last mot_cluster / mot_frame call. */
int mot_cluster_stats(mot_handle* h, mot_cluster_stat* stats, size_t capacity);

/* Replaces ObstacleTrack::getCentroid (MOT.cpp:708-822; decl MOT.h:184-187) for the clusters of the last
 * mot_cluster / mot_frame call: farthest pair, farthest point from its XY line, XY circumcentre.
 * out_xyzi: n_clusters x 4 floats (x, y, z = 0, intensity = stamp_minus_time_init), i.e. the payload of
 * the pcl::PointXYZI the reference returns. */
int mot_get_centroid(mot_handle* h, double stamp_minus_time_init, float* out_xyzi, size_t capacity);

/* Fused frame step: removeStatic -> clustering -> table (+ optional circumcentres) without host round trips
 * of the cloud (ObstacleTrack::clusterPointCloud minus VoxelGrid, MOT.cpp:461-491).
 * Optional outputs may be NULL: kept_xyz16 (capacity kept_capacity points), stats, centroids_xyzi
 * (both with table_capacity rows).  Indices refer to positions in the kept cloud, as in the reference. */
int mot_frame(mot_handle* h, const float* xyz16, size_t n, double stamp_minus_time_init, float* kept_xyz16,
              size_t kept_capacity, size_t* m, int32_t* cluster_offsets, size_t offsets_capacity,
              int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters, mot_cluster_stat* stats,
              float* centroids_xyzi, size_t table_capacity);

/* ---- device-resident entry points (inputs already in HBM; used by bench.py's `value` leg and by callers that
 * keep the cloud on the GPU).  d_xyz16 is a device pointer on the handle's device.  Results stay on the
 * device; only the 64-byte result header crosses PCIe.  do_remove_static = 0 clusters the cloud as is. */
int mot_frame_device(mot_handle* h, const float* d_xyz16, size_t n, int do_remove_static, int with_centroids,
                     double stamp_minus_time_init);
/* Counts of the last result: kept points, clusters, total indices. */
int mot_result_counts(mot_handle* h, size_t* m, int32_t* n_clusters, size_t* n_indices);
/* Grid statistics of the last result: occupied fine (clique) cells, occupied coarse cells, voxel key width. */
int mot_result_grid(mot_handle* h, int32_t* fine_cells, int32_t* coarse_cells, int32_t* key_bits);
/* Internal device counters of the last result (diagnostics for the benches and tests, no reference counterpart): [0] fine
 * cells, [1] coarse cells, [2] clusters, [3] indices, [4] flags, [5] kept points, [9]/[10] fine-cell pairs handed to the
 * cooperative witness search (ring 1 / ring 2), [12] pairs searched serially because that list was full; up to 16 ints. */
int mot_result_counters(mot_handle* h, int32_t* out, int capacity);
/* Speculative grid plan (no reference counterpart; PCL builds its KD-tree from scratch every frame, MOT.cpp:472-473).  A
 * clustering call on an already compacted cloud (mot_cluster*, mot_frame* without removeStatic) first tries the voxel grid of
 * the handle's previous call, padded to the range of its key bits: no bounding-box pass and no host round trip before the keys.
 * The key kernel checks every point; a cloud that left the planned grid is clustered again from its own bounding box (a miss
 * costs one wasted pass, never a wrong result).  enable: 1 / 0 switch it (and drop the current plan), -1 leaves it; hits /
 * misses (may be NULL) receive the counts since mot_create.  Default on; MOT_PLAN_SPEC=0 in the environment turns it off. */
int mot_grid_plan(mot_handle* h, int enable, int* hits, int* misses);
/* Single-launch path for small frames (no reference counterpart; it exists because a frame of the size the reference tracker
 * sees -- 65,536 points, MOT.cpp:461-491 -- spends its time in launches, not in kernels).  Single-frame calls (mot_cluster,
 * mot_frame, mot_frame_device, mot_cluster_pointcloud2) with at most max_points input points (default 131072; environment
 * MOT_SMALL_POINTS) run removeStatic + clustering + tables as ONE kernel on one thread-block cluster and synchronise once.
 * A frame that kernel cannot take (a crowded cell, > 4096 clusters, very large clusters) is handed back and clustered by the
 * general path in the same call -- same results either way.  max_points: >= 0 sets the limit (0 = never), < 0 leaves it;
 * hits / misses (may be NULL): frames served / handed back since mot_create.  Returns the cluster size in CTAs (0 = the device
 * cannot schedule the cluster: the path is off) or a negative error. */
int mot_small_frames(mot_handle* h, int max_points, int* hits, int* misses);
/* Diagnostics: the GPU's %globaltimer (ns) at the end of every phase of the last small-frame launch, [0] = kernel start, [1..13] =
 * phases A..M as listed in csrc/frame_small.cuh (taken by one thread of the first CTA; 16 values). */
int mot_small_frame_phases(mot_handle* h, uint64_t* ns, int capacity);
/* Union-find diagnostics accumulated since the last call (finds, parent hops, unions, ...; cell_uf.cuh ST_*).  Only a
 * library built with -DMOT_UF_STATS counts (returns 1); the product build returns 0 and zeroes. */
int mot_debug_stats(mot_handle* h, uint64_t* out, int capacity);
/* Device pointers of the last result (valid until the next call on the handle). */
int mot_result_device_ptrs(mot_handle* h, const float** d_kept_xyz16, const int32_t** d_cluster_offsets,
                           const int32_t** d_point_indices, const mot_cluster_stat** d_stats,
                           const float** d_centroids_xyzi);
/* Copies the last result to caller buffers (any may be NULL).  Destinations may be host or device pointers
 * (cudaMemcpyDefault), so a rank can land its table straight in a buffer it is about to gather. */
int mot_result_fetch(mot_handle* h, float* kept_xyz16, size_t kept_capacity, int32_t* cluster_offsets,
                     size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity,
                     mot_cluster_stat* stats, float* centroids_xyzi, size_t table_capacity);
/* Component label of every kept point after union-find (smallest point index of its component; no size
 * filter).  Debug/parity aid for partition checks at sizes where CSR comparison is unwieldy. */
int mot_result_labels(mot_handle* h, int32_t* labels, size_t capacity);
int mot_last_timings(mot_handle* h, mot_timings* t);

/* Number of kernels the last frame / batch / IHGP call launched. */
int mot_last_launches(mot_handle* h);

/* Per-kernel profile: when on, every launch is bracketed by a CUDA event pair on the handle's stream and the
 * durations accumulate per kernel id until the next mot_set_profiling call.  Off by default (events cost time). */
int mot_set_profiling(mot_handle* h, int on);
int mot_profile_kernels(void);
const char* mot_profile_kernel_name(int kernel_id);
int mot_profile_read(mot_handle* h, float* ms_total, int32_t* launches, int capacity);

/* CUDA-event stopwatch on the handle's stream (the stream every kernel of this handle is launched on).
 * mot_timer_stop first waits for all work of the process on the device, so it also brackets other handles' streams. */
int mot_timer_start(mot_handle* h);
int mot_timer_stop(mot_handle* h, float* ms);

/* Pins / unpins a caller-owned host buffer so copies run at full PCIe speed (optional). */
int mot_host_register(void* ptr, size_t bytes);
int mot_host_unregister(void* ptr);

/* ---- frame batches (BASELINE config 3: multi-LiDAR / multi-sequence streams).  n_frames clouds are
 * concatenated in xyz16; frame f owns points [frame_offsets[f], frame_offsets[f+1]).  Frames never share
 * clusters.  cluster c belongs to frame f iff frame_cluster_offsets[f] <= c < frame_cluster_offsets[f+1];
 * point_indices are positions inside the owning frame's cloud.  Clustering only; mot_frame_batch is the full path.
 * mot_cluster_stats / mot_result_fetch give the per-cluster table of the whole batch (same cluster order). */
int mot_cluster_batch(mot_handle* h, const float* xyz16, const int64_t* frame_offsets, int n_frames,
                      int32_t* frame_cluster_offsets /* n_frames+1 */, int32_t* cluster_offsets,
                      size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters);
int mot_cluster_batch_device(mot_handle* h, const float* d_xyz16, const int64_t* frame_offsets, int n_frames);

/* The whole per-frame path on a batch (BASELINE config 3 with the map): removeStatic with the handle's map (optional) ->
 * clustering -> table (+ circumcentres when centroids_xyzi != NULL).  What ObstacleTrack::clusterPointCloud (MOT.cpp:461-491)
 * does for one cloud, for n_frames clouds in one pass.  xyz: concatenated clouds, point_stride_bytes = 16 (pcl::PointXYZ)
 * or 12 (packed x, y, z: a quarter fewer PCIe bytes, expanded on the device).  frame_stamps (optional, n_frames floats) are
 * the centroid intensities (stamp - time_init per frame, MOT.cpp:491).  frame_kept_offsets (optional, n_frames+1): boundaries
 * of the frames inside the kept cloud after removeStatic; point_indices are positions inside the owning frame's kept cloud.
 * Output pointers may be host or device pointers. */
int mot_frame_batch(mot_handle* h, const float* xyz, int point_stride_bytes, const int64_t* frame_offsets, int n_frames,
                    int do_remove_static, const float* frame_stamps, int32_t* frame_kept_offsets,
                    int32_t* frame_cluster_offsets, int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices,
                    size_t indices_capacity, int32_t* n_clusters, mot_cluster_stat* stats, float* centroids_xyzi,
                    size_t table_capacity);
/* Same with the clouds already on the device (16-byte points); results stay on the device (mot_result_*). */
int mot_frame_batch_device(mot_handle* h, const float* d_xyz16, const int64_t* frame_offsets, int n_frames, int do_remove_static,
                           int with_centroids, const float* frame_stamps);

/* the same batch call over several handles = several GPUs of one box (one handle per GPU; more than one handle on
 * a GPU overlaps one range's copies with another's kernels).  One host thread per handle; handle g takes the contiguous frame
 * range [n_frames*g/n, n_frames*(g+1)/n); there is no exchange between the ranges.  The merged tables land in the caller's
 * buffers -- host memory, or device memory of any GPU (the copies are cudaMemcpyDefault) -- in frame order, exactly as one
 * mot_frame_batch call over all frames would leave them.  Every handle must hold the same map / parameters.  The handles'
 * own results are consumed by the merge. */
int mot_batch_run(mot_handle* const* handles, int n_handles, const float* xyz, int point_stride_bytes, const int64_t* frame_offsets,
                  int n_frames, int do_remove_static, const float* frame_stamps, int32_t* frame_kept_offsets,
                  int32_t* frame_cluster_offsets, int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices,
                  size_t indices_capacity, int32_t* n_clusters, mot_cluster_stat* stats, float* centroids_xyzi,
                  size_t table_capacity);

/* ObstacleTrack::clusterPointCloud in one call (MOT.cpp:444-505): pcl::fromROSMsg (:448-449) -> VoxelGrid with leaf
 * (L, L, 20 L) (:452-456, skipped when voxel_leaf_size <= 0) -> removeStatic (:461, optional) -> EuclideanClusterExtraction
 * (:472-488) -> getCentroid (:491, when centroids_xyzi != NULL).  The cloud never returns to the host between the stages
 * (only the stage sizes do).  data: the PointCloud2 payload (host or device pointer); non-finite points are dropped after the
 * unpack, as PCL's filters do for a non-dense cloud.  kept_xyz16 (optional) receives the cloud that was clustered; indices
 * refer to it. */
int mot_cluster_pointcloud2(mot_handle* h, const uint8_t* data, size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                            uint32_t off_z, int is_bigendian, float voxel_leaf_size, int do_remove_static,
                            double stamp_minus_time_init, float* kept_xyz16, size_t kept_capacity, size_t* m,
                            int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity,
                            int32_t* n_clusters, mot_cluster_stat* stats, float* centroids_xyzi, size_t table_capacity);

/* ---- IHGP track filter.  Replaces Matern32model + InfiniteHorizonGP construction in
 * registerNewObstacle (MOT.cpp:521-534; IHGP.cpp:12-37; M32.cpp:15-24): hyp = {sigma2, magnSigma2,
 * lengthScale} already exponentiated as at MOT.cpp:524-530; dt = dt_gp = 1/frequency (MOT.cpp:159);
 * data_length = launch parameter data_length (MOT.cpp:113). */
int mot_ihgp_configure(mot_handle* h, double dt, float lpf_tau, const double hyp_x[3], const double hyp_y[3],
                       int data_length);
/* 16 doubles per axis: A[4], AKHA[4], K[2], G[4], S, lambda (row-major 2x2); axis 0 = x, 1 = y. */
int mot_ihgp_constants(mot_handle* h, int axis, double* consts16);
/* Replaces the per-track loop of ObstacleTrack::callIHGP (MOT.cpp:621-662): LPF_pos (MOT.cpp:824-833),
 * IHGP_fixed_vel (MOT.cpp:871-920) = init_step + (L-1) x update + getEft (IHGP.cpp:108-196), +-1.5 m/s
 * clamp (MOT.cpp:649-654).  rings: T x L x 4 floats (x, y, z, intensity=time), oldest sample first
 * (stack_obj, MOT.h:107).  m_state: T x 4 doubles (m_x[2], m_y[2]), in/out -- the cross-frame carry the
 * reference keeps inside each InfiniteHorizonGP object.  pos_vel: T x 8 floats (pos xyzi, vel xyzi). */
int mot_ihgp_step(mot_handle* h, const float* rings, int n_tracks, double* m_state, float* pos_vel);
/* Same step, and additionally the packed obstacle table of SURVEY 8f-4 (one mot_obstacle per track; track_ids may be
 * NULL, then id = row index): the payload publishObstacles (MOT.cpp:253-295) puts on the wire. */
int mot_ihgp_step_obstacles(mot_handle* h, const float* rings, int n_tracks, const int32_t* track_ids, double* m_state,
                            float* pos_vel, mot_obstacle* obstacles);

/* ---- SURVEY 8f-2 ("next" row): data association + track lifecycle on the device.  Replaces the non-first-frame body of
 * ObstacleTrack::cloudCallback (MOT.cpp:176-233): first-match association in registration order (non exclusive;
 * tracks registered earlier in the same frame are matchable), fill_with_linear_interpolation (:593-619),
 * updateObstacleQueue (:586-591), registerNewObstacle (:507-543), callIHGP over this_objIDs (:621-662, a track matched
 * twice is filtered twice) and unregisterOldObstacle (:545-584, every 5*frequency callbacks, tracks unseen for > 5 s).
 * The track table (ids, rings of data_length centroids, IHGP state) lives on the device between calls.
 * centroids_xyzi: n_centroids x 4 floats (x, y, z, intensity = stamp - time_init) in cluster order -- the output of
 * mot_get_centroid.  now = stamp - time_init.  The first call after mot_tracks_reset (or mot_create) is the reference's
 * first frame: every centroid registers a track and *produced = 0 (nothing is filtered or published, MOT.cpp:126-161).
 * Outputs per centroid: this_obj_ids, pos_vel (8 floats, as mot_ihgp_step) and optionally the obstacle table. */
int mot_tracks_reset(mot_handle* h);
/* Test aid (host only, no device needed): the double S for which the reference's association test
 * (float)sqrt(dx*dx + dy*dy + 0) < id_threshold (euc_dist, MOT.cpp:1025-1028, used at :184-207) equals  dx*dx + dy*dy + 0 < S.
 * Both roundings are monotone, so S exists and is found by bisection over the bit patterns; the device compares against it
 * instead of taking a 64-bit square root per track (csrc/tracks.cuh). */
double mot_assoc_match_below(float id_threshold);
int mot_tracks_step(mot_handle* h, const float* centroids_xyzi, int n_centroids, double now, float id_threshold,
                    float frequency, int32_t* this_obj_ids, float* pos_vel, mot_obstacle* obstacles, int32_t* n_tracks,
                    int32_t* produced);
/* Copies the track table out (ids, rings n x data_length x 4 floats, m_state n x 4 doubles); any pointer may be NULL. */
int mot_tracks_get(mot_handle* h, int32_t* ids, float* rings, double* m_state, size_t capacity, int32_t* n_tracks);

#ifdef __cplusplus
}
#endif
#endif /* MOT_B200_H_ */
