/*
 * lego_loam_b200.h -- C ABI of the B200 (sm_100a) implementation of LeGO-LOAM-BOR's
 * per-scan hot path.  Plain pointers and sizes only; no C++/torch types cross it.
 *
 * The reference has no FFI: its boundary is three C++ classes, two payload structs
 * and the yaml keys.  Each entry point below names the reference
 * interface it stands behind (paths relative to /root/reference/LeGO-LOAM):
 *
 *   ll_image_projection      ImageProjection::cloudHandler          src/imageProjection.cpp:153-174
 *   ll_feature_association   FeatureAssociation::runFeatureAssociation loop body
 *                                                                    src/featureAssociation.cpp:1386-1450
 *   ll_scan_to_map           MapOptimization::scan2MapOptimization  src/mapOptmization.cpp:1315-1332
 *   ll_map_save_keyframe / ll_map_extract_surrounding_keyframes / ll_mapping_cycle
 *                            MapOptimization::saveKeyFramesAndFactor / extractSurroundingKeyFrames / run
 *                                                                    src/mapOptmization.cpp:1335-1474, 856-996, 1545-1560
 *   ll_set_scans_pointcloud2_host   pcl::fromROSMsg + removeNaNFromPointCloud  src/imageProjection.cpp:159-161
 *   LegoLoamParams           config/loam_config.yaml:1-35 (same key names)
 *   LL_BUF_SEG_* / ll_cloud_info   cloud_msgs/msg/cloud_info.msg:1-13
 *
 * A handle owns all device memory for `batch` independent sequences that advance
 * in lock step (one scan per sequence per call).  One CUDA stream per handle; a handle
 * is not thread-safe (the reference drives each stage object from exactly one thread).
 * Every function returns 0 on success or a negative ll_status; nothing throws.
 * There is no CPU fallback: if no CUDA device is usable, ll_create fails.
 */
#ifndef LEGO_LOAM_B200_H
#define LEGO_LOAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ll_handle ll_handle;

typedef enum ll_status {
  LL_OK = 0,
  LL_ERR_INVALID_ARG = -1,
  LL_ERR_CUDA = -2,
  LL_ERR_NO_DEVICE = -3,
  LL_ERR_CAPACITY = -4,
  LL_ERR_STATE = -5
} ll_status;

/* Fields are the 21 keys of config/loam_config.yaml, same names and units. */
typedef struct LegoLoamParams {
  /* laser: */
  int32_t num_vertical_scans;
  int32_t num_horizontal_scans;
  int32_t ground_scan_index;
  float vertical_angle_bottom; /* degrees */
  float vertical_angle_top;    /* degrees */
  float sensor_mount_angle;    /* degrees */
  float scan_period;           /* seconds */
  /* imageProjection: */
  int32_t segment_valid_point_num;
  int32_t segment_valid_line_num;
  float segment_theta; /* degrees */
  /* featureAssociation: */
  float edge_threshold;
  float surf_threshold;
  float nearest_feature_search_distance;
  /* mapping: */
  int32_t enable_loop_closure;
  int32_t mapping_frequency_divider;
  float surrounding_keyframe_search_radius;
  int32_t surrounding_keyframe_search_num;
  float history_keyframe_search_radius;
  int32_t history_keyframe_search_num;
  float history_keyframe_fitness_score;
  float global_map_visualization_search_radius;
} LegoLoamParams;

/* Fills *p with the values of the reference's config/loam_config.yaml (VLP-16). */
void ll_default_params(LegoLoamParams* p);

/* Device-resident arrays that ll_download / ll_upload can address, per sequence.
 * Element types are given on the right; "pt" = 4 floats (x, y, z, intensity), the
 * 16-byte form of pcl::PointXYZI (utility.h:46). */
typedef enum ll_buffer {
  /* ImageProjection state (imageProjection.h:30-81) */
  LL_BUF_RANGE_MAT = 0,   /* f32[N]  _range_mat, row-major, FLT_MAX = no return */
  LL_BUF_FULL_CLOUD = 1,  /* pt[N]   _full_cloud (NaN xyz where empty) */
  LL_BUF_GROUND_MAT = 2,  /* i8[N]   _ground_mat */
  LL_BUF_LABEL_MAT = 3,   /* i32[N]  _label_mat (private to ImageProjection, imageProjection.h; the per-scan path decides everything
                             cloudSegmentation reads from it -- > 0 / == 999999 -- from the component forest and numbers the roots;
                             the numbers of the other cells are filled in when this buffer is first read after a scan) */
  /* ProjectionOut (utility.h:64-70) and cloud_info (cloud_info.msg:1-13) */
  LL_BUF_SEG_CLOUD = 4,        /* pt[S]  segmented_cloud (after ll_feature_association: axis-swapped + time-tagged) */
  LL_BUF_SEG_GROUND_FLAG = 5,  /* u8[S]  segmentedCloudGroundFlag */
  LL_BUF_SEG_COL_IND = 6,      /* u32[S] segmentedCloudColInd */
  LL_BUF_SEG_RANGE = 7,        /* f32[S] segmentedCloudRange */
  LL_BUF_START_RING_INDEX = 8, /* i32[V] */
  LL_BUF_END_RING_INDEX = 9,   /* i32[V] */
  LL_BUF_ORIENTATION = 10,     /* f32[3] startOrientation, endOrientation, orientationDiff */
  LL_BUF_OUTLIER_CLOUD = 11,   /* pt[O]  outlier_cloud */
  /* FeatureAssociation state (featureAssociation.h:66-70) */
  LL_BUF_CLOUD_CURVATURE = 12,  /* f32[N] */
  LL_BUF_NEIGHBOR_PICKED = 13,  /* i32[N] */
  LL_BUF_CLOUD_LABEL = 14,      /* i32[N] */
  LL_BUF_CORNER_SHARP = 15,     /* pt[..] cornerPointsSharp */
  LL_BUF_CORNER_LESS_SHARP = 16,
  LL_BUF_SURF_FLAT = 17,
  LL_BUF_SURF_LESS_FLAT = 18,
  LL_BUF_CORNER_SHARP_IND = 19, /* i32[..] index into the segmented cloud of every picked point */
  LL_BUF_CORNER_LESS_SHARP_IND = 20,
  LL_BUF_SURF_FLAT_IND = 21,
  LL_BUF_CORNER_LAST = 22, /* pt[..] laserCloudCornerLast */
  LL_BUF_SURF_LAST = 23,   /* pt[..] laserCloudSurfLast */
  LL_BUF_TRANSFORM_CUR = 24, /* f32[6] */
  LL_BUF_TRANSFORM_SUM = 25, /* f32[6] */
  LL_BUF_ODOM_ITERS = 26,    /* i32[2] LM iterations actually run (surf, corner) in the last frame */
  /* MapOptimization scan-to-map slice (mapOptimization.h:120-150,190-210) */
  LL_BUF_MAP_CORNER = 27,    /* pt[..] laserCloudCornerFromMapDS */
  LL_BUF_MAP_SURF = 28,      /* pt[..] laserCloudSurfFromMapDS */
  LL_BUF_SCAN_CORNER_DS = 29,     /* pt[..] laserCloudCornerLastDS */
  LL_BUF_SCAN_SURF_TOTAL_DS = 30, /* pt[..] laserCloudSurfTotalLastDS */
  LL_BUF_TRANSFORM_TOBE_MAPPED = 31, /* f32[6] */
  LL_BUF_MAP_ITERS = 32,             /* i32[2] iterations run, rows in the last iteration */
  LL_BUF_OUTLIER_LAST = 33,          /* pt[O] outlier cloud after adjustOutlierCloud (featureAssociation.cpp:1273-1283) */
  LL_BUF_SURF_LESS_FLAT_RAW_COUNT = 34, /* i32[V] less-flat points per ring before the 0.2 m VoxelGrid */
  LL_BUF_MAP_TRACE = 35, /* f64[10][34] per scan-to-map iteration: 21 J^T J (upper), 6 J^T r, rows, 6 step X */
  LL_BUF_TRANSFORM_BEF_MAPPED = 36, /* f32[6] */
  LL_BUF_TRANSFORM_AFT_MAPPED = 37, /* f32[6] */
  LL_BUF_SCAN_SURF_DS = 38,    /* pt[..] laserCloudSurfLastDS (before the outlier cloud is appended) */
  LL_BUF_SCAN_OUTLIER_DS = 39, /* pt[..] laserCloudOutlierLastDS */
  LL_BUF_STAGE_CLOCKS = 40,    /* i64[16] profiling aid: nanoseconds the two scan-to-scan LM stages of the last frame spent per phase
                                  (stage*8 + {0 total, 1 prologue, 2 re-search, 3 rows + reduction, 4 solve, 5 re-searched points}); the per-phase
                                  entries only after ll_enable_stage_timing(h, 1) */
  LL_BUF_RING_CLOCKS = 47,     /* i64[V][10] profiling aid: SM cycles the feature extraction of every ring of the last frame spent per
                                  phase ({0 total, 1 load + keys, 2 speculative picks, 3 boundary check + re-runs, 4 persist + less-flat
                                  collection, 5 bounding box + voxel keys, 6 run heads, 7 run sort, 8 centroids, 9 span | n_raw << 16 |
                                  runs << 32 | re-run rounds << 48}) */
  /* MapOptimization key frames and local map (mapOptimization.h:96-135; needs ll_map_enable_keyframes) */
  LL_BUF_KEYFRAME_STATE = 41,      /* i32[4] key frames stored (cloudKeyPoses3D->size()), surroundingExistingKeyPosesID.size(),
                                      1 if the last extractSurroundingKeyFrames erased a key frame, capacity error bits (0 = none:
                                      1 key-frame slots, 2 point pool, 4 voxel table, 8 local-map output, 16 voxel index range) */
  LL_BUF_KEY_POSES_6D = 42,        /* f32[K][6] cloudKeyPoses6D as (roll, pitch, yaw, x, y, z) */
  LL_BUF_SURROUNDING_KEY_IDS = 43, /* i32[..] surroundingExistingKeyPosesID */
  LL_BUF_INPUT_CLOUD = 44,         /* pt[n] the scan staged by the last ll_set_scans_* call (after NaN removal) */
  /* index-level parity aids, filled only after ll_enable_index_trace(h, 1) */
  LL_BUF_MAP_KNN_IDX = 45,         /* i32[10][Q][5] the 5 nearest map points (indices into MAP_CORNER for the Qc corner queries, which come
                                      first, into MAP_SURF for the Qs surf queries; ascending distance) of every query in every
                                      scan-to-map LM iteration of the last ll_scan_to_map; -1 x 5: fifth neighbour not closer than 1 m
                                      (mapOptmization.cpp:1036,1144) or iteration not run.  Q = Qc + Qs */
  LL_BUF_ODOM_SEARCH_IDX = 46,     /* i32[2][5][24V][3] scan-to-scan correspondences (pointSearchSurfInd1/2/3, pointSearchCornerInd1/2,
                                      featureAssociation.h:87-93) of stage 0 surf / 1 corner in search round r (LM iteration 5r) for every
                                      feature point of the last frame: closest, ind2, ind3 (-1: none / round not run) */
  LL_BUF_COUNT_
} ll_buffer;

/* ---- lifetime ---------------------------------------------------------------- */

/* batch: sequences processed per call (>=1).  max_points: capacity of one input
 * scan (points).  cuda_stream: a cudaStream_t to enqueue on, or NULL to let the
 * handle create its own.  device: CUDA ordinal. */
int ll_create(const LegoLoamParams* params, int batch, int max_points, int device,
              void* cuda_stream, ll_handle** out);
int ll_destroy(ll_handle* h);
/* Forget all per-sequence history (first-frame state, transformCur/Sum, last clouds). */
int ll_reset(ll_handle* h);
/* Only the FeatureAssociation side, as if that object had just been constructed (featureAssociation.cpp:96-157): the
 * next scan is a first frame again.  MapOptimization state (poses, key frames, local map) is kept.  Used to store
 * key frames from scans that are not consecutive in time (the 500-key-frame map). */
int ll_reset_feature_association(ll_handle* h);
const char* ll_last_error(const ll_handle* h);
/* Number of kernels this handle has launched since creation (bench "gpu_launches"). */
int64_t ll_kernel_launches(const ll_handle* h);

/* ---- input (replaces sensor_msgs::PointCloud2 of imageProjection.cpp:153-161) -- */

/* Host scans: xyzi[batch][stride_points][4] floats, n_points[batch] valid points each
 * (NaN points must already be removed, as pcl::removeNaNFromPointCloud does).
 * Copies host->device on the handle's own copy stream into one of two input buffers (async if the
 * memory is pinned): the next ll_image_projection consumes it, and calling ll_set_scans_host for scan
 * f+1 before reading back the results of scan f overlaps that copy with the kernels of scan f. */
int ll_set_scans_host(ll_handle* h, const float* xyzi, const int32_t* n_points, int stride_points);
/* Same with packed 12-byte points, xyz[batch][stride_points][3] floats: three quarters of the host->device bytes.
 * The path never reads the intensity a sensor reports -- projectPointCloud overwrites it with row + col / 10000
 * (imageProjection.cpp:216) before anything uses it -- so every output is identical to ll_set_scans_host's.
 * LL_BUF_INPUT_CLOUD then returns 3 floats per point. */
int ll_set_scans_xyz_host(ll_handle* h, const float* xyz, const int32_t* n_points, int stride_points);
/* sensor_msgs/PointCloud2 ingest: pcl::fromROSMsg + pcl::removeNaNFromPointCloud of
 * imageProjection.cpp:159-161 on the device.  data: the messages' `data` arrays, [batch][stride_bytes] host bytes;
 * n_points[batch] = width * height; point_step and the byte offsets of the FLOAT32 fields x, y, z, intensity as listed
 * in msg.fields (off_intensity < 0: the message has no such field, intensity = 0).  is_dense is the message flag:
 * when set PCL copies the cloud unchecked, otherwise every point with a non-finite x, y or z is dropped, order kept.
 * Staged like ll_set_scans_host (copy stream, double buffered); the next ll_image_projection consumes it. */
int ll_set_scans_pointcloud2_host(ll_handle* h, const uint8_t* data, const int32_t* n_points, size_t stride_bytes, int point_step,
                                  int off_x, int off_y, int off_z, int off_intensity, int is_dense);
/* Same, for scans that already live in device memory (no copy is made; the buffer
 * must stay valid until the next ll_image_projection has been enqueued). */
int ll_set_scans_device(ll_handle* h, const float* xyzi_dev, const int32_t* n_points, int stride_points);

/* ---- the hot path ------------------------------------------------------------- */

/* resetParameters + findStartEndAngle + projectPointCloud + groundRemoval +
 * cloudSegmentation/labelComponents (imageProjection.cpp:107-150,178-308,352-496). */
int ll_image_projection(ll_handle* h);
/* adjustDistortion, calculateSmoothness, markOccludedPoints, extractFeatures, then
 * (first frame) checkSystemInitialization, (later frames) updateTransformation,
 * integrateTransformation, publishCloudsLast (featureAssociation.cpp:161-383,
 * 503-1032,1181-1270,1329-1359). */
int ll_feature_association(ll_handle* h);

/* Local map of one sequence for scan-to-map, already down-sampled
 * (laserCloudCornerFromMapDS / laserCloudSurfFromMapDS, mapOptmization.cpp:989-995).
 * Host pointers; copied to the device. */
int ll_map_set_local(ll_handle* h, int seq, const float* corner_xyzi, int n_corner,
                     const float* surf_xyzi, int n_surf);
/* Down-sampled current scan of one sequence (laserCloudCornerLastDS and
 * laserCloudSurfTotalLastDS, mapOptmization.cpp:999-1026), host pointers. */
int ll_map_set_scan(ll_handle* h, int seq, const float* corner_xyzi, int n_corner,
                    const float* surf_total_xyzi, int n_surf_total);
/* On-device downsampleCurrentScan (mapOptmization.cpp:999-1026): VoxelGrid 0.2 m of
 * corner_last, 0.4 m of surf_last and outlier_last, concat, 0.4 m again -- for all
 * sequences, from the clouds ll_feature_association left on the device. */
int ll_map_downsample_current_scan(ll_handle* h);
/* transformTobeMapped initial guess, f32[batch][6] host (result of
 * transformAssociateToMap, mapOptmization.cpp:264-387, computed by the host class). */
int ll_map_set_initial_guess(ll_handle* h, const float* transform_tobe_mapped);
/* Same as ll_map_set_initial_guess but only enqueues the copy (the host buffer must stay valid and
 * unchanged until the stream has consumed it; use pinned memory to keep it asynchronous). */
int ll_map_set_initial_guess_async(ll_handle* h, const float* transform_tobe_mapped);
/* Odometry -> map pose chain kept on the device (MapOptimization::transformAssociateToMap and
 * transformUpdate, mapOptmization.cpp:264-395).  ll_map_set_poses seeds transformAftMapped /
 * transformBefMapped (f32[batch][6] host); ll_map_predict_pose computes transformTobeMapped from them and
 * the current odometry pose; ll_scan_to_map ends with transformUpdate. */
int ll_map_set_poses(ll_handle* h, const float* transform_aft_mapped, const float* transform_bef_mapped);
int ll_map_predict_pose(ll_handle* h);
/* MapOptimization's copy of the odometry pose (transformSum of mapOptmization.cpp:1539, filled there from
 * AssociationOut::laser_odometry), f32[batch][6] host.  ll_map_downsample_current_scan -- the hand-over of a scan to
 * MapOptimization -- already takes this copy on the device, so a FeatureAssociation thread may integrate further scans
 * before ll_map_predict_pose / ll_scan_to_map run; call this only to override it (a caller that carries the payload
 * through its own channel). */
int ll_map_set_odometry(ll_handle* h, const float* transform_sum);
/* kd-tree replacement build + <=10 x (cornerOptimization, surfOptimization,
 * LMOptimization) (mapOptmization.cpp:1028-1332) for all sequences. */
int ll_scan_to_map(ll_handle* h);

/* ---- key frames and the local map on the device ----
 * MapOptimization::saveKeyFramesAndFactor (mapOptmization.cpp:1335-1474, iSAM2 == identity because loop closure is
 * off), extractSurroundingKeyFrames (:856-996, the enable_loop_closure == false branch) and transformPointCloud
 * (:443-473) for all sequences, without host round trips.
 *
 * ll_map_enable_keyframes allocates the per-sequence stores: max_keyframes key poses (<= 32768), pool_points points
 * for all key-frame clouds of one sequence (corner + surf + outlier, already down-sampled), and local maps of up to
 * max_map_corner / max_map_surf points (laserCloudCornerFromMapDS / laserCloudSurfFromMapDS).  A sequence that
 * outgrows a capacity stops saving key frames / truncates its map and raises a bit in LL_BUF_KEYFRAME_STATE[3]; the
 * next ll_mapping_cycle then returns LL_ERR_CAPACITY (the bits are read back one cycle late, no call waits for the device).
 * Fails with LL_ERR_STATE when params.enable_loop_closure is set (the loop-closure branch is not built). */
int ll_map_enable_keyframes(ll_handle* h, int max_keyframes, int pool_points, int max_map_corner, int max_map_surf);
/* extractSurroundingKeyFrames: radius search over the key poses around currentRobotPosPoint, 1 m VoxelGrid of those
 * poses, update of surroundingExistingKeyPosesID, concatenation of the surrounding key-frame clouds and the 0.2 m /
 * 0.4 m VoxelGrid that gives the local map used by ll_scan_to_map. */
int ll_map_extract_surrounding_keyframes(ll_handle* h);
/* saveKeyFramesAndFactor: the 0.3 m rule, key pose, copies of laserCloudCornerLastDS / SurfLastDS / OutlierLastDS
 * (left on the device by ll_map_downsample_current_scan), stored transformed by the key pose. */
int ll_map_save_keyframe(ll_handle* h);
/* One body of MapOptimization::run (:1526-1562): ll_map_downsample_current_scan (the hand-over), ll_map_predict_pose,
 * ll_map_extract_surrounding_keyframes, ll_scan_to_map, ll_map_save_keyframe.  Returns LL_ERR_CAPACITY when a
 * sequence outgrew a capacity of the key-frame store in the previous cycle. */
int ll_mapping_cycle(ll_handle* h);
/* One stored key-frame cloud (which: 0 corner, 1 surf, 2 outlier) of one sequence, transformed by its key pose, in
 * the point order of the down-sampled scan cloud it was copied from.  Synchronises the stream. */
int ll_map_download_keyframe(ll_handle* h, int seq, int keyframe, int which, void* dst, size_t dst_bytes, size_t* n_elems);

/* One full frame for every sequence: ll_image_projection + ll_feature_association,
 * and every mapping_frequency_divider-th odometry frame also ll_map_downsample_current_scan +
 * ll_map_predict_pose + ll_scan_to_map when a local map is set -- or, after ll_map_enable_keyframes, ll_mapping_cycle.
 * Returns 1 on such frames, else 0. */
int ll_process_scans(ll_handle* h);

/* ---- results ------------------------------------------------------------------- */

/* Poses of all sequences, f32[batch][6] each, host pointers (any may be NULL).
 * Synchronises the stream. */
int ll_get_poses(ll_handle* h, float* transform_sum, float* transform_cur, float* transform_tobe_mapped);
/* The same copies enqueued behind the work already submitted, without waiting: the caller may submit the next scans
 * first and collect these poses later with ll_wait_poses (a pipelined consumer, like the reference's stage threads
 * behind their Channels, main.cpp:37-38).  The destination buffers must stay valid until ll_wait_poses returns; use
 * page-locked memory for the copies to be truly asynchronous.  At most two read-backs may be in flight;
 * ll_wait_poses waits for the oldest one. */
int ll_get_poses_async(ll_handle* h, float* transform_sum, float* transform_cur, float* transform_tobe_mapped);
int ll_wait_poses(ll_handle* h);
/* nav_msgs/Odometry form of a pose.  pose7 = position x, y, z, orientation x, y, z, w as
 * FeatureAssociation::publishOdometry (featureAssociation.cpp:1286-1298) and MapOptimization::publishTF
 * (mapOptmization.cpp:510-522) fill them: tf::createQuaternionMsgFromRollPitchYaw(t[2], -t[0], -t[1]) in double, then
 * (x, y, z, w) = (-q.y, -q.z, q.x, q.w).  ll_odometry_to_transform is the consumer's side, OdometryToTransform of
 * utility.h:96-110 (tf::Matrix3x3(tf::Quaternion(z, -x, -y, w)).getRPY).  Pure host functions, no handle needed.
 * tf is not vendored in the reference tree: setRPY / getRPY are restated from its published LinearMath sources. */
void ll_transform_to_odometry(const float* transform6, double* pose7);
void ll_odometry_to_transform(const double* pose7, float* transform6);
/* The two odometry messages of the path for every sequence (either pointer may be NULL): laser_odometry f64[batch][7]
 * from transformSum (/laser_odom_to_init), odom_aft_mapped f64[batch][13] = pose7 of transformAftMapped followed by
 * twist.angular = transformBefMapped[0..2], twist.linear = transformBefMapped[3..5] (mapOptmization.cpp:524-529).
 * Synchronises the stream. */
int ll_get_odometry(ll_handle* h, double* laser_odometry, double* odom_aft_mapped);
/* Copy one array of one sequence to host.  *n_elems receives the element count
 * (points for "pt" buffers).  dst may be NULL to query the count only.
 * Synchronises the stream. */
int ll_download(ll_handle* h, int seq, int buffer, void* dst, size_t dst_bytes, size_t* n_elems);
/* Overwrite persistent state of one sequence (teacher-forced parity runs):
 * LL_BUF_TRANSFORM_CUR, LL_BUF_TRANSFORM_SUM, LL_BUF_TRANSFORM_TOBE_MAPPED, LL_BUF_TRANSFORM_BEF_MAPPED,
 * LL_BUF_TRANSFORM_AFT_MAPPED. */
int ll_upload(ll_handle* h, int seq, int buffer, const void* src, size_t n_elems);
int ll_synchronize(ll_handle* h);
/* MapOptimization's work is enqueued on a stream of its own (the reference runs it on a thread of its own,
 * mapOptmization.cpp:122) and overlaps the following scans.  ll_join_mapping makes the handle's frame stream wait, on the
 * device, for every mapping cycle enqueued so far -- no host synchronisation; an event recorded on the frame stream
 * afterwards marks the completion of both.  ll_synchronize waits on the host for both streams. */
int ll_join_mapping(ll_handle* h);

/* Index-level parity aid: when enabled, ll_scan_to_map records LL_BUF_MAP_KNN_IDX and ll_feature_association records
 * LL_BUF_ODOM_SEARCH_IDX (extra kernels / stores; off by default). */
int ll_enable_index_trace(ll_handle* h, int enable);

/* Per-stage device time of the last ll_process_scans call, in milliseconds:
 * [0] projection+ground, [1] segmentation, [2] feature extraction,
 * [3] scan-to-scan LM + glue, [4] scan-to-map (0 if not run).  Needs
 * ll_enable_stage_timing(h, 1) before the call; synchronises the stream. */
int ll_enable_stage_timing(ll_handle* h, int enable);
int ll_get_stage_times_ms(ll_handle* h, float* ms5);

/* Per-kernel device time: select one kernel by name (e.g. "k_extract_features"; NULL = none) and every
 * later launch of it is bracketed by a CUDA event pair on the handle's stream (up to 4096 launches).
 * ll_get_kernel_time returns the summed duration and the number of launches since ll_time_kernel. */
int ll_time_kernel(ll_handle* h, const char* kernel_name);
int ll_get_kernel_time(ll_handle* h, double* total_ms, int* launches);
/* With ll_time_kernel(h, "*") every kernel is timed; this writes one text line per kernel name,
 * "name total_ms launches\n", into buf (NUL-terminated). */
int ll_get_kernel_time_table(ll_handle* h, char* buf, size_t cap);

/* ---- rosbag ingest (replaces rosbag::Bag / rosbag::View of main.cpp:26-35,60-76) ---------- */

/* ROS-less reader of rosbag v2.0 files: the sensor_msgs/PointCloud2 messages of one topic in record-time order, as
 * rosbag::View hands them to ImageProjection::cloudHandler.  Uncompressed, lz4 and bz2 chunks (anything else, or a
 * damaged chunk: LL_ERR_INVALID_ARG, see ll_bag_last_error).  Pure host code; the bag format and the message layout are restated from their published
 * specifications (ROS is not vendored in the reference tree: parity unpinned). */
typedef struct ll_bag ll_bag;
typedef struct {
  uint64_t bag_time_ns;            /* record time (what rosbag::View sorts by) */
  uint32_t stamp_sec, stamp_nsec;  /* header.stamp */
  uint32_t height, width, point_step, row_step;
  int32_t is_bigendian, is_dense;
  int32_t off_x, off_y, off_z, off_intensity; /* byte offsets of the FLOAT32 fields of that name, -1 if absent */
  const uint8_t* data;             /* points into the bag's buffer: valid until ll_bag_close */
  uint64_t data_len;
} ll_pointcloud2_view;
/* topic NULL or "": the first sensor_msgs/PointCloud2 topic found in the bag (ll_bag_topic tells which). */
int ll_bag_open(const char* path, const char* topic, ll_bag** out);
int ll_bag_num_messages(const ll_bag* bag);
const char* ll_bag_topic(const ll_bag* bag);
/* Message `index` (0 .. ll_bag_num_messages-1) ready for ll_set_scans_pointcloud2_host:
 * n_points = width * height, stride_bytes = data_len, point_step and the offsets as given. */
int ll_bag_get_pointcloud2(const ll_bag* bag, int index, ll_pointcloud2_view* view);
void ll_bag_close(ll_bag* bag);
const char* ll_bag_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* LEGO_LOAM_B200_H */
