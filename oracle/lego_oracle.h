/*
 * lego_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * C interface of the CPU oracle: a from-scratch restatement of the reference's hot path
 * (ImageProjection, FeatureAssociation, MapOptimization::scan2MapOptimization) used only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under lego_loam_bor_b200/ may include, link or call this.
 *
 * It mirrors include/lego_loam_b200.h (same LegoLoamParams, same ll_buffer ids) for ONE
 * sequence, so a parity test is "run both, download the same buffer id, compare".
 */
#ifndef LEGO_ORACLE_H
#define LEGO_ORACLE_H

#include "../include/lego_loam_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lo_handle lo_handle;

/* Global backend switches (affect objects created afterwards and all math calls). */
void lo_set_math_backend(int use_libm);      /* 0 = portable (default), 1 = glibc float libm */
void lo_set_knn_backend(int use_nanoflann);  /* 1 = reference nanoflann (if compiled in), 0 = port kd-tree */
int lo_has_nanoflann(void);
/* J^T J / J^T r accumulation of the three LM solves of objects created afterwards: 0 = exact double products summed in
 * double (default, what the CUDA kernels do), 1 = float row by row, 2 = float with eight interleaved partial sums; 1 and 2
 * bracket what the reference's Eigen float GEMM does (featureAssociation.cpp:860-866, mapOptmization.cpp:1258). */
void lo_set_accum_backend(int float_accum);
/* extractFeatures' per-sextant sort of objects created afterwards: 0 = std::sort as in the reference (featureAssociation.cpp:285;
 * the order of EQUAL curvatures is whatever libstdc++'s introsort leaves), 1 = std::stable_sort (equal curvatures in
 * position order: one of the outcomes the C++ standard permits, and the one a parallel (value, position) sort gives). */
void lo_set_sort_backend(int stable);

lo_handle* lo_create(const LegoLoamParams* p);
void lo_destroy(lo_handle* h);
void lo_reset(lo_handle* h);
/* a freshly constructed FeatureAssociation (first-frame state, transformCur/Sum, last clouds); MapOptimization state is kept */
void lo_reset_feature_association(lo_handle* h);

int lo_image_projection(lo_handle* h, const float* xyzi, int n_points);
/* returns 1 when this frame would be handed to MapOptimization (featureAssociation.cpp:1432), else 0 */
int lo_feature_association(lo_handle* h);
int lo_map_set_local(lo_handle* h, const float* corner, int nc, const float* surf, int ns);
int lo_map_set_scan(lo_handle* h, const float* corner, int nc, const float* surf_total, int ns);
int lo_map_downsample_current_scan(lo_handle* h);
int lo_map_set_initial_guess(lo_handle* h, const float* t6);
/* transformAftMapped / transformBefMapped (mapOptimization.h), then transformAssociateToMap */
int lo_map_set_poses(lo_handle* h, const float* aft6, const float* bef6);
int lo_map_predict_pose(lo_handle* h);
int lo_scan_to_map(lo_handle* h);
/* Key frames and the local map, loop closure off (mapOptmization.cpp:856-996, 1335-1474, 1526-1562). */
int lo_map_extract_surrounding_keyframes(lo_handle* h);
int lo_map_save_keyframe(lo_handle* h);
/* transformAssociateToMap, extractSurroundingKeyFrames, downsampleCurrentScan, scan2MapOptimization, saveKeyFramesAndFactor */
int lo_mapping_cycle(lo_handle* h);
int lo_map_download_keyframe(lo_handle* h, int kf, int which, void* dst, size_t dst_bytes, size_t* n_elems);
double lo_get_timer_map_assembly(lo_handle* h);
int lo_download(lo_handle* h, int buffer, void* dst, size_t dst_bytes, size_t* n_elems);
int lo_upload(lo_handle* h, int buffer, const void* src, size_t n_elems);

/* sensor_msgs/PointCloud2 -> xyzi with NaN removal (imageProjection.cpp:159-161); returns the number of points kept. */
int lo_decode_pointcloud2(const unsigned char* data, int n_points, int point_step, int off_x, int off_y, int off_z,
                          int off_intensity, int is_dense, float* out_xyzi);
/* pcl::VoxelGrid restatement on its own: out must hold n points; returns output count. */
int lo_voxel_grid(const float* xyzi, int n, float leaf, float* out_xyzi);
/* k-NN on its own (for kd-tree pin tests): idx[nq*k], d2[nq*k]. */
int lo_knn(const float* cloud_xyzi, int n, const float* query_xyzi, int nq, int k, int* idx, float* d2);

/* The reference's own threading (main.cpp:37-47): ImageProjection on the caller's thread, FeatureAssociation and
 * MapOptimization on a thread each, one-slot blocking channels in between.  Three stage objects of one sequence; see
 * lego_oracle.cpp for the arguments. */
int lo_run_pipeline(lo_handle* h_ip, lo_handle* h_fa, lo_handle* h_mo, const float* const* scans, const int* counts, int n_frames,
                    int first_timed_frame, double* stage_ms, double* wall_s);

/* Wall-clock seconds spent inside each stage since the last lo_reset_timers:
 * [0] image projection, [1] feature extraction (adjust..extract), [2] scan-to-scan LM + glue,
 * [3] scan-to-map incl. tree builds, [4] downsampleCurrentScan. */
void lo_get_timers(lo_handle* h, double* sec5);
void lo_reset_timers(lo_handle* h);

#ifdef __cplusplus
}
#endif
#endif
