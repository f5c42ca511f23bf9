// ll_device.cuh -- device-side state shared by all kernels of the hot path.
//
// Data layout in HBM (DESIGN.md section "Layout"): everything is batched over B independent
// sequences; per-sequence arrays are contiguous slabs of N = V*H elements (range image order,
// row-major: cell = row*H + col) or of a fixed capacity, so one launch covers all sequences
// with blockIdx.y (or a flattened index) selecting the sequence.  Points are 16-byte float4
// (x, y, z, intensity) -- the packed form of pcl::PointXYZI (utility.h:46).
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/lego_loam_b200.h"
#include "../../include/ll_portable_math.h"

#define LL_MAX_RINGS 128
// scan-to-map k-NN: map points recorded per query by a full search (mapping.cu, k_map_knn); <= 16 (4-bit positions)
#ifndef LL_KNN_K
#define LL_KNN_K 8
#endif
#define LL_INVALID_LABEL 999999

// hash grid over a point cloud (replaces nanoflann::KdTreeFLANN, nanoflann_pcl.h:54-152)
struct HashGrid {
  int* cell_start;   // [B][tbl+4] exclusive prefix of bucket counts (tbl + 1 entries; rows padded to 16 bytes: grid_cs_stride)
  int* cnt;          // [B][tbl]   bucket counters (zero at rest: the fill counts them back down)
  unsigned* occ;     // [B][tbl/32] bit h set <=> bucket h is not empty (tested before touching cell_start)
  unsigned* sig;     // [B][tbl] or null: bit ((ring id + 1) & 31) set for every ring id present in the bucket
  int* tile_tot;     // [B][ntiles] per-tile counter sums (accumulated by the count kernel, zero at rest)
  int ntiles;        // tbl / 4096
  float4* sorted;    // [B][cap]   points bucket by bucket; .w carries the original index (int bits)
  int* count;        // [B]        points indexed
  float inv_cell;    // 1 / cell size
  float cell;        // cell size (m)
  int tbl;           // buckets per sequence (power of two)
  int cap;           // point capacity per sequence
};
__host__ __device__ __forceinline__ size_t grid_cs_stride(const HashGrid& g) { return (size_t)g.tbl + 4; }

// Sparse voxel accumulator of one local map (keyframes.cu): open-addressing table keyed by the absolute voxel
// coordinates, split into `parts` independent sub-tables (a voxel belongs to exactly one, chosen by a hash of its key).
// Every voxel also heads a chain of the key-frame runs summed into it, in concatenation order (KeyframeStore::pool_link),
// so that a key frame that leaves the surrounding set is taken out by re-summing only the voxels it touches.
struct VoxTable {
  unsigned long long* key;  // [B][parts][sub_cap] packed voxel (iz, iy, ix), all ones = empty
  float4* sum;              // [B][parts][sub_cap] running sums of x, y, z, intensity in concatenation order
  int* cnt;                 // [B][parts][sub_cap] points summed (0: a voxel all of whose key frames have left)
  int* head;                // [B][parts][sub_cap] pool index of the first run of the chain, -1 = none
  int* tail;                // [B][parts][sub_cap] pool index of the last run
  int* claim;               // [B][parts][sub_cap] stamp of the erase sweep that last re-summed the voxel
  unsigned* list;           // [B][parts][sub_cap] slots in use, in insertion order
  int* list_n;              // [B][parts]
  int* list_done;           // [B][parts] entries of `list` already merged into the sorted order (xs_*)
  int parts, sub_cap;       // sub_cap is a power of two
  float leaf;               // VoxelGrid leaf size (m)
};

// MapOptimization's key-frame state for every sequence (mapOptimization.h:96-135), SURVEY.md section 8 f2.
struct KeyframeStore {
  int enabled;
  int kf_cap;                    // key frames per sequence
  int pool_cap;                  // points per sequence over all key-frame clouds
  float radius2;                 // (float)(surrounding_keyframe_search_radius^2), squared in double (nanoflann_pcl.h:163)
  int* kf_count;                 // [B] cloudKeyPoses3D->size()
  float* kf_pose;                // [B][kf_cap][6] cloudKeyPoses6D: roll, pitch, yaw, x, y, z
  int* kf_off;                   // [B][kf_cap][4] pool offsets of the corner, surf, outlier cloud and the end
  int* pool_used;                // [B]
  float4* pool_pts;              // [B][pool_cap] key-frame clouds transformed by their key pose; each cloud stable-sorted by voxel
  unsigned long long* pool_key;  // [B][pool_cap] packed absolute voxel of every stored point (leaf of the map it goes into)
  int* pool_perm;                // [B][pool_cap] position of every stored point in its down-sampled scan cloud
  unsigned long long* pool_link; // [B][pool_cap] at the first point of a run that is summed into a voxel: (next run of that voxel's chain << 32) | run length
  int* kf_new;                   // [B] index of the key frame ll_map_save_keyframe is storing, or -1
  float* robot_pos;              // [B][8] currentRobotPosPoint xyz, pad, previousRobotPosPoint xyz, pad
  float* transform_last;         // [B][6]
  int* sur_ids;                  // [B][kf_cap] surroundingExistingKeyPosesID
  int* sur_n;                    // [B]
  int* sur_first;                // [B] first list position whose clouds still have to be summed into the voxel tables
  int* sur_rebuild;              // [B] 1: the tables are rebuilt from list position 0 (table pressure after erases)
  int* sur_valid;                // [B] 0: no key frames yet (extractSurroundingKeyFrames returns early, maps stay empty)
  int* sur_erased;               // [B][kf_cap] key frames the last extractSurroundingKeyFrames erased from the list
  int* sur_n_erased;             // [B]
  int* sur_last_erased;          // [B] 1 if the last extractSurroundingKeyFrames erased a key frame (LL_BUF_KEYFRAME_STATE[2])
  int* err;                      // [B] capacity error bits (LL_BUF_KEYFRAME_STATE[3])
  VoxTable tbl[2];               // 0: corner map (leaf 0.2), 1: surf map (leaf 0.4)
  // k_kf_select scratch
  unsigned *sel_k0, *sel_k1, *sel_v0, *sel_v1;  // [B][kf_cap] radix sort ping-pong
  int* sel_rank_idx;             // [B][kf_cap] key-frame index of the r-th nearest selected pose
  int* sel_ds_ids;               // [B][kf_cap] ids of the down-sampled surrounding poses, in voxel order
  int* sel_first_pos;            // [B][kf_cap] first position of an id in sel_ds_ids
  // sorted order of the voxels of a map (ascending PCL voxel index = lexicographic (iz, iy, ix)), kept across cycles
  unsigned long long* xs_key[2]; // ping-pong: [B][2 maps][sort_cap]
  unsigned* xs_slot[2];          // slot over the whole table (part * sub_cap + slot)
  int* xs_cur;                   // [B][2] which ping-pong buffer is current
  int* xs_n;                     // [B][2] voxels in the sorted order (alive or not)
  int* xs_merge;                 // [B][2][4] merge job of this cycle: source buffer, old count, new count, pad
  unsigned long long* xn_key;    // [B][2][sort_cap] new voxels of this cycle, sorted
  unsigned* xn_slot;             // [B][2][sort_cap]
  int* xs_tile;                  // [B][2][sort_cap / 4096 + 1] alive voxels per tile of the sorted order
  int* xs_bbox;                  // [B][2][8] running voxel bounding box of a map: min xyz, max xyz
  unsigned *sk0, *sk1, *sv0, *sv1;  // [B][2][sort_cap] radix sort scratch for the new voxels
  int sort_cap;
  int stamp;                     // host counter of extractSurroundingKeyFrames calls (erase sweeps claim voxels with it)
};

struct DevParams {
  int B, V, H, N, max_pts;
  // ImageProjection derived constants (imageProjection.cpp:64-84), computed on the host with
  // the same portable math the kernels use
  float ang_bottom, ang_res_x, ang_res_y, sensor_mount_angle;
  float seg_tan_theta, sin_ax, cos_ax, sin_ay, cos_ay;
  int gsi, seg_valid_point_num, seg_valid_line_num;
  // FeatureAssociation (featureAssociation.cpp:69-81)
  float scan_period, edge_threshold, surf_threshold, nearest_feature_dist_sqr;
  // capacities
  int cap_sharp, cap_less_sharp, cap_flat;  // 12V, 120V, 24V
};

struct DevState {
  DevParams p;
  // ---- input ----
  const float4* in_pts;  // [B][in_stride] (x, y, z, intensity), or packed (x, y, z) floats when in_xyz3 is set
  int in_stride;
  int in_xyz3;           // ll_set_scans_xyz_host: 12-byte points (the path never reads the input intensity)
  int* n_in;             // [B]
  uint32_t frame_tag;    // increases every frame; makes `winner` self-clearing
  // ---- ImageProjection ----
  unsigned long long* winner;  // [B][N] (frame_tag << 32) | input index of the last writer
  float* range_mat;            // [B][N]
  float4* full_cloud;          // [B][N]
  int8_t* ground_mat;          // [B][N]
  int* label_mat;              // [B][N]
  int* parent;                 // [B][N] union-find forest, -1 = not a segmentation candidate
  int* comp_size;              // [B][N] at roots: pixel count
  unsigned* comp_rows;         // [B][N] at roots: bit (row - root row) for every non-seed pixel
  int* tile_counts;            // [B][V][4] per row: feasible roots, kept pixels, outliers
  float* orientation;          // [B][4] start, end, diff
  int* half_idx;               // [B] first segmented index whose branch-1 orientation passes pi (adjustDistortion)
  // ProjectionOut / cloud_info
  float4* seg_cloud;           // [B][N]
  float* seg_range;            // [B][N]
  uint32_t* seg_col;           // [B][N]
  uint8_t* seg_ground;         // [B][N]
  uint8_t* seg_class;          // [B][N] per CELL: classify_cell flags of this frame (k_seg_count -> k_seg_emit)
  float* seg_ori;              // [B][N] -atan2(y, x) of every segmented point (adjustDistortion's raw orientation)
  int* start_ring;             // [B][V]
  int* end_ring;               // [B][V]
  int* seg_count;              // [B]
  float4* outlier_cloud;       // [B][N/5+V]
  int* outlier_count;          // [B]
  int cap_outlier;
  // ---- FeatureAssociation ----
  float* curvature;            // [B][N] cloudCurvature (persistent)
  int* picked;                 // [B][N] cloudNeighborPicked (persistent)
  int* cloud_label;            // [B][N] cloudLabel (persistent)
  unsigned* slot4;             // [B][2] the one stale cloudSmoothness entry that is ever read (position 4): value bits, ind
  unsigned* scan_list;         // [B][N] per ring: its 12 candidate lists (edge/flat per sextant) in visiting order
  int* sext_off;               // [B][V][16] offsets of those 12 lists inside the ring's range (+ end)
  // per-ring staging written by the extraction kernel
  int* st_sharp_ind;       // [B][V][12]
  int* st_less_sharp_ind;  // [B][V][120]
  int* st_flat_ind;        // [B][V][24]
  float4* st_less_flat;                       // [B][V][H]
  int* ring_counts;                           // [B][V][8]: sharp, lessSharp, flat, lessFlatDS, lessFlatRaw
  // compacted feature clouds of the current frame
  float4* corner_sharp; int* corner_sharp_ind;            // [B][12V]
  float4* corner_less_sharp; int* corner_less_sharp_ind;  // [B][120V]
  float4* surf_flat; int* surf_flat_ind;                  // [B][24V]
  float4* surf_less_flat;                                 // [B][N]
  int* feat_counts;                                       // [B][4]
  // last-frame clouds (double buffered by pointer swap like featureAssociation.cpp:1344-1350)
  float4* corner_last;  // [B][120V]
  float4* surf_last;    // [B][N]
  int* last_counts;     // [B][2]
  int* win_first;       // [2 frame parities][B][2 clouds][LL_MAX_RINGS+8] first index of a run of each ring id
  int* win_last;        // same shape: last index of a run of each ring id
  // correspondences of LM iteration 0 of the current stage (k_odom_search -> k_odom_stage), stride 24V
  float4* odom_ga;      // [B][24V] SURF: unit plane through the three correspondences; CORNER: first line point
  float4* odom_gb;      // [B][12V] CORNER: second line point
  float4* odom_s0;      // [B][24V] transformed feature point at search time (x, y, z), slack
  uint8_t* odom_ok;     // [B][24V]
  int* odom_cl;         // [B][24V] closest point (index into the last-frame cloud) or -1
  long long* stage_clocks;  // [B][16] see LL_BUF_STAGE_CLOCKS
  int stage_clocks_on;      // per-phase clocks of the LM stage kernels (ll_enable_stage_timing); the total is always written
  long long* ring_clocks;   // [B][V][10] see LL_BUF_RING_CLOCKS
  float4* outlier_last; // [B][cap_outlier]
  HashGrid grid_corner_last, grid_surf_last;  // index the clouds the "kd-trees" were last built on
  float* transform_cur;  // [B][6]
  float* transform_sum;  // [B][6]
  int* odom_iters;       // [B][2]
  int* odom_flags;       // [B][4]: 0 isDegenerate, 2 grids index the current last-frame clouds, 3 outlier count
  float* odom_matP;      // [B][9]
  // ---- MapOptimization scan-to-map ----
  float4* map_corner; float4* map_surf;   // [B][cap_map_corner], [B][cap_map_surf]
  int* map_counts;                        // [B][2]
  int cap_map_corner, cap_map_surf;
  HashGrid grid_map_corner, grid_map_surf;
  float4* scan_corner_ds; float4* scan_surf_ds;  // [B][120V], [B][N]
  int* scan_ds_counts;                           // [B][2]
  float* transform_tobe_mapped;                  // [B][6]
  float* map_odom;                               // [B][6] MapOptimization::transformSum: the odometry pose handed over with the scan (mapOptmization.cpp:1539)
  float* transform_bef_mapped;                   // [B][6]
  float* transform_aft_mapped;                   // [B][6]
  int* map_iters;                                // [B][2]
  int* map_flags;                                // [B][4]: 0 isDegenerate, 1 converged/done
  float* map_matP;                               // [B][36]
  // VoxelGrid scratch for downsampleCurrentScan (voxelgrid.cu)
  unsigned *vox_key0, *vox_key1, *vox_val0, *vox_val1;  // [B][3][vox_cap]
  float4* vox_tmp_surf;                          // [B][N]  laserCloudSurfLastDS
  float4* vox_tmp_out;                           // [B][cap_outlier] laserCloudOutlierLastDS
  int* vox_tmp_counts;                           // [B][2]
  int vox_cap;
  float4* map_knn_rec;                           // [B][map_knn_cap][11] per query: 10 nearest map points (x,y,z,index) + (p0, bound)
  int* map_knn_sel;                              // [B][map_knn_cap] positions of the current 5-NN inside the record (5 x 4 bits), bit 31 valid
  int map_knn_cap;
  float4* map_fit;                               // [B][map_knn_cap][2] cached neighbour-only part of the line / plane fit
  int* map_ticket;                               // [B] blocks of k_map_iter that have delivered their partial sums
  double* map_partials;                          // [B][max_blocks][28]
  double* map_trace;                             // [B][10][34] per-iteration normal equations + step (parity/debug)
  int* map_rows;                                 // [B][max_blocks]
  int map_max_blocks;
  // index-level parity aids (ll_enable_index_trace); null when disabled
  int* map_knn_trace;                            // [B][10][map_knn_cap][5]
  int* odom_trace;                               // [B][2 stages][5 rounds][24V][3]
  // ---- MapOptimization key frames / local map ----
  KeyframeStore kf;
};

// ---- small device helpers --------------------------------------------------------------

__device__ __forceinline__ float4 ld_pt(const float4* p) { return __ldg(p); }

__device__ __forceinline__ float warp_min_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// `warp_tot` is a 33-int shared scratch.  Returns the exclusive prefix; *total gets the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < nw ? warp_tot[lane] : 0;
    int winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) warp_tot[lane] = winc - w;
    if (lane == 31) warp_tot[32] = winc;
  }
  __syncthreads();
  const int res = warp_tot[wid] + inc - v;
  *total = warp_tot[32];
  __syncthreads();
  return res;
}

__device__ __forceinline__ uint32_t grid_hash(int ix, int iy, int iz, int tbl) {
  // multiply-add of the three cell coordinates followed by a full avalanche (lowbias32): neighbouring
  // cells land in unrelated buckets, so an empty neighbour cell rarely aliases an occupied bucket
  uint32_t h = (uint32_t)ix * 0x9E3779B1u + (uint32_t)iy * 0x85EBCA77u + (uint32_t)iz * 0xC2B2AE3Du;
  h ^= h >> 16;
  h *= 0x7feb352du;
  h ^= h >> 15;
  h *= 0x846ca68bu;
  h ^= h >> 16;
  return h & (uint32_t)(tbl - 1);
}
