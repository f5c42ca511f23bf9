// ll_kernels.h -- host-side launch entry points of the kernel files (internal to the library).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct DevState;
struct HashGrid;

// Every kernel launch of the library goes through LL_LAUNCH: it counts launches (the bench's
// "gpu_launches"), records the first launch error, and -- for ONE selected kernel name -- brackets
// each launch with a CUDA event pair on the launching stream so that bench.py can report that
// kernel's average duration inside the timed region (roofline.achieved).
struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  cudaError_t first_error = cudaSuccess;
  const char* first_error_kernel = nullptr;
  // per-kernel timing
  static const int kMaxTimed = 4096;
  char timed_name[64] = {0};
  cudaEvent_t* ev_start = nullptr;
  cudaEvent_t* ev_stop = nullptr;
  const char** ev_name = nullptr;  // kernel name of every recorded launch
  int timed_used = 0;
  bool timing_now = false;
  // set by the hand-over to MapOptimization (api.cu): the next k_publish_clouds_last overwrites the last-frame clouds that
  // downsampleCurrentScan is still reading on the mapping stream and has to wait for this event first
  cudaEvent_t wait_before_publish = nullptr;
  inline bool is_timed(const char* name) const {
    if (!timed_name[0] || !ev_start) return false;
    if (timed_name[0] == '*' && timed_name[1] == 0) return true;  // "*": every kernel
    int i = 0;
    while (timed_name[i] && name[i] && timed_name[i] == name[i]) ++i;
    return timed_name[i] == 0 && name[i] == 0;
  }
  inline void pre(const char* name) {
    timing_now = is_timed(name) && timed_used < kMaxTimed;
    if (timing_now) cudaEventRecord(ev_start[timed_used], stream);
  }
  inline void count(const char* name) {
    ++launches;
    if (timing_now) {
      cudaEventRecord(ev_stop[timed_used], stream);
      ev_name[timed_used] = name;
      ++timed_used;
      timing_now = false;
    }
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && first_error == cudaSuccess) {
      first_error = e;
      first_error_kernel = name;
    }
  }
};

#define LL_LAUNCH(ctx, name, ...) \
  do {                            \
    (ctx).pre(name);              \
    __VA_ARGS__;                  \
    (ctx).count(name);            \
  } while (0)

// projection.cu: projectPointCloud + groundRemoval (+ resets, start/end angle)
void launch_projection(LaunchCtx& ctx, DevState& st);
// segmentation.cu: cloudSegmentation / labelComponents
void launch_segmentation(LaunchCtx& ctx, DevState& st);
void launch_label_final(LaunchCtx& ctx, DevState& st);  // numeric labels of the non-root cells, on demand (LL_BUF_LABEL_MAT)
// features.cu: adjustDistortion, calculateSmoothness, markOccludedPoints, extractFeatures
void launch_feature_extraction(LaunchCtx& ctx, DevState& st);
// odometry.cu: first-frame initialisation / updateTransformation + integrateTransformation + publishCloudsLast
void launch_odometry(LaunchCtx& ctx, DevState& st, bool first_frame);
// hashgrid.cu: build the k-NN structures over two clouds at once ([B][stride] points, counts[s*cstride+coff])
void launch_grid_build2(LaunchCtx& ctx, int B, HashGrid& g0, const float4* pts0, int stride0, const int* counts0,
                        int cstride0, int coff0, HashGrid& g1, const float4* pts1, int stride1, const int* counts1,
                        int cstride1, int coff1, const int* enable /* [B*enable_stride] or null */, int enable_stride,
                        bool pack_ring = false /* sorted.w = index | (int(intensity) + 1) << 24 */);
// mapping.cu: scan2MapOptimization and downsampleCurrentScan
void launch_scan_to_map(LaunchCtx& ctx, DevState& st);
void launch_map_predict_pose(LaunchCtx& ctx, DevState& st);
void launch_downsample_current_scan(LaunchCtx& ctx, DevState& st);
// keyframes.cu: saveKeyFramesAndFactor / extractSurroundingKeyFrames on the device
void launch_extract_surrounding_keyframes(LaunchCtx& ctx, DevState& st);
void launch_save_keyframe(LaunchCtx& ctx, DevState& st);
// ingest.cu: sensor_msgs/PointCloud2 decode + removeNaNFromPointCloud (imageProjection.cpp:159-161)
struct Pc2Args {
  const uint8_t* raw;   // [B][raw_stride] message data
  size_t raw_stride;
  const int* n_raw;     // [B] width * height
  int point_step, off_x, off_y, off_z, off_intensity, is_dense;
  float4* out;          // [B][out_stride]
  int out_stride;
  int* n_out;           // [B]
  int* tile_cnt;        // [B][ntiles]
  int ntiles;
};
void launch_decode_pointcloud2(LaunchCtx& ctx, cudaStream_t stream, int B, const Pc2Args& a);
int pc2_tile_points();
