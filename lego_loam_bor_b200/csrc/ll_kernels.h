// ll_kernels.h -- host-side launch entry points of the kernel files (internal to the library).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct DevState;
struct HashGrid;

struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  cudaError_t first_error = cudaSuccess;
  const char* first_error_kernel = nullptr;
  inline void count(const char* name) {
    ++launches;
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && first_error == cudaSuccess) {
      first_error = e;
      first_error_kernel = name;
    }
  }
};

// projection.cu: projectPointCloud + groundRemoval (+ resets, start/end angle)
void launch_projection(LaunchCtx& ctx, DevState& st);
// segmentation.cu: cloudSegmentation / labelComponents
void launch_segmentation(LaunchCtx& ctx, DevState& st);
// features.cu: adjustDistortion, calculateSmoothness, markOccludedPoints, extractFeatures
void launch_feature_extraction(LaunchCtx& ctx, DevState& st);
// odometry.cu: first-frame initialisation / updateTransformation + integrateTransformation + publishCloudsLast
void launch_odometry(LaunchCtx& ctx, DevState& st, bool first_frame);
// hashgrid.cu: build the k-NN structure over `pts` ([B][stride] points, counts[B*count_stride + count_off])
void launch_grid_build(LaunchCtx& ctx, HashGrid& g, int B, const float4* pts, int stride, const int* counts,
                       int count_stride, int count_off, const int* enable /* [B] or null */, int enable_stride);
// mapping.cu: scan2MapOptimization and downsampleCurrentScan
void launch_scan_to_map(LaunchCtx& ctx, DevState& st);
void launch_downsample_current_scan(LaunchCtx& ctx, DevState& st);
