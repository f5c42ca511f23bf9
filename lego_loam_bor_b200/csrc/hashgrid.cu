// hashgrid.cu -- build of the voxel hash grid (see hashgrid.cuh): clear, count, scan, fill.
// Replaces KdTreeFLANN::setInputCloud / buildIndex (nanoflann_pcl.h:131-138), which the reference
// calls for both last-frame clouds every frame (featureAssociation.cpp:1356-1359) and for both local
// maps every mapping cycle (mapOptmization.cpp:1317-1318).
#include "hashgrid.cuh"
#include "ll_kernels.h"

namespace {

struct BuildArgs {
  HashGrid g;
  const float4* pts;
  int stride;
  const int* counts;
  int count_stride, count_off;
  const int* enable;
  int enable_stride;
};

__device__ __forceinline__ bool seq_enabled(const BuildArgs& a, int s) { return a.enable == nullptr || a.enable[s * a.enable_stride] != 0; }

__global__ void __launch_bounds__(256) k_grid_clear(BuildArgs a) {
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.g.tbl) a.g.cursor[(size_t)s * a.g.tbl + i] = 0;
}

__global__ void __launch_bounds__(256) k_grid_count(BuildArgs a) {
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const int n = min(a.counts[s * a.count_stride + a.count_off], a.g.cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = a.pts[(size_t)s * a.stride + i];
  const uint32_t h = grid_hash(grid_cell(q.x, a.g.inv_cell), grid_cell(q.y, a.g.inv_cell), grid_cell(q.z, a.g.inv_cell), a.g.tbl);
  atomicAdd(a.g.cursor + (size_t)s * a.g.tbl + h, 1);
}

__global__ void __launch_bounds__(1024) k_grid_scan(BuildArgs a) {
  __shared__ int warp_tot[33];
  const int s = blockIdx.x;
  if (!seq_enabled(a, s)) return;
  int* cur = a.g.cursor + (size_t)s * a.g.tbl;
  int* cs = a.g.cell_start + (size_t)s * (a.g.tbl + 1);
  const int ipt = (a.g.tbl + blockDim.x - 1) / blockDim.x;
  const int i0 = threadIdx.x * ipt, i1 = min(a.g.tbl, i0 + ipt);
  int sum = 0;
  for (int i = i0; i < i1; ++i) sum += cur[i];
  int total;
  int run = block_exclusive_scan(sum, warp_tot, &total);
  for (int i = i0; i < i1; ++i) {
    const int c = cur[i];
    cs[i] = run;
    cur[i] = run;
    run += c;
  }
  if (threadIdx.x == 0) {
    cs[a.g.tbl] = total;
    a.g.count[s] = total;
  }
}

__global__ void __launch_bounds__(256) k_grid_fill(BuildArgs a) {
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const int n = min(a.counts[s * a.count_stride + a.count_off], a.g.cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = a.pts[(size_t)s * a.stride + i];
  const uint32_t h = grid_hash(grid_cell(q.x, a.g.inv_cell), grid_cell(q.y, a.g.inv_cell), grid_cell(q.z, a.g.inv_cell), a.g.tbl);
  const int pos = atomicAdd(a.g.cursor + (size_t)s * a.g.tbl + h, 1);
  a.g.sorted[(size_t)s * a.g.cap + pos] = make_float4(q.x, q.y, q.z, __int_as_float(i));
}

}  // namespace

void launch_grid_build(LaunchCtx& ctx, HashGrid& g, int B, const float4* pts, int stride, const int* counts,
                       int count_stride, int count_off, const int* enable, int enable_stride) {
  BuildArgs a{g, pts, stride, counts, count_stride, count_off, enable, enable_stride};
  LL_LAUNCH(ctx, "k_grid_clear", k_grid_clear<<<dim3((g.tbl + 255) / 256, B), 256, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_count", k_grid_count<<<dim3((g.cap + 255) / 256, B), 256, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_scan", k_grid_scan<<<B, 1024, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_fill", k_grid_fill<<<dim3((g.cap + 255) / 256, B), 256, 0, ctx.stream>>>(a));
}
