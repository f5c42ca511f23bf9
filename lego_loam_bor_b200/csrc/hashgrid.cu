// hashgrid.cu -- build of the voxel hash grid (see hashgrid.cuh): count, tile sums, scan, fill.
// Replaces KdTreeFLANN::setInputCloud / buildIndex (nanoflann_pcl.h:131-138), which the reference
// calls for both last-frame clouds every frame (featureAssociation.cpp:1356-1359) and for both local
// maps every mapping cycle (mapOptmization.cpp:1317-1318).  Two grids (corner + surf) are built by the
// same four launches (blockIdx.z selects the grid).  The bucket counters are zero at rest: the scan
// kernel clears them after reading, so no separate clear pass is needed.
#include "hashgrid.cuh"
#include "ll_kernels.h"

namespace {

#define GS_THREADS 1024
#define GS_TILE (GS_THREADS * 4)

struct BuildJob {
  HashGrid g;
  const float4* pts;
  int stride;
  const int* counts;
  int count_stride, count_off;
};

struct BuildArgs {
  BuildJob job[2];
  const int* enable;
  int enable_stride;
  int pack_ring;  // sorted.w = index | (ring id + 1) << 24 instead of the plain index (scan-to-scan clouds)
};

__device__ __forceinline__ bool seq_enabled(const BuildArgs& a, int s) { return a.enable == nullptr || a.enable[s * a.enable_stride] != 0; }

__global__ void __launch_bounds__(256) k_grid_count(BuildArgs a) {
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const BuildJob& j = a.job[blockIdx.z];
  const int n = min(j.counts[s * j.count_stride + j.count_off], j.g.cap);
  // grid-stride: the launch is sized for a typical cloud, not for the capacity (most capacity-sized blocks would be empty)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = j.pts[(size_t)s * j.stride + i];
    const uint32_t h = grid_hash(grid_cell(q.x, j.g.inv_cell), grid_cell(q.y, j.g.inv_cell), grid_cell(q.z, j.g.inv_cell), j.g.tbl);
    atomicAdd(j.g.cnt + (size_t)s * j.g.tbl + h, 1);
  }
}

// per-tile totals of the bucket counters
__global__ void __launch_bounds__(GS_THREADS) k_grid_tile_sums(BuildArgs a) {
  __shared__ int sh[GS_THREADS / 32];
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const BuildJob& j = a.job[blockIdx.z];
  const int tile = blockIdx.x;
  if (tile * GS_TILE >= j.g.tbl) return;
  const int4 v = reinterpret_cast<const int4*>(j.g.cnt + (size_t)s * j.g.tbl + (size_t)tile * GS_TILE)[threadIdx.x];
  int sum = warp_sum_i(v.x + v.y + v.z + v.w);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = threadIdx.x < GS_THREADS / 32 ? sh[threadIdx.x] : 0;
    t = warp_sum_i(t);
    if (threadIdx.x == 0) j.g.tile_tot[(size_t)s * j.g.ntiles + tile] = t;
  }
}

// exclusive scan: cell_start / cursor from the counters; counters are cleared for the next build
__global__ void __launch_bounds__(GS_THREADS) k_grid_scan(BuildArgs a) {
  __shared__ int warp_tot[33];
  __shared__ int sh_pre;
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const BuildJob& j = a.job[blockIdx.z];
  const int tile = blockIdx.x;
  if (tile * GS_TILE >= j.g.tbl) return;
  if (threadIdx.x < 32) {
    int t = 0;
    for (int k = threadIdx.x; k < tile; k += 32) t += j.g.tile_tot[(size_t)s * j.g.ntiles + k];
    t = warp_sum_i(t);
    if (threadIdx.x == 0) sh_pre = t;
  }
  int* cnt = j.g.cnt + (size_t)s * j.g.tbl + (size_t)tile * GS_TILE;
  const int4 v = reinterpret_cast<const int4*>(cnt)[threadIdx.x];
  reinterpret_cast<int4*>(cnt)[threadIdx.x] = make_int4(0, 0, 0, 0);
  if (j.g.sig) reinterpret_cast<int4*>(j.g.sig + (size_t)s * j.g.tbl + (size_t)tile * GS_TILE)[threadIdx.x] = make_int4(0, 0, 0, 0);
  {
    // occupancy bitmap: 4 buckets per thread, 8 threads per 32-bit word
    unsigned bits = (v.x > 0 ? 1u : 0u) | (v.y > 0 ? 2u : 0u) | (v.z > 0 ? 4u : 0u) | (v.w > 0 ? 8u : 0u);
    bits <<= 4 * (threadIdx.x & 7);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
    if ((threadIdx.x & 7) == 0) j.g.occ[(size_t)s * (j.g.tbl / 32) + (size_t)tile * (GS_TILE / 32) + (threadIdx.x >> 3)] = bits;
  }
  int total;
  const int ex = block_exclusive_scan(v.x + v.y + v.z + v.w, warp_tot, &total);  // syncs: sh_pre is visible after it
  const int b0 = sh_pre + ex;
  const int4 o = make_int4(b0, b0 + v.x, b0 + v.x + v.y, b0 + v.x + v.y + v.z);
  const size_t off = (size_t)tile * GS_TILE;
  reinterpret_cast<int4*>(j.g.cursor + (size_t)s * j.g.tbl + off)[threadIdx.x] = o;
  // cell_start has tbl + 1 entries per sequence, so its rows are not 16-byte aligned: scalar stores
  int* cs = j.g.cell_start + (size_t)s * (j.g.tbl + 1) + off + threadIdx.x * 4;
  cs[0] = o.x; cs[1] = o.y; cs[2] = o.z; cs[3] = o.w;
  if (tile == j.g.ntiles - 1 && threadIdx.x == GS_THREADS - 1) {
    const int n = sh_pre + total;
    j.g.cell_start[(size_t)s * (j.g.tbl + 1) + j.g.tbl] = n;
    j.g.count[s] = n;
  }
}

__global__ void __launch_bounds__(256) k_grid_fill(BuildArgs a) {
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const BuildJob& j = a.job[blockIdx.z];
  const int n = min(j.counts[s * j.count_stride + j.count_off], j.g.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = j.pts[(size_t)s * j.stride + i];
    const uint32_t h = grid_hash(grid_cell(q.x, j.g.inv_cell), grid_cell(q.y, j.g.inv_cell), grid_cell(q.z, j.g.inv_cell), j.g.tbl);
    const int pos = atomicAdd(j.g.cursor + (size_t)s * j.g.tbl + h, 1);
    int w = i;
    if (a.pack_ring) {
      const int idp = min(max((int)q.w + 1, 0), 255);  // ring id = int(intensity), -1..254
      w |= idp << 24;
      if (j.g.sig) atomicOr(j.g.sig + (size_t)s * j.g.tbl + h, 1u << (idp & 31));
    }
    j.g.sorted[(size_t)s * j.g.cap + pos] = make_float4(q.x, q.y, q.z, __int_as_float(w));
  }
}

}  // namespace

void launch_grid_build2(LaunchCtx& ctx, int B, HashGrid& g0, const float4* pts0, int stride0, const int* counts0,
                        int cstride0, int coff0, HashGrid& g1, const float4* pts1, int stride1, const int* counts1,
                        int cstride1, int coff1, const int* enable, int enable_stride, bool pack_ring) {
  BuildArgs a;
  a.job[0] = BuildJob{g0, pts0, stride0, counts0, cstride0, coff0};
  a.job[1] = BuildJob{g1, pts1, stride1, counts1, cstride1, coff1};
  a.enable = enable;
  a.enable_stride = enable_stride;
  a.pack_ring = pack_ring ? 1 : 0;
  const int cap = g0.cap > g1.cap ? g0.cap : g1.cap;
  const int ntiles = g0.ntiles > g1.ntiles ? g0.ntiles : g1.ntiles;
  // enough blocks to fill the GPU at small batches, few enough that large batches do not launch thousands of idle ones
  const int pblocks = (cap + 255) / 256 < 64 ? (cap + 255) / 256 : 64;
  LL_LAUNCH(ctx, "k_grid_count", k_grid_count<<<dim3(pblocks, B, 2), 256, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_tile_sums", k_grid_tile_sums<<<dim3(ntiles, B, 2), GS_THREADS, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_scan", k_grid_scan<<<dim3(ntiles, B, 2), GS_THREADS, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_fill", k_grid_fill<<<dim3(pblocks, B, 2), 256, 0, ctx.stream>>>(a));
}
