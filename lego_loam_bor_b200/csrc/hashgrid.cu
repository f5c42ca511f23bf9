// hashgrid.cu -- build of the voxel hash grid (see hashgrid.cuh): count (+ tile totals), scan, fill.
// Replaces KdTreeFLANN::setInputCloud / buildIndex (nanoflann_pcl.h:131-138), which the reference
// calls for both last-frame clouds every frame (featureAssociation.cpp:1356-1359) and for both local
// maps every mapping cycle (mapOptmization.cpp:1317-1318).  Two grids (corner + surf) are built by the
// same three launches (blockIdx.z selects the grid).  The bucket counters and tile totals are zero at
// rest: the fill counts the buckets back down, so no clear pass and no separate cursor array is needed.
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

// bucket counters + per-tile totals (a tile = GS_TILE consecutive buckets: the unit of the scan kernel)
__global__ void __launch_bounds__(256) k_grid_count(BuildArgs a) {
  extern __shared__ int sh_tile[];  // [ntiles of this launch]
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const BuildJob& j = a.job[blockIdx.z];
  const int n = min(j.counts[s * j.count_stride + j.count_off], j.g.cap);
  if ((int)(blockIdx.x * blockDim.x) >= n) return;
  for (int t = threadIdx.x; t < j.g.ntiles; t += blockDim.x) sh_tile[t] = 0;
  __syncthreads();
  // grid-stride: the launch is sized for a typical cloud, not for the capacity (most capacity-sized blocks would be empty)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = j.pts[(size_t)s * j.stride + i];
    const uint32_t h = grid_hash(grid_cell(q.x, j.g.inv_cell), grid_cell(q.y, j.g.inv_cell), grid_cell(q.z, j.g.inv_cell), j.g.tbl);
    atomicAdd(j.g.cnt + (size_t)s * j.g.tbl + h, 1);
    atomicAdd(&sh_tile[h / GS_TILE], 1);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < j.g.ntiles; t += blockDim.x)
    if (sh_tile[t]) atomicAdd(j.g.tile_tot + (size_t)s * j.g.ntiles + t, sh_tile[t]);
}

// exclusive scan: cell_start from the counters (which stay: the fill counts them back down to zero)
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
  const size_t off = (size_t)tile * GS_TILE;
  const int4 v = reinterpret_cast<const int4*>(j.g.cnt + (size_t)s * j.g.tbl + off)[threadIdx.x];
  if (j.g.sig) reinterpret_cast<int4*>(j.g.sig + (size_t)s * j.g.tbl + off)[threadIdx.x] = make_int4(0, 0, 0, 0);
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
  int* cs_row = j.g.cell_start + (size_t)s * grid_cs_stride(j.g);   // rows are 16-byte aligned (grid_cs_stride)
  reinterpret_cast<int4*>(cs_row + off)[threadIdx.x] = make_int4(b0, b0 + v.x, b0 + v.x + v.y, b0 + v.x + v.y + v.z);
  if (tile == j.g.ntiles - 1 && threadIdx.x == GS_THREADS - 1) {
    const int n = sh_pre + total;
    cs_row[j.g.tbl] = n;
    j.g.count[s] = n;
  }
}

// counting-sort scatter: a point takes the LAST free slot of its bucket (the counter counts down and is zero again when
// the bucket is full; the order inside a bucket is arbitrary either way, every query breaks ties by index)
__global__ void __launch_bounds__(256) k_grid_fill(BuildArgs a) {
  const int s = blockIdx.y;
  if (!seq_enabled(a, s)) return;
  const BuildJob& j = a.job[blockIdx.z];
  if (blockIdx.x == 0)   // the tile totals are zero at rest (every scan block of this build has read them)
    for (int t = threadIdx.x; t < j.g.ntiles; t += blockDim.x) j.g.tile_tot[(size_t)s * j.g.ntiles + t] = 0;
  const int n = min(j.counts[s * j.count_stride + j.count_off], j.g.cap);
  const int* cs_row = j.g.cell_start + (size_t)s * grid_cs_stride(j.g);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = j.pts[(size_t)s * j.stride + i];
    const uint32_t h = grid_hash(grid_cell(q.x, j.g.inv_cell), grid_cell(q.y, j.g.inv_cell), grid_cell(q.z, j.g.inv_cell), j.g.tbl);
    const int pos = cs_row[h] + atomicSub(j.g.cnt + (size_t)s * j.g.tbl + h, 1) - 1;
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
  LL_LAUNCH(ctx, "k_grid_count", k_grid_count<<<dim3(pblocks, B, 2), 256, (size_t)ntiles * sizeof(int), ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_scan", k_grid_scan<<<dim3(ntiles, B, 2), GS_THREADS, 0, ctx.stream>>>(a));
  LL_LAUNCH(ctx, "k_grid_fill", k_grid_fill<<<dim3(pblocks, B, 2), 256, 0, ctx.stream>>>(a));
}
