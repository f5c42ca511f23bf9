// hashgrid.cuh -- exact nearest-neighbour queries on a GPU-built voxel hash grid.
//
// Stands in for nanoflann::KdTreeFLANN<PointType> (include/lego_loam/nanoflann_pcl.h:54-152,
// nanoflann.hpp:857-1003,1190-1241,1346-1409): both are EXACT k-NN under squared L2 in float
// with d = query - point accumulated ((0+dx^2)+dy^2)+dz^2 (nanoflann.hpp:432-440), so they return the
// same neighbours; only the order among exactly equal distances could differ (lowest index here).
//
// Layout: points are counting-sorted by hash bucket of their cell (cell size `cell`), so a bucket
// is one contiguous run of float4 (x, y, z, original index).  A query visits the (2r+1)^3 block of
// cells around its own cell, r = 0, 1, ...; every point closer than r*cell is guaranteed to be inside
// that block, which gives an exact termination test.  Hash collisions only add candidates.
#pragma once

#include "ll_device.cuh"

__device__ __forceinline__ int grid_cell(float v, float inv_cell) { return (int)floorf(v * inv_cell); }

__device__ __forceinline__ float nn_dist2(float qx, float qy, float qz, const float4 p) {
  float r = 0.f;
  float d = qx - p.x;
  r += d * d;
  d = qy - p.y;
  r += d * d;
  d = qz - p.z;
  r += d * d;
  return r;
}

// Warp-cooperative exact 1-NN: all 32 lanes pass the same query; returns (in every lane) the
// nearest point with d2 < max_d2, or idx = -1.
__device__ __forceinline__ void warp_nn1(const HashGrid& g, int s, float qx, float qy, float qz, float max_d2,
                                         float* out_d2, int* out_idx) {
  const int lane = threadIdx.x & 31;
  const int* cs = g.cell_start + (size_t)s * grid_cs_stride(g);
  const unsigned* occ = g.occ + (size_t)s * (g.tbl / 32);
  const float4* pts = g.sorted + (size_t)s * g.cap;
  const int cx = grid_cell(qx, g.inv_cell), cy = grid_cell(qy, g.inv_cell), cz = grid_cell(qz, g.inv_cell);
  float best = max_d2;
  int bidx = 0x7fffffff;
  const int rmax = (int)ceilf(sqrtf(max_d2) * g.inv_cell);
  for (int r = 0; r <= rmax; ++r) {
    const int side = 2 * r + 1, ncell = side * side * side;
    for (int t = lane; t < ncell; t += 32) {
      const int dz = t / (side * side) - r;
      const int rem = t % (side * side);
      const int dy = rem / side - r, dx = rem % side - r;
      if (max(max(abs(dx), abs(dy)), abs(dz)) != r) continue;  // inner cells were visited at smaller r
      const uint32_t h = grid_hash(cx + dx, cy + dy, cz + dz, g.tbl);
      if (!((occ[h >> 5] >> (h & 31)) & 1u)) continue;  // empty bucket
      const int b0 = cs[h], b1 = cs[h + 1];
      for (int k = b0; k < b1; ++k) {
        const float4 q = pts[k];
        const float d2 = nn_dist2(qx, qy, qz, q);
        const int idx = __float_as_int(q.w);
        if (d2 < best || (d2 == best && idx < bidx)) { best = d2; bidx = idx; }
      }
    }
    const float wbest = warp_min_f(best);
    const float reach = (float)r * g.cell;
    if (wbest < reach * reach * 0.9999f) break;  // nothing outside the visited block can be closer
  }
  // lexicographic (d2, idx) minimum over the warp
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (od < best || (od == best && oi < bidx)) { best = od; bidx = oi; }
  }
  *out_d2 = best;
  *out_idx = (bidx == 0x7fffffff) ? -1 : bidx;
}

// Thread-level exact 5-NN restricted to d2 < max_d2 where max_d2 <= cell^2 (so the 27-cell block is
// exact).  Results ascending by (d2, idx); returns the number found (<= 5).
__device__ __forceinline__ int thread_knn5(const HashGrid& g, int s, float qx, float qy, float qz, float max_d2,
                                           float* d2o, int* idxo) {
  const int* cs = g.cell_start + (size_t)s * grid_cs_stride(g);
  const float4* pts = g.sorted + (size_t)s * g.cap;
  const int cx = grid_cell(qx, g.inv_cell), cy = grid_cell(qy, g.inv_cell), cz = grid_cell(qz, g.inv_cell);
  int n = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) { d2o[i] = max_d2; idxo[i] = 0x7fffffff; }
  for (int t = 0; t < 27; ++t) {
    const int dz = t / 9 - 1, dy = (t % 9) / 3 - 1, dx = t % 3 - 1;
    const int ix = cx + dx, iy = cy + dy, iz = cz + dz;
    const uint32_t h = grid_hash(ix, iy, iz, g.tbl);
    const int b0 = cs[h], b1 = cs[h + 1];
    for (int k = b0; k < b1; ++k) {
      const float4 q = pts[k];
      // hash collisions: only accept points whose own cell is the one being visited (no duplicates)
      if (grid_cell(q.x, g.inv_cell) != ix || grid_cell(q.y, g.inv_cell) != iy || grid_cell(q.z, g.inv_cell) != iz) continue;
      const float d2 = nn_dist2(qx, qy, qz, q);
      if (!(d2 < max_d2)) continue;
      const int idx = __float_as_int(q.w);
      if (d2 < d2o[4] || (d2 == d2o[4] && idx < idxo[4])) {
        // insertion into the sorted list of 5
        float cd = d2;
        int ci = idx;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          if (cd < d2o[i] || (cd == d2o[i] && ci < idxo[i])) {
            const float td = d2o[i]; const int ti = idxo[i];
            d2o[i] = cd; idxo[i] = ci;
            cd = td; ci = ti;
          }
        }
        if (n < 5) ++n;
      }
    }
  }
  return n;
}
