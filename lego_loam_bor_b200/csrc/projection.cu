// projection.cu -- ImageProjection::resetParameters / findStartEndAngle / projectPointCloud /
// groundRemoval (reference: LeGO-LOAM/src/imageProjection.cpp:107-150,178-308).
//
// Two kernels, both HBM-bound:
//   k_project_scatter : one thread per input point; computes (row, col) exactly as the reference and
//       resolves the reference's sequential last-writer-wins rule (imageProjection.cpp:214-222) with a
//       64-bit atomicMax on (frame_tag << 32 | point index), so no per-frame clear pass is needed.
//   k_gather_ground   : one thread per (column, 8-row chunk); gathers the winning point of every cell,
//       writes _range_mat / _full_cloud (the fills of resetParameters are folded in here), evaluates the
//       ground test on vertically adjacent cells from registers, and initialises _ground_mat,
//       _label_mat and the union-find forest used by the segmentation kernels.
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

__device__ __forceinline__ bool project_cell(const DevParams& p, const float4 pt, int* row, int* col, float* range_out) {
  const float range = sqrtf(pt.x * pt.x + pt.y * pt.y + pt.z * pt.z);
  const float va = ll_asinf(pt.z / range);
  const float rowf = (va + p.ang_bottom) / p.ang_res_y;
  if (!(rowf == rowf)) return false;
  if (!(rowf > -2147483648.f && rowf < 2147483648.f)) return false;
  const int r = (int)rowf;  // C truncation toward zero (imageProjection.cpp:193)
  if (r < 0 || r >= p.V) return false;
  const float ha = ll_atan2f(pt.x, pt.y);
  // imageProjection.cpp:200, evaluated in double because of M_PI_2 and * 0.5
  const double cd = -round(((double)ha - LL_PI_2) / (double)p.ang_res_x) + p.H * 0.5;
  int c = (int)cd;
  if (c >= p.H) c -= p.H;
  if (c < 0 || c >= p.H) return false;
  if ((double)range < 0.1) return false;
  *row = r;
  *col = c;
  *range_out = range;
  return true;
}

__global__ void __launch_bounds__(256) k_project_scatter(DevState st) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n_in[s]) return;
  const float4 pt = ld_pt(st.in_pts + (size_t)s * st.in_stride + i);
  int row, col;
  float range;
  if (!project_cell(st.p, pt, &row, &col, &range)) return;
  const unsigned long long key = ((unsigned long long)st.frame_tag << 32) | (unsigned)i;
  atomicMax(st.winner + (size_t)s * st.p.N + row * st.p.H + col, key);
}

struct CellVal {
  float4 pt;
  float range;
};

__device__ __forceinline__ CellVal load_cell(const DevState& st, int s, int row, int col) {
  CellVal c;
  const unsigned long long w = st.winner[(size_t)s * st.p.N + row * st.p.H + col];
  if ((uint32_t)(w >> 32) == st.frame_tag) {
    const float4 q = ld_pt(st.in_pts + (size_t)s * st.in_stride + (uint32_t)w);
    c.range = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
    // intensity = (float)row + (float)col / 10000.0 in double (imageProjection.cpp:216)
    c.pt = make_float4(q.x, q.y, q.z, (float)((double)(float)row + (double)(float)col / 10000.0));
  } else {
    const float nanv = __int_as_float(0x7fc00000);
    c.pt = make_float4(nanv, nanv, nanv, 0.f);
    c.range = FLT_MAX;
  }
  return c;
}

// ground test between two vertically adjacent cells (imageProjection.cpp:271-285)
__device__ __forceinline__ bool ground_pair(const DevParams& p, const float4 lower, const float4 upper) {
  const float dX = upper.x - lower.x;
  const float dY = upper.y - lower.y;
  const float dZ = upper.z - lower.z;
  const float ang = ll_atan2f(dZ, sqrtf(dX * dX + dY * dY + dZ * dZ));
  return (double)(ang - p.sensor_mount_angle) <= 10.0 * (LL_PI / 180.0);
}

#define GG_ROWS 8

__global__ void __launch_bounds__(128) k_gather_ground(DevState st) {
  const DevParams& p = st.p;
  const int s = blockIdx.z;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * GG_ROWS;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    // findStartEndAngle (imageProjection.cpp:234-249) + per-frame scalars
    const int n = st.n_in[s];
    float so = 0.f, eo = 0.f, diff = 0.f;
    if (n > 0) {
      const float4 a = ld_pt(st.in_pts + (size_t)s * st.in_stride);
      const float4 b = ld_pt(st.in_pts + (size_t)s * st.in_stride + (n - 1));
      so = -ll_atan2f(a.y, a.x);
      eo = (float)((double)(-ll_atan2f(b.y, b.x)) + 2.0 * LL_PI);
      if ((double)(eo - so) > 3.0 * LL_PI) {
        eo = (float)((double)eo - 2.0 * LL_PI);
      } else if ((double)(eo - so) < LL_PI) {
        eo = (float)((double)eo + 2.0 * LL_PI);
      }
      diff = eo - so;
    }
    st.orientation[s * 4 + 0] = so;
    st.orientation[s * 4 + 1] = eo;
    st.orientation[s * 4 + 2] = diff;
    st.half_idx[s] = 0x7fffffff;
  }
  if (col >= p.H) return;
  const int r1 = min(p.V, r0 + GG_ROWS);
  const size_t base = (size_t)s * p.N;
  // halo below: pair (r0-1, r0) decides whether row r0 is ground
  bool mark_prev = false;  // pair (row-1, row) marked
  CellVal prev;
  prev.range = FLT_MAX;
  prev.pt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r0 >= 1 && r0 <= p.gsi) prev = load_cell(st, s, r0 - 1, col);
#pragma unroll 2
  for (int row = r0; row <= r1; ++row) {
    // row == r1 is the halo above (only needed for its pair with r1-1)
    const bool need = (row < r1) || (row < p.V && row <= p.gsi);
    CellVal cur;
    cur.range = FLT_MAX;
    cur.pt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (need) cur = load_cell(st, s, row, col);
    bool mark = false;
    if (need && row >= 1 && row <= p.gsi) mark = ground_pair(p, prev.pt, cur.pt);
    if (row > r0) {
      // finalise row-1: ground if either adjacent pair was marked
      const int cell = (row - 1) * p.H + col;
      const bool g = mark_prev || mark;
      const bool no_label = g || (prev.range == FLT_MAX);
      st.range_mat[base + cell] = prev.range;
      st.full_cloud[base + cell] = prev.pt;
      st.ground_mat[base + cell] = g ? 1 : 0;
      st.label_mat[base + cell] = no_label ? -1 : 0;
      st.parent[base + cell] = no_label ? -1 : cell;
      st.comp_size[base + cell] = 0;
      st.comp_rows[base + cell] = 0u;
    }
    mark_prev = mark;
    prev = cur;
  }
}

}  // namespace

void launch_projection(LaunchCtx& ctx, DevState& st) {
  const DevParams& p = st.p;
  {
    dim3 grid((p.max_pts + 255) / 256, p.B);
    LL_LAUNCH(ctx, "k_project_scatter", k_project_scatter<<<grid, 256, 0, ctx.stream>>>(st));
  }
  {
    dim3 grid((p.H + 127) / 128, (p.V + GG_ROWS - 1) / GG_ROWS, p.B);
    LL_LAUNCH(ctx, "k_gather_ground", k_gather_ground<<<grid, 128, 0, ctx.stream>>>(st));
  }
}
