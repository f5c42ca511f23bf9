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

// ---- discrete decisions without the exact (double-precision) inverse trigonometry -----------------------------
// Row, column and the ground flag are step functions of one angle each, and every step of their evaluation (float
// add / divide, double subtract / divide / round, truncation) is monotonic in that angle.  CUDA's asinf / atan2f are
// within a few ulp of the true value (2 and 3 ulp in the CUDA C Programming Guide) and the reference value -- the
// correctly rounded float, ll_asinf / ll_atan2f -- within 0.5 ulp, so if the step function gives the same result FAST_ULPS
// ulps below and above the fast value, it gives that result for the reference value too.  Only the (rare) angles next to
// a step pay for the exact evaluation.
#define FAST_ULPS 16

__device__ __forceinline__ float step_ulps(float v, int n) {
  int b = __float_as_int(v);
  int o = b >= 0 ? b : (int)(0x80000000u - (unsigned)b);  // order-preserving integer image of the float
  o += n;
  b = o >= 0 ? o : (int)(0x80000000u - (unsigned)o);
  return __int_as_float(b);
}

// imageProjection.cpp:193-196 as a monotonic class: -1 below the image, V above, else the row
__device__ __forceinline__ int row_class(const DevParams& p, float va) {
  const float rowf = (va + p.ang_bottom) / p.ang_res_y;
  if (!(rowf > -2147483648.f)) return -1;
  if (!(rowf < 2147483648.f)) return p.V;
  const int r = (int)rowf;  // C truncation toward zero (imageProjection.cpp:193)
  return r < 0 ? -1 : (r >= p.V ? p.V : r);
}

// imageProjection.cpp:200 before the wrap, evaluated in double because of M_PI_2 and * 0.5 (non-increasing in ha)
__device__ __forceinline__ double col_value(const DevParams& p, float ha) {
  return -round(((double)ha - LL_PI_2) / (double)p.ang_res_x) + p.H * 0.5;
}

__device__ __forceinline__ bool project_cell(const DevParams& p, const float4 pt, int* row, int* col, float* range_out) {
  const float range = sqrtf(pt.x * pt.x + pt.y * pt.y + pt.z * pt.z);
  const float q = pt.z / range;
  int r;
  {
    const float vf = asinf(q);
    const int r_lo = row_class(p, step_ulps(vf, -FAST_ULPS)), r_hi = row_class(p, step_ulps(vf, FAST_ULPS));
    if (vf == vf && r_lo == r_hi) {
      r = r_lo;
    } else {
      const float va = ll_asinf(q);
      const float rowf = (va + p.ang_bottom) / p.ang_res_y;
      if (!(rowf == rowf)) return false;
      r = row_class(p, va);
    }
  }
  if (r < 0 || r >= p.V) return false;
  double cd;
  {
    const float hf = atan2f(pt.x, pt.y);
    const double c_lo = col_value(p, step_ulps(hf, FAST_ULPS)), c_hi = col_value(p, step_ulps(hf, -FAST_ULPS));
    cd = (hf == hf && c_lo == c_hi) ? c_lo : col_value(p, ll_atan2f(pt.x, pt.y));
  }
  int c = (int)cd;
  if (c >= p.H) c -= p.H;
  if (c < 0 || c >= p.H) return false;
  if ((double)range < 0.1) return false;
  *row = r;
  *col = c;
  *range_out = range;
  return true;
}

// input point i of sequence s; the intensity the sensor reported is never read (imageProjection.cpp:216 overwrites it)
__device__ __forceinline__ float4 ld_in(const DevState& st, int s, uint32_t i) {
  if (st.in_xyz3) {
    const float* q = reinterpret_cast<const float*>(st.in_pts) + ((size_t)s * st.in_stride + i) * 3;
    return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
  }
  return ld_pt(st.in_pts + (size_t)s * st.in_stride + i);
}

__global__ void __launch_bounds__(256) k_project_scatter(DevState st) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n_in[s]) return;
  const float4 pt = ld_in(st, s, (uint32_t)i);
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
    const float4 q = ld_in(st, s, (uint32_t)w);
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
  const float nrm = sqrtf(dX * dX + dY * dY + dZ * dZ);
  const double thr = 10.0 * (LL_PI / 180.0);
  // the test is monotonic in the angle: decide from the fast atan2f when both ends of its error interval agree
  const float af = atan2f(dZ, nrm);
  const bool g_lo = (double)(step_ulps(af, -FAST_ULPS) - p.sensor_mount_angle) <= thr;
  const bool g_hi = (double)(step_ulps(af, FAST_ULPS) - p.sensor_mount_angle) <= thr;
  if (af == af && g_lo == g_hi) return g_lo;
  const float ang = ll_atan2f(dZ, nrm);
  return (double)(ang - p.sensor_mount_angle) <= thr;
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
      const float4 a = ld_in(st, s, 0u);
      const float4 b = ld_in(st, s, (uint32_t)(n - 1));
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
