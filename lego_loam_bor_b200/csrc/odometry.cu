// odometry.cu -- scan-to-scan LM odometry of FeatureAssociation
// (reference: LeGO-LOAM/src/featureAssociation.cpp:388-500, 503-1032, 1181-1270, 1329-1359).
//
//   k_odom_search<SURF|CORNER>  one warp per feature point: TransformToStart, exact 1-NN in the
//       last-frame cloud (hash grid instead of the kd-tree) and the reference's ring-window scans,
//       with its loop bounds and visiting-order tie-breaks.  Runs for LM
//       iteration 0 of each stage for all sequences at once.
//   k_odom_lm<SURF|CORNER>      one block per sequence: five LM iterations of one stage per launch
//       (5 launches per stage, no-ops once converged), no host round trips.  Per iteration: residual + Jacobian row per feature, J^T J / J^T r
//       as exact double products reduced by warp shuffles + a fixed-order cross-warp sum, 3x3
//       column-pivoted Householder solve, degeneracy test at iteration 0, convergence test.
//       Correspondences are refreshed by k_odom_search before iterations 0, 5, 10, 15, 20 (featureAssociation.cpp:511).
//   k_odom_finish               integrateTransformation (one thread per sequence)
//   k_publish_clouds_last       TransformToEnd on the less-sharp / less-flat clouds into the
//       "last" buffers, adjustOutlierCloud, counts and the kd-tree-rebuild condition.
#include "../../include/ll_smallmat.h"
#include "hashgrid.cuh"
#include "ll_kernels.h"

namespace {

enum { STAGE_SURF = 0, STAGE_CORNER = 1 };

__device__ __forceinline__ float4 transform_to_start(const float4 pi, const float* T) {
  // featureAssociation.cpp:388-418
  const float s = 10 * (pi.w - (float)(int)pi.w);
  const float ry = s * T[1], rx = s * T[0], rz = s * T[2];
  const float tx = s * T[3], ty = s * T[4], tz = s * T[5];
  float srz, crz, srx, crx, sry, cry;
  ll_sincosf(rz, &srz, &crz);
  ll_sincosf(rx, &srx, &crx);
  ll_sincosf(ry, &sry, &cry);
  const float x1 = crz * (pi.x - tx) + srz * (pi.y - ty);
  const float y1 = -srz * (pi.x - tx) + crz * (pi.y - ty);
  const float z1 = (pi.z - tz);
  const float x2 = x1;
  const float y2 = crx * y1 + srx * z1;
  const float z2 = -srx * y1 + crx * z1;
  return make_float4(cry * x2 - sry * z2, y2, sry * x2 + cry * z2, pi.w);
}

__device__ __forceinline__ float4 transform_to_end(const float4 pi, const float* T) {
  // featureAssociation.cpp:422-471
  const float s = 10 * (pi.w - (float)(int)pi.w);
  float rx = s * T[0], ry = s * T[1], rz = s * T[2];
  float tx = s * T[3], ty = s * T[4], tz = s * T[5];
  float srz, crz, srx, crx, sry, cry;
  ll_sincosf(rz, &srz, &crz);
  ll_sincosf(rx, &srx, &crx);
  ll_sincosf(ry, &sry, &cry);
  const float x1 = crz * (pi.x - tx) + srz * (pi.y - ty);
  const float y1 = -srz * (pi.x - tx) + crz * (pi.y - ty);
  const float z1 = (pi.z - tz);
  const float x2 = x1;
  const float y2 = crx * y1 + srx * z1;
  const float z2 = -srx * y1 + crx * z1;
  const float x3 = cry * x2 - sry * z2;
  const float y3 = y2;
  const float z3 = sry * x2 + cry * z2;
  rx = T[0]; ry = T[1]; rz = T[2];
  tx = T[3]; ty = T[4]; tz = T[5];
  ll_sincosf(rz, &srz, &crz);
  ll_sincosf(rx, &srx, &crx);
  ll_sincosf(ry, &sry, &cry);
  const float x4 = cry * x3 + sry * z3;
  const float y4 = y3;
  const float z4 = -sry * x3 + cry * z3;
  const float x5 = x4;
  const float y5 = crx * y4 - srx * z4;
  const float z5 = srx * y4 + crx * z4;
  const float x6 = crz * x5 - srz * y5 + tx;
  const float y6 = srz * x5 + crz * y5 + ty;
  const float z6 = z5 + tz;
  return make_float4(x6, y6, z6, (float)(int)pi.w);
}

__device__ __forceinline__ float sq_dist_ref(const float4 a, const float4 b) {
  // (a.x - b.x)^2 + (a.y - b.y)^2 + (a.z - b.z)^2, left to right (featureAssociation.cpp:528-533)
  return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z);
}

// running minimum with the reference's "first strictly smaller wins" rule, expressed as a
// lexicographic (distance, visiting order) minimum so that a warp can evaluate it in parallel
struct OrdMin {
  float d2;
  int ord;
  int idx;
};
__device__ __forceinline__ void ordmin_update(OrdMin& m, float d2, int ord, int idx) {
  if (d2 < m.d2 || (d2 == m.d2 && m.idx >= 0 && ord < m.ord)) { m.d2 = d2; m.ord = ord; m.idx = idx; }
}
__device__ __forceinline__ void ordmin_warp_reduce(OrdMin& m) {
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, m.d2, o);
    const int oo = __shfl_xor_sync(0xffffffffu, m.ord, o);
    const int oi = __shfl_xor_sync(0xffffffffu, m.idx, o);
    if (oi >= 0 && (m.idx < 0 || od < m.d2 || (od == m.d2 && oo < m.ord))) { m.d2 = od; m.ord = oo; m.idx = oi; }
  }
}

// Correspondence search for one feature point, executed by one warp.
// SURF: featureAssociation.cpp:649-718; CORNER: featureAssociation.cpp:511-568.
template <int STAGE>
__device__ __forceinline__ void warp_find_correspondence(const DevState& st, int s, int i, const float* T, int lane) {
  const DevParams& p = st.p;
  const bool surf = (STAGE == STAGE_SURF);
  const float4* cur = surf ? st.surf_flat + (size_t)s * p.cap_flat : st.corner_sharp + (size_t)s * p.cap_sharp;
  const float4* last = surf ? st.surf_last + (size_t)s * p.N : st.corner_last + (size_t)s * p.cap_less_sharp;
  const int cur_n = st.feat_counts[s * 4 + (surf ? 2 : 0)];
  const int last_n = st.last_counts[s * 2 + (surf ? 1 : 0)];
  const HashGrid& g = surf ? st.grid_surf_last : st.grid_corner_last;
  const float4 sel = transform_to_start(cur[i], T);
  float nd2;
  int nidx;
  warp_nn1(g, s, sel.x, sel.y, sel.z, p.nearest_feature_dist_sqr, &nd2, &nidx);
  int closest = -1, ind2 = -1, ind3 = -1;
  if (nidx >= 0 && nidx < last_n) {  // nd2 < nearest_feature_dist_sqr by construction
    closest = nidx;
    const int cs = (int)last[closest].w;
    OrdMin m2{p.nearest_feature_dist_sqr, 0, -1}, m3{p.nearest_feature_dist_sqr, 0, -1};
    // The reference scans upwards from closest+1 until the first point whose ring id exceeds cs + 2.5 and
    // downwards from closest-1 until the first id below cs - 2.5 (featureAssociation.cpp:522-563,661-712).
    // Ring ids are non-decreasing along the cloud up to a -1 wobble (negative relTime truncates to ring-1),
    // which is enough to make those break positions a function of cs alone: the first index of the WHOLE
    // cloud with id >= cs+3 and the last index with id <= cs-3 (tables built by k_window_tables).
    const int* tab = st.win_tab + ((size_t)s * 2 + (surf ? 1 : 0)) * 2 * (LL_MAX_RINGS + 8);
    const int up_break = tab[min(max(cs + 3, 0), LL_MAX_RINGS + 7)];
    const int dn_break = (cs - 3 >= 0) ? tab[(LL_MAX_RINGS + 8) + min(cs - 3, LL_MAX_RINGS + 7)] : -1;
    // upward scan; also bounded by the CURRENT frame's feature count (sic, featureAssociation.cpp:522,661)
    const int jend = min(min(cur_n, last_n), up_break);
    for (int b0 = closest + 1; b0 < jend; b0 += 32) {
      const int j = b0 + lane;
      if (j < jend) {
        const float4 q = last[j];
        const int id = (int)q.w;
        const float d2 = sq_dist_ref(q, sel);
        const int ord = j - closest;
        if (surf) {
          if (id <= cs) ordmin_update(m2, d2, ord, j); else ordmin_update(m3, d2, ord, j);
        } else {
          if (id > cs) ordmin_update(m2, d2, ord, j);
        }
      }
    }
    for (int b0 = closest - 1; b0 > dn_break; b0 -= 32) {
      const int j = b0 - lane;
      if (j > dn_break) {
        const float4 q = last[j];
        const int id = (int)q.w;
        const float d2 = sq_dist_ref(q, sel);
        const int ord = 0x40000000 + (closest - j);
        if (surf) {
          if (id >= cs) ordmin_update(m2, d2, ord, j); else ordmin_update(m3, d2, ord, j);
        } else {
          if (id < cs) ordmin_update(m2, d2, ord, j);
        }
      }
    }
    ordmin_warp_reduce(m2);
    ind2 = m2.idx;
    if (surf) {
      ordmin_warp_reduce(m3);
      ind3 = m3.idx;
    }
  }
  if (lane == 0) {
    if (surf) {
      int* c = st.corr_surf + ((size_t)s * p.cap_flat + i) * 3;
      c[0] = closest; c[1] = ind2; c[2] = ind3;
    } else {
      int* c = st.corr_corner + ((size_t)s * p.cap_sharp + i) * 2;
      c[0] = closest; c[1] = ind2;
    }
  }
}

__device__ __forceinline__ bool odom_guard(const DevState& st, int s) {
  // featureAssociation.cpp:1214
  return !(st.last_counts[s * 2 + 0] < 10 || st.last_counts[s * 2 + 1] < 100);
}

template <int STAGE>
__global__ void __launch_bounds__(256) k_odom_search(DevState st) {
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  if (!odom_guard(st, s) || st.odom_flags[s * 4 + 1]) return;  // flag 1: this stage has converged
  const int n = st.feat_counts[s * 4 + (STAGE == STAGE_SURF ? 2 : 0)];
  float T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) T[k] = st.transform_cur[s * 6 + k];
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps)
    warp_find_correspondence<STAGE>(st, s, i, T, threadIdx.x & 31);
}

// One accepted correspondence -> one row [a0 a1 a2 | b] of the 3-column system.
struct Row3 {
  float a0, a1, a2, b;
  bool ok;
};

struct SurfCoef {  // featureAssociation.cpp:797-832
  float a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, b1, b2, b5, b6, c1, c2, c3, c4, c5, c6, c7, c8, c9, crx;
};
struct CornerCoef {  // featureAssociation.cpp:939-958
  float b1, b2, b3, b4, b5, b6, b7, b8, c5, srx;
};

__device__ __forceinline__ SurfCoef make_surf_coef(const float* T) {
  SurfCoef c;
  float srx, crx, sry, cry, srz, crz;
  ll_sincosf(T[0], &srx, &crx);
  ll_sincosf(T[1], &sry, &cry);
  ll_sincosf(T[2], &srz, &crz);
  const float tx = T[3], ty = T[4], tz = T[5];
  c.crx = crx;
  c.a1 = crx * sry * srz; c.a2 = crx * crz * sry; c.a3 = srx * sry;
  c.a4 = tx * c.a1 - ty * c.a2 - tz * c.a3;
  c.a5 = srx * srz; c.a6 = crz * srx;
  c.a7 = ty * c.a6 - tz * crx - tx * c.a5;
  c.a8 = crx * cry * srz; c.a9 = crx * cry * crz; c.a10 = cry * srx;
  c.a11 = tz * c.a10 + ty * c.a9 - tx * c.a8;
  c.b1 = -crz * sry - cry * srx * srz;
  c.b2 = cry * crz * srx - sry * srz;
  c.b5 = cry * crz - srx * sry * srz;
  c.b6 = cry * srz + crz * srx * sry;
  c.c1 = -c.b6; c.c2 = c.b5;
  c.c3 = tx * c.b6 - ty * c.b5;
  c.c4 = -crx * crz; c.c5 = crx * srz;
  c.c6 = ty * c.c5 + tx * -c.c4;
  c.c7 = c.b2; c.c8 = -c.b1;
  c.c9 = tx * -c.b2 - ty * -c.b1;
  return c;
}

__device__ __forceinline__ CornerCoef make_corner_coef(const float* T) {
  CornerCoef c;
  float srx, crx, sry, cry, srz, crz;
  ll_sincosf(T[0], &srx, &crx);
  ll_sincosf(T[1], &sry, &cry);
  ll_sincosf(T[2], &srz, &crz);
  const float tx = T[3], ty = T[4], tz = T[5];
  c.srx = srx;
  c.b1 = -crz * sry - cry * srx * srz;
  c.b2 = cry * crz * srx - sry * srz;
  c.b3 = crx * cry;
  c.b4 = tx * -c.b1 + ty * -c.b2 + tz * c.b3;
  c.b5 = cry * crz - srx * sry * srz;
  c.b6 = cry * srz + crz * srx * sry;
  c.b7 = crx * sry;
  c.b8 = tz * c.b7 - ty * c.b6 - tx * c.b5;
  c.c5 = crx * srz;
  return c;
}

// featureAssociation.cpp:721-777 + 834-857
__device__ __forceinline__ Row3 surf_row(const DevState& st, int s, int i, const float* T, const SurfCoef& k, int iter) {
  const DevParams& p = st.p;
  Row3 r;
  r.ok = false;
  const int* c = st.corr_surf + ((size_t)s * p.cap_flat + i) * 3;
  const int i1 = c[0], i2 = c[1], i3 = c[2];
  if (!(i2 >= 0 && i3 >= 0)) return r;
  const float4 ori = st.surf_flat[(size_t)s * p.cap_flat + i];
  const float4 sel = transform_to_start(ori, T);
  const float4* last = st.surf_last + (size_t)s * p.N;
  const float4 t1 = last[i1], t2 = last[i2], t3 = last[i3];
  float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
  float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
  float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
  float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
  const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
  float w = 1;
  if (iter >= 5) {
    w = (float)(1 - 1.8 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
  }
  if (!((double)w > 0.1 && pd2 != 0)) return r;
  const float cx = w * pa, cy = w * pb, cz = w * pc, ci = w * pd2;
  const float arx = (-k.a1 * ori.x + k.a2 * ori.y + k.a3 * ori.z + k.a4) * cx +
                    (k.a5 * ori.x - k.a6 * ori.y + k.crx * ori.z + k.a7) * cy +
                    (k.a8 * ori.x - k.a9 * ori.y - k.a10 * ori.z + k.a11) * cz;
  const float arz = (k.c1 * ori.x + k.c2 * ori.y + k.c3) * cx + (k.c4 * ori.x - k.c5 * ori.y + k.c6) * cy +
                    (k.c7 * ori.x + k.c8 * ori.y + k.c9) * cz;
  const float aty = -k.b6 * cx + k.c4 * cy + k.b2 * cz;
  r.a0 = arx; r.a1 = arz; r.a2 = aty;
  r.b = (float)(-0.05 * (double)ci);
  r.ok = true;
  return r;
}

// featureAssociation.cpp:571-635 + 960-978
__device__ __forceinline__ Row3 corner_row(const DevState& st, int s, int i, const float* T, const CornerCoef& k, int iter) {
  const DevParams& p = st.p;
  Row3 r;
  r.ok = false;
  const int* c = st.corr_corner + ((size_t)s * p.cap_sharp + i) * 2;
  const int i1 = c[0], i2 = c[1];
  if (!(i2 >= 0)) return r;
  const float4 ori = st.corner_sharp[(size_t)s * p.cap_sharp + i];
  const float4 sel = transform_to_start(ori, T);
  const float4* last = st.corner_last + (size_t)s * p.cap_less_sharp;
  const float4 t1 = last[i1], t2 = last[i2];
  const float x0 = sel.x, y0 = sel.y, z0 = sel.z;
  const float x1 = t1.x, y1 = t1.y, z1 = t1.z;
  const float x2 = t2.x, y2 = t2.y, z2 = t2.z;
  const float m11 = ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1));
  const float m22 = ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1));
  const float m33 = ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1));
  const float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
  const float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
  const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
  const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
  const float ld2 = a012 / l12;
  float w = 1;
  if (iter >= 5) w = (float)(1 - 1.8 * (double)fabsf(ld2));
  if (!((double)w > 0.1 && ld2 != 0)) return r;
  const float cx = w * la, cy = w * lb, cz = w * lc, ci = w * ld2;
  const float ary = (k.b1 * ori.x + k.b2 * ori.y - k.b3 * ori.z + k.b4) * cx +
                    (k.b5 * ori.x + k.b6 * ori.y - k.b7 * ori.z + k.b8) * cz;
  const float atx = -k.b5 * cx + k.c5 * cy + k.b1 * cz;
  const float atz = k.b7 * cx - k.srx * cy - k.b3 * cz;
  r.a0 = ary; r.a1 = atx; r.a2 = atz;
  r.b = (float)(-0.05 * (double)ci);
  r.ok = true;
  return r;
}

#define LM_THREADS 512
#define LM_WARPS (LM_THREADS / 32)

// Iterations [it_begin, it_begin + 5) of one LM stage; the correspondences were refreshed by
// k_odom_search right before (featureAssociation.cpp:511: every 5th iteration).  A sequence that has
// converged (or failed the guard) turns the remaining launches of its stage into no-ops.
template <int STAGE>
__global__ void __launch_bounds__(LM_THREADS) k_odom_lm(DevState st, int it_begin) {
  __shared__ float sT[6];
  __shared__ double sh_part[LM_WARPS][10];
  __shared__ int sh_state[4];  // 0: stop flag, 1: iterations run
  const DevParams& p = st.p;
  const int s = blockIdx.x;
  const bool surf = (STAGE == STAGE_SURF);
  if (!odom_guard(st, s)) {
    if (threadIdx.x == 0 && it_begin == 0) st.odom_iters[s * 2 + STAGE] = 0;
    return;
  }
  if (st.odom_flags[s * 4 + 1]) return;
  const int n = st.feat_counts[s * 4 + (surf ? 2 : 0)];
  if (threadIdx.x < 6) sT[threadIdx.x] = st.transform_cur[s * 6 + threadIdx.x];
  if (threadIdx.x == 0) { sh_state[0] = 0; sh_state[1] = it_begin; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int iter = it_begin; iter < it_begin + 5; ++iter) {
    float T[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) T[k] = sT[k];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    SurfCoef ks;
    CornerCoef kc;
    if (surf) ks = make_surf_coef(T); else kc = make_corner_coef(T);
    for (int i = threadIdx.x; i < n; i += LM_THREADS) {
      const Row3 r = surf ? surf_row(st, s, i, T, ks, iter) : corner_row(st, s, i, T, kc, iter);
      if (r.ok) {
        const double a0 = r.a0, a1 = r.a1, a2 = r.a2, b = r.b;
        acc[0] += a0 * a0; acc[1] += a0 * a1; acc[2] += a0 * a2;
        acc[3] += a1 * a1; acc[4] += a1 * a2; acc[5] += a2 * a2;
        acc[6] += a0 * b; acc[7] += a1 * b; acc[8] += a2 * b;
        acc[9] += 1.0;
      }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = warp_sum_d(acc[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 10; ++k) sh_part[wid][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot[10];
      for (int k = 0; k < 10; ++k) {
        double v = 0.0;
        for (int w = 0; w < LM_WARPS; ++w) v += sh_part[w][k];
        tot[k] = v;
      }
      sh_state[1] = iter + 1;
      const int rows = (int)tot[9];
      if (rows >= 10) {  // featureAssociation.cpp:1222,1232
        float AtA[9] = {(float)tot[0], (float)tot[1], (float)tot[2], (float)tot[1], (float)tot[3],
                        (float)tot[4], (float)tot[2], (float)tot[4], (float)tot[5]};
        float AtB[3] = {(float)tot[6], (float)tot[7], (float)tot[8]};
        float A2[9], X[3];
        for (int k = 0; k < 9; ++k) A2[k] = AtA[k];
        llm::colpiv_qr_solve<3, 3>(A2, AtB, X);
        float* matP = st.odom_matP + s * 9;
        if (iter == 0) st.odom_flags[s * 4 + 0] = llm::degeneracy_projector<3>(AtA, 10.f, matP) ? 1 : 0;
        if (st.odom_flags[s * 4 + 0]) {
          const float X2[3] = {X[0], X[1], X[2]};
          for (int r = 0; r < 3; ++r) X[r] = matP[r * 3 + 0] * X2[0] + matP[r * 3 + 1] * X2[1] + matP[r * 3 + 2] * X2[2];
        }
        if (surf) {
          sT[0] += X[0]; sT[2] += X[1]; sT[4] += X[2];
        } else {
          sT[1] += X[0]; sT[3] += X[1]; sT[5] += X[2];
        }
        for (int k = 0; k < 6; ++k)
          if (sT[k] != sT[k]) sT[k] = 0;
        const float RAD2DEG = (float)(180.0 / LL_PI);
        float deltaR, deltaT;
        if (surf) {
          const double r0 = (double)(RAD2DEG * X[0]), r1 = (double)(RAD2DEG * X[1]);
          deltaR = (float)sqrt(r0 * r0 + r1 * r1);
          const double t0 = (double)(X[2] * 100);
          deltaT = (float)sqrt(t0 * t0);
        } else {
          const double r0 = (double)(RAD2DEG * X[0]);
          deltaR = (float)sqrt(r0 * r0);
          const double t0 = (double)(X[1] * 100), t1 = (double)(X[2] * 100);
          deltaT = (float)sqrt(t0 * t0 + t1 * t1);
        }
        if ((double)deltaR < 0.1 && (double)deltaT < 0.1) sh_state[0] = 1;
      }
    }
    __syncthreads();
    if (sh_state[0]) break;
  }
  if (threadIdx.x < 6) st.transform_cur[s * 6 + threadIdx.x] = sT[threadIdx.x];
  if (threadIdx.x == 0) {
    st.odom_iters[s * 2 + STAGE] = sh_state[1];
    if (sh_state[0]) st.odom_flags[s * 4 + 1] = 1;
  }
}

__global__ void k_odom_stage_begin(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.p.B) st.odom_flags[s * 4 + 1] = 0;
}

__device__ __forceinline__ void accumulate_rotation(float cx, float cy, float cz, float lx, float ly, float lz,
                                                    float* ox, float* oy, float* oz) {
  // featureAssociation.cpp:474-500
  float slx, clx, sly, cly, slz, clz, scx, ccx, scy, ccy, scz, ccz;
  ll_sincosf(lx, &slx, &clx); ll_sincosf(ly, &sly, &cly); ll_sincosf(lz, &slz, &clz);
  ll_sincosf(cx, &scx, &ccx); ll_sincosf(cy, &scy, &ccy); ll_sincosf(cz, &scz, &ccz);
  const float srx = clx * ccx * sly * scz - ccx * ccz * slx - clx * cly * scx;
  *ox = -ll_asinf(srx);
  const float srycrx = slx * (ccy * scz - ccz * scx * scy) + clx * sly * (ccy * ccz + scx * scy * scz) + clx * cly * ccx * scy;
  const float crycrx = clx * cly * ccx * ccy - clx * sly * (ccz * scy - ccy * scx * scz) - slx * (scy * scz + ccy * ccz * scx);
  const float cox = ll_cosf(*ox);
  *oy = ll_atan2f(srycrx / cox, crycrx / cox);
  const float srzcrx = scx * (clz * sly - cly * slx * slz) + ccx * scz * (cly * clz + slx * sly * slz) + clx * ccx * ccz * slz;
  const float crzcrx = clx * clz * ccx * ccz - ccx * scz * (cly * slz - clz * slx * sly) - scx * (sly * slz + cly * clz * slx);
  *oz = ll_atan2f(srzcrx / cox, crzcrx / cox);
}

__global__ void k_odom_finish(DevState st) {
  // integrateTransformation, featureAssociation.cpp:1241-1270
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B) return;
  float* Tc = st.transform_cur + s * 6;
  float* Ts = st.transform_sum + s * 6;
  float rx, ry, rz;
  accumulate_rotation(Ts[0], Ts[1], Ts[2], -Tc[0], -Tc[1], -Tc[2], &rx, &ry, &rz);
  float srz, crz, srx, crx, sry, cry;
  ll_sincosf(rz, &srz, &crz);
  ll_sincosf(rx, &srx, &crx);
  ll_sincosf(ry, &sry, &cry);
  const float x1 = crz * (Tc[3]) - srz * (Tc[4]);
  const float y1 = srz * (Tc[3]) + crz * (Tc[4]);
  const float z1 = Tc[5];
  const float x2 = x1;
  const float y2 = crx * y1 - srx * z1;
  const float z2 = srx * y1 + crx * z1;
  const float tx = Ts[3] - (cry * x2 + sry * z2);
  const float ty = Ts[4] - y2;
  const float tz = Ts[5] - (-sry * x2 + cry * z2);
  Ts[0] = rx; Ts[1] = ry; Ts[2] = rz;
  Ts[3] = tx; Ts[4] = ty; Ts[5] = tz;
}

__global__ void __launch_bounds__(256) k_publish_clouds_last(DevState st, int first_frame) {
  // first frame: checkSystemInitialization (featureAssociation.cpp:1181-1209): plain hand-over
  // later: publishCloudsLast (featureAssociation.cpp:1329-1363)
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) T[k] = st.transform_cur[s * 6 + k];
  const int n_corner = st.feat_counts[s * 4 + 1];
  const int n_surf = st.feat_counts[s * 4 + 3];
  if (i < n_corner) {
    const float4 q = st.corner_less_sharp[(size_t)s * p.cap_less_sharp + i];
    st.corner_last[(size_t)s * p.cap_less_sharp + i] = first_frame ? q : transform_to_end(q, T);
  }
  if (i < n_surf) {
    const float4 q = st.surf_less_flat[(size_t)s * p.N + i];
    st.surf_last[(size_t)s * p.N + i] = first_frame ? q : transform_to_end(q, T);
  }
  if (!first_frame) {
    const int n_out = st.outlier_count[s];
    if (i < n_out) {  // adjustOutlierCloud, featureAssociation.cpp:1273-1283
      const float4 q = st.outlier_cloud[(size_t)s * st.cap_outlier + i];
      st.outlier_last[(size_t)s * st.cap_outlier + i] = make_float4(q.y, q.z, q.x, q.w);
    }
  }
  if (i == 0) {
    st.last_counts[s * 2 + 0] = n_corner;
    st.last_counts[s * 2 + 1] = n_surf;
    // kd-trees are rebuilt only when both clouds are large enough (featureAssociation.cpp:1356)
    st.odom_flags[s * 4 + 2] = (first_frame || (n_corner > 10 && n_surf > 100)) ? 1 : 0;
    if (!first_frame) st.odom_flags[s * 4 + 3] = st.outlier_count[s];
  }
}

// Break positions of the ring-window scans, per (sequence, cloud): tab[v] = first index whose ring id
// is >= v, tab2[v] = last index whose ring id is <= v (see warp_find_correspondence).
__global__ void __launch_bounds__(256) k_window_tables(DevState st) {
  __shared__ int first_eq[LL_MAX_RINGS + 8], last_eq[LL_MAX_RINGS + 8];
  const DevParams& p = st.p;
  const int s = blockIdx.x, cloud = blockIdx.y;  // cloud 0: corner_last, 1: surf_last
  const int R = LL_MAX_RINGS + 8;
  const float4* pts = cloud ? st.surf_last + (size_t)s * p.N : st.corner_last + (size_t)s * p.cap_less_sharp;
  const int n = st.last_counts[s * 2 + cloud];
  for (int t = threadIdx.x; t < R; t += blockDim.x) { first_eq[t] = 0x7fffffff; last_eq[t] = -1; }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int id = min(max((int)pts[j].w, 0), R - 1);
    // neighbours with the same id make most of these atomics redundant: only run boundaries matter
    const int idp = j > 0 ? min(max((int)pts[j - 1].w, 0), R - 1) : -1;
    const int idn = j + 1 < n ? min(max((int)pts[j + 1].w, 0), R - 1) : -1;
    if (idp != id) atomicMin(&first_eq[id], j);
    if (idn != id) atomicMax(&last_eq[id], j);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int* tab = st.win_tab + ((size_t)s * 2 + cloud) * 2 * R;
    int run = n;  // "no such index": the scan runs to the end of the cloud
    for (int v = R - 1; v >= 0; --v) { run = min(run, first_eq[v]); tab[v] = run; }
    int runl = -1;
    for (int v = 0; v < R; ++v) { runl = max(runl, last_eq[v]); tab[R + v] = runl; }
  }
}

}  // namespace

void launch_odometry(LaunchCtx& ctx, DevState& st, bool first_frame) {
  const DevParams& p = st.p;
  if (!first_frame) {
    const dim3 g_surf(24, p.B), g_corner(24, p.B);  // 24 x 8 warps per sequence, each looping over its feature points
    const int g_seq = (p.B + 63) / 64;
    // surf stage: <= 25 iterations = 5 x (search, 5 LM iterations)   (featureAssociation.cpp:1216-1224)
    LL_LAUNCH(ctx, "k_odom_stage_begin", k_odom_stage_begin<<<g_seq, 64, 0, ctx.stream>>>(st));
    for (int c = 0; c < 5; ++c) {
      LL_LAUNCH(ctx, "k_odom_search_surf", k_odom_search<STAGE_SURF><<<g_surf, 256, 0, ctx.stream>>>(st));
      LL_LAUNCH(ctx, "k_odom_lm_surf", k_odom_lm<STAGE_SURF><<<p.B, LM_THREADS, 0, ctx.stream>>>(st, 5 * c));
    }
    // corner stage (featureAssociation.cpp:1226-1234)
    LL_LAUNCH(ctx, "k_odom_stage_begin", k_odom_stage_begin<<<g_seq, 64, 0, ctx.stream>>>(st));
    for (int c = 0; c < 5; ++c) {
      LL_LAUNCH(ctx, "k_odom_search_corner", k_odom_search<STAGE_CORNER><<<g_corner, 256, 0, ctx.stream>>>(st));
      LL_LAUNCH(ctx, "k_odom_lm_corner", k_odom_lm<STAGE_CORNER><<<p.B, LM_THREADS, 0, ctx.stream>>>(st, 5 * c));
    }
    LL_LAUNCH(ctx, "k_odom_finish", k_odom_finish<<<(p.B + 63) / 64, 64, 0, ctx.stream>>>(st));
  }
  LL_LAUNCH(ctx, "k_publish_clouds_last", k_publish_clouds_last<<<dim3((p.N + 255) / 256, p.B), 256, 0, ctx.stream>>>(st, first_frame ? 1 : 0));
  LL_LAUNCH(ctx, "k_window_tables", k_window_tables<<<dim3(p.B, 2), 256, 0, ctx.stream>>>(st));
  launch_grid_build2(ctx, p.B, st.grid_corner_last, st.corner_last, p.cap_less_sharp, st.last_counts, 2, 0,
                     st.grid_surf_last, st.surf_last, p.N, st.last_counts, 2, 1, st.odom_flags + 2, 4);
}
