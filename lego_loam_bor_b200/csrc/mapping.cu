// mapping.cu -- MapOptimization::scan2MapOptimization (reference: LeGO-LOAM/src/mapOptmization.cpp:
// 397-426 pointAssociateToMap, 1028-1134 cornerOptimization, 1136-1197 surfOptimization,
// 1199-1312 LMOptimization, 1315-1332 scan2MapOptimization) and downsampleCurrentScan (:999-1026).
//
//   grid build (hashgrid.cu)  replaces the two kd-tree builds of every mapping cycle (:1317-1318);
//       cell = 1 m = the acceptance radius of :1036,1144, so a 27-cell lookup is exact.
//   k_map_knn    one thread per down-sampled scan point: pointAssociateToMap + exact 5-NN with provable reuse
//   k_map_iter   one thread per down-sampled scan point: line fit
//       (3x3 symmetric eigen-solve) or plane fit (5x3 column-pivoted QR), residual + 6-column
//       Jacobian row, J^T J / J^T r partial sums (exact double products) per block.
//   k_map_solve  one block per sequence: fixed-order sum of the partials, 6x6 solve, degeneracy
//       projection, pose update and convergence flag.  Ten (knn, iter, solve) triples are enqueued; a
//       converged sequence makes its remaining launches no-ops, so there is no host round trip.
//   k_voxel_*    pcl::VoxelGrid on whole clouds for downsampleCurrentScan (one block per sequence
//       and cloud: stable LSD radix sort by voxel index in shared/global memory, sequential centroids).
#include "../../include/ll_smallmat.h"
#include "hashgrid.cuh"
#include "ll_kernels.h"

namespace {

#define MAP_BLOCKS 96
#define KNN_BLOCKS 112
#define MAP_THREADS 128
#define MAP_USE_FIT_CACHE 1
#define MAP_NACC 28  // 21 upper-triangle J^T J + 6 J^T r + row count

__device__ __forceinline__ bool map_guard(const DevState& st, int s) {
  return st.map_counts[s * 2 + 0] > 10 && st.map_counts[s * 2 + 1] > 100;  // mapOptmization.cpp:1316
}

struct MapPose {
  float cRoll, sRoll, cPitch, sPitch, cYaw, sYaw, tX, tY, tZ;
};

__device__ __forceinline__ MapPose make_map_pose(const float* T) {  // mapOptmization.cpp:397-410
  MapPose m;
  ll_sincosf(T[0], &m.sRoll, &m.cRoll);
  ll_sincosf(T[1], &m.sPitch, &m.cPitch);
  ll_sincosf(T[2], &m.sYaw, &m.cYaw);
  m.tX = T[3]; m.tY = T[4]; m.tZ = T[5];
  return m;
}

// updatePointAssociateToMapSinCos once per block: one warp-0 thread evaluates the six values, everybody reads them
__device__ __forceinline__ MapPose block_map_pose(const float* T_global, MapPose* sh) {
  if (threadIdx.x == 0) {
    float T[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) T[k] = T_global[k];
    *sh = make_map_pose(T);
  }
  __syncthreads();
  return *sh;
}

__device__ __forceinline__ float4 point_associate_to_map(const MapPose& m, const float4 pi) {  // :412-426
  const float x1 = m.cYaw * pi.x - m.sYaw * pi.y;
  const float y1 = m.sYaw * pi.x + m.cYaw * pi.y;
  const float z1 = pi.z;
  const float x2 = x1;
  const float y2 = m.cRoll * y1 - m.sRoll * z1;
  const float z2 = m.sRoll * y1 + m.cRoll * z1;
  return make_float4(m.cPitch * x2 + m.sPitch * z2 + m.tX, y2 + m.tY, -m.sPitch * x2 + m.cPitch * z2 + m.tZ, pi.w);
}

// cornerOptimization / surfOptimization are split in two: the part that only depends on the five neighbours (the line
// through the cluster / the plane fit and its validity) and the part that depends on the transformed query point.
// Between LM iterations the ordered neighbour set of most points does not change, so the first part is cached.

// cornerOptimization, neighbours only (mapOptmization.cpp:1037-1091): g0 = (x1, y1, z1, valid), g1 = (x2, y2, z2, -)
__device__ __forceinline__ void corner_geom(const float4* q, float4* g0, float4* g1) {
  float cx = 0, cy = 0, cz = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) { cx += q[j].x; cy += q[j].y; cz += q[j].z; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const float ax = q[j].x - cx, ay = q[j].y - cy, az = q[j].z - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
    a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  const float matA1[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
  float matD1[3], matV1[9];
  llm::self_adjoint_eigen<3>(matA1, matD1, matV1);
  const bool valid = matD1[2] > 3 * matD1[1];
  // row 0 of the eigenvector matrix (sic), mapOptmization.cpp:1086-1091; 0.1 * v in double
  *g0 = make_float4((float)((double)cx + 0.1 * (double)matV1[0]), (float)((double)cy + 0.1 * (double)matV1[1]),
                    (float)((double)cz + 0.1 * (double)matV1[2]), valid ? 1.f : 0.f);
  *g1 = make_float4((float)((double)cx - 0.1 * (double)matV1[0]), (float)((double)cy - 0.1 * (double)matV1[1]),
                    (float)((double)cz - 0.1 * (double)matV1[2]), 0.f);
}

// cornerOptimization, query-dependent part (mapOptmization.cpp:1093-1131); returns false if rejected
__device__ __forceinline__ bool corner_eval(const float4 sel, const float4 g0, const float4 g1, float4* coeff) {
  if (g0.w == 0.f) return false;
  const float x0 = sel.x, y0 = sel.y, z0 = sel.z;
  const float x1 = g0.x, y1 = g0.y, z1 = g0.z;
  const float x2 = g1.x, y2 = g1.y, z2 = g1.z;
  const float a012 = sqrtf(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                           ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                           ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
  const float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  const float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                    (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
  const float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) -
                     (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  const float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                     (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  const float ld2 = a012 / l12;
  const float w = (float)(1 - 0.9 * (double)fabsf(ld2));
  *coeff = make_float4(w * la, w * lb, w * lc, w * ld2);
  return (double)w > 0.1;
}

// surfOptimization, neighbours only (mapOptmization.cpp:1145-1175): g0 = unit plane (pa, pb, pc, pd), g1.x = planeValid
__device__ __forceinline__ void surf_geom(const float4* q, float4* g0, float4* g1) {
  float matA0[15], matX0[3];
  const float matB0[5] = {-1, -1, -1, -1, -1};
#pragma unroll
  for (int j = 0; j < 5; j++) { matA0[j * 3 + 0] = q[j].x; matA0[j * 3 + 1] = q[j].y; matA0[j * 3 + 2] = q[j].z; }
  llm::colpiv_qr_solve<5, 3>(matA0, matB0, matX0);
  float pa = matX0[0], pb = matX0[1], pc = matX0[2], pd = 1;
  const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  bool planeValid = true;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    if ((double)fabsf(pa * q[j].x + pb * q[j].y + pc * q[j].z + pd) > 0.2) { planeValid = false; break; }
  }
  *g0 = make_float4(pa, pb, pc, pd);
  *g1 = make_float4(planeValid ? 1.f : 0.f, 0.f, 0.f, 0.f);
}

// surfOptimization, query-dependent part (mapOptmization.cpp:1177-1194)
__device__ __forceinline__ bool surf_eval(const float4 sel, const float4 g0, const float4 g1, float4* coeff) {
  if (g1.x == 0.f) return false;
  const float pa = g0.x, pb = g0.y, pc = g0.z, pd = g0.w;
  const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
  const float w = (float)(1 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
  *coeff = make_float4(w * pa, w * pb, w * pc, w * pd2);
  return (double)w > 0.1;
}

// Exact 5-NN of every down-sampled scan point in the local map, one THREAD per query.
//
// Per query a record of KNN_K + 1 float4 is kept: the KNN_K nearest map points of the last full search with their
// coordinates (x, y, z, index), ascending by (d2, index), and (p0, bound): the query position of that search and the
// squared distance to the (KNN_K+1)-th nearest map point there (or 1 m, the reach of the 27-cell block).
// In later LM iterations the query has moved by |delta| since p0, so every map point outside the record is at least
// sqrt(bound) - |delta| away; if the five nearest of the RECORDED points are closer than that (minus a float-safety
// margin) they are the exact 5-NN (triangle inequality) and no search is needed.  Their positions inside the record,
// in ascending (d2, index) order, go to the selection word k_map_iter reads (5 x 4 bits, bit 31 = all five closer
// than 1 m, mapOptmization.cpp:1036,1144).
// The full search visits the 27 buckets around the query (occupancy bitmap first) and keeps a sorted top-(KNN_K+1)
// in registers; consecutive queries are consecutive voxels of the VoxelGrid output, so neighbouring threads walk the
// same buckets and share them through L1.
#define KNN_THREADS 128
#define KNN_K LL_KNN_K
#define KNN_REC (KNN_K + 1)

// The 27 cells around a query, nearest first (centre, 6 faces, 12 edges, 8 corners).  The visiting order does not
// change the result of the exact search, but the nearest points now arrive first, the acceptance threshold drops at once
// and few later candidates pay for a sorted insertion.  n is a compile-time constant at every call site (unrolled loops).
__device__ __forceinline__ void knn_cell_offset(int n, int* dx, int* dy, int* dz) {
  switch (n) {
    case 0: *dx = 0; *dy = 0; *dz = 0; break;
    case 1: *dx = 0; *dy = 0; *dz = -1; break;
    case 2: *dx = 0; *dy = -1; *dz = 0; break;
    case 3: *dx = -1; *dy = 0; *dz = 0; break;
    case 4: *dx = 1; *dy = 0; *dz = 0; break;
    case 5: *dx = 0; *dy = 1; *dz = 0; break;
    case 6: *dx = 0; *dy = 0; *dz = 1; break;
    case 7: *dx = 0; *dy = -1; *dz = -1; break;
    case 8: *dx = -1; *dy = 0; *dz = -1; break;
    case 9: *dx = 1; *dy = 0; *dz = -1; break;
    case 10: *dx = 0; *dy = 1; *dz = -1; break;
    case 11: *dx = -1; *dy = -1; *dz = 0; break;
    case 12: *dx = 1; *dy = -1; *dz = 0; break;
    case 13: *dx = -1; *dy = 1; *dz = 0; break;
    case 14: *dx = 1; *dy = 1; *dz = 0; break;
    case 15: *dx = 0; *dy = -1; *dz = 1; break;
    case 16: *dx = -1; *dy = 0; *dz = 1; break;
    case 17: *dx = 1; *dy = 0; *dz = 1; break;
    case 18: *dx = 0; *dy = 1; *dz = 1; break;
    case 19: *dx = -1; *dy = -1; *dz = -1; break;
    case 20: *dx = 1; *dy = -1; *dz = -1; break;
    case 21: *dx = -1; *dy = 1; *dz = -1; break;
    case 22: *dx = 1; *dy = 1; *dz = -1; break;
    case 23: *dx = -1; *dy = -1; *dz = 1; break;
    case 24: *dx = 1; *dy = -1; *dz = 1; break;
    case 25: *dx = -1; *dy = 1; *dz = 1; break;
    case 26: *dx = 1; *dy = 1; *dz = 1; break;
    default: *dx = 0; *dy = 0; *dz = 0; break;
  }
}

__device__ __forceinline__ bool cand_less(float d2a, int ia, float d2b, int ib) { return d2a < d2b || (d2a == d2b && ia < ib); }

// Full search of ONE query by a whole warp, for the launches in which only a few queries of a block fail the reuse test
// (LM iterations >= 2): a single thread walks its 27 cells as a chain of dependent loads (tens of microseconds) while the
// rest of its block waits at the barrier; here lanes 0..26 take one cell each (occupancy word, then bucket bounds), the
// points of all non-empty buckets are spread over the lanes, and every lane keeps the same sorted top-(KNN_K+1) in
// registers: a candidate below the current threshold is broadcast and inserted by all lanes alike.  Same candidates
// (points of the 27 cells closer than 1 m), same (d2, index) order, so the record is the one the thread search writes
// (up to the choice among exactly equal distances at the last recorded position, which either way keeps the record valid).
#ifndef KNN_COOP_MAX
#define KNN_COOP_MAX 24
#endif
#ifndef KNN_SEED_PREV
#define KNN_SEED_PREV 0
#endif

__device__ __forceinline__ void knn_search_warp(const DevState& st, int s, int q, int nc, const MapPose& mp, int* sh /* [3][32] */) {
  const DevParams& p = st.p;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool corner = q < nc;
  const float4 ori = corner ? st.scan_corner_ds[(size_t)s * p.cap_less_sharp + q] : st.scan_surf_ds[(size_t)s * p.N + (q - nc)];
  const float4 sel = point_associate_to_map(mp, ori);
  const HashGrid& g = corner ? st.grid_map_corner : st.grid_map_surf;
  const int* cs = g.cell_start + (size_t)s * grid_cs_stride(g);
  const unsigned* occ = g.occ + (size_t)s * (g.tbl / 32);
  const float4* pts = g.sorted + (size_t)s * g.cap;
  const int cx = grid_cell(sel.x, g.inv_cell), cy = grid_cell(sel.y, g.inv_cell), cz = grid_cell(sel.z, g.inv_cell);
  const float fx = sel.x - (float)cx * g.cell, fy = sel.y - (float)cy * g.cell, fz = sel.z - (float)cz * g.cell;
  // one cell per lane
  uint32_t hh = 0x80000000u | (uint32_t)lane;  // lanes without a bucket keep distinct values for the match below
  bool occd = false;
  if (lane < 27) {
    const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
    const float gx = dx < 0 ? fx : (dx > 0 ? g.cell - fx : 0.f);
    const float gy = dy < 0 ? fy : (dy > 0 ? g.cell - fy : 0.f);
    const float gz = dz < 0 ? fz : (dz > 0 ? g.cell - fz : 0.f);
    // a cell whose nearest corner is not closer than 1 m holds no candidate (0.999: never skip on a rounding error)
    if ((dx == 0 && dy == 0 && dz == 0) || (gx * gx + gy * gy + gz * gz) * 0.999f <= 1.0f) {
      const uint32_t h = grid_hash(cx + dx, cy + dy, cz + dz, g.tbl);
      if ((occ[h >> 5] >> (h & 31)) & 1u) { occd = true; hh = h; }
    }
  }
  // two of the 27 cells can share a bucket (hash collision): the lowest lane keeps it
  const unsigned same = __match_any_sync(full, hh);
  if (occd && (__ffs(same) - 1) != lane) occd = false;
  int b0 = 0, cnt = 0;
  if (occd) { b0 = cs[hh]; cnt = cs[hh + 1] - b0; }
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(full, inc, o);
    if (lane >= o) inc += x;
  }
  const int total = __shfl_sync(full, inc, 31);
  __syncwarp();
  sh[lane] = inc - cnt;  // first flat position of this lane's bucket
  sh[32 + lane] = inc;   // one past its last
  sh[64 + lane] = b0;
  __syncwarp();
  unsigned long long bk[KNN_REC];
#pragma unroll
  for (int i = 0; i < KNN_REC; ++i) bk[i] = ((unsigned long long)__float_as_uint(1.0f) << 32) | 0x7fffffffull;
  // flat position j -> point: the bucket whose [first, last) holds j
  auto fetch = [&](int j, float4* out) -> bool {
    if (j >= total) return false;
    int L = 0;
    while (sh[32 + L] <= j) ++L;
    *out = pts[sh[64 + L] + (j - sh[L])];
    return true;
  };
  float4 nxt;
  bool have_nxt = fetch(lane, &nxt);
  for (int r0 = 0; r0 < total; r0 += 32) {
    const float4 cur = nxt;
    const bool have = have_nxt;
    have_nxt = fetch(r0 + 32 + lane, &nxt);  // in flight while this round is merged
    unsigned long long key = ~0ull;
    bool cand = false;
    if (have) {
      const float cd = nn_dist2(sel.x, sel.y, sel.z, cur);
      key = ((unsigned long long)__float_as_uint(cd) << 32) | (unsigned)__float_as_int(cur.w);
      cand = cd < 1.0f && key < bk[KNN_K];
    }
    unsigned bal = __ballot_sync(full, cand);
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      unsigned long long k2 = __shfl_sync(full, key, src);
      if (k2 < bk[KNN_K]) {
#pragma unroll
        for (int i = 0; i < KNN_REC; ++i) {
          const unsigned long long b = bk[i];
          const bool lt = k2 < b;
          bk[i] = lt ? k2 : b;
          k2 = lt ? b : k2;
        }
      }
    }
  }
  // record: lane i writes candidate i, lane KNN_K the query position + the (KNN_K+1)-th distance
  const float4* mpts = corner ? st.map_corner + (size_t)s * st.cap_map_corner : st.map_surf + (size_t)s * st.cap_map_surf;
  float4* rec = st.map_knn_rec + ((size_t)s * st.map_knn_cap + q) * KNN_REC;
  unsigned long long mine = 0ull;
#pragma unroll
  for (int i = 0; i < KNN_REC; ++i) mine = (lane == i) ? bk[i] : mine;
  if (lane < KNN_K) {
    const int bi = (int)(unsigned)mine;
    float4 c = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (bi != 0x7fffffff) { c = mpts[bi]; c.w = __int_as_float(bi); }
    rec[lane] = c;
  } else if (lane == KNN_K) {
    rec[KNN_K] = make_float4(sel.x, sel.y, sel.z, __uint_as_float((unsigned)(mine >> 32)));
    st.map_knn_sel[(size_t)s * st.map_knn_cap + q] = (int)(((int)(unsigned)bk[4] != 0x7fffffff ? 0x80000000u : 0u) | 0x43210u);
  }
}

#ifndef KNN_MIN_BLOCKS
#define KNN_MIN_BLOCKS 5   // 6 (80 registers) and 8 (64, spills) measured alike
#endif
__global__ void __launch_bounds__(KNN_THREADS, KNN_MIN_BLOCKS) k_map_knn(DevState st, int iter) {
  __shared__ int sh_need[KNN_THREADS];
  __shared__ int sh_coop[KNN_THREADS / 32][96];
  __shared__ int sh_n;
  __shared__ int2 sh_bucket[9][KNN_THREADS];  // per thread: the non-empty buckets (start, end) of the current batch of nine cells, column-major: no bank conflicts
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  if (!map_guard(st, s) || st.map_flags[s * 4 + 1]) return;
  const int nc = st.scan_ds_counts[s * 2 + 0], ns = st.scan_ds_counts[s * 2 + 1];
  const int nq = nc + ns;
  __shared__ MapPose sh_pose;
  const MapPose mp = block_map_pose(st.transform_tobe_mapped + s * 6, &sh_pose);
  for (int qbase = blockIdx.x * KNN_THREADS; qbase < nq; qbase += gridDim.x * KNN_THREADS) {
    if (threadIdx.x == 0) sh_n = 0;
    __syncthreads();
    // ---- phase 1: one thread per query, try to select the five nearest among the recorded candidates ----
    {
      const int q = qbase + threadIdx.x;
      if (q < nq) {
        bool need = true;
        if (iter > 0) {
          const bool corner = q < nc;
          const float4 ori = corner ? st.scan_corner_ds[(size_t)s * p.cap_less_sharp + q] : st.scan_surf_ds[(size_t)s * p.N + (q - nc)];
          const float4 sel = point_associate_to_map(mp, ori);
          const float4* rec = st.map_knn_rec + ((size_t)s * st.map_knn_cap + q) * KNN_REC;
          const float4 stt = rec[KNN_K];
          float d5[5];
          int i5[5], p5[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) { d5[i] = FLT_MAX; i5[i] = 0x7fffffff; p5[i] = 0; }
#pragma unroll
          for (int c = 0; c < KNN_K; ++c) {
            const float4 cp = rec[c];
            const int ci0 = __float_as_int(cp.w);
            if (ci0 < 0) continue;  // fewer than KNN_K points within reach at p0
            float cd = nn_dist2(sel.x, sel.y, sel.z, cp);
            int ci = ci0, cpos = c;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              if (cand_less(cd, ci, d5[i], i5[i])) {
                const float td = d5[i]; const int ti = i5[i], tp = p5[i];
                d5[i] = cd; i5[i] = ci; p5[i] = cpos;
                cd = td; ci = ti; cpos = tp;
              }
            }
          }
          if (i5[4] != 0x7fffffff) {
            const float dx = sel.x - stt.x, dy = sel.y - stt.y, dz = sel.z - stt.z;
            const float delta = sqrtf(dx * dx + dy * dy + dz * dz);
            if (sqrtf(d5[4]) + delta + 1e-4f < sqrtf(stt.w)) {
              // d5[4] < bound <= 1: all five are inside the acceptance radius
              const unsigned neww = 0x80000000u | (unsigned)p5[0] | ((unsigned)p5[1] << 4) | ((unsigned)p5[2] << 8) | ((unsigned)p5[3] << 12) | ((unsigned)p5[4] << 16);
              int* selp = st.map_knn_sel + (size_t)s * st.map_knn_cap + q;
              // bit 30: the ordered neighbours are the ones k_map_iter last fitted, its cached line / plane still holds
              const unsigned oldw = (unsigned)*selp;
              *selp = (int)(((oldw & 0xbfffffffu) == neww) ? (neww | 0x40000000u) : neww);
              need = false;
            }
          }
        }
        if (need) sh_need[atomicAdd(&sh_n, 1)] = q;
      }
    }
    __syncthreads();
    // ---- phase 2: full search of the remaining queries, packed onto consecutive threads ----
    const int n_need = sh_n;
    if (n_need <= KNN_COOP_MAX) {
      // few stragglers: a warp per query
      for (int w = threadIdx.x >> 5; w < n_need; w += KNN_THREADS / 32) knn_search_warp(st, s, sh_need[w], nc, mp, sh_coop[threadIdx.x >> 5]);
    } else
    for (int w = threadIdx.x; w < n_need; w += KNN_THREADS) {
      const int q = sh_need[w];
      const bool corner = q < nc;
      const float4 ori = corner ? st.scan_corner_ds[(size_t)s * p.cap_less_sharp + q] : st.scan_surf_ds[(size_t)s * p.N + (q - nc)];
      const float4 sel = point_associate_to_map(mp, ori);
      const HashGrid& g = corner ? st.grid_map_corner : st.grid_map_surf;
      const int* cs = g.cell_start + (size_t)s * grid_cs_stride(g);
      const unsigned* occ = g.occ + (size_t)s * (g.tbl / 32);
      const float4* pts = g.sorted + (size_t)s * g.cap;
      const int cx = grid_cell(sel.x, g.inv_cell), cy = grid_cell(sel.y, g.inv_cell), cz = grid_cell(sel.z, g.inv_cell);
      // sorted top-(KNN_K+1) as 64-bit keys (d2 bits << 32 | index): for the non-negative d2 an unsigned compare of
      // the keys is the lexicographic (d2, index) order of cand_less, and an insertion step is one compare + selects
      unsigned long long bk[KNN_REC];
#pragma unroll
      for (int i = 0; i < KNN_REC; ++i) bk[i] = ((unsigned long long)__float_as_uint(1.0f) << 32) | 0x7fffffffull;
#if KNN_SEED_PREV
      // Experiment for the next round (off by default, not yet measured on the GPU): a repeated search starts from the
      // previous record's candidates.  They are map points closer than 1 m if their new distance says so, hence inside the
      // 27 cells and found again below (skipped there by the duplicate test): the result is the same, but the threshold is
      // tight from the first cell on, so far fewer candidates pay for a sorted insertion and more cells are pruned.
      if (iter > 0) {
        const float4* prev = st.map_knn_rec + ((size_t)s * st.map_knn_cap + q) * KNN_REC;
#pragma unroll
        for (int c = 0; c < KNN_K; ++c) {
          const float4 cp = prev[c];
          const int ci0 = __float_as_int(cp.w);
          const float cd = nn_dist2(sel.x, sel.y, sel.z, cp);
          if (ci0 >= 0 && cd < 1.0f) {
            unsigned long long key = ((unsigned long long)__float_as_uint(cd) << 32) | (unsigned)ci0;
#pragma unroll
            for (int i = 0; i < KNN_REC; ++i) {
              const unsigned long long b = bk[i];
              const bool lt = key < b;
              bk[i] = lt ? key : b;
              key = lt ? b : key;
            }
          }
        }
      }
#endif
      // position of the query inside its cell: the gap to a neighbouring cell along an axis is f or cell - f
      const float fx = sel.x - (float)cx * g.cell, fy = sel.y - (float)cy * g.cell, fz = sel.z - (float)cz * g.cell;
      // Three batches of nine cells, nearest first.  2a: the non-empty buckets of the batch; the bitmap loads (and then the
      // cell_start loads of the occupied ones) are independent and in flight together: this kernel is bound by memory
      // latency.  From the second batch on, a cell whose nearest corner is farther than the current (KNN_K+1)-th best
      // cannot contribute (the threshold only drops), so it is skipped without touching the bitmap.
#pragma unroll
      for (int z = 0; z < 3; ++z) {
        int nb = 0;
        uint32_t hh[9];
        unsigned occw[9];
        const float thr = __uint_as_float((unsigned)(bk[KNN_K] >> 32));
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          int dx, dy, dz;
          knn_cell_offset(z * 9 + t, &dx, &dy, &dz);
          const float gx = dx < 0 ? fx : (dx > 0 ? g.cell - fx : 0.f);
          const float gy = dy < 0 ? fy : (dy > 0 ? g.cell - fy : 0.f);
          const float gz = dz < 0 ? fz : (dz > 0 ? g.cell - fz : 0.f);
          // 0.999: the gaps and the candidate distances are rounded differently; never skip a cell on a rounding error
          const bool reach = z == 0 || (gx * gx + gy * gy + gz * gz) * 0.999f <= thr;
          hh[t] = grid_hash(cx + dx, cy + dy, cz + dz, g.tbl);
          occw[t] = reach ? occ[hh[t] >> 5] : 0u;
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          if ((occw[t] >> (hh[t] & 31)) & 1u) sh_bucket[nb++][threadIdx.x] = make_int2(cs[hh[t]], cs[hh[t] + 1]);
        }
        // 2b: candidates of the batch's buckets, four independent loads per step.  A point is a candidate if it is
        // closer than 1 m (such a point necessarily lies in one of the 27 cells); two of the 27 cells can share a
        // bucket (hash collision), so an index that is already in the list is skipped.
        for (int u = 0; u < nb; ++u) {
          const int2 be = sh_bucket[u][threadIdx.x];
          for (int k = be.x; k < be.y; k += 4) {
            float4 cpt[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) cpt[v] = pts[min(k + v, be.y - 1)];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (k + v >= be.y) continue;
              const float cd = nn_dist2(sel.x, sel.y, sel.z, cpt[v]);
              if (cd < __uint_as_float((unsigned)(bk[KNN_K] >> 32))) {
                const unsigned ci = (unsigned)__float_as_int(cpt[v].w);
                bool dup = false;
#pragma unroll
                for (int i = 0; i < KNN_REC; ++i) dup = dup || ((unsigned)bk[i] == ci);
                if (!dup) {
                  unsigned long long key = ((unsigned long long)__float_as_uint(cd) << 32) | ci;
#pragma unroll
                  for (int i = 0; i < KNN_REC; ++i) {
                    const unsigned long long b = bk[i];
                    const bool lt = key < b;
                    bk[i] = lt ? key : b;
                    key = lt ? b : key;
                  }
                }
              }
            }
          }
        }
      }
      float bd[KNN_REC];
      int bi[KNN_REC];
#pragma unroll
      for (int i = 0; i < KNN_REC; ++i) { bd[i] = __uint_as_float((unsigned)(bk[i] >> 32)); bi[i] = (int)(unsigned)bk[i]; }
      const float4* mpts = corner ? st.map_corner + (size_t)s * st.cap_map_corner : st.map_surf + (size_t)s * st.cap_map_surf;
      float4* rec = st.map_knn_rec + ((size_t)s * st.map_knn_cap + q) * KNN_REC;
#pragma unroll
      for (int i = 0; i < KNN_K; ++i) {
        float4 c = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (bi[i] != 0x7fffffff) { c = mpts[bi[i]]; c.w = __int_as_float(bi[i]); }
        rec[i] = c;
      }
      rec[KNN_K] = make_float4(sel.x, sel.y, sel.z, bd[KNN_K]);  // bd[KNN_K] = min((KNN_K+1)-th nearest d2, 1)
      // a fresh record is already in ascending order
      st.map_knn_sel[(size_t)s * st.map_knn_cap + q] = (int)((bi[4] != 0x7fffffff ? 0x80000000u : 0u) | 0x43210u);
    }
    __syncthreads();
  }
}

// parity aid (ll_enable_index_trace): the ordered 5-NN indices every query has after k_map_knn of this iteration
__global__ void __launch_bounds__(256) k_map_knn_trace(DevState st, int iter) {
  const int s = blockIdx.y;
  if (!map_guard(st, s) || st.map_flags[s * 4 + 1]) return;
  const int nq = st.scan_ds_counts[s * 2 + 0] + st.scan_ds_counts[s * 2 + 1];
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const unsigned w = (unsigned)st.map_knn_sel[(size_t)s * st.map_knn_cap + q];
    const float4* rec = st.map_knn_rec + ((size_t)s * st.map_knn_cap + q) * KNN_REC;
    int* out = st.map_knn_trace + (((size_t)s * 10 + iter) * st.map_knn_cap + q) * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) out[j] = (w & 0x80000000u) ? __float_as_int(rec[(w >> (4 * j)) & 15u].w) : -1;
  }
}

__device__ __noinline__ void map_solve_body(float* Tm, int* map_iters, double* trace, float* matP, int* map_flags, int iter, const double* tot);

__global__ void __launch_bounds__(MAP_THREADS) k_map_iter(DevState st, int iter) {
  __shared__ double sh_part[MAP_THREADS / 32][MAP_NACC];
  __shared__ double sh_tot[MAP_NACC];
  __shared__ int sh_last;
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  double* out = st.map_partials + ((size_t)s * MAP_BLOCKS + blockIdx.x) * MAP_NACC;
  if (!map_guard(st, s) || st.map_flags[s * 4 + 1]) return;
  __shared__ MapPose sh_pose;
  const MapPose mp = block_map_pose(st.transform_tobe_mapped + s * 6, &sh_pose);
  const float srx = mp.sRoll, crx = mp.cRoll, sry = mp.sPitch, cry = mp.cPitch, srz = mp.sYaw, crz = mp.cYaw;
  const int nc = st.scan_ds_counts[s * 2 + 0], ns = st.scan_ds_counts[s * 2 + 1];
  double acc[MAP_NACC];
#pragma unroll
  for (int k = 0; k < MAP_NACC; ++k) acc[k] = 0.0;
  for (int q = blockIdx.x * MAP_THREADS + threadIdx.x; q < nc + ns; q += MAP_BLOCKS * MAP_THREADS) {
    const bool corner = q < nc;
    const float4 ori = corner ? st.scan_corner_ds[(size_t)s * p.cap_less_sharp + q] : st.scan_surf_ds[(size_t)s * p.N + (q - nc)];
    const float4 sel = point_associate_to_map(mp, ori);
    float4 cf;
    const unsigned selw = (unsigned)st.map_knn_sel[(size_t)s * st.map_knn_cap + q];
    if (!(selw >> 31)) continue;  // fewer than five map points within 1 m (pointSearchSqDis[4] < 1.0 fails)
    float4* fit = st.map_fit + ((size_t)s * st.map_knn_cap + q) * 2;
    float4 g0, g1;
    if ((selw & 0x40000000u) && MAP_USE_FIT_CACHE) {
      g0 = fit[0]; g1 = fit[1];
    } else {
      float4 nb[5];
      const float4* rec = st.map_knn_rec + ((size_t)s * st.map_knn_cap + q) * KNN_REC;
#pragma unroll
      for (int i = 0; i < 5; ++i) nb[i] = rec[(selw >> (4 * i)) & 15u];
      if (corner) corner_geom(nb, &g0, &g1); else surf_geom(nb, &g0, &g1);
      fit[0] = g0; fit[1] = g1;
    }
    const bool ok = corner ? corner_eval(sel, g0, g1, &cf) : surf_eval(sel, g0, g1, &cf);
    if (!ok) continue;
    // mapOptmization.cpp:1223-1255
    const float arx = (crx * sry * srz * ori.x + crx * crz * sry * ori.y - srx * sry * ori.z) * cf.x +
                      (-srx * srz * ori.x - crz * srx * ori.y - crx * ori.z) * cf.y +
                      (crx * cry * srz * ori.x + crx * cry * crz * ori.y - cry * srx * ori.z) * cf.z;
    const float ary = ((cry * srx * srz - crz * sry) * ori.x + (sry * srz + cry * crz * srx) * ori.y + crx * cry * ori.z) * cf.x +
                      ((-cry * crz - srx * sry * srz) * ori.x + (cry * srz - crz * srx * sry) * ori.y - crx * sry * ori.z) * cf.z;
    const float arz = ((crz * srx * sry - cry * srz) * ori.x + (-cry * crz - srx * sry * srz) * ori.y) * cf.x +
                      (crx * crz * ori.x - crx * srz * ori.y) * cf.y +
                      ((sry * srz + cry * crz * srx) * ori.x + (crz * sry - cry * srx * srz) * ori.y) * cf.z;
    const double a[6] = {arx, ary, arz, cf.x, cf.y, cf.z};
    const double b = (double)(-cf.w);
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) acc[k++] += a[r] * a[c];
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[21 + r] += a[r] * b;
    acc[27] += 1.0;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < MAP_NACC; ++k) {
    const double v = warp_sum_d(acc[k]);
    if (lane == 0) sh_part[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < MAP_NACC) {
    double v = 0.0;
    for (int w = 0; w < MAP_THREADS / 32; ++w) v += sh_part[w][threadIdx.x];
    out[threadIdx.x] = v;
  }
  // The block that delivers the last partial sums of its sequence finishes the iteration: fixed-order sum of the
  // partials, then LMOptimization's solve / pose update / convergence flag (no separate launch).
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sh_last = (atomicAdd(st.map_ticket + s, 1) == MAP_BLOCKS - 1) ? 1 : 0;
  __syncthreads();
  if (!sh_last) return;
  __threadfence();
  if (threadIdx.x < MAP_NACC) {
    const double* part = st.map_partials + (size_t)s * MAP_BLOCKS * MAP_NACC;
    double v = 0.0;
    for (int b = 0; b < MAP_BLOCKS; ++b) v += __ldcg(part + b * MAP_NACC + threadIdx.x);
    sh_tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  st.map_ticket[s] = 0;
  double tot[MAP_NACC];
  for (int i = 0; i < MAP_NACC; ++i) tot[i] = sh_tot[i];
  map_solve_body(st.transform_tobe_mapped + s * 6, st.map_iters + s * 2, st.map_trace + ((size_t)s * 10 + iter) * 34,
                 st.map_matP + s * 36, st.map_flags + s * 4, iter, tot);
}

// LMOptimization after the normal equations are summed (mapOptmization.cpp:1257-1312): 6x6 solve, degeneracy projection,
// pose update, convergence flag.  One thread.  Kept out of line and fed from a private copy of the totals: nvcc 12.9 -O3
// miscompiles this code when it is inlined next to code that produces the totals in shared memory (see ll_smallmat.h).
__device__ __noinline__ void map_solve_body(float* Tm, int* map_iters, double* trace, float* matP, int* map_flags, int iter, const double* tot) {
  const int rows = (int)tot[27];
  map_iters[0] = iter + 1;
  map_iters[1] = rows;
  for (int i = 0; i < 28; ++i) trace[i] = tot[i];
  for (int i = 28; i < 34; ++i) trace[i] = 0.0;
  if (rows < 50) return;  // LMOptimization returns false: keep iterating (mapOptmization.cpp:1208-1210)
  float AtA[36], AtB[6], A2[36], X[6];
  int k = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) { AtA[r * 6 + c] = AtA[c * 6 + r] = (float)tot[k]; ++k; }
  for (int r = 0; r < 6; ++r) AtB[r] = (float)tot[21 + r];
  for (int i = 0; i < 36; ++i) A2[i] = AtA[i];
  llm::colpiv_qr_solve<6, 6>(A2, AtB, X);
  if (iter == 0) map_flags[0] = llm::certainly_not_degenerate<6>(AtA, 100.f) ? 0 : (llm::degeneracy_projector<6>(AtA, 100.f, matP) ? 1 : 0);
  if (map_flags[0]) {
    float X2[6];
    for (int i = 0; i < 6; ++i) X2[i] = X[i];
    for (int r = 0; r < 6; ++r) {
      float v = 0.f;
      for (int c = 0; c < 6; ++c) v += matP[r * 6 + c] * X2[c];
      X[r] = v;
    }
  }
  for (int i = 0; i < 6; ++i) Tm[i] += X[i];
  for (int i = 0; i < 6; ++i) trace[28 + i] = (double)X[i];
  const float r2d = 57.29578f;  // pcl::rad2deg(float)
  const double r0 = (double)(X[0] * r2d), r1 = (double)(X[1] * r2d), r2 = (double)(X[2] * r2d);
  const double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
  const float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
  const float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
  if ((double)deltaR < 0.05 && (double)deltaT < 0.05) map_flags[1] = 1;
}

__global__ void k_map_begin(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B) return;
  st.map_flags[s * 4 + 1] = 0;
  st.map_ticket[s] = 0;
  st.map_iters[s * 2 + 0] = 0;
  st.map_iters[s * 2 + 1] = 0;
}

// MapOptimization::transformAssociateToMap (mapOptmization.cpp:264-387), one thread per sequence.
__global__ void k_map_associate(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B) return;
  const float* tS = st.map_odom + s * 6;  // the odometry of the handed-over scan, not FeatureAssociation's live pose
  const float* tB = st.transform_bef_mapped + s * 6;
  const float* tA = st.transform_aft_mapped + s * 6;
  float* tT = st.transform_tobe_mapped + s * 6;
  float sbcx, cbcx, sbcy, cbcy, sbcz, cbcz, sblx, cblx, sbly, cbly, sblz, cblz, salx, calx, saly, caly, salz, calz;
  ll_sincosf(tS[0], &sbcx, &cbcx); ll_sincosf(tS[1], &sbcy, &cbcy); ll_sincosf(tS[2], &sbcz, &cbcz);
  ll_sincosf(tB[0], &sblx, &cblx); ll_sincosf(tB[1], &sbly, &cbly); ll_sincosf(tB[2], &sblz, &cblz);
  ll_sincosf(tA[0], &salx, &calx); ll_sincosf(tA[1], &saly, &caly); ll_sincosf(tA[2], &salz, &calz);
  float x1 = cbcy * (tB[3] - tS[3]) - sbcy * (tB[5] - tS[5]);
  float y1 = tB[4] - tS[4];
  float z1 = sbcy * (tB[3] - tS[3]) + cbcy * (tB[5] - tS[5]);
  float x2 = x1;
  float y2 = cbcx * y1 + sbcx * z1;
  float z2 = -sbcx * y1 + cbcx * z1;
  const float inc3 = cbcz * x2 + sbcz * y2;
  const float inc4 = -sbcz * x2 + cbcz * y2;
  const float inc5 = z2;
  const float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
                    cbcx * sbcy * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                    cbcx * cbcy * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
  const float t0 = -ll_asinf(srx);
  const float srycrx = sbcx * (cblx * cblz * (caly * salz - calz * salx * saly) - cblx * sblz * (caly * calz + salx * saly * salz) + calx * saly * sblx) -
                       cbcx * cbcy * ((caly * calz + salx * saly * salz) * (cblz * sbly - cbly * sblx * sblz) +
                                      (caly * salz - calz * salx * saly) * (sbly * sblz + cbly * cblz * sblx) - calx * cblx * cbly * saly) +
                       cbcx * sbcy * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) +
                                      (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) + calx * cblx * saly * sbly);
  const float crycrx = sbcx * (cblx * sblz * (calz * saly - caly * salx * salz) - cblx * cblz * (saly * salz + caly * calz * salx) + calx * caly * sblx) +
                       cbcx * cbcy * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) +
                                      (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) + calx * caly * cblx * cbly) -
                       cbcx * sbcy * ((saly * salz + caly * calz * salx) * (cbly * sblz - cblz * sblx * sbly) +
                                      (calz * saly - caly * salx * salz) * (cbly * cblz + sblx * sbly * sblz) - calx * caly * cblx * sbly);
  float st0, ct0;
  ll_sincosf(t0, &st0, &ct0);
  const float t1 = ll_atan2f(srycrx / ct0, crycrx / ct0);
  const float srzcrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) -
                       (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) +
                       cbcx * sbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  const float crzcrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                       (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) +
                       cbcx * cbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  const float t2 = ll_atan2f(srzcrx / ct0, crzcrx / ct0);
  float st1, ct1, st2, ct2;
  ll_sincosf(t1, &st1, &ct1);
  ll_sincosf(t2, &st2, &ct2);
  x1 = ct2 * inc3 - st2 * inc4;
  y1 = st2 * inc3 + ct2 * inc4;
  z1 = inc5;
  x2 = x1;
  y2 = ct0 * y1 - st0 * z1;
  z2 = st0 * y1 + ct0 * z1;
  tT[0] = t0; tT[1] = t1; tT[2] = t2;
  tT[3] = tA[3] - (ct1 * x2 + st1 * z2);
  tT[4] = tA[4] - y2;
  tT[5] = tA[5] - (-st1 * x2 + ct1 * z2);
}

// MapOptimization::transformUpdate (mapOptmization.cpp:389-395), inside the guard of :1316
__global__ void k_map_update(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B || !map_guard(st, s)) return;
  for (int i = 0; i < 6; ++i) {
    st.transform_bef_mapped[s * 6 + i] = st.map_odom[s * 6 + i];
    st.transform_aft_mapped[s * 6 + i] = st.transform_tobe_mapped[s * 6 + i];
  }
}

}  // namespace

void launch_map_predict_pose(LaunchCtx& ctx, DevState& st) {
  LL_LAUNCH(ctx, "k_map_associate", k_map_associate<<<(st.p.B + 63) / 64, 64, 0, ctx.stream>>>(st));
}

void launch_scan_to_map(LaunchCtx& ctx, DevState& st) {
  const DevParams& p = st.p;
  // kd-tree builds of every mapping cycle (mapOptmization.cpp:1317-1318)
  launch_grid_build2(ctx, p.B, st.grid_map_corner, st.map_corner, st.cap_map_corner, st.map_counts, 2, 0,
                     st.grid_map_surf, st.map_surf, st.cap_map_surf, st.map_counts, 2, 1, nullptr, 0);
  LL_LAUNCH(ctx, "k_map_begin", k_map_begin<<<(p.B + 63) / 64, 64, 0, ctx.stream>>>(st));
  if (st.map_knn_trace) cudaMemsetAsync(st.map_knn_trace, 0xff, (size_t)p.B * 10 * st.map_knn_cap * 5 * sizeof(int), ctx.stream);
#ifdef MAP_ITER_NAMES   // measurement build (LEGO_LOAM_B200_NVCC_EXTRA=-DMAP_ITER_NAMES): one timing name per LM iteration
  static const char* const knn_names[10] = {"k_map_knn", "k_map_knn_r1", "k_map_knn_r2", "k_map_knn_r3", "k_map_knn_r4", "k_map_knn_r5", "k_map_knn_r6", "k_map_knn_r7", "k_map_knn_r8", "k_map_knn_r9"};
  static const char* const iter_names[10] = {"k_map_iter", "k_map_iter_c1", "k_map_iter_c2", "k_map_iter_c3", "k_map_iter_c4", "k_map_iter_c5", "k_map_iter_c6", "k_map_iter_c7", "k_map_iter_c8", "k_map_iter_c9"};
#endif
  for (int iter = 0; iter < 10; ++iter) {
    // (two names so that per-kernel timing tells the full search of iteration 0 from the reuse / re-search launches)
#ifdef MAP_ITER_NAMES
    const char* kn = knn_names[iter];
    const char* in = iter_names[iter];
#else
    const char* kn = iter == 0 ? "k_map_knn" : "k_map_knn_reuse";
    const char* in = iter == 0 ? "k_map_iter" : "k_map_iter_cached";
#endif
    LL_LAUNCH(ctx, kn, k_map_knn<<<dim3(KNN_BLOCKS, p.B), KNN_THREADS, 0, ctx.stream>>>(st, iter));
    if (st.map_knn_trace) LL_LAUNCH(ctx, "k_map_knn_trace", k_map_knn_trace<<<dim3(32, p.B), 256, 0, ctx.stream>>>(st, iter));
    LL_LAUNCH(ctx, in, k_map_iter<<<dim3(MAP_BLOCKS, p.B), MAP_THREADS, 0, ctx.stream>>>(st, iter));
  }
  LL_LAUNCH(ctx, "k_map_update", k_map_update<<<(p.B + 63) / 64, 64, 0, ctx.stream>>>(st));
}

