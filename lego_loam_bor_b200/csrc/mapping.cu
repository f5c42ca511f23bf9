// mapping.cu -- MapOptimization::scan2MapOptimization (reference: LeGO-LOAM/src/mapOptmization.cpp:
// 397-426 pointAssociateToMap, 1028-1134 cornerOptimization, 1136-1197 surfOptimization,
// 1199-1312 LMOptimization, 1315-1332 scan2MapOptimization) and downsampleCurrentScan (:999-1026).
//
//   grid build (hashgrid.cu)  replaces the two kd-tree builds of every mapping cycle (:1317-1318);
//       cell = 1 m = the acceptance radius of :1036,1144, so a 27-cell lookup is exact.
//   k_map_knn    one warp per down-sampled scan point: pointAssociateToMap + exact 5-NN (27 cells in parallel)
//   k_map_iter   one thread per down-sampled scan point: line fit
//       (3x3 symmetric eigen-solve) or plane fit (5x3 column-pivoted QR), residual + 6-column
//       Jacobian row, J^T J / J^T r partial sums (exact double products) per block.
//   k_map_solve  one block per sequence: fixed-order sum of the partials, 6x6 solve, degeneracy
//       projection, pose update and convergence flag.  Ten (iter, solve) pairs are enqueued; a
//       converged sequence makes its remaining pairs no-ops, so there is no host round trip.
//   k_voxel_*    pcl::VoxelGrid on whole clouds for downsampleCurrentScan (one block per sequence
//       and cloud: stable LSD radix sort by voxel index in shared/global memory, sequential centroids).
#include "../../include/ll_smallmat.h"
#include "hashgrid.cuh"
#include "ll_kernels.h"

namespace {

#define MAP_BLOCKS 96
#define KNN_BLOCKS 64
#define MAP_THREADS 128
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

__device__ __forceinline__ float4 point_associate_to_map(const MapPose& m, const float4 pi) {  // :412-426
  const float x1 = m.cYaw * pi.x - m.sYaw * pi.y;
  const float y1 = m.sYaw * pi.x + m.cYaw * pi.y;
  const float z1 = pi.z;
  const float x2 = x1;
  const float y2 = m.cRoll * y1 - m.sRoll * z1;
  const float z2 = m.sRoll * y1 + m.cRoll * z1;
  return make_float4(m.cPitch * x2 + m.sPitch * z2 + m.tX, y2 + m.tY, -m.sPitch * x2 + m.cPitch * z2 + m.tZ, pi.w);
}

// cornerOptimization body for one point; returns false if rejected
__device__ __forceinline__ bool corner_fit(const DevState& st, int s, const float4 sel, const int* idx, float4* coeff) {
  if (idx[4] < 0) return false;  // pointSearchSqDis[4] < 1.0 failed
  const float4* mp = st.map_corner + (size_t)s * st.cap_map_corner;
  float4 q[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) q[j] = mp[idx[j]];
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
  if (!(matD1[2] > 3 * matD1[1])) return false;
  const float x0 = sel.x, y0 = sel.y, z0 = sel.z;
  // row 0 of the eigenvector matrix (sic), mapOptmization.cpp:1086-1091; 0.1 * v in double
  const float x1 = (float)((double)cx + 0.1 * (double)matV1[0]);
  const float y1 = (float)((double)cy + 0.1 * (double)matV1[1]);
  const float z1 = (float)((double)cz + 0.1 * (double)matV1[2]);
  const float x2 = (float)((double)cx - 0.1 * (double)matV1[0]);
  const float y2 = (float)((double)cy - 0.1 * (double)matV1[1]);
  const float z2 = (float)((double)cz - 0.1 * (double)matV1[2]);
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

// surfOptimization body for one point
__device__ __forceinline__ bool surf_fit(const DevState& st, int s, const float4 sel, const int* idx, float4* coeff) {
  if (idx[4] < 0) return false;
  const float4* mp = st.map_surf + (size_t)s * st.cap_map_surf;
  float4 q[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) q[j] = mp[idx[j]];
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
  if (!planeValid) return false;
  const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
  const float w = (float)(1 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
  *coeff = make_float4(w * pa, w * pb, w * pc, w * pd2);
  return (double)w > 0.1;
}


// Exact 5-NN of every down-sampled scan point in the local map.
//
// Phase 1 (one thread per query) tries to REUSE the neighbours of the previous LM iteration: the last
// full search left the query position p0 and the squared distance d6 to the 6th nearest point (or 1 m,
// the reach of the 27-cell block).  If the query has moved by |delta| and the five old neighbours are
// now all closer than sqrt(d6) - |delta| (minus a float-safety margin), no other map point can have
// entered the top five (triangle inequality), so only their order has to be recomputed.
// Phase 2 (one warp per remaining query) is the full search: the candidates of the 27 buckets are
// flattened over the 32 lanes, every lane keeps its six best, six rounds of warp arg-min merge them.
#define KNN_THREADS 256
#define KNN_WARPS (KNN_THREADS / 32)

struct Cand {
  float d2;
  int idx;
};
__device__ __forceinline__ bool cand_less(float d2a, int ia, float d2b, int ib) { return d2a < d2b || (d2a == d2b && ia < ib); }

#define KNN_SLOTS 8                 // candidates a lane can hold per pass (32 * 8 = 256 per pass)
#define KNN_NS (KNN_SLOTS + 1)      // + one slot that carries the best six between passes (lanes 0..5)

// Removes and returns the warp-wide minimum (d2, idx) over all slots of all lanes; every slot holding
// that index is invalidated, which also removes the duplicates that hash collisions can produce.
// Squared distances are >= 0, so their bit patterns order like unsigned integers and the minimum is
// one __reduce_min_sync.
__device__ __forceinline__ void knn_extract_min(unsigned (&kd)[KNN_NS], int (&ki)[KNN_NS], unsigned* out_d, int* out_i) {
  unsigned m = kd[0];
#pragma unroll
  for (int i = 1; i < KNN_NS; ++i) m = min(m, kd[i]);
  const unsigned wm = __reduce_min_sync(0xffffffffu, m);
  int mi = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < KNN_NS; ++i)
    if (kd[i] == wm) mi = min(mi, ki[i]);
  const int wi = __reduce_min_sync(0xffffffffu, mi);
#pragma unroll
  for (int i = 0; i < KNN_NS; ++i)
    if (ki[i] == wi) { kd[i] = 0xffffffffu; ki[i] = 0x7fffffff; }
  *out_d = wm;
  *out_i = wi;
}

__device__ __forceinline__ void knn_full_search(const DevState& st, int s, int q, bool corner, const float4 sel, int lane) {
  const HashGrid& g = corner ? st.grid_map_corner : st.grid_map_surf;
  const int* cs = g.cell_start + (size_t)s * (g.tbl + 1);
  const unsigned* occ = g.occ + (size_t)s * (g.tbl / 32);
  const float4* pts = g.sorted + (size_t)s * g.cap;
  const int cx = grid_cell(sel.x, g.inv_cell), cy = grid_cell(sel.y, g.inv_cell), cz = grid_cell(sel.z, g.inv_cell);
  int b0 = 0, len = 0;
  if (lane < 27) {
    const uint32_t h = grid_hash(cx + (lane % 3 - 1), cy + ((lane % 9) / 3 - 1), cz + (lane / 9 - 1), g.tbl);
    if ((occ[h >> 5] >> (h & 31)) & 1u) {
      b0 = cs[h];
      len = cs[h + 1] - b0;
    }
  }
  // inclusive scan of the bucket lengths over the lanes: candidate t lives in the first lane with inc > t
  int inc = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  const int total = __shfl_sync(0xffffffffu, inc, 31);
  const int start_of_mine = b0 - (inc - len);  // bucket start minus exclusive prefix: element k = start + t
  unsigned kd[KNN_NS];
  int ki[KNN_NS];
#pragma unroll
  for (int i = 0; i < KNN_NS; ++i) { kd[i] = 0xffffffffu; ki[i] = 0x7fffffff; }
  unsigned res_d = 0xffffffffu;  // lane r < 6 ends up with the r-th nearest
  int res_i = 0x7fffffff;
  for (int pass0 = 0; pass0 == 0 || pass0 < total; pass0 += 32 * KNN_SLOTS) {
#pragma unroll
    for (int i = 0; i < KNN_SLOTS; ++i) {
      const int t = pass0 + i * 32 + lane;
      int lo = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int probe = __shfl_sync(0xffffffffu, inc, lo + step - 1);
        if (probe <= t) lo += step;
      }
      const int base_k = __shfl_sync(0xffffffffu, start_of_mine, lo);
      if (t < total) {
        const float4 c = pts[base_k + t];
        const float dd = nn_dist2(sel.x, sel.y, sel.z, c);
        if (dd < 1.0f) { kd[i] = __float_as_uint(dd); ki[i] = __float_as_int(c.w); }
      }
    }
    // carried best six of the earlier passes re-enter through the spare slot of lanes 0..5
    kd[KNN_SLOTS] = res_d;
    ki[KNN_SLOTS] = res_i;
    res_d = 0xffffffffu;
    res_i = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      unsigned wd;
      int wi;
      knn_extract_min(kd, ki, &wd, &wi);
      if (lane == r) { res_d = wd; res_i = wi; }
    }
#pragma unroll
    for (int i = 0; i < KNN_NS; ++i) { kd[i] = 0xffffffffu; ki[i] = 0x7fffffff; }
  }
  const size_t qo = (size_t)s * st.map_knn_cap + q;
  if (lane < 5) st.map_knn[qo * 5 + lane] = (res_i == 0x7fffffff) ? -1 : res_i;
  const unsigned d6u = __shfl_sync(0xffffffffu, res_d, 5);
  const int i6 = __shfl_sync(0xffffffffu, res_i, 5);
  if (lane == 0) {
    const float d6 = (i6 == 0x7fffffff) ? 1.0f : fminf(__uint_as_float(d6u), 1.0f);
    st.map_knn_state[qo] = make_float4(sel.x, sel.y, sel.z, d6);
  }
}

__global__ void __launch_bounds__(KNN_THREADS) k_map_knn(DevState st, int iter) {
  __shared__ int sh_need[KNN_THREADS];
  __shared__ int sh_n;
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  if (!map_guard(st, s) || st.map_flags[s * 4 + 1]) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nc = st.scan_ds_counts[s * 2 + 0], ns = st.scan_ds_counts[s * 2 + 1];
  const int nq = nc + ns;
  float T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) T[k] = st.transform_tobe_mapped[s * 6 + k];
  const MapPose mp = make_map_pose(T);
  for (int qbase = blockIdx.x * KNN_THREADS; qbase < nq; qbase += gridDim.x * KNN_THREADS) {
    if (threadIdx.x == 0) sh_n = 0;
    __syncthreads();
    const int q = qbase + threadIdx.x;
    if (q < nq) {
      bool need = true;
      if (iter > 0) {
        const bool corner = q < nc;
        const float4 ori = corner ? st.scan_corner_ds[(size_t)s * p.cap_less_sharp + q] : st.scan_surf_ds[(size_t)s * p.N + (q - nc)];
        const float4 sel = point_associate_to_map(mp, ori);
        const size_t qo = (size_t)s * st.map_knn_cap + q;
        const float4 stt = st.map_knn_state[qo];
        int idx[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) idx[i] = st.map_knn[qo * 5 + i];
        if (idx[4] >= 0) {
          const float4* mpts = corner ? st.map_corner + (size_t)s * st.cap_map_corner : st.map_surf + (size_t)s * st.cap_map_surf;
          float d2[5];
          float dmax = 0.f;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            d2[i] = nn_dist2(sel.x, sel.y, sel.z, mpts[idx[i]]);
            dmax = fmaxf(dmax, d2[i]);
          }
          const float dx = sel.x - stt.x, dy = sel.y - stt.y, dz = sel.z - stt.z;
          const float delta = sqrtf(dx * dx + dy * dy + dz * dz);
          if (sqrtf(dmax) + delta + 1e-4f < sqrtf(stt.w)) {
            // same five neighbours; restore the ascending (d2, idx) order of a fresh search
#pragma unroll
            for (int a = 1; a < 5; ++a) {
#pragma unroll
              for (int b = a; b > 0; --b) {
                if (cand_less(d2[b], idx[b], d2[b - 1], idx[b - 1])) {
                  const float td = d2[b]; d2[b] = d2[b - 1]; d2[b - 1] = td;
                  const int ti = idx[b]; idx[b] = idx[b - 1]; idx[b - 1] = ti;
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) st.map_knn[qo * 5 + i] = idx[i];
            need = false;
          }
        }
      }
      if (need) sh_need[atomicAdd(&sh_n, 1)] = q;
    }
    __syncthreads();
    const int n_need = sh_n;
    for (int i = wid; i < n_need; i += KNN_WARPS) {
      const int qq = sh_need[i];
      const bool corner = qq < nc;
      const float4 ori = corner ? st.scan_corner_ds[(size_t)s * p.cap_less_sharp + qq] : st.scan_surf_ds[(size_t)s * p.N + (qq - nc)];
      const float4 sel = point_associate_to_map(mp, ori);
      knn_full_search(st, s, qq, corner, sel, lane);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(MAP_THREADS) k_map_iter(DevState st) {
  __shared__ double sh_part[MAP_THREADS / 32][MAP_NACC];
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  double* out = st.map_partials + ((size_t)s * MAP_BLOCKS + blockIdx.x) * MAP_NACC;
  if (!map_guard(st, s) || st.map_flags[s * 4 + 1]) return;
  float T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) T[k] = st.transform_tobe_mapped[s * 6 + k];
  const MapPose mp = make_map_pose(T);
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
    int idx[5];
    {
      const int* kn = st.map_knn + ((size_t)s * st.map_knn_cap + q) * 5;
#pragma unroll
      for (int i = 0; i < 5; ++i) idx[i] = kn[i];
    }
    const bool ok = corner ? corner_fit(st, s, sel, idx, &cf) : surf_fit(st, s, sel, idx, &cf);
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
}

__global__ void __launch_bounds__(32) k_map_solve(DevState st, int iter) {
  __shared__ double tot[MAP_NACC];
  const int s = blockIdx.x;
  if (!map_guard(st, s) || st.map_flags[s * 4 + 1]) return;
  if (threadIdx.x < MAP_NACC) {
    const double* part = st.map_partials + (size_t)s * MAP_BLOCKS * MAP_NACC;
    double v = 0.0;
    for (int b = 0; b < MAP_BLOCKS; ++b) v += part[b * MAP_NACC + threadIdx.x];
    tot[threadIdx.x] = v;
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  float* T = st.transform_tobe_mapped + s * 6;
  const int rows = (int)tot[27];
  st.map_iters[s * 2 + 0] = iter + 1;
  st.map_iters[s * 2 + 1] = rows;
  double* trace = st.map_trace + ((size_t)s * 10 + iter) * 34;
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
  float* matP = st.map_matP + s * 36;
  if (iter == 0) st.map_flags[s * 4 + 0] = llm::degeneracy_projector<6>(AtA, 100.f, matP) ? 1 : 0;
  if (st.map_flags[s * 4 + 0]) {
    float X2[6];
    for (int i = 0; i < 6; ++i) X2[i] = X[i];
    for (int r = 0; r < 6; ++r) {
      float v = 0.f;
      for (int c = 0; c < 6; ++c) v += matP[r * 6 + c] * X2[c];
      X[r] = v;
    }
  }
  for (int i = 0; i < 6; ++i) T[i] += X[i];
  for (int i = 0; i < 6; ++i) trace[28 + i] = (double)X[i];
  const float r2d = 57.29578f;  // pcl::rad2deg(float)
  const double r0 = (double)(X[0] * r2d), r1 = (double)(X[1] * r2d), r2 = (double)(X[2] * r2d);
  const double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
  const float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
  const float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
  if ((double)deltaR < 0.05 && (double)deltaT < 0.05) st.map_flags[s * 4 + 1] = 1;
}

__global__ void k_map_begin(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B) return;
  st.map_flags[s * 4 + 1] = 0;
  st.map_iters[s * 2 + 0] = 0;
  st.map_iters[s * 2 + 1] = 0;
}

// MapOptimization::transformAssociateToMap (mapOptmization.cpp:264-387), one thread per sequence.
__global__ void k_map_associate(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B) return;
  const float* tS = st.transform_sum + s * 6;
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
    st.transform_bef_mapped[s * 6 + i] = st.transform_sum[s * 6 + i];
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
  for (int iter = 0; iter < 10; ++iter) {
    LL_LAUNCH(ctx, "k_map_knn", k_map_knn<<<dim3(KNN_BLOCKS, p.B), KNN_THREADS, 0, ctx.stream>>>(st, iter));
    LL_LAUNCH(ctx, "k_map_iter", k_map_iter<<<dim3(MAP_BLOCKS, p.B), MAP_THREADS, 0, ctx.stream>>>(st));
    LL_LAUNCH(ctx, "k_map_solve", k_map_solve<<<p.B, 32, 0, ctx.stream>>>(st, iter));
  }
  LL_LAUNCH(ctx, "k_map_update", k_map_update<<<(p.B + 63) / 64, 64, 0, ctx.stream>>>(st));
}

