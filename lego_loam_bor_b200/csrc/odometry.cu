// odometry.cu -- scan-to-scan LM odometry of FeatureAssociation
// (reference: LeGO-LOAM/src/featureAssociation.cpp:388-500, 503-1032, 1181-1270, 1329-1359).
//
//   k_odom_search_surf          wide correspondence search of LM iteration 0 of the surf stage: a group of 8 lanes per
//       flat point, all sequences at once.  TransformToStart, exact 1-NN in the last-frame cloud (hash grid
//       instead of the kd-tree), then the reference's ring-window scans restated as a second walk over the
//       same hash grid with the window as a predicate (index range from the ring break tables, ring-id group,
//       buckets pre-filtered by a ring signature) and the reference's "first strictly smaller wins" rule as
//       a lexicographic (distance, visiting order) minimum.
//   k_odom_stage<SURF|CORNER>   one block per sequence runs a whole LM stage of updateTransformation
//       (featureAssociation.cpp:1213-1235): up to 25 iterations with the 3x3 solve, the degeneracy test and
//       the convergence test, without a host or kernel boundary in between.  The features of the sequence
//       and the geometry derived from their correspondences (plane / line end points) live in shared memory
//       for the whole stage.  The CORNER stage searches in-block (its clouds are small: 1-NN on the hash
//       grid, window scans linearly); the SURF stage starts from k_odom_search_surf's result and searches
//       in-block only at iterations 5, 10, ... (rare).
//       LM iteration: residual + Jacobian row per feature, J^T J / J^T r as exact double products,
//       warp shuffles + one cross-warp sum, 3x3 column-pivoted Householder solve by one thread.
//       The CORNER launch ends with integrateTransformation (featureAssociation.cpp:1241-1270).
//   k_publish_clouds_last       TransformToEnd on the less-sharp / less-flat clouds into the
//       "last" buffers, adjustOutlierCloud, counts, the kd-tree-rebuild condition and the ring run boundaries
//       the window break tables are made of.
#include "../../include/ll_smallmat.h"
#include "hashgrid.cuh"
#include "ll_kernels.h"
#include "shell_offsets.cuh"

namespace {

enum { STAGE_SURF = 0, STAGE_CORNER = 1 };

// (out of line: the double-precision sin/cos expansions are long, and the stage kernels call this from several places)
__device__ __noinline__ float4 transform_to_start(const float4 pi, const float* T) {
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

// sin / cos of the three whole-sweep angles transformCur[0..2]: the same for every point of a sequence
struct EndTrig { float srx, crx, sry, cry, srz, crz; };

__device__ __forceinline__ float4 transform_to_end(const float4 pi, const float* T, const EndTrig& e) {
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
  tx = T[3]; ty = T[4]; tz = T[5];
  srz = e.srz; crz = e.crz; srx = e.srx; crx = e.crx; sry = e.sry; cry = e.cry;  // of T[2], T[0], T[1] (:452-460)
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

// Running minimum with a tie-break key, plus the smallest distance among all OTHER candidates seen (`second`,
// not limited by the acceptance cap the minimum starts from).  With key = visiting order this is the reference's
// "first strictly smaller wins" rule as a lexicographic (distance, order) minimum that lanes can evaluate in
// parallel; with key = point index it is the 1-NN with ties to the lowest index.  `second` is what lets a later
// LM iteration prove that the minimum cannot have changed (see CorrS::slack).
#define BEST_NONE 0x7fffffff
struct Best {
  float d2;      // starts at the acceptance cap: only strictly smaller distances are accepted
  int key;       // BEST_NONE: no candidate accepted yet
  int w;         // payload (packed grid word or cloud index)
  float second;
};
__device__ __forceinline__ Best best_init(float cap) { return Best{cap, BEST_NONE, -1, FLT_MAX}; }
__device__ __forceinline__ void best_update(Best& m, float d2, int key, int w) {
  if (d2 < m.d2 || (d2 == m.d2 && m.key != BEST_NONE && key < m.key)) {
    if (m.key != BEST_NONE) m.second = fminf(m.second, m.d2);
    m.d2 = d2; m.key = key; m.w = w;
  } else {
    m.second = fminf(m.second, d2);
  }
}
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool odom_guard(const DevState& st, int s) {
  // featureAssociation.cpp:1214
  return !(st.last_counts[s * 2 + 0] < 10 || st.last_counts[s * 2 + 1] < 100);
}

#ifndef LM_THREADS
#ifndef LM_THREADS
#define LM_THREADS 768
#endif
#endif
#define LM_WARPS (LM_THREADS / 32)
#define WIN_R (LL_MAX_RINGS + 8)

// Break positions of the ring-window scans.  The reference walks upwards from closest+1 until the first
// point whose ring id exceeds cs + 2.5 and downwards from closest-1 until the first id below cs - 2.5
// (featureAssociation.cpp:522-563,661-712).  Ring ids are non-decreasing along the cloud up to a -1 wobble
// (a negative relTime truncates to ring-1), which makes those break positions a function of cs alone: the
// first index of the WHOLE cloud with id >= cs+3 and the last index with id <= cs-3.
struct WinTables {
  int up_tab[WIN_R];   // up_tab[v]: first index of the last-frame cloud whose ring id is >= v (cloud size if none)
  int dn_tab[WIN_R];   // dn_tab[v]: last index whose ring id is <= v (-1 if none)
};

// Shared state of one stage of one sequence.
struct StageShared {
  WinTables win;
  float T[6];
  double part[LM_WARPS][10];
  int stop, iters;
};

// win_first / win_last hold, per ring id, the first / last index of a run of that id (k_publish_clouds_last);
// the tables are their suffix minimum / prefix maximum.  Called by all threads of the block, ends with a barrier.
__device__ __forceinline__ void build_window_tables(WinTables& w, const DevState& st, int s, int cloud, int n) {
  const int par = (int)((st.frame_tag - 1u) & 1u);  // written while the previous frame was published
  const int* first = st.win_first + (((size_t)par * st.p.B + s) * 2 + cloud) * WIN_R;
  const int* lastv = st.win_last + (((size_t)par * st.p.B + s) * 2 + cloud) * WIN_R;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int carry = n;
    for (int base = WIN_R - 1; base >= 0; base -= 32) {
      const int v = base - lane;
      int x = v >= 0 ? first[v] : 0x7fffffff;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = min(x, y);
      }
      x = min(x, carry);
      if (v >= 0) w.up_tab[v] = x;
      carry = __shfl_sync(0xffffffffu, x, 31);
    }
    carry = -1;
    for (int base = 0; base < WIN_R; base += 32) {
      const int v = base + lane;
      int x = v < WIN_R ? lastv[v] : -1;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = max(x, y);
      }
      x = max(x, carry);
      if (v < WIN_R) w.dn_tab[v] = x;
      carry = __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
}

#define SG 32                      // lanes that cooperate on the search of one feature point (a warp: sub-warp groups
                                   // would run their data-dependent loops one after the other anyway)
#define SLACK_HORIZON 0.2f         // cells farther than the current minimum + this are not visited; caps CorrS::slack at half of it
#define LM_GROUPS (LM_THREADS / SG)
#define SEARCH_THREADS 256
#define SEARCH_GROUPS (SEARCH_THREADS / SG)
#ifndef SEARCH_QPW_CORNER
#define SEARCH_QPW_CORNER 4   // corner feature points per warp of k_odom_search (1 / 2 / 4 measured alike)
#endif

// Warp-wide reduction of Best (all 32 lanes; SG == 32): hardware integer min-reductions on the distance bits
// (non-negative floats order like unsigned ints) and on the tie-break key, payload from the winning lane.
template <int G>
__device__ __forceinline__ void best_group_reduce(unsigned gmask, Best& m) {
  if (G == 1) return;  // one thread per feature point: nothing to merge
  const unsigned mybits = m.key != BEST_NONE ? __float_as_uint(m.d2) : 0xffffffffu;
  const unsigned minbits = __reduce_min_sync(gmask, mybits);
  if (minbits == 0xffffffffu) {  // no lane has a candidate: only the runner-up distances need merging
    m.second = __uint_as_float(__reduce_min_sync(gmask, __float_as_uint(m.second)));
    return;
  }
  const unsigned mykey = mybits == minbits ? (unsigned)m.key : 0xffffffffu;
  const unsigned minkey = __reduce_min_sync(gmask, mykey);
  const bool holds = mybits == minbits && (unsigned)m.key == minkey;  // lanes holding the winning candidate (same key => same candidate)
  const int src = __ffs(__ballot_sync(gmask, holds)) - 1;
  const int ww = __shfl_sync(gmask, m.w, src);
  // runner-up: every lane's `second`, plus the minimum of the lanes that hold a different candidate
  const float mine = (holds || m.key == BEST_NONE) ? m.second : fminf(m.second, m.d2);
  m.second = __uint_as_float(__reduce_min_sync(gmask, __float_as_uint(mine)));
  m.d2 = __uint_as_float(minbits); m.key = (int)minkey; m.w = ww;
}

// The SG lanes of a group visit the buckets of shell r (cells at Chebyshev distance r from (cx,cy,cz)).  Cells are
// tested SG at a time, one per lane (occupancy bit, optional ring signature, bucket range); the non-empty buckets
// are then walked one after the other with their points spread over the lanes, so that long buckets (clustered
// edge points) do not serialise on one lane.  `want` != 0 restricts the visit to buckets whose ring signature
// intersects it.  f(point) is called for every stored point of the visited buckets (by some lane of the group);
// aliasing buckets only add candidates.
template <int G, typename F>
__device__ __forceinline__ void group_visit_shell(const HashGrid& g, const int* __restrict__ cs, const unsigned* __restrict__ occ,
                                                  const unsigned* __restrict__ sig, unsigned want, const float4* __restrict__ pts,
                                                  int cx, int cy, int cz, int r_lo, int r, int gl, unsigned gmask, float qx, float qy,
                                                  float qz, float bound2, F&& f) {
  // shells r_lo .. r in one pass (r_lo < r only inside the offset table)
  // bound2: cells whose box is farther than sqrt(bound2) from the query (qx, qy, qz) are skipped
  const int side = 2 * r + 1;
  const bool tabled = r <= LL_SHELL_TABLE_R;
  const int t_begin = tabled ? k_shell_start[r_lo] : 0;
  const int t_end = tabled ? k_shell_start[r + 1] : side * side * side;
  for (int t0 = t_begin; t0 < t_end; t0 += G) {  // uniform over the group
    const int t = t0 + gl;
    int b0 = 0, b1 = 0;
    if (t < t_end) {
      int dx, dy, dz;
      if (tabled) {
        const int o = k_shell_ofs[t];
        dx = (o & 15) - 8; dy = ((o >> 4) & 15) - 8; dz = (o >> 8) - 8;
      } else {  // beyond the table: enumerate the cube and skip its interior
        dz = t / (side * side) - r;
        const int rem = t % (side * side);
        dy = rem / side - r; dx = rem % side - r;
      }
      bool visit = tabled || max(max(abs(dx), abs(dy)), abs(dz)) == r;
      if (visit && (dx | dy | dz) != 0) {
        // distance from the query to the cell's box, shrunk a little so that rounding in grid_cell() cannot matter
        const float tol = 1e-3f * g.cell;
        const float lx = (float)(cx + dx) * g.cell, ly = (float)(cy + dy) * g.cell, lz = (float)(cz + dz) * g.cell;
        const float gx = fmaxf(fmaxf(lx - qx, qx - (lx + g.cell)) - tol, 0.f);
        const float gy = fmaxf(fmaxf(ly - qy, qy - (ly + g.cell)) - tol, 0.f);
        const float gz = fmaxf(fmaxf(lz - qz, qz - (lz + g.cell)) - tol, 0.f);
        visit = gx * gx + gy * gy + gz * gz < bound2;
      }
      if (visit) {
        const uint32_t h = grid_hash(cx + dx, cy + dy, cz + dz, g.tbl);
        if ((__ldg(occ + (h >> 5)) >> (h & 31)) & 1u) {
          if (want == 0u || (__ldg(sig + h) & want)) { b0 = __ldg(cs + h); b1 = __ldg(cs + h + 1); }
        }
      }
    }
    if (G == 1) {
      for (int k = b0; k < b1; ++k) f(__ldg(pts + k));
    } else {
      unsigned m = __ballot_sync(gmask, b1 > b0) & gmask;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int sb0 = __shfl_sync(gmask, b0, src), sb1 = __shfl_sync(gmask, b1, src);
        for (int k = sb0 + gl; k < sb1; k += G) f(__ldg(pts + k));
      }
    }
  }
}

// Result of a correspondence search.  slack (metres): if the transformed feature point moves by less than
// `slack` from where it was searched, closest / ind2 / ind3 provably stay what they are: every other candidate
// of each of the three minima is more than 2 * slack farther away than the winner (triangle inequality on both),
// the winners stay inside the acceptance radius, an empty minimum stays empty, and the window (a function of the
// closest point's ring) stays the same.  slack <= 0: nothing can be proved.
struct CorrS {
  int closest, ind2, ind3;
  float slack;
};

// Correspondence search for one feature point by a group of SG lanes (all lanes pass the same `sel` and get the
// same result).  SURF: featureAssociation.cpp:649-718; CORNER: featureAssociation.cpp:511-568.
// fresh: the hash grid indexes the current last-frame cloud and its sorted points carry index | (ring id + 1) << 24
// in .w, so the ring-window scans of the SURF stage become a second walk over the grid with the window as a
// predicate; the CORNER windows are a few hundred points and are scanned linearly like the reference does.
// !fresh: the grid still indexes an OLDER cloud (the reference rebuilds its kd-trees only when both last-frame
// clouds are large enough, featureAssociation.cpp:1356, while the clouds are always swapped): the 1-NN runs on
// the stale grid and the window scans run linearly over the current cloud, as in the reference.  Rare.
template <int STAGE, int G>
__device__ __forceinline__ CorrS group_search(const DevState& st, const WinTables& win, int s, const float4 sel, int cur_n, int last_n,
                                              const float4* __restrict__ last, bool fresh, int gl, unsigned gmask, int seed = -1) {
  const DevParams& p = st.p;
  const bool surf = (STAGE == STAGE_SURF);
  const HashGrid& g = surf ? st.grid_surf_last : st.grid_corner_last;
  const int* cs_tab = g.cell_start + (size_t)s * grid_cs_stride(g);
  const unsigned* occ = g.occ + (size_t)s * (g.tbl / 32);
  const unsigned* sig = g.sig + (size_t)s * g.tbl;
  const float4* pts = g.sorted + (size_t)s * g.cap;
  const int cx = grid_cell(sel.x, g.inv_cell), cy = grid_cell(sel.y, g.inv_cell), cz = grid_cell(sel.z, g.inv_cell);
  const float cap = p.nearest_feature_dist_sqr;
  const float lim = sqrtf(cap);
  const int rmax = (int)ceilf(lim * g.inv_cell);
  CorrS c{-1, -1, -1, 0.f};
  // ---- exact 1-NN with d2 < nearest_feature_dist_sqr; ties: lowest index ----
  Best nn = best_init(cap);
  if (fresh && seed >= 0 && seed < last_n) {
    // re-search: the previous closest point is a candidate that usually is, or is next to, the new one; starting
    // from it lets the walk skip every cell that is farther away
    const float4 q = last[seed];
    best_update(nn, nn_dist2(sel.x, sel.y, sel.z, q), seed, seed | (min(max((int)q.w + 1, 0), 255) << 24));
  }
  float seen2 = 0.f;  // everything closer than sqrt(seen2) has been visited
  for (int r = 0; r <= rmax; ++r) {
    Best t = best_init(cap);
    const float hb = sqrtf(nn.d2) + SLACK_HORIZON;
    // a seeded search already has a tight bound: its first pass covers the whole 3x3x3 block (one round of dependent loads less)
    const int r_lo = r;
    if (r == 0 && nn.key != BEST_NONE && rmax >= 1 && LL_SHELL_TABLE_R >= 1) r = 1;
    group_visit_shell<G>(g, cs_tab, occ, sig, 0u, pts, cx, cy, cz, r_lo, r, gl, gmask, sel.x, sel.y, sel.z, hb * hb, [&](const float4 q) {
      const int w = __float_as_int(q.w);
      best_update(t, nn_dist2(sel.x, sel.y, sel.z, q), w & 0xffffff, w);
    });
    best_group_reduce<G>(gmask, t);
    // merge the shell into the running result (both are uniform over the group)
    float sec = fminf(nn.second, t.second);
    if (t.key != BEST_NONE && (nn.key == BEST_NONE || t.d2 < nn.d2 || (t.d2 == nn.d2 && t.key < nn.key))) {
      if (nn.key != BEST_NONE) sec = fminf(sec, nn.d2);
      nn.d2 = t.d2; nn.key = t.key; nn.w = t.w;
    } else if (t.key != BEST_NONE && !(t.key == nn.key && t.w == nn.w)) {
      sec = fminf(sec, t.d2);
    }
    nn.second = sec;
    const float reach = (float)r * g.cell;
    seen2 = reach * reach * 0.9999f;
    if (nn.d2 < seen2) break;  // nothing outside the visited block can be closer
  }
  const int closest = nn.key;
  if (nn.key == BEST_NONE || closest >= last_n) return c;
  c.closest = closest;
  const float sq1 = sqrtf(nn.d2);
  // runner-ups closer than min(reach of the visited block, the final pruning horizon) have all been visited
  float slack = fminf(lim - sq1, 0.5f * (fminf(sqrtf(fminf(nn.second, seen2)), sq1 + SLACK_HORIZON - 0.01f) - sq1));
  const int csr = fresh ? (int)((unsigned)nn.w >> 24) - 1 : (int)last[closest].w;  // closestPointScan
  const int up_break = win.up_tab[min(max(csr + 3, 0), WIN_R - 1)];
  const int dn_break = (csr - 3 >= 0) ? win.dn_tab[min(csr - 3, WIN_R - 1)] : -1;
  // the upward scan is also bounded by the CURRENT frame's feature count (sic, featureAssociation.cpp:522,661)
  const int jend = min(min(cur_n, last_n), up_break);
  Best m2 = best_init(cap), m3 = best_init(cap);
  float wseen2 = FLT_MAX;  // window candidates closer than sqrt(wseen2) have all been visited
  if (fresh && surf) {
    // SURF windows span thousands of points: walk the grid instead.  Ring ids csr-2 .. csr+2 as signature bits of (id + 1) & 31
    unsigned want = 0u;
#pragma unroll
    for (int d = -2; d <= 2; ++d) want |= 1u << ((csr + 1 + d) & 31);
    float wb = lim + SLACK_HORIZON;  // pruning horizon: the farther of the two current minima + SLACK_HORIZON
    for (int r = 0; r <= rmax; ++r) {
      group_visit_shell<G>(g, cs_tab, occ, sig, want, pts, cx, cy, cz, r, r, gl, gmask, sel.x, sel.y, sel.z, wb * wb, [&](const float4 q) {
        const int w = __float_as_int(q.w);
        const int j = w & 0xffffff;
        const int id = (int)((unsigned)w >> 24) - 1;
        if (j == closest) return;
        const bool up = j > closest;
        if (up ? (j >= jend) : (j <= dn_break)) return;
        const float d2 = nn_dist2(sel.x, sel.y, sel.z, q);  // == (q - sel)^2 summed left to right
        const int ord = up ? (j - closest) : (0x40000000 + (closest - j));
        const bool same = up ? (id <= csr) : (id >= csr);
        if (same) best_update(m2, d2, ord, j); else best_update(m3, d2, ord, j);
      });
      Best t2 = m2, t3 = m3;
      best_group_reduce<G>(gmask, t2);
      best_group_reduce<G>(gmask, t3);
      const float reach = (float)r * g.cell;
      wseen2 = reach * reach * 0.9999f;
      wb = sqrtf(fmaxf(t2.d2, t3.d2)) + SLACK_HORIZON;  // d2 is the cap while a minimum is still empty
      const bool done2 = t2.key != BEST_NONE && t2.d2 < wseen2;
      const bool done3 = t3.key != BEST_NONE && t3.d2 < wseen2;
      if ((done2 && done3) || r == rmax) {
        m2 = t2; m3 = t3;
        // skipped cells were farther than (the larger minimum at that time + SLACK_HORIZON) >= the final horizon
        wseen2 = fminf(wseen2, (wb - 0.5f * SLACK_HORIZON) * (wb - 0.5f * SLACK_HORIZON));
        break;
      }
    }
  } else {
    // candidates of the upward scan have id > csr (CORNER) and those of the downward scan id < csr, so the same-ring
    // runs next to the closest point can be skipped there: ids below csr + 1 end before up_tab[csr + 1]
    const int up0 = surf ? closest + 1 : max(closest + 1, win.up_tab[min(max(csr + 1, 0), WIN_R - 1)]);
    const int dn0 = surf ? closest - 1 : min(closest - 1, (csr - 1 >= 0) ? win.dn_tab[min(csr - 1, WIN_R - 1)] : -1);
    for (int j = up0 + gl; j < jend; j += G) {
      const float4 q = last[j];
      const int id = (int)q.w;
      const float d2 = sq_dist_ref(q, sel);
      const int ord = j - closest;
      if (surf) {
        if (id <= csr) best_update(m2, d2, ord, j); else best_update(m3, d2, ord, j);
      } else if (id > csr) {
        best_update(m2, d2, ord, j);
      }
    }
    for (int j = dn0 - gl; j > dn_break; j -= G) {
      const float4 q = last[j];
      const int id = (int)q.w;
      const float d2 = sq_dist_ref(q, sel);
      const int ord = 0x40000000 + (closest - j);
      if (surf) {
        if (id >= csr) best_update(m2, d2, ord, j); else best_update(m3, d2, ord, j);
      } else if (id < csr) {
        best_update(m2, d2, ord, j);
      }
    }
    best_group_reduce<G>(gmask, m2);
    if (surf) best_group_reduce<G>(gmask, m3);
  }
  // slack of the window minima: a found one must keep its lead and stay inside the radius, an empty one must stay empty
  if (m2.key != BEST_NONE) {
    const float sq = sqrtf(m2.d2);
    slack = fminf(slack, fminf(lim - sq, 0.5f * (sqrtf(fminf(m2.second, wseen2)) - sq)));
    c.ind2 = m2.w;
  } else {
    slack = fminf(slack, sqrtf(fminf(m2.second, wseen2)) - lim);
  }
  if (surf) {
    if (m3.key != BEST_NONE) {
      const float sq = sqrtf(m3.d2);
      slack = fminf(slack, fminf(lim - sq, 0.5f * (sqrtf(fminf(m3.second, wseen2)) - sq)));
      c.ind3 = m3.w;
    } else {
      slack = fminf(slack, sqrtf(fminf(m3.second, wseen2)) - lim);
    }
  }
  c.slack = fresh ? slack : 0.f;
  return c;
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


// featureAssociation.cpp:721-777 + 834-857; `pl` = the unit plane (pa, pb, pc, pd) through the three
// correspondences, which the reference recomputes every iteration from the same three points
__device__ __forceinline__ Row3 surf_row(const float4 ori, const float4 sel, const float4 pl, const SurfCoef& k, int iter) {
  Row3 r;
  r.ok = false;
  const float pa = pl.x, pb = pl.y, pc = pl.z, pd = pl.w;
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

__device__ __forceinline__ float4 surf_plane(const float4 t1, const float4 t2, const float4 t3) {
  // featureAssociation.cpp:734-747
  float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
  float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
  float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
  float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
  const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  return make_float4(pa, pb, pc, pd);
}

// featureAssociation.cpp:571-635 + 960-978
__device__ __forceinline__ Row3 corner_row(const float4 ori, const float4 sel, const float4 t1, const float4 t2, const CornerCoef& k, int iter) {
  Row3 r;
  r.ok = false;
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

// integrateTransformation, featureAssociation.cpp:1241-1270 (one thread)
__device__ __forceinline__ void integrate_transformation(const DevState& st, int s) {
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
  {
    // the grid covers the capacity (N points); blocks past the longest of the three clouds have nothing to do
    const int longest = max(max(st.feat_counts[s * 4 + 1], st.feat_counts[s * 4 + 3]), max(first_frame ? 0 : st.outlier_count[s], 2 * WIN_R));
    if ((int)(blockIdx.x * blockDim.x) >= longest) return;
  }
  float T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) T[k] = st.transform_cur[s * 6 + k];
  __shared__ EndTrig sh_end;
  if (threadIdx.x == 0) {
    ll_sincosf(T[0], &sh_end.srx, &sh_end.crx);
    ll_sincosf(T[1], &sh_end.sry, &sh_end.cry);
    ll_sincosf(T[2], &sh_end.srz, &sh_end.crz);
  }
  __syncthreads();
  const EndTrig et = sh_end;
  const int n_corner = st.feat_counts[s * 4 + 1];
  const int n_surf = st.feat_counts[s * 4 + 3];
  // window tables are double buffered by frame parity: this frame fills one half (reset while the previous
  // frame was published, or by ll_reset) and resets the other half for the next frame
  const int par = (int)(st.frame_tag & 1u);
  int* wfirst = st.win_first + ((size_t)par * p.B + s) * 2 * WIN_R;
  int* wlast = st.win_last + ((size_t)par * p.B + s) * 2 * WIN_R;
  if (i < 2 * WIN_R) {
    st.win_first[((size_t)(par ^ 1) * p.B + s) * 2 * WIN_R + i] = 0x7fffffff;
    st.win_last[((size_t)(par ^ 1) * p.B + s) * 2 * WIN_R + i] = -1;
  }
  if (i < n_corner) {
    const float4 q = st.corner_less_sharp[(size_t)s * p.cap_less_sharp + i];
    st.corner_last[(size_t)s * p.cap_less_sharp + i] = first_frame ? q : transform_to_end(q, T, et);
    // ring run boundaries of the new last-frame cloud: id = int(intensity) is the same before and after TransformToEnd
    const float4* src = st.corner_less_sharp + (size_t)s * p.cap_less_sharp;
    const int id = min(max((int)q.w, 0), WIN_R - 1);
    const int idp = i > 0 ? min(max((int)src[i - 1].w, 0), WIN_R - 1) : -1;
    const int idn = i + 1 < n_corner ? min(max((int)src[i + 1].w, 0), WIN_R - 1) : -1;
    if (idp != id) atomicMin(wfirst + id, i);
    if (idn != id) atomicMax(wlast + id, i);
  }
  if (i < n_surf) {
    const float4 q = st.surf_less_flat[(size_t)s * p.N + i];
    st.surf_last[(size_t)s * p.N + i] = first_frame ? q : transform_to_end(q, T, et);
    const float4* src = st.surf_less_flat + (size_t)s * p.N;
    const int id = min(max((int)q.w, 0), WIN_R - 1);
    const int idp = i > 0 ? min(max((int)src[i - 1].w, 0), WIN_R - 1) : -1;
    const int idn = i + 1 < n_surf ? min(max((int)src[i + 1].w, 0), WIN_R - 1) : -1;
    if (idp != id) atomicMin(wfirst + WIN_R + id, i);
    if (idn != id) atomicMax(wlast + WIN_R + id, i);
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

// Geometry of one feature point's correspondences as the LM rows need it: SURF: the unit plane through the
// three points (featureAssociation.cpp:734-747, recomputed by the reference in every iteration from the same
// three points); CORNER: the two line points.
template <int STAGE>
__device__ __forceinline__ bool corr_geometry(const CorrS& c, const float4* __restrict__ last, float4* ga, float4* gb) {
  if (STAGE == STAGE_SURF) {
    if (!(c.ind2 >= 0 && c.ind3 >= 0)) return false;
    *ga = surf_plane(last[c.closest], last[c.ind2], last[c.ind3]);
    return true;
  }
  if (!(c.ind2 >= 0)) return false;
  *ga = last[c.closest];
  *gb = last[c.ind2];
  return true;
}

// Correspondences of LM iteration 0 of one stage for all sequences (a group of SG lanes per feature point).
template <int STAGE>
__global__ void __launch_bounds__(SEARCH_THREADS) k_odom_search(DevState st) {
  // feature points per warp: the long SURF searches want as many warps in flight as possible
  constexpr int SEARCH_QPW = STAGE == STAGE_SURF ? 1 : SEARCH_QPW_CORNER;
  __shared__ WinTables win;
  const DevParams& p = st.p;
  const bool surf = (STAGE == STAGE_SURF);
  const int s = blockIdx.y;
  if (!odom_guard(st, s)) return;
  const int n = st.feat_counts[s * 4 + (surf ? 2 : 0)];
  if (blockIdx.x * SEARCH_GROUPS * SEARCH_QPW >= n) return;
  const int last_n = st.last_counts[s * 2 + (surf ? 1 : 0)];
  const bool fresh = st.odom_flags[s * 4 + 2] != 0;
  build_window_tables(win, st, s, surf ? 1 : 0, last_n);
  float T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) T[k] = st.transform_cur[s * 6 + k];
  const int lane = threadIdx.x & 31;
  const int gl = threadIdx.x & (SG - 1), grp = threadIdx.x / SG;
  const unsigned gmask = SG == 32 ? 0xffffffffu : ((1u << (SG & 31)) - 1u) << (lane & ~(SG - 1));
  const int cap = surf ? p.cap_flat : p.cap_sharp;
  const float4* cur = surf ? st.surf_flat + (size_t)s * p.cap_flat : st.corner_sharp + (size_t)s * p.cap_sharp;
  const float4* last = surf ? st.surf_last + (size_t)s * p.N : st.corner_last + (size_t)s * p.cap_less_sharp;
  // a warp takes SEARCH_QPW consecutive feature points: their TransformToStart (long double-precision chains) run
  // side by side on the first lanes, then the warp searches them one after the other
  const int i_base = (blockIdx.x * SEARCH_GROUPS + grp) * SEARCH_QPW;
  if (i_base >= n) return;
  float4 mysel = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gl < SEARCH_QPW && i_base + gl < n) mysel = transform_to_start(cur[i_base + gl], T);
  for (int q = 0; q < SEARCH_QPW; ++q) {
    const int i = i_base + q;
    if (i >= n) break;
    float4 sel;
    sel.x = __shfl_sync(gmask, mysel.x, q, SG);
    sel.y = __shfl_sync(gmask, mysel.y, q, SG);
    sel.z = __shfl_sync(gmask, mysel.z, q, SG);
    sel.w = 0.f;
    const CorrS c = group_search<STAGE, SG>(st, win, s, sel, n, last_n, last, fresh, gl, gmask);
    if (gl == 0) {
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
      const bool ok = corr_geometry<STAGE>(c, last, &ga, &gb);
      st.odom_ok[(size_t)s * p.cap_flat + i] = ok ? 1 : 0;
      st.odom_ga[(size_t)s * p.cap_flat + i] = ga;
      if (!surf) st.odom_gb[(size_t)s * cap + i] = gb;
      st.odom_s0[(size_t)s * p.cap_flat + i] = make_float4(sel.x, sel.y, sel.z, c.slack);
      st.odom_cl[(size_t)s * p.cap_flat + i] = c.closest;
      if (st.odom_trace) {  // parity aid: search round 0
        int* tr = st.odom_trace + ((((size_t)s * 2 + STAGE) * 5 + 0) * p.cap_flat + i) * 3;
        tr[0] = c.closest; tr[1] = c.ind2; tr[2] = surf ? c.ind3 : -1;
      }
    }
  }
}

// One LM stage of one sequence.  Dynamic shared memory: float4 ori[cap], ga[cap], s0[cap], gb[cap] (corner only),
// unsigned short list[cap], unsigned char ok[cap]; cap = 24V (SURF) or 12V (CORNER).
template <int STAGE>
__global__ void __launch_bounds__(LM_THREADS, 1) k_odom_stage(DevState st) {
  extern __shared__ float4 sh_dyn[];
  __shared__ StageShared sh;
  __shared__ int sh_nlist, sh_next;
  const DevParams& p = st.p;
  const int s = blockIdx.x;
  const bool surf = (STAGE == STAGE_SURF);
  const int cap = surf ? p.cap_flat : p.cap_sharp;
  float4* sh_ori = sh_dyn;
  float4* sh_ga = sh_dyn + cap;                       // SURF: unit plane; CORNER: tripod1
  float4* sh_s0 = sh_dyn + 2 * cap;                   // where the point was when it was searched (x, y, z), slack
  float4* sh_gb = sh_dyn + 3 * cap;                   // CORNER: tripod2
  int* sh_cl = reinterpret_cast<int*>(sh_dyn + (surf ? 3 : 4) * cap);  // closest point of the last search (-1: none)
  unsigned short* sh_list = reinterpret_cast<unsigned short*>(sh_cl + cap);
  unsigned char* sh_ok = reinterpret_cast<unsigned char*>(sh_list + cap);
  const bool guard = odom_guard(st, s);
  long long clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // thread 0 only: LL_BUF_STAGE_CLOCKS
  const long long t_begin = global_ns();
  long long t_mark = t_begin;
  // (%globaltimer reads on the solving thread cost 5 us of a 130 us launch: only with ll_enable_stage_timing)
  const bool clocks_on = st.stage_clocks_on != 0 && threadIdx.x == 0;
#define STAGE_CLOCK(slot) do { if (clocks_on) { const long long t__ = global_ns(); clk[slot] += t__ - t_mark; t_mark = t__; } } while (0)
  if (!guard) {
    if (threadIdx.x == 0) {
      st.odom_iters[s * 2 + STAGE] = 0;
      if (!surf) integrate_transformation(st, s);  // runs whether or not updateTransformation returned early
    }
    return;
  }
  const float4* cur = surf ? st.surf_flat + (size_t)s * p.cap_flat : st.corner_sharp + (size_t)s * p.cap_sharp;
  const float4* last = surf ? st.surf_last + (size_t)s * p.N : st.corner_last + (size_t)s * p.cap_less_sharp;
  const int n = st.feat_counts[s * 4 + (surf ? 2 : 0)];
  const int last_n = st.last_counts[s * 2 + (surf ? 1 : 0)];
  const bool fresh = st.odom_flags[s * 4 + 2] != 0;  // the grids index the current last-frame clouds
  if (threadIdx.x < 6) sh.T[threadIdx.x] = st.transform_cur[s * 6 + threadIdx.x];
  if (threadIdx.x == 0) { sh.stop = 0; sh.iters = 0; }
  // iteration 0 correspondences come from k_odom_search
  for (int i = threadIdx.x; i < n; i += LM_THREADS) {
    sh_ori[i] = cur[i];
    sh_ga[i] = st.odom_ga[(size_t)s * p.cap_flat + i];
    sh_s0[i] = st.odom_s0[(size_t)s * p.cap_flat + i];
    sh_ok[i] = st.odom_ok[(size_t)s * p.cap_flat + i];
    sh_cl[i] = st.odom_cl[(size_t)s * p.cap_flat + i];
    if (!surf) sh_gb[i] = st.odom_gb[(size_t)s * cap + i];
  }
  build_window_tables(sh.win, st, s, surf ? 1 : 0, last_n);  // ends with a barrier
  STAGE_CLOCK(1);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gl = threadIdx.x & (SG - 1);
  const unsigned gmask = SG == 32 ? 0xffffffffu : ((1u << (SG & 31)) - 1u) << (lane & ~(SG - 1));
  // isDegenerate (a member shared by both stages, featureAssociation.h:115): read once, kept in a register by the
  // thread that solves; iteration 0 overwrites it when it has enough rows
  int degenerate = st.odom_flags[s * 4 + 0];
  for (int iter = 0; iter < 25; ++iter) {
    float T[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) T[k] = sh.T[k];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    SurfCoef ks;
    CornerCoef kc;
    if (surf) ks = make_surf_coef(T); else kc = make_corner_coef(T);
    if (iter % 5 == 0 && iter > 0) {
      // Correspondences are refreshed every 5th iteration (featureAssociation.cpp:511,649).  A point that has moved by
      // less than its slack since it was searched keeps its correspondences (CorrS::slack); the others are searched again.
      if (threadIdx.x == 0) { sh_nlist = 0; sh_next = 0; }
      if (st.odom_trace) {  // parity aid: a point that keeps its correspondences keeps the indices of the previous round
        int* tr = st.odom_trace + (((size_t)s * 2 + STAGE) * 5 + iter / 5) * p.cap_flat * 3;
        for (int i = threadIdx.x; i < n * 3; i += LM_THREADS) tr[i] = tr[i - p.cap_flat * 3];
      }
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += LM_THREADS) {
        const float4 sel = transform_to_start(sh_ori[i], T);
        const float4 s0 = sh_s0[i];
        const float dx = sel.x - s0.x, dy = sel.y - s0.y, dz = sel.z - s0.z;
        const float moved = sqrtf(dx * dx + dy * dy + dz * dz);
        if (!(moved + 1e-4f < s0.w)) {
          sh_list[atomicAdd(&sh_nlist, 1)] = (unsigned short)i;
          sh_s0[i] = sel;  // position and slack of the old search are dead now: carry the transformed point to the new one
        }
      }
      __syncthreads();
      const int nlist = sh_nlist;
      if (threadIdx.x == 0) clk[5] += nlist;
      // a warp per point; searches differ a lot in length, so warps take the next list entry when they are done with one
      // (one THREAD per point was tried for the seeded re-search: the few points without a nearby seed walk every shell
      // and hold their whole warp, which made the stage several times slower)
      for (;;) {
        int l = 0;
        if (gl == 0) l = atomicAdd(&sh_next, 1);
        l = __shfl_sync(gmask, l, 0, SG);
        if (l >= nlist) break;
        const int i = sh_list[l];
        const float4 sel = sh_s0[i];
        __syncwarp(gmask);  // every lane has read it before lane 0 overwrites it below
        const CorrS c = group_search<STAGE, SG>(st, sh.win, s, sel, n, last_n, last, fresh, gl, gmask, sh_cl[i]);
        if (gl == 0) {
          float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
          sh_cl[i] = c.closest;
          sh_ok[i] = corr_geometry<STAGE>(c, last, &ga, &gb) ? 1 : 0;
          sh_ga[i] = ga;
          if (!surf) sh_gb[i] = gb;
          sh_s0[i] = make_float4(sel.x, sel.y, sel.z, c.slack);
          if (st.odom_trace) {
            int* tr = st.odom_trace + ((((size_t)s * 2 + STAGE) * 5 + iter / 5) * p.cap_flat + i) * 3;
            tr[0] = c.closest; tr[1] = c.ind2; tr[2] = surf ? c.ind3 : -1;
          }
        }
      }
      __syncthreads();
      STAGE_CLOCK(2);
    }
    for (int i = threadIdx.x; i < n; i += LM_THREADS) {
      if (!sh_ok[i]) continue;
      const float4 ori = sh_ori[i];
      const float4 sel = transform_to_start(ori, T);
      const Row3 r = surf ? surf_row(ori, sel, sh_ga[i], ks, iter) : corner_row(ori, sel, sh_ga[i], sh_gb[i], kc, iter);
      if (r.ok) {
        const double a0 = r.a0, a1 = r.a1, a2 = r.a2, b = r.b;
        acc[0] += a0 * a0; acc[1] += a0 * a1; acc[2] += a0 * a2;
        acc[3] += a1 * a1; acc[4] += a1 * a2; acc[5] += a2 * a2;
        acc[6] += a0 * b; acc[7] += a1 * b; acc[8] += a2 * b;
        acc[9] += 1.0;
      }
    }
    STAGE_CLOCK(6);
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = warp_sum_d(acc[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 10; ++k) sh.part[wid][k] = acc[k];
    }
    __syncthreads();
    STAGE_CLOCK(3);
    if (wid == 0) {
      double v = 0.0;
      if (lane < 10)
        for (int w = 0; w < LM_WARPS; ++w) v += sh.part[w][lane];
      double tot[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) tot[k] = __shfl_sync(0xffffffffu, v, k);
      if (lane == 0) {
        sh.iters = iter + 1;
        const int rows = (int)tot[9];
        if (rows >= 10) {  // featureAssociation.cpp:1222,1232
          float AtA[9] = {(float)tot[0], (float)tot[1], (float)tot[2], (float)tot[1], (float)tot[3],
                          (float)tot[4], (float)tot[2], (float)tot[4], (float)tot[5]};
          float AtB[3] = {(float)tot[6], (float)tot[7], (float)tot[8]};
          float X[3];
          llm::colpiv_qr_solve3(AtA, AtB, X);  // == colpiv_qr_solve<3, 3>, bit for bit (tests/csrc/check_qr3.cpp)
          float* matP = st.odom_matP + s * 9;
          if (iter == 0) {
            degenerate = llm::certainly_not_degenerate<3>(AtA, 10.f) ? 0 : (llm::degeneracy_projector<3>(AtA, 10.f, matP) ? 1 : 0);
            st.odom_flags[s * 4 + 0] = degenerate;
          }
          if (degenerate) {
            const float X2[3] = {X[0], X[1], X[2]};
            for (int r = 0; r < 3; ++r) X[r] = matP[r * 3 + 0] * X2[0] + matP[r * 3 + 1] * X2[1] + matP[r * 3 + 2] * X2[2];
          }
          if (surf) {
            sh.T[0] += X[0]; sh.T[2] += X[1]; sh.T[4] += X[2];
          } else {
            sh.T[1] += X[0]; sh.T[3] += X[1]; sh.T[5] += X[2];
          }
          for (int k = 0; k < 6; ++k)
            if (sh.T[k] != sh.T[k]) sh.T[k] = 0;
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
          if ((double)deltaR < 0.1 && (double)deltaT < 0.1) sh.stop = 1;
        }
      }
    }
    __syncthreads();
    STAGE_CLOCK(4);
    if (sh.stop) break;
  }
  if (threadIdx.x == 0) {
    long long* out = st.stage_clocks + (size_t)s * 16 + STAGE * 8;
    out[0] = global_ns() - t_begin;
    for (int k = 1; k < 8; ++k) out[k] = clk[k];
  }
  if (threadIdx.x < 6) st.transform_cur[s * 6 + threadIdx.x] = sh.T[threadIdx.x];
  if (threadIdx.x == 0) st.odom_iters[s * 2 + STAGE] = sh.iters;
  if (!surf) {
    __syncthreads();  // transform_cur of this sequence is complete (same block)
    if (threadIdx.x == 0) integrate_transformation(st, s);
  }
}

}  // namespace

static size_t stage_smem(const DevParams& p, bool surf) {
  const size_t cap = surf ? p.cap_flat : p.cap_sharp;
  return cap * (surf ? 3 : 4) * sizeof(float4) + cap * 4 + cap * 2 + ((cap + 15) / 16) * 16;
}

void launch_odometry(LaunchCtx& ctx, DevState& st, bool first_frame) {
  const DevParams& p = st.p;
  if (!first_frame) {
    if (st.odom_trace) cudaMemsetAsync(st.odom_trace, 0xff, (size_t)p.B * 2 * 5 * p.cap_flat * 3 * sizeof(int), ctx.stream);
    // the opt-in shared-memory size is a per-device attribute of the kernel
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
      cudaFuncSetAttribute(k_odom_stage<STAGE_SURF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_odom_stage<STAGE_CORNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_done[dev] = true;
    }
    // surf stage then corner stage (featureAssociation.cpp:1216-1234), then integrateTransformation
    LL_LAUNCH(ctx, "k_odom_search_surf", k_odom_search<STAGE_SURF><<<dim3((p.cap_flat + SEARCH_GROUPS - 1) / SEARCH_GROUPS, p.B), SEARCH_THREADS, 0, ctx.stream>>>(st));
    LL_LAUNCH(ctx, "k_odom_stage_surf", k_odom_stage<STAGE_SURF><<<p.B, LM_THREADS, stage_smem(p, true), ctx.stream>>>(st));
    LL_LAUNCH(ctx, "k_odom_search_corner", k_odom_search<STAGE_CORNER><<<dim3((p.cap_sharp + SEARCH_GROUPS * SEARCH_QPW_CORNER - 1) / (SEARCH_GROUPS * SEARCH_QPW_CORNER), p.B), SEARCH_THREADS, 0, ctx.stream>>>(st));
    LL_LAUNCH(ctx, "k_odom_stage_corner", k_odom_stage<STAGE_CORNER><<<p.B, LM_THREADS, stage_smem(p, false), ctx.stream>>>(st));
  }
  if (ctx.wait_before_publish) {
    cudaStreamWaitEvent(ctx.stream, ctx.wait_before_publish, 0);
    ctx.wait_before_publish = nullptr;
  }
  LL_LAUNCH(ctx, "k_publish_clouds_last", k_publish_clouds_last<<<dim3((p.N + 255) / 256, p.B), 256, 0, ctx.stream>>>(st, first_frame ? 1 : 0));
  launch_grid_build2(ctx, p.B, st.grid_corner_last, st.corner_last, p.cap_less_sharp, st.last_counts, 2, 0,
                     st.grid_surf_last, st.surf_last, p.N, st.last_counts, 2, 1, st.odom_flags + 2, 4, /*pack_ring=*/true);
}
