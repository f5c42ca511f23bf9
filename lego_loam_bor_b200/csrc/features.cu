// features.cu -- FeatureAssociation::adjustDistortion / calculateSmoothness / markOccludedPoints /
// extractFeatures (reference: LeGO-LOAM/src/featureAssociation.cpp:161-383).
//
//   (cloudSmoothness is not materialised: see k_feature_sort)
//   k_feature_prep     one thread per segmented point.  adjustDistortion with the sequential
//       halfPassed flag resolved from the first-trigger index found by k_seg_emit; the 11-tap
//       curvature stencil over the FLATTENED range array from a shared-memory tile with a 6-point
//       halo; markOccludedPoints turned into a gather (all its writes are idempotent "= 1").
//       The persistent arrays (cloudCurvature / cloudNeighborPicked / cloudLabel /
//       cloudSmoothness) are only rewritten on [5, S-5) like the reference, stale values elsewhere
//       survive across frames.
//   k_feature_ring     one block per (ring, sequence).  Sorts the six sextants of the ring in shared
//       memory (bitonic, key = (curvature, index)), runs the two greedy pick scans with one warp
//       (32 sorted candidates tested per step), collects the less-flat points and applies the
//       per-ring 0.2 m VoxelGrid (sort by voxel index, sequential centroid per voxel) -- all on the
//       ring's span held in shared memory.
//   k_feature_compact  one block per (ring, sequence): concatenates the per-ring results in ring order.
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

#define FP_THREADS 256
#define FP_HALO 8

__global__ void __launch_bounds__(FP_THREADS) k_feature_prep(DevState st) {
  __shared__ float sh_r[FP_THREADS + 2 * FP_HALO];
  __shared__ int sh_c[FP_THREADS + 2 * FP_HALO];
  __shared__ unsigned char sh_f[FP_THREADS + 2 * FP_HALO];
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  const int S = st.seg_count[s];
  const int b0 = blockIdx.x * FP_THREADS;
  if (b0 >= S) return;
  const size_t base = (size_t)s * p.N;
  for (int t = threadIdx.x; t < FP_THREADS + 2 * FP_HALO; t += FP_THREADS) {
    const int idx = b0 - FP_HALO + t;
    const bool ok = idx >= 0 && idx < S;
    sh_r[t] = ok ? st.seg_range[base + idx] : 0.f;
    sh_c[t] = ok ? (int)st.seg_col[base + idx] : 0;
  }
  __syncthreads();
  // markOccludedPoints (featureAssociation.cpp:226-262), step 1: the depth test of every neighbouring pair (k, k+1),
  // evaluated once per pair: 1 = k is the far side (marks k-5 .. k), 2 = k+1 is the far side (marks k+1 .. k+6)
  for (int t = threadIdx.x; t < FP_THREADS + 2 * FP_HALO - 1; t += FP_THREADS) {
    const int k = b0 - FP_HALO + t;  // loop index of the reference
    unsigned char f = 0;
    if (k >= 5 && k < S - 6) {
      const float depth1 = sh_r[t], depth2 = sh_r[t + 1];
      const int cdiff = abs(sh_c[t + 1] - sh_c[t]);
      if (cdiff < 10) {
        const bool far_near = (double)(depth1 - depth2) > 0.3;
        const bool near_far = !far_near && (double)(depth2 - depth1) > 0.3;
        f = far_near ? 1 : (near_far ? 2 : 0);
      }
    }
    sh_f[t] = f;
  }
  __syncthreads();
  const int i = b0 + threadIdx.x;
  if (i >= S) return;
  // ---- adjustDistortion (featureAssociation.cpp:161-197) ----
  {
    const float start_ori = st.orientation[s * 4 + 0];
    const float end_ori = st.orientation[s * 4 + 1];
    const float ori_diff = st.orientation[s * 4 + 2];
    const int half = st.half_idx[s];
    const float4 sp = st.seg_cloud[base + i];
    const float px = sp.y, py = sp.z, pz = sp.x;
    // -atan2(point.x, point.z) of the axis-swapped point = -atan2(seg.y, seg.x): k_seg_emit needed the same value to find
    // the halfPassed trigger and left it in seg_ori (a double-precision atan2 per point is most of this kernel otherwise)
    float ori = st.seg_ori[base + i];
    if (i <= half) {
      if ((double)ori < (double)start_ori - LL_PI / 2)
        ori = (float)((double)ori + 2 * LL_PI);
      else if ((double)ori > (double)start_ori + LL_PI * 3 / 2)
        ori = (float)((double)ori - 2 * LL_PI);
    } else {
      ori = (float)((double)ori + 2 * LL_PI);
      if ((double)ori < (double)end_ori - LL_PI * 3 / 2)
        ori = (float)((double)ori + 2 * LL_PI);
      else if ((double)ori > (double)end_ori + LL_PI / 2)
        ori = (float)((double)ori - 2 * LL_PI);
    }
    const float rel = (ori - start_ori) / ori_diff;
    const float inten = (float)(int)sp.w + p.scan_period * rel;
    st.seg_cloud[base + i] = make_float4(px, py, pz, inten);
  }
  const float* r = sh_r + FP_HALO + threadIdx.x;  // r[k] = range of point i + k
  const int* c = sh_c + FP_HALO + threadIdx.x;
  const bool interior = (i >= 5 && i < S - 5);
  // ---- markOccludedPoints, step 2: gather (all its writes are idempotent "= 1") ----
  // picked[i] is written by the far-side-left test of pairs i .. i+5 and by the far-side-right test of pairs i-6 .. i-1
  const unsigned char* f = sh_f + FP_HALO + threadIdx.x;
  bool mark = (((f[0] | f[1] | f[2] | f[3] | f[4] | f[5]) & 1) | ((f[-6] | f[-5] | f[-4] | f[-3] | f[-2] | f[-1]) & 2)) != 0;
  if (i >= 5 && i < S - 6) {
    const float diff1 = fabsf(r[-1] - r[0]);
    const float diff2 = fabsf(r[1] - r[0]);
    if ((double)diff1 > 0.02 * (double)r[0] && (double)diff2 > 0.02 * (double)r[0]) mark = true;
  }
  if (interior) {
    // ---- calculateSmoothness (featureAssociation.cpp:200-223), strictly left to right ----
    const float d = r[-5] + r[-4] + r[-3] + r[-2] + r[-1] - r[0] * 10 + r[1] + r[2] + r[3] + r[4] + r[5];
    const float curv = d * d;
    st.curvature[base + i] = curv;
    st.picked[base + i] = mark ? 1 : 0;
    st.cloud_label[base + i] = 0;
  } else if (mark) {
    st.picked[base + i] = 1;  // outside [5, S-5) the array is never reset
  }
}

// ---------------------------------------------------------------------------------------------

// Direction-free bitonic network, ascending, over every aligned block of n keys of key[0 .. total) in shared memory
// (n a power of two, total a multiple of n, blockDim.x a multiple of 32).  One compare-exchange per loop trip: pair
// t owns the elements i (t with a zero bit inserted at the stride) and its partner, so no thread idles on the upper
// half of a pair.  All strides <= 32 of a merge phase touch one aligned 64-key block per 32 consecutive pairs, i.e.
// per warp, so they run back to back with warp-level synchronisation only; a block-wide barrier is needed only
// around strides >= 64.
template <typename T>
__device__ __forceinline__ void block_bitonic_sort(T* key, int n, int total) {
  const int half = total >> 1;
  const int lane = threadIdx.x & 31;
  for (int k = 2; k <= n; k <<= 1) {
    int j = k >> 1;
    for (; j > 32; j >>= 1) {
      const int jm = j - 1;
      const bool flip = (j == (k >> 1));
      for (int t = threadIdx.x; t < half; t += blockDim.x) {
        const int i = ((t & ~jm) << 1) | (t & jm);
        const int l = flip ? (i ^ (k - 1)) : (i | j);
        const T a = key[i], b = key[l];
        if (a > b) { key[i] = b; key[l] = a; }
      }
      __syncthreads();
    }
    for (int t0 = threadIdx.x - lane; t0 < half; t0 += blockDim.x) {  // warp-uniform trip count
      const int t = t0 + lane;
      for (int jj = j; jj > 0; jj >>= 1) {
        if (t < half) {
          const int jm = jj - 1;
          const int i = ((t & ~jm) << 1) | (t & jm);
          const int l = (jj == (k >> 1)) ? (i ^ (k - 1)) : (i | jj);
          const T a = key[i], b = key[l];
          if (a > b) { key[i] = b; key[l] = a; }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* key, int n /* pow2 */, int total /* multiple of n */) {
  block_bitonic_sort<unsigned long long>(key, n, total);
}

template <typename T>
__device__ __forceinline__ void bitonic_sort_t(T* key, int n /* pow2 */) {
  block_bitonic_sort<T>(key, n, n);
}

// order-preserving float <-> int map (an involution), for shared-memory atomicMin/Max on floats
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

#define FR_THREADS 256
#define FR_WARPS (FR_THREADS / 32)
#define FR_RADIX_MIN 512    // voxel runs of a ring: above this many, a stable radix sort instead of the bitonic network
#define FR_RADIX_MAX 2048   // 8 warps x 8 trips of 32

struct RingSortSmem {
  unsigned short whist[FR_WARPS][256];
  unsigned base[256];
};

// Stable LSD radix sort of n <= FR_RADIX_MAX (key, value) pairs in shared memory by the low `bits` bits of the key, 8 bits a
// pass, ping-pong between (k0, v0) and (k1, v1); returns the buffer (0 / 1) that holds the result.  Every warp owns a
// contiguous chunk of the input and ranks it 32 keys at a time (match_any: peers of equal digit, in lane order), carrying
// its own digit counters from trip to trip, so a pass needs no barrier per trip: count, one scan over (digit, warp), scatter.
// Called by all FR_THREADS threads; ends with a barrier.
__device__ __noinline__ int ring_radix_sort(unsigned* k0, unsigned short* v0, unsigned* k1, unsigned short* v1, int n, int bits,
                                               RingSortSmem& sm, int* warp_tot) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int chunk = (((n + FR_WARPS - 1) / FR_WARPS) + 31) & ~31;   // keys per warp: a multiple of 32, <= 256
  const int c0 = wid * chunk;
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    const unsigned* kin = cur ? k1 : k0;
    const unsigned short* vin = cur ? v1 : v0;
    unsigned* kout = cur ? k0 : k1;
    unsigned short* vout = cur ? v0 : v1;
    for (int d = threadIdx.x; d < FR_WARPS * 256; d += FR_THREADS) (&sm.whist[0][0])[d] = 0;
    __syncthreads();
    unsigned rpack[4] = {0u, 0u, 0u, 0u};   // rank of this lane's key of trip t inside the warp's chunk and digit, 16 bits each
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t * 32 < chunk) {   // warp-uniform
        const int i = c0 + t * 32 + lane;
        const bool act = i < n;
        const unsigned amask = __ballot_sync(0xffffffffu, act);
        unsigned dg = 0u, peers = 0u, prev = 0u, r = 0u;
        if (act) {
          dg = (kin[i] >> shift) & 255u;
          peers = __match_any_sync(amask, dg);
          r = __popc(peers & ((1u << lane) - 1u));
          prev = sm.whist[wid][dg];
        }
        __syncwarp();
        if (act && r == 0u) sm.whist[wid][dg] = (unsigned short)(prev + __popc(peers));
        __syncwarp();
        rpack[t >> 1] |= (prev + r) << (16 * (t & 1));
      }
    }
    __syncthreads();
    {
      // per digit: exclusive scan over the warps, then over the digits
      unsigned run = 0u;
#pragma unroll
      for (int w = 0; w < FR_WARPS; ++w) {
        const unsigned c = sm.whist[w][threadIdx.x];
        sm.whist[w][threadIdx.x] = (unsigned short)run;
        run += c;
      }
      int total;
      sm.base[threadIdx.x] = (unsigned)block_exclusive_scan((int)run, warp_tot, &total);   // FR_THREADS == 256 digits
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t * 32 < chunk) {
        const int i = c0 + t * 32 + lane;
        if (i < n) {
          const unsigned k = kin[i];
          const unsigned dg = (k >> shift) & 255u;
          const unsigned pos = sm.base[dg] + sm.whist[wid][dg] + ((rpack[t >> 1] >> (16 * (t & 1))) & 0xffffu);
          kout[pos] = k;
          vout[pos] = vin[i];
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}

// ---- per-ring feature extraction: candidate sort, greedy picks, less-flat collection + VoxelGrid, in ONE block ---------
// One block per (ring, sequence); everything between the loads of the ring's span and its results lives in shared
// memory (extractFeatures, featureAssociation.cpp:265-383).
//
// 1. Candidate sort.  The reference sorts each sextant of cloudSmoothness by curvature (std::sort(begin + sp,
//    begin + ep), :275-286) and then scans it downwards for edge points and upwards for flat points.  Only elements that
//    can pass the tests of :291-293 / :333-335 (curvature vs threshold, ground flag) can ever be picked, so only those
//    are sorted, all twelve lists of the ring in ONE bitonic network with the composite key
//      [list = sextant*2 + (0 edge | 1 flat)] [visit order] [index tie-break]
//    where the visit order is ~curvature for edge lists (descending scan), curvature for flat lists, and the unsorted
//    element at position ep (the reference's sort excludes ep but both scans include it)
//    gets the first / last slot.  cloudSmoothness itself is not stored: for k in [5, S-5) its entry is
//    (curvature[k], k) by construction (:220-221); the one stale entry that can ever be read, position 4 (SURVEY.md
//    section 9 item 7), is carried in `slot4` and updated to the minimum of the sorted range that contains it, which is
//    what the reference's in-place sort leaves there.
// 2. Greedy picks by warp 0; the six sextants of a ring must run in order because a pick suppresses up to 5 neighbours on
//    either side, across sextant boundaries.  Each step tests 32 consecutive candidates of the sorted list at once; the
//    only mutable state, cloudNeighborPicked and cloudLabel of the ring's span, are bytes in shared memory.  Rings do not
//    interact: candidates lie in [start, end], their +-5 neighbours inside the ring's own points [start-5, end+5].
// 3. Less-flat collection by POSITION (:370-374) and the per-ring 0.2 m pcl::VoxelGrid (:101,377-381): consecutive points
//    of a ring mostly share a voxel, so the sort runs over RUNS of equal voxel index; runs of the same voxel end up
//    adjacent and in input order, so each voxel's float sums keep the input order.
__global__ void __launch_bounds__(FR_THREADS, 5) k_feature_ring(DevState st, int cap2) {
  extern __shared__ unsigned long long fr_keys[];  // [cap2] u64, then int vox[H], u16 lfpos / col [H+32] each, u8 picked, i8 label [H+32] each
  __shared__ int warp_tot[33];
  __shared__ RingSortSmem sh_sort;
  __shared__ int sh_sp[6], sh_ep[6];
  __shared__ int sh_imn[3], sh_imx[3];
  __shared__ int sh_stale, sh_has_stale;
  __shared__ int sh_counts[3];
  __shared__ unsigned short sh_pick[6][24];   // per sextant: up to 20 edge picks, then up to 4 flat picks (positions from the ring's first)
  __shared__ int sh_ne[6], sh_nf[6], sh_spill_r[6], sh_spill_l[6];
  __shared__ unsigned long long sh_slot_min;
  const DevParams& p = st.p;
  const int s = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // Longest ring first: a block lasts about as long as its ring has points (10 .. 130 us), and the launch lasts as long as
  // its last block.  Blocks are dispatched in grid order (x = sequence fastest), so the block with blockIdx.y = r takes the
  // ring with the r-th largest span of its sequence: the long rings of all sequences start at once, the short ones fill in.
  __shared__ int sh_span[LL_MAX_RINGS];
  __shared__ int sh_ring;
  for (int r = threadIdx.x; r < p.V; r += FR_THREADS) sh_span[r] = st.end_ring[s * p.V + r] - st.start_ring[s * p.V + r];
  __syncthreads();
  for (int r = threadIdx.x; r < p.V; r += FR_THREADS) {
    const int mine = sh_span[r];
    int rank = 0;
    for (int q = 0; q < p.V; ++q) rank += (sh_span[q] > mine || (sh_span[q] == mine && q < r)) ? 1 : 0;
    if (rank == (int)blockIdx.y) sh_ring = r;
  }
  __syncthreads();
  const int ring = sh_ring;
  const size_t base = (size_t)s * p.N;
  const size_t rs = (size_t)s * p.V + ring;
  const int HP = p.H + 32;
  unsigned long long* keys = fr_keys;
  int* sm_vox = reinterpret_cast<int*>(fr_keys + cap2);
  // (col, picked and label are dead once the less-flat positions are collected: 4 (H + 32) contiguous bytes, the second
  // key buffer of the radix sort of the voxel runs)
  unsigned short* sm_lfpos = reinterpret_cast<unsigned short*>(sm_vox + p.H);
  unsigned short* sm_col = sm_lfpos + HP;
  unsigned char* sm_picked = reinterpret_cast<unsigned char*>(sm_col + HP);
  signed char* sm_label = reinterpret_cast<signed char*>(sm_picked + HP);
  // LL_BUF_RING_CLOCKS (profiling aid, thread 0): SM cycles per phase of this block (every slot is passed at most once)
  long long* ring_clk = st.ring_clocks + rs * 10;
  if (threadIdx.x < 10) ring_clk[threadIdx.x] = 0;
  const long long rc_begin = clock64();
  long long rc_mark = rc_begin;
#define RING_CLOCK(slot) do { if (threadIdx.x == 0) { const long long t__ = clock64(); ring_clk[slot] = t__ - rc_mark; rc_mark = t__; } } while (0)
  const int start = st.start_ring[s * p.V + ring], end = st.end_ring[s * p.V + ring];
  const int a = start - 4, b = end + 6;  // this ring's points are [a - 1, b)
  const int span_lo = max(0, a - 8), span_hi = min(p.N, b + 8);
  const int L = max(0, span_hi - span_lo);
  if (threadIdx.x < 6) {
    const int j = threadIdx.x;
    sh_sp[j] = (start * (6 - j) + end * j) / 6;
    sh_ep[j] = (start * (5 - j) + end * (j + 1)) / 6 - 1;
  }
  if (threadIdx.x == 0) { sh_slot_min = ~0ull; sh_stale = 0; sh_has_stale = 0; }
  for (int t = threadIdx.x; t < L; t += FR_THREADS) {
    sm_picked[t] = (unsigned char)(st.picked[base + span_lo + t] != 0);
    sm_col[t] = (unsigned short)st.seg_col[base + span_lo + t];
    sm_label[t] = (signed char)max(-2, min(3, st.cloud_label[base + span_lo + t]));
  }
  __syncthreads();
  // ---- 1. one key per position of the ring ----
  // The reference sorts every sextant by curvature and scans it downwards for edge points (first unpicked candidate,
  // 20 times) and upwards for flat points (4 times).  "The next unpicked candidate in sorted order" is the arg-max (arg-min)
  // over the candidates that are still unpicked, so nothing is sorted here: a warp keeps the sextant's candidates spread
  // over its lanes and selects with two hardware warp reductions per pick.
  //   ckey[u]  edge list: curvature bits (position ep: 0xffffffff, scanned first); flat list: 0xfffffffe - curvature bits
  //            (position ep: 1, scanned last); 0: not a candidate.  ctype[u]: 0 edge, 1 flat, 3 neither.
  const unsigned slot_val = st.slot4[s * 2 + 0];
  const int slot_ind = (int)st.slot4[s * 2 + 1];
  const int k_first = sh_sp[0];
  unsigned* ckey = reinterpret_cast<unsigned*>(keys);             // [H]
  unsigned char* ctype = reinterpret_cast<unsigned char*>(ckey + p.H);  // [H]
  for (int k0 = k_first; k0 <= sh_ep[5]; k0 += FR_THREADS) {
    const int k = k0 + threadIdx.x;
    if (k > sh_ep[5]) break;
    int j = 0;
#pragma unroll
    for (int q = 1; q < 6; ++q) j += (k >= sh_sp[q]) ? 1 : 0;   // sextant ranges tile [sp_0, ep_5]: sp_{q+1} = ep_q + 1
    const int sp = sh_sp[j], ep = sh_ep[j];
    unsigned key = 0u;
    unsigned char type = 3;
    if (sp < ep && k >= 0 && k < p.N) {
      unsigned vb;
      int ind;
      if (k == 4) { vb = slot_val; ind = slot_ind; }
      else { vb = __float_as_uint(st.curvature[base + k]); ind = k; }
      if (sp == 4 && k < ep) atomicMin(&sh_slot_min, ((unsigned long long)vb << 32) | (unsigned)ind);
      const float c = (k == 4) ? st.curvature[base + ind] : __uint_as_float(vb);
      const bool ground = st.seg_ground[base + ind] != 0;
      const bool edge_ok = c > p.edge_threshold && !ground;   // featureAssociation.cpp:291-293
      const bool flat_ok = c < p.surf_threshold && ground;    // featureAssociation.cpp:333-335
      if (edge_ok) { type = 0; key = (k == ep) ? 0xffffffffu : vb; }
      else if (flat_ok) { type = 1; key = (k == ep) ? 1u : 0xfffffffeu - vb; }
      if (k == 4 && (edge_ok || flat_ok)) { sh_stale = ind; sh_has_stale = 1; }
    }
    ckey[k - k_first] = key;
    ctype[k - k_first] = type;
  }
  bool slot_in_sorted_range = false;
#pragma unroll
  for (int j = 0; j < 6; ++j) slot_in_sorted_range = slot_in_sorted_range || (sh_sp[j] < sh_ep[j] && sh_sp[j] == 4);
  __syncthreads();
  if (threadIdx.x == 0 && slot_in_sorted_range && sh_slot_min != ~0ull) {
    st.slot4[s * 2 + 0] = (unsigned)(sh_slot_min >> 32);
    st.slot4[s * 2 + 1] = (unsigned)(sh_slot_min & 0xffffffffull);
  }
  RING_CLOCK(1);
  // ---- 2. greedy picks ----
  // The six sextants of a ring run in order in the reference because a pick suppresses up to 5 neighbours on either side,
  // across a sextant boundary.  Only the first five positions of a sextant can be reached from the previous one, so the
  // sextants are picked SPECULATIVELY in parallel, one warp each, on the picked flags of their own range; marks that fall
  // outside the own range are kept as two 5-bit spill masks.  Then the boundaries are checked in order: if none of the
  // positions sextant j-1 (final) marks at the start of sextant j was picked by sextant j's speculative run, that run is
  // exactly what the sequential scan does (a candidate that was never picked never influenced a decision); otherwise
  // sextant j is reset and run again with those marks in place.  Ring 0 (the stale cloudSmoothness entry may point
  // anywhere) and rings with sextants shorter than 6 points take the sequential path below.
  int n_rounds_dbg = 0;
  bool speculative = sh_has_stale == 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) speculative = speculative && (sh_sp[j] >= sh_ep[j] ? false : (sh_ep[j] - sh_sp[j] + 1 >= 6));
  if (speculative) {
    // one sextant by one warp; own range [own_lo, own_hi]: marks inside go to sm_picked, marks outside to the spill masks
    auto run_sextant = [&](int j, unsigned premark) {
      const int sp = sh_sp[j], ep = sh_ep[j];
      const int len = ep - sp + 1;
      const unsigned* kq = ckey + (sp - k_first);
      const unsigned char* tq = ctype + (sp - k_first);
      const int items = (len + 31) >> 5;
      unsigned spill_r = 0u, spill_l = 0u;
      int n_edge = 0, n_flat = 0;
      if (lane < 5 && ((premark >> lane) & 1u)) sm_picked[sp + lane - span_lo] = 1;
      __syncwarp();
      for (int type = 0; type < 2; ++type) {
        const int quota = type == 0 ? 20 : 4;
        unsigned elig = 0u;
        for (int i = 0; i < items; ++i) {
          const int u = lane + 32 * i;
          if (u < len && tq[u] == type && sm_picked[sp + u - span_lo] == 0) elig |= 1u << i;
        }
        int done = 0;
        while (done < quota) {
          unsigned bkey = 0u;
          int bu = -1;
          unsigned e = elig;
          while (e) {
            const int i = __ffs(e) - 1;
            e &= e - 1;
            const int u = lane + 32 * i;
            const unsigned kv = kq[u];
            // equal keys: the edge scan meets the higher index first, the flat scan the lower one
            if (kv > bkey || (kv == bkey && type == 0)) { bkey = kv; bu = u; }
          }
          const unsigned top = __reduce_max_sync(0xffffffffu, bu >= 0 ? bkey : 0u);
          if (top == 0u) break;
          const int myind = (bu >= 0 && bkey == top) ? sp + bu : -1;
          int pick;
          if (type == 0) pick = (int)__reduce_max_sync(0xffffffffu, (unsigned)(myind + 1)) - 1;
          else pick = (int)__reduce_min_sync(0xffffffffu, myind >= 0 ? (unsigned)myind : 0xffffffffu);
          ++done;
          if (lane == 0) {
            sh_pick[j][(type == 0 ? 0 : 20) + done - 1] = (unsigned short)(pick - k_first);
            sm_label[pick - span_lo] = (signed char)(type == 0 ? (done <= 2 ? 2 : 1) : -1);
          }
          if (type == 0) n_edge = done; else n_flat = done;
          if (type == 1 && done >= 4) break;  // featureAssociation.cpp:339-342: the fourth flat pick suppresses nothing
          // featureAssociation.cpp:306-326 / 344-366: lane 0 the pick, lanes 1..5 ind+1..ind+5, lanes 6..10 ind-1..ind-5
          const int off = lane <= 5 ? lane : 5 - lane;
          const int g = pick + off;
          const int l = g - span_lo;
          bool bad = false;
          // l < 0 only for ind - k < 0 at the very start of the cloud: the reference `continue`s there (:317)
          if (lane >= 1 && lane <= 10 && l >= 0) bad = abs((int)sm_col[l] - (int)sm_col[l + (lane <= 5 ? -1 : 1)]) > 10;
          const unsigned bm = __ballot_sync(0xffffffffu, bad);
          const unsigned fwd = bm & 0x3eu, bwd = bm & 0x7c0u;
          const int stop_f = fwd ? __ffs(fwd) - 1 : 32, stop_b = bwd ? __ffs(bwd) - 1 : 32;
          const bool mark = l >= 0 && (lane == 0 || (lane >= 1 && lane <= 5 && lane < stop_f) || (lane >= 6 && lane <= 10 && lane < stop_b));
          if (mark && g >= sp && g <= ep) sm_picked[l] = 1;
          spill_r |= __reduce_or_sync(0xffffffffu, (mark && g > ep) ? 1u << (g - ep - 1) : 0u);
          spill_l |= __reduce_or_sync(0xffffffffu, (mark && g < sp) ? 1u << (sp - 1 - g) : 0u);
          __syncwarp();
          // the positions pick-5 .. pick+5 may have been marked: each belongs to exactly one lane
          const int lo = max(pick - 5, sp), hi = min(pick + 5, ep);
          const int first = lo + ((lane - (lo - sp)) & 31);
          if (first <= hi && sm_picked[first - span_lo]) elig &= ~(1u << ((first - sp) >> 5));
        }
      }
      if (lane == 0) { sh_ne[j] = n_edge; sh_nf[j] = n_flat; sh_spill_r[j] = spill_r; sh_spill_l[j] = spill_l; }
    };
    // One pick of the predecessor is known in advance: the reference's sort leaves out position ep but both scans include
    // it, so an edge candidate at ep is met FIRST by the descending scan (key 0xffffffff) and, unless it is occluded, picked
    // whatever else happens -- nothing of its own sextant has been picked yet and the marks of the sextant before reach
    // only five positions, less than a speculative sextant is long.  Its marks on the first positions of this sextant are
    // applied before the speculative run; without them about every fourth boundary of a ring of noisy far points (nearly
    // every point an edge candidate) forced a re-run.
    unsigned pre = 0u;
    if (wid >= 1 && wid < 6) {
      const int e = sh_ep[wid - 1];
      const int le = e - span_lo;
      const bool elig_e = ctype[e - k_first] == 0 && ckey[e - k_first] == 0xffffffffu && sm_picked[le] == 0;
      bool bad = false;
      if (lane >= 1 && lane <= 5) bad = abs((int)sm_col[le + lane] - (int)sm_col[le + lane - 1]) > 10;   // featureAssociation.cpp:306-316
      const unsigned bm = __ballot_sync(0xffffffffu, bad) & 0x3eu;
      const int stop = bm ? __ffs(bm) - 1 : 6;   // ep + stop is the first position the marking loop does not reach
      if (elig_e) pre = (1u << (stop - 1)) - 1u;
    }
    if (wid < 6) run_sextant(wid, pre);
    __syncthreads();
    RING_CLOCK(2);
    // Boundaries.  Sextant j ran with the marks `used` of its predecessor in place (none, at first); the predecessor's
    // run says `r`.  The run stands if r == used, or if r only ADDS marks at positions sextant j never picked (a candidate
    // that was never picked never influenced a decision); otherwise sextant j is reset and run again with r in place.
    // All sextants decide on the same snapshot of the spill masks and re-run in parallel, a warp each; sextants 0 .. k are
    // final after round k (sextant 0 has no predecessor), so this ends after at most five rounds -- usually one.
    {
      unsigned used = pre;   // warp j: the marks its last run started with
      for (int round = 0; round < 6; ++round) {
        n_rounds_dbg = round;
        bool rerun = false;
        unsigned r = 0u;
        if (wid >= 1 && wid < 6) {
          const int j = wid;
          r = (unsigned)sh_spill_r[j - 1];
          if (r != used) {
            bool hit = (used & ~r) != 0u;   // a mark this run assumed has gone: a position it avoided may be free again
            if (lane < 24) {
              const bool mine = lane < 20 ? lane < sh_ne[j] : (lane - 20) < sh_nf[j];
              if (mine) {
                const int q = k_first + (int)sh_pick[j][lane] - sh_sp[j];
                hit = hit || (q < 5 && (((r & ~used) >> q) & 1u));
              }
            }
            rerun = __any_sync(0xffffffffu, hit);
            if (!rerun) used = r;   // the added marks change nothing for this run; they reach sm_picked with the spills below
          }
        }
        if (!__syncthreads_or(rerun ? 1 : 0)) break;   // (also orders the reads of the spill masks before the re-runs write them)
        if (rerun) {
          const int j = wid;
          // reset sextant j to the state before any pick (picked flags from global memory, labels 0: its candidates lie
          // in [5, S-5), where calculateSmoothness has just reset cloudLabel) and run it with the marks in place
          for (int t = sh_sp[j] + lane; t <= sh_ep[j]; t += 32) sm_picked[t - span_lo] = (unsigned char)(st.picked[base + t] != 0);
          if (lane < 24) {
            const bool mine = lane < 20 ? lane < sh_ne[j] : (lane - 20) < sh_nf[j];
            if (mine) sm_label[k_first + (int)sh_pick[j][lane] - span_lo] = 0;
          }
          // marks sextant j itself put into its predecessor's range are dropped with the run; the new run records its own
          __syncwarp();
          run_sextant(j, r);
          used = r;
        }
        __syncthreads();
      }
    }
    __syncthreads();
    // spills into the neighbours' ranges and the ring's margins (cloudNeighborPicked is only ever set), outputs in scan order
    if (threadIdx.x < 60) {
      const int j = threadIdx.x / 10, i = threadIdx.x % 10;
      const int sp = sh_sp[j], ep = sh_ep[j];
      if (i < 5) { if ((sh_spill_r[j] >> i) & 1) { const int g = ep + 1 + i; if (g >= span_lo && g < span_hi) sm_picked[g - span_lo] = 1; } }
      else { if ((sh_spill_l[j] >> (i - 5)) & 1) { const int g = sp - 1 - (i - 5); if (g >= span_lo && g < span_hi) sm_picked[g - span_lo] = 1; } }
    }
    // outputs in scan order: sextant by sextant, edge picks (the first two of a sextant are also the sharp ones), flat picks
    if (threadIdx.x < 6 * 24) {
      const int j = threadIdx.x / 24, t = threadIdx.x % 24;
      int o_sharp = 0, o_lsharp = 0, o_flat = 0;
      for (int q = 0; q < j; ++q) { o_sharp += min(sh_ne[q], 2); o_lsharp += sh_ne[q]; o_flat += sh_nf[q]; }
      if (t < 20) {
        if (t < sh_ne[j]) {
          const int pick = k_first + (int)sh_pick[j][t];
          if (t < 2) st.st_sharp_ind[rs * 12 + o_sharp + t] = pick;
          st.st_less_sharp_ind[rs * 120 + o_lsharp + t] = pick;
        }
      } else if (t - 20 < sh_nf[j]) {
        st.st_flat_ind[rs * 24 + o_flat + (t - 20)] = k_first + (int)sh_pick[j][t];
      }
      if (j == 5 && t == 0) {
        sh_counts[0] = o_sharp + min(sh_ne[5], 2); sh_counts[1] = o_lsharp + sh_ne[5]; sh_counts[2] = o_flat + sh_nf[5];
      }
    }
  } else
  if (wid == 0) {
    int* o_sharp_i = st.st_sharp_ind + rs * 12;
    int* o_lsharp_i = st.st_less_sharp_ind + rs * 120;
    int* o_flat_i = st.st_flat_ind + rs * 24;
    int n_sharp = 0, n_lsharp = 0, n_flat = 0;
    const int colsz = p.N;  // segInfo.segmentedCloudColInd.size()
    const int stale_ind = sh_stale;
    const bool has_stale = sh_has_stale != 0;
    auto in_span = [&](int g) { return g >= span_lo && g < span_hi; };
    auto get_picked = [&](int g) -> int { return in_span(g) ? (int)sm_picked[g - span_lo] : st.picked[base + g]; };
    auto set_picked = [&](int g) { if (in_span(g)) sm_picked[g - span_lo] = 1; else st.picked[base + g] = 1; };
    auto get_col = [&](int g) -> int { return in_span(g) ? (int)sm_col[g - span_lo] : (int)st.seg_col[base + g]; };
    // (general path: the stale entry may point outside this ring's own range, so the label goes to global memory at once)
    auto set_label = [&](int g, int v) { if (in_span(g)) sm_label[g - span_lo] = (signed char)v; st.cloud_label[base + g] = v; };
    // featureAssociation.cpp:306-326 / 344-366: lanes 1..5 handle ind+1..ind+5, lanes 6..10 handle ind-1..ind-5
    auto mark_neighbors = [&](int ind) {
      int g = -1;          // index this lane may mark
      bool bad = false;    // column gap > 10 right before it (stops the loop)
      bool skip = true;    // `continue` cases: out of range
      if (lane >= 1 && lane <= 5) {
        g = ind + lane;
        if (g < colsz) { skip = false; bad = abs(get_col(g) - get_col(g - 1)) > 10; }
      } else if (lane >= 6 && lane <= 10) {
        g = ind - (lane - 5);
        if (g >= 0) { skip = false; bad = abs(get_col(g) - get_col(g + 1)) > 10; }
      }
      const unsigned bm = __ballot_sync(0xffffffffu, bad);
      const unsigned fwd = bm & 0x3eu, bwd = bm & 0x7c0u;
      const int stop_f = fwd ? __ffs(fwd) - 1 : 32, stop_b = bwd ? __ffs(bwd) - 1 : 32;
      if (lane == 0) set_picked(ind);
      if (!skip && lane >= 1 && lane <= 5 && lane < stop_f) set_picked(g);
      if (!skip && lane >= 6 && lane <= 10 && lane < stop_b) set_picked(g);
      __syncwarp();
    };
    // Fast accessors: every candidate of a ring lies in [start, end], so the pick and its +-5 neighbours are inside the
    // span held in shared memory and inside [0, colsz): no range checks, no global fall-back.  Only the stale
    // cloudSmoothness entry of position 4 (ring 0) can point elsewhere; a ring that holds such an entry uses the general
    // accessors above.
    auto mark_neighbors_fast = [&](int ind) {
      const int off = lane <= 5 ? lane : 5 - lane;  // lane 0: the pick; 1..5: ind+1..ind+5; 6..10: ind-1..ind-5
      const int l = ind - span_lo + off;
      bool bad = false;
      // l < 0 only for ind - k < 0 at the very start of the cloud: the reference `continue`s there (:317)
      if (lane >= 1 && lane <= 10 && l >= 0) bad = abs((int)sm_col[l] - (int)sm_col[l + (lane <= 5 ? -1 : 1)]) > 10;
      const unsigned bm = __ballot_sync(0xffffffffu, bad);
      const unsigned fwd = bm & 0x3eu, bwd = bm & 0x7c0u;
      const int stop_f = fwd ? __ffs(fwd) - 1 : 32, stop_b = bwd ? __ffs(bwd) - 1 : 32;
      if (l >= 0 && (lane == 0 || (lane <= 5 && lane < stop_f) || (lane >= 6 && lane <= 10 && lane < stop_b))) sm_picked[l] = 1;
      __syncwarp();
    };
    for (int j = 0; j < 6; ++j) {
      const int sp = sh_sp[j], ep = sh_ep[j];
      if (sp >= ep) continue;
      const int len = ep - sp + 1;
      const unsigned* kq = ckey + (sp - k_first);
      const unsigned char* tq = ctype + (sp - k_first);
      const int items = (len + 31) >> 5;   // candidates per lane: positions sp + lane + 32 i
      for (int type = 0; type < 2; ++type) {  // edge list, then flat list (featureAssociation.cpp:289-330, 331-368)
        const int quota = type == 0 ? 20 : 4;
        // eligibility bits of this lane's candidates (fast path): of this list and not picked yet
        unsigned elig = 0u;
        if (!has_stale)
          for (int i = 0; i < items; ++i) {
            const int u = lane + 32 * i;
            if (u < len && tq[u] == type && sm_picked[sp + u - span_lo] == 0) elig |= 1u << i;
          }
        int done = 0;
        while (done < quota) {
          // this lane's best remaining candidate
          unsigned bkey = 0u;
          int bu = -1;
          if (!has_stale) {
            unsigned e = elig;
            while (e) {
              const int i = __ffs(e) - 1;
              e &= e - 1;
              const int u = lane + 32 * i;
              const unsigned kv = kq[u];
              // equal keys: the edge scan meets the higher index first, the flat scan the lower one
              if (kv > bkey || (kv == bkey && type == 0)) { bkey = kv; bu = u; }
            }
          } else {
            for (int i = 0; i < items; ++i) {
              const int u = lane + 32 * i;
              if (u >= len || tq[u] != type) continue;
              const int ind = (sp + u == 4) ? stale_ind : sp + u;
              if (get_picked(ind) != 0) continue;
              const unsigned kv = kq[u];
              if (bu < 0 || kv > bkey || (kv == bkey && type == 0)) { bkey = kv; bu = u; }
            }
          }
          const unsigned top = __reduce_max_sync(0xffffffffu, bu >= 0 ? bkey : 0u);
          if (top == 0u) break;  // no unpicked candidate left in this list
          // among equal keys: highest index (edge) / lowest index (flat); ties need the candidates' own indices
          const int myind = (bu >= 0 && bkey == top) ? ((sp + bu == 4 && has_stale) ? stale_ind : sp + bu) : -1;
          int pick;
          if (type == 0) pick = (int)__reduce_max_sync(0xffffffffu, (unsigned)(myind + 1)) - 1;
          else pick = (int)__reduce_min_sync(0xffffffffu, myind >= 0 ? (unsigned)myind : 0xffffffffu);
          ++done;
          if (type == 0) {
            if (lane == 0) {
              if (done <= 2) o_sharp_i[n_sharp] = pick;
              o_lsharp_i[n_lsharp] = pick;
            }
            if (done <= 2) n_sharp++;
            n_lsharp++;
          } else {
            if (lane == 0) o_flat_i[n_flat] = pick;
            n_flat++;
          }
          const int lab = type == 0 ? (done <= 2 ? 2 : 1) : -1;
          if (!has_stale) {
            if (lane == 0) sm_label[pick - span_lo] = (signed char)lab;
            if (type == 1 && done >= 4) break;  // featureAssociation.cpp:339-342: the fourth flat pick suppresses nothing
            mark_neighbors_fast(pick);
            // the positions pick-5 .. pick+5 may have been marked: each belongs to exactly one lane
            const int lo = max(pick - 5, sp), hi = min(pick + 5, ep);
            const int first = lo + ((lane - (lo - sp)) & 31);   // the position of this lane inside the window, if any
            if (first <= hi && sm_picked[first - span_lo]) elig &= ~(1u << ((first - sp) >> 5));
          } else {
            if (lane == 0) set_label(pick, lab);
            if (type == 1 && done >= 4) break;
            mark_neighbors(pick);
          }
        }
      }
    }
    if (lane == 0) { sh_counts[0] = n_sharp; sh_counts[1] = n_lsharp; sh_counts[2] = n_flat; }
  }
  __syncthreads();
  RING_CLOCK(3);
  // persist cloudNeighborPicked (entries are only ever set to 1 here) and cloudLabel for this ring's own range
  for (int t = threadIdx.x; t < L; t += FR_THREADS) {
    const int g = span_lo + t;
    if (g >= a - 1 && g < b && sm_picked[t]) st.picked[base + g] = 1;
    if (g >= start && g <= end) st.cloud_label[base + g] = (int)sm_label[t];
  }
  // ---- 3. less-flat collection (featureAssociation.cpp:370-374): by POSITION k over the active sextants ----
  float4* o_lflat = st.st_less_flat + rs * p.H;
  int n_raw = 0;
  const int k_lo = sh_sp[0];
  {
    const int k_hi = sh_ep[5];
    int run = 0;
    for (int k0 = k_lo; k0 <= k_hi; k0 += FR_THREADS) {
      const int k = k0 + threadIdx.x;
      int f = 0;
      if (k <= k_hi && k >= 0 && k < p.N) {
        bool active = false;
#pragma unroll
        for (int j = 0; j < 6; ++j) active = active || (sh_sp[j] < sh_ep[j] && k >= sh_sp[j] && k <= sh_ep[j]);
        const int lab = (k >= span_lo && k < span_hi) ? (int)sm_label[k - span_lo] : st.cloud_label[base + k];
        if (active && lab <= 0) f = 1;
      }
      int total;
      const int ex = block_exclusive_scan(f, warp_tot, &total);
      if (f) sm_lfpos[run + ex] = (unsigned short)(k - k_lo);
      run += total;
    }
    n_raw = run;
  }
  __syncthreads();
  RING_CLOCK(4);
  // pcl::VoxelGrid, leaf 0.2 (featureAssociation.cpp:101,377-381; algorithm: SURVEY.md section 8 f1)
  int n_ds = 0;
  int n_runs_dbg = 0;
  if (n_raw > 0) {
    const float inv = 1.0f / 0.2f;
    if (threadIdx.x < 3) { sh_imn[threadIdx.x] = f2ord(FLT_MAX); sh_imx[threadIdx.x] = f2ord(-FLT_MAX); }
    __syncthreads();
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int t = threadIdx.x; t < n_raw; t += FR_THREADS) {
      const float4 q = st.seg_cloud[base + k_lo + sm_lfpos[t]];
      mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
      mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
      mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
      if (lane == 0) {
        atomicMin(&sh_imn[d], f2ord(mn[d]));
        atomicMax(&sh_imx[d], f2ord(mx[d]));
      }
    }
    __syncthreads();
    float bmn[3], bmx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      bmn[d] = ord2f(sh_imn[d]);
      bmx[d] = ord2f(sh_imx[d]);
    }
    const long long dx = (long long)((bmx[0] - bmn[0]) * inv) + 1;
    const long long dy = (long long)((bmx[1] - bmn[1]) * inv) + 1;
    const long long dz = (long long)((bmx[2] - bmn[2]) * inv) + 1;
    if (dx * dy * dz > 2147483647LL) {
      // PCL refuses to filter (index overflow) and returns the input unchanged
      for (int t = threadIdx.x; t < n_raw; t += FR_THREADS) o_lflat[t] = st.seg_cloud[base + k_lo + sm_lfpos[t]];
      n_ds = n_raw;
    } else {
      const int minb0 = (int)floorf(bmn[0] * inv), minb1 = (int)floorf(bmn[1] * inv), minb2 = (int)floorf(bmn[2] * inv);
      const int div0 = (int)floorf(bmx[0] * inv) - minb0 + 1, div1 = (int)floorf(bmx[1] * inv) - minb1 + 1;
      int pos_bits = 0;
      while ((1 << pos_bits) < n_raw) ++pos_bits;
      int* sm_runend = sm_vox;  // [H] end position (exclusive) of the run starting at position t; written only after the
                                // last read of sm_vox (barrier after the run keys), so the two share their space
      for (int t = threadIdx.x; t < n_raw; t += FR_THREADS) {
        const float4 q = st.seg_cloud[base + k_lo + sm_lfpos[t]];
        const int i0 = (int)floorf(q.x * inv) - minb0;
        const int i1 = (int)floorf(q.y * inv) - minb1;
        const int i2 = (int)floorf(q.z * inv) - minb2;
        sm_vox[t] = i0 + i1 * div0 + i2 * div0 * div1;
      }
      __syncthreads();
      RING_CLOCK(5);
      // run heads -> compact run keys
      int n_runs = 0;
      {
        int run = 0;
        for (int t0 = 0; t0 < n_raw; t0 += FR_THREADS) {
          const int t = t0 + threadIdx.x;
          int head = 0;
          if (t < n_raw) head = (t == 0) || (sm_vox[t] != sm_vox[t - 1]);
          int total;
          const int ex = block_exclusive_scan(head, warp_tot, &total);
          if (head) keys[run + ex] = ((unsigned long long)(unsigned)sm_vox[t] << pos_bits) | (unsigned)t;
          run += total;
        }
        n_runs = run;
      }
      __syncthreads();
      // run ends: the next run head in input order (keys are still in input order here)
      for (int r = threadIdx.x; r < n_runs; r += FR_THREADS) {
        const unsigned pos_mask0 = (1u << pos_bits) - 1u;
        const int startp = (int)((unsigned)keys[r] & pos_mask0);
        const int endp = (r + 1 < n_runs) ? (int)((unsigned)keys[r + 1] & pos_mask0) : n_raw;
        sm_runend[startp] = endp;
      }
      __syncthreads();
      RING_CLOCK(6);
      n_runs_dbg = n_runs;
      // sort the runs by (voxel, first position)
      const unsigned pos_mask = (1u << pos_bits) - 1u;
      // sorted runs: (voxel, first position) of run t; the radix path keeps them as two arrays
      const unsigned* rk = nullptr;
      const unsigned short* rv = nullptr;
      unsigned short* sm_slot = sm_col;  // output slot of every voxel, see below (sm_col is dead by now)
      if (n_runs > FR_RADIX_MIN && n_runs <= FR_RADIX_MAX && p.H <= FR_RADIX_MAX) {
        // Many runs (a ring of far, sparse points: nearly every point a voxel of its own): stable LSD radix sort by voxel
        // -- the runs are in order of first position already.  Key buffers 0 / 1: the two halves of the u64 key area; value
        // buffers 0 / 1: the dead col / picked / label area.  The u64 keys are split in place, 256 at a time in ascending
        // order: the 32-bit keys of a chunk land in bytes whose u64 keys have been read already.
        unsigned* k0 = reinterpret_cast<unsigned*>(keys);
        unsigned* k1 = k0 + p.H;
        unsigned short* v0 = sm_col;
        unsigned short* v1 = v0 + p.H;
        for (int r0 = 0; r0 < n_runs; r0 += FR_THREADS) {
          const int r = r0 + threadIdx.x;
          const unsigned long long mine = r < n_runs ? keys[r] : 0ull;
          __syncthreads();
          if (r < n_runs) { k0[r] = (unsigned)(mine >> pos_bits); v0[r] = (unsigned short)((unsigned)mine & pos_mask); }
        }
        int vbits = 1;
        while (vbits < 31 && ((dx * dy * dz - 1) >> vbits) > 0) ++vbits;
        const int res = ring_radix_sort(k0, v0, k1, v1, n_runs, vbits, sh_sort, warp_tot);  // starts and ends with a barrier
        rk = res ? k1 : k0;
        rv = res ? v1 : v0;
        sm_slot = res ? v0 : v1;
      } else if (n_runs <= 96) {
        // a few dozen keys: rank by counting, no synchronised passes
        unsigned long long* sorted = keys + cap2 / 2;
        for (int r = threadIdx.x; r < n_runs; r += FR_THREADS) {
          const unsigned long long key = keys[r];
          int rank = 0;
          int u = 0;
          for (; u + 4 <= n_runs; u += 4) {
            rank += keys[u] < key ? 1 : 0;
            rank += keys[u + 1] < key ? 1 : 0;
            rank += keys[u + 2] < key ? 1 : 0;
            rank += keys[u + 3] < key ? 1 : 0;
          }
          for (; u < n_runs; ++u) rank += keys[u] < key ? 1 : 0;
          sorted[rank] = key;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < n_runs; r += FR_THREADS) keys[r] = sorted[r];
        __syncthreads();
      } else {
        int r2 = 1;
        while (r2 < n_runs) r2 <<= 1;
        for (int t = n_runs + threadIdx.x; t < r2; t += FR_THREADS) keys[t] = ~0ull;
        __syncthreads();
        bitonic_sort_t<unsigned long long>(keys, r2);
      }
      RING_CLOCK(7);
      auto vox_of = [&](int t) -> unsigned { return rk ? rk[t] : (unsigned)(keys[t] >> pos_bits); };
      auto pos_of = [&](int t) -> int { return rk ? (int)rv[t] : (int)((unsigned)keys[t] & pos_mask); };
      // output slot of every voxel (= of the first run of every group of equal voxel)
      int run = 0;
      for (int t0 = 0; t0 < n_runs; t0 += FR_THREADS) {
        const int t = t0 + threadIdx.x;
        int head = 0;
        if (t < n_runs) head = (t == 0) || (vox_of(t) != vox_of(t - 1));
        int total;
        const int ex = block_exclusive_scan(head, warp_tot, &total);
        if (t < n_runs) sm_slot[t] = head ? (unsigned short)(run + ex) : (unsigned short)0xffff;
        run += total;
      }
      // centroids: no barrier in this loop, so the point loads of different voxels overlap across the warps
      for (int t = threadIdx.x; t < n_runs; t += FR_THREADS) {
        const unsigned slot = sm_slot[t];
        if (slot == 0xffffu) continue;
        const unsigned vox = vox_of(t);
        float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
        int cnt = 0;
        for (int u = t; u < n_runs && vox_of(u) == vox; ++u) {
          const int ps = pos_of(u), pe = sm_runend[ps];
          for (int v = ps; v < pe; ++v) {
            const float4 q = st.seg_cloud[base + k_lo + sm_lfpos[v]];
            cx += q.x; cy += q.y; cz += q.z; ci += q.w;
            ++cnt;
          }
        }
        const float fc = (float)cnt;
        o_lflat[slot] = make_float4(cx / fc, cy / fc, cz / fc, ci / fc);
      }
      n_ds = run;
    }
  }
  RING_CLOCK(8);
  if (threadIdx.x == 0) {
    ring_clk[0] = clock64() - rc_begin;
    ring_clk[9] = (long long)L | ((long long)n_raw << 16) | ((long long)n_runs_dbg << 32) | ((long long)n_rounds_dbg << 48);
    int* o_counts = st.ring_counts + rs * 8;
    o_counts[0] = sh_counts[0];
    o_counts[1] = sh_counts[1];
    o_counts[2] = sh_counts[2];
    o_counts[3] = n_ds;
    o_counts[4] = n_raw;
  }
}

__global__ void __launch_bounds__(256) k_feature_compact(DevState st) {
  __shared__ int sh_off[4];
  const DevParams& p = st.p;
  const int ring = blockIdx.x, s = blockIdx.y;
  const size_t base = (size_t)s * p.N;
  if (threadIdx.x < 4) sh_off[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < ring) {
    const int* rc = st.ring_counts + ((size_t)s * p.V + threadIdx.x) * 8;
    atomicAdd(&sh_off[0], rc[0]);
    atomicAdd(&sh_off[1], rc[1]);
    atomicAdd(&sh_off[2], rc[2]);
    atomicAdd(&sh_off[3], rc[3]);
  }
  __syncthreads();
  const int* rc = st.ring_counts + ((size_t)s * p.V + ring) * 8;
  const size_t rs = (size_t)s * p.V + ring;
  for (int t = threadIdx.x; t < rc[0]; t += blockDim.x) {
    const int ind = st.st_sharp_ind[rs * 12 + t];
    st.corner_sharp[(size_t)s * p.cap_sharp + sh_off[0] + t] = st.seg_cloud[base + ind];
    st.corner_sharp_ind[(size_t)s * p.cap_sharp + sh_off[0] + t] = ind;
  }
  for (int t = threadIdx.x; t < rc[1]; t += blockDim.x) {
    const int ind = st.st_less_sharp_ind[rs * 120 + t];
    st.corner_less_sharp[(size_t)s * p.cap_less_sharp + sh_off[1] + t] = st.seg_cloud[base + ind];
    st.corner_less_sharp_ind[(size_t)s * p.cap_less_sharp + sh_off[1] + t] = ind;
  }
  for (int t = threadIdx.x; t < rc[2]; t += blockDim.x) {
    const int ind = st.st_flat_ind[rs * 24 + t];
    st.surf_flat[(size_t)s * p.cap_flat + sh_off[2] + t] = st.seg_cloud[base + ind];
    st.surf_flat_ind[(size_t)s * p.cap_flat + sh_off[2] + t] = ind;
  }
  for (int t = threadIdx.x; t < rc[3]; t += blockDim.x)
    st.surf_less_flat[(size_t)s * p.N + sh_off[3] + t] = st.st_less_flat[rs * p.H + t];
  if (ring == p.V - 1 && threadIdx.x < 4) st.feat_counts[s * 4 + threadIdx.x] = sh_off[threadIdx.x] + rc[threadIdx.x];
}

int next_pow2(int v) {
  int n = 1;
  while (n < v) n <<= 1;
  return n;
}

}  // namespace

void launch_feature_extraction(LaunchCtx& ctx, DevState& st) {
  const DevParams& p = st.p;
  {
    dim3 grid((p.N + FP_THREADS - 1) / FP_THREADS, p.B);
    LL_LAUNCH(ctx, "k_feature_prep", k_feature_prep<<<grid, FP_THREADS, 0, ctx.stream>>>(st));
  }
  const dim3 grid_rings(p.V, p.B);
  {
    // 37 KB at H = 2048: six blocks per SM; grid (sequence, rank of the ring by length)
    const int cap2 = next_pow2(p.H);
    const size_t smem = (size_t)cap2 * 8 + (size_t)p.H * 4 + (size_t)(p.H + 32) * (2 * 2 + 2);
    // the opt-in shared-memory size is a per-device attribute of the kernel
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && smem > configured[dev]) {
      cudaFuncSetAttribute(k_feature_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_feature_ring, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      configured[dev] = smem;
    }
    LL_LAUNCH(ctx, "k_feature_ring", k_feature_ring<<<dim3(p.B, p.V), FR_THREADS, smem, ctx.stream>>>(st, cap2));
  }
  LL_LAUNCH(ctx, "k_feature_compact", k_feature_compact<<<grid_rings, 256, 0, ctx.stream>>>(st));
}
