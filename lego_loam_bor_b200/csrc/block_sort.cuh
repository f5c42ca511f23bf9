// block_sort.cuh -- block-local stable LSD radix sort of (32-bit key, 32-bit value) pairs that ping-pongs through
// two global scratch arrays.  Stability is what the VoxelGrid restatements need: PCL sums the points of a voxel in
// sorted order, and the oracle (oracle/lego_oracle.cpp: voxel_grid) fixes that order as "input order inside a voxel".
// Used by k_voxel_grid (voxelgrid.cu) and the key-frame / local-map kernels (keyframes.cu).
#pragma once

#include "ll_device.cuh"

#define BS_THREADS 1024
#define BS_WARPS (BS_THREADS / 32)

struct BlockSortSmem {
  int warp_tot[33];
  unsigned hist[256];
  unsigned base[256];
  unsigned short whist[BS_WARPS][256];
};

// Sorts key[0][0..n) / val[0][0..n) ascending by key, 8 bits per pass, only as many passes as `max_key`
// (exclusive upper bound of the keys) needs.  Returns which of the two buffers (0 or 1) holds the result.
// Must be called by all BS_THREADS threads of the block; ends with a __syncthreads().
__device__ __forceinline__ int block_radix_sort(unsigned* const key[2], unsigned* const val[2], int n, long long max_key,
                                                BlockSortSmem& sm) {
  int passes = 1;
  while (passes < 4 && (max_key >> (8 * passes)) > 0) ++passes;
  int cur = 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = 8 * pass;
    const unsigned* kin = key[cur];
    const unsigned* vin = val[cur];
    unsigned* kout = key[cur ^ 1];
    unsigned* vout = val[cur ^ 1];
    if (threadIdx.x < 256) sm.hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += BS_THREADS) atomicAdd(&sm.hist[(kin[i] >> shift) & 255u], 1u);
    __syncthreads();
    {
      int total;
      const int v = threadIdx.x < 256 ? (int)sm.hist[threadIdx.x] : 0;
      const int ex = block_exclusive_scan(v, sm.warp_tot, &total);
      if (threadIdx.x < 256) sm.base[threadIdx.x] = (unsigned)ex;
    }
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += BS_THREADS) {
      const int i = t0 + threadIdx.x;
      for (int d = lane; d < 256; d += 32) sm.whist[wid][d] = 0;
      __syncwarp();
      unsigned k = 0, v = 0, dg = 0, rank = 0;
      const bool active = i < n;
      if (active) { k = kin[i]; v = vin[i]; dg = (k >> shift) & 255u; }
      const unsigned amask = __ballot_sync(0xffffffffu, active);
      if (active) {
        const unsigned peers = __match_any_sync(amask, dg);
        rank = __popc(peers & ((1u << lane) - 1u));
        if (rank == 0) sm.whist[wid][dg] = (unsigned short)__popc(peers);
      }
      __syncthreads();
      // exclusive scan over warps for every digit, and advance the bin bases
      if (threadIdx.x < 256) {
        unsigned run = 0;
        for (int w = 0; w < BS_WARPS; ++w) {
          const unsigned c = sm.whist[w][threadIdx.x];
          sm.whist[w][threadIdx.x] = (unsigned short)run;
          run += c;
        }
        sm.hist[threadIdx.x] = run;  // tile total of this digit
      }
      __syncthreads();
      if (active) {
        const unsigned pos = sm.base[dg] + sm.whist[wid][dg] + rank;
        kout[pos] = k;
        vout[pos] = v;
      }
      __syncthreads();
      if (threadIdx.x < 256) sm.base[threadIdx.x] += sm.hist[threadIdx.x];
      __syncthreads();
    }
    cur ^= 1;
  }
  return cur;
}
