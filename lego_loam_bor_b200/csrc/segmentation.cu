// segmentation.cu -- ImageProjection::cloudSegmentation / labelComponents
// (reference: LeGO-LOAM/src/imageProjection.cpp:352-496).
//
// The reference flood-fills with a BFS per seed.  Its edge predicate (imageProjection.cpp:457-465)
// is symmetric in the two ranges, so a BFS component is a connected component of an undirected
// graph on the label==0 cells (4-neighbourhood, columns wrap, rows do not) and union-find gives the
// identical partition.  Roots are forced to be the smallest row-major index of a component, which
// is exactly the BFS seed (imageProjection.cpp:354-356), so
//   * "rows touched by non-seed pixels" (lineCountFlag, imageProjection.cpp:469) is the set of rows
//     of all component pixels except the root pixel, and
//   * the reference's label numbering (valid components only, in seed order, :489-490) is an
//     exclusive prefix count of feasible roots in row-major order.
//
// Kernels (all HBM/L2-bound integer work):
//   k_ccl_rows      one block per (sequence, row): horizontal runs by a block-wide max-scan
//   k_ccl_merge     one thread per cell: vertical edges + the column wrap edge, lock-free union (atomicMin)
//   k_ccl_flatten   one thread per cell: final root, component size / row mask (warp-aggregated atomics)
//   k_seg_count     one block per (sequence, row): feasible roots, kept cells, outlier cells
//   k_seg_emit      one block per (sequence, row): ordered compaction into segmented cloud + cloud_info
//   k_label_final   one thread per cell: numeric labels of non-root cells / 999999 (imageProjection.cpp:489-495); on demand
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

__device__ __forceinline__ bool seg_edge(float ra, float rb, float sin_a, float cos_a, float thr) {
  const float d1 = fmaxf(ra, rb);
  const float d2 = fminf(ra, rb);
  const float tang = (d2 * sin_a / (d1 - d2 * cos_a));
  return tang > thr;
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = __ldcg(parent + x);
  while (p != x) {
    const int gp = __ldcg(parent + p);
    if (gp != p) parent[x] = gp;  // path halving; any ancestor is a valid parent
    x = p;
    p = gp;
  }
  return x;
}

// both roots at once: the two chains of dependent loads are walked in lock step, so a union waits for the longer chain
// instead of the sum of the two (this kernel is bound by the latency of these loads)
__device__ __forceinline__ void uf_find2(int* parent, int& a, int& b) {
  int pa = __ldcg(parent + a), pb = __ldcg(parent + b);
  while (pa != a || pb != b) {
    const int ga = __ldcg(parent + pa), gb = __ldcg(parent + pb);
    if (pa != a && ga != pa) parent[a] = ga;  // path halving; any ancestor is a valid parent
    if (pb != b && gb != pb) parent[b] = gb;
    a = pa; pa = ga;
    b = pb; pb = gb;
  }
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    uf_find2(parent, a, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(parent + b, a);  // link the larger root under the smaller
    if (old == b) return;
    b = old;
  }
}

// block-wide exclusive max-scan, one value per thread; scratch: 33 ints
__device__ __forceinline__ int block_exclusive_max_scan(int v, int* warp_tot) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, t);
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int w = lane < nw ? warp_tot[lane] : -1;
    int winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc = max(winc, t);
    }
    int wex = __shfl_up_sync(0xffffffffu, winc, 1);
    if (lane == 0) wex = -1;
    if (lane < nw) warp_tot[lane] = wex;
  }
  __syncthreads();
  int ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) ex = -1;
  const int res = max(ex, warp_tot[wid]);
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(256) k_ccl_rows(DevState st) {
  extern __shared__ int sh_head[];  // [H] inclusive max of run starts within a thread's chunk
  __shared__ int warp_tot[33];
  const DevParams& p = st.p;
  const int row = blockIdx.x, s = blockIdx.y;
  const size_t base = (size_t)s * p.N + (size_t)row * p.H;
  const int ipt = (p.H + blockDim.x - 1) / blockDim.x;
  const int c0 = threadIdx.x * ipt, c1 = min(p.H, c0 + ipt);
  int running = -1;
  bool left_valid = false;
  float left_r = 0.f;
  if (c0 >= 1 && c0 < p.H) {
    left_valid = st.parent[base + c0 - 1] >= 0;
    left_r = st.range_mat[base + c0 - 1];
  }
  for (int c = c0; c < c1; ++c) {
    const bool valid = st.parent[base + c] >= 0;
    const float r = st.range_mat[base + c];
    if (valid) {
      const bool joined = left_valid && seg_edge(left_r, r, p.sin_ax, p.cos_ax, p.seg_tan_theta);
      if (!joined) running = c;
    }
    sh_head[c] = running;
    left_valid = valid;
    left_r = r;
  }
  const int carry = block_exclusive_max_scan(running, warp_tot);
  for (int c = c0; c < c1; ++c) {
    if (st.parent[base + c] >= 0) st.parent[base + c] = row * p.H + max(sh_head[c], carry);
  }
}

__global__ void __launch_bounds__(256) k_ccl_merge(DevState st) {
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= p.N) return;
  int* parent = st.parent + (size_t)s * p.N;
  const float* range = st.range_mat + (size_t)s * p.N;
  const int row = cell / p.H, col = cell - row * p.H;
  const bool has_up = row + 1 < p.V;
  const bool has_left = col > 0;
  // all first-touch loads are issued together (coalesced rows); stale values are harmless, see below
  const int up = has_up ? cell + p.H : cell;
  const int pc = parent[cell];
  const int pu = parent[up];
  const float r = range[cell];
  const float ru = range[up];
  const int pl = has_left ? parent[cell - 1] : -1;
  const int plu = has_left ? parent[up - 1] : -1;
  const float rl = has_left ? range[cell - 1] : 0.f;
  const float rlu = has_left ? range[up - 1] : 0.f;
  if (pc < 0) return;
  if (has_up && pu >= 0 && seg_edge(r, ru, p.sin_ay, p.cos_ay, p.seg_tan_theta)) {
    // The same two horizontal runs usually overlap over many columns; only the leftmost column of an
    // overlap has to union them.  Skip when the left neighbours are in the same two components (equal
    // parents, at whatever time they were read, imply equal components) and are vertically connected
    // themselves: that column, or one further left, does the union.
    const bool skip = pl >= 0 && plu >= 0 && pl == pc && plu == pu && seg_edge(rl, rlu, p.sin_ay, p.cos_ay, p.seg_tan_theta);
    // the finds start from the parents already loaded (the run heads written by k_ccl_rows, or better): any node of a
    // component stands for it, and this saves the two dependent loads of starting at the cells themselves
    if (!skip) uf_union(parent, pc, pu);
  }
  if (col == p.H - 1 && p.H > 1) {
    const int w = cell - col;  // column 0 of the same row (imageProjection.cpp:446-451)
    const int pw = __ldcg(parent + w);
    if (pw >= 0 && seg_edge(r, range[w], p.sin_ax, p.cos_ax, p.seg_tan_theta)) uf_union(parent, pc, pw);
  }
}

__global__ void __launch_bounds__(256) k_ccl_flatten(DevState st) {
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  int* parent = st.parent + (size_t)s * p.N;
  int root = -1;
  if (cell < p.N && parent[cell] >= 0) {
    int x = cell;
    int q = parent[x];
    while (q != x) { x = q; q = __ldcg(parent + x); }
    root = x;
    parent[cell] = root;
  }
  // warp-aggregated statistics per root
  const unsigned grp = __match_any_sync(0xffffffffu, root);
  unsigned bit = 0u;
  if (root >= 0 && cell != root) {
    const int d = cell / p.H - root / p.H;
    bit = 1u << min(d, 31);
  }
  const unsigned bits = __reduce_or_sync(grp, bit);
  if (root >= 0 && (int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) {
    atomicAdd(st.comp_size + (size_t)s * p.N + root, __popc(grp));
    if (bits) atomicOr(st.comp_rows + (size_t)s * p.N + root, bits);
  }
}

__device__ __forceinline__ bool feasible_root(const DevState& st, size_t base, int root) {
  const int size = st.comp_size[base + root];
  if (size >= 30) return true;
  if (size >= st.p.seg_valid_point_num) return __popc(st.comp_rows[base + root]) >= st.p.seg_valid_line_num;
  return false;
}

// classification of one cell for the ordered compaction of imageProjection.cpp:360-396
// bit0: feasible root, bit1: kept in the segmented cloud, bit2: outlier cloud
__device__ __forceinline__ int classify_cell(const DevState& st, size_t base, int row, int col) {
  const DevParams& p = st.p;
  const int cell = row * p.H + col;
  const int par = st.parent[base + cell];
  const bool ground = st.ground_mat[base + cell] == 1;
  int f = 0;
  const bool lab_pos = par >= 0;
  if (lab_pos || ground) {
    const bool feas = lab_pos && feasible_root(st, base, par);
    if (lab_pos && par == cell && feas) f |= 1;
    if (lab_pos && !feas) {  // label 999999
      if (row > p.gsi && col % 5 == 0) f |= 4;
      return f;
    }
    if (ground) {
      if (col % 5 != 0 && col > 5 && col < p.H - 5) return f;
    }
    f |= 2;
  }
  return f;
}

__global__ void __launch_bounds__(256) k_seg_count(DevState st) {
  __shared__ int sh_cnt[3];
  const DevParams& p = st.p;
  const int row = blockIdx.x, s = blockIdx.y;
  const size_t base = (size_t)s * p.N;
  if (threadIdx.x < 3) sh_cnt[threadIdx.x] = 0;
  __syncthreads();
  int c_root = 0, c_keep = 0, c_out = 0;
  for (int col = threadIdx.x; col < p.H; col += blockDim.x) {
    const int f = classify_cell(st, base, row, col);
    st.seg_class[base + (size_t)row * p.H + col] = (uint8_t)f;  // k_seg_emit orders the cells by these flags
    c_root += f & 1;
    c_keep += (f >> 1) & 1;
    c_out += (f >> 2) & 1;
  }
  c_root = warp_sum_i(c_root);
  c_keep = warp_sum_i(c_keep);
  c_out = warp_sum_i(c_out);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sh_cnt[0], c_root);
    atomicAdd(&sh_cnt[1], c_keep);
    atomicAdd(&sh_cnt[2], c_out);
  }
  __syncthreads();
  if (threadIdx.x < 3) st.tile_counts[((size_t)s * p.V + row) * 4 + threadIdx.x] = sh_cnt[threadIdx.x];
}

#ifndef SEG_TRIPS
#define SEG_TRIPS 4
#endif

// SEG_TRIPS block-wide exclusive scans at once (256 threads): one pair of barriers for all of them
__device__ __forceinline__ void block_exclusive_scan_trips(const int (&v)[SEG_TRIPS], int (&ex)[SEG_TRIPS], int (&total)[SEG_TRIPS],
                                                           int (*sh)[SEG_TRIPS]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc[SEG_TRIPS];
#pragma unroll
  for (int t = 0; t < SEG_TRIPS; ++t) {
    inc[t] = v[t];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, inc[t], o);
      if (lane >= o) inc[t] += x;
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int t = 0; t < SEG_TRIPS; ++t) sh[wid][t] = inc[t];
  }
  __syncthreads();
  // the nw x SEG_TRIPS warp totals are one value per lane of warp 0: exclusive prefix over the warps of every trip with
  // shuffles of stride SEG_TRIPS (a loop over the warps in every thread was a sixth of this kernel's instructions)
  static_assert(8 * SEG_TRIPS <= 32, "one warp total per lane");
  if (wid == 0) {
    const int w = lane / SEG_TRIPS, t = lane % SEG_TRIPS;
    const int x0 = w < nw ? sh[w][t] : 0;
    int incl = x0;
#pragma unroll
    for (int o = SEG_TRIPS; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (w < nw) sh[w][t] = incl - x0;
    if (w == nw - 1) sh[8][t] = incl;   // total of the trip
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < SEG_TRIPS; ++t) {
    ex[t] = sh[wid][t] + inc[t] - v[t];
    total[t] = sh[8][t];
  }
  __syncthreads();
}

// One block per (sequence, row).  The row is handled SEG_TRIPS x 256 columns at a time: the classification of all those
// cells (left by k_seg_count, one byte per cell), then the points of the kept cells and their orientation, are in flight
// together, and one multi-scan orders them; the kernel is a single wave of blocks, so its
// duration is the latency of one block.
__global__ void __launch_bounds__(256) k_seg_emit(DevState st) {
  __shared__ int sh_pre[3];
  __shared__ int sh_scan[9][SEG_TRIPS];   // warp totals -> exclusive prefixes; row 8: the totals
  const DevParams& p = st.p;
  const int row = blockIdx.x, s = blockIdx.y;
  const size_t base = (size_t)s * p.N;
  if (threadIdx.x < 3) sh_pre[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < row) {
    const int* tc = st.tile_counts + ((size_t)s * p.V + threadIdx.x) * 4;
    atomicAdd(&sh_pre[0], tc[0]);
    atomicAdd(&sh_pre[1], tc[1]);
    atomicAdd(&sh_pre[2], tc[2]);
  }
  __syncthreads();
  int run_root = sh_pre[0], run_keep = sh_pre[1], run_out = sh_pre[2];
  const int keep_before_row = run_keep;
  const float start_ori = st.orientation[s * 4 + 0];
  for (int c0 = 0; c0 < p.H; c0 += 256 * SEG_TRIPS) {
    int f[SEG_TRIPS], packed[SEG_TRIPS];
#pragma unroll
    for (int t = 0; t < SEG_TRIPS; ++t) {
      const int col = c0 + t * 256 + threadIdx.x;
      f[t] = col < p.H ? (int)st.seg_class[base + (size_t)row * p.H + col] : 0;
      // one packed scan for the three flags (each partial sum <= 256 < 2^10)
      packed[t] = (f[t] & 1) | (((f[t] >> 1) & 1) << 10) | (((f[t] >> 2) & 1) << 20);
    }
    float4 pt[SEG_TRIPS];
    float rng[SEG_TRIPS], ori_raw[SEG_TRIPS];
    bool gnd[SEG_TRIPS];
#pragma unroll
    for (int t = 0; t < SEG_TRIPS; ++t) {
      const int cell = row * p.H + c0 + t * 256 + threadIdx.x;
      pt[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      rng[t] = 0.f;
      gnd[t] = false;
      if (f[t] & 6) pt[t] = st.full_cloud[base + cell];
      if (f[t] & 2) {
        rng[t] = st.range_mat[base + cell];
        gnd[t] = st.ground_mat[base + cell] == 1;
      }
    }
#pragma unroll
    for (int t = 0; t < SEG_TRIPS; ++t) {
      ori_raw[t] = 0.f;
      if (f[t] & 2) ori_raw[t] = -ll_atan2f(pt[t].y, pt[t].x);  // point.x = seg.y, point.z = seg.x after the axis swap
    }
    int ex[SEG_TRIPS], total[SEG_TRIPS];
    block_exclusive_scan_trips(packed, ex, total, sh_scan);
#pragma unroll
    for (int t = 0; t < SEG_TRIPS; ++t) {
      const int col = c0 + t * 256 + threadIdx.x;
      const int cell = row * p.H + col;
      if (f[t] & 1) st.label_mat[base + cell] = run_root + (ex[t] & 1023) + 1;
      if (f[t] & 2) {
        const int pos = run_keep + ((ex[t] >> 10) & 1023);
        st.seg_cloud[base + pos] = pt[t];
        st.seg_range[base + pos] = rng[t];
        st.seg_col[base + pos] = (uint32_t)col;
        st.seg_ground[base + pos] = gnd[t] ? 1 : 0;
        // adjustDistortion's sequential halfPassed flag (featureAssociation.cpp:162,173-187): the flag
        // flips at the FIRST segmented point whose branch-1 orientation is more than pi past the start;
        // record candidates here, k_feature_prep reads the minimum.
        float ori = ori_raw[t];
        st.seg_ori[base + pos] = ori;        // adjustDistortion's raw orientation, reused by k_feature_prep
        if ((double)ori < (double)start_ori - LL_PI / 2)
          ori = (float)((double)ori + 2 * LL_PI);
        else if ((double)ori > (double)start_ori + LL_PI * 3 / 2)
          ori = (float)((double)ori - 2 * LL_PI);
        if ((double)(ori - start_ori) > LL_PI) atomicMin(st.half_idx + s, pos);
      }
      if (f[t] & 4) {
        const int pos = run_out + ((ex[t] >> 20) & 1023);
        if (pos < st.cap_outlier) st.outlier_cloud[(size_t)s * st.cap_outlier + pos] = pt[t];
      }
      run_root += total[t] & 1023;
      run_keep += (total[t] >> 10) & 1023;
      run_out += (total[t] >> 20) & 1023;
    }
  }
  if (threadIdx.x == 0) {
    st.start_ring[s * p.V + row] = keep_before_row - 1 + 5;  // imageProjection.cpp:361
    st.end_ring[s * p.V + row] = run_keep - 1 - 5;           // imageProjection.cpp:395
    if (row == p.V - 1) {
      st.seg_count[s] = run_keep;
      st.outlier_count[s] = min(run_out, st.cap_outlier);
    }
  }
}

__global__ void __launch_bounds__(256) k_label_final(DevState st) {
  const DevParams& p = st.p;
  const int s = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= p.N) return;
  const size_t base = (size_t)s * p.N;
  const int par = st.parent[base + cell];
  if (par < 0) return;  // label stays -1
  const bool feas = feasible_root(st, base, par);
  if (par == cell) {
    if (!feas) st.label_mat[base + cell] = LL_INVALID_LABEL;  // feasible roots were numbered by k_seg_emit
  } else {
    // labels of feasible roots are not written in this kernel, so this read does not race
    st.label_mat[base + cell] = feas ? st.label_mat[base + par] : LL_INVALID_LABEL;
  }
}

}  // namespace

void launch_segmentation(LaunchCtx& ctx, DevState& st) {
  const DevParams& p = st.p;
  const dim3 grid_rows(p.V, p.B);
  const dim3 grid_cells((p.N + 255) / 256, p.B);
  LL_LAUNCH(ctx, "k_ccl_rows", k_ccl_rows<<<grid_rows, 256, p.H * sizeof(int), ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_ccl_merge", k_ccl_merge<<<grid_cells, 256, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_ccl_flatten", k_ccl_flatten<<<grid_cells, 256, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_seg_count", k_seg_count<<<grid_rows, 256, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_seg_emit", k_seg_emit<<<grid_rows, 256, 0, ctx.stream>>>(st));
}

// Not part of the per-scan path: the reference never publishes labelMat, and everything cloudSegmentation reads from it is
// decided above from the forest (feasible or not); the numbers of the non-root cells are written when a caller asks for the
// matrix (ll_download(LL_BUF_LABEL_MAT)), before the next scan re-initialises the forest.
void launch_label_final(LaunchCtx& ctx, DevState& st) {
  const DevParams& p = st.p;
  const dim3 grid_cells((p.N + 255) / 256, p.B);
  LL_LAUNCH(ctx, "k_label_final", k_label_final<<<grid_cells, 256, 0, ctx.stream>>>(st));
}
