// voxelgrid.cu -- pcl::VoxelGrid<PointXYZI> on whole clouds, on the device, for
// MapOptimization::downsampleCurrentScan (reference: LeGO-LOAM/src/mapOptmization.cpp:999-1026; leaf
// sizes :71-73).  PCL itself is not vendored in the reference; the algorithm restated here (and in
// oracle/lego_oracle.cpp: voxel_grid) is SURVEY.md section 8 f1 / 11.3: float inverse leaf, voxel index from
// floor(p * inv) - min_b, output one centroid (x, y, z, intensity) per occupied voxel in ascending voxel
// index, centroid summed in float in input order and divided by the count.
//
// One block per (sequence, cloud).  The voxel indices are sorted with a block-local stable LSD radix
// sort (8-bit digits, only as many passes as the index range needs) that ping-pongs through global
// scratch; stability keeps the input order inside a voxel, so the sequential float sums match the CPU.
#include "block_sort.cuh"
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

#define VG_THREADS BS_THREADS

struct VoxJob {
  const float4* a;   // first input cloud [B][stride_a]
  const int* na;     // count of a: na[s * na_stride]
  int stride_a, na_stride;
  const float4* b;   // optional second cloud appended after a (virtual concat), may be null
  const int* nb;
  int stride_b, nb_stride;
  float leaf;
  float4* out;       // [B][stride_out]
  int* nout;         // nout[s * nout_stride]
  int stride_out, nout_stride;
};

struct VoxArgs {
  VoxJob job[3];
  int njobs;
  unsigned* key0; unsigned* key1; unsigned* val0; unsigned* val1;  // [B][3][cap]
  int cap;
};

__device__ __forceinline__ float4 vox_point(const VoxJob& j, int s, int i, int na) {
  return i < na ? j.a[(size_t)s * j.stride_a + i] : j.b[(size_t)s * j.stride_b + (i - na)];
}

__device__ __forceinline__ int f2ord_vg(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f_vg(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(VG_THREADS) k_voxel_grid(VoxArgs args) {
  __shared__ int sh_imn[3], sh_imx[3];
  __shared__ BlockSortSmem sort_sm;
  int* const warp_tot = sort_sm.warp_tot;
  const int s = blockIdx.x;
  const int jid = blockIdx.y;
  const VoxJob& job = args.job[jid];
  const int na = job.na[s * job.na_stride];
  const int nb = job.b ? job.nb[s * job.nb_stride] : 0;
  const int n = min(na + nb, args.cap);
  float4* out = job.out + (size_t)s * job.stride_out;
  if (n == 0) {
    if (threadIdx.x == 0) job.nout[s * job.nout_stride] = 0;
    return;
  }
  const size_t soff = ((size_t)s * 3 + jid) * args.cap;
  unsigned* const key[2] = {args.key0 + soff, args.key1 + soff};
  unsigned* const val[2] = {args.val0 + soff, args.val1 + soff};
  const float inv = 1.0f / job.leaf;
  // ---- bounding box ----
  if (threadIdx.x < 3) { sh_imn[threadIdx.x] = f2ord_vg(FLT_MAX); sh_imx[threadIdx.x] = f2ord_vg(-FLT_MAX); }
  __syncthreads();
  {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = threadIdx.x; i < n; i += VG_THREADS) {
      const float4 q = vox_point(job, s, i, na);
      if (!isfinite(q.x) || !isfinite(q.y) || !isfinite(q.z)) continue;
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
      if ((threadIdx.x & 31) == 0) {
        atomicMin(&sh_imn[d], f2ord_vg(mn[d]));
        atomicMax(&sh_imx[d], f2ord_vg(mx[d]));
      }
    }
  }
  __syncthreads();
  float bmn[3], bmx[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { bmn[d] = ord2f_vg(sh_imn[d]); bmx[d] = ord2f_vg(sh_imx[d]); }
  const long long dx = (long long)((bmx[0] - bmn[0]) * inv) + 1;
  const long long dy = (long long)((bmx[1] - bmn[1]) * inv) + 1;
  const long long dz = (long long)((bmx[2] - bmn[2]) * inv) + 1;
  if (dx * dy * dz > 2147483647LL) {
    // PCL: "leaf size is too small for the input dataset" -> output = input
    const int nc = min(n, job.stride_out);  // never past the output cloud's capacity
    for (int i = threadIdx.x; i < nc; i += VG_THREADS) out[i] = vox_point(job, s, i, na);
    if (threadIdx.x == 0) job.nout[s * job.nout_stride] = nc;
    return;
  }
  const int minb0 = (int)floorf(bmn[0] * inv), minb1 = (int)floorf(bmn[1] * inv), minb2 = (int)floorf(bmn[2] * inv);
  const int div0 = (int)floorf(bmx[0] * inv) - minb0 + 1, div1 = (int)floorf(bmx[1] * inv) - minb1 + 1;
  const int div2 = (int)floorf(bmx[2] * inv) - minb2 + 1;
  const long long max_idx = (long long)div0 * div1 * div2;  // exclusive upper bound of the voxel index
  for (int i = threadIdx.x; i < n; i += VG_THREADS) {
    const float4 q = vox_point(job, s, i, na);
    unsigned k = 0xffffffffu;  // non-finite points sort last and are dropped
    if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
      const int i0 = (int)floorf(q.x * inv) - minb0;
      const int i1 = (int)floorf(q.y * inv) - minb1;
      const int i2 = (int)floorf(q.z * inv) - minb2;
      k = (unsigned)(i0 + i1 * div0 + i2 * div0 * div1);
    }
    key[0][i] = k;
    val[0][i] = (unsigned)i;
  }
  __syncthreads();
  // ---- stable LSD radix sort (block_sort.cuh) ----
  const int cur = block_radix_sort(key, val, n, max_idx, sort_sm);
  const unsigned* ks = key[cur];
  const unsigned* vs = val[cur];
  // ---- one centroid per run of equal voxel index ----
  int run_out = 0;
  for (int t0 = 0; t0 < n; t0 += VG_THREADS) {
    const int t = t0 + threadIdx.x;
    int head = 0;
    if (t < n && ks[t] != 0xffffffffu) head = (t == 0) || (ks[t] != ks[t - 1]);
    int total;
    const int ex = block_exclusive_scan(head, warp_tot, &total);
    if (head) {
      const unsigned vox = ks[t];
      float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
      int cnt = 0;
      for (int u = t; u < n && ks[u] == vox; ++u) {
        const float4 q = vox_point(job, s, (int)vs[u], na);
        cx += q.x; cy += q.y; cz += q.z; ci += q.w;
        ++cnt;
      }
      const float fc = (float)cnt;
      if (run_out + ex < job.stride_out) out[run_out + ex] = make_float4(cx / fc, cy / fc, cz / fc, ci / fc);
    }
    run_out += total;
  }
  if (threadIdx.x == 0) job.nout[s * job.nout_stride] = min(run_out, job.stride_out);
}

}  // namespace

void launch_downsample_current_scan(LaunchCtx& ctx, DevState& st) {
  const DevParams& p = st.p;
  VoxArgs a;
  a.key0 = st.vox_key0; a.key1 = st.vox_key1; a.val0 = st.vox_val0; a.val1 = st.vox_val1; a.cap = st.vox_cap;
  // downSizeFilterCorner 0.2 on corner_last; downSizeFilterSurf 0.4 on surf_last; downSizeFilterOutlier 0.4 on outlier_last
  a.njobs = 3;
  a.job[0] = VoxJob{st.corner_last, st.last_counts + 0, p.cap_less_sharp, 2, nullptr, nullptr, 0, 0, 0.2f,
                    st.scan_corner_ds, st.scan_ds_counts + 0, p.cap_less_sharp, 2};
  a.job[1] = VoxJob{st.surf_last, st.last_counts + 1, p.N, 2, nullptr, nullptr, 0, 0, 0.4f,
                    st.vox_tmp_surf, st.vox_tmp_counts + 0, p.N, 2};
  a.job[2] = VoxJob{st.outlier_last, st.odom_flags + 3, st.cap_outlier, 4, nullptr, nullptr, 0, 0, 0.4f,
                    st.vox_tmp_out, st.vox_tmp_counts + 1, st.cap_outlier, 2};
  LL_LAUNCH(ctx, "k_voxel_grid", k_voxel_grid<<<dim3(p.B, 3), VG_THREADS, 0, ctx.stream>>>(a));
  // laserCloudSurfTotalLast = surfDS + outlierDS, then downSizeFilterSurf again (mapOptmization.cpp:1019-1025)
  VoxArgs b = a;
  b.njobs = 1;
  b.job[0] = VoxJob{st.vox_tmp_surf, st.vox_tmp_counts + 0, p.N, 2, st.vox_tmp_out, st.vox_tmp_counts + 1, st.cap_outlier, 2,
                    0.4f, st.scan_surf_ds, st.scan_ds_counts + 1, p.N, 2};
  LL_LAUNCH(ctx, "k_voxel_grid_total", k_voxel_grid<<<dim3(p.B, 1), VG_THREADS, 0, ctx.stream>>>(b));
}
