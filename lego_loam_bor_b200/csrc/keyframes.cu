// keyframes.cu -- MapOptimization's key frames and local map on the device, loop closure off.
// Reference (LeGO-LOAM/src/mapOptmization.cpp): saveKeyFramesAndFactor :1335-1474, extractSurroundingKeyFrames :856-996
// (the enable_loop_closure == false branch :915-987), transformPointCloud :428-473, clearCloud :1518-1523.
// iSAM2 (GTSAM, not vendored) is the identity here: without loop closures the graph is an odometry chain whose
// optimum is the inserted initial values (SURVEY.md sections 8c, 11.5).
//
// The reference concatenates the transformed clouds of all surrounding key frames (10.7 M points for the 500-key-frame
// map of BASELINE.json configs[3]) and runs pcl::VoxelGrid over them every mapping cycle.  A VoxelGrid centroid is a
// left-to-right float sum over the voxel's points in concatenation order, so it is kept as a running sum per voxel, and
// the cycle only touches what changed:
//
//   k_kf_decide   thread / sequence : the 0.3 m rule, key pose, pool allocation
//   k_kf_store    block / (sequence, cloud): transform by the (never changing) key pose, voxel keys, stable radix sort,
//                                     write to the pool: "the points of this key frame in that voxel, in order" is a
//                                     contiguous run
//   k_kf_select   block / sequence  : radius search over the key poses, 1 m VoxelGrid of the poses (its centroid
//                                     intensity, cast to int, is the key-frame id -- sic), update of
//                                     surroundingExistingKeyPosesID (erase + append, order preserved); any number of
//                                     key frames (block radix sorts through global scratch)
//   k_kf_accumulate block / (sequence, table partition), thread / run:
//                     append  a key frame that joined the list adds its runs to the sums of their voxels and links
//                             them at the tail of the voxels' run chains (= concatenation order);
//                     erase   a key frame that left the list is unlinked from the chains of the voxels it touches and
//                             those voxels are re-summed from their remaining runs, in chain order -- bit-identical
//                             to re-summing the whole concatenation, at the cost of the touched voxels only;
//                     rebuild everything from list position 0 (only under table pressure).
//                   A voxel belongs to one partition and a key frame has one run per voxel and cloud, so no two threads
//                   ever touch the same voxel.
//   k_kfx_sort_new / k_kfx_merge / k_kfx_alive / k_kfx_output
//                   the voxels in ascending PCL voxel index (= lexicographic (iz, iy, ix)) are kept as a sorted array
//                   across cycles: the few voxels a cycle creates are sorted by one block and merged in by binary
//                   search (thread / element, all tiles in parallel); the centroids of the voxels that still have
//                   points are written in that order -> laserCloudCornerFromMapDS / laserCloudSurfFromMapDS.
#include "block_sort.cuh"
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

#define KF_THREADS BS_THREADS
#define KF_MAX_CAP 32768   // key frames per sequence the select kernel's bitmaps cover
#define KF_NONE 0xFFFFFFFFu
#define XS_TILE 1024
#define KF_BIAS (1 << 20)
#define KF_EMPTY 0xFFFFFFFFFFFFFFFFull
#define KF_INVALID 0xFFFFFFFFFFFFFFFEull

enum { KF_ERR_SLOTS = 1, KF_ERR_POOL = 2, KF_ERR_TABLE = 4, KF_ERR_MAP = 8, KF_ERR_RANGE = 16 };

__device__ __forceinline__ unsigned long long kf_pack(int ix, int iy, int iz) {
  const int bx = ix + KF_BIAS, by = iy + KF_BIAS, bz = iz + KF_BIAS;
  if (((bx | by | bz) >> 21) != 0) return KF_INVALID;  // also catches negatives
  return ((unsigned long long)bz << 42) | ((unsigned long long)by << 21) | (unsigned long long)bx;
}
__device__ __forceinline__ void kf_unpack(unsigned long long k, int* ix, int* iy, int* iz) {
  *ix = (int)(k & 0x1FFFFFu) - KF_BIAS;
  *iy = (int)((k >> 21) & 0x1FFFFFu) - KF_BIAS;
  *iz = (int)((k >> 42) & 0x1FFFFFu) - KF_BIAS;
}
__device__ __forceinline__ unsigned long long kf_mix(unsigned long long x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

// ---------------------------------------------------------------------------------------------------------------
// saveKeyFramesAndFactor, the scalar part (mapOptmization.cpp:1335-1458)
__global__ void k_kf_decide(DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.p.B) return;
  KeyframeStore& kf = st.kf;
  float* aft = st.transform_aft_mapped + s * 6;
  float* tobe = st.transform_tobe_mapped + s * 6;
  float* rp = kf.robot_pos + s * 8;
  const float cx = aft[3], cy = aft[4], cz = aft[5];
  rp[0] = cx; rp[1] = cy; rp[2] = cz;  // currentRobotPosPoint
  const float px = rp[4], py = rp[5], pz = rp[6];
  const float d = sqrtf((px - cx) * (px - cx) + (py - cy) * (py - cy) + (pz - cz) * (pz - cz));
  const bool save = !((double)d < 0.3);
  const int n = kf.kf_count[s];
  kf.kf_new[s] = -1;
  if (!save && n > 0) return;
  const int nc = st.scan_ds_counts[s * 2 + 0], ns = st.vox_tmp_counts[s * 2 + 0], no = st.vox_tmp_counts[s * 2 + 1];
  if (n >= kf.kf_cap) { kf.err[s] |= KF_ERR_SLOTS; return; }
  const int used = kf.pool_used[s];
  if (used + nc + ns + no > kf.pool_cap) { kf.err[s] |= KF_ERR_POOL; return; }
  rp[4] = cx; rp[5] = cy; rp[6] = cz;  // previousRobotPosPoint = currentRobotPosPoint
  // first key frame: prior on transformTobeMapped (:1362-1376); later ones: transformAftMapped (:1391-1400)
  const float* est = n == 0 ? tobe : aft;
  float* pose = kf.kf_pose + ((size_t)s * kf.kf_cap + n) * 6;
  float e[6];
  for (int i = 0; i < 6; ++i) e[i] = est[i];
  for (int i = 0; i < 6; ++i) pose[i] = e[i];  // roll, pitch, yaw = T[0..2]; x, y, z = T[3..5] (:1419-1437)
  float* last = kf.transform_last + s * 6;
  if (n == 0) {
    for (int i = 0; i < 6; ++i) last[i] = tobe[i];
  } else {
    for (int i = 0; i < 6; ++i) { last[i] = aft[i]; tobe[i] = aft[i]; }  // :1440-1452
  }
  int* off = kf.kf_off + ((size_t)s * kf.kf_cap + n) * 4;
  off[0] = used; off[1] = used + nc; off[2] = used + nc + ns; off[3] = used + nc + ns + no;
  kf.pool_used[s] = used + nc + ns + no;
  kf.kf_count[s] = n + 1;
  kf.kf_new[s] = n;
}

struct KfPoseTrig { float cr, sr, cp, sp, cy, sy, tx, ty, tz; };

// transformPointCloud (mapOptmization.cpp:443-473)
__device__ __forceinline__ float4 kf_transform(const KfPoseTrig& t, const float4 p) {
  const float x1 = t.cy * p.x - t.sy * p.y;
  const float y1 = t.sy * p.x + t.cy * p.y;
  const float z1 = p.z;
  const float x2 = x1;
  const float y2 = t.cr * y1 - t.sr * z1;
  const float z2 = t.sr * y1 + t.cr * z1;
  float4 o;
  o.x = t.cp * x2 + t.sp * z2 + t.tx;
  o.y = y2 + t.ty;
  o.z = -t.sp * x2 + t.cp * z2 + t.tz;
  o.w = p.w;
  return o;
}

// copies of laserCloudCornerLastDS / SurfLastDS / OutlierLastDS (:1461-1474), stored transformed and voxel-sorted
__global__ void __launch_bounds__(KF_THREADS) k_kf_store(DevState st) {
  __shared__ BlockSortSmem sort_sm;
  __shared__ int sh_mn[3], sh_mx[3];
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.x, j = blockIdx.y;
  const int id = kf.kf_new[s];
  if (id < 0) return;
  const DevParams& p = st.p;
  const float4* src;
  int n;
  if (j == 0) { src = st.scan_corner_ds + (size_t)s * p.cap_less_sharp; n = st.scan_ds_counts[s * 2 + 0]; }
  else if (j == 1) { src = st.vox_tmp_surf + (size_t)s * p.N; n = st.vox_tmp_counts[s * 2 + 0]; }
  else { src = st.vox_tmp_out + (size_t)s * st.cap_outlier; n = st.vox_tmp_counts[s * 2 + 1]; }
  if (n == 0) return;
  const int off = kf.kf_off[((size_t)s * kf.kf_cap + id) * 4 + j];
  const float inv = 1.0f / (j == 0 ? kf.tbl[0].leaf : kf.tbl[1].leaf);
  const float* pose = kf.kf_pose + ((size_t)s * kf.kf_cap + id) * 6;
  KfPoseTrig t;
  ll_sincosf(pose[0], &t.sr, &t.cr);
  ll_sincosf(pose[1], &t.sp, &t.cp);
  ll_sincosf(pose[2], &t.sy, &t.cy);
  t.tx = pose[3]; t.ty = pose[4]; t.tz = pose[5];
  const size_t soff = ((size_t)s * 3 + j) * st.vox_cap;
  unsigned* const key[2] = {st.vox_key0 + soff, st.vox_key1 + soff};
  unsigned* const val[2] = {st.vox_val0 + soff, st.vox_val1 + soff};
  if (threadIdx.x < 3) { sh_mn[threadIdx.x] = INT_MAX; sh_mx[threadIdx.x] = INT_MIN; }
  __syncthreads();
  // voxel bounding box of the valid points
  {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = threadIdx.x; i < n; i += KF_THREADS) {
      const float4 q = kf_transform(t, src[i]);
      if (!isfinite(q.x) || !isfinite(q.y) || !isfinite(q.z)) continue;
      const int v[3] = {(int)floorf(q.x * inv), (int)floorf(q.y * inv), (int)floorf(q.z * inv)};
      if (kf_pack(v[0], v[1], v[2]) == KF_INVALID) continue;
#pragma unroll
      for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], v[d]); mx[d] = max(mx[d], v[d]); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
      if ((threadIdx.x & 31) == 0) { atomicMin(&sh_mn[d], mn[d]); atomicMax(&sh_mx[d], mx[d]); }
    }
  }
  __syncthreads();
  const int mn0 = sh_mn[0], mn1 = sh_mn[1], mn2 = sh_mn[2];
  long long max_idx = 1;
  int div0 = 1, div1 = 1;
  bool range_ok = true;
  if (sh_mx[0] >= mn0) {
    div0 = sh_mx[0] - mn0 + 1; div1 = sh_mx[1] - mn1 + 1;
    max_idx = (long long)div0 * div1 * (sh_mx[2] - mn2 + 1);
    range_ok = max_idx < 0xffffffffLL;
  }
  bool dropped = false;
  for (int i = threadIdx.x; i < n; i += KF_THREADS) {
    const float4 q = kf_transform(t, src[i]);
    unsigned k = 0xffffffffu;  // dropped points sort last
    if (range_ok && isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
      const int v0 = (int)floorf(q.x * inv), v1 = (int)floorf(q.y * inv), v2 = (int)floorf(q.z * inv);
      if (kf_pack(v0, v1, v2) != KF_INVALID) k = (unsigned)((v0 - mn0) + (v1 - mn1) * div0 + (v2 - mn2) * div0 * div1);
    }
    if (k == 0xffffffffu) dropped = true;
    key[0][i] = k;
    val[0][i] = (unsigned)i;
  }
  if (dropped) atomicOr(&kf.err[s], KF_ERR_RANGE);
  const int cur = block_radix_sort(key, val, n, max_idx, sort_sm);
  const unsigned* vs = val[cur];
  const unsigned* ks = key[cur];
  float4* dpts = kf.pool_pts + (size_t)s * kf.pool_cap + off;
  unsigned long long* dkey = kf.pool_key + (size_t)s * kf.pool_cap + off;
  int* dperm = kf.pool_perm + (size_t)s * kf.pool_cap + off;
  for (int u = threadIdx.x; u < n; u += KF_THREADS) {
    const int i = (int)vs[u];
    const float4 q = kf_transform(t, src[i]);
    unsigned long long k = KF_INVALID;
    if (ks[u] != 0xffffffffu) k = kf_pack((int)floorf(q.x * inv), (int)floorf(q.y * inv), (int)floorf(q.z * inv));
    dpts[u] = q;
    dkey[u] = k;
    dperm[u] = i;
  }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool kf_bit(const unsigned* bm, int i) { return (bm[i >> 5] >> (i & 31)) & 1u; }

// extractSurroundingKeyFrames, the key-pose part (mapOptmization.cpp:915-980)
__global__ void __launch_bounds__(KF_THREADS) k_kf_select(DevState st) {
  __shared__ BlockSortSmem sort_sm;
  __shared__ unsigned in_ds[KF_MAX_CAP / 32], in_sur[KF_MAX_CAP / 32];
  __shared__ int sh_mn[3], sh_mx[3];
  __shared__ int sh_cnt;
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.x, tid = threadIdx.x;
  const int n = kf.kf_count[s];
  if (n == 0) {  // :857 -- nothing to extract yet, the maps stay empty
    if (tid == 0) {
      kf.sur_valid[s] = 0; kf.sur_rebuild[s] = 0; kf.sur_first[s] = kf.sur_n[s];
      kf.sur_n_erased[s] = 0; kf.sur_last_erased[s] = 0;
    }
    return;
  }
  const size_t sbase = (size_t)s * kf.kf_cap;
  unsigned* const key[2] = {kf.sel_k0 + sbase, kf.sel_k1 + sbase};
  unsigned* const val[2] = {kf.sel_v0 + sbase, kf.sel_v1 + sbase};
  int* rank_idx = kf.sel_rank_idx + sbase;
  int* ds_ids = kf.sel_ds_ids + sbase;
  int* first_pos = kf.sel_first_pos + sbase;
  const float* poses = kf.kf_pose + sbase * 6;
  const float* rp = kf.robot_pos + s * 8;
  const float qx = rp[0], qy = rp[1], qz = rp[2];
  if (tid == 0) sh_cnt = 0;
  if (tid < 3) { sh_mn[tid] = INT_MAX; sh_mx[tid] = INT_MIN; }
  __syncthreads();
  // radiusSearch (nanoflann_pcl.h:155-175): d2 < r2, ascending by distance (ties: ascending index)
  {
    int local = 0;
    for (int i = tid; i < n; i += KF_THREADS) {
      const float* pose = poses + (size_t)i * 6;
      float d2 = 0.f, df;
      df = qx - pose[3]; d2 += df * df;
      df = qy - pose[4]; d2 += df * df;
      df = qz - pose[5]; d2 += df * df;
      const bool in = d2 < kf.radius2;
      key[0][i] = in ? __float_as_uint(d2) : 0xffffffffu;
      val[0][i] = (unsigned)i;
      local += in ? 1 : 0;
    }
    local = warp_sum_i(local);
    if ((tid & 31) == 0 && local) atomicAdd(&sh_cnt, local);
  }
  __syncthreads();
  const int n_sel = sh_cnt;
  {
    const int cur = block_radix_sort(key, val, n, 0x100000000LL, sort_sm);
    for (int t = tid; t < n_sel; t += KF_THREADS) rank_idx[t] = (int)val[cur][t];
  }
  __syncthreads();
  // downSizeFilterSurroundingKeyPoses (leaf 1.0, :78): voxel of every selected pose
  {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int t = tid; t < n_sel; t += KF_THREADS) {
      const float* pose = poses + (size_t)rank_idx[t] * 6;
      const int v[3] = {(int)floorf(pose[3]), (int)floorf(pose[4]), (int)floorf(pose[5])};
#pragma unroll
      for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], v[d]); mx[d] = max(mx[d], v[d]); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
      if ((tid & 31) == 0) { atomicMin(&sh_mn[d], mn[d]); atomicMax(&sh_mx[d], mx[d]); }
    }
  }
  __syncthreads();
  // all selected poses lie within 2 * radius of each other, so the index fits 32 bits (no PCL overflow branch)
  const int mn0 = sh_mn[0], mn1 = sh_mn[1], mn2 = sh_mn[2];
  const int div0 = n_sel ? sh_mx[0] - mn0 + 1 : 1, div1 = n_sel ? sh_mx[1] - mn1 + 1 : 1, div2 = n_sel ? sh_mx[2] - mn2 + 1 : 1;
  for (int t = tid; t < n_sel; t += KF_THREADS) {
    const float* pose = poses + (size_t)rank_idx[t] * 6;
    const int v0 = (int)floorf(pose[3]), v1 = (int)floorf(pose[4]), v2 = (int)floorf(pose[5]);
    key[0][t] = (unsigned)((v0 - mn0) + (v1 - mn1) * div0 + (v2 - mn2) * div0 * div1);
    val[0][t] = (unsigned)t;  // rank in distance order: the stable sort keeps it inside a voxel
  }
  const int cur2 = block_radix_sort(key, val, n_sel, (long long)div0 * div1 * div2, sort_sm);
  const unsigned* ks = key[cur2];
  const unsigned* vs = val[cur2];
  // one output pose per voxel; its id is (int)(mean intensity) = (int)(mean key-frame index) (:940, :962, :968)
  int n_ds = 0;
  for (int t0 = 0; t0 < n_sel; t0 += KF_THREADS) {
    const int t = t0 + tid;
    int head = 0, idv = 0;
    if (t < n_sel) {
      const unsigned vox = ks[t];
      head = t == 0 || ks[t - 1] != vox;
      if (head) {
        float sum = 0.f;
        int cnt = 0;
        for (int u = t; u < n_sel && ks[u] == vox; ++u) { sum += (float)rank_idx[vs[u]]; ++cnt; }
        idv = (int)(sum / (float)cnt);
      }
    }
    int total;
    const int ex = block_exclusive_scan(head, sort_sm.warp_tot, &total);
    if (head) ds_ids[n_ds + ex] = idv;
    n_ds += total;
  }
  for (int w = tid; w < (n + 31) / 32; w += KF_THREADS) { in_ds[w] = 0u; in_sur[w] = 0u; }
  for (int i = tid; i < n; i += KF_THREADS) first_pos[i] = INT_MAX;
  __syncthreads();
  for (int j = tid; j < n_ds; j += KF_THREADS) {
    const int id = ds_ids[j];
    atomicOr(&in_ds[id >> 5], 1u << (id & 31));
    atomicMin(&first_pos[id], j);
  }
  __syncthreads();
  // erase the key frames that left the surrounding set, keeping the order of the others (:935-955)
  const int m = kf.sur_n[s];
  int* sur = kf.sur_ids + sbase;
  int* erased = kf.sur_erased + sbase;
  int n_keep = 0, n_erased = 0;
  for (int t0 = 0; t0 < m; t0 += KF_THREADS) {
    const int t = t0 + tid;
    int keep = 0, gone = 0, sid = 0;
    if (t < m) { sid = sur[t]; keep = kf_bit(in_ds, sid) ? 1 : 0; gone = 1 - keep; }
    int tk, tg;
    const int exk = block_exclusive_scan(keep, sort_sm.warp_tot, &tk);
    const int exg = block_exclusive_scan(gone, sort_sm.warp_tot, &tg);
    if (keep) { sur[n_keep + exk] = sid; atomicOr(&in_sur[sid >> 5], 1u << (sid & 31)); }  // in place: n_keep + exk <= t
    if (gone) erased[n_erased + exg] = sid;
    n_keep += tk;
    n_erased += tg;
    __syncthreads();
  }
  // append the ids that are not in the list yet, in the order of the down-sampled poses (:957-980)
  int n_add = 0;
  for (int j0 = 0; j0 < n_ds; j0 += KF_THREADS) {
    const int j = j0 + tid;
    int add = 0, aid = 0;
    if (j < n_ds) { aid = ds_ids[j]; add = (!kf_bit(in_sur, aid) && first_pos[aid] == j) ? 1 : 0; }
    int total;
    const int ex = block_exclusive_scan(add, sort_sm.warp_tot, &total);
    if (add && n_keep + n_add + ex < kf.kf_cap) sur[n_keep + n_add + ex] = aid;
    n_add += total;
  }
  if (tid == 0) {
    const int m_new = min(n_keep + n_add, kf.kf_cap);
    // erased key frames leave dead voxels behind; when a table runs full, rebuild it from the list instead
    bool pressure = false;
    for (int mp = 0; mp < 2; ++mp)
      for (int part = 0; part < kf.tbl[mp].parts; ++part)
        pressure |= kf.tbl[mp].list_n[s * kf.tbl[mp].parts + part] > (kf.tbl[mp].sub_cap >> 1);
    const int rebuild = (n_erased > 0 && pressure) ? 1 : 0;
    kf.sur_n[s] = m_new;
    kf.sur_n_erased[s] = n_erased;
    kf.sur_last_erased[s] = n_erased > 0;
    kf.sur_rebuild[s] = rebuild;
    kf.sur_first[s] = rebuild ? 0 : n_keep;
    kf.sur_valid[s] = 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the concatenation of the surrounding clouds (:982-986) + the per-voxel sums of pcl::VoxelGrid, as running sums

__device__ __forceinline__ int kf_find(const unsigned long long* keys, unsigned mask, int sub_cap, unsigned long long k,
                                       unsigned long long hsh) {
  unsigned h = (unsigned)hsh & mask;
  for (int probe = 0; probe < sub_cap; ++probe) {
    const unsigned long long cur = keys[h];
    if (cur == k) return (int)h;
    if (cur == KF_EMPTY) return -1;
    h = (h + 1u) & mask;
  }
  return -1;
}

#define KF_DEAD 0x80000000u   // in the run-length half of pool_link: the run belongs to a key frame that left the list

// ---- erase, step 1: thread / stored point of an erased key frame: flag its runs ----
__global__ void __launch_bounds__(256) k_kf_erase_mark(DevState st) {
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.y;
  if (!kf.sur_valid[s] || kf.sur_rebuild[s]) return;
  const int n_erased = kf.sur_n_erased[s];
  const unsigned long long* pkey = kf.pool_key + (size_t)s * kf.pool_cap;
  unsigned long long* link = kf.pool_link + (size_t)s * kf.pool_cap;
  for (int e = blockIdx.z; e < n_erased; e += gridDim.z) {
    const int* off = kf.kf_off + ((size_t)s * kf.kf_cap + kf.sur_erased[(size_t)s * kf.kf_cap + e]) * 4;
    const int o0 = off[0], o1 = off[1], o2 = off[2], o3 = off[3];
    for (int i = o0 + blockIdx.x * blockDim.x + threadIdx.x; i < o3; i += gridDim.x * blockDim.x) {
      const int lo = i < o1 ? o0 : (i < o2 ? o1 : o2);
      const unsigned long long k = pkey[i];
      if (k == KF_INVALID || (i > lo && pkey[i - 1] == k)) continue;  // not the head of a run
      link[i] |= (unsigned long long)KF_DEAD;
    }
  }
}

// ---- erase, step 2: thread / run of an erased key frame.  The first one to claim the run's voxel walks its chain once:
// flagged runs are unlinked, the others re-summed in chain order (= the concatenation order of the remaining key
// frames), which is bit-identical to summing the whole concatenation again. ----
__global__ void __launch_bounds__(256) k_kf_erase_sweep(DevState st, int stamp) {
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.y;
  if (!kf.sur_valid[s] || kf.sur_rebuild[s]) return;
  const int n_erased = kf.sur_n_erased[s];
  const float4* pts = kf.pool_pts + (size_t)s * kf.pool_cap;
  const unsigned long long* pkey = kf.pool_key + (size_t)s * kf.pool_cap;
  unsigned long long* link = kf.pool_link + (size_t)s * kf.pool_cap;
  for (int e = blockIdx.z; e < n_erased; e += gridDim.z) {
    const int* off = kf.kf_off + ((size_t)s * kf.kf_cap + kf.sur_erased[(size_t)s * kf.kf_cap + e]) * 4;
    const int o0 = off[0], o1 = off[1], o2 = off[2], o3 = off[3];
    for (int i = o0 + blockIdx.x * blockDim.x + threadIdx.x; i < o3; i += gridDim.x * blockDim.x) {
      const int lo = i < o1 ? o0 : (i < o2 ? o1 : o2);
      const unsigned long long k = pkey[i];
      if (k == KF_INVALID || (i > lo && pkey[i - 1] == k)) continue;  // not the head of a run
      const VoxTable& tb = kf.tbl[i < o1 ? 0 : 1];
      const unsigned long long hsh = kf_mix(k);
      const int part = tb.parts > 1 ? (int)((hsh >> 40) % (unsigned)tb.parts) : 0;
      const size_t tbase = ((size_t)s * tb.parts + part) * tb.sub_cap;
      const int slot = kf_find(tb.key + tbase, (unsigned)tb.sub_cap - 1u, tb.sub_cap, k, hsh);
      if (slot < 0) { atomicOr(&kf.err[s], KF_ERR_TABLE); continue; }
      if (atomicExch(tb.claim + tbase + slot, stamp) == stamp) continue;  // another run of this voxel got there first
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      unsigned cur = (unsigned)tb.head[tbase + slot], prev = KF_NONE, first = KF_NONE;
      unsigned long long prev_lk = 0ull;  // link word of `prev` as it stands in memory
      while (cur != KF_NONE) {
        const unsigned long long lk = link[cur];
        const float4 q0 = pts[cur];  // every run has at least one point: fetched together with its link
        const unsigned nxt = (unsigned)(lk >> 32);
        const unsigned lw = (unsigned)(lk & 0xffffffffu);
        if (lw & KF_DEAD) {
          link[cur] = lk & ~(unsigned long long)KF_DEAD;
        } else {
          const int len = (int)lw;
          acc.x += q0.x; acc.y += q0.y; acc.z += q0.z; acc.w += q0.w;
          for (int u = 1; u < len; ++u) {
            const float4 q = pts[cur + u];
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
          }
          cnt += len;
          if (prev == KF_NONE) first = cur;
          else if ((unsigned)(prev_lk >> 32) != cur) link[prev] = ((unsigned long long)cur << 32) | (prev_lk & 0xffffffffull);
          prev = cur;
          prev_lk = lk;
        }
        cur = nxt;
      }
      if (prev != KF_NONE && (unsigned)(prev_lk >> 32) != KF_NONE) link[prev] = ((unsigned long long)KF_NONE << 32) | (prev_lk & 0xffffffffull);
      tb.head[tbase + slot] = (int)first;
      tb.tail[tbase + slot] = (int)prev;
      tb.sum[tbase + slot] = acc;
      tb.cnt[tbase + slot] = cnt;
    }
  }
}

__global__ void __launch_bounds__(KF_THREADS) k_kf_accumulate(DevState st) {
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.x;
  const int map = blockIdx.y == 0 ? 0 : 1;
  const int part = blockIdx.y == 0 ? 0 : blockIdx.y - 1;
  if (!kf.sur_valid[s]) return;
  const VoxTable& tb = kf.tbl[map];
  const size_t tbase = ((size_t)s * tb.parts + part) * tb.sub_cap;
  unsigned long long* keys = tb.key + tbase;
  float4* sums = tb.sum + tbase;
  int* cnts = tb.cnt + tbase;
  int* heads = tb.head + tbase;
  int* tails = tb.tail + tbase;
  unsigned* list = tb.list + tbase;
  int* list_n = tb.list_n + s * tb.parts + part;
  const unsigned mask = (unsigned)tb.sub_cap - 1u;
  const int max_fill = tb.sub_cap - (tb.sub_cap >> 2);
  const int* sur = kf.sur_ids + (size_t)s * kf.kf_cap;
  const float4* pts = kf.pool_pts + (size_t)s * kf.pool_cap;
  const unsigned long long* pkey = kf.pool_key + (size_t)s * kf.pool_cap;
  unsigned long long* link = kf.pool_link + (size_t)s * kf.pool_cap;
  const int c_first = map == 0 ? 0 : 1, c_last = map == 0 ? 1 : 3;  // corner | surf then outlier (:983-985)
  if (kf.sur_rebuild[s]) {
    const int ln = min(*list_n, tb.sub_cap);
    for (int i = threadIdx.x; i < ln; i += KF_THREADS) keys[list[i]] = KF_EMPTY;
    __syncthreads();
    if (threadIdx.x == 0) {
      *list_n = 0;
      tb.list_done[s * tb.parts + part] = 0;
      if (part == 0) kf.xs_n[s * 2 + map] = 0;
    }
    __syncthreads();
  }
  // ---- append: the key frames from list position sur_first on, in list order ----
  const int first = kf.sur_first[s], last = kf.sur_n[s];
  for (int r = first; r < last; ++r) {
    const int* off = kf.kf_off + ((size_t)s * kf.kf_cap + sur[r]) * 4;
    for (int c = c_first; c < c_last; ++c) {
      const int lo = off[c], hi = off[c + 1];
      for (int i = lo + threadIdx.x; i < hi; i += KF_THREADS) {
        const unsigned long long k = pkey[i];
        if (k == KF_INVALID || (i > lo && pkey[i - 1] == k)) continue;  // not the head of a run
        const unsigned long long hsh = kf_mix(k);
        if (tb.parts > 1 && (int)((hsh >> 40) % (unsigned)tb.parts) != part) continue;
        // find or insert the voxel
        unsigned h = (unsigned)hsh & mask;
        int slot = -1;
        bool fresh = false;
        for (int probe = 0; probe < tb.sub_cap; ++probe) {
          unsigned long long cur = *((volatile unsigned long long*)(keys + h));
          if (cur == KF_EMPTY) {
            if (*((volatile int*)list_n) >= max_fill) break;
            cur = atomicCAS(keys + h, KF_EMPTY, k);
            if (cur == KF_EMPTY) {
              const int pos = atomicAdd(list_n, 1);
              list[pos] = h;  // pos < sub_cap: at most one entry per slot
              slot = (int)h; fresh = true;
              break;
            }
          }
          if (cur == k) { slot = (int)h; break; }
          h = (h + 1u) & mask;
        }
        if (slot < 0) { atomicOr(&kf.err[s], KF_ERR_TABLE); continue; }
        float4 acc = fresh ? make_float4(0.f, 0.f, 0.f, 0.f) : sums[slot];
        int cnt = fresh ? 0 : cnts[slot];
        int len = 0;
        for (int u = i; u < hi && pkey[u] == k; ++u) {
          const float4 q = pts[u];
          acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
          ++len;
        }
        sums[slot] = acc;
        cnts[slot] = cnt + len;
        // the run goes to the tail of the voxel's chain
        link[i] = ((unsigned long long)KF_NONE << 32) | (unsigned)len;
        const int tl = (fresh || cnt == 0) ? -1 : tails[slot];
        if (tl < 0) heads[slot] = i;
        else link[tl] = ((unsigned long long)(unsigned)i << 32) | (link[tl] & 0xffffffffull);
        tails[slot] = i;
      }
      __syncthreads();  // the next cloud may continue the sums of this one
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// downSizeFilterCorner / downSizeFilterSurf output (:989-995): one centroid per voxel, ascending voxel index.

// lower bound: entries of a[0..n) smaller than k
__device__ __forceinline__ int kfx_lower_bound(const unsigned long long* a, int n, unsigned long long k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// block / (sequence, map): the voxels created since the last cycle, sorted; sets up this cycle's merge
__global__ void __launch_bounds__(KF_THREADS) k_kfx_sort_new(DevState st) {
  __shared__ BlockSortSmem sort_sm;
  __shared__ int sh_mn[3], sh_mx[3];
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.x, map = blockIdx.y;
  if (!kf.sur_valid[s]) return;
  const VoxTable& tb = kf.tbl[map];
  const size_t soff = ((size_t)s * 2 + map) * kf.sort_cap;
  unsigned* const key[2] = {kf.sk0 + soff, kf.sk1 + soff};
  unsigned* const val[2] = {kf.sv0 + soff, kf.sv1 + soff};
  const size_t tbase = (size_t)s * tb.parts * tb.sub_cap;
  int* merge = kf.xs_merge + (s * 2 + map) * 4;
  const int n_old = kf.xs_n[s * 2 + map];
  if (threadIdx.x < 3) { sh_mn[threadIdx.x] = INT_MAX; sh_mx[threadIdx.x] = INT_MIN; }
  __syncthreads();
  int n = 0;
  bool overflow = false;
  {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int part = 0; part < tb.parts; ++part) {
      const int done = tb.list_done[s * tb.parts + part];
      const int ln = min(tb.list_n[s * tb.parts + part], tb.sub_cap);
      const int take = max(0, min(ln - done, kf.sort_cap - n_old - n));
      const unsigned* list = tb.list + tbase + (size_t)part * tb.sub_cap + done;
      for (int i = threadIdx.x; i < take; i += KF_THREADS) {
        const unsigned g = (unsigned)(part * tb.sub_cap) + list[i];
        val[0][n + i] = g;
        int v[3];
        kf_unpack(tb.key[tbase + g], &v[0], &v[1], &v[2]);
#pragma unroll
        for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], v[d]); mx[d] = max(mx[d], v[d]); }
      }
      if (take < ln - done) overflow = true;
      n += take;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
      if ((threadIdx.x & 31) == 0) { atomicMin(&sh_mn[d], mn[d]); atomicMax(&sh_mx[d], mx[d]); }
    }
  }
  __syncthreads();
  if (overflow && threadIdx.x == 0) atomicOr(&kf.err[s], KF_ERR_MAP);
  if (n == 0) {
    if (threadIdx.x == 0) merge[2] = 0;
    return;
  }
  const int mn0 = sh_mn[0], mn1 = sh_mn[1], mn2 = sh_mn[2];
  const int div0 = sh_mx[0] - mn0 + 1, div1 = sh_mx[1] - mn1 + 1, div2 = sh_mx[2] - mn2 + 1;
  const long long max_idx = (long long)div0 * div1 * div2;   // of the new voxels only: their order is what is needed
  const bool range_ok = max_idx < 0xffffffffLL;
  if (!range_ok && threadIdx.x == 0) atomicOr(&kf.err[s], KF_ERR_RANGE);
  for (int i = threadIdx.x; i < n; i += KF_THREADS) {
    int v0, v1, v2;
    kf_unpack(tb.key[tbase + val[0][i]], &v0, &v1, &v2);
    key[0][i] = range_ok ? (unsigned)((v0 - mn0) + (v1 - mn1) * div0 + (v2 - mn2) * div0 * div1) : 0u;
  }
  const int cur = block_radix_sort(key, val, n, range_ok ? max_idx : 1, sort_sm);
  const unsigned* vs = val[cur];
  unsigned long long* nk = kf.xn_key + soff;
  unsigned* ns = kf.xn_slot + soff;
  for (int i = threadIdx.x; i < n; i += KF_THREADS) {
    const unsigned g = vs[i];
    nk[i] = tb.key[tbase + g];
    ns[i] = g;
  }
  if (threadIdx.x == 0) {
    for (int part = 0; part < tb.parts; ++part) tb.list_done[s * tb.parts + part] = min(tb.list_n[s * tb.parts + part], tb.sub_cap);
    const int src = kf.xs_cur[s * 2 + map];
    merge[0] = src; merge[1] = n_old; merge[2] = n;
    kf.xs_cur[s * 2 + map] = src ^ 1;
    kf.xs_n[s * 2 + map] = n_old + n;
    // running voxel bounding box of the map, for PCL's index overflow test (never shrinks: conservative after erases)
    int* bb = kf.xs_bbox + (s * 2 + map) * 8;
    if (n_old == 0) { bb[0] = mn0; bb[1] = mn1; bb[2] = mn2; bb[3] = sh_mx[0]; bb[4] = sh_mx[1]; bb[5] = sh_mx[2]; }
    else {
      bb[0] = min(bb[0], mn0); bb[1] = min(bb[1], mn1); bb[2] = min(bb[2], mn2);
      bb[3] = max(bb[3], sh_mx[0]); bb[4] = max(bb[4], sh_mx[1]); bb[5] = max(bb[5], sh_mx[2]);
    }
  }
}

// thread / element of the old sorted order or of the new voxels: its place in the merged order
__global__ void __launch_bounds__(XS_TILE) k_kfx_merge(DevState st) {
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.y, map = blockIdx.z;
  if (!kf.sur_valid[s]) return;
  const int* merge = kf.xs_merge + (s * 2 + map) * 4;
  const int n_new = merge[2];
  if (n_new == 0) return;
  const int src = merge[0], n_old = merge[1];
  const int j = blockIdx.x * XS_TILE + threadIdx.x;
  if (j >= n_old + n_new) return;
  const size_t soff = ((size_t)s * 2 + map) * kf.sort_cap;
  const unsigned long long* ok = kf.xs_key[src] + soff;
  const unsigned* os = kf.xs_slot[src] + soff;
  const unsigned long long* nk = kf.xn_key + soff;
  const unsigned* ns = kf.xn_slot + soff;
  unsigned long long* dk = kf.xs_key[src ^ 1] + soff;
  unsigned* dsl = kf.xs_slot[src ^ 1] + soff;
  if (j < n_old) {
    const unsigned long long k = ok[j];
    const int pos = j + kfx_lower_bound(nk, n_new, k);
    dk[pos] = k; dsl[pos] = os[j];
  } else {
    const int i = j - n_old;
    const unsigned long long k = nk[i];
    const int pos = i + kfx_lower_bound(ok, n_old, k);
    dk[pos] = k; dsl[pos] = ns[i];
  }
}

// voxels that still hold points, per tile of the sorted order
__global__ void __launch_bounds__(XS_TILE) k_kfx_alive(DevState st) {
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.y, map = blockIdx.z;
  if (!kf.sur_valid[s]) return;
  const int n = kf.xs_n[s * 2 + map];
  const int j = blockIdx.x * XS_TILE + threadIdx.x;
  if (blockIdx.x * XS_TILE >= n) return;
  const VoxTable& tb = kf.tbl[map];
  const size_t soff = ((size_t)s * 2 + map) * kf.sort_cap;
  const size_t tbase = (size_t)s * tb.parts * tb.sub_cap;
  const int cur = kf.xs_cur[s * 2 + map];
  int alive = 0;
  if (j < n) alive = tb.cnt[tbase + kf.xs_slot[cur][soff + j]] > 0 ? 1 : 0;
  const int c = __syncthreads_count(alive);
  if (threadIdx.x == 0) kf.xs_tile[(size_t)(s * 2 + map) * (kf.sort_cap / XS_TILE + 1) + blockIdx.x] = c;
}

__global__ void __launch_bounds__(XS_TILE) k_kfx_output(DevState st) {
  __shared__ int warp_tot[33];
  __shared__ int sh_pre;
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.y, map = blockIdx.z;
  int* out_n = st.map_counts + s * 2 + map;
  if (!kf.sur_valid[s]) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_n = 0;
    return;
  }
  const int n = kf.xs_n[s * 2 + map];
  if (n == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_n = 0;
    return;
  }
  if (blockIdx.x * XS_TILE >= n) return;
  const int* bb = kf.xs_bbox + (s * 2 + map) * 8;
  const long long max_idx = (long long)(bb[3] - bb[0] + 1) * (bb[4] - bb[1] + 1) * (bb[5] - bb[2] + 1);
  if (max_idx > 2147483647LL) {
    // PCL would warn "leaf size is too small" and return the undecimated cloud; that cloud is not materialised here
    if (blockIdx.x == 0 && threadIdx.x == 0) { *out_n = 0; atomicOr(&kf.err[s], KF_ERR_RANGE); }
    return;
  }
  const VoxTable& tb = kf.tbl[map];
  const size_t soff = ((size_t)s * 2 + map) * kf.sort_cap;
  const size_t tbase = (size_t)s * tb.parts * tb.sub_cap;
  const int cur = kf.xs_cur[s * 2 + map];
  const int* tiles = kf.xs_tile + (size_t)(s * 2 + map) * (kf.sort_cap / XS_TILE + 1);
  if (threadIdx.x < 32) {
    int t = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += 32) t += tiles[k];
    t = warp_sum_i(t);
    if (threadIdx.x == 0) sh_pre = t;
  }
  const int j = blockIdx.x * XS_TILE + threadIdx.x;
  unsigned g = 0;
  int cnt = 0;
  if (j < n) { g = kf.xs_slot[cur][soff + j]; cnt = tb.cnt[tbase + g]; }
  int total;
  const int ex = block_exclusive_scan(cnt > 0 ? 1 : 0, warp_tot, &total);  // syncs: sh_pre is visible after it
  float4* out = map == 0 ? st.map_corner + (size_t)s * st.cap_map_corner : st.map_surf + (size_t)s * st.cap_map_surf;
  const int cap_out = map == 0 ? st.cap_map_corner : st.cap_map_surf;
  const int pos = sh_pre + ex;
  if (cnt > 0 && pos < cap_out) {
    const float4 sum = tb.sum[tbase + g];
    const float fc = (float)cnt;
    out[pos] = make_float4(sum.x / fc, sum.y / fc, sum.z / fc, sum.w / fc);
  }
  if ((blockIdx.x + 1) * XS_TILE >= n && threadIdx.x == 0) {  // the last tile knows the total
    const int n_out = sh_pre + total;
    *out_n = min(n_out, cap_out);
    if (n_out > cap_out) atomicOr(&kf.err[s], KF_ERR_MAP);
  }
}

}  // namespace

void launch_extract_surrounding_keyframes(LaunchCtx& ctx, DevState& st) {
  const int B = st.p.B;
  const int tiles = (st.kf.sort_cap + XS_TILE - 1) / XS_TILE;
  LL_LAUNCH(ctx, "k_kf_select", k_kf_select<<<B, KF_THREADS, 0, ctx.stream>>>(st));
  // (both erase kernels return at once for the sequences that erased nothing, which is most cycles of most sequences)
  st.kf.stamp += 1;
  LL_LAUNCH(ctx, "k_kf_erase_mark", k_kf_erase_mark<<<dim3(16, B, 2), 256, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kf_erase_sweep", k_kf_erase_sweep<<<dim3(64, B, 2), 256, 0, ctx.stream>>>(st, st.kf.stamp));
  LL_LAUNCH(ctx, "k_kf_accumulate", k_kf_accumulate<<<dim3(B, 1 + st.kf.tbl[1].parts), KF_THREADS, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kfx_sort_new", k_kfx_sort_new<<<dim3(B, 2), KF_THREADS, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kfx_merge", k_kfx_merge<<<dim3(tiles, B, 2), XS_TILE, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kfx_alive", k_kfx_alive<<<dim3(tiles, B, 2), XS_TILE, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kfx_output", k_kfx_output<<<dim3(tiles, B, 2), XS_TILE, 0, ctx.stream>>>(st));
}

void launch_save_keyframe(LaunchCtx& ctx, DevState& st) {
  const int B = st.p.B;
  LL_LAUNCH(ctx, "k_kf_decide", k_kf_decide<<<(B + 63) / 64, 64, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kf_store", k_kf_store<<<dim3(B, 3), KF_THREADS, 0, ctx.stream>>>(st));
}
