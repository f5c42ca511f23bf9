// keyframes.cu -- MapOptimization's key frames and local map on the device, loop closure off.
// Reference (LeGO-LOAM/src/mapOptmization.cpp): saveKeyFramesAndFactor :1335-1474, extractSurroundingKeyFrames :856-996
// (the enable_loop_closure == false branch :915-987), transformPointCloud :428-473, clearCloud :1518-1523.
// iSAM2 (GTSAM, not vendored) is the identity here: without loop closures the graph is an odometry chain whose
// optimum is the inserted initial values (SURVEY.md sections 8c, 11.5).
//
// The reference concatenates the transformed clouds of all surrounding key frames (hundreds of thousands of points)
// and runs pcl::VoxelGrid over them every mapping cycle.  A VoxelGrid centroid is a left-to-right float sum over the
// voxel's points in concatenation order, so it can be kept as a running sum: appending a key frame adds its points to
// the sums of the voxels it touches, in order.  What is stored per key frame is therefore its three clouds already
// transformed by the (never changing) key pose and stable-sorted by voxel, so that "the points of this key frame in
// that voxel, in order" is a contiguous run:
//
//   k_kf_decide   thread / sequence : the 0.3 m rule, key pose, pool allocation
//   k_kf_store    block / (sequence, cloud): transform, voxel keys, stable radix sort, write to the pool
//   k_kf_select   block / sequence  : radius search over the key poses, 1 m VoxelGrid of the poses (its centroid
//                                     intensity, cast to int, is the key-frame id -- sic), update of
//                                     surroundingExistingKeyPosesID (erase + append, order preserved)
//   k_kf_accumulate block / (sequence, table partition): for every pending key frame in list order, thread per run:
//                                     find-or-insert the voxel, add the run's points to its sums.  Only appended key
//                                     frames are pending unless one was erased; then the tables are rebuilt.
//                                     A voxel belongs to one partition, so no two threads ever touch the same sums.
//   k_kf_extract  block / (sequence, map): occupied voxels sorted by PCL's voxel index -> centroids -> the local map
//                                     (laserCloudCornerFromMapDS / laserCloudSurfFromMapDS) ll_scan_to_map reads.
#include "block_sort.cuh"
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

#define KF_THREADS BS_THREADS
#define KF_MAX 1024
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
  rp[4] = cx; rp[5] = cy; rp[6] = cz;  // previousRobotPosPoint = currentRobotPosPoint
  const int nc = st.scan_ds_counts[s * 2 + 0], ns = st.vox_tmp_counts[s * 2 + 0], no = st.vox_tmp_counts[s * 2 + 1];
  if (n >= kf.kf_cap) { kf.err[s] |= KF_ERR_SLOTS; return; }
  const int used = kf.pool_used[s];
  if (used + nc + ns + no > kf.pool_cap) { kf.err[s] |= KF_ERR_POOL; return; }
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
// bitonic sort of KF_MAX 64-bit keys in shared memory, ascending; all KF_THREADS threads
__device__ __forceinline__ void kf_bitonic(unsigned long long* a) {
  const int i = threadIdx.x;
  for (int k = 2; k <= KF_MAX; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int x = i ^ j;
      if (x > i) {
        const unsigned long long u = a[i], v = a[x];
        const bool asc = (i & k) == 0;
        if ((u > v) == asc) { a[i] = v; a[x] = u; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ bool kf_bit(const unsigned* bm, int i) { return (bm[i >> 5] >> (i & 31)) & 1u; }

// extractSurroundingKeyFrames, the key-pose part (mapOptmization.cpp:915-980)
__global__ void __launch_bounds__(KF_THREADS) k_kf_select(DevState st) {
  __shared__ unsigned long long a[KF_MAX];
  __shared__ int idx_by_rank[KF_MAX];
  __shared__ int ds_ids[KF_MAX];
  __shared__ int first_pos[KF_MAX];
  __shared__ int sur_sm[KF_MAX];
  __shared__ unsigned in_ds[KF_MAX / 32], in_sur[KF_MAX / 32];
  __shared__ int warp_tot[33];
  __shared__ int sh_mn[3], sh_mx[3];
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.x, tid = threadIdx.x;
  const int n = kf.kf_count[s];
  if (n == 0) {  // :857 -- nothing to extract yet, the maps stay empty
    if (tid == 0) { kf.sur_valid[s] = 0; kf.sur_rebuild[s] = 0; kf.sur_first[s] = kf.sur_n[s]; }
    return;
  }
  const float* rp = kf.robot_pos + s * 8;
  const float qx = rp[0], qy = rp[1], qz = rp[2];
  // radiusSearch (nanoflann_pcl.h:155-175): d2 < r2, ascending by distance (ties: ascending index)
  float px = 0.f, py = 0.f, pz = 0.f;
  unsigned long long e = KF_EMPTY;
  if (tid < n) {
    const float* pose = kf.kf_pose + ((size_t)s * kf.kf_cap + tid) * 6;
    px = pose[3]; py = pose[4]; pz = pose[5];
    float d2 = 0.f, df;
    df = qx - px; d2 += df * df;
    df = qy - py; d2 += df * df;
    df = qz - pz; d2 += df * df;
    if (d2 < kf.radius2) e = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)tid;
  }
  a[tid] = e;
  if (tid < 3) { sh_mn[tid] = INT_MAX; sh_mx[tid] = INT_MIN; }
  if (tid < KF_MAX / 32) { in_ds[tid] = 0u; in_sur[tid] = 0u; }
  first_pos[tid] = INT_MAX;
  __syncthreads();
  const int n_sel = __syncthreads_count(e != KF_EMPTY);
  kf_bitonic(a);
  // downSizeFilterSurroundingKeyPoses (leaf 1.0, :78): voxel of every selected pose
  int v0 = 0, v1 = 0, v2 = 0, idx = 0;
  if (tid < n_sel) {
    idx = (int)(a[tid] & 0xffffffffu);
    idx_by_rank[tid] = idx;
    const float* pose = kf.kf_pose + ((size_t)s * kf.kf_cap + idx) * 6;
    const float inv = 1.0f / 1.0f;
    v0 = (int)floorf(pose[3] * inv); v1 = (int)floorf(pose[4] * inv); v2 = (int)floorf(pose[5] * inv);
    atomicMin(&sh_mn[0], v0); atomicMax(&sh_mx[0], v0);
    atomicMin(&sh_mn[1], v1); atomicMax(&sh_mx[1], v1);
    atomicMin(&sh_mn[2], v2); atomicMax(&sh_mx[2], v2);
  }
  __syncthreads();
  e = KF_EMPTY;
  if (tid < n_sel) {
    // all selected poses lie within 2 * 50 m of each other, so the index fits 32 bits (no PCL overflow branch)
    const int div0 = sh_mx[0] - sh_mn[0] + 1, div1 = sh_mx[1] - sh_mn[1] + 1;
    const unsigned vox = (unsigned)((v0 - sh_mn[0]) + (v1 - sh_mn[1]) * div0 + (v2 - sh_mn[2]) * div0 * div1);
    e = ((unsigned long long)vox << 32) | (unsigned)tid;  // (voxel, rank in distance order): stable
  }
  __syncthreads();
  a[tid] = e;
  __syncthreads();
  kf_bitonic(a);
  // one output pose per voxel; its id is (int)(mean intensity) = (int)(mean key-frame index) (:940, :962, :968)
  int head = 0, idv = 0;
  if (tid < n_sel) {
    const unsigned vox = (unsigned)(a[tid] >> 32);
    head = tid == 0 || (unsigned)(a[tid - 1] >> 32) != vox;
    if (head) {
      float sum = 0.f;
      int cnt = 0;
      for (int u = tid; u < n_sel && (unsigned)(a[u] >> 32) == vox; ++u) {
        sum += (float)idx_by_rank[(int)(a[u] & 0xffffffffu)];
        ++cnt;
      }
      idv = (int)(sum / (float)cnt);
    }
  }
  int n_ds;
  {
    const int ex = block_exclusive_scan(head, warp_tot, &n_ds);
    if (head) ds_ids[ex] = idv;
  }
  __syncthreads();
  if (tid < n_ds) {
    const int id = ds_ids[tid];
    atomicOr(&in_ds[id >> 5], 1u << (id & 31));
    atomicMin(&first_pos[id], tid);
  }
  __syncthreads();
  // erase the key frames that left the surrounding set, keeping the order of the others (:935-955)
  const int m = kf.sur_n[s];
  int* sur = kf.sur_ids + (size_t)s * kf.kf_cap;
  int keep = 0, sid = 0;
  if (tid < m) { sid = sur[tid]; keep = kf_bit(in_ds, sid) ? 1 : 0; }
  int n_keep;
  {
    const int ex = block_exclusive_scan(keep, warp_tot, &n_keep);
    if (keep) { sur_sm[ex] = sid; atomicOr(&in_sur[sid >> 5], 1u << (sid & 31)); }
  }
  __syncthreads();
  // append the ids that are not in the list yet, in the order of the down-sampled poses (:957-980)
  int add = 0, aid = 0;
  if (tid < n_ds) { aid = ds_ids[tid]; add = (!kf_bit(in_sur, aid) && first_pos[aid] == tid) ? 1 : 0; }
  int n_add;
  {
    const int ex = block_exclusive_scan(add, warp_tot, &n_add);
    if (add) sur_sm[n_keep + ex] = aid;
  }
  __syncthreads();
  const int m_new = n_keep + n_add;
  if (tid < m_new) sur[tid] = sur_sm[tid];
  if (tid == 0) {
    kf.sur_n[s] = m_new;
    kf.sur_rebuild[s] = n_keep != m;
    kf.sur_first[s] = n_keep != m ? 0 : n_keep;
    kf.sur_valid[s] = 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the concatenation of the surrounding clouds (:982-986) + the per-voxel sums of pcl::VoxelGrid, as running sums
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
  unsigned* list = tb.list + tbase;
  int* list_n = tb.list_n + s * tb.parts + part;
  const unsigned mask = (unsigned)tb.sub_cap - 1u;
  const int max_fill = tb.sub_cap - (tb.sub_cap >> 2);
  if (kf.sur_rebuild[s]) {
    const int ln = min(*list_n, tb.sub_cap);
    for (int i = threadIdx.x; i < ln; i += KF_THREADS) keys[list[i]] = KF_EMPTY;
    __syncthreads();
    if (threadIdx.x == 0) *list_n = 0;
    __syncthreads();
  }
  const int first = kf.sur_first[s], last = kf.sur_n[s];
  const int* sur = kf.sur_ids + (size_t)s * kf.kf_cap;
  const float4* pts = kf.pool_pts + (size_t)s * kf.pool_cap;
  const unsigned long long* pkey = kf.pool_key + (size_t)s * kf.pool_cap;
  const int c_first = map == 0 ? 0 : 1, c_last = map == 0 ? 1 : 3;  // corner | surf then outlier (:983-985)
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
        for (int u = i; u < hi && pkey[u] == k; ++u) {
          const float4 q = pts[u];
          acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
          ++cnt;
        }
        sums[slot] = acc;
        cnts[slot] = cnt;
      }
      __syncthreads();  // the next cloud may continue the sums of this one
    }
  }
}

// downSizeFilterCorner / downSizeFilterSurf output (:989-995): one centroid per voxel, ascending voxel index
__global__ void __launch_bounds__(KF_THREADS) k_kf_extract(DevState st) {
  __shared__ BlockSortSmem sort_sm;
  __shared__ int sh_mn[3], sh_mx[3];
  KeyframeStore& kf = st.kf;
  const int s = blockIdx.x, map = blockIdx.y;
  int* out_n = st.map_counts + s * 2 + map;
  if (!kf.sur_valid[s]) {
    if (threadIdx.x == 0) *out_n = 0;
    return;
  }
  const VoxTable& tb = kf.tbl[map];
  float4* out = map == 0 ? st.map_corner + (size_t)s * st.cap_map_corner : st.map_surf + (size_t)s * st.cap_map_surf;
  const int cap_out = min(map == 0 ? st.cap_map_corner : st.cap_map_surf, kf.sort_cap);
  const size_t soff = ((size_t)s * 2 + map) * kf.sort_cap;
  unsigned* const key[2] = {kf.sk0 + soff, kf.sk1 + soff};
  unsigned* const val[2] = {kf.sv0 + soff, kf.sv1 + soff};
  const size_t tbase = (size_t)s * tb.parts * tb.sub_cap;
  if (threadIdx.x < 3) { sh_mn[threadIdx.x] = INT_MAX; sh_mx[threadIdx.x] = INT_MIN; }
  __syncthreads();
  // gather the occupied slots of all partitions (val = slot over the whole table) and their voxel bounding box
  int n = 0;
  {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int part = 0; part < tb.parts; ++part) {
      const int ln = min(tb.list_n[s * tb.parts + part], tb.sub_cap);
      const int take = min(ln, kf.sort_cap - n);
      const unsigned* list = tb.list + tbase + (size_t)part * tb.sub_cap;
      for (int i = threadIdx.x; i < take; i += KF_THREADS) {
        const unsigned g = (unsigned)(part * tb.sub_cap) + list[i];
        val[0][n + i] = g;
        int v[3];
        kf_unpack(tb.key[tbase + g], &v[0], &v[1], &v[2]);
#pragma unroll
        for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], v[d]); mx[d] = max(mx[d], v[d]); }
      }
      if (take < ln && threadIdx.x == 0) atomicOr(&kf.err[s], KF_ERR_MAP);
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
  if (n == 0) {
    if (threadIdx.x == 0) *out_n = 0;
    return;
  }
  const int mn0 = sh_mn[0], mn1 = sh_mn[1], mn2 = sh_mn[2];
  const int div0 = sh_mx[0] - mn0 + 1, div1 = sh_mx[1] - mn1 + 1, div2 = sh_mx[2] - mn2 + 1;
  const long long max_idx = (long long)div0 * div1 * div2;
  if (max_idx > 2147483647LL) {
    // PCL would warn "leaf size is too small" and return the undecimated cloud; that cloud is not materialised here
    if (threadIdx.x == 0) { *out_n = 0; atomicOr(&kf.err[s], KF_ERR_RANGE); }
    return;
  }
  for (int i = threadIdx.x; i < n; i += KF_THREADS) {
    int v0, v1, v2;
    kf_unpack(tb.key[tbase + val[0][i]], &v0, &v1, &v2);
    key[0][i] = (unsigned)((v0 - mn0) + (v1 - mn1) * div0 + (v2 - mn2) * div0 * div1);
  }
  const int cur = block_radix_sort(key, val, n, max_idx, sort_sm);
  const unsigned* vs = val[cur];
  const int n_out = min(n, cap_out);
  for (int i = threadIdx.x; i < n_out; i += KF_THREADS) {
    const unsigned g = vs[i];
    const float4 sum = tb.sum[tbase + g];
    const float fc = (float)tb.cnt[tbase + g];
    out[i] = make_float4(sum.x / fc, sum.y / fc, sum.z / fc, sum.w / fc);
  }
  if (threadIdx.x == 0) {
    *out_n = n_out;
    if (n_out < n) atomicOr(&kf.err[s], KF_ERR_MAP);
  }
}

}  // namespace

void launch_extract_surrounding_keyframes(LaunchCtx& ctx, DevState& st) {
  const int B = st.p.B;
  LL_LAUNCH(ctx, "k_kf_select", k_kf_select<<<B, KF_THREADS, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kf_accumulate", k_kf_accumulate<<<dim3(B, 1 + st.kf.tbl[1].parts), KF_THREADS, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kf_extract", k_kf_extract<<<dim3(B, 2), KF_THREADS, 0, ctx.stream>>>(st));
}

void launch_save_keyframe(LaunchCtx& ctx, DevState& st) {
  const int B = st.p.B;
  LL_LAUNCH(ctx, "k_kf_decide", k_kf_decide<<<(B + 63) / 64, 64, 0, ctx.stream>>>(st));
  LL_LAUNCH(ctx, "k_kf_store", k_kf_store<<<dim3(B, 3), KF_THREADS, 0, ctx.stream>>>(st));
}
