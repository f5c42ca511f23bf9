// ingest.cu -- sensor_msgs/PointCloud2 decode + NaN removal on the device.
// Reference: ImageProjection::cloudHandler, LeGO-LOAM/src/imageProjection.cpp:159-161:
//     pcl::fromROSMsg(*laserCloudMsg, *_laser_cloud_in);
//     pcl::removeNaNFromPointCloud(*_laser_cloud_in, *_laser_cloud_in, indices);
// fromROSMsg copies the FLOAT32 fields x, y, z, intensity of every point_step-sized record (a field the message
// does not carry keeps the PointXYZI default, 0); removeNaNFromPointCloud keeps the order and drops every point
// whose x, y or z is not finite -- unless the cloud is flagged dense, in which case PCL copies it unchecked.
// (PCL is not vendored in the reference; both behaviours are restated from upstream: parity unpinned.)
//
// The ordered compaction is three launches over 2048-point tiles: per-tile counts, a scan of the tile counts per
// sequence, and the write of every kept point at (tile offset + rank inside the tile).
#include "ll_device.cuh"
#include "ll_kernels.h"

namespace {

#define PC2_THREADS 256
#define PC2_PER_THREAD 8
#define PC2_TILE (PC2_THREADS * PC2_PER_THREAD)

// layout classes of a message, decided on the host: 2 = the packed x y z intensity record of 16 bytes (one 16-byte load per
// point), 1 = every field 4-byte aligned (one 32-bit load per field), 0 = anything else (assembled bytewise)
template <int MODE>
__device__ __forceinline__ float pc2_field(const uint8_t* rec, int off) {
  if (MODE >= 1) return __ldg(reinterpret_cast<const float*>(rec + off));
  // records are point_step apart and fields sit at arbitrary byte offsets: assemble the little-endian float bytewise
  const unsigned b = (unsigned)rec[off] | ((unsigned)rec[off + 1] << 8) | ((unsigned)rec[off + 2] << 16) | ((unsigned)rec[off + 3] << 24);
  return __uint_as_float(b);
}

template <int MODE>
__device__ __forceinline__ float4 pc2_point(const Pc2Args& a, const uint8_t* rec) {
  if (MODE == 2) return __ldg(reinterpret_cast<const float4*>(rec));
  return make_float4(pc2_field<MODE>(rec, a.off_x), pc2_field<MODE>(rec, a.off_y), pc2_field<MODE>(rec, a.off_z),
                     a.off_intensity >= 0 ? pc2_field<MODE>(rec, a.off_intensity) : 0.f);
}

template <int MODE>
__device__ __forceinline__ bool pc2_keep(const Pc2Args& a, const uint8_t* rec) {
  if (a.is_dense) return true;
  float x, y, z;
  if (MODE == 2) { const float4 q = __ldg(reinterpret_cast<const float4*>(rec)); x = q.x; y = q.y; z = q.z; }
  else { x = pc2_field<MODE>(rec, a.off_x); y = pc2_field<MODE>(rec, a.off_y); z = pc2_field<MODE>(rec, a.off_z); }
  return isfinite(x) && isfinite(y) && isfinite(z);
}

template <int MODE>
__global__ void __launch_bounds__(PC2_THREADS) k_pc2_count(Pc2Args a) {
  __shared__ int warp_tot[33];
  const int s = blockIdx.y, tile = blockIdx.x;
  const int n = a.n_raw[s];
  const int t0 = tile * PC2_TILE;
  if (t0 >= n) {
    if (threadIdx.x == 0 && tile < a.ntiles) a.tile_cnt[s * a.ntiles + tile] = 0;
    return;
  }
  const uint8_t* raw = a.raw + (size_t)s * a.raw_stride;
  int c = 0;
  for (int k = 0; k < PC2_PER_THREAD; ++k) {
    const int i = t0 + threadIdx.x * PC2_PER_THREAD + k;
    if (i < n && pc2_keep<MODE>(a, raw + (size_t)i * a.point_step)) ++c;
  }
  int total;
  block_exclusive_scan(c, warp_tot, &total);
  if (threadIdx.x == 0) a.tile_cnt[s * a.ntiles + tile] = total;
}

__global__ void __launch_bounds__(1024) k_pc2_scan(Pc2Args a) {
  __shared__ int warp_tot[33];
  const int s = blockIdx.x;
  int run = 0;
  for (int t0 = 0; t0 < a.ntiles; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    const int v = t < a.ntiles ? a.tile_cnt[s * a.ntiles + t] : 0;
    int total;
    const int ex = block_exclusive_scan(v, warp_tot, &total);
    if (t < a.ntiles) a.tile_cnt[s * a.ntiles + t] = run + ex;
    run += total;
  }
  if (threadIdx.x == 0) a.n_out[s] = min(run, a.out_stride);
}

template <int MODE>
__global__ void __launch_bounds__(PC2_THREADS) k_pc2_write(Pc2Args a) {
  __shared__ int warp_tot[33];
  const int s = blockIdx.y, tile = blockIdx.x;
  const int n = a.n_raw[s];
  const int t0 = tile * PC2_TILE;
  if (t0 >= n) return;
  const uint8_t* raw = a.raw + (size_t)s * a.raw_stride;
  unsigned keep = 0u;
  int c = 0;
  for (int k = 0; k < PC2_PER_THREAD; ++k) {
    const int i = t0 + threadIdx.x * PC2_PER_THREAD + k;
    if (i < n && pc2_keep<MODE>(a, raw + (size_t)i * a.point_step)) { keep |= 1u << k; ++c; }
  }
  int total;
  int pos = a.tile_cnt[s * a.ntiles + tile] + block_exclusive_scan(c, warp_tot, &total);
  float4* out = a.out + (size_t)s * a.out_stride;
  for (int k = 0; k < PC2_PER_THREAD; ++k) {
    if (!((keep >> k) & 1u)) continue;
    const uint8_t* rec = raw + (size_t)(t0 + threadIdx.x * PC2_PER_THREAD + k) * a.point_step;
    if (pos < a.out_stride) out[pos] = pc2_point<MODE>(a, rec);
    ++pos;
  }
}

}  // namespace

void launch_decode_pointcloud2(LaunchCtx& ctx, cudaStream_t stream, int B, const Pc2Args& a) {
  cudaStream_t keep = ctx.stream;
  ctx.stream = stream;  // the decode runs on the copy stream, behind the H2D copy of the raw message
  const bool al4 = a.point_step % 4 == 0 && a.raw_stride % 4 == 0 && a.off_x % 4 == 0 && a.off_y % 4 == 0 && a.off_z % 4 == 0 &&
                   (a.off_intensity < 0 || a.off_intensity % 4 == 0) && ((uintptr_t)a.raw % 16 == 0);
  const bool packed16 = al4 && a.point_step == 16 && a.raw_stride % 16 == 0 && a.off_x == 0 && a.off_y == 4 && a.off_z == 8 && a.off_intensity == 12;
  const int mode = packed16 ? 2 : (al4 ? 1 : 0);
  const dim3 grid(a.ntiles, B);
  if (mode == 2) LL_LAUNCH(ctx, "k_pc2_count", k_pc2_count<2><<<grid, PC2_THREADS, 0, stream>>>(a));
  else if (mode == 1) LL_LAUNCH(ctx, "k_pc2_count", k_pc2_count<1><<<grid, PC2_THREADS, 0, stream>>>(a));
  else LL_LAUNCH(ctx, "k_pc2_count", k_pc2_count<0><<<grid, PC2_THREADS, 0, stream>>>(a));
  LL_LAUNCH(ctx, "k_pc2_scan", k_pc2_scan<<<B, 1024, 0, stream>>>(a));
  if (mode == 2) LL_LAUNCH(ctx, "k_pc2_write", k_pc2_write<2><<<grid, PC2_THREADS, 0, stream>>>(a));
  else if (mode == 1) LL_LAUNCH(ctx, "k_pc2_write", k_pc2_write<1><<<grid, PC2_THREADS, 0, stream>>>(a));
  else LL_LAUNCH(ctx, "k_pc2_write", k_pc2_write<0><<<grid, PC2_THREADS, 0, stream>>>(a));
  ctx.stream = keep;
}

int pc2_tile_points() { return PC2_TILE; }
