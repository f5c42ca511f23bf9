// synth_arena_cuda.cu -- the arena scan generator of synth_arena.h on the device, one thread per beam.
// Inputs only (bench / test data), not part of the hot path and not linked into liblego_loam_b200.so.
// Compiled with -fmad=false: every beam is computed by exactly the double operations of the host generator
// (synth_arena.c), so the two produce bit-identical scans.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "synth_arena.h"

namespace {

__global__ void __launch_bounds__(256) k_arena_scan(const ArenaScanCtx* ctxs, const ArenaBox* boxes, const int* nbs, int max_boxes,
                                                    float4* out) {
  __shared__ ArenaBox sh_boxes[ARENA_MAX_BOXES];
  __shared__ ArenaScanCtx k;
  const int b = blockIdx.y;
  const int nb = nbs[b];
  {
    const double* src = reinterpret_cast<const double*>(boxes + (size_t)b * max_boxes);
    double* dst = reinterpret_cast<double*>(sh_boxes);
    for (int i = threadIdx.x; i < nb * 6; i += blockDim.x) dst[i] = src[i];
    const uint64_t* cs = reinterpret_cast<const uint64_t*>(ctxs + b);
    uint64_t* cd = reinterpret_cast<uint64_t*>(&k);
    for (int i = threadIdx.x; i < (int)(sizeof(ArenaScanCtx) / 8); i += blockDim.x) cd[i] = cs[i];
  }
  __syncthreads();
  const int N = k.V * k.H;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // firing order: column H-1..0, ring 0..V-1
  if (j >= N) return;
  const int c = k.H - 1 - j / k.V, r = j % k.V;
  uint64_t h;
  double ds[3], dw[3];
  arena_beam(&k, r, c, &h, ds, dw);
  float p[3];
  const float nanv = __int_as_float(0x7fc00000);
  float4 o = make_float4(nanv, nanv, nanv, 0.f);
  if (arena_ray(&k, sh_boxes, nullptr, nb, h, ds, dw, p)) o = make_float4(p[0], p[1], p[2], 0.f);
  out[(size_t)b * N + j] = o;
}

}  // namespace

// ctxs [B], boxes [B][max_boxes], nbs [B]: DEVICE pointers; out: device float4 [B][V*H], NaN xyz where a beam has no
// return, in firing order.  Returns a cudaError_t.
extern "C" int arena_scans_device(const void* ctxs, const void* boxes, const int* nbs, int B, int max_boxes, int n_cells,
                                  void* out, void* stream) {
  if (B <= 0 || max_boxes > ARENA_MAX_BOXES) return (int)cudaErrorInvalidValue;
  dim3 grid((n_cells + 255) / 256, B);
  k_arena_scan<<<grid, 256, 0, (cudaStream_t)stream>>>((const ArenaScanCtx*)ctxs, (const ArenaBox*)boxes, nbs, max_boxes,
                                                       (float4*)out);
  return (int)cudaGetLastError();
}
