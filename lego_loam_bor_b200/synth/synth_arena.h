/*
 * synth_arena.h -- the "arena" synthetic world of BASELINE.json configs[3]/[4] (way 1):
 * a 120 m x 120 m walled arena with pillars and low walls, 500 key-frame poses >= 1 m apart on an Archimedean
 * spiral inside a 50 m radius, and a driving circle through the middle of it for the timed sequence.
 *
 * Inputs only -- not part of the hot path.  Single source for gcc (libsynth_lidar.so, host) and nvcc
 * (libsynth_cuda.so, one thread per ray): everything a ray needs is built from IEEE-754 double + - * / sqrt in one
 * fixed order (ll_portable_math.h for sin / cos, an Irwin-Hall sum of twelve 16-bit uniforms instead of Box-Muller
 * for the range noise), so the host and the device produce bit-identical scans (tests/test_synth_arena.py).
 *
 * Beam geometry, jitter, firing order and the no-return rule are those of synth_lidar.c.
 */
#ifndef SYNTH_ARENA_H
#define SYNTH_ARENA_H

#include <stdint.h>

#include "../../include/ll_portable_math.h"

#define ARENA_MAX_BOXES 320
#define ARENA_GROUND_Z (-1.5)

typedef struct ArenaConfig {
  int32_t V, H;
  float bottom_deg, top_deg;
  uint64_t seed;
  float range_sigma;          /* 0.01 m */
  float jitter_cells;         /* 0.3 */
  float min_range, max_range; /* 0.5, 100 */
  float half;                 /* arena half size: 60 m */
  int32_t n_pillars;          /* 180 */
  int32_t n_walls;            /* 16 low interior walls */
  int32_t n_keyframes;        /* 500 */
  float spiral_pitch;         /* arm spacing = arc spacing between key-frame poses: 2.4 m */
  float spiral_r0;            /* radius of the first key-frame pose: 3 m */
  float speed, radius, dt;    /* driving circle: 1 m/s, 10 m, 0.1 s */
} ArenaConfig;

typedef struct ArenaBox { double lo[3], hi[3]; } ArenaBox;

/* everything one scan needs besides the boxes */
typedef struct ArenaScanCtx {
  double R[9];       /* world = R * sensor (Rz(yaw) Ry(pitch) Rx(roll)), row-major */
  double o[3];       /* sensor origin in the world */
  uint64_t seed, seq, key; /* noise key: (seed, sequence, frame key, ring, column) */
  int32_t V, H;
  double bottom_deg, top_deg, jitter_cells, range_sigma, min_range, max_range;
} ArenaScanCtx;

LL_HD uint64_t arena_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
LL_HD double arena_u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

LL_HD uint64_t arena_key5(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  uint64_t h = arena_splitmix64(seed);
  h = arena_splitmix64(h ^ (a * 0x100000001B3ULL + 1));
  h = arena_splitmix64(h ^ (b * 0x100000001B3ULL + 2));
  h = arena_splitmix64(h ^ (c * 0x100000001B3ULL + 3));
  h = arena_splitmix64(h ^ (d * 0x100000001B3ULL + 4));
  return h;
}

/* approximately N(0, 1): sum of twelve 16-bit uniforms, integer arithmetic up to the last step */
LL_HD double arena_gauss(uint64_t h) {
  const uint64_t a = arena_splitmix64(h ^ 0xD1B54A32D192ED03ULL), b = arena_splitmix64(a), c = arena_splitmix64(b);
  uint64_t s = 0;
  for (int k = 0; k < 4; ++k) s += ((a >> (16 * k)) & 0xFFFFu) + ((b >> (16 * k)) & 0xFFFFu) + ((c >> (16 * k)) & 0xFFFFu);
  return ((double)s + 6.0) * (1.0 / 65536.0) - 6.0;
}

LL_HD double arena_sqrt(double v) {
#if defined(__CUDA_ARCH__)
  return sqrt(v);
#else
  return __builtin_sqrt(v);
#endif
}

/* slab test; returns the entry distance (> 0) or -1 */
LL_HD double arena_ray_box(const double* o, const double* inv, const ArenaBox* b, double tbest) {
  double tmin = 0.0, tmax = tbest;
  for (int a = 0; a < 3; ++a) {
    double t1 = (b->lo[a] - o[a]) * inv[a], t2 = (b->hi[a] - o[a]) * inv[a];
    if (t1 > t2) { const double t = t1; t1 = t2; t2 = t; }
    if (t1 > tmin) tmin = t1;
    if (t2 < tmax) tmax = t2;
    if (!(tmin <= tmax)) return -1.0;
  }
  return tmin > 0.0 ? tmin : -1.0;
}

/* sensor-frame unit direction and world direction of beam (ring r, column c) */
LL_HD void arena_beam(const ArenaScanCtx* k, int r, int c, uint64_t* h_out, double* ds, double* dw) {
  const uint64_t h = arena_key5(k->seed, k->seq, k->key, (uint64_t)r, (uint64_t)c);
  const double jit = (arena_u01(h) * 2.0 - 1.0) * k->jitter_cells;
  const double elev = (k->bottom_deg + (double)r * (k->top_deg - k->bottom_deg) / (double)(k->V - 1)) * (LL_PI / 180.0);
  const double ha = LL_PI_2 - ((double)(c - k->H / 2) + jit) * (2.0 * LL_PI / (double)k->H);
  double se, ce, sh, ch;
  ll_sincosd(elev, &se, &ce);
  ll_sincosd(ha, &sh, &ch);
  ds[0] = ce * sh; ds[1] = ce * ch; ds[2] = se;  /* atan2(x, y) = ha */
  for (int a = 0; a < 3; ++a) dw[a] = k->R[3 * a + 0] * ds[0] + k->R[3 * a + 1] * ds[1] + k->R[3 * a + 2] * ds[2];
  *h_out = h;
}

/* One beam against `nb` candidate boxes (idx == NULL: boxes[0..nb-1]; else boxes[idx[0..nb-1]]; the result does not
 * depend on which superset of the boxes the ray can hit is given, nor on their order).  Returns 1 and the point
 * (sensor frame) when the beam has a return. */
LL_HD int arena_ray(const ArenaScanCtx* k, const ArenaBox* boxes, const int* idx, int nb, uint64_t h, const double* ds,
                    const double* dw, float* out_xyz) {
  double t = 1e30;
  if (dw[2] < -1e-9) t = (ARENA_GROUND_Z - k->o[2]) / dw[2];
  double inv[3];
  for (int a = 0; a < 3; ++a) {
    const double ad = dw[a] < 0.0 ? -dw[a] : dw[a];
    inv[a] = ad < 1e-12 ? (dw[a] < 0.0 ? -1e300 : 1e300) : 1.0 / dw[a];
  }
  for (int b = 0; b < nb; ++b) {
    const double tb = arena_ray_box(k->o, inv, &boxes[idx ? idx[b] : b], t);
    if (tb > 0.0 && tb < t) t = tb;
  }
  if (t >= 1e29) return 0;
  const double rng = t + k->range_sigma * arena_gauss(h);
  if (rng > k->max_range || rng < k->min_range) return 0;
  out_xyz[0] = (float)(rng * ds[0]);
  out_xyz[1] = (float)(rng * ds[1]);
  out_xyz[2] = (float)(rng * ds[2]);
  return 1;
}

#endif /* SYNTH_ARENA_H */
