/*
 * synth_lidar.c -- deterministic synthetic spinning-lidar sequences.
 *
 * Not part of the hot path: it only manufactures the inputs that the tests, the CPU oracle
 * and bench.py feed to ImageProjection (the reference ships no data and no tests).
 *
 * World: ground plane z = -1.5 m (sensor frame is level, 1.5 m above ground), a rectangular
 * room of 8 m high walls, square pillars and a few low interior walls, all axis-aligned
 * boxes.  The sensor drives a circle (1 m/s, radius 10 m) so every frame moves 0.1 m and
 * yaws 0.01 rad.  Beam (ring r, column c): elevation bottom + r*(top-bottom)/(V-1), azimuth
 * such that atan2(x, y) = pi/2 - (c - H/2 + jitter) * 2*pi/H with |jitter| <= 0.3 cell, which
 * keeps the reference's round() at imageProjection.cpp:200 away from a cell boundary.
 * Range = exact hit + N(0, sigma) from a counter-based generator keyed
 * (seed, sequence, frame, ring, column).  Hits beyond max_range or below min_range give no
 * return.  No motion distortion inside a frame.  Points are emitted in firing order
 * (column H-1 .. 0, ring 0 .. V-1), so the first / last point define start / end orientation
 * (imageProjection.cpp:236-240).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct SynthConfig {
  int32_t V, H;
  float bottom_deg, top_deg;
  uint64_t seed;
  float range_sigma;      /* 0.01 */
  float jitter_cells;     /* 0.3  */
  float min_range, max_range; /* 0.5, 100 */
  float room_half_x, room_half_y; /* 30, 20 */
  int32_t n_pillars;      /* 30 */
  float speed, radius;    /* 1.0 m/s, 10 m */
  float dt;               /* 0.1 s */
} SynthConfig;

typedef struct Box { double lo[3], hi[3]; } Box;

#define MAX_BOXES 128
#define GROUND_Z (-1.5)

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
static double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

static uint64_t key5(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  uint64_t h = splitmix64(seed);
  h = splitmix64(h ^ (a * 0x100000001B3ULL + 1));
  h = splitmix64(h ^ (b * 0x100000001B3ULL + 2));
  h = splitmix64(h ^ (c * 0x100000001B3ULL + 3));
  h = splitmix64(h ^ (d * 0x100000001B3ULL + 4));
  return h;
}

static int build_world(const SynthConfig* cfg, int seq, Box* boxes) {
  int n = 0;
  const double hx = cfg->room_half_x, hy = cfg->room_half_y;
  const double top = GROUND_Z + 8.0, th = 0.5;
  /* four outer walls as thick slabs */
  Box w;
  w.lo[2] = GROUND_Z; w.hi[2] = top;
  w.lo[0] = hx; w.hi[0] = hx + th; w.lo[1] = -hy - th; w.hi[1] = hy + th; boxes[n++] = w;
  w.lo[0] = -hx - th; w.hi[0] = -hx; boxes[n++] = w;
  w.lo[0] = -hx - th; w.hi[0] = hx + th; w.lo[1] = hy; w.hi[1] = hy + th; boxes[n++] = w;
  w.lo[1] = -hy - th; w.hi[1] = -hy; boxes[n++] = w;
  /* pillars: 0.5 m square, 5.5 m tall, kept off the driving circle */
  int placed = 0, tries = 0;
  while (placed < cfg->n_pillars && tries < 10000 && n < MAX_BOXES - 8) {
    const uint64_t h = key5(cfg->seed, (uint64_t)seq, 0xA11CEULL, (uint64_t)tries, 0);
    const double px = (u01(h) * 2.0 - 1.0) * (hx - 2.0);
    const double py = (u01(splitmix64(h)) * 2.0 - 1.0) * (hy - 2.0);
    ++tries;
    const double rr = sqrt(px * px + py * py);
    if (fabs(rr - cfg->radius) < 2.5) continue;
    Box b;
    b.lo[0] = px - 0.25; b.hi[0] = px + 0.25; b.lo[1] = py - 0.25; b.hi[1] = py + 0.25;
    b.lo[2] = GROUND_Z; b.hi[2] = GROUND_Z + 5.5;
    boxes[n++] = b;
    ++placed;
  }
  /* four low interior walls, 0.3 m thick, 6 m long, 2.5 m tall */
  for (int k = 0; k < 4; ++k) {
    const uint64_t h = key5(cfg->seed, (uint64_t)seq, 0xBA11ULL, (uint64_t)k, 0);
    const double ang = (k + u01(h)) * (M_PI / 2.0);
    const double rad = (k % 2 == 0) ? cfg->radius + 5.0 : cfg->radius - 5.0;
    const double cx = rad * cos(ang), cy = rad * sin(ang);
    Box b;
    if (k % 2 == 0) { b.lo[0] = cx - 3.0; b.hi[0] = cx + 3.0; b.lo[1] = cy - 0.15; b.hi[1] = cy + 0.15; }
    else { b.lo[0] = cx - 0.15; b.hi[0] = cx + 0.15; b.lo[1] = cy - 3.0; b.hi[1] = cy + 3.0; }
    if (fabs(b.lo[0]) > hx - 1 || fabs(b.hi[0]) > hx - 1 || fabs(b.lo[1]) > hy - 1 || fabs(b.hi[1]) > hy - 1) continue;
    b.lo[2] = GROUND_Z; b.hi[2] = GROUND_Z + 2.5;
    boxes[n++] = b;
  }
  return n;
}

/* Sensor pose of `frame` in sequence `seq`: x, y, z, roll, pitch, yaw (world = Rz Ry Rx).
 * A slow roll / pitch / height wobble keeps the ground-plane LM stage (rx, rz, ty) busy. */
void synth_pose(const SynthConfig* cfg, int seq, int frame, double* pose6) {
  const double phi0 = u01(key5(cfg->seed, (uint64_t)seq, 0x9057ULL, 0, 0)) * 2.0 * M_PI;
  const double w = cfg->speed / cfg->radius;
  const double t = cfg->dt * frame;
  const double phi = phi0 + w * t;
  pose6[0] = cfg->radius * cos(phi);
  pose6[1] = cfg->radius * sin(phi);
  pose6[2] = 0.03 * sin(1.3 * t + phi0);
  pose6[3] = 0.010 * sin(0.9 * t + 2.0 * phi0);
  pose6[4] = 0.008 * sin(1.1 * t + 3.0 * phi0);
  pose6[5] = phi + M_PI / 2.0;
}

/* slab test with a precomputed reciprocal direction (inf where the direction component is 0) */
static inline double ray_box(const double o[3], const double inv[3], const Box* b, double tbest) {
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

/* Writes up to V*H points (x, y, z, intensity=0) in firing order; returns the count. */
int synth_scan(const SynthConfig* cfg, int seq, int frame, float* out_xyzi) {
  Box boxes[MAX_BOXES];
  const int nb = build_world(cfg, seq, boxes);
  double pose[6];
  synth_pose(cfg, seq, frame, pose);
  const double cr = cos(pose[3]), sr = sin(pose[3]), cp = cos(pose[4]), sp = sin(pose[4]);
  const double cy = cos(pose[5]), sy = sin(pose[5]);
  /* R = Rz(yaw) Ry(pitch) Rx(roll) */
  const double R[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
                          {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
                          {-sp, cp * sr, cp * cr}};
  const double o[3] = {pose[0], pose[1], pose[2]};
  const int V = cfg->V, H = cfg->H;
  const double resx = 2.0 * M_PI / H;
  int n = 0;
  for (int c = H - 1; c >= 0; --c) {
    for (int r = 0; r < V; ++r) {
      const uint64_t h = key5(cfg->seed, (uint64_t)seq, (uint64_t)frame, (uint64_t)r, (uint64_t)c);
      const double jit = (u01(h) * 2.0 - 1.0) * cfg->jitter_cells;
      const double elev = (cfg->bottom_deg + r * (double)(cfg->top_deg - cfg->bottom_deg) / (V - 1)) * (M_PI / 180.0);
      const double ha = M_PI / 2.0 - (c - H / 2 + jit) * resx;
      /* sensor-frame unit direction: atan2(x, y) = ha */
      const double ds[3] = {cos(elev) * sin(ha), cos(elev) * cos(ha), sin(elev)};
      const double dw[3] = {R[0][0] * ds[0] + R[0][1] * ds[1] + R[0][2] * ds[2],
                            R[1][0] * ds[0] + R[1][1] * ds[1] + R[1][2] * ds[2],
                            R[2][0] * ds[0] + R[2][1] * ds[1] + R[2][2] * ds[2]};
      double t = 1e30;
      if (dw[2] < -1e-9) t = (GROUND_Z - o[2]) / dw[2];
      double inv[3];
      for (int a = 0; a < 3; ++a) inv[a] = fabs(dw[a]) < 1e-12 ? (dw[a] < 0 ? -1e300 : 1e300) : 1.0 / dw[a];
      for (int b = 0; b < nb; ++b) {
        const double tb = ray_box(o, inv, &boxes[b], t);
        if (tb > 0.0 && tb < t) t = tb;
      }
      if (t >= 1e29) continue;
      /* Box-Muller on two further hashes */
      const uint64_t h2 = splitmix64(h ^ 0xD1B54A32D192ED03ULL);
      const double g = sqrt(-2.0 * log(u01(h2))) * cos(2.0 * M_PI * u01(splitmix64(h2)));
      const double rng = t + cfg->range_sigma * g;
      if (rng > cfg->max_range || rng < cfg->min_range) continue;
      out_xyzi[4 * n + 0] = (float)(rng * ds[0]);
      out_xyzi[4 * n + 1] = (float)(rng * ds[1]);
      out_xyzi[4 * n + 2] = (float)(rng * ds[2]);
      out_xyzi[4 * n + 3] = 0.0f;
      ++n;
    }
  }
  return n;
}

/* Points sampled on the world's surfaces on a regular lattice (spacing `step`), in the world
 * frame mapped to the reference's camera axes (x<-y, y<-z, z<-x), with N(0, sigma) noise:
 * a stand-in for a down-sampled local map assembled from many key-frames.
 * kind 0: planar surfaces (ground + faces); kind 1: vertical edges of pillars and walls.
 * Returns the count written (<= cap). */
int synth_local_map(const SynthConfig* cfg, int seq, int kind, float step, float sigma, float radius_limit,
                    float* out_xyzi, int cap) {
  Box boxes[MAX_BOXES];
  const int nb = build_world(cfg, seq, boxes);
  int n = 0;
  uint64_t ctr = 0;
#define EMIT(X, Y, Z)                                                                    \
  do {                                                                                   \
    const double ex = (X), ey = (Y), ez = (Z);                                           \
    if (n < cap && ex * ex + ey * ey <= (double)radius_limit * radius_limit) {           \
      const uint64_t hh = key5(cfg->seed, (uint64_t)seq, 0x3A9ULL + (uint64_t)kind, ctr, 7); \
      const double g1 = sqrt(-2.0 * log(u01(hh))) * cos(2.0 * M_PI * u01(splitmix64(hh)));   \
      const double g2 = sqrt(-2.0 * log(u01(hh))) * sin(2.0 * M_PI * u01(splitmix64(hh)));   \
      const uint64_t h3 = splitmix64(hh ^ 0x5851F42D4C957F2DULL);                          \
      const double g3 = sqrt(-2.0 * log(u01(h3))) * cos(2.0 * M_PI * u01(splitmix64(h3)));   \
      const double wx = ex + sigma * g1, wy = ey + sigma * g2, wz = ez + sigma * g3;      \
      out_xyzi[4 * n + 0] = (float)wy;                                                   \
      out_xyzi[4 * n + 1] = (float)wz;                                                   \
      out_xyzi[4 * n + 2] = (float)wx;                                                   \
      out_xyzi[4 * n + 3] = 0.0f;                                                        \
      ++n;                                                                               \
    }                                                                                    \
    ++ctr;                                                                               \
  } while (0)
  if (kind == 0) {
    const double hx = cfg->room_half_x, hy = cfg->room_half_y;
    for (double x = -hx + step * 0.5; x < hx; x += step)
      for (double y = -hy + step * 0.5; y < hy; y += step) EMIT(x, y, GROUND_Z);
    for (int b = 0; b < nb; ++b) {
      const Box* B = &boxes[b];
      for (double z = B->lo[2] + step * 0.5; z < B->hi[2]; z += step) {
        for (double x = B->lo[0] + step * 0.5; x < B->hi[0]; x += step) { EMIT(x, B->lo[1], z); EMIT(x, B->hi[1], z); }
        for (double y = B->lo[1] + step * 0.5; y < B->hi[1]; y += step) { EMIT(B->lo[0], y, z); EMIT(B->hi[0], y, z); }
      }
    }
  } else {
    for (int b = 4; b < nb; ++b) { /* skip the outer walls: their corners are the room corners */
      const Box* B = &boxes[b];
      for (double z = B->lo[2] + step * 0.5; z < B->hi[2]; z += step) {
        EMIT(B->lo[0], B->lo[1], z); EMIT(B->lo[0], B->hi[1], z);
        EMIT(B->hi[0], B->lo[1], z); EMIT(B->hi[0], B->hi[1], z);
      }
    }
    const double hx = cfg->room_half_x, hy = cfg->room_half_y;
    for (double z = GROUND_Z + step * 0.5; z < GROUND_Z + 8.0; z += step) {
      EMIT(hx, hy, z); EMIT(-hx, hy, z); EMIT(hx, -hy, z); EMIT(-hx, -hy, z);
    }
  }
#undef EMIT
  return n;
}
