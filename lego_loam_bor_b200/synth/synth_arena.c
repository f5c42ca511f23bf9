/*
 * synth_arena.c -- host side of the arena world (synth_arena.h): world layout, key-frame / driving poses, and the
 * scan generator on the CPU.  Inputs only, not part of the hot path.
 */
#include "synth_arena.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- poses -------------------------------------------------------------------------------------------------- */

static double arena_phi0(const ArenaConfig* cfg, int seq) {
  return arena_u01(arena_key5(cfg->seed, (uint64_t)seq, 0x9057ULL, 0, 0)) * 2.0 * LL_PI;
}

static void wobble(double t, double phi0, double* pose6) {
  double s, c;
  ll_sincosd(1.3 * t + phi0, &s, &c); pose6[2] = 0.03 * s;
  ll_sincosd(0.9 * t + 2.0 * phi0, &s, &c); pose6[3] = 0.010 * s;
  ll_sincosd(1.1 * t + 3.0 * phi0, &s, &c); pose6[4] = 0.008 * s;
}

/* Key-frame pose i of sequence seq: x, y, z, roll, pitch, yaw.  Archimedean spiral r = a * theta with arm spacing
 * `spiral_pitch`, walked in arc steps of `spiral_pitch` (so neighbours along and across the arms are ~pitch apart),
 * rotated by the sequence's phase; heading along the tangent. */
void arena_keyframe_pose(const ArenaConfig* cfg, int seq, int i, double* pose6) {
  const double a = (double)cfg->spiral_pitch / (2.0 * LL_PI);
  double theta = (double)cfg->spiral_r0 / a;
  for (int k = 0; k < i; ++k) {
    const double r = a * theta;
    theta += (double)cfg->spiral_pitch / arena_sqrt(r * r + a * a);
  }
  const double phi0 = arena_phi0(cfg, seq);
  const double r = a * theta;
  double s, c;
  ll_sincosd(theta + phi0, &s, &c);
  pose6[0] = r * c;
  pose6[1] = r * s;
  const double dx = a * c - r * s, dy = a * s + r * c;
  pose6[5] = ll_atan2d(dy, dx);
  wobble(0.37 * (double)i, phi0, pose6);
}

/* Pose of frame f of the driving sequence: the circle of synth_lidar.c (radius 10 m, 1 m/s), through the spiral. */
void arena_drive_pose(const ArenaConfig* cfg, int seq, int frame, double* pose6) {
  const double phi0 = arena_phi0(cfg, seq);
  const double w = (double)cfg->speed / (double)cfg->radius;
  const double t = (double)cfg->dt * (double)frame;
  const double phi = phi0 + w * t;
  double s, c;
  ll_sincosd(phi, &s, &c);
  pose6[0] = (double)cfg->radius * c;
  pose6[1] = (double)cfg->radius * s;
  pose6[5] = phi + LL_PI_2;
  wobble(t, phi0, pose6);
}

/* ---- world -------------------------------------------------------------------------------------------------- */

static double dist_point_box_xy(double x, double y, const ArenaBox* b) {
  const double dx = x < b->lo[0] ? b->lo[0] - x : (x > b->hi[0] ? x - b->hi[0] : 0.0);
  const double dy = y < b->lo[1] ? b->lo[1] - y : (y > b->hi[1] ? y - b->hi[1] : 0.0);
  return arena_sqrt(dx * dx + dy * dy);
}

/* 1 when the box keeps `clear` metres from every key-frame pose and from the driving circle */
static int box_is_clear(const ArenaConfig* cfg, const double* kf_xy, const ArenaBox* b, double clear) {
  for (int i = 0; i < cfg->n_keyframes; ++i)
    if (dist_point_box_xy(kf_xy[2 * i], kf_xy[2 * i + 1], b) < clear) return 0;
  for (int k = 0; k < 720; ++k) {
    double s, c;
    ll_sincosd(2.0 * LL_PI * (double)k / 720.0, &s, &c);
    if (dist_point_box_xy((double)cfg->radius * c, (double)cfg->radius * s, b) < clear) return 0;
  }
  return 1;
}

/* boxes[0..3] are the outer walls; returns the number of boxes (<= ARENA_MAX_BOXES) */
int arena_build_world(const ArenaConfig* cfg, int seq, ArenaBox* boxes) {
  int n = 0;
  const double hx = cfg->half, hy = cfg->half, top = ARENA_GROUND_Z + 8.0, th = 0.5;
  ArenaBox w;
  w.lo[2] = ARENA_GROUND_Z; w.hi[2] = top;
  w.lo[0] = hx; w.hi[0] = hx + th; w.lo[1] = -hy - th; w.hi[1] = hy + th; boxes[n++] = w;
  w.lo[0] = -hx - th; w.hi[0] = -hx; boxes[n++] = w;
  w.lo[0] = -hx - th; w.hi[0] = hx + th; w.lo[1] = hy; w.hi[1] = hy + th; boxes[n++] = w;
  w.lo[1] = -hy - th; w.hi[1] = -hy; boxes[n++] = w;
  double* kf_xy = (double*)malloc(sizeof(double) * 2 * (size_t)(cfg->n_keyframes > 0 ? cfg->n_keyframes : 1));
  for (int i = 0; i < cfg->n_keyframes; ++i) {
    double p[6];
    arena_keyframe_pose(cfg, seq, i, p);
    kf_xy[2 * i] = p[0]; kf_xy[2 * i + 1] = p[1];
  }
  /* pillars: 0.5 m square, 5.5 m tall */
  int placed = 0, tries = 0;
  while (placed < cfg->n_pillars && tries < 100000 && n < ARENA_MAX_BOXES) {
    const uint64_t h = arena_key5(cfg->seed, (uint64_t)seq, 0xA11CEULL, (uint64_t)tries, 1);
    const double px = (arena_u01(h) * 2.0 - 1.0) * (hx - 2.0);
    const double py = (arena_u01(arena_splitmix64(h)) * 2.0 - 1.0) * (hy - 2.0);
    ++tries;
    ArenaBox b;
    b.lo[0] = px - 0.25; b.hi[0] = px + 0.25; b.lo[1] = py - 0.25; b.hi[1] = py + 0.25;
    b.lo[2] = ARENA_GROUND_Z; b.hi[2] = ARENA_GROUND_Z + 5.5;
    if (!box_is_clear(cfg, kf_xy, &b, 0.9)) continue;
    boxes[n++] = b;
    ++placed;
  }
  /* low interior walls: 0.3 m thick, 6 m long, 2.5 m tall, alternating orientation */
  placed = 0; tries = 0;
  while (placed < cfg->n_walls && tries < 100000 && n < ARENA_MAX_BOXES) {
    const uint64_t h = arena_key5(cfg->seed, (uint64_t)seq, 0xBA11ULL, (uint64_t)tries, 1);
    const double cx = (arena_u01(h) * 2.0 - 1.0) * (hx - 5.0);
    const double cy = (arena_u01(arena_splitmix64(h)) * 2.0 - 1.0) * (hy - 5.0);
    ++tries;
    ArenaBox b;
    if (placed % 2 == 0) { b.lo[0] = cx - 3.0; b.hi[0] = cx + 3.0; b.lo[1] = cy - 0.15; b.hi[1] = cy + 0.15; }
    else { b.lo[0] = cx - 0.15; b.hi[0] = cx + 0.15; b.lo[1] = cy - 3.0; b.hi[1] = cy + 3.0; }
    b.lo[2] = ARENA_GROUND_Z; b.hi[2] = ARENA_GROUND_Z + 2.5;
    if (!box_is_clear(cfg, kf_xy, &b, 1.2)) continue;
    boxes[n++] = b;
    ++placed;
  }
  free(kf_xy);
  return n;
}

/* ---- scans -------------------------------------------------------------------------------------------------- */

void arena_scan_ctx(const ArenaConfig* cfg, int seq, uint64_t key, const double* pose, ArenaScanCtx* k) {
  double sr, cr, sp, cp, sy, cy;
  ll_sincosd(pose[3], &sr, &cr);
  ll_sincosd(pose[4], &sp, &cp);
  ll_sincosd(pose[5], &sy, &cy);
  const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                       sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                       -sp, cp * sr, cp * cr};
  memcpy(k->R, R, sizeof(R));
  k->o[0] = pose[0]; k->o[1] = pose[1]; k->o[2] = pose[2];
  k->seed = cfg->seed; k->seq = (uint64_t)seq; k->key = key;
  k->V = cfg->V; k->H = cfg->H;
  k->bottom_deg = cfg->bottom_deg; k->top_deg = cfg->top_deg; k->jitter_cells = cfg->jitter_cells;
  k->range_sigma = cfg->range_sigma; k->min_range = cfg->min_range; k->max_range = cfg->max_range;
}

/* frame key of the noise stream: driving frames use their frame number, key frames 1 000 000 + index */
uint64_t arena_frame_key(int kind, int index) { return kind == 0 ? 1000000ull + (uint64_t)index : (uint64_t)index; }

void arena_pose(const ArenaConfig* cfg, int seq, int kind, int index, double* pose6) {
  if (kind == 0) arena_keyframe_pose(cfg, seq, index, pose6);
  else arena_drive_pose(cfg, seq, index, pose6);
}

#define ARENA_SECTORS 512

/* Scan from `pose` against the given world.  Writes up to V*H points (x, y, z, 0) in firing order (column H-1..0,
 * ring 0..V-1); returns the count.  Boxes are culled per azimuth sector of the ray (conservatively: the result is the
 * same as testing every box, which is what the device generator does). */
int arena_scan_boxes(const ArenaConfig* cfg, const ArenaBox* boxes, int nb, int seq, uint64_t key, const double* pose,
                     float* out_xyzi) {
  ArenaScanCtx k;
  arena_scan_ctx(cfg, seq, key, pose, &k);
  /* sector lists: bearing interval of every box seen from the sensor, widened by two sectors */
  static __thread int* lists = NULL;
  static __thread int* counts = NULL;
  if (!lists) {
    lists = (int*)malloc(sizeof(int) * ARENA_SECTORS * ARENA_MAX_BOXES);
    counts = (int*)malloc(sizeof(int) * ARENA_SECTORS);
  }
  memset(counts, 0, sizeof(int) * ARENA_SECTORS);
  const double sec = 2.0 * M_PI / ARENA_SECTORS;
  for (int b = 0; b < nb; ++b) {
    const ArenaBox* B = &boxes[b];
    int all = k.o[0] >= B->lo[0] - 0.05 && k.o[0] <= B->hi[0] + 0.05 && k.o[1] >= B->lo[1] - 0.05 && k.o[1] <= B->hi[1] + 0.05;
    int s_lo = 0, s_hi = ARENA_SECTORS - 1;
    if (!all) {
      const double cxm = 0.5 * (B->lo[0] + B->hi[0]), cym = 0.5 * (B->lo[1] + B->hi[1]);
      const double bc = atan2(cym - k.o[1], cxm - k.o[0]);
      double dmax = 0.0;
      for (int q = 0; q < 4; ++q) {
        const double x = (q & 1) ? B->hi[0] : B->lo[0], y = (q & 2) ? B->hi[1] : B->lo[1];
        double d = atan2(y - k.o[1], x - k.o[0]) - bc;
        while (d > M_PI) d -= 2.0 * M_PI;
        while (d < -M_PI) d += 2.0 * M_PI;
        if (fabs(d) > dmax) dmax = fabs(d);
      }
      if (dmax > 1.5) all = 1;
      else {
        s_lo = (int)floor((bc - dmax + M_PI) / sec) - 2;
        s_hi = (int)floor((bc + dmax + M_PI) / sec) + 2;
      }
    }
    if (all) { s_lo = 0; s_hi = ARENA_SECTORS - 1; }
    if (s_hi - s_lo >= ARENA_SECTORS - 1) { s_lo = 0; s_hi = ARENA_SECTORS - 1; }
    for (int s = s_lo; s <= s_hi; ++s) {
      const int w = ((s % ARENA_SECTORS) + ARENA_SECTORS) % ARENA_SECTORS;
      lists[w * ARENA_MAX_BOXES + counts[w]++] = b;
    }
  }
  int n = 0;
  for (int c = k.H - 1; c >= 0; --c) {
    for (int r = 0; r < k.V; ++r) {
      uint64_t h;
      double ds[3], dw[3];
      arena_beam(&k, r, c, &h, ds, dw);
      int w = (int)floor((atan2(dw[1], dw[0]) + M_PI) / sec);
      w = ((w % ARENA_SECTORS) + ARENA_SECTORS) % ARENA_SECTORS;
      if (arena_ray(&k, boxes, lists + w * ARENA_MAX_BOXES, counts[w], h, ds, dw, out_xyzi + 4 * n)) {
        out_xyzi[4 * n + 3] = 0.0f;
        ++n;
      }
    }
  }
  return n;
}

/* kind 0: key frame `index`; kind 1: driving frame `index` */
int arena_scan(const ArenaConfig* cfg, int seq, int kind, int index, float* out_xyzi) {
  static __thread ArenaBox* boxes = NULL;
  static __thread int cached_seq = -1, cached_nb = 0;
  static __thread ArenaConfig cached_cfg;
  if (!boxes) boxes = (ArenaBox*)malloc(sizeof(ArenaBox) * ARENA_MAX_BOXES);
  if (cached_seq != seq || memcmp(&cached_cfg, cfg, sizeof(ArenaConfig)) != 0) {
    cached_nb = arena_build_world(cfg, seq, boxes);
    cached_seq = seq;
    cached_cfg = *cfg;
  }
  double pose[6];
  arena_pose(cfg, seq, kind, index, pose);
  return arena_scan_boxes(cfg, boxes, cached_nb, seq, arena_frame_key(kind, index), pose, out_xyzi);
}

/* reference-free check helper: the same scan with every box tested for every ray (what the device does) */
int arena_scan_bruteforce(const ArenaConfig* cfg, int seq, int kind, int index, float* out_xyzi) {
  ArenaBox* boxes = (ArenaBox*)malloc(sizeof(ArenaBox) * ARENA_MAX_BOXES);
  const int nb = arena_build_world(cfg, seq, boxes);
  double pose[6];
  arena_pose(cfg, seq, kind, index, pose);
  ArenaScanCtx k;
  arena_scan_ctx(cfg, seq, arena_frame_key(kind, index), pose, &k);
  int n = 0;
  for (int c = k.H - 1; c >= 0; --c)
    for (int r = 0; r < k.V; ++r) {
      uint64_t h;
      double ds[3], dw[3];
      arena_beam(&k, r, c, &h, ds, dw);
      if (arena_ray(&k, boxes, NULL, nb, h, ds, dw, out_xyzi + 4 * n)) { out_xyzi[4 * n + 3] = 0.0f; ++n; }
    }
  free(boxes);
  return n;
}
