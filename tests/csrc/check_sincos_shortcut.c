/* CPU check of the certified short cut of ll_sincosf (include/ll_portable_math.h, device-only there): the same
 * fused-multiply-add chains evaluated with the host's fma(), compared with the portable routine over a dense sample
 * of float angles in (-0.25, 0.25).  Prints "accepted rejected mismatches". */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ll_portable_math.h"

static int shortcut(float x, float* s, float* c) {
  if (!(fabsf(x) < 0.25f)) return 0;
  const double r = (double)x, z = r * r;
  double ps = -1.0 / 39916800.0;
  ps = fma(z, ps, 1.0 / 362880.0);
  ps = fma(z, ps, -1.0 / 5040.0);
  ps = fma(z, ps, 1.0 / 120.0);
  ps = fma(z, ps, -1.0 / 6.0);
  const double ds = fma(r * z, ps, r);
  double pc = 1.0 / 479001600.0;
  pc = fma(z, pc, -1.0 / 3628800.0);
  pc = fma(z, pc, 1.0 / 40320.0);
  pc = fma(z, pc, -1.0 / 720.0);
  pc = fma(z, pc, 1.0 / 24.0);
  pc = fma(z, pc, -0.5);
  const double dc = fma(z, pc, 1.0);
  const float sf = (float)ds, cf = (float)dc;
  const double es = fabs(ds) * 1e-15, ec = 1e-15;
  if ((float)(ds - es) != sf || (float)(ds + es) != sf || (float)(dc - ec) != cf || (float)(dc + ec) != cf) return 0;
  *s = sf;
  *c = cf;
  return 1;
}

int main(void) {
  long accepted = 0, rejected = 0, mismatches = 0;
  double max_rel = 0.0;
  /* every 37th float bit pattern of [2^-40, 0.25), both signs, plus zero and the smallest normals */
  const uint32_t lo = 0x2B800000u, hi = 0x3E800000u;
  for (uint32_t b = lo; b < hi; b += 37) {
    for (int sign = 0; sign < 2; ++sign) {
      uint32_t bits = b | (sign ? 0x80000000u : 0u);
      float x, s1, c1, s2, c2;
      memcpy(&x, &bits, 4);
      ll_sincosf(x, &s2, &c2);
      if (!shortcut(x, &s1, &c1)) { ++rejected; continue; }
      ++accepted;
      if (memcmp(&s1, &s2, 4) || memcmp(&c1, &c2, 4)) ++mismatches;
    }
  }
  const float specials[] = {0.0f, -0.0f, 1e-38f, -1e-38f, 1e-45f, 0.24999999f, -0.24999999f};
  for (unsigned i = 0; i < sizeof(specials) / sizeof(specials[0]); ++i) {
    float s1, c1, s2, c2;
    ll_sincosf(specials[i], &s2, &c2);
    if (shortcut(specials[i], &s1, &c1)) { ++accepted; if (memcmp(&s1, &s2, 4) || memcmp(&c1, &c2, 4)) ++mismatches; } else ++rejected;
  }
  (void)max_rel;
  printf("%ld %ld %ld\n", accepted, rejected, mismatches);
  return mismatches ? 1 : 0;
}
