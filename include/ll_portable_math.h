/*
 * ll_portable_math.h -- deterministic float libm replacements shared by the CUDA
 * kernels (nvcc, -fmad=false) and by the CPU oracle (gcc, -ffp-contract=off).
 *
 * Why this exists (hard part 1): the reference decides range-image
 * rows/columns, ground flags and the half-scan flag with std::asin / std::atan2 /
 * atan2 on float arguments (imageProjection.cpp:190,198,237-240,278;
 * featureAssociation.cpp:172).  glibc's and CUDA's float libm disagree in the last
 * ulp, which flips discrete decisions at thresholds.  Every function below is
 * built only from IEEE-754 double + - * / and sqrt, evaluated in one fixed order,
 * so gcc and nvcc produce the same bits.  The double result is rounded once to
 * float, which makes the value correctly rounded except in ~1e-8 of cases --
 * i.e. within the same <=1 ulp envelope glibc's own float functions have.
 *
 * The oracle can also be built with -DORACLE_LIBM to call glibc instead; the CPU
 * tests check that both oracle builds agree on every discrete output of the
 * fixtures (tests/test_oracle_pins.py).
 */
#ifndef LL_PORTABLE_MATH_H
#define LL_PORTABLE_MATH_H

#if defined(__CUDACC__)
#define LL_HD __host__ __device__ __forceinline__
#else
#define LL_HD static inline
#endif

#define LL_PI 3.14159265358979323846
#define LL_PI_2 1.57079632679489661923
#define LL_PI_4 0.78539816339744830962

/* atan(t) for |t| <= tan(pi/8): odd Taylor series in Horner form, truncated after
 * t^29/29 (remainder < 0.4143^31/31 = 4e-14). */
LL_HD double ll_atan_small(double t) {
  const double z = t * t;
  double p = 1.0 / 29.0;
  p = 1.0 / 27.0 - z * p;
  p = 1.0 / 25.0 - z * p;
  p = 1.0 / 23.0 - z * p;
  p = 1.0 / 21.0 - z * p;
  p = 1.0 / 19.0 - z * p;
  p = 1.0 / 17.0 - z * p;
  p = 1.0 / 15.0 - z * p;
  p = 1.0 / 13.0 - z * p;
  p = 1.0 / 11.0 - z * p;
  p = 1.0 / 9.0 - z * p;
  p = 1.0 / 7.0 - z * p;
  p = 1.0 / 5.0 - z * p;
  p = 1.0 / 3.0 - z * p;
  p = 1.0 - z * p;
  return t * p;
}

/* atan(a) for a in [0, 1]. */
LL_HD double ll_atan01(double a) {
  if (a > 0.41421356237309503) {
    const double t = (a - 1.0) / (a + 1.0);
    return LL_PI_4 + ll_atan_small(t);
  }
  return ll_atan_small(a);
}

/* atan2 in double for finite inputs; the float wrappers below round it once. */
LL_HD double ll_atan2d(double y, double x) {
  if (x != x || y != y) return x + y; /* NaN in, NaN out */
  const double ax = x < 0.0 ? -x : x;
  const double ay = y < 0.0 ? -y : y;
  double r;
  if (ax == 0.0 && ay == 0.0) {
    r = 0.0;
  } else if (ay <= ax) {
    r = ll_atan01(ay / ax);
  } else {
    r = LL_PI_2 - ll_atan01(ax / ay);
  }
  if (x < 0.0) r = LL_PI - r;
  if (y < 0.0) r = -r;
  return r;
}

LL_HD float ll_atan2f(float y, float x) { return (float)ll_atan2d((double)y, (double)x); }

/* asin(v) = atan2(v, sqrt((1-v)(1+v))); NaN outside [-1,1] like libm. */
LL_HD float ll_asinf(float v) {
  const double d = (double)v;
  const double w = (1.0 - d) * (1.0 + d);
  if (!(w >= 0.0)) return (float)((d - d) / (d - d)); /* NaN */
#if defined(__CUDA_ARCH__)
  const double s = sqrt(w);
#else
  const double s = __builtin_sqrt(w);
#endif
  return (float)ll_atan2d(d, s);
}

/* sin and cos of a float angle.  Cody-Waite reduction by pi/2 (the high part
 * has 33 significant bits, so k*hi is exact for |k| < 2^20), then Taylor
 * polynomials on [-pi/4, pi/4].  Intended for |x| < 1e5. */
LL_HD void ll_sincosd(double x, double* s_out, double* c_out) {
  const double two_over_pi = 0.63661977236758134308;
  const double pio2_hi = 1.57079632673412561417e+00; /* 0x3FF921FB54400000 */
  const double pio2_lo = 6.07710050650619224932e-11; /* pi/2 - pio2_hi */
  const double v = x * two_over_pi;
  const long long k = (long long)(v + (v >= 0.0 ? 0.5 : -0.5));
  const double kd = (double)k;
  const double r = (x - kd * pio2_hi) - kd * pio2_lo;
  const double z = r * r;
  /* sin(r) = r - r^3/3! + ... - r^15/15! */
  double ps = -1.0 / 1307674368000.0;
  ps = 1.0 / 6227020800.0 + z * ps;
  ps = -1.0 / 39916800.0 + z * ps;
  ps = 1.0 / 362880.0 + z * ps;
  ps = -1.0 / 5040.0 + z * ps;
  ps = 1.0 / 120.0 + z * ps;
  ps = -1.0 / 6.0 + z * ps;
  const double sr = r + r * (z * ps);
  /* cos(r) = 1 - r^2/2! + ... + r^16/16! */
  double pc = 1.0 / 20922789888000.0;
  pc = -1.0 / 87178291200.0 + z * pc;
  pc = 1.0 / 479001600.0 + z * pc;
  pc = -1.0 / 3628800.0 + z * pc;
  pc = 1.0 / 40320.0 + z * pc;
  pc = -1.0 / 720.0 + z * pc;
  pc = 1.0 / 24.0 + z * pc;
  pc = -0.5 + z * pc;
  const double cr = 1.0 + z * pc;
  const int q = (int)(k & 3LL);
  double s, c;
  if (q == 0) {
    s = sr; c = cr;
  } else if (q == 1) {
    s = cr; c = -sr;
  } else if (q == 2) {
    s = -sr; c = -cr;
  } else {
    s = -cr; c = sr;
  }
  *s_out = s;
  *c_out = c;
}

LL_HD void ll_sincosf(float x, float* s, float* c) {
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}
LL_HD float ll_sinf(float x) {
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
  return (float)sd;
}
LL_HD float ll_cosf(float x) {
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
  return (float)cd;
}
LL_HD float ll_tanf(float x) {
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
  return (float)(sd / cd);
}

#endif /* LL_PORTABLE_MATH_H */
