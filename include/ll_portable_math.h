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

#if defined(__CUDACC__)
/* out-of-line copy of the full routine for the (rare) device fall-back, so that the short cut below does not drag the
 * whole reduction + two long polynomials into every call site */
static __device__ __noinline__ void ll_sincosd_outofline(double x, double* s_out, double* c_out) { ll_sincosd(x, s_out, c_out); }
#endif
#if defined(__CUDA_ARCH__)
/* Device-only certified short cut for |x| < 0.25 (every per-point angle of TransformToStart / TransformToEnd and most
 * poses): sin and cos from short fused-multiply-add Taylor chains in double.  For such x the reduction above gives k = 0
 * and r = x exactly, and both this chain and ll_sincosd are within 1.5e-16 relative of the true value (truncation
 * < 1e-17; the leading term is exact, the correction term is < 0.032 of it), so the two doubles differ by less than
 * 3e-16 relative.  The float results can only differ if a float rounding boundary lies inside that gap: the value is
 * accepted only when rounding the double moved by -/+ 1e-15 relative gives the same float, otherwise the caller falls
 * back to ll_sincosd.  The result is therefore bit-identical to the portable definition by construction. */
__device__ __forceinline__ bool ll_sincosf_small(float x, float* s, float* c) {
  if (!(fabsf(x) < 0.25f)) return false;
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
  if ((float)(ds - es) != sf || (float)(ds + es) != sf || (float)(dc - ec) != cf || (float)(dc + ec) != cf) return false;
  *s = sf;
  *c = cf;
  return true;
}
#endif

LL_HD void ll_sincosf(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  if (ll_sincosf_small(x, s, c)) return;
  double sd, cd;
  ll_sincosd_outofline((double)x, &sd, &cd);
#else
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
#endif
  *s = (float)sd;
  *c = (float)cd;
}
LL_HD float ll_sinf(float x) {
#if defined(__CUDA_ARCH__)
  { float sf, cf; if (ll_sincosf_small(x, &sf, &cf)) return sf; }
  double sd, cd;
  ll_sincosd_outofline((double)x, &sd, &cd);
#else
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
#endif
  return (float)sd;
}
LL_HD float ll_cosf(float x) {
#if defined(__CUDA_ARCH__)
  { float sf, cf; if (ll_sincosf_small(x, &sf, &cf)) return cf; }
  double sd, cd;
  ll_sincosd_outofline((double)x, &sd, &cd);
#else
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
#endif
  return (float)cd;
}
LL_HD float ll_tanf(float x) {
  double sd, cd;
  ll_sincosd((double)x, &sd, &cd);
  return (float)(sd / cd);
}

#endif /* LL_PORTABLE_MATH_H */
