/*
 * oracle_math.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Float math backend of the oracle.  Two run-time selectable backends:
 *   portable (default): include/ll_portable_math.h, bit-identical to the CUDA kernels;
 *   libm: glibc float overloads, i.e. what the reference gets from std::asin,
 *         std::atan2 and (with <math.h> visible, g++ >= 6) unqualified sin/cos/tan/
 *         atan2/asin on float arguments (SURVEY.md section 7 hard part 1, section 10).
 * tests/test_oracle_pins.py checks that both backends give identical discrete outputs
 * on the fixtures, which is what ties the portable backend to the reference's libm.
 */
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <cmath>

#include "../include/ll_portable_math.h"

namespace om {

extern int g_use_libm;

inline float asin_(float v) { return g_use_libm ? ::asinf(v) : ll_asinf(v); }
inline float atan2_(float y, float x) { return g_use_libm ? ::atan2f(y, x) : ll_atan2f(y, x); }
inline float sin_(float v) { return g_use_libm ? ::sinf(v) : ll_sinf(v); }
inline float cos_(float v) { return g_use_libm ? ::cosf(v) : ll_cosf(v); }
inline float tan_(float v) { return g_use_libm ? ::tanf(v) : ll_tanf(v); }
inline float sqrt_(float v) { return ::sqrtf(v); } /* IEEE exact on both sides */
inline float fabs_(float v) { return ::fabsf(v); }

}  // namespace om

#endif
