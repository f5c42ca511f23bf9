// CPU check: llm::colpiv_qr_solve3 (register-only specialisation) against the generic llm::colpiv_qr_solve<3, 3>, bit for
// bit, on random symmetric positive (normal-equation like), general, rank-deficient, tiny and tied-norm systems.
// Prints "cases mismatches".
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ll_smallmat.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t next() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static float uni(float lo, float hi) { return lo + (hi - lo) * (float)((next() >> 11) * (1.0 / 9007199254740992.0)); }

int main() {
  long cases = 0, bad = 0;
  for (long it = 0; it < 3000000; ++it) {
    float A[9], b[3];
    const int kind = (int)(next() % 8);
    if (kind < 3) {  // J^T J of a few random rows, the shape the LM feeds
      float J[12][3];
      const int rows = 3 + (int)(next() % 10);
      const float sc = kind == 0 ? 1.f : (kind == 1 ? 1e3f : 1e-3f);
      for (int r = 0; r < rows; ++r) for (int c = 0; c < 3; ++c) J[r][c] = uni(-1.f, 1.f) * sc * (c == 2 && kind == 2 ? 1e-4f : 1.f);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { float s = 0.f; for (int r = 0; r < rows; ++r) s += J[r][i] * J[r][j]; A[i * 3 + j] = s; }
    } else if (kind == 3) {  // general
      for (int i = 0; i < 9; ++i) A[i] = uni(-10.f, 10.f);
    } else if (kind == 4) {  // rank 1 or 2
      float u[3] = {uni(-1, 1), uni(-1, 1), uni(-1, 1)}, v[3] = {uni(-1, 1), uni(-1, 1), uni(-1, 1)};
      const bool two = next() & 1;
      float u2[3] = {uni(-1, 1), uni(-1, 1), uni(-1, 1)}, v2[3] = {uni(-1, 1), uni(-1, 1), uni(-1, 1)};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i * 3 + j] = u[i] * v[j] + (two ? u2[i] * v2[j] : 0.f);
    } else if (kind == 5) {  // zero / duplicated / tied-norm columns
      for (int i = 0; i < 9; ++i) A[i] = uni(-1.f, 1.f);
      const int c = (int)(next() % 3), c2 = (int)(next() % 3);
      const int mode = (int)(next() % 3);
      for (int r = 0; r < 3; ++r) A[r * 3 + c] = mode == 0 ? 0.f : (mode == 1 ? A[r * 3 + c2] : -A[r * 3 + c2]);
    } else if (kind == 6) {  // tiny and huge magnitudes
      const float sc = (next() & 1) ? 1e-20f : 1e15f;
      for (int i = 0; i < 9; ++i) A[i] = uni(-1.f, 1.f) * sc;
    } else {  // diagonal / triangular / all zero
      for (int i = 0; i < 9; ++i) A[i] = 0.f;
      const int mode = (int)(next() % 3);
      if (mode >= 1) for (int i = 0; i < 3; ++i) A[i * 3 + i] = uni(-2.f, 2.f);
      if (mode == 2) { A[1] = uni(-1, 1); A[2] = uni(-1, 1); A[5] = uni(-1, 1); }
    }
    for (int i = 0; i < 3; ++i) b[i] = uni(-5.f, 5.f);
    float Ag[9], xg[3], xs[3];
    std::memcpy(Ag, A, sizeof(A));
    llm::colpiv_qr_solve<3, 3>(Ag, b, xg);
    llm::colpiv_qr_solve3(A, b, xs);
    ++cases;
    if (std::memcmp(xg, xs, sizeof(xg)) != 0) {
      if (bad < 5) std::fprintf(stderr, "mismatch kind %d: generic %.9g %.9g %.9g  special %.9g %.9g %.9g\n", kind, xg[0], xg[1], xg[2], xs[0], xs[1], xs[2]);
      ++bad;
    }
  }
  std::printf("%ld %ld\n", cases, bad);
  return bad ? 1 : 0;
}
