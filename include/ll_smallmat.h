/*
 * ll_smallmat.h -- fixed-size float linear algebra used by the LM stages, shared by
 * the CUDA kernels and the CPU oracle, and validated on its own against numpy
 * (tests/test_smallmat.py), because the reference takes these from Eigen, which is
 * NOT vendored in /root/reference ("parity unpinned"):
 *
 *   colPivHouseholderQr().solve   featureAssociation.cpp:866,983  mapOptmization.cpp:1153,1260
 *   SelfAdjointEigenSolver        featureAssociation.cpp:874,990  mapOptmization.cpp:1077,1267
 *   .inverse()                    featureAssociation.cpp:891,1007 mapOptmization.cpp:1285
 *
 * The algorithms restate Eigen 3.3's published ones (column-pivoted Householder QR
 * with norm down-dating; scale + Householder tridiagonalisation + implicit
 * Wilkinson-shift symmetric QR + ascending selection sort of eigenvalues with column
 * swaps).  Sums run strictly left to right.  Matrices are row-major: A[r*C + c].
 */
#ifndef LL_SMALLMAT_H
#define LL_SMALLMAT_H

#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define LLM_HD __host__ __device__ inline
/* The three solvers are kept out of line in device code: nvcc 12.9's optimiser (cicc -O3) was
 * observed to miscompile colpiv_qr_solve<6,6> when it is inlined into a kernel that builds the
 * matrix from shared-memory doubles (tools/test_smallmat2.cu reproduces it: wrong X at -O3,
 * right X with -Xcicc -O1, with -G, or out of line).  Out of line the body is compiled once,
 * in one context, which is the context tests/test_gpu_mapping.py checks against the CPU. */
#define LLM_SOLVER __host__ __device__ __noinline__
#else
#define LLM_HD inline
#define LLM_SOLVER inline
#endif

namespace llm {

LLM_HD float fabs_(float v) { return v < 0.f ? -v : v; }
LLM_HD float sqrt_(float v) { return sqrtf(v); }

/* Householder reflector of x[0..n) (stride st): on exit x[0] would become beta,
 * essential part (x[1..n) / (x0 - beta)) is written back in place of x[1..n). */
LLM_HD void make_householder(float* x, int n, int st, float* tau, float* beta) {
  float tail_sq = 0.f;
  for (int i = 1; i < n; ++i) tail_sq += x[i * st] * x[i * st];
  const float c0 = x[0];
  if (n == 1 || tail_sq <= FLT_MIN) {
    *tau = 0.f;
    *beta = c0;
    for (int i = 1; i < n; ++i) x[i * st] = 0.f;
  } else {
    float b = sqrt_(c0 * c0 + tail_sq);
    if (c0 >= 0.f) b = -b;
    const float d = c0 - b;
    for (int i = 1; i < n; ++i) x[i * st] = x[i * st] / d;
    *tau = (b - c0) / b;
    *beta = b;
  }
}

/* Least-squares / square solve of A(RxC) x = b by column-pivoted Householder QR
 * (Eigen ColPivHouseholderQR::compute + solve).  A is destroyed. */
template <int R, int C>
LLM_SOLVER void colpiv_qr_solve(float* A, const float* b_in, float* x) {
  const int size = R < C ? R : C;
  const float eps = FLT_EPSILON;
  float h[C];
  float norm_upd[C], norm_dir[C];
  int perm[C];
  float maxnorm = 0.f;
  for (int k = 0; k < C; ++k) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += A[r * C + k] * A[r * C + k];
    norm_dir[k] = norm_upd[k] = sqrt_(s);
    if (norm_upd[k] > maxnorm) maxnorm = norm_upd[k];
    perm[k] = k;
  }
  const float thr_helper = (maxnorm * eps) * (maxnorm * eps) / (float)R;
  const float downdate_thr = sqrt_(eps);
  int nonzero_pivots = size;
  float maxpivot = 0.f;
  for (int k = 0; k < size; ++k) {
    int big = k;
    float bigv = norm_upd[k];
    for (int j = k + 1; j < C; ++j)
      if (norm_upd[j] > bigv) { bigv = norm_upd[j]; big = j; }
    const float big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < thr_helper * (float)(R - k)) nonzero_pivots = k;
    if (big != k) {
      for (int r = 0; r < R; ++r) { float t = A[r * C + k]; A[r * C + k] = A[r * C + big]; A[r * C + big] = t; }
      float t = norm_upd[k]; norm_upd[k] = norm_upd[big]; norm_upd[big] = t;
      t = norm_dir[k]; norm_dir[k] = norm_dir[big]; norm_dir[big] = t;
      int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
    }
    float beta;
    make_householder(&A[k * C + k], R - k, C, &h[k], &beta);
    A[k * C + k] = beta;
    if (fabs_(beta) > maxpivot) maxpivot = fabs_(beta);
    /* apply H_k to the trailing columns */
    if (R - k == 1) {
      for (int j = k + 1; j < C; ++j) A[k * C + j] *= (1.f - h[k]);
    } else if (h[k] != 0.f) {
      for (int j = k + 1; j < C; ++j) {
        float t = 0.f;
        for (int r = k + 1; r < R; ++r) t += A[r * C + k] * A[r * C + j];
        t += A[k * C + j];
        A[k * C + j] -= h[k] * t;
        for (int r = k + 1; r < R; ++r) A[r * C + j] -= h[k] * A[r * C + k] * t;
      }
    }
    for (int j = k + 1; j < C; ++j) {
      if (norm_upd[j] != 0.f) {
        float t = fabs_(A[k * C + j]) / norm_upd[j];
        t = (1.f + t) * (1.f - t);
        if (t < 0.f) t = 0.f;
        const float q = norm_upd[j] / norm_dir[j];
        const float t2 = t * (q * q);
        if (t2 <= downdate_thr) {
          float s = 0.f;
          for (int r = k + 1; r < R; ++r) s += A[r * C + j] * A[r * C + j];
          norm_dir[j] = sqrt_(s);
          norm_upd[j] = norm_dir[j];
        } else {
          norm_upd[j] *= sqrt_(t);
        }
      }
    }
  }
  /* rank with Eigen's default threshold eps * diagonalSize */
  const float premul = fabs_(maxpivot) * (eps * (float)size);
  int rank = 0;
  for (int i = 0; i < nonzero_pivots; ++i) rank += (fabs_(A[i * C + i]) > premul) ? 1 : 0;
  float c[R];
  for (int r = 0; r < R; ++r) c[r] = b_in[r];
  /* c = Q^T b, Q = H_0 H_1 ... (only the first nonzero_pivots reflectors) */
  for (int k = 0; k < nonzero_pivots; ++k) {
    if (R - k == 1) {
      c[k] *= (1.f - h[k]);
    } else if (h[k] != 0.f) {
      float t = 0.f;
      for (int r = k + 1; r < R; ++r) t += A[r * C + k] * c[r];
      t += c[k];
      c[k] -= h[k] * t;
      for (int r = k + 1; r < R; ++r) c[r] -= h[k] * A[r * C + k] * t;
    }
  }
  /* back substitution on the leading rank x rank upper triangle */
  for (int i = rank - 1; i >= 0; --i) {
    float s = c[i];
    for (int j = i + 1; j < rank; ++j) s -= A[i * C + j] * c[j];
    c[i] = s / A[i * C + i];
  }
  for (int i = 0; i < C; ++i) x[i] = 0.f;
  for (int i = 0; i < rank; ++i) x[perm[i]] = c[i];
}

/* colpiv_qr_solve<3, 3> with every array index a compile-time constant (column swaps, the pivot count, the rank and
 * the final permutation are resolved with explicit comparisons), so that on the device everything stays in registers:
 * the generic routine above is out of line and indexes its arrays dynamically, which costs the single thread that
 * solves the 3x3 normal equations of a scan-to-scan LM iteration several microseconds.  Same operations in the same
 * order as the generic code -- tests/csrc/check_qr3.cpp compares the two bit for bit on random and degenerate systems. */
#define LLM_SWAPF(a, b) do { const float t__ = (a); (a) = (b); (b) = t__; } while (0)
LLM_HD void colpiv_qr_solve3(const float* A_in, const float* b_in, float* x) {
  const float eps = FLT_EPSILON;
  float a00 = A_in[0], a01 = A_in[1], a02 = A_in[2], a10 = A_in[3], a11 = A_in[4], a12 = A_in[5], a20 = A_in[6], a21 = A_in[7],
        a22 = A_in[8];
  float nu0, nu1, nu2, nd0, nd1, nd2;
  int p0 = 0, p1 = 1, p2 = 2;
  float maxnorm = 0.f;
  { float s = 0.f; s += a00 * a00; s += a10 * a10; s += a20 * a20; nd0 = nu0 = sqrt_(s); if (nu0 > maxnorm) maxnorm = nu0; }
  { float s = 0.f; s += a01 * a01; s += a11 * a11; s += a21 * a21; nd1 = nu1 = sqrt_(s); if (nu1 > maxnorm) maxnorm = nu1; }
  { float s = 0.f; s += a02 * a02; s += a12 * a12; s += a22 * a22; nd2 = nu2 = sqrt_(s); if (nu2 > maxnorm) maxnorm = nu2; }
  const float thr_helper = (maxnorm * eps) * (maxnorm * eps) / 3.f;
  const float downdate_thr = sqrt_(eps);
  int nonzero_pivots = 3;
  float maxpivot = 0.f;
  float h0, h1, h2;
  /* ---- k = 0 ---- */
  {
    int big = 0;
    float bigv = nu0;
    if (nu1 > bigv) { bigv = nu1; big = 1; }
    if (nu2 > bigv) { bigv = nu2; big = 2; }
    const float big_sq = bigv * bigv;
    if (nonzero_pivots == 3 && big_sq < thr_helper * 3.f) nonzero_pivots = 0;
    if (big == 1) {
      LLM_SWAPF(a00, a01); LLM_SWAPF(a10, a11); LLM_SWAPF(a20, a21); LLM_SWAPF(nu0, nu1); LLM_SWAPF(nd0, nd1);
      const int t = p0; p0 = p1; p1 = t;
    } else if (big == 2) {
      LLM_SWAPF(a00, a02); LLM_SWAPF(a10, a12); LLM_SWAPF(a20, a22); LLM_SWAPF(nu0, nu2); LLM_SWAPF(nd0, nd2);
      const int t = p0; p0 = p2; p2 = t;
    }
    /* make_householder on (a00, a10, a20) */
    float tail_sq = 0.f;
    tail_sq += a10 * a10;
    tail_sq += a20 * a20;
    const float c0 = a00;
    float beta;
    if (tail_sq <= FLT_MIN) {
      h0 = 0.f; beta = c0; a10 = 0.f; a20 = 0.f;
    } else {
      float b = sqrt_(c0 * c0 + tail_sq);
      if (c0 >= 0.f) b = -b;
      const float d = c0 - b;
      a10 = a10 / d; a20 = a20 / d;
      h0 = (b - c0) / b;
      beta = b;
    }
    a00 = beta;
    if (fabs_(beta) > maxpivot) maxpivot = fabs_(beta);
    if (h0 != 0.f) {
      { float t = 0.f; t += a10 * a11; t += a20 * a21; t += a01; a01 -= h0 * t; a11 -= h0 * a10 * t; a21 -= h0 * a20 * t; }
      { float t = 0.f; t += a10 * a12; t += a20 * a22; t += a02; a02 -= h0 * t; a12 -= h0 * a10 * t; a22 -= h0 * a20 * t; }
    }
    if (nu1 != 0.f) {
      float t = fabs_(a01) / nu1;
      t = (1.f + t) * (1.f - t);
      if (t < 0.f) t = 0.f;
      const float q = nu1 / nd1;
      const float t2 = t * (q * q);
      if (t2 <= downdate_thr) { float s = 0.f; s += a11 * a11; s += a21 * a21; nd1 = sqrt_(s); nu1 = nd1; } else { nu1 *= sqrt_(t); }
    }
    if (nu2 != 0.f) {
      float t = fabs_(a02) / nu2;
      t = (1.f + t) * (1.f - t);
      if (t < 0.f) t = 0.f;
      const float q = nu2 / nd2;
      const float t2 = t * (q * q);
      if (t2 <= downdate_thr) { float s = 0.f; s += a12 * a12; s += a22 * a22; nd2 = sqrt_(s); nu2 = nd2; } else { nu2 *= sqrt_(t); }
    }
  }
  /* ---- k = 1 ---- */
  {
    int big = 1;
    float bigv = nu1;
    if (nu2 > bigv) { bigv = nu2; big = 2; }
    const float big_sq = bigv * bigv;
    if (nonzero_pivots == 3 && big_sq < thr_helper * 2.f) nonzero_pivots = 1;
    if (big == 2) {
      LLM_SWAPF(a01, a02); LLM_SWAPF(a11, a12); LLM_SWAPF(a21, a22); LLM_SWAPF(nu1, nu2); LLM_SWAPF(nd1, nd2);
      const int t = p1; p1 = p2; p2 = t;
    }
    /* make_householder on (a11, a21) */
    float tail_sq = 0.f;
    tail_sq += a21 * a21;
    const float c0 = a11;
    float beta;
    if (tail_sq <= FLT_MIN) {
      h1 = 0.f; beta = c0; a21 = 0.f;
    } else {
      float b = sqrt_(c0 * c0 + tail_sq);
      if (c0 >= 0.f) b = -b;
      const float d = c0 - b;
      a21 = a21 / d;
      h1 = (b - c0) / b;
      beta = b;
    }
    a11 = beta;
    if (fabs_(beta) > maxpivot) maxpivot = fabs_(beta);
    if (h1 != 0.f) {
      float t = 0.f; t += a21 * a22; t += a12; a12 -= h1 * t; a22 -= h1 * a21 * t;
    }
    if (nu2 != 0.f) {
      float t = fabs_(a12) / nu2;
      t = (1.f + t) * (1.f - t);
      if (t < 0.f) t = 0.f;
      const float q = nu2 / nd2;
      const float t2 = t * (q * q);
      if (t2 <= downdate_thr) { float s = 0.f; s += a22 * a22; nd2 = sqrt_(s); nu2 = nd2; } else { nu2 *= sqrt_(t); }
    }
  }
  /* ---- k = 2 ---- */
  {
    const float bigv = nu2;
    const float big_sq = bigv * bigv;
    if (nonzero_pivots == 3 && big_sq < thr_helper * 1.f) nonzero_pivots = 2;
    /* make_householder with n == 1 */
    h2 = 0.f;
    const float beta = a22;
    a22 = beta;
    if (fabs_(beta) > maxpivot) maxpivot = fabs_(beta);
  }
  const float premul = fabs_(maxpivot) * (eps * 3.f);
  int rank = 0;
  if (0 < nonzero_pivots) rank += (fabs_(a00) > premul) ? 1 : 0;
  if (1 < nonzero_pivots) rank += (fabs_(a11) > premul) ? 1 : 0;
  if (2 < nonzero_pivots) rank += (fabs_(a22) > premul) ? 1 : 0;
  float c0 = b_in[0], c1 = b_in[1], c2 = b_in[2];
  if (0 < nonzero_pivots && h0 != 0.f) {
    float t = 0.f; t += a10 * c1; t += a20 * c2; t += c0; c0 -= h0 * t; c1 -= h0 * a10 * t; c2 -= h0 * a20 * t;
  }
  if (1 < nonzero_pivots && h1 != 0.f) {
    float t = 0.f; t += a21 * c2; t += c1; c1 -= h1 * t; c2 -= h1 * a21 * t;
  }
  if (2 < nonzero_pivots) c2 *= (1.f - h2);
  /* back substitution on the leading rank x rank upper triangle */
  if (rank > 2) { const float s = c2; c2 = s / a22; }
  if (rank > 1) { float s = c1; if (rank > 2) s -= a12 * c2; c1 = s / a11; }
  if (rank > 0) { float s = c0; if (rank > 1) s -= a01 * c1; if (rank > 2) s -= a02 * c2; c0 = s / a00; }
  float x0 = 0.f, x1 = 0.f, x2 = 0.f;
  if (rank > 0) { if (p0 == 0) x0 = c0; else if (p0 == 1) x1 = c0; else x2 = c0; }
  if (rank > 1) { if (p1 == 0) x0 = c1; else if (p1 == 1) x1 = c1; else x2 = c1; }
  if (rank > 2) { if (p2 == 0) x0 = c2; else if (p2 == 1) x1 = c2; else x2 = c2; }
  x[0] = x0; x[1] = x1; x[2] = x2;
}
#undef LLM_SWAPF

LLM_HD void make_givens(float p, float q, float* c, float* s) {
  if (q == 0.f) {
    *c = p < 0.f ? -1.f : 1.f;
    *s = 0.f;
  } else if (p == 0.f) {
    *c = 0.f;
    *s = q < 0.f ? 1.f : -1.f;
  } else if (fabs_(p) > fabs_(q)) {
    const float t = q / p;
    float u = sqrt_(1.f + t * t);
    if (p < 0.f) u = -u;
    *c = 1.f / u;
    *s = -t * (*c);
  } else {
    const float t = p / q;
    float u = sqrt_(1.f + t * t);
    if (q < 0.f) u = -u;
    *s = -1.f / u;
    *c = -t * (*s);
  }
}

/* Symmetric eigen-decomposition (Eigen SelfAdjointEigenSolver::compute).  Reads the
 * lower triangle of M (row-major NxN).  evals ascending; eigenvector k is COLUMN k
 * of V, i.e. V[r*N + k]. */
template <int N>
LLM_SOLVER void self_adjoint_eigen(const float* M, float* evals, float* V) {
  float a[N * N];
  float scale = 0.f;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c <= r; ++c) {
      const float v = fabs_(M[r * N + c]);
      if (v > scale) scale = v;
    }
  if (scale == 0.f) scale = 1.f;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) a[r * N + c] = (c <= r ? M[r * N + c] : M[c * N + r]) / scale;
  float diag[N], sub[N > 1 ? N - 1 : 1];
  if (N == 3) {
    /* closed-form 3x3 tridiagonalisation */
    diag[0] = a[0];
    const float v1norm2 = a[2 * N + 0] * a[2 * N + 0];
    if (v1norm2 <= FLT_MIN) {
      diag[1] = a[1 * N + 1];
      diag[2] = a[2 * N + 2];
      sub[0] = a[1 * N + 0];
      sub[1] = a[2 * N + 1];
      for (int i = 0; i < 9; ++i) V[i] = 0.f;
      V[0] = V[4] = V[8] = 1.f;
    } else {
      const float beta = sqrt_(a[1 * N + 0] * a[1 * N + 0] + v1norm2);
      const float invb = 1.f / beta;
      const float m01 = a[1 * N + 0] * invb;
      const float m02 = a[2 * N + 0] * invb;
      const float q = 2.f * m01 * a[2 * N + 1] + m02 * (a[2 * N + 2] - a[1 * N + 1]);
      diag[1] = a[1 * N + 1] + m02 * q;
      diag[2] = a[2 * N + 2] - m02 * q;
      sub[0] = beta;
      sub[1] = a[2 * N + 1] - m01 * q;
      V[0] = 1.f; V[1] = 0.f; V[2] = 0.f;
      V[3] = 0.f; V[4] = m01; V[5] = m02;
      V[6] = 0.f; V[7] = m02; V[8] = -m01;
    }
  } else {
    /* Householder tridiagonalisation on the lower triangle, then Q = H_0 ... H_{N-2} */
    float hc[N];
    for (int i = 0; i < N - 1; ++i) {
      const int rem = N - i - 1;
      float tau, beta;
      make_householder(&a[(i + 1) * N + i], rem, N, &tau, &beta);
      a[(i + 1) * N + i] = 1.f;
      float p[N];
      for (int r = 0; r < rem; ++r) {
        float s = 0.f;
        for (int c = 0; c < rem; ++c) {
          const int rr = i + 1 + r, cc = i + 1 + c;
          const float m = (cc <= rr) ? a[rr * N + cc] : a[cc * N + rr];
          s += m * (tau * a[cc * N + i]);
        }
        p[r] = s;
      }
      float dot = 0.f;
      for (int r = 0; r < rem; ++r) dot += p[r] * a[(i + 1 + r) * N + i];
      const float alpha = tau * (-0.5f) * dot;
      for (int r = 0; r < rem; ++r) p[r] += alpha * a[(i + 1 + r) * N + i];
      for (int r = 0; r < rem; ++r)
        for (int c = 0; c <= r; ++c) {
          const int rr = i + 1 + r, cc = i + 1 + c;
          a[rr * N + cc] -= a[rr * N + i] * p[c] + p[r] * a[cc * N + i];
        }
      a[(i + 1) * N + i] = beta;
      hc[i] = tau;
    }
    for (int i = 0; i < N; ++i) diag[i] = a[i * N + i];
    for (int i = 0; i < N - 1; ++i) sub[i] = a[(i + 1) * N + i];
    for (int i = 0; i < N * N; ++i) V[i] = 0.f;
    for (int i = 0; i < N; ++i) V[i * N + i] = 1.f;
    for (int k = N - 2; k >= 0; --k) {
      /* V <- H_k V with v = [0.. , 1 at k+1, essential below] */
      const float tau = hc[k];
      if (tau == 0.f) continue;
      for (int c = 0; c < N; ++c) {
        float t = V[(k + 1) * N + c];
        for (int r = k + 2; r < N; ++r) t += a[r * N + k] * V[r * N + c];
        V[(k + 1) * N + c] -= tau * t;
        for (int r = k + 2; r < N; ++r) V[r * N + c] -= tau * a[r * N + k] * t;
      }
    }
  }
  /* implicit symmetric QR on the tridiagonal */
  int end = N - 1, start = 0, iter = 0;
  const float tiny = FLT_MIN;
  const float prec = 2.f * FLT_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (fabs_(sub[i]) <= (fabs_(diag[i]) + fabs_(diag[i + 1])) * prec || fabs_(sub[i]) <= tiny) sub[i] = 0.f;
    while (end > 0 && sub[end - 1] == 0.f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * N) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.f) start--;
    /* one QR step on [start, end] */
    const float td = (diag[end - 1] - diag[end]) * 0.5f;
    const float e = sub[end - 1];
    float mu = diag[end];
    if (td == 0.f) {
      mu -= fabs_(e);
    } else {
      const float e2 = e * e;
      const float hyp = sqrt_(td * td + e * e);
      if (e2 == 0.f)
        mu -= (e / (td + (td > 0.f ? 1.f : -1.f))) * (e / hyp);
      else
        mu -= e2 / (td + (td > 0.f ? hyp : -hyp));
    }
    float x = diag[start] - mu;
    float z = sub[start];
    for (int k = start; k < end; ++k) {
      float gc, gs;
      make_givens(x, z, &gc, &gs);
      const float sdk = gs * diag[k] + gc * sub[k];
      const float dkp1 = gs * sub[k] + gc * diag[k + 1];
      diag[k] = gc * (gc * diag[k] - gs * sub[k]) - gs * (gc * sub[k] - gs * diag[k + 1]);
      diag[k + 1] = gs * sdk + gc * dkp1;
      sub[k] = gc * sdk - gs * dkp1;
      if (k > start) sub[k - 1] = gc * sub[k - 1] - gs * z;
      x = sub[k];
      if (k < end - 1) {
        z = -gs * sub[k + 1];
        sub[k + 1] = gc * sub[k + 1];
      }
      for (int r = 0; r < N; ++r) {
        const float xi = V[r * N + k], yi = V[r * N + k + 1];
        V[r * N + k] = gc * xi - gs * yi;
        V[r * N + k + 1] = gs * xi + gc * yi;
      }
    }
  }
  for (int i = 0; i < N - 1; ++i) {
    int k = i;
    for (int j = i + 1; j < N; ++j)
      if (diag[j] < diag[k]) k = j;
    if (k != i) {
      float t = diag[i]; diag[i] = diag[k]; diag[k] = t;
      for (int r = 0; r < N; ++r) { t = V[r * N + i]; V[r * N + i] = V[r * N + k]; V[r * N + k] = t; }
    }
  }
  for (int i = 0; i < N; ++i) evals[i] = diag[i] * scale;
}

/* Inverse by Gauss-Jordan with partial pivoting (only reached when an LM stage is
 * degenerate: matP = V^-1 * V2, featureAssociation.cpp:891). Returns false if singular. */
template <int N>
LLM_SOLVER bool invert(const float* M, float* Inv) {
  float a[N * N];
  for (int i = 0; i < N * N; ++i) { a[i] = M[i]; Inv[i] = 0.f; }
  for (int i = 0; i < N; ++i) Inv[i * N + i] = 1.f;
  for (int c = 0; c < N; ++c) {
    int p = c;
    for (int r = c + 1; r < N; ++r)
      if (fabs_(a[r * N + c]) > fabs_(a[p * N + c])) p = r;
    if (a[p * N + c] == 0.f) return false;
    if (p != c)
      for (int k = 0; k < N; ++k) {
        float t = a[c * N + k]; a[c * N + k] = a[p * N + k]; a[p * N + k] = t;
        t = Inv[c * N + k]; Inv[c * N + k] = Inv[p * N + k]; Inv[p * N + k] = t;
      }
    const float d = a[c * N + c];
    for (int k = 0; k < N; ++k) { a[c * N + k] /= d; Inv[c * N + k] /= d; }
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const float f = a[r * N + c];
      if (f == 0.f) continue;
      for (int k = 0; k < N; ++k) { a[r * N + k] -= f * a[c * N + k]; Inv[r * N + k] -= f * Inv[c * N + k]; }
    }
  }
  return true;
}

/* The degeneracy handling common to featureAssociation.cpp:869-898,985-1014 and
 * mapOptmization.cpp:1262-1292: at iteration 0 find eigenvalues below `thr` scanning
 * from the largest down, zero the matching ROWS of V2 (sic), P = V^-1 * V2. */
template <int N>
LLM_SOLVER bool degeneracy_projector(const float* AtA, float thr, float* P) {
  float ev[N], V[N * N], V2[N * N], Vi[N * N];
  self_adjoint_eigen<N>(AtA, ev, V);
  for (int i = 0; i < N * N; ++i) V2[i] = V[i];
  bool degenerate = false;
  for (int i = N - 1; i >= 0; --i) {
    if (ev[i] < thr) {
      for (int j = 0; j < N; ++j) V2[i * N + j] = 0.f;
      degenerate = true;
    } else {
      break;
    }
  }
  if (!invert<N>(V, Vi)) {
    for (int i = 0; i < N * N; ++i) P[i] = 0.f;
    for (int i = 0; i < N; ++i) P[i * N + i] = 1.f;
    return degenerate;
  }
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      float s = 0.f;
      for (int k = 0; k < N; ++k) s += Vi[r * N + k] * V2[k * N + c];
      P[r * N + c] = s;
    }
  return degenerate;
}

/* Certified skip of the degeneracy test: the largest eigenvalue of a symmetric matrix is at least trace / N, and the
 * float symmetric-QR eigenvalues are within ~1e-5 relative of the true ones, so trace / N >= 2 * thr proves that
 * the largest COMPUTED eigenvalue is >= thr, i.e. that degeneracy_projector would return false (its P is then never
 * read: featureAssociation.cpp:893-899, mapOptmization.cpp:1287-1292 only use matP when isDegenerate).  NaN / Inf
 * fail the comparison and take the full path. */
template <int N>
LLM_HD bool certainly_not_degenerate(const float* AtA, float thr) {
  float tr = 0.f;
  for (int i = 0; i < N; ++i) tr += AtA[i * N + i];
  return tr >= 2.f * thr * (float)N && tr < FLT_MAX;
}

}  // namespace llm

#endif /* LL_SMALLMAT_H */
