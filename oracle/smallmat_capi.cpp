// TEST INFRASTRUCTURE: C wrappers around include/ll_smallmat.h so that tests/test_smallmat.py can
// check the shared small-matrix solvers against numpy (they restate un-vendored Eigen algorithms).
#include "../include/ll_smallmat.h"
extern "C" {
void sm_qr_solve_3x3(const float* A, const float* b, float* x) { float a[9]; for (int i = 0; i < 9; ++i) a[i] = A[i]; llm::colpiv_qr_solve<3, 3>(a, b, x); }
void sm_qr_solve_6x6(const float* A, const float* b, float* x) { float a[36]; for (int i = 0; i < 36; ++i) a[i] = A[i]; llm::colpiv_qr_solve<6, 6>(a, b, x); }
void sm_qr_solve_5x3(const float* A, const float* b, float* x) { float a[15]; for (int i = 0; i < 15; ++i) a[i] = A[i]; llm::colpiv_qr_solve<5, 3>(a, b, x); }
void sm_eigen_3(const float* M, float* ev, float* V) { llm::self_adjoint_eigen<3>(M, ev, V); }
void sm_eigen_6(const float* M, float* ev, float* V) { llm::self_adjoint_eigen<6>(M, ev, V); }
int sm_invert_6(const float* M, float* Inv) { return llm::invert<6>(M, Inv) ? 1 : 0; }
int sm_degeneracy_3(const float* AtA, float thr, float* P) { return llm::degeneracy_projector<3>(AtA, thr, P) ? 1 : 0; }
int sm_degeneracy_6(const float* AtA, float thr, float* P) { return llm::degeneracy_projector<6>(AtA, thr, P) ? 1 : 0; }
}
