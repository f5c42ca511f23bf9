"""The shared fixed-size solvers of include/ll_smallmat.h against numpy (float64).  They restate Eigen
algorithms the reference calls (colPivHouseholderQr().solve, SelfAdjointEigenSolver, .inverse()) and
are used by BOTH the CUDA kernels and the CPU oracle, so they get an independent check here."""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "build", "libsmallmat_capi.so")


@pytest.fixture(scope="module")
def sm(built):
    lib = C.CDLL(LIB)
    for n in ("sm_qr_solve_3x3", "sm_qr_solve_6x6", "sm_qr_solve_5x3", "sm_eigen_3", "sm_eigen_6"):
        getattr(lib, n).argtypes = [C.c_void_p] * 3
        getattr(lib, n).restype = None
    lib.sm_invert_6.argtypes = [C.c_void_p, C.c_void_p]
    lib.sm_degeneracy_3.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    lib.sm_degeneracy_6.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    return lib


def _spd(rng, n, cond=1e3):
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ev = np.exp(rng.uniform(0, np.log(cond), n))
    return (q * ev) @ q.T


@pytest.mark.parametrize("n,fn", [(3, "sm_qr_solve_3x3"), (6, "sm_qr_solve_6x6")])
def test_square_solve_matches_numpy(sm, n, fn):
    rng = np.random.default_rng(n)
    for _ in range(200):
        A = _spd(rng, n).astype(np.float32)
        b = rng.normal(size=n).astype(np.float32)
        x = np.zeros(n, np.float32)
        getattr(sm, fn)(A.ctypes.data, b.ctypes.data, x.ctypes.data)
        ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        assert np.allclose(x, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max()), (x, ref)


def test_least_squares_5x3_matches_numpy(sm):
    rng = np.random.default_rng(7)
    for _ in range(200):
        # five points near a plane n.p + d = 0, as in surfOptimization (mapOptmization.cpp:1146-1153)
        nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
        P = rng.normal(size=(5, 3)) * 0.5
        P -= np.outer(P @ nrm, nrm)
        P += nrm * rng.uniform(2, 20) + rng.normal(size=(5, 3)) * 0.01
        A = P.astype(np.float32)
        b = -np.ones(5, np.float32)
        x = np.zeros(3, np.float32)
        sm.sm_qr_solve_5x3(A.ctypes.data, b.ctypes.data, x.ctypes.data)
        ref = np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=None)[0]
        assert np.allclose(x, ref, rtol=5e-3, atol=5e-4 * np.abs(ref).max()), (x, ref)


@pytest.mark.parametrize("n,fn", [(3, "sm_eigen_3"), (6, "sm_eigen_6")])
def test_symmetric_eigen_matches_numpy(sm, n, fn):
    rng = np.random.default_rng(10 + n)
    for _ in range(200):
        M = _spd(rng, n, 1e4).astype(np.float32)
        M = ((M + M.T) / 2).astype(np.float32)
        ev = np.zeros(n, np.float32)
        V = np.zeros((n, n), np.float32)
        getattr(sm, fn)(M.ctypes.data, ev.ctypes.data, V.ctypes.data)
        ref = np.linalg.eigvalsh(M.astype(np.float64))
        assert np.all(np.diff(ev) >= 0), "eigenvalues must be ascending (Eigen convention)"
        assert np.allclose(ev, ref, rtol=1e-4, atol=1e-4 * ref.max())
        # columns are unit eigenvectors: M V = V diag(ev), V^T V = I
        Vd = V.astype(np.float64)
        assert np.allclose(Vd.T @ Vd, np.eye(n), atol=1e-4)
        assert np.allclose(M.astype(np.float64) @ Vd, Vd * ev, atol=2e-4 * ref.max())


def test_eigen_3x3_diagonal_and_degenerate_inputs(sm):
    for M in (np.diag([3.0, 1.0, 2.0]), np.zeros((3, 3)), np.eye(3) * 5):
        M = M.astype(np.float32)
        ev = np.zeros(3, np.float32)
        V = np.zeros((3, 3), np.float32)
        sm.sm_eigen_3(M.ctypes.data, ev.ctypes.data, V.ctypes.data)
        assert np.allclose(np.sort(np.diag(M)), ev)
        assert np.allclose(np.abs(np.linalg.det(V.astype(np.float64))), 1.0, atol=1e-5)


def test_degeneracy_projector_semantics(sm):
    """featureAssociation.cpp:869-891 / mapOptmization.cpp:1262-1285: eigenvalues are tested from the
    largest down and the scan stops at the first one above the threshold, so a matrix is only flagged
    degenerate when its LARGEST eigenvalue is below the threshold; then P = V^-1 * V2 with rows of V2 zeroed."""
    rng = np.random.default_rng(3)
    P = np.zeros((3, 3), np.float32)
    well = _spd(rng, 3, 10).astype(np.float32) * 1000
    assert sm.sm_degeneracy_3(well.ctypes.data, C.c_float(10.0), P.ctypes.data) == 0
    assert np.allclose(P, np.eye(3), atol=1e-4)
    one_small = np.diag([1.0, 500.0, 900.0]).astype(np.float32)   # smallest < thr but largest above: not flagged
    assert sm.sm_degeneracy_3(one_small.ctypes.data, C.c_float(10.0), P.ctypes.data) == 0
    tiny = np.diag([1.0, 2.0, 3.0]).astype(np.float32)            # all below: flagged, everything projected out
    assert sm.sm_degeneracy_3(tiny.ctypes.data, C.c_float(10.0), P.ctypes.data) == 1
    assert np.allclose(P, 0, atol=1e-6)
    P6 = np.zeros((6, 6), np.float32)
    well6 = (_spd(rng, 6, 10) * 1000).astype(np.float32)
    assert sm.sm_degeneracy_6(well6.ctypes.data, C.c_float(100.0), P6.ctypes.data) == 0
    assert np.allclose(P6, np.eye(6), atol=1e-3)


def test_invert_6(sm):
    rng = np.random.default_rng(5)
    M = _spd(rng, 6, 100).astype(np.float32)
    Inv = np.zeros((6, 6), np.float32)
    assert sm.sm_invert_6(M.ctypes.data, Inv.ctypes.data) == 1
    assert np.allclose(M.astype(np.float64) @ Inv, np.eye(6), atol=1e-3)
