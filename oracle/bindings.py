"""ctypes bindings for the checkers -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It loads

* ``oracle/_build/libchefsi_oracle.so``  -- our plain-C restatement ("port"), built on demand
  with gcc (works on the GPU box too);
* ``oracle/_ref/libref_harness.so``      -- the UNMODIFIED reference routines (built in the
  dev container from /root/reference, travels to the GPU box as a prebuilt file).

Array convention: a block of ``ncol`` columns is a C-contiguous numpy array of shape
``(ncol, ld)`` (so column n starts at n*ld, exactly the reference's layout).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libchefsi_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_harness.so")

_dp = C.POINTER(C.c_double)


def build_port(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("chefsi_oracle.c", "chefsi_oracle.h", "chefsi_oracle_impl.inc")]
    stale = force or not os.path.exists(PORT_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(PORT_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return PORT_SO


def reference_available() -> bool:
    return os.path.exists(REF_SO)


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _kvec(k):
    return (C.c_double * 3)(*[float(v) for v in k])


class Port:
    """The C restatement (oracle/chefsi_oracle.c)."""

    def __init__(self):
        self.lib = C.CDLL(build_port())
        self.lib.oracle_random_value.restype = C.c_double
        self.lib.oracle_random_value.argtypes = [C.c_ulonglong, C.c_longlong, C.c_longlong]

    # -- helpers -----------------------------------------------------------------
    @staticmethod
    def _args(grid, proj):
        g = grid.to_c()
        nl = proj.to_c() if proj is not None else None
        return g, nl

    def lap_plus_diag(self, grid, a, b, c, v, x, kvec=None):
        g, _ = self._args(grid, None)
        x = np.ascontiguousarray(x)
        y = np.empty_like(x)
        ncol, ld = x.shape
        vp = _ptr(v) if v is not None else None
        if np.iscomplexobj(x):
            self.lib.oracle_lap_plus_diag_kpt(C.byref(g), _kvec(kvec), C.c_int(ncol), C.c_double(a),
                                              C.c_double(b), C.c_double(c), vp, _ptr(x),
                                              C.c_size_t(ld), _ptr(y), C.c_size_t(ld))
        else:
            self.lib.oracle_lap_plus_diag(C.byref(g), C.c_int(ncol), C.c_double(a), C.c_double(b),
                                          C.c_double(c), vp, _ptr(x), C.c_size_t(ld), _ptr(y),
                                          C.c_size_t(ld))
        return y

    def vnl_mult(self, grid, proj, x, Hx, kvec=None):
        g, nl = self._args(grid, proj)
        ncol, ld = x.shape
        if np.iscomplexobj(x):
            self.lib.oracle_vnl_mult_kpt(C.byref(g), C.byref(nl), _kvec(kvec), C.c_int(ncol), _ptr(x),
                                         C.c_size_t(ld), _ptr(Hx), C.c_size_t(ld))
        else:
            self.lib.oracle_vnl_mult(C.byref(g), C.byref(nl), C.c_int(ncol), _ptr(x), C.c_size_t(ld),
                                     _ptr(Hx), C.c_size_t(ld))
        return Hx

    def hamiltonian_mult(self, grid, proj, veff, c, x, kvec=None):
        g, nl = self._args(grid, proj)
        x = np.ascontiguousarray(x)
        Hx = np.empty_like(x)
        ncol, ld = x.shape
        nlp = C.byref(nl) if nl is not None else None
        if np.iscomplexobj(x):
            self.lib.oracle_hamiltonian_mult_kpt(C.byref(g), nlp, _ptr(veff), _kvec(kvec), C.c_int(ncol),
                                                 C.c_double(c), _ptr(x), C.c_size_t(ld), _ptr(Hx),
                                                 C.c_size_t(ld))
        else:
            self.lib.oracle_hamiltonian_mult(C.byref(g), nlp, _ptr(veff), C.c_int(ncol), C.c_double(c),
                                             _ptr(x), C.c_size_t(ld), _ptr(Hx), C.c_size_t(ld))
        return Hx

    def chebyshev_filter(self, grid, proj, veff, X, m, a, b, a0, kvec=None):
        """Returns (X_out, Y); the input array is not modified."""
        g, nl = self._args(grid, proj)
        X = np.array(X, copy=True, order="C")
        Y = np.empty_like(X)
        ncol, ld = X.shape
        nlp = C.byref(nl) if nl is not None else None
        if np.iscomplexobj(X):
            self.lib.oracle_chebyshev_filter_kpt(C.byref(g), nlp, _ptr(veff), _kvec(kvec), _ptr(X),
                                                 C.c_size_t(ld), _ptr(Y), C.c_size_t(ld), C.c_int(ncol),
                                                 C.c_int(m), C.c_double(a), C.c_double(b), C.c_double(a0))
        else:
            self.lib.oracle_chebyshev_filter(C.byref(g), nlp, _ptr(veff), _ptr(X), C.c_size_t(ld), _ptr(Y),
                                             C.c_size_t(ld), C.c_int(ncol), C.c_int(m), C.c_double(a),
                                             C.c_double(b), C.c_double(a0))
        return X, Y

    def fill_random(self, n_per_col, ncol, first_col=0, seed=1):
        out = np.empty((ncol, n_per_col))
        self.lib.oracle_fill_random(_ptr(out), C.c_size_t(n_per_col), C.c_size_t(n_per_col), C.c_int(ncol),
                                    C.c_longlong(first_col), C.c_ulonglong(seed))
        return out


class Reference:
    """The reference's own compiled routines behind oracle/ref_harness.c."""

    def __init__(self, grid, proj=None, veff=None, kvec=None):
        if not reference_available():
            raise RuntimeError("oracle/_ref/libref_harness.so not built (needs /root/reference)")
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_problem_create.restype = C.c_void_p
        self.lib.ref_chebyshev_filter.restype = C.c_double
        self.lib.ref_chebyshev_filter_kpt.restype = C.c_double
        self.grid = grid
        if proj is None:
            from sparc_b200.problem import make_projectors
            proj = make_projectors(grid, np.zeros((0, 3)), 1.0, 1)
        self._g = grid.to_c()
        self._nl = proj.to_c()
        self._veff = np.ascontiguousarray(veff) if veff is not None else None
        self.h = C.c_void_p(self.lib.ref_problem_create(
            C.byref(self._g), C.byref(self._nl),
            _ptr(self._veff) if self._veff is not None else None,
            _kvec(kvec) if kvec is not None else None))

    def __del__(self):
        try:
            self.lib.ref_problem_destroy(self.h)
        except Exception:
            pass

    def lap_plus_diag(self, a, b, c, use_v, x):
        x = np.ascontiguousarray(x)
        y = np.empty_like(x)
        ncol, ld = x.shape
        fn = self.lib.ref_lap_plus_diag_kpt if np.iscomplexobj(x) else self.lib.ref_lap_plus_diag
        fn(self.h, C.c_int(ncol), C.c_double(a), C.c_double(b), C.c_double(c), C.c_int(int(use_v)),
           _ptr(x), C.c_int(ld), _ptr(y), C.c_int(ld))
        return y

    def vnl_mult(self, x, Hx):
        ncol, ld = x.shape
        fn = self.lib.ref_vnl_mult_kpt if np.iscomplexobj(x) else self.lib.ref_vnl_mult
        fn(self.h, C.c_int(ncol), _ptr(x), C.c_int(ld), _ptr(Hx), C.c_int(ld))
        return Hx

    def hamiltonian_mult(self, c, x):
        x = np.ascontiguousarray(x)
        Hx = np.empty_like(x)
        ncol, ld = x.shape
        fn = self.lib.ref_hamiltonian_mult_kpt if np.iscomplexobj(x) else self.lib.ref_hamiltonian_mult
        fn(self.h, C.c_int(ncol), C.c_double(c), _ptr(x), C.c_int(ld), _ptr(Hx), C.c_int(ld))
        return Hx

    def chebyshev_filter(self, X, m, a, b, a0):
        X = np.array(X, copy=True, order="C")
        Y = np.empty_like(X)
        ncol, ld = X.shape
        fn = self.lib.ref_chebyshev_filter_kpt if np.iscomplexobj(X) else self.lib.ref_chebyshev_filter
        fn(self.h, _ptr(X), C.c_int(ld), _ptr(Y), C.c_int(ld), C.c_int(ncol), C.c_int(m), C.c_double(a),
           C.c_double(b), C.c_double(a0))
        return X, Y

    def lanczos(self, x0, tol_min, tol_max, maxit=1000):
        """The reference's own Lanczos (src/eigenSolver.c:1920; complex x0: Lanczos_kpt, src/eigenSolverKpt.c:1361)
        on this problem: (eigmin, eigmax)."""
        cplx = np.iscomplexobj(x0)
        x0 = np.array(x0, dtype=np.complex128 if cplx else np.float64, copy=True, order="C").reshape(-1)
        lo, hi = C.c_double(0), C.c_double(0)
        (self.lib.ref_lanczos_kpt if cplx else self.lib.ref_lanczos)(self.h, _ptr(x0), C.c_double(tol_min), C.c_double(tol_max), C.c_int(maxit),
                             C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def subspace_eig(self, Hp, Mp):
        """The reference's own DP_Solve_Generalized_EigenProblem[_kpt] (src/eigenSolver.c:1262, src/eigenSolverKpt.c:836)
        on column-major n x n Hp, Mp (numpy [n, m] = element (m, n)): (lambda, Q)."""
        cplx = np.iscomplexobj(Hp)
        dt = np.complex128 if cplx else np.float64
        n = Hp.shape[0]
        Hp = np.array(Hp, dtype=dt, copy=True, order="C")
        Mp = np.array(Mp, dtype=dt, copy=True, order="C")
        lam, Q = np.zeros(n), np.zeros((n, n), dtype=dt)
        (self.lib.ref_subspace_eig_kpt if cplx else self.lib.ref_subspace_eig)(self.h, C.c_int(n), _ptr(Hp), _ptr(Mp), _ptr(lam), _ptr(Q))
        return lam, Q

    def density(self, X, occ, occfac=2.0, kptwt=1.0):
        """The reference's own CalculateDensity_psi (src/electronDensity.c:104) for one k-point: rho[Nd]."""
        cplx = np.iscomplexobj(X)
        X = np.array(X, dtype=np.complex128 if cplx else np.float64, copy=True, order="C")
        occ = np.array(occ, dtype=np.float64, copy=True, order="C")
        rho = np.zeros(X.shape[1])
        self.lib.ref_density(self.h, C.c_int(X.shape[0]), _ptr(X), C.c_int(int(cplx)), _ptr(occ), C.c_double(occfac),
                             C.c_double(kptwt), _ptr(rho))
        return rho

    def aar(self, c, x, b, omega=0.6, beta=0.6, m=7, p=6, tol=1e-8, max_iter=1000):
        """The reference's own AAR (src/linearSolver.c:38) with poisson_residual + Jacobi_preconditioner: returns x."""
        x = np.array(x, dtype=np.float64, copy=True, order="C").reshape(-1)
        b = np.array(b, dtype=np.float64, copy=True, order="C").reshape(-1)
        self.lib.ref_aar(self.h, C.c_double(c), _ptr(x), _ptr(b), C.c_double(omega), C.c_double(beta), C.c_int(m),
                         C.c_int(p), C.c_double(tol), C.c_int(max_iter))
        return x

    def gradient_dir(self, c, x, dir, kdir=0.0):
        """The reference's own Gradient_vectors_dir (src/gradVecRoutines.c:32) / Gradient_vectors_dir_kpt
        (src/gradVecRoutinesKpt.c:35; kdir = *kpt_vec) on this problem's grid: Dx."""
        x = np.ascontiguousarray(x)
        Dx = np.empty_like(x)
        ncol, ld = x.shape
        if np.iscomplexobj(x):
            self.lib.ref_gradient_dir_kpt(self.h, C.c_int(ncol), C.c_double(c), _ptr(x), C.c_int(ld), _ptr(Dx), C.c_int(ld),
                                          C.c_int(dir), C.c_double(kdir))
        else:
            self.lib.ref_gradient_dir(self.h, C.c_int(ncol), C.c_double(c), _ptr(x), C.c_int(ld), _ptr(Dx), C.c_int(ld), C.c_int(dir))
        return Dx
