"""CPU restatements (numpy) of the rows around the filter -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module.  Each function cites the reference lines it follows; the Hamiltonian / Laplacian
applications inside them are the oracle's C restatement (`oracle.bindings.Port`), which is itself pinned to the compiled
reference.  These restatements are pinned in turn by tests/test_next_rows_oracle.py against the reference's OWN
`Lanczos` and `AAR` (oracle/_ref/libref_harness.so: ref_lanczos, ref_aar) on the same inputs.
"""
from __future__ import annotations

import numpy as np


def lanczos(port, grid, proj, veff, x0, tol_min, tol_max, maxit=1000, kvec=None):
    """Extreme eigenvalues of H by Lanczos -- src/eigenSolver.c:1920-2129 (line numbers below), and with kvec
    Lanczos_kpt, src/eigenSolverKpt.c:1361-1566: the same loop on complex vectors whose dot product keeps only the
    REAL part (VectorDotProduct_complex accumulates conj(a)*b into a double, src/tools.c:815-826) -- for a Hermitian
    H the imaginary part is rounding noise anyway.  Returns (eigmin, eigmax, iterations)."""
    H = lambda v: port.hamiltonian_mult(grid, proj, veff, 0.0, v[None, :].copy(), kvec=kvec)[0]
    dot = (lambda u, v: u @ v) if kvec is None else (lambda u, v: float(np.real(np.vdot(u, v))))
    vjm1 = x0 / np.linalg.norm(x0)                      # :1986-1992
    vj = H(vjm1)                                        # :2003
    a = [dot(vjm1, vj)]                                 # :2012
    vj = vj - a[0] * vjm1                               # :2014-2015
    b = [np.linalg.norm(vj)]                            # :2017
    vj = vj / b[0]                                      # :2036-2038
    emin = emax = emin_pre = emax_pre = 0.0
    j = 0
    while True:
        vjp1 = H(vj)                                    # :2048
        a.append(dot(vj, vjp1))                         # :2054
        vjp1 = vjp1 - (a[j + 1] * vj + b[j] * vjm1)     # :2056-2061
        vjm1 = vj
        b.append(np.linalg.norm(vjp1))                  # :2063
        if b[j + 1] == 0.0:
            break
        vj = vjp1 / b[j + 1]                            # :2068-2071
        T = np.diag(a[:j + 2]) + np.diag(b[:j + 1], 1) + np.diag(b[:j + 1], -1)
        ev = np.linalg.eigvalsh(T)                      # LAPACKE_dsterf, :2075
        emin, emax = ev[0], ev[-1]
        err_min, err_max = abs(emin - emin_pre), abs(emax - emax_pre)
        emin_pre, emax_pre = emin, emax
        j += 1
        if not ((err_min > tol_min or err_max > tol_max) and j < maxit):   # :2044
            break
    return emin, emax, j


def aar(port, grid, c, x, b, omega=0.6, beta=0.6, m=7, p=6, tol=1e-8, max_iter=1000):
    """Alternating Anderson-Richardson solve of -(Lap + c) x = b -- src/linearSolver.c:38-146 with
    res_fun = poisson_residual (src/lapVecRoutines.c:61-79: r = b + (Lap + c) x) and precond_fun =
    Jacobi_preconditioner (src/electrostatics.c:1682-1700); Anderson step src/mixing.c:48-140 (minimum-norm least
    squares, LAPACKE_dgelsd).  Returns (x, iterations, ||r||)."""
    lap = lambda v: port.lap_plus_diag(grid, 1.0, 0.0, c, None, v[None, :].copy())[0]
    N = x.size
    m_inv = grid.coefs["D2_x"][0] + grid.coefs["D2_y"][0] + grid.coefs["D2_z"][0] + c
    m_inv = -1.0 / (1.0 if abs(m_inv) < 1e-14 else m_inv)
    x = x.copy()
    x_old, f_old = x.copy(), np.zeros(N)
    X, F = np.zeros((m, N)), np.zeros((m, N))          # calloc'd histories, :66-67
    r = b + lap(x)                                     # :82
    tol = tol * np.linalg.norm(b)                      # :84
    r_2norm, it = tol + 1.0, 0
    while r_2norm > tol and it < max_iter:
        f = m_inv * r                                  # :88
        if it > 0:                                     # :89-95
            h = (it - 1) % m
            X[h], F[h] = x - x_old, f - f_old
        x_old, f_old = x.copy(), f.copy()
        if (it + 1) % p == 0 and it > 0:               # Anderson, :102-121
            G = np.linalg.lstsq(F @ F.T, F @ f, rcond=None)[0]
            x = x_old - G @ X + beta * (f - G @ F)
            r = b + lap(x)
            r_2norm = np.linalg.norm(r)
        else:                                          # Richardson, :122-133
            x = x_old + omega * f
            r = b + lap(x)
        it += 1
    return x, it, r_2norm


def project(port, grid, proj, veff, Y, kvec=None):
    """Hp = Y^H H Y, Mp = Y^H Y in the reference's column-major storage (numpy [n, m] = element (m, n)) --
    src/eigenSolver.c:939-1086, src/eigenSolverKpt.c:676-790."""
    HY = port.hamiltonian_mult(grid, proj, veff, 0.0, Y, kvec=kvec)
    return HY @ Y.conj().T, Y @ Y.conj().T


def rotate(Y, Q):
    """X = Y Q with Q in column-major storage (numpy Q[n, m] = element (m, n)) -- src/eigenSolver.c:1386-1443."""
    return Q @ Y


def subspace_eig(Hp, Mp):
    """Hp q = lambda Mp q, eigenvalues ascending, Q^H Mp Q = I -- DP_Solve_Generalized_EigenProblem, src/eigenSolver.c:
    1262-1375 (LAPACKE_dsygvd, itype 1) and its k-point twin, src/eigenSolverKpt.c:836-930 (LAPACKE_zhegvd).  Input and
    output in the reference's column-major storage (numpy [n, m] = element (m, n); row n of Q = eigenvector n).
    scipy's driver "gvd" is the same LAPACK routine.  Eigenvectors are defined up to a sign / phase each."""
    import scipy.linalg
    lam, V = scipy.linalg.eigh(Hp.T, Mp.T, driver="gvd", type=1)
    return lam, np.ascontiguousarray(V.T)


def density(X, g):
    """rho[i] = sum_n g[n] |X[n, i]|^2 -- the loop body of CalculateDensity_psi, src/electronDensity.c:135-156 (g[n] =
    occfac * kptWts_loc[k] / Nkpts * occ[n]; the caller scales by 1/dV, :190-196)."""
    return np.einsum("n,ni->i", np.asarray(g, dtype=np.float64), (X.real ** 2 + X.imag ** 2) if np.iscomplexobj(X) else X * X)


def gradient_dir(grid, c, x, dir, kdir=0.0):
    """(D_dir + c) x for a block x[ncol, Nd] -- Gradient_vec_dir, src/gradVecRoutines.c:59-311 (np = 1 branch): x is
    extended by FDn points along `dir` only, wrapped on a periodic axis (:262-284) or zero on a Dirichlet axis (:285-298),
    then Calc_DX (:318-409): temp = c x; temp += (x[+r] - x[-r]) w[r], r = 1..FDn, w = D1_stencil_coeffs_{x,y,z}.
    Complex x: Gradient_vec_dir_kpt, src/gradVecRoutinesKpt.c:63-340: wrapped values from beyond the low face are
    multiplied by cos(k L) - i sin(k L), from beyond the high face by its conjugate (:179-191,301-311)."""
    F = grid.FDn
    N3 = (int(grid.N[2]), int(grid.N[1]), int(grid.N[0]))
    ax = 3 - dir                       # numpy axis of the lattice direction in x.reshape(ncol, Nz, Ny, Nx)
    N = N3[2 - dir]
    bc = grid.BC[dir]
    L = grid.L[dir]
    w = grid.coefs[("D1_x", "D1_y", "D1_z")[dir]]
    v = x.reshape((x.shape[0],) + N3)
    lo = np.take(v, range(N - F, N), axis=ax)
    hi = np.take(v, range(0, F), axis=ax)
    if bc:
        lo, hi = np.zeros_like(lo), np.zeros_like(hi)
    elif np.iscomplexobj(x):
        ph = np.cos(kdir * L) - 1j * np.sin(kdir * L)
        lo, hi = lo * ph, hi * np.conj(ph)
    ex = np.concatenate([lo, v, hi], axis=ax)
    sl = lambda s: np.take(ex, range(F + s, F + s + N), axis=ax)
    out = v * c
    for r in range(1, F + 1):
        out = out + (sl(r) - sl(-r)) * w[r]
    return out.reshape(x.shape)
