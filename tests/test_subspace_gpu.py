"""Rayleigh-Ritz projection and subspace rotation with the block resident on the device:
chefsi_subspace_project <- DP_Project_Hamiltonian (eigenSolver.c:939-1086), chefsi_subspace_rotate <- DP_Subspace_Rotation
(:1386-1443).  Checked against the oracle's Hamiltonian apply + numpy's FP64 products (the reference calls cblas_dgemm
for the same three GEMMs; only the summation order differs)."""
import numpy as np
import pytest

from sparc_b200 import problem as P
from tests.cases import overlap_case, rel_fro, small_case

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    from sparc_b200.chefsi import ChefsiContext
    c = ChefsiContext(0)
    yield c
    c.close()


def _setup(ctx, g, veff, proj):
    ctx.set_grid(g)
    ctx.set_veff(veff)
    ctx.set_projectors(proj)
    ctx.set_kpoint((0, 0, 0))


def _check(ctx, port, g, veff, proj, y, resident_from_filter=None):
    ncol = y.shape[0]
    Hp, Mp = np.zeros((ncol, ncol + 3)), np.zeros((ncol, ncol + 3))   # ld > ncol
    ctx.DP_Project_Hamiltonian(y, Hp, Mp)
    hy = port.hamiltonian_mult(g, proj, veff, 0.0, y)
    assert rel_fro(Mp[:, :ncol], y @ y.T) < TOL
    assert rel_fro(Hp[:, :ncol], hy @ y.T) < TOL                      # element (m, n) = y_m . (H y_n), stored column-major
    assert (Hp[:, ncol:] == 0).all() and (Mp[:, ncol:] == 0).all()
    rng = np.random.default_rng(5)
    Q = np.ascontiguousarray(rng.standard_normal((ncol, ncol)))        # Q[n, m] = element (m, n) of the column-major matrix
    X = np.full((ncol, g.Nd + 5), 3.0)
    ctx.DP_Subspace_Rotation(Q, X)
    assert rel_fro(X[:, :g.Nd], Q @ y) < TOL
    assert (X[:, g.Nd:] == 3.0).all()


@pytest.mark.parametrize("cell_typ,ncol", [(0, 9), (17, 30), (17, 1)])
def test_project_and_rotate_small(ctx, port, cell_typ, ncol):
    g, veff, proj, y = small_case(cell_typ, ncol=ncol)
    _setup(ctx, g, veff, proj)
    _check(ctx, port, g, veff, proj, y)


def test_project_and_rotate_streaming_kernel_many_columns(ctx, port):
    """More columns than one 64 x 64 GEMM tile, streaming stencil kernel, overlapping spheres, several K slabs."""
    g, veff, proj, y = overlap_case("stream", ncol=70)
    _setup(ctx, g, veff, proj)
    _check(ctx, port, g, veff, proj, y)


def test_filter_keeps_y_resident(ctx, port):
    """ChebyshevFiltering with KEEP_Y and no Y copy-back, then projection and rotation from the device copy: the
    sequence CheFSI runs (eigenSolver.c:325-420); Y never visits the host."""
    g, veff, proj, x = small_case(17, ncol=12)
    _setup(ctx, g, veff, proj)
    a, b, a0 = 0.5, 40.0, -0.6
    ctx.subspace_reserve(12)
    X = x.copy()
    Y = np.full_like(x, np.nan)
    ctx.ChebyshevFiltering(X, Y, 7, a, b, a0, copy_back_x=False, keep_y=True, copy_back_y=False)
    assert np.isnan(Y).all()                                           # untouched on the host
    _, Yw = port.chebyshev_filter(g, proj, veff, x, 7, a, b, a0)
    Hp, Mp = np.zeros((12, 12)), np.zeros((12, 12))
    ctx.DP_Project_Hamiltonian(Y, Hp, Mp)                              # same host address: the device copy is used
    assert rel_fro(Mp, Yw @ Yw.T) < TOL
    assert rel_fro(Hp, port.hamiltonian_mult(g, proj, veff, 0.0, Yw) @ Yw.T) < TOL
    Q = np.ascontiguousarray(np.random.default_rng(1).standard_normal((12, 12)))
    Xr = np.empty_like(x)
    ctx.DP_Subspace_Rotation(Q, Xr)
    assert rel_fro(Xr, Q @ Yw) < TOL


@pytest.mark.parametrize("cell_typ,BC", [(0, (0, 0, 0)), (17, (0, 0, 0)), (0, (1, 1, 1))])
def test_lanczos_extreme_eigenvalues(ctx, port, cell_typ, BC):
    """chefsi_lanczos (Lanczos, eigenSolver.c:1920-2129) against the same iteration done on the host with the oracle's
    H apply and numpy's symmetric tridiagonal eigenvalues (the reference calls LAPACKE_dsterf): same stopping step,
    eigenvalues equal to rounding."""
    g, veff, proj, x = small_case(cell_typ, BC, ncol=1)
    _setup(ctx, g, veff, proj)
    tol = 1e-2
    lo, hi, it = ctx.Lanczos(x[0], tol, tol, maxit=300)
    H = lambda v: port.hamiltonian_mult(g, proj, veff, 0.0, v[None, :].copy())[0]
    vjm1 = x[0] / np.linalg.norm(x[0])
    vj = H(vjm1)
    a = [vjm1 @ vj]
    vj = vj - a[0] * vjm1
    b = [np.linalg.norm(vj)]
    vj = vj / b[0]
    emin_pre = emax_pre = 0.0
    j = 0
    while True:
        vjp1 = H(vj)
        a.append(vj @ vjp1)
        vjp1 = vjp1 - (a[j + 1] * vj + b[j] * vjm1)
        vjm1 = vj
        b.append(np.linalg.norm(vjp1))
        vj = vjp1 / b[j + 1]
        T = np.diag(a[:j + 2]) + np.diag(b[:j + 1], 1) + np.diag(b[:j + 1], -1)
        ev = np.linalg.eigvalsh(T)
        emin, emax = ev[0], ev[-1]
        done = abs(emin - emin_pre) <= tol and abs(emax - emax_pre) <= tol
        emin_pre, emax_pre = emin, emax
        j += 1
        if done or j >= 300:
            break
    assert it == j
    assert abs(lo - emin) < 1e-9 * max(1.0, abs(emax)) and abs(hi - emax) < 1e-9 * max(1.0, abs(emax))
