"""Rayleigh-Ritz projection and subspace rotation with the block resident on the device:
chefsi_subspace_project <- DP_Project_Hamiltonian (eigenSolver.c:939-1086), chefsi_subspace_rotate <- DP_Subspace_Rotation
(:1386-1443).  Checked against the oracle's Hamiltonian apply + numpy's FP64 products (the reference calls cblas_dgemm
for the same three GEMMs; only the summation order differs)."""
import numpy as np
import pytest

from oracle import next_rows  # checker only
from sparc_b200 import problem as P
from tests.cases import KVEC, overlap_case, rel_fro, small_case

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


def _check(ctx, port, g, veff, proj, y, kvec=(0, 0, 0)):
    ncol = y.shape[0]
    Hp, Mp = np.zeros((ncol, ncol + 3), dtype=y.dtype), np.zeros((ncol, ncol + 3), dtype=y.dtype)   # ld > ncol
    ctx.DP_Project_Hamiltonian(y, Hp, Mp)
    Hp_w, Mp_w = next_rows.project(port, g, proj, veff, y, kvec=kvec)   # element (m, n) = conj(y_m) . (H y_n) at numpy [n, m]
    assert rel_fro(Mp[:, :ncol], Mp_w) < TOL
    assert rel_fro(Hp[:, :ncol], Hp_w) < TOL
    assert (Hp[:, ncol:] == 0).all() and (Mp[:, ncol:] == 0).all()
    rng = np.random.default_rng(5)
    Q = rng.standard_normal((ncol, ncol))                              # Q[n, m] = element (m, n) of the column-major matrix
    if np.iscomplexobj(y):
        Q = Q + 1j * rng.standard_normal((ncol, ncol))
    Q = np.ascontiguousarray(Q)
    X = np.full((ncol, g.Nd + 5), 3.0, dtype=y.dtype)
    ctx.DP_Subspace_Rotation(Q, X)
    assert rel_fro(X[:, :g.Nd], next_rows.rotate(y, Q)) < TOL
    assert (X[:, g.Nd:] == 3.0).all()


@pytest.mark.parametrize("cell_typ,ncol", [(0, 9), (17, 30), (17, 1)])
def test_project_and_rotate_small(ctx, port, cell_typ, ncol):
    g, veff, proj, y = small_case(cell_typ, ncol=ncol)
    _setup(ctx, g, veff, proj)
    _check(ctx, port, g, veff, proj, y)


@pytest.mark.parametrize("cell_typ,ncol", [(0, 9), (17, 12)])
def test_project_and_rotate_kpt(ctx, port, cell_typ, ncol):
    """Complex (k-point) variants: Hp = Y^H H Y, Mp = Y^H Y (zgemm ConjTrans, eigenSolverKpt.c:749-770), X = Y Q."""
    from tests.cases import KVEC
    g, veff, proj, y = small_case(cell_typ, ncol=ncol, complex_=True)
    _setup(ctx, g, veff, proj)
    ctx.set_kpoint(KVEC)
    _check(ctx, port, g, veff, proj, y, kvec=KVEC)


def test_project_and_rotate_kpt_streaming_kernel(ctx, port):
    from tests.cases import KVEC
    g, veff, proj, y = overlap_case("stream", ncol=70, complex_=True)
    _setup(ctx, g, veff, proj)
    ctx.set_kpoint(KVEC)
    _check(ctx, port, g, veff, proj, y, kvec=KVEC)


def test_project_and_rotate_streaming_kernel_many_columns(ctx, port):
    """More columns than one 64 x 64 GEMM tile, streaming stencil kernel, overlapping spheres, several K slabs."""
    g, veff, proj, y = overlap_case("stream", ncol=70)
    _setup(ctx, g, veff, proj)
    _check(ctx, port, g, veff, proj, y)


@pytest.mark.parametrize("complex_,ncol", [(False, 150), (False, 129), (True, 97)])
def test_project_and_rotate_several_big_tiles(ctx, port, complex_, ncol):
    """More than 64 columns take the 128 x 128 CTA tiles: two tiles per side with a ragged edge, several K slabs."""
    g, veff, proj, y = overlap_case("stream", ncol=ncol, complex_=complex_)
    _setup(ctx, g, veff, proj)
    if complex_:
        ctx.set_kpoint(KVEC)
    _check(ctx, port, g, veff, proj, y, kvec=KVEC if complex_ else (0, 0, 0))


@pytest.mark.parametrize("complex_", [False, True])
def test_filter_keeps_y_resident(ctx, port, complex_):
    """ChebyshevFiltering with KEEP_Y and no Y copy-back, then projection and rotation from the device copy: the
    sequence CheFSI[_kpt] runs (eigenSolver.c:325-420, eigenSolverKpt.c:246-313); Y never visits the host."""
    from tests.cases import KVEC
    g, veff, proj, x = small_case(17, ncol=12, complex_=complex_)
    _setup(ctx, g, veff, proj)
    kvec = KVEC if complex_ else (0, 0, 0)
    ctx.set_kpoint(kvec)
    a, b, a0 = 0.5, 40.0, -0.6
    ctx.subspace_reserve(12, is_complex=complex_)
    X = x.copy()
    Y = np.full_like(x, np.nan)
    ctx.ChebyshevFiltering(X, Y, 7, a, b, a0, copy_back_x=False, keep_y=True, copy_back_y=False)
    assert np.isnan(Y).all()                                           # untouched on the host
    _, Yw = port.chebyshev_filter(g, proj, veff, x, 7, a, b, a0, kvec=kvec)
    Hp, Mp = np.zeros((12, 12), dtype=x.dtype), np.zeros((12, 12), dtype=x.dtype)
    ctx.DP_Project_Hamiltonian(Y, Hp, Mp)                              # same host address: the device copy is used
    assert rel_fro(Mp, Yw @ Yw.conj().T) < TOL
    assert rel_fro(Hp, port.hamiltonian_mult(g, proj, veff, 0.0, Yw, kvec=kvec) @ Yw.conj().T) < TOL
    rng = np.random.default_rng(1)
    Q = rng.standard_normal((12, 12)) + (1j * rng.standard_normal((12, 12)) if complex_ else 0.0)
    Q = np.ascontiguousarray(Q.astype(x.dtype))
    Xr = np.empty_like(x)
    ctx.DP_Subspace_Rotation(Q, Xr)
    assert rel_fro(Xr, Q @ Yw) < TOL


@pytest.mark.parametrize("cell_typ,BC", [(0, (0, 0, 0)), (17, (0, 0, 0)), (0, (1, 1, 1))])
def test_lanczos_extreme_eigenvalues(ctx, port, cell_typ, BC):
    """chefsi_lanczos (Lanczos, eigenSolver.c:1920-2129) against the oracle's restatement (oracle/next_rows.py, pinned to
    the reference's own Lanczos by tests/test_next_rows_oracle.py): same stopping step, eigenvalues equal to rounding."""
    g, veff, proj, x = small_case(cell_typ, BC, ncol=1)
    _setup(ctx, g, veff, proj)
    tol = 1e-2
    lo, hi, it = ctx.Lanczos(x[0], tol, tol, maxit=300)
    emin, emax, j = next_rows.lanczos(port, g, proj, veff, x[0], tol, tol, 300)
    assert it == j
    assert abs(lo - emin) < 1e-9 * max(1.0, abs(emax)) and abs(hi - emax) < 1e-9 * max(1.0, abs(emax))


@pytest.mark.parametrize("cell_typ,BC", [(0, (0, 0, 0)), (17, (0, 0, 0)), (0, (0, 0, 1))])
def test_lanczos_kpt_extreme_eigenvalues(ctx, port, cell_typ, BC):
    """chefsi_lanczos_kpt (Lanczos_kpt, eigenSolverKpt.c:1361-1566: complex vectors, real-part dot products) against the
    oracle's restatement, itself pinned to the reference's own Lanczos_kpt by tests/test_next_rows_oracle.py."""
    g, veff, proj, x = small_case(cell_typ, BC, ncol=1, complex_=True)
    kvec = tuple(kk if bc == 0 else 0.0 for kk, bc in zip(KVEC, g.BC))
    _setup(ctx, g, veff, proj)
    ctx.set_kpoint(kvec)
    tol = 1e-2
    lo, hi, it = ctx.Lanczos(x[0], tol, tol, maxit=300)
    emin, emax, j = next_rows.lanczos(port, g, proj, veff, x[0], tol, tol, 300, kvec=kvec)
    assert it == j
    assert abs(lo - emin) < 1e-9 * max(1.0, abs(emax)) and abs(hi - emax) < 1e-9 * max(1.0, abs(emax))


@pytest.mark.parametrize("cell_typ,BC,c", [(0, (1, 1, 1), 0.0), (17, (0, 0, 0), -0.35), (0, (0, 1, 0), -0.2)])
def test_aar_poisson_solve(ctx, port, cell_typ, BC, c):
    """chefsi_poisson_aar against the oracle's restatement of the reference's AAR (oracle/next_rows.py, pinned to the
    reference's own AAR by tests/test_next_rows_oracle.py; same operator pair, parameters and stopping rule): same
    iteration count up to one stopping test, solution equal to solver tolerance, residual of the returned x small."""
    g, veff, proj, _ = small_case(cell_typ, BC, ncol=1)
    _setup(ctx, g, veff, proj)
    rng = np.random.default_rng(3)
    b = rng.standard_normal(g.Nd)
    if c == 0.0 and BC == (0, 0, 0):
        b -= b.mean()
    x0 = np.zeros(g.Nd)
    x = x0.copy()
    it, rn = ctx.AAR(c, x, b, tol=1e-8, max_iter=600)
    xh, ith, rnh = next_rows.aar(port, g, c, x0, b, 0.6, 0.6, 7, 6, 1e-8, 600)
    assert abs(it - ith) <= 6 and it < 600   # the norm is tested every p = 6 steps; rounding may move the stop by one test
    assert np.linalg.norm(x - xh) <= 1e-5 * np.linalg.norm(xh)
    r = b + port.lap_plus_diag(g, 1.0, 0.0, c, None, x[None, :].copy())[0]
    assert np.linalg.norm(r) <= 1.0001e-8 * np.linalg.norm(b)
    assert abs(rn - np.linalg.norm(r)) <= 1e-6 * np.linalg.norm(r)
