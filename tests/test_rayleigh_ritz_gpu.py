"""The steps that close the device-resident loop of one SCF iteration:
chefsi_subspace_eig[_kpt] <- DP_Solve_Generalized_EigenProblem[_kpt] (eigenSolver.c:1262-1375, eigenSolverKpt.c:836-930),
chefsi_density_accumulate[_kpt] <- the loop body of CalculateDensity_psi (electronDensity.c:104-200).
Checked against oracle/next_rows.py, which tests/test_next_rows_oracle.py pins to the reference's own compiled routines.
Eigenvectors are defined up to a sign / phase each, so they are compared after aligning that factor, and through the
quantities that do not depend on it (eigenvalues, residual, Mp-orthonormality, the density)."""
import numpy as np
import pytest

from oracle import next_rows  # checker only
from tests.cases import KVEC, rel_fro, small_case
from tests.test_next_rows_oracle import _rand_pencil

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    from sparc_b200.chefsi import ChefsiContext
    c = ChefsiContext(0)
    yield c
    c.close()


def _setup(ctx, g, veff, proj, kvec=(0, 0, 0)):
    ctx.set_grid(g)
    ctx.set_veff(veff)
    ctx.set_projectors(proj)
    ctx.set_kpoint(kvec)


def _check_pencil(Hp, Mp, lam, Q, lam_w):
    """Column-major storage: numpy [n, m] = element (m, n); row n of Q = eigenvector n."""
    H, M, V = Hp.T, Mp.T, Q.T                                          # V[:, n] = eigenvector n
    assert np.abs(lam - lam_w).max() <= TOL * np.abs(lam_w).max()
    assert np.all(np.diff(lam) >= 0)
    assert np.abs(V.conj().T @ M @ V - np.eye(len(lam))).max() < 1e-10
    assert np.linalg.norm(H @ V - M @ V * lam[None, :]) <= 1e-10 * np.linalg.norm(H) * np.sqrt(len(lam))


@pytest.mark.parametrize("n,complex_", [(9, False), (30, False), (257, False), (12, True), (129, True)])
def test_subspace_eig_from_host_matrices(ctx, n, complex_):
    g, veff, proj, _ = small_case(0, ncol=1)
    _setup(ctx, g, veff, proj)
    Hp, Mp = _rand_pencil(n, complex_)
    Hp_pad = np.zeros((n, n + 3), dtype=Hp.dtype)
    Mp_pad = np.zeros((n, n + 3), dtype=Hp.dtype)                       # ld > n
    Hp_pad[:, :n], Mp_pad[:, :n] = Hp, Mp
    lam, Q = ctx.DP_Solve_Generalized_EigenProblem(n, Hp_pad, Mp_pad)
    lam_w, Q_w = next_rows.subspace_eig(Hp, Mp)
    _check_pencil(Hp, Mp, lam, Q, lam_w)
    ov = np.einsum("ni,ij,nj->n", Q_w.conj(), Mp.T, Q)                  # random spectrum: no degeneracies
    assert np.abs(np.abs(ov) - 1).max() < 1e-9
    assert np.abs(Q - Q_w * ov[:, None]).max() < 1e-8


def test_subspace_eig_rejects_an_indefinite_mp(ctx):
    from sparc_b200.capi import ChefsiError
    g, veff, proj, _ = small_case(0, ncol=1)
    _setup(ctx, g, veff, proj)
    Hp, Mp = _rand_pencil(8, False)
    Mp[3, 3] = -5.0
    with pytest.raises(ChefsiError, match="positive definite"):
        ctx.DP_Solve_Generalized_EigenProblem(8, Hp, Mp)


@pytest.mark.parametrize("cell_typ,ncol,complex_", [(0, 9, False), (17, 30, False), (0, 12, True)])
def test_project_eig_rotate_density_all_on_the_device(ctx, port, cell_typ, ncol, complex_):
    """DP_Project_Hamiltonian -> DP_Solve_Generalized_EigenProblem -> DP_Subspace_Rotation -> CalculateDensity_psi
    (eigenSolver.c:349-420, electronDensity.c:47) with Hp, Mp, Q and the rotated block never leaving the device between
    the steps: only lambda and rho (and the rotated block the caller asked for) come back."""
    g, veff, proj, y = small_case(cell_typ, ncol=ncol, complex_=complex_)
    kvec = KVEC if complex_ else (0, 0, 0)
    _setup(ctx, g, veff, proj, kvec)
    ctx.band_store(2)
    Hp, Mp = np.zeros((ncol, ncol), dtype=y.dtype), np.zeros((ncol, ncol), dtype=y.dtype)
    ctx.DP_Project_Hamiltonian(y, Hp, Mp)
    lam, _ = ctx.DP_Solve_Generalized_EigenProblem(ncol, is_complex=complex_, want_Q=False)   # device Hp, Mp
    X = np.empty_like(y)
    ctx.DP_Subspace_Rotation(None, X)                                                          # device Q
    Hp_w, Mp_w = next_rows.project(port, g, proj, veff, y, kvec=kvec if complex_ else None)
    lam_w, Q_w = next_rows.subspace_eig(Hp_w, Mp_w)
    X_w = next_rows.rotate(y, Q_w)
    assert np.abs(lam - lam_w).max() <= TOL * np.abs(lam_w).max()
    ov = np.einsum("ni,ni->n", X_w.conj(), X) / np.einsum("ni,ni->n", X_w.conj(), X_w).real
    assert np.abs(np.abs(ov) - 1).max() < 1e-8
    assert rel_fro(X, X_w * ov[:, None]) < 1e-8                     # eigenvector conditioning, not kernel rounding
    # the rotated block spans the same subspace exactly: X^H-orthonormal and H-diagonal to rounding
    HX = port.hamiltonian_mult(g, proj, veff, 0.0, X, kvec=kvec if complex_ else None)
    assert np.abs(X.conj() @ X.T - np.eye(ncol)).max() < 1e-10
    assert np.abs(X.conj() @ HX.T - np.diag(lam)).max() < 1e-9 * max(1.0, np.abs(lam).max())
    # density from the device copy of X
    occ = np.random.default_rng(2).uniform(0, 1, ncol)
    gw = 2.0 * 0.5 * occ
    rho = np.full(g.Nd, 0.125)
    st0 = ctx.stats()
    ctx.CalculateDensity_psi(X, gw, rho)
    st1 = ctx.stats()
    assert st1["density_resident_blocks"] == st0["density_resident_blocks"] + 1
    assert st1["density_uploaded_blocks"] == st0["density_uploaded_blocks"]
    assert np.abs(rho - 0.125 - next_rows.density(X, gw)).max() <= TOL * np.abs(rho).max()
    # the copy was consumed: the same call now uploads the host block, with the same result
    rho2 = np.zeros(g.Nd)
    ctx.CalculateDensity_psi(X, gw, rho2)
    assert ctx.stats()["density_uploaded_blocks"] == st1["density_uploaded_blocks"] + 1
    assert np.abs(rho2 - next_rows.density(X, gw)).max() <= TOL * np.abs(rho2).max()
    ctx.band_store(0)


def test_filtering_a_block_drops_its_device_copy(ctx, port):
    """The band store is keyed by host address: a ChebyshevFiltering call on that block (its contents change) must
    invalidate the copy, or the density would be the previous iteration's."""
    g, veff, proj, y = small_case(0, ncol=6)
    _setup(ctx, g, veff, proj)
    ctx.band_store(1)
    Hp, Mp = np.zeros((6, 6)), np.zeros((6, 6))
    X = y.copy()
    ctx.DP_Project_Hamiltonian(X, Hp, Mp)
    ctx.DP_Solve_Generalized_EigenProblem(6, want_Q=False)
    ctx.DP_Subspace_Rotation(None, X)                                   # X <- rotated block, copy kept
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, 3, 0.5, 40.0, -0.6, copy_back_x=True)  # X now holds p_{m-1}(H) X0
    rho = np.zeros(g.Nd)
    up0 = ctx.stats()["density_uploaded_blocks"]
    ctx.CalculateDensity_psi(X, np.ones(6), rho)
    assert ctx.stats()["density_uploaded_blocks"] == up0 + 1
    assert np.abs(rho - next_rows.density(X, np.ones(6))).max() <= TOL * np.abs(rho).max()
    ctx.band_store(0)


@pytest.mark.parametrize("complex_", [False, True])
def test_density_many_columns_streaming_grid(ctx, complex_):
    """Several column slabs and a grid large enough for one slab per launch row."""
    from tests.cases import overlap_case
    g, veff, proj, x = overlap_case("stream", ncol=70, complex_=complex_)
    _setup(ctx, g, veff, proj)
    gw = np.random.default_rng(4).uniform(0, 2, 70)
    xp = np.zeros((70, g.Nd + 7), dtype=x.dtype)                        # ld > Nd
    xp[:, :g.Nd] = x
    rho = np.zeros(g.Nd)
    ctx.CalculateDensity_psi(xp, gw, rho)
    assert np.abs(rho - next_rows.density(x, gw)).max() <= TOL * np.abs(rho).max()
