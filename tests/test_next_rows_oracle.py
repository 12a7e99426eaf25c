"""Pins the numpy restatements of the rows around the filter (oracle/next_rows.py: Lanczos, AAR) to the reference's
OWN compiled routines (oracle/_ref: ref_lanczos -> Lanczos eigenSolver.c:1920, ref_aar -> AAR linearSolver.c:38 with
poisson_residual + Jacobi_preconditioner) on the same inputs.  CPU only."""
import numpy as np
import pytest

from oracle import next_rows
from tests.cases import KVEC, small_case


@pytest.fixture(scope="module")
def ref_cls(have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle.bindings import Reference
    return Reference


@pytest.mark.parametrize("cell_typ,BC", [(0, (0, 0, 0)), (17, (0, 0, 0)), (0, (1, 1, 1))])
def test_lanczos_restatement_vs_reference(port, ref_cls, cell_typ, BC):
    g, veff, proj, x = small_case(cell_typ, BC, ncol=1)
    ref = ref_cls(g, proj, veff)
    tol = 1e-2
    lo_r, hi_r = ref.lanczos(x[0], tol, tol, 300)
    lo, hi, it = next_rows.lanczos(port, g, proj, veff, x[0], tol, tol, 300)
    assert it < 300
    assert abs(lo - lo_r) < 1e-9 * max(1.0, abs(hi_r)) and abs(hi - hi_r) < 1e-9 * max(1.0, abs(hi_r))


@pytest.mark.parametrize("cell_typ,BC", [(0, (0, 0, 0)), (17, (0, 0, 0)), (0, (0, 0, 1))])
def test_lanczos_kpt_restatement_vs_reference(port, ref_cls, cell_typ, BC):
    """Lanczos_kpt (eigenSolverKpt.c:1361): complex vectors, real-part dot products (tools.c:815)."""
    g, veff, proj, x = small_case(cell_typ, BC, ncol=1, complex_=True)
    kvec = tuple(kk if bc == 0 else 0.0 for kk, bc in zip(KVEC, g.BC))
    ref = ref_cls(g, proj, veff, kvec=kvec)
    tol = 1e-2
    lo_r, hi_r = ref.lanczos(x[0], tol, tol, 300)
    lo, hi, it = next_rows.lanczos(port, g, proj, veff, x[0], tol, tol, 300, kvec=kvec)
    assert it < 300
    assert abs(lo - lo_r) < 1e-9 * max(1.0, abs(hi_r)) and abs(hi - hi_r) < 1e-9 * max(1.0, abs(hi_r))


@pytest.mark.parametrize("cell_typ,BC,c", [(0, (1, 1, 1), 0.0), (17, (0, 0, 0), -0.35), (0, (0, 1, 0), -0.2)])
def test_aar_restatement_vs_reference(port, ref_cls, cell_typ, BC, c):
    g, veff, proj, _ = small_case(cell_typ, BC, ncol=1)
    ref = ref_cls(g, proj, veff)
    b = np.random.default_rng(3).standard_normal(g.Nd)
    x0 = np.zeros(g.Nd)
    x_r = ref.aar(c, x0, b, tol=1e-8, max_iter=600)
    x, it, rn = next_rows.aar(port, g, c, x0, b, tol=1e-8, max_iter=600)
    assert it < 600
    assert np.linalg.norm(x - x_r) <= 1e-6 * np.linalg.norm(x_r)
    r = b + port.lap_plus_diag(g, 1.0, 0.0, c, None, x_r[None, :].copy())[0]
    assert np.linalg.norm(r) <= 1.0001e-8 * np.linalg.norm(b)     # the reference's solution meets its own tolerance


def _rand_pencil(n, complex_, seed=5):
    """A Hermitian Hp and a Hermitian positive definite Mp (column-major storage: numpy [n, m] = element (m, n))."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if complex_ else 0)
    Bm = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if complex_ else 0)
    Hp = (A + A.conj().T) / 2
    Mp = Bm @ Bm.conj().T / n + np.eye(n)
    return np.ascontiguousarray(Hp.T), np.ascontiguousarray(Mp.T)


@pytest.mark.parametrize("n,complex_", [(9, False), (30, False), (9, True), (40, True)])
def test_subspace_eig_restatement_vs_reference(ref_cls, n, complex_):
    """next_rows.subspace_eig against the reference's own DP_Solve_Generalized_EigenProblem[_kpt]: eigenvalues equal,
    eigenvectors equal up to a sign / phase per vector."""
    g, veff, proj, _ = small_case(0, (0, 0, 0), ncol=1)
    ref = ref_cls(g, proj, veff)
    Hp, Mp = _rand_pencil(n, complex_)
    lam_r, Q_r = ref.subspace_eig(Hp, Mp)
    lam, Q = next_rows.subspace_eig(Hp, Mp)
    assert np.abs(lam - lam_r).max() < 1e-12 * np.abs(lam_r).max()
    M = Mp.T
    ov = np.einsum("ni,ij,nj->n", Q_r.conj(), M, Q)           # q_r^H Mp q per vector: unit modulus
    assert np.abs(np.abs(ov) - 1).max() < 1e-10
    assert np.abs(Q - Q_r * ov[:, None]).max() < 1e-9


@pytest.mark.parametrize("complex_", [False, True])
def test_density_restatement_vs_reference(ref_cls, complex_):
    g, veff, proj, x = small_case(0, (0, 0, 0), ncol=7, complex_=complex_)
    ref = ref_cls(g, proj, veff)
    occ = np.random.default_rng(1).uniform(0, 1, 7)
    rho_r = ref.density(x, occ, occfac=2.0, kptwt=0.25)
    rho = next_rows.density(x, 2.0 * 0.25 * occ) / g.dV
    assert np.abs(rho - rho_r).max() <= 1e-14 * np.abs(rho_r).max()


@pytest.mark.parametrize("cell_typ,BC,complex_", [(0, (0, 0, 0), False), (17, (0, 0, 0), False), (0, (1, 0, 1), False),
                                                  (0, (0, 0, 0), True), (14, (0, 0, 0), True), (0, (0, 1, 0), True)])
def test_gradient_restatement_vs_reference(ref_cls, cell_typ, BC, complex_):
    """Gradient_vectors_dir[_kpt] (gradVecRoutines.c:32, gradVecRoutinesKpt.c:35): all three directions, with and
    without the diagonal term; for k-points the Bloch phase of the wrapped halo."""
    g, veff, proj, x = small_case(cell_typ, BC, ncol=2, complex_=complex_, with_proj=False)
    ref = ref_cls(g, None, veff)
    for dir in range(3):
        for c in (0.0, -0.37):
            kdir = KVEC[dir] if (complex_ and BC[dir] == 0) else 0.0
            want = ref.gradient_dir(c, x, dir, kdir)
            got = next_rows.gradient_dir(g, c, x, dir, kdir)
            assert np.linalg.norm(got - want) <= 1e-14 * np.linalg.norm(want), (dir, c)
