"""Parity tests proper: the CUDA path, called through the C ABI (libchefsi_b200.so), against
the oracle on the same seeded inputs, against the committed reference-generated vectors, and --
at sizes the oracle cannot reach in seconds -- through size-independent properties.

Tolerance: north_star's <= 1e-10 relative Frobenius error on filtered vectors (FP64 throughout;
only summation order differs from the reference).
"""
import numpy as np
import pytest

from sparc_b200 import problem as P
from tests.cases import (BOUNDS, GOLDEN, KVEC, OVERLAP_CASES, SPARC_GOLDEN, load_golden, overlap_case, rel_fro, small_case,
                         sphere_overlap_count)

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    from sparc_b200.chefsi import ChefsiContext
    c = ChefsiContext(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_zmarch():
    """Context that keeps launches with few CTAs on the z-march kernel (by default single columns and other tiny
    launches take the 3-D brick kernel, which also cuts z: CHEFSI_B200_SMALL_BRICK)."""
    import os
    from sparc_b200.chefsi import ChefsiContext
    os.environ["CHEFSI_B200_SMALL_BRICK"] = "0"
    try:
        c = ChefsiContext(0)
    finally:
        del os.environ["CHEFSI_B200_SMALL_BRICK"]
    yield c
    c.close()


def _setup(ctx, g, veff, proj, kvec=None):
    ctx.set_grid(g)
    ctx.set_veff(veff)
    ctx.set_projectors(proj)
    ctx.set_kpoint(kvec if kvec is not None else (0, 0, 0))


# ---------------------------------------------------------------- golden vectors (reference-made)
@pytest.mark.parametrize("name", GOLDEN)
def test_golden_vectors(ctx, name):
    g, veff, proj, d = load_golden(name)
    _setup(ctx, g, veff, proj, tuple(d["kvec"]))
    a, b, a0 = d["bounds"]
    x = np.ascontiguousarray(d["X0"])
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(float(d["c_shift"]), x, Hx)
    assert rel_fro(Hx, d["Hx"]) < TOL
    X = x.copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, int(d["m"]), a, b, a0)
    assert rel_fro(Y, d["Y_out"]) < TOL
    assert rel_fro(X, d["X_out"]) < TOL
    if name.startswith("stream_"):  # the two fixtures sized for the TMA streaming kernels (real / k-point)
        assert ctx.stats()["last_path"] == 1


@pytest.mark.parametrize("name", SPARC_GOLDEN)
def test_real_sparc_filter_calls(ctx, name):
    """Dumps of real ChebyshevFiltering[_kpt] calls from the reference's SCF runs of Si8 (cell_typ 17), BaTiO3
    (orthogonal, 32 projectors on Ba, overlapping spheres) and Si8_kpt (complex, cell_typ 17): real psp8/spline Chi,
    real Gamma, the SCF's Veff and bounds; outputs are the reference routine's own (tests/golden/make_sparc_dumps.py)."""
    g, veff, proj, d = load_golden(name)
    _setup(ctx, g, veff, proj, tuple(d["kvec"]))
    a, b, a0 = d["bounds"]
    X = np.ascontiguousarray(d["X0"]).copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, int(d["m"]), a, b, a0)
    assert rel_fro(Y, d["Y_out"]) < TOL
    assert rel_fro(X, d["X_out"]) < TOL
    if name == "sparc_batio3":
        assert ctx.stats()["last_nloc_atomic"] == 1   # real overlapping spheres: the scatter-add branch


# ---------------------------------------------------------------- every cell type / BC vs the oracle
@pytest.mark.parametrize("cell_typ", [0, 11, 12, 13, 14, 15, 16, 17])
@pytest.mark.parametrize("BC", [(0, 0, 0), (0, 1, 0), (1, 1, 1)])
@pytest.mark.parametrize("complex_", [False, True])
def test_hamiltonian_all_cell_types(ctx_zmarch, port, cell_typ, BC, complex_):
    ctx = ctx_zmarch
    g, veff, proj, x = small_case(cell_typ, BC, complex_=complex_)
    _setup(ctx, g, veff, proj, KVEC)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(-0.3, x, Hx)
    assert ctx.stats()["last_path"] == (0 if (cell_typ == 15 and complex_) else 2)
    want = port.hamiltonian_mult(g, proj, veff, -0.3, x, kvec=KVEC)
    assert rel_fro(Hx, want) < TOL


@pytest.mark.parametrize("cell_typ", [0, 14, 17])
@pytest.mark.parametrize("complex_", [False, True])
def test_chebyshev_filter_small(ctx_zmarch, port, cell_typ, complex_):
    ctx = ctx_zmarch
    g, veff, proj, x = small_case(cell_typ, complex_=complex_, ncol=5)
    _setup(ctx, g, veff, proj, KVEC)
    a, b, a0 = BOUNDS
    X = x.copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, 12, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 12, a, b, a0, kvec=KVEC)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


def test_leading_dimension_larger_than_grid(ctx, port):
    """ld != rows (collinear spin: ldi = DMnd * Nspinor_spincomm, eigenSolver.c:325-328)."""
    g, veff, proj, x = small_case(0, ncol=3)
    _setup(ctx, g, veff, proj)
    ld = 2 * g.Nd + 3
    X = np.full((3, ld), 7.0)
    X[:, :g.Nd] = x
    Y = np.full((3, ld), -3.0)
    a, b, a0 = BOUNDS
    ctx.ChebyshevFiltering(X, Y, 6, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 6, a, b, a0)
    assert rel_fro(Y[:, :g.Nd], Yw) < TOL and rel_fro(X[:, :g.Nd], Xw) < TOL
    assert (X[:, g.Nd:] == 7.0).all() and (Y[:, g.Nd:] == -3.0).all()


def test_degree_one_and_no_x_copyback(ctx, port):
    g, veff, proj, x = small_case(0, ncol=2)
    _setup(ctx, g, veff, proj)
    a, b, a0 = BOUNDS
    X = x.copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, 1, a, b, a0, copy_back_x=False)
    _, Yw = port.chebyshev_filter(g, proj, veff, x, 1, a, b, a0)
    assert rel_fro(Y, Yw) < TOL
    assert np.array_equal(X, x)  # untouched on the host when copy-back is off


def test_no_projectors_no_veff_lanczos_style_call(ctx, port):
    """ncol = 1, c = 0 as Lanczos calls Hamiltonian_vectors_mult (eigenSolver.c:2002)."""
    g, veff, proj, x = small_case(17, ncol=1, with_proj=False)
    _setup(ctx, g, veff, None)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(0.0, x, Hx)
    assert rel_fro(Hx, port.hamiltonian_mult(g, None, veff, 0.0, x)) < TOL


@pytest.mark.parametrize("cell_typ,N,BC,path", [(0, (14, 13, 15), (0, 0, 0), 0), (17, (14, 13, 15), (0, 1, 0), 0),
                                                (0, (32, 32, 16), (0, 0, 0), 1), (0, (40, 39, 12), (1, 1, 1), 1),
                                                (17, (32, 32, 14), (0, 0, 0), 3), (14, (64, 40, 13), (0, 0, 1), 3)])
def test_lap_vec_mult(ctx, port, cell_typ, N, BC, path):
    """Lap_vec_mult (lapVecRoutines.c:37-58): (Lap + c) x without potential and projectors -- the operator of the
    Poisson residual / Kerker preconditioner -- on the kernels the default dispatch picks (tiny launches: the 3-D
    brick kernel; streaming grids: K1 / K2m), single column and a small block, with a potential and projectors set on
    the context (they must not be applied)."""
    g = P.make_grid(N, tuple(0.45 * n for n in N), BC=BC, latvec=P.LATVEC_BY_CELL_TYP[cell_typ])
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array([[0.3, 0.5, 0.6]]), rc=[2.0], nproj=[5])
    _setup(ctx, g, veff, proj)
    for ncol, c in ((1, 0.0), (3, -0.02)):
        x = P.random_columns(g.Nd, ncol, seed=23)
        y = np.empty_like(x)
        ctx.Lap_vec_mult(c, x, y)
        assert ctx.stats()["last_path"] == path
        assert rel_fro(y, port.lap_plus_diag(g, 1.0, 0.0, c, None, x)) < TOL


@pytest.mark.parametrize("cell_typ,N,BC,FDn", [(0, (14, 13, 15), (0, 0, 0), 6), (17, (14, 13, 15), (0, 1, 0), 6),
                                               (0, (40, 39, 21), (1, 0, 1), 6), (14, (64, 40, 33), (0, 0, 0), 6),
                                               (0, (14, 13, 15), (0, 0, 1), 4), (0, (37, 14, 13), (0, 0, 0), 6),
                                               (0, (131, 9, 8), (1, 0, 0), 6)])
@pytest.mark.parametrize("complex_", [False, True])
def test_gradient_vectors_dir(ctx, cell_typ, N, BC, FDn, complex_):
    """Gradient_vectors_dir[_kpt] (gradVecRoutines.c:32, gradVecRoutinesKpt.c:35): (D_dir + c) x along each lattice
    direction -- the marching kernel (y, z at FD radius 6, several chunks along the marching axis), the shared-memory
    row kernel (x at radius 6; even and odd Nx, rows longer than two warp passes) and the gather kernel (any other
    radius) -- periodic / Dirichlet faces, Bloch phases for complex columns, single column and a small
    block, vs the numpy restatement pinned to the compiled reference (tests/test_next_rows_oracle.py)."""
    from oracle import next_rows
    g = P.make_grid(N, tuple(0.45 * n for n in N), BC=BC, FDn=FDn, latvec=P.LATVEC_BY_CELL_TYP[cell_typ])
    _setup(ctx, g, P.synthetic_veff(g), None)
    for ncol, c in ((1, 0.0), (3, -0.37)):
        x = P.random_columns(g.Nd, ncol, seed=29)
        if complex_:
            x = np.ascontiguousarray(x + 1j * P.random_columns(g.Nd, ncol, first_col=500, seed=29))
        for dir in range(3):
            kdir = KVEC[dir] if (complex_ and BC[dir] == 0) else 0.0
            Dx = np.empty_like(x)
            ctx.Gradient_vectors_dir(c, x, Dx, dir, kdir)
            assert rel_fro(Dx, next_rows.gradient_dir(g, c, x, dir, kdir)) < 1e-13, (dir, ncol)


def test_fd_radius_four(ctx, port):
    g, veff, proj, x = small_case(17, FDn=4)
    _setup(ctx, g, veff, proj)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(0.1, x, Hx)
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, 0.1, x)) < TOL


# ---------------------------------------------------------------- overlapping rc-spheres
@pytest.mark.parametrize("name", sorted(OVERLAP_CASES))
@pytest.mark.parametrize("complex_", [False, True])
def test_overlapping_spheres(ctx, port, name, complex_):
    """Atoms closer than rc1 + rc2 and atoms whose own periodic images overlap: the unchained
    PROJECT-per-step sequence with the FP64-atomic expand (nloc.cu MODE_EXPAND_ATOMIC) and, where an atom has
    more than 8 alpha partials, alpha_reduce_kernel -- the branch Si8 and BaTiO3 take inside SPARC -- against
    the oracle at 1e-10 (reference: beta = 1 accumulation over images nlocVecRoutines.c:821-827, overlapping
    scatter-add :866-881)."""
    g, veff, proj, x = overlap_case(name, complex_=complex_, ncol=5)
    assert sphere_overlap_count(proj, g.Nd) > 0
    kvec = tuple(kk if bc == 0 else 0.0 for kk, bc in zip(KVEC, g.BC))
    _setup(ctx, g, veff, proj, kvec)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(-0.3, x, Hx)
    st = ctx.stats()
    assert st["last_nloc_atomic"] == 1
    assert st["last_path"] == (1 if name.startswith("stream") else 0)   # 5 columns on a 14 x 13 x 15 grid: a tiny launch -> brick kernel
    if name in ("general_small", "general_typ17", "stream"):
        assert st["last_alpha_reduced"] == 1   # 9 alpha partials on one atom
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, -0.3, x, kvec=kvec)) < TOL
    a, b, a0 = 0.5, (1.01 * g.max_eig_mhalf_lap() + 0.5) if g.cell_typ == 0 else 40.0, -0.6
    X, Y = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X, Y, 9, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 9, a, b, a0, kvec=kvec)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


def test_overlapping_spheres_many_columns(ctx, port):
    """The overlapping-sphere branch on more than one 32-column projector group (70 columns) and with the
    alpha partial threshold forced both ways."""
    import os
    from sparc_b200.chefsi import ChefsiContext
    g, veff, proj, x = overlap_case("stream", ncol=70)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 5, a, b, a0)
    for thr in ("0", "1000"):
        os.environ["CHEFSI_B200_ALPHA_REDUCE_MIN"] = thr
        try:
            c = ChefsiContext(0)
        finally:
            del os.environ["CHEFSI_B200_ALPHA_REDUCE_MIN"]
        _setup(c, g, veff, proj)
        X, Y = x.copy(), np.empty_like(x)
        c.ChebyshevFiltering(X, Y, 5, a, b, a0)
        assert c.stats()["last_alpha_reduced"] == (1 if thr == "0" else 0)
        c.close()
        assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


# ---------------------------------------------------------------- streaming orthogonal kernel
@pytest.fixture(scope="module")
def ctx_dense(ctx):
    """The dense-layout TMA streaming kernel (stencil_stream_dense.cu) is the only real orthogonal streaming kernel."""
    return ctx


@pytest.mark.parametrize("N,BC", [((32, 32, 24), (0, 0, 0)), ((48, 40, 20), (0, 0, 0)), ((36, 38, 16), (0, 0, 0)),
                                   ((96, 96, 13), (0, 0, 0)), ((32, 32, 16), (1, 0, 1)), ((64, 32, 16), (0, 1, 0)),
                                   ((32, 64, 12), (1, 1, 1)), ((40, 70, 12), (0, 0, 1)), ((68, 36, 12), (0, 1, 0)),
                                   ((40, 39, 12), (0, 0, 0)), ((44, 37, 14), (0, 1, 0)),
                                   ((38, 40, 12), (0, 0, 0)), ((70, 33, 12), (1, 1, 0)),
                                   ((34, 32, 12), (0, 0, 0)), ((36, 64, 12), (0, 0, 1))])
def test_dense_stream_kernel_vs_oracle(ctx_dense, port, N, BC):
    """Streaming kernel: single tiles that wrap on both sides, shifted (overlapping) last
    tiles, interior tiles (one TMA box), periodic-x strips, split periodic-y boxes, Dirichlet faces
    (TMA zero fill)."""
    ctx = ctx_dense
    g = P.make_grid(N, tuple(0.45 * n for n in N), BC=BC)
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array([[0.02, 0.5, 0.97], [0.5, 0.5, 0.5]]), rc=[2.4, 2.0], nproj=[18, 7])
    x = P.random_columns(g.Nd, 3, seed=11)
    _setup(ctx, g, veff, proj)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(-0.3, x, Hx)
    assert ctx.stats()["last_path"] == 1
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, -0.3, x)) < TOL
    X = x.copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, 8, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 8, a, b, a0)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


# disjoint spheres on grids the dense streaming kernel takes: several atoms per tile, spheres cut by tile edges,
# by a shifted (overlapping) last tile and by periodic / Dirichlet faces, segmented and whole sphere images
DISJOINT_CASES = {
    # name: (N, BC, frac, rc, nproj, ncol, degree)
    "one_tile": ((32, 32, 24), (0, 0, 0), [[0.02, 0.5, 0.97], [0.5, 0.5, 0.5]], [2.4, 2.0], [18, 7], 3, 8),
    "shifted_tiles": ((48, 40, 20), (0, 0, 0), [[0.64, 0.78, 0.3], [0.2, 0.2, 0.8], [0.97, 0.03, 0.5]], [2.3, 2.6, 2.0], [18, 13, 7], 5, 6),
    "dirichlet": ((64, 32, 16), (1, 0, 1), [[0.05, 0.5, 0.1], [0.5, 0.97, 0.5], [0.8, 0.4, 0.9]], [2.4, 2.0, 2.2], [18, 7, 9], 4, 5),
    "many_atoms": ((96, 96, 13), (0, 0, 0), [[(i + 0.37) / 6, (j + 0.61) / 6, 0.21 + 0.5 * ((i + j) % 2)]
                                            for i in range(6) for j in range(6)], 2.1, 18, 6, 4),
    "degree_two": ((36, 64, 12), (0, 0, 1), [[0.3, 0.3, 0.5], [0.7, 0.8, 0.4]], [2.2, 2.5], [26, 18], 2, 2),
    "many_columns": ((64, 64, 12), (0, 0, 0), [[0.25, 0.25, 0.5], [0.75, 0.75, 0.1], [0.5, 0.02, 0.6]], [2.4, 2.4, 2.0], [18, 18, 5], 70, 3),
}


def _disjoint_case(name):
    N, BC, frac, rc, nproj, ncol, m = DISJOINT_CASES[name]
    g = P.make_grid(N, tuple(0.45 * n for n in N), BC=BC)
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array(frac), rc=rc, nproj=nproj, seed=5)
    assert sphere_overlap_count(proj, g.Nd) == 0
    x = P.random_columns(g.Nd, ncol, seed=17)
    return g, veff, proj, x, m


@pytest.mark.parametrize("name", sorted(DISJOINT_CASES))
def test_fused_projector_chain_disjoint_spheres(ctx, port, name):
    """Disjoint spheres on the dense streaming path: one stencil launch + one FUSED projector launch per degree
    (expand with the input's alpha, project the final output for the next degree), alpha partials per image /
    segment summed in a fixed order, first degree PROJECT, last degree EXPAND."""
    g, veff, proj, x, m = _disjoint_case(name)
    _setup(ctx, g, veff, proj)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    X, Y = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X, Y, m, a, b, a0)
    st = ctx.stats()
    assert st["last_path"] == 1 and st["last_nloc_atomic"] == 0
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, m, a, b, a0)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


@pytest.mark.parametrize("cell_typ", [11, 12, 13, 14, 15, 16, 17])
@pytest.mark.parametrize("N,BC", [((32, 32, 14), (0, 0, 0)), ((64, 40, 13), (0, 0, 0)), ((38, 70, 12), (0, 0, 0)),
                                   ((32, 32, 16), (1, 0, 1)), ((36, 64, 12), (0, 1, 0)), ((40, 39, 12), (1, 1, 1)),
                                   ((96, 32, 12), (0, 0, 1))])
def test_mixed_stream_kernel_vs_oracle(ctx, port, cell_typ, N, BC):
    """Non-orthogonal TMA streaming kernel (stencil_stream_mixed.cu): every mixed-derivative flavour (x-y in-plane
    two-stage term through the per-warp D tile, x-z / y-z terms through the register z-queue), single tiles that wrap
    on both sides (corner halos from the strips), shifted last tiles, interior tiles, Dirichlet faces, with
    projectors; H apply and a degree-8 filter against the oracle's two-stage composition."""
    g = P.make_grid(N, tuple(0.45 * n for n in N), BC=BC, latvec=P.LATVEC_BY_CELL_TYP[cell_typ])
    assert g.cell_typ == cell_typ
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array([[0.02, 0.5, 0.97], [0.5, 0.5, 0.5]]), rc=[2.4, 2.0], nproj=[18, 7])
    x = P.random_columns(g.Nd, 3, seed=17)
    _setup(ctx, g, veff, proj)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(-0.3, x, Hx)
    assert ctx.stats()["last_path"] == 3
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, -0.3, x)) < TOL
    a, b, a0 = 0.5, 150.0, -0.6
    X, Y = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X, Y, 8, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 8, a, b, a0)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


def test_mixed_stream_many_columns(ctx, port):
    """More work items than SMs on the non-orthogonal streaming kernel (persistent CTAs, round barrier)."""
    g = P.make_grid((64, 64, 12), (28.8, 28.8, 5.4), latvec=P.SI8_LATVEC)
    veff = P.synthetic_veff(g)
    x = P.random_columns(g.Nd, 80, seed=5)
    _setup(ctx, g, veff, None)
    a, b, a0 = 0.5, 150.0, -0.6
    X, Y = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X, Y, 4, a, b, a0)
    assert ctx.stats()["last_path"] == 3
    Xw, Yw = port.chebyshev_filter(g, None, veff, x, 4, a, b, a0)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


@pytest.mark.parametrize("N,BC", [((16, 32, 12), (0, 0, 0)), ((32, 32, 16), (0, 0, 0)), ((40, 38, 14), (0, 0, 0)),
                                   ((48, 64, 12), (0, 0, 0)), ((18, 39, 13), (0, 0, 0)), ((34, 70, 12), (0, 0, 0)),
                                   ((32, 32, 16), (1, 0, 1)), ((36, 32, 12), (0, 1, 0)), ((32, 40, 12), (1, 1, 1)),
                                   ((38, 33, 12), (0, 1, 1))])
def test_kpt_stream_kernel_vs_oracle(ctx, port, N, BC):
    """k-point (complex) streaming kernel: Bloch phases on periodic-x strips, wrapped y boxes and z wrap planes,
    shifted last tiles, Dirichlet faces, with projectors; H apply and a degree-8 filter against the oracle."""
    g = P.make_grid(N, tuple(0.45 * n for n in N), BC=BC)
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array([[0.02, 0.5, 0.97], [0.5, 0.5, 0.5]]), rc=[2.4, 2.0], nproj=[18, 7])
    x = P.random_columns(g.Nd, 3, seed=21) + 1j * P.random_columns(g.Nd, 3, first_col=500, seed=21)
    x = np.ascontiguousarray(x)
    kvec = tuple(kk if bc == 0 else 0.0 for kk, bc in zip(KVEC, BC))  # SPARC has no k component along Dirichlet axes
    _setup(ctx, g, veff, proj, kvec)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(-0.3, x, Hx)
    assert ctx.stats()["last_path"] == 1
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, -0.3, x, kvec=kvec)) < TOL
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    X = x.copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, 8, a, b, a0)
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 8, a, b, a0, kvec=kvec)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


def test_kpt_stream_many_columns(ctx, port):
    """More k-point work items than SMs (round barrier, ring wrap-around across items)."""
    g = P.make_grid((32, 64, 12), (14.4, 28.8, 5.4))
    veff = P.synthetic_veff(g)
    x = np.ascontiguousarray(P.random_columns(g.Nd, 50, seed=6) + 1j * P.random_columns(g.Nd, 50, first_col=900, seed=6))
    _setup(ctx, g, veff, None, KVEC)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    X, Y = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X, Y, 4, a, b, a0)
    assert ctx.stats()["last_path"] == 1
    Xw, Yw = port.chebyshev_filter(g, None, veff, x, 4, a, b, a0, kvec=KVEC)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


def test_dense_stream_many_columns_round_barrier(ctx_dense, port):
    """More work items than SMs: the persistent CTAs loop over items and the producers' round barrier runs."""
    ctx = ctx_dense
    g = P.make_grid((64, 64, 12), (28.8, 28.8, 5.4))
    veff = P.synthetic_veff(g)
    x = P.random_columns(g.Nd, 80, seed=5)
    _setup(ctx, g, veff, None)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    X, Y = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X, Y, 4, a, b, a0)
    assert ctx.stats()["last_path"] == 1
    Xw, Yw = port.chebyshev_filter(g, None, veff, x, 4, a, b, a0)
    assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL


def test_stream_zmarch_and_brick_kernels_agree(ctx):
    """The three stencil kernels (TMA streaming, z-march, 3-D brick) on the same orthogonal problem."""
    import os
    from sparc_b200.chefsi import ChefsiContext
    g = P.make_grid((64, 48, 40), (20.0, 15.0, 12.5))
    veff = P.synthetic_veff(g)
    x = P.random_columns(g.Nd, 4, seed=2)
    _setup(ctx, g, veff, None)
    a, b, a0 = P.chebyshev_bounds(g)
    X1, Y1 = x.copy(), np.empty_like(x)
    ctx.ChebyshevFiltering(X1, Y1, 10, a, b, a0)
    assert ctx.stats()["last_path"] == 1
    for level, path, tol in ((1, 2, 1e-12), (2, 0, 1e-12)):
        os.environ["CHEFSI_B200_FORCE_GENERAL"] = str(level)
        os.environ["CHEFSI_B200_SMALL_BRICK"] = "0"
        try:
            c2 = ChefsiContext(0)
        finally:
            del os.environ["CHEFSI_B200_FORCE_GENERAL"]
            del os.environ["CHEFSI_B200_SMALL_BRICK"]
        _setup(c2, g, veff, None)
        X2, Y2 = x.copy(), np.empty_like(x)
        c2.ChebyshevFiltering(X2, Y2, 10, a, b, a0)
        assert c2.stats()["last_path"] == path
        c2.close()
        assert rel_fro(Y1, Y2) < tol and rel_fro(X1, X2) < tol


@pytest.fixture(scope="module")
def ctx_brick():
    """Context restricted to the 3-D brick kernel (CHEFSI_B200_FORCE_GENERAL=2), the first general kernel."""
    import os
    from sparc_b200.chefsi import ChefsiContext
    os.environ["CHEFSI_B200_FORCE_GENERAL"] = "2"
    try:
        c = ChefsiContext(0)
    finally:
        del os.environ["CHEFSI_B200_FORCE_GENERAL"]
    yield c
    c.close()


@pytest.mark.parametrize("cell_typ", [0, 11, 12, 13, 14, 15, 16, 17])
@pytest.mark.parametrize("complex_", [False, True])
def test_brick_kernel_all_cell_types(ctx_brick, port, cell_typ, complex_):
    """The brick kernel stays the fallback (FD radius != 6, complex cell_typ 15): keep it pinned to the oracle."""
    g, veff, proj, x = small_case(cell_typ, (0, 0, 0), complex_=complex_)
    _setup(ctx_brick, g, veff, proj, KVEC)
    Hx = np.empty_like(x)
    ctx_brick.Hamiltonian_vectors_mult(0.25, x, Hx)
    assert ctx_brick.stats()["last_path"] == 0
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, 0.25, x, kvec=KVEC)) < TOL


@pytest.mark.parametrize("cell_typ", [0, 12, 15, 17])
@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("N", [(45, 21, 19), (70, 35, 14)])
def test_zmarch_kernel_multi_tile(ctx_zmarch, port, cell_typ, complex_, N):
    """z-march kernel on grids with several (ragged) tiles per plane, all three kinds of mixed-derivative
    components (x-, y- and z-extended), real and complex."""
    ctx = ctx_zmarch
    g, veff, proj, x = small_case(cell_typ, (0, 0, 0), N=N, L=tuple(0.5 * n for n in N), complex_=complex_)
    _setup(ctx, g, veff, proj, KVEC)
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(0.25, x, Hx)
    # complex cell_typ 15 does not fit the z-march kernel's shared memory and takes the brick kernel
    assert ctx.stats()["last_path"] == (0 if (cell_typ == 15 and complex_) else 2)
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, 0.25, x, kvec=KVEC)) < TOL


def test_host_pipeline_ramped_chunks(port):
    """Host-buffer entry point on a block big enough for the chunk pipeline (three buffer trios, chunk schedule
    8, 16, 32, .., 16, 8), X copy-back on: every column must come back filtered exactly once."""
    import os
    from sparc_b200.chefsi import ChefsiContext
    os.environ["CHEFSI_B200_HOST_CHUNK"] = "32"
    try:
        c = ChefsiContext(0)
        g = P.make_grid((64, 64, 64), (28.8, 28.8, 28.8))
        veff = P.synthetic_veff(g)
        x = P.random_columns(g.Nd, 200, seed=9)
        _setup(c, g, veff, None)
        a, b, a0 = P.chebyshev_bounds(g)
        X, Y = x.copy(), np.empty_like(x)
        c.ChebyshevFiltering(X, Y, 3, a, b, a0)
        Xw, Yw = port.chebyshev_filter(g, None, veff, x, 3, a, b, a0)
        assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL
        c.close()
    finally:
        del os.environ["CHEFSI_B200_HOST_CHUNK"]


# ---------------------------------------------------------------- full-size properties (160^3)
def test_full_size_plane_wave_and_linearity(ctx):
    """BASELINE.json's grid (160^3, FD order 12): a plane wave is an eigenvector of H when Veff is
    constant, so the filter must return the scalar Chebyshev recurrence times the input; plus
    linearity filter(x + 2y) = filter(x) + 2 filter(y) on random columns with the real Veff."""
    import torch
    N = 160
    g = P.make_grid((N, N, N), (45.9, 45.9, 45.9))
    ctx.set_grid(g)
    ctx.set_projectors(None)
    ctx.set_kpoint((0, 0, 0))
    ax = 2 * np.pi * np.arange(N) / N
    pw = np.cos(3 * ax[None, None, :] + 5 * ax[None, :, None] + 2 * ax[:, None, None]).reshape(1, -1)
    lam = 0.0
    for d, (name, mm) in enumerate((("D2_x", 3), ("D2_y", 5), ("D2_z", 2))):
        w = g.coefs[name]
        lam += w[0] + sum(2 * w[p] * np.cos(p * 2 * np.pi * mm / N) for p in range(1, 7))
    lam = -0.5 * lam - 0.37
    ctx.set_veff(np.full(g.Nd, -0.37))
    a, b, a0 = P.chebyshev_bounds(g)
    m = 20
    e, c = 0.5 * (b - a), 0.5 * (b + a)
    sigma = sigma1 = e / (a0 - c)
    gamma = 2.0 / sigma1
    t_prev, t = 1.0, (sigma1 / e) * (lam - c)
    for _ in range(1, m):
        sigma2 = 1.0 / (gamma - sigma)
        t_prev, t = t, (2 * sigma2 / e) * (lam - c) * t - sigma * sigma2 * t_prev
        sigma = sigma2
    X = pw.copy()
    Y = np.empty_like(X)
    ctx.ChebyshevFiltering(X, Y, m, a, b, a0)
    assert ctx.stats()["last_path"] == 1
    assert rel_fro(Y, t * pw) < TOL and rel_fro(X, t_prev * pw) < TOL

    ctx.set_veff(P.synthetic_veff(g))
    r = P.random_columns(g.Nd, 2, seed=4)
    Xs = np.ascontiguousarray(np.stack([r[0], r[1], r[0] + 2.0 * r[1]]))
    Ys = np.empty_like(Xs)
    ctx.ChebyshevFiltering(Xs, Ys, m, a, b, a0)
    assert rel_fro(Ys[2], Ys[0] + 2.0 * Ys[1]) < TOL


def test_bench_problem_parity(ctx, port):
    """The exact bench.py workload (BASELINE.json configs[4]: 160^3, 864 Al atoms x 18 projectors = whole,
    unsegmented sphere images, NP = 20 FUSED projector mode, degree 20) filtered inside ONE 128-column launch
    group (> 148 work items: persistent CTAs, round barrier) through the device-resident entry point bench.py
    times; three of its columns are compared with the reference's own ChebyshevFiltering (the compiled reference
    when its prebuilt library travelled, else the C restatement)."""
    import argparse
    import torch
    import bench
    from oracle.bindings import Reference, reference_available
    args = argparse.Namespace(grid=160, cell_typ=0, no_nloc=False, ncell=6)
    g, veff, proj, (a, b, a0) = bench.build_problem(args)
    assert proj.n_atom == 864 and proj.n_img >= 864
    _setup(ctx, g, veff, proj)
    ld, ncol, m = ctx.device_ld, 128, 20
    first = bench.CPU_FIRST_COL
    bufs = [torch.empty(ncol * ld, dtype=torch.float64, device="cuda") for _ in range(3)]
    ctx.fill_random_device(bufs[0], ncol, first_col=first, seed=1)
    ys, xs = ctx.filter_device(bufs[0], bufs[1], bufs[2], ncol, m, a, b, a0)
    ctx.synchronize()
    st = ctx.stats()
    assert st["last_path"] == 1 and st["last_nloc_atomic"] == 0
    cols = [0, 77, 127]
    x = np.concatenate([P.random_columns(g.Nd, 1, first_col=first + k, seed=1) for k in cols])
    if reference_available():
        Xw, Yw = Reference(g, proj, veff).chebyshev_filter(x, m, a, b, a0)
    else:
        Xw, Yw = port.chebyshev_filter(g, proj, veff, x, m, a, b, a0)
    for q, k in enumerate(cols):
        y = bufs[ys][k * ld:k * ld + g.Nd].cpu().numpy()
        xo = bufs[xs][k * ld:k * ld + g.Nd].cpu().numpy()
        assert rel_fro(y, Yw[q]) < TOL and rel_fro(xo, Xw[q]) < TOL
    del bufs
    torch.cuda.empty_cache()


def test_empty_block_and_degenerate_projector_tables(ctx, port):
    """Edge cases the reference's loops pass through silently: a block with no columns (every entry point returns
    without touching its output), an atom type without projectors (nlocVecRoutines.c:811 `if (!nproj) continue`), a
    sphere image with no grid points in the domain (`if (!ndc) continue`, :816), degree 1."""
    g = P.make_grid((32, 32, 16), (14.4, 14.4, 7.2))
    veff = P.synthetic_veff(g)
    pr = P.make_projectors(g, np.array([[0.2, 0.3, 0.4], [0.7, 0.7, 0.6], [0.5, 0.1, 0.2]]), rc=[2.0, 2.2, 1.8],
                           nproj=[0, 7, 5], seed=3)
    assert pr.IP_displ[1] == 0 and sphere_overlap_count(pr, g.Nd) == 0
    proj = P.Projectors(n_atom=pr.n_atom, IP_displ=pr.IP_displ, gamma=pr.gamma,
                        img_atom=np.append(pr.img_atom, 1).astype(np.int32), img_ndc=np.append(pr.img_ndc, 0).astype(np.int32),
                        img_coords=np.append(pr.img_coords, [99.0, 99.0, 99.0]),
                        pos_off=np.append(pr.pos_off, pr.pos_off[-1]).astype(np.int64),
                        chi_off=np.append(pr.chi_off, pr.chi_off[-1]).astype(np.int64), grid_pos=pr.grid_pos, chi=pr.chi)
    _setup(ctx, g, veff, proj)
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    x = P.random_columns(g.Nd, 3, seed=23)
    for m in (1, 4):
        X, Y = x.copy(), np.empty_like(x)
        ctx.ChebyshevFiltering(X, Y, m, a, b, a0)
        Xw, Yw = port.chebyshev_filter(g, proj, veff, x, m, a, b, a0)
        assert rel_fro(Y, Yw) < TOL and rel_fro(X, Xw) < TOL
    Hx = np.empty_like(x)
    ctx.Hamiltonian_vectors_mult(0.2, x, Hx)
    assert rel_fro(Hx, port.hamiltonian_mult(g, proj, veff, 0.2, x)) < TOL
    # no columns
    launches = ctx.stats()["kernel_launches"]
    e_in, e_out = np.empty((0, g.Nd)), np.empty((0, g.Nd))
    ctx.ChebyshevFiltering(e_in, e_out, 4, a, b, a0)
    ctx.Hamiltonian_vectors_mult(0.2, e_in, e_out)
    ctx.Lap_vec_mult(0.0, e_in, e_out)
    ctx.Gradient_vectors_dir(0.0, e_in, e_out, 1)
    assert ctx.stats()["kernel_launches"] == launches


def test_device_resident_entry_point_and_rng(ctx, port):
    import torch
    g, veff, proj, _ = small_case(0, N=(32, 32, 16), L=(14.0, 14.0, 7.0))
    _setup(ctx, g, veff, proj)
    ld = ctx.device_ld
    ncol = 4
    bufs = [torch.zeros(ncol * ld, dtype=torch.float64, device="cuda") for _ in range(3)]
    dense = torch.zeros((ncol, g.Nd), dtype=torch.float64, device="cuda")

    def unpack(buf):
        ctx.unpack_device(buf, dense, g.Nd, ncol)
        ctx.synchronize()
        return dense.cpu().numpy()

    ctx.fill_random_device(bufs[0], ncol, first_col=7, seed=3)
    x = unpack(bufs[0])
    assert np.array_equal(x, P.random_columns(g.Nd, ncol, first_col=7, seed=3))
    a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
    ys, xs = ctx.filter_device(bufs[0], bufs[1], bufs[2], ncol, 7, a, b, a0)
    ctx.synchronize()
    Xw, Yw = port.chebyshev_filter(g, proj, veff, x, 7, a, b, a0)
    assert rel_fro(unpack(bufs[ys]), Yw) < TOL
    assert rel_fro(unpack(bufs[xs]), Xw) < TOL
    # pack is the inverse of unpack
    ctx.pack_device(dense, g.Nd, bufs[2], ncol)
    assert np.array_equal(unpack(bufs[2]), dense.cpu().numpy())
