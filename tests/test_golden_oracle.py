"""The C restatement against (i) the committed reference-generated vectors in tests/golden/
(made by tests/golden/make_golden.py from the unmodified reference routines) and (ii) analytic
known answers (plane waves, Bloch waves, Chebyshev polynomial of an eigenvector).
Runs without /root/reference and without a GPU."""
import numpy as np
import pytest

from sparc_b200 import problem as P
from tests.cases import GOLDEN, SPARC_GOLDEN, load_golden, rel_fro


@pytest.mark.parametrize("name", GOLDEN)
def test_port_reproduces_reference_vectors(port, name):
    g, veff, proj, d = load_golden(name)
    k = tuple(d["kvec"])
    a, b, a0 = d["bounds"]
    Hx = port.hamiltonian_mult(g, proj, veff, float(d["c_shift"]), d["X0"], kvec=k)
    assert rel_fro(Hx, d["Hx"]) < 1e-13
    Xo, Yo = port.chebyshev_filter(g, proj, veff, d["X0"], int(d["m"]), a, b, a0, kvec=k)
    assert rel_fro(Yo, d["Y_out"]) < 1e-12
    assert rel_fro(Xo, d["X_out"]) < 1e-12


@pytest.mark.parametrize("name", SPARC_GOLDEN)
def test_port_reproduces_real_sparc_filter_calls(port, name):
    """One mid-SCF ChebyshevFiltering[_kpt] call of the reference's own Si8 / BaTiO3 / Si8_kpt runs (dumped at the
    function's entry and exit): real pseudopotential projectors (BaTiO3: overlapping spheres)."""
    g, veff, proj, d = load_golden(name)
    a, b, a0 = d["bounds"]
    Xo, Yo = port.chebyshev_filter(g, proj, veff, d["X0"], int(d["m"]), a, b, a0, kvec=tuple(d["kvec"]))
    assert rel_fro(Yo, d["Y_out"]) < 1e-11
    assert rel_fro(Xo, d["X_out"]) < 1e-11


def _plane_wave(g, mvec, kfrac=(0, 0, 0)):
    i, j, k = np.meshgrid(*[np.arange(n) for n in g.N[::-1]], indexing="ij")  # z, y, x order
    k_, j_, i_ = i, j, k
    ph = 2 * np.pi * ((mvec[0] + kfrac[0]) * i_ / g.N[0] + (mvec[1] + kfrac[1]) * j_ / g.N[1]
                      + (mvec[2] + kfrac[2]) * k_ / g.N[2])
    return ph.reshape(-1)


def _lap_eig_orth(g, mvec, kfrac=(0, 0, 0)):
    lam = 0.0
    for d, name in enumerate(("D2_x", "D2_y", "D2_z")):
        w = g.coefs[name]
        th = 2 * np.pi * (mvec[d] + kfrac[d]) / g.N[d]
        lam += w[0] + sum(2 * w[p] * np.cos(p * th) for p in range(1, g.FDn + 1))
    return lam


def test_plane_wave_is_eigenvector_of_fd_laplacian(port):
    g = P.make_grid((16, 12, 20), (8.0, 6.0, 10.0))
    mvec = (2, 1, 3)
    x = np.cos(_plane_wave(g, mvec))[None, :].copy()
    y = port.lap_plus_diag(g, 1.0, 0.0, 0.0, None, x)
    assert rel_fro(y, _lap_eig_orth(g, mvec) * x) < 1e-12


def test_bloch_wave_is_eigenvector_with_phase_halos(port):
    g = P.make_grid((16, 12, 20), (8.0, 6.0, 10.0))
    mvec, kfrac = (1, 2, 1), (0.25, 0.1, -0.3)
    kvec = tuple(2 * np.pi * kfrac[d] / g.L[d] for d in range(3))
    x = np.exp(1j * _plane_wave(g, mvec, kfrac))[None, :].copy()
    y = port.lap_plus_diag(g, 1.0, 0.0, 0.0, None, x, kvec=kvec)
    assert rel_fro(y, _lap_eig_orth(g, mvec, kfrac) * x) < 1e-12


def test_nonorthogonal_plane_wave(port):
    """Lattice-coordinate plane wave: eigenvalue sum_ij T_ij d_i d_j with FD symbols."""
    g = P.make_grid((16, 16, 16), (8.0, 8.0, 8.0), latvec=P.SI8_LATVEC)
    assert g.cell_typ == 17
    mvec = (1, 2, 1)
    x = np.exp(1j * _plane_wave(g, mvec))[None, :].copy()
    w1, w2 = P.fd_weights(g.FDn)
    th = [2 * np.pi * mvec[d] / g.N[d] for d in range(3)]
    d1 = [1j * sum(2 * w1[p] * np.sin(p * th[d]) for p in range(1, g.FDn + 1)) / g.h[d] for d in range(3)]
    d2 = [(w2[0] + sum(2 * w2[p] * np.cos(p * th[d]) for p in range(1, g.FDn + 1))) / g.h[d] ** 2 for d in range(3)]
    _, _, _, T, _ = P.lattice_transforms(P.SI8_LATVEC)
    lam = sum(T[d, d] * d2[d] for d in range(3)) + 2 * T[0, 1] * d1[0] * d1[1] + 2 * T[0, 2] * d1[0] * d1[2] \
        + 2 * T[1, 2] * d1[1] * d1[2]
    y = port.lap_plus_diag(g, 1.0, 0.0, 0.0, None, x, kvec=(0, 0, 0))
    assert rel_fro(y, lam * x) < 1e-12


def test_chebyshev_of_eigenvector_is_scalar_recurrence(port):
    g = P.make_grid((12, 12, 12), (6.0, 6.0, 6.0))
    mvec = (1, 0, 2)
    x = np.cos(_plane_wave(g, mvec))[None, :].copy()
    veff = np.full(g.Nd, -0.2)
    lam = -0.5 * _lap_eig_orth(g, mvec) - 0.2
    a, b, a0, m = 0.5, 45.0, -0.6, 11
    e, c = 0.5 * (b - a), 0.5 * (b + a)
    sigma = sigma1 = e / (a0 - c)
    gamma = 2.0 / sigma1
    t_prev, t = 1.0, (sigma1 / e) * (lam - c)
    for _ in range(1, m):
        sigma2 = 1.0 / (gamma - sigma)
        t_prev, t = t, (2 * sigma2 / e) * (lam - c) * t - sigma * sigma2 * t_prev
        sigma = sigma2
    Xo, Yo = port.chebyshev_filter(g, None, veff, x, m, a, b, a0)
    assert rel_fro(Yo, t * x) < 1e-11
    assert rel_fro(Xo, t_prev * x) < 1e-11


def test_single_projector_is_rank_one_update(port):
    g = P.make_grid((12, 12, 12), (6.0, 6.0, 6.0), BC=(1, 1, 1))
    proj = P.make_projectors(g, np.array([[0.5, 0.5, 0.5]]), rc=2.0, nproj=1)
    assert proj.n_img == 1
    x = P.random_columns(g.Nd, 2, seed=5)
    out = port.vnl_mult(g, proj, x, np.zeros_like(x))
    chi = np.zeros(g.Nd)
    chi[proj.grid_pos] = proj.chi
    want = np.outer(x @ chi * g.dV * proj.gamma[0], chi)
    assert rel_fro(out, want) < 1e-13
