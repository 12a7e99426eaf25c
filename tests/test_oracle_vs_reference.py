"""Pins the C restatement (oracle/chefsi_oracle.c) to the reference's own compiled routines.

The reference functions are the UNMODIFIED sources under /root/reference/src, built by
oracle/Makefile into oracle/_ref/ (they travel to the GPU box as prebuilt files).  Every cell
type the filter supports, both boundary conditions, real and complex data, with and without
nonlocal projectors are covered; agreement is to rounding (1e-13), far inside the 1e-10 the
CUDA path is held to.
"""
import numpy as np
import pytest

from tests.cases import BOUNDS, KVEC, OVERLAP_CASES, overlap_case, rel_fro, small_case, sphere_overlap_count

CELL_TYPES = [0, 11, 12, 13, 14, 15, 16, 17]
BCS = [(0, 0, 0), (0, 1, 0), (1, 1, 1)]
TOL = 1e-13


@pytest.fixture(scope="module")
def ref_cls(have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle.bindings import Reference
    return Reference


@pytest.mark.parametrize("cell_typ", CELL_TYPES)
@pytest.mark.parametrize("BC", BCS)
def test_hamiltonian_real(port, ref_cls, cell_typ, BC):
    g, veff, proj, x = small_case(cell_typ, BC)
    ref = ref_cls(g, proj, veff)
    want = ref.hamiltonian_mult(-0.3, x)
    got = port.hamiltonian_mult(g, proj, veff, -0.3, x)
    assert rel_fro(got, want) < TOL


@pytest.mark.parametrize("cell_typ", CELL_TYPES)
@pytest.mark.parametrize("BC", BCS)
def test_hamiltonian_complex(port, ref_cls, cell_typ, BC):
    g, veff, proj, x = small_case(cell_typ, BC, complex_=True)
    ref = ref_cls(g, proj, veff, kvec=KVEC)
    want = ref.hamiltonian_mult(0.2, x)
    got = port.hamiltonian_mult(g, proj, veff, 0.2, x, kvec=KVEC)
    assert rel_fro(got, want) < TOL


@pytest.mark.parametrize("cell_typ", [0, 13, 17])
def test_laplacian_only_no_potential(port, ref_cls, cell_typ):
    """b = 0 / v = NULL branch (lapVecRoutines.c:321-322), arbitrary a and c."""
    g, veff, proj, x = small_case(cell_typ, with_proj=False)
    ref = ref_cls(g, None, veff)
    want = ref.lap_plus_diag(1.0, 0.0, 0.7, False, x)
    got = port.lap_plus_diag(g, 1.0, 0.0, 0.7, None, x)
    assert rel_fro(got, want) < TOL


@pytest.mark.parametrize("cell_typ", [0, 17])
@pytest.mark.parametrize("complex_", [False, True])
def test_vnl_only(port, ref_cls, cell_typ, complex_):
    g, veff, proj, x = small_case(cell_typ, complex_=complex_)
    ref = ref_cls(g, proj, veff, kvec=KVEC)
    want = ref.vnl_mult(x, np.zeros_like(x))
    got = port.vnl_mult(g, proj, x, np.zeros_like(x), kvec=KVEC)
    assert np.linalg.norm(want) > 0
    assert rel_fro(got, want) < TOL


@pytest.mark.parametrize("cell_typ", CELL_TYPES)
@pytest.mark.parametrize("complex_", [False, True])
def test_chebyshev_filter(port, ref_cls, cell_typ, complex_):
    g, veff, proj, x = small_case(cell_typ, complex_=complex_, ncol=2)
    a, b, a0 = BOUNDS
    ref = ref_cls(g, proj, veff, kvec=KVEC)
    Xr, Yr = ref.chebyshev_filter(x, 9, a, b, a0)
    Xp, Yp = port.chebyshev_filter(g, proj, veff, x, 9, a, b, a0, kvec=KVEC)
    assert rel_fro(Yp, Yr) < 1e-12
    assert rel_fro(Xp, Xr) < 1e-12


def test_fd_radius_other_than_six(port, ref_cls):
    """generic-radius branch of the reference (stencil_3axis_thread_variable_radius)."""
    for cell_typ in (0, 17):
        g, veff, proj, x = small_case(cell_typ, FDn=4)
        ref = ref_cls(g, proj, veff)
        assert rel_fro(port.hamiltonian_mult(g, proj, veff, 0.1, x), ref.hamiltonian_mult(0.1, x)) < TOL


@pytest.mark.parametrize("name", sorted(OVERLAP_CASES))
@pytest.mark.parametrize("complex_", [False, True])
def test_overlapping_spheres(port, ref_cls, name, complex_):
    """Overlapping rc-spheres: atoms closer than rc1 + rc2 and atoms whose own periodic images overlap.  The
    reference accumulates the images of an atom into one alpha block with beta = 1 (nlocVecRoutines.c:821-827) and
    scatter-adds every image's Chi alpha onto Hx (:866-881), so a grid point in two spheres receives both."""
    g, veff, proj, x = overlap_case(name, complex_=complex_, ncol=2)
    assert sphere_overlap_count(proj, g.Nd) > 0
    kvec = tuple(kk if bc == 0 else 0.0 for kk, bc in zip(KVEC, g.BC))
    ref = ref_cls(g, proj, veff, kvec=kvec)
    want = ref.vnl_mult(x, np.zeros_like(x))
    got = port.vnl_mult(g, proj, x, np.zeros_like(x), kvec=kvec)
    assert rel_fro(got, want) < TOL
    a, b, a0 = 0.5, (1.01 * g.max_eig_mhalf_lap() + 0.5) if g.cell_typ == 0 else 40.0, -0.6
    Xr, Yr = ref.chebyshev_filter(x, 7, a, b, a0)
    Xp, Yp = port.chebyshev_filter(g, proj, veff, x, 7, a, b, a0, kvec=kvec)
    assert rel_fro(Yp, Yr) < 1e-12 and rel_fro(Xp, Xr) < 1e-12
