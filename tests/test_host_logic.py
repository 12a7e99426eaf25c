"""Host-side logic that needs neither a GPU nor the reference: SPARC's FD tables, cell-type
classification, the counter-based start-vector generator, band partitioning."""
import numpy as np
import pytest

from sparc_b200 import problem as P
from sparc_b200.partition import band_partition


def test_fd_weights_order12():
    w1, w2 = P.fd_weights(6)
    # classic 12th-order central weights
    assert w2[0] == pytest.approx(-5369.0 / 1800.0, rel=1e-14)
    assert w2[1] == pytest.approx(12.0 / 7.0, rel=1e-14)
    assert w2[6] == pytest.approx(-1.0 / 16632.0, rel=1e-13)
    assert w1[1] == pytest.approx(6.0 / 7.0, rel=1e-14)
    assert abs(w2[0] + 2 * w2[1:].sum()) < 1e-14  # constants are annihilated


@pytest.mark.parametrize("ct", [0, 11, 12, 13, 14, 15, 16, 17])
def test_cell_type_classification(ct):
    g = P.make_grid((12, 12, 12), (6, 6, 6), latvec=P.LATVEC_BY_CELL_TYP[ct])
    assert g.cell_typ == ct


def test_si8_lattice_is_type_17():
    """tests/Si8/standard/Si8.inpt:3-6 (Si8 is triclinic, cell_typ 17)."""
    assert P.lattice_transforms(P.SI8_LATVEC)[4] == 17


def test_random_columns_match_c_generator(port):
    a = P.random_columns(1000, 3, first_col=5, seed=9)
    b = port.fill_random(1000, 3, first_col=5, seed=9)
    assert np.array_equal(a, b)
    assert a.min() >= -0.5 and a.max() < 0.5 and abs(a.mean()) < 0.02
    # column blocks are reproducible independently of how they are split
    c = P.random_columns(1000, 1, first_col=6, seed=9)
    assert np.array_equal(a[1], c[0])


def test_band_partition_matches_sparc_npband_rule():
    """parallelization.c:403-428: NB = ceil(Ns/npband); rank r owns [r*NB, min((r+1)NB, Ns))."""
    for ns, p in [(4096, 8), (30, 6), (29, 8), (9, 4), (5, 8), (27, 9)]:
        parts = [band_partition(ns, p, r) for r in range(p)]
        nb = -(-ns // p)
        covered = []
        for r, (s, n) in enumerate(parts):
            assert s == min(r * nb, ns)
            assert n == max(0, min((r + 1) * nb, ns) - r * nb)
            covered += list(range(s, s + n))
        assert covered == list(range(ns))


def test_projector_tables_are_consistent():
    g = P.make_grid((14, 13, 15), (7.0, 6.5, 7.5), latvec=P.SI8_LATVEC)
    pr = P.make_projectors(g, np.array([[0.05, 0.5, 0.5]]), rc=2.2, nproj=4)
    assert pr.n_img >= 2  # atom near the x face: its periodic image reaches into the cell
    assert pr.pos_off[-1] == pr.grid_pos.size and pr.chi_off[-1] == pr.chi.size
    assert (pr.grid_pos >= 0).all() and (pr.grid_pos < g.Nd).all()


def test_shared_hermitian_projection_parts_cover_every_element_pair_once_and_mirror_back():
    """Band-parallel projection with every Hermitian element pair formed once (sparc_b200/band_parallel.py, the rule of
    rank_block_part in chefsi_internal.h): of each mirrored pair of off-diagonal elements exactly one is formed, the
    per-rank share is balanced (antipodal blocks of an even rank count are split half and half), and assemble_hermitian
    rebuilds the full matrix from the zero-padded column blocks."""
    import ctypes as C
    import numpy as np
    from sparc_b200 import capi
    from sparc_b200.band_parallel import assemble_hermitian, rank_block_part
    lib = capi.load_library()
    for P, ncJ, ncI in ((2, 5, 4), (4, 3, 3), (5, 2, 7), (8, 6, 5)):           # the host restatement == the library's rule
        for I in range(P):
            for J in range(P):
                out = [C.c_int() for _ in range(4)]
                lib.chefsi_rank_block_part(J, I, P, ncJ, ncI, *[C.byref(v) for v in out])
                got, want = tuple(v.value for v in out), rank_block_part(J, I, P, ncJ, ncI)
                assert got == want or (max(got[1] - got[0], 0) * max(got[3] - got[2], 0) == 0 and (want[1] - want[0]) * (want[3] - want[2]) == 0)
    rng = np.random.default_rng(3)
    cases = [([3, 2], False), ([2, 0, 3, 1], True), ([1, 1, 1, 1, 1], True), ([4, 3, 5], False), ([5, 4, 4, 5, 3, 4, 5, 4], False),
             ([7, 6], True), ([1, 1], False)]
    for P in range(1, 10):
        cases.append(([4] * P, False))
    for ncols, cplx in cases:
        ns, P = sum(ncols), len(ncols)
        off = np.concatenate([[0], np.cumsum(ncols)]).astype(int)
        formed = np.zeros((ns, ns), dtype=int)                                  # [row, col]
        work = []
        for I in range(P):
            w = 0.0
            for J in range(P):
                r0, r1, c0, c1 = rank_block_part(J, I, P, ncols[J], ncols[I])
                if r1 > r0 and c1 > c0:
                    formed[off[J] + r0:off[J] + r1, off[I] + c0:off[I] + c1] += 1
                    w += (r1 - r0) * (c1 - c0) * (0.5 if J == I else 1.0)
            work.append(w)
        offdiag = ~np.zeros((ns, ns), dtype=bool)
        for I in range(P):
            offdiag[off[I]:off[I + 1], off[I]:off[I + 1]] = False
        assert np.all((formed + formed.T)[offdiag] == 1)                        # exactly one of each mirrored pair
        if len(set(ncols)) == 1 and P > 1:
            assert max(work) - min(work) <= ncols[0] + 1e-9     # balanced share (an odd column count splits into h and h + 1 columns)
        A = rng.standard_normal((ns, ns)) + (1j * rng.standard_normal((ns, ns)) if cplx else 0)
        H = A + A.conj().T                                                      # H[row, col]
        G = np.ascontiguousarray(H.T)                                           # column-major: G[col, row]
        blocks = []
        for I in range(P):
            b = np.zeros_like(G[off[I]:off[I + 1]])
            for J in range(P):
                r0, r1, c0, c1 = rank_block_part(J, I, P, ncols[J], ncols[I])
                b[c0:c1, off[J] + r0:off[J] + r1] = G[off[I] + c0:off[I] + c1, off[J] + r0:off[J] + r1]
            blocks.append(b)
        assert np.array_equal(assemble_hermitian(blocks, ncols), G)
