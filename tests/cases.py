"""Shared problem builders for the parity tests (small, seconds on a CPU)."""
import numpy as np

from sparc_b200 import problem as P

KVEC = (0.11, -0.07, 0.2)


def rel_fro(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def small_case(cell_typ=0, BC=(0, 0, 0), N=(14, 13, 15), L=(7.0, 6.5, 7.5), ncol=3, complex_=False, seed=0,
               with_proj=True, FDn=6):
    g = P.make_grid(N, L, BC=BC, FDn=FDn, latvec=P.LATVEC_BY_CELL_TYP[cell_typ])
    assert g.cell_typ == cell_typ
    veff = P.synthetic_veff(g)
    proj = None
    if with_proj:
        frac = np.array([[0.1, 0.2, 0.3], [0.6, 0.55, 0.9]])
        proj = P.make_projectors(g, frac, rc=[2.1, 1.7], nproj=[5, 9])
    x = P.random_columns(g.Nd, ncol, first_col=0, seed=seed + 1)
    if complex_:
        x = x + 1j * P.random_columns(g.Nd, ncol, first_col=1000, seed=seed + 1)
    return g, veff, proj, np.ascontiguousarray(x)


# (a, b, a0) used with the small cases: b above the spectrum of -1/2 Lap + Veff for h ~ 0.5
BOUNDS = (0.5, 40.0, -0.6)


def load_golden(name):
    """Rebuild (grid, veff, proj, data) from a tests/golden/*.npz fixture (all inputs are stored)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    ct = int(d["cell_typ"])
    g = P.make_grid(tuple(d["N"]), tuple(d["L"]), BC=tuple(d["BC"]), FDn=int(d["FDn"]), latvec=P.LATVEC_BY_CELL_TYP[ct])
    # use the STORED tables, not recomputed ones
    for i, n in enumerate(P._COEF_NAMES):
        g.coefs[n] = d["coefs"][i].copy()
    g.dV = float(d["dV"])
    g.cell_typ = ct
    proj = P.Projectors(
        n_atom=len(d["IP_displ"]) - 1, IP_displ=d["IP_displ"].astype(np.int32), gamma=d["gamma"].copy(),
        img_atom=d["img_atom"].astype(np.int32), img_ndc=d["img_ndc"].astype(np.int32),
        img_coords=d["img_coords"].copy(), pos_off=d["pos_off"].astype(np.int64),
        chi_off=d["chi_off"].astype(np.int64), grid_pos=d["grid_pos"].astype(np.int32), chi=d["chi"].copy())
    return g, d["veff"].copy(), proj, d


GOLDEN = ["orth_gamma", "orth_dirichlet_gamma", "si8lat_gamma", "orth_kpt", "si8lat_kpt", "type14_mixedbc_gamma",
          "stream_gamma", "stream_kpt"]
