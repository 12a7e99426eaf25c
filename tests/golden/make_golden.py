"""Generate tests/golden/*.npz from the reference's own compiled routines.

Run in the dev container (needs oracle/_ref/, i.e. /root/reference):

    python tests/golden/make_golden.py

Each file stores the complete inputs (grid tables, Veff, projector tables, start vectors,
bounds, k-point) together with the outputs of the UNMODIFIED reference functions
Hamiltonian_vectors_mult[_kpt] and ChebyshevFiltering[_kpt] on them, so the vectors can be
checked anywhere (the GPU box has no /root/reference).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bindings import Reference  # noqa: E402
from sparc_b200 import problem as P  # noqa: E402
from tests.cases import BOUNDS, KVEC, overlap_case, small_case, sphere_overlap_count  # noqa: E402

# grids: the small one exercises the general kernels, the two "stream_*" ones are large enough for the TMA
# streaming kernels (real: >= 32 x 32 x 12, k-point: >= 16 x 32 x 12), so those are pinned to reference-made
# vectors as well and not only to the oracle restatement
SMALL = ((12, 10, 14), (6.0, 5.2, 7.0))
CASES = {
    # name: (cell_typ, BC, complex, m[, (N, L)])
    "stream_gamma": (0, (0, 0, 0), False, 6, ((32, 32, 14), (16.0, 16.0, 7.0))),
    "stream_kpt": (0, (0, 0, 0), True, 6, ((16, 32, 14), (8.0, 16.0, 7.0))),
    "orth_gamma": (0, (0, 0, 0), False, 8),
    "orth_dirichlet_gamma": (0, (1, 0, 1), False, 5),
    "si8lat_gamma": (17, (0, 0, 0), False, 8),
    "orth_kpt": (0, (0, 0, 0), True, 6),
    "si8lat_kpt": (17, (0, 0, 0), True, 6),
    "type14_mixedbc_gamma": (14, (0, 1, 0), False, 5),
    # overlapping rc-spheres (tests.cases.OVERLAP_CASES["stream"]): two atoms closer than rc1 + rc2, one atom whose
    # own periodic images overlap along z, 9 alpha partials on one atom; sized for the TMA streaming kernels
    "overlap_gamma": ("overlap:stream", None, False, 6),
    "overlap_kpt": ("overlap:stream", None, True, 6),
    # disjoint spheres on a multi-tile grid whose last tiles are shifted inwards (48 x 40: tiles of 32 overlap), spheres cut by
    # tile edges and by the periodic faces: the streaming kernel with its fused projector chain, pinned to reference-made vectors
    "stream_tiles_gamma": ("disjoint", (0, 0, 0), False, 5),
}

DISJOINT = dict(N=(48, 40, 14), frac=[[0.64, 0.78, 0.3], [0.2, 0.2, 0.8], [0.97, 0.03, 0.5]], rc=[2.3, 2.6, 2.0], nproj=[18, 13, 7])


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        ct, BC, cplx, m = spec[:4]
        if isinstance(ct, str) and ct.startswith("overlap:"):
            g, veff, proj, x = overlap_case(ct.split(":")[1], complex_=cplx, ncol=2, seed=3)
            assert sphere_overlap_count(proj, g.Nd) > 0
            a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
        elif ct == "disjoint":
            g = P.make_grid(DISJOINT["N"], tuple(0.45 * n for n in DISJOINT["N"]), BC=BC)
            veff = P.synthetic_veff(g)
            proj = P.make_projectors(g, np.array(DISJOINT["frac"]), rc=DISJOINT["rc"], nproj=DISJOINT["nproj"], seed=5)
            assert sphere_overlap_count(proj, g.Nd) == 0
            x = P.random_columns(g.Nd, 2, seed=19)
            a, b, a0 = 0.5, 1.01 * g.max_eig_mhalf_lap() + 0.5, -0.6
        else:
            N, L = spec[4] if len(spec) > 4 else SMALL
            g, veff, proj, x = small_case(ct, BC, N=N, L=L, ncol=2, complex_=cplx, seed=3)
            a, b, a0 = BOUNDS
        ref = Reference(g, proj, veff, kvec=KVEC)
        Hx = ref.hamiltonian_mult(-0.25, x)
        Xo, Yo = ref.chebyshev_filter(x, m, a, b, a0)
        coefs = np.stack([g.coefs[n] for n in P._COEF_NAMES])
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            N=np.array(g.N), BC=np.array(g.BC), L=np.array(g.L), FDn=g.FDn, cell_typ=g.cell_typ, dV=g.dV,
            coefs=coefs, veff=veff, kvec=np.array(KVEC), bounds=np.array([a, b, a0]), m=m, c_shift=-0.25,
            IP_displ=proj.IP_displ, gamma=proj.gamma, img_atom=proj.img_atom, img_ndc=proj.img_ndc,
            img_coords=proj.img_coords, pos_off=proj.pos_off, chi_off=proj.chi_off, grid_pos=proj.grid_pos,
            chi=proj.chi, X0=x, Hx=Hx, X_out=Xo, Y_out=Yo)
        print(name, "cell_typ", g.cell_typ, "n_img", proj.n_img, "|Y|", np.linalg.norm(Yo))


if __name__ == "__main__":
    main()
