"""Where does a ChebyshevFiltering call on a small (SCF test system sized) problem spend its time?
BaTiO3-like: 39^3 grid, 5 atoms with 32/18/13 projectors and overlapping spheres, 29 columns, degree 36."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_b200 import problem as P
from sparc_b200.chefsi import ChefsiContext

def run(N, L, pos, rc, nproj, ncol, m, label):
    g = P.make_grid(N, L)
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array(pos), rc=rc, nproj=nproj)
    x = P.random_columns(g.Nd, ncol, seed=3)
    a, b, a0 = P.chebyshev_bounds(g)
    ctx = ChefsiContext(0)
    ctx.set_grid(g); ctx.set_veff(veff); ctx.set_projectors(proj)
    X, Y = x.copy(), np.empty_like(x)
    for _ in range(2):
        X[:] = x; ctx.ChebyshevFiltering(X, Y, m, a, b, a0)
    t0 = time.perf_counter()
    for _ in range(5):
        X[:] = x; ctx.ChebyshevFiltering(X, Y, m, a, b, a0)
    dt = (time.perf_counter() - t0) / 5
    ctx.set_profiling(True)
    X[:] = x; ctx.ChebyshevFiltering(X, Y, m, a, b, a0)
    s = ctx.stats()
    print(f"{label}: {1e3*dt:.2f} ms per call  (stencil {s['last_stencil_ms']:.2f} ms in {s['last_stencil_launches']} launches, "
          f"nloc {s['last_nloc_ms']:.2f} ms, path {s['last_path']}, kernel launches total {s['kernel_launches']})")
    ctx.close()

run((39, 39, 39), (7.63, 7.63, 7.63), [[0, 0, 0], [0.5, 0.5, 0.5], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]],
    [3.2, 2.6, 1.6, 1.6, 1.6], [32, 18, 13, 13, 13], 29, 36, "BaTiO3-like 39^3 x 29, m=36")
run((25, 25, 25), (10.26, 10.26, 10.26), [[0, 0, 0], [0.25, 0.25, 0.25], [0.5, 0.5, 0], [0.75, 0.75, 0.25], [0.5, 0, 0.5], [0.75, 0.25, 0.75], [0, 0.5, 0.5], [0.25, 0.75, 0.75]],
    [2.4] * 8, [18] * 8, 30, 21, "Si8-like 25^3 x 30, m=21")
