"""Per-call latency of the host-buffer entry points on SCF-test-system-sized problems (single columns as Lanczos /
the Poisson residual make them, and a 30-column filter), pageable vs pinned host memory."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparc_b200 import problem as P
from sparc_b200.chefsi import ChefsiContext

def bench(fn, n=200):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3

def run(N, L, latvec, label):
    g = P.make_grid(N, L, latvec=latvec)
    veff = P.synthetic_veff(g)
    proj = P.make_projectors(g, np.array([[0, 0, 0], [0.25, 0.25, 0.25], [0.5, 0.5, 0], [0.75, 0.75, 0.25]]), rc=[2.4] * 4, nproj=[18] * 4)
    ctx = ChefsiContext(0)
    ctx.set_grid(g); ctx.set_veff(veff); ctx.set_projectors(proj)
    a, b, a0 = 0.5, 60.0, -0.6
    for kind in ("pageable", "pinned"):
        def alloc(ncol):
            t = torch.empty((ncol, g.Nd), dtype=torch.float64)
            if kind == "pinned": t = t.pin_memory()
            return t
        x1, y1 = alloc(1), alloc(1)
        x1.numpy()[:] = P.random_columns(g.Nd, 1, seed=3)
        x30, y30 = alloc(30), alloc(30)
        x30.numpy()[:] = P.random_columns(g.Nd, 30, seed=3)
        t_lap = bench(lambda: ctx.Lap_vec_mult(0.0, x1, y1))
        t_h = bench(lambda: ctx.Hamiltonian_vectors_mult(0.0, x1, y1))
        t_f = bench(lambda: ctx.ChebyshevFiltering(x30, y30, 21, a, b, a0, copy_back_x=False), n=30)
        st = ctx.stats()
        print(f"{label} {kind}: Lap_vec_mult {t_lap:.3f} ms, Hamiltonian_vectors_mult {t_h:.3f} ms, ChebyshevFiltering(30 cols, m=21) {t_f:.3f} ms "
              f"(device {st['last_filter_ms']:.3f} ms, path {st['last_path']})", flush=True)
    ctx.close()

run((25, 25, 25), (10.26, 10.26, 10.26), P.SI8_LATVEC, "Si8-like 25^3 typ17")
run((39, 39, 39), (7.63, 7.63, 7.63), None, "BaTiO3-like 39^3 orth")
run((77, 39, 39), (15.0, 7.6, 7.6), P.SI8_LATVEC, "Au-like 77x39x39 typ17")
