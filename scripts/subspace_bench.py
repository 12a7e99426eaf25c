"""Throughput of the Rayleigh-Ritz steps on the resident block (chefsi_subspace_project / _rotate): FP64 DMMA GEMMs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparc_b200 import problem as P
from sparc_b200.chefsi import ChefsiContext

CASES = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(96, 512), (128, 256), (64, 1024)]
for n, ncol in CASES:
    L = 45.9 * n / 160.0
    g = P.make_grid((n, n, n), (L, L, L))
    ctx = ChefsiContext(0)
    ctx.set_grid(g); ctx.set_veff(P.synthetic_veff(g)); ctx.set_projectors(None)
    y = torch.empty((ncol, g.Nd), dtype=torch.float64).pin_memory()
    y.numpy()[:] = P.random_columns(g.Nd, ncol, seed=2)
    Hp, Mp = np.zeros((ncol, ncol)), np.zeros((ncol, ncol))
    Q = np.ascontiguousarray(np.random.default_rng(0).standard_normal((ncol, ncol)))
    X = torch.empty((ncol, g.Nd), dtype=torch.float64).pin_memory()
    ctx.subspace_reserve(ncol)
    ctx.set_profiling(False)
    for rep in range(2):
        t0 = time.perf_counter(); ctx.DP_Project_Hamiltonian(y, Hp, Mp); t1 = time.perf_counter()
        ctx.DP_Subspace_Rotation(Q, X); t2 = time.perf_counter()
    fl = 2.0 * g.Nd * ncol * ncol
    ref = y.numpy()[:8] @ y.numpy().T
    err = np.abs(Mp[:8] - ref).max() / np.abs(ref).max()
    print(f"{n}^3 x {ncol}: project (H2D Y + H Y + 2 GEMMs) {1e3*(t1-t0):.1f} ms, rotate (GEMM + D2H) {1e3*(t2-t1):.1f} ms; "
          f"GEMM flops each {fl:.2e}; Mp check {err:.1e}", flush=True)
    # device-only timing of the GEMMs through repeated project calls with Y resident (same host address)
    t0 = time.perf_counter()
    for _ in range(3): ctx.DP_Project_Hamiltonian(y, Hp, Mp)
    dt = (time.perf_counter() - t0) / 3
    print(f"    resident project {1e3*dt:.1f} ms -> >= {2*fl/dt/1e12:.1f} TFLOP/s over the two A^T B products (incl. the H apply)")
    ctx.close()
