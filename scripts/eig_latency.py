"""Subspace eigenproblem Hp q = lambda Mp q: chefsi_subspace_eig (cuSOLVER Dsygvd on the device, matrices uploaded per
call) against LAPACK dsygvd on the host (scipy, the routine the reference calls), per size.  Decides the default of
CHEFSI_B200_EIG_MIN_N in sparc_shim.c.  Run on a GPU box: python scripts/eig_latency.py"""
import os
import sys
import time

import numpy as np
import scipy.linalg

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_b200 import problem as P  # noqa: E402
from sparc_b200.chefsi import ChefsiContext  # noqa: E402


def pencil(n, cplx, seed=5):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
    B = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
    return np.ascontiguousarray(((A + A.conj().T) / 2).T), np.ascontiguousarray((B @ B.conj().T / n + np.eye(n)).T)


ctx = ChefsiContext(0)
ctx.set_grid(P.make_grid((16, 16, 16), (8.0, 8.0, 8.0)))
for cplx in (False, True):
    for n in (9, 30, 64, 128, 256, 512, 1024, 2048):
        Hp, Mp = pencil(n, cplx)
        reps = 20 if n <= 256 else 3
        ctx.DP_Solve_Generalized_EigenProblem(n, Hp, Mp)
        t0 = time.perf_counter()
        for _ in range(reps):
            lam, Q = ctx.DP_Solve_Generalized_EigenProblem(n, Hp, Mp)
        t_gpu = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            lam_h, _ = scipy.linalg.eigh(Hp.T, Mp.T, driver="gvd")
        t_cpu = (time.perf_counter() - t0) / reps
        print(f"{'complex' if cplx else 'real':8s} n={n:5d}  device {1e3 * t_gpu:9.3f} ms   host LAPACK ({os.environ.get('OMP_NUM_THREADS', 'all')} threads) "
              f"{1e3 * t_cpu:9.3f} ms   max |dlambda| {np.abs(lam - lam_h).max():.1e}", flush=True)
ctx.close()
