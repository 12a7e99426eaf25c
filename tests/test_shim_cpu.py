"""The reference-side binding on a CPU-only box: with CHEFSI_B200_DISABLE=1 every replaced routine
(ChebyshevFiltering, Hamiltonian_vectors_mult, Lap_vec_mult, AAR, Lanczos[_kpt], DP_Project_Hamiltonian, DP_Subspace_Rotation,
DP_Solve_Generalized_EigenProblem, CalculateDensity_psi)
forwards to the reference's own definition kept in the executable as *_ref, no CUDA context is created, and the SCF
energy of tests/Si8 is the reference's.  Checks the link-time substitution (integration/Makefile) without a GPU.
Skipped where integration/_build is absent (it is derived from /root/reference)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")
EXE = os.path.join(BUILD, "sparc_b200")


def test_every_replaced_symbol_has_a_ref_twin():
    if not os.path.exists(EXE):
        pytest.skip("integration/_build not present")
    syms = subprocess.run(["nm", EXE], capture_output=True, text=True).stdout
    for name in ("ChebyshevFiltering", "ChebyshevFiltering_kpt", "Hamiltonian_vectors_mult", "Hamiltonian_vectors_mult_kpt",
                 "Lap_vec_mult", "AAR", "Lanczos", "Lanczos_kpt", "DP_Project_Hamiltonian", "DP_Subspace_Rotation",
                 "DP_Project_Hamiltonian_kpt", "DP_Subspace_Rotation_kpt", "DP_Solve_Generalized_EigenProblem",
                 "DP_Solve_Generalized_EigenProblem_kpt", "CalculateDensity_psi"):
        assert re.search(rf" T {name}$", syms, re.M), name
        assert re.search(rf" T {name}_ref$", syms, re.M), name + "_ref"


def test_disabled_shim_forwards_everything_to_the_reference(tmp_path):
    src = os.path.join(BUILD, "cases")
    if not (os.path.exists(EXE) and os.path.isdir(src)):
        pytest.skip("integration/_build not present")
    work = os.path.join(str(tmp_path), "cases")
    shutil.copytree(src, work)
    cwd = os.path.join(work, "tests", "Si8", "standard")
    env = dict(os.environ, CHEFSI_B200_DISABLE="1", CHEFSI_B200_SHIM_VERBOSE="1", OMP_NUM_THREADS="1")
    r = subprocess.run([EXE, "-name", "Si8"], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1500:]
    e = float(re.findall(r"Free energy per atom\s*:\s*([-+0-9.Ee]+)", open(os.path.join(cwd, "Si8.out")).read())[-1])
    e_ref = float(re.findall(r"Free energy per atom\s*:\s*([-+0-9.Ee]+)", open(os.path.join(cwd, "Si8.refout")).read())[-1])
    assert abs(e - e_ref) <= 1e-6
    m = re.search(r"(\d+) ChebyshevFiltering calls .*?(\d+) calls forwarded to the reference \(of which (\d+) ChebyshevFiltering", r.stderr)
    assert m and int(m.group(1)) == 0 and int(m.group(3)) > 0, r.stderr[-800:]
    assert "context creation 0.000 s" in r.stderr      # no device was touched
