"""Golden vectors from REAL SPARC runs ("debug hook dump (X_in, Veff, bounds, Y_out) at
ChebyshevFiltering entry/exit").

Runs the unmodified reference program (integration/_build/sparc_b200 with every call forwarded to the reference's
own routines, CHEFSI_B200_DISABLE=1 -- no GPU involved) on the reference's own test systems and lets the shim's dump
hook (sparc_shim.c, CHEFSI_B200_DUMP_DIR) write, for one mid-SCF ChebyshevFiltering[_kpt] call, the flattened inputs
exactly as the CUDA library would receive them -- the real psp8/spline Chi tables with their overlapping spheres,
the real Gamma, the SCF's Veff and eigenvalue bounds, the current orbitals -- and the outputs of the REFERENCE
routine on them.  The first few columns are kept (the filter is column-independent).

    python tests/golden/make_sparc_dumps.py            # needs /root/reference (dev container)
"""
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BUILD = os.path.join(ROOT, "integration", "_build")
OUT = os.path.dirname(os.path.abspath(__file__))

# system -> (dump call index, columns kept)
SYSTEMS = {"Si8": (6, 3), "BaTiO3": (6, 2), "Si8_kpt": (9, 2)}
_KINDS = {0: np.int32, 1: np.int64, 2: np.float64}


def read_blob(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            head = f.read(32)
            if len(head) < 32:
                break
            name = head.split(b"\0", 1)[0].decode()
            kind, = struct.unpack("i", f.read(4))
            count, = struct.unpack("q", f.read(8))
            out[name] = np.frombuffer(f.read(count * np.dtype(_KINDS[kind]).itemsize), dtype=_KINDS[kind]).copy()
    return out


def main():
    only = set(sys.argv[1:])
    for name, (call, ncol) in SYSTEMS.items():
        if only and name not in only:
            continue
        work = tempfile.mkdtemp()
        shutil.copytree(os.path.join(BUILD, "cases"), os.path.join(work, "cases"))
        cwd = os.path.join(work, "cases", "tests", name, "standard")
        env = dict(os.environ, CHEFSI_B200_DISABLE="1", CHEFSI_B200_DUMP_DIR=work, CHEFSI_B200_DUMP_CALL=str(call),
                   CHEFSI_B200_DUMP_NCOL=str(ncol), CHEFSI_B200_DUMP_EXIT="1", OMP_NUM_THREADS="1")
        subprocess.run([os.path.join(BUILD, "sparc_b200"), "-name", name], cwd=cwd, env=env, check=True,
                       stdout=subprocess.DEVNULL)
        kpt = name.endswith("_kpt")
        d = read_blob(os.path.join(work, "filter_call_kpt.bin" if kpt else "filter_call.bin"))
        Nx, Ny, Nz, BCx, BCy, BCz, FDn, cell_typ, is_kpt, m, nd, ncol_all, n_atom, n_img = d["ints"]
        dV, Lx, Ly, Lz, a, b, a0, k1, k2, k3 = d["doubles"]
        Nd = Nx * Ny * Nz

        def cols(prefix):
            arr = np.stack([d[f"{prefix}_{n}"] for n in range(nd)])
            return arr.view(np.complex128) if is_kpt else arr

        X0, Xo, Yo = cols("X0"), cols("Xout"), cols("Yout")
        assert X0.shape == (nd, Nd)
        np.savez_compressed(
            os.path.join(OUT, f"sparc_{name.lower()}.npz"),
            N=np.array([Nx, Ny, Nz]), BC=np.array([BCx, BCy, BCz]), L=np.array([Lx, Ly, Lz]), FDn=FDn, cell_typ=cell_typ,
            dV=dV, coefs=d["coefs"].reshape(15, -1), veff=d["veff"], kvec=np.array([k1, k2, k3]),
            bounds=np.array([a, b, a0]), m=m, IP_displ=d["IP_displ"], gamma=d["gamma"], img_atom=d["img_atom"],
            img_ndc=d["img_ndc"], img_coords=d["img_coords"], pos_off=d["pos_off"], chi_off=d["chi_off"],
            grid_pos=d["grid_pos"], chi=d["chi"], X0=X0, X_out=Xo, Y_out=Yo,
            provenance=f"{name}: ChebyshevFiltering{'_kpt' if kpt else ''} call #{call} of the reference SCF run "
                       f"(tests/{name}/standard), {nd} of {ncol_all} columns")
        overlap = int((np.bincount(d["grid_pos"], minlength=Nd) > 1).sum())
        print(f"{name}: {Nx}x{Ny}x{Nz} cell_typ {cell_typ} m={m} atoms {n_atom} images {n_img} sphere points "
              f"{len(d['grid_pos'])} overlapping grid points {overlap} |Y| {np.linalg.norm(Yo):.6e} "
              f"-> {os.path.getsize(os.path.join(OUT, f'sparc_{name.lower()}.npz')) / 1e6:.1f} MB")
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
