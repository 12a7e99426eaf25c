#!/usr/bin/env python
"""bench.py -- Chebyshev-filter throughput on the BASELINE.json workload.

Workload (BASELINE.json configs[4]): synthetic Al supercell, orthogonal 160^3 grid
(h = 45.9/160 Bohr, FD order 12), 864 atoms x 18 Kleinman-Bylander projectors, 4096 states,
Chebyshev degree 20, real FP64.  A "step" is one full degree-20 ChebyshevFiltering pass over ALL
4096 columns (the columns are split over the ranks like SPARC's npband axis; the filter has no
inter-band communication, so the timed region has no collective).

    python bench.py [--gpus N --steps K --warmup W]                  # our CUDA path
    python bench.py --impl reference [--gpus N --steps K --warmup W] # the reference's own C routines on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field's definition.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "chebyshev_filter_gridpt_vectors_per_s"
UNIT = "grid-pt*vectors/s"
# DRAM bytes of one 128-column launch of the dense streaming kernel on the 160^3 grid (ncu capture, profiles/)
NCU_TRAFFIC_BYTES = 13.78e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload overrides (debugging only; the defaults are the BASELINE.json configuration)
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--ncol", type=int, default=4096)
    ap.add_argument("--degree", type=int, default=20)
    ap.add_argument("--block", type=int, default=128, help="columns filtered per launch group")
    ap.add_argument("--ncell", type=int, default=6, help="fcc conventional cells per axis (4 atoms each)")
    ap.add_argument("--no-nloc", action="store_true", help="stencil + Veff only (roofline study)")
    # SURVEY.md 8d "secondary synthetic" configurations (not the headline line): the same workload on the k-point
    # (complex, Bloch-phase halo) path and on a non-orthogonal lattice
    ap.add_argument("--kpt", action="store_true", help="complex orbitals at k = (0.25, 0.25, 0.25) * 2 pi / L")
    ap.add_argument("--cell-typ", type=int, default=0, help="lattice flavour (0 orthogonal; 11..17: problem.LATVEC_BY_CELL_TYP)")
    ap.add_argument("--no-veff", action="store_true", help="skip the local potential (experiment: cost of the Veff tile stream)")
    ap.add_argument("--e2e-cols", type=int, default=256)
    ap.add_argument("--cpu-cols-per-core", type=int, default=1)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def build_problem(args):
    from sparc_b200 import problem as P
    n = args.grid
    L = 45.9 * n / 160.0  # keep h = 0.286875 Bohr when the grid is shrunk for debugging
    g = P.make_grid((n, n, n), (L, L, L), latvec=P.LATVEC_BY_CELL_TYP[args.cell_typ] if args.cell_typ else None)
    veff = P.synthetic_veff(g)
    proj = None
    if not args.no_nloc:
        ncell = max(1, round(args.ncell * n / 160.0))
        proj = P.make_projectors(g, P.fcc_positions(ncell), rc=1.95, nproj=18, seed=7)
    a, b, a0 = P.chebyshev_bounds(g)
    return g, veff, proj, (a, b, a0)


def workload_name(args, proj):
    nat = proj.n_atom if proj is not None else 0
    kind = "complex FP64 (k-point)" if args.kpt else "real FP64"
    lat = f", cell_typ {args.cell_typ} lattice" if args.cell_typ else ""
    return (f"synthetic Al fcc supercell {args.grid}^3 grid x {args.ncol} states, Chebyshev degree {args.degree}, "
            f"FD order 12, {nat} atoms x 18 KB projectors{lat}, {kind}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w": float(np.median(pw)) if pw else None}


# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """One host process = one of the reference's band communicators (NP_BAND_PARAL = nproc,
    domain unsplit): filter `ncols` columns with the reference's ChebyshevFiltering."""
    (g, veff, proj, bounds, m, first_col, ncols, kind) = job
    from sparc_b200 import problem as P
    a, b, a0 = bounds
    x = P.random_columns(g.Nd, ncols, first_col=first_col, seed=1)
    if kind == "reference":
        from oracle.bindings import Reference
        ref = Reference(g, proj, veff)
        t0 = time.perf_counter()
        ref.chebyshev_filter(x, m, a, b, a0)
        return time.perf_counter() - t0
    from oracle.bindings import Port
    os.environ["OMP_NUM_THREADS"] = "1"
    port = Port()
    t0 = time.perf_counter()
    port.chebyshev_filter(g, proj, veff, x, m, a, b, a0)
    return time.perf_counter() - t0


def cpu_filter_rate(g, veff, proj, bounds, m, cols_per_core, reps=1):
    """Throughput of the reference CPU path on this host: N = #cores independent processes, each
    filtering `cols_per_core` columns (== the reference's NP_BAND_PARAL=N layout, which has zero
    communication inside the filter).  Returns (rate, cores, kind, sample description)."""
    import multiprocessing as mp
    from oracle.bindings import build_port, reference_available
    kind = "reference" if reference_available() else "port"
    if kind == "port":
        build_port()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = max(1, min(cores, int(os.environ.get("CHEFSI_BENCH_MAX_PROCS", "64"))))
    jobs = [(g, veff, proj, bounds, m, 100000 + r * cols_per_core, cols_per_core, kind) for r in range(cores)]
    ctx = mp.get_context("fork")
    best = None
    with ctx.Pool(cores) as pool:
        for _ in range(reps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    ncols = cores * cols_per_core
    sample = (f"{ncols} columns of the same workload ({cores} processes x {cols_per_core} column(s), "
              f"full degree-{m} filter incl. projectors), wall time incl. start-vector generation excluded")
    return g.Nd * ncols / best, cores, kind, sample, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, veff, proj, bounds = build_problem(args)
    m = args.degree
    times = []
    rate = cores = kind = sample = None
    for i in range(args.warmup + args.steps):
        rate, cores, kind, sample, dt = cpu_filter_rate(g, veff, proj, bounds, m, args.cpu_cols_per_core)
        if i >= args.warmup:
            times.append(dt)
        if i == 0 and dt * (args.warmup + args.steps) > 600:  # keep the whole run within minutes
            args.warmup, args.steps = 0, 1
            times = [dt]
            break
    ncols = cores * args.cpu_cols_per_core
    t = float(np.mean(times))
    value = g.Nd * ncols / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, proj), "step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sparc_b200.chefsi import ChefsiContext
    from sparc_b200.partition import band_partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    m = args.degree
    # rank 0 builds the problem; Veff and the projector tables are replicated with an NCCL broadcast
    # (mirrors Transfer_Veff_loc's MPI_Bcast, electronicGroundState.c:1313-1385)
    g, veff, proj, bounds = build_problem(args)
    if world > 1:
        from sparc_b200.replicate import broadcast_problem
        veff, proj = broadcast_problem(g, veff, proj, src=0, device=torch.device("cuda", local_rank))
    a, b, a0 = bounds

    ctx = ChefsiContext(local_rank)
    ctx.set_grid(g)
    if not args.no_veff:
        ctx.set_veff(veff)
    ctx.set_projectors(proj)
    cplx = bool(args.kpt)
    words = 2 if cplx else 1
    if cplx:
        ctx.set_kpoint(tuple(0.25 * 2 * np.pi / Lk for Lk in g.L))
    ld = ctx.device_ld
    first_col, ncol_local = band_partition(args.ncol, world, rank)
    block = min(args.block, max(ncol_local, 1))
    nblocks = (ncol_local + block - 1) // block
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # all local columns resident in HBM: nblocks slots + 2 spare slots that rotate through the recurrence
    slots = [torch.empty(block * ld * words, dtype=torch.float64, device="cuda") for _ in range(nblocks + 2)]
    where = list(range(nblocks))           # slot holding block i
    spare = [nblocks, nblocks + 1]
    for i in range(nblocks):
        nc = min(block, ncol_local - i * block)
        ctx.fill_random_device(slots[where[i]], nc, first_col=first_col + i * block, seed=1, is_complex=cplx)
    ctx.synchronize()

    def one_step():
        for i in range(nblocks):
            nc = min(block, ncol_local - i * block)
            trio = [where[i], spare[0], spare[1]]
            ys, xs = ctx.filter_device(slots[trio[0]], slots[trio[1]], slots[trio[2]], nc, m, a, b, a0, is_complex=cplx)
            new_where = trio[ys]
            rest = [t for t in trio if t != new_where]
            where[i], spare[0], spare[1] = new_where, rest[0], rest[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    ctx.synchronize()

    # ---- timed region: exactly K steps, device time on the launching stream, max over ranks ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.stats()["kernel_launches"]
    barrier()
    with ClockSampler(local_rank) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            one_step()
        ev1.record(stream)
        ctx.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.stats()["kernel_launches"] - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = g.Nd * args.ncol / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (fused stencil step): one more full step, back to back with the timed
    # region (same sustained clocks), with a CUDA-event pair around every kernel launch on the library's stream ----
    ctx.set_profiling(True)
    st_ms = st_n = 0.0
    nl_ms = 0.0
    alg_bytes = 0.0
    for i in range(nblocks):
        nc = min(block, ncol_local - i * block)
        trio = [where[i], spare[0], spare[1]]
        ys, xs = ctx.filter_device(slots[trio[0]], slots[trio[1]], slots[trio[2]], nc, m, a, b, a0, is_complex=cplx)
        new_where = trio[ys]
        rest = [t for t in trio if t != new_where]
        where[i], spare[0], spare[1] = new_where, rest[0], rest[1]
        s = ctx.stats()   # filter_device synchronises when profiling is on
        st_ms += s["last_stencil_ms"]
        st_n += s["last_stencil_launches"]
        nl_ms += s["last_nloc_ms"]
        alg_bytes += 8.0 * words * (3 * m - 1) * g.Nd * nc   # 16 B first step, 24 B the others (x2 complex)
    ctx.set_profiling(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = alg_bytes / (st_ms * 1e-3) / 1e9 if st_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": ("stream_kpt_kernel" if cplx else "stream_dense_kernel" if os.environ.get("CHEFSI_B200_DENSE", "1") != "0" else "stream_orth_kernel") + " (fused stencil + Veff + recurrence)" if ctx.stats()["last_path"] == 1 else ("stencil_zmarch_kernel" if ctx.stats()["last_path"] == 2 else "stencil_general_kernel"),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES if (args.grid == 160 and block == 128 and ctx.stats()["last_path"] == 1 and os.environ.get("CHEFSI_B200_DENSE", "1") != "0") else None,
        "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 128-column launch (profiles/r1_ncu_dense_final.txt: 9.61 GB read + 4.17 GB write)",
        "peak_source": peak_src, "avg_launch_ms": st_ms / st_n if st_n else None,
        "algorithmic_bytes_per_launch": 8.0 * words * (3 * m - 1) * g.Nd * block / m,
        "timing": "per-launch CUDA events over one full step run back to back with the timed steps (sustained clocks)",
        "stencil_share_of_filter": st_ms / (st_ms + nl_ms) if st_ms + nl_ms > 0 else None,
        "nloc_ms_per_degree": nl_ms / st_n if st_n else None,
    }

    # ---- e2e: the same metric through the host-buffer C-ABI call (H2D + D2H inside the timed region) ----
    e2e_cols = max(1, min(args.e2e_cols // world, ncol_local))
    hdt = torch.complex128 if cplx else torch.float64
    xh = torch.empty((e2e_cols, g.Nd), dtype=hdt).pin_memory()
    yh = torch.empty((e2e_cols, g.Nd), dtype=hdt).pin_memory()
    from sparc_b200 import problem as P
    xh.numpy()[:] = P.random_columns(g.Nd, 1, first_col=first_col, seed=1)[0]  # same column replicated: content is irrelevant to timing
    del slots  # free HBM for the host entry point's own buffers
    torch.cuda.empty_cache()
    ctx.ChebyshevFiltering(xh, yh, m, a, b, a0, copy_back_x=False)  # warm-up (allocations, page registration)
    barrier()
    t0 = time.perf_counter()
    ctx.ChebyshevFiltering(xh, yh, m, a, b, a0, copy_back_x=False)
    checksum = float(yh[0, :8].sum().real)  # device->host result is read on the host
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_total_cols = e2e_cols * world
    e2e = {"value": g.Nd * e2e_total_cols / t_e2e, "unit": UNIT,
           "h2d_bytes_per_step": int(e2e_cols * g.Nd * 8 * words), "d2h_bytes_per_step": int(e2e_cols * g.Nd * 8 * words),
           "columns": e2e_total_cols, "seconds": t_e2e, "checksum": checksum,
           "api": "chefsi_chebyshev_filter (host buffers, pinned), X copy-back off as in the SPARC shim"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        rate, cores, kind, sample, _ = cpu_filter_rate(g, veff, proj, bounds, m, args.cpu_cols_per_core)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, proj), "parallelism": f"band-split x{world} (npband)",
                       "columns_per_launch": block, "per_h_apply_value": value * m,
                       "l2": "inputs per launch (>= 8 GB) far exceed the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
