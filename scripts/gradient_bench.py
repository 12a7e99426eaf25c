"""Bandwidth of the first-derivative kernels (gradient.cu) on a device-resident block: 16 B per grid-pt*column
(8 B read + 8 B written; x2 for complex) against the measured HBM peak."""
import ctypes as C
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparc_b200 import problem as P
from sparc_b200.chefsi import ChefsiContext

peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
n, ncol = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 64)
L = 45.9 * n / 160.0
g = P.make_grid((n, n, n), (L, L, L))
ctx = ChefsiContext(0)
ctx.set_grid(g); ctx.set_veff(P.synthetic_veff(g)); ctx.set_projectors(None)
ld = ctx.device_ld
stream = torch.cuda.ExternalStream(ctx.stream)
for cplx in (False, True):
    words = 2 if cplx else 1
    x = torch.rand((ncol, ld * words), dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    for dir in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.cuda.stream(stream):
            for rep in range(3):
                ctx._check(ctx._lib.chefsi_gradient_mult_device(ctx._h, ncol, 0.0, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), dir, 0.1, int(cplx)))
            ev[0].record(stream)
            for rep in range(10):
                ctx._check(ctx._lib.chefsi_gradient_mult_device(ctx._h, ncol, 0.0, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), dir, 0.1, int(cplx)))
            ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        gbs = 16.0 * words * g.Nd * ncol / ms / 1e6
        print(f"{n}^3 x {ncol} {'complex' if cplx else 'real'} dir {dir}: {ms:.3f} ms, {gbs:.0f} GB/s = {gbs/peak:.2f} of the measured HBM peak ({peak:.0f})", flush=True)
ctx.close()
