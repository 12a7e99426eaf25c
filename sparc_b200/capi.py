"""ctypes binding of ``libchefsi_b200.so`` (the C ABI in ``include/chefsi_b200.h``).

This is the product path: it loads the hand-written sm_100a CUDA library that lives in-tree
next to this file and fails loudly if it is missing or if no B200 is visible.  There is no CPU
fallback and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHEFSI_B200_LIB") or os.path.join(HERE, "libchefsi_b200.so")  # override: A/B builds only

# every symbol include/chefsi_b200.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "chefsi_device_count", "chefsi_create", "chefsi_create_multi", "chefsi_multi_info", "chefsi_destroy", "chefsi_last_error", "chefsi_version",
    "chefsi_set_grid", "chefsi_set_projectors", "chefsi_set_veff", "chefsi_set_kpoint",
    "chefsi_chebyshev_filter", "chefsi_chebyshev_filter_kpt",
    "chefsi_hamiltonian_mult", "chefsi_hamiltonian_mult_kpt", "chefsi_laplacian_mult", "chefsi_laplacian_mult_kpt",
    "chefsi_gradient_mult", "chefsi_gradient_mult_kpt", "chefsi_gradient_mult_device",
    "chefsi_lanczos", "chefsi_lanczos_kpt", "chefsi_subspace_eig", "chefsi_subspace_eig_kpt", "chefsi_band_store", "chefsi_density_accumulate", "chefsi_density_accumulate_kpt", "chefsi_poisson_aar", "chefsi_subspace_reserve", "chefsi_subspace_reserve_kpt", "chefsi_subspace_project_kpt", "chefsi_subspace_rotate_kpt", "chefsi_subspace_project", "chefsi_subspace_rotate",
    "chefsi_rank_load", "chefsi_resident_ptr", "chefsi_ipc_export", "chefsi_ipc_open", "chefsi_ipc_close",
    "chefsi_rank_project", "chefsi_rank_project_shared", "chefsi_rank_forms_block", "chefsi_rank_block_part", "chefsi_rank_rotate_prepare", "chefsi_rank_rotate",
    "chefsi_device_ld",
    "chefsi_chebyshev_filter_device", "chefsi_chebyshev_filter_kpt_device",
    "chefsi_hamiltonian_mult_device", "chefsi_hamiltonian_mult_kpt_device",
    "chefsi_synchronize", "chefsi_fill_random_device", "chefsi_pack_device", "chefsi_unpack_device",
    "chefsi_get_stats", "chefsi_set_profiling", "chefsi_stream", "chefsi_host_register", "chefsi_host_unregister",
    "chefsi_stencil_step_device", "chefsi_nloc_project_device", "chefsi_nloc_expand_device",
)


class ChefsiStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_ulonglong),
        ("last_filter_ms", C.c_double),
        ("last_stencil_ms", C.c_double),
        ("last_nloc_ms", C.c_double),
        ("last_stencil_launches", C.c_int),
        ("last_path", C.c_int),
        ("last_nloc_atomic", C.c_int),
        ("last_alpha_reduced", C.c_int),
        ("round_barrier_timeouts", C.c_uint),
        ("reserved_", C.c_int),
        ("density_resident_blocks", C.c_uint),
        ("density_uploaded_blocks", C.c_uint),
        ("band_store_misses", C.c_uint),
        ("reserved2_", C.c_uint),
    ]


class ChefsiError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ChefsiError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, dp, sz, i, d = C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double
    ip = C.POINTER(C.c_int)
    lib.chefsi_device_count.argtypes = []
    lib.chefsi_create.argtypes = [C.POINTER(vp), i]
    lib.chefsi_create_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i]
    lib.chefsi_multi_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.chefsi_destroy.argtypes = [vp]
    lib.chefsi_destroy.restype = None
    lib.chefsi_last_error.argtypes = [vp]
    lib.chefsi_last_error.restype = C.c_char_p
    lib.chefsi_version.restype = C.c_char_p
    lib.chefsi_set_grid.argtypes = [vp, vp]
    lib.chefsi_set_projectors.argtypes = [vp, vp]
    lib.chefsi_set_veff.argtypes = [vp, dp]
    lib.chefsi_set_kpoint.argtypes = [vp, d, d, d]
    for name in ("chefsi_chebyshev_filter", "chefsi_chebyshev_filter_kpt"):
        getattr(lib, name).argtypes = [vp, dp, sz, dp, sz, i, i, d, d, d, i]
    for name in ("chefsi_hamiltonian_mult", "chefsi_hamiltonian_mult_kpt"):
        getattr(lib, name).argtypes = [vp, i, d, dp, sz, dp, sz]
    for name in ("chefsi_laplacian_mult", "chefsi_laplacian_mult_kpt"):
        getattr(lib, name).argtypes = [vp, i, d, d, dp, sz, dp, sz]
    lib.chefsi_gradient_mult.argtypes = [vp, i, d, dp, sz, dp, sz, i]
    lib.chefsi_gradient_mult_kpt.argtypes = [vp, i, d, dp, sz, dp, sz, i, d]
    lib.chefsi_gradient_mult_device.argtypes = [vp, i, d, vp, vp, i, d, i]
    lib.chefsi_rank_load.argtypes = [vp, vp, sz, i, i]
    lib.chefsi_resident_ptr.argtypes = [vp, i]
    lib.chefsi_resident_ptr.restype = C.c_void_p
    lib.chefsi_ipc_export.argtypes = [vp, i, vp]
    lib.chefsi_ipc_open.argtypes = [vp, vp, C.POINTER(C.c_void_p)]
    lib.chefsi_ipc_close.argtypes = [vp, vp]
    lib.chefsi_rank_project.argtypes = [vp, i, i, i, ip, C.POINTER(C.c_void_p), vp, vp, sz]
    lib.chefsi_rank_project_shared.argtypes = [vp, i, i, i, ip, C.POINTER(C.c_void_p), vp, vp, sz]
    lib.chefsi_rank_forms_block.argtypes = [i, i, i]
    lib.chefsi_rank_block_part.argtypes = [i, i, i, i, i, ip, ip, ip, ip]
    lib.chefsi_rank_block_part.restype = None
    lib.chefsi_rank_rotate_prepare.argtypes = [vp, i]
    lib.chefsi_rank_rotate.argtypes = [vp, i, i, i, ip, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), vp, sz, vp, sz]
    lib.chefsi_lanczos.argtypes = [vp, dp, d, d, i, C.POINTER(C.c_double), C.POINTER(C.c_double), ip]
    lib.chefsi_lanczos_kpt.argtypes = [vp, vp, d, d, i, C.POINTER(C.c_double), C.POINTER(C.c_double), ip]
    for name in ("chefsi_subspace_eig", "chefsi_subspace_eig_kpt"):
        getattr(lib, name).argtypes = [vp, i, dp, dp, sz, dp, dp, sz]
    lib.chefsi_band_store.argtypes = [vp, i]
    for name in ("chefsi_density_accumulate", "chefsi_density_accumulate_kpt"):
        getattr(lib, name).argtypes = [vp, dp, sz, i, dp, dp]
    lib.chefsi_poisson_aar.argtypes = [vp, d, dp, dp, d, d, i, i, d, i, ip, C.POINTER(C.c_double)]
    lib.chefsi_subspace_reserve.argtypes = [vp, i]
    lib.chefsi_subspace_reserve_kpt.argtypes = [vp, i]
    lib.chefsi_subspace_project_kpt.argtypes = [vp, dp, sz, i, dp, dp, sz]
    lib.chefsi_subspace_rotate_kpt.argtypes = [vp, dp, sz, i, dp, sz]
    lib.chefsi_subspace_project.argtypes = [vp, dp, sz, i, dp, dp, sz]
    lib.chefsi_subspace_rotate.argtypes = [vp, dp, sz, i, dp, sz]
    lib.chefsi_device_ld.argtypes = [vp]
    lib.chefsi_device_ld.restype = sz
    for name in ("chefsi_chebyshev_filter_device", "chefsi_chebyshev_filter_kpt_device"):
        getattr(lib, name).argtypes = [vp, dp, dp, dp, i, i, d, d, d, ip, ip]
    for name in ("chefsi_hamiltonian_mult_device", "chefsi_hamiltonian_mult_kpt_device"):
        getattr(lib, name).argtypes = [vp, i, d, dp, dp]
    lib.chefsi_synchronize.argtypes = [vp]
    lib.chefsi_fill_random_device.argtypes = [vp, dp, i, C.c_longlong, C.c_ulonglong, i]
    lib.chefsi_pack_device.argtypes = [vp, dp, sz, dp, i, i]
    lib.chefsi_unpack_device.argtypes = [vp, dp, dp, sz, i, i]
    lib.chefsi_get_stats.argtypes = [vp, C.POINTER(ChefsiStats)]
    lib.chefsi_set_profiling.argtypes = [vp, i]
    lib.chefsi_host_register.argtypes = [vp, vp, sz]
    lib.chefsi_host_unregister.argtypes = [vp, vp]
    lib.chefsi_stencil_step_device.argtypes = [vp, dp, dp, dp, i, d, d, d]
    lib.chefsi_nloc_project_device.argtypes = [vp, dp, i, dp]
    lib.chefsi_nloc_expand_device.argtypes = [vp, dp, i, d, dp]
    lib.chefsi_stream.argtypes = [vp]
    lib.chefsi_stream.restype = vp
    _lib = lib
    return lib
