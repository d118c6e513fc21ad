"""ctypes loader for libfen_gpu.so (the C ABI declared in include/fen_gpu.h).

There is no fallback: if the CUDA library has not been built (``python -m fen_b200.build``) or no
CUDA device is present, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FEN_GPU_LIB: load an A/B build variant of the library (fen_b200/build.py) instead of the default one
LIB_PATH = os.environ.get("FEN_GPU_LIB") or os.path.join(HERE, "libfen_gpu.so")


class FenError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfen_gpu error %d: %s" % (code, msg))
        self.code = code


class GridDesc(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("ndim", C.c_int),
                ("delta", C.c_double), ("bc", C.c_int * 6), ("rank", C.c_int), ("nranks", C.c_int),
                ("device", C.c_int)]


class NsParams(C.Structure):
    _fields_ = [("density", C.c_double), ("viscosity", C.c_double), ("g", C.c_double * 3),
                ("CFL", C.c_double), ("dt_o", C.c_double), ("dt_visc", C.c_double),
                ("dt_conv", C.c_double), ("constant_CFL", C.c_int)]


class MfParams(C.Structure):
    _fields_ = [("rho_0", C.c_double), ("rho_1", C.c_double), ("mu_0", C.c_double), ("mu_1", C.c_double),
                ("sigma", C.c_double), ("beta", C.c_double), ("cut", C.c_double),
                ("quadratic", C.c_int), ("x_first", C.c_int),
                ("dt_surf", C.c_double), ("rhomin", C.c_double), ("irhomin", C.c_double)]


_P = C.c_void_p
_I = C.c_int
_D = C.c_double
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)

FORCING_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double)     # fen_forcing_fn
DISTANCE_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double, C.c_double)   # fen_distance_fn

# name -> (restype, argtypes); every symbol include/fen_gpu.h declares
SIGNATURES = {
    "fen_gpu_last_error": (C.c_char_p, []),
    "fen_gpu_version": (_I, []),
    "fen_gpu_create": (_I, [C.POINTER(GridDesc), C.POINTER(_P)]),
    "fen_gpu_destroy": (_I, [_P]),
    "fen_gpu_synchronize": (_I, [_P]),
    "fen_gpu_local_bounds": (_I, [_P, _PI, _PI]),
    "fen_gpu_comm_handle_bytes": (_I, []),
    "fen_gpu_comm_export": (_I, [_P, _P]),
    "fen_gpu_comm_connect": (_I, [_P, _P]),
    "fen_gpu_scalar_allocate": (_I, [_P, _I, _I, _PI]),
    "fen_gpu_scalar_destroy": (_I, [_P, _I]),
    "fen_gpu_push": (_I, [_P, _I, _P, _I]),
    "fen_gpu_pull": (_I, [_P, _I, _P, _I]),
    "fen_gpu_pull_async": (_I, [_P, _I, _P, _I]),
    "fen_gpu_pull_wait": (_I, [_P]),
    "fen_gpu_set_to_value": (_I, [_P, _I, _D]),
    "fen_gpu_set_bc_type": (_I, [_P, _I, _I, _I]),
    "fen_gpu_get_bc_type": (_I, [_P, _I, _I, _PI]),
    "fen_gpu_set_bc_plane": (_I, [_P, _I, _I, _P, _I]),
    "fen_gpu_update_ghost_nodes": (_I, [_P, _I, _I]),
    "fen_gpu_update_halos": (_I, [_P, _I]),
    "fen_gpu_max_value": (_I, [_P, _I, _PD]),
    "fen_gpu_integral": (_I, [_P, _I, _PD]),
    "fen_gpu_gradient": (_I, [_P, _I, _I]),
    "fen_gpu_divergence": (_I, [_P, _I, _I]),
    "fen_gpu_laplacian": (_I, [_P, _I, _I]),
    "fen_gpu_center_to_face": (_I, [_P, _I, _I]),
    "fen_gpu_laplacian_scalar": (_I, [_P, _I, _I]),
    "fen_gpu_face_to_center": (_I, [_P, _I, _I, _I]),
    "fen_gpu_curl": (_I, [_P, _I, _I]),
    "fen_gpu_init_poisson_solver": (_I, [_P]),
    "fen_gpu_solve_poisson": (_I, [_P, _I]),
    "fen_gpu_destroy_poisson_solver": (_I, [_P]),
    "fen_gpu_poisson_variant": (C.c_char_p, [_P]),
    "fen_gpu_init_solver": (_I, [_P]),
    "fen_gpu_destroy_solver": (_I, [_P]),
    "fen_gpu_get_params": (_I, [_P, C.POINTER(NsParams)]),
    "fen_gpu_set_params": (_I, [_P, C.POINTER(NsParams)]),
    "fen_gpu_set_timestep": (_I, [_P, _D, _PD]),
    "fen_gpu_navier_stokes_solver": (_I, [_P, _I, _PD]),
    "fen_gpu_get_status": (_I, [_P, _PD, _PD]),
    "fen_gpu_status_line": (_I, [_P, _I, _D, _D, C.c_char_p, _I]),
    "fen_gpu_add_advection": (_I, [_P, _I]),
    "fen_gpu_compute_explicit_terms": (_I, [_P, _I]),
    "fen_gpu_predicted_velocity_field": (_I, [_P, _D]),
    "fen_gpu_correct_velocity_field": (_I, [_P, _D]),
    "fen_gpu_update_pressure": (_I, [_P]),
    "fen_gpu_checks": (_I, [_P, _D]),
    "fen_gpu_scalar_write": (_I, [_P, _I, C.c_char_p]),
    "fen_gpu_scalar_read": (_I, [_P, _I, C.c_char_p]),
    "fen_gpu_save_state": (_I, [_P, C.c_char_p]),
    "fen_gpu_load_state": (_I, [_P, C.c_char_p]),
    "fen_gpu_save_fields": (_I, [_P, _I, C.c_char_p]),
    "fen_gpu_set_forcing_hook": (_I, [_P, FORCING_FN, _P]),
    "fen_gpu_mf_get_params": (_I, [_P, C.POINTER(MfParams)]),
    "fen_gpu_mf_set_params": (_I, [_P, C.POINTER(MfParams)]),
    "fen_gpu_allocate_vof_fields": (_I, [_P]),
    "fen_gpu_get_vof_from_distance": (_I, [_P, DISTANCE_FN, _P, _D, _D]),
    "fen_gpu_get_h_from_vof": (_I, [_P]),
    "fen_gpu_advect_vof": (_I, [_P, _I, _D]),
    "fen_gpu_check_vof_integral": (_I, [_P, _PD, _PD]),
    "fen_gpu_destroy_vof": (_I, [_P]),
    "fen_gpu_update_material_properties": (_I, [_P]),
    "fen_gpu_init_solver_mf": (_I, [_P, DISTANCE_FN, _P, _D, _D]),
    "fen_gpu_profile_enable": (_I, [_P, _I]),
    "fen_gpu_profile_read": (_I, [_P, _I, _P, _PD, _PI, _PI]),
    "fen_gpu_launch_count": (C.c_longlong, [_P]),
    "fen_gpu_stream": (_P, [_P]),
}

_lib = None


def load():
    """Load libfen_gpu.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FenError(-1, "%s is missing: build it with `python -m fen_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise FenError(code, load().fen_gpu_last_error().decode())
