"""ctypes binding of include/ocean_b200.h -- the same C ABI a Julia `ccall` would bind.

The product path has NO fallback: if libocean_b200.so is missing or fails to load,
`load_library()` raises.  (The CPU oracle under oracle/ is test infrastructure and is
never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OB200_LIB", os.path.join(HERE, "csrc", "libocean_b200.so"))   # OB200_LIB: tuning variants
# the reference's PRECISION (experiments/run.jl:38): the Float32 build of the same sources, same entry points
LIB_PATH_F32 = os.environ.get("OB200_LIB_F32", os.path.join(HERE, "csrc", "libocean_b200_f32.so"))
PRECISIONS = {"Float64": (LIB_PATH, np.float64), "Float32": (LIB_PATH_F32, np.float32)}

OB200_HALO = 7
OB200_EOS_LINEAR, OB200_EOS_TEOS10 = 0, 1
OB200_BC_QUIESCENT, OB200_BC_DOUBLEDRAKE, OB200_BC_REALISTIC = 0, 1, 2
OB200_BAROX_AUTO, OB200_BAROX_WIDE, OB200_BAROX_SUBSTEP = 0, 1, 2

FIELDS = ["U", "V", "W", "T", "S", "P", "KAPPA_C", "KAPPA_U", "GU", "GV", "GT", "GS",
          "GU_PREV", "GV_PREV", "GT_PREV", "GS_PREV", "ETA", "UBAR", "VBAR", "UTRANS", "VTRANS",
          "GUBT", "GVBT", "NU_LEITH", "QX", "QY", "LD"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}
STAGES = ["BAROTROPIC_FORCING", "AB2_IMPLICIT", "BAROTROPIC_INIT", "BAROTROPIC_ETA", "BAROTROPIC_VEL",
          "BAROTROPIC_FINISH", "CORRECT_AND_STORE", "MASK_AND_LOCAL_HALOS", "AUXILIARIES", "TENDENCIES",
          "BAROTROPIC_LOOP", "EXCHANGE_HALOS"]
STAGE_ID = {n: i for i, n in enumerate(STAGES)}

_Y_FACE = {"V", "GV", "GV_PREV", "VBAR", "VTRANS", "GVBT"}
_Z_FACE = {"W", "KAPPA_C", "KAPPA_U", "QX", "QY"}
_TWO_D = {"ETA", "UBAR", "VBAR", "UTRANS", "VTRANS", "GUBT", "GVBT", "LD"}


def field_shape(name: str, Nx: int, Ny: int, Nz: int):
    """Interior extents of a field as a C-ordered numpy shape (k, j, i) / (j, i)."""
    ny = Ny + 1 if name in _Y_FACE else Ny
    if name in _TWO_D:
        return (ny, Nx)
    nz = Nz + 1 if name in _Z_FACE else Nz
    return (nz, ny, Nx)


class ob200_config(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
        ("Nx_global", C.c_int32), ("i_offset", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("lon_west", C.c_double), ("lon_east", C.c_double),
        ("lat_south", C.c_double), ("lat_north", C.c_double),
        ("radius", C.c_double), ("gravity", C.c_double), ("rotation_rate", C.c_double),
        ("eos_kind", C.c_int32),
        ("eos_alpha", C.c_double), ("eos_beta", C.c_double), ("eos_rho0", C.c_double),
        ("nu_z", C.c_double), ("kappa_z", C.c_double),
        ("ri_enabled", C.c_int32),
        ("ri_nu0", C.c_double), ("ri_kappa0", C.c_double), ("ri_kappa_ca", C.c_double), ("ri_Cen", C.c_double),
        ("ri_Cav", C.c_double), ("ri_Ri0", C.c_double), ("ri_Ridelta", C.c_double),
        ("order_vorticity", C.c_int32), ("order_divergence", C.c_int32), ("order_tracer", C.c_int32),
        ("ab2_chi", C.c_double),
        ("n_substeps", C.c_int32), ("frac_dtau", C.c_double),
        ("bc_kind", C.c_int32),
        ("wind_coeffs", C.c_double * 24), ("salt_coeffs", C.c_double * 16),
        ("T_restoring_lambda", C.c_double), ("drag_mu", C.c_double),
        ("qgleith_enabled", C.c_int32),
        ("leith_C", C.c_double), ("leith_min_N2", C.c_double), ("leith_Vscale", C.c_double),
        ("bottom_halo_columns", C.c_int32),
        ("barotropic_exchange", C.c_int32),
        ("z_faces", C.POINTER(C.c_double)),
        ("bottom_height", C.POINTER(C.c_double)),
        ("averaging_weights", C.POINTER(C.c_double)),
    ]


class ob200_timing(C.Structure):
    _fields_ = [("n", C.c_int32), ("ms", C.c_float * 16), ("launches", C.c_int64)]


def as_double_ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


# every symbol include/ocean_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "ob200_create", "ob200_destroy", "ob200_last_error", "ob200_version",
    "ob200_set_field", "ob200_get_field", "ob200_get_level", "ob200_set_forcing",
    "ob200_update_state", "ob200_time_step", "ob200_time_steps", "ob200_run_stage", "ob200_sync",
    "ob200_get_clock", "ob200_set_clock", "ob200_diagnostics", "ob200_get_timing", "ob200_set_stream",
    "ob200_set_option", "ob200_comm_unique_id", "ob200_comm_init",
    "ob200_bind", "ob200_unbind", "ob200_sync_outputs", "ob200_sync_inputs", "ob200_real_bytes",
]

_libs = {}


def load_library(path: str | None = None, precision: str = "Float64"):
    """dlopen libocean_b200.so (or its Float32 build) and declare the prototypes.  Raises if the library is absent."""
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)} (experiments/run.jl:38), got {precision!r}")
    if path is None and precision in _libs:
        return _libs[precision]
    p = path or PRECISIONS[precision][0]
    if not os.path.exists(p):
        raise RuntimeError(
            f"{os.path.basename(p)} not found at {p}: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (there is no CPU fallback for the product path)")
    try:
        import torch  # noqa: F401  -- loads the bundled libnccl.so.2 / libcudart first so sonames resolve to them
    except Exception:
        pass
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    dp = C.c_void_p  # raw address: host numpy pointer or device pointer
    lib.ob200_create.argtypes = [C.POINTER(ob200_config), C.c_int32, C.POINTER(H)]
    lib.ob200_create.restype = C.c_int
    lib.ob200_destroy.argtypes = [H]
    lib.ob200_destroy.restype = None
    lib.ob200_last_error.argtypes = [H]
    lib.ob200_last_error.restype = C.c_char_p
    lib.ob200_version.argtypes = []
    lib.ob200_version.restype = C.c_char_p
    lib.ob200_real_bytes.argtypes = []
    lib.ob200_real_bytes.restype = C.c_int32
    if path is None and lib.ob200_real_bytes() != np.dtype(PRECISIONS[precision][1]).itemsize:
        raise RuntimeError(f"{p} is not the {precision} build (ob200_real_bytes() = {lib.ob200_real_bytes()})")
    for f in (lib.ob200_set_field, lib.ob200_get_field):
        f.argtypes = [H, C.c_int32, dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        f.restype = C.c_int
    lib.ob200_get_level.argtypes = [H, C.c_int32, C.c_int32, dp, C.c_int32, C.c_int32, C.c_int32]
    lib.ob200_get_level.restype = C.c_int
    lib.ob200_set_forcing.argtypes = [H, C.c_int32, dp, C.c_int32]
    lib.ob200_set_forcing.restype = C.c_int
    lib.ob200_update_state.argtypes = [H]
    lib.ob200_update_state.restype = C.c_int
    lib.ob200_time_step.argtypes = [H, C.c_double]
    lib.ob200_time_step.restype = C.c_int
    lib.ob200_time_steps.argtypes = [H, C.c_double, C.c_int32]
    lib.ob200_time_steps.restype = C.c_int
    lib.ob200_run_stage.argtypes = [H, C.c_int32, C.c_double, C.c_int32]
    lib.ob200_run_stage.restype = C.c_int
    lib.ob200_sync.argtypes = [H]
    lib.ob200_sync.restype = C.c_int
    lib.ob200_get_clock.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.ob200_get_clock.restype = C.c_int
    lib.ob200_set_clock.argtypes = [H, C.c_double, C.c_int64, C.c_double]
    lib.ob200_set_clock.restype = C.c_int
    lib.ob200_diagnostics.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ob200_diagnostics.restype = C.c_int
    lib.ob200_get_timing.argtypes = [H, C.POINTER(ob200_timing)]
    lib.ob200_get_timing.restype = C.c_int
    lib.ob200_set_stream.argtypes = [H, C.c_void_p]
    lib.ob200_set_stream.restype = C.c_int
    lib.ob200_set_option.argtypes = [H, C.c_char_p, C.c_int32]
    lib.ob200_set_option.restype = C.c_int
    lib.ob200_bind.argtypes = [H, C.c_int32, dp, C.c_int32, C.c_int32, C.c_int32]
    lib.ob200_bind.restype = C.c_int
    lib.ob200_unbind.argtypes = [H, C.c_int32]
    lib.ob200_unbind.restype = C.c_int
    lib.ob200_sync_outputs.argtypes = [H]
    lib.ob200_sync_outputs.restype = C.c_int
    lib.ob200_sync_inputs.argtypes = [H]
    lib.ob200_sync_inputs.restype = C.c_int
    lib.ob200_comm_unique_id.argtypes = [C.c_void_p]
    lib.ob200_comm_unique_id.restype = C.c_int
    lib.ob200_comm_init.argtypes = [H, C.c_void_p]
    lib.ob200_comm_init.restype = C.c_int
    if path is None:
        _libs[precision] = lib
    return lib
