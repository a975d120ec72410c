"""ctypes wrapper of the CPU oracle (oracle/libocean_oracle.so).  TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never from the product package.  `OracleBackend` offers the
same methods as the product's `B200Backend`, so a `HydrostaticFreeSurfaceModel` can be
built on either and the parity tests read the same for both.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libocean_oracle.so")
# the checker of libocean_b200_f32.so: the same restatement built with real = float (its C interface stays Float64)
LIB_F32 = os.path.join(HERE, "libocean_oracle_f32.so")


def build(force: bool = False) -> str:
    """Build the oracle if its .so is absent (the prebuilt .so travels to the GPU box; file times there are
    not meaningful, so staleness is only checked by `make` when forced, i.e. from __graft_entry__.build())."""
    if force or not os.path.exists(LIB) or not os.path.exists(LIB_F32):
        subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)
    return LIB


_libs = {}


def load(precision: str = "Float64"):
    if precision in _libs:
        return _libs[precision]
    path = {"Float64": LIB, "Float32": LIB_F32}[precision]
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    H, dp = C.c_void_p, C.c_void_p
    lib.oracle_create.argtypes = [C.c_void_p]
    lib.oracle_create.restype = H
    lib.oracle_destroy.argtypes = [H]
    lib.oracle_destroy.restype = None
    for f in (lib.oracle_set_field, lib.oracle_get_field):
        f.argtypes = [H, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        f.restype = C.c_int
    for f in (lib.oracle_get_strip, lib.oracle_set_strip):
        f.argtypes = [H, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        f.restype = C.c_int
    lib.oracle_set_forcing.argtypes = [H, C.c_int, dp]
    lib.oracle_set_forcing.restype = C.c_int
    lib.oracle_set_option.argtypes = [H, C.c_char_p, C.c_int]
    lib.oracle_set_option.restype = C.c_int
    lib.oracle_update_state.argtypes = [H]
    lib.oracle_time_step.argtypes = [H, C.c_double]
    lib.oracle_run_stage.argtypes = [H, C.c_int, C.c_double, C.c_int]
    lib.oracle_get_clock.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.oracle_set_clock.argtypes = [H, C.c_double, C.c_int64, C.c_double]
    lib.oracle_tick.argtypes = [H, C.c_double]
    lib.oracle_zero_prev_tendencies.argtypes = [H]
    lib.oracle_weno.argtypes = [C.c_int, dp, dp, dp]
    lib.oracle_weno.restype = C.c_double
    lib.oracle_teos10_rprime.argtypes = [C.c_double] * 3
    lib.oracle_teos10_rprime.restype = C.c_double
    lib.oracle_det_tanh.argtypes = [C.c_double]
    lib.oracle_det_tanh.restype = C.c_double
    lib.oracle_set_num_threads.argtypes = [C.c_int]
    lib.oracle_set_num_threads.restype = None
    lib.oracle_get_max_threads.restype = C.c_int
    lib.oracle_threads_in_parallel_region.restype = C.c_int
    lib.oracle_teos10_rprime_derivatives.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.oracle_teos10_rprime_derivatives.restype = C.c_double
    lib.oracle_real_bytes.restype = C.c_int
    assert lib.oracle_real_bytes() == (8 if precision == "Float64" else 4), path
    _libs[precision] = lib
    return lib


class OracleBackend:
    name = "oracle"
    precision = "Float64"

    def __init__(self, cfg):
        self.lib = load(self.precision)
        self.h = self.lib.oracle_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.Nx, self.Ny, self.Nz = cfg.Nx, cfg.Ny, cfg.Nz
        self.grid_for_diagnostics = None      # set by tests that drive the TimeStepWizard on the oracle

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle error {rc}")

    def set_field(self, fid, arr, halo=(0, 0, 0)):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._check(self.lib.oracle_set_field(self.h, fid, arr.ctypes.data, *halo))

    def get_field(self, fid, shape, halo=(0, 0, 0)):
        out = np.zeros(shape, dtype=np.float64) if any(halo) else np.empty(shape, dtype=np.float64)
        self._check(self.lib.oracle_get_field(self.h, fid, out.ctypes.data, *halo))
        return out

    def get_level(self, fid, k, shape2d, halo=(0, 0)):
        hx, hy = halo
        full = self.get_field(fid, (self._nz(fid), shape2d[0], shape2d[1]), (hx, hy, 0))
        return np.ascontiguousarray(full[k])

    def _nz(self, fid):
        return self.Nz + 1 if fid in (2, 6, 7, 24, 25) else self.Nz      # W, KAPPA_C, KAPPA_U, QX, QY

    def diagnostics(self, grid=None):
        """max |u|, |v|, |w|, |eta|, |T|, |S| and the advective CFL time scale 1 / max(|u|/dx + |v|/dy + |w|/dz)
        (the progress callback and TimeStepWizard of near_global_simulation.jl:138-162), with numpy."""
        g = self.grid_for_diagnostics if grid is None else grid
        Nx, Ny, Nz = self.Nx, self.Ny, self.Nz
        u = self.get_field(0, (Nz, Ny, Nx)); v = self.get_field(1, (Nz, Ny + 1, Nx)); w = self.get_field(2, (Nz + 1, Ny, Nx))
        T = self.get_field(3, (Nz, Ny, Nx)); S = self.get_field(4, (Nz, Ny, Nx)); eta = self.get_field(16, (Ny, Nx))
        mx = [float(np.max(np.abs(a))) for a in (u, v, w, eta, T, S)]
        rad = np.pi / 180.0
        dx = g.radius * np.cos(g.phi_c() * rad) * g.dlam * rad
        dy = g.radius * g.dphi * rad
        dz = g.z_faces[1:] - g.z_faces[:-1]
        rate = np.abs(u) / dx[None, :, None] + np.abs(v[:, :-1, :]) / dy + np.abs(w[:-1]) / dz[:, None, None]
        m = float(np.max(rate))
        return mx, (1.0 / m if m > 0 else float("inf"))

    def set_option(self, name, value):
        self._check(self.lib.oracle_set_option(self.h, name.encode(), int(value)))

    def get_strip(self, fid, n, side, width, interior_edge):
        out = np.empty(n, dtype=np.float64)
        self._check(self.lib.oracle_get_strip(self.h, fid, out.ctypes.data, side, width, int(interior_edge)))
        return out

    def set_strip(self, fid, arr, side, width, interior_edge):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._check(self.lib.oracle_set_strip(self.h, fid, arr.ctypes.data, side, width, int(interior_edge)))

    def set_forcing(self, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._check(self.lib.oracle_set_forcing(self.h, which, arr.ctypes.data))

    def update_state(self):
        self._check(self.lib.oracle_update_state(self.h))

    def time_step(self, dt):
        self._check(self.lib.oracle_time_step(self.h, float(dt)))

    def time_steps(self, dt, n):
        for _ in range(n):
            self.time_step(dt)

    def run_stage(self, stage, dt=0.0, arg=0):
        self._check(self.lib.oracle_run_stage(self.h, stage, float(dt), int(arg)))

    def sync(self):
        pass

    def clock(self):
        t, it, ldt = C.c_double(), C.c_int64(), C.c_double()
        self.lib.oracle_get_clock(self.h, C.byref(t), C.byref(it), C.byref(ldt))
        return t.value, it.value, ldt.value

    def set_clock(self, time, iteration, last_dt):
        self.lib.oracle_set_clock(self.h, float(time), int(iteration), float(last_dt))

    def tick(self, dt):
        self.lib.oracle_tick(self.h, float(dt))

    def zero_prev_tendencies(self):
        self.lib.oracle_zero_prev_tendencies(self.h)

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleBackendF32(OracleBackend):
    """The Float32 build of the oracle.  Buffers at its interface are Float64 arrays holding Float32 values exactly."""
    name = "oracle_f32"
    precision = "Float32"


def weno(N, q, sa=None, sb=None) -> float:
    lib = load()
    q = np.ascontiguousarray(q, dtype=np.float64)
    sa = q if sa is None else np.ascontiguousarray(sa, dtype=np.float64)
    sbp = None if sb is None else np.ascontiguousarray(sb, dtype=np.float64).ctypes.data
    return lib.oracle_weno(N, q.ctypes.data, sa.ctypes.data, sbp)


def set_num_threads(n: int) -> int:
    """Size of the oracle's OpenMP team (overrides an inherited OMP_NUM_THREADS); returns the threads a parallel region
    actually gets."""
    lib = load()
    lib.oracle_set_num_threads(int(n))
    return int(lib.oracle_threads_in_parallel_region())
