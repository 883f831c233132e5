"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle_sfm.so, the CPU restatement of the Rust
reference (see oracle/sfm_oracle.hpp). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"
LIB = ORACLE_DIR / "liboracle_sfm.so"

fp = C.POINTER(C.c_float)
up = C.POINTER(C.c_uint32)


class FieldArgs(C.Structure):
    _fields_ = [("unit", C.c_float), ("fy", C.c_int), ("fx", C.c_int), ("n_maps", C.c_int),
                ("distance_map", fp), ("potential_maps", fp)]


_lib = None


def build(force: bool = False) -> Path:
    srcs = list(ORACLE_DIR.glob("*.cpp")) + list(ORACLE_DIR.glob("*.hpp"))
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.oracle_bilinear.restype = C.c_float
        L.oracle_bilinear.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float]
        L.oracle_sobel_filter.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float, fp]
        L.oracle_distance_from_line.argtypes = [C.c_float] * 6 + [fp]
        L.oracle_get_potential.restype = C.c_float
        L.oracle_get_potential.argtypes = [C.POINTER(FieldArgs), C.c_uint, C.c_float, C.c_float]
        L.oracle_model_new.restype = C.c_void_p
        L.oracle_model_new.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, fp]
        L.oracle_model_free.argtypes = [C.c_void_p]
        L.oracle_model_grid_shape.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_model_spawn.argtypes = [C.c_void_p, C.POINTER(FieldArgs), C.c_int, fp, up, fp]
        L.oracle_model_update.argtypes = [C.c_void_p, C.POINTER(FieldArgs)]
        L.oracle_model_count.argtypes = [C.c_void_p]
        L.oracle_model_get.argtypes = [C.c_void_p, fp, up, fp, fp]
        L.oracle_model_set.argtypes = [C.c_void_p, C.c_int, fp, up, fp, fp]
        L.oracle_model_indices_len.argtypes = [C.c_void_p]
        L.oracle_model_indices.argtypes = [C.c_void_p, up]
        L.oracle_model_accelerations.argtypes = [C.c_void_p, fp]
        L.oracle_model_run.restype = C.c_longlong
        L.oracle_model_run.argtypes = [C.c_void_p, C.POINTER(FieldArgs), C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_field_shape.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_field_build.restype = C.c_int
        L.oracle_field_build.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, fp, C.c_int, fp,
                                         C.POINTER(C.c_uint8), fp, fp]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(fp)


def _u(a):
    return a.ctypes.data_as(up)


def bilinear(grid: np.ndarray, x: float, y: float) -> float:
    g = np.ascontiguousarray(grid, np.float32)
    return lib().oracle_bilinear(_f(g), g.shape[0], g.shape[1], x, y)


def sobel_filter(grid: np.ndarray, x: float, y: float):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.zeros(2, np.float32)
    lib().oracle_sobel_filter(_f(g), g.shape[0], g.shape[1], x, y, _f(out))
    return out


def distance_from_line(p, a, b):
    out = np.zeros(2, np.float32)
    lib().oracle_distance_from_line(p[0], p[1], a[0], a[1], b[0], b[1], _f(out))
    return out


def field_build(size, unit, obstacles: np.ndarray, waypoints: np.ndarray):
    """field.rs:220-232 via oracle/field_oracle.cpp. obstacles/waypoints: (n, 5) x0,y0,x1,y1,width."""
    L = lib()
    fy, fx = C.c_int(), C.c_int()
    L.oracle_field_shape(size[0], size[1], unit, C.byref(fy), C.byref(fx))
    fy, fx = fy.value, fx.value
    obstacles = np.ascontiguousarray(obstacles, np.float32).reshape(-1, 5)
    waypoints = np.ascontiguousarray(waypoints, np.float32).reshape(-1, 5)
    obs = np.zeros((fy, fx), np.uint8)
    dist = np.zeros((fy, fx), np.float32)
    pots = np.zeros((len(waypoints), fy, fx), np.float32)
    rc = L.oracle_field_build(size[0], size[1], unit, len(obstacles), _f(obstacles), len(waypoints), _f(waypoints),
                              obs.ctypes.data_as(C.POINTER(C.c_uint8)), _f(dist), _f(pots))
    assert rc == 0
    return obs.astype(bool), dist, pots


class OracleModel:
    """`SocialForceModel` (sfm.rs) restated in C++."""

    def __init__(self, size, neighbor_unit, field_unit, distance_map, potential_maps, obstacles=None,
                 use_neighbor_grid=True, use_distance_map=True):
        self.L = lib()
        self.dist = np.ascontiguousarray(distance_map, np.float32)
        self.pots = np.ascontiguousarray(potential_maps, np.float32)
        fy, fx = self.dist.shape
        self.fa = FieldArgs(field_unit, fy, fx, self.pots.shape[0], _f(self.dist), _f(self.pots))
        obstacles = np.zeros((0, 5), np.float32) if obstacles is None else \
            np.ascontiguousarray(obstacles, np.float32).reshape(-1, 5)
        self.h = self.L.oracle_model_new(size[0], size[1], neighbor_unit, int(use_neighbor_grid),
                                         int(use_distance_map), len(obstacles), _f(obstacles))

    def grid_shape(self):
        ny, nx = C.c_int(), C.c_int()
        self.L.oracle_model_grid_shape(self.h, C.byref(ny), C.byref(nx))
        return ny.value, nx.value

    def spawn(self, pos=None, dest=None, v0=None):
        if pos is None or len(dest) == 0:
            self.L.oracle_model_spawn(self.h, C.byref(self.fa), 0, None, None, None)
            return
        pos = np.ascontiguousarray(pos, np.float32)
        dest = np.ascontiguousarray(dest, np.uint32)
        v0 = np.ascontiguousarray(v0, np.float32)
        self.L.oracle_model_spawn(self.h, C.byref(self.fa), len(dest), _f(pos), _u(dest), _f(v0))

    def update(self):
        self.L.oracle_model_update(self.h, C.byref(self.fa))

    def count(self):
        return self.L.oracle_model_count(self.h)

    def get(self):
        n = self.count()
        pos = np.empty((n, 2), np.float32)
        dest = np.empty(n, np.uint32)
        vel = np.empty((n, 2), np.float32)
        v0 = np.empty(n, np.float32)
        self.L.oracle_model_get(self.h, _f(pos), _u(dest), _f(vel), _f(v0))
        return pos, dest, vel, v0

    def set(self, pos, dest, vel, v0):
        pos = np.ascontiguousarray(pos, np.float32)
        dest = np.ascontiguousarray(dest, np.uint32)
        vel = np.ascontiguousarray(vel, np.float32)
        v0 = np.ascontiguousarray(v0, np.float32)
        self.L.oracle_model_set(self.h, len(dest), _f(pos), _u(dest), _f(vel), _f(v0))

    def indices(self):
        n = self.L.oracle_model_indices_len(self.h)
        out = np.empty(n, np.uint32)
        self.L.oracle_model_indices(self.h, _u(out))
        return out

    def accelerations(self):
        out = np.empty((self.count(), 2), np.float32)
        self.L.oracle_model_accelerations(self.h, _f(out))
        return out

    def run(self, steps: int):
        ts, tc = C.c_double(), C.c_double()
        updates = self.L.oracle_model_run(self.h, C.byref(self.fa), steps, C.byref(ts), C.byref(tc))
        return updates, ts.value, tc.value

    def __del__(self):
        try:
            self.L.oracle_model_free(self.h)
        except Exception:
            pass
