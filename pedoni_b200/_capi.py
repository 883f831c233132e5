"""ctypes binding of include/pedoni_cuda.h (libpedoni_cuda.so).

This is the harness-side stub of the C ABI: the same entry points the Rust shim in ffi/ binds. There
is no fallback — if the shared library is missing or a symbol is absent, import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libpedoni_cuda.so"

PEDONI_ABI_VERSION = 3
PEDONI_OK = 0
PEDONI_ERR_INVALID = -1
PEDONI_ERR_CUDA = -2
PEDONI_ERR_STATE = -3
PEDONI_ERR_CAPACITY = -4
PEDONI_ERR_UNSUPPORTED = -5
PEDONI_ERR_COMM = -6
PEDONI_MATH_STRICT = 0
PEDONI_MATH_FAST = 1
PEDONI_COMM_ID_BYTES = 128

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)


class PedoniConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("field_size_x", C.c_float),
        ("field_size_y", C.c_float),
        ("neighbor_grid_unit", C.c_float),
        ("field_grid_unit", C.c_float),
        ("use_neighbor_grid", C.c_int32),
        ("use_distance_map", C.c_int32),
        ("field_ny", C.c_int32),
        ("field_nx", C.c_int32),
        ("n_potential_maps", C.c_int32),
        ("n_obstacles", C.c_int32),
        ("distance_map", c_float_p),
        ("potential_maps", c_float_p),
        ("obstacles", c_float_p),
        ("capacity", C.c_uint32),
        ("math_mode", C.c_int32),
        ("slab_rank", C.c_int32),
        ("slab_count", C.c_int32),
        ("stream", C.c_void_p),
        ("halo_capacity", C.c_uint32),
    ]


class PedoniSpawnGroup(C.Structure):
    _fields_ = [("p1_x", C.c_float), ("p1_y", C.c_float), ("p2_x", C.c_float), ("p2_y", C.c_float),
                ("destination", C.c_uint32), ("count", C.c_uint32)]


class PedoniSpawnRate(C.Structure):
    _fields_ = [("p1_x", C.c_float), ("p1_y", C.c_float), ("p2_x", C.c_float), ("p2_y", C.c_float),
                ("destination", C.c_uint32), ("frequency", C.c_double)]


class PedoniObservables(C.Structure):
    _fields_ = [("count", C.c_uint32), ("mean_speed", C.c_float), ("per_destination", C.c_uint32 * 16),
                ("arrived", C.c_uint64 * 16), ("n_bins", C.c_uint32), ("bin_count", C.c_uint32 * 64),
                ("bin_mean_vx", C.c_float * 64)]


class PedoniKernelTimes(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("key_ms", "histogram_ms", "scan_ms", "scatter_ms", "gather_ms", "force_ms", "comm_ms")] + \
               [(n, C.c_uint64) for n in
                ("key_launches", "histogram_launches", "scan_launches", "scatter_launches", "gather_launches",
                 "force_launches", "comm_launches", "force_agents")] + \
               [("force_edge_ms", C.c_double), ("pack_ms", C.c_double), ("force_edge_launches", C.c_uint64),
                ("pack_launches", C.c_uint64)]


class PedoniLaunchRecord(C.Structure):
    _fields_ = [("kind", C.c_int32), ("stream", C.c_int32), ("start_ms", C.c_float), ("stop_ms", C.c_float)]


# name -> (restype, argtypes): every symbol include/pedoni_cuda.h declares.
SIGNATURES = {
    "pedoni_abi_version": (C.c_int, []),
    "pedoni_create": (C.c_int, [C.POINTER(PedoniConfig), C.POINTER(C.c_void_p)]),
    "pedoni_destroy": (None, [C.c_void_p]),
    "pedoni_last_error": (C.c_char_p, [C.c_void_p]),
    "pedoni_spawn": (C.c_int, [C.c_void_p, C.c_uint32, c_float_p, c_u32_p, c_float_p]),
    "pedoni_spawn_groups": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PedoniSpawnGroup), C.c_uint64, C.c_uint64]),
    "pedoni_spawn_stream_seek": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "pedoni_spawn_stream_tell": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pedoni_spawn_poisson": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PedoniSpawnRate)]),
    "pedoni_rebuild": (C.c_int, [C.c_void_p]),
    "pedoni_step": (C.c_int, [C.c_void_p]),
    "pedoni_count": (C.c_int32, [C.c_void_p]),
    "pedoni_count_published": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), c_u32_p]),
    "pedoni_download": (C.c_int, [C.c_void_p, c_float_p, c_u32_p, c_float_p, c_float_p, C.c_uint32, c_u32_p]),
    "pedoni_download_begin": (C.c_int, [C.c_void_p, c_float_p, c_u32_p, C.c_uint32]),
    "pedoni_download_end": (C.c_int, [C.c_void_p, c_u32_p]),
    "pedoni_download_begin_u8": (C.c_int, [C.c_void_p, c_float_p, C.POINTER(C.c_uint8), C.c_uint32]),
    "pedoni_observe": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_uint32, C.POINTER(PedoniObservables)]),
    "pedoni_upload_state": (C.c_int, [C.c_void_p, C.c_uint32, c_float_p, c_u32_p, c_float_p, c_float_p]),
    "pedoni_grid_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pedoni_cell_table": (C.c_int, [C.c_void_p, c_u32_p, C.c_uint32, c_u32_p]),
    "pedoni_synchronize": (C.c_int, [C.c_void_p]),
    "pedoni_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "pedoni_profile_reset": (C.c_int, [C.c_void_p]),
    "pedoni_profile_read": (C.c_int, [C.c_void_p, C.POINTER(PedoniKernelTimes)]),
    "pedoni_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pedoni_timer_begin": (C.c_int, [C.c_void_p]),
    "pedoni_timer_end": (C.c_int, [C.c_void_p, c_float_p]),
    "pedoni_field_shape": (C.c_int, [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pedoni_field_build": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_int32, c_float_p, C.c_int32, c_float_p,
                                     C.POINTER(C.c_uint8), c_float_p, c_float_p]),
    "pedoni_field_build_device": (C.c_int, [C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, c_float_p, C.c_int32,
                                            c_float_p, C.POINTER(C.c_uint8), c_float_p, c_float_p, C.POINTER(C.c_int32)]),
    "pedoni_slab_rows": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pedoni_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pedoni_comm_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pedoni_slab_exchange_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32]),
    "pedoni_slab_transport": (C.c_char_p, [C.c_void_p]),
    "pedoni_halo_capacity": (C.c_int, [C.c_void_p, c_u32_p]),
    "pedoni_field_textures": (C.c_int, [C.c_void_p]),
    "pedoni_wall_far_cells": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pedoni_download_wire_bytes": (C.c_int, [C.c_void_p]),
    "pedoni_profile_timeline": (C.c_int, [C.c_void_p, C.POINTER(PedoniLaunchRecord), C.c_uint32, c_u32_p]),
    "pedoni_host_alloc": (C.c_void_p, [C.c_size_t]),
    "pedoni_host_free": (None, [C.c_void_p]),
}

_lib = None


class PedoniError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"pedoni_cuda error {code}: {message}")
        self.code = code
        self.message = message


def load() -> C.CDLL:
    """Load libpedoni_cuda.so (built in-tree by `python -m pedoni_b200.build`). No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("PEDONI_CUDA_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: build the CUDA extension first (python -m pedoni_b200.build). "
            "pedoni_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pedoni_abi_version() != PEDONI_ABI_VERSION:
        raise ImportError(f"ABI version mismatch: library {lib.pedoni_abi_version()} != binding {PEDONI_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, handle=None) -> int:
    if rc < 0:
        msg = load().pedoni_last_error(handle)
        raise PedoniError(rc, msg.decode() if msg else "")
    return rc
