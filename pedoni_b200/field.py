"""`Field` — mirrors pedoni-simulator/src/field.rs:194-205: the one-time precompute whose OUTPUTS
(distance map + one potential map per waypoint) are inputs of the per-timestep hot path.

Building it is the step BEFORE the path (lib.rs:30 builds the field, then hands `&Field` to
`PedestrianModel::new`). `Field.from_scenario` runs the host-side C++ restatement of
`Field::from_scenario` (field.rs:220-232) in libpedoni_cuda.so (csrc/host/field_builder.cpp, SURVEY.md
section 8 row f1) so the shipped scenario TOMLs run without the Rust side; any other builder's arrays
(the reference's own, a closed form for an open domain) can be passed instead; `from_scenario(..., device=k)` runs
the device builder (csrc/field_device.cu), which bench.py uses for the synthetic crowd's 12 656^2 maps."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np


@dataclass
class Field:
    unit: float                  # field.rs:196
    shape: tuple                 # (fy, fx), field.rs:198
    obstacle_exist: np.ndarray   # bool (fy, fx), field.rs:200
    distance_map: np.ndarray     # f32 (fy, fx), field.rs:202
    potential_maps: np.ndarray   # f32 (n_waypoints, fy, fx), field.rs:204

    @staticmethod
    def from_scenario(scenario, unit: float, device=None) -> "Field":
        """field.rs:220-232. device=None: the host builder (the reference's marching restated, works without a GPU);
        device=k: the block-iterative eikonal solver on GPU k (pedoni_field_build_device), for large domains."""
        from . import _capi
        lib = _capi.load()
        fy, fx = C.c_int32(), C.c_int32()
        sx, sy = float(scenario.field.size[0]), float(scenario.field.size[1])
        _capi.check(lib.pedoni_field_shape(sx, sy, unit, C.byref(fy), C.byref(fx)))
        fy, fx = fy.value, fx.value
        pack = lambda items: np.ascontiguousarray(  # noqa: E731
            [[*i.line[0], *i.line[1], i.width] for i in items], dtype=np.float32).reshape(-1, 5)
        obs, wps = pack(scenario.obstacles), pack(scenario.waypoints)
        exist = np.zeros((fy, fx), np.uint8)
        dist = np.zeros((fy, fx), np.float32)
        pots = np.zeros((len(wps), fy, fx), np.float32)
        fp = lambda a: a.ctypes.data_as(_capi.c_float_p)  # noqa: E731
        if device is None:
            _capi.check(lib.pedoni_field_build(sx, sy, unit, len(obs), fp(obs), len(wps), fp(wps),
                                               exist.ctypes.data_as(C.POINTER(C.c_uint8)), fp(dist), fp(pots)))
        else:
            passes = C.c_int32()
            rc = lib.pedoni_field_build_device(int(device), sx, sy, unit, len(obs), fp(obs), len(wps), fp(wps),
                                               exist.ctypes.data_as(C.POINTER(C.c_uint8)), fp(dist), fp(pots),
                                               C.byref(passes))
            if rc < 0:
                raise _capi.PedoniError(rc, "pedoni_field_build_device failed (needs a CUDA device; see stderr)")
        return Field(unit=unit, shape=(fy, fx), obstacle_exist=exist.astype(bool), distance_map=dist,
                     potential_maps=pots)
