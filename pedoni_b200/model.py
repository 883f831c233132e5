"""`SocialForceModelCuda` — harness-side mirror of the reference's `PedestrianModel` plugin trait
(pedoni-simulator/src/models/mod.rs:13-25) over the C ABI in include/pedoni_cuda.h.

Same method names, argument meaning and error behaviour as the Rust trait: methods do not return
errors; a failing C-ABI call raises (the Rust shim panics, like the reference's unwrap()s at
sfm_gpu.rs:51,69,79,127).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _capi


@dataclass
class Pedestrian:
    """models/mod.rs:28-32 `Pedestrian { pos: Vec2, destination: usize }`."""
    pos: tuple
    destination: int = 0


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _fp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_capi.c_float_p) if a is not None else None


def _up(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_capi.c_u32_p) if a is not None else None


class SocialForceModelCuda:
    """Third `PedestrianModel` implementation (after sfm.rs `SocialForceModel` and sfm_gpu.rs
    `SocialForceModelGpu`), running on one B200."""

    def __init__(self, options, scenario, field, *, device: int = 0, math_mode: int = _capi.PEDONI_MATH_STRICT,
                 capacity: int = 0, slab_rank: int = 0, slab_count: int = 1, stream: int = 0,
                 halo_capacity: int = 0):
        """`PedestrianModel::new(&SimulatorOptions, &Scenario, &Field)` (mod.rs:14)."""
        self._lib = _capi.load()
        self._h = C.c_void_p()
        dist = _f32(field.distance_map)
        pots = _f32(field.potential_maps)
        fy, fx = dist.shape
        assert pots.ndim == 3 and pots.shape[1:] == (fy, fx), "potential_maps must be (n_maps, fy, fx)"
        obstacles = _f32([[*o.line[0], *o.line[1], o.width] for o in scenario.obstacles]).reshape(-1, 5)
        cfg = _capi.PedoniConfig()
        cfg.struct_size = C.sizeof(_capi.PedoniConfig)
        cfg.device = device
        cfg.field_size_x, cfg.field_size_y = float(scenario.field.size[0]), float(scenario.field.size[1])
        cfg.neighbor_grid_unit = float(options.neighbor_grid_unit)
        cfg.field_grid_unit = float(field.unit)
        cfg.use_neighbor_grid = int(bool(options.use_neighbor_grid))
        cfg.use_distance_map = int(bool(options.use_distance_map))
        cfg.field_ny, cfg.field_nx = fy, fx
        cfg.n_potential_maps = pots.shape[0]
        cfg.n_obstacles = obstacles.shape[0]
        cfg.distance_map = _fp(dist)
        cfg.potential_maps = _fp(pots)
        cfg.obstacles = _fp(obstacles) if obstacles.shape[0] else None
        cfg.capacity = capacity
        cfg.math_mode = math_mode
        cfg.slab_rank, cfg.slab_count = slab_rank, slab_count
        cfg.stream = stream or None
        cfg.halo_capacity = halo_capacity
        _capi.check(self._lib.pedoni_create(C.byref(cfg), C.byref(self._h)))
        self.n_maps = pots.shape[0]

    # -- trait methods ---------------------------------------------------------------------------
    @classmethod
    def new(cls, options, scenario, field, **kw) -> "SocialForceModelCuda":
        return cls(options, scenario, field, **kw)

    def spawn_pedestrians(self, field, new_pedestrians: Sequence[Pedestrian], desired_speeds=None) -> None:
        """mod.rs:19 / sfm.rs:48-89: append, then rebuild the neighbor grid (despawn + reorder).

        `desired_speeds` are the N(1.34, 0.26) draws of sfm.rs:54, made by the caller (the
        Simulator's seeded stream) so that device code has no RNG."""
        n = len(new_pedestrians)
        if n:
            pos = _f32([p.pos for p in new_pedestrians]).reshape(n, 2)
            dest = np.ascontiguousarray([p.destination for p in new_pedestrians], dtype=np.uint32)
            if desired_speeds is None:
                raise ValueError("desired_speeds must accompany spawned pedestrians")
            self.spawn_arrays(pos, dest, _f32(desired_speeds))
        self.rebuild()

    def update_states(self, scenario=None, field=None) -> None:
        """mod.rs:21 / sfm.rs:91-255."""
        _capi.check(self._lib.pedoni_step(self._h), self._h)

    def list_pedestrians(self) -> list:
        """mod.rs:23 / sfm.rs:257-265."""
        pos, dest = self.download(vel=False, v0=False)[:2]
        return [Pedestrian(pos=(float(p[0]), float(p[1])), destination=int(d)) for p, d in zip(pos, dest)]

    def get_pedestrian_count(self) -> int:
        """mod.rs:25 / sfm.rs:267-269."""
        return _capi.check(self._lib.pedoni_count(self._h), self._h)

    def count_published(self):
        """(population, rebuild ordinal) as last published by the device; never blocks."""
        n, t = C.c_int32(), C.c_uint32()
        _capi.check(self._lib.pedoni_count_published(self._h, C.byref(n), C.byref(t)), self._h)
        return n.value, t.value

    # -- array-level entry points (what the trait methods are made of) ----------------------------
    def spawn_arrays(self, pos: np.ndarray, dest: np.ndarray, v0: np.ndarray) -> None:
        pos, v0 = _f32(pos), _f32(v0)
        dest = np.ascontiguousarray(dest, dtype=np.uint32)
        n = dest.shape[0]
        assert pos.size == 2 * n and v0.size == n
        _capi.check(self._lib.pedoni_spawn(self._h, n, _fp(pos), _up(dest), _fp(v0)), self._h)

    def spawn_groups(self, groups, seed: int, counter: int) -> int:
        """Device-side spawn (pedoni_spawn_groups). groups: [(p1, p2, destination, count)]. Returns how many
        stream numbers were consumed (3 per pedestrian)."""
        arr = (_capi.PedoniSpawnGroup * len(groups))(*[
            _capi.PedoniSpawnGroup(float(np.float32(p1[0])), float(np.float32(p1[1])), float(np.float32(p2[0])),
                                   float(np.float32(p2[1])), int(d), int(c)) for p1, p2, d, c in groups])
        _capi.check(self._lib.pedoni_spawn_groups(self._h, len(groups), arr, C.c_uint64(seed & (2 ** 64 - 1)),
                                                  C.c_uint64(counter)), self._h)
        return 3 * sum(int(g[3]) for g in groups)

    def spawn_stream_seek(self, seed: int, counter: int) -> None:
        _capi.check(self._lib.pedoni_spawn_stream_seek(self._h, C.c_uint64(seed & (2 ** 64 - 1)), C.c_uint64(counter)),
                    self._h)

    def spawn_stream_tell(self):
        """(stream position, pedestrians drawn so far) of the device-side arrivals; blocks."""
        k, n = C.c_uint64(), C.c_uint64()
        _capi.check(self._lib.pedoni_spawn_stream_tell(self._h, C.byref(k), C.byref(n)), self._h)
        return k.value, n.value

    def spawn_poisson(self, rates) -> None:
        """Device-side arrivals of one tick (pedoni_spawn_poisson). rates: [(p1, p2, destination, frequency)]."""
        arr = (_capi.PedoniSpawnRate * len(rates))(*[
            _capi.PedoniSpawnRate(float(np.float32(p1[0])), float(np.float32(p1[1])), float(np.float32(p2[0])),
                                  float(np.float32(p2[1])), int(d), float(f)) for p1, p2, d, f in rates])
        _capi.check(self._lib.pedoni_spawn_poisson(self._h, len(rates), arr), self._h)

    def rebuild(self) -> None:
        _capi.check(self._lib.pedoni_rebuild(self._h), self._h)

    def step(self) -> None:
        _capi.check(self._lib.pedoni_step(self._h), self._h)

    def upload_state(self, pos, dest, vel, v0) -> None:
        pos, vel, v0 = _f32(pos), _f32(vel), _f32(v0)
        dest = np.ascontiguousarray(dest, dtype=np.uint32)
        n = dest.shape[0]
        assert pos.size == 2 * n and vel.size == 2 * n and v0.size == n
        _capi.check(self._lib.pedoni_upload_state(self._h, n, _fp(pos), _up(dest), _fp(vel), _fp(v0)), self._h)

    def download(self, vel: bool = True, v0: bool = True, out=None):
        """Returns (pos[n,2], dest[n], vel[n,2] | None, v0[n] | None) in the model's current order."""
        n = self.get_pedestrian_count()
        if out is None:
            pos = np.empty((n, 2), np.float32)
            dest = np.empty(n, np.uint32)
            velo = np.empty((n, 2), np.float32) if vel else None
            v0o = np.empty(n, np.float32) if v0 else None
        else:
            pos, dest, velo, v0o = out
        n_out = C.c_uint32()
        _capi.check(self._lib.pedoni_download(self._h, _fp(pos), _up(dest), _fp(velo), _fp(v0o), dest.shape[0],
                                              C.byref(n_out)), self._h)
        k = n_out.value
        return pos[:k], dest[:k], (velo[:k] if velo is not None else None), (v0o[:k] if v0o is not None else None)

    def download_begin(self, pos: np.ndarray, dest: np.ndarray) -> None:
        """Pipelined `list_pedestrians`: snapshot on the device, copy to (pinned) `pos`/`dest` in the background.
        `dest` may be uint32 (the trait's type) or uint8 (destinations travel and arrive as bytes)."""
        if dest.dtype == np.uint8:
            _capi.check(self._lib.pedoni_download_begin_u8(self._h, _fp(pos), dest.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                           dest.shape[0]), self._h)
        else:
            _capi.check(self._lib.pedoni_download_begin(self._h, _fp(pos), _up(dest), dest.shape[0]), self._h)
        if not hasattr(self, "_dl"):
            self._dl = []
        self._dl.append((pos, dest))  # keep the buffers alive until download_end; up to two may be in flight

    def download_end(self):
        """Completes the OLDEST pipelined download in flight."""
        n = C.c_uint32()
        _capi.check(self._lib.pedoni_download_end(self._h, C.byref(n)), self._h)
        pos, dest = self._dl.pop(0)
        return pos[: n.value], dest[: n.value]

    def observe(self, y_range=(0.0, 1.0), bins: int = 0) -> dict:
        """Device-side observables (pedoni_observe): no pedestrian leaves the GPU."""
        o = _capi.PedoniObservables()
        _capi.check(self._lib.pedoni_observe(self._h, float(y_range[0]), float(y_range[1]), int(bins), C.byref(o)),
                    self._h)
        return {"count": o.count, "mean_speed": o.mean_speed,
                "per_destination": np.array(o.per_destination[:], np.uint32),
                "arrived": np.array(o.arrived[:], np.uint64),
                "bin_count": np.array(o.bin_count[:bins], np.uint32),
                "bin_mean_vx": np.array(o.bin_mean_vx[:bins], np.float32)}

    def grid_shape(self):
        ny, nx = C.c_int32(), C.c_int32()
        _capi.check(self._lib.pedoni_grid_shape(self._h, C.byref(ny), C.byref(nx)), self._h)
        return ny.value, nx.value

    def cell_table(self) -> np.ndarray:
        """`neighbor_grid_indices` (sfm.rs:22) of the last rebuild."""
        ny, nx = self.grid_shape()
        buf = np.empty(ny * nx + 1, np.uint32)
        n = C.c_uint32()
        _capi.check(self._lib.pedoni_cell_table(self._h, _up(buf), buf.shape[0], C.byref(n)), self._h)
        return buf[: n.value]

    def synchronize(self) -> None:
        _capi.check(self._lib.pedoni_synchronize(self._h), self._h)

    # -- measurement ----------------------------------------------------------------------------
    def profile_enable(self, on: bool = True) -> None:
        _capi.check(self._lib.pedoni_profile_enable(self._h, int(on)), self._h)

    def profile_reset(self) -> None:
        _capi.check(self._lib.pedoni_profile_reset(self._h), self._h)

    def profile_read(self) -> dict:
        t = _capi.PedoniKernelTimes()
        _capi.check(self._lib.pedoni_profile_read(self._h, C.byref(t)), self._h)
        return {name: getattr(t, name) for name, _ in t._fields_}

    KIND_NAMES = {0: "key", 4: "sort", 5: "force", 6: "exchange+unpack", 7: "force_edge", 8: "pack"}

    def profile_timeline(self) -> list:
        """[(kernel, stream, start_ms, stop_ms)] of the launches timed since the last timer_begin (profiling on)."""
        n = C.c_uint32()
        _capi.check(self._lib.pedoni_profile_timeline(self._h, None, 0, C.byref(n)), self._h)
        buf = (_capi.PedoniLaunchRecord * max(n.value, 1))()
        _capi.check(self._lib.pedoni_profile_timeline(self._h, buf, n.value, C.byref(n)), self._h)
        return [(self.KIND_NAMES.get(r.kind, str(r.kind)), "edge" if r.stream else "main", r.start_ms, r.stop_ms)
                for r in buf[: n.value]]

    def counters(self):
        """(kernel launches, pedestrian-updates) since creation."""
        a, b = C.c_uint64(), C.c_uint64()
        _capi.check(self._lib.pedoni_counters(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def timer_begin(self) -> None:
        _capi.check(self._lib.pedoni_timer_begin(self._h), self._h)

    def timer_end(self) -> float:
        ms = C.c_float()
        _capi.check(self._lib.pedoni_timer_end(self._h, C.byref(ms)), self._h)
        return ms.value

    # -- slabs ------------------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes) -> None:
        assert len(unique_id) == _capi.PEDONI_COMM_ID_BYTES
        buf = C.create_string_buffer(unique_id, _capi.PEDONI_COMM_ID_BYTES)
        _capi.check(self._lib.pedoni_comm_init(self._h, buf), self._h)

    def slab_transport(self) -> str:
        return self._lib.pedoni_slab_transport(self._h).decode()

    def download_wire_bytes(self) -> int:
        """Bytes per pedestrian the pipelined download moves over PCIe (12, or 9 with byte-sized destinations)."""
        return int(self._lib.pedoni_download_wire_bytes(self._h))

    def field_textures(self) -> bool:
        """True if the force kernel fetches the field maps with texture gathers (fast math only)."""
        return bool(self._lib.pedoni_field_textures(self._h))

    def wall_far_cells(self):
        """(marked, total) blocks of 8 x 8 distance-map texels in which the fast-math force kernel skips the wall
        term (pedoni_wall_far_cells: more than 8 m from every obstacle, no ridge of the distance map)."""
        far, total = C.c_uint64(), C.c_uint64()
        _capi.check(self._lib.pedoni_wall_far_cells(self._h, C.byref(far), C.byref(total)), self._h)
        return int(far.value), int(total.value)

    def halo_capacity(self) -> int:
        h = C.c_uint32()
        _capi.check(self._lib.pedoni_halo_capacity(self._h, C.byref(h)), self._h)
        return h.value

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.pedoni_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SlabGroup:
    """`slab_count` slab handles living in ONE process, exchanging ghost rows with the in-process
    transport (pedoni_slab_exchange_local). Presents the same trait-shaped surface as a single
    `SocialForceModelCuda`; outputs are the rank-order concatenation, which equals the whole-domain
    order. One-process-per-GPU programs (bench.py under torchrun) use one handle + `comm_init` instead."""

    def __init__(self, options, scenario, field, slab_count: int, *, devices=None, **kw):
        devices = devices or [kw.pop("device", 0)] * slab_count
        kw.pop("device", None)
        self.slabs = [SocialForceModelCuda(options, scenario, field, device=devices[r], slab_rank=r,
                                           slab_count=slab_count, **kw) for r in range(slab_count)]
        self._lib = _capi.load()

    def _exchange(self) -> None:
        arr = (C.c_void_p * len(self.slabs))(*[s._h.value for s in self.slabs])
        _capi.check(self._lib.pedoni_slab_exchange_local(arr, len(self.slabs)), None)

    def spawn_arrays(self, pos, dest, v0) -> None:  # replicated list; each slab keeps the rows it owns
        for s in self.slabs:
            s.spawn_arrays(pos, dest, v0)

    def spawn_groups(self, groups, seed: int, counter: int) -> int:
        return [s.spawn_groups(groups, seed, counter) for s in self.slabs][0]

    def spawn_stream_seek(self, seed: int, counter: int) -> None:
        for s in self.slabs:
            s.spawn_stream_seek(seed, counter)

    def spawn_stream_tell(self):
        return [s.spawn_stream_tell() for s in self.slabs][0]  # every slab draws the same (replicated) arrivals

    def spawn_poisson(self, rates) -> None:
        for s in self.slabs:
            s.spawn_poisson(rates)

    def upload_state(self, pos, dest, vel, v0) -> None:
        for s in self.slabs:
            s.upload_state(pos, dest, vel, v0)

    def rebuild(self) -> None:
        for s in self.slabs:
            s.rebuild()
        self._exchange()

    def step(self) -> None:
        for s in self.slabs:
            s.step()

    update_states = step

    def get_pedestrian_count(self) -> int:
        return sum(s.get_pedestrian_count() for s in self.slabs)

    def download(self, vel: bool = True, v0: bool = True):
        parts = [s.download(vel=vel, v0=v0) for s in self.slabs]
        cat = lambda k: None if parts[0][k] is None else np.concatenate([p[k] for p in parts])  # noqa: E731
        return cat(0), cat(1), cat(2), cat(3)

    def observe(self, y_range=(0.0, 1.0), bins: int = 0) -> dict:
        """Device-side observables of every slab, combined (counts add; means are population-weighted)."""
        parts = [s.observe(y_range, bins) for s in self.slabs]
        count = sum(p["count"] for p in parts)
        bin_count = sum(p["bin_count"].astype(np.int64) for p in parts)
        vx_sum = sum(p["bin_mean_vx"].astype(np.float64) * p["bin_count"] for p in parts)
        return {"count": count,
                "mean_speed": float(sum(p["mean_speed"] * p["count"] for p in parts) / max(count, 1)),
                "per_destination": sum(p["per_destination"].astype(np.int64) for p in parts),
                "arrived": sum(p["arrived"].astype(np.int64) for p in parts),
                "bin_count": bin_count,
                "bin_mean_vx": np.where(bin_count > 0, vx_sum / np.maximum(bin_count, 1), 0.0).astype(np.float32)}

    def cell_table(self) -> np.ndarray:
        """Whole-domain `neighbor_grid_indices` stitched from the slabs' owned rows."""
        out, base = [np.zeros(1, np.uint32)], 0
        for s in self.slabs:
            t = s.cell_table().astype(np.uint64)
            out.append((t[1:] + base).astype(np.uint32))
            base += int(t[-1])
        return np.concatenate(out)

    def synchronize(self) -> None:
        for s in self.slabs:
            s.synchronize()

    def counters(self):
        c = [s.counters() for s in self.slabs]
        return sum(x[0] for x in c), sum(x[1] for x in c)

    def close(self) -> None:
        for s in self.slabs:
            s.close()


def comm_unique_id() -> bytes:
    lib = _capi.load()
    buf = C.create_string_buffer(_capi.PEDONI_COMM_ID_BYTES)
    _capi.check(lib.pedoni_comm_unique_id(buf))
    return buf.raw


def slab_rows(ny: int, count: int, rank: int):
    lib = _capi.load()
    r0, r1 = C.c_int32(), C.c_int32()
    _capi.check(lib.pedoni_slab_rows(ny, count, rank, C.byref(r0), C.byref(r1)))
    return r0.value, r1.value
