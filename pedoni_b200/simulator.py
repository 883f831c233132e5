"""`Simulator` — harness-side mirror of the reference's step API (pedoni-simulator/src/lib.rs:16-105):
`Simulator::new` (spawn the "once" groups), `tick` (Poisson spawns -> model.spawn_pedestrians ->
model.update_states -> StepMetrics) and `list_pedestrians`.

The reference draws its random numbers from fastrand's global, never-seeded generator (lib.rs:43,76;
util.rs:80,84; sfm.rs:54), so no two runs of it agree. Here every draw comes from `SpawnStream`, a
counter-based seeded generator on the CALLER side of the plugin boundary: the model (CUDA or oracle)
receives positions, destinations and desired speeds as inputs, which is what makes the two comparable
bit for bit.

The `Field` is an input as well (lib.rs:30 builds it once before the model exists); pass the arrays the
reference's `Field::from_scenario` produces, or any other builder's.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field as dc_field
from typing import List, Optional

import numpy as np

from .synthetic import splitmix64

_MASK32 = np.uint64(0xFFFFFFFF)


class SpawnStream:
    """Seeded stand-in for the fastrand calls on the caller side of the plugin.

    u64 draw k of the stream is splitmix64(seed ^ k) — reproducible, cheap, independent of call batching."""

    def __init__(self, seed: int = 0x5EED0001):
        self.seed = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        self.k = 0

    def _u64(self, n: int) -> np.ndarray:
        ctr = np.arange(self.k, self.k + n, dtype=np.uint64)
        self.k += n
        with np.errstate(over="ignore"):
            return splitmix64(self.seed ^ (ctr * np.uint64(0x2545F4914F6CDD1D)))

    def f32(self, n: int) -> np.ndarray:
        """fastrand::f32(): uniform in [0, 1) with 24 random mantissa bits."""
        return ((self._u64(n) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))).astype(np.float32)

    def f64(self) -> float:
        """fastrand::f64(): uniform in [0, 1) with 53 random bits."""
        return float(self._u64(1)[0] >> np.uint64(11)) * (1.0 / (1 << 53))

    def poisson(self, lam: float) -> int:
        """util.rs:78-89 (Knuth's multiplication method), same control flow."""
        y = 0
        x = self.f64()
        exp_lambda = math.exp(-lam)
        while x >= exp_lambda:
            x *= self.f64()
            y += 1
        return y

    def normal_approx(self, n: int, mu: float, sigma: float) -> np.ndarray:
        """fastrand_contrib::f32_normal_approx(mu, sigma) (sfm.rs:54): the popcount + triangular
        approximation of a standard normal (binomial(64, 1/2) centred, plus the difference of two
        32-bit uniforms), scaled to unit variance: var = 16 + 1/6."""
        a = self._u64(n)
        b = self._u64(n)
        # popcount via bytes
        pop = np.unpackbits(a.view(np.uint8).reshape(-1, 8), axis=1).sum(1).astype(np.float64) - 32.0
        tri = ((b & _MASK32).astype(np.float64) - (b >> np.uint64(32)).astype(np.float64)) * (1.0 / (1 << 32))
        z = (pop + tri) * (1.0 / math.sqrt(16.0 + 1.0 / 6.0))
        return (np.float32(mu) + np.float32(sigma) * z.astype(np.float32)).astype(np.float32)


@dataclass
class StepMetrics:
    """diagnostic.rs:45-50."""
    active_ped_count: int
    time_spawn: float
    time_calc_state: float
    time_calc_state_kernel: Optional[float] = None


@dataclass
class StepMetricsCollection:
    """diagnostic.rs:22-37: the four per-step series of the reference's headless JSON log."""
    active_ped_count: List[int] = dc_field(default_factory=list)
    time_spawn: List[float] = dc_field(default_factory=list)
    time_calc_state: List[float] = dc_field(default_factory=list)
    time_calc_state_kernel: List[Optional[float]] = dc_field(default_factory=list)

    def push(self, m: StepMetrics) -> None:
        self.active_ped_count.append(m.active_ped_count)
        self.time_spawn.append(m.time_spawn)
        self.time_calc_state.append(m.time_calc_state)
        self.time_calc_state_kernel.append(m.time_calc_state_kernel)


class Simulator:
    """lib.rs:16-105. `model` is any `PedestrianModel`-shaped object: `SocialForceModelCuda`, `SlabGroup`, or
    the tests' oracle adapter. It must offer spawn_arrays(pos, dest, v0) + rebuild() (the two halves of
    `spawn_pedestrians`), step() (`update_states`), get_pedestrian_count() and download()."""

    def __init__(self, options, scenario, field, model, seed: int = 0x5EED0001, count_every: int = 1,
                 device_spawn: bool = False):
        self.options, self.scenario, self.field, self.model = options, scenario, field, model
        self.step = 0
        self.rng = SpawnStream(seed)
        self.spawned_total = 0
        # draw positions and desired speeds on the device (pedoni_spawn_groups) instead of uploading them;
        # same stream numbers either way, so the two modes give identical runs
        self.device_spawn = bool(device_spawn) and hasattr(model, "spawn_groups")
        # device_spawn="poisson": the per-group Poisson counts are drawn on the device as well
        # (pedoni_spawn_poisson); the model then owns the stream position, and this object's `rng.k` /
        # `spawned_total` are refreshed from it by sync_spawn_stream()
        self.device_poisson = device_spawn == "poisson" and hasattr(model, "spawn_poisson")  # (any other true value: positions)
        self._rates = None
        self.count_every = max(1, int(count_every))  # counting blocks; headless runs may sample it
        self._last_count = 0
        # lib.rs:37-52: "once" groups are spawned at construction
        self._spawn(kind="once")

    def _draw_group(self, ped, count: int):
        (p1, p2) = self.scenario.waypoints[ped.origin].line
        t = self.rng.f32(count)[:, None]
        p1 = np.asarray(p1, np.float32)[None, :]
        p2 = np.asarray(p2, np.float32)[None, :]
        # glam lerp: self + (rhs - self) * s   (f32)
        pos = (p1 + (p2 - p1) * t).astype(np.float32)
        dest = np.full(count, ped.destination, np.uint32)
        return pos, dest

    def _spawn(self, kind: str) -> int:
        """Stream order per call: every group's count first (Poisson draws), then one uniform per pedestrian,
        group after group, then the desired speeds (sfm.rs:54, in push order) — the same on the host path
        and on the device path (pedoni_spawn_groups), so the two give bit-identical runs."""
        if kind == "periodic" and self.device_poisson:
            if self._rates is None:  # first periodic tick: hand the stream over to the device
                self._rates = [(*self.scenario.waypoints[p.origin].line, p.destination, float(p.spawn.frequency))
                               for p in self.scenario.pedestrians if p.spawn.kind == "periodic"]
                self.model.spawn_stream_seek(int(self.rng.seed), self.rng.k)
                self._spawned_before_device = self.spawned_total
            if self._rates:
                self.model.spawn_poisson(self._rates)
            self.model.rebuild()
            return 0
        groups = []
        for ped in self.scenario.pedestrians:
            if ped.spawn.kind != kind:
                continue
            # lib.rs:74: poisson(frequency / 10.0) new pedestrians per 0.1 s tick
            count = ped.spawn.count if kind == "once" else self.rng.poisson(ped.spawn.frequency / 10.0)
            if count > 0:
                groups.append((ped, count))
        n = sum(c for _, c in groups)
        if n and self.device_spawn:
            table = [(*self.scenario.waypoints[ped.origin].line, ped.destination, c) for ped, c in groups]
            self.rng.k += self.model.spawn_groups(table, int(self.rng.seed), self.rng.k)
        elif n:
            drawn = [self._draw_group(ped, c) for ped, c in groups]
            pos, dest = np.concatenate([d[0] for d in drawn]), np.concatenate([d[1] for d in drawn])
            v0 = self.rng.normal_approx(n, 1.34, 0.26)
            self.model.spawn_arrays(pos, dest, v0)
        self.model.rebuild()  # spawn_pedestrians rebuilds the grid even with no newcomers (lib.rs:85, sfm.rs:58)
        self.spawned_total += n
        return n

    def tick(self) -> StepMetrics:
        self.step += 1
        t0 = time.perf_counter()
        self._spawn(kind="periodic")
        t1 = time.perf_counter()
        self.model.step()
        if self.step % self.count_every == 0:
            self._last_count = self.model.get_pedestrian_count()
        t2 = time.perf_counter()
        return StepMetrics(self._last_count, t1 - t0, t2 - t1, None)

    def sync_spawn_stream(self) -> None:
        """device_spawn="poisson": read the stream position and the number of arrivals back from the device (blocks)."""
        if self.device_poisson and self._rates is not None:
            self.rng.k, drawn = self.model.spawn_stream_tell()
            self.spawned_total = self._spawned_before_device + drawn

    def list_pedestrians(self):
        return self.model.list_pedestrians()

    def run(self, max_steps: int, until_empty: bool = False) -> StepMetricsCollection:
        """main.rs:81-104 headless loop without its rate limiter; `until_empty` is the reference's
        commented-out evacuation experiment (main.rs:58-77): stop at the first tick with nobody left."""
        log = StepMetricsCollection()
        for _ in range(max_steps):
            m = self.tick()
            log.push(m)
            if until_empty and m.active_ped_count <= 0:
                break
        return log
