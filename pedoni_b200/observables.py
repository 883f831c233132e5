"""Aggregate observables named by BASELINE.json's north star. The reference computes none of them (its
only whole-run measure is the commented-out loop at main.rs:58-77 counting ticks until
`active_ped_count <= 0`); they are defined here on the reference's own per-step series
(`active_ped_count`, diagnostic.rs:22-27) plus downloaded state, so that any backend can be compared.
"""
from __future__ import annotations

import numpy as np

DT = 0.1  # s per tick (main.rs:28 DELTA_TIME, sfm.rs:249)


def evacuation_time(active_ped_count, fraction: float = 1.0, initial=None, dt: float = DT):
    """Seconds until the first tick at which `fraction` of the initial population has left; with the
    default 1.0 this is the reference's own experiment, ticks until `active_ped_count <= 0`
    (main.rs:58-77). None if never reached."""
    a = np.asarray(active_ped_count)
    n0 = max(a) if initial is None else initial
    hit = np.nonzero(a <= (1.0 - fraction) * n0 + 1e-9)[0]
    return None if len(hit) == 0 else float(hit[0] + 1) * dt


def flow_rate(active_ped_count, spawned_cumulative, window=(0.5, 1.0), dt: float = DT) -> float:
    """Pedestrians per second leaving the scene (despawned at their destination, sfm.rs:69) averaged over
    the given fraction of the run: d/dt (spawned_total - active)."""
    a = np.asarray(active_ped_count, np.float64)
    s = np.asarray(spawned_cumulative, np.float64)
    out = s - a
    i0, i1 = int(window[0] * (len(a) - 1)), int(window[1] * (len(a) - 1))
    if i1 <= i0:
        return 0.0
    return float(out[i1] - out[i0]) / ((i1 - i0) * dt)


def lane_count(pos, vel, y_range, bins: int = 32, min_agents: int = 2) -> int:
    """Number of lanes in a counter-flow corridor: sign changes of the mean x-velocity across y-bins
    (+1). Bins with fewer than `min_agents` pedestrians are skipped."""
    pos, vel = np.asarray(pos), np.asarray(vel)
    if len(pos) == 0:
        return 0
    edges = np.linspace(y_range[0], y_range[1], bins + 1)
    idx = np.clip(np.digitize(pos[:, 1], edges) - 1, 0, bins - 1)
    signs = []
    for b in range(bins):
        m = idx == b
        if m.sum() >= min_agents:
            v = vel[m, 0].mean()
            if v != 0:
                signs.append(np.sign(v))
    if not signs:
        return 0
    signs = np.asarray(signs)
    return int((signs[1:] != signs[:-1]).sum()) + 1


def lane_count_from_bins(bin_mean_vx, bin_count, min_agents: int = 2) -> int:
    """`lane_count` on the histogram pedoni_observe reduces on the device."""
    signs = [np.sign(v) for v, c in zip(bin_mean_vx, bin_count) if c >= min_agents and v != 0]
    if not signs:
        return 0
    signs = np.asarray(signs)
    return int((signs[1:] != signs[:-1]).sum()) + 1


def mean_speed(vel) -> float:
    vel = np.asarray(vel)
    return float(np.linalg.norm(vel, axis=1).mean()) if len(vel) else 0.0
