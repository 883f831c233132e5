"""CPU: the Simulator facade (lib.rs:27-105 mirror) and the observables, driven with the oracle model."""
import numpy as np

import helpers
from pedoni_b200 import SimulatorOptions, observables
from pedoni_b200.simulator import Simulator, SpawnStream


def test_spawn_stream_is_reproducible_and_sane():
    a, b = SpawnStream(7), SpawnStream(7)
    assert np.array_equal(a.f32(100), b.f32(100))
    u = SpawnStream(1).f32(20000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    z = SpawnStream(2).normal_approx(50000, 1.34, 0.26)
    assert abs(z.mean() - 1.34) < 0.01 and abs(z.std() - 0.26) < 0.01
    s = SpawnStream(3)
    k = np.array([s.poisson(0.4) for _ in range(20000)])
    assert abs(k.mean() - 0.4) < 0.02 and abs(k.var() - 0.4) < 0.03


def test_narrow_gap_evacuates_and_once_groups_spawn_at_construction():
    sc = helpers.load_scenario("narrow-gap")
    opts = SimulatorOptions()
    field = helpers.oracle_field(sc, opts.field_grid_unit)
    sim = Simulator(opts, sc, field, helpers.OracleAdapter(opts, sc, field), seed=3)
    assert sim.model.get_pedestrian_count() == 50 and sim.spawned_total == 50  # lib.rs:37-52
    log = sim.run(2000, until_empty=True)
    t = observables.evacuation_time(log.active_ped_count)
    assert t is not None and 10.0 < t < 120.0, t
    assert log.active_ped_count[-1] == 0 and max(log.active_ped_count) == 50


def test_periodic_spawn_rate_and_flow():
    sc = helpers.load_scenario("lanes")  # 1.04 + 1.04 pedestrians/s
    opts = SimulatorOptions()
    field = helpers.oracle_field(sc, opts.field_grid_unit)
    sim = Simulator(opts, sc, field, helpers.OracleAdapter(opts, sc, field), seed=5)
    spawned = []
    log_counts = []
    for _ in range(1500):
        m = sim.tick()
        spawned.append(sim.spawned_total)
        log_counts.append(m.active_ped_count)
    rate = sim.spawned_total / 150.0
    assert 1.6 < rate < 2.6, rate
    # steady state: what walks in walks out
    flow = observables.flow_rate(log_counts, spawned, window=(0.6, 1.0))
    assert 1.2 < flow < 3.0, flow
    pos, dest, vel, _ = sim.model.download()
    assert observables.lane_count(pos, vel, (0.0, 8.0), bins=8, min_agents=1) >= 1
    assert 0.5 < observables.mean_speed(vel) < 2.0
