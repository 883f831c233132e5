"""GPU: the field precompute on the device (second half) against the host builder, which
restates the reference's marching (field.rs:118-192) and is bit-equal to the oracle's (tests/test_field_builder.py).

pedoni_field_build_device solves the per-cell upwind update as a fixed point (block-iterative, csrc/field_device.cu),
stated so that its fixed point is the marching result. Stated tolerance, on every shipped scenario:
  obstacle mask, waypoint outlines   identical
  distance map; potential maps on    |device - host| <= 2e-3 field cells + 2e-4 relative: rounding only — a potential is a
  the cells one can stand on         sum of thousands of fp32 steps, formed in a different order (measured: <= 2e-3 cells
                                     on maps up to 800^2, 0.5 cells = 1e-4 relative on default10's 4000^2)
  potential INSIDE obstacles         same magnitude (factor 2): a wall cell costs 1e6 cells, where the reference's
                                     two-axis test `2 f^2 - (u1 - u2)^2 >= 0` and the causal one differ (csrc/field_device.cu);
                                     pedestrians only ever see these values as "through a wall" (force.cuh)
and a scenario run on the device-built field gives the same aggregate observables as on the host-built one."""
import numpy as np
import pytest

import helpers
from pedoni_b200 import PEDONI_MATH_FAST, Field, SimulatorOptions, SocialForceModelCuda, observables
from pedoni_b200.simulator import Simulator

pytestmark = pytest.mark.gpu

UNIT = 0.25
TOL_CELLS, TOL_REL = 2e-3, 2e-4


@pytest.mark.parametrize("name", helpers.scenario_names())
def test_device_field_matches_host_builder(name):
    sc = helpers.load_scenario(name)
    host = Field.from_scenario(sc, UNIT)
    dev = Field.from_scenario(sc, UNIT, device=0)
    np.testing.assert_array_equal(host.obstacle_exist, dev.obstacle_exist)
    free = ~host.obstacle_exist
    close = lambda g, h: (np.abs(g - h) <= TOL_CELLS * UNIT + TOL_REL * np.abs(h)).all()  # noqa: E731
    assert close(dev.distance_map, host.distance_map), np.abs(dev.distance_map - host.distance_map).max()
    assert (dev.distance_map[~free] == 0).all()
    identical = int(np.array_equal(dev.distance_map, host.distance_map))
    for k in range(host.potential_maps.shape[0]):
        h, g = host.potential_maps[k], dev.potential_maps[k]
        assert close(g[free], h[free]), f"{name}: potential map {k}: {np.abs(g - h)[free].max()}"
        wall = ~free & (h > 0)
        assert ((g[wall] >= 0.5 * h[wall]) & (g[wall] <= 2.0 * h[wall])).all(), f"{name}: map {k} inside obstacles"
        np.testing.assert_array_equal(g == 0, h == 0)  # the waypoint outline itself
        identical += int(np.array_equal(g[free], h[free]))
    print(f"{name}: {identical} of {1 + host.potential_maps.shape[0]} maps bit-identical to the host builder's off the obstacles")


def test_scenario_run_on_device_built_field_gives_the_same_observables():
    """narrow-gap.toml (50 pedestrians through a 3 m gap): evacuation time over 12 seeds on the device-built field vs
    the host-built field, fast math; means within 2 standard errors + one tick."""
    sc = helpers.load_scenario("narrow-gap")
    opts = SimulatorOptions()
    fields = [Field.from_scenario(sc, opts.field_grid_unit), Field.from_scenario(sc, opts.field_grid_unit, device=0)]
    times = ([], [])
    for seed in range(12):
        for k, field in enumerate(fields):
            sim = Simulator(opts, sc, field, SocialForceModelCuda(opts, sc, field, math_mode=PEDONI_MATH_FAST), seed=500 + seed)
            n0 = sim.model.get_pedestrian_count()
            counts = sim.run(3000, until_empty=True).active_ped_count
            times[k].append(observables.evacuation_time(counts, fraction=1.0, initial=n0))
            sim.model.close()
    assert None not in times[0] and None not in times[1]
    a, b = np.asarray(times[0], float), np.asarray(times[1], float)
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    print(f"narrow-gap evacuation time: host-built field {a.mean():.2f} s, device-built {b.mean():.2f} s")
    assert abs(a.mean() - b.mean()) <= 2 * se + observables.DT
