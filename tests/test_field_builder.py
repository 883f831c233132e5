"""CPU: the product's host field builder (csrc/host/field_builder.cpp behind pedoni_field_build) against the
oracle's restatement of field.rs:16-192 — bit for bit on the shipped scenarios — and the closed-form
open-domain field of pedoni_b200/synthetic.py against both."""
import numpy as np
import pytest

import helpers
from pedoni_b200 import Field
from pedoni_b200.synthetic import SyntheticCrowd


@pytest.mark.parametrize("name", ["default", "narrow-gap", "bottleneck", "evacuation", "lanes", "random", "straight"])
def test_product_field_builder_equals_oracle(name):
    sc = helpers.load_scenario(name)
    want = helpers.oracle_field(sc, 0.25)
    got = Field.from_scenario(sc, 0.25)
    assert got.shape == tuple(want.distance_map.shape)
    np.testing.assert_array_equal(got.obstacle_exist, want.obstacle_exist)
    np.testing.assert_array_equal(helpers.bits(got.distance_map), helpers.bits(want.distance_map))
    np.testing.assert_array_equal(helpers.bits(got.potential_maps), helpers.bits(want.potential_maps))


def test_other_units_and_degenerate_inputs():
    sc = helpers.load_scenario("narrow-gap")
    for unit in (0.5, 0.2):
        want, got = helpers.oracle_field(sc, unit), Field.from_scenario(sc, unit)
        np.testing.assert_array_equal(helpers.bits(got.distance_map), helpers.bits(want.distance_map))
        np.testing.assert_array_equal(helpers.bits(got.potential_maps), helpers.bits(want.potential_maps))
    empty = helpers.scenario_of((10.0, 6.0))  # no obstacles, no waypoints: border ring only
    f = Field.from_scenario(empty, 0.25)
    assert f.potential_maps.shape == (0, 24, 40) and f.obstacle_exist[0].all() and f.distance_map[12, 20] > 0


def test_closed_form_synthetic_field_matches_the_builders():
    """bench.py's open-domain field (pedoni_b200/synthetic.py) away from the border ring."""
    crowd = SyntheticCrowd(n=400)  # 16 cells -> 22.4 m -> 90 x 90 texels
    got = crowd.field()
    want = Field.from_scenario(crowd.scenario(), crowd.field_unit)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got.obstacle_exist, want.obstacle_exist)
    # (the marching rounds the corners where the waypoint outline meets the border ring: skip 4 texels = 1 m;
    #  the synthetic crowd is seeded 2 m inside and despawns at the waypoint before it gets there)
    inner = (slice(4, -4), slice(4, -4))
    np.testing.assert_allclose(got.potential_maps[:, inner[0], inner[1]], want.potential_maps[:, inner[0], inner[1]],
                               rtol=0, atol=5e-3)  # 2 % of a texel
    # Distance map: the closed form is the exact distance to the ring along the axes; the marching adds its
    # first-order error where two fronts meet (up to 0.1 m around the diagonals and the centre). The wall
    # force is 2 exp(-d / 0.2): it is 3e-7 m/s^2 at d = 3 m, so only the first 3 m from the ring matter.
    c = got.shape[0] // 2
    near = slice(1, 13)
    np.testing.assert_allclose(got.distance_map[c, near], want.distance_map[c, near], rtol=0, atol=1e-4)
    np.testing.assert_allclose(got.distance_map[near, c], want.distance_map[near, c], rtol=0, atol=1e-4)
    assert np.abs(got.distance_map - want.distance_map)[4:-4, 4:-4].max() < 0.15
