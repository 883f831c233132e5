"""Scenario TOML: the reference's serde structs (scenario.rs:10-66) and its shipped scenario files."""
import numpy as np
import pytest

import helpers
from pedoni_b200 import Scenario

LANES_LIKE = """
[field]
size = [60, 30]
unit = 0.5            # unknown key: serde ignores it (random.toml:3 carries one)

[[waypoints]]
line = [[5, 0.5], [5, 7.5]]

[[waypoints]]
line = [[55, 0.5], [55, 7.5]]
width = 2.0

[[obstacles]]
line = [[0, 0], [60, 0]]
width = 0.01

[[obstacles]]
line = [[0, 8], [60, 8]]

[[pedestrians]]
origin = 0
destination = 1
spawn = { kind = "periodic", frequency = 1.04 }

[[pedestrians]]
origin = 1
destination = 0
spawn = { kind = "once", count = 25 }
"""


def test_parse_matches_serde_shapes():
    sc = Scenario.from_toml_str(LANES_LIKE)
    assert sc.field.size == (60.0, 30.0)
    assert [w.width for w in sc.waypoints] == [1.0, 2.0]          # default width 1.0 (scenario.rs:4-6)
    assert [o.width for o in sc.obstacles] == [0.01, 1.0]
    assert sc.obstacles[1].line == ((0.0, 8.0), (60.0, 8.0))
    assert (sc.pedestrians[0].spawn.kind, sc.pedestrians[0].spawn.frequency) == ("periodic", 1.04)
    assert (sc.pedestrians[1].spawn.kind, sc.pedestrians[1].spawn.count) == ("once", 25)


@pytest.mark.parametrize("broken", [
    LANES_LIKE.replace("[field]\nsize = [60, 30]", "[field]"),            # missing required key
    LANES_LIKE.replace('kind = "once", count = 25', 'kind = "never"'),    # unknown enum tag
    LANES_LIKE.replace("line = [[5, 0.5], [5, 7.5]]", "line = [[5, 0.5]]"),
])
def test_malformed_scenarios_are_rejected(broken):
    with pytest.raises((KeyError, ValueError)):
        Scenario.from_toml_str(broken)


def test_shipped_scenarios_fixture():
    names = helpers.scenario_names()
    assert {"default", "narrow-gap", "bottleneck", "evacuation", "lanes", "random"} <= set(names)
    ev = helpers.load_scenario("evacuation")
    assert len(ev.obstacles) == 100 and len(ev.waypoints) == 33
    assert sum(p.spawn.count for p in ev.pedestrians if p.spawn.kind == "once") == 84
    rnd = helpers.load_scenario("random")
    assert len(rnd.obstacles) == 1004 and rnd.field.size == (200.0, 200.0)
    lanes = helpers.load_scenario("lanes")
    assert [p.spawn.frequency for p in lanes.pedestrians] == [1.04, 1.04]
    assert helpers.load_scenario("s-shape").field.size == helpers.load_scenario("default").field.size  # alias
