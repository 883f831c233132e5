"""CPU: the oracle against the committed golden step vectors (tests/golden/step_vectors.npz, made by
tests/golden/make_step_vectors.py). Integer outputs (cell table, destinations, population) are exact;
floats may move in the last bits with the host's libm expf variant, hence the 1e-5 bound."""
import numpy as np
import pytest

import helpers
import oracle
from pedoni_b200 import SimulatorOptions

GOLD = np.load(helpers.GOLDEN / "step_vectors.npz")


@pytest.mark.parametrize("case,use_map", [("distance_map", True), ("segments", False)])
def test_oracle_reproduces_golden_vectors(case, use_map):
    oracle.lib().oracle_set_threads(2)
    sc = helpers.corridor_scenario()
    field = helpers.oracle_field(sc)
    m = helpers.OracleAdapter(SimulatorOptions(use_distance_map=use_map), sc, field)
    m.spawn_arrays(GOLD[f"{case}/in_pos"], GOLD[f"{case}/in_dest"], GOLD[f"{case}/in_v0"])
    m.rebuild()
    for tick in range(11):
        if tick in (0, 1, 5, 10):
            p, d, v, s = m.download()
            np.testing.assert_array_equal(m.cell_table(), GOLD[f"{case}/t{tick}_table"])
            np.testing.assert_array_equal(d, GOLD[f"{case}/t{tick}_dest"])
            np.testing.assert_allclose(p, GOLD[f"{case}/t{tick}_pos"], atol=1e-5, rtol=0)
            np.testing.assert_allclose(v, GOLD[f"{case}/t{tick}_vel"], atol=1e-5, rtol=0)
        m.step()
        m.rebuild()
