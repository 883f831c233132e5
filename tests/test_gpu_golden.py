"""GPU: the CUDA path through the C ABI against the committed golden step vectors (no oracle call)."""
import numpy as np
import pytest

import helpers
from pedoni_b200 import PEDONI_MATH_FAST, PEDONI_MATH_STRICT, SimulatorOptions, SocialForceModelCuda

pytestmark = pytest.mark.gpu
GOLD = np.load(helpers.GOLDEN / "step_vectors.npz")


@pytest.mark.parametrize("mode", [PEDONI_MATH_STRICT, PEDONI_MATH_FAST])
@pytest.mark.parametrize("case,use_map", [("distance_map", True), ("segments", False)])
def test_cuda_reproduces_golden_vectors(case, use_map, mode):
    sc = helpers.corridor_scenario()
    field = helpers.oracle_field(sc)  # the field is an input of the path; both sides share the arrays
    cu = SocialForceModelCuda(SimulatorOptions(use_distance_map=use_map), sc, field, math_mode=mode)
    cu.spawn_arrays(GOLD[f"{case}/in_pos"], GOLD[f"{case}/in_dest"], GOLD[f"{case}/in_v0"])
    cu.rebuild()
    tol_p, tol_v = helpers.tolerances(mode)
    for tick in range(11):
        if tick in (0, 1, 5, 10):
            p, d, v, s = cu.download()
            np.testing.assert_array_equal(cu.cell_table(), GOLD[f"{case}/t{tick}_table"])  # bit-exact cells
            np.testing.assert_array_equal(d, GOLD[f"{case}/t{tick}_dest"])
            np.testing.assert_array_equal(helpers.bits(s), helpers.bits(GOLD[f"{case}/t{tick}_v0"]))
            assert np.abs(p - GOLD[f"{case}/t{tick}_pos"]).max() <= tol_p
            assert np.abs(v - GOLD[f"{case}/t{tick}_vel"]).max() <= tol_v
        cu.step()
        cu.rebuild()
    cu.close()
