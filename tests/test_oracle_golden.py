"""The reference's own known-answer tests for this path, ported verbatim as pins of the oracle:
util.rs:148-154 (test_distance_from_line) and util.rs:156-163 (test_bilinear). These are ALL the
assertions the reference holds for the hot path."""
import numpy as np

import oracle

ABS_TOL = 1e-6  # assert_float_absolute_eq! default


def test_distance_from_line_util_rs_148():
    a, b = (1.0, 1.0), (4.0, 1.0)
    assert abs(np.hypot(*oracle.distance_from_line((2.0, 3.0), a, b)) - 2.0) <= ABS_TOL
    assert abs(np.hypot(*oracle.distance_from_line((0.0, 0.25), a, b)) - 1.25) <= ABS_TOL


def test_bilinear_util_rs_156():
    grid = np.array([[1.0, 0.0, 4.0], [3.0, 1.0, -1.0]], np.float32)
    assert abs(oracle.bilinear(grid, 0.0, 0.0) - 1.0) <= ABS_TOL
    assert abs(oracle.bilinear(grid, 0.5, 0.0) - 0.5) <= ABS_TOL
    assert abs(oracle.bilinear(grid, 0.0, 0.25) - 1.5) <= ABS_TOL
    assert abs(oracle.bilinear(grid, 0.5, 0.5) - 1.25) <= ABS_TOL


def test_bilinear_out_of_bounds_tap_is_1e12():
    # util.rs:45,53-56: a missing tap contributes 1e12 (weighted).
    grid = np.ones((2, 2), np.float32)
    assert oracle.bilinear(grid, -1.0, 0.0) == np.float32(1e12)  # all weight on an OOB tap
    assert oracle.bilinear(grid, 1.0, 0.0) == 1.0  # weight-0 taps are still added: 0 * 1e12 = 0
    v = oracle.bilinear(grid, 1.5, 0.0)
    assert abs(v - (0.5 * 1.0 + 0.5 * 1e12)) / 1e12 < 1e-6


def test_sobel_of_a_ramp():
    # u(x, y) = x  ->  gx = (u00+2u10+u20) - (u02+2u12+u22) = -8, gy = 0 (util.rs:72-73)
    grid = np.tile(np.arange(8, dtype=np.float32), (8, 1))
    g = oracle.sobel_filter(grid, 3.25, 3.5)
    assert abs(g[0] + 8.0) < 1e-5 and abs(g[1]) < 1e-5
