"""Checks shared by the user-defined-property tests (tests/scripts/props_script.py): decomposition-independent invariants that
hold only if every property stayed attached to ITS particle through the cell-order sort, the periodic wrap, capacity growth and
the migration between ranks."""
import numpy as np

COEF = (1.0, 0.3719, 0.1137)       # init_scale() of props_script.py


def expected_scale(x0, xlen):
    return 1.0 + 0.25 * (x0[:, 0] * COEF[0] + COEF[1] * x0[:, 1] + COEF[2] * x0[:, 2]) / xlen


def check_identity(position, path, scale, box, xlen):
    """`scale` was written once, from the INITIAL position; `path` integrates the displacement every step.  So position - path
    (wrapped back into the box) is the particle's lattice site, and the scale computed from that site must be the scale the
    particle carries -- any mix-up of rows between particles breaks this by at least the spacing of the scale values (~1e-5)."""
    x0 = position - path
    x0 = x0 - box * np.floor((x0 + 1e-6) / box)
    err = np.abs(expected_scale(x0, xlen) - scale).max()
    assert err <= 1e-9, err
    return err
