"""Shared helpers of the DEM tests: parameters of the reference's examples/dem.py and golden access."""
import math
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dem_t1.npz")
# examples/dem.py:100-122
DOMAIN = (0.1, 0.015, 0.04)          # variant dem_t1 (the stock script uses 0.8 x 0.015 x 0.2)
DIAMETER, SPACING, V0, RHO_P, RHO_F, G = 0.0029, 0.005, 1, 2550, 1000, 9.81
DT, FRICTION, RESTITUTION, COLLISION_TIME, POISSON = 5e-5, 0.5, 0.1, 5e-4, 0.22
KAPPA = 2.0 * (1.0 - POISSON) / (2.0 - POISSON)
LN_DRY = math.log(RESTITUTION)
MIN_D, MAX_D = DIAMETER * 0.9, DIAMETER * 1.1
CELL = 1.01 * MAX_D
NTYPES, C = 1, 20
FS, FD = [0.0], [FRICTION]
# data/planes.input of the reference, with the upper plane at the corner of THIS box (uid,type,mass,pos,normal,flags)
PLANES = [(100000, 0, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 13), (100001, 0, 1.0, (0.8, 0.015, 0.2), (0.0, 0.0, -1.0), 13)]


def gold():
    return np.load(GOLD)


def state(z, tag, ts, names):
    return {k: np.ascontiguousarray(z[f"{tag}_{ts}_{k}"]) for k in names}


def contact_sets(num, uid, stick, tsd, ivm, n):
    """Per particle: {partner uid: (is_sticking, tsd bits, ivm bits)} -- slot order is traversal-dependent, the set is not."""
    out = []
    for i in range(n):
        out.append({int(uid[i, c]): (int(stick[i, c]), tuple(np.asarray(tsd[i, c]).view(np.int64)), int(np.float64(ivm[i, c]).view(np.int64)))
                    for c in range(int(num[i]))})
    return out
