"""Driver for ncu captures of the DEM contact kernels in the settled state (998,402 particles):
   ncu --set full --clock-control none --import-source on -k regex:pb_k_dem_(detect|force) -s 8000 -c 2 python tools/profile_dem.py 4000"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from pairs_b200.backend import Context  # noqa: E402
from tests import dem_common as dc  # noqa: E402

settle = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
DOMAIN = (0.8, 0.8, 0.2)
ctx = Context(0)
ctx.init_domain([0.0, DOMAIN[0], 0.0, DOMAIN[1], 0.0, DOMAIN[2]], pbc=(1, 1, 0), partitioner=1)
ctx.dem_enable(dc.C)
ctx.dem_set_params(dc.DT, math.pi, dc.KAPPA, dc.LN_DRY, dc.COLLISION_TIME, dc.RHO_P, dc.RHO_F, dc.G, dc.NTYPES, dc.FS, dc.FD)
ctx.setup_cells(dc.CELL)
g = ctx.dem_sc_grid(DOMAIN[0], DOMAIN[1], DOMAIN[2], dc.SPACING, dc.DIAMETER, dc.MIN_D, dc.MAX_D, dc.V0, dc.RHO_P, dc.NTYPES)
ns = len(g["uid"])
n = ns + 2
pos, vel, normal = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
mass, radius = np.ones(n), np.zeros(n)
uid, typ, flags, shape = (np.zeros(n, np.int32) for _ in range(4))
pos[:ns], vel[:ns], mass[:ns], radius[:ns], uid[:ns], typ[:ns] = g["position"], g["linear_velocity"], g["mass"], g["radius"], g["uid"], g["type"]
for k, (u, p, nrm) in enumerate([(100000000, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), (100000001, DOMAIN, (0.0, 0.0, -1.0))]):
    uid[ns + k], pos[ns + k], normal[ns + k], flags[ns + k], shape[ns + k] = u, p, nrm, 13, 1
ctx.upload(pos, vel, mass, typ, flags, uid, shape)
ctx.dem_upload("radius", radius)
ctx.dem_upload("normal", normal)
ctx.dem_stage("update_mass_and_inertia")
ctx.dem_run(dc.CELL, 0, settle + 3)
ctx.sync()
print("mean contacts", float(ctx.dem_download_contacts(n)["num_contacts"].mean()))
