"""Which force is right?  State after 41 iterations; forces (1) as left by the loop, (2) after upload + md_run(0, 1), (3) after upload +
module-by-module calls on per-particle lists, against a brute-force numpy evaluation with the minimum-image convention."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairs_b200.backend import Context  # noqa: E402

nx = 6
L = nx * pow(4.0 / 0.8442, 1.0 / 3.0)


def new():
    c = Context(0)
    c.init_domain([0, L, 0, L, 0, L])
    return c


def brute(x):
    f = np.zeros_like(x)
    for i in range(len(x)):
        d = x[i] - x
        d -= L * np.round(d / L)
        r2 = (d * d).sum(axis=1)
        m = (r2 < 6.25) & (r2 > 0)
        sr2 = 1.0 / r2[m]
        sr6 = sr2 ** 3
        f[i] = (d[m] * (48.0 * sr6 * (sr6 - 0.5) * sr2)[:, None]).sum(axis=0)
    return f


for nsteps in (1, 21, 41):
    a = new()
    a.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
    a.adjust_thermo(1.44)
    a.set_lj_params(4, [1.0] * 16, [1.0] * 16)
    run = lambda c, b, e: c.md_run(b, e, 0.005, 2.5, 2.8, 2.8, 20, 0)  # noqa: E731
    run(a, 0, nsteps)
    x, v, m, t, f_loop = a.real("position"), a.real("linear_velocity"), a.real("mass"), a.ints("type"), a.real("force")
    fb = brute(x)
    a2 = new()
    a2.set_lj_params(4, [1.0] * 16, [1.0] * 16)
    a2.setup_cells(2.8)
    a2.upload(x, v, m, t)
    run(a2, 0, 1)
    f2 = a2.real("force")[np.argsort(a2.ints("tag"))]
    a3 = new()
    a3.set_option("tile_lists", 0)
    a3.set_lj_params(4, [1.0] * 16, [1.0] * 16)
    a3.setup_cells(2.8)
    a3.upload(x, v, m, t)
    a3.exchange(); a3.borders(); a3.build_cell_lists(); a3.build_neighbor_lists(2.8); a3.reset_volatile(); a3.lennard_jones(2.5)
    f3 = a3.real("force")[np.argsort(a3.ints("tag"))]
    sc = np.abs(fb).max()
    print(nsteps, "outside box:", int(((x < 0) | (x >= L)).any(axis=1).sum()), "loop", np.abs(f_loop - fb).max() / sc, "upload+md_run", np.abs(f2 - fb).max() / sc,
          "upload+modules(per-particle lists)", np.abs(f3 - fb).max() / sc)
