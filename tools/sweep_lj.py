"""Sweep of the force-kernel tuning knobs on the bench workload (GPU box only): prints ms per launch."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairs_b200.backend import Context  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
RHO = 0.8442
L = nx * pow((4.0 / RHO), (1.0 / 3.0))
ctx = Context(0)
ctx.init_domain([0.0, L, 0.0, L, 0.0, L])
ctx.copper_fcc_lattice(nx, nx, nx, RHO, 4)
ctx.adjust_thermo(1.44)
ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
ctx.md_run(0, 60, 0.005, 2.5, 2.8, 2.8, 20, 100)       # melt a little so lists look like steady state
res = []
for lanes in (1, 2, 4):
    ctx.set_option("lanes_per_particle", lanes)
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(2.8)
    for unroll in (2, 4, 8):
        ctx.set_option("lj_unroll", unroll)
        for _ in range(3):
            ctx.reset_volatile(); ctx.lennard_jones(2.5)
        ctx.sync()
        ctx.timers_reset(); ctx.timers_enable(True)
        for _ in range(20):
            ctx.reset_volatile(); ctx.lennard_jones(2.5)
        ms, calls = ctx.timer("lennard_jones")
        ctx.timers_enable(False)
        res.append({"lanes": lanes, "unroll": unroll, "ms": ms / calls})
        print(json.dumps(res[-1]), flush=True)
ctx.timers_reset(); ctx.timers_enable(True)
ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(2.8)
for name in ("exchange", "borders", "build_cell_lists", "build_neighbor_lists"):
    print(name, ctx.timer(name))
