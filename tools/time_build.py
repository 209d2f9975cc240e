"""Times the reneighbouring stages on the bench workload (GPU box only)."""
import os
import sys
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
ctx.md_run(0, 60, 0.005, 2.5, 2.8, 2.8, 20, 100)
ctx.timers_reset(); ctx.timers_enable(True)
for _ in range(5):
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(2.8)
for name in ("exchange", "borders", "build_cell_lists", "build_neighbor_lists"):
    ms, c = ctx.timer(name)
    print(f"{name:24s} {ms / c:8.3f} ms")
