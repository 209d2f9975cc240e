"""Small driver for ncu captures of the MD hot path: python tools/prof_md.py [nx=63] [steps=45] [tile=1] [fma=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairs_b200 import backend  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 63
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 45
L = nx * pow(4.0 / 0.8442, 1.0 / 3.0)
ctx = backend.Context(0)
ctx.init_domain([0, L, 0, L, 0, L])
ctx.set_option("tile_lists", int(sys.argv[3]) if len(sys.argv) > 3 else 1)
ctx.set_option("lj_fma", int(sys.argv[4]) if len(sys.argv) > 4 else 1)
ctx.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
ctx.adjust_thermo(1.44)
ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
ctx.md_run(0, steps, 0.005, 2.5, 2.8, 2.8, 20, 0)
ctx.sync()
print("done", ctx.counts(), ctx.kernel_launches())
