"""Small runs of the paths added late in round 2, meant to be executed under compute-sanitizer (memcheck / racecheck):
python tools/sanitize_probe.py [md|dem]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "scripts"))
from pairs_b200 import backend  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "md"
if what == "md":
    nx = 8
    L = nx * pow(4.0 / 0.8442, 1.0 / 3.0)
    ctx = backend.Context(0)
    ctx.init_domain([0, L, 0, L, 0, L])
    ctx.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
    ctx.adjust_thermo(1.44)
    ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
    args = (0.005, 2.5, 2.8, 2.8, 20, 10)
    ctx.md_run(0, 25, *args)                                   # tile build with the fp32 pre-filter + reorder, fused force kernel
    state = [np.ascontiguousarray(a) for a in (ctx.real("position"), ctx.real("linear_velocity"), ctx.real("mass"), ctx.ints("type"))]
    for a in state:
        ctx.host_register(a)
    th = ctx.md_run_from_host(*state, 0, 25, *args)           # overlapped upload, split reorder, ghost refill
    for a in state:
        ctx.host_unregister(a)
    ctx.set_option("tile_prefilter", 0)
    ctx.md_run(25, 45, *args)                                  # fp64 build kernel
    print("md probe ok", th[-1])
else:
    import dem_script
    ctx = dem_script.build("gpu", (0.1, 0.015, 0.04), 330, more_contact_props=True).generate()
    n = ctx.counts()[0]
    cx = ctx.dem_download_contact_extras(n)
    print("dem probe ok", n, float(cx[..., 3].max()))
