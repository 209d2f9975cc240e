"""Where does the end-to-end leg of bench.py spend its time?  python tools/e2e_probe.py [nx] [K]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairs_b200 import backend  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
L = nx * pow(4.0 / 0.8442, 1.0 / 3.0)
ctx = backend.Context(0)
ctx.init_domain([0, L, 0, L, 0, L])
ctx.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
ctx.adjust_thermo(1.44)
ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
run = lambda a, b: ctx.md_run(a, b, 0.005, 2.5, 2.8, 2.8, 20, 100)  # noqa: E731
run(0, 20)
ctx.sync()
t0 = time.perf_counter(); run(20, 20 + K); ctx.sync(); t_dev = time.perf_counter() - t0
pos, vel, mass, typ = ctx.real("position"), ctx.real("linear_velocity"), ctx.real("mass"), ctx.ints("type")
out_pos, out_vel = np.empty((len(pos) + 4096, 3)), np.empty((len(pos) + 4096, 3))
for a in (pos, vel, mass, typ, out_pos, out_vel):
    ctx.host_register(a)
ctx.sync()
res = {"wall_continue_s": t_dev, "ncap": ctx.lib.pb_neighbor_capacity(ctx.h)}
ctx.timers_reset(); ctx.timers_enable(True)
t0 = time.perf_counter(); ctx.upload(pos, vel, mass, typ); ctx.sync(); res["upload_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); run(0, K); ctx.sync(); res["run_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); ctx.real_into("position", out_pos); ctx.real_into("linear_velocity", out_vel); ctx.sync(); res["download_s"] = time.perf_counter() - t0
res["stages"] = {k: ctx.timer(k) for k in ("lennard_jones", "build_neighbor_lists", "build_cell_lists", "exchange", "borders", "synchronize", "compute_thermo")}
res["ncap_after"] = ctx.lib.pb_neighbor_capacity(ctx.h)
print(json.dumps(res))
