"""Secondary benchmark: DEM (examples/dem.py scaled to the 0.8 x 0.8 x 0.2 box = 998,400 spheres + 2 half-spaces, BASELINE.json
configs[2]) on one B200, particle-steps/s in a falling window (no contacts yet) and in a settled window (dense contacts), with the
reference's generated serial C++ timed on a host core next to it.  GPU box only:  python tools/bench_dem.py [settle_steps]"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from pairs_b200.backend import Context  # noqa: E402
from tests import dem_common as dc  # noqa: E402

settle = int(sys.argv[1]) if len(sys.argv) > 1 else 8000          # settled window = iterations 8000..9000
DOMAIN = (0.8, 0.8, 0.2)
ctx = Context(0)
ctx.init_domain([0.0, DOMAIN[0], 0.0, DOMAIN[1], 0.0, DOMAIN[2]], pbc=(1, 1, 0), partitioner=1)
ctx.dem_enable(dc.C)
if os.environ.get("PB_DEM_MAXREG"):            # occupancy experiment: contact kernel re-built at run time with a register cap
    ctx.set_option("dem_force_maxreg", int(os.environ["PB_DEM_MAXREG"]))
ctx.dem_set_params(dc.DT, math.pi, dc.KAPPA, dc.LN_DRY, dc.COLLISION_TIME, dc.RHO_P, dc.RHO_F, dc.G, dc.NTYPES, dc.FS, dc.FD)
ctx.setup_cells(dc.CELL)
g = ctx.dem_sc_grid(DOMAIN[0], DOMAIN[1], DOMAIN[2], dc.SPACING, dc.DIAMETER, dc.MIN_D, dc.MAX_D, dc.V0, dc.RHO_P, dc.NTYPES)
ns = len(g["uid"])
n = ns + 2
pos, vel, normal = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
mass, radius = np.ones(n), np.zeros(n)
uid, typ, flags, shape = (np.zeros(n, np.int32) for _ in range(4))
pos[:ns], vel[:ns], mass[:ns], radius[:ns], uid[:ns], typ[:ns] = g["position"], g["linear_velocity"], g["mass"], g["radius"], g["uid"], g["type"]
planes = [(100000, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), (100001, DOMAIN, (0.0, 0.0, -1.0))]
for k, (u, p, nrm) in enumerate(planes):
    uid[ns + k], pos[ns + k], normal[ns + k], flags[ns + k], shape[ns + k] = u, p, nrm, 13, 1
ctx.upload(pos, vel, mass, typ, flags, uid, shape)
ctx.dem_upload("radius", radius)
ctx.dem_upload("normal", normal)
ctx.dem_stage("update_mass_and_inertia")
res = {"particles": n}


def window(name, a, b):
    ctx.timers_reset(); ctx.timers_enable(True)
    ctx.sync()
    ctx.stream_timer_start()
    ctx.dem_run(dc.CELL, a, b)
    ms = ctx.stream_timer_stop()
    ctx.timers_enable(False)
    c = ctx.dem_download_contacts(n)
    res[name] = {"steps": b - a, "ms_per_step": ms / (b - a), "particle_steps_per_s": n * (b - a) / (ms * 1e-3),
                 "mean_contacts": float(c["num_contacts"].mean()), "nghost": ctx.counts()[1],
                 "stages_ms_per_step": {k: ctx.timer(k)[0] / (b - a) for k in ("exchange", "borders", "build_cell_lists", "gravity",
                                                                                 "linear_spring_dashpot", "euler", "reset_contact_history_usage_status",
                                                                                 "clear_unused_contact_history")}}


ctx.dem_run(dc.CELL, 0, 20)
window("falling", 20, 220)
t0 = time.time()
ctx.dem_run(dc.CELL, 220, settle)
res["settle_wall_s"] = time.time() - t0
window("settled", settle, settle + 1000)
try:
    from oracle import ref, ref_worker
    if ref.available("dem_bench"):
        r = ref_worker.bench_many("dem_bench", 2, 8, 1)[0]
        res["cpu_reference"] = {"particle_steps_per_s": r["n"] * r["steps"] / r["seconds"], "cores": 1,
                                "sample": "same 998,402-particle box, loop iterations 3..10 (falling phase), serial target, g++ -O3 -ffp-contract=off"}
except Exception as e:      # noqa: BLE001
    res["cpu_reference"] = {"error": str(e)}
print(json.dumps(res))
