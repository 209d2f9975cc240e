"""Tile lists vs per-particle lists, fused-multiply-add vs the reference's arithmetic, on the 4 M-atom bench workload: ms per step and
per-stage times.  Usage (GPU box): python tools/bench_tiles.py [nx]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pairs_b200 import backend  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
L = nx * pow(4.0 / 0.8442, 1.0 / 3.0)
out = {"atoms": 4 * nx ** 3}
for tile, fma in ((0, 0), (1, 0), (1, 1)):
    ctx = backend.Context(0)
    ctx.init_domain([0, L, 0, L, 0, L])
    ctx.set_option("tile_lists", tile)
    ctx.set_option("lj_fma", fma)
    ctx.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
    ctx.adjust_thermo(1.44)
    ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
    ctx.md_run(0, 40, 0.005, 2.5, 2.8, 2.8, 20, 0)
    ctx.timers_enable(True)
    ctx.timers_reset()
    ctx.stream_timer_start()
    ctx.md_run(40, 140, 0.005, 2.5, 2.8, 2.8, 20, 0)
    ms = ctx.stream_timer_stop()
    stages = {}
    for name in ("lennard_jones", "build_neighbor_lists", "build_cell_lists", "exchange", "borders", "synchronize"):
        t, c = ctx.timer(name)
        if c:
            stages[name] = {"ms_per_call": t / c, "calls": c}
    out[f"tile{tile}_fma{fma}"] = {"ms_per_step": ms / 100, "atom_steps_per_s": out["atoms"] * 100 / (ms * 1e-3), "stages": stages}
    ctx.close()
print(json.dumps(out))
