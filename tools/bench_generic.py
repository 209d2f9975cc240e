"""What the generic (generated + NVRTC) path costs next to the hand-written kernels, and what user-defined properties add to a
reneighbouring: the md.py procedure list on a 4 nx^3 FCC lattice run three ways through the DSL --
  native    recognised kernels, native loop (pb_md_run: fused integrators)
  generic   the same three kernels forced through kernelgen (staged Python loop, no fusion)
  props     tests/scripts/props_script.py: six extra properties (10 rows) carried through sort / wrap / ghosts
ms per iteration from the wall clock around generate() minus set-up, and the per-stage CUDA-event timers.
Usage (GPU box): python tools/bench_generic.py [nx=63] [steps=100]"""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "scripts"))
from pairs_b200 import dsl  # noqa: E402
import lj_script  # noqa: E402
import props_script  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 63
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = {"atoms": 4 * nx ** 3, "steps": steps}


def run(name, build, force_generic=False):
    res = {}
    for nsteps in (0, steps):                 # the 0-step run measures set-up + the first iteration (list build, JIT compilation)
        dsl.FORCE_GENERIC = force_generic
        try:
            psim = build("gpu", nx, nsteps, 20, 0)
        finally:
            dsl.FORCE_GENERIC = False
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ctx = psim.generate()
        res[nsteps] = (time.perf_counter() - t0, ctx, psim)
    wall = (res[steps][0] - res[0][0]) / steps
    ctx, psim = res[steps][1], res[steps][2]
    stages = {}
    for stage in ["exchange", "borders", "build_cell_lists", "build_neighbor_lists", "synchronize", "lennard_jones", "initial_integrate",
                  "final_integrate"] + [f"user_{e['name']}" for e in psim.pre_step + psim.functions]:
        ms, calls = ctx.timer(stage)
        if calls:
            stages[stage] = {"ms_per_call": ms / calls, "calls": calls}
    out[name] = {"ms_per_step_wall": wall * 1e3, "atom_steps_per_s": out["atoms"] / wall, "stages": stages}


run("native", lj_script.build)
run("generic", lj_script.build, force_generic=True)
run("props", props_script.build)
print(json.dumps(out))
