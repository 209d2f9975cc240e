"""DEM weak scaling (BASELINE.json configs[4]): examples/dem.py's system on a box of (0.8 m x 0.8 m x 0.2 m) per GPU, RegularXY
partitioner (1x2x1, 2x2x1, 2x4x1 ranks), particles migrate WITH their contact history every iteration (dem.py reneighbours
every step).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dem_multi.py [settle_steps]

Prints one JSON line on rank 0: particle-steps/s (all ranks) in a falling window and in a settled window, max over ranks of
the device time."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from tests import dem_common as dc  # noqa: E402

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}       # what Regular6DStencil::setConfig picks for these boxes


def main():
    import torch
    import torch.distributed as dist
    from pairs_b200 import backend
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    settle = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    gx, gy = GRIDS[world]
    domain = (0.8 * gx, 0.8 * gy, 0.2)
    if len(sys.argv) > 2 and sys.argv[2] == "c5":
        # BASELINE.json configs[4] / SURVEY.md C5: the 3.2 x 3.2 x 0.2 box = 15,974,400 spheres on 8 GPUs (2 x 4 x 1 ranks)
        assert world == 8, "config C5 is defined for 8 GPUs"
        domain = (3.2, 3.2, 0.2)
    ctx = backend.Context(local)
    ctx.init_domain([0.0, domain[0], 0.0, domain[1], 0.0, domain[2]], pbc=(1, 1, 0), partitioner=1, world_size=world, rank=rank)
    dec = ctx.decomposition()
    assert tuple(dec["nranks"]) == (gx, gy, 1), dec["nranks"]
    if world > 1:
        ids = [backend.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.nccl_init(ids[0])
    ctx.dem_enable(dc.C)
    ctx.dem_set_params(dc.DT, math.pi, dc.KAPPA, dc.LN_DRY, dc.COLLISION_TIME, dc.RHO_P, dc.RHO_F, dc.G, dc.NTYPES, dc.FS, dc.FD)
    ctx.setup_cells(dc.CELL)
    # runtime/dem_sc_grid.hpp: every rank walks the whole grid and keeps the points inside its sub-box
    g = ctx.dem_sc_grid(domain[0], domain[1], domain[2], dc.SPACING, dc.DIAMETER, dc.MIN_D, dc.MAX_D, dc.V0, dc.RHO_P, dc.NTYPES)
    ns = len(g["uid"])
    n = ns + 2
    pos, vel, normal = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    mass, radius = np.ones(n), np.zeros(n)
    uid, typ, flags, shape = (np.zeros(n, np.int32) for _ in range(4))
    pos[:ns], vel[:ns], mass[:ns], radius[:ns], uid[:ns], typ[:ns] = g["position"], g["linear_velocity"], g["mass"], g["radius"], g["uid"], g["type"]
    planes = [(100000000, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), (100000001, domain, (0.0, 0.0, -1.0))]
    for k, (u, p, nrm) in enumerate(planes):
        uid[ns + k], pos[ns + k], normal[ns + k], flags[ns + k], shape[ns + k] = u, p, nrm, 13, 1
    ctx.reserve(int(1.25 * n) + 65536)
    ctx.upload(pos, vel, mass, typ, flags, uid, shape)
    ctx.dem_upload("radius", radius)
    ctx.dem_upload("normal", normal)
    ctx.dem_stage("update_mass_and_inertia")

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    n_global = int(allsum(ns))
    res = {"n_gpus": world, "rank_grid": [gx, gy, 1], "spheres": n_global, "domain": domain}

    def window(name, a, b):
        ctx.timers_reset()
        ctx.timers_enable(True)
        ctx.sync()
        if world > 1:
            dist.barrier()
        ctx.stream_timer_start()
        ctx.dem_run(dc.CELL, a, b)
        ms = allmax(ctx.stream_timer_stop())
        ctx.timers_enable(False)
        nl, ng = ctx.counts()
        c = ctx.dem_download_contacts(nl)
        contacts = allsum(float(c["num_contacts"].sum()))
        stages = {k: allmax(ctx.timer(k)[0] / (b - a)) for k in ("exchange", "borders", "build_cell_lists", "gravity", "linear_spring_dashpot",
                                                                  "euler", "reset_contact_history_usage_status", "clear_unused_contact_history")}
        res[name] = {"steps": b - a, "ms_per_step": ms / (b - a), "particle_steps_per_s": n_global * (b - a) / (ms * 1e-3),
                     "mean_contacts": contacts / n_global, "locals_min_max": [int(-allmax(-nl)), int(allmax(nl))], "ghosts_max": int(allmax(ng)),
                     "stages_ms_per_step_max_over_ranks": stages}

    ctx.dem_run(dc.CELL, 0, 20)
    window("falling", 20, 220)
    ctx.dem_run(dc.CELL, 220, settle)
    window("settled", settle, settle + 200)
    total = int(allsum(ctx.counts()[0])) - 2 * world
    res["spheres_at_end"] = total
    assert total == n_global, (total, n_global)          # nothing lost or duplicated by thousands of migrations
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
