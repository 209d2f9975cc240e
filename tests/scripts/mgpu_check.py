"""Multi-GPU parity check, one process per GPU (launch with torchrun --nproc-per-node N):
every rank runs the MD loop on its sub-box through the C-ABI with NCCL halo exchange / migration; the comparator is the
oracle restatement running the SAME N-rank decomposition in one process (what the reference does under MPI).
Checks per step: per-rank nlocal / nghost identical, global temperature within 1e-9; at the end per-particle positions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from pairs_b200 import backend
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = pow(4.0 / 0.8442, 1.0 / 3.0)
    grid = [0.0, nx * a, 0.0, nx * a, 0.0, nx * a]
    ctx = backend.Context(local)
    ctx.init_domain(grid, world_size=world, rank=rank)
    ids = [backend.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.nccl_init(ids[0])
    def run_case(overlap):
        ctx.set_option("overlap_comm", overlap)
        n = ctx.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
        ctx.adjust_thermo(1.44)
        ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
        log = {}
        chunk = 7      # multi-step calls: fused integrators + (N > 1) ghost refresh overlapped with the interior force
        for c in range(0, steps, chunk):
            th = ctx.md_run(c, min(c + chunk, steps), 0.005, 2.5, 2.8, 2.8, 20, chunk)
            for row in th:
                log[int(row[0])] = float(row[1])
            log[("counts", min(c + chunk, steps) - 1)] = ctx.counts()
        return n, log, ctx.ints("tag"), ctx.real("position"), ctx.real("linear_velocity")

    n, log, tag, pos, vel = run_case(1)
    n2, log2, tag2, pos2, vel2 = run_case(0)
    # overlap only reorders independent work: results must be bit-identical
    assert log == log2 and np.array_equal(tag, tag2) and np.array_equal(pos, pos2) and np.array_equal(vel, vel2)
    state = {"log": log, "tag": tag, "pos": pos, "n0": n, "decomp": ctx.decomposition()}
    gathered = [None] * world
    dist.all_gather_object(gathered, state)
    ok = True
    if rank == 0:
        from oracle import port
        sim = port.md_example(nx, world_size=world, reneigh_every=20, particle_capacity=200000, send_capacity=200000)
        assert tuple(gathered[0]["decomp"]["nranks"]) == sim.nranks
        for k, r in enumerate(sim.ranks):
            d = r.decomposition()
            assert np.array_equal(d["neighbor_ranks"], gathered[k]["decomp"]["neighbor_ranks"])
            assert np.array_equal(d["pbc"], gathered[k]["decomp"]["pbc"]) and np.array_equal(d["subdom"], gathered[k]["decomp"]["subdom"])
            assert r.nlocal == gathered[k]["n0"]
        worst = 0.0
        checked = 0
        for ts in range(steps):
            sim.step(ts)
            t = sim.thermo()[0]
            for k, r in enumerate(sim.ranks):
                lg = gathered[k]["log"]
                if ("counts", ts) in lg:
                    assert lg[("counts", ts)] == (r.nlocal, r.nghost), (ts, k, lg[("counts", ts)], r.nlocal, r.nghost)
                if ts in lg:
                    worst = max(worst, abs(lg[ts] - t) / t)
                    checked += 1
        assert checked >= world * (steps // 7)
        assert worst <= 1e-9, worst
        # per-particle end state: match through exact lattice identity = sorted coordinates per rank
        for k, r in enumerate(sim.ranks):
            pg = np.sort(gathered[k]["pos"], axis=0)
            po = np.sort(r.real("position"), axis=0)
            assert np.abs(pg - po).max() <= 1e-9
        print(f"mgpu_check ok: world {world}, grid {sim.nranks}, {steps} steps, worst thermo rel err {worst:.2e}")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
