"""world_size-2 host-side test on CPU (gloo): the launcher plumbing bench.py / the DSL use for N > 1 -- rendezvous on
127.0.0.1, broadcast of the 128-byte communicator id from rank 0, max-over-ranks of the timed region -- and the per-rank
domain decomposition (neighbour ranks, PBC multipliers, sub-box bounds) that each rank derives on its own."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # each rank computes the decomposition of the weak-scaling box on its own (pure host code of the shim)
    from pairs_b200 import backend
    a = pow(4.0 / 0.8442, 1.0 / 3.0)
    grid = [0.0, 20 * a, 0.0, 20 * a, 0.0, 40 * a]
    nranks = backend.rank_grid(world, grid)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "nranks": nranks, "id_ok": ids[0] == bytes(range(128)), "tmax": float(t[0])})
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_rendezvous_id_broadcast_and_decomposition():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(r["id_ok"] for r in res) and all(r["tmax"] == 11.0 for r in res)
    assert all(tuple(r["nranks"]) == (1, 1, 2) for r in res)
    # the oracle's view of the same decomposition: prev/next consistency and PBC multipliers at the global faces
    from oracle import port
    a = pow(4.0 / 0.8442, 1.0 / 3.0)
    sim = port.OracleSim([0.0, 20 * a, 0.0, 20 * a, 0.0, 40 * a], world_size=2, particle_capacity=1000, send_capacity=1000)
    d0, d1 = sim.ranks[0].decomposition(), sim.ranks[1].decomposition()
    assert list(d0["neighbor_ranks"]) == [0, 0, 0, 0, 1, 1] and list(d1["neighbor_ranks"]) == [1, 1, 1, 1, 0, 0]
    assert list(d0["pbc"]) == [1, -1, 1, -1, 1, 0] and list(d1["pbc"]) == [1, -1, 1, -1, 0, -1]
    assert d0["subdom"][5] == d1["subdom"][4] == 20 * a


def _id_worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    from pairs_b200 import dsl

    class FakeBackend:                       # pb_nccl_unique_id needs no GPU, but a fixed pattern shows WHICH call's id arrived
        calls = 0

        @classmethod
        def nccl_unique_id(cls):
            cls.calls += 1
            return bytes([cls.calls]) * 128

    got = []
    for k in range(3):
        if rank == 1:
            time.sleep(0.3)                  # the reader arrives late: rank 0 has long returned from the previous call
        got.append(dsl._broadcast_nccl_id(FakeBackend, rank, world))
    out.put((rank, got))
    time.sleep(1.0 if rank == 0 else 0.0)    # (rank 0 hosts the store; in a real run the NCCL initialisation keeps it alive)


@pytest.mark.timeout(120)
def test_world2_dsl_id_broadcast_survives_late_readers_and_repeated_generate():
    """dsl._broadcast_nccl_id (one call per generate()): three calls in two processes, the reader 0.3 s late each time -- every
    call delivers ITS id to both ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port_no, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res[0] == res[1] == [bytes([k]) * 128 for k in (1, 2, 3)]
