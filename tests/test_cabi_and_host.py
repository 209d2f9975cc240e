"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/pairs_b200.h declares, the pure host
entry points agree with the oracle, the product fails loudly without a device, and nothing under pairs_b200/ touches oracle/."""
import os
import re

import numpy as np
import pytest

from tests.conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pairs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pairs_b200 import backend
    lib = backend.load()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pairs_b200.h but not exported"
    assert sorted(backend.SIGNATURES) == syms, "ctypes signature table and header disagree"
    assert lib.pb_version().decode().startswith("pairs_b200")


def test_rank_grid_matches_reference_factorisation():
    """pb_rank_grid == Regular6DStencil::setConfig (oracle restatement, pinned): cubes, slabs, RegularXY."""
    from oracle import port
    from pairs_b200 import backend
    a = pow(4.0 / 0.8442, 1.0 / 3.0)
    boxes = [[0, 100 * a, 0, 100 * a, 0, 100 * a], [0, 100 * a, 0, 100 * a, 0, 200 * a], [0, 100 * a, 0, 200 * a, 0, 200 * a],
             [0, 0.8, 0, 0.015, 0, 0.2], [0, 0.8, 0, 0.8, 0, 0.2], [0, 3.2, 0, 3.2, 0, 0.2], [0, 1, 0, 2, 0, 3]]
    for box in boxes:
        for world in (1, 2, 3, 4, 6, 8, 12, 16):
            assert backend.rank_grid(world, box, 0) == port.set_config(world, box, (1, 1, 1)), (box, world)
            assert backend.rank_grid(world, box, 1) == port.set_config(world, box, (1, 1, 0)), (box, world)
    assert backend.rank_grid(8, boxes[0]) == (2, 2, 2) and backend.rank_grid(2, boxes[0]) == (1, 1, 2)
    assert backend.rank_grid(8, boxes[5], 1) == (2, 4, 1)          # DEM square box, RegularXY


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device failure mode")
def test_fails_loudly_without_a_device():
    from pairs_b200 import backend
    with pytest.raises(backend.BackendError, match="no CUDA device|no CPU fallback"):
        backend.Context(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: no file of the product package may import, load or execute anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pairs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/_ref|pairs_oracle|libref_|_build/libpairs_oracle", text):
                    if f == "dsl.py" and "parity oracle under oracle/" in text and not re.search(r"(from|import)\s+oracle", text):
                        continue
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    for f in os.listdir(os.path.join(ROOT, "pairs")):
        assert "oracle" not in open(os.path.join(ROOT, "pairs", f)).read()


def test_timestep_guards_match_reference():
    """sim/timestep.py:47: every-n procedures run when ((ts+1) % n == 0) || ts == 0; 201 loop iterations for timesteps=200."""
    hits = [ts for ts in range(201) if ((ts + 1) % 20 == 0) or ts == 0]
    assert hits[:4] == [0, 19, 39, 59] and len(hits) == 11      # 11 rebuilds in examples/md.py
    thermo = [ts for ts in range(201) if ((ts + 1) % 100 == 0) or ts == 0]
    assert thermo == [0, 99, 199]                               # the three thermo lines of the reference's stdout


def _board_worker(args):
    name, world, rank, rounds = args
    from pairs_b200 import backend
    return backend.load().pb_board_selftest(name.encode(), world, rank, rounds)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_count_board_protocol_between_processes(world):
    """The shared-memory board that carries message counts between the ranks of a node (csrc/comm_nccl.cu): `world` real
    processes on a periodic ring exchange 2000 messages per dimension; every received count is checked, a lost or reordered
    post shows up as a wrong value or a timeout."""
    import multiprocessing as mp
    import os
    from pairs_b200 import backend
    name = f"/pairs_b200_test_{os.getpid()}_{world}"
    ctx = mp.get_context("spawn")
    try:
        with ctx.Pool(world) as pool:
            res = pool.map(_board_worker, [(name, world, r, 2000) for r in range(world)])
    finally:
        backend.load().pb_board_unlink(name.encode())
    assert res == [0] * world
