"""CPU unit test of the PRODUCT's MD arithmetic (pairs_b200/csrc/md_math.h -- the functions the CUDA kernels call -- compiled for
the host) against the modules of the reference's own generated C++ (oracle/_ref, called directly) and against the restatement:
cell indices and Lennard-Jones forces are identical bits (the force loop runs over the reference's lists in the reference's order,
so even the summation order is the same)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = ctypes.POINTER(ctypes.c_double)
I = ctypes.POINTER(ctypes.c_int)


def dp(a):
    return a.ctypes.data_as(D)


def ip(a):
    return a.ctypes.data_as(I)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mdhost") / "libmd_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "pairs_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "md_host.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def melted(nx, seed, ntypes=4, eps=None, sig6=None):
    from oracle import port
    sim = port.md_example(nx, reneigh_every=20, ntypes=ntypes, particle_capacity=60000, send_capacity=60000)
    if eps is not None:
        sim.set_params(2.8, 2.8, 2.5, 0.005, ntypes, eps, sig6, 20)
    r = sim.ranks[0]
    rng = np.random.default_rng(seed)
    r.real("position", r.nlocal, view=True)[:] += 0.2 * (rng.random((r.nlocal, 3)) - 0.5)
    sim.step(0)
    return sim, r


def test_cell_index_is_the_reference_value_for_every_particle(host):
    sim, r = melted(7, 1)
    tot = r.nlocal + r.nghost
    d = r.decomposition()
    lo = np.array([d["subdom"][0] - 2.8, d["subdom"][2] - 2.8, d["subdom"][4] - 2.8])
    pos = r.real("position", tot)
    flags = r.ints("flags", tot)
    # extra probes: coordinates exactly on cell edges, on and beyond the grid, an INFINITE particle
    extra = np.array([[lo[0] + 3 * 2.8, lo[1], lo[2] + 2.8], [lo[0] - 5.0, lo[1] + 1e3, lo[2]], [np.nextafter(lo[0] + 2 * 2.8, -np.inf), 0.0, 0.0],
                      [1.0, 1.0, 1.0]])
    eflags = np.array([0, 0, 0, 1], np.int32)
    out = np.zeros(tot + len(extra), np.int32)
    allpos = np.ascontiguousarray(np.vstack([pos, extra]))
    allflags = np.ascontiguousarray(np.concatenate([flags, eflags]))
    dim = np.ascontiguousarray(d["dim_cells"], np.int32)
    host.host_md_cell_index(dp(lo), ctypes.c_double(2.8), ip(dim), len(out), dp(allpos), ip(allflags), ip(out))
    assert np.array_equal(out[:tot], r.ints("particle_cell", tot))            # restatement (pinned to the reference bit for bit)
    assert out[-1] == 0                                                       # INFINITE -> cell 0
    from oracle import ref
    if ref.available("md"):
        prog = ref.RefProgram("md")
        res = prog.build_lists(allpos, allflags, np.full(len(out), 2, np.int32), len(out), 0, d["subdom"])
        assert np.array_equal(out, res["particle_cell"])                      # the reference's generated build_cell_lists itself


@pytest.mark.parametrize("uniform", [True, False])
def test_lj_pair_arithmetic_equals_the_reference_module(host, uniform):
    rng = np.random.default_rng(4)
    nt = 4
    eps = np.ones(16) if uniform else 0.8 + 0.4 * rng.random(16)
    sig6 = np.ones(16) if uniform else 0.9 + 0.2 * rng.random(16)
    sim, r = melted(6, 2, nt, list(eps), list(sig6))
    n, tot = r.nlocal, r.nlocal + r.nghost
    nn, nl = r.neighbor_sets()
    force = np.zeros((n, 3))
    host.host_md_lennard_jones(n, r.neighbor_capacity, ip(nn.astype(np.int32)), ip(np.ascontiguousarray(nl, np.int32)), ip(r.ints("flags", tot)),
                               dp(r.real("position", tot)), ip(r.ints("type", tot)), nt, dp(sig6), dp(eps), ctypes.c_double(6.25), dp(force))
    assert np.abs(force).max() > 1.0
    assert np.array_equal(force, r.real("force"))                             # restatement
    from oracle import ref
    if ref.available("md"):
        prog = ref.RefProgram("md")
        f_ref = np.zeros((n, 3))
        prog.lennard_jones(r.neighbor_capacity, n, nn.astype(np.int32), np.ascontiguousarray(nl, np.int32), r.ints("flags", tot),
                           r.real("position", tot), r.ints("type", tot), f_ref, sig6, eps)
        assert np.array_equal(force, f_ref)                                   # the reference's generated lennard_jones itself
