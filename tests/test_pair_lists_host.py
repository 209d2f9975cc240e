"""Pair lists (pairs_b200/csrc/pair_lists.h, option "pair_lists"): one neighbour list per pair of consecutive particles.  The two
kernel bodies are host+device functions; here they are compiled for the host and run on the oracle's state and lists."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("pl") / "pair_lists_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "pairs_b200", "csrc"),
                    os.path.join(HERE, "host", "pair_lists_host.cpp"), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    P, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.host_pairlist_merge_all.argtypes = [I, I, I, D, P, P, P, P, P, P]
    lib.host_lj_pairs_all.argtypes = [I, I, I, I, D, D, P, P, P, P, P, P, P, P, P, P, I, I]
    return lib


def _p(a):
    return a.ctypes.data


def _system(ntypes_uniform, odd=False):
    from oracle import port
    nx = 6
    eps = [1.0] * 16 if ntypes_uniform else [1.0 + 0.05 * ((k % 4) + (k // 4)) for k in range(16)]
    lattice = pow(4.0 / 0.8442, 1.0 / 3.0)
    L = nx * lattice
    sim = port.OracleSim([0.0, L, 0.0, L, 0.0, L], particle_capacity=60000, send_capacity=60000)
    sim.set_params(2.8, 2.8, 2.5, 0.005, 4, eps, [1.0] * 16, 20)
    sim.copper_fcc_lattice(nx, nx, nx, 0.8442, 1.44, 4)
    r = sim.ranks[0]
    n = r.nlocal
    rng = np.random.default_rng(12)
    r.real("position", n, view=True)[:] += 0.2 * (rng.random((n, 3)) - 0.5)
    r.ints("flags", n, view=True)[::23] |= 4
    sim.step(0)
    return sim, r, np.array(eps)


def _device_lists(r):
    n, tot = r.nlocal, r.nlocal + r.nghost
    nn, nl = r.neighbor_sets()
    T = int(nn.max()) + 3
    neigh = np.zeros(((n + 31) // 32, T, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = r.real("position", tot)
    pos4[:, 3] = r.ints("type", tot).astype(np.int64).view(np.float64)
    return nn, nl, T, neigh, numneigh, pos4


@pytest.mark.parametrize("uniform", [True, False])
def test_pair_lists_give_the_oracles_forces_and_integrator_updates(host, uniform):
    sim, r, eps = _system(uniform)
    n, tot = r.nlocal, r.nlocal + r.nghost
    nn, nl, T, neigh, numneigh, pos4 = _device_lists(r)
    npairs = (n + 1) // 2
    T2 = (2 * int(nn.max()) + 3) // 4 * 4
    pneigh = np.full(((npairs + 31) // 32, T2, 32), -1, np.int32)
    pnum = np.zeros(npairs, np.int32)
    flags = r.ints("flags", tot).copy()
    host.host_pairlist_merge_all(n, T, T2, 2.8 * 2.8, _p(pos4), _p(flags), _p(numneigh), _p(neigh), _p(pneigh), _p(pnum))
    # the union list of a pair is the union of the two lists without the owners, every partner once
    for p in range(npairs):
        i0, i1 = 2 * p, 2 * p + 1
        want = set(nl[i0, :nn[i0]].tolist()) | (set(nl[i1, :nn[i1]].tolist()) if i1 < n else set())
        want -= {i0, i1}
        got = pneigh[p // 32, :pnum[p], p % 32].tolist()
        assert len(got) == len(set(got)) and set(got) == want, p
    assert pnum.mean() < 1.8 * nn.mean()                     # partners are shared even in the oracle's (unsorted) particle order
    mass = r.real("mass", tot).copy()
    sig6 = np.ones(16)
    f_oracle = r.real("force").copy()
    scale = np.abs(f_oracle).max()
    fixed = (flags[:n] & 4) != 0
    assert scale > 1.0 and fixed.sum() > 10 and not f_oracle[fixed].any()
    # (a) plain evaluation with the reset folded in, (b) accumulation onto an existing force
    for mode_acc, start in ((0, 7.0), (2, 0.25)):
        force = np.full((3, tot), start)
        vel = np.ascontiguousarray(r.real("linear_velocity", tot).T)
        nxt = pos4.copy()
        host.host_lj_pairs_all(n, T2, tot, 4, 2.5 * 2.5, 0.005, _p(eps), _p(sig6), _p(pos4), _p(flags), _p(pnum), _p(pneigh), _p(force), _p(mass),
                               _p(vel), _p(nxt), (1 if uniform else 0) | mode_acc, 0)
        base = np.where(fixed[:, None], start if mode_acc else 0.0, start if mode_acc else 0.0)
        assert np.abs(force[:, :n].T - (f_oracle + base)).max() <= 2e-13 * scale
        assert np.all(force[:, n:] == start)                 # ghosts are not touched
    # (c) both integrator halves fused into the epilogue: final_integrate of this step + initial_integrate of the next one
    v0 = r.real("linear_velocity", tot).copy()
    force = np.zeros((3, tot))
    vel = np.ascontiguousarray(v0.T)
    nxt = np.zeros_like(pos4)
    host.host_lj_pairs_all(n, T2, tot, 4, 2.5 * 2.5, 0.005, _p(eps), _p(sig6), _p(pos4), _p(flags), _p(pnum), _p(pneigh), _p(force), _p(mass),
                           _p(vel), _p(nxt), (1 if uniform else 0), 3)
    sim.final_integrate()
    sim.initial_integrate()
    assert np.abs(vel[:, :n].T - r.real("linear_velocity")).max() <= 1e-14
    assert np.abs(nxt[:n, :3] - r.real("position")).max() <= 1e-14
    assert np.array_equal(nxt[:n, 3].view(np.int64), pos4[:n, 3].view(np.int64))          # the type rides along
    assert np.array_equal(vel[:, :n].T[fixed], v0[:n][fixed]) and np.array_equal(nxt[:n, :3][fixed], pos4[:n, :3][fixed])
