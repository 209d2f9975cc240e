"""ctypes access to oracle/pairs_oracle.c (the CPU restatement of the reference's MD path).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from pairs_b200/.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pairs_oracle.c")
LIB = os.path.join(HERE, "_build", "libpairs_oracle.so")


def build(force=False):
    """gcc -O3 -ffp-contract=off (no FMA contraction: results comparable op by op)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-ffp-contract=off", "-std=c99", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def _load():
    lib = ctypes.CDLL(build())
    P = ctypes.c_void_p
    I = ctypes.c_int
    D = ctypes.c_double
    IP = ctypes.POINTER(ctypes.c_int)
    DP = ctypes.POINTER(ctypes.c_double)
    lib.po_create.argtypes = [I, DP, IP, IP, I, I, I, I]
    lib.po_create.restype = P
    lib.po_destroy.argtypes = [P]
    lib.po_set_params.argtypes = [P, D, D, D, D, I, DP, DP, I, I]
    lib.po_set_compute_half.argtypes = [P, I]
    lib.po_copper_fcc_lattice.argtypes = [P, I, I, I, D, I]
    lib.po_adjust_thermo.argtypes = [P, D]
    lib.po_compute_thermo.argtypes = [P, DP]
    lib.po_compute_thermo.restype = D
    lib.po_setup_cells.argtypes = [P]
    lib.po_md_step.argtypes = [P, I]
    for f in ("po_exchange", "po_borders", "po_synchronize"):
        getattr(lib, f).argtypes = [P]
    lib.po_rank_ptr.argtypes = [P, I]
    lib.po_rank_ptr.restype = P
    for f in ("po_nlocal", "po_nghost", "po_ncells", "po_neighbor_capacity", "po_cell_capacity", "po_nsend_all"):
        getattr(lib, f).argtypes = [P]
        getattr(lib, f).restype = I
    lib.po_get_nranks.argtypes = [P, IP]
    lib.po_get_decomposition.argtypes = [P, IP, IP, DP, IP]
    lib.po_int_array.argtypes = [P, ctypes.c_char_p]
    lib.po_int_array.restype = IP
    lib.po_real_array.argtypes = [P, ctypes.c_char_p]
    lib.po_real_array.restype = DP
    lib.po_set_counts.argtypes = [P, I, I]
    lib.po_set_config.argtypes = [I, DP, IP, IP]
    for f in ("po_build_cell_lists", "po_build_neighbor_lists"):
        getattr(lib, f).argtypes = [P, P]
        getattr(lib, f).restype = I
    lib.po_build_cell_lists_stencil.argtypes = [P, P]
    lib.po_partition_cell_lists.argtypes = [P, I]
    for f in ("po_lennard_jones", "po_initial_integrate", "po_final_integrate"):
        getattr(lib, f).argtypes = [P, P]
    lib.po_reset_volatile.argtypes = [P]
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load()
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def set_config(world_size, grid, part_flags=(1, 1, 1)):
    """Rank grid the reference picks (runtime/domain/regular_6d_stencil.cpp:10-54)."""
    g = np.ascontiguousarray(grid, np.float64)
    pf = np.ascontiguousarray(part_flags, np.int32)
    out = np.zeros(3, np.int32)
    lib().po_set_config(world_size, _dp(g), _ip(pf), _ip(out))
    return tuple(int(x) for x in out)


class Rank:
    WIDTH = {"position": 3, "linear_velocity": 3, "force": 3, "mass": 1}

    def __init__(self, sim, k):
        self.sim = sim
        self.k = k
        self.p = lib().po_rank_ptr(sim.p, k)

    nlocal = property(lambda self: lib().po_nlocal(self.p))
    nghost = property(lambda self: lib().po_nghost(self.p))
    ncells = property(lambda self: lib().po_ncells(self.p))
    neighbor_capacity = property(lambda self: lib().po_neighbor_capacity(self.p))
    cell_capacity = property(lambda self: lib().po_cell_capacity(self.p))
    nsend_all = property(lambda self: lib().po_nsend_all(self.p))

    def decomposition(self):
        nr = np.zeros(6, np.int32)
        pbc = np.zeros(6, np.int32)
        sub = np.zeros(6, np.float64)
        dc = np.zeros(3, np.int32)
        lib().po_get_decomposition(self.p, _ip(nr), _ip(pbc), _dp(sub), _ip(dc))
        return {"neighbor_ranks": nr, "pbc": pbc, "subdom": sub, "dim_cells": dc}

    def real(self, name, n=None, view=False):
        n = self.nlocal if n is None else n
        w = self.WIDTH[name]
        ptr = lib().po_real_array(self.p, name.encode())
        a = np.ctypeslib.as_array(ptr, shape=(n * w,))
        a = a if view else a.copy()
        return a.reshape(n, w) if w > 1 else a

    def ints(self, name, n=None, view=False):
        n = self.nlocal if n is None else n
        ptr = lib().po_int_array(self.p, name.encode())
        a = np.ctypeslib.as_array(ptr, shape=(n,))
        return a if view else a.copy()

    def neighbor_sets(self):
        """list over local i of the neighbour index arrays (reference order)."""
        n = self.nlocal
        cap = self.neighbor_capacity
        nn = self.ints("numneighs", n)
        nl = self.ints("neighborlists", n * cap).reshape(n, cap)
        return nn, nl

    def set_particles(self, position, velocity, mass, type_, flags=None, shape=None, uid=None, nghost=0):
        n = len(position)
        nl = n - nghost
        lib().po_set_counts(self.p, nl, nghost)
        self.real("position", n, view=True)[:] = position
        self.real("linear_velocity", n, view=True)[:] = velocity
        self.real("mass", n, view=True)[:] = mass
        self.ints("type", n, view=True)[:] = type_
        self.ints("flags", n, view=True)[:] = 0 if flags is None else flags
        self.ints("shape", n, view=True)[:] = 2 if shape is None else shape
        self.ints("uid", n, view=True)[:] = 0 if uid is None else uid


class OracleSim:
    """R in-process ranks of the restated reference MD program (examples/md.py semantics)."""

    def __init__(self, grid, world_size=1, part_flags=(1, 1, 1), pbc=(1, 1, 1), particle_capacity=800000,
                 neighbor_capacity=100, cell_capacity=64, send_capacity=200000):
        g = np.ascontiguousarray(grid, np.float64)   # xmin,xmax,ymin,ymax,zmin,zmax
        pf = np.ascontiguousarray(part_flags, np.int32)
        pb = np.ascontiguousarray(pbc, np.int32)
        self.p = lib().po_create(world_size, _dp(g), _ip(pf), _ip(pb), particle_capacity, neighbor_capacity,
                                 cell_capacity, send_capacity)
        self.world = world_size
        self.grid = g
        self.ranks = [Rank(self, k) for k in range(world_size)]
        nr = np.zeros(3, np.int32)
        lib().po_get_nranks(self.p, _ip(nr))
        self.nranks = tuple(int(x) for x in nr)

    def close(self):
        if self.p:
            lib().po_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, cell_spacing, cutoff_lists, cutoff_force, dt, ntypes, epsilon, sigma6, reneigh_every, thermo_every=100):
        e = np.ascontiguousarray(epsilon, np.float64)
        s6 = np.ascontiguousarray(sigma6, np.float64)
        lib().po_set_params(self.p, cell_spacing, cutoff_lists, cutoff_force, dt, ntypes, _dp(e), _dp(s6), reneigh_every, thermo_every)
        lib().po_setup_cells(self.p)

    def compute_half(self, on=True):
        """Simulation.compute_half() (sim/simulation.py:119-120): half neighbour lists, partner updated with the opposite term."""
        lib().po_set_compute_half(self.p, 1 if on else 0)

    def copper_fcc_lattice(self, nx, ny, nz, rho, temp, ntypes):
        lib().po_copper_fcc_lattice(self.p, nx, ny, nz, rho, ntypes)
        lib().po_adjust_thermo(self.p, temp)

    def thermo(self):
        pr = ctypes.c_double(0.0)
        t = lib().po_compute_thermo(self.p, ctypes.byref(pr))
        return t, pr.value

    def step(self, ts):
        lib().po_md_step(self.p, ts)

    def exchange(self):
        lib().po_exchange(self.p)

    def borders(self):
        lib().po_borders(self.p)

    def synchronize(self):
        lib().po_synchronize(self.p)

    # single modules on rank k
    def build_cell_lists(self, k=0):
        return lib().po_build_cell_lists(self.p, self.ranks[k].p)

    def partition_cell_lists(self, k=0):
        lib().po_partition_cell_lists(self.ranks[k].p, 2)

    def build_neighbor_lists(self, k=0):
        return lib().po_build_neighbor_lists(self.p, self.ranks[k].p)

    def lennard_jones(self, k=0):
        lib().po_lennard_jones(self.p, self.ranks[k].p)

    def reset_volatile(self, k=0):
        lib().po_reset_volatile(self.ranks[k].p)

    def initial_integrate(self, k=0):
        lib().po_initial_integrate(self.p, self.ranks[k].p)

    def final_integrate(self, k=0):
        lib().po_final_integrate(self.p, self.ranks[k].p)


def md_example(nx, world_size=1, reneigh_every=20, ntypes=4, rho=0.8442, temp=1.44, **caps):
    """The system examples/md.py sets up (lines 25-63), with nx = ny = nz cells."""
    lattice = pow((4.0 / rho), (1.0 / 3.0))      # sim/copper_fcc_lattice.py:19-23
    L = nx * lattice
    sim = OracleSim([0.0, L, 0.0, L, 0.0, L], world_size=world_size, **caps)
    sim.set_params(2.5 + 0.3, 2.5 + 0.3, 2.5, 0.005, ntypes, [1.0] * (ntypes * ntypes), [1.0] * (ntypes * ntypes), reneigh_every)
    sim.copper_fcc_lattice(nx, nx, nx, rho, temp, ntypes)
    return sim
