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


def test_fp32_prefilter_band_of_the_tile_build_never_misclassifies():
    """The tile build (csrc/tile_lists.cu pb_k_tile_build32) tests list candidates in fp32 first and lets fp32 decide only outside a
    band of +- 2 err around the squared cutoff, err = 2 sqrt(3) rc (2^-23 A + 2^-24 rc) + 8 * 2^-24 rc^2 + 3 (2^-23 A)^2 with A = the
    particle's largest coordinate plus the cutoff.  Host restatement of that arithmetic (positions rounded to fp32, differences, one
    product and two fused multiply-adds in fp32) on ten million random pairs placed close to the cutoff, at coordinate magnitudes
    from 10 to 10^5: the fp32 value is within err of the reference's fp64 expression, hence whatever fp32 decides agrees with
    rsq < cutsq in fp64.  (The GPU side -- identical lists with and without the pre-filter on adversarial pairs -- is
    tests/test_gpu_tiles.py.)"""
    rc = 2.8
    cutsq = rc * rc
    u = 2.0 ** -24
    rng = np.random.default_rng(11)
    for scale in (10.0, 170.0, 4000.0, 1.0e5):
        n = 2_500_000
        xi = rng.uniform(-scale, scale, (n, 3))
        d = rng.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        r = rc * (1.0 + rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-9.0, -1.0, n))        # |r - rc| / rc from 1e-9 to 0.1
        xj = xi + d * r[:, None]
        # the reference's fp64 expression (md_math.h pb_pair_rsq): (dx*dx + dy*dy) + dz*dz
        e = xi - xj
        rsq64 = (e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]) + e[:, 2] * e[:, 2]
        # the kernel's fp32 arithmetic: fmaf(dz, dz, fmaf(dy, dy, dx * dx)) on rounded positions
        f = (xi.astype(np.float32) - xj.astype(np.float32)).astype(np.float32)
        p = (f[:, 0] * f[:, 0]).astype(np.float32)
        q = (f[:, 1].astype(np.float64) * f[:, 1].astype(np.float64) + p.astype(np.float64)).astype(np.float32)
        rsq32 = (f[:, 2].astype(np.float64) * f[:, 2].astype(np.float64) + q.astype(np.float64)).astype(np.float32)
        A = np.abs(xi).max(axis=1) + rc
        err = 2.0 * np.sqrt(3.0) * rc * (2.0 * u * A + u * rc) + 8.0 * u * cutsq + 3.0 * (2.0 * u * A) ** 2
        near = np.abs(r - rc) < 0.02 * rc                                       # the bound is stated for pairs around the cutoff
        assert (np.abs(rsq32.astype(np.float64) - rsq64)[near] <= err[near]).all(), scale
        cut_lo = np.nextafter((cutsq - 2.0 * err).astype(np.float32), np.float32(-np.inf))      # rounded outwards, as the kernel does
        cut_hi = np.nextafter((cutsq + 2.0 * err).astype(np.float32), np.float32(np.inf))
        decided_in, decided_out = rsq32 <= cut_lo, rsq32 >= cut_hi
        assert (rsq64[decided_in] < cutsq).all() and (rsq64[decided_out] >= cutsq).all(), scale
        assert decided_in.sum() > 1000 and decided_out.sum() > 1000             # (both outcomes occur at every magnitude)
        if scale <= 170.0:                                                      # at the benchmark's box size fp32 leaves only a thin
            far = np.abs(r - rc) > 1e-4 * rc                                    # shell to fp64: |r - rc| < 1e-4 rc, ~0.02 entries per atom
            assert (decided_in | decided_out)[far].all()
