"""Pins the oracle (CPU, no GPU needed):
  1. oracle/pairs_oracle.c (the restatement) against the committed golden fixtures tests/golden/*.npz, which were produced by
     the reference's own generated C++ (tests/golden/make_golden_md.py);
  2. where oracle/_ref exists (built from /root/reference by oracle/build_ref.py), the restatement against the reference's
     generated code directly: every step of the run, bit for bit, plus single modules on identical arrays."""
import os

import numpy as np
import pytest

from oracle import port, ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_port(nx, reneigh, steps, keep, half=False):
    sim = port.md_example(nx, reneigh_every=reneigh, particle_capacity=60000, send_capacity=60000)
    if half:
        sim.compute_half()
    r = sim.ranks[0]
    temps, kept, counts, types0 = [], {}, [], None
    for ts in range(steps + 1):
        sim.step(ts)
        if ts == 0:
            types0 = r.ints("type")
            kept["lists"] = (r.ints("numneighs", r.nlocal), r.ints("neighborlists", r.nlocal * r.neighbor_capacity).reshape(r.nlocal, -1))
        temps.append(sim.thermo()[0])
        counts.append((r.nlocal, r.nghost))
        if ts in keep:
            kept[ts] = {n: r.real(n) for n in ("position", "linear_velocity", "force")}
    return np.array(temps), counts, kept, types0


@pytest.mark.parametrize("variant,nx,reneigh,steps", [("md_t1", 8, 20, 100), ("md_t2", 12, 5, 60)])
def test_restatement_matches_reference_golden_bit_for_bit(variant, nx, reneigh, steps):
    z = np.load(os.path.join(GOLD, f"{variant}.npz"))
    keep = [int(k) for k in z["steps_kept"]]
    temps, counts, kept, types = run_port(nx, reneigh, steps, keep)
    assert np.array_equal(temps, z["temperature"])            # serial left-to-right sums: identical bits
    assert [c[0] for c in counts] == list(z["nlocal"]) and [c[1] for c in counts] == list(z["nghost"])
    assert np.array_equal(types, z["type"])                   # glibc rand() stream from seed 1
    for k in keep:
        for name in ("position", "linear_velocity", "force"):
            if f"{name}_{k}" in z:
                assert np.array_equal(kept[k][name], z[f"{name}_{k}"]), (k, name)


def test_half_list_restatement_matches_reference_golden_bit_for_bit():
    """compute_half() (sim/interaction.py:107-113, ir/apply.py:111-125): the serial reference applies the partner updates in
    loop order, the restatement does the same -> identical bits, including the half lists themselves."""
    z = np.load(os.path.join(GOLD, "md_half_t1.npz"))
    keep = [int(k) for k in z["steps_kept"]]
    temps, counts, kept, types = run_port(8, 20, 100, keep, half=True)
    assert np.array_equal(temps, z["temperature"])
    assert [c[0] for c in counts] == list(z["nlocal"]) and [c[1] for c in counts] == list(z["nghost"])
    for k in keep:
        for name in ("position", "linear_velocity", "force"):
            assert np.array_equal(kept[k][name], z[f"{name}_{k}"]), (k, name)
    nn, lists = kept["lists"]
    assert np.array_equal(nn, z["numneighs_0"])
    w = z["neighborlists_0"].shape[1]
    for i in range(len(nn)):
        assert np.array_equal(lists[i, :nn[i]], z["neighborlists_0"][i, :nn[i]]) and nn[i] <= w
    # half lists: every stored partner has a larger index; the full-list run of the same system sees the same physics
    assert all((lists[i, :nn[i]] > i).all() for i in range(len(nn)))
    full = np.load(os.path.join(GOLD, "md_t1.npz"))
    assert np.abs(temps - full["temperature"]).max() <= 1e-12


def test_golden_thermo_reproduces_reference_stdout():
    """The 6-digit thermo lines the stock reference program prints for examples/md.py start at T = 1.44;
    the nx = 8 golden starts from the same adjust_thermo target."""
    z = np.load(os.path.join(GOLD, "md_t1.npz"))
    assert f"{z['temperature'][0]:.6g}" == "1.44"


@pytest.mark.skipif(not ref.available("md_t1"), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_reference_run_every_step(tmp_path):
    from oracle import ref_worker
    snaps = ref_worker.dump("md_t1", str(tmp_path / "t1.npz"))
    sim = port.md_example(8, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    for ts, s in enumerate(snaps):
        sim.step(ts)
        assert (r.nlocal, r.nghost) == (s["nlocal"], s["nghost"])
        for name in ("position", "linear_velocity", "force"):
            assert np.array_equal(r.real(name), s[name]), (ts, name)


@pytest.mark.skipif(not ref.available("md_t1"), reason="oracle/_ref not built (needs /root/reference)")
def test_single_modules_match_reference_on_identical_arrays():
    """build_cell_lists / partition_cell_lists / build_neighbor_lists / lennard_jones / integrators of the generated code,
    called directly on the restatement's arrays (ghosts included)."""
    prog = ref.RefProgram("md_t1")
    sim = port.md_example(8, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    for ts in range(25):
        sim.step(ts)
    # state right after a reneighbouring step (ts = 19) + 5 steps: rebuild lists from the current arrays on both sides
    sim.exchange()
    sim.borders()
    assert sim.build_cell_lists() == 0
    sim.partition_cell_lists()
    assert sim.build_neighbor_lists() == 0
    n, ng = r.nlocal, r.nghost
    tot = n + ng
    d = r.decomposition()
    pos, flags, shape = r.real("position", tot), r.ints("flags", tot), r.ints("shape", tot)
    out = prog.build_lists(pos, flags, shape, n, ng, d["subdom"])
    assert out["ncells"] == r.ncells and np.array_equal(out["dim_cells"], d["dim_cells"])
    assert np.array_equal(out["stencil"], r.ints("stencil", 27))
    assert np.array_equal(out["particle_cell"], r.ints("particle_cell", tot))
    nn, nl = r.neighbor_sets()
    assert np.array_equal(out["numneighs"], nn)
    for i in range(0, n, 97):
        assert np.array_equal(out["neighborlists"][i, :nn[i]], nl[i, :nn[i]])
    # force + integrators
    types, mass, vel = r.ints("type", tot), r.real("mass", tot), r.real("linear_velocity", tot)
    f_ref = np.zeros((tot, 3))
    ones = np.ones(16)
    prog.lennard_jones(out["neighbor_capacity"], n, out["numneighs"], np.ascontiguousarray(out["neighborlists"]), flags, pos, types,
                       f_ref, ones, ones)
    sim.reset_volatile()
    sim.lennard_jones()
    assert np.array_equal(r.real("force"), f_ref[:n])
    p2, v2 = pos.copy(), vel.copy()
    prog.initial_integrate(n, flags, f_ref, mass, v2, p2)
    sim.initial_integrate()
    assert np.array_equal(r.real("position"), p2[:n]) and np.array_equal(r.real("linear_velocity"), v2[:n])
    prog.final_integrate(n, flags, f_ref, mass, v2)
    sim.final_integrate()
    assert np.array_equal(r.real("linear_velocity"), v2[:n])


def test_multirank_emulation_consistent_with_single_rank():
    """The R-rank in-process emulation (what the reference does under MPI) conserves the global system.  It is NOT
    bit-equal to the single-rank run, by construction of the reference: Comm.synchronize refreshes forwarded (edge / corner)
    ghosts one step late (see po_synchronize), and WHICH images are forwarded depends on the decomposition.  So: identical
    at ts = 0 (all ghosts fresh after borders), then close (the lag perturbs forces at the 1e-2 level for a few
    boundary pairs), particle set conserved, every particle near its owner's sub-box."""
    nx, steps = 8, 45
    for world in (2, 4, 8):
        many = port.md_example(nx, world_size=world, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
        ref1 = port.md_example(nx, world_size=1, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
        assert sum(r.nlocal for r in many.ranks) == ref1.ranks[0].nlocal == 4 * nx ** 3
        assert many.nranks == {2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
        # initial state: same atoms, velocities equal up to the summation order of the drift / temperature reductions
        v_m = np.sort(np.concatenate([r.real("linear_velocity") for r in many.ranks]), axis=0)
        v_1 = np.sort(ref1.ranks[0].real("linear_velocity"), axis=0)
        assert np.abs(v_m - v_1).max() <= 1e-13
        for ts in range(steps):
            many.step(ts)
            ref1.step(ts)
            t_m, t_1 = many.thermo()[0], ref1.thermo()[0]
            tol = 1e-13 if ts == 0 else 1e-2
            assert abs(t_m - t_1) <= tol * t_1, (world, ts, t_m, t_1)
        assert sum(r.nlocal for r in many.ranks) == 4 * nx ** 3
        for r in many.ranks:      # ts = 39 was a reneighbouring step; 5 steps later particles are within skin of their box
            sub = r.decomposition()["subdom"]
            p = r.real("position")
            for d in range(3):
                assert np.all(p[:, d] >= sub[2 * d] - 0.3) and np.all(p[:, d] <= sub[2 * d + 1] + 0.3)


@pytest.mark.skipif(not (ref.available("dem_t1") and ref.available("dem_cn_t1")), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_per_cell_neighbor_lists_change_no_bit(tmp_path):
    """build_cell_lists(..., store_neighbors_per_cell=True) (sim/cell_lists.py:174-206) is a storage variant of the same
    traversal: the reference's generated dem.cpp produces identical state with and without it.  This is what allows the
    backend to accept the flag without a second code path."""
    from oracle import ref_worker
    keep = [0, 100, 250]
    a = ref_worker.dump_dem("dem_t1", str(tmp_path / "a.npz"), 251, keep)
    b = ref_worker.dump_dem("dem_cn_t1", str(tmp_path / "b.npz"), 251, keep)
    assert sorted(a.files) == sorted(b.files) and len(a.files) > 100
    assert all(np.array_equal(a[k], b[k]) for k in a.files)
