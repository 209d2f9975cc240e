"""BASELINE.json's configurations at FULL size on the GPU.

C1 (examples/md.py as shipped: 32^3 cells = 131,072 atoms, 200 steps) is small enough for the oracle: thermo of the reference's
three output lines to 1e-9 and to the six digits the reference prints.
C2 (4,000,000 atoms) is checked through size-independent properties: total momentum and total force vanish (Newton's third
law over 3 x 10^8 pair terms), the total energy is conserved by velocity-Verlet, the half and the full neighbour lists describe
the same set of pairs (identical potential energy and virial), and a run is reproducible bit for bit."""
import numpy as np
import pytest

from tests.test_gpu_md import CUT, DT, SKIN, make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def test_config_c1_stock_md_py_against_the_oracle():
    ctx, n = make_gpu(32)
    assert n == 131072
    sim = make_oracle(32)
    th = ctx.md_run(0, 201, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 100)
    assert [int(r[0]) for r in th] == [0, 99, 199]
    ref = {}
    for ts in range(200):
        sim.step(ts)
        if ts in (0, 99, 199):
            ref[ts] = sim.thermo()
    for row in th:
        t_o, p_o = ref[int(row[0])]
        assert abs(row[1] - t_o) <= 1e-9 * t_o and abs(row[2] - p_o) <= 1e-9 * p_o, row
    # the lines the stock reference program prints (BASELINE.md / SURVEY.md section 6), six significant digits
    assert [f"{r[1]:.6g}\t{r[2]:.6g}" for r in th] == ["1.44\t1.21564", "0.820671\t0.692805", "0.79644\t0.67235"]
    r = sim.ranks[0]
    assert ctx.counts() == (r.nlocal, r.nghost)


def test_config_c2_four_million_atoms_invariants():
    ctx, n = make_gpu(100)
    assert n == 4000000
    ctx.md_run(0, 21, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 0)          # melt a little; ends one step after a reneighbouring

    def totals():
        m = ctx.real("mass")
        v = ctx.real("linear_velocity")
        f = ctx.real("force")
        ekin = 0.5 * float((m * (v * v).sum(axis=1)).sum())
        return (m[:, None] * v).sum(axis=0), np.abs(m[:, None] * v).sum(), f.sum(axis=0), np.abs(f).sum(), ekin

    p0, pabs, _, _, ek0 = totals()
    # adjust_thermo removed the centre-of-mass motion.  Momentum is conserved only approximately afterwards -- in the reference
    # too: between two reneighbourings the edge and corner ghosts lag one step behind (Comm.synchronize packs every send entry
    # before it unpacks, sim/comm.py:45-54), so the few pairs across box edges are not exactly antisymmetric.
    assert np.abs(p0).max() <= 1e-5 * pabs, (p0, pabs)
    # right after a reneighbouring every ghost is a fresh copy of its source: the pair forces cancel to round-off
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
    ctx.reset_volatile(); ctx.lennard_jones(CUT)
    f = ctx.real("force")
    assert np.abs(f.sum(axis=0)).max() <= 1e-9 * np.abs(f).sum()
    ep0, w0 = ctx.lj_energy_virial(CUT)
    # half lists describe the same pairs
    ctx.set_option("compute_half", 1)
    ctx.build_neighbor_lists(CUT + SKIN)
    ep_h, w_h = ctx.lj_energy_virial(CUT)
    assert abs(ep_h - ep0) <= 1e-11 * abs(ep0) and abs(w_h - w0) <= 1e-11 * abs(w0)
    assert int(ctx.ints("numneighs").sum()) < 0.56 * 78 * n
    ctx.set_option("compute_half", 0)
    ctx.build_neighbor_lists(CUT + SKIN)
    # from here on: a state that can be re-created exactly (upload in a fixed order + fresh lists)
    order = np.argsort(ctx.ints("tag"))
    state = (ctx.real("position")[order], ctx.real("linear_velocity")[order], ctx.real("mass")[order], ctx.ints("type")[order])

    def restart():
        ctx.upload(*state)
        ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
        ctx.reset_volatile(); ctx.lennard_jones(CUT)          # f(x_20): the first half kick of iteration 21 needs it

    restart()
    th_a = ctx.md_run(21, 60, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 1)
    _, _, _, _, ek1 = totals()
    end_a = (ctx.ints("tag"), ctx.real("position"))
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
    ep1, _ = ctx.lj_energy_virial(CUT)
    e0, e1 = ek0 + ep0, ek1 + ep1
    # velocity-Verlet with dt = 0.005 on the truncated, unshifted potential while the lattice is still melting: the total
    # energy moves by a small fraction of the kinetic energy (a wrong force or integrator shows up as O(1))
    print(f"C2 energy: E(20) = {e0:.9e}, E(59) = {e1:.9e}, drift / E_kin = {(e1 - e0) / ek0:.3e}")
    assert abs(e1 - e0) <= 2e-2 * abs(ek0), (e0, e1, ek0)
    p1 = totals()[0]
    assert np.abs(p1).max() <= 1e-5 * pabs, (p1, pabs)
    # reproducibility: the same 39 iterations from the same state give the same bits (no atomics on the force path, ordered
    # compactions everywhere)
    restart()
    th_b = ctx.md_run(21, 60, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 1)
    assert np.array_equal(th_a, th_b)
    assert np.array_equal(end_a[0], ctx.ints("tag")) and np.array_equal(end_a[1], ctx.real("position"))


def _mix(tag, pos):
    """order-independent identity key of a particle image: its tag and the exact bits of its coordinates"""
    k = tag.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for d in range(3):
        b = (np.ascontiguousarray(pos[:, d]) + 0.0).view(np.uint64)
        k = (k ^ (b + np.uint64(0x7F4A7C159E3779B9) + (k << np.uint64(6)) + (k >> np.uint64(2)))) * np.uint64(0xBF58476D1CE4E5B9)
    return k


def _row_hashes(keys, nn, nb, chunk=250000):
    """per particle: the sum (mod 2^64) of the keys of its neighbours -- equal for equal neighbour SETS, whatever their order"""
    out = np.zeros(len(nn), np.uint64)
    cols = np.arange(nb.shape[1])[None, :]
    for a in range(0, len(nn), chunk):
        b = min(a + chunk, len(nn))
        m = cols < nn[a:b, None]
        k = keys[np.where(m, nb[a:b], 0)]
        k[~m] = 0
        out[a:b] = k.sum(axis=1, dtype=np.uint64)
    return out


def test_config_c2_four_million_atoms_against_the_oracle():
    """C2 at size, iterations 0 and 1, against the oracle restatement (bit-identical to the reference's generated C++ on the small
    cases, tests/test_oracle_pin.py): cell index of every particle, ghost set, the neighbour SET of every particle (a 64-bit
    order-independent hash over (partner tag, exact partner coordinates), 3.1 x 10^8 entries), forces of both iterations to 1e-12
    (max-norm relative), positions and velocities after iteration 1."""
    from oracle import port
    from tests.util import by_id
    nx = 100
    ctx, n = make_gpu(nx)
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=4700000, send_capacity=600000)
    r = sim.ranks[0]
    r.ints("uid", r.nlocal, view=True)[:] = np.arange(r.nlocal)
    assert n == r.nlocal == 4000000
    # iteration 0, module by module on the GPU
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
    ctx.reset_volatile(); ctx.lennard_jones(CUT)
    sim.step(0)
    nl, ng = ctx.counts()
    assert (nl, ng) == (r.nlocal, r.nghost)
    tag_all = ctx.ints("tag", with_ghosts=True)
    order = np.argsort(tag_all[:nl])
    assert np.array_equal(tag_all[:nl][order], np.arange(nl))
    # cells (locals by tag; ghosts as a multiset of (tag, exact coordinates, cell))
    pc_g, pc_o = ctx.ints("particle_cell", with_ghosts=True), r.ints("particle_cell", nl + ng)
    uid_o = r.ints("uid", nl + ng)
    assert np.array_equal(pc_g[:nl][order], pc_o[:nl][np.argsort(uid_o[:nl])])
    pos_g, pos_o = ctx.real("position", with_ghosts=True), r.real("position", nl + ng)
    key_g, key_o = _mix(tag_all, pos_g), _mix(uid_o, pos_o)
    gg = np.sort(key_g[nl:] ^ pc_g[nl:].astype(np.uint64))
    go = np.sort(key_o[nl:] ^ pc_o[nl:].astype(np.uint64))
    assert np.array_equal(gg, go)                                    # the same ghost images in the same cells
    # neighbour sets
    nn_g, nb_g = ctx.ints("numneighs"), ctx.neighbors()
    nn_o, nb_o = r.neighbor_sets()
    assert np.array_equal(nn_g[order], nn_o[np.argsort(uid_o[:nl])]) and 70 < nn_g.mean() < 80
    h_g = _row_hashes(key_g, nn_g, nb_g)
    del nb_g
    h_o = _row_hashes(key_o, nn_o, nb_o)
    assert np.array_equal(h_g[order], h_o[np.argsort(uid_o[:nl])])
    # forces of iteration 0 (the perfect lattice: every component cancels to ~1e-14) and, after one full step, of iteration 1
    f_o = by_id(uid_o[:nl], r.real("force"))
    assert np.abs(ctx.real("force")[order] - f_o).max() <= 1e-12
    ctx.md_run(1, 2, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 0)
    sim.step(1)
    o2, u2 = np.argsort(ctx.ints("tag")), np.argsort(r.ints("uid", nl))
    for name, tol in (("position", 1e-13), ("linear_velocity", 1e-12), ("force", 1e-12)):
        a, b = ctx.real(name)[o2], r.real(name)[u2]
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300) or np.abs(a - b).max() <= 1e-12, name


def _dem_c3(plane_uids=(100000, 100001), ceiling_at=(0.8, 0.015, 0.2)):
    """BASELINE configs[2]: examples/dem.py on the 0.8 x 0.8 x 0.2 m box, 998,400 spheres + 2 half-spaces (bench.py --workload dem)."""
    import math
    from pairs_b200.backend import Context
    from tests import dem_common as dc
    domain = (0.8, 0.8, 0.2)
    ctx = Context(0)
    ctx.init_domain([0.0, domain[0], 0.0, domain[1], 0.0, domain[2]], pbc=(1, 1, 0), partitioner=1)
    ctx.dem_enable(dc.C)
    ctx.dem_set_params(dc.DT, math.pi, dc.KAPPA, dc.LN_DRY, dc.COLLISION_TIME, dc.RHO_P, dc.RHO_F, dc.G, dc.NTYPES, dc.FS, dc.FD)
    ctx.setup_cells(dc.CELL)
    g = ctx.dem_sc_grid(domain[0], domain[1], domain[2], dc.SPACING, dc.DIAMETER, dc.MIN_D, dc.MAX_D, dc.V0, dc.RHO_P, dc.NTYPES)
    ns = len(g["uid"])
    n = ns + 2
    pos, vel, normal = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    mass, radius = np.ones(n), np.zeros(n)
    uid, typ, flags, shape = (np.zeros(n, np.int32) for _ in range(4))
    pos[:ns], vel[:ns], mass[:ns], radius[:ns], uid[:ns], typ[:ns] = g["position"], g["linear_velocity"], g["mass"], g["radius"], g["uid"], g["type"]
    # data/planes.input of the reference as it is (oracle/build_ref.py copies it): the ceiling keeps the y of the stock thin box
    # (at this size the uids of the plane file, 100000 and 100001, also belong to two spheres -- in the reference as well)
    for k, (u, p_, nrm) in enumerate([(plane_uids[0], (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), (plane_uids[1], ceiling_at, (0.0, 0.0, -1.0))]):
        uid[ns + k], pos[ns + k], normal[ns + k], flags[ns + k], shape[ns + k] = u, p_, nrm, 13, 1
    ctx.upload(pos, vel, mass, typ, flags, uid, shape)
    ctx.dem_upload("radius", radius)
    ctx.dem_upload("normal", normal)
    ctx.dem_stage("update_mass_and_inertia")
    return ctx, n, dc


def test_config_c3_one_million_spheres_against_the_reference(tmp_path):
    """C3 at size against the reference's OWN generated C++ (oracle/_ref variant dem_bench, run here on a host core): after 21
    iterations of the loop the same particles in the same order, bit-identical positions and linear velocities, the same ghost
    count, the same cell of every particle, empty contact rows on both sides (the bed has not formed yet: the settled bed is
    covered by invariants below and, against the reference, on the 420-sphere case of tests/test_gpu_dem.py)."""
    from oracle import ref, ref_worker
    if not ref.available("dem_bench"):
        pytest.skip("oracle/_ref variant dem_bench not built")
    last = 20
    z = ref_worker.dump_dem_end("dem_bench", str(tmp_path / "c3.npz"), last, ["position", "linear_velocity", "uid", "num_contacts", "particle_cell"])
    ctx, n, dc = _dem_c3()
    ctx.dem_run(dc.CELL, 0, last + 1)
    nl, ng = ctx.counts()
    assert (nl, ng) == (int(z["nlocal"][0]), int(z["nghost"][0])) and nl == n == 998402
    assert np.array_equal(ctx.ints("uid"), z["uid"])                        # DEM keeps the particle order (no re-sort before 200)
    dx = np.abs(ctx.real("position") - z["position"]).max()
    dv = np.abs(ctx.real("linear_velocity") - z["linear_velocity"]).max() / np.abs(z["linear_velocity"]).max()
    print(f"C3 after {last + 1} iterations: max |dx| = {dx:.3e} m, max |dv| / max |v| = {dv:.3e}")
    assert dx == 0.0 and dv == 0.0                                           # bit for bit: gravity + the linear part of euler
    assert np.array_equal(ctx.ints("particle_cell"), z["particle_cell"])
    c = ctx.dem_download_contacts(nl)
    assert not c["num_contacts"].any() and not z["num_contacts"].any()


def test_config_c3_settled_bed_invariants():
    """C3 at size, 4000 iterations (the bed has formed: ~5 contacts per sphere).  Properties that hold whatever the size: nothing is
    lost or duplicated, every sphere lies between the two planes, contact rows are symmetric (j in the row of i <=> i in the row
    of j, for sphere-sphere contacts), partners of a contact touch or nearly touch, no row exceeds the capacity."""
    ctx, n, dc = _dem_c3(plane_uids=(100000000, 100000001))
    ctx.dem_run(dc.CELL, 0, 4000)
    nl, _ = ctx.counts()
    assert nl == n
    uid = ctx.ints("uid")
    assert len(np.unique(uid)) == n
    pos, shape = ctx.real("position"), ctx.ints("shape")
    sph = shape == 0
    radius = ctx.dem_download("radius", nl)
    assert np.isfinite(pos[sph]).all() and (pos[sph, 2] > -1e-4).all() and (pos[sph, 2] < 0.2).all()
    c = ctx.dem_download_contacts(nl)
    num, lists = c["num_contacts"], c["contact_lists"]
    assert num.max() <= dc.C and 3.0 < num[sph].mean() < 8.0
    order = np.argsort(uid)
    row_of = lambda u: order[np.searchsorted(uid[order], u)]      # noqa: E731  (uid -> row; every listed uid exists)
    m = np.arange(lists.shape[1])[None, :] < num[:, None]
    i_idx = np.broadcast_to(np.arange(n)[:, None], lists.shape)[m]
    j_idx = row_of(lists[m])
    assert np.array_equal(uid[j_idx], lists[m])
    both = sph[i_idx] & sph[j_idx]                                            # (the half-spaces are FIXED: they keep no rows)
    a, b = i_idx[both], j_idx[both]
    fwd = np.unique(a.astype(np.int64) * n + b)
    bwd = np.unique(b.astype(np.int64) * n + a)
    assert np.array_equal(fwd, bwd)
    d = pos[a] - pos[b]
    d[:, 0] -= 0.8 * np.round(d[:, 0] / 0.8)                                  # periodic in x and y
    d[:, 1] -= 0.8 * np.round(d[:, 1] / 0.8)
    gap = np.sqrt((d * d).sum(axis=1)) - radius[a] - radius[b]
    assert gap.max() < 0.05 * dc.DIAMETER                                      # a row entry outlives the contact by at most one step
