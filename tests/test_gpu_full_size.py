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
