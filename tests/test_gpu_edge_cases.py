"""Edge cases of the MD path on the GPU against the oracle restatement: ragged (non-cubic) boxes with mixed periodicity, random
gases with particles exactly on box faces and cell edges, a dense cluster that overflows the reference's default cell and
neighbour capacities, two-particle and empty systems.  Same bars as everywhere: cell assignment, ghosts and neighbour sets
bit-exact, forces 1e-12, short trajectories 1e-9."""
import numpy as np
import pytest

from tests.util import by_id, f2i, neighbor_rows, rel_err_force, rows_sorted

pytestmark = pytest.mark.gpu

CUT, SKIN, DT = 2.5, 0.3, 0.005


def gas(rng, grid, n, min_dist, edge_values=True):
    """n random points in the box, pairwise at least min_dist apart (incl. periodic images: measured in the box metric); the
    first seven carry coordinates exactly on box faces and on cell edges of the binning grid (origin lo - 2.8, spacing 2.8)."""
    lo, hi = np.array(grid[0::2]), np.array(grid[1::2])
    L = hi - lo
    pts = []

    def far_enough(p):
        for q in pts:
            d = np.abs(p - q)
            d = np.minimum(d, L - d)
            if (d * d).sum() < min_dist * min_dist:
                return False
        return True

    special = []
    if edge_values and n >= 8:
        special = [(0, lo[0]), (0, hi[0]), (1, lo[1]), (2, hi[2]), (0, (lo[0] - 2.8) + 3 * 2.8), (1, (lo[1] - 2.8) + 2 * 2.8),
                   (2, np.nextafter((lo[2] - 2.8) + 2 * 2.8, -np.inf))]
    while len(pts) < n:
        p = lo + rng.random(3) * L
        if len(pts) < len(special):
            axis, value = special[len(pts)]
            p[axis] = value
        if far_enough(p):
            pts.append(p)
    return np.array(pts)


def make_pair(grid, pbc, ntypes, eps, sig6, pos, vel, mass, typ):
    from oracle import port
    from pairs_b200.backend import Context
    sim = port.OracleSim(grid, pbc=pbc, particle_capacity=60000, send_capacity=60000)
    sim.set_params(CUT + SKIN, CUT + SKIN, CUT, DT, ntypes, eps, sig6, 5, 1)
    r = sim.ranks[0]
    r.set_particles(pos, vel, mass, typ, uid=np.arange(len(pos)))
    ctx = Context(0)
    ctx.init_domain(grid, pbc=pbc)
    ctx.setup_cells(CUT + SKIN)
    ctx.set_lj_params(ntypes, eps, sig6)
    if len(pos):
        ctx.upload(pos, vel, mass, typ)
    return sim, r, ctx


def compare_step0(sim, r, ctx):
    sim.step(0)
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
    nl, ng = ctx.counts()
    assert (nl, ng) == (r.nlocal, r.nghost)
    tot = nl + ng
    if tot == 0:
        return
    g_rows = rows_sorted(ctx.ints("tag", True).astype(np.int64), *f2i(ctx.real("position", True)).T, ctx.ints("particle_cell", True).astype(np.int64))
    o_rows = rows_sorted(r.ints("uid", tot).astype(np.int64), *f2i(r.real("position", tot)).T, r.ints("particle_cell", tot).astype(np.int64))
    assert np.array_equal(g_rows, o_rows)                    # identity, exact coordinates (after wrap / shift) and cell of everything
    nn_o, nl_o = r.neighbor_sets()
    o = neighbor_rows(r.ints("uid", tot), r.real("position", tot), nn_o, nl_o, nl)
    g = neighbor_rows(ctx.ints("tag", True), ctx.real("position", True), ctx.ints("numneighs"), ctx.neighbors(), nl)
    assert g.shape == o.shape and np.array_equal(g, o)       # neighbour sets
    ctx.reset_volatile()
    ctx.lennard_jones(CUT)
    assert rel_err_force(by_id(ctx.ints("tag"), ctx.real("force")), by_id(r.ints("uid"), r.real("force"))) <= 1e-12


@pytest.mark.parametrize("grid,pbc,n,seed", [
    ([0.0, 11.3, 0.0, 17.9, 0.0, 23.1], (1, 1, 1), 900, 1),          # ragged box, fully periodic
    ([-3.0, 9.5, 2.0, 14.7, -8.0, 6.3], (1, 0, 1), 700, 2),          # shifted origin, open in y
    ([0.0, 30.0, 0.0, 8.5, 0.0, 8.5], (0, 0, 0), 600, 3),            # a slab, no periodicity at all (no ghosts)
    ([0.0, 8.41, 0.0, 8.41, 0.0, 8.41], (1, 1, 1), 300, 4),          # the smallest box three cells wide
])
def test_ragged_boxes_random_gas(grid, pbc, n, seed):
    rng = np.random.default_rng(seed)
    ntypes = 3
    eps = list(0.8 + 0.4 * rng.random(9))
    sig6 = list(0.9 + 0.2 * rng.random(9))
    pos = gas(rng, grid, n, 0.85)
    vel = rng.normal(0.0, 1.0, (n, 3))
    mass = 0.5 + rng.random(n)
    typ = rng.integers(0, ntypes, n).astype(np.int32)
    sim, r, ctx = make_pair(grid, pbc, ntypes, eps, sig6, pos, vel, mass, typ)
    compare_step0(sim, r, ctx)
    # a short free run: positions leave through open faces / wrap through periodic ones
    ctx.upload(pos, vel, mass, typ)
    worst = 0.0
    for ts in range(1, 16):
        sim.step(ts)
    th = ctx.md_run(0, 16, DT, CUT, CUT + SKIN, CUT + SKIN, 5, 1)
    t_o = sim.thermo()[0]
    worst = abs(th[-1, 1] - t_o) / t_o
    assert worst <= 1e-9, worst
    assert ctx.counts() == (r.nlocal, r.nghost)
    p_o = by_id(r.ints("uid"), r.real("position"))
    assert (np.abs(by_id(ctx.ints("tag"), ctx.real("position")) - p_o) <= 1e-9 * np.maximum(1.0, np.abs(p_o))).all()


def test_dense_cluster_overflows_reference_capacities():
    """130 particles inside one cell-sized blob: more than the reference's cell_capacity (64) and neighbor_capacity (100) ->
    its resize protocol runs (oracle), the CSR cell list / grown ELLPACK lists here."""
    rng = np.random.default_rng(7)
    grid = [0.0, 14.0, 0.0, 14.0, 0.0, 14.0]
    blob = gas(rng, [5.7, 8.3, 5.7, 8.3, 5.7, 8.3], 130, 0.45, edge_values=False)
    rest = gas(rng, grid, 200, 1.0, edge_values=False)
    rest = rest[np.abs(rest - 7.0).max(axis=1) > 2.4]
    pos = np.vstack([blob, rest])
    n = len(pos)
    sim, r, ctx = make_pair(grid, (1, 1, 1), 1, [1.0], [1.0], pos, np.zeros((n, 3)), np.ones(n), np.zeros(n, np.int32))
    compare_step0(sim, r, ctx)
    assert ctx.ints("numneighs").max() > 100 and r.neighbor_capacity > 100 and r.cell_capacity > 64


def test_two_particles_and_empty_system():
    grid = [0.0, 9.0, 0.0, 9.0, 0.0, 9.0]
    pos = np.array([[0.4, 4.5, 4.5], [8.8, 4.5, 4.5]])       # interact only through the periodic face
    vel = np.array([[0.3, 0.0, 0.0], [-0.3, 0.1, 0.0]])
    sim, r, ctx = make_pair(grid, (1, 1, 1), 1, [1.0], [1.0], pos, vel, np.ones(2), np.zeros(2, np.int32))
    compare_step0(sim, r, ctx)
    f = by_id(ctx.ints("tag"), ctx.real("force"))
    # Newton's third law through the periodic image: the two shifted coordinates (x - L, x + L) round differently, so the two
    # evaluations agree to round-off, not bit for bit -- in the reference as well
    assert abs(f[0, 0]) > 0.1 and np.abs(f[0] + f[1]).max() <= 1e-12 * np.abs(f[0]).max()
    # empty: every stage and the native loop are no-ops that do not fail
    sim, r, ctx = make_pair(grid, (1, 1, 1), 1, [1.0], [1.0], np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0), np.zeros(0, np.int32))
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
    ctx.reset_volatile(); ctx.lennard_jones(CUT); ctx.initial_integrate(DT); ctx.final_integrate(DT)
    assert ctx.counts() == (0, 0)
    ctx.md_run(0, 6, DT, CUT, CUT + SKIN, CUT + SKIN, 5, 0)
    assert ctx.counts() == (0, 0)
