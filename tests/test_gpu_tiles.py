"""Tile lists (csrc/tile_lists.cu: cell tiles staged in shared memory, 16-bit tile-relative neighbour lists) against the
per-particle 32-bit lists they replace on the hot path, and the fused-multiply-add arithmetic against the reference's expression
tree.  With the reference's arithmetic (option "lj_fma" = 0) and the builder's own list order (option "tile_reorder" = 0) the two
list formats must agree BIT FOR BIT: same neighbour sets in the same order, same forces, same trajectories.  The production
default stores every list in a shared-memory-conflict-aware order (same sets, another summation order: forces to 1e-12)."""
import numpy as np
import pytest

from tests.test_gpu_md import CUT, DT, SKIN, box, make_gpu, make_oracle, _reneighbor_gpu
from tests.util import by_id, rel_err_force

pytestmark = pytest.mark.gpu

EPS4 = [1.0 + 0.05 * ((k % 4) + (k // 4)) for k in range(16)]      # a symmetric, non-uniform table
SIG4 = [1.0 - 0.02 * abs((k % 4) - (k // 4)) for k in range(16)]


def _run(nx, tile, fma, steps, eps=None, sig6=None, thermo=1, reorder=0):
    ctx, n = make_gpu(nx, eps=eps, sig6=sig6)
    ctx.set_option("tile_lists", tile)
    ctx.set_option("lj_fma", fma)
    ctx.set_option("tile_reorder", reorder)
    th = ctx.md_run(0, steps, DT, CUT, CUT + SKIN, CUT + SKIN, 20, thermo)
    tag = ctx.ints("tag")
    return ctx, th, by_id(tag, ctx.real("position")), by_id(tag, ctx.real("linear_velocity")), by_id(tag, ctx.real("force"))


@pytest.mark.parametrize("nx,eps,sig6", [(6, None, None), (10, None, None), (10, EPS4, SIG4), (13, EPS4, SIG4)])
def test_tile_lists_equal_the_per_particle_lists_bit_for_bit(nx, eps, sig6):
    ctx_t, n = make_gpu(nx, eps=eps, sig6=sig6)
    ctx_p, _ = make_gpu(nx, eps=eps, sig6=sig6)
    ctx_p.set_option("tile_lists", 0)
    for c in (ctx_t, ctx_p):
        c.set_option("lj_fma", 0)
        c.set_option("tile_reorder", 0)
        _reneighbor_gpu(c)
        c.reset_volatile()
        c.lennard_jones(CUT)
    assert np.array_equal(ctx_t.ints("numneighs"), ctx_p.ints("numneighs")) and ctx_t.ints("numneighs").min() > 40
    assert np.array_equal(ctx_t.neighbors(), ctx_p.neighbors())                 # same sets, same order
    assert np.array_equal(ctx_t.real("force"), ctx_p.real("force"))
    # the 32-bit lists derived from the tiles serve the kernels that walk lists by particle
    assert ctx_t.lj_energy_virial(CUT) == ctx_p.lj_energy_virial(CUT)
    # 45 iterations of the fused loop (three list builds, wraps): identical bits
    a = _run(nx, 1, 0, 45, eps, sig6)
    b = _run(nx, 0, 0, 45, eps, sig6)
    assert np.array_equal(a[1], b[1]) and len(a[1]) == 45
    for k in (2, 3, 4):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("eps,sig6", [(None, None), (EPS4, SIG4)])
def test_fma_arithmetic_stays_inside_the_parity_budget(eps, sig6):
    """Option "lj_fma" (default on): forces of one evaluation within 1e-12 (max-norm relative) of the reference's expression tree,
    thermo of 100 iterations within 1e-9; and both against the oracle."""
    nx = 8
    a = _run(nx, 1, 1, 100, eps, sig6, reorder=1)        # the production default
    b = _run(nx, 1, 0, 100, eps, sig6)
    assert np.abs(a[1][:, 1:] - b[1][:, 1:]).max() <= 1e-9 * np.abs(b[1][:, 1:]).max()
    # one evaluation on the SAME positions (a molten state: on the initial lattice every force is zero by symmetry)
    from pairs_b200.backend import Context
    fs = []
    for fma in (1, 0):
        ctx = Context(0)
        ctx.init_domain(box(nx))
        ctx.set_option("lj_fma", fma)
        ctx.setup_cells(CUT + SKIN)
        ctx.set_lj_params(4, eps or [1.0] * 16, sig6 or [1.0] * 16)
        ctx.upload(b[2], b[3], np.ones(len(b[2])), by_id(b[0].ints("tag"), b[0].ints("type")))
        _reneighbor_gpu(ctx)
        ctx.reset_volatile()
        ctx.lennard_jones(CUT)
        fs.append(by_id(ctx.ints("tag"), ctx.real("force")))
    assert np.abs(fs[1]).max() > 10.0 and 0.0 < rel_err_force(fs[0], fs[1]) <= 1e-12
    sim = make_oracle(nx, eps=eps, sig6=sig6)
    for ts in range(100):
        sim.step(ts)
        t, p = sim.thermo()
        assert abs(a[1][ts, 1] - t) <= 1e-9 * t and abs(a[1][ts, 2] - p) <= 1e-9 * abs(p), ts


def test_fixed_particles_and_small_capacity():
    """FIXED particles get no list and no force update; a neighbour capacity that is far too small grows (resize protocol)."""
    from pairs_b200.backend import Context
    nx = 7
    ctx0, n = make_gpu(nx)
    pos, vel, typ = ctx0.real("position"), ctx0.real("linear_velocity"), ctx0.ints("type")
    flags = np.zeros(n, np.int32)
    flags[::17] = 4
    res = []
    for tile in (1, 0):
        ctx = Context(0)
        ctx.init_domain(box(nx))
        ctx.set_option("tile_lists", tile)
        ctx.set_option("lj_fma", 0)
        ctx.set_option("tile_reorder", 0)
        ctx.reserve(0, 8)
        ctx.setup_cells(CUT + SKIN)
        ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
        ctx.upload(pos, vel, np.ones(n), typ, flags)
        th = ctx.md_run(0, 25, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 1)
        tag = ctx.ints("tag")
        res.append((th, by_id(tag, ctx.real("position")), by_id(tag, ctx.real("force")), by_id(tag, ctx.ints("numneighs"))))
        assert ctx.lib.pb_neighbor_capacity(ctx.h) >= 78
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)
    assert np.array_equal(res[0][1][::17], pos[::17]) and not res[0][3][::17].any() and res[0][3][1] > 40


def test_ragged_box_and_thin_slab():
    """Boxes whose edges are no multiple of the cell spacing, fewer cells than a super-column is wide, a thin periodic slab."""
    from pairs_b200.backend import Context
    rng = np.random.default_rng(11)
    for dims in ((9.1, 17.3, 30.2), (5.7, 5.7, 5.7), (40.0, 6.1, 8.9)):
        n = int(0.8 * dims[0] * dims[1] * dims[2])
        pos = rng.random((n, 3)) * np.array(dims) * (1 - 1e-9)
        res = []
        for tile in (1, 0):
            ctx = Context(0)
            ctx.init_domain([0.0, dims[0], 0.0, dims[1], 0.0, dims[2]])
            ctx.set_option("tile_lists", tile)
            ctx.set_option("lj_fma", 0)
            ctx.set_option("tile_reorder", 0)
            ctx.setup_cells(CUT + SKIN)
            ctx.set_lj_params(1, [1.0], [1.0])
            ctx.upload(pos, np.zeros((n, 3)), np.ones(n), np.zeros(n, np.int32))
            _reneighbor_gpu(ctx)
            tag = ctx.ints("tag")
            nb = ctx.neighbors()
            tags_all = ctx.ints("tag", with_ghosts=True)
            # compare as (identity of i, sorted identities / ghost indices of its neighbours): ghosts are ordered alike in both runs
            res.append((by_id(tag, ctx.ints("numneighs")), by_id(tag, np.where(nb >= 0, nb, -1)), tags_all))
        assert np.array_equal(res[0][0], res[1][0]) and res[0][0].sum() > 0, dims
        assert np.array_equal(res[0][1], res[1][1]), dims


@pytest.mark.parametrize("nx,eps,sig6", [(6, None, None), (11, EPS4, SIG4)])
def test_conflict_aware_list_order_keeps_the_sets(nx, eps, sig6):
    """Option "tile_reorder" (default on): every list holds the same neighbours as the builder's own order -- as sets, per
    particle -- the 32-bit lists derived from the tiles follow the stored order, forces agree to 1e-12 (summation order), and 45
    iterations of the loop (three builds) stay within 1e-9."""
    from pairs_b200.backend import Context
    m = _run(nx, 1, 0, 12, eps, sig6, thermo=0)             # a molten state, not the lattice (same positions for both orders)
    typ = by_id(m[0].ints("tag"), m[0].ints("type"))
    ctxs = []
    for reorder in (1, 0):
        ctx = Context(0)
        ctx.init_domain(box(nx))
        ctx.set_option("lj_fma", 0)
        ctx.set_option("tile_reorder", reorder)
        ctx.setup_cells(CUT + SKIN)
        ctx.set_lj_params(4, eps or [1.0] * 16, sig6 or [1.0] * 16)
        ctx.upload(m[2], m[3], np.ones(len(m[2])), typ)
        _reneighbor_gpu(ctx)
        ctx.reset_volatile()
        ctx.lennard_jones(CUT)
        ctxs.append(ctx)
    a, b = ctxs
    nn_a, nn_b = a.ints("numneighs"), b.ints("numneighs")
    assert np.array_equal(nn_a, nn_b) and nn_a.min() > 40
    nb_a, nb_b = a.neighbors(), b.neighbors()
    assert not np.array_equal(nb_a, nb_b)                                      # the order did change ...
    assert np.array_equal(np.sort(nb_a, axis=1), np.sort(nb_b, axis=1))        # ... the sets did not (-1 pads sort first in both)
    fa, fb = a.real("force"), b.real("force")
    assert np.abs(fb).max() > 10.0 and 0.0 < rel_err_force(fa, fb) <= 1e-12
    assert abs(a.lj_energy_virial(CUT)[0] - b.lj_energy_virial(CUT)[0]) <= 1e-12 * abs(b.lj_energy_virial(CUT)[0])
    x = _run(nx, 1, 0, 45, eps, sig6, reorder=1)
    y = _run(nx, 1, 0, 45, eps, sig6, reorder=0)
    assert np.abs(x[1][:, 1:] - y[1][:, 1:]).max() <= 1e-9 * np.abs(y[1][:, 1:]).max()
    assert np.abs(x[2] - y[2]).max() <= 1e-9


@pytest.mark.parametrize("origin", [0.0, 4000.0])
def test_fp32_prefilter_of_the_tile_build_leaves_the_lists_unchanged(origin):
    """Option "tile_prefilter" (default on): the tile build tests candidates on fp32 copies first and decides only what fp32 cannot
    get wrong; pairs within the error band of the cutoff go through the reference's fp64 expression.  The adversarial case: 400
    particles moved to a distance of cutoff * (1 +- 0, 1e-15 ... 1e-5) from another one, in a box whose coordinates start at `origin`
    (large coordinates = coarse fp32 grid = wide band).  Lists with the pre-filter, without it, and the per-particle builder's must
    be identical, entry for entry."""
    from pairs_b200.backend import Context
    nx = 10
    src, n = make_gpu(nx)
    src.md_run(0, 30, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 0)                  # a molten state
    pos, vel, mass, typ = src.real("position"), src.real("linear_velocity"), src.real("mass"), src.ints("type")
    L = box(nx)[1]
    rng = np.random.default_rng(5)
    rc = CUT + SKIN
    inner = np.nonzero(np.all((pos > rc + 0.5) & (pos < L - rc - 0.5), axis=1))[0]
    picks = rng.choice(inner, 800, replace=False)
    eps = [0.0, 1e-15, -1e-15, 1e-13, -1e-13, 1e-11, -1e-11, 1e-9, -1e-9, 1e-7, -1e-7, 1e-5, -1e-5]
    pos = pos + origin
    for k in range(400):
        i, j = picks[2 * k], picks[2 * k + 1]
        u = rng.standard_normal(3)
        u /= np.linalg.norm(u)
        pos[j] = pos[i] + rc * (1.0 + eps[k % len(eps)]) * u
    d = pos[picks[0::2]] - pos[picks[1::2]]
    near = np.abs((d * d).sum(axis=1) - rc * rc) < 1e-3
    assert near.sum() >= 390
    res = []
    for tiles, prefilter in ((1, 1), (1, 0), (0, 0)):
        ctx = Context(0)
        ctx.init_domain([origin, origin + L] * 3)
        ctx.upload(pos, vel, mass, typ)
        ctx.setup_cells(rc)
        ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
        ctx.set_option("tile_lists", tiles)
        ctx.set_option("tile_prefilter", prefilter)
        ctx.set_option("tile_reorder", 0)
        _reneighbor_gpu(ctx)
        tag = ctx.ints("tag", True)
        nn, rows = ctx.ints("numneighs"), ctx.neighbors()
        order = np.argsort(tag[:len(nn)])
        res.append((nn[order], [tuple(tag[rows[i, :nn[i]]]) for i in order]))
    for other in res[1:]:
        assert np.array_equal(res[0][0], other[0])
        assert res[0][1] == other[1]
