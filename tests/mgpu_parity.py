"""N-rank parity of the MD path against the SINGLE-RANK oracle of the same global system ("single-rank CPU run of
the same global system: compare thermo, per-uid positions/forces, global neighbour-pair set").

Run by every rank of a torchrun job -- tests/scripts/mgpu_check.py (pytest, needs >= 2 GPUs) and bench.py (N > 1, before its timed
region, so that the driver's scaling run carries the evidence).  TEST INFRASTRUCTURE: the oracle is the checker here, never the
thing measured.

What can and what cannot be compared across decompositions.  The reference's `Comm.synchronize` packs every send entry before it
unpacks any (sim/comm.py:45-54), so a ghost that was forwarded through two or three dimensions is one step stale per forwarding
level on the steps BETWEEN reneighbourings; which particles see such ghosts depends on where the sub-box corners are, i.e. a
different rank grid is a (slightly) different trajectory -- in the reference itself (tests/test_oracle_pin.py).  Ghosts made by
`Comm.borders` are always fresh.  Hence three cases:

  A  reneighbour EVERY step (exchange + borders + cell lists + neighbour lists each iteration, no synchronize): the N-rank run is
     the single-rank run up to summation order.  30 iterations: thermo every step <= 1e-9, every particle's end position <= 1e-9
     (particles identified by their exact initial lattice position), migration included.
  B  one reneighbouring + force evaluation on IDENTICAL inputs: the single-rank oracle's (molten) state after 25 iterations is
     dealt out to the ranks by sub-box; GLOBAL directed neighbour-pair set identical, forces per particle <= 1e-12 (max-norm
     relative) against the oracle's own modules run on the same state.
  C  45 iterations of the standard loop, overlap of halo refresh and interior forces on and off: bit-identical to each other, and
     per-rank counts / thermo <= 1e-9 / end positions against the oracle restatement holding the SAME N-rank decomposition in one
     process (the restatement's single-rank mode is pinned bit for bit to the reference's generated C++, tests/test_oracle_pin.py).
"""
import numpy as np

RHO, TEMP, NTYPES = 0.8442, 1.44, 4
DT, CUT, SKIN = 0.005, 2.5, 0.3


def _rows_sorted(a):
    """lexicographic row order of an [n][3] array"""
    return np.lexsort((a[:, 2], a[:, 1], a[:, 0]))


def _new_ctx(backend, dist, rank, world, local, grid):
    ctx = backend.Context(local)
    ctx.init_domain(grid, world_size=world, rank=rank)
    ids = [backend.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.nccl_init(ids[0])
    return ctx


def _setup(ctx, nx):
    n = ctx.copper_fcc_lattice(nx, nx, nx, RHO, NTYPES)
    ctx.adjust_thermo(TEMP)
    ctx.set_lj_params(NTYPES, [1.0] * (NTYPES * NTYPES), [1.0] * (NTYPES * NTYPES))
    return n


def _by_tag(tag_now, tag0):
    """index into the initial arrays for every current particle (tags are unique per rank set)"""
    order = np.argsort(tag0)
    pos = np.searchsorted(tag0[order], tag_now)
    assert np.array_equal(tag0[order][pos], tag_now)
    return order[pos]


def check(backend, dist, rank, world, local, nx=12, steps_a=30, steps_c=45):
    """-> dict (rank 0; other ranks get {'ok': ...} broadcast).  Raises AssertionError on any mismatch."""
    a = pow(4.0 / RHO, 1.0 / 3.0)
    L = nx * a
    grid = [0.0, L, 0.0, L, 0.0, L]
    report = {}

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # ---- the initial lattice, as every rank generated its part: the union must be the single-rank lattice, bit for bit ----
    ctx = _new_ctx(backend, dist, rank, world, local, grid)
    _setup(ctx, nx)
    x0, tag0 = ctx.real("position"), ctx.ints("tag")
    v0 = ctx.real("linear_velocity")

    # ---- case A: reneighbour every step, against the single-rank oracle ----
    th = ctx.md_run(0, steps_a, DT, CUT, CUT + SKIN, CUT + SKIN, 1, 1)
    x1, tag1 = ctx.real("position"), ctx.ints("tag")
    # a migrated particle keeps its tag; tags are unique across ranks (rank stride), so gather everything and match on rank 0
    ga = gather({"x0": x0, "v0": v0, "tag0": tag0, "x1": x1, "tag1": tag1, "th": th, "counts": ctx.counts(),
                 "decomp": ctx.decomposition()})
    ctx.close()

    # ---- case B: the oracle's state after 25 iterations, dealt out by sub-box; one reneighbouring + one force evaluation ----
    state = [None]
    if rank == 0:
        from oracle import port
        n_glob_b = 4 * nx ** 3
        sim_b = port.md_example(nx, world_size=1, reneigh_every=20, particle_capacity=4 * n_glob_b + 4096, send_capacity=2 * n_glob_b + 4096)
        rb = sim_b.ranks[0]
        rb.ints("uid", rb.nlocal, view=True)[:] = np.arange(rb.nlocal)
        for ts in range(25):
            sim_b.step(ts)
        ob = np.argsort(rb.ints("uid"))
        state[0] = {"pos": rb.real("position")[ob], "vel": rb.real("linear_velocity")[ob], "mass": rb.real("mass")[ob], "type": rb.ints("type")[ob]}
    dist.broadcast_object_list(state, src=0)
    sb = state[0]
    ctx = _new_ctx(backend, dist, rank, world, local, grid)
    sub = ctx.decomposition()["subdom"]
    mine = np.ones(len(sb["pos"]), bool)
    for d in range(3):
        # 5 iterations after the last reneighbouring some particles sit just outside the global box (the wrap happens in
        # exchange): they belong to the rank at that face, whose exchange then wraps / hands them over like any other leaver
        lo = -np.inf if sub[2 * d] <= grid[2 * d] + 1e-9 * L else sub[2 * d]
        hi = np.inf if sub[2 * d + 1] >= grid[2 * d + 1] - 1e-9 * L else sub[2 * d + 1]
        mine &= (sb["pos"][:, d] >= lo) & (sb["pos"][:, d] < hi)
    gid_mine = np.nonzero(mine)[0].astype(np.int32)
    ctx.setup_cells(CUT + SKIN)
    ctx.set_lj_params(NTYPES, [1.0] * (NTYPES * NTYPES), [1.0] * (NTYPES * NTYPES))
    ctx.upload(sb["pos"][mine], sb["vel"][mine], sb["mass"][mine], sb["type"][mine], None, gid_mine)
    ctx.exchange(); ctx.borders(); ctx.build_cell_lists(); ctx.build_neighbor_lists(CUT + SKIN)
    ctx.reset_volatile(); ctx.lennard_jones(CUT)
    nl, ng = ctx.counts()
    uid_all = ctx.ints("uid", with_ghosts=True)
    nb = ctx.neighbors()
    nn = ctx.ints("numneighs")
    f = ctx.real("force")
    mask = np.arange(nb.shape[1])[None, :] < nn[:, None]
    ii = np.broadcast_to(np.arange(nl)[:, None], nb.shape)[mask]
    jj = nb[mask]
    gb = gather({"uid_i": uid_all[ii], "uid_j": uid_all[jj], "uid": uid_all[:nl], "force": f, "counts": (nl, ng), "dealt": int(mine.sum())})
    ctx.close()

    # ---- case C: 45 iterations, overlap on / off, against the N-rank restatement ----
    ctx = _new_ctx(backend, dist, rank, world, local, grid)

    def run_c(overlap):
        ctx.set_option("overlap_comm", overlap)
        n = _setup(ctx, nx)
        log = {}
        chunk = 7      # multi-step calls: fused integrators + ghost refresh overlapped with the interior force
        for c in range(0, steps_c, chunk):
            e = min(c + chunk, steps_c)
            for row in ctx.md_run(c, e, DT, CUT, CUT + SKIN, CUT + SKIN, 20, chunk):
                log[int(row[0])] = float(row[1])
            log[("counts", e - 1)] = ctx.counts()
        return n, log, ctx.ints("tag"), ctx.real("position"), ctx.real("linear_velocity")

    n_c, log, tag_c, pos_c, vel_c = run_c(1)
    n_c2, log2, tag_c2, pos_c2, vel_c2 = run_c(0)
    same = log == log2 and np.array_equal(tag_c, tag_c2) and np.array_equal(pos_c, pos_c2) and np.array_equal(vel_c, vel_c2)
    gc = gather({"log": log, "pos": pos_c, "tag": tag_c, "n0": n_c, "same": same, "decomp": ctx.decomposition()})
    ctx.close()

    if rank == 0:
      try:
        from oracle import port
        n_glob = 4 * nx ** 3
        # -- single-rank oracle, case A
        sim = port.md_example(nx, world_size=1, reneigh_every=1, particle_capacity=4 * n_glob + 4096, send_capacity=2 * n_glob + 4096)
        r = sim.ranks[0]
        assert r.nlocal == n_glob
        r.ints("uid", r.nlocal, view=True)[:] = np.arange(r.nlocal)
        ox0, ov0 = r.real("position"), r.real("linear_velocity")
        gx0 = np.concatenate([g["x0"] for g in ga])
        gv0 = np.concatenate([g["v0"] for g in ga])
        gtag0 = np.concatenate([g["tag0"] for g in ga])
        assert len(gx0) == n_glob and len(np.unique(gtag0)) == n_glob
        og, oo = _rows_sorted(gx0), _rows_sorted(ox0)
        assert np.array_equal(gx0[og], ox0[oo]), "the ranks' lattice parts are not the single-rank lattice"
        # adjust_thermo subtracts the GLOBAL mean velocity and rescales to the target temperature: over N ranks the global sums are
        # sums of per-rank partial sums (an MPI_Allreduce in the reference), so the factors differ from the single-rank ones in the
        # last bits
        dv0 = float(np.abs(gv0[og] - ov0[oo]).max() / np.abs(ov0).max())
        assert dv0 <= 1e-13, f"initial velocities differ from the single-rank set-up by {dv0}"
        gid_of_tag0 = np.empty(n_glob, np.int64)       # global id (= oracle uid) of the k-th gathered initial particle
        gid_of_tag0[og] = oo
        ox0_sorted = ox0[oo]
        tag2gid = lambda tg: gid_of_tag0[_by_tag(tg, gtag0)]      # noqa: E731  (the lattice set-up gives the same tags in every case)

        def oo_sorted_lookup(px):
            """global ids of particles given by their exact initial positions"""
            o = _rows_sorted(px)
            key = lambda q: np.ascontiguousarray(q).view([("", np.float64)] * 3).ravel()      # noqa: E731
            idx = np.searchsorted(key(ox0_sorted), key(px[o]))
            assert np.array_equal(ox0_sorted[idx], px[o])
            out = np.empty(len(px), np.int64)
            out[o] = oo[idx]
            return out
        worst_t = worst_p = 0.0
        for ts in range(steps_a):
            sim.step(ts)
            t, p = sim.thermo()
            worst_t = max(worst_t, abs(ga[0]["th"][ts, 1] - t) / t)
            worst_p = max(worst_p, abs(ga[0]["th"][ts, 2] - p) / abs(p))
        assert len(ga[0]["th"]) == steps_a and worst_t <= 1e-9 and worst_p <= 1e-9, (worst_t, worst_p)
        gx1 = np.concatenate([g["x1"] for g in ga])
        gtag1 = np.concatenate([g["tag1"] for g in ga])
        assert len(gx1) == n_glob and len(np.unique(gtag1)) == n_glob, "particles lost or duplicated by the migration"
        gid1 = gid_of_tag0[_by_tag(gtag1, gtag0)]
        ox1 = np.empty((n_glob, 3))
        ox1[r.ints("uid")] = r.real("position")
        d = gx1 - ox1[gid1]
        d -= L * np.round(d / L)          # a particle within rounding of a periodic face may be wrapped in one run only
        worst_x = float(np.abs(d).max())
        assert worst_x <= 1e-9, worst_x
        moved = int(sum(abs(g["counts"][0] - len(g["x0"])) for g in ga))
        report["setup"] = {"lattice_positions_bit_identical": True, "initial_velocity_rel": dv0}
        report["A"] = {"iterations": steps_a, "reneighbor_every": 1, "thermo_rel": worst_t, "pressure_rel": worst_p, "position_abs": worst_x,
                       "nlocal_per_rank": [g["counts"][0] for g in ga], "net_migration": moved}
        sim.close()

        # -- single-rank oracle, case B: its own modules on the state the ranks were dealt
        sim_b.exchange(); sim_b.borders()
        sim_b.build_cell_lists(); sim_b.partition_cell_lists(); sim_b.build_neighbor_lists()
        sim_b.reset_volatile(); sim_b.lennard_jones()
        assert sum(g["dealt"] for g in gb) == n_glob, ("dealt", [g["dealt"] for g in gb], n_glob)
        assert sum(g["counts"][0] for g in gb) == n_glob, ("owned after exchange", [g["counts"] for g in gb], n_glob)
        onn, onl = rb.neighbor_sets()
        ouid = rb.ints("uid", rb.nlocal + rb.nghost)
        omask = np.arange(onl.shape[1])[None, :] < onn[:, None]
        oi = np.broadcast_to(np.arange(rb.nlocal)[:, None], onl.shape)[omask]
        opairs = np.sort(ouid[oi].astype(np.int64) * n_glob + ouid[onl[omask]])
        gpairs = np.sort(np.concatenate([g["uid_i"].astype(np.int64) * n_glob + g["uid_j"] for g in gb]))
        assert np.array_equal(gpairs, opairs), f"global neighbour-pair set differs ({len(gpairs)} vs {len(opairs)} pairs)"
        of = np.empty((n_glob, 3))
        of[rb.ints("uid")] = rb.real("force")
        gf = np.concatenate([g["force"] for g in gb])
        guid = np.concatenate([g["uid"] for g in gb])
        assert np.array_equal(np.sort(guid), np.arange(n_glob))
        worst_f = float(np.abs(gf - of[guid]).max() / np.abs(of).max())
        assert np.abs(of).max() > 10.0 and worst_f <= 1e-12, worst_f
        report["B"] = {"state": "single-rank oracle after 25 iterations", "directed_pairs": int(len(opairs)), "pair_set_identical": True,
                       "force_rel": worst_f, "nghost_per_rank": [g["counts"][1] for g in gb]}
        sim_b.close()

        # -- N-rank restatement, case C
        assert all(g["same"] for g in gc), "overlap on / off are not bit-identical"
        sim = port.md_example(nx, world_size=world, reneigh_every=20, particle_capacity=200000, send_capacity=200000)
        assert tuple(gc[0]["decomp"]["nranks"]) == sim.nranks
        for k, rk in enumerate(sim.ranks):
            dcp = rk.decomposition()
            assert np.array_equal(dcp["neighbor_ranks"], gc[k]["decomp"]["neighbor_ranks"])
            assert np.array_equal(dcp["pbc"], gc[k]["decomp"]["pbc"]) and np.array_equal(dcp["subdom"], gc[k]["decomp"]["subdom"])
            assert rk.nlocal == gc[k]["n0"]
        o_gid_c = []
        for k, rk in enumerate(sim.ranks):
            rk.ints("uid", rk.nlocal, view=True)[:] = np.arange(rk.nlocal) + k * n_glob
            px = rk.real("position")
            o_gid_c.append(oo_sorted_lookup(px))
        worst, checked = 0.0, 0
        for ts in range(steps_c):
            sim.step(ts)
            t = sim.thermo()[0]
            for k, rk in enumerate(sim.ranks):
                lg = gc[k]["log"]
                if ("counts", ts) in lg:
                    assert lg[("counts", ts)] == (rk.nlocal, rk.nghost), (ts, k, lg[("counts", ts)], rk.nlocal, rk.nghost)
                if ts in lg:
                    worst = max(worst, abs(lg[ts] - t) / t)
                    checked += 1
        assert checked >= world * (steps_c // 7) and worst <= 1e-9, (checked, worst)
        # per-particle end state: identity through the exact initial lattice position (o_uid was set before the first step)
        worst_xc = 0.0
        for k, rk in enumerate(sim.ranks):
            gid = tag2gid(gc[k]["tag"])
            ox = np.full((n_glob, 3), np.nan)
            u = rk.ints("uid").astype(np.int64)                # a migrated particle carries the uid its ORIGIN rank gave it
            ogid = np.array([o_gid_c[q][j] for q, j in zip(u // n_glob, u % n_glob)], dtype=np.int64)
            ox[ogid] = rk.real("position")
            dd = gc[k]["pos"] - ox[gid]            # NaN where the two runs disagree on the owner of a particle
            assert np.all(np.isfinite(dd)), f"rank {k}: particle sets differ from the restatement's"
            worst_xc = max(worst_xc, float(np.abs(dd).max()))
        assert worst_xc <= 1e-9, worst_xc
        report["C"] = {"iterations": steps_c, "reneighbor_every": 20, "overlap_on_off_bit_identical": True, "thermo_rel_vs_nrank_restatement": worst,
                       "position_abs_vs_nrank_restatement": worst_xc, "rank_grid": list(sim.nranks)}
        sim.close()
        report.update({"ok": True, "world": world, "atoms": n_glob, "comparator": "single-rank oracle (A, B); N-rank restatement (C)"})
      except Exception as e:      # noqa: BLE001  -- the peers wait in the broadcast below: report, then everybody raises
        import traceback
        report.update({"ok": False, "error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-1500:]})
    out = [report]
    dist.broadcast_object_list(out, src=0)
    if not out[0].get("ok"):
        raise AssertionError("N-rank parity failed: " + str(out[0].get("error")) + "\n" + str(out[0].get("traceback", "")))
    return out[0]
