// The generated timestep loop of examples/md.py, run natively: sim/timestep.py:9-72 (loop of nsteps+1 iterations,
// guards `every` -> ((ts+1) % n == 0) || ts == 0 and `skip_first` -> ts > 0) around the per-step procedure list of
// sim/simulation.py:387-417:  pre_step kernels (initial_integrate) -> exchange + borders | synchronize ->
// build_cell_lists -> partition_cell_lists -> build_neighbor_lists -> reset_volatile_properties -> lennard_jones ->
// final_integrate -> compute_thermo.  Everything stays on the device; only thermo scalars (every `thermo_every`
// steps) and capacity counters (every reneighbouring) are read back.
#include <algorithm>

#include "ctx.cuh"

int pb_lennard_jones_fused(pb_ctx *ctx, double cutoff, double dt, int fuse, int part);
int pb_lj_finish_split(pb_ctx *ctx, int fuse);
int pb_upload_join(pb_ctx *ctx);      // ctx.cu: the deferred half of pb_md_run_from_host's upload
int pb_borders_refill(pb_ctx *ctx);   // comm.cu: ... and the ghosts' copies of those arrays

extern "C" int pb_md_run(pb_ctx *ctx, const pb_md_params *p, int ts_begin, int ts_end, double *thermo_out, int thermo_cap, int *n_thermo) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->cells_set || ctx->spacing != p->cell_spacing) { PB_TRY(pb_setup_cells(ctx, p->cell_spacing)); }
    int nt = 0;
    bool initial_done = false;     // initial_integrate of this iteration was already applied by the previous force kernel
    // The mirror of the tile lists (positions in CSR order, tile_lists.cu) is trusted only inside this loop, where every writer of
    // positions is known: the fused force kernel and the ghost refresh keep it current, everything else marks it stale.
    struct MirrorScope {
        pb_ctx *c;
        explicit MirrorScope(pb_ctx *c_) : c(c_) { c->mirror_scope = true; c->mirror_fresh = false; }
        ~MirrorScope() { c->mirror_scope = false; c->mirror_fresh = false; }
    } mirror_scope(ctx);
    for(int ts = ts_begin; ts < ts_end; ts++) {
        const bool reneigh = (((ts + 1) % p->reneighbor_every) == 0) || (ts == 0);
        if(!reneigh || ts > 0) { PB_TRY(pb_upload_join(ctx)); }      // (only the list build of iteration 0 runs without velocities)
        if(ts > 0 && !initial_done) {
            PB_TRY(pb_initial_integrate(ctx, p->dt));
            ctx->mirror_fresh = false;
        }
        initial_done = false;
        if(reneigh) {
            ctx->mirror_fresh = false;      // (the tile build writes the mirror anew)
            PB_TRY(pb_exchange(ctx));
            PB_TRY(pb_borders(ctx));
            PB_TRY(pb_build_cell_lists(ctx));
            PB_TRY(pb_build_neighbor_lists(ctx, p->cutoff_lists));
            if(ctx->upload_pending) {          // pb_md_run_from_host: velocities and masses have arrived meanwhile
                PB_TRY(pb_upload_join(ctx));
                PB_TRY(pb_borders_refill(ctx));
            }
        }
        // Multi-rank steps without reneighbouring: the ghost refresh (pack -> NCCL -> unpack) runs on comm_stream while the
        // interior warp groups -- no ghost neighbour, no halo source -- already compute on the main stream; the boundary
        // groups follow once the refresh has landed.  (The reference's communication is blocking)
        // (half lists: a particle's force is complete only after the whole grid, so nothing is fused or split)
        const bool fusing = ctx->fuse_integrate && !ctx->half_lists;
        // (tile lists: only with a current mirror -- otherwise this step rebuilds it, in order, on one stream)
        const bool overlap = !reneigh && ctx->world > 1 && ctx->overlap_comm && fusing &&
                             ((ctx->tiles_n == ctx->nlocal && ctx->tile_split_valid && ctx->mirror_fresh) ||
                              (ctx->tiles_n != ctx->nlocal && ctx->groups_valid && ctx->neigh_n == ctx->nlocal));
        if(!reneigh && !overlap) { PB_TRY(pb_synchronize(ctx)); }
        PB_TRY(pb_reset_volatile(ctx));
        const bool thermo_now = p->thermo_every > 0 && ((((ts + 1) % p->thermo_every) == 0) || ts == 0);
        if(fusing) {
            // fold final_integrate(ts) and -- unless thermo must see the velocities in between, or the call ends here --
            // initial_integrate(ts + 1) into the force kernel
            int fuse = (ts > 0) ? 1 : 0;
            if(!thermo_now && ts + 1 < ts_end) { fuse |= 2; initial_done = true; }
            if(overlap) {
                // comm stream (high priority, issued first so its small kernels are scheduled ahead of the big one):
                // pack -> NCCL -> unpack -> boundary groups.  main stream: interior groups.  The two force launches touch
                // disjoint particles; the streams join before the next iteration.
                PB_CHECK(cudaEventRecord(ctx->ev_prev, ctx->stream));
                PB_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_prev, 0));
                std::swap(ctx->stream, ctx->comm_stream);
                int rc = pb_synchronize(ctx);
                if(rc >= 0) { rc = pb_lennard_jones_fused(ctx, p->cutoff_force, p->dt, fuse, 2); }
                if(rc >= 0) { cudaEventRecord(ctx->ev_sync, ctx->stream); }
                std::swap(ctx->stream, ctx->comm_stream);
                PB_TRY(rc);
                PB_TRY(pb_lennard_jones_fused(ctx, p->cutoff_force, p->dt, fuse, 1));
                PB_TRY(pb_lj_finish_split(ctx, fuse));
                PB_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_sync, 0));
            } else {
                PB_TRY(pb_lennard_jones_fused(ctx, p->cutoff_force, p->dt, fuse, 0));
            }
        } else {
            PB_TRY(pb_lennard_jones(ctx, p->cutoff_force));
            if(ts > 0) { PB_TRY(pb_final_integrate(ctx, p->dt)); }
        }
        if(thermo_now) {
            double t = 0.0, pr = 0.0;
            PB_TRY(pb_compute_thermo(ctx, &t, &pr));
            if(thermo_out != nullptr && nt < thermo_cap) {
                thermo_out[nt * 3 + 0] = (double) ts;
                thermo_out[nt * 3 + 1] = t;
                thermo_out[nt * 3 + 2] = pr;
            }
            nt++;
        }
    }
    if(n_thermo != nullptr) { *n_thermo = nt; }
    return 0;
}
