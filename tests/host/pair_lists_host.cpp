// CPU build of the PRODUCT's pair-list code (pairs_b200/csrc/pair_lists.h): the merge of two per-particle lists into the union list
// of a pair and the Lennard-Jones evaluation over those lists, one "thread" after the other.  Test code only.
#include <cstddef>
#include "pair_lists.h"

extern "C" {

void host_pairlist_merge_all(int nlocal, int T, int T2, double cutsq_lists, const double *pos4, const int *flags, const int *numneigh,
                             const int *neigh, int *pneigh, int *pnum) {
    const int npairs = (nlocal + 1) / 2;
    for(int p = 0; p < npairs; p++) {
        pnum[p] = pb_pairlist_merge(p, nlocal, T, T2, cutsq_lists, (const PbPos4 *) pos4, flags, numneigh, neigh, pneigh);
    }
}

// mode bits: 1 = uniform tables, 2 = accumulate onto the force array; fuse = 0..3
void host_lj_pairs_all(int nlocal, int T2, int cap, int ntypes, double cutsq, double dt, const double *eps_t, const double *sig6_t,
                       const double *pos4, const int *flags, const int *pnum, const int *pneigh, double *force, const double *mass, double *vel,
                       double *pos_next4, int mode, int fuse) {
    PbLjPairArgs a;
    a.nlocal = nlocal; a.T2 = T2; a.cap = cap; a.ntypes = ntypes; a.cutsq = cutsq; a.eps_u = eps_t[0]; a.sig6_u = sig6_t[0];
    a.dt = dt; a.half_dt = dt * 0.5; a.eps_t = eps_t; a.sig6_t = sig6_t; a.pos = (const PbPos4 *) pos4; a.flags = flags; a.pnum = pnum;
    a.pneigh = pneigh; a.force = force; a.mass = mass; a.vel = vel; a.pos_next = (PbPos4 *) pos_next4;
    const int npairs = (nlocal + 1) / 2;
    for(int p = 0; p < npairs; p++) {
#define RUN(U, A)                                                              \
        switch(fuse) {                                                         \
            case 1: pb_lj_pairs_thread<U, A, 1>(a, p); break;                  \
            case 2: pb_lj_pairs_thread<U, A, 2>(a, p); break;                  \
            case 3: pb_lj_pairs_thread<U, A, 3>(a, p); break;                  \
            default: pb_lj_pairs_thread<U, A, 0>(a, p); break;                 \
        }
        if(mode & 1) { if(mode & 2) { RUN(true, true) } else { RUN(true, false) } }
        else { if(mode & 2) { RUN(false, true) } else { RUN(false, false) } }
#undef RUN
    }
}

}
