/* Single-rank MPI stand-in for building the reference's serial CPU target as a
 * parity oracle (there is no MPI in this image).  TEST INFRASTRUCTURE ONLY.
 *
 * Semantics: world size 1, rank 0, every Cartesian neighbour is the rank itself.
 * The reference runtime (runtime/domain/regular_6d_stencil.cpp:113-238) never
 * reaches Send/Recv/Sendrecv when prev == next == rank, so those are no-ops.
 * MPI_Allreduce over one rank is a copy; datatype handles encode the element size. */
#ifndef PAIRS_B200_MPI_STUB_H
#define PAIRS_B200_MPI_STUB_H
#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int unused; } MPI_Status;
typedef int MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_SUM 0
#define MPI_MIN 1
#define MPI_MAX 2
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL 0
#define MPI_SUCCESS 0

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Initialized(int *flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Cart_create(MPI_Comm c, int nd, const int *dims, const int *periods, int reorder, MPI_Comm *out) {
    (void)c; (void)nd; (void)dims; (void)periods; (void)reorder; *out = 1; return MPI_SUCCESS;
}
static inline int MPI_Cart_get(MPI_Comm c, int nd, int *dims, int *periods, int *coords) {
    (void)c; (void)dims; (void)periods;
    for(int d = 0; d < nd; d++) { coords[d] = 0; }
    return MPI_SUCCESS;
}
static inline int MPI_Cart_shift(MPI_Comm c, int dir, int disp, int *src, int *dst) {
    (void)c; (void)dir; (void)disp; *src = 0; *dst = 0; return MPI_SUCCESS;
}
static inline int MPI_Comm_free(MPI_Comm *c) { *c = 0; return MPI_SUCCESS; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return MPI_SUCCESS;
}
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s) {
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return MPI_SUCCESS;
}
static inline int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int dst, int stag,
                               void *rb, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status *s) {
    (void)dst; (void)stag; (void)rn; (void)rt; (void)src; (void)rtag; (void)c; (void)s;
    memcpy(rb, sb, (size_t)sn * (size_t)st);
    return MPI_SUCCESS;
}
static inline int MPI_Allreduce(const void *sb, void *rb, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    (void)op; (void)c;
    memcpy(rb, sb, (size_t)n * (size_t)t);
    return MPI_SUCCESS;
}
#endif
