// Inter-rank transport for the comm stages: replaces Regular6DStencil::communicateSizes / communicateData /
// communicateAllData (runtime/domain/regular_6d_stencil.cpp:113-238, blocking MPI_Send/Recv/Sendrecv of host or
// CUDA-aware device buffers) by NCCL point-to-point over NVLink 5 / NVSwitch, one ncclGroup per phase on the
// context's stream (device buffers end to end, no host staging, no global synchronisation).
//
// NCCL is loaded with dlopen at pb_nccl_init time, so a single-GPU process needs no NCCL at all.  If the process
// already holds a libnccl.so.2 (e.g. PyTorch's bundled one, used by the launcher for rendezvous) the same copy is
// reused.
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>

#include "ctx.cuh"

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::string g_nccl_error;

static bool pb_nccl_load() {
    if(g_nccl.handle != nullptr) { return true; }
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for(const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if(h != nullptr) { break; }
    }
    if(h == nullptr) { g_nccl_error = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define PB_SYM(field, name)                                                     \
    *(void **) (&g_nccl.field) = dlsym(h, name);                                \
    if(g_nccl.field == nullptr) { g_nccl_error = std::string("missing NCCL symbol ") + name; return false; }
    PB_SYM(GetUniqueId, "ncclGetUniqueId");
    PB_SYM(CommInitRank, "ncclCommInitRank");
    PB_SYM(CommDestroy, "ncclCommDestroy");
    PB_SYM(Send, "ncclSend");
    PB_SYM(Recv, "ncclRecv");
    PB_SYM(GroupStart, "ncclGroupStart");
    PB_SYM(GroupEnd, "ncclGroupEnd");
    PB_SYM(AllReduce, "ncclAllReduce");
    PB_SYM(GetErrorString, "ncclGetErrorString");
#undef PB_SYM
    g_nccl.handle = h;
    return true;
}

// Message counts travel through a POSIX shared-memory board instead of the device: all ranks of a job live on one node
// (one process per GPU of an 8xB200 box), so "how many records will I get from you" is a host-to-host question.  Each rank
// owns one post per (dimension, parity): {sequence number, count to prev, count to next}; the receiver spins on the sequence
// number.  Compared with a 1-int ncclSend/ncclRecv group plus the stream synchronisation needed to read the result this
// removes one NCCL launch and one device round trip from every communication phase (12 per reneighbouring, 4 per DEM step).
// Parity double-buffering is enough: a rank posts message m+2 only after it has read both neighbours' message m+1, which they
// posted after reading its message m.
struct PbShmPost {
    std::atomic<unsigned long long> seq;
    int count[2];
    int pad[2];
};

struct NcclState {
    ncclComm_t comm = nullptr;
    double *d_red = nullptr;   // allreduce scratch
    int *d_counts = nullptr;   // size exchange scratch: [0..1] send, [2..3] recv
    PbShmPost *board = nullptr;                  // [world][3 dims][2 parities], null -> counts go through NCCL
    size_t board_bytes = 0;
    char board_name[64] = {0};
    unsigned long long msg[3] = {0, 0, 0};       // messages posted so far per dimension
};

static PbShmPost *pb_post(PbShmPost *board, int rank, int dim, unsigned long long m) { return board + ((size_t) rank * 3 + dim) * 2 + (m & 1ULL); }

// maps (creating it if needed) the board `name` for `world` ranks; nullptr if shared memory is not available
static PbShmPost *pb_board_map(const char *name, int world, size_t *bytes_out) {
    static_assert(sizeof(PbShmPost) == 24, "PbShmPost layout");
    const size_t bytes = sizeof(PbShmPost) * (size_t) world * 3 * 2;
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if(fd < 0) { return nullptr; }
    if(ftruncate(fd, (off_t) bytes) != 0) { close(fd); return nullptr; }        // new segments are zero-filled: seq = 0 everywhere
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if(p == MAP_FAILED) { return nullptr; }
    *bytes_out = bytes;
    return (PbShmPost *) p;
}

// message m of dimension `dim`: post my two counts, wait for the same message of both neighbours, read theirs.
// recv[0] = what `next` sends towards its prev (= me), recv[1] = what `prev` sends towards its next (= me).  false on timeout.
static bool pb_board_exchange(PbShmPost *board, int rank, int prev, int next, int dim, unsigned long long m, const int send[2], int recv[2],
                              int timeout_s) {
    PbShmPost *mine = pb_post(board, rank, dim, m);
    mine->count[0] = send[0];
    mine->count[1] = send[1];
    mine->seq.store(m, std::memory_order_release);
    PbShmPost *from_next = pb_post(board, next, dim, m), *from_prev = pb_post(board, prev, dim, m);
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while(from_next->seq.load(std::memory_order_acquire) != m || from_prev->seq.load(std::memory_order_acquire) != m) {
        if((++spins & 0xfffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(timeout_s)) { return false; }
    }
    recv[0] = from_next->count[0];
    recv[1] = from_prev->count[1];
    return true;
}

static void pb_board_open(pb_ctx *ctx, NcclState *st, const void *id128) {
    if(getenv("PB_NO_SHM_BOARD") != nullptr) { return; }
    unsigned long long h = 1469598103934665603ULL;           // FNV-1a of the job's NCCL id: the same on every rank, unique per job
    for(int k = 0; k < 128; k++) { h = (h ^ ((const unsigned char *) id128)[k]) * 1099511628211ULL; }
    snprintf(st->board_name, sizeof(st->board_name), "/pairs_b200_%016llx", h);
    st->board = pb_board_map(st->board_name, ctx->world, &st->board_bytes);
}

// Host-only self-test of the board protocol (no GPU, no NCCL): `rounds` messages per dimension on a periodic ring of `world`
// processes, every count checked against what the neighbour must have posted.  Run by tests/test_cabi_and_host.py with several
// processes.  Returns 0, or the (1-based) round in which a wrong value / a timeout was seen (negative for timeouts).
extern "C" int pb_board_selftest(const char *name, int world, int rank, int rounds) {
    size_t bytes = 0;
    PbShmPost *board = pb_board_map(name, world, &bytes);
    if(board == nullptr) { return -1000000; }
    const int prev = (rank + world - 1) % world, next = (rank + 1) % world;
    auto value = [](int r, int round, int dim, int side) { return ((r * 1000 + round) * 3 + dim) * 2 + side; };
    int rc = 0;
    for(int round = 1; round <= rounds && rc == 0; round++) {
        for(int dim = 0; dim < 3 && rc == 0; dim++) {
            const int send[2] = {value(rank, round, dim, 0), value(rank, round, dim, 1)};
            int recv[2] = {-1, -1};
            if(!pb_board_exchange(board, rank, prev, next, dim, (unsigned long long) round, send, recv, 20)) { rc = -round; break; }
            if(recv[0] != value(next, round, dim, 0) || recv[1] != value(prev, round, dim, 1)) { rc = round; }
        }
    }
    munmap(board, bytes);
    return rc;
}

extern "C" int pb_board_unlink(const char *name) { return shm_unlink(name); }

#define PB_NCCL(call)                                                                                       \
    do {                                                                                                    \
        ncclResult_t r_ = (call);                                                                           \
        if(r_ != ncclSuccess) {                                                                             \
            ctx->set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r_));                          \
            return -1;                                                                                      \
        }                                                                                                   \
    } while(0)

extern "C" int pb_nccl_unique_id(void *id128) {
    if(!pb_nccl_load()) { return -1; }
    ncclUniqueId id;
    if(g_nccl.GetUniqueId(&id) != ncclSuccess) { return -1; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id128, &id, 128);
    return 0;
}

int pb_allreduce_sum(pb_ctx *ctx, double *vals, int n);

extern "C" int pb_nccl_init(pb_ctx *ctx, const void *id128) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!pb_nccl_load()) { ctx->set_error(g_nccl_error); return -1; }
    if(!ctx->domain_set) { ctx->set_error("pb_nccl_init: call pb_init_domain first"); return -1; }
    NcclState *st = new NcclState();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&st->comm, ctx->world, id, ctx->rank);
    if(r != ncclSuccess) {
        ctx->set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
        delete st;
        return -1;
    }
    PB_CHECK(cudaMalloc(&st->d_red, sizeof(double) * 4));
    PB_CHECK(cudaMalloc(&st->d_counts, sizeof(int) * 8));
    pb_board_open(ctx, st, id128);
    // every rank must agree on the way counts travel: sum of "I have a board" over the ranks -- and the board is a POSIX
    // shared-memory segment, i.e. one per NODE: ranks on different hosts would each map their own and wait for posts that never
    // arrive.  All ranks on one host <=> world * sum(h^2) == sum(h)^2 for a 20-bit hash h of the host name (exact in doubles).
    char host[256] = {0};
    gethostname(host, sizeof(host) - 1);
    unsigned hh = 2166136261u;
    for(const char *c = host; *c != 0; c++) { hh = (hh ^ (unsigned char) *c) * 16777619u; }
    const double hv = (double) (hh & 0xfffffu);
    double agree[3] = {st->board != nullptr ? 1.0 : 0.0, hv, hv * hv};
    ctx->nccl = st;
    PB_TRY(pb_allreduce_sum(ctx, agree, 3));
    const bool one_host = (double) ctx->world * agree[2] == agree[1] * agree[1];
    const double have = one_host ? agree[0] : 0.0;
    if(have < ctx->world - 0.5 && st->board != nullptr) {
        munmap(st->board, st->board_bytes);
        st->board = nullptr;
    }
    return 0;
}

void pb_nccl_destroy(pb_ctx *ctx) {
    NcclState *st = (NcclState *) ctx->nccl;
    if(st == nullptr) { return; }
    if(st->comm != nullptr) { g_nccl.CommDestroy(st->comm); }
    if(st->board != nullptr) {
        munmap(st->board, st->board_bytes);
        shm_unlink(st->board_name);          // every rank has mapped it by now (agreement allreduce in pb_nccl_init); repeated unlinks just fail
    }
    cudaFree(st->d_red);
    cudaFree(st->d_counts);
    delete st;
    ctx->nccl = nullptr;
}

static int pb_require_comm(pb_ctx *ctx, NcclState **st) {
    *st = (NcclState *) ctx->nccl;
    if(*st == nullptr) {
        ctx->set_error("multi-rank domain but no communicator attached (pb_nccl_init)");
        return -1;
    }
    return 0;
}

// communicateSizes: what I send to prev arrives in prev's "from next" slot (2d+0), and vice versa.
int pb_transport_sizes(pb_ctx *ctx, int dim) {
    const int prev = ctx->neighbor_ranks[dim * 2], next = ctx->neighbor_ranks[dim * 2 + 1];
    if(prev == ctx->rank && next == ctx->rank) {
        ctx->nrecv[dim * 2] = ctx->nsend[dim * 2];
        ctx->nrecv[dim * 2 + 1] = ctx->nsend[dim * 2 + 1];
        return 0;
    }
    NcclState *st;
    PB_TRY(pb_require_comm(ctx, &st));
    if(st->board != nullptr) {
        const int send[2] = {ctx->nsend[dim * 2], ctx->nsend[dim * 2 + 1]};
        int recv[2] = {0, 0};
        if(!pb_board_exchange(st->board, ctx->rank, prev, next, dim, ++st->msg[dim], send, recv, 120)) {
            ctx->set_error("size exchange timed out: a neighbouring rank did not reach the same communication phase");
            return -1;
        }
        ctx->nrecv[dim * 2] = recv[0];
        ctx->nrecv[dim * 2 + 1] = recv[1];
        return 0;
    }
    ctx->h_scalars[4] = ctx->nsend[dim * 2];
    ctx->h_scalars[5] = ctx->nsend[dim * 2 + 1];
    PB_CHECK(cudaMemcpyAsync(st->d_counts, ctx->h_scalars + 4, sizeof(int) * 2, cudaMemcpyHostToDevice, ctx->stream));
    PB_NCCL(g_nccl.GroupStart());
    PB_NCCL(g_nccl.Send(st->d_counts + 0, 1, ncclInt32, prev, st->comm, ctx->stream));
    PB_NCCL(g_nccl.Recv(st->d_counts + 2, 1, ncclInt32, next, st->comm, ctx->stream));
    PB_NCCL(g_nccl.Send(st->d_counts + 1, 1, ncclInt32, next, st->comm, ctx->stream));
    PB_NCCL(g_nccl.Recv(st->d_counts + 3, 1, ncclInt32, prev, st->comm, ctx->stream));
    PB_NCCL(g_nccl.GroupEnd());
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 6, st->d_counts + 2, sizeof(int) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->nrecv[dim * 2] = ctx->h_scalars[6];
    ctx->nrecv[dim * 2 + 1] = ctx->h_scalars[7];
    return 0;
}

// communicateData (one dim) / communicateAllData (all dims): moves the packed records.  *recv_src is the buffer the
// unpack kernel must read (indexed with recv_offsets): on a single rank that is the send buffer itself.
int pb_transport_data(pb_ctx *ctx, int dim_begin, int dim_end, int elem, const double **recv_src) {
    if(ctx->world == 1) {
        *recv_src = ctx->send_buf;
        return 0;
    }
    NcclState *st;
    PB_TRY(pb_require_comm(ctx, &st));
    int total_recv = 0;
    for(int j = 0; j < 6; j++) { total_recv = std::max(total_recv, ctx->recv_offsets[j] + ctx->nrecv[j]); }
    if(total_recv > ctx->recv_cap) {
        if(ctx->recv_buf != nullptr) { PB_CHECK(cudaFree(ctx->recv_buf)); }
        ctx->recv_cap = std::max(total_recv + total_recv / 4, 1024);
        PB_CHECK(cudaMalloc(&ctx->recv_buf, sizeof(double) * (size_t) ctx->recv_cap * pb_record_elems(ctx)));
    }
    bool grouped = false;
    for(int d = dim_begin; d < dim_end; d++) {
        const int prev = ctx->neighbor_ranks[d * 2], next = ctx->neighbor_ranks[d * 2 + 1];
        const double *send_prev = ctx->send_buf + (size_t) ctx->send_offsets[d * 2] * elem;
        const double *send_next = ctx->send_buf + (size_t) ctx->send_offsets[d * 2 + 1] * elem;
        double *recv_prev = ctx->recv_buf + (size_t) ctx->recv_offsets[d * 2] * elem;
        double *recv_next = ctx->recv_buf + (size_t) ctx->recv_offsets[d * 2 + 1] * elem;
        if(prev == ctx->rank && next == ctx->rank) {
            // copy_in_device branch of the reference
            if(ctx->nsend[d * 2] > 0) {
                PB_CHECK(cudaMemcpyAsync(recv_prev, send_prev, sizeof(double) * (size_t) ctx->nsend[d * 2] * elem,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
            }
            if(ctx->nsend[d * 2 + 1] > 0) {
                PB_CHECK(cudaMemcpyAsync(recv_next, send_next, sizeof(double) * (size_t) ctx->nsend[d * 2 + 1] * elem,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
            }
            continue;
        }
        if(!grouped) { PB_NCCL(g_nccl.GroupStart()); grouped = true; }
        // to prev / from next fills slot 2d+0; to next / from prev fills slot 2d+1 (MPI_Sendrecv pairs of the reference)
        PB_NCCL(g_nccl.Send(send_prev, (size_t) ctx->nsend[d * 2] * elem, ncclFloat64, prev, st->comm, ctx->stream));
        PB_NCCL(g_nccl.Recv(recv_prev, (size_t) ctx->nrecv[d * 2] * elem, ncclFloat64, next, st->comm, ctx->stream));
        PB_NCCL(g_nccl.Send(send_next, (size_t) ctx->nsend[d * 2 + 1] * elem, ncclFloat64, next, st->comm, ctx->stream));
        PB_NCCL(g_nccl.Recv(recv_next, (size_t) ctx->nrecv[d * 2 + 1] * elem, ncclFloat64, prev, st->comm, ctx->stream));
    }
    if(grouped) { PB_NCCL(g_nccl.GroupEnd()); }
    *recv_src = ctx->recv_buf;
    return 0;
}

// MPI_Allreduce(SUM) of runtime/thermo.hpp:18,40
int pb_allreduce_thermo(pb_ctx *ctx, double *sum_mv2, long *natoms) {
    NcclState *st;
    PB_TRY(pb_require_comm(ctx, &st));
    double h[2] = {*sum_mv2, (double) *natoms};
    PB_CHECK(cudaMemcpyAsync(st->d_red, h, sizeof(double) * 2, cudaMemcpyHostToDevice, ctx->stream));
    PB_NCCL(g_nccl.AllReduce(st->d_red, st->d_red + 2, 2, ncclFloat64, ncclSum, st->comm, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(h, st->d_red + 2, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    *sum_mv2 = h[0];
    *natoms = (long) (h[1] + 0.5);
    return 0;
}

// generic small sum-allreduce of host doubles (set-up paths: adjust_thermo)
int pb_allreduce_sum(pb_ctx *ctx, double *vals, int n) {
    NcclState *st;
    PB_TRY(pb_require_comm(ctx, &st));
    if(n > 2) {
        // scratch holds 4 doubles: reduce in chunks of 2
        for(int k = 0; k < n; k += 2) {
            const int c = std::min(2, n - k);
            PB_CHECK(cudaMemcpyAsync(st->d_red, vals + k, sizeof(double) * c, cudaMemcpyHostToDevice, ctx->stream));
            PB_NCCL(g_nccl.AllReduce(st->d_red, st->d_red + 2, c, ncclFloat64, ncclSum, st->comm, ctx->stream));
            PB_CHECK(cudaMemcpyAsync(vals + k, st->d_red + 2, sizeof(double) * c, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CHECK(cudaStreamSynchronize(ctx->stream));
        }
        return 0;
    }
    PB_CHECK(cudaMemcpyAsync(st->d_red, vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    PB_NCCL(g_nccl.AllReduce(st->d_red, st->d_red + 2, n, ncclFloat64, ncclSum, st->comm, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(vals, st->d_red + 2, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
