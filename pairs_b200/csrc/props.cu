// User-defined particle properties: everything Simulation.add_property() declares beyond the arrays the MD path keeps in
// dedicated storage (position, one velocity, one force, mass, type / flags / uid / shape).  The reference allocates one array per
// declared property (sim/properties.py:13-36) and lets every generated module take the arrays it touches as arguments; its
// communication treats them by volatility: Comm.exchange carries EVERY non-volatile property with a migrating particle
// (sim/comm.py:100-151, prop_list = properties.non_volatiles()), reset_volatile_properties zeroes every volatile one
// (sim/properties.py:61-70).  Here they are rows of ONE SoA block xdata[xrows][pcap] (a real = 1 row, a vector = 3 rows), so a
// warp's access to component d of a property is coalesced, and every structural operation of the pipeline is a loop over rows:
//   cell-order sort      gather through the permutation the counting sort leaves behind (binning.cu)
//   migration            non-volatile rows appended to the exchange record, hole filling moves them too (migrate.cu)
//   ghost creation       non-volatile rows appended to the border record (comm.cu) -- the reference leaves user properties of
//                        ghosts undefined (Comm.borders sends a fixed name list, sim/comm.py:56-72); here a ghost carries the
//                        values its source had at the last reneighbouring, like mass
//   capacity growth      rows re-strided, new slots = declared default
// Generated kernels (kernelgen.py) address a property as a.xdata[(row0 + d) * cap + i].
#include <algorithm>
#include <cstring>

#include "ctx.cuh"

PbXRows pb_xprops_nv_rows(const pb_ctx *ctx) {
    PbXRows r;
    r.n = 0;
    for(const auto &p : ctx->xprops) {
        if(p.is_volatile) { continue; }
        for(int d = 0; d < p.comps; d++) { r.row[r.n++] = p.row0 + d; }
    }
    return r;
}

__global__ void __launch_bounds__(256) pb_k_xfill(size_t first, size_t count, double *__restrict__ row, double v) {
    const size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(k < count) { row[first + k] = v; }
}

// slots [first, cap) of every row of every property <- its default
static int pb_xprops_fill(pb_ctx *ctx, double *data, size_t cap, size_t first) {
    if(cap <= first) { return 0; }
    for(const auto &p : ctx->xprops) {
        for(int d = 0; d < p.comps; d++) {
            PB_LAUNCH(pb_k_xfill, pb_blocks((long) (cap - first), 256), 256, first, cap - first, data + (size_t) (p.row0 + d) * cap, p.dflt[d]);
        }
    }
    return 0;
}

extern "C" int pb_add_property(pb_ctx *ctx, const char *name, int ncomps, int is_volatile, const double *defaults, int *prop_id) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(name == nullptr || ncomps < 1 || ncomps > PB_XPROP_MAX_COMPS) { ctx->set_error("pb_add_property: a property has 1 to 9 components"); return -1; }
    if(ctx->xrows + ncomps > PB_XPROP_MAX_ROWS) {
        ctx->set_error("pb_add_property: more than " + std::to_string(PB_XPROP_MAX_ROWS) + " rows of user-defined properties");
        return -1;
    }
    for(const auto &p : ctx->xprops) {
        if(p.name == name) { ctx->set_error(std::string("pb_add_property: '") + name + "' is already defined"); return -1; }
    }
    pb_ctx::XProp np;
    np.name = name;
    np.comps = ncomps;
    np.row0 = ctx->xrows;
    np.is_volatile = is_volatile != 0;
    for(int d = 0; d < ncomps; d++) { np.dflt[d] = (defaults != nullptr) ? defaults[d] : 0.0; }
    const size_t cap = (size_t) ctx->pcap;
    const int new_rows = ctx->xrows + ncomps;
    if(cap > 0) {
        // rows are contiguous blocks of `cap` doubles: the existing rows are a prefix of the new allocation
        PbScratch fresh, fresh_alt;          // released to the context only when every step has succeeded
        PB_CHECK(fresh.alloc(sizeof(double) * (size_t) new_rows * cap));
        PB_CHECK(fresh_alt.alloc(sizeof(double) * (size_t) new_rows * cap));
        double *q = fresh.as<double>();
        if(ctx->xdata != nullptr && ctx->xrows > 0) {
            PB_CHECK(cudaMemcpyAsync(q, ctx->xdata, sizeof(double) * (size_t) ctx->xrows * cap, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        for(int d = 0; d < ncomps; d++) {
            PB_LAUNCH(pb_k_xfill, pb_blocks((long) cap, 256), 256, (size_t) 0, cap, q + (size_t) (np.row0 + d) * cap, np.dflt[d]);
        }
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        if(ctx->xdata != nullptr) { PB_CHECK(cudaFree(ctx->xdata)); ctx->xdata = nullptr; }
        if(ctx->xdata_alt != nullptr) { PB_CHECK(cudaFree(ctx->xdata_alt)); ctx->xdata_alt = nullptr; }
        ctx->xdata = fresh.release<double>();
        ctx->xdata_alt = fresh_alt.release<double>();
    }
    ctx->xprops.push_back(np);
    ctx->xrows = new_rows;
    if(!np.is_volatile) { ctx->xrows_nv += ncomps; }
    if(ctx->send_cap > 0) {      // wire records got longer: the send buffer is sized per record
        const int keep = ctx->send_cap;
        ctx->send_cap = 0;
        PB_TRY(pb_ensure_send_capacity(ctx, keep));
    }
    if(ctx->recv_buf != nullptr) { PB_CHECK(cudaFree(ctx->recv_buf)); ctx->recv_buf = nullptr; ctx->recv_cap = 0; }
    if(prop_id != nullptr) { *prop_id = (int) ctx->xprops.size() - 1; }
    return 0;
}

extern "C" int pb_property_info(const pb_ctx *ctx, int prop_id, int *ncomps, int *row0, int *is_volatile) {
    if(prop_id < 0 || prop_id >= (int) ctx->xprops.size()) { return -1; }
    const auto &p = ctx->xprops[prop_id];
    if(ncomps != nullptr) { *ncomps = p.comps; }
    if(row0 != nullptr) { *row0 = p.row0; }
    if(is_volatile != nullptr) { *is_volatile = p.is_volatile ? 1 : 0; }
    return 0;
}

extern "C" int pb_property_count(const pb_ctx *ctx) { return (int) ctx->xprops.size(); }

// capacity change (pb_ensure_particle_capacity): rows keep their first `used` entries, the rest takes the default
int pb_xprops_grow(pb_ctx *ctx, size_t oldcap, size_t newcap, size_t used) {
    if(ctx->xrows == 0) { return 0; }
    PbScratch fresh, fresh_alt;              // released to the context only when every step has succeeded
    PB_CHECK(fresh.alloc(sizeof(double) * (size_t) ctx->xrows * newcap));
    PB_CHECK(fresh_alt.alloc(sizeof(double) * (size_t) ctx->xrows * newcap));
    double *q = fresh.as<double>();
    if(ctx->xdata != nullptr && used > 0) {
        PB_CHECK(cudaMemcpy2DAsync(q, newcap * sizeof(double), ctx->xdata, oldcap * sizeof(double), used * sizeof(double), (size_t) ctx->xrows,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    }
    PB_TRY(pb_xprops_fill(ctx, q, newcap, used));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    if(ctx->xdata != nullptr) { PB_CHECK(cudaFree(ctx->xdata)); ctx->xdata = nullptr; }
    if(ctx->xdata_alt != nullptr) { PB_CHECK(cudaFree(ctx->xdata_alt)); ctx->xdata_alt = nullptr; }
    ctx->xdata = fresh.release<double>();
    ctx->xdata_alt = fresh_alt.release<double>();
    return 0;
}

int pb_xprops_defaults(pb_ctx *ctx) {
    if(ctx->xrows == 0 || ctx->pcap == 0) { return 0; }
    return pb_xprops_fill(ctx, ctx->xdata, (size_t) ctx->pcap, 0);
}

__global__ void __launch_bounds__(256) pb_k_xgather(int n, size_t cap, const int *__restrict__ perm, const double *__restrict__ src,
                                                    double *__restrict__ dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) { return; }
    const size_t row = blockIdx.y;
    dst[row * cap + k] = src[row * cap + perm[k]];
}

int pb_xprops_permute(pb_ctx *ctx, const int *perm, int n) {
    if(ctx->xrows == 0 || n == 0) { return 0; }
    pb_k_xgather<<<dim3(pb_blocks(n, 256), ctx->xrows), 256, 0, ctx->stream>>>(n, (size_t) ctx->pcap, perm, ctx->xdata, ctx->xdata_alt);
    ctx->launches++;
    PB_CHECK(cudaGetLastError());
    std::swap(ctx->xdata, ctx->xdata_alt);
    return 0;
}

int pb_xprops_reset_volatile(pb_ctx *ctx) {
    if(ctx->nlocal == 0) { return 0; }
    for(const auto &p : ctx->xprops) {
        if(!p.is_volatile) { continue; }
        // the rows of one property are adjacent: one strided memset (0.0 is all-zero bits)
        PB_CHECK(cudaMemset2DAsync(ctx->xdata + (size_t) p.row0 * ctx->pcap, sizeof(double) * (size_t) ctx->pcap, 0,
                                   sizeof(double) * (size_t) ctx->nlocal, (size_t) p.comps, ctx->stream));
    }
    return 0;
}

// ---- wire records -----------------------------------------------------------------------------------------------------
// entry e of the send list (source particle send_map[e]) -> elements [offset, offset + rows.n) of record e
__global__ void __launch_bounds__(256) pb_k_xpack(int first, int count, size_t cap, int stride, int offset, PbXRows rows,
                                                  const int *__restrict__ send_map, const double *__restrict__ xdata, double *__restrict__ buf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const int e = first + k;
    const int p = send_map[e];
    double *b = buf + (size_t) e * stride + offset;
    for(int r = 0; r < rows.n; r++) { b[r] = xdata[(size_t) rows.row[r] * cap + p]; }
}

// migration: particle i leaves with wire record rec[i] (>= 0)
__global__ void __launch_bounds__(256) pb_k_xpack_leavers(int n, size_t cap, int stride, int offset, int buf_cap, PbXRows rows,
                                                          const int *__restrict__ rec, const double *__restrict__ xdata,
                                                          double *__restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const int e = rec[i];
    if(e < 0 || e >= buf_cap) { return; }
    double *b = buf + (size_t) e * stride + offset;
    for(int r = 0; r < rows.n; r++) { b[r] = xdata[(size_t) rows.row[r] * cap + i]; }
}

__global__ void __launch_bounds__(256) pb_k_xunpack(int first_rec, int count, int dst0, size_t cap, int stride, int offset, PbXRows rows,
                                                    const double *__restrict__ buf, double *__restrict__ xdata) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const double *b = buf + (size_t) (first_rec + k) * stride + offset;
    const int p = dst0 + k;
    for(int r = 0; r < rows.n; r++) { xdata[(size_t) rows.row[r] * cap + p] = b[r]; }
}

// hole filling of the migration: the k-th filler moves into the k-th hole, all rows (volatile ones included: cheap, and the
// particle keeps whatever a kernel accumulated since the last reset)
__global__ void __launch_bounds__(256) pb_k_xmove(const int *__restrict__ count, size_t cap, int nrows, const int *__restrict__ src_idx,
                                                  const int *__restrict__ dst_idx, double *__restrict__ xdata) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= *count) { return; }
    const int s = src_idx[k], t = dst_idx[k];
    for(int r = 0; r < nrows; r++) { xdata[(size_t) r * cap + t] = xdata[(size_t) r * cap + s]; }
}

int pb_xprops_pack(pb_ctx *ctx, int first, int count, int stride, int offset, const int *send_map, double *buf) {
    if(ctx->xrows_nv == 0 || count == 0) { return 0; }
    PB_LAUNCH(pb_k_xpack, pb_blocks(count, 256), 256, first, count, (size_t) ctx->pcap, stride, offset, pb_xprops_nv_rows(ctx), send_map,
              ctx->xdata, buf);
    return 0;
}

int pb_xprops_pack_leavers(pb_ctx *ctx, int n, int stride, int offset, const int *rec, double *buf) {
    if(ctx->xrows_nv == 0 || n == 0) { return 0; }
    PB_LAUNCH(pb_k_xpack_leavers, pb_blocks(n, 256), 256, n, (size_t) ctx->pcap, stride, offset, ctx->send_cap, pb_xprops_nv_rows(ctx), rec,
              ctx->xdata, buf);
    return 0;
}

int pb_xprops_unpack(pb_ctx *ctx, int first_rec, int count, int dst0, int stride, int offset, const double *buf) {
    if(ctx->xrows_nv == 0 || count == 0) { return 0; }
    PB_LAUNCH(pb_k_xunpack, pb_blocks(count, 256), 256, first_rec, count, dst0, (size_t) ctx->pcap, stride, offset, pb_xprops_nv_rows(ctx), buf,
              ctx->xdata);
    return 0;
}

int pb_xprops_move(pb_ctx *ctx, int max_count, const int *count, const int *src_idx, const int *dst_idx) {
    if(ctx->xrows == 0 || max_count == 0) { return 0; }
    PB_LAUNCH(pb_k_xmove, pb_blocks(max_count, 256), 256, count, (size_t) ctx->pcap, ctx->xrows, src_idx, dst_idx, ctx->xdata);
    return 0;
}

// ---- upload / download (host layout of the reference: [n][ncomps] doubles) -------------------------------------------
__global__ void __launch_bounds__(256) pb_k_x_aos_to_rows(int n, size_t cap, int comps, const double *__restrict__ aos, double *__restrict__ rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    for(int d = 0; d < comps; d++) { rows[(size_t) d * cap + i] = aos[(size_t) i * comps + d]; }
}

__global__ void __launch_bounds__(256) pb_k_x_rows_to_aos(int n, size_t cap, int comps, const double *__restrict__ rows, double *__restrict__ aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    for(int d = 0; d < comps; d++) { aos[(size_t) i * comps + d] = rows[(size_t) d * cap + i]; }
}

extern "C" int pb_upload_property(pb_ctx *ctx, int prop_id, int n, const double *values) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(prop_id < 0 || prop_id >= (int) ctx->xprops.size()) { ctx->set_error("pb_upload_property: unknown property id"); return -1; }
    if(n < 0 || n > ctx->nlocal) { ctx->set_error("pb_upload_property: n exceeds the number of local particles"); return -1; }
    if(n == 0) { return 0; }
    const auto &p = ctx->xprops[prop_id];
    PbScratch stage;
    PB_CHECK(stage.alloc(sizeof(double) * (size_t) n * p.comps));
    PB_CHECK(cudaMemcpyAsync(stage.p, values, sizeof(double) * (size_t) n * p.comps, cudaMemcpyHostToDevice, ctx->stream));
    PB_LAUNCH(pb_k_x_aos_to_rows, pb_blocks(n, 256), 256, n, (size_t) ctx->pcap, p.comps, stage.as<double>(),
              ctx->xdata + (size_t) p.row0 * ctx->pcap);
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int pb_download_property(pb_ctx *ctx, int prop_id, double *out, int with_ghosts) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(prop_id < 0 || prop_id >= (int) ctx->xprops.size()) { ctx->set_error("pb_download_property: unknown property id"); return -1; }
    const int n = ctx->nlocal + (with_ghosts ? ctx->nghost : 0);
    if(n == 0) { return 0; }
    const auto &p = ctx->xprops[prop_id];
    PbScratch stage;
    PB_CHECK(stage.alloc(sizeof(double) * (size_t) n * p.comps));
    PB_LAUNCH(pb_k_x_rows_to_aos, pb_blocks(n, 256), 256, n, (size_t) ctx->pcap, p.comps, ctx->xdata + (size_t) p.row0 * ctx->pcap,
              stage.as<double>());
    PB_CHECK(cudaMemcpyAsync(out, stage.p, sizeof(double) * (size_t) n * p.comps, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
