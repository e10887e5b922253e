// comm.cu — NCCL inside the boundary: communicator set-up and the row-sharded lookup columns with
// both exchanges (bin-histogram all-reduce, partial-product all-gather) issued by the library.
//
// The reference has no communication layer ("distributed communication backend: none");
// this is SURVEY 8(b) item 8 / 8(e): shard g of G owns rows [g n/G, (g+1) n/G) of every column of
// ONE circuit, A' is a global sort (hence the histogram sum) and Z a global running product
// (hence the carry).  Both messages are small (6.4 MiB and G x 800 bytes at K=22, N=5) and
// latency-bound on NVLink 5 / NVSwitch: plain NCCL collectives on the compute stream, one call
// each for all lookups — there is no bulk transfer to fuse a kernel with.
//
// libnccl.so.2 is resolved with dlopen on first use: under PyTorch that is the NCCL torch already
// loaded (same SONAME), in a Rust/C++ host the system library.  No link-time dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "internal.h"

using namespace h2;

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};
NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!api.handle) {
            api.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        auto sym = [&](const char *s) -> void * {
            void *p = dlsym(api.handle, s);
            if (!p && api.why.empty()) api.why = std::string("libnccl lacks ") + s;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return api;
}
int32_t nccl_fail(h2aes_ctx *ctx, const NcclApi &api, ncclResult_t r, const char *what) {
    if (ctx) {
        ctx->last_error = std::string(what) + ": " +
                          (api.GetErrorString ? api.GetErrorString(r) : "NCCL error");
    }
    return H2AES_ERR_NCCL;
}
static_assert(H2AES_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "the id is an ncclUniqueId");
}  // namespace

#define NC(ctx, api, call)                                          \
    do {                                                            \
        ncclResult_t r__ = (call);                                  \
        if (r__ != ncclSuccess) return nccl_fail(ctx, api, r__, #call); \
    } while (0)

extern "C" int32_t h2aes_internal_set_beta_gamma(h2aes_circuit *c, const uint64_t *beta,
                                                 const uint64_t *gamma);

extern "C" {

int32_t h2aes_comm_unique_id(uint8_t *id) {
    if (!id) return H2AES_ERR_ARG;
    NcclApi &api = nccl_api();
    if (!api.why.empty()) return H2AES_ERR_NCCL;
    ncclUniqueId u;
    if (api.GetUniqueId(&u) != ncclSuccess) return H2AES_ERR_NCCL;
    memcpy(id, u.internal, H2AES_COMM_ID_BYTES);
    return H2AES_OK;
}

int32_t h2aes_comm_init(h2aes_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world) {
    if (!ctx || !id) return H2AES_ERR_ARG;
    if (world < 1 || world > 64 || (world & (world - 1)) || rank < 0 || rank >= world) {
        ctx->last_error = "world must be a power of two <= 64, 0 <= rank < world";
        return H2AES_ERR_ARG;
    }
    NcclApi &api = nccl_api();
    if (!api.why.empty()) {
        ctx->last_error = api.why;
        return H2AES_ERR_NCCL;
    }
    if (ctx->comm) {
        ctx->last_error = "the context already has a communicator";
        return H2AES_ERR_STATE;
    }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return H2AES_ERR_CUDA;
    ncclUniqueId u;
    memcpy(u.internal, id, H2AES_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    NC(ctx, api, api.CommInitRank(&comm, world, u, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return H2AES_OK;
}

int32_t h2aes_comm_destroy(h2aes_ctx *ctx) {
    if (!ctx) return H2AES_ERR_ARG;
    if (!ctx->comm) return H2AES_OK;
    NcclApi &api = nccl_api();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    NC(ctx, api, api.CommDestroy((ncclComm_t)ctx->comm));
    ctx->comm = nullptr;
    ctx->comm_world = 1;
    ctx->comm_rank = 0;
    return H2AES_OK;
}

int32_t h2aes_lookup_columns_sharded_dev(h2aes_circuit *c, const uint64_t *theta,
                                         const uint64_t *beta, const uint64_t *gamma,
                                         uint64_t *Ap_dev, uint64_t *Sp_dev, uint64_t *Z_dev) {
    if (!c) return H2AES_ERR_ARG;
    h2aes_ctx *ctx = c->ctx;
    auto bad = [&](int32_t code, const char *msg) {
        ctx->last_error = msg;
        return code;
    };
    if (!theta) return bad(H2AES_ERR_ARG, "null pointer");
    if ((Ap_dev == nullptr) != (Sp_dev == nullptr)) return bad(H2AES_ERR_ARG, "A' and S' go together");
    if ((beta == nullptr) != (gamma == nullptr) || (beta && !Z_dev) || (!beta && !Ap_dev))
        return bad(H2AES_ERR_ARG, "give A', S' buffers, or beta, gamma and a Z buffer, or all of them");
    if (c->n_shards > 1 && (!ctx->comm || ctx->comm_world != c->n_shards || ctx->comm_rank != c->shard))
        return bad(H2AES_ERR_STATE, "h2aes_comm_init(rank, world) must match h2aes_circuit_set_shard");
    NcclApi &api = nccl_api();
    const bool multi = c->n_shards > 1;
    const bool fresh = !c->have_bins || !c->have_theta || memcmp(c->theta, theta, 32) != 0;
    bool forked = false;
    int32_t rc;
    if (Ap_dev || fresh) {
        // exchange 1: the permuted columns are a GLOBAL sort -> sum of the bin histograms
        uint32_t *hist = nullptr;
        uint64_t words = 0;
        if ((rc = h2aes_lookup_hist_dev(c, theta, &hist, &words))) return rc;
        if (beta) {
            // beta, gamma are known up front here: their per-bin inverse tables only need the bins
            // (just enqueued), so they are built on a side stream while the histogram is reduced
            // and the permuted columns are expanded — serial work that does not shrink with G
            if (!ctx->side) {
                if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess)
                    return bad(H2AES_ERR_CUDA, "side stream");
            }
            cudaStream_t main_stream = ctx->stream;
            cudaEventRecord(ctx->ev_fork, main_stream);
            cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0);
            ctx->stream = ctx->side;
            rc = h2aes_internal_set_beta_gamma(c, beta, gamma);
            ctx->stream = main_stream;
            if (rc) return rc;
            cudaEventRecord(ctx->ev_join, ctx->side);
            forked = true;
        }
        if (multi)
            NC(ctx, api, api.AllReduce(hist, hist, words, ncclUint32, ncclSum, (ncclComm_t)ctx->comm,
                                       ctx->stream));
        if ((rc = h2aes_lookup_permuted_shard_dev(c, nullptr, nullptr, Ap_dev, Sp_dev))) return rc;
    }
    if (!beta) return H2AES_OK;
    // exchange 2: one partial product per lookup per shard -> carry of the shards before this one
    uint64_t *partial = nullptr;
    if (forked) cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    if ((rc = h2aes_lookup_partial_dev(c, beta, gamma, Z_dev, &partial))) return rc;
    uint64_t pwords = 0;  // one Fr per (lookup, segment of this shard)
    h2aes_circuit_shard_query(c, nullptr, &pwords);
    if (!c->d_gathered) {
        const size_t L = (size_t)c->sh.n_lookups();
        cudaError_t e = cudaMalloc((void **)&c->d_gathered, L * (zscan_tiles(c->sh.n()) + 64) * 32);
        if (e != cudaSuccess) return bad(H2AES_ERR_NOMEM, "partial-product gather buffer");
    }
    if (multi)
        NC(ctx, api, api.AllGather(partial, c->d_gathered, pwords, ncclUint64, (ncclComm_t)ctx->comm,
                                   ctx->stream));
    else if (cudaMemcpyAsync(c->d_gathered, partial, pwords * 8, cudaMemcpyDeviceToDevice, ctx->stream) !=
             cudaSuccess)
        return bad(H2AES_ERR_CUDA, "partial-product copy");
    return h2aes_lookup_product_shard_dev(c, c->d_gathered, Z_dev);
}

}  // extern "C"
