// apc_comm.cu — the multi-GPU entry points of libapc (include/apc.h): one NCCL rank per context, the
// per-k-mer count vectors of the read shards summed with one small all-reduce on the context's stream.
//
// Replaces the reference's OpenMP team over k-mers and its `omp critical` merge of the results
// (/root/reference/approx_counter.cpp:547-599, :595-596): reads are independent and a count is a sum over
// reads (:589-596), so sharding the reads needs no other exchange.
//
// libnccl.so.2 is opened with dlopen on first use, so libapc.so has no link-time dependency on it: a
// single-GPU host needs no NCCL at all, and inside a process that already loaded an NCCL (torch ships its
// own) the same copy is used.  Only the five entry points below are bound; their signatures and the two
// enum values are those of nccl.h 2.x (ncclUint64 = 5, ncclSum = 0, 128-byte unique id passed by value).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "apc_internal.h"

namespace {

struct NcclId {
    char internal[APC_COMM_ID_BYTES];
};
typedef void *NcclComm;
constexpr int kNcclUint64 = 5, kNcclSum = 0;

struct NcclApi {
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error; // why loading failed ("" = loaded)
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = nullptr;
        // APC_NCCL_LIB: explicit path (tests, unusual installs)
        const char *names[] = {getenv("APC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
            api.error = "libnccl.so.2 lacks an expected entry point";
            return;
        }
        api.ok = true;
    });
    return api;
}

int comm_fail(apc::Ctx *c, const char *what, int rc) {
    std::string msg = what;
    NcclApi &n = nccl();
    if (!n.ok) msg += ": " + n.error;
    else if (rc != 0) msg += std::string(": ") + n.GetErrorString(rc);
    return apc::fail(c, APC_ERR_COMM, msg.c_str());
}

} // namespace

extern "C" {

int apc_comm_unique_id(uint8_t id_out[APC_COMM_ID_BYTES]) {
    if (!id_out) return APC_ERR_INVALID;
    NcclApi &n = nccl();
    if (!n.ok) return APC_ERR_COMM;
    NcclId id;
    if (n.GetUniqueId(&id) != 0) return APC_ERR_COMM;
    std::memcpy(id_out, id.internal, APC_COMM_ID_BYTES);
    return APC_OK;
}

int apc_comm_init_rank(apc_ctx *c, int n_ranks, int rank, const uint8_t id[APC_COMM_ID_BYTES]) {
    if (!c) return APC_ERR_INVALID;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id) return apc::fail(c, APC_ERR_INVALID, "apc_comm_init_rank: bad rank / size / id");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return apc::fail(c, APC_ERR_CUDA, "cudaSetDevice", e);
    NcclApi &n = nccl();
    if (!n.ok) return comm_fail(c, "apc_comm_init_rank", 0);
    apc_comm_destroy(c);
    NcclId nid;
    std::memcpy(nid.internal, id, APC_COMM_ID_BYTES);
    NcclComm comm = nullptr;
    const int rc = n.CommInitRank(&comm, n_ranks, nid, rank);
    if (rc != 0 || !comm) return comm_fail(c, "ncclCommInitRank", rc);
    c->nccl_comm = comm;
    c->comm_rank = rank;
    c->comm_size = n_ranks;
    return APC_OK;
}

int apc_comm_destroy(apc_ctx *c) {
    if (!c) return APC_ERR_INVALID;
    if (c->nccl_comm) {
        cudaSetDevice(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        nccl().CommDestroy((NcclComm)c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_size = 1;
    return APC_OK;
}

int apc_comm_info(const apc_ctx *c, int *rank, int *n_ranks) {
    if (!c) return APC_ERR_INVALID;
    if (rank) *rank = c->comm_rank;
    if (n_ranks) *n_ranks = c->comm_size;
    return APC_OK;
}

int apc_allreduce_counts(apc_ctx *c, uint64_t *d_counts, uint64_t n) {
    if (!c) return APC_ERR_INVALID;
    if (!c->nccl_comm) return APC_OK; // no communicator: one GPU, the sum is the vector itself
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return apc::fail(c, APC_ERR_CUDA, "cudaSetDevice", e);
    void *buf = d_counts;
    if (!buf) {
        buf = c->d_counts;
        n = c->n_kmers;
    }
    if (n == 0) return APC_OK;
    if (!buf) return apc::fail(c, APC_ERR_NO_QUERIES, "apc_allreduce_counts: no count buffer");
    const int rc = nccl().AllReduce(buf, buf, (size_t)n, kNcclUint64, kNcclSum, (NcclComm)c->nccl_comm, c->stream);
    if (rc != 0) return comm_fail(c, "ncclAllReduce", rc);
    return APC_OK;
}

} // extern "C"
