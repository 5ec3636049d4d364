// comm.cu — the one exchange step of the sharded direct Fock build behind the C ABI: FP64 (or, for deterministic
// builds, 64-bit integer) sum all-reduce of the partial G matrices over NCCL / NVLink.  A non-Python binding of
// include/mmdb200.h gets shards from mmdb_fock_direct(shard, nshards) and the reduction from mmdb_allreduce_G; the
// NCCL unique id travels through whatever rendezvous the host program has (file, MPI, torch.distributed, ...).
//
// NCCL is resolved at run time (dlopen of the libnccl the process already carries, e.g. torch's bundled copy), so
// libmmdb200.so loads — and its symbols can be checked — on machines without NCCL or without a GPU.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "handle.h"

namespace {
typedef struct { char internal[128]; } nccl_uid;           // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void *nccl_comm;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(nccl_comm *, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm);
typedef const char *(*fn_errstr)(int);
constexpr int NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0;      // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct NcclApi {
    void *so = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
    std::string why;
};

NcclApi &nccl()
{
    static NcclApi api;
    if (api.so || !api.why.empty()) return api;
    const char *names[] = {getenv("MMDB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.so) break;
    }
    if (!api.so) {
        api.why = "libnccl.so.2 not found (set MMDB_NCCL_LIB or import torch first, which carries one)";
        return api;
    }
    api.get_uid = (fn_get_uid)dlsym(api.so, "ncclGetUniqueId");
    api.init_rank = (fn_init_rank)dlsym(api.so, "ncclCommInitRank");
    api.allreduce = (fn_allreduce)dlsym(api.so, "ncclAllReduce");
    api.destroy = (fn_destroy)dlsym(api.so, "ncclCommDestroy");
    api.errstr = (fn_errstr)dlsym(api.so, "ncclGetErrorString");
    if (!api.get_uid || !api.init_rank || !api.allreduce || !api.destroy) {
        api.why = "libnccl lacks ncclGetUniqueId/ncclCommInitRank/ncclAllReduce/ncclCommDestroy";
        api.so = nullptr;
    }
    return api;
}

int nccl_fail(const char *what, int rc)
{
    NcclApi &a = nccl();
    return fail(MMDB_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(rc) + (a.errstr ? std::string(" (") + a.errstr(rc) + ")" : ""));
}
}  // namespace

struct mmdb_comm {
    nccl_comm comm = nullptr;
    int device = 0, rank = 0, nranks = 1;
};

extern "C" int mmdb_comm_unique_id(unsigned char *id128)
{
    NcclApi &a = nccl();
    if (!a.so) return fail(MMDB_ERR_UNSUPPORTED, "mmdb_comm_unique_id: " + a.why);
    nccl_uid u;
    const int rc = a.get_uid(&u);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    std::memcpy(id128, u.internal, 128);
    return MMDB_OK;
}

extern "C" int mmdb_comm_init(int device, int nranks, int rank, const unsigned char *id128, mmdb_comm **out)
{
    if (!out || nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(MMDB_ERR_INVALID, "mmdb_comm_init: bad arguments");
    *out = nullptr;
    NcclApi &a = nccl();
    if (!a.so) return fail(MMDB_ERR_UNSUPPORTED, "mmdb_comm_init: " + a.why);
    CU(cudaSetDevice(device));
    nccl_uid u;
    std::memcpy(u.internal, id128, 128);
    mmdb_comm *c = new mmdb_comm();
    c->device = device; c->rank = rank; c->nranks = nranks;
    const int rc = a.init_rank(&c->comm, nranks, u, rank);
    if (rc != 0) {
        delete c;
        return nccl_fail("ncclCommInitRank", rc);
    }
    *out = c;
    return MMDB_OK;
}

extern "C" int mmdb_allreduce_G(mmdb_comm *c, double *G_dev, int64_t n, int fixed_point, void *stream)
{
    if (!c || !c->comm) return fail(MMDB_ERR_INVALID, "mmdb_allreduce_G: null communicator");
    if (c->nranks == 1) return MMDB_OK;                      // nothing to exchange
    NcclApi &a = nccl();
    CU(cudaSetDevice(c->device));
    // deterministic builds hold 2^50-scaled 64-bit integers in G: integer sums do not depend on the reduction order
    const int rc = a.allreduce(G_dev, G_dev, (size_t)n, fixed_point ? NCCL_INT64 : NCCL_FLOAT64, NCCL_SUM, c->comm, (cudaStream_t)stream);
    if (rc != 0) return nccl_fail("ncclAllReduce", rc);
    return MMDB_OK;
}

extern "C" int mmdb_comm_destroy(mmdb_comm *c)
{
    if (!c) return MMDB_OK;
    NcclApi &a = nccl();
    if (c->comm && a.so) a.destroy(c->comm);
    delete c;
    return MMDB_OK;
}
