// comm.cc - multi-GPU plumbing of the C ABI: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference is single process (README.md:81); sharding is along the leading axis of the traversal:
//   stencils exchange one halo plane with each slab neighbour (ncclSend/ncclRecv, grouped),
//   folds combine their per-GPU partials with one ncclAllReduce.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "special.h"

struct racu_comm
{
    ncclComm_t comm;
    int nranks, rank;
};

namespace racu {
static int nccl_fail(ncclResult_t r, char const * where)
{
    fail(RACU_ERR_NCCL, std::string(where) + ": " + ncclGetErrorString(r));
    return RACU_ERR_NCCL;
}
static bool nccl_type(int dtype, ncclDataType_t & t)
{
    switch (dtype) {
    case RACU_BOOL: t = ncclUint8; return true;
    case RACU_I32: t = ncclInt32; return true;
    case RACU_I64: t = ncclInt64; return true;
    case RACU_F32: t = ncclFloat32; return true;
    case RACU_F64: t = ncclFloat64; return true;
    }
    return false;
}
} // namespace racu

using namespace racu;

extern "C" {

static_assert(sizeof(ncclUniqueId)<=RACU_UNIQUE_ID_BYTES, "unique id size");

int racu_comm_unique_id(void * id_out)
{
    if (!id_out) return fail(RACU_ERR_BAD_DESC, "null id");
    ncclUniqueId id;
    if (ncclResult_t r = ncclGetUniqueId(&id)) return nccl_fail(r, "ncclGetUniqueId");
    std::memset(id_out, 0, RACU_UNIQUE_ID_BYTES);
    std::memcpy(id_out, &id, sizeof(id));
    return RACU_OK;
}

int racu_comm_init_rank(racu_comm ** comm, const void * id_in, int nranks, int rank)
{
    if (!comm || !id_in || nranks<1 || rank<0 || rank>=nranks) return fail(RACU_ERR_BAD_DESC, "bad communicator arguments");
    ncclUniqueId id;
    std::memcpy(&id, id_in, sizeof(id));
    racu_comm * c = new racu_comm {nullptr, nranks, rank};
    if (ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank)) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    *comm = c;
    return RACU_OK;
}

int racu_comm_destroy(racu_comm * comm)
{
    if (!comm) return RACU_OK;
    ncclResult_t r = ncclCommDestroy(comm->comm);
    delete comm;
    if (r) return nccl_fail(r, "ncclCommDestroy");
    return RACU_OK;
}

int racu_allreduce(racu_comm * comm, void * buf, size_t count, int dtype, int redop, void * stream)
{
    if (!comm || !buf) return fail(RACU_ERR_BAD_DESC, "null argument");
    ncclDataType_t t;
    if (!nccl_type(dtype, t)) return fail(RACU_ERR_BAD_DESC, "bad dtype");
    ncclRedOp_t op;
    switch (redop) {
    case RACU_RED_SUM: op = ncclSum; break;
    case RACU_RED_PROD: op = ncclProd; break;
    case RACU_RED_AMIN: op = ncclMin; break;
    case RACU_RED_AMAX: op = ncclMax; break;
    default: return fail(RACU_ERR_BAD_DESC, "bad redop");
    }
    if (ncclResult_t r = ncclAllReduce(buf, buf, count, t, op, comm->comm, cudaStream_t(stream))) return nccl_fail(r, "ncclAllReduce");
    return RACU_OK;
}

int racu_halo_exchange(racu_comm * comm, const void * send_lo, void * recv_lo, const void * send_hi, void * recv_hi,
                       size_t count, int dtype, void * stream)
{
    if (!comm) return fail(RACU_ERR_BAD_DESC, "null communicator");
    ncclDataType_t t;
    if (!nccl_type(dtype, t)) return fail(RACU_ERR_BAD_DESC, "bad dtype");
    cudaStream_t st = cudaStream_t(stream);
    bool lo = comm->rank>0, hi = comm->rank+1<comm->nranks;
    ncclResult_t r = ncclGroupStart();
    if (!r && lo) r = ncclSend(send_lo, count, t, comm->rank-1, comm->comm, st);
    if (!r && lo) r = ncclRecv(recv_lo, count, t, comm->rank-1, comm->comm, st);
    if (!r && hi) r = ncclSend(send_hi, count, t, comm->rank+1, comm->comm, st);
    if (!r && hi) r = ncclRecv(recv_hi, count, t, comm->rank+1, comm->comm, st);
    ncclResult_t r2 = ncclGroupEnd();
    if (r) return nccl_fail(r, "halo send/recv");
    if (r2) return nccl_fail(r2, "ncclGroupEnd");
    return RACU_OK;
}

} // extern "C"
