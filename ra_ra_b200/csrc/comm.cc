// comm.cc - multi-GPU plumbing of the C ABI: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference is single process (README.md:81); sharding is along the leading axis of the traversal:
//   stencils exchange one halo plane with each slab neighbour (ncclSend/ncclRecv, grouped),
//   folds combine their per-GPU partials with one ncclAllReduce.
//
// NCCL is bound at first use with dlopen("libnccl.so.2") instead of at link time: a process that already carries an NCCL
// (PyTorch bundles its own, newer one under the same soname) must keep using that one; a plain C++ process gets the system's.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "special.h"

struct racu_comm
{
    ncclComm_t comm;
    int nranks, rank;
};

namespace racu {

static struct Nccl
{
    decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&::ncclCommInitRank) CommInitRank = nullptr;
    decltype(&::ncclCommDestroy) CommDestroy = nullptr;
    decltype(&::ncclAllReduce) AllReduce = nullptr;
    decltype(&::ncclSend) Send = nullptr;
    decltype(&::ncclRecv) Recv = nullptr;
    decltype(&::ncclGroupStart) GroupStart = nullptr;
    decltype(&::ncclGroupEnd) GroupEnd = nullptr;
    decltype(&::ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string why;
} nccl;

static bool
nccl_load()
{
    static std::once_flag once;
    std::call_once(once, []{
        void * h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { nccl.why = std::string("dlopen libnccl.so.2: ") + dlerror(); return; }
        bool all = true;
#define RACU_SYM(NAME) do { nccl.NAME = reinterpret_cast<decltype(nccl.NAME)>(dlsym(h, "nccl" #NAME)); all = all && nccl.NAME; } while (0)
        RACU_SYM(GetUniqueId); RACU_SYM(CommInitRank); RACU_SYM(CommDestroy); RACU_SYM(AllReduce); RACU_SYM(Send); RACU_SYM(Recv);
        RACU_SYM(GroupStart); RACU_SYM(GroupEnd); RACU_SYM(GetErrorString);
#undef RACU_SYM
        nccl.ok = all;
        if (!all) nccl.why = "libnccl.so.2 lacks a required symbol";
    });
    if (!nccl.ok) fail(RACU_ERR_NCCL, nccl.why);
    return nccl.ok;
}

static int nccl_fail(ncclResult_t r, char const * where)
{
    fail(RACU_ERR_NCCL, std::string(where) + ": " + nccl.GetErrorString(r));
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
    if (!nccl_load()) return RACU_ERR_NCCL;
    ncclUniqueId id;
    if (ncclResult_t r = nccl.GetUniqueId(&id)) return nccl_fail(r, "ncclGetUniqueId");
    std::memset(id_out, 0, RACU_UNIQUE_ID_BYTES);
    std::memcpy(id_out, &id, sizeof(id));
    return RACU_OK;
}

int racu_comm_init_rank(racu_comm ** comm, const void * id_in, int nranks, int rank)
{
    if (!comm || !id_in || nranks<1 || rank<0 || rank>=nranks) return fail(RACU_ERR_BAD_DESC, "bad communicator arguments");
    if (!nccl_load()) return RACU_ERR_NCCL;
    ncclUniqueId id;
    std::memcpy(&id, id_in, sizeof(id));
    racu_comm * c = new racu_comm {nullptr, nranks, rank};
    if (ncclResult_t r = nccl.CommInitRank(&c->comm, nranks, id, rank)) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    *comm = c;
    return RACU_OK;
}

int racu_comm_destroy(racu_comm * comm)
{
    if (!comm) return RACU_OK;
    ncclResult_t r = nccl.CommDestroy(comm->comm);
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
    if (ncclResult_t r = nccl.AllReduce(buf, buf, count, t, op, comm->comm, cudaStream_t(stream))) return nccl_fail(r, "ncclAllReduce");
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
    ncclResult_t r = nccl.GroupStart();
    if (!r && lo) r = nccl.Send(send_lo, count, t, comm->rank-1, comm->comm, st);
    if (!r && lo) r = nccl.Recv(recv_lo, count, t, comm->rank-1, comm->comm, st);
    if (!r && hi) r = nccl.Send(send_hi, count, t, comm->rank+1, comm->comm, st);
    if (!r && hi) r = nccl.Recv(recv_hi, count, t, comm->rank+1, comm->comm, st);
    ncclResult_t r2 = nccl.GroupEnd();
    if (r) return nccl_fail(r, "halo send/recv");
    if (r2) return nccl_fail(r2, "ncclGroupEnd");
    return RACU_OK;
}

// ---- peer memory for racu_stencil_push: CUDA IPC handles of racu_alloc'ed buffers

static_assert(sizeof(cudaIpcMemHandle_t)<=RACU_IPC_HANDLE_BYTES, "ipc handle size");

int racu_ipc_export(void * dev_ptr, void * handle_out)
{
    if (!dev_ptr || !handle_out) return fail(RACU_ERR_BAD_DESC, "null argument");
    cudaIpcMemHandle_t h;
    if (cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr)) return cuda_fail(e, "cudaIpcGetMemHandle");
    std::memset(handle_out, 0, RACU_IPC_HANDLE_BYTES);
    std::memcpy(handle_out, &h, sizeof(h));
    return RACU_OK;
}

int racu_ipc_open(const void * handle, void ** dev_ptr_out)
{
    if (!handle || !dev_ptr_out) return fail(RACU_ERR_BAD_DESC, "null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    if (cudaError_t e = cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess)) return cuda_fail(e, "cudaIpcOpenMemHandle");
    return RACU_OK;
}

int racu_ipc_close(void * dev_ptr)
{
    if (!dev_ptr) return RACU_OK;
    if (cudaError_t e = cudaIpcCloseMemHandle(dev_ptr)) return cuda_fail(e, "cudaIpcCloseMemHandle");
    return RACU_OK;
}

} // extern "C"
