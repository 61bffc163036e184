// racu_shim.cc - the racu_* C ABI (include/racu.h) over HOST memory, executed by the CPU oracle.
// TEST INFRASTRUCTURE ONLY. It exists so that the C++ known-answer tests of the host header (tests/cxx/*.cc) can be
// linked once against this file (no GPU: checks the header's flattening, slicing and agreement logic against the oracle)
// and once against the CUDA library (GPU: the parity run). The product never links or loads it.
#include "../include/racu.h"
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

// RACU_SHIM_PREPARE=<path of libracu.so>: every descriptor this stand-in executes is also handed to the product's racu_prepare (plan +
// run-time specialisation, no GPU, nothing launched: include/racu.h), so that a CPU run of the known-answer tests leaves the kernels the GPU
// run will need in the specialiser's disk cache. Allocations are 256-byte aligned like cudaMalloc's so that the planner sees the same
// alignment classes.
static void
shim_prepare(racu_desc const * d, int redop)
{
    using Prepare = int (*)(racu_desc const *, int, char const **);
    static Prepare fn = [] {
        char const * path = std::getenv("RACU_SHIM_PREPARE");
        void * h = path && *path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
        return h ? reinterpret_cast<Prepare>(dlsym(h, "racu_prepare")) : nullptr;
    }();
    if (fn) { char const * k = nullptr; fn(d, redop, &k); }
}

extern "C" {
const char * oracle_last_error_string(void);
int oracle_agree(racu_desc const *, int32_t *, int64_t *);
int oracle_map(racu_desc const *);
int oracle_reduce(racu_desc const *, int, void *);
int oracle_stencil_raw(racu_stencil_desc const *);
int oracle_gemm_raw(racu_gemm_desc const *);

int racu_version(void) { return -RACU_VERSION; } // negative: "this is the oracle stand-in"
const char * racu_last_error_string(void) { return oracle_last_error_string(); }
int racu_device_count(void) { return 0; }
int racu_set_device(int) { return RACU_OK; }
int racu_alloc(void ** p, size_t n) { *p = std::aligned_alloc(256, (n+255)/256*256 + 256); return *p ? RACU_OK : RACU_ERR_ALLOC; }
int racu_free(void * p) { std::free(p); return RACU_OK; }
int racu_memcpy_h2d(void * d, const void * s, size_t n, void *) { std::memcpy(d, s, n); return RACU_OK; }
int racu_memcpy_d2h(void * d, const void * s, size_t n, void *) { std::memcpy(d, s, n); return RACU_OK; }
int racu_memcpy_d2d(void * d, const void * s, size_t n, void *) { std::memmove(d, s, n); return RACU_OK; }
int racu_fill(void * d, int dtype, racu_base v, size_t count, void *)
{
    for (size_t i=0; i<count; ++i) {
        switch (dtype) {
        case RACU_BOOL: static_cast<bool *>(d)[i] = v.i!=0; break;
        case RACU_I32: static_cast<int32_t *>(d)[i] = int32_t(v.i); break;
        case RACU_I64: static_cast<int64_t *>(d)[i] = v.i; break;
        case RACU_F32: static_cast<float *>(d)[i] = float(v.f); break;
        case RACU_F64: static_cast<double *>(d)[i] = v.f; break;
        default: return RACU_ERR_BAD_DESC;
        }
    }
    return RACU_OK;
}
int racu_stream_sync(void *) { return RACU_OK; }
int racu_host_alloc(void ** p, size_t n) { return racu_alloc(p, n); }
int racu_host_free(void * p) { return racu_free(p); }
int racu_agree(const racu_desc * d, int32_t * r, int64_t * l) { return oracle_agree(d, r, l); }
int racu_map(const racu_desc * d, void *) { shim_prepare(d, 0); return oracle_map(d); }
int racu_reduce(const racu_desc * d, int redop, void * out, void *) { shim_prepare(d, redop); return oracle_reduce(d, redop, out); }
int racu_reduce_dev(const racu_desc * d, int redop, void * out, void *) { shim_prepare(d, redop); return oracle_reduce(d, redop, out); }
int racu_reduce_async(const racu_desc * d, int redop, void * out, void *) { shim_prepare(d, redop); return oracle_reduce(d, redop, out); }
void racu_plan_cache_stats(int64_t * h, int64_t * m) { if (h) *h = 0; if (m) *m = 0; }
int racu_prepare(const racu_desc *, int, const char ** k) { if (k) *k = "oracle"; return RACU_OK; }
int64_t racu_launch_count(void) { return 0; }
void racu_launch_count_reset(void) {}
const char * racu_last_kernel(void) { return "oracle"; }
int racu_stencil(const racu_stencil_desc * s, void *) { return oracle_stencil_raw(s); }
// the meaning of racu_stencil_steps: `steps` single sweeps with the roles of the two buffers swapped after each (bench/bench-stencil3.cc:117-121)
int racu_stencil_steps(const racu_stencil_desc * s, int32_t steps, void *, void *)
{
    if (steps<0 || 0!=s->lo0 || 0!=s->hi0) return RACU_ERR_BAD_DESC;
    for (int t=0; t<steps; ++t) {
        racu_stencil_desc d = *s;
        if (t&1) { d.src = s->dst; d.dst = const_cast<void *>(s->src); for (int k=0; k<3; ++k) { d.src_stride[k] = s->dst_stride[k]; d.dst_stride[k] = s->src_stride[k]; } }
        if (int e = oracle_stencil_raw(&d)) return e;
    }
    return RACU_OK;
}
int racu_gemm(const racu_gemm_desc * g, void *) { return oracle_gemm_raw(g); }
int racu_gemv(int dtype, int trans, int64_t M, int64_t N, const void * a, int64_t lda, const void * x, int64_t incx,
              void * y, int64_t incy, void *)
{
    auto run = [&](auto t) {
        using T = decltype(t);
        T const * A = static_cast<T const *>(a); T const * X = static_cast<T const *>(x); T * Y = static_cast<T *>(y);
        if (!trans) { for (int64_t i=0; i<M; ++i) for (int64_t j=0; j<N; ++j) Y[i*incy] += A[i*lda+j]*X[j*incx]; }
        else { for (int64_t i=0; i<M; ++i) for (int64_t j=0; j<N; ++j) Y[j*incy] += X[i*incx]*A[i*lda+j]; }
    };
    if (dtype==RACU_F32) run(float {}); else if (dtype==RACU_F64) run(double {}); else return RACU_ERR_UNSUPPORTED;
    return RACU_OK;
}
int racu_comm_unique_id(void *) { return RACU_ERR_UNSUPPORTED; }
int racu_comm_init_rank(racu_comm **, const void *, int, int) { return RACU_ERR_UNSUPPORTED; }
int racu_comm_destroy(racu_comm *) { return RACU_ERR_UNSUPPORTED; }
int racu_allreduce(racu_comm *, void *, size_t, int, int, void *) { return RACU_ERR_UNSUPPORTED; }
int racu_halo_exchange(racu_comm *, const void *, void *, const void *, void *, size_t, int, void *) { return RACU_ERR_UNSUPPORTED; }
int racu_ipc_export(void *, void *) { return RACU_ERR_UNSUPPORTED; }
int racu_ipc_open(const void *, void **) { return RACU_ERR_UNSUPPORTED; }
int racu_ipc_close(void *) { return RACU_ERR_UNSUPPORTED; }
// Host-memory meaning of the fused step (all "ranks" live in one process and are stepped one after the other): sweep, then
// copy the interior of the first / last owned plane into the neighbour's ghost plane and raise its flag.
int racu_stencil_push(const racu_stencil_desc * s, const racu_halo_push * h, void *)
{
    if (int e = oracle_stencil_raw(s)) return e;
    size_t const es = s->dtype==RACU_F32 ? 4 : 8;
    for (int side=0; side<2; ++side) {
        char * peer = static_cast<char *>(side ? h->push_hi : h->push_lo);
        if (!peer) continue;
        int64_t const i = side ? s->len[0]-2 : 1;
        char const * src = static_cast<char const *>(s->dst) + i*s->dst_stride[0]*es;
        for (int64_t j=1; j<s->len[1]-1; ++j)
            for (int64_t k=1; k<s->len[2]-1; ++k)
                std::memcpy(peer + (j*s->dst_stride[1]+k*s->dst_stride[2])*es, src + (j*s->dst_stride[1]+k*s->dst_stride[2])*es, es);
        *(side ? h->signal_hi : h->signal_lo) = h->step+1;
    }
    return RACU_OK;
}
}
