// api.cc - the C ABI (include/racu.h) of the CUDA library: storage, traversal, reductions.
// Stencil, contraction and communicator entry points live in stencil.cu, gemm.cu and comm.cc.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "launch.h"
#include "special.h"

namespace racu {

thread_local std::string g_error;
thread_local const char * g_kernel = "";

int fail(int status, std::string const & msg) { g_error = msg; return status; }
int cuda_fail(cudaError_t e, char const * where)
{
    g_error = std::string(where) + ": " + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
    return RACU_ERR_CUDA;
}
void set_kernel(const char * k) { g_kernel = k; }

static int
run_plan(racu_desc const * d, int redop, void * reduce_out, cudaStream_t stream)
{
    Plan plan;
    std::string err;
    if (int e = make_plan(d, redop, reduce_out, plan, err)) {
        if (e==RACU_ERR_SEQ) { // sequential semantics are observable: the folded axes one index at a time, in row-major order
            Shape sh;
            if (int e2 = agree(d, sh, err)) return fail(e2, err);
            int const k = first_folded_axis(d, sh);
            int64_t steps = 1;
            for (int j=0; j<sh.rank; ++j) { if (sh.len[j]>1 && (j>=d->opnd[d->dst].rank || 0==d->opnd[d->dst].stride[j])) steps *= sh.len[j]; }
            if (k<0 || steps>(int64_t(1)<<16)) return fail(RACU_ERR_UNSUPPORTED, err + ": more than 65536 sequential steps");
            racu_desc sub;
            for (int64_t i=0; i<sh.len[k]; ++i) {
                peel_axis(*d, k, i, sub);
                if (int e2 = run_plan(&sub, 0, nullptr, stream)) return e2;
            }
            return RACU_OK;
        }
        if (e!=RACU_ERR_PEEL) return fail(e, err);
        // more uncollapsible axes than one launch handles: peel the outermost expression axis here
        if (reduce_out) { // as `out OP= expr` onto a rank-0 destination, which composes slice by slice
            racu_desc d2;
            int t; uint64_t bits;
            if (!peel_reduction(*d, redop, reduce_out, d2, t, bits)) return fail(RACU_ERR_UNSUPPORTED, err);
            racu_base v;
            if (t==RACU_F32) v.f = bits_f32(uint32_t(bits)); else if (t==RACU_F64) v.f = bits_f64(bits);
            else if (t==RACU_I32) v.i = int32_t(uint32_t(bits)); else v.i = int64_t(bits);
            if (cudaError_t e3 = launch_fill(reduce_out, t, v, 1, stream)) return cuda_fail(e3, "fill");
            return run_plan(&d2, 0, nullptr, stream);
        }
        if (!can_peel(d, reduce_out!=nullptr)) return fail(RACU_ERR_UNSUPPORTED, err);
        Shape sh;
        if (int e2 = agree(d, sh, err)) return fail(e2, err);
        racu_desc sub;
        for (int64_t i = peel_last_only(d) ? sh.len[0]-1 : 0; i<sh.len[0]; ++i) {
            peel_axis0(*d, i, sub);
            if (int e2 = run_plan(&sub, redop, reduce_out, stream)) return e2;
        }
        return RACU_OK;
    }
    if (plan.empty) {
        g_kernel = "none (empty)";
        if (plan.fill_identity) { // empty reduction: the start value (test/reduction.cc:199-201, :236-238)
            racu_base v;
            int t = plan.kp.acc_type;
            if (t==RACU_F32) v.f = bits_f32(uint32_t(plan.kp.identity)); else if (t==RACU_F64) v.f = bits_f64(plan.kp.identity);
            else if (t==RACU_I32) v.i = int32_t(uint32_t(plan.kp.identity)); else v.i = int64_t(plan.kp.identity);
            if (cudaError_t e = launch_fill(reduce_out, t, v, 1, stream)) return cuda_fail(e, "fill");
        }
        return RACU_OK;
    }
    void * partial = nullptr;
    if (plan.partial_bytes) {
        if (cudaError_t e = cudaMallocAsync(&partial, plan.partial_bytes, stream)) return cuda_fail(e, "cudaMallocAsync");
        plan.kp.partial = partial; plan.fp.partial = partial; plan.sp.partial = partial;
    }
    plan.fp.identity = plan.kp.identity;
    g_kernel = plan.kernel;
    cudaError_t e = cudaSuccess;
    if (!try_transpose(plan, stream, e)) e = launch_plan(plan, stream);
    if (partial) { cudaError_t e2 = cudaFreeAsync(partial, stream); if (!e) e = e2; }
    if (e) return cuda_fail(e, plan.kernel);
    return RACU_OK;
}

} // namespace racu

using namespace racu;

extern "C" {

int racu_version(void) { return RACU_VERSION; }
const char * racu_last_error_string(void) { return g_error.c_str(); }
int racu_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n)) { cudaGetLastError(); return 0; } return n; }
int racu_set_device(int dev) { if (cudaError_t e = cudaSetDevice(dev)) return cuda_fail(e, "cudaSetDevice"); return RACU_OK; }

int racu_alloc(void ** ptr, size_t bytes)
{
    if (!ptr) return fail(RACU_ERR_BAD_DESC, "null pointer");
    if (cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16)) { cudaGetLastError(); g_error = std::string("cudaMalloc: ") + cudaGetErrorString(e); return RACU_ERR_ALLOC; }
    return RACU_OK;
}
int racu_free(void * ptr) { if (cudaError_t e = cudaFree(ptr)) return cuda_fail(e, "cudaFree"); return RACU_OK; }
int racu_memcpy_h2d(void * d, const void * s, size_t n, void * st)
{
    if (cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, cudaStream_t(st))) return cuda_fail(e, "memcpy h2d");
    return RACU_OK;
}
int racu_memcpy_d2h(void * d, const void * s, size_t n, void * st)
{
    if (cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, cudaStream_t(st))) return cuda_fail(e, "memcpy d2h");
    if (cudaError_t e = cudaStreamSynchronize(cudaStream_t(st))) return cuda_fail(e, "memcpy d2h sync");
    return RACU_OK;
}
int racu_memcpy_d2d(void * d, const void * s, size_t n, void * st)
{
    if (cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, cudaStream_t(st))) return cuda_fail(e, "memcpy d2d");
    return RACU_OK;
}
int racu_fill(void * d, int dtype, racu_base v, size_t count, void * st)
{
    if (dtype<0 || dtype>=RACU_NDTYPE) return fail(RACU_ERR_BAD_DESC, "bad dtype");
    if (cudaError_t e = launch_fill(d, dtype, v, count, cudaStream_t(st))) return cuda_fail(e, "fill");
    return RACU_OK;
}
int racu_stream_sync(void * st) { if (cudaError_t e = cudaStreamSynchronize(cudaStream_t(st))) return cuda_fail(e, "stream sync"); return RACU_OK; }
int racu_host_alloc(void ** ptr, size_t bytes) { if (cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 16)) { cudaGetLastError(); g_error = cudaGetErrorString(e); return RACU_ERR_ALLOC; } return RACU_OK; }
int racu_host_free(void * ptr) { if (cudaError_t e = cudaFreeHost(ptr)) return cuda_fail(e, "cudaFreeHost"); return RACU_OK; }

int racu_agree(const racu_desc * d, int32_t * rank_out, int64_t * len_out)
{
    if (!d) return fail(RACU_ERR_BAD_DESC, "null descriptor");
    Shape s; std::string err;
    if (int e = agree(d, s, err)) return fail(e, err);
    if (rank_out) *rank_out = s.rank;
    if (len_out) { for (int k=0; k<s.rank; ++k) len_out[k] = s.len[k]; }
    return RACU_OK;
}

int racu_map(const racu_desc * d, void * stream)
{
    { std::string err; if (int e = validate_desc(d, false, err)) return fail(e, err); } // before anything indexes into the descriptor
    int handled = 0;
    if (int e = try_special_map(d, cudaStream_t(stream), &handled)) return e;
    if (handled) return RACU_OK;
    return run_plan(d, 0, nullptr, cudaStream_t(stream));
}

int racu_prepare(const racu_desc * d, int redop, const char ** kernel_out)
{
    if (!d) return fail(RACU_ERR_BAD_DESC, "null descriptor");
    Plan plan;
    std::string err;
    if (int e = validate_desc(d, redop!=0, err)) return fail(e, err);
    racu_desc sub;
    racu_desc const * cur = d;
    void * const out = redop ? reinterpret_cast<void *>(uintptr_t(256)) : nullptr; // planning only looks at its alignment
    for (int depth=0; depth<RACU_MAX_RANK; ++depth) {
        int e = make_plan(cur, redop, out, plan, err);
        if (RACU_OK==e) break;
        if (e==RACU_ERR_SEQ) { // every step of the sequential schedule runs the same map
            Shape sh;
            if (int e2 = agree(cur, sh, err)) return fail(e2, err);
            int const k = first_folded_axis(cur, sh);
            if (k<0) return fail(RACU_ERR_UNSUPPORTED, err);
            racu_desc next;
            peel_axis(*cur, k, 0, next);
            sub = next; cur = &sub;
            continue;
        }
        if (e!=RACU_ERR_PEEL) return fail(e, err);
        if (!can_peel(cur, out!=nullptr)) return fail(RACU_ERR_UNSUPPORTED, err);
        racu_desc next;
        peel_axis0(*cur, 0, next); // every slice of the peeled axis runs the same program
        sub = next; cur = &sub;
    }
    if (kernel_out) *kernel_out = plan.empty ? "none (empty)" : plan.kernel;
    return RACU_OK;
}

int racu_reduce_dev(const racu_desc * d, int redop, void * out_dev, void * stream)
{
    if (!d || !out_dev) return fail(RACU_ERR_BAD_DESC, "null argument");
    { std::string err; if (int e = validate_desc(d, true, err)) return fail(e, err); }
    return run_plan(d, redop, out_dev, cudaStream_t(stream));
}

int racu_reduce(const racu_desc * d, int redop, void * out_host, void * stream)
{
    if (!d || !out_host) return fail(RACU_ERR_BAD_DESC, "null argument");
    std::string err;
    if (int e = validate_desc(d, true, err)) return fail(e, err);
    int rt = check_program(d, err);
    if (rt<0) return fail(RACU_ERR_BAD_DESC, err);
    void * tmp = nullptr;
    cudaStream_t st = cudaStream_t(stream);
    if (cudaError_t e = cudaMallocAsync(&tmp, 16, st)) return cuda_fail(e, "cudaMallocAsync");
    int r = run_plan(d, redop, tmp, st);
    if (RACU_OK==r) {
        cudaError_t e = cudaMemcpyAsync(out_host, tmp, dtype_size(rt), cudaMemcpyDeviceToHost, st);
        if (!e) e = cudaStreamSynchronize(st);
        if (e) r = cuda_fail(e, "reduce readback");
    }
    cudaFreeAsync(tmp, st);
    return r;
}

int64_t racu_launch_count(void) { return launch_count(); }
void racu_launch_count_reset(void) { launch_count_reset(); }
const char * racu_last_kernel(void) { return g_kernel; }

} // extern "C"
