// api.cc - the C ABI (include/racu.h) of the CUDA library: storage, traversal, reductions.
// Stencil, contraction and communicator entry points live in stencil.cu, gemm.cu and comm.cc.
#include <cuda_runtime.h>

#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

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

// ---- per-call overhead: plans are cached, fold scratch is kept.
//
// Plan cache (per thread): key = the bytes of the descriptor that are in use + redop + the reduce destination + the device. Base
// pointers are part of the key on purpose: alignment and aliasing decide the plan. A loop that issues the same expression again
// (benchmarks, the CG iteration, a pointer-swapping sweep = two entries) skips make_plan and the specialiser's lookup.
// RACU_NO_PLAN_CACHE=1 plans every call (the tuning knobs are read at plan time).
struct PlanKey
{
    std::string bytes;
    bool operator==(PlanKey const & o) const { return bytes==o.bytes; }
};
static void
make_key(racu_desc const * d, int redop, void * reduce_out, PlanKey & k)
{
    int dev = -1;
    if (cudaGetDevice(&dev)) { cudaGetLastError(); dev = -1; }
    // the planner's environment knobs are read at plan time: their values are part of the key, so changing one takes effect at once
    uint64_t env = 1469598103934665603ull;
    for (char const * name: {"RACU_NO_STATIC", "RACU_NO_JIT", "RACU_U", "RACU_NO_TRANSPOSE", "RACU_ROWS_MINROWS", "RACU_ROWS_MAXVEC", "RACU_OUTER_CTAS"}) {
        char const * v = std::getenv(name);
        env = (env ^ (v ? 0x9e : 0x3b))*1099511628211ull;
        for (; v && *v; ++v) env = (env ^ uint8_t(*v))*1099511628211ull;
    }
    int64_t const head[5] = {(int64_t(d->nopnd)<<32) | uint32_t(d->nins), (int64_t(d->dst)<<32) | uint32_t(d->assign), (int64_t(redop)<<32) | uint32_t(dev), int64_t(uintptr_t(reduce_out)), int64_t(env)};
    k.bytes.assign(reinterpret_cast<char const *>(head), sizeof(head));
    k.bytes.append(reinterpret_cast<char const *>(d->opnd), sizeof(racu_operand)*size_t(d->nopnd));
    k.bytes.append(reinterpret_cast<char const *>(d->ins), sizeof(racu_ins)*size_t(d->nins));
}
struct PlanCache
{
    static constexpr int N = 64;
    struct Entry { PlanKey key; Plan plan; uint64_t used = 0; bool valid = false; };
    Entry e[N];
    uint64_t tick = 0;
    Plan * find(PlanKey const & k)
    {
        for (auto & x: e) { if (x.valid && x.key==k) { x.used = ++tick; return &x.plan; } }
        return nullptr;
    }
    Plan * insert(PlanKey && k, Plan const & p)
    {
        Entry * v = &e[0];
        for (auto & x: e) { if (!x.valid) { v = &x; break; } if (x.used<v->used) v = &x; }
        v->key = std::move(k); v->plan = p; v->used = ++tick; v->valid = true;
        return &v->plan;
    }
};
static PlanCache & plan_cache() { static thread_local PlanCache * c = new PlanCache; return *c; } // (leaked at thread exit: no CUDA calls in a destructor)
static thread_local int64_t g_plan_hits = 0, g_plan_misses = 0;

// Fold scratch (per thread, device and stream): the partials of a chunked fold. Work on one stream is ordered, so consecutive folds
// share the allocation; it only grows. Replaces a cudaMallocAsync / cudaFreeAsync pair per fold.
struct Scratch { int dev; cudaStream_t stream; void * ptr; size_t bytes; };
static void *
fold_scratch(size_t bytes, cudaStream_t stream, cudaError_t & err)
{
    static thread_local std::vector<Scratch> * all = new std::vector<Scratch>;
    int dev = 0;
    if ((err = cudaGetDevice(&dev))) return nullptr;
    Scratch * s = nullptr;
    for (auto & x: *all) { if (x.dev==dev && x.stream==stream) { s = &x; break; } }
    if (!s) { all->push_back(Scratch {dev, stream, nullptr, 0}); s = &all->back(); }
    if (s->bytes<bytes) {
        if (s->ptr) { if ((err = cudaStreamSynchronize(stream))) return nullptr; cudaFree(s->ptr); s->ptr = nullptr; s->bytes = 0; }
        size_t const want = std::max<size_t>(bytes+bytes/2, size_t(1)<<20);
        if ((err = cudaMalloc(&s->ptr, want))) { s->ptr = nullptr; return nullptr; }
        s->bytes = want;
    }
    return s->ptr;
}

static int
run_plan(racu_desc const * d, int redop, void * reduce_out, cudaStream_t stream)
{
    static bool const no_cache = nullptr!=std::getenv("RACU_NO_PLAN_CACHE");
    PlanKey key;
    Plan * cached = nullptr;
    if (!no_cache) { make_key(d, redop, reduce_out, key); cached = plan_cache().find(key); }
    Plan local;
    std::string err;
    if (cached) ++g_plan_hits;
    else ++g_plan_misses;
    if (int e = cached ? RACU_OK : make_plan(d, redop, reduce_out, local, err)) {
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
    if (!cached && !no_cache) cached = plan_cache().insert(std::move(key), local);
    Plan & plan = cached ? *cached : local;
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
    cudaError_t e = cudaSuccess;
    if (plan.partial_bytes) {
        void * partial = fold_scratch(plan.partial_bytes, stream, e);
        if (!partial) return cuda_fail(e, "fold scratch");
        plan.kp.partial = partial; plan.fp.partial = partial; plan.sp.partial = partial;
    }
    plan.fp.identity = plan.kp.identity;
    g_kernel = plan.kernel;
    if (!try_transpose(plan, stream, e)) e = launch_plan(plan, stream);
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
int racu_host_alloc(void ** ptr, size_t bytes) { if (cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocPortable | cudaHostAllocMapped)) { cudaGetLastError(); g_error = cudaGetErrorString(e); return RACU_ERR_ALLOC; } return RACU_OK; }
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

// The scalar goes straight into page-locked host memory that the device writes over PCIe (one slot per thread): no allocation, no
// copy call, one stream synchronise. (Round 1: cudaMallocAsync + kernel + finish + cudaMemcpyAsync + sync + cudaFreeAsync per call.)
static void *
result_slot()
{
    static thread_local void * slot = nullptr;
    if (!slot) { if (cudaError_t e = cudaHostAlloc(&slot, 64, cudaHostAllocPortable | cudaHostAllocMapped)) { (void)e; cudaGetLastError(); slot = nullptr; } }
    return slot;
}

int racu_reduce(const racu_desc * d, int redop, void * out_host, void * stream)
{
    if (!d || !out_host) return fail(RACU_ERR_BAD_DESC, "null argument");
    std::string err;
    if (int e = validate_desc(d, true, err)) return fail(e, err);
    int rt = check_program(d, err);
    if (rt<0) return fail(RACU_ERR_BAD_DESC, err);
    void * slot = result_slot();
    if (!slot) return fail(RACU_ERR_ALLOC, "cudaHostAlloc failed");
    cudaStream_t st = cudaStream_t(stream);
    int r = run_plan(d, redop, slot, st);
    if (RACU_OK==r) {
        if (cudaError_t e = cudaStreamSynchronize(st)) return cuda_fail(e, "reduce sync");
        std::memcpy(out_host, slot, dtype_size(rt));
    }
    return r;
}

int racu_reduce_async(const racu_desc * d, int redop, void * out_pinned, void * stream)
{
    if (!d || !out_pinned) return fail(RACU_ERR_BAD_DESC, "null argument");
    { std::string err; if (int e = validate_desc(d, true, err)) return fail(e, err); }
    cudaPointerAttributes at;
    if (cudaError_t e = cudaPointerGetAttributes(&at, out_pinned)) { cudaGetLastError(); return cuda_fail(e, "racu_reduce_async"); }
    if (at.type!=cudaMemoryTypeHost && at.type!=cudaMemoryTypeManaged && at.type!=cudaMemoryTypeDevice)
        return fail(RACU_ERR_BAD_DESC, "racu_reduce_async: the result must go to page-locked host memory (racu_host_alloc) or device memory");
    return run_plan(d, redop, at.type==cudaMemoryTypeHost ? at.devicePointer : out_pinned, cudaStream_t(stream));
}

void racu_plan_cache_stats(int64_t * hits, int64_t * misses) { if (hits) *hits = g_plan_hits; if (misses) *misses = g_plan_misses; }

int64_t racu_launch_count(void) { return launch_count(); }
void racu_launch_count_reset(void) { launch_count_reset(); }
const char * racu_last_kernel(void) { return g_kernel; }

} // extern "C"
