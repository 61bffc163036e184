// ply_kernels.cu - the fused traversal kernels for sm_100a: map, fold-inner, fold-outer and the fold finish.
//
// One code path serves every expression: operands are staged with cp.async into per-thread 16-byte shared memory
// slots (all loads of a thread in flight before the first use), then the postfix program is interpreted over the
// V lanes of the vector (kcore.h). The kernels differ only in how vectors map to threads and what happens to the
// result: stored (map), accumulated along the vectorised axis (fold-inner: sum, dot, sqrm, row-wise reductions),
// or accumulated per lane while looping over the folded axes (fold-outer: reductions of the leading axes).
// Folds are deterministic: fixed tree inside the CTA, per-chunk partials combined in chunk order by ply_finish.
#include <cuda_runtime.h>
#include "kbody.h"
#include "launch.h"

namespace racu {

struct CpAsync
{
    static __device__ __forceinline__ uint32_t saddr(void * p) { return uint32_t(__cvta_generic_to_shared(p)); }
    __device__ __forceinline__ void copy16(void * s, void const * g) const { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr(s)), "l"(g) : "memory"); }
    __device__ __forceinline__ void copy8(void * s, void const * g) const { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr(s)), "l"(g) : "memory"); }
    __device__ __forceinline__ void copy4(void * s, void const * g) const { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr(s)), "l"(g) : "memory"); }
};
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

template <class W> __device__ __forceinline__ W shfl_xor(W v, int o);
template <> __device__ __forceinline__ uint32_t shfl_xor<uint32_t>(uint32_t v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ uint64_t shfl_xor<uint64_t>(uint64_t v, int o) { return __shfl_xor_sync(0xffffffffu, (unsigned long long)v, o); }

// ---- map: dst[...] = program(operands)
template <class W, int U> __global__ void __launch_bounds__(KBLOCK)
ply_map_kernel(const __grid_constant__ KParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int V = Lanes<W>::V;
    int const tid = threadIdx.x;
    constexpr int sstride = KBLOCK*KSLOT;
    unsigned char * const stage = smem + tid*KSLOT;
    unsigned char * const stack = smem + size_t(U)*p.nstage*sstride + tid*KSLOT;
    constexpr int64_t per = int64_t(KBLOCK)*U;
    for (int64_t g0 = int64_t(blockIdx.x)*per; g0<p.nvecA; g0 += int64_t(gridDim.x)*per) {
        KIdx a[U];
        int nvalid[U];
#pragma unroll
        for (int u=0; u<U; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK + tid;
            nvalid[u] = 0;
            if (g<p.nvecA) { decompA(p, g, a[u]); int64_t const rem = p.len0-a[u].e0; nvalid[u] = rem<V ? int(rem) : V; }
        }
        stage_group<W, U, U, 0>(p, a, nullptr, nvalid, stage, sstride, CpAsync {});
        cp_async_wait_all();
        W t[U][V];
        interp<W, U>(p, stage, stack, sstride, t);
#pragma unroll
        for (int u=0; u<U; ++u) {
            if (nvalid[u]<=0) continue;
            if (p.scatter) scatter_vector<W>(p, reinterpret_cast<W const *>(stack + u*sstride), t[u], nvalid[u]); // stack level 0: the offsets
            else store_vector<W>(p, a[u], t[u]);
        }
    }
}

template <class W> __device__ __forceinline__ W
block_combine(KParams const & p, W r, W * red)
{
    for (int o=16; o>0; o>>=1) { W other = shfl_xor<W>(r, o); r = combine<W>(p.fold_op, p.acc_type, r, other); }
    int const lane = threadIdx.x&31, warp = threadIdx.x>>5;
    if (0==lane) red[warp] = r;
    __syncthreads();
    if (0==threadIdx.x) {
        r = red[0];
        for (int w=1; w<KBLOCK/32; ++w) r = combine<W>(p.fold_op, p.acc_type, r, red[w]);
    }
    return r; // valid in thread 0
}

// ---- fold-inner: CTA = (kept element kb, chunk c of the folded vectors); threads stride over the chunk
template <class W, int U> __global__ void __launch_bounds__(KBLOCK)
ply_fold_inner_kernel(const __grid_constant__ KParams p, const __grid_constant__ FinishParams f)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ W red[KBLOCK/32];
    constexpr int V = Lanes<W>::V;
    int const tid = threadIdx.x;
    constexpr int sstride = KBLOCK*KSLOT;
    unsigned char * const stage = smem + tid*KSLOT;
    unsigned char * const stack = smem + size_t(U)*p.nstage*sstride + tid*KSLOT;
    uint32_t const kb = blockIdx.x/uint32_t(p.nchunks), c = blockIdx.x-kb*uint32_t(p.nchunks);
    KIdx b; b.i[0] = 0;
    if (p.rankB>0) decompB(p, kb, b);
    int64_t const start = int64_t(c)*p.chunk_len;
    int64_t const end = start+p.chunk_len<p.nvecA ? start+p.chunk_len : p.nvecA;
    constexpr int64_t per = int64_t(KBLOCK)*U;
    W acc[V];
#pragma unroll
    for (int j=0; j<V; ++j) acc[j] = W(p.identity);
    for (int64_t g0 = start; g0<end; g0 += per) {
        KIdx a[U];
        int nvalid[U];
#pragma unroll
        for (int u=0; u<U; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK + tid;
            nvalid[u] = 0;
            if (g<end) { decompA(p, g, a[u]); int64_t const rem = p.len0-a[u].e0; nvalid[u] = rem<V ? int(rem) : V; }
        }
        stage_group<W, U, U, 1>(p, a, &b, nvalid, stage, sstride, CpAsync {});
        cp_async_wait_all();
        W t[U][V];
        interp<W, U>(p, stage, stack, sstride, t);
        group_combine<W, U>(p.fold_op, p.acc_type, acc, t, [&](int u, int j) { return j<nvalid[u]; });
    }
    W r = acc[0];
#pragma unroll
    for (int j=1; j<V; ++j) r = combine<W>(p.fold_op, p.acc_type, r, acc[j]);
    r = block_combine<W>(p, r, red);
    if (0==tid) {
        if (p.nchunks>1) static_cast<W *>(p.partial)[int64_t(c)*p.nk+kb] = r;
        else finish_apply<W>(f, kb, r);
    }
}

// ---- fold-outer: thread = kept vector (lanes are distinct kept elements); blockIdx.y = chunk of the folded index space
template <class W, int U> __global__ void __launch_bounds__(KBLOCK)
ply_fold_outer_kernel(const __grid_constant__ KParams p, const __grid_constant__ FinishParams f)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int V = Lanes<W>::V;
    int const tid = threadIdx.x;
    constexpr int sstride = KBLOCK*KSLOT;
    unsigned char * const stage = smem + tid*KSLOT;
    unsigned char * const stack = smem + size_t(U)*p.nstage*sstride + tid*KSLOT;
    int64_t const g = int64_t(blockIdx.x)*KBLOCK + tid;
    bool const valid = g<p.nvecA;
    KIdx a; a.e0 = 0;
    if (valid) decompA(p, g, a);
    uint32_t const c = blockIdx.y;
    int64_t const r0 = int64_t(c)*p.chunk_len;
    int64_t const r1 = r0+p.chunk_len<p.nB ? r0+p.chunk_len : p.nB;
    W acc[V];
#pragma unroll
    for (int j=0; j<V; ++j) acc[j] = W(p.identity);
    if (valid) {
        int64_t const rem0 = p.len0-a.e0;
        int const nv = rem0<V ? int(rem0) : V;
        for (int64_t r = r0; r<r1; r += U) {
            KIdx b[U];
            int nvalid[U];
#pragma unroll
            for (int u=0; u<U; ++u) { nvalid[u] = 0; if (r+u<r1) { decompB(p, r+u, b[u]); nvalid[u] = nv; } }
            stage_group<W, U, 1, U>(p, &a, b, nvalid, stage, sstride, CpAsync {});
            cp_async_wait_all();
            W t[U][V];
            interp<W, U>(p, stage, stack, sstride, t);
            group_combine<W, U>(p.fold_op, p.acc_type, acc, t, [&](int u, int) { return nvalid[u]>0; });
        }
        int64_t const rem = p.len0-a.e0;
        int64_t row = 0;
        if (p.rankA>1) row = fastdiv(uint32_t(g), p.vprdiv);
        int64_t const n0 = row*p.len0+a.e0;
#pragma unroll
        for (int j=0; j<V; ++j) {
            if (j<rem) {
                if (p.nchunks>1) static_cast<W *>(p.partial)[int64_t(c)*p.nk+n0+j] = acc[j];
                else finish_apply<W>(f, n0+j, acc[j]);
            }
        }
    }
}

// ---- finish: combine the per-chunk partials in chunk order and apply to the destination
template <class W> __global__ void __launch_bounds__(KBLOCK)
ply_finish_kernel(const __grid_constant__ FinishParams f)
{
    W const * part = static_cast<W const *>(f.partial);
    if (f.per_warp) {
        int64_t const n = (int64_t(blockIdx.x)*KBLOCK + threadIdx.x)>>5;
        int const lane = threadIdx.x&31;
        if (n>=f.nk) return; // whole warp
        W r = W(f.identity);
        for (int c=lane; c<f.nchunks; c+=32) r = combine<W>(f.fold_op, f.acc_type, r, part[int64_t(c)*f.nk+n]);
        for (int o=16; o>0; o>>=1) { W other = shfl_xor<W>(r, o); r = combine<W>(f.fold_op, f.acc_type, r, other); }
        if (0==lane) finish_apply<W>(f, n, r);
    } else {
        int64_t const n = int64_t(blockIdx.x)*KBLOCK + threadIdx.x;
        if (n>=f.nk) return;
        W r = part[n];
        for (int c=1; c<f.nchunks; ++c) r = combine<W>(f.fold_op, f.acc_type, r, part[int64_t(c)*f.nk+n]);
        finish_apply<W>(f, n, r);
    }
}

// ---- fill
template <class T> __global__ void __launch_bounds__(KBLOCK)
fill_kernel(T * dst, T v, size_t n)
{
    for (size_t i = size_t(blockIdx.x)*KBLOCK+threadIdx.x; i<n; i += size_t(gridDim.x)*KBLOCK) dst[i] = v;
}

thread_local int64_t g_launches = 0;
int64_t launch_count() { return g_launches; }
void launch_count_reset() { g_launches = 0; }
void launch_count_add(int64_t n) { g_launches += n; }

template <class K> static cudaError_t
set_smem(K kernel, size_t smem)
{
    if (smem>32*1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    return cudaSuccess;
}

cudaError_t
launch_plan(Plan const & plan, cudaStream_t stream)
{
    KParams const & p = plan.kp;
    cudaError_t e = cudaSuccess;
    dim3 grid(plan.grid_x, plan.grid_y), block(KBLOCK);
    if (plan.is_static) {
        if ((e = launch_static(plan, stream))) return e;
        ++g_launches;
    } else {
#define RACU_LAUNCH_U(KERNEL, WT, UU, ...) \
    do { if ((e = set_smem(KERNEL<WT, UU>, plan.smem))) return e; KERNEL<WT, UU><<<grid, block, plan.smem, stream>>>(__VA_ARGS__); } while (0)
#define RACU_LAUNCH_W(KERNEL, WT, ...) \
    do { switch (p.U) { case 1: RACU_LAUNCH_U(KERNEL, WT, 1, __VA_ARGS__); break; case 2: RACU_LAUNCH_U(KERNEL, WT, 2, __VA_ARGS__); break; \
         default: RACU_LAUNCH_U(KERNEL, WT, 4, __VA_ARGS__); break; } } while (0)
#define RACU_LAUNCH(KERNEL, ...) \
    do { if (p.wide) RACU_LAUNCH_W(KERNEL, uint64_t, __VA_ARGS__); else RACU_LAUNCH_W(KERNEL, uint32_t, __VA_ARGS__); } while (0)
    switch (p.mode) {
    case K_MAP: RACU_LAUNCH(ply_map_kernel, p); break;
    case K_FOLD_INNER: RACU_LAUNCH(ply_fold_inner_kernel, p, plan.fp); break;
    default: RACU_LAUNCH(ply_fold_outer_kernel, p, plan.fp); break;
    }
#undef RACU_LAUNCH_U
#undef RACU_LAUNCH_W
#undef RACU_LAUNCH
    ++g_launches;
    if ((e = cudaGetLastError())) return e;
    }
    if (plan.need_finish) {
        FinishParams const & f = plan.fp;
        int64_t threads = f.per_warp ? f.nk*32 : f.nk;
        unsigned nb = unsigned((threads+KBLOCK-1)/KBLOCK);
        if (f.wide) ply_finish_kernel<uint64_t><<<nb, KBLOCK, 0, stream>>>(f); else ply_finish_kernel<uint32_t><<<nb, KBLOCK, 0, stream>>>(f);
        ++g_launches;
        e = cudaGetLastError();
    }
    return e;
}

cudaError_t
launch_fill(void * dst, int dtype, racu_base v, size_t n, cudaStream_t stream)
{
    if (0==n) return cudaSuccess;
    unsigned nb = unsigned(std::min<size_t>((n+KBLOCK-1)/KBLOCK, 148*16));
    switch (dtype) {
    case RACU_BOOL: fill_kernel<uint8_t><<<nb, KBLOCK, 0, stream>>>(static_cast<uint8_t *>(dst), uint8_t(v.i!=0), n); break;
    case RACU_I32: fill_kernel<int32_t><<<nb, KBLOCK, 0, stream>>>(static_cast<int32_t *>(dst), int32_t(v.i), n); break;
    case RACU_I64: fill_kernel<int64_t><<<nb, KBLOCK, 0, stream>>>(static_cast<int64_t *>(dst), v.i, n); break;
    case RACU_F32: fill_kernel<float><<<nb, KBLOCK, 0, stream>>>(static_cast<float *>(dst), float(v.f), n); break;
    case RACU_F64: fill_kernel<double><<<nb, KBLOCK, 0, stream>>>(static_cast<double *>(dst), v.f, n); break;
    default: return cudaErrorInvalidValue;
    }
    ++g_launches;
    return cudaGetLastError();
}

} // namespace racu
