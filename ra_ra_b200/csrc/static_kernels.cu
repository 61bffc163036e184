// static_kernels.cu - compile-time specialisations of the ply kernels for the hottest program shapes.
//
// The generic kernels (ply_kernels.cu) interpret the postfix program; on flat traversals (one collapsed contiguous axis,
// ra/ply.hh:42-46 with `keep` true on every leaf) of a short single-type program the interpretive overhead, not HBM, is
// the limit. Here the SAME postfix program is a constexpr array: the evaluator loop is fully unrolled by the compiler,
// the value stack lives in registers, operands are read with 128-bit loads issued UNR deep before the first use.
// Same ops, same order, same types, no contraction (-fmad=false): results are bit-identical to the interpreter's.
//
// The planner (plan.cc: match_static) picks a specialisation when
//   * every instruction has one type T in {f32, f64, i32} and there is no CAST,
//   * every operand is a 16-byte aligned unit-stride vector (S_VEC) of T, or a scalar of T,
//   * the traversal is flat: map with rankA==1; fold-inner with rankA==1 (any kept rows); fold-outer with rankA==rankB==1.
#include <cuda_runtime.h>
#include "kbody.h"
#include "launch.h"
#include "static_progs.h"

namespace racu {

template <class T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int V = 4; };
template <> struct VecOf<int32_t> { using type = int4; static constexpr int V = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };

template <class T> __device__ __forceinline__ void
ld16(T * out, void const * p) // streaming 128-bit load: read once, do not keep in L1
{
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    uint32_t w[4] = {a, b, c, d};
    memcpy(out, w, 16);
}
template <class T> __device__ __forceinline__ void
ld16_reused(T * out, void const * p) // operand re-read by every kept row (the vector of gemv): let L1 keep it
{
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    uint32_t w[4] = {a, b, c, d};
    memcpy(out, w, 16);
}
template <class T> __device__ __forceinline__ void
ld16_rw(T * out, void const * p) // destination operand (read-modify-write): ordinary load
{
    uint4 q = *reinterpret_cast<uint4 const *>(p);
    memcpy(out, &q, 16);
}
template <class T> __device__ __forceinline__ void
st16(void * p, T const * v)
{
    uint4 q;
    memcpy(&q, v, 16);
    *reinterpret_cast<uint4 *>(p) = q;
}

// The constexpr postfix evaluator. x[k][j]: lane j of input k (role SP_Y is the destination's old value).
template <class T, int ID> __device__ __forceinline__ void
eval_static(T const (*x)[VecOf<T>::V], T * out)
{
    constexpr int V = VecOf<T>::V;
    constexpr SProg P = static_prog(ID);
    T st[P.n>0 ? P.n : 1][V];
    int sp = 0;
#pragma unroll
    for (int pc=0; pc<P.n; ++pc) {
        int const op = P.ins[pc].op, arg = P.ins[pc].arg;
        if (op==RACU_OP_PUSH) {
#pragma unroll
            for (int j=0; j<V; ++j) st[sp][j] = x[arg][j];
            ++sp;
        } else if (op==RACU_OP_NEG || op==RACU_OP_SQR || op==RACU_OP_ABS) {
#pragma unroll
            for (int j=0; j<V; ++j) { T a = st[sp-1][j]; st[sp-1][j] = op==RACU_OP_NEG ? T(-a) : op==RACU_OP_SQR ? T(a*a) : t_abs<T>(a); }
        } else {
#pragma unroll
            for (int j=0; j<V; ++j) {
                T a = st[sp-2][j], b = st[sp-1][j];
                st[sp-2][j] = op==RACU_OP_ADD ? T(a+b) : op==RACU_OP_SUB ? T(a-b) : op==RACU_OP_MUL ? T(a*b) : op==RACU_OP_DIV ? T(a/b)
                    : op==RACU_OP_MIN ? (b<a ? b : a) : (a<b ? b : a);
            }
            --sp;
        }
    }
#pragma unroll
    for (int j=0; j<V; ++j) out[j] = st[0][j];
}

template <class T, int FOLD> __device__ __forceinline__ T
comb(T c, T x)
{
    if constexpr (FOLD==RACU_ASSIGN_ADD || FOLD==RACU_ASSIGN_SUB) return T(c+x);
    else if constexpr (FOLD==RACU_ASSIGN_MUL || FOLD==RACU_ASSIGN_DIV) return T(c*x);
    else if constexpr (FOLD==RACU_ASSIGN_AMIN) return x<c ? x : c;
    else return c<x ? x : c;
}

template <class T> __device__ __forceinline__ T from_bits(uint64_t b);
template <> __device__ __forceinline__ float from_bits<float>(uint64_t b) { return __uint_as_float(uint32_t(b)); }
template <> __device__ __forceinline__ int32_t from_bits<int32_t>(uint64_t b) { return int32_t(uint32_t(b)); }
template <> __device__ __forceinline__ double from_bits<double>(uint64_t b) { return __longlong_as_double((long long)b); }
template <class T> __device__ __forceinline__ uint64_t to_bits(T v);
template <> __device__ __forceinline__ uint64_t to_bits<float>(float v) { return __float_as_uint(v); }
template <> __device__ __forceinline__ uint64_t to_bits<int32_t>(int32_t v) { return uint32_t(v); }
template <> __device__ __forceinline__ uint64_t to_bits<double>(double v) { return uint64_t(__double_as_longlong(v)); }

constexpr int SUNR = 4; // vectors in flight per thread per input

// Fill x[k] for one vector at byte offset `off` from each input's base (+ rowoff[k]). Scalars are pre-broadcast in xs.
// REUSE is a compile-time switch: a run-time choice between the two load forms makes the compiler load into temporaries
// and select right after each load, which serialises the loads on their latency (measured: 7.0 -> 3.0 TB/s on dot).
template <class T, int ID, bool REUSE=false> __device__ __forceinline__ void
load_inputs(SParams const & p, T (*x)[VecOf<T>::V], T const (*xs)[VecOf<T>::V], int64_t const * rowoff, int64_t off)
{
    constexpr int V = VecOf<T>::V;
    constexpr SProg P = static_prog(ID);
#pragma unroll
    for (int k=0; k<P.nin; ++k) {
        if (p.scalar[k]) {
#pragma unroll
            for (int j=0; j<V; ++j) x[k][j] = xs[k][j];
        } else if (SK_ROW==p.kind[k]) { // step 0 along the vectorised axis: one value per row (gevm: c(j) += a(i)*b(i,j))
            T v = __ldg(reinterpret_cast<T const *>(reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[k]));
#pragma unroll
            for (int j=0; j<V; ++j) x[k][j] = v;
        } else if (k==P.yrole) {
            ld16_rw<T>(x[k], reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[k] + off);
        } else {
            if constexpr (REUSE) {
                if (p.reused[k]) ld16_reused<T>(x[k], reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[k] + off);
                else ld16<T>(x[k], reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[k] + off);
            } else {
                ld16<T>(x[k], reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[k] + off);
            }
        }
    }
}

// ---- map: dst[...] = P(inputs[...]) over up to KGR uncollapsible axes; axis 0 is vectorised (16 bytes per access).
// Inputs per role: aligned unit-step vectors (forward or reversed), one value per row (prefix rank extension: the
// `v` of B += A*C + v, examples/read-me.cc:32), or scalars. A flat traversal (rankA==1) may have a scalar tail.
// GENERAL=false: every non-scalar input is a forward unit-step vector. The kind is then known at compile time, which matters:
// a run-time choice between load forms makes the compiler select right after each load and serialises them.
template <class T, int ID, bool GENERAL> __global__ void __launch_bounds__(KBLOCK)
sflat_map_kernel(const __grid_constant__ SParams p)
{
    constexpr int V = VecOf<T>::V;
    constexpr SProg P = static_prog(ID);
    T xs[P.nin][V];
#pragma unroll
    for (int k=0; k<P.nin; ++k) {
#pragma unroll
        for (int j=0; j<V; ++j) xs[k][j] = from_bits<T>(p.in[k]);
    }
    int64_t const per = int64_t(KBLOCK)*SUNR;
    for (int64_t g0 = int64_t(blockIdx.x)*per + threadIdx.x; g0<p.nvec; g0 += int64_t(gridDim.x)*per) {
        T x[SUNR][P.nin][V];
        int64_t doff[SUNR];
#pragma unroll
        for (int u=0; u<SUNR; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK;
            if (g<p.nvec) {
                int64_t e0;
                uint32_t i1 = 0, i2 = 0, i3 = 0;
                if (1==p.rankA) { e0 = g*V; }
                else {
                    uint32_t n = uint32_t(g), q = fastdiv(n, p.vprdiv);
                    e0 = int64_t(n-q*uint32_t(p.vpr))*V; n = q;
                    if (2==p.rankA) i1 = n;
                    else { q = fastdiv(n, p.dA[1]); i1 = n-q*p.dA[1].len; n = q;
                        if (3==p.rankA) i2 = n; else { q = fastdiv(n, p.dA[2]); i2 = n-q*p.dA[2].len; i3 = q; } }
                }
                doff[u] = e0 + int64_t(i1)*p.dsA[1] + int64_t(i2)*p.dsA[2] + int64_t(i3)*p.dsA[3];
#pragma unroll
                for (int k=0; k<P.nin; ++k) {
                    int const kd = GENERAL ? int(p.kind[k]) : (p.scalar[k] ? int(SK_SCALAR) : int(SK_VEC));
                    if (SK_SCALAR==kd) {
#pragma unroll
                        for (int j=0; j<V; ++j) x[u][k][j] = xs[k][j];
                    } else {
                        int64_t off = int64_t(i1)*p.sA[k][1] + int64_t(i2)*p.sA[k][2] + int64_t(i3)*p.sA[k][3];
                        T const * base = reinterpret_cast<T const *>(p.in[k]);
                        if (!GENERAL || SK_VEC==kd) {
                            if (k==P.yrole) ld16_rw<T>(x[u][k], base + off + e0); else ld16<T>(x[u][k], base + off + e0);
                        } else if (SK_VECREV==kd) {
                            T r[V];
                            ld16<T>(r, base + off - e0 - (V-1));
#pragma unroll
                            for (int j=0; j<V; ++j) x[u][k][j] = r[V-1-j];
                        } else { // SK_ROW
                            T v = __ldg(base + off);
#pragma unroll
                            for (int j=0; j<V; ++j) x[u][k][j] = v;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u=0; u<SUNR; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK;
            if (g<p.nvec) { T out[V]; eval_static<T, ID>(x[u], out); st16<T>(static_cast<T *>(p.dst) + doff[u], out); }
        }
    }
    if (1==p.rankA && 0==blockIdx.x && 0==threadIdx.x) { // tail elements of a flat traversal, scalar
        for (int64_t e = p.nvec*V; e<p.n; ++e) {
            T x[P.nin][V], out[V];
#pragma unroll
            for (int k=0; k<P.nin; ++k) {
                T const * base = reinterpret_cast<T const *>(p.in[k]);
                x[k][0] = SK_SCALAR==p.kind[k] ? xs[k][0] : SK_VEC==p.kind[k] ? base[e] : SK_VECREV==p.kind[k] ? base[-e] : base[0];
            }
#pragma unroll
            for (int k=0; k<P.nin; ++k) { for (int j=1; j<V; ++j) x[k][j] = x[k][0]; }
            eval_static<T, ID>(x, out);
            static_cast<T *>(p.dst)[e] = out[0];
        }
    }
}

template <class T, int FOLD> __device__ __forceinline__ T
block_comb(T r, T * red)
{
    for (int o=16; o>0; o>>=1) { T other = __shfl_xor_sync(0xffffffffu, r, o); r = comb<T, FOLD>(r, other); }
    int const lane = threadIdx.x&31, warp = threadIdx.x>>5;
    if (0==lane) red[warp] = r;
    __syncthreads();
    if (0==threadIdx.x) {
        r = red[0];
        for (int w=1; w<KBLOCK/32; ++w) r = comb<T, FOLD>(r, red[w]);
    }
    return r;
}

// ---- fold-inner, flat rows: CTA = (RB consecutive rows starting at kb*RB, chunk c); row length len0 (multiple of V unless
// there is a single row). RB > 1 (rankB == 1 only) blocks rows in registers so that operands that do not move along the
// kept axis (the vector of gemv: c(i) += a(i,j)*b(j), ra/ra.hh:418) are loaded once per RB rows instead of once per row.
template <class T, int ID, int FOLD, bool REUSE> __global__ void __launch_bounds__(KBLOCK)
sflat_fold_inner_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f)
{
    using W = std::conditional_t<sizeof(T)==8, uint64_t, uint32_t>;
    constexpr int V = VecOf<T>::V;
    constexpr SProg P = static_prog(ID);
    constexpr int RB = 1;               // (register row blocking measured slower than L1 reuse: kept at 1)
    constexpr int UN = SUNR;            // vectors in flight per input per row
    __shared__ T red[KBLOCK/32];
    // chunk-major order (gemv): CTAs resident together work on the SAME column chunk of consecutive rows, so the operand
    // every row re-reads stays in L1; row-major order otherwise
    uint32_t kb, c;
    if (p.chunk_major) { c = blockIdx.x/uint32_t(p.nk/RB); kb = blockIdx.x-c*uint32_t(p.nk/RB); }
    else { kb = blockIdx.x/uint32_t(p.nchunks); c = blockIdx.x-kb*uint32_t(p.nchunks); }
    T xs[P.nin][V];
    int64_t rowoff[RB][P.nin];
    {
        // row kb -> group-B position -> per input byte offset
        uint32_t n = kb*RB, q, ib[KGR] = {0, 0, 0, 0};
        if (p.rankB<=1) ib[0] = n;
        else {
            q = fastdiv(n, p.dB[0]); ib[0] = n-q*p.dB[0].len; n = q;
            if (2==p.rankB) ib[1] = n;
            else { q = fastdiv(n, p.dB[1]); ib[1] = n-q*p.dB[1].len; n = q;
                if (3==p.rankB) ib[2] = n; else { q = fastdiv(n, p.dB[2]); ib[2] = n-q*p.dB[2].len; ib[3] = q; } }
        }
#pragma unroll
        for (int k=0; k<P.nin; ++k) {
            int64_t off = 0;
#pragma unroll
            for (int j=0; j<KGR; ++j) { if (j<p.rankB) off += int64_t(ib[j])*p.sB[k][j]; }
#pragma unroll
            for (int r=0; r<RB; ++r) rowoff[r][k] = (off + int64_t(r)*p.sB[k][0])*int64_t(sizeof(T));
#pragma unroll
            for (int j=0; j<V; ++j) xs[k][j] = from_bits<T>(p.in[k]);
        }
    }
    int64_t const nfull = p.n/V;
    int64_t const start = int64_t(c)*p.chunk_len;
    int64_t const end = start+p.chunk_len<nfull ? start+p.chunk_len : nfull;
    T acc[RB][V];
#pragma unroll
    for (int r=0; r<RB; ++r)
#pragma unroll
        for (int j=0; j<V; ++j) acc[r][j] = from_bits<T>(p.identity);
    int64_t const per = int64_t(KBLOCK)*UN;
    for (int64_t g0 = start + threadIdx.x; g0<end; g0 += per) {
        T x[RB][UN][P.nin][V];
#pragma unroll
        for (int u=0; u<UN; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK;
            if (g<end) {
#pragma unroll
                for (int r=0; r<RB; ++r) {
                    if (0==r) load_inputs<T, ID, REUSE>(p, x[0][u], xs, rowoff[0], g*16);
                    else {
#pragma unroll
                        for (int k=0; k<P.nin; ++k) {
                            if (0==p.sB[k][0] || p.scalar[k]) { // does not move along the kept axis: reuse row 0's registers
#pragma unroll
                                for (int j=0; j<V; ++j) x[r][u][k][j] = x[0][u][k][j];
                            } else if (SK_ROW==p.kind[k]) {
                                T v = __ldg(reinterpret_cast<T const *>(reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[r][k]));
#pragma unroll
                                for (int j=0; j<V; ++j) x[r][u][k][j] = v;
                            } else {
                                ld16<T>(x[r][u][k], reinterpret_cast<unsigned char const *>(p.in[k]) + rowoff[r][k] + g*16);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u=0; u<UN; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK;
            if (g<end) {
#pragma unroll
                for (int r=0; r<RB; ++r) {
                    T out[V];
                    eval_static<T, ID>(x[r][u], out);
#pragma unroll
                    for (int j=0; j<V; ++j) acc[r][j] = comb<T, FOLD>(acc[r][j], out[j]);
                }
            }
        }
    }
#pragma unroll
    for (int r=0; r<RB; ++r) {
        T v = acc[r][0];
#pragma unroll
        for (int j=1; j<V; ++j) v = comb<T, FOLD>(v, acc[r][j]);
        if (1==RB && c+1==uint32_t(p.nchunks) && 0==threadIdx.x) { // tail elements of the row (single flat row only)
            for (int64_t e = nfull*V; e<p.n; ++e) {
                T x[P.nin][V], out[V];
#pragma unroll
                for (int k=0; k<P.nin; ++k) x[k][0] = p.scalar[k] ? xs[k][0] : reinterpret_cast<T const *>(reinterpret_cast<unsigned char const *>(p.in[k])+rowoff[0][k])[SK_ROW==p.kind[k] ? 0 : e];
#pragma unroll
                for (int k=0; k<P.nin; ++k) { for (int j=1; j<V; ++j) x[k][j] = x[k][0]; }
                eval_static<T, ID>(x, out);
                v = comb<T, FOLD>(v, out[0]);
            }
        }
        if (r>0) __syncthreads(); // red[] reuse
        v = block_comb<T, FOLD>(v, red);
        if (0==threadIdx.x) {
            int64_t const row = int64_t(kb)*RB + r;
            if (p.nchunks>1) static_cast<W *>(p.partial)[int64_t(c)*p.nk+row] = W(to_bits<T>(v));
            else finish_apply<W>(f, row, W(to_bits<T>(v)));
        }
    }
}

// ---- fold-outer, flat: thread = kept vector g (lanes are kept elements), blockIdx.y = chunk of the folded rows
template <class T, int ID, int FOLD> __global__ void __launch_bounds__(KBLOCK)
sflat_fold_outer_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f)
{
    using W = std::conditional_t<sizeof(T)==8, uint64_t, uint32_t>;
    constexpr int V = VecOf<T>::V;
    constexpr SProg P = static_prog(ID);
    int64_t const g = int64_t(blockIdx.x)*KBLOCK + threadIdx.x;
    if (g>=p.n/V) return;
    uint32_t const c = blockIdx.y;
    int64_t const r0 = int64_t(c)*p.chunk_len;
    int64_t const r1 = r0+p.chunk_len<p.nB ? r0+p.chunk_len : p.nB;
    T xs[P.nin][V];
#pragma unroll
    for (int k=0; k<P.nin; ++k) {
#pragma unroll
        for (int j=0; j<V; ++j) xs[k][j] = from_bits<T>(p.in[k]);
    }
    T acc[V];
#pragma unroll
    for (int j=0; j<V; ++j) acc[j] = from_bits<T>(p.identity);
    for (int64_t r = r0; r<r1; r += SUNR) {
        T x[SUNR][P.nin][V];
#pragma unroll
        for (int u=0; u<SUNR; ++u) {
            if (r+u<r1) {
                int64_t rowoff[P.nin];
#pragma unroll
                for (int k=0; k<P.nin; ++k) rowoff[k] = (r+u)*p.sB[k][0]*int64_t(sizeof(T));
                load_inputs<T, ID>(p, x[u], xs, rowoff, g*16);
            }
        }
#pragma unroll
        for (int u=0; u<SUNR; ++u) {
            if (r+u<r1) {
                T out[V];
                eval_static<T, ID>(x[u], out);
#pragma unroll
                for (int j=0; j<V; ++j) acc[j] = comb<T, FOLD>(acc[j], out[j]);
            }
        }
    }
#pragma unroll
    for (int j=0; j<V; ++j) {
        if (p.nchunks>1) static_cast<W *>(p.partial)[int64_t(c)*p.nk+g*V+j] = W(to_bits<T>(acc[j]));
        else finish_apply<W>(f, g*V+j, W(to_bits<T>(acc[j])));
    }
}

// ---------------- dispatch ----------------

template <class T, int ID> static cudaError_t
launch_one(Plan const & plan, dim3 grid, cudaStream_t stream)
{
    SParams const & p = plan.sp;
    constexpr SProg P = static_prog(ID);
    if (p.mode==K_MAP) {
        if constexpr (!P.fold) {
            bool general = false;
            for (int k=0; k<P.nin; ++k) general = general || p.kind[k]==SK_VECREV || p.kind[k]==SK_ROW;
            if (general) sflat_map_kernel<T, ID, true><<<grid, KBLOCK, 0, stream>>>(p); else sflat_map_kernel<T, ID, false><<<grid, KBLOCK, 0, stream>>>(p);
        }
    } else if constexpr (P.fold) {
#define RACU_SF(KERNEL, ...) \
        switch (p.fold_op) { \
        case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: KERNEL<T, ID, RACU_ASSIGN_ADD __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        case RACU_ASSIGN_AMIN: KERNEL<T, ID, RACU_ASSIGN_AMIN __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        case RACU_ASSIGN_AMAX: KERNEL<T, ID, RACU_ASSIGN_AMAX __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        default: KERNEL<T, ID, RACU_ASSIGN_MUL __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        }
        if (p.mode==K_FOLD_INNER) {
            bool any = false;
            for (int k=0; k<SP_MAXIN; ++k) any = any || p.reused[k];
            if (any) { RACU_SF(sflat_fold_inner_kernel, , true) } else { RACU_SF(sflat_fold_inner_kernel, , false) }
        } else { RACU_SF(sflat_fold_outer_kernel) }
#undef RACU_SF
    }
    return cudaGetLastError();
}

template <int ID> static cudaError_t
launch_typed(Plan const & plan, dim3 grid, cudaStream_t stream)
{
    switch (plan.sp.dtype) {
    case RACU_F32: return launch_one<float, ID>(plan, grid, stream);
    case RACU_F64: return launch_one<double, ID>(plan, grid, stream);
    default: return launch_one<int32_t, ID>(plan, grid, stream);
    }
}

template <int ID=0> static cudaError_t
launch_id(Plan const & plan, dim3 grid, cudaStream_t stream)
{
    if constexpr (ID<SP_COUNT) {
        if (plan.sp.id==ID) return launch_typed<ID>(plan, grid, stream);
        return launch_id<ID+1>(plan, grid, stream);
    } else {
        return cudaErrorInvalidValue;
    }
}

cudaError_t
launch_static(Plan const & plan, cudaStream_t stream)
{
    dim3 grid(plan.grid_x, plan.grid_y);
    return launch_id<0>(plan, grid, stream);
}

} // namespace racu
