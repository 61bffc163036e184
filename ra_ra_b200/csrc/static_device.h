// static_device.h - device side of the specialised ply kernels: the compile-time typed postfix evaluator and the map /
// fold-inner / fold-outer bodies built on it. Included by static_kernels.cu (the table of pre-compiled program shapes, nvcc)
// and, verbatim, by the run-time specialiser (jit.cc: the same bodies instantiated by NVRTC for a program that is not in
// the table). Device code only: no host headers beyond what kcore.h needs.
#pragma once
#include "kbody.h"
#include "static_progs.h"

namespace racu {

template <int ID> struct SPT
{
    static constexpr SProg P = static_prog(ID);
    static constexpr bool wide = prog_wide(P);
    using W = std::conditional_t<wide, uint64_t, uint32_t>;
};

__host__ __device__ constexpr int es_of(int t) { return t==RACU_BOOL ? 1 : (t==RACU_I32 || t==RACU_F32) ? 4 : 8; }

enum LdForm { LD_STREAM = 0, LD_REUSED = 1, LD_RW = 2 };

template <int FORM> __device__ __forceinline__ uint4
ld16(void const * p)
{
    uint4 r;
    if constexpr (FORM==LD_STREAM) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if constexpr (FORM==LD_REUSED) asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else r = *reinterpret_cast<uint4 const *>(p);
    return r;
}

// 64-bit lanes are loaded as 64-bit registers: packing two 32-bit halves after the load is an ALU op that depends on the load
// and stalls on it, which serialises the loads.
template <int FORM> __device__ __forceinline__ void
ld16x2(uint64_t & a, uint64_t & b, void const * p)
{
    if constexpr (FORM==LD_STREAM) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    else if constexpr (FORM==LD_REUSED) asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    else { ulonglong2 q = *reinterpret_cast<ulonglong2 const *>(p); a = q.x; b = q.y; }
}

// Operand vectors are held as RAW 32-bit words exactly as loaded (4 words for SV 4-byte elements, 8 for 8-byte ones, 1 for
// bools); they become typed slot values only at PUSH time inside the evaluator. Anything computed from a loaded register
// right after the load (packing halves, widening) would stall on that load and serialise the loads.
template <int FORM> __device__ __forceinline__ void
load_words(uint32_t * x, unsigned char const * p, int t)
{
    int const es = es_of(t);
    if (4==es) { uint4 q = ld16<FORM>(p); x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w; }
    else if (8==es) {
        uint4 q0 = ld16<FORM>(p), q1 = ld16<FORM>(p+16);
        x[0] = q0.x; x[1] = q0.y; x[2] = q0.z; x[3] = q0.w; x[4] = q1.x; x[5] = q1.y; x[6] = q1.z; x[7] = q1.w;
    } else { x[0] = *reinterpret_cast<uint32_t const *>(p); }
}
// one element broadcast to every lane, in raw words
__device__ __forceinline__ void
splat_words(uint32_t * x, unsigned char const * p, int t)
{
    int const es = es_of(t);
    if (4==es) { uint32_t v = *reinterpret_cast<uint32_t const *>(p); x[0] = v; x[1] = v; x[2] = v; x[3] = v; }
    else if (8==es) { uint2 v = *reinterpret_cast<uint2 const *>(p); x[0] = v.x; x[1] = v.y; x[2] = v.x; x[3] = v.y; x[4] = v.x; x[5] = v.y; x[6] = v.x; x[7] = v.y; }
    else { uint32_t v = *p ? 0x01010101u : 0u; x[0] = v; }
}
__device__ __forceinline__ void
splat_bits(uint32_t * x, uint64_t bits, int t)
{
    int const es = es_of(t);
    if (4==es) { uint32_t v = uint32_t(bits); x[0] = v; x[1] = v; x[2] = v; x[3] = v; }
    else if (8==es) { uint32_t lo = uint32_t(bits), hi = uint32_t(bits>>32); x[0] = lo; x[1] = hi; x[2] = lo; x[3] = hi; x[4] = lo; x[5] = hi; x[6] = lo; x[7] = hi; }
    else { x[0] = bits ? 0x01010101u : 0u; }
}
// SEQ role: lane j holds origin + (off + j*s0) computed in the role's type, exactly like seq_value (kcore.h) / Seq (ra/expr.hh:86-91)
__device__ __forceinline__ void
seq_words(uint32_t * x, uint64_t org, int64_t off, int64_t s0, int t)
{
#pragma unroll
    for (int j=0; j<SV; ++j) {
        int64_t const o = off + j*s0;
        if (t==RACU_I32) x[j] = uint32_t(int32_t(int32_t(int64_t(org)) + int32_t(o)));
        else if (t==RACU_F32) x[j] = __float_as_uint(float(__longlong_as_double((long long)org)) + float(o));
        else if (t==RACU_I64) { uint64_t const v = uint64_t(int64_t(org) + o); x[2*j] = uint32_t(v); x[2*j+1] = uint32_t(v>>32); }
        else if (t==RACU_F64) { uint64_t const v = uint64_t(__double_as_longlong(__longlong_as_double((long long)org) + double(o))); x[2*j] = uint32_t(v); x[2*j+1] = uint32_t(v>>32); }
    }
}
template <class W> __device__ __forceinline__ W
word_lane(uint32_t const * x, int j, int t) // lane j of a raw vector as slot bits
{
    int const es = es_of(t);
    if (4==es) return W(x[j]);
    if (8==es) { if constexpr (sizeof(W)==8) return uint64_t(x[2*j]) | (uint64_t(x[2*j+1])<<32); else return W(0); }
    return W((x[0]>>(8*j))&0xffu ? 1 : 0);
}
template <class W> __device__ __forceinline__ void
store_vec(unsigned char * p, W const * v, int t)
{
    int const es = es_of(t);
    if (4==es) { uint4 q; q.x = uint32_t(v[0]); q.y = uint32_t(v[1]); q.z = uint32_t(v[2]); q.w = uint32_t(v[3]); *reinterpret_cast<uint4 *>(p) = q; }
    else if (8==es) {
        if constexpr (sizeof(W)==8) {
            ulonglong2 q0, q1;
            q0.x = v[0]; q0.y = v[1]; q1.x = v[2]; q1.y = v[3];
            *reinterpret_cast<ulonglong2 *>(p) = q0; *reinterpret_cast<ulonglong2 *>(p+16) = q1;
        }
    } else { *reinterpret_cast<uint32_t *>(p) = (uint32_t(v[0])&1u) | ((uint32_t(v[1])&1u)<<8) | ((uint32_t(v[2])&1u)<<16) | ((uint32_t(v[3])&1u)<<24); }
}

// words per raw operand vector, and vectors in flight per thread: about 48 registers of operand data
template <int ID> struct SPX
{
    static constexpr int XW = SPT<ID>::wide ? 2*SV : SV;
    static constexpr int words() { int w = 0; for (int r=0; r<SPT<ID>::P.nin; ++r) w += SV*es_of(SPT<ID>::P.rtype[r])/4 + (es_of(SPT<ID>::P.rtype[r])<4); return w>0 ? w : 1; }
    static constexpr int UNR = 48/words()>=4 ? 4 : 48/words()>=2 ? 48/words() : 1;
};

// Called between a group of loads and the first use of their results. The assembler's scheduler otherwise sinks the later loads below
// the arithmetic on the earlier ones to save registers, and since issue is in order the first dependent instruction then holds back
// the loads behind it (two to five loads in flight per thread instead of all of them: dot ran at 6.3 instead of 7.1 TB/s; moving the
// arithmetic behind a branch did not help, the loads were sunk into it). Here one word of every loaded vector is folded into a value
// the compiler cannot know to be zero (SParams::zero, a kernel parameter) and that value is xor-ed into every lane of role 0: every
// result lane now depends on every load, so every load is issued before any arithmetic. One integer instruction per 16 bytes loaded
// plus one per lane: noise for kernels that issue a few percent of what the SM can.
template <int ID, int NU> __device__ __forceinline__ void
loads_issued(SParams const & p, uint32_t (*x)[SPT<ID>::P.nin][SPX<ID>::XW])
{
    constexpr SProg P = SPT<ID>::P;
    uint32_t t = 0;
#pragma unroll
    for (int u=0; u<NU; ++u)
#pragma unroll
        for (int k=0; k<P.nin; ++k) t |= x[u][k][0];
    t &= uint32_t(p.zero);
    constexpr int es0 = es_of(P.rtype[0]), wpl = es0>=4 ? es0/4 : 0; // words per lane of role 0 (bools: 4 lanes in word 0)
#pragma unroll
    for (int u=0; u<NU; ++u) {
        if constexpr (0==wpl) x[u][0][0] ^= t;
        else {
#pragma unroll
            for (int j=0; j<SV; ++j) x[u][0][j*wpl] ^= t;
        }
    }
}

// The constexpr typed postfix evaluator. x[k]: raw words of role k's vector.
template <int ID> __device__ __forceinline__ void
eval_static(SParams const & p, uint32_t const (*x)[SPX<ID>::XW], typename SPT<ID>::W * out)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    W st[P.n>0 ? P.n : 1][SV];
    int sp = 0;
#pragma unroll
    for (int pc=0; pc<P.n; ++pc) {
        int const op = P.ins[pc].op, arg = P.ins[pc].arg, type = P.ins[pc].type, aux = P.ins[pc].aux;
        if (op==RACU_OP_PUSH) {
#pragma unroll
            for (int j=0; j<SV; ++j) st[sp][j] = word_lane<W>(x[arg], j, type);
            ++sp;
        } else if (op==RACU_OP_CAST) {
#pragma unroll
            for (int j=0; j<SV; ++j) st[sp-1][j] = cast_any<W>(aux, type, st[sp-1][j]);
        } else if ((op>=RACU_OP_NEG && op<=RACU_OP_NOT) || (op>=RACU_OP_TAN && op<=RACU_OP_ISINF)) {
#pragma unroll
            for (int j=0; j<SV; ++j) st[sp-1][j] = un_any<W>(op, type, st[sp-1][j]);
        } else if (op==RACU_OP_LOAD) { // gather (ra/ply.hh:461-499): role arg is the array (SK_PTR), the offset is clamped to its valid range
            unsigned char const * const gb = reinterpret_cast<unsigned char const *>(p.in[arg]);
#pragma unroll
            for (int j=0; j<SV; ++j) {
                int64_t off = aux==RACU_I32 ? int64_t(int32_t(uint32_t(st[sp-1][j]))) : int64_t(st[sp-1][j]);
                off = off<p.sA[arg][0] ? p.sA[arg][0] : off>p.sA[arg][1] ? p.sA[arg][1] : off;
                st[sp-1][j] = load_elem<W>(gb + off*es_of(type), type);
            }
        } else if (op==RACU_OP_PICK) { // stack: sel a0 .. a(n-1) -> a(sel); n = arg, selector type = aux (ra/expr.hh:686-728)
#pragma unroll
            for (int j=0; j<SV; ++j) {
                W sel = st[sp-1-arg][j];
                if (aux==RACU_BOOL) sel = W(uint32_t(sel)&0xffu); else if (aux==RACU_I32) sel = W(uint32_t(sel));
                W r = st[sp-arg][j];             // out of contract selectors take branch 0, like the oracle
#pragma unroll
                for (int b=1; b<arg; ++b) r = sel==W(b) ? st[sp-arg+b][j] : r;
                st[sp-1-arg][j] = r;
            }
            sp -= arg;
        } else if (op==RACU_OP_FMA || op==RACU_OP_CLAMP || op==RACU_OP_LERP) {
#pragma unroll
            for (int j=0; j<SV; ++j) st[sp-3][j] = tern_any<W>(op, type, st[sp-3][j], st[sp-2][j], st[sp-1][j]);
            sp -= 2;
        } else { // binary
#pragma unroll
            for (int j=0; j<SV; ++j) st[sp-2][j] = bin_any<W>(op, type, st[sp-2][j], st[sp-1][j]);
            --sp;
        }
    }
#pragma unroll
    for (int j=0; j<SV; ++j) out[j] = st[0][j];
}

// role k of the vector at (e0, off) -> raw words. The operand kind is a compile-time constant when KINDS says so (run-time specialised
// kernels, and the table's `B += A*C + v` layouts); else it is read from p.kind[k] (GENERAL) or is scalar / forward vector (!GENERAL).
template <int ID, bool GENERAL, int FORM, uint64_t KINDS> __device__ __forceinline__ void
fetch_role(SParams const & p, int k, uint32_t * x, int64_t off, int64_t e0)
{
    constexpr SProg P = SPT<ID>::P;
    int const t = P.rtype[k], es = es_of(t);
    int const kd = KINDS!=SK_RUNTIME ? int((KINDS>>(4*k))&15u) : GENERAL ? int(p.kind[k]) : (SK_SCALAR==p.kind[k] ? int(SK_SCALAR) : int(SK_VEC));
    constexpr bool any = GENERAL || KINDS!=SK_RUNTIME;
    unsigned char const * base = reinterpret_cast<unsigned char const *>(p.in[k]);
    if (SK_SCALAR==kd) splat_bits(x, p.in[k], t);
    else if (any && SK_PTR==kd) {} // addressed by RACU_OP_LOAD inside the evaluator
    else if (any && SK_SEQ==kd) seq_words(x, p.in[k], off + e0*p.sA[k][0], p.sA[k][0], t);
    else if (!any || SK_VEC==kd) {
        if (k==P.yrole) load_words<LD_RW>(x, base + (off+e0)*es, t); else load_words<FORM>(x, base + (off+e0)*es, t);
    } else if (SK_VECREV==kd) { // lanes reversed: swap whole elements
        uint32_t r[SPX<ID>::XW];
        load_words<FORM>(r, base + (off-e0-(SV-1))*es, t);
        if (4==es) { x[0] = r[3]; x[1] = r[2]; x[2] = r[1]; x[3] = r[0]; }
        else if (8==es) { x[0] = r[6]; x[1] = r[7]; x[2] = r[4]; x[3] = r[5]; x[4] = r[2]; x[5] = r[3]; x[6] = r[0]; x[7] = r[1]; }
        else { x[0] = __byte_perm(r[0], 0, 0x0123); }
    } else splat_words(x, base + off*es, t); // SK_ROW
}

// ---- map: dst[...] = P(inputs[...]) over up to KGR uncollapsible axes; axis 0 is vectorised (SV elements per access).
// Inputs per role: aligned unit-step vectors (forward or reversed), one value per row (prefix rank extension: the
// `v` of B += A*C + v, examples/read-me.cc:32), or scalars. A flat traversal (rankA==1) may have a scalar tail.
// GENERAL=false: every non-scalar input is a forward unit-step vector.
template <int ID, bool GENERAL, uint64_t KINDS=SK_RUNTIME> __device__ __forceinline__ void
smap_body(SParams const & p)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    constexpr int des = es_of(P.otype), UNR = SPX<ID>::UNR, XW = SPX<ID>::XW;
    int64_t const per = int64_t(KBLOCK)*UNR;
    // vector g -> element index along axis 0 and the destination / input row offsets
    auto locate = [&](int64_t g, int64_t & e0, int64_t & doff, int64_t * off) {
        uint32_t i1 = 0, i2 = 0, i3 = 0;
        if (1==p.rankA) { e0 = g*SV; }
        else {
            uint32_t n = uint32_t(g), q = fastdiv(n, p.vprdiv);
            e0 = int64_t(n-q*uint32_t(p.vpr))*SV; n = q;
            if (2==p.rankA) i1 = n;
            else { q = fastdiv(n, p.dA[1]); i1 = n-q*p.dA[1].len; n = q;
                if (3==p.rankA) i2 = n; else { q = fastdiv(n, p.dA[2]); i2 = n-q*p.dA[2].len; i3 = q; } }
        }
        doff = e0 + int64_t(i1)*p.dsA[1] + int64_t(i2)*p.dsA[2] + int64_t(i3)*p.dsA[3];
#pragma unroll
        for (int k=0; k<P.nin; ++k) off[k] = 1==p.rankA ? 0 : int64_t(i1)*p.sA[k][1] + int64_t(i2)*p.sA[k][2] + int64_t(i3)*p.sA[k][3];
    };
    for (int64_t g0 = int64_t(blockIdx.x)*per + threadIdx.x; g0<p.nvec; g0 += int64_t(gridDim.x)*per) {
        uint32_t x[UNR][P.nin][XW];
        int64_t doff[UNR];
        if (g0 + int64_t(UNR-1)*KBLOCK<p.nvec) { // whole group: every load is issued before the first use (loads_issued)
#pragma unroll
            for (int u=0; u<UNR; ++u) {
                int64_t e0, off[P.nin];
                locate(g0 + int64_t(u)*KBLOCK, e0, doff[u], off);
#pragma unroll
                for (int k=0; k<P.nin; ++k) fetch_role<ID, GENERAL, LD_STREAM, KINDS>(p, k, x[u][k], off[k], e0);
            }
            loads_issued<ID, UNR>(p, x);
#pragma unroll
            for (int u=0; u<UNR; ++u) { W out[SV]; eval_static<ID>(p, x[u], out); store_vec<W>(static_cast<unsigned char *>(p.dst) + doff[u]*des, out, P.otype); }
            continue;
        }
#pragma unroll
        for (int u=0; u<UNR; ++u) {
            int64_t g = g0 + int64_t(u)*KBLOCK;
            if (g<p.nvec) {
                int64_t e0, off[P.nin];
                W out[SV];
                locate(g, e0, doff[u], off);
#pragma unroll
                for (int k=0; k<P.nin; ++k) fetch_role<ID, GENERAL, LD_STREAM, KINDS>(p, k, x[u][k], off[k], e0);
                eval_static<ID>(p, x[u], out);
                store_vec<W>(static_cast<unsigned char *>(p.dst) + doff[u]*des, out, P.otype);
            }
        }
    }
    if (1==p.rankA && 0==blockIdx.x && 0==threadIdx.x) { // tail elements of a flat traversal, one at a time
        for (int64_t e = p.nvec*SV; e<p.n; ++e) {
            uint32_t x[P.nin][XW];
            W out[SV];
#pragma unroll
            for (int k=0; k<P.nin; ++k) {
                int const es = es_of(P.rtype[k]);
                unsigned char const * base = reinterpret_cast<unsigned char const *>(p.in[k]);
                if (SK_SCALAR==p.kind[k]) splat_bits(x[k], p.in[k], P.rtype[k]);
                else if (SK_PTR==p.kind[k]) {}
                else if (SK_SEQ==p.kind[k]) seq_words(x[k], p.in[k], e*p.sA[k][0], 0, P.rtype[k]);
                else splat_words(x[k], base + (SK_VEC==p.kind[k] ? e : SK_VECREV==p.kind[k] ? -e : 0)*es, P.rtype[k]);
            }
            eval_static<ID>(p, x, out);
            store_elem<W>(static_cast<unsigned char *>(p.dst) + e*des, P.otype, out[0]);
        }
    }
}

// One group of UNR warp tiles of a row for smap_lanes_body.
// FULL: every element of the UNR warp tiles is inside the row (all but the last group of a row). The loads and stores are then
// unpredicated and the four accesses of a role are ONE address plus immediates; with the per-element `live` predicates the compiler
// recomputed a 64-bit address under every predicate: 39 instructions per element on the 5-point stencil, issue bound next to the L1 limit.
template <int ID, uint64_t KINDS, bool FULL> __device__ __forceinline__ void
smap_lanes_group(SParams const & p, int lane, int64_t wt, int64_t wt1, unsigned char const * const * rowp, unsigned char * drow)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    constexpr int des = es_of(P.otype), UNR = SPX<ID>::UNR>2 ? 2 : SPX<ID>::UNR, XW = SPX<ID>::XW;
    uint32_t x[UNR][P.nin][XW];
    int64_t e0[UNR];
#pragma unroll
    for (int u=0; u<UNR; ++u) {
        e0[u] = (FULL || wt+u<wt1) ? (wt+u)*(32*SV) + lane : p.n;   // this lane's first element of the warp tile; the others are 32 apart
        bool live[SV];                                    // which of this lane's elements are inside the row: once for every role
#pragma unroll
        for (int j=0; j<SV; ++j) live[j] = FULL || e0[u] + 32*j<p.n;
#pragma unroll
        for (int k=0; k<P.nin; ++k) {
            int const t = P.rtype[k], es = es_of(t);
            int const kd = KINDS!=SK_RUNTIME ? int((KINDS>>(4*k))&15u) : int(p.kind[k]);
            uint32_t * xv = x[u][k];
            if (SK_SCALAR==kd) splat_bits(xv, p.in[k], t);
            else if (SK_PTR==kd) {}
            else if (SK_SEQ==kd) seq_words(xv, p.in[k], int64_t(reinterpret_cast<intptr_t>(rowp[k])) + e0[u]*p.sA[k][0], 32*p.sA[k][0], t);
            else if (SK_ROW==kd) splat_words(xv, rowp[k], t);
            else { // the four loads at fixed distances from one pointer
                bool const rev = SK_VECREV==kd || SK_VECREVU==kd;
                unsigned char const * const q0 = rowp[k] + (rev ? -e0[u] : e0[u])*es;
                int const step = rev ? -32*es : 32*es;
                if (1==es) xv[0] = 0;
#pragma unroll
                for (int j=0; j<SV; ++j) {
                    unsigned char const * q = q0 + j*step;
                    if (4==es) { uint32_t v = 0; if (live[j]) v = k==P.yrole ? *reinterpret_cast<uint32_t const *>(q) : __ldg(reinterpret_cast<uint32_t const *>(q)); xv[j] = v; }
                    else if (8==es) { uint2 v = make_uint2(0, 0); if (live[j]) v = k==P.yrole ? *reinterpret_cast<uint2 const *>(q) : __ldg(reinterpret_cast<uint2 const *>(q)); xv[2*j] = v.x; xv[2*j+1] = v.y; }
                    else { if (live[j] && *q) xv[0] |= 1u<<(8*j); }
                }
            }
        }
    }
#pragma unroll
    for (int u=0; u<UNR; ++u) {
        if (FULL || e0[u]<p.n) {
            W out[SV];
            eval_static<ID>(p, x[u], out);
            unsigned char * const d0 = drow + e0[u]*des;
#pragma unroll
            for (int j=0; j<SV; ++j) { if (FULL || e0[u] + 32*j<p.n) store_elem<W>(d0 + j*32*des, P.otype, out[j]); }
        }
    }
}

// ---- map, lane-wise: the same traversal for rows that are NOT 16-byte aligned or not a whole number of vectors - slices shifted by
// one element (the user-written stencils of examples/laplace2d.cc:36,48 and bench/bench-stencil1.cc:45, any A(I, J+1)), odd row lengths.
// A warp owns 32*SV consecutive elements of a row and lane l holds elements l, l+32, l+64, l+96 of them, so every load and store is
// ONE coalesced 4- or 8-byte access per lane (128 / 256 contiguous bytes per warp) whatever the alignment; rows are padded to whole
// warp tiles and the lanes past the row end do nothing. (A thread owning 4 consecutive unaligned elements instead reads 16-byte
// strided words: 4x the L1 wavefronts, measured 0.3 of the HBM peak on the 5-point stencil.) Loads allocate in L1: the shifted
// slices of one array share their lines. p.n = row length, p.vpr = vectors per padded row (a multiple of 32).
template <int ID, uint64_t KINDS> __device__ __forceinline__ void
smap_lanes_body(SParams const & p)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    constexpr int des = es_of(P.otype), UNR = SPX<ID>::UNR>2 ? 2 : SPX<ID>::UNR, XW = SPX<ID>::XW, NW = KBLOCK/32, CT = SLANES_CT;
    int const lane = threadIdx.x&31, warp = threadIdx.x>>5;
    // Work units. Traversals over rows take PATCHES of NW rows x CT warp tiles (warp w owns row w of the patch and walks its CT tiles):
    // the rows i-1, i, i+1 that a stencil expression reads are then loaded by the same CTA and the re-reads hit L1 (measured: 72 %
    // L1 hits on the 5-point stencil), and everything that depends on the row only - the index decomposition, one pointer per role -
    // is computed once per CT*128 elements: the first version of this kernel spent 78 instructions per element, most of them on
    // that, and was issue bound (83 % issue slots busy at 0.5 of the HBM peak). Flat traversals: a unit is NW*CT consecutive tiles.
    bool const rows = p.rankA>1;
    int64_t const wpr = p.vpr/32, nrows = rows ? p.nvec/p.vpr : 1;   // warp tiles per (padded) row
    uint32_t const ppb = uint32_t((wpr+CT-1)/CT);                     // patches per band of NW rows
    int64_t const nunits = rows ? ((nrows+NW-1)/NW)*int64_t(ppb) : (wpr+int64_t(NW)*CT-1)/(int64_t(NW)*CT);
    for (int64_t unit = blockIdx.x; unit<nunits; unit += gridDim.x) {
        int64_t row = 0, wt0, wt1;
        if (rows) {
            uint32_t const band = fastdiv(uint32_t(unit), p.vprdiv), patch = uint32_t(unit)-band*ppb; // (vprdiv divides by ppb here)
            row = int64_t(band)*NW + warp;
            if (row>=nrows) continue; // (the kernel has no barrier)
            wt0 = int64_t(patch)*CT;
        } else wt0 = (unit*NW + warp)*CT;
        wt1 = wt0+CT<wpr ? wt0+CT : wpr;
        if (wt0>=wt1) continue;
        // once per row: where the row starts in the destination and in every role
        uint32_t i1 = 0, i2 = 0, i3 = 0;
        if (rows) {
            uint32_t n = uint32_t(row), q;
            if (2==p.rankA) i1 = n;
            else { q = fastdiv(n, p.dA[1]); i1 = n-q*p.dA[1].len; n = q;
                if (3==p.rankA) i2 = n; else { q = fastdiv(n, p.dA[2]); i2 = n-q*p.dA[2].len; i3 = q; } }
        }
        unsigned char * const drow = static_cast<unsigned char *>(p.dst) + (int64_t(i1)*p.dsA[1] + int64_t(i2)*p.dsA[2] + int64_t(i3)*p.dsA[3])*des;
        unsigned char const * rowp[P.nin]; // MEM roles: address of the row's element 0. SEQ: the row's element offset
#pragma unroll
        for (int k=0; k<P.nin; ++k) {
            int const kd = KINDS!=SK_RUNTIME ? int((KINDS>>(4*k))&15u) : int(p.kind[k]);
            int64_t const off = int64_t(i1)*p.sA[k][1] + int64_t(i2)*p.sA[k][2] + int64_t(i3)*p.sA[k][3];
            rowp[k] = SK_SEQ==kd ? reinterpret_cast<unsigned char const *>(intptr_t(off)) : reinterpret_cast<unsigned char const *>(p.in[k]) + off*es_of(P.rtype[k]);
        }
        for (int64_t wt = wt0; wt<wt1; wt += UNR) {
            if (wt+UNR<=wt1 && (wt+UNR)*int64_t(32*SV)<=p.n) smap_lanes_group<ID, KINDS, true>(p, lane, wt, wt1, rowp, drow);
            else smap_lanes_group<ID, KINDS, false>(p, lane, wt, wt1, rowp, drow);
        }
    }
}

template <class W> __device__ __forceinline__ W shfl_xor_w(W v, int o);
template <> __device__ __forceinline__ uint32_t shfl_xor_w<uint32_t>(uint32_t v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ uint64_t shfl_xor_w<uint64_t>(uint64_t v, int o) { return __shfl_xor_sync(0xffffffffu, (unsigned long long)v, o); }

template <class W, int FOLD, int TYPE> __device__ __forceinline__ W
block_comb(W r, W * red)
{
    for (int o=16; o>0; o>>=1) { W other = shfl_xor_w<W>(r, o); r = combine<W>(FOLD, TYPE, r, other); }
    int const lane = threadIdx.x&31, warp = threadIdx.x>>5;
    if (0==lane) red[warp] = r;
    __syncthreads();
    if (0==threadIdx.x) {
        r = red[0];
        for (int w=1; w<KBLOCK/32; ++w) r = combine<W>(FOLD, TYPE, r, red[w]);
    }
    return r;
}

template <class W> __device__ __forceinline__ void
emit_partial(SParams const & p, FinishParams const & f, int64_t chunk, int64_t n, W v)
{
    if (p.nchunks>1) {
        if (p.wide_partials) static_cast<uint64_t *>(p.partial)[chunk*p.nk+n] = uint64_t(v);
        else static_cast<uint32_t *>(p.partial)[chunk*p.nk+n] = uint32_t(v);
    } else {
        if (p.wide_partials) finish_apply<uint64_t>(f, n, uint64_t(v)); else finish_apply<uint32_t>(f, n, uint32_t(v));
    }
}

// ---- folds. KINDS: the operand kind of every role, 4 bits each (role k at bits 4k..4k+3), or SK_RUNTIME = read p.kind[] at run time.
// RMASK: bit k set = role k does not move along the kept rows (the vector of gemv, ra/ra.hh:418: every row re-reads it) and is loaded
// through L1; everything else streams past L1. Both are compile-time so that, after unrolling over the roles, every load is ONE
// instruction of a fixed form: measured (round 1) on the gevm kernel with run-time kinds, the per-row switch made a 4096-instruction
// body with jump tables that ran at 0.36 of the HBM peak.
template <uint64_t KINDS> __device__ __forceinline__ int
role_kind(SParams const & p, int k) { if constexpr (KINDS==SK_RUNTIME) return int(p.kind[k]); else return int((KINDS>>(4*k))&15u); }

__device__ __forceinline__ void
load_words_form(int form, uint32_t * x, unsigned char const * p, int t)
{
    if (LD_STREAM==form) load_words<LD_STREAM>(x, p, t); else if (LD_REUSED==form) load_words<LD_REUSED>(x, p, t); else load_words<LD_RW>(x, p, t);
}

// role k of vector g of the row at rowbase -> raw words (SEQ roles carry the row's element offset in rowbase: no memory)
template <int ID, uint64_t KINDS> __device__ __forceinline__ void
fetch_fold_role(SParams const & p, int k, int form, uint32_t * x, unsigned char const * rowbase, int64_t g)
{
    constexpr SProg P = SPT<ID>::P;
    int const t = P.rtype[k], es = es_of(t), kd = role_kind<KINDS>(p, k);
    if (SK_SCALAR==kd) splat_bits(x, p.in[k], t);
    else if (SK_PTR==kd) {}
    else if (SK_SEQ==kd) seq_words(x, p.in[k], int64_t(reinterpret_cast<intptr_t>(rowbase)) + g*SV*p.sA[k][0], p.sA[k][0], t);
    else if (SK_ROW==kd) splat_words(x, rowbase, t);
    else load_words_form(form, x, rowbase + g*SV*es, t);
}

// kept row kb -> per role: where the row starts (bytes; SEQ: the element offset itself)
template <int ID, uint64_t KINDS> __device__ __forceinline__ void
fold_row_bases(SParams const & p, uint32_t kb, unsigned char const ** rowbase)
{
    constexpr SProg P = SPT<ID>::P;
    uint32_t n = kb, q, ib[KGR] = {0, 0, 0, 0};
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
        if (SK_SEQ==role_kind<KINDS>(p, k)) rowbase[k] = reinterpret_cast<unsigned char const *>(intptr_t(off)); // no memory: carry the offset itself
        else rowbase[k] = reinterpret_cast<unsigned char const *>(p.in[k]) + off*int64_t(es_of(P.rtype[k]));
    }
}

// ---- fold-inner, flat rows: CTA = (row kb, chunk c); row length n (multiple of SV unless there is a single row).
template <int ID, int FOLD, int RMASK, uint64_t KINDS> __device__ __forceinline__ void
sfold_inner_body(SParams const & p, FinishParams const & f)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    constexpr int AT = P.otype, UNR = SPX<ID>::UNR, XW = SPX<ID>::XW;
    __shared__ W red[KBLOCK/32];
    uint32_t const kb = blockIdx.x/uint32_t(p.nchunks), c = blockIdx.x-kb*uint32_t(p.nchunks);
    unsigned char const * rowbase[P.nin];
    fold_row_bases<ID, KINDS>(p, kb, rowbase);
    int64_t const nfull = p.n/SV;
    int64_t const start = int64_t(c)*p.chunk_len;
    int64_t const end = start+p.chunk_len<nfull ? start+p.chunk_len : nfull;
    W acc[SV];
#pragma unroll
    for (int j=0; j<SV; ++j) acc[j] = W(p.identity);
    int64_t const per = int64_t(KBLOCK)*UNR;
    for (int64_t g0 = start + threadIdx.x; g0<end; g0 += per) {
        uint32_t x[UNR][P.nin][XW];
        // Whole groups of UNR vectors take a path WITHOUT per-vector predicates: with them the compiler pairs every load with its use
        // (load, multiply, add, next load ...), two loads in flight per thread instead of 2*UNR: measured on dot, 6.3 instead of 7.1 TB/s.
        if (g0 + int64_t(UNR-1)*KBLOCK<end) {
#pragma unroll
            for (int u=0; u<UNR; ++u) {
#pragma unroll
                for (int k=0; k<P.nin; ++k) fetch_fold_role<ID, KINDS>(p, k, (RMASK>>k)&1 ? LD_REUSED : LD_STREAM, x[u][k], rowbase[k], g0 + int64_t(u)*KBLOCK);
            }
            loads_issued<ID, UNR>(p, x);
#pragma unroll
            for (int u=0; u<UNR; ++u) {
                W out[SV];
                eval_static<ID>(p, x[u], out);
#pragma unroll
                for (int j=0; j<SV; ++j) acc[j] = combine<W>(FOLD, AT, acc[j], out[j]);
            }
        } else {
#pragma unroll
            for (int u=0; u<UNR; ++u) {
                int64_t g = g0 + int64_t(u)*KBLOCK;
                if (g<end) {
                    W out[SV];
#pragma unroll
                    for (int k=0; k<P.nin; ++k) fetch_fold_role<ID, KINDS>(p, k, (RMASK>>k)&1 ? LD_REUSED : LD_STREAM, x[u][k], rowbase[k], g);
                    eval_static<ID>(p, x[u], out);
#pragma unroll
                    for (int j=0; j<SV; ++j) acc[j] = combine<W>(FOLD, AT, acc[j], out[j]);
                }
            }
        }
    }
    W r = acc[0];
#pragma unroll
    for (int j=1; j<SV; ++j) r = combine<W>(FOLD, AT, r, acc[j]);
    if (c+1==uint32_t(p.nchunks) && 0==threadIdx.x) { // tail elements of the row (single flat row only)
        for (int64_t e = nfull*SV; e<p.n; ++e) {
            uint32_t x[P.nin][XW];
            W out[SV];
#pragma unroll
            for (int k=0; k<P.nin; ++k) {
                int const kd = role_kind<KINDS>(p, k);
                if (SK_SCALAR==kd) splat_bits(x[k], p.in[k], P.rtype[k]);
                else if (SK_PTR==kd) {}
                else if (SK_SEQ==kd) seq_words(x[k], p.in[k], int64_t(reinterpret_cast<intptr_t>(rowbase[k])) + e*p.sA[k][0], 0, P.rtype[k]);
                else splat_words(x[k], rowbase[k] + (SK_ROW==kd ? 0 : e)*es_of(P.rtype[k]), P.rtype[k]);
            }
            eval_static<ID>(p, x, out);
            r = combine<W>(FOLD, AT, r, out[0]);
        }
    }
    r = block_comb<W, FOLD, AT>(r, red);
    if (0==threadIdx.x) emit_partial<W>(p, f, c, kb, r);
}

// ---- fold-inner, one WARP per kept row (gemv ra/ra.hh:414-425, bench-sum-cols.cc:62-65 when the rows are many and of moderate
// length): no CTA barrier and no shared memory, a row is 32 lanes x UNR vectors in flight until it ends, then five shuffles. With a
// CTA per 64 KiB row (round 1) a CTA lived for four iterations and most of its time went into starting and draining: 0.56 of the HBM peak.
// Rows are whole vectors (n % SV == 0), one chunk (the result goes straight to the destination).
template <int ID, int FOLD, int RMASK, uint64_t KINDS> __device__ __forceinline__ void
sfold_inner_rows_body(SParams const & p, FinishParams const & f)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    constexpr int AT = P.otype, UNR = SPX<ID>::UNR, XW = SPX<ID>::XW;
    int const lane = threadIdx.x&31;
    int64_t const row = int64_t(blockIdx.x)*(KBLOCK/32) + (threadIdx.x>>5);
    if (row>=p.nk) return; // whole warp
    unsigned char const * rowbase[P.nin];
    fold_row_bases<ID, KINDS>(p, uint32_t(row), rowbase);
    int64_t const nfull = p.n/SV;
    W acc[SV];
#pragma unroll
    for (int j=0; j<SV; ++j) acc[j] = W(p.identity);
    for (int64_t g0 = lane; g0<nfull; g0 += 32*UNR) {
        uint32_t x[UNR][P.nin][XW];
        if (g0 + (UNR-1)*32<nfull) { // whole group: every load issued before the first use (see sfold_inner_body)
#pragma unroll
            for (int u=0; u<UNR; ++u) {
#pragma unroll
                for (int k=0; k<P.nin; ++k) fetch_fold_role<ID, KINDS>(p, k, (RMASK>>k)&1 ? LD_REUSED : LD_STREAM, x[u][k], rowbase[k], g0 + u*32);
            }
            loads_issued<ID, UNR>(p, x);
#pragma unroll
            for (int u=0; u<UNR; ++u) {
                W out[SV];
                eval_static<ID>(p, x[u], out);
#pragma unroll
                for (int j=0; j<SV; ++j) acc[j] = combine<W>(FOLD, AT, acc[j], out[j]);
            }
        } else {
#pragma unroll
            for (int u=0; u<UNR; ++u) {
                int64_t g = g0 + u*32;
                if (g<nfull) {
                    W out[SV];
#pragma unroll
                    for (int k=0; k<P.nin; ++k) fetch_fold_role<ID, KINDS>(p, k, (RMASK>>k)&1 ? LD_REUSED : LD_STREAM, x[u][k], rowbase[k], g);
                    eval_static<ID>(p, x[u], out);
#pragma unroll
                    for (int j=0; j<SV; ++j) acc[j] = combine<W>(FOLD, AT, acc[j], out[j]);
                }
            }
        }
    }
    W r = acc[0];
#pragma unroll
    for (int j=1; j<SV; ++j) r = combine<W>(FOLD, AT, r, acc[j]);
    for (int o=16; o>0; o>>=1) { W other = shfl_xor_w<W>(r, o); r = combine<W>(FOLD, AT, r, other); }
    if (0==lane) emit_partial<W>(p, f, 0, row, r);
}

// ---- fold-outer, flat (gevm ra/ra.hh:427-438: c(j) += a(i)*b(i,j), a is one value per folded row, SK_ROW; bench-sum-rows.cc:58-81).
// A CTA owns a tile of 32*KV kept vectors (a lane owns KV of them: lanes are kept elements, 32 bytes of the destination row per thread)
// and chunk blockIdx.y of the folded rows; its 8 warps take the rows of the chunk in turns, UNR rows in flight each, addresses advancing
// by pointer bumps, and combine their accumulators through shared memory in warp order (deterministic). Splitting the rows inside
// the CTA keeps the number of chunks - partials written, read again and combined by ply_finish - 8x lower for the same number of
// threads: with a thread per kept vector and 150-300 chunks (round 1) the partials pass cost as much as the sweep on 8192^2.
template <int ID> struct SPO
{
    static constexpr int KV = SPT<ID>::wide ? 1 : 2;
    static constexpr int UNR = SPX<ID>::UNR; // folded rows in flight per warp (x KV vectors each)
    static constexpr int TILE = 32*KV;       // kept vectors per CTA
};

template <int ID, int FOLD, uint64_t KINDS> __device__ __forceinline__ void
sfold_outer_body(SParams const & p, FinishParams const & f)
{
    using W = typename SPT<ID>::W;
    constexpr SProg P = SPT<ID>::P;
    constexpr int AT = P.otype, XW = SPX<ID>::XW, KV = SPO<ID>::KV, UNR = SPO<ID>::UNR, NW = KBLOCK/32;
    __shared__ W sacc[NW][KV*SV][32];
    int const lane = threadIdx.x&31, warp = threadIdx.x>>5;
    int64_t const nvec = p.n/SV;
    int64_t const g0 = int64_t(blockIdx.x)*SPO<ID>::TILE + lane;
    bool ok[KV];
#pragma unroll
    for (int v=0; v<KV; ++v) ok[v] = g0 + v*32<nvec;
    uint32_t const c = blockIdx.y;
    int64_t const r0 = int64_t(c)*p.chunk_len;
    int64_t const r1 = r0+p.chunk_len<p.nB ? r0+p.chunk_len : p.nB;
    // per role: the address (SEQ: element offset) of this warp's first row, and what one row adds to it
    unsigned char const * rowp[P.nin];
    int64_t rowstep[P.nin];
#pragma unroll
    for (int k=0; k<P.nin; ++k) {
        int const kd = role_kind<KINDS>(p, k);
        int64_t const scale = SK_SEQ==kd ? 1 : es_of(P.rtype[k]);
        rowstep[k] = p.sB[k][0]*scale;
        rowp[k] = (SK_SEQ==kd ? static_cast<unsigned char const *>(nullptr) : reinterpret_cast<unsigned char const *>(p.in[k])) + (r0 + warp*UNR)*rowstep[k];
    }
    W acc[KV][SV];
#pragma unroll
    for (int v=0; v<KV; ++v)
#pragma unroll
        for (int j=0; j<SV; ++j) acc[v][j] = W(p.identity);
    bool all_ok = true;
#pragma unroll
    for (int v=0; v<KV; ++v) all_ok = all_ok && ok[v];
    for (int64_t r = r0 + warp*UNR; r<r1; r += NW*UNR) {
        uint32_t x[UNR][KV][P.nin][XW];
        if (all_ok && r+UNR<=r1) { // whole group of rows, whole tile: every load issued before the first use (see sfold_inner_body)
#pragma unroll
            for (int u=0; u<UNR; ++u) {
#pragma unroll
                for (int k=0; k<P.nin; ++k) {
                    int const kd = role_kind<KINDS>(p, k);
                    bool const per_row = SK_SCALAR==kd || SK_ROW==kd || SK_PTR==kd;
#pragma unroll
                    for (int v=0; v<KV; ++v) {
                        if (per_row && v>0) {
#pragma unroll
                            for (int q=0; q<XW; ++q) x[u][v][k][q] = x[u][0][k][q];
                        } else fetch_fold_role<ID, KINDS>(p, k, LD_STREAM, x[u][v][k], rowp[k] + u*rowstep[k], g0 + v*32);
                    }
                }
            }
            loads_issued<ID, UNR*KV>(p, reinterpret_cast<uint32_t (*)[P.nin][XW]>(x));
#pragma unroll
            for (int u=0; u<UNR; ++u) {
#pragma unroll
                for (int v=0; v<KV; ++v) {
                    W out[SV];
                    eval_static<ID>(p, x[u][v], out);
#pragma unroll
                    for (int j=0; j<SV; ++j) acc[v][j] = combine<W>(FOLD, AT, acc[v][j], out[j]);
                }
            }
#pragma unroll
            for (int k=0; k<P.nin; ++k) rowp[k] += NW*UNR*rowstep[k];
            continue;
        }
#pragma unroll
        for (int u=0; u<UNR; ++u) {
            if (r+u<r1) {
#pragma unroll
                for (int k=0; k<P.nin; ++k) {
                    int const kd = role_kind<KINDS>(p, k);
                    bool const per_row = SK_SCALAR==kd || SK_ROW==kd || SK_PTR==kd; // the same for every kept vector: fetch once
#pragma unroll
                    for (int v=0; v<KV; ++v) {
                        if (per_row && v>0) {
#pragma unroll
                            for (int q=0; q<XW; ++q) x[u][v][k][q] = x[u][0][k][q];
                        } else if (ok[v] || per_row) fetch_fold_role<ID, KINDS>(p, k, LD_STREAM, x[u][v][k], rowp[k] + u*rowstep[k], g0 + v*32);
                    }
                }
            }
        }
#pragma unroll
        for (int u=0; u<UNR; ++u) {
            if (r+u<r1) {
#pragma unroll
                for (int v=0; v<KV; ++v) {
                    if (ok[v]) {
                        W out[SV];
                        eval_static<ID>(p, x[u][v], out);
#pragma unroll
                        for (int j=0; j<SV; ++j) acc[v][j] = combine<W>(FOLD, AT, acc[v][j], out[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int k=0; k<P.nin; ++k) rowp[k] += NW*UNR*rowstep[k];
    }
#pragma unroll
    for (int v=0; v<KV; ++v)
#pragma unroll
        for (int j=0; j<SV; ++j) sacc[warp][v*SV+j][lane] = acc[v][j];
    __syncthreads();
    // thread t combines kept element (e = t / 32, lane t % 32) over the warps, in warp order
    if (int(threadIdx.x)<KV*SV*32) {
        int const e = threadIdx.x>>5, v = e/SV, j = e-v*SV;
        int64_t const g = g0 + v*32;
        if (g<nvec) {
            W r = sacc[0][e][lane];
#pragma unroll
            for (int w=1; w<NW; ++w) r = combine<W>(FOLD, AT, r, sacc[w][e][lane]);
            emit_partial<W>(p, f, c, g*SV+j, r);
        }
    }
}

} // namespace racu
