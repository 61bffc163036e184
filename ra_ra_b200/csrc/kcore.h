// kcore.h - kernel-side data layout and the per-element machinery of the fused ply kernels:
// fast index decomposition, operand offsets, the typed op table and the postfix interpreter.
//
// Everything here is RACU_HD so that (a) the sm_100a kernels (ply_kernels.cu) and (b) the planner's CPU
// simulator (sim/plan_sim.cc, built only for tests, never linked into libracu.so) run the SAME code.
//
// Execution model. A thread owns "vectors": V consecutive elements along the innermost canonical dimension
// (V = 4 with 32-bit value slots when every type in the program is <= 4 bytes, V = 2 with 64-bit slots otherwise),
// so a full vector of a 4- or 8-byte operand is one 16-byte access. Operand elements are first STAGED into a
// per-thread 16-byte slot (shared memory on the device; cp.async, so all operand loads of a thread are in flight
// at once), then the postfix program runs over the V lanes with the top of stack in registers and the rest of the
// value stack in per-thread shared memory rows (level `lvl`, fixed per instruction by the planner).
#pragma once
#include <stdint.h>
#include <math.h>
#include <type_traits>
#include "../../include/racu.h"

#if defined(__CUDACC__)
#define RACU_HD __host__ __device__ __forceinline__
#define RACU_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define RACU_HD inline
#define RACU_HD_NOINLINE inline
#endif

#if !defined(__CUDACC__) && !defined(__VECTOR_TYPES_H__)
// host stand-ins for the CUDA vector types used by the shared kernel bodies (CPU plan simulator only)
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
#endif

namespace racu {

constexpr int KBLOCK = 256;      // threads per CTA of the ply kernels
constexpr int KMAXR = RACU_MAX_RANK;
constexpr int KGR = 4;           // axes per group (A, B) after collapsing; longer traversals are peeled on the host
constexpr int KSLOT = 16;        // bytes of one staged vector / one stack row per thread

enum KMode { K_MAP = 0, K_FOLD_INNER = 1, K_FOLD_OUTER = 2 };
enum KStage { S_NONE = 0 /* scalar immediate */, S_VEC = 1, S_GATHER = 2, S_BCAST = 3, S_SEQ = 4 };

struct KDim { uint32_t len, mul, shift, pad; }; // n / len == (umulhi(n, mul) + n) >> shift for n < 2^31

struct KOpnd
{
    uint64_t base;            // MEM: address of element [0..0] after canonicalisation. SEQ/SCALAR: value bits (i64 or f64)
    int64_t sA[KGR];          // element strides over group A dims (sA[0]: innermost, the vectorised one)
    int64_t sB[KGR];          // element strides over group B dims
    uint8_t kind, dtype, stage, esize;
    uint8_t slot;             // staging slot index (S_NONE: unused)
    uint8_t pad[3];
};

struct KIns { uint8_t op, type, arg, aux, lvl, pad[3]; };

struct KParams
{
    int32_t mode;             // KMode
    int32_t wide;             // 0: 32-bit slots, V=4.  1: 64-bit slots, V=2
    int32_t rankA, rankB;
    int32_t nopnd, nins, nstage, depth;
    int32_t U;                // vectors staged per thread per iteration
    int32_t dst;              // MAP: operand index of the destination. folds: operand index of dst, or -1
    int32_t fold_op;          // racu_assign of a fold
    int32_t acc_type;         // racu_dtype of the accumulator
    int32_t overwrite;        // folds: result replaces dst instead of combining with its old value (racu_reduce)
    int32_t nchunks;
    int32_t scatter;          // MAP with a RACU_PTR destination: the program leaves [I64 offset (stack level 0), value]; fold_op = the assign op,
    int32_t pad0;             // acc_type = the value's type (common type of destination and right hand side)
    uint64_t identity;        // accumulator start value, slot bits
    int64_t len0;             // length of A dim 0 in elements
    int64_t vpr;              // vectors per row = ceil(len0 / V)
    int64_t nvecA;            // vpr * prod(lenA[1..])
    int64_t nB;               // prod(lenB)
    int64_t chunk_len;        // FOLD_INNER: A-vectors per chunk. FOLD_OUTER: B indices per chunk
    int64_t nk;               // folds: number of kept elements
    void * partial;           // folds: [nchunks][nk] accumulator slots
    KDim vprdiv;
    KDim dA[KGR];             // dA[0] unused (see vpr)
    KDim dB[KGR];
    KOpnd opnd[RACU_MAX_OPND];
    KIns ins[RACU_MAX_INS];
};

struct FinishParams
{
    int32_t wide, fold_op, acc_type, dst_dtype, overwrite, rank, nchunks, per_warp;
    uint64_t identity;
    int64_t nk;
    void * partial;
    void * dst;
    KDim d[KGR];
    int64_t len0;         // length of kept dim 0 (not limited to 2^31 when rank==1)
    int64_t stride[KGR];
};

// ---------------- index math ----------------

RACU_HD uint32_t umulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return uint32_t((uint64_t(a)*uint64_t(b))>>32);
#endif
}
RACU_HD uint32_t fastdiv(uint32_t n, KDim const & d) { return (umulhi32(n, d.mul)+n)>>d.shift; }

struct KIdx { int64_t e0; uint32_t i[KGR]; }; // group A: e0 = element index along dim 0, i[1..] the others. group B: i[0..]

// Ranks are uniform across the grid: nested uniform branches cost one compare each, where an unrolled predicated
// loop over the maximum rank costs the full chain for every operand of every vector.

// vector id g (< nvecA) -> position in group A
RACU_HD void decompA(KParams const & p, int64_t g, KIdx & x)
{
    int const V = p.wide ? 2 : 4;
    if (1==p.rankA) { x.e0 = g*V; return; }
    uint32_t n = uint32_t(g);
    uint32_t q = fastdiv(n, p.vprdiv);
    x.e0 = int64_t(n-q*uint32_t(p.vpr))*V;
    n = q;
    if (2==p.rankA) { x.i[1] = n; return; }
    q = fastdiv(n, p.dA[1]); x.i[1] = n-q*p.dA[1].len; n = q;
    if (3==p.rankA) { x.i[2] = n; return; }
    q = fastdiv(n, p.dA[2]); x.i[2] = n-q*p.dA[2].len; x.i[3] = q;
}

// linear index b (< nB) -> position in group B
RACU_HD void decompB(KParams const & p, int64_t b, KIdx & x)
{
    uint32_t n = uint32_t(b), q;
    if (p.rankB<=1) { x.i[0] = n; return; }
    q = fastdiv(n, p.dB[0]); x.i[0] = n-q*p.dB[0].len; n = q;
    if (2==p.rankB) { x.i[1] = n; return; }
    q = fastdiv(n, p.dB[1]); x.i[1] = n-q*p.dB[1].len; n = q;
    if (3==p.rankB) { x.i[2] = n; return; }
    q = fastdiv(n, p.dB[2]); x.i[2] = n-q*p.dB[2].len; x.i[3] = q;
}

RACU_HD int64_t offsetA(KParams const & p, KOpnd const & o, KIdx const & a)
{
    int64_t off = a.e0*o.sA[0];
    if (p.rankA>1) {
        off += int64_t(a.i[1])*o.sA[1];
        if (p.rankA>2) {
            off += int64_t(a.i[2])*o.sA[2];
            if (p.rankA>3) off += int64_t(a.i[3])*o.sA[3];
        }
    }
    return off;
}
RACU_HD int64_t offsetB(KParams const & p, KOpnd const & o, KIdx const & b)
{
    int64_t off = 0;
    if (p.rankB>0) {
        off = int64_t(b.i[0])*o.sB[0];
        if (p.rankB>1) {
            off += int64_t(b.i[1])*o.sB[1];
            if (p.rankB>2) {
                off += int64_t(b.i[2])*o.sB[2];
                if (p.rankB>3) off += int64_t(b.i[3])*o.sB[3];
            }
        }
    }
    return off;
}

// ---------------- value slots ----------------
// A slot holds one value as raw bits: BOOL 0/1, I32 / F32 in the low 32 bits, I64 / F64 in 64 bits.

template <class W> struct Lanes { static constexpr int V = sizeof(W)==4 ? 4 : 2; };

RACU_HD float bits_f32(uint32_t a)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(a);
#else
    union { uint32_t u; float f; } c; c.u = a; return c.f;
#endif
}
RACU_HD uint32_t f32_bits(float a)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(a);
#else
    union { uint32_t u; float f; } c; c.f = a; return c.u;
#endif
}
RACU_HD double bits_f64(uint64_t a)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)a);
#else
    union { uint64_t u; double f; } c; c.u = a; return c.f;
#endif
}
RACU_HD uint64_t f64_bits(double a)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(a);
#else
    union { uint64_t u; double f; } c; c.f = a; return c.u;
#endif
}

template <class T> struct Ty;
template <> struct Ty<bool> { template <class W> static RACU_HD bool get(W a) { return (uint32_t(a)&0xffu)!=0; } template <class W> static RACU_HD W put(bool x) { return W(x ? 1 : 0); } };
template <> struct Ty<int32_t> { template <class W> static RACU_HD int32_t get(W a) { return int32_t(uint32_t(a)); } template <class W> static RACU_HD W put(int32_t x) { return W(uint32_t(x)); } };
template <> struct Ty<float> { template <class W> static RACU_HD float get(W a) { return bits_f32(uint32_t(a)); } template <class W> static RACU_HD W put(float x) { return W(f32_bits(x)); } };
template <> struct Ty<int64_t> { template <class W> static RACU_HD int64_t get(W a) { return int64_t(uint64_t(a)); } template <class W> static RACU_HD W put(int64_t x) { return W(uint64_t(x)); } };
template <> struct Ty<double> { template <class W> static RACU_HD double get(W a) { return bits_f64(uint64_t(a)); } template <class W> static RACU_HD W put(double x) { return W(f64_bits(x)); } };

// Transcendentals and other long sequences are kept out of line: they would otherwise be unrolled V times per case.
RACU_HD_NOINLINE float slow_un_f32(int op, float a)
{
    switch (op) {
    case RACU_OP_EXP: return expf(a); case RACU_OP_LOG: return logf(a); case RACU_OP_SIN: return sinf(a); case RACU_OP_COS: return cosf(a);
    case RACU_OP_TAN: return tanf(a); case RACU_OP_SINH: return sinhf(a); case RACU_OP_COSH: return coshf(a); case RACU_OP_TANH: return tanhf(a);
    case RACU_OP_ASIN: return asinf(a); case RACU_OP_ACOS: return acosf(a); case RACU_OP_ATAN: return atanf(a);
    case RACU_OP_EXPM1: return expm1f(a); case RACU_OP_LOG1P: return log1pf(a); default: return log10f(a);
    }
}
RACU_HD_NOINLINE double slow_un_f64(int op, double a)
{
    switch (op) {
    case RACU_OP_EXP: return exp(a); case RACU_OP_LOG: return log(a); case RACU_OP_SIN: return sin(a); case RACU_OP_COS: return cos(a);
    case RACU_OP_TAN: return tan(a); case RACU_OP_SINH: return sinh(a); case RACU_OP_COSH: return cosh(a); case RACU_OP_TANH: return tanh(a);
    case RACU_OP_ASIN: return asin(a); case RACU_OP_ACOS: return acos(a); case RACU_OP_ATAN: return atan(a);
    case RACU_OP_EXPM1: return expm1(a); case RACU_OP_LOG1P: return log1p(a); default: return log10(a);
    }
}
// isfinite / isnan / isinf on the bit patterns (the same answer on the host simulator and on the device, whatever the math flags)
RACU_HD bool classify_f32(int op, float a)
{
    uint32_t const m = f32_bits(a) & 0x7fffffffu;
    return op==RACU_OP_ISFINITE ? m<0x7f800000u : op==RACU_OP_ISNAN ? m>0x7f800000u : m==0x7f800000u;
}
RACU_HD bool classify_f64(int op, double a)
{
    uint64_t const m = f64_bits(a) & 0x7fffffffffffffffull;
    return op==RACU_OP_ISFINITE ? m<0x7ff0000000000000ull : op==RACU_OP_ISNAN ? m>0x7ff0000000000000ull : m==0x7ff0000000000000ull;
}
// std::lerp as libstdc++ computes it (<cmath> __lerp): exact at t = 0 and t = 1, monotonic, no contraction (the library is built -fmad=false)
template <class T> RACU_HD T t_lerp(T a, T b, T t)
{
    if ((a<=T(0) && b>=T(0)) || (a>=T(0) && b<=T(0))) return t*b + (T(1)-t)*a;
    if (t==T(1)) return b;
    T const x = a + t*(b-a);
    return (t>T(1))==(b>a) ? (b<x ? x : b) : (b>x ? x : b);
}
RACU_HD_NOINLINE float slow_bin_f32(int op, float a, float b)
{
    switch (op) { case RACU_OP_POW: return powf(a, b); case RACU_OP_ATAN2: return atan2f(a, b); default: return fmodf(a, b); }
}
RACU_HD_NOINLINE double slow_bin_f64(int op, double a, double b)
{
    switch (op) { case RACU_OP_POW: return pow(a, b); case RACU_OP_ATAN2: return atan2(a, b); case RACU_OP_MOD: return fmod(a, b); default: return a/b; }
}
RACU_HD_NOINLINE int64_t slow_div_i64(int op, int64_t a, int64_t b) { return 0==b ? 0 : (op==RACU_OP_DIV ? a/b : a%b); }

template <class T> RACU_HD T t_abs(T a) { return a<T(0) ? T(-a) : a; }
template <> RACU_HD float t_abs<float>(float a) { return fabsf(a); }
template <> RACU_HD double t_abs<double>(double a) { return fabs(a); }

// unary, one lane. T is the operand type.
template <class W, class T> RACU_HD W un1(int op, W a)
{
    T x = Ty<T>::template get<W>(a);
    if (op==RACU_OP_NOT) return Ty<bool>::template put<W>(!x);
    if constexpr (std::is_same_v<T, float>) { if (op>=RACU_OP_ISFINITE && op<=RACU_OP_ISINF) return Ty<bool>::template put<W>(classify_f32(op, x)); }
    if constexpr (std::is_same_v<T, double>) { if (op>=RACU_OP_ISFINITE && op<=RACU_OP_ISINF) return Ty<bool>::template put<W>(classify_f64(op, x)); }
    if constexpr (!std::is_same_v<T, bool>) {
        switch (op) {
        case RACU_OP_NEG: return Ty<T>::template put<W>(T(-x));
        case RACU_OP_ABS: return Ty<T>::template put<W>(t_abs<T>(x));
        case RACU_OP_SQR: return Ty<T>::template put<W>(T(x*x));
        }
        if constexpr (std::is_same_v<T, float>) {
            if (op==RACU_OP_SQRT) return Ty<T>::template put<W>(sqrtf(x));
            return Ty<T>::template put<W>(slow_un_f32(op, x));
        } else if constexpr (std::is_same_v<T, double>) {
            if (op==RACU_OP_SQRT) return Ty<T>::template put<W>(sqrt(x));
            return Ty<T>::template put<W>(slow_un_f64(op, x));
        }
    }
    return a;
}

// binary, one lane.
template <class W, class T> RACU_HD W bin1(int op, W a, W b)
{
    T x = Ty<T>::template get<W>(a), y = Ty<T>::template get<W>(b);
    switch (op) {
    case RACU_OP_EQ: return Ty<bool>::template put<W>(x==y);
    case RACU_OP_NE: return Ty<bool>::template put<W>(x!=y);
    case RACU_OP_LT: return Ty<bool>::template put<W>(x<y);
    case RACU_OP_LE: return Ty<bool>::template put<W>(x<=y);
    case RACU_OP_GT: return Ty<bool>::template put<W>(x>y);
    case RACU_OP_GE: return Ty<bool>::template put<W>(x>=y);
    }
    if constexpr (std::is_same_v<T, bool>) {
        switch (op) {
        case RACU_OP_BAND: return Ty<bool>::template put<W>(x & y);
        case RACU_OP_BOR: return Ty<bool>::template put<W>(x | y);
        case RACU_OP_BXOR: return Ty<bool>::template put<W>(x ^ y);
        }
    } else {
        switch (op) {
        case RACU_OP_ADD: return Ty<T>::template put<W>(T(x+y));
        case RACU_OP_SUB: return Ty<T>::template put<W>(T(x-y));
        case RACU_OP_MUL: return Ty<T>::template put<W>(T(x*y));
        case RACU_OP_MIN: return Ty<T>::template put<W>(y<x ? y : x);
        case RACU_OP_MAX: return Ty<T>::template put<W>(x<y ? y : x);
        }
        if constexpr (std::is_same_v<T, int32_t>) {
            switch (op) {
            case RACU_OP_DIV: return Ty<T>::template put<W>(0==y ? 0 : (y==-1 ? T(-x) : x/y));
            case RACU_OP_MOD: return Ty<T>::template put<W>((0==y || y==-1) ? 0 : x%y);
            case RACU_OP_BAND: return Ty<T>::template put<W>(x & y);
            case RACU_OP_BOR: return Ty<T>::template put<W>(x | y);
            case RACU_OP_BXOR: return Ty<T>::template put<W>(x ^ y);
            }
        } else if constexpr (std::is_same_v<T, int64_t>) {
            switch (op) {
            case RACU_OP_DIV: case RACU_OP_MOD: return Ty<T>::template put<W>(slow_div_i64(op, x, y));
            case RACU_OP_BAND: return Ty<T>::template put<W>(x & y);
            case RACU_OP_BOR: return Ty<T>::template put<W>(x | y);
            case RACU_OP_BXOR: return Ty<T>::template put<W>(x ^ y);
            }
        } else if constexpr (std::is_same_v<T, float>) {
            if (op==RACU_OP_DIV) return Ty<T>::template put<W>(x/y);
            return Ty<T>::template put<W>(slow_bin_f32(op, x, y));
        } else {
            return Ty<T>::template put<W>(slow_bin_f64(op, x, y));
        }
    }
    return a;
}

template <class W, class S, class D> RACU_HD W cast1(W a)
{
    S x = Ty<S>::template get<W>(a);
    if constexpr (std::is_same_v<D, bool>) return Ty<bool>::template put<W>(x!=S(0));
    else return Ty<D>::template put<W>(static_cast<D>(x));
}

template <class W, class S> RACU_HD W cast_from(int dt, W a)
{
    switch (dt) {
    case RACU_BOOL: return cast1<W, S, bool>(a);
    case RACU_I32: return cast1<W, S, int32_t>(a);
    case RACU_F32: return cast1<W, S, float>(a);
    default:
        if constexpr (sizeof(W)==8) { return dt==RACU_I64 ? cast1<W, S, int64_t>(a) : cast1<W, S, double>(a); }
        return a;
    }
}

template <class W> RACU_HD W cast_any(int st, int dt, W a)
{
    switch (st) {
    case RACU_BOOL: return cast_from<W, bool>(dt, a);
    case RACU_I32: return cast_from<W, int32_t>(dt, a);
    case RACU_F32: return cast_from<W, float>(dt, a);
    default:
        if constexpr (sizeof(W)==8) { return st==RACU_I64 ? cast_from<W, int64_t>(dt, a) : cast_from<W, double>(dt, a); }
        return a;
    }
}

template <class W> RACU_HD W un_any(int op, int type, W a)
{
    switch (type) {
    case RACU_BOOL: return un1<W, bool>(op, a);
    case RACU_I32: return un1<W, int32_t>(op, a);
    case RACU_F32: return un1<W, float>(op, a);
    default:
        if constexpr (sizeof(W)==8) { return type==RACU_I64 ? un1<W, int64_t>(op, a) : un1<W, double>(op, a); }
        return a;
    }
}

template <class W> RACU_HD W bin_any(int op, int type, W a, W b)
{
    switch (type) {
    case RACU_BOOL: return bin1<W, bool>(op, a, b);
    case RACU_I32: return bin1<W, int32_t>(op, a, b);
    case RACU_F32: return bin1<W, float>(op, a, b);
    default:
        if constexpr (sizeof(W)==8) { return type==RACU_I64 ? bin1<W, int64_t>(op, a, b) : bin1<W, double>(op, a, b); }
        return a;
    }
}

template <class W> RACU_HD W fma_any(int type, W a, W b, W c)
{
    if (type==RACU_F32) return Ty<float>::template put<W>(fmaf(Ty<float>::template get<W>(a), Ty<float>::template get<W>(b), Ty<float>::template get<W>(c)));
    if constexpr (sizeof(W)==8) { if (type==RACU_F64) return Ty<double>::template put<W>(fma(Ty<double>::template get<W>(a), Ty<double>::template get<W>(b), Ty<double>::template get<W>(c))); }
    return c;
}
// ternary ops: fma(a, b, c), clamp(v, lo, hi), lerp(a, b, t)
template <class W, class T> RACU_HD W tern1(int op, W a, W b, W c)
{
    T const x = Ty<T>::template get<W>(a), y = Ty<T>::template get<W>(b), z = Ty<T>::template get<W>(c);
    if (op==RACU_OP_CLAMP) return Ty<T>::template put<W>(x<y ? y : z<x ? z : x); // std::clamp
    if constexpr (std::is_same_v<T, float> || std::is_same_v<T, double>) { if (op==RACU_OP_LERP) return Ty<T>::template put<W>(t_lerp<T>(x, y, z)); }
    return a;
}
template <class W> RACU_HD W tern_any(int op, int type, W a, W b, W c)
{
    if (op==RACU_OP_FMA) return fma_any<W>(type, a, b, c);
    switch (type) {
    case RACU_I32: return tern1<W, int32_t>(op, a, b, c);
    case RACU_F32: return tern1<W, float>(op, a, b, c);
    default:
        if constexpr (sizeof(W)==8) { return type==RACU_I64 ? tern1<W, int64_t>(op, a, b, c) : tern1<W, double>(op, a, b, c); }
        return a;
    }
}

// Fold combiners over accumulator type `type`. SUB and DIV accumulate the sum / product; the finish applies dst - s, dst / s.
template <class W> RACU_HD W combine(int fold_op, int type, W acc, W x)
{
    switch (fold_op) {
    case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: return bin_any<W>(RACU_OP_ADD, type, acc, x);
    case RACU_ASSIGN_MUL: case RACU_ASSIGN_DIV: return bin_any<W>(RACU_OP_MUL, type, acc, x);
    case RACU_ASSIGN_AMIN: return Ty<bool>::template get<W>(bin_any<W>(RACU_OP_LT, type, x, acc)) ? x : acc; // if (x<c) c=x
    default: return Ty<bool>::template get<W>(bin_any<W>(RACU_OP_LT, type, acc, x)) ? x : acc;              // if (c<x) c=x
    }
}
// dst_old OP total
template <class W> RACU_HD W finish_op(int fold_op, int type, W old, W total)
{
    switch (fold_op) {
    case RACU_ASSIGN_SUB: return bin_any<W>(RACU_OP_SUB, type, old, total);
    case RACU_ASSIGN_DIV: return bin_any<W>(RACU_OP_DIV, type, old, total);
    default: return combine<W>(fold_op, type, old, total);
    }
}

// ---------------- element <-> slot ----------------

template <class W> RACU_HD W load_elem(void const * ptr, int dtype)
{
    switch (dtype) {
    case RACU_BOOL: return W(*static_cast<uint8_t const *>(ptr) ? 1 : 0);
    case RACU_I32: case RACU_F32: return W(*static_cast<uint32_t const *>(ptr));
    default:
        if constexpr (sizeof(W)==8) return W(*static_cast<uint64_t const *>(ptr));
        return W(0);
    }
}
template <class W> RACU_HD void store_elem(void * ptr, int dtype, W v)
{
    switch (dtype) {
    case RACU_BOOL: *static_cast<uint8_t *>(ptr) = uint8_t(uint32_t(v)&1u); break;
    case RACU_I32: case RACU_F32: *static_cast<uint32_t *>(ptr) = uint32_t(v); break;
    default:
        if constexpr (sizeof(W)==8) *static_cast<uint64_t *>(ptr) = uint64_t(v);
    }
}

// SEQ value at element offset `off`: Seq: *p = i, p + d = i + I(d)  (ra/expr.hh:86-91).
template <class W> RACU_HD W seq_value(KOpnd const & o, int64_t off)
{
    switch (o.dtype) {
    case RACU_I32: return Ty<int32_t>::template put<W>(int32_t(int32_t(int64_t(o.base))+int32_t(off)));
    case RACU_F32: return Ty<float>::template put<W>(float(bits_f64(o.base))+float(off));
    case RACU_BOOL: return Ty<bool>::template put<W>((int64_t(o.base)+off)!=0);
    default:
        if constexpr (sizeof(W)==8) {
            if (o.dtype==RACU_I64) return Ty<int64_t>::template put<W>(int64_t(o.base)+off);
            return Ty<double>::template put<W>(bits_f64(o.base)+double(off));
        }
        return W(0);
    }
}
template <class W> RACU_HD W scalar_value(KOpnd const & o)
{
    switch (o.dtype) {
    case RACU_I32: return Ty<int32_t>::template put<W>(int32_t(int64_t(o.base)));
    case RACU_F32: return Ty<float>::template put<W>(float(bits_f64(o.base)));
    case RACU_BOOL: return Ty<bool>::template put<W>(int64_t(o.base)!=0);
    default:
        if constexpr (sizeof(W)==8) {
            if (o.dtype==RACU_I64) return Ty<int64_t>::template put<W>(int64_t(o.base));
            return Ty<double>::template put<W>(bits_f64(o.base));
        }
        return W(0);
    }
}

// ---------------- the postfix interpreter over U x V lanes ----------------
// One decode per instruction serves the U vectors a thread has staged (U*V lanes): the dispatch is a two level switch
// (type, then opcode) OUTSIDE the lane loops, so its cost is amortised over every lane.
//
// stage: this thread's staged operand slots; slot s of vector u at stage + (u*nstage + s)*sstride, holding V raw
//        elements of the operand's dtype.
// stack: this thread's value stack rows; level l of vector u at stack + (l*U + u)*sstride, V slots of W.
// Result in t[u][j].

#define RACU_LANES for (int u=0; u<U; ++u) _Pragma("unroll") for (int j=0; j<V; ++j)

template <class W, int U, class T> RACU_HD void
group_unary(int op, W (*t)[Lanes<W>::V])
{
    constexpr int V = Lanes<W>::V;
    switch (op) {
    case RACU_OP_NOT:
#pragma unroll
        RACU_LANES { t[u][j] = Ty<bool>::template put<W>(!Ty<T>::template get<W>(t[u][j])); } break;
    case RACU_OP_NEG:
        if constexpr (!std::is_same_v<T, bool>) {
#pragma unroll
            RACU_LANES { t[u][j] = Ty<T>::template put<W>(T(-Ty<T>::template get<W>(t[u][j]))); }
        } break;
    case RACU_OP_ABS:
        if constexpr (!std::is_same_v<T, bool>) {
#pragma unroll
            RACU_LANES { t[u][j] = Ty<T>::template put<W>(t_abs<T>(Ty<T>::template get<W>(t[u][j]))); }
        } break;
    case RACU_OP_SQR:
        if constexpr (!std::is_same_v<T, bool>) {
#pragma unroll
            RACU_LANES { T x = Ty<T>::template get<W>(t[u][j]); t[u][j] = Ty<T>::template put<W>(T(x*x)); }
        } break;
    default:
#pragma unroll
        RACU_LANES { t[u][j] = un1<W, T>(op, t[u][j]); } break; // sqrt and the out-of-line transcendentals
    }
}

template <class W, int U, class T> RACU_HD void
group_binary(int op, unsigned char const * lv, int sstride, W (*t)[Lanes<W>::V])
{
    constexpr int V = Lanes<W>::V;
#define RACU_A(u, j) Ty<T>::template get<W>(reinterpret_cast<W const *>(lv + (u)*sstride)[j])
#define RACU_B(u, j) Ty<T>::template get<W>(t[u][j])
#define RACU_BINCASE(OPC, R, EXPR) case OPC: _Pragma("unroll") RACU_LANES { T x = RACU_A(u, j), y = RACU_B(u, j); t[u][j] = Ty<R>::template put<W>(EXPR); } break;
    switch (op) {
    RACU_BINCASE(RACU_OP_EQ, bool, x==y) RACU_BINCASE(RACU_OP_NE, bool, x!=y) RACU_BINCASE(RACU_OP_LT, bool, x<y)
    RACU_BINCASE(RACU_OP_LE, bool, x<=y) RACU_BINCASE(RACU_OP_GT, bool, x>y) RACU_BINCASE(RACU_OP_GE, bool, x>=y)
    default:
        if constexpr (std::is_same_v<T, bool>) {
#pragma unroll
            RACU_LANES { t[u][j] = bin1<W, T>(op, reinterpret_cast<W const *>(lv + u*sstride)[j], t[u][j]); }
        } else {
            switch (op) {
            RACU_BINCASE(RACU_OP_ADD, T, T(x+y)) RACU_BINCASE(RACU_OP_SUB, T, T(x-y)) RACU_BINCASE(RACU_OP_MUL, T, T(x*y))
            RACU_BINCASE(RACU_OP_MIN, T, (y<x ? y : x)) RACU_BINCASE(RACU_OP_MAX, T, (x<y ? y : x))
            default:
#pragma unroll
                RACU_LANES { t[u][j] = bin1<W, T>(op, reinterpret_cast<W const *>(lv + u*sstride)[j], t[u][j]); } break; // div, mod, bit ops, pow, atan2
            }
        }
        break;
    }
#undef RACU_BINCASE
#undef RACU_A
#undef RACU_B
}

template <class W, int U, class S> RACU_HD void
group_cast_from(int dt, W (*t)[Lanes<W>::V])
{
    constexpr int V = Lanes<W>::V;
#define RACU_CASTCASE(D) _Pragma("unroll") RACU_LANES { t[u][j] = cast1<W, S, D>(t[u][j]); } break;
    switch (dt) {
    case RACU_BOOL: RACU_CASTCASE(bool)
    case RACU_I32: RACU_CASTCASE(int32_t)
    case RACU_F32: RACU_CASTCASE(float)
    default:
        if constexpr (sizeof(W)==8) { if (dt==RACU_I64) { RACU_CASTCASE(int64_t) } else { RACU_CASTCASE(double) } }
        break;
    }
#undef RACU_CASTCASE
}

#define RACU_TYPE_SWITCH(TYPE, CALL)                                            \
    switch (TYPE) {                                                             \
    case RACU_BOOL: { using T = bool; CALL; } break;                            \
    case RACU_I32: { using T = int32_t; CALL; } break;                          \
    case RACU_F32: { using T = float; CALL; } break;                            \
    default:                                                                    \
        if constexpr (sizeof(W)==8) {                                           \
            if ((TYPE)==RACU_I64) { using T = int64_t; CALL; } else { using T = double; CALL; } \
        }                                                                       \
        break;                                                                  \
    }

template <class W, int U> RACU_HD void
interp(KParams const & p, unsigned char const * stage, unsigned char * stack, int sstride, W (*t)[Lanes<W>::V])
{
    constexpr int V = Lanes<W>::V;
    int const nstage = p.nstage;
    for (int pc=0; pc<p.nins; ++pc) {
        KIns const in = p.ins[pc];
        unsigned char * lv = stack + int(in.lvl)*U*sstride;
        switch (in.op) {
        case RACU_OP_PUSH: {
            if (in.lvl!=255) {
#pragma unroll
                RACU_LANES { reinterpret_cast<W *>(lv + u*sstride)[j] = t[u][j]; }
            }
            KOpnd const & o = p.opnd[in.arg];
            if (S_NONE==o.stage) {
                W s = scalar_value<W>(o);
#pragma unroll
                RACU_LANES { t[u][j] = s; }
            } else {
                unsigned char const * sl = stage + int(o.slot)*sstride;
                int const us = nstage*sstride;
                if (4==o.esize) {
#pragma unroll
                    RACU_LANES { t[u][j] = W(reinterpret_cast<uint32_t const *>(sl + u*us)[j]); }
                } else if (1==o.esize) {
#pragma unroll
                    RACU_LANES { t[u][j] = W((sl + u*us)[j] ? 1 : 0); }
                } else {
                    if constexpr (sizeof(W)==8) {
#pragma unroll
                        RACU_LANES { t[u][j] = reinterpret_cast<uint64_t const *>(sl + u*us)[j]; }
                    }
                }
            }
            break;
        }
        case RACU_OP_CAST:
            RACU_TYPE_SWITCH(in.aux, (group_cast_from<W, U, T>(in.type, t)))
            break;
        case RACU_OP_NEG: case RACU_OP_ABS: case RACU_OP_SQRT: case RACU_OP_SQR: case RACU_OP_EXP:
        case RACU_OP_LOG: case RACU_OP_SIN: case RACU_OP_COS: case RACU_OP_NOT:
        case RACU_OP_TAN: case RACU_OP_SINH: case RACU_OP_COSH: case RACU_OP_TANH: case RACU_OP_ASIN: case RACU_OP_ACOS: case RACU_OP_ATAN:
        case RACU_OP_EXPM1: case RACU_OP_LOG1P: case RACU_OP_LOG10: case RACU_OP_ISFINITE: case RACU_OP_ISNAN: case RACU_OP_ISINF:
            RACU_TYPE_SWITCH(in.type, (group_unary<W, U, T>(in.op, t)))
            break;
        case RACU_OP_FMA: case RACU_OP_CLAMP: case RACU_OP_LERP: {
            RACU_LANES {
                W const * la = reinterpret_cast<W const *>(lv + u*sstride);
                W const * lb = reinterpret_cast<W const *>(lv + (U+u)*sstride);
                t[u][j] = tern_any<W>(in.op, in.type, la[j], lb[j], t[u][j]);
            }
            break;
        }
        case RACU_OP_LOAD: { // gather: t = opnd[arg][clamp(t)] (unbeaten subscripts, ra/ply.hh:461-499); lanes past the end hold
            KOpnd const & o = p.opnd[in.arg]; // anything, the clamp keeps every access inside the array
            unsigned char const * const gb = reinterpret_cast<unsigned char const *>(o.base);
            RACU_LANES {
                int64_t off = in.aux==RACU_I32 ? int64_t(int32_t(uint32_t(t[u][j]))) : int64_t(t[u][j]);
                off = off<o.sA[0] ? o.sA[0] : off>o.sA[1] ? o.sA[1] : off;
                t[u][j] = load_elem<W>(gb + off*int64_t(o.esize), o.dtype);
            }
            break;
        }
        case RACU_OP_PICK: {
            int n = in.arg;
            RACU_LANES {
                W sel = reinterpret_cast<W const *>(lv + u*sstride)[j];
                if (in.aux==RACU_BOOL) sel = W(uint32_t(sel)&0xffu);
                else if (in.aux==RACU_I32) sel = W(uint32_t(sel));
                if (!(sel<W(n))) sel = 0; // out of contract selectors take branch 0, like the oracle
                if (sel+1<W(n)) t[u][j] = reinterpret_cast<W const *>(lv + ((1+int(sel))*U+u)*sstride)[j];
            }
            break;
        }
        default:
            RACU_TYPE_SWITCH(in.type, (group_binary<W, U, T>(in.op, lv, sstride, t)))
            break;
        }
    }
}

// acc[j] = combine(acc[j], t[u][j]) over the lanes selected by `valid(u, j)`; dispatch hoisted out of the lane loops.
template <class W, int U, class T, class Valid> RACU_HD void
group_combine_t(int fold_op, W * acc, W (*t)[Lanes<W>::V], Valid valid)
{
    constexpr int V = Lanes<W>::V;
#define RACU_ACC(EXPR) _Pragma("unroll") RACU_LANES { if (valid(u, j)) { T c = Ty<T>::template get<W>(acc[j]), x = Ty<T>::template get<W>(t[u][j]); acc[j] = Ty<T>::template put<W>(EXPR); } } break;
    switch (fold_op) {
    case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB:
        if constexpr (!std::is_same_v<T, bool>) { RACU_ACC(T(c+x)) } break;
    case RACU_ASSIGN_MUL: case RACU_ASSIGN_DIV:
        if constexpr (!std::is_same_v<T, bool>) { RACU_ACC(T(c*x)) } break;
    case RACU_ASSIGN_AMIN: RACU_ACC((x<c ? x : c))
    default: RACU_ACC((c<x ? x : c))
    }
#undef RACU_ACC
}
template <class W, int U, class Valid> RACU_HD void
group_combine(int fold_op, int type, W * acc, W (*t)[Lanes<W>::V], Valid valid)
{
    RACU_TYPE_SWITCH(type, (group_combine_t<W, U, T>(fold_op, acc, t, valid)))
}

} // namespace racu
