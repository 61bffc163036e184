// ply_oracle.cc - CPU restatement of ra-ra's traversal hot path. TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the CUDA back end. Nothing under ra_ra_b200/ or include/ may call,
// link or load it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs do, and only as the checker. The reference itself (header-only C++23, needs gcc >= 14: deducing
// this ra/base.hh:123, std::ranges::to ra/base.hh:427, <print> ra/base.hh:462) cannot be compiled with the
// g++ 13.3 of this image, so the algorithm is restated here, literally and sequentially, and PINNED by the
// reference's own known-answer tests (tests/test_oracle_golden.py lists them with file:line).
//
// Each function cites the reference lines it follows. Citations are relative to the reference root.
// Build: g++ -O2 -std=c++20 -ffp-contract=off (no FMA contraction, no -ffast-math; the reference never
// uses -ffast-math, config/ra.py:30).

#include "../include/racu.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char * msg)
{
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

using dim_t = int64_t; // ra/base.hh:237
constexpr dim_t UNB = RACU_UNB, MIS = RACU_MIS;

// ra/expr.hh:529. finite before UNB; two different finite -> MIS.
constexpr dim_t colen(dim_t a, dim_t b) { return a>=0 ? (a==b ? a : b>=0 ? MIS : a) : UNB==a ? b : UNB==b ? a : b; }

template <class F> decltype(auto)
with_type(int t, F && f)
{
    switch (t) {
    case RACU_BOOL: return f(bool {});
    case RACU_I32: return f(int32_t {});
    case RACU_I64: return f(int64_t {});
    case RACU_F32: return f(float {});
    default: return f(double {});
    }
}

struct Val
{
    int type;
    union { bool b; int32_t i32; int64_t i64; float f32; double f64; };
};

template <class T> T get(Val const & v)
{
    if constexpr (std::is_same_v<T, bool>) return v.b;
    else if constexpr (std::is_same_v<T, int32_t>) return v.i32;
    else if constexpr (std::is_same_v<T, int64_t>) return v.i64;
    else if constexpr (std::is_same_v<T, float>) return v.f32;
    else return v.f64;
}

template <class T> Val mk(T x)
{
    Val v;
    if constexpr (std::is_same_v<T, bool>) { v.type=RACU_BOOL; v.i64=0; v.b=x; }
    else if constexpr (std::is_same_v<T, int32_t>) { v.type=RACU_I32; v.i64=0; v.i32=x; }
    else if constexpr (std::is_same_v<T, int64_t>) { v.type=RACU_I64; v.i64=x; }
    else if constexpr (std::is_same_v<T, float>) { v.type=RACU_F32; v.i64=0; v.f32=x; }
    else { v.type=RACU_F64; v.f64=x; }
    return v;
}

// Static check of the postfix program: simulates the type stack. Returns result type or -1.
int
check_program(racu_desc const * d)
{
    int st[RACU_MAX_INS+1], sp=0;
    if (d->nins<=0 || d->nins>RACU_MAX_INS || d->nopnd<0 || d->nopnd>RACU_MAX_OPND) return -1;
    for (int pc=0; pc<d->nins; ++pc) {
        racu_ins const & in = d->ins[pc];
        if (in.type>=RACU_NDTYPE) return -1;
        bool isf = in.type==RACU_F32 || in.type==RACU_F64, isb = in.type==RACU_BOOL;
        switch (in.op) {
        case RACU_OP_PUSH:
            if (in.arg>=d->nopnd || d->opnd[in.arg].dtype!=in.type) return -1;
            st[sp++] = in.type; break;
        case RACU_OP_CAST:
            if (sp<1 || st[sp-1]!=in.aux) return -1;
            st[sp-1] = in.type; break;
        case RACU_OP_NEG: case RACU_OP_ABS: case RACU_OP_SQR:
            if (sp<1 || st[sp-1]!=in.type || isb) return -1;
            break;
        case RACU_OP_SQRT: case RACU_OP_EXP: case RACU_OP_LOG: case RACU_OP_SIN: case RACU_OP_COS:
        case RACU_OP_TAN: case RACU_OP_SINH: case RACU_OP_COSH: case RACU_OP_TANH: case RACU_OP_ASIN: case RACU_OP_ACOS: case RACU_OP_ATAN:
        case RACU_OP_EXPM1: case RACU_OP_LOG1P: case RACU_OP_LOG10:
            if (sp<1 || st[sp-1]!=in.type || !isf) return -1;
            break;
        case RACU_OP_ISFINITE: case RACU_OP_ISNAN: case RACU_OP_ISINF:
            if (sp<1 || st[sp-1]!=in.type || !isf) return -1;
            st[sp-1] = RACU_BOOL; break;
        case RACU_OP_CLAMP: case RACU_OP_LERP:
            if (sp<3 || st[sp-1]!=in.type || st[sp-2]!=in.type || st[sp-3]!=in.type || isb || (in.op==RACU_OP_LERP && !isf)) return -1;
            sp -= 2; break;
        case RACU_OP_NOT:
            if (sp<1 || st[sp-1]!=in.type) return -1;
            st[sp-1] = RACU_BOOL; break;
        case RACU_OP_ADD: case RACU_OP_SUB: case RACU_OP_MUL: case RACU_OP_DIV: case RACU_OP_MOD:
        case RACU_OP_MIN: case RACU_OP_MAX:
            if (sp<2 || st[sp-1]!=in.type || st[sp-2]!=in.type || isb) return -1;
            --sp; break;
        case RACU_OP_POW: case RACU_OP_ATAN2:
            if (sp<2 || st[sp-1]!=in.type || st[sp-2]!=in.type || !isf) return -1;
            --sp; break;
        case RACU_OP_BAND: case RACU_OP_BOR: case RACU_OP_BXOR:
            if (sp<2 || st[sp-1]!=in.type || st[sp-2]!=in.type || isf) return -1;
            --sp; break;
        case RACU_OP_EQ: case RACU_OP_NE: case RACU_OP_LT: case RACU_OP_LE: case RACU_OP_GT: case RACU_OP_GE:
            if (sp<2 || st[sp-1]!=in.type || st[sp-2]!=in.type) return -1;
            --sp; st[sp-1] = RACU_BOOL; break;
        case RACU_OP_FMA:
            if (sp<3 || st[sp-1]!=in.type || st[sp-2]!=in.type || st[sp-3]!=in.type || !isf) return -1;
            sp -= 2; break;
        case RACU_OP_PICK: {
            int n = in.arg;
            if (n<1 || sp<n+1) return -1;
            for (int j=0; j<n; ++j) { if (st[sp-1-j]!=in.type) return -1; }
            int sel = st[sp-1-n];
            if (sel!=in.aux || sel==RACU_F32 || sel==RACU_F64) return -1;
            sp -= n; st[sp-1] = in.type; break;
        }
        case RACU_OP_LOAD:
            if (in.arg>=d->nopnd || d->opnd[in.arg].kind!=RACU_PTR || d->opnd[in.arg].dtype!=in.type) return -1;
            if (sp<1 || st[sp-1]!=in.aux || (in.aux!=RACU_I32 && in.aux!=RACU_I64)) return -1;
            st[sp-1] = in.type; break;
        default: return -1;
        }
    }
    // scatter (dst is a RACU_PTR operand): the program leaves [element offset, value]
    if (d->dst>=0 && d->dst<d->nopnd && d->opnd[d->dst].kind==RACU_PTR) return (2==sp && (st[0]==RACU_I32 || st[0]==RACU_I64)) ? st[1] : -1;
    return 1==sp ? st[0] : -1;
}

// Match restated: ra/expr.hh:532-608. rank = max over operands (corank :530); len(k) folds colen over the
// operands with rank > k (:597-608). Any len(k)<0 after matching is an error at validate() (:509-526, :555-567),
// which also covers the wholly unbounded expression (test/checks.cc:159-193).
int
agree(racu_desc const * d, int32_t * rank_out, dim_t * len_out)
{
    if (d->nopnd<=0 || d->nopnd>RACU_MAX_OPND) return fail(RACU_ERR_BAD_DESC, "bad operand count");
    int rank = 0;
    for (int o=0; o<d->nopnd; ++o) {
        racu_operand const & p = d->opnd[o];
        if (p.rank<0 || p.rank>RACU_MAX_RANK) return fail(RACU_ERR_BAD_DESC, "bad operand rank");
        if (p.kind==RACU_SCALAR && p.rank!=0) return fail(RACU_ERR_BAD_DESC, "scalar operand must have rank 0");
        rank = std::max(rank, p.rank);
    }
    for (int k=0; k<rank; ++k) {
        dim_t s = UNB;
        for (int o=0; o<d->nopnd && MIS!=s; ++o) {
            racu_operand const & p = d->opnd[o];
            if (k<p.rank) {
                dim_t sk = p.len[k];
                if (sk<0 && sk!=UNB) return fail(RACU_ERR_BAD_DESC, "bad operand len");
                s = colen(sk, s);
            }
        }
        if (s<0) {
            std::snprintf(g_err, sizeof(g_err), "Bad shapes: axis %d is %s.", k, MIS==s ? "mismatched" : "unbounded");
            return RACU_ERR_SHAPE;
        }
        if (len_out) len_out[k] = s;
    }
    if (rank_out) *rank_out = rank;
    return RACU_OK;
}

// One leaf during traversal: the Iterator state of Cell / Scalar (ra/expr.hh:104-120,266-312).
// `pos` is the pointer as an element offset for MEM, and the Seq value offset for SEQ (Seq::operator+=, ra/expr.hh:96).
struct Leaf
{
    racu_operand const * p;
    dim_t pos = 0;
    dim_t step(int k) const { return k<p->rank ? p->stride[k] : 0; } // ra/expr.hh:294-295; Scalar :112
    bool keep(dim_t st, int z, int j) const { return RACU_SCALAR==p->kind ? true : st*step(z)==step(j); } // :296-297, :114
    void adv(int k, dim_t d) { pos += step(k)*d; } // :300
    void mov(dim_t d) { pos += d; } // :310
};

Val
load(Leaf const & l)
{
    racu_operand const & p = *l.p;
    return with_type(p.dtype, [&](auto t) {
        using T = decltype(t);
        switch (p.kind) {
        case RACU_MEM: return mk<T>(static_cast<T const *>(p.base.ptr)[l.pos]);
        case RACU_SEQ: // Seq: *p = i; p+d = i+I(d). ra/expr.hh:86-91
            if constexpr (std::is_floating_point_v<T>) return mk<T>(T(p.base.f) + T(l.pos));
            else return mk<T>(T(T(p.base.i) + T(l.pos)));
        default:
            if constexpr (std::is_floating_point_v<T>) return mk<T>(T(p.base.f));
            else return mk<T>(T(p.base.i));
        }
    });
}

template <class T> Val
unary(int op, T a)
{
    switch (op) {
    case RACU_OP_NEG: if constexpr (!std::is_same_v<T, bool>) return mk<T>(T(-a)); break;
    case RACU_OP_ABS: if constexpr (!std::is_same_v<T, bool>) return mk<T>(T(std::abs(a))); break;
    case RACU_OP_SQR: if constexpr (!std::is_same_v<T, bool>) return mk<T>(T(a*a)); break; // ra/ra.hh:65
    case RACU_OP_NOT: return mk<bool>(!a);
    default:
        if constexpr (std::is_floating_point_v<T>) {
            switch (op) {
            case RACU_OP_SQRT: return mk<T>(std::sqrt(a));
            case RACU_OP_EXP: return mk<T>(std::exp(a));
            case RACU_OP_LOG: return mk<T>(std::log(a));
            case RACU_OP_SIN: return mk<T>(std::sin(a));
            case RACU_OP_COS: return mk<T>(std::cos(a));
            // the using-declared <cmath> functions of ra/ra.hh:221-222
            case RACU_OP_TAN: return mk<T>(std::tan(a));
            case RACU_OP_SINH: return mk<T>(std::sinh(a));
            case RACU_OP_COSH: return mk<T>(std::cosh(a));
            case RACU_OP_TANH: return mk<T>(std::tanh(a));
            case RACU_OP_ASIN: return mk<T>(std::asin(a));
            case RACU_OP_ACOS: return mk<T>(std::acos(a));
            case RACU_OP_ATAN: return mk<T>(std::atan(a));
            case RACU_OP_EXPM1: return mk<T>(std::expm1(a));
            case RACU_OP_LOG1P: return mk<T>(std::log1p(a));
            case RACU_OP_LOG10: return mk<T>(std::log10(a));
            case RACU_OP_ISFINITE: return mk<bool>(std::isfinite(a));
            case RACU_OP_ISNAN: return mk<bool>(std::isnan(a));
            case RACU_OP_ISINF: return mk<bool>(std::isinf(a));
            }
        }
    }
    std::abort();
}

template <class T> Val
binary(int op, T a, T b)
{
    switch (op) {
    case RACU_OP_EQ: return mk<bool>(a==b);
    case RACU_OP_NE: return mk<bool>(a!=b);
    case RACU_OP_LT: return mk<bool>(a<b);
    case RACU_OP_LE: return mk<bool>(a<=b);
    case RACU_OP_GT: return mk<bool>(a>b);
    case RACU_OP_GE: return mk<bool>(a>=b);
    }
    if constexpr (!std::is_floating_point_v<T>) {
        switch (op) {
        case RACU_OP_BAND: return mk<T>(T(a & b));
        case RACU_OP_BOR: return mk<T>(T(a | b));
        case RACU_OP_BXOR: return mk<T>(T(a ^ b));
        }
    }
    if constexpr (!std::is_same_v<T, bool>) {
        switch (op) {
        case RACU_OP_ADD: return mk<T>(T(a+b));
        case RACU_OP_SUB: return mk<T>(T(a-b));
        case RACU_OP_MUL: return mk<T>(T(a*b));
        case RACU_OP_MIN: return mk<T>(b<a ? b : a); // std::min
        case RACU_OP_MAX: return mk<T>(a<b ? b : a); // std::max
        case RACU_OP_DIV:
            if constexpr (std::is_integral_v<T>) return mk<T>(0==b ? T(0) : T(a/b)); // C++ UB; defined as 0 on both sides
            else return mk<T>(a/b);
        case RACU_OP_MOD:
            if constexpr (std::is_integral_v<T>) return mk<T>(0==b ? T(0) : T(a%b));
            else return mk<T>(std::fmod(a, b));
        }
        if constexpr (std::is_floating_point_v<T>) {
            switch (op) {
            case RACU_OP_POW: return mk<T>(std::pow(a, b));
            case RACU_OP_ATAN2: return mk<T>(std::atan2(a, b));
            }
        }
    }
    std::abort();
}

// Map::operator* (ra/expr.hh:661): op(*leaf...). Pick::operator* (:720) and pick_star (:701-710).
Val
eval(racu_desc const * d, Leaf const * leaf, Val * second=nullptr)
{
    Val st[RACU_MAX_INS+1];
    int sp = 0;
    for (int pc=0; pc<d->nins; ++pc) {
        racu_ins const & in = d->ins[pc];
        switch (in.op) {
        case RACU_OP_PUSH: st[sp++] = load(leaf[in.arg]); break;
        case RACU_OP_CAST:
            st[sp-1] = with_type(in.aux, [&](auto s) {
                return with_type(in.type, [&](auto t) { return mk<decltype(t)>(static_cast<decltype(t)>(get<decltype(s)>(st[sp-1]))); });
            });
            break;
        case RACU_OP_NEG: case RACU_OP_ABS: case RACU_OP_SQRT: case RACU_OP_SQR: case RACU_OP_EXP:
        case RACU_OP_LOG: case RACU_OP_SIN: case RACU_OP_COS: case RACU_OP_NOT:
        case RACU_OP_TAN: case RACU_OP_SINH: case RACU_OP_COSH: case RACU_OP_TANH: case RACU_OP_ASIN: case RACU_OP_ACOS: case RACU_OP_ATAN:
        case RACU_OP_EXPM1: case RACU_OP_LOG1P: case RACU_OP_LOG10: case RACU_OP_ISFINITE: case RACU_OP_ISNAN: case RACU_OP_ISINF:
            st[sp-1] = with_type(in.type, [&](auto t) { return unary<decltype(t)>(in.op, get<decltype(t)>(st[sp-1])); });
            break;
        case RACU_OP_CLAMP: // std::clamp, ra/ra.hh:222
            st[sp-3] = with_type(in.type, [&](auto t) {
                using T = decltype(t);
                if constexpr (!std::is_same_v<T, bool>) return mk<T>(std::clamp(get<T>(st[sp-3]), get<T>(st[sp-2]), get<T>(st[sp-1])));
                else return st[sp-3];
            });
            sp -= 2; break;
        case RACU_OP_LERP: // std::lerp, ra/ra.hh:222
            st[sp-3] = with_type(in.type, [&](auto t) {
                using T = decltype(t);
                if constexpr (std::is_floating_point_v<T>) return mk<T>(std::lerp(get<T>(st[sp-3]), get<T>(st[sp-2]), get<T>(st[sp-1])));
                else return st[sp-3];
            });
            sp -= 2; break;
        case RACU_OP_FMA:
            st[sp-3] = with_type(in.type, [&](auto t) {
                using T = decltype(t);
                if constexpr (std::is_floating_point_v<T>) return mk<T>(std::fma(get<T>(st[sp-3]), get<T>(st[sp-2]), get<T>(st[sp-1])));
                else return st[sp-3];
            });
            sp -= 2; break;
        case RACU_OP_PICK: {
            int n = in.arg;
            int64_t sel = with_type(in.aux, [&](auto s) { return int64_t(get<decltype(s)>(st[sp-1-n])); });
            if (sel<0 || sel>=n) { sel = 0; } // RA_CK "Bad pick" ra/expr.hh:708: out of contract
            st[sp-1-n] = st[sp-n+sel];
            sp -= n; break;
        }
        case RACU_OP_LOAD: { // unbeaten subscripts: *frompl(a.data(), a, i ...) (ra/ply.hh:531-549) with the offset already summed
            racu_operand const & a = d->opnd[in.arg];
            int64_t off = with_type(in.aux, [&](auto s) { return int64_t(get<decltype(s)>(st[sp-1])); });
            off = off<a.len[0] ? a.len[0] : off>a.len[1] ? a.len[1] : off; // out of contract (RA_CK "Bad index", ra/ply.hh:395): clamp
            st[sp-1] = with_type(in.type, [&](auto t) { return mk<decltype(t)>(static_cast<decltype(t) const *>(a.base.ptr)[off]); });
            break;
        }
        default:
            st[sp-2] = with_type(in.type, [&](auto t) { return binary<decltype(t)>(in.op, get<decltype(t)>(st[sp-2]), get<decltype(t)>(st[sp-1])); });
            --sp; break;
        }
    }
    if (second && sp>=2) *second = st[1];
    return st[0];
}

// y OP= x with C++'s own conversions. ra/expr.hh:51: [](auto && y, auto && x){ y OP x; }.
// AMIN/AMAX: ra/ra.hh:311,320.
template <class Ty, class Tx> void
assign1(int as, Ty & y, Tx x)
{
    switch (as) {
    case RACU_ASSIGN_SET: y = Ty(x); return;
    case RACU_ASSIGN_AMIN: if (x<y) { y = Ty(x); } return;
    case RACU_ASSIGN_AMAX: if (y<x) { y = Ty(x); } return;
    }
    if constexpr (!std::is_same_v<Ty, bool> && !std::is_same_v<Tx, bool>) {
        switch (as) {
        case RACU_ASSIGN_ADD: y += x; return;
        case RACU_ASSIGN_SUB: y -= x; return;
        case RACU_ASSIGN_MUL: y *= x; return;
        case RACU_ASSIGN_DIV:
            if constexpr (std::is_integral_v<Ty> && std::is_integral_v<Tx>) { y = (0==x) ? Ty(0) : Ty(y/x); }
            else { y /= x; }
            return;
        }
    }
    std::abort();
}

void
assign(int as, int dtype, void * base, dim_t pos, Val const & x)
{
    with_type(dtype, [&](auto ty) {
        using Ty = decltype(ty);
        Ty & y = static_cast<Ty *>(base)[pos];
        with_type(x.type, [&](auto tx) { assign1<Ty, decltype(tx)>(as, y, get<decltype(tx)>(x)); return 0; });
        return 0;
    });
}

int
check_desc(racu_desc const * d, bool need_dst)
{
    if (!d) return fail(RACU_ERR_BAD_DESC, "null descriptor");
    int rt = check_program(d);
    if (rt<0) return fail(RACU_ERR_BAD_DESC, "bad program");
    if (need_dst) {
        if (d->dst<0 || d->dst>=d->nopnd || (d->opnd[d->dst].kind!=RACU_MEM && d->opnd[d->dst].kind!=RACU_PTR)) return fail(RACU_ERR_BAD_DESC, "bad destination");
        if (d->assign<RACU_ASSIGN_SET || d->assign>RACU_ASSIGN_AMAX) return fail(RACU_ERR_BAD_DESC, "bad assign op");
        int dt = d->opnd[d->dst].dtype;
        if (d->assign>=RACU_ASSIGN_ADD && d->assign<=RACU_ASSIGN_DIV && (dt==RACU_BOOL || rt==RACU_BOOL))
            return fail(RACU_ERR_BAD_DESC, "arithmetic assign on bool");
    }
    return RACU_OK;
}

// ply_ravel restated: ra/ply.hh:21-86. `sink(leaves, value)` is the body of the op for one element.
template <class Sink> int
ply_ravel(racu_desc const * d, Sink && sink)
{
    int32_t rank;
    dim_t len[RACU_MAX_RANK];
    if (int e = agree(d, &rank, len)) return e; // validate(a) :25
    std::vector<Leaf> leaf(d->nopnd);
    for (int o=0; o<d->nopnd; ++o) { leaf[o].p = &d->opnd[o]; }
    if (0==rank) { // :27-33
        sink(leaf.data(), eval(d, leaf.data()));
        return RACU_OK;
    }
    struct z_t { int order; dim_t sha, ind=0; };
    std::vector<z_t> z(rank);
    for (int i=0; i<rank; ++i) { z[i].order = rank-1-i; } // :38-40 inside first
    // find outermost compact dim :42-46
    int ocd = 0;
    dim_t ss = len[z[ocd].order];
    int r = rank;
    auto keep = [&](dim_t st, int zz, int j) { for (auto const & l: leaf) { if (!l.keep(st, zz, j)) return false; } return true; }; // Match::keep :609-613
    for (--r, ++ocd; r>0 && keep(ss, z[0].order, z[ocd].order); --r, ++ocd) {
        ss *= len[z[ocd].order];
    }
    for (int k=0; k<r; ++k) { // :47-56
        if (0>=(z[k].sha=len[z[ocd+k].order])) return RACU_OK;
    }
    std::vector<dim_t> ss0(d->nopnd), place(d->nopnd);
    for (int o=0; o<d->nopnd; ++o) { ss0[o] = leaf[o].step(z[0].order); } // :57
    for (;;) {
        for (int o=0; o<d->nopnd; ++o) { place[o] = leaf[o].pos; } // save :59
        for (dim_t s=ss; --s>=0;) { // :60
            sink(leaf.data(), eval(d, leaf.data()));
            for (int o=0; o<d->nopnd; ++o) { leaf[o].mov(ss0[o]); }
        }
        for (int o=0; o<d->nopnd; ++o) { leaf[o].pos = place[o]; } // load :69
        for (int k=0; ; ++k) { // :70-84
            if (k>=r) {
                return RACU_OK;
            } else if (++z[k].ind<z[k].sha) {
                for (auto & l: leaf) { l.adv(z[ocd+k].order, 1); }
                break;
            } else {
                z[k].ind = 0;
                for (auto & l: leaf) { l.adv(z[ocd+k].order, 1-z[k].sha); }
            }
        }
    }
}

// ply_fixed / subply restated: ra/ply.hh:90-153. Same row-major order as ply_ravel but with NO collapsing of axes: a recursion
// over the outer axes (subply, :105-118) and a flat loop over the innermost one only (:95-104). The reference selects it for static
// shapes (:161-165) and its own tests run both traversals side by side (TEST(plier), test/ply.cc:35-44); so does this oracle
// (oracle_set_traversal), which pins the collapse rule of ply_ravel against the traversal that does not need it.
template <class Sink> int
ply_fixed(racu_desc const * d, Sink && sink)
{
    int32_t rank;
    dim_t len[RACU_MAX_RANK];
    if (int e = agree(d, &rank, len)) return e; // validate(a) :131
    std::vector<Leaf> leaf(d->nopnd);
    for (int o=0; o<d->nopnd; ++o) { leaf[o].p = &d->opnd[o]; }
    if (0==rank) { sink(leaf.data(), eval(d, leaf.data())); return RACU_OK; } // :134-139
    auto order = [&](int k) { return rank-1-k; }; // inside first :142
    std::vector<dim_t> ss0(d->nopnd);
    for (int o=0; o<d->nopnd; ++o) { ss0[o] = leaf[o].step(order(0)); } // :145
    auto subply = [&](auto && self, int k) -> void {
        if (k<1) { // :94-104
            std::vector<dim_t> place(d->nopnd);
            for (int o=0; o<d->nopnd; ++o) { place[o] = leaf[o].pos; }
            for (dim_t s=len[order(0)]; --s>=0;) {
                sink(leaf.data(), eval(d, leaf.data()));
                for (int o=0; o<d->nopnd; ++o) { leaf[o].mov(ss0[o]); }
            }
            for (int o=0; o<d->nopnd; ++o) { leaf[o].pos = place[o]; }
        } else { // :105-118
            dim_t const size = len[order(k)];
            for (dim_t i=0; i<size; ++i) {
                self(self, k-1);
                for (auto & l: leaf) { l.adv(order(k), 1); }
            }
            for (auto & l: leaf) { l.adv(order(k), -size); }
        }
    };
    subply(subply, rank-1);
    return RACU_OK;
}

int g_traversal = 0; // 0: ply_ravel, 1: ply_fixed

template <class Sink> int
traverse(racu_desc const * d, Sink && sink)
{
    return g_traversal ? ply_fixed(d, sink) : ply_ravel(d, sink);
}

} // namespace

extern "C" {

const char * oracle_last_error_string(void) { return g_err; }
// Which restated traversal the oracle uses: 0 = ply_ravel (ra/ply.hh:21-86, default), 1 = ply_fixed (ra/ply.hh:90-153).
void oracle_set_traversal(int fixed) { g_traversal = fixed ? 1 : 0; }

int oracle_agree(racu_desc const * d, int32_t * rank_out, int64_t * len_out) { return agree(d, rank_out, len_out); }

// dst OP= program, sequentially, in the reference's row-major order. When dst has step 0 on an axis this IS the
// reference's fold (test/ra-12.cc:28, test/reduction.cc:155-158), including "last one wins" for SET (docs/ra-ra.texi:1986-2007).
int
oracle_map(racu_desc const * d)
{
    if (int e = check_desc(d, true)) return e;
    racu_operand const & y = d->opnd[d->dst];
    int dst = d->dst, as = d->assign;
    if (y.kind==RACU_PTR) { // a(i ...) OP= x with index arrays (test/fromu.cc:53-61,84-92): y = a[offset], sequentially: repeated indices accumulate
        return traverse(d, [&](Leaf const * lf, Val const & off) {
            Val x;
            eval(d, lf, &x); // (the traversal's own eval returned only the bottom of the stack)
            int64_t o = off.type==RACU_I32 ? int64_t(off.i32) : off.i64;
            o = o<y.len[0] ? y.len[0] : o>y.len[1] ? y.len[1] : o;
            assign(as, y.dtype, y.base.ptr, o, x);
        });
    }
    return traverse(d, [&](Leaf const * leaf, Val const & x) { assign(as, y.dtype, y.base.ptr, leaf[dst].pos, x); });
}

// sum / prod / amin / amax / dot / reduce_sqrm: ra/ra.hh:306-322,348-401. The accumulator has the program's result type
// (sum of float stays float, :351) and is updated sequentially. dot/reduce_sqrm are SUM over a*b / a*a, that is
// RA_FMA==0 (:364-374); with RA_FMA the program may use RACU_OP_FMA-free accumulation only, so FMA builds are covered by tolerance.
int
oracle_reduce(racu_desc const * d, int redop, void * out)
{
    if (int e = check_desc(d, false)) return e;
    if (d->dst!=-1) return fail(RACU_ERR_BAD_DESC, "reduce needs dst = -1");
    int rt = check_program(d);
    int as;
    switch (redop) {
    case RACU_RED_SUM: as = RACU_ASSIGN_ADD; break;
    case RACU_RED_PROD: as = RACU_ASSIGN_MUL; break;
    case RACU_RED_AMIN: as = RACU_ASSIGN_AMIN; break;
    case RACU_RED_AMAX: as = RACU_ASSIGN_AMAX; break;
    default: return fail(RACU_ERR_BAD_DESC, "bad redop");
    }
    if (rt==RACU_BOOL && (as==RACU_ASSIGN_ADD || as==RACU_ASSIGN_MUL)) return fail(RACU_ERR_BAD_DESC, "arithmetic reduction of bool");
    with_type(rt, [&](auto t) {
        using T = decltype(t);
        T & c = *static_cast<T *>(out);
        switch (redop) {
        case RACU_RED_SUM: c = T(0); break; // :351
        case RACU_RED_PROD: c = T(1); break; // :359
        case RACU_RED_AMIN: c = std::numeric_limits<T>::has_infinity ? std::numeric_limits<T>::infinity() : std::numeric_limits<T>::max(); break; // :310
        case RACU_RED_AMAX: c = std::numeric_limits<T>::has_infinity ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::lowest(); break; // :319
        }
        return 0;
    });
    return traverse(d, [&](Leaf const *, Val const & x) { assign(as, rt, out, 0, x); });
}

// ---------------- view algebra (host only in the reference; O(rank)) ----------------

typedef struct { int64_t len, step; } oracle_dim; // ra/expr.hh:183

// reverse(a, k): ra/arrays.hh:408-428. Returns the pointer offset in elements.
int
oracle_reverse(oracle_dim * dv, int rank, int k, int64_t * off)
{
    if (k<0 || k>=rank) return fail(RACU_ERR_BOUNDS, "Bad axis for rank."); // :417
    if (UNB==dv[k].len) return fail(RACU_ERR_BOUNDS, "Cannot reverse unbounded axis."); // :416
    *off = 0==dv[k].len ? 0 : dv[k].step*(dv[k].len-1); // revp :419
    dv[k].step *= -1; // revd :420
    return RACU_OK;
}

// transpose(a, s): axis k of the source goes to axis s[k] of the result. ra/arrays.hh:431-457.
int
oracle_transpose(oracle_dim const * src, int srank, int32_t const * s, oracle_dim * dst, int32_t * drank)
{
    int r = 0;
    for (int k=0; k<srank; ++k) {
        if (s[k]<0 || s[k]>=RACU_MAX_RANK) return fail(RACU_ERR_BOUNDS, "Bad transposed axis.");
        r = std::max(r, s[k]+1); // trank :447
    }
    for (int k=0; k<r; ++k) { dst[k] = oracle_dim {UNB, 0}; } // :436
    for (int k=0; k<srank; ++k) { // :437-441
        int sk = s[k];
        dst[sk].step += src[k].step;
        dst[sk].len = dst[sk].len>=0 ? std::min(dst[sk].len, src[k].len) : src[k].len;
    }
    *drank = r;
    return RACU_OK;
}

// Beaten subscripts: ra/ply.hh:367-435. ra::len inside a subscript is resolved by the caller (wlen, ra/ply.hh:243-295).
enum { ORACLE_SUB_SCALAR=0, ORACLE_SUB_IOTA=1, ORACLE_SUB_DOTS=2, ORACLE_SUB_INSERT=3 };
typedef struct { int32_t kind; int32_t n /* dots / insert count; -1 = dots<> stretch */; int64_t len, org, step; } oracle_sub;

static bool inside(dim_t i, dim_t b) { return 0<=i && i<b; } // ra/base.hh

int
oracle_from(oracle_dim const * a, int arank, oracle_sub const * sub, int nsub, oracle_dim * b, int32_t * brank, int64_t * off)
{
    int ak=0, bk=0;
    dim_t pl=0;
    auto fsrc = [](oracle_sub const & i) { return i.kind==ORACLE_SUB_DOTS ? i.n : i.kind==ORACLE_SUB_INSERT ? 0 : 1; }; // :344-348
    for (int q=0; q<nsub; ++q) {
        oracle_sub const & i0 = sub[q];
        switch (i0.kind) {
        case ORACLE_SUB_SCALAR: { // :390-398
            if (ak>=arank) return fail(RACU_ERR_BOUNDS, "Too many subscripts.");
            dim_t la = a[ak].len, i = i0.org;
            if (!(inside(i, la) || (UNB==la && 0==a[ak].step))) return fail(RACU_ERR_BOUNDS, "Bad index."); // :395
            pl += i*a[ak].step;
            ++ak;
            break;
        }
        case ORACLE_SUB_IOTA: { // :399-415 (rank 1 iota)
            if (ak>=arank) return fail(RACU_ERR_BOUNDS, "Too many subscripts.");
            dim_t la = a[ak].len;
            if (!(0==i0.len || (inside(i0.org, la) && inside(i0.org+(i0.len-1)*i0.step, la)))) return fail(RACU_ERR_BOUNDS, "Bad iota."); // :407
            if (bk>=RACU_MAX_RANK) return fail(RACU_ERR_BOUNDS, "Rank too large.");
            b[bk++] = oracle_dim {i0.len, i0.step*a[ak].step}; // :403
            pl += i0.org*a[ak].step; // :409
            ++ak;
            break;
        }
        case ORACLE_SUB_DOTS: { // :416-422
            int n = i0.n;
            if (n<0) { n = arank-ak; for (int r=q+1; r<nsub; ++r) { n -= (sub[r].kind==ORACLE_SUB_DOTS && sub[r].n<0) ? 0 : fsrc(sub[r]); } } // :417
            if (n<0 || ak+n>arank) return fail(RACU_ERR_BOUNDS, "Too many subscripts.");
            for (int r=0; r<n; ++r) { if (bk>=RACU_MAX_RANK) return fail(RACU_ERR_BOUNDS, "Rank too large."); b[bk++] = a[ak++]; }
            break;
        }
        case ORACLE_SUB_INSERT: // :423-427
            for (int r=0; r<i0.n; ++r) { if (bk>=RACU_MAX_RANK) return fail(RACU_ERR_BOUNDS, "Rank too large."); b[bk++] = oracle_dim {UNB, 0}; }
            break;
        default: return fail(RACU_ERR_BAD_DESC, "bad subscript kind");
        }
    }
    for (; ak<arank; ++ak) { if (bk>=RACU_MAX_RANK) return fail(RACU_ERR_BOUNDS, "Rank too large."); b[bk++] = a[ak]; } // trailing axes :374-381
    *brank = bk;
    *off = pl;
    return RACU_OK;
}

// ---------------- raw-loop variants the reference's benchmarks check their array forms against ----------------

// f_raw: bench/bench-stencil3.cc:37-52, bench-stencil2.cc:35-47, bench-stencil1.cc:30-38; examples/laplace2d.cc:36 as a loop.
// Element (i,j,k) of a rank-3 array with steps st[]; lower ranks use the leading entries.
} // extern "C"

template <class T> static void
stencil_raw(int kind, T const * A, T * An, dim_t const * n, dim_t const * sa, dim_t const * sn, dim_t lo0, dim_t hi0)
{
    dim_t i0 = std::max<dim_t>(1, lo0), i1 = std::min<dim_t>(n[0]-1, hi0);
    switch (kind) {
    case RACU_STENCIL_3D7:
        for (dim_t i=i0; i<i1; ++i) for (dim_t j=1; j+1<n[1]; ++j) for (dim_t k=1; k+1<n[2]; ++k) {
            auto a = [&](dim_t x, dim_t y, dim_t z) { return A[x*sa[0]+y*sa[1]+z*sa[2]]; };
            An[i*sn[0]+j*sn[1]+k*sn[2]] = -6*a(i, j, k) + a(i+1, j, k) + a(i, j+1, k) + a(i, j, k+1) + a(i-1, j, k) + a(i, j-1, k) + a(i, j, k-1);
        }
        break;
    case RACU_STENCIL_2D5:
        for (dim_t i=i0; i<i1; ++i) for (dim_t j=1; j+1<n[1]; ++j) {
            auto a = [&](dim_t x, dim_t y) { return A[x*sa[0]+y*sa[1]]; };
            An[i*sn[0]+j*sn[1]] = -4*a(i, j) + a(i+1, j) + a(i, j+1) + a(i-1, j) + a(i, j-1);
        }
        break;
    case RACU_STENCIL_2D5_STIFF:
        for (dim_t i=i0; i<i1; ++i) for (dim_t j=1; j+1<n[1]; ++j) {
            auto a = [&](dim_t x, dim_t y) { return A[x*sa[0]+y*sa[1]]; };
            An[i*sn[0]+j*sn[1]] = 4*a(i, j) - a(i-1, j) - a(i, j-1) - a(i+1, j) - a(i, j+1);
        }
        break;
    case RACU_STENCIL_1D3:
        for (dim_t i=i0; i<i1; ++i) { An[i*sn[0]] = -2*A[i*sa[0]] + A[(i+1)*sa[0]] + A[(i-1)*sa[0]]; }
        break;
    }
}

extern "C" int
oracle_stencil_raw(racu_stencil_desc const * s)
{
    if (!s || s->rank<1 || s->rank>3) return fail(RACU_ERR_BAD_DESC, "bad stencil rank");
    dim_t lo = s->lo0, hi = s->hi0;
    if (0==lo && 0==hi) { hi = s->len[0]; }
    if (s->dtype==RACU_F32) stencil_raw(s->kind, static_cast<float const *>(s->src), static_cast<float *>(s->dst), s->len, s->src_stride, s->dst_stride, lo, hi);
    else if (s->dtype==RACU_F64) stencil_raw(s->kind, static_cast<double const *>(s->src), static_cast<double *>(s->dst), s->len, s->src_stride, s->dst_stride, lo, hi);
    else return fail(RACU_ERR_UNSUPPORTED, "stencil dtype");
    return RACU_OK;
}

// gemm as the k-ascending triple loop that `c(all, insert<1>) += a * b(insert<1>)` performs per c[i,j]
// (ra/ra.hh:407; traversal i, k, j with j innermost; per element c += a*b, no fma on this path).
template <class T> static void
gemm_raw(dim_t M, dim_t N, dim_t K, T const * a, dim_t lda, T const * b, dim_t ldb, T * c, dim_t ldc)
{
    for (dim_t i=0; i<M; ++i) for (dim_t k=0; k<K; ++k) for (dim_t j=0; j<N; ++j) { c[i*ldc+j] += a[i*lda+k]*b[k*ldb+j]; }
}

extern "C" int
oracle_gemm_raw(racu_gemm_desc const * g)
{
    if (g->dtype==RACU_F32) gemm_raw(g->M, g->N, g->K, static_cast<float const *>(g->a), g->lda, static_cast<float const *>(g->b), g->ldb, static_cast<float *>(g->c), g->ldc);
    else if (g->dtype==RACU_F64) gemm_raw(g->M, g->N, g->K, static_cast<double const *>(g->a), g->lda, static_cast<double const *>(g->b), g->ldb, static_cast<double *>(g->c), g->ldc);
    else return fail(RACU_ERR_UNSUPPORTED, "gemm dtype");
    return RACU_OK;
}
