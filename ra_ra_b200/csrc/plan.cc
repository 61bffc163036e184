// plan.cc - the flattener's second half: racu_desc -> canonical kernel parameters.
//
// The host header (include/ra/cuda.hh) turns an expression tree into a racu_desc: leaves + postfix program, in the
// reference's own terms (per-operand rank/len/step, prefix matching). This file does what ra::ply_ravel does before its
// hot loop (ra/ply.hh:21-57) - validate, agree shapes, find what can be raveled - but for a machine where the order is
// free (docs/ra-ra.texi:2142: "shouldn't be relied upon") and the inner loop is 148 SMs wide:
//   * Match::len / check (ra/expr.hh:555-608) -> agree()
//   * the collapse rule `ss*step(z)==step(j)` (ra/ply.hh:42-46, Cell::keep ra/expr.hh:296-297) applied to ANY adjacent
//     pair after sorting the axes by destination stride (maps) or source stride (folds)
//   * negative destination steps are flipped so the innermost destination step is +1 where possible
//   * `dst OP= rhs` is folded into the program (maps), or becomes a tree fold when dst has step 0 on an axis (ra/expr.hh:50-53)
#include "plan.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace racu {

size_t dtype_size(int t) { return t==RACU_BOOL ? 1 : (t==RACU_I32 || t==RACU_F32) ? 4 : 8; }

static bool is_float(int t) { return t==RACU_F32 || t==RACU_F64; }
static int promote(int t) { return t==RACU_BOOL ? RACU_I32 : t; }
// C++ usual arithmetic conversions over the five element types.
static int common_type(int a, int b)
{
    a = promote(a); b = promote(b);
    if (a==RACU_F64 || b==RACU_F64) return RACU_F64;
    if (a==RACU_F32 || b==RACU_F32) return RACU_F32;
    if (a==RACU_I64 || b==RACU_I64) return RACU_I64;
    return RACU_I32;
}

// Type-checks the postfix program by simulating its type stack. Returns the result dtype, or -1.
static int
check_ins(racu_ins const * ins, int nins, racu_operand const * opnd, int nopnd, std::string & err, bool scatter=false)
{
    int st[RACU_MAX_INS+1], sp = 0;
    auto bad = [&](int pc, char const * what) { err = "bad program at instruction " + std::to_string(pc) + ": " + what; return -1; };
    if (nins<=0 || nins>RACU_MAX_INS) { err = "bad program length"; return -1; }
    for (int pc=0; pc<nins; ++pc) {
        racu_ins const & in = ins[pc];
        if (in.type>=RACU_NDTYPE) return bad(pc, "type");
        bool isf = is_float(in.type), isb = in.type==RACU_BOOL;
        auto need = [&](int n) { if (sp<n) return false; for (int j=0; j<n; ++j) { if (st[sp-1-j]!=in.type) return false; } return true; };
        switch (in.op) {
        case RACU_OP_PUSH:
            if (in.arg>=nopnd || opnd[in.arg].dtype!=in.type || opnd[in.arg].kind==RACU_PTR) return bad(pc, "push");
            st[sp++] = in.type; break;
        case RACU_OP_CAST:
            if (sp<1 || st[sp-1]!=in.aux || in.aux>=RACU_NDTYPE) return bad(pc, "cast");
            st[sp-1] = in.type; break;
        case RACU_OP_NEG: case RACU_OP_ABS: case RACU_OP_SQR:
            if (!need(1) || isb) return bad(pc, "unary");
            break;
        case RACU_OP_SQRT: case RACU_OP_EXP: case RACU_OP_LOG: case RACU_OP_SIN: case RACU_OP_COS:
        case RACU_OP_TAN: case RACU_OP_SINH: case RACU_OP_COSH: case RACU_OP_TANH: case RACU_OP_ASIN: case RACU_OP_ACOS: case RACU_OP_ATAN:
        case RACU_OP_EXPM1: case RACU_OP_LOG1P: case RACU_OP_LOG10:
            if (!need(1) || !isf) return bad(pc, "float unary");
            break;
        case RACU_OP_ISFINITE: case RACU_OP_ISNAN: case RACU_OP_ISINF:
            if (!need(1) || !isf) return bad(pc, "float classification");
            st[sp-1] = RACU_BOOL; break;
        case RACU_OP_CLAMP:
            if (!need(3) || isb) return bad(pc, "clamp");
            sp -= 2; break;
        case RACU_OP_LERP:
            if (!need(3) || !isf) return bad(pc, "lerp");
            sp -= 2; break;
        case RACU_OP_NOT:
            if (!need(1)) return bad(pc, "not");
            st[sp-1] = RACU_BOOL; break;
        case RACU_OP_ADD: case RACU_OP_SUB: case RACU_OP_MUL: case RACU_OP_DIV: case RACU_OP_MOD: case RACU_OP_MIN: case RACU_OP_MAX:
            if (!need(2) || isb) return bad(pc, "binary");
            --sp; break;
        case RACU_OP_POW: case RACU_OP_ATAN2:
            if (!need(2) || !isf) return bad(pc, "float binary");
            --sp; break;
        case RACU_OP_BAND: case RACU_OP_BOR: case RACU_OP_BXOR:
            if (!need(2) || isf) return bad(pc, "bit op");
            --sp; break;
        case RACU_OP_EQ: case RACU_OP_NE: case RACU_OP_LT: case RACU_OP_LE: case RACU_OP_GT: case RACU_OP_GE:
            if (!need(2)) return bad(pc, "comparison");
            --sp; st[sp-1] = RACU_BOOL; break;
        case RACU_OP_FMA:
            if (!need(3) || !isf) return bad(pc, "fma");
            sp -= 2; break;
        case RACU_OP_PICK: {
            int n = in.arg;
            if (n<1 || sp<n+1 || !need(n)) return bad(pc, "pick");
            int sel = st[sp-1-n];
            if (sel!=in.aux || is_float(sel)) return bad(pc, "pick selector");
            sp -= n; st[sp-1] = in.type; break;
        }
        case RACU_OP_LOAD: // gather: integer offset -> element of a RACU_PTR operand
            if (in.arg>=nopnd || opnd[in.arg].kind!=RACU_PTR || opnd[in.arg].dtype!=in.type) return bad(pc, "load operand");
            if (sp<1 || st[sp-1]!=in.aux || (in.aux!=RACU_I32 && in.aux!=RACU_I64)) return bad(pc, "load offset");
            st[sp-1] = in.type; break;
        default: return bad(pc, "opcode");
        }
    }
    if (scatter) { // the destination is a RACU_PTR operand: [element offset, value]
        if (2!=sp || st[0]!=RACU_I64) { err = "bad scatter program: it must leave an I64 element offset and a value"; return -1; }
        return st[1];
    }
    if (1!=sp) { err = "bad program: stack not reduced to one value"; return -1; }
    return st[0];
}

int
check_program(racu_desc const * d, std::string & err)
{
    if (d->nopnd<=0 || d->nopnd>RACU_MAX_OPND) { err = "bad operand count"; return -1; }
    for (int o=0; o<d->nopnd; ++o) {
        racu_operand const & p = d->opnd[o];
        if (p.dtype<0 || p.dtype>=RACU_NDTYPE || p.kind<RACU_MEM || p.kind>RACU_PTR) { err = "bad operand kind/dtype"; return -1; }
        if (p.kind==RACU_PTR && (0!=p.rank || p.len[0]>p.len[1] || !p.base.ptr)) { err = "bad pointer operand"; return -1; }
    }
    bool const scatter = d->dst>=0 && d->dst<d->nopnd && d->opnd[d->dst].kind==RACU_PTR;
    return check_ins(d->ins, d->nins, d->opnd, d->nopnd, err, scatter);
}

// Match::rank / len / check: ra/expr.hh:529-608.
static int64_t colen(int64_t a, int64_t b) { return a>=0 ? (a==b ? a : b>=0 ? RACU_MIS : a) : RACU_UNB==a ? b : RACU_UNB==b ? a : b; }

int
agree(racu_desc const * d, Shape & s, std::string & err)
{
    if (d->nopnd<=0 || d->nopnd>RACU_MAX_OPND) { err = "bad operand count"; return RACU_ERR_BAD_DESC; }
    s.rank = 0;
    for (int o=0; o<d->nopnd; ++o) {
        racu_operand const & p = d->opnd[o];
        if (p.rank<0 || p.rank>RACU_MAX_RANK || (p.kind==RACU_SCALAR && p.rank!=0)) { err = "bad operand rank"; return RACU_ERR_BAD_DESC; }
        s.rank = std::max(s.rank, int(p.rank));
    }
    for (int k=0; k<s.rank; ++k) {
        int64_t l = RACU_UNB;
        for (int o=0; o<d->nopnd && l!=RACU_MIS; ++o) {
            racu_operand const & p = d->opnd[o];
            if (k<p.rank) {
                if (p.len[k]<0 && p.len[k]!=RACU_UNB) { err = "bad operand len"; return RACU_ERR_BAD_DESC; }
                l = colen(p.len[k], l);
            }
        }
        if (l<0) {
            err = "Bad shapes: axis " + std::to_string(k) + (l==RACU_MIS ? " is mismatched." : " is unbounded.");
            return RACU_ERR_SHAPE;
        }
        s.len[k] = l;
    }
    return RACU_OK;
}

static KDim
make_div(int64_t len)
{
    KDim d {};
    d.len = uint32_t(len);
    uint32_t sh = 0;
    while (sh<32 && (uint64_t(1)<<sh)<uint64_t(len)) ++sh;
    d.shift = sh;
    d.mul = uint32_t(((uint64_t(1)<<32)*((uint64_t(1)<<sh)-uint64_t(len)))/uint64_t(len)+1);
    return d;
}

static uint64_t
identity_bits(int fold_op, int t)
{
    auto pack = [&](double f, int64_t i) -> uint64_t {
        switch (t) {
        case RACU_F32: return f32_bits(float(f));
        case RACU_F64: return f64_bits(f);
        case RACU_I32: return uint32_t(int32_t(i));
        case RACU_BOOL: return i ? 1 : 0;
        default: return uint64_t(i);
        }
    };
    switch (fold_op) {
    case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: return pack(0, 0);
    case RACU_ASSIGN_MUL: case RACU_ASSIGN_DIV: return pack(1, 1);
    case RACU_ASSIGN_AMIN: // ra/ra.hh:310
        return pack(std::numeric_limits<double>::infinity(), t==RACU_I32 ? std::numeric_limits<int32_t>::max() : t==RACU_BOOL ? 1 : std::numeric_limits<int64_t>::max());
    default: // AMAX ra/ra.hh:319
        return pack(-std::numeric_limits<double>::infinity(), t==RACU_I32 ? std::numeric_limits<int32_t>::lowest() : t==RACU_BOOL ? 0 : std::numeric_limits<int64_t>::lowest());
    }
}

namespace {

struct WOp // working operand
{
    int kind, dtype;
    uint64_t base;   // MEM / PTR: byte address; else value bits
    int64_t st[KMAXR];
    int64_t lo = 0, hi = 0; // PTR: valid element offsets
};

struct Work
{
    int rank = 0;
    int64_t len[KMAXR];
    std::vector<WOp> op;
    std::vector<racu_ins> ins;
    void drop(int k)
    {
        for (int j=k; j+1<rank; ++j) { len[j] = len[j+1]; for (auto & o: op) o.st[j] = o.st[j+1]; }
        --rank;
    }
    void flip(int k) // traverse axis k backwards
    {
        for (auto & o: op) {
            if (o.kind==RACU_MEM) o.base += uint64_t((len[k]-1)*o.st[k]*int64_t(dtype_size(o.dtype)));
            else if (o.kind==RACU_SEQ) { // value offset
                int64_t d = (len[k]-1)*o.st[k];
                if (is_float(o.dtype)) o.base = f64_bits(bits_f64(o.base)+double(d)); else o.base = uint64_t(int64_t(o.base)+d);
            }
            o.st[k] = -o.st[k];
        }
    }
    void swap(int a, int b)
    {
        std::swap(len[a], len[b]);
        for (auto & o: op) std::swap(o.st[a], o.st[b]);
    }
    // merge axis `outer` into axis `inner` when every operand has st[outer] == len[inner]*st[inner] (Cell::keep generalised)
    bool can_merge(int inner, int outer) const
    {
        for (auto const & o: op) { if (o.st[outer]!=len[inner]*o.st[inner]) return false; }
        return true;
    }
};

// byte range touched by a MEM operand
void
extent(Work const & w, WOp const & o, uint64_t & lo, uint64_t & hi)
{
    int64_t mn = 0, mx = 0;
    for (int k=0; k<w.rank; ++k) { int64_t d = (w.len[k]-1)*o.st[k]; if (d<0) mn += d; else mx += d; }
    int64_t es = int64_t(dtype_size(o.dtype));
    lo = o.base + uint64_t(mn*es);
    hi = o.base + uint64_t(mx*es) + uint64_t(es);
}

bool
same_view(Work const & w, WOp const & a, WOp const & b)
{
    if (a.base!=b.base || a.dtype!=b.dtype) return false;
    for (int k=0; k<w.rank; ++k) { if (a.st[k]!=b.st[k]) return false; }
    return true;
}

// Two views with the SAME steps whose bounding ranges overlap may still be disjoint (even / odd, red / black updates:
// A(iota(n/2, 0, 2)) = A(iota(n/2, 1, 2)) is legal and order independent in the reference). They share an element iff the offset
// between their origins is sum d_k*st_k with |d_k| < len_k: a bounded search over the axes, largest step first, pruned by what the
// remaining axes can still reach (non self-overlapping views leave at most two candidates per axis). Inconclusive => not disjoint.
bool
provably_disjoint(Work const & w, WOp const & a, WOp const & b)
{
    if (a.dtype!=b.dtype) return false;
    for (int k=0; k<w.rank; ++k) { if (a.st[k]!=b.st[k]) return false; }
    int64_t const es = int64_t(dtype_size(a.dtype));
    int64_t const dbytes = int64_t(a.base-b.base);
    if (0!=dbytes%es) return false;
    int ax[KMAXR], n = 0;
    for (int k=0; k<w.rank; ++k) { if (0!=a.st[k] && w.len[k]>1) ax[n++] = k; }
    for (int i=1; i<n; ++i) { for (int j=i; j>0 && std::llabs(a.st[ax[j]])>std::llabs(a.st[ax[j-1]]); --j) std::swap(ax[j], ax[j-1]); } // largest step first
    int64_t reach[KMAXR+1]; // reach[i]: largest |offset| axes i.. can produce
    reach[n] = 0;
    for (int i=n-1; i>=0; --i) reach[i] = reach[i+1] + (w.len[ax[i]]-1)*std::llabs(a.st[ax[i]]);
    int budget = 4096;
    bool inconclusive = false;
    auto rec = [&](auto && self, int i, int64_t r) -> bool { // can r be written with axes i.. ?
        if (--budget<0) { inconclusive = true; return true; }
        if (i==n) return 0==r;
        if (std::llabs(r)>reach[i]) return false;
        int64_t const st = std::llabs(a.st[ax[i]]), L = w.len[ax[i]]-1;
        // d*st must land within reach[i+1] of r
        auto fdiv = [](int64_t x, int64_t y) { int64_t q = x/y; return (x%y!=0 && ((x<0)!=(y<0))) ? q-1 : q; };
        int64_t lo = -fdiv(-(r-reach[i+1]), st) /* ceil */, hi = fdiv(r+reach[i+1], st);
        lo = std::max(lo, -L); hi = std::min(hi, L);
        for (int64_t d=lo; d<=hi; ++d) { if (self(self, i+1, r-d*st)) return true; }
        return false;
    };
    bool const hit = rec(rec, 0, dbytes/es);
    return !hit && !inconclusive;
}

void
emit(std::vector<racu_ins> & v, int op, int type, int arg=0, int aux=0)
{
    racu_ins i; i.op = uint8_t(op); i.type = uint8_t(type); i.arg = uint8_t(arg); i.aux = uint8_t(aux);
    v.push_back(i);
}
void emit_cast(std::vector<racu_ins> & v, int from, int to) { if (from!=to) emit(v, RACU_OP_CAST, to, 0, from); }

// orders axes [lo, hi) of w by key ascending (stable), key(k) computed once
template <class Key> void
sort_axes(Work & w, int lo, int hi, Key && key)
{
    for (int i=lo+1; i<hi; ++i) {
        for (int j=i; j>lo && key(j)<key(j-1); --j) w.swap(j, j-1);
    }
}

// merges adjacent axes within [lo, hi); returns new hi
int
merge_axes(Work & w, int lo, int hi)
{
    for (int k=lo; k+1<hi;) {
        if (w.can_merge(k, k+1)) { w.len[k] *= w.len[k+1]; w.drop(k+1); --hi; } else { ++k; }
    }
    return hi;
}

} // namespace

bool
can_peel(racu_desc const * d, bool reducing)
{
    if (reducing || d->dst<0) return false;
    racu_operand const & y = d->opnd[d->dst];
    if (y.kind==RACU_PTR) return d->assign!=RACU_ASSIGN_SET; // scatter: accumulating updates compose in sequence (`=` with repeats: any order anyway, but keep one launch)
    if (y.rank>=1 && y.stride[0]!=0) return true;        // independent slices of the destination
    return true;                                          // accumulating folds compose in sequence; `=` keeps the last slice (peel_last_only)
}

// `dst = expr` where dst does not move along expression axis 0: sequentially every slice overwrites the previous one ("last one
// wins" in row-major order, docs/ra-ra.texi:1986-2007), so only the last slice of the peeled axis is evaluated.
bool
peel_last_only(racu_desc const * d)
{
    if (d->dst<0 || d->assign!=RACU_ASSIGN_SET) return false;
    racu_operand const & y = d->opnd[d->dst];
    if (y.kind!=RACU_MEM || (y.rank>=1 && y.stride[0]!=0)) return false;
    // `dst = dst OP rest` (a running sum / max, test/reduction.cc:160-164) reads what the previous slice wrote: every slice counts
    for (int pc=0; pc<d->nins; ++pc) {
        if (d->ins[pc].op==RACU_OP_PUSH && d->ins[pc].arg<d->nopnd) {
            racu_operand const & x = d->opnd[d->ins[pc].arg];
            if (int(d->ins[pc].arg)==d->dst || (x.kind==RACU_MEM && x.base.ptr==y.base.ptr)) return false;
        }
    }
    return true;
}

// A whole-expression reduction with more uncollapsible axes than one launch handles: the same reduction written as the fold by
// assignment `out OP= expr` onto a rank-0 destination at reduce_out (which the caller first fills with the identity); that form
// composes slice by slice (can_peel). Returns false if the descriptor has no room for the extra operand.
bool
peel_reduction(racu_desc const & d, int redop, void * reduce_out, racu_desc & out, int & acc_type, uint64_t & identity)
{
    std::string err;
    int const rt = check_program(&d, err);
    if (rt<0 || d.nopnd>=RACU_MAX_OPND || d.dst!=-1) return false;
    int assign;
    switch (redop) {
    case RACU_RED_SUM: assign = RACU_ASSIGN_ADD; break;
    case RACU_RED_PROD: assign = RACU_ASSIGN_MUL; break;
    case RACU_RED_AMIN: assign = RACU_ASSIGN_AMIN; break;
    case RACU_RED_AMAX: assign = RACU_ASSIGN_AMAX; break;
    default: return false;
    }
    out = d;
    racu_operand & y = out.opnd[out.nopnd];
    std::memset(&y, 0, sizeof(y));
    y.kind = RACU_MEM; y.dtype = rt; y.rank = 0; y.base.ptr = reduce_out;
    out.dst = out.nopnd++;
    out.assign = assign;
    acc_type = rt;
    identity = identity_bits(assign, rt);
    return true;
}

// Slice every operand at index i of expression axis 0 (prefix matching aligns operand axes from axis 0).
void
peel_axis0(racu_desc const & d, int64_t i, racu_desc & out)
{
    out = d;
    for (int o=0; o<out.nopnd; ++o) {
        racu_operand & p = out.opnd[o];
        if (p.rank<1 || p.kind==RACU_SCALAR) continue;
        int64_t off = i*p.stride[0];
        if (p.kind==RACU_MEM) p.base.ptr = static_cast<char *>(p.base.ptr) + off*int64_t(dtype_size(p.dtype));
        else if (is_float(p.dtype)) p.base.f += double(off);
        else p.base.i += off;
        for (int k=0; k+1<p.rank; ++k) { p.len[k] = p.len[k+1]; p.stride[k] = p.stride[k+1]; }
        --p.rank;
    }
}

void
peel_axis(racu_desc const & d, int k, int64_t i, racu_desc & out)
{
    out = d;
    for (int o=0; o<out.nopnd; ++o) {
        racu_operand & p = out.opnd[o];
        if (p.rank<=k || p.kind==RACU_SCALAR || p.kind==RACU_PTR) continue;
        int64_t off = i*p.stride[k];
        if (p.kind==RACU_MEM) p.base.ptr = static_cast<char *>(p.base.ptr) + off*int64_t(dtype_size(p.dtype));
        else if (is_float(p.dtype)) p.base.f += double(off);
        else p.base.i += off;
        for (int j=k; j+1<p.rank; ++j) { p.len[j] = p.len[j+1]; p.stride[j] = p.stride[j+1]; }
        --p.rank;
    }
}

int
first_folded_axis(racu_desc const * d, Shape const & sh)
{
    if (d->dst<0 || d->dst>=d->nopnd) return -1;
    racu_operand const & y = d->opnd[d->dst];
    for (int k=0; k<sh.rank; ++k) { if (sh.len[k]>1 && (k>=y.rank || 0==y.stride[k])) return k; }
    return -1;
}

int
validate_desc(racu_desc const * d, bool reducing, std::string & err)
{
    if (!d) { err = "null descriptor"; return RACU_ERR_BAD_DESC; }
    if (d->nopnd<=0 || d->nopnd>RACU_MAX_OPND) { err = "bad operand count"; return RACU_ERR_BAD_DESC; }
    if (d->nins<=0 || d->nins>RACU_MAX_INS) { err = "bad program length"; return RACU_ERR_BAD_DESC; }
    if (reducing ? d->dst!=-1 : (d->dst<0 || d->dst>=d->nopnd)) { err = reducing ? "reduce needs dst = -1" : "bad destination"; return RACU_ERR_BAD_DESC; }
    for (int o=0; o<d->nopnd; ++o) {
        racu_operand const & p = d->opnd[o];
        if (p.rank<0 || p.rank>RACU_MAX_RANK || (p.kind==RACU_SCALAR && p.rank!=0)) { err = "bad operand rank"; return RACU_ERR_BAD_DESC; }
    }
    if (check_program(d, err)<0) return RACU_ERR_BAD_DESC;
    return RACU_OK;
}

// Does the canonical program have a compile-time specialised kernel (static_progs.h)? Fills plan.sp if so.
static void
match_static_(Plan & plan, std::vector<racu_ins> const & prog, int dst, bool fold)
{
    KParams & kp = plan.kp;
    plan.is_static = false;
    if (std::getenv("RACU_NO_STATIC") || kp.scatter) return; // (scatters run on the interpreting kernel)
    if (int(prog.size())>SP_MAXINS) return;
    bool const map = kp.mode==K_MAP;
    if (!map && kp.rankA!=1) return;
    if (map && kp.rankB!=0) return;
    if (kp.mode==K_FOLD_OUTER && kp.rankB!=1) return;
    // how the static kernels can read operand k (SV-element vectors, 16-byte accesses): from the raw strides
    auto aligned_outer = [&](KOpnd const & k) {
        int64_t const es = k.esize;
        for (int j=1; j<kp.rankA; ++j) { if (0!=(k.sA[j]*es)%16) return false; }
        for (int j=0; j<kp.rankB; ++j) { if (0!=(k.sB[j]*es)%16) return false; }
        return true;
    };
    auto kind_of = [&](KOpnd const & k) -> int {
        if (k.kind==RACU_SCALAR) return SK_SCALAR;
        if (k.kind==RACU_PTR) return SK_PTR;
        if (k.kind==RACU_SEQ) return k.dtype==RACU_BOOL ? -1 : SK_SEQ;
        if (k.kind!=RACU_MEM) return -1;
        int64_t const es = k.esize;
        if (1==k.sA[0]) return (0==k.base%16 && aligned_outer(k)) ? SK_VEC : (map && 0==k.base%es) ? SK_VECU : -1; // unaligned rows: the lane-wise map kernel
        if (0==k.sA[0]) return SK_ROW;
        if (-1==k.sA[0] && map) return (0==(k.base-uint64_t((SV-1)*es))%16 && aligned_outer(k)) ? SK_VECREV : 0==k.base%es ? SK_VECREVU : -1;
        return -1;
    };
    // roles by first appearance
    int role_of[RACU_MAX_OPND], nrole = 0, yrole = -1;
    for (int o=0; o<RACU_MAX_OPND; ++o) role_of[o] = -1;
    SIns seq[SP_MAXINS];
    uint8_t rtype[SP_MAXIN] = {};
    for (size_t pc=0; pc<prog.size(); ++pc) {
        racu_ins const & i = prog[pc];
        if (i.op==RACU_OP_PUSH) {
            if (role_of[i.arg]<0) {
                if (nrole>=SP_MAXIN) return;
                KOpnd const & k = kp.opnd[i.arg];
                int kd = kind_of(k);
                if (kd<0 || (!map && kd==SK_VECREV)) return;
                if (!fold && int(i.arg)==dst) { if (kd!=SK_VEC && kd!=SK_VECU) return; yrole = nrole; }
                rtype[nrole] = k.dtype;
                role_of[i.arg] = nrole++;
            }
            seq[pc] = SIns {RACU_OP_PUSH, uint8_t(role_of[i.arg]), i.type, 0};
        } else if (i.op==RACU_OP_LOAD) { // the gathered array is a role of kind SK_PTR (never fetched as a vector)
            if (role_of[i.arg]<0) {
                if (nrole>=SP_MAXIN) return;
                rtype[nrole] = kp.opnd[i.arg].dtype;
                role_of[i.arg] = nrole++;
            }
            seq[pc] = SIns {RACU_OP_LOAD, uint8_t(role_of[i.arg]), i.type, i.aux};
        } else {
            seq[pc] = SIns {i.op, uint8_t(i.op==RACU_OP_PICK ? i.arg : 0), i.type, uint8_t(i.op==RACU_OP_CAST || i.op==RACU_OP_PICK ? i.aux : 0)};
        }
    }
    int const otype = fold ? kp.acc_type : kp.opnd[dst].dtype;
    int id = -1;
    for (int c=0; c<SP_COUNT && id<0; ++c) {
        SProg const P = static_prog(c);
        if (P.n!=int(prog.size()) || P.nin!=nrole || P.yrole!=yrole || P.fold!=fold || P.otype!=otype) continue;
        bool same = true;
        for (int r=0; r<nrole && same; ++r) same = P.rtype[r]==rtype[r];
        for (int pc=0; pc<P.n && same; ++pc) same = P.ins[pc].op==seq[pc].op && P.ins[pc].type==seq[pc].type && P.ins[pc].aux==seq[pc].aux && P.ins[pc].arg==seq[pc].arg;
        if (same) id = c;
    }
    if (id<0) { // not in the table: specialise at run time (jit.cc) if the library carries the specialiser
        if (!g_jit_prepare || std::getenv("RACU_NO_JIT")) return;
        SProg & P = plan.jit_prog;
        P = SProg {};
        P.n = int(prog.size()); P.nin = nrole; P.yrole = yrole; P.fold = fold; P.otype = uint8_t(otype);
        for (int pc=0; pc<P.n; ++pc) P.ins[pc] = seq[pc];
        for (int r=0; r<nrole; ++r) P.rtype[r] = rtype[r];
        id = SP_JIT;
    }
    SParams & sp = plan.sp;
    std::memset(&sp, 0, sizeof(sp));
    int64_t const nvec0 = kp.len0/SV;
    int const target = 148*8;
    bool const pwide = prog_wide(SP_JIT==id ? plan.jit_prog : static_prog(id));
    auto knob = [](char const * name, int64_t dflt) { char const * e = std::getenv(name); return e ? int64_t(std::atoll(e)) : dflt; }; // tuning knobs
    if (fold) {
        if (kp.fold_op==RACU_ASSIGN_SET) return;
        if ((kp.nk>1 || kp.mode==K_FOLD_OUTER) && 0!=kp.len0%SV) return;
        if (kp.mode==K_FOLD_INNER && kp.nk>=knob("RACU_ROWS_MINROWS", 148*16) && nvec0<=knob("RACU_ROWS_MAXVEC", 4096)) {
            // many kept rows of moderate length (gemv, sum-cols): one warp per row, no chunks, no finish pass
            if ((kp.nk+KBLOCK/32-1)/(KBLOCK/32)>=(int64_t(1)<<31)) return;
            sp.inner_rows = 1;
            kp.nchunks = 1; kp.chunk_len = nvec0;
            plan.grid_x = unsigned((kp.nk+KBLOCK/32-1)/(KBLOCK/32)); plan.grid_y = 1;
        } else if (kp.mode==K_FOLD_INNER) { // chunk the row in units of SV-element vectors
            int64_t per = int64_t(KBLOCK)*4; // upper bound of the kernel's vectors per iteration
            int64_t maxchunks = std::max<int64_t>(1, nvec0/(per*2));
            int64_t want = std::max<int64_t>(1, (int64_t(target)*2)/std::max<int64_t>(1, kp.nk));
            int64_t nch = std::min(maxchunks, want);
            int64_t cl = std::max<int64_t>(1, (nvec0+nch-1)/nch);
            cl = ((cl+per-1)/per)*per;
            nch = std::max<int64_t>(1, (nvec0+cl-1)/cl);
            if (kp.nk*nch>=(int64_t(1)<<31)) return;
            kp.nchunks = int(nch); kp.chunk_len = cl;
            plan.grid_x = unsigned(kp.nk*nch); plan.grid_y = 1;
        } else {
            // geometry below (outer_geometry), once the kernel is known
            int64_t const per = 32*(pwide ? 1 : 2);
            int64_t nbx = (nvec0+per-1)/per;
            if (nbx<1 || nbx>=(int64_t(1)<<31)) return;
            kp.nchunks = 1; kp.chunk_len = kp.nB;
            plan.grid_x = unsigned(nbx); plan.grid_y = 1;
        }
        plan.fp.nchunks = kp.nchunks;
        plan.need_finish = kp.nchunks>1;
        plan.partial_bytes = plan.need_finish ? size_t(kp.nchunks)*size_t(kp.nk)*(plan.fp.wide ? 8 : 4) : 0;
        sp.wide_partials = plan.fp.wide;
    } else {
        KOpnd const & y = kp.opnd[dst];
        if (1!=y.sA[0] || 0!=y.base%y.esize) return;
        // rows that are not 16-byte aligned (destination or any input) or not whole vectors: the lane-wise kernel (smap_lanes_body)
        bool lanes = 0!=y.base%16 || !aligned_outer(y) || (kp.rankA>1 && 0!=kp.len0%SV);
        for (int o=0; o<kp.nopnd; ++o) { if (role_of[o]>=0) { int const kd = kind_of(kp.opnd[o]); lanes = lanes || SK_VECU==kd || SK_VECREVU==kd; } }
        sp.lanes = lanes;
        int64_t const nvrow = lanes ? ((kp.len0+32*SV-1)/(32*SV))*32 : nvec0; // lane-wise: rows padded to whole warp tiles of 32*SV elements
        sp.dst = reinterpret_cast<void *>(uintptr_t(y.base));
        sp.vpr = nvrow;
        sp.nvec = nvrow;
        for (int j=1; j<kp.rankA; ++j) sp.nvec *= kp.dA[j].len;
        if (kp.rankA>1) { if (nvrow<1 || sp.nvec>=(int64_t(1)<<31)) return; sp.vprdiv = make_div(nvrow); }
        for (int j=0; j<KGR; ++j) { sp.dA[j] = kp.dA[j]; sp.dsA[j] = j<kp.rankA ? y.sA[j] : 0; }
        int64_t nb = (sp.nvec+int64_t(KBLOCK)*4-1)/(int64_t(KBLOCK)*4);
        if (lanes) { // work units of the lane-wise kernel: patches of 8 rows x SLANES_CT warp tiles (rows), or 8*SLANES_CT consecutive tiles (flat)
            int64_t const wpr = nvrow/32, ppb = (wpr+SLANES_CT-1)/SLANES_CT, nrows = kp.rankA>1 ? sp.nvec/nvrow : 1;
            nb = kp.rankA>1 ? ((nrows+7)/8)*ppb : (wpr+8*SLANES_CT-1)/(8*SLANES_CT);
            if (nb>=(int64_t(1)<<31)) return;
            if (kp.rankA>1) sp.vprdiv = make_div(ppb); // the kernel divides unit ids by the patches per band, not vectors by the row length
        }
        plan.grid_x = unsigned(std::max<int64_t>(1, std::min<int64_t>(nb, int64_t(target)*4))); plan.grid_y = 1;
    }
    sp.id = id; sp.mode = kp.mode; sp.fold_op = kp.fold_op; sp.nchunks = kp.nchunks; sp.rankB = kp.rankB; sp.rankA = kp.rankA;
    sp.identity = kp.identity; sp.n = kp.len0; sp.nB = kp.nB; sp.chunk_len = kp.chunk_len; sp.nk = kp.nk;
    for (int j=0; j<KGR; ++j) sp.dB[j] = kp.dB[j];
    for (int o=0; o<kp.nopnd; ++o) {
        int r = role_of[o];
        if (r<0) continue;
        KOpnd const & k = kp.opnd[o];
        sp.kind[r] = uint8_t(kind_of(k));
        if (sp.kind[r]==SK_VECREV || sp.kind[r]==SK_VECU || sp.kind[r]==SK_VECREVU || sp.kind[r]==SK_SEQ || sp.kind[r]==SK_PTR || (map && sp.kind[r]==SK_ROW)) sp.general = 1;
        if (k.kind==RACU_SCALAR) { // value bits in the role's type
            switch (k.dtype) {
            case RACU_F32: sp.in[r] = f32_bits(float(bits_f64(k.base))); break;
            case RACU_I32: sp.in[r] = uint32_t(int32_t(int64_t(k.base))); break;
            case RACU_BOOL: sp.in[r] = k.base!=0; break;
            default: sp.in[r] = k.base; break;
            }
        } else {
            sp.in[r] = k.base;
        }
        for (int j=0; j<KGR; ++j) { sp.sB[r][j] = j<kp.rankB ? k.sB[j] : 0; sp.sA[r][j] = j<kp.rankA ? k.sA[j] : 0; }
        if (k.kind==RACU_PTR) { sp.sA[r][0] = k.sA[0]; sp.sA[r][1] = k.sA[1]; } // valid offsets of a gathered array
    }
    if (fold && kp.mode==K_FOLD_INNER && kp.rankB>=1 && kp.nk>1) {
        for (int r=0; r<nrole; ++r) {
            bool moves = false;
            for (int j=0; j<kp.rankB; ++j) moves = moves || 0!=sp.sB[r][j];
            sp.reused[r] = sp.kind[r]==SK_VEC && !moves;
        }
    }
    if (SP_JIT==id) {
        std::string err;
        if (!g_jit_prepare(plan, err)) { if (std::getenv("RACU_JIT_VERBOSE")) std::fprintf(stderr, "racu jit: %s\n", err.c_str()); return; }
    }
    if (fold && kp.mode==K_FOLD_OUTER) {
        // A CTA owns 32*KV kept vectors (SPO<ID>::TILE) and one chunk of the folded rows, which its 8 warps share (UNR rows each per turn).
        // Chunks cost a partials pass, and these CTAs all take the same time, so the grid is sized in WHOLE WAVES of the kernel's own
        // occupancy: measured on 16384^2, gevm f32 (3 CTAs / SM) ran at 0.97 of the HBM peak with 888 CTAs and 0.90 with 1184, gevm f64
        // (4 / SM) at 1.02 with 1184 and 0.89 with 888.
        int64_t const nbx = plan.grid_x;
        int occ = g_static_occupancy ? g_static_occupancy(plan) : 0;
        if (occ<=0) occ = 3;
        int64_t const slots = int64_t(148)*occ;
        int64_t const maxchunks = std::min<int64_t>(std::max<int64_t>(1, kp.nB/64), 65535);
        int64_t nch = 1;
        if (int64_t forced = knob("RACU_OUTER_CTAS", 0)) nch = std::max<int64_t>(1, forced/nbx);
        else {
            double best = -1;
            for (int64_t c = 1; c<=maxchunks && nbx*c<=slots*3; ++c) { // up to three waves; prefer fuller last waves, then more CTAs per slot
                int64_t const ctas = nbx*c, waves = (ctas+slots-1)/slots;
                double const eff = double(ctas)/double(waves*slots) - 0.02/double(waves); // a single wave has no one to hide its start / drain behind
                if (eff>best) { best = eff; nch = c; }
            }
        }
        nch = std::min(nch, maxchunks);
        int64_t cl = (kp.nB+nch-1)/nch;
        cl = ((cl+31)/32)*32;
        nch = (kp.nB+cl-1)/cl;
        kp.nchunks = int(nch); kp.chunk_len = cl;
        plan.grid_y = unsigned(nch);
        plan.fp.nchunks = kp.nchunks;
        plan.need_finish = kp.nchunks>1;
        plan.partial_bytes = plan.need_finish ? size_t(kp.nchunks)*size_t(kp.nk)*(plan.fp.wide ? 8 : 4) : 0;
        sp.nchunks = kp.nchunks; sp.chunk_len = kp.chunk_len;
    }
    plan.is_static = true;
    static const char * names[6] = {"sflat_map", "sflat_fold_inner", "sflat_fold_outer", "jit_map", "jit_fold_inner", "jit_fold_outer"};
    plan.kernel = names[kp.mode + (SP_JIT==id ? 3 : 0)];
}

bool (*g_jit_prepare)(Plan &, std::string &) = nullptr;
int (*g_static_occupancy)(Plan const &) = nullptr;

// The static / run-time specialised kernels use their own launch geometry: if no specialisation applies in the end (or the
// specialiser fails), the interpreter's plan must come back untouched.
static void
match_static(Plan & plan, std::vector<racu_ins> const & prog, int dst, bool fold)
{
    Plan const saved = plan;
    match_static_(plan, prog, dst, fold);
    if (!plan.is_static) plan = saved;
}

int
make_plan(racu_desc const * d, int redop, void * reduce_out, Plan & plan, std::string & err)
{
    if (!d) { err = "null descriptor"; return RACU_ERR_BAD_DESC; }
    int rt = check_program(d, err);
    if (rt<0) return RACU_ERR_BAD_DESC;
    Shape sh;
    if (int e = agree(d, sh, err)) return e;

    bool const reducing = reduce_out!=nullptr;
    bool const scatter = !reducing && d->dst>=0 && d->dst<d->nopnd && d->opnd[d->dst].kind==RACU_PTR; // a(i ...) OP= x with index arrays
    int assign = d->assign;
    if (reducing) {
        if (d->dst!=-1) { err = "reduce needs dst = -1"; return RACU_ERR_BAD_DESC; }
        switch (redop) {
        case RACU_RED_SUM: assign = RACU_ASSIGN_ADD; break;
        case RACU_RED_PROD: assign = RACU_ASSIGN_MUL; break;
        case RACU_RED_AMIN: assign = RACU_ASSIGN_AMIN; break;
        case RACU_RED_AMAX: assign = RACU_ASSIGN_AMAX; break;
        default: err = "bad redop"; return RACU_ERR_BAD_DESC;
        }
        if (rt==RACU_BOOL && (assign==RACU_ASSIGN_ADD || assign==RACU_ASSIGN_MUL)) { err = "arithmetic reduction of bool"; return RACU_ERR_BAD_DESC; }
    } else {
        if (d->dst<0 || d->dst>=d->nopnd || (d->opnd[d->dst].kind!=RACU_MEM && d->opnd[d->dst].kind!=RACU_PTR)) { err = "bad destination"; return RACU_ERR_BAD_DESC; }
        if (assign<RACU_ASSIGN_SET || assign>RACU_ASSIGN_AMAX) { err = "bad assign op"; return RACU_ERR_BAD_DESC; }
        int dt = d->opnd[d->dst].dtype;
        if (assign>=RACU_ASSIGN_ADD && assign<=RACU_ASSIGN_DIV && (dt==RACU_BOOL || rt==RACU_BOOL)) { err = "arithmetic assign on bool"; return RACU_ERR_BAD_DESC; }
        if ((assign==RACU_ASSIGN_AMIN || assign==RACU_ASSIGN_AMAX) && dt!=rt) { err = "amin/amax need matching types"; return RACU_ERR_BAD_DESC; }
    }

    // ---- working copy: expression axes, outermost first as in the reference
    Work w;
    w.rank = sh.rank;
    for (int k=0; k<sh.rank; ++k) w.len[k] = sh.len[k];
    for (int o=0; o<d->nopnd; ++o) {
        racu_operand const & p = d->opnd[o];
        WOp x {};
        x.kind = p.kind; x.dtype = p.dtype;
        if (p.kind==RACU_MEM || p.kind==RACU_PTR) x.base = uint64_t(uintptr_t(p.base.ptr));
        else if (is_float(p.dtype)) x.base = f64_bits(p.base.f);
        else x.base = uint64_t(p.base.i);
        for (int k=0; k<KMAXR; ++k) x.st[k] = (k<p.rank && p.kind!=RACU_SCALAR) ? p.stride[k] : 0; // Cell::step ra/expr.hh:294-295
        if (p.kind==RACU_PTR) { x.lo = p.len[0]; x.hi = p.len[1]; }
        w.op.push_back(x);
    }
    w.ins.assign(d->ins, d->ins+d->nins);
    int dst = d->dst;
    if (reducing) { // synthetic rank-0 destination holding the result
        if (int(w.op.size())>=RACU_MAX_OPND) { err = "too many operands"; return RACU_ERR_UNSUPPORTED; }
        WOp x {};
        x.kind = RACU_MEM; x.dtype = rt; x.base = uint64_t(uintptr_t(reduce_out));
        w.op.push_back(x);
        dst = int(w.op.size())-1;
    }
    int const D = w.op[dst].dtype;

    // ---- empty: the op never runs (ra/ply.hh:47-56)
    for (int k=0; k<w.rank; ++k) {
        if (0==w.len[k]) {
            plan.empty = true;
            plan.fill_identity = reducing;
            plan.kp.identity = identity_bits(assign, rt);
            plan.kp.acc_type = rt;
            return RACU_OK;
        }
    }
    for (auto const & o: w.op) { if (o.kind==RACU_MEM && 0==o.base) { err = "null operand pointer"; return RACU_ERR_BAD_DESC; } }
    // ---- drop axes of length 1
    for (int k=0; k<w.rank;) { if (1==w.len[k]) w.drop(k); else ++k; }

    // ---- destination with step 0 on a live axis => fold over those axes
    bool red[KMAXR];
    int nred = 0;
    for (int k=0; k<w.rank; ++k) { red[k] = !scatter && 0==w.op[dst].st[k]; nred += red[k]; }
    bool fold = nred>0 || reducing;

    // aliasing between dst and sources
    bool alias_same = false, alias_other = false;
    {
        uint64_t dlo, dhi; extent(w, w.op[dst], dlo, dhi);
        if (scatter) { uint64_t const es = dtype_size(D); dlo = w.op[dst].base + uint64_t(w.op[dst].lo*int64_t(es)); dhi = w.op[dst].base + uint64_t((w.op[dst].hi+1)*int64_t(es)); }
        for (int o=0; o<int(w.op.size()); ++o) {
            if (o!=dst && w.op[o].kind==RACU_PTR) { // a gathered array: everything it may address
                uint64_t const es = dtype_size(w.op[o].dtype), lo = w.op[o].base + uint64_t(w.op[o].lo*int64_t(es)), hi = w.op[o].base + uint64_t((w.op[o].hi+1)*int64_t(es));
                if (lo<dhi && dlo<hi) alias_other = true;
                continue;
            }
            if (o==dst || w.op[o].kind!=RACU_MEM) continue;
            uint64_t lo, hi; extent(w, w.op[o], lo, hi);
            if (lo<dhi && dlo<hi) {
                if (same_view(w, w.op[o], w.op[dst])) alias_same = true;
                else if (!scatter && provably_disjoint(w, w.op[o], w.op[dst])) {} // interleaved, never the same element
                else alias_other = true;
            }
        }
    }
    for (int pc=0; pc<d->nins; ++pc) { if (d->ins[pc].op==RACU_OP_PUSH && int(d->ins[pc].arg)==dst) alias_same = true; } // the program reads the destination operand itself
    if (alias_other) { err = "destination overlaps a source that is not the same view: traversal order would be observable"; return RACU_ERR_UNSUPPORTED; }

    if (fold && assign==RACU_ASSIGN_SET) {
        if (!alias_same) {
            // "last one wins" in row-major order (docs/ra-ra.texi:1986-2007): evaluate at the last index of the folded axes
            for (int k=0; k<w.rank;) {
                if (red[k]) {
                    for (auto & o: w.op) {
                        if (o.kind==RACU_MEM) o.base += uint64_t((w.len[k]-1)*o.st[k]*int64_t(dtype_size(o.dtype)));
                        else if (o.kind==RACU_SEQ) { int64_t dd = (w.len[k]-1)*o.st[k]; if (is_float(o.dtype)) o.base = f64_bits(bits_f64(o.base)+double(dd)); else o.base = uint64_t(int64_t(o.base)+dd); }
                    }
                    for (int j=k; j+1<w.rank; ++j) red[j] = red[j+1];
                    w.drop(k);
                } else ++k;
            }
            fold = false;
        } else {
            // dst = dst OP rest  (test/reduction.cc:160-164,230-234): running sum / product / max / min
            // program must be: PUSH dst' [CAST] <rest> OP [CAST]   with dst' the same view as dst and no other use of it
            std::vector<racu_ins> & in = w.ins;
            int n = int(in.size());
            int last = n-1;
            if (last>=0 && in[last].op==RACU_OP_CAST) --last;
            int newop = -1;
            if (last>=1) {
                switch (in[last].op) {
                case RACU_OP_ADD: newop = RACU_ASSIGN_ADD; break;
                case RACU_OP_MUL: newop = RACU_ASSIGN_MUL; break;
                case RACU_OP_MAX: newop = RACU_ASSIGN_AMAX; break;
                case RACU_OP_MIN: newop = RACU_ASSIGN_AMIN; break;
                }
            }
            int first = 0;
            bool ok = newop>=0 && in[0].op==RACU_OP_PUSH && same_view(w, w.op[in[0].arg], w.op[dst]);
            if (ok) {
                first = 1;
                if (in[1].op==RACU_OP_CAST) first = 2;
                for (int pc=first; pc<last; ++pc) { if (in[pc].op==RACU_OP_PUSH && w.op[in[pc].arg].kind==RACU_MEM && same_view(w, w.op[in[pc].arg], w.op[dst])) ok = false; }
                // <rest> must be a complete subexpression: its stack effect is exactly +1
                std::string e2;
                std::vector<racu_operand> tmp(w.op.size());
                for (size_t o=0; o<w.op.size(); ++o) tmp[o].dtype = w.op[o].dtype;
                if (ok && check_ins(in.data()+first, last-first, tmp.data(), int(tmp.size()), e2)<0) ok = false;
            }
            if (!ok) { err = "assignment to a broadcast destination that reads itself is only supported as dst = dst OP expr with OP in + * max min"; return RACU_ERR_UNSUPPORTED; }
            if ((newop==RACU_ASSIGN_AMIN || newop==RACU_ASSIGN_AMAX) && in[last].type!=D) { err = "running max/min needs matching types"; return RACU_ERR_UNSUPPORTED; }
            std::vector<racu_ins> rest(in.begin()+first, in.begin()+last);
            w.ins = rest;
            assign = newop;
            std::string e2;
            std::vector<racu_operand> tmp(w.op.size());
            for (size_t o=0; o<w.op.size(); ++o) tmp[o].dtype = w.op[o].dtype;
            rt = check_ins(w.ins.data(), int(w.ins.size()), tmp.data(), int(tmp.size()), e2);
        }
    }
    if (fold && alias_same && assign!=RACU_ASSIGN_SET && d->assign!=RACU_ASSIGN_SET) {
        // c += f(c, a) with broadcast c: sequential dependence through c
        for (auto const & i: w.ins) {
            if (i.op==RACU_OP_PUSH && w.op[i.arg].kind==RACU_MEM && (int(i.arg)==dst || same_view(w, w.op[i.arg], w.op[dst]))) {
                err = "fold whose right hand side reads the destination"; return RACU_ERR_UNSUPPORTED;
            }
        }
    }

    KParams & kp = plan.kp;
    std::memset(&kp, 0, sizeof(kp));

    // ---- program: fold `dst OP= rhs` into it (maps) or end in the accumulator type (folds)
    std::vector<racu_ins> prog;
    int A = rt; // accumulator type
    if (scatter) { // the kernel applies `a[off] OP= value` itself (atomically: offsets may repeat); value in the common type, C++ style
        A = (assign==RACU_ASSIGN_SET || assign==RACU_ASSIGN_AMIN || assign==RACU_ASSIGN_AMAX) ? D : common_type(D, rt);
        prog = w.ins;
        emit_cast(prog, rt, A);
    } else if (!fold) {
        if (assign==RACU_ASSIGN_SET) {
            prog = w.ins;
            emit_cast(prog, rt, D);
        } else {
            int C = (assign==RACU_ASSIGN_AMIN || assign==RACU_ASSIGN_AMAX) ? D : common_type(D, rt);
            emit(prog, RACU_OP_PUSH, D, dst);
            emit_cast(prog, D, C);
            prog.insert(prog.end(), w.ins.begin(), w.ins.end());
            emit_cast(prog, rt, C);
            int op = assign==RACU_ASSIGN_ADD ? RACU_OP_ADD : assign==RACU_ASSIGN_SUB ? RACU_OP_SUB : assign==RACU_ASSIGN_MUL ? RACU_OP_MUL
                : assign==RACU_ASSIGN_DIV ? RACU_OP_DIV : assign==RACU_ASSIGN_AMIN ? RACU_OP_MIN : RACU_OP_MAX;
            emit(prog, op, C);
            emit_cast(prog, C, D);
        }
    } else {
        A = (assign==RACU_ASSIGN_AMIN || assign==RACU_ASSIGN_AMAX) ? rt : common_type(D, rt);
        if ((assign==RACU_ASSIGN_AMIN || assign==RACU_ASSIGN_AMAX) && rt!=D) { err = "amin/amax need matching types"; return RACU_ERR_BAD_DESC; }
        // for_each(y OP= x) narrows to the destination at EVERY step (ra/expr.hh:50-53). A tree that accumulates in the common type and
        // narrows once is the same computation when the narrowing commutes with the op: same type, integer -> narrower integer (modular),
        // or floating point -> narrower floating point (a rounding difference, within the fold tolerance: the order is unspecified anyway).
        // It is NOT for an integer destination with a floating point right hand side (int c += 0.5 four times is 0, not 2): those run
        // the folded axes one index at a time.
        if (!reducing && is_float(A) && !is_float(D)) { err = "fold that truncates to an integer destination at every step"; return RACU_ERR_SEQ; }
        prog = w.ins;
        emit_cast(prog, rt, A);
    }
    if (int(prog.size())>RACU_MAX_INS) { err = "program too long"; return RACU_ERR_UNSUPPORTED; }

    // ---- axis order. Work axes are outermost-first; build fastest-first.
    // reverse to fastest-first
    for (int a=0, b=w.rank-1; a<b; ++a, --b) { w.swap(a, b); std::swap(red[a], red[b]); }
    int nA = 0, nB = 0; // group sizes after ordering: axes [0,nA) = group A, [nA, nA+nB) = group B
    int mode = K_MAP;
    if (scatter) { // no destination strides: order the axes by the source that moves along most of them
        int dom = -1, best = -1;
        for (int o=0; o<int(w.op.size()); ++o) {
            if (w.op[o].kind!=RACU_MEM) continue;
            int c = 0; for (int k=0; k<w.rank; ++k) c += w.op[o].st[k]!=0;
            if (c>best) { best = c; dom = o; }
        }
        if (dom>=0) {
            for (int k=0; k<w.rank; ++k) { if (w.op[dom].st[k]<0) w.flip(k); }
            sort_axes(w, 0, w.rank, [&](int k) -> int64_t { int64_t s = w.op[dom].st[k]; return 0==s ? std::numeric_limits<int64_t>::max() : s; });
        }
        nA = merge_axes(w, 0, w.rank);
        nB = 0;
        if (0==nA) { w.len[0] = 1; for (auto & o: w.op) o.st[0] = 0; w.rank = 1; nA = 1; }
    } else if (!fold) {
        for (int k=0; k<w.rank; ++k) { if (w.op[dst].st[k]<0) w.flip(k); }
        sort_axes(w, 0, w.rank, [&](int k) { return w.op[dst].st[k]; });
        nA = merge_axes(w, 0, w.rank);
        nB = 0;
        if (0==nA) { w.len[0] = 1; for (auto & o: w.op) o.st[0] = 0; w.rank = 1; nA = 1; } // rank 0: one element
    } else {
        // dominant source: the MEM operand (not dst) moving along most axes
        int dom = -1, best = -1;
        for (int o=0; o<int(w.op.size()); ++o) {
            if (o==dst || w.op[o].kind!=RACU_MEM) continue;
            int c = 0; for (int k=0; k<w.rank; ++k) c += w.op[o].st[k]!=0;
            if (c>best) { best = c; dom = o; }
        }
        auto skey = [&](int k) -> int64_t {
            int64_t s = dom>=0 ? w.op[dom].st[k] : 0;
            if (0==s) s = w.op[dst].st[k];
            if (0==s) return std::numeric_limits<int64_t>::max();
            return s<0 ? -s : s;
        };
        for (int k=0; k<w.rank; ++k) { int64_t s = dom>=0 && w.op[dom].st[k]!=0 ? w.op[dom].st[k] : w.op[dst].st[k]; if (s<0) w.flip(k); }
        // kept axes first, then folded ones; each sorted fastest-first
        int nk = 0;
        for (int k=0; k<w.rank; ++k) {
            if (!red[k]) { for (int j=k; j>nk; --j) { w.swap(j, j-1); std::swap(red[j], red[j-1]); } ++nk; }
        }
        sort_axes(w, 0, nk, skey);
        sort_axes(w, nk, w.rank, skey);
        int nr = w.rank-nk;
        int nk2 = merge_axes(w, 0, nk);
        int nr2 = merge_axes(w, nk2, nk2+nr) - nk2;
        nk = nk2; nr = nr2;
        int64_t NK = 1, NR = 1;
        for (int k=0; k<nk; ++k) NK *= w.len[k];
        for (int k=nk; k<nk+nr; ++k) NR *= w.len[k];
        bool wide_guess = false;
        for (auto const & o: w.op) wide_guess |= dtype_size(o.dtype)==8;
        int V = wide_guess ? 2 : 4;
        int64_t nkvec = nk>0 ? ((w.len[0]+V-1)/V)*(NK/w.len[0]) : 1;
        bool red_fastest = nr>0 && (0==nk || skey(nk)<=skey(0));
        bool inner = 0==nk || nkvec<64 || (red_fastest && NR>=256);
        if (0==nr) { // reducing over nothing (rank 0 expression, or every folded axis had length 1): one element per kept
            w.len[w.rank] = 1; for (auto & o: w.op) o.st[w.rank] = 0; ++w.rank; nr = 1; inner = false; if (0==nk) inner = true;
        }
        if (inner) {
            // group A = folded axes, group B = kept axes: rotate folded axes to the front
            Work r = w;
            for (int k=0; k<nr; ++k) { r.len[k] = w.len[nk+k]; for (size_t o=0; o<w.op.size(); ++o) r.op[o].st[k] = w.op[o].st[nk+k]; }
            for (int k=0; k<nk; ++k) { r.len[nr+k] = w.len[k]; for (size_t o=0; o<w.op.size(); ++o) r.op[o].st[nr+k] = w.op[o].st[k]; }
            w = r;
            nA = nr; nB = nk; mode = K_FOLD_INNER;
        } else {
            nA = nk; nB = nr; mode = K_FOLD_OUTER;
        }
        kp.nk = NK;
    }

    // ---- slot width
    bool wide = dtype_size(A)==8;
    for (auto const & o: w.op) wide |= dtype_size(o.dtype)==8;
    for (auto const & i: prog) wide |= dtype_size(i.type)==8 || (i.op==RACU_OP_CAST && dtype_size(i.aux)==8) || ((i.op==RACU_OP_PICK || i.op==RACU_OP_LOAD) && dtype_size(i.aux)==8);
    int const V = wide ? 2 : 4;

    if (nA>KGR || nB>KGR) { err = "more than 4 uncollapsible axes in one group"; return RACU_ERR_PEEL; }
    // ---- limits of the 32-bit index decomposition
    for (int k=(1==nA ? 1 : 0); k<nA+nB; ++k) { if (w.len[k]>=(int64_t(1)<<31)) { err = "axis longer than 2^31 after collapsing"; return RACU_ERR_UNSUPPORTED; } }
    kp.mode = mode; kp.wide = wide; kp.rankA = nA; kp.rankB = nB;
    kp.len0 = w.len[0];
    kp.vpr = (w.len[0]+V-1)/V;
    kp.nvecA = kp.vpr;
    for (int k=1; k<nA; ++k) kp.nvecA *= w.len[k];
    kp.nB = 1;
    for (int k=nA; k<nA+nB; ++k) kp.nB *= w.len[k];
    if ((nA>1 && kp.nvecA>=(int64_t(1)<<31)) || kp.nB>=(int64_t(1)<<31)) { err = "more than 2^31 rows in one traversal"; return RACU_ERR_UNSUPPORTED; }
    if (nA>1) kp.vprdiv = make_div(kp.vpr);
    for (int k=1; k<nA; ++k) kp.dA[k] = make_div(w.len[k]);
    for (int k=0; k<nB; ++k) kp.dB[k] = make_div(w.len[nA+k]);

    // ---- operands: staging mode and slot
    kp.nopnd = int(w.op.size());
    kp.dst = dst;
    kp.scatter = scatter;
    int nstage = 0, nasync = 0;
    bool used[RACU_MAX_OPND] = {};
    for (auto const & i: prog) { if (i.op==RACU_OP_PUSH) used[i.arg] = true; }
    for (int o=0; o<kp.nopnd; ++o) {
        WOp const & x = w.op[o];
        KOpnd & k = kp.opnd[o];
        k.base = x.base; k.kind = uint8_t(x.kind); k.dtype = uint8_t(x.dtype); k.esize = uint8_t(dtype_size(x.dtype));
        k.slot = 255;
        for (int j=0; j<nA; ++j) k.sA[j] = x.st[j];
        for (int j=0; j<nB; ++j) k.sB[j] = x.st[nA+j];
        if (x.kind==RACU_SCALAR) { k.stage = S_NONE; continue; }
        if (x.kind==RACU_PTR) { k.stage = S_NONE; k.sA[0] = x.lo; k.sA[1] = x.hi; continue; } // addressed by RACU_OP_LOAD only
        if (x.kind==RACU_SEQ) { k.stage = S_SEQ; continue; }
        if (0==x.st[0]) { k.stage = S_BCAST; continue; }
        bool vec = 1==x.st[0] && 0==(x.base % uint64_t(V*k.esize));
        for (int j=1; j<nA+nB && vec; ++j) vec = 0==(x.st[j] % V);
        k.stage = vec ? S_VEC : S_GATHER;
    }
    for (int pass=0; pass<2; ++pass) { // cp.async operands first so their loads are in flight before the synchronous ones
        for (int o=0; o<kp.nopnd; ++o) {
            KOpnd & k = kp.opnd[o];
            if (!used[o] || S_NONE==k.stage) continue;
            bool async = (k.stage==S_VEC || k.stage==S_GATHER) && k.esize>=4;
            if (async==(0==pass)) { k.slot = uint8_t(nstage++); nasync += async; }
        }
    }
    kp.nstage = nstage;

    // ---- stack levels
    kp.nins = int(prog.size());
    int depth = 0, sd = 0;
    for (int pc=0; pc<kp.nins; ++pc) {
        KIns & k = kp.ins[pc];
        racu_ins const & i = prog[pc];
        k.op = i.op; k.type = i.type; k.arg = i.arg; k.aux = i.aux; k.lvl = 255;
        switch (i.op) {
        case RACU_OP_PUSH: if (sd>=1) { k.lvl = uint8_t(sd-1); depth = std::max(depth, sd); } ++sd; break;
        case RACU_OP_CAST: case RACU_OP_NEG: case RACU_OP_ABS: case RACU_OP_SQRT: case RACU_OP_SQR: case RACU_OP_EXP:
        case RACU_OP_LOG: case RACU_OP_SIN: case RACU_OP_COS: case RACU_OP_NOT: case RACU_OP_LOAD:
        case RACU_OP_TAN: case RACU_OP_SINH: case RACU_OP_COSH: case RACU_OP_TANH: case RACU_OP_ASIN: case RACU_OP_ACOS: case RACU_OP_ATAN:
        case RACU_OP_EXPM1: case RACU_OP_LOG1P: case RACU_OP_LOG10: case RACU_OP_ISFINITE: case RACU_OP_ISNAN: case RACU_OP_ISINF: break;
        case RACU_OP_FMA: case RACU_OP_CLAMP: case RACU_OP_LERP: k.lvl = uint8_t(sd-3); sd -= 2; break;
        case RACU_OP_PICK: k.lvl = uint8_t(sd-1-i.arg); sd -= i.arg; break;
        default: k.lvl = uint8_t(sd-2); --sd; break;
        }
    }
    kp.depth = depth;

    // ---- launch geometry
    int U = nasync>0 ? std::max(1, std::min(4, 8/nasync)) : 1;
    auto smem_for = [&](int u) { return size_t(u)*size_t(nstage+depth)*KBLOCK*KSLOT; }; // U staged vectors + U stack rows per level
    if (3==U) U = 2;
    if (char const * e = std::getenv("RACU_U")) { int u = std::atoi(e); if (1==u || 2==u || 4==u) U = u; } // tuning knob
    while (U>1 && smem_for(U)>size_t(100*1024)) U /= 2;
    if (smem_for(U)>size_t(220*1024)) { err = "expression needs too much staging memory"; return RACU_ERR_UNSUPPORTED; }
    kp.U = U;
    plan.smem = smem_for(U);
    kp.fold_op = assign; kp.acc_type = A; kp.overwrite = reducing;
    kp.identity = identity_bits(assign, A);
    int const target = 148*8;
    if (mode==K_MAP) {
        int64_t nb = (kp.nvecA+int64_t(KBLOCK)*U-1)/(int64_t(KBLOCK)*U);
        plan.grid_x = unsigned(std::min<int64_t>(nb, int64_t(target)*4));
        kp.nchunks = 1;
        plan.kernel = wide ? "ply_map<u64>" : "ply_map<u32>";
    } else if (mode==K_FOLD_INNER) {
        int64_t per = int64_t(KBLOCK)*U;
        int64_t maxchunks = std::max<int64_t>(1, kp.nvecA/(per*2));
        int64_t want = std::max<int64_t>(1, (int64_t(target)*2)/std::max<int64_t>(1, kp.nk));
        int64_t nch = std::min(maxchunks, want);
        kp.chunk_len = (kp.nvecA+nch-1)/nch;
        kp.chunk_len = ((kp.chunk_len+per-1)/per)*per;
        nch = (kp.nvecA+kp.chunk_len-1)/kp.chunk_len;
        kp.nchunks = int(nch);
        if (kp.nk*nch>=(int64_t(1)<<31)) { err = "too many fold groups"; return RACU_ERR_UNSUPPORTED; }
        plan.grid_x = unsigned(kp.nk*nch);
        plan.kernel = wide ? "ply_fold_inner<u64>" : "ply_fold_inner<u32>";
    } else {
        int64_t nbx = (kp.nvecA+KBLOCK-1)/KBLOCK;
        int64_t want = std::max<int64_t>(1, (int64_t(target)*2)/nbx);
        int64_t maxchunks = std::max<int64_t>(1, kp.nB/(int64_t(U)*8));
        int64_t nch = std::min<int64_t>(std::min(maxchunks, want), 65535);
        kp.chunk_len = (kp.nB+nch-1)/nch;
        kp.chunk_len = ((kp.chunk_len+U-1)/U)*U;
        nch = (kp.nB+kp.chunk_len-1)/kp.chunk_len;
        kp.nchunks = int(nch);
        if (nbx>=(int64_t(1)<<31)) { err = "too many fold groups"; return RACU_ERR_UNSUPPORTED; }
        plan.grid_x = unsigned(nbx); plan.grid_y = unsigned(nch);
        plan.kernel = wide ? "ply_fold_outer<u64>" : "ply_fold_outer<u32>";
    }
    if (mode!=K_MAP) {
        plan.need_finish = kp.nchunks>1;
        plan.partial_bytes = plan.need_finish ? size_t(kp.nchunks)*size_t(kp.nk)*(wide ? 8 : 4) : 0;
        FinishParams & f = plan.fp;
        std::memset(&f, 0, sizeof(f));
        f.wide = wide; f.fold_op = assign; f.acc_type = A; f.dst_dtype = D; f.overwrite = reducing; f.nchunks = kp.nchunks;
        f.nk = kp.nk; f.dst = reinterpret_cast<void *>(uintptr_t(w.op[dst].base));
        f.per_warp = mode==K_FOLD_INNER;
        // kept axes, fastest first, with the destination's strides
        int k0 = mode==K_FOLD_INNER ? nA : 0, kn = mode==K_FOLD_INNER ? nB : nA;
        f.rank = kn;
        for (int k=0; k<kn; ++k) { f.stride[k] = w.op[dst].st[k0+k]; if (k>0 || kn>1) f.d[k] = make_div(w.len[k0+k]); }
        f.len0 = kn>0 ? w.len[k0] : 1;
        if (kn>1 && f.len0>=(int64_t(1)<<31)) { err = "axis longer than 2^31 after collapsing"; return RACU_ERR_UNSUPPORTED; }
    }
    match_static(plan, prog, dst, fold);
    return RACU_OK;
}

} // namespace racu
