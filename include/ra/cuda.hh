// -*- mode: c++; coding: utf-8 -*-
// ra/cuda.hh - header-only C++23 host side of the B200 back end for ra-ra's traversal hot path.
//
// This header keeps ra-ra's array surface for the path that funnels into ra::ply (ra/ply.hh:157-166):
//   Big / View of static rank or ANY        ra/arrays.hh:84-149,252-327
//   operators & named functions -> Map trees  ra/ra.hh:181-224, ra/expr.hh:652-678
//   rank extension by PREFIX matching         ra/expr.hh:532-634 (checked by the library: racu_agree)
//   iota / all / dots / insert / len slicing  ra/ply.hh:314-330,367-435,503-551 (beaten subscripts only)
//   transpose / reverse / diag                ra/arrays.hh:408-462
//   sum prod amin amax dot reduce_sqrm norm2  ra/ra.hh:306-322,348-401
//   where / pick / && / ||                    ra/ra.hh:252-272, ra/expr.hh:686-728
//   gemm / gemv / gevm                        ra/ra.hh:403-456
// but the arrays live in device memory, and instead of traversing, ra::cuda::ply FLATTENS the expression tree into a
// racu_desc (leaves: base pointer + per-axis len/step with step 0 on broadcast axes; a short postfix program) and
// passes it through the C ABI (include/racu.h) to hand-written sm_100a kernels. There is no CPU fallback: an
// expression with a leaf that is not device resident (other than scalars and rank-1 host vectors, which are uploaded)
// does not compile.
//
// Differences from the reference, all forced by the device boundary:
//   * ops are an enumerated set of named functor types (no user lambdas in map());
//   * a rank-0 subscript a(i, j) is a rank-0 View; it converts to T by copying from the device;
//   * RA_CK failures throw ra::cuda::error by default (define RACU_ASSERT to override, like RA_ASSERT, ra/expr.hh:23-24).
// Written for g++ >= 13 -std=c++23 (no deducing-this, <print> or ranges::to). Include only this file and link libracu.so.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../racu.h"

namespace ra::cuda {

using rank_t = int;              // ra/base.hh:236
using dim_t = std::ptrdiff_t;    // ra/base.hh:237
constexpr dim_t ANY = RACU_ANY, UNB = RACU_UNB, MIS = RACU_MIS; // ra/base.hh:232-234

struct error: std::runtime_error
{
    int status;
    error(int status_, std::string const & msg): std::runtime_error(msg), status(status_) {}
};

#ifndef RACU_ASSERT
#define RACU_ASSERT(cond, status, msg) do { if (!(cond)) throw ::ra::cuda::error((status), (msg)); } while (0)
#endif
#define RACU_CK(cond, msg) RACU_ASSERT((cond), RACU_ERR_BOUNDS, std::string(msg))

inline void ck_status(int st) { RACU_ASSERT(RACU_OK==st, st, std::string(racu_last_error_string())); }

// The stream every traversal issued from this thread is enqueued on (a cudaStream_t; nullptr = default stream).
inline void * & stream() { thread_local void * s = nullptr; return s; }
inline void sync() { ck_status(racu_stream_sync(stream())); }

struct none_t {};
constexpr none_t none {};

struct Dim { dim_t len, step; };  // ra/expr.hh:183

template <class T> constexpr int dtype_of = -1;
template <> inline constexpr int dtype_of<bool> = RACU_BOOL;
template <> inline constexpr int dtype_of<int> = RACU_I32;
template <> inline constexpr int dtype_of<long> = RACU_I64;
template <> inline constexpr int dtype_of<long long> = RACU_I64;
template <> inline constexpr int dtype_of<float> = RACU_F32;
template <> inline constexpr int dtype_of<double> = RACU_F64;
template <class T> concept is_elem = (dtype_of<std::remove_cv_t<T>> >= 0);


// ---------------------
// subscripts: ra/ply.hh:314-330, ra/base.hh:278-291
// ---------------------

// ra::len and arithmetic on it; resolved against the axis length by View::operator() (wlen, ra/ply.hh:243-295).
struct lenx
{
    dim_t v = 0;
    std::function<dim_t(dim_t)> f;
    lenx() = default;
    template <std::integral I> lenx(I i): v(dim_t(i)) {}
    explicit lenx(std::function<dim_t(dim_t)> f_): f(std::move(f_)) {}
    bool symbolic() const { return bool(f); }
    dim_t operator()(dim_t ln) const { return f ? f(ln) : v; }
};
struct Len { operator lenx() const { return lenx([](dim_t n){ return n; }); } };
constexpr Len len {};

template <class A> concept is_lenlike = std::is_same_v<std::decay_t<A>, Len> || std::is_same_v<std::decay_t<A>, lenx>;
template <class A> concept is_lenarg = is_lenlike<A> || std::integral<std::decay_t<A>>;
#define RACU_LENOP(OP)                                                  \
    template <class A, class B> requires ((is_lenlike<A> && is_lenarg<B>) || (is_lenarg<A> && is_lenlike<B>)) \
    lenx operator OP(A const & a, B const & b)                          \
    {                                                                   \
        lenx x = a, y = b;                                              \
        return lenx([x, y](dim_t n){ return x(n) OP y(n); });           \
    }
RACU_LENOP(+) RACU_LENOP(-) RACU_LENOP(*) RACU_LENOP(/)
#undef RACU_LENOP
template <is_lenlike A> lenx operator-(A const & a) { lenx x = a; return lenx([x](dim_t n){ return -x(n); }); }

template <int n> struct insert_t { constexpr static int N = n; static_assert(n>=0); };
template <int n=1> constexpr insert_t<n> insert = insert_t<n>();
template <int n> struct dots_t { constexpr static int N = n; static_assert(n>=0 || UNB==n); };
template <int n=int(UNB)> constexpr dots_t<n> dots = dots_t<n>();
constexpr auto all = dots<1>;

struct node_tag {}; // every expression node derives from this
template <class A> concept is_node = std::is_base_of_v<node_tag, std::decay_t<A>>;

class Builder;

// ra::iota(n, o, s): rank-1 sequence o + k*s (ra/ply.hh:320-321). A subscript when used in operator(), a SEQ leaf in expressions.
template <class T=dim_t>
struct Iota: node_tag
{
    using value_type = T;
    lenx n, o, s;
    inline void flatten(Builder & b) const;
};
template <class A> constexpr bool is_iota = false;
template <class T> constexpr bool is_iota<Iota<T>> = true;

inline auto iota() { return Iota<dim_t> { {}, lenx(UNB), lenx(0), lenx(1) }; }
template <class N> requires (is_lenarg<N>) auto iota(N const & n) { return Iota<dim_t> { {}, lenx(n), lenx(0), lenx(1) }; }
template <class N, class O, class S=int> requires (is_lenarg<N> && is_lenarg<S>)
auto iota(N const & n, O const & o, S const & s=1)
{
    if constexpr (is_lenlike<O>) return Iota<dim_t> { {}, lenx(n), lenx(o), lenx(s) };
    else if constexpr (std::integral<O>) return Iota<O> { {}, lenx(n), lenx(o), lenx(s) };
    else static_assert(std::integral<O>, "iota origin must be integral on the device path");
}

// iota arithmetic stays iota: opt(), ra/ra.hh:129-143. Needed so that A(I+1, J, K) is a view (bench/bench-stencil3.cc:59-61).
template <class T, class K> requires (is_lenarg<K>) auto operator+(Iota<T> const & i, K const & k) { return Iota<T> { {}, i.n, i.o+lenx(k), i.s }; }
template <class T, class K> requires (is_lenarg<K>) auto operator+(K const & k, Iota<T> const & i) { return Iota<T> { {}, i.n, lenx(k)+i.o, i.s }; }
template <class T, class K> requires (is_lenarg<K>) auto operator-(Iota<T> const & i, K const & k) { return Iota<T> { {}, i.n, i.o-lenx(k), i.s }; }
template <class T, class K> requires (is_lenarg<K>) auto operator-(K const & k, Iota<T> const & i) { return Iota<T> { {}, i.n, lenx(k)-i.o, -i.s }; }
template <class T, class K> requires (std::integral<K>) auto operator*(Iota<T> const & i, K const & k) { return Iota<T> { {}, i.n, i.o*lenx(k), i.s*lenx(k) }; }
template <class T, class K> requires (std::integral<K>) auto operator*(K const & k, Iota<T> const & i) { return Iota<T> { {}, i.n, lenx(k)*i.o, lenx(k)*i.s }; }
template <class T> auto operator-(Iota<T> const & i) { return Iota<T> { {}, i.n, -i.o, -i.s }; }


// ---------------------
// descriptor builder
// ---------------------

class Builder
{
public:
    racu_desc d;
    std::vector<void *> temps; // device copies of host vectors, freed after the traversal
    int shift = 0;             // leaves flattened now are placed `shift` axes further in (dead leading axes): index arrays of a(i ...)

    Builder() { std::memset(&d, 0, sizeof(d)); d.dst = -1; }
    Builder(Builder const &) = delete;
    ~Builder() { if (!temps.empty()) { racu_stream_sync(stream()); for (void * p: temps) racu_free(p); } }

    int operand(racu_operand const & o)
    {
        for (int k=0; k<d.nopnd; ++k) { if (0==std::memcmp(&d.opnd[k], &o, sizeof(o))) return k; } // one load per distinct leaf
        RACU_ASSERT(d.nopnd<RACU_MAX_OPND, RACU_ERR_UNSUPPORTED, "too many operands in one expression");
        d.opnd[d.nopnd] = o;
        return d.nopnd++;
    }
    void emit(int op, int type, int arg=0, int aux=0)
    {
        RACU_ASSERT(d.nins<RACU_MAX_INS, RACU_ERR_UNSUPPORTED, "expression too long");
        racu_ins & i = d.ins[d.nins++];
        i.op = uint8_t(op); i.type = uint8_t(type); i.arg = uint8_t(arg); i.aux = uint8_t(aux);
    }
    template <class From, class To> void cast()
    {
        if constexpr (dtype_of<From> != dtype_of<To>) emit(RACU_OP_CAST, dtype_of<To>, 0, dtype_of<From>);
    }
    template <class T> int mem(T const * p, Dim const * dv, rank_t rank)
    {
        RACU_CK(rank<=RACU_MAX_RANK, "Rank too large.");
        racu_operand o;
        std::memset(&o, 0, sizeof(o));
        RACU_CK(rank+shift<=RACU_MAX_RANK, "Rank too large.");
        o.kind = RACU_MEM; o.dtype = dtype_of<std::remove_cv_t<T>>; o.rank = rank+shift;
        o.base.ptr = const_cast<std::remove_cv_t<T> *>(p);
        for (rank_t k=0; k<shift; ++k) { o.len[k] = UNB; o.stride[k] = 0; }
        for (rank_t k=0; k<rank; ++k) { o.len[shift+k] = dv[k].len; o.stride[shift+k] = dv[k].step; }
        return operand(o);
    }
    // an array addressed by computed offsets (RACU_OP_LOAD): base = element at offset 0, [lo, hi] = valid element offsets
    template <class T> int ptr(T const * p, dim_t lo, dim_t hi)
    {
        racu_operand o;
        std::memset(&o, 0, sizeof(o));
        o.kind = RACU_PTR; o.dtype = dtype_of<std::remove_cv_t<T>>; o.rank = 0;
        o.base.ptr = const_cast<std::remove_cv_t<T> *>(p);
        o.len[0] = lo; o.len[1] = hi;
        return operand(o);
    }
    template <class T> int scalar(T v)
    {
        racu_operand o;
        std::memset(&o, 0, sizeof(o));
        o.kind = RACU_SCALAR; o.dtype = dtype_of<T>; o.rank = 0;
        if constexpr (std::is_floating_point_v<T>) o.base.f = double(v); else o.base.i = int64_t(v);
        return operand(o);
    }
    template <class T> int seq(T org, rank_t rank, dim_t const * len, dim_t const * step)
    {
        racu_operand o;
        std::memset(&o, 0, sizeof(o));
        RACU_CK(rank+shift<=RACU_MAX_RANK, "Rank too large.");
        o.kind = RACU_SEQ; o.dtype = dtype_of<T>; o.rank = rank+shift;
        if constexpr (std::is_floating_point_v<T>) o.base.f = double(org); else o.base.i = int64_t(org);
        for (rank_t k=0; k<shift; ++k) { o.len[k] = UNB; o.stride[k] = 0; }
        for (rank_t k=0; k<rank; ++k) { o.len[shift+k] = len[k]; o.stride[shift+k] = step[k]; }
        return operand(o);
    }
};

template <class T> void Iota<T>::flatten(Builder & b) const
{
    RACU_CK(!n.symbolic() && !o.symbolic() && !s.symbolic(), "Stray ra::len."); // ra/expr.hh:513
    dim_t ln = n(0), st = s(0);
    b.emit(RACU_OP_PUSH, dtype_of<T>, b.seq<T>(T(o(0)), 1, &ln, &st));
}


// ---------------------
// leaves
// ---------------------

template <class T> struct Scalar: node_tag // ra/expr.hh:104-120
{
    using value_type = T;
    T c;
    void flatten(Builder & b) const { b.emit(RACU_OP_PUSH, dtype_of<T>, b.scalar<T>(c)); }
};

// Index along axis w (ra::_0 ...: tindex, ra/arrays.hh:361-364): dead on the axes before w.
template <int w> struct TIndex: node_tag
{
    using value_type = dim_t;
    void flatten(Builder & b) const
    {
        dim_t ln[w+1], st[w+1];
        for (int k=0; k<=w; ++k) { ln[k] = UNB; st[k] = (k==w); }
        b.emit(RACU_OP_PUSH, dtype_of<dim_t>, b.seq<dim_t>(0, w+1, ln, st));
    }
};
constexpr TIndex<0> _0 {};
constexpr TIndex<1> _1 {};
constexpr TIndex<2> _2 {};
constexpr TIndex<3> _3 {};
constexpr TIndex<4> _4 {};

// A foreign rank-1 range (std::vector / std::array; is_fov ra/base.hh:332, iter(is_fov) ra/expr.hh:385): uploaded at ply time.
template <class T> struct HostVec: node_tag
{
    using value_type = T;
    T const * p;
    dim_t n;
    void flatten(Builder & b) const
    {
        void * dev = nullptr;
        ck_status(racu_alloc(&dev, sizeof(T)*size_t(n)));
        b.temps.push_back(dev);
        ck_status(racu_memcpy_h2d(dev, p, sizeof(T)*size_t(n), stream()));
        Dim dv {n, 1};
        b.emit(RACU_OP_PUSH, dtype_of<T>, b.mem<T>(static_cast<T *>(dev), &dv, 1));
    }
};

template <class A> constexpr bool is_fov = false;
template <class T, class Al> requires (is_elem<T> && !std::is_same_v<T, bool>) constexpr bool is_fov<std::vector<T, Al>> = true;
template <class T, std::size_t N> requires (is_elem<T>) constexpr bool is_fov<std::array<T, N>> = true;

template <class A> concept is_scalar = is_elem<std::decay_t<A>>;
template <class A> concept is_ra = is_node<A>;                                  // device expression nodes
template <class A> concept is_operand = is_ra<A> || is_scalar<A> || is_fov<std::decay_t<A>>;

// iter(x): make a node out of anything that can appear in an expression (ra/expr.hh:128-134,317-328,385-387).
template <class A> requires (is_operand<A>)
constexpr decltype(auto) iter(A && a)
{
    using D = std::decay_t<A>;
    if constexpr (requires { a.view(); }) return a.view(); // owning arrays enter expressions as their (non owning) view
    else if constexpr (is_node<A>) return std::forward<A>(a);
    else if constexpr (is_scalar<A>) return Scalar<D> { {}, a };
    else return HostVec<typename D::value_type> { {}, a.data(), dim_t(a.size()) };
}
template <class A> using node_t = std::decay_t<decltype(iter(std::declval<A>()))>;
template <class A> using value_t = typename node_t<A>::value_type;


// ---------------------
// ops: the opcode set of the postfix program (ra/ra.hh:181-224). Each functor is a distinct named type.
// ---------------------

#define RACU_ARITH(NAME, OPCODE)                                        \
    struct NAME                                                         \
    {                                                                   \
        template <class A, class B> using common = decltype(std::declval<A>()+std::declval<B>()); \
        template <class A, class B> using result = common<A, B>;        \
        static constexpr int opcode = OPCODE;                           \
    };
#define RACU_COMPARE(NAME, OPCODE)                                      \
    struct NAME                                                         \
    {                                                                   \
        template <class A, class B> using common = decltype(std::declval<A>()+std::declval<B>()); \
        template <class A, class B> using result = bool;                \
        static constexpr int opcode = OPCODE;                           \
    };
RACU_ARITH(plus, RACU_OP_ADD)       RACU_ARITH(minus, RACU_OP_SUB)      RACU_ARITH(times, RACU_OP_MUL)
RACU_ARITH(divides, RACU_OP_DIV)    RACU_ARITH(modulus, RACU_OP_MOD)    RACU_ARITH(bit_and, RACU_OP_BAND)
RACU_ARITH(bit_or, RACU_OP_BOR)     RACU_ARITH(bit_xor, RACU_OP_BXOR)   RACU_ARITH(max_fn, RACU_OP_MAX)
RACU_ARITH(min_fn, RACU_OP_MIN)
RACU_COMPARE(equal_to, RACU_OP_EQ)  RACU_COMPARE(not_equal_to, RACU_OP_NE) RACU_COMPARE(less, RACU_OP_LT)
RACU_COMPARE(less_equal, RACU_OP_LE) RACU_COMPARE(greater, RACU_OP_GT)  RACU_COMPARE(greater_equal, RACU_OP_GE)
#undef RACU_ARITH
#undef RACU_COMPARE

// binary node: operands are converted to their common type first (usual arithmetic conversions), as C++ would.
template <class Op, class A, class B>
struct Map2: node_tag
{
    using TA = typename A::value_type;
    using TB = typename B::value_type;
    using common = typename Op::template common<TA, TB>;
    using value_type = typename Op::template result<TA, TB>;
    static_assert(is_elem<common>, "unsupported operand types");
    A a; B b;
    void flatten(Builder & bb) const
    {
        a.flatten(bb); bb.template cast<TA, common>();
        b.flatten(bb); bb.template cast<TB, common>();
        bb.emit(Op::opcode, dtype_of<common>);
    }
};

// pow / atan2: std::pow(float, float) is float, anything else promotes to double.
template <int OPCODE, class A, class B>
struct MapF2: node_tag
{
    using TA = typename A::value_type;
    using TB = typename B::value_type;
    using value_type = std::conditional_t<std::is_same_v<TA, float> && std::is_same_v<TB, float>, float, double>;
    A a; B b;
    void flatten(Builder & bb) const
    {
        a.flatten(bb); bb.template cast<TA, value_type>();
        b.flatten(bb); bb.template cast<TB, value_type>();
        bb.emit(OPCODE, dtype_of<value_type>);
    }
};

enum class un_kind { same, floating, logical, classify }; // result type rule (classify: floating operand, bool result)
template <int OPCODE, un_kind K, class A>
struct Map1: node_tag
{
    using TA = typename A::value_type;
    using promoted = decltype(+std::declval<TA>());
    using operand_t = std::conditional_t<K==un_kind::floating || K==un_kind::classify, std::conditional_t<std::is_floating_point_v<TA>, TA, double>,
                                         std::conditional_t<K==un_kind::logical, TA, promoted>>;
    using value_type = std::conditional_t<K==un_kind::logical || K==un_kind::classify, bool, operand_t>;
    A a;
    void flatten(Builder & bb) const
    {
        a.flatten(bb); bb.template cast<TA, operand_t>();
        bb.emit(OPCODE, dtype_of<operand_t>);
    }
};

template <class T, class A>
struct CastNode: node_tag // ra::cast<T>, ra/expr.hh:680-681
{
    using value_type = T;
    A a;
    void flatten(Builder & bb) const { a.flatten(bb); bb.template cast<typename A::value_type, T>(); }
};

template <class A, class B, class C, int OPCODE=RACU_OP_FMA>
struct Fma: node_tag
{
    // fma / lerp: the floating point type the three promote to; clamp: their common type
    using common3 = decltype(std::declval<typename A::value_type>()+std::declval<typename B::value_type>()+std::declval<typename C::value_type>());
    using value_type = std::conditional_t<OPCODE==RACU_OP_CLAMP, common3, std::conditional_t<std::is_floating_point_v<common3>, common3, double>>;
    A a; B b; C c;
    void flatten(Builder & bb) const
    {
        a.flatten(bb); bb.template cast<typename A::value_type, value_type>();
        b.flatten(bb); bb.template cast<typename B::value_type, value_type>();
        c.flatten(bb); bb.template cast<typename C::value_type, value_type>();
        bb.emit(OPCODE, dtype_of<value_type>);
    }
};

// ra::pick(sel, a0, a1, ...): ra/expr.hh:686-728. Only the chosen branch is observable (the opcode set has no side effects).
template <class S, class ... P>
struct Pick: node_tag
{
    using value_type = std::common_type_t<typename P::value_type ...>;
    static_assert(std::is_integral_v<typename S::value_type>, "pick selector must be integral or bool");
    S sel;
    std::tuple<P ...> p;
    void flatten(Builder & bb) const
    {
        sel.flatten(bb);
        std::apply([&](auto const & ... q) { ((q.flatten(bb), bb.template cast<typename std::decay_t<decltype(q)>::value_type, value_type>()), ...); }, p);
        bb.emit(RACU_OP_PICK, dtype_of<value_type>, int(sizeof...(P)), dtype_of<typename S::value_type>);
    }
    // pick / where as an lvalue (test/where.cc:148-157): `pick(sel, a0, a1, ...) OP= x` updates, per element, the branch the selector
    // names; every branch must be an array. On the device: one masked assignment per branch, a_k = pick(sel, a_k, ..., a_k OP x, ..., a_k).
    template <std::size_t K, class X> void assign_branch(int op, X const & x) const
    {
        auto const & y = std::get<K>(p);
        using Y = std::decay_t<decltype(y)>;
        using TY = typename Y::value_type;
        using TX = typename X::value_type;
        static_assert(requires { y.cp; y.dimv; } && !std::is_const_v<std::remove_pointer_t<decltype(y.cp)>>, "pick / where as an lvalue needs array branches");
        Builder b;
        b.d.dst = b.template mem<TY>(y.cp, y.dimv.data(), y.rank());
        b.d.assign = RACU_ASSIGN_SET;
        sel.flatten(b);
        for (std::size_t j=0; j<sizeof...(P); ++j) {
            if (j!=K) { y.flatten(b); continue; }
            if (RACU_ASSIGN_SET==op) { x.flatten(b); b.template cast<TX, TY>(); continue; }
            using C = decltype(std::declval<TY>()+std::declval<TX>());
            y.flatten(b); b.template cast<TY, C>();
            x.flatten(b); b.template cast<TX, C>();
            b.emit(RACU_ASSIGN_ADD==op ? RACU_OP_ADD : RACU_ASSIGN_SUB==op ? RACU_OP_SUB : RACU_ASSIGN_MUL==op ? RACU_OP_MUL : RACU_OP_DIV, dtype_of<C>);
            b.template cast<C, TY>();
        }
        b.emit(RACU_OP_PICK, dtype_of<TY>, int(sizeof...(P)), dtype_of<typename S::value_type>);
        ck_status(racu_map(&b.d, stream()));
    }
    template <class X> void assign(int op, X const & x) const
    {
        [&]<std::size_t ... K>(std::index_sequence<K ...>) { (assign_branch<K>(op, x), ...); }(std::index_sequence_for<P ...> {});
    }
#define RACU_ASSIGNOP(OP, CODE) \
    template <class X> requires (std::is_base_of_v<node_tag, std::decay_t<X>> || is_elem<std::decay_t<X>>) \
    Pick const & operator OP(X const & x) const { if constexpr (is_elem<std::decay_t<X>>) assign(CODE, Scalar<std::decay_t<X>> { {}, x }); else assign(CODE, x); return *this; }
    RACU_ASSIGNOP(=, RACU_ASSIGN_SET) RACU_ASSIGNOP(+=, RACU_ASSIGN_ADD) RACU_ASSIGNOP(-=, RACU_ASSIGN_SUB)
    RACU_ASSIGNOP(*=, RACU_ASSIGN_MUL) RACU_ASSIGNOP(/=, RACU_ASSIGN_DIV)
#undef RACU_ASSIGNOP
};

template <class ... A> concept tomap = ((is_operand<A> && ...) && (is_ra<A> || ...) && !(is_iota<std::decay_t<A>> && ...));

#define RACU_BINOP(OP, OPNAME)                                          \
    template <class A, class B> requires (tomap<A, B> && !(is_iota<std::decay_t<A>> && is_lenarg<B>) && !(is_lenarg<A> && is_iota<std::decay_t<B>>)) \
    constexpr auto operator OP(A && a, B && b) { return Map2<OPNAME, node_t<A>, node_t<B>> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)) }; }
RACU_BINOP(+, plus)       RACU_BINOP(-, minus)          RACU_BINOP(*, times)
RACU_BINOP(/, divides)    RACU_BINOP(==, equal_to)      RACU_BINOP(<=, less_equal)
RACU_BINOP(<, less)       RACU_BINOP(>, greater)        RACU_BINOP(>=, greater_equal)
RACU_BINOP(|, bit_or)     RACU_BINOP(&, bit_and)        RACU_BINOP(!=, not_equal_to)
RACU_BINOP(^, bit_xor)    RACU_BINOP(%, modulus)
#undef RACU_BINOP

template <class A> requires (tomap<A> && !is_iota<std::decay_t<A>>)
constexpr auto operator-(A && a) { return Map1<RACU_OP_NEG, un_kind::same, node_t<A>> { {}, iter(std::forward<A>(a)) }; }
template <class A> requires (tomap<A>)
constexpr auto operator+(A && a) { return CastNode<decltype(+std::declval<value_t<A>>()), node_t<A>> { {}, iter(std::forward<A>(a)) }; }
template <class A> requires (tomap<A>)
constexpr auto operator!(A && a) { return Map1<RACU_OP_NOT, un_kind::logical, node_t<A>> { {}, iter(std::forward<A>(a)) }; }

#define RACU_UNFN(NAME, OPCODE, KIND)                                   \
    template <class A> requires (tomap<A>)                              \
    constexpr auto NAME(A && a) { return Map1<OPCODE, KIND, node_t<A>> { {}, iter(std::forward<A>(a)) }; }
RACU_UNFN(abs, RACU_OP_ABS, un_kind::same)   RACU_UNFN(sqr, RACU_OP_SQR, un_kind::same)
RACU_UNFN(sqrt, RACU_OP_SQRT, un_kind::floating) RACU_UNFN(exp, RACU_OP_EXP, un_kind::floating)
RACU_UNFN(log, RACU_OP_LOG, un_kind::floating)   RACU_UNFN(sin, RACU_OP_SIN, un_kind::floating)
RACU_UNFN(cos, RACU_OP_COS, un_kind::floating)
// the rest of the using-declared <cmath> functions of ra/ra.hh:221-222
RACU_UNFN(tan, RACU_OP_TAN, un_kind::floating)     RACU_UNFN(sinh, RACU_OP_SINH, un_kind::floating)
RACU_UNFN(cosh, RACU_OP_COSH, un_kind::floating)   RACU_UNFN(tanh, RACU_OP_TANH, un_kind::floating)
RACU_UNFN(asin, RACU_OP_ASIN, un_kind::floating)   RACU_UNFN(acos, RACU_OP_ACOS, un_kind::floating)
RACU_UNFN(atan, RACU_OP_ATAN, un_kind::floating)   RACU_UNFN(expm1, RACU_OP_EXPM1, un_kind::floating)
RACU_UNFN(log1p, RACU_OP_LOG1P, un_kind::floating) RACU_UNFN(log10, RACU_OP_LOG10, un_kind::floating)
RACU_UNFN(isfinite, RACU_OP_ISFINITE, un_kind::classify) RACU_UNFN(isnan, RACU_OP_ISNAN, un_kind::classify)
RACU_UNFN(isinf, RACU_OP_ISINF, un_kind::classify)
#undef RACU_UNFN

template <class A, class B> requires (tomap<A, B>)
constexpr auto max(A && a, B && b) { return Map2<max_fn, node_t<A>, node_t<B>> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)) }; }
template <class A, class B> requires (tomap<A, B>)
constexpr auto min(A && a, B && b) { return Map2<min_fn, node_t<A>, node_t<B>> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)) }; }
template <class A, class B> requires (tomap<A, B>)
constexpr auto pow(A && a, B && b) { return MapF2<RACU_OP_POW, node_t<A>, node_t<B>> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)) }; }
template <class A, class B> requires (tomap<A, B>)
constexpr auto atan2(A && a, B && b) { return MapF2<RACU_OP_ATAN2, node_t<A>, node_t<B>> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)) }; }
template <class A, class B, class C> requires (tomap<A, B, C>)
constexpr auto fma(A && a, B && b, C && c) { return Fma<node_t<A>, node_t<B>, node_t<C>> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)), iter(std::forward<C>(c)) }; }

// std::clamp / std::lerp (ra/ra.hh:222)
template <class A, class B, class C> requires (tomap<A, B, C>)
constexpr auto clamp(A && a, B && b, C && c) { return Fma<node_t<A>, node_t<B>, node_t<C>, RACU_OP_CLAMP> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)), iter(std::forward<C>(c)) }; }
template <class A, class B, class C> requires (tomap<A, B, C>)
constexpr auto lerp(A && a, B && b, C && c) { return Fma<node_t<A>, node_t<B>, node_t<C>, RACU_OP_LERP> { {}, iter(std::forward<A>(a)), iter(std::forward<B>(b)), iter(std::forward<C>(c)) }; }

// sqrm(x) = x*x, sqrm(x, y) = sqr(x-y): ra/ra.hh:71-72.
template <class A> requires (tomap<A>) constexpr auto sqrm(A && a) { return sqr(std::forward<A>(a)); }
template <class A, class B> requires (tomap<A, B>) constexpr auto sqrm(A && a, B && b) { return sqr(std::forward<A>(a)-std::forward<B>(b)); }
template <class T> requires (std::is_arithmetic_v<T>) constexpr T sqr(T x) { return x*x; }   // ra/ra.hh:65
template <class T> requires (std::is_arithmetic_v<T>) constexpr T sqrm(T x) { return x*x; }

template <class T, class A> requires (tomap<A>)
constexpr auto cast(A && a) { return CastNode<T, node_t<A>> { {}, iter(std::forward<A>(a)) }; }

template <class S, class ... P> requires (tomap<S, P ...> && sizeof...(P)>=1)
constexpr auto pick(S && s, P && ... p) { return Pick<node_t<S>, node_t<P> ...> { {}, iter(std::forward<S>(s)), std::tuple<node_t<P> ...>(iter(std::forward<P>(p)) ...) }; }

// where(w, t, f) = pick(cast<bool>(w), f, t): ra/ra.hh:255-256.
template <class W, class T, class F> requires (tomap<W, T, F>)
constexpr auto where(W && w, T && t, F && f) { return pick(cast<bool>(std::forward<W>(w)), std::forward<F>(f), std::forward<T>(t)); }
// a && b = where(a, cast<bool>(b), false);  a || b = where(a, true, cast<bool>(b)): ra/ra.hh:262-266.
template <class A, class B> requires (tomap<A, B>)
constexpr auto operator&&(A && a, B && b) { return where(std::forward<A>(a), cast<bool>(std::forward<B>(b)), false); }
template <class A, class B> requires (tomap<A, B>)
constexpr auto operator||(A && a, B && b) { return where(std::forward<A>(a), true, cast<bool>(std::forward<B>(b))); }


// odd(x) (ra/ra.hh:61: N & 1), rel_error(a, b) (ra/ra.hh:94: D = |a|+|b|, D==0 ? 0. : 2.*|a-b|/D, with the double literals) and
// operator<=> (ra/ra.hh:190) are written with the opcodes above. `<=>` yields an int: -1, 0, +1, and 2 where the operands are
// unordered (the reference yields std::partial_ordering objects, which are not array elements on the device).
template <class A> requires (tomap<A> && std::is_integral_v<value_t<A>>)
constexpr auto odd(A && a) { return (std::forward<A>(a) & 1) != 0; }
template <class A, class B> requires (tomap<A, B>)
constexpr auto rel_error(A && a, B && b)
{
    using C0 = decltype(std::declval<value_t<A>>()+std::declval<value_t<B>>());
    using R = std::conditional_t<std::is_floating_point_v<C0>, C0, double>;
    auto x = cast<R>(std::forward<A>(a));
    auto y = cast<R>(std::forward<B>(b));
    auto D = abs(x) + abs(y);
    return cast<R>(where(D==R(0), 0., 2.*abs(x-y)/D));
}
template <class A, class B> requires (tomap<A, B>)
constexpr auto operator<=>(A && a, B && b)
{
    auto x = iter(std::forward<A>(a));
    auto y = iter(std::forward<B>(b));
    return where(x<y, -1, where(x>y, 1, where(x==y, 0, 2)));
}

// a(i ...) with unbeaten subscripts (index arrays / index expressions): ra/ply.hh:461-499,531-549. The reference maps
// a-as-a-function over the subscripts, each on its own axes (fromu_loop / wrank); here the element offset
// sum(index_k * step_k) is an ordinary I64 expression and RACU_OP_LOAD reads a[offset]. Offsets are clamped to the array
// (the reference asserts on a bad index, ra/ply.hh:395).
template <class T> struct Gather: node_tag
{
    using value_type = T;
    struct Term { std::function<void (Builder &)> flat; int shift; dim_t step; }; // flat pushes the index as I64
    T const * base = nullptr;
    dim_t lo = 0, hi = 0;
    std::vector<Term> terms;
    Gather() = default;
    Gather(Gather const &) = default;
    void flatten_offset(Builder & b) const
    {
        if (terms.empty()) b.emit(RACU_OP_PUSH, dtype_of<dim_t>, b.scalar<dim_t>(0));
        for (std::size_t q=0; q<terms.size(); ++q) {
            int const old = b.shift;
            b.shift += terms[q].shift;
            terms[q].flat(b);
            b.shift = old;
            b.emit(RACU_OP_PUSH, dtype_of<dim_t>, b.scalar<dim_t>(terms[q].step));
            b.emit(RACU_OP_MUL, dtype_of<dim_t>);
            if (q>0) b.emit(RACU_OP_ADD, dtype_of<dim_t>);
        }
    }
    void flatten(Builder & b) const
    {
        flatten_offset(b);
        b.emit(RACU_OP_LOAD, dtype_of<T>, b.ptr<T>(base, lo, hi), dtype_of<dim_t>);
    }
    // as an lvalue (test/fromu.cc:53-61,84-92): a[off] OP= x over the frame of the subscripts and x: a scatter. The program leaves
    // [offset, value] and the destination is the RACU_PTR operand; repeated indices combine in an unspecified order.
    template <class X> void assign(int op, X const & x) const
    {
        Builder b;
        flatten_offset(b);
        b.d.dst = b.ptr<T>(base, lo, hi);
        b.d.assign = op;
        iter(x).flatten(b);
        ck_status(racu_map(&b.d, stream()));
    }
    Gather const & operator=(Gather const & x) const { assign(RACU_ASSIGN_SET, x); return *this; }
#define RACU_ASSIGNOP(OP, CODE) \
    template <class X> requires (is_operand<X>) Gather const & operator OP(X const & x) const { assign(CODE, x); return *this; }
    RACU_ASSIGNOP(=, RACU_ASSIGN_SET) RACU_ASSIGNOP(+=, RACU_ASSIGN_ADD) RACU_ASSIGNOP(-=, RACU_ASSIGN_SUB)
    RACU_ASSIGNOP(*=, RACU_ASSIGN_MUL) RACU_ASSIGNOP(/=, RACU_ASSIGN_DIV)
#undef RACU_ASSIGNOP
};

// ---------------------
// ply: flatten and launch (replaces ra::ply, ra/ply.hh:157-166, for device resident expressions)
// ---------------------

// Shape of an expression by the prefix agreement rule (Match::len, ra/expr.hh:597-608).
template <class E> requires (is_ra<E>)
std::vector<dim_t> shape(E const & e)
{
    Builder b;
    e.flatten(b);
    int32_t rank = 0;
    int64_t ln[RACU_MAX_RANK];
    ck_status(racu_agree(&b.d, &rank, ln));
    return std::vector<dim_t>(ln, ln+rank);
}

template <class E> requires (is_ra<E>)
typename E::value_type reduce(E const & e, int redop)
{
    Builder b;
    e.flatten(b);
    typename E::value_type out {};
    if constexpr (std::is_same_v<typename E::value_type, bool>) {
        unsigned char o = 0; ck_status(racu_reduce(&b.d, redop, &o, stream())); return o!=0;
    } else {
        ck_status(racu_reduce(&b.d, redop, &out, stream()));
        return out;
    }
}

// Plan an expression exactly as evaluating it would - a map when `dst` is given (dst OP= e), a whole-array reduction otherwise - and
// make sure its kernel exists, without launching: a program outside the pre-compiled table is specialised and cached now
// (racu_prepare). The device analogue of the compiler instantiating ply_ravel for the expression's type. Returns the kernel family.
template <class E> requires (is_ra<E>)
std::string prepare(E const & e, int redop=RACU_RED_SUM)
{
    Builder b;
    e.flatten(b);
    char const * k = "";
    ck_status(racu_prepare(&b.d, redop, &k));
    return k ? k : "";
}

// whole-array reductions: ra/ra.hh:306-322,348-401
template <class A> requires (is_ra<A>) auto sum(A const & a) { return reduce(iter(a), RACU_RED_SUM); }
template <class A> requires (is_ra<A>) auto prod(A const & a) { return reduce(iter(a), RACU_RED_PROD); }
template <class A> requires (is_ra<A>) auto amin(A const & a) { return reduce(iter(a), RACU_RED_AMIN); }
template <class A> requires (is_ra<A>) auto amax(A const & a) { return reduce(iter(a), RACU_RED_AMAX); }
template <class A, class B> requires (tomap<A, B>) auto dot(A const & a, B const & b) { return reduce(a*b, RACU_RED_SUM); }
template <class A> requires (is_ra<A>) auto reduce_sqrm(A const & a) { return reduce(sqrm(a), RACU_RED_SUM); }
template <class A> requires (is_ra<A>) auto norm2(A const & a) { return std::sqrt(reduce_sqrm(a)); }

// any / every / index / lexical_compare: ra/ra.hh:279-304. The reference short-circuits a sequential traversal (early(),
// ra/ply.hh:169); the order of traversal is unspecified (docs/ra-ra.texi:3143-3148), so on the device they are max / min
// folds over the truth values and over the position of the first hit. Empty: false / true / -1 / false, like the reference's
// defaults.
template <class A> requires (is_ra<A>) bool any(A const & a) { return reduce(cast<int>(cast<bool>(a)), RACU_RED_AMAX)>0; }
template <class A> requires (is_ra<A>) bool every(A const & a) { return reduce(cast<int>(cast<bool>(a)), RACU_RED_AMIN)>0; }
// index of the first true element along axis 0 (the reference matches a rank-1 iota() against axis 0: rank > 1 "checks
// elements but only returns the first index", test/early.cc:77-82): the smallest axis-0 index that holds a true element.
template <class A> requires (is_ra<A>) dim_t index(A const & a)
{
    constexpr dim_t none = std::numeric_limits<dim_t>::max();
    dim_t const i = reduce(where(a, _0, none), RACU_RED_AMIN);
    return none==i ? dim_t(-1) : i;
}
// first position, in row-major order, where a != b decides: a < b there (false if there is none).
// key = 2 * (row-major position) + (a < b ? 0 : 1) where a != b: the smallest key is the first difference.
template <int R, class A, class B> bool lexical_compare_rank(A const & a, B const & b, std::vector<dim_t> const & sh)
{
    constexpr dim_t none = std::numeric_limits<dim_t>::max();
    auto lt01 = where(a<b, dim_t(0), dim_t(1));
    if constexpr (0==R) {
        return reduce(where(a!=b, lt01, none), RACU_RED_AMIN)==0;
    } else {
        dim_t st[R], acc = 1;
        for (int k=R-1; k>=0; --k) { st[k] = acc; acc *= sh[k]; }
        auto pos = [&]<int ... k>(std::integer_sequence<int, k ...>) { return ((TIndex<k> {}*st[k]) + ...); }(std::make_integer_sequence<int, R> {});
        dim_t const key = reduce(where(a!=b, pos*dim_t(2) + lt01, none), RACU_RED_AMIN);
        return none!=key && 0==key%2;
    }
}
template <class A, class B> requires (tomap<A, B>) bool lexical_compare(A const & a, B const & b)
{
    auto const sh = shape(a!=b);
    switch (sh.size()) {
    case 0: return lexical_compare_rank<0>(a, b, sh);
    case 1: return lexical_compare_rank<1>(a, b, sh);
    case 2: return lexical_compare_rank<2>(a, b, sh);
    case 3: return lexical_compare_rank<3>(a, b, sh);
    case 4: return lexical_compare_rank<4>(a, b, sh);
    case 5: return lexical_compare_rank<5>(a, b, sh);
    default: RACU_ASSERT(false, RACU_ERR_UNSUPPORTED, std::string("lexical_compare: rank > 5")); return false;
    }
}


// ---------------------
// views: ra/arrays.hh:84-149; ViewBig ra/expr.hh:233
// ---------------------

template <rank_t R> using Dimv = std::conditional_t<ANY==R, std::vector<Dim>, std::array<Dim, (ANY==R ? 0 : R)>>;

template <class I> constexpr int fsrc = 1;                                  // ra/ply.hh:344-348
template <int n> constexpr int fsrc<dots_t<n>> = n;
template <int n> constexpr int fsrc<insert_t<n>> = 0;
template <class I> constexpr int fdst = 1;
template <class I> requires (std::integral<I> || is_lenlike<I>) constexpr int fdst<I> = 0;
template <int n> constexpr int fdst<dots_t<n>> = n;
template <int n> constexpr int fdst<insert_t<n>> = n;

template <rank_t R, class ... I> constexpr rank_t
frombrank() // ra/ply.hh:351-357
{
    if (ANY==R) return ANY;
    constexpr int nstretch = (0 + ... + (int(UNB)==fsrc<I>));
    static_assert(nstretch<=1, "at most one dots<> in a subscript");
    int src = (0 + ... + (int(UNB)==fsrc<I> ? 0 : fsrc<I>));
    int dst = (0 + ... + (int(UNB)==fdst<I> ? 0 : fdst<I>));
    return R - src + dst; // the stretch and the implied trailing axes map 1:1
}

template <class T, rank_t R=ANY>
struct View: node_tag
{
    static_assert(is_elem<T>, "unsupported element type on the device path");
    using value_type = std::remove_cv_t<T>;
    T * cp = nullptr;
    Dimv<R> dimv {};

    constexpr View() = default;
    constexpr View(Dimv<R> const & dv, T * p): cp(p), dimv(dv) {}
    constexpr View(View const &) = default;
    template <rank_t RR> requires (RR!=R && (ANY==R || ANY==RR))
    View(View<T, RR> const & x): cp(x.cp)
    {
        if constexpr (ANY==R) dimv.assign(x.dimv.begin(), x.dimv.end());
        else { RACU_CK(x.rank()==R, "Bad rank."); for (rank_t k=0; k<R; ++k) dimv[k] = x.dimv[k]; }
    }

    constexpr rank_t rank() const { return rank_t(dimv.size()); }
    constexpr dim_t len(int k) const { return dimv[k].len; }
    constexpr dim_t step(int k) const { return k<rank() ? dimv[k].step : 0; } // ra/expr.hh:294-295
    constexpr dim_t size() const { dim_t s = 1; for (auto const & d: dimv) { if (d.len<0) return d.len; s *= d.len; } return s; }
    constexpr T * data() const { return cp; }
    std::vector<dim_t> shape() const { std::vector<dim_t> s; for (auto const & d: dimv) s.push_back(d.len); return s; }

    void flatten(Builder & b) const { b.emit(RACU_OP_PUSH, dtype_of<value_type>, b.template mem<T>(cp, dimv.data(), rank())); }

    // ---- unbeaten subscripts (index arrays, index expressions, std::vector<int>): from() -> fromu(), ra/ply.hh:461-499
    template <class I0> static constexpr bool is_unbeaten = (is_ra<I0> && !is_iota<std::decay_t<I0>>) || is_fov<std::decay_t<I0>>;
    template <class ... I>
    Gather<value_type> gather(I const & ... i) const
    {
        std::vector<Dim> bv;                                   // the source after its beaten subscripts; gathered axes kept whole
        std::vector<std::function<void (Builder &)>> fl;       // per axis of bv: the index expression (empty: the axis is traversed)
        std::vector<int> rk;                                   // per axis of bv: expression axes it takes
        dim_t pl = 0;
        int ak = 0, seen = 0;
        int const nstatic = (0 + ... + (int(UNB)==fsrc<std::decay_t<I>> ? 0 : fsrc<std::decay_t<I>>));
        auto one = [&](auto const & i0) {
            using I0 = std::decay_t<decltype(i0)>;
            if constexpr (std::integral<I0> || is_lenlike<I0>) {
                RACU_CK(ak<rank(), "Too many subscripts.");
                dim_t la = len(ak), ii = lenx(i0)(la);
                RACU_CK(0<=ii && ii<la, "Bad index " + std::to_string(ii) + " in len[" + std::to_string(ak) + "]=" + std::to_string(la) + ".");
                pl += ii*dimv[ak].step;
                ++ak;
            } else if constexpr (is_iota<I0>) {
                RACU_CK(ak<rank(), "Too many subscripts.");
                dim_t la = len(ak), n = i0.n(la), o = i0.o(la), s = i0.s(la);
                RACU_ASSERT(UNB!=n, RACU_ERR_UNSUPPORTED, "unbounded iota subscripts are not beatable (ra/ply.hh:341)");
                RACU_CK(0==n || (0<=o && o<la && 0<=o+(n-1)*s && o+(n-1)*s<la), "Bad iota subscript.");
                bv.push_back(Dim {n, s*dimv[ak].step}); fl.emplace_back(); rk.push_back(1);
                pl += o*dimv[ak].step;
                ++ak;
            } else if constexpr (requires { []<int N>(dots_t<N>){}(i0); }) {
                int n = I0::N;
                if (int(UNB)==n) n = rank()-ak-(nstatic-seen);
                RACU_CK(n>=0 && ak+n<=rank(), "Too many subscripts.");
                for (int q=0; q<n; ++q) { bv.push_back(dimv[ak++]); fl.emplace_back(); rk.push_back(1); }
            } else if constexpr (is_unbeaten<I0>) {
                RACU_CK(ak<rank(), "Too many subscripts.");
                auto node = iter(i0);
                using N = std::decay_t<decltype(node)>;
                static_assert(std::is_integral_v<typename N::value_type> && !std::is_same_v<typename N::value_type, bool>, "subscripts must be integers");
                int r = 1;
                if constexpr (!is_fov<I0>) r = int(ra::cuda::shape(node).size());
                fl.emplace_back([node](Builder & b) { node.flatten(b); b.template cast<typename N::value_type, dim_t>(); });
                bv.push_back(dimv[ak]); rk.push_back(r);
                ++ak;
            } else {
                static_assert(std::integral<I0>, "this subscript cannot be combined with index arrays on the device path");
            }
            if constexpr (int(UNB)!=fsrc<I0>) seen += fsrc<I0>;
        };
        (one(i), ...);
        for (; ak<rank(); ++ak) { bv.push_back(dimv[ak]); fl.emplace_back(); rk.push_back(1); } // trailing axes implied
        Gather<value_type> g;
        g.base = cp+pl;
        int pos = 0;
        for (std::size_t q=0; q<bv.size(); ++q) {
            dim_t const ln = bv[q].len, st = bv[q].step;
            RACU_ASSERT(ln>=0, RACU_ERR_UNSUPPORTED, "dead axes next to an index array subscript");
            if (fl[q]) g.terms.push_back({fl[q], pos, st});
            else g.terms.push_back({[ln](Builder & b) { dim_t one_ = 1; b.emit(RACU_OP_PUSH, dtype_of<dim_t>, b.seq<dim_t>(0, 1, &ln, &one_)); }, pos, st});
            pos += rk[q];
            if (ln>0) { g.lo += std::min<dim_t>(0, (ln-1)*st); g.hi += std::max<dim_t>(0, (ln-1)*st); }
        }
        RACU_CK(pos<=RACU_MAX_RANK, "Rank too large.");
        return g;
    }

    // ---- beaten subscripts: from() / fromb(), ra/ply.hh:367-435,503-551
    template <class ... I> requires ((is_unbeaten<I> || ...))
    auto operator()(I const & ... i) const { return gather(i ...); }
    template <class ... I> requires (!(is_unbeaten<I> || ...))
    auto operator()(I const & ... i) const
    {
        constexpr rank_t BR = frombrank<R, std::decay_t<I> ...>();
        View<T, BR> b;
        std::vector<Dim> bv;
        dim_t pl = 0;
        int ak = 0, seen = 0; // seen: source axes consumed by the fixed-size subscripts so far (sizes dots<>: ra/ply.hh:416-422)
        int const nstatic = (0 + ... + (int(UNB)==fsrc<std::decay_t<I>> ? 0 : fsrc<std::decay_t<I>>));
        auto one = [&](auto const & i0) {
            using I0 = std::decay_t<decltype(i0)>;
            if constexpr (std::integral<I0> || is_lenlike<I0>) { // :390-398
                RACU_CK(ak<rank(), "Too many subscripts.");
                dim_t la = len(ak), ii = lenx(i0)(la);
                RACU_CK((0<=ii && ii<la) || (UNB==la && 0==dimv[ak].step), "Bad index " + std::to_string(ii) + " in len[" + std::to_string(ak) + "]=" + std::to_string(la) + ".");
                pl += ii*dimv[ak].step;
                ++ak;
            } else if constexpr (is_iota<I0>) { // :399-415
                RACU_CK(ak<rank(), "Too many subscripts.");
                dim_t la = len(ak), n = i0.n(la), o = i0.o(la), s = i0.s(la);
                RACU_ASSERT(UNB!=n, RACU_ERR_UNSUPPORTED, "unbounded iota subscripts are not beatable (ra/ply.hh:341)");
                RACU_CK(0==n || (0<=o && o<la && 0<=o+(n-1)*s && o+(n-1)*s<la),
                        "Bad iota len " + std::to_string(n) + " step " + std::to_string(s) + " in len[" + std::to_string(ak) + "]=" + std::to_string(la) + "."); // :407
                bv.push_back(Dim {n, s*dimv[ak].step});
                pl += o*dimv[ak].step;
                ++ak;
            } else if constexpr (requires { []<int N>(dots_t<N>){}(i0); }) { // :416-422
                int n = I0::N;
                if (int(UNB)==n) n = rank()-ak-(nstatic-seen);
                RACU_CK(n>=0 && ak+n<=rank(), "Too many subscripts.");
                for (int q=0; q<n; ++q) bv.push_back(dimv[ak++]);
            } else if constexpr (requires { []<int N>(insert_t<N>){}(i0); }) { // :423-427
                for (int q=0; q<I0::N; ++q) bv.push_back(Dim {UNB, 0});
            } else {
                static_assert(std::integral<I0>, "unbeaten (gather) subscripts are not on the device path");
            }
            if constexpr (int(UNB)!=fsrc<I0>) seen += fsrc<I0>;
        };
        (one(i), ...);
        for (; ak<rank(); ++ak) bv.push_back(dimv[ak]); // trailing axes implied, :374-381
        RACU_CK(rank_t(bv.size())<=RACU_MAX_RANK, "Rank too large.");
        if constexpr (ANY==BR) b.dimv = bv;
        else { RACU_CK(rank_t(bv.size())==BR, "Bad subscripts."); for (rank_t k=0; k<BR; ++k) b.dimv[k] = bv[k]; }
        b.cp = cp+pl;
        return b;
    }

    // ---- assignment ops: for_each(y OP= x) with the destination in the agreement (ra/expr.hh:50-60, ra/arrays.hh:131-135)
    template <class X> void assign(int op, X const & x) const
    {
        static_assert(!std::is_const_v<T>, "assignment to const view");
        Builder b;
        b.d.dst = b.template mem<T>(cp, dimv.data(), rank());
        b.d.assign = op;
        iter(x).flatten(b);
        ck_status(racu_map(&b.d, stream()));
    }
    View const & operator=(View const & x) const { assign(RACU_ASSIGN_SET, x); return *this; } // views copy contents: ra/arrays.hh:133
#define RACU_ASSIGNOP(OP, CODE) \
    template <class X> requires (is_operand<X>) View const & operator OP(X const & x) const { assign(CODE, x); return *this; }
    RACU_ASSIGNOP(=, RACU_ASSIGN_SET) RACU_ASSIGNOP(+=, RACU_ASSIGN_ADD) RACU_ASSIGNOP(-=, RACU_ASSIGN_SUB)
    RACU_ASSIGNOP(*=, RACU_ASSIGN_MUL) RACU_ASSIGNOP(/=, RACU_ASSIGN_DIV)
#undef RACU_ASSIGNOP

    // ---- host access (copies through the device boundary; for tests and small results)
    std::vector<value_type> to_host() const; // row-major ravel of the view
    operator value_type () const requires (0==R) { value_type v; ck_status(racu_memcpy_d2h(&v, cp, sizeof(v), stream())); return v; }
};


// ---------------------
// owning arrays: ra/arrays.hh:252-327
// ---------------------

template <class T, rank_t R=ANY>
struct Big: View<T, R>
{
    using V = View<T, R>;
    using V::cp, V::dimv;

    Big() = default;
    ~Big() { release(); }
    Big(Big && x) noexcept: V(x.dimv, x.cp) { x.cp = nullptr; }
    Big(Big const & x): V() { init(x.shape()); if (V::size()>0) ck_status(racu_memcpy_d2d(cp, x.cp, sizeof(T)*size_t(V::size()), stream())); } // defaulted copy = new storage: ra/arrays.hh:275
    Big & operator=(Big && x) noexcept { if (this!=&x) { release(); cp = x.cp; dimv = x.dimv; x.cp = nullptr; } return *this; }
    Big & operator=(Big const & x) { if (this!=&x) { Big y(x); *this = std::move(y); } return *this; }

    // Array(shape, none | value | expr): ra/arrays.hh:296-298
    template <class S> requires (requires (S s) { s.begin(); s.size(); }) Big(S const & s, none_t) { init(s); }
    Big(std::initializer_list<dim_t> s, none_t) { init(s); }
    template <class X> requires (is_operand<X>) Big(std::initializer_list<dim_t> s, X const & x) { init(s); view() = x; }
    template <class S, class X> requires (requires (S s) { s.begin(); s.size(); } && is_operand<X>) Big(S const & s, X const & x) { init(s); view() = x; }
    // from row-major host data
    Big(std::initializer_list<dim_t> s, std::initializer_list<T> x) { init(s); from_host(x.begin(), dim_t(x.size())); }
    // from an expression: shape by agreement (ra/arrays.hh:297)
    template <class X> requires (is_ra<X> && !std::is_base_of_v<Big, std::decay_t<X>>) Big(X const & x) { init(ra::cuda::shape(iter(x))); view() = x; }
    // rank 1 from values
    Big(std::initializer_list<T> x) requires (1==R || ANY==R) { init(std::array<dim_t, 1> {dim_t(x.size())}); from_host(x.begin(), dim_t(x.size())); }

    V const & view() const { return *this; }
    template <class X> requires (is_operand<X> && !std::is_base_of_v<Big, std::decay_t<X>>) Big & operator=(X const & x) { view() = x; return *this; } // contents, like a view: ra/arrays.hh:276-281
#define RACU_ASSIGNOP(OP) template <class X> requires (is_operand<X>) Big & operator OP(X const & x) { view() OP x; return *this; }
    RACU_ASSIGNOP(+=) RACU_ASSIGNOP(-=) RACU_ASSIGNOP(*=) RACU_ASSIGNOP(/=)
#undef RACU_ASSIGNOP

    void from_host(T const * p, dim_t n)
    {
        RACU_CK(n==V::size(), "Bad size " + std::to_string(n) + ", need " + std::to_string(V::size()) + ".");
        if (n>0) { ck_status(racu_memcpy_h2d(cp, p, sizeof(T)*size_t(n), stream())); ck_status(racu_stream_sync(stream())); }
    }

private:
    template <class S> void init(S const & s)
    {
        if constexpr (ANY==R) dimv.resize(s.size()); else RACU_CK(rank_t(s.size())==R, "Bad rank for shape.");
        rank_t k = 0;
        for (auto l: s) { RACU_CK(l>=0, "Bad len."); dimv[k++].len = dim_t(l); }
        dim_t st = 1; // row-major steps: filldimv, ra/expr.hh:199-219
        for (rank_t j=V::rank(); --j>=0;) { dimv[j].step = st; st *= dimv[j].len; }
        void * p = nullptr;
        ck_status(racu_alloc(&p, sizeof(T)*size_t(st>0 ? st : 1)));
        cp = static_cast<T *>(p);
    }
    void release() { if (cp) { racu_stream_sync(stream()); racu_free(cp); cp = nullptr; } }
};

template <class T, rank_t R>
std::vector<typename View<T, R>::value_type>
View<T, R>::to_host() const
{
    dim_t n = size();
    RACU_CK(n>=0, "Unbounded size.");
    std::vector<dim_t> sh = shape();
    Big<value_type, ANY> tmp(sh, none);
    View<value_type, ANY> dst = tmp;
    dst = View<T, ANY>(*this);
    std::vector<value_type> out;
    if constexpr (std::is_same_v<value_type, bool>) {
        std::vector<unsigned char> raw(n);
        if (n>0) ck_status(racu_memcpy_d2h(raw.data(), tmp.cp, size_t(n), stream()));
        out.assign(raw.begin(), raw.end());
    } else {
        out.resize(n);
        if (n>0) ck_status(racu_memcpy_d2h(out.data(), tmp.cp, sizeof(value_type)*size_t(n), stream()));
    }
    return out;
}

// Materialise an expression (ra::concrete, ra/arrays.hh:371-401).
template <class E> requires (is_ra<E>)
auto concrete(E const & e) { return Big<typename node_t<E>::value_type, ANY>(e); }


// ---------------------
// view ops: ra/arrays.hh:408-462
// ---------------------

template <class T, rank_t R>
View<T, R> reverse(View<T, R> const & a, int k=0) // :408-428
{
    RACU_CK(0<=k && k<a.rank(), "Bad axis " + std::to_string(k) + " for rank " + std::to_string(a.rank()) + ".");
    RACU_CK(UNB!=a.len(k), "Cannot reverse unbounded axis " + std::to_string(k) + ".");
    View<T, R> b = a;
    b.cp = a.cp + (0==a.len(k) ? 0 : a.dimv[k].step*(a.len(k)-1));
    b.dimv[k].step *= -1;
    return b;
}

inline void
transpose_dims(int const * s, rank_t ns, Dim const * src, Dim * dst, rank_t nd) // :431-442
{
    for (rank_t k=0; k<nd; ++k) dst[k] = Dim {UNB, 0};
    for (rank_t k=0; k<ns; ++k) {
        dst[s[k]].step += src[k].step;
        dst[s[k]].len = dst[s[k]].len>=0 ? std::min(dst[s[k]].len, src[k].len) : src[k].len;
    }
}

// transpose<1, 0>(a): axis k of a goes to axis s[k] of the result; repeated targets give diagonals (:444-461).
// ra::stencil(a, lo, hi) (ra/arrays.hh:615-630): a view of rank 2*rank(a) over the same memory; axes [0, r) run over the positions
// whose neighbourhood [-lo, +hi] is inside a, axes [r, 2r) over the neighbourhood: s[i..., d...] = a[i+d ...].
template <class T, rank_t R>
View<T, ANY> stencil(View<T, R> const & a, dim_t lo, dim_t hi)
{
    RACU_CK(lo>=0 && hi>=0, "Bad stencil bounds.");
    View<T, ANY> s;
    s.cp = a.cp;
    rank_t const r = a.rank();
    RACU_CK(2*r<=RACU_MAX_RANK, "Rank too large.");
    s.dimv.resize(2*r);
    for (rank_t k=0; k<r; ++k) {
        RACU_CK(a.len(k)>=lo+hi, "Stencil is too large for array.");
        s.dimv[k] = Dim {a.len(k)-lo-hi, a.dimv[k].step};
        s.dimv[r+k] = Dim {lo+hi+1, a.dimv[k].step};
    }
    return s;
}
template <class A> requires (requires (A const & a) { a.view(); })
auto stencil(A const & a, dim_t lo, dim_t hi) { return stencil(a.view(), lo, hi); }

// The time loop of bench/bench-stencil3.cc (`Anext(I, J, K) = -6*A(I, J, K) + A(I+1, J, K) + ...; std::swap(A.cp, Anext.cp);`, :55-64 and
// :117-121, T times) handed to the library as ONE call (racu_stencil_steps): on return a and anext hold exactly what the loop leaves in the two
// buffers - state T in anext if T is odd, else in a; the state before it in the other; boundary cells untouched - but the steps were taken
// two per pass over the grid where the arrays allow it. Returns true when the last state is in anext (the caller swaps like the loop would).
template <class T, rank_t R>
bool stencil_steps(View<T, R> const & a, View<T, R> const & anext, int steps)
{
    static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "stencil sweeps are for float and double");
    rank_t const r = a.rank();
    RACU_CK(r==anext.rank() && r>=1 && r<=3, "Bad shapes for stencil_steps.");
    racu_stencil_desc s {};
    s.kind = 3==r ? RACU_STENCIL_3D7 : 2==r ? RACU_STENCIL_2D5 : RACU_STENCIL_1D3;
    s.dtype = dtype_of<T>; s.rank = int32_t(r);
    s.src = a.cp; s.dst = anext.cp;
    for (rank_t k=0; k<r; ++k) {
        RACU_CK(a.len(k)==anext.len(k), "Bad shapes for stencil_steps.");
        s.len[k] = a.len(k); s.src_stride[k] = a.dimv[k].step; s.dst_stride[k] = anext.dimv[k].step;
    }
    ck_status(racu_stencil_steps(&s, int32_t(steps), nullptr, stream()));
    return 0!=(steps&1);
}
template <class A, class B> requires (requires (A & a, B & b) { a.view(); b.view(); })
bool stencil_steps(A & a, B & anext, int steps) { return stencil_steps(a.view(), anext.view(), steps); }

// refmin / refmax (ra/ra.hh:326-346): a REFERENCE to the first smallest / largest element in row-major order - here a rank-0 view of
// it, assignable like the reference's `double &`. Two folds: the extreme value, then the smallest row-major position holding it.
template <class T, rank_t R>
View<T, 0> ref_extreme(View<T, R> const & a, int redop)
{
    RACU_CK(a.size()>0, "refmin requires nonempty argument.");
    using V0 = std::remove_cv_t<T>;
    V0 const m = reduce(a, redop);
    rank_t const r = a.rank();
    RACU_ASSERT(r<=5, RACU_ERR_UNSUPPORTED, std::string("refmin / refmax: rank > 5"));
    constexpr dim_t none = std::numeric_limits<dim_t>::max();
    dim_t st[5] = {0, 0, 0, 0, 0}, acc = 1, first = 0;
    for (rank_t k=r-1; k>=0; --k) { st[k] = acc; acc *= a.len(k); }
    auto hit = [&](auto const & pos) { return reduce(where(a==m, pos, none), RACU_RED_AMIN); };
    switch (r) {
    case 0: break;
    case 1: first = hit(_0*st[0]); break;
    case 2: first = hit(_0*st[0] + _1*st[1]); break;
    case 3: first = hit(_0*st[0] + _1*st[1] + _2*st[2]); break;
    case 4: first = hit(_0*st[0] + _1*st[1] + _2*st[2] + _3*st[3]); break;
    default: first = hit(_0*st[0] + _1*st[1] + _2*st[2] + _3*st[3] + _4*st[4]); break;
    }
    View<T, 0> v;
    dim_t off = 0;
    for (rank_t k=0; k<r; ++k) { off += (first/st[k])*a.dimv[k].step; first %= st[k]; }
    v.cp = a.cp + off;
    return v;
}
template <class A> requires (requires (A const & a) { a.view(); }) auto refmin(A const & a) { return ref_extreme(a.view(), RACU_RED_AMIN); }
template <class A> requires (requires (A const & a) { a.view(); }) auto refmax(A const & a) { return ref_extreme(a.view(), RACU_RED_AMAX); }
template <class T, rank_t R> auto refmin(View<T, R> const & a) { return ref_extreme(a, RACU_RED_AMIN); }
template <class T, rank_t R> auto refmax(View<T, R> const & a) { return ref_extreme(a, RACU_RED_AMAX); }

// Whole-expression reduction whose result STAYS on the device: dst is a rank-0 view of the expression's type. The reference
// returns the scalar by value (ra/ra.hh:348-401); chains of reductions and maps then need no host round trip per scalar.
template <class T, rank_t R, class E> requires (is_ra<E>)
void reduce_into(View<T, R> const & dst, E const & e, int redop=RACU_RED_SUM)
{
    static_assert(std::is_same_v<std::remove_cv_t<T>, typename E::value_type>, "reduce_into: destination type must be the expression's");
    RACU_CK(0==dst.rank(), "reduce_into needs a rank-0 destination.");
    Builder b;
    e.flatten(b);
    ck_status(racu_reduce_dev(&b.d, redop, dst.cp, stream()));
}

// Conjugate gradient after Hestenes and Stiefel: examples/cghs.hh:15-45 with rho, sig, tau, t, gam resident on the device
// (rank-0 views filled by reduce_into, combined by rank-0 maps): one host read per iteration, the convergence test the
// reference makes at the top of the loop. mult(v, w): w = A v. work: 3 x shape(b). Returns the iteration count.
template <class Mult, class B, class X, class W>
int cghs(Mult && mult, B const & b, X & x, W & work, double eps, int max_its=std::numeric_limits<int>::max())
{
    Big<double, 1> scal({8}, 0.);
    work = 0.;
    auto g = work(0), r = work(1), p = work(2);
    auto err = scal(0), gg = scal(1), rho = scal(2), sig = scal(3), tau = scal(4), t = scal(5), gam = scal(6);
    reduce_into(err, sqrm(b));
    err *= eps*eps;                                          // :26
    mult(x, g);
    g = b-g;
    r = g;
    int its = 0;
    for (; its<max_its; ++its) {
        reduce_into(gg, sqrm(g));
        auto const e_gg = scal(iota(2)).to_host();           // reduce_sqrm(g) > err   (:33)
        if (!(e_gg[1]>e_gg[0])) break;
        mult(r, p);
        reduce_into(rho, sqrm(p));                           // :35-37
        reduce_into(sig, r*p);
        reduce_into(tau, g*r);
        t = tau/sig;                                         // :38
        gam = (t*t*rho - tau)/tau;                           // :39
        x += t*r;                                            // :40-42
        g -= t*p;
        r = r*gam + g;
    }
    return its;
}

// at(a, i) (ra/ra.hh:228-239): i holds index TUPLES along its last axis (rank(a) coordinates each); the result has the shape of
// i without that axis: at(a, i)[...] = a[i[..., 0], i[..., 1], ...].
template <class T, rank_t R, class J, rank_t RI> requires (std::is_integral_v<J>)
Gather<std::remove_cv_t<T>> at(View<T, R> const & a, View<J, RI> const & i)
{
    RACU_CK(i.rank()>=1 && i.len(i.rank()-1)==a.rank(), "at(): the last axis of i must hold rank(a) coordinates");
    Gather<std::remove_cv_t<T>> g;
    g.base = a.cp;
    for (rank_t k=0; k<a.rank(); ++k) {
        View<J, ANY> ik;                                     // i(dots, k): the k-th coordinate of every tuple
        ik.cp = i.cp + k*i.dimv[i.rank()-1].step;
        ik.dimv.assign(i.dimv.begin(), i.dimv.end()-1);
        dim_t const ln = a.len(k), st = a.dimv[k].step;
        g.terms.push_back({[ik](Builder & b) { ik.flatten(b); b.template cast<std::remove_cv_t<J>, dim_t>(); }, 0, st});
        if (ln>0) { g.lo += std::min<dim_t>(0, (ln-1)*st); g.hi += std::max<dim_t>(0, (ln-1)*st); }
    }
    return g;
}
template <class A, class I> requires (requires (A const & a, I const & i) { a.view(); i.view(); })
auto at(A const & a, I const & i) { return at(a.view(), i.view()); }

template <int ... s, class T, rank_t R> requires (sizeof...(s)>0)
auto transpose(View<T, R> const & a)
{
    constexpr int ns = sizeof...(s);
    constexpr int sv[ns>0 ? ns : 1] = { s ... };
    constexpr rank_t nd = []{ int m = 0; for (int k=0; k<ns; ++k) m = std::max(m, sv[k]+1); return m; }();
    RACU_CK(a.rank()==ns, "Bad rank " + std::to_string(a.rank()) + " for " + std::to_string(ns) + " transposed axes.");
    View<T, nd> b;
    b.cp = a.cp;
    transpose_dims(sv, ns, a.dimv.data(), b.dimv.data(), nd);
    return b;
}
template <class T, rank_t R>
auto transpose(View<T, R> const & a, std::initializer_list<int> s)
{
    RACU_CK(a.rank()==rank_t(s.size()), "Bad rank " + std::to_string(a.rank()) + " for " + std::to_string(s.size()) + " transposed axes.");
    int nd = 0;
    for (int k: s) { RACU_CK(k>=0 && k<RACU_MAX_RANK, "Bad transposed axis."); nd = std::max(nd, k+1); }
    View<T, ANY> b;
    b.cp = a.cp;
    b.dimv.resize(nd);
    transpose_dims(s.begin(), rank_t(s.size()), a.dimv.data(), b.dimv.data(), nd);
    return b;
}
template <class T, rank_t R> auto transpose(View<T, R> const & a) { return transpose<1, 0>(a); } // :461
template <class T, rank_t R> auto diag(View<T, R> const & a) { return transpose<0, 0>(a); }       // :462


// ---------------------
// cell iterators and rank conjunction: iter<k>(a) (ra/expr.hh:317-328), scalar(a) (ra/expr.hh:128), wrank<cr ...>(op) (ra/expr.hh:469-499).
//
// In the reference these build iterators over the k-cells of an array / apply an op to cells of given ranks with the FRAMES (the axes
// in front of the cells) matched by prefix, and any callable may sit inside. On the device the callable has to be one of the
// enumerated assignment functors below, and each spelling LOWERS to the one fused descriptor the planner already knows: every argument
// gets dead (inserted) axes where its frame is shorter than the longest one, and the op is then the ordinary prefix-matched
// assignment. So the seven spellings of "sum the rows" of test/reduction.cc:182-222 and bench/bench-sum-rows.cc:62-81
//     sum(iter<1>(a));  iter<1>(c) += iter<1>(a);  scalar(c) += iter<1>(a);  for_each(wrank<1, 1>(add_assign), c, a);
//     for_each(wrank<1, 1>(wrank<0, 0>(add_assign)), c, a);  c += transpose(a);  c(insert<1>) += a
// all are the fold `c(insert<1>) += a` (sflat_fold_outer), and gemm1 of bench/bench-gemm.cc:25-30,
//     for_each(wrank<1, 2, 1>(wrank<0, 1, 1>(maybe_fma)), a, b, c)
// is c(all, insert<1>) += a * b(insert<1>): the contraction kernel.
// ---------------------

template <class T> using ViewAny = View<T, ANY>;

// `frame` axes of v are kept, then `extra` dead axes, then its cells.
template <class T, rank_t R>
ViewAny<T> insert_after_frame(View<T, R> const & v, int frame, int extra)
{
    ViewAny<T> out;
    out.cp = v.cp;
    RACU_CK(0<=frame && frame<=v.rank() && extra>=0, "Bad cell rank.");
    out.dimv.assign(v.dimv.begin(), v.dimv.begin()+frame);
    for (int q=0; q<extra; ++q) out.dimv.push_back(Dim {UNB, 0});
    out.dimv.insert(out.dimv.end(), v.dimv.begin()+frame, v.dimv.end());
    RACU_CK(out.rank()<=RACU_MAX_RANK, "Rank too large.");
    return out;
}

template <class T, rank_t R, int k>
struct CellIter // iter<k>(a): the k-cells of a; k < 0 counts frame axes (ra/expr.hh:323: cell rank = rank + k)
{
    View<T, R> v;
    int cell_rank() const { int c = k>=0 ? k : v.rank()+k; RACU_CK(0<=c && c<=v.rank(), "Bad cell rank."); return c; }
    int frame_rank() const { return v.rank()-cell_rank(); }
    template <class T2, rank_t R2, int k2> void assign(int op, CellIter<T2, R2, k2> const & x) const
    {
        int const F = std::max(frame_rank(), x.frame_rank());
        insert_after_frame(v, frame_rank(), F-frame_rank()).assign(op, insert_after_frame(x.v, x.frame_rank(), F-x.frame_rank()));
    }
#define RACU_ASSIGNOP(OP, CODE) \
    template <class T2, rank_t R2, int k2> CellIter const & operator OP(CellIter<T2, R2, k2> const & x) const { assign(CODE, x); return *this; }
    RACU_ASSIGNOP(=, RACU_ASSIGN_SET) RACU_ASSIGNOP(+=, RACU_ASSIGN_ADD) RACU_ASSIGNOP(-=, RACU_ASSIGN_SUB)
    RACU_ASSIGNOP(*=, RACU_ASSIGN_MUL) RACU_ASSIGNOP(/=, RACU_ASSIGN_DIV)
#undef RACU_ASSIGNOP
};

template <int k, class T, rank_t R> auto iter(View<T, R> const & a) { return CellIter<T, R, k> {a}; }
template <int k, class A> requires (requires (A const & a) { a.view(); }) auto iter(A const & a) { return iter<k>(a.view()); }
// ra::scalar(c): c as ONE cell (every cell of the other side is combined into the whole of it): scalar(c) += iter<1>(a)
template <class T, rank_t R> auto scalar(View<T, R> const & a) { return CellIter<T, R, (ANY==R ? 0 : R)> {a}; }
template <class A> requires (requires (A const & a) { a.view(); }) auto scalar(A const & a) { return scalar(a.view()); }
template <class T> struct CellIter<T, ANY, 0> // (dynamic rank: scalar() means the whole array, iter<0> its elements; told apart by a flag)
{
    View<T, ANY> v;
    bool whole = true;
    int cell_rank() const { return whole ? v.rank() : 0; }
    int frame_rank() const { return v.rank()-cell_rank(); }
    template <class X> void assign(int op, X const & x) const
    {
        int const F = std::max(frame_rank(), x.frame_rank());
        insert_after_frame(v, frame_rank(), F-frame_rank()).assign(op, insert_after_frame(x.v, x.frame_rank(), F-x.frame_rank()));
    }
#define RACU_ASSIGNOP(OP, CODE) template <class X> CellIter const & operator OP(X const & x) const { assign(CODE, x); return *this; }
    RACU_ASSIGNOP(+=, RACU_ASSIGN_ADD) RACU_ASSIGNOP(-=, RACU_ASSIGN_SUB) RACU_ASSIGNOP(*=, RACU_ASSIGN_MUL) RACU_ASSIGNOP(/=, RACU_ASSIGN_DIV)
#undef RACU_ASSIGNOP
};

// sum / prod over the cells (test/reduction.cc:189,197-201): an array of the cell's shape; no cells: 0 / 1.
template <class T, rank_t R, int k>
auto reduce_cells(CellIter<T, R, k> const & it, int op, std::remove_cv_t<T> start)
{
    using V0 = std::remove_cv_t<T>;
    std::vector<dim_t> sh;
    for (int q=it.frame_rank(); q<it.v.rank(); ++q) sh.push_back(it.v.len(q));
    Big<V0, ANY> c(sh, start);
    CellIter<V0, ANY, 0> whole {c.view(), true};
    whole.assign(op, it);
    return c;
}
template <class T, rank_t R, int k> auto sum(CellIter<T, R, k> const & it) { return reduce_cells(it, RACU_ASSIGN_ADD, std::remove_cv_t<T>(0)); }
template <class T, rank_t R, int k> auto prod(CellIter<T, R, k> const & it) { return reduce_cells(it, RACU_ASSIGN_MUL, std::remove_cv_t<T>(1)); }

// The callables a device for_each takes: y OP= x on cells (destination FIRST, as in for_each([](auto & c, auto a) { c += a; }, c, a)),
// and maybe_fma(a, b, c): c += a*b (ra/ra.hh:364-374; destination LAST, as bench/bench-gemm.cc:25-30 writes it).
template <int CODE> struct assign_fn { static constexpr int code = CODE; };
constexpr assign_fn<RACU_ASSIGN_SET> set_assign {};
constexpr assign_fn<RACU_ASSIGN_ADD> add_assign {};
constexpr assign_fn<RACU_ASSIGN_SUB> sub_assign {};
constexpr assign_fn<RACU_ASSIGN_MUL> mul_assign {};
constexpr assign_fn<RACU_ASSIGN_DIV> div_assign {};
struct maybe_fma_fn {};
constexpr maybe_fma_fn maybe_fma {};

template <class Op, int ... cr> struct Wrank { Op op; };
template <int ... cr, class Op> constexpr auto wrank(Op const & op) { return Wrank<Op, cr ...> {op}; }

namespace detail {
// one rank-conjunction level: argument i keeps pos[i] axes as frame already; its cells at this level have rank cr[i]
template <std::size_t N> void
conjoin(std::array<std::vector<Dim>, N> & dv, std::array<int, N> & pos, std::array<int, N> const & cr)
{
    int f[N], F = 0;
    for (std::size_t i=0; i<N; ++i) {
        int const left = int(dv[i].size())-pos[i];
        int const c = cr[i]>=0 ? cr[i] : left+cr[i];
        RACU_CK(0<=c && c<=left, "Bad cell rank.");
        f[i] = left-c; F = std::max(F, f[i]);
    }
    for (std::size_t i=0; i<N; ++i) {
        dv[i].insert(dv[i].begin()+pos[i]+f[i], std::size_t(F-f[i]), Dim {UNB, 0});
        pos[i] += F;
        RACU_CK(int(dv[i].size())<=RACU_MAX_RANK, "Rank too large.");
    }
}
template <class Op> struct final_op { using type = Op; };
template <class Op, int ... cr> struct final_op<Wrank<Op, cr ...>> { using type = typename final_op<Op>::type; };
template <std::size_t N, class Op> void conjoin_all(Op const &, std::array<std::vector<Dim>, N> &, std::array<int, N> &) {}
template <std::size_t N, class Op, int ... cr> void
conjoin_all(Wrank<Op, cr ...> const & w, std::array<std::vector<Dim>, N> & dv, std::array<int, N> & pos)
{
    static_assert(sizeof...(cr)==N, "wrank<> needs one cell rank per argument");
    conjoin<N>(dv, pos, std::array<int, N> {cr ...});
    conjoin_all<N>(w.op, dv, pos);
}
} // namespace detail

// for_each(wrank<cr ...>(op), args ...) for the enumerated ops (ra/ply.hh:168 + ra/expr.hh:469-499).
template <class Op, int ... cr, class ... A>
void for_each(Wrank<Op, cr ...> const & w, A const & ... a)
{
    constexpr std::size_t N = sizeof...(A);
    using F = typename detail::final_op<Op>::type;
    auto views = std::make_tuple(ViewAny<typename node_t<A>::value_type>(iter(a)) ...);
    std::array<std::vector<Dim>, N> dv;
    std::array<int, N> pos {};
    [&]<std::size_t ... I>(std::index_sequence<I ...>) { ((dv[I] = std::get<I>(views).dimv), ...); }(std::make_index_sequence<N> {});
    detail::conjoin_all<N>(w, dv, pos);
    [&]<std::size_t ... I>(std::index_sequence<I ...>) { ((std::get<I>(views).dimv = dv[I]), ...); }(std::make_index_sequence<N> {});
    if constexpr (std::is_same_v<F, maybe_fma_fn>) {
        static_assert(3==N, "maybe_fma(a, b, c)");
        std::get<2>(views) += std::get<0>(views) * std::get<1>(views);
    } else {
        static_assert(2==N, "y OP= x");
        std::get<0>(views).assign(F::code, std::get<1>(views));
    }
}
// without a rank conjunction the op applies element by element
template <int CODE, class Y, class X> void for_each(assign_fn<CODE>, Y const & y, X const & x) { iter(y).assign(CODE, x); }

// ---------------------
// BLAS-like: ra/ra.hh:403-456. Same expressions as the reference; the library recognises them and runs the contraction kernels.
// ---------------------

template <class A, class B, class C_>
decltype(auto) gemm(A const & a, B const & b, C_ && c) { c(all, insert<1>) += a * b(insert<1>); return std::forward<C_>(c); } // :407
template <class A, class B, class C_>
decltype(auto) gemv(A const & a, B const & b, C_ && c) { c.view() += a * b(insert<1>); return std::forward<C_>(c); }          // :418
template <class A, class B, class C_>
decltype(auto) gevm(A const & a, B const & b, C_ && c) { c(insert<1>) += a * b; return std::forward<C_>(c); }                 // :431

template <class A, class B> auto gemm(A const & a, B const & b) // :440-444
{
    using T = decltype(std::declval<typename A::value_type>()*std::declval<typename B::value_type>());
    Big<T, 2> c({a.len(0), b.len(1)}, T());
    gemm(a, b, c);
    return c;
}
template <class A, class B> auto gemv(A const & a, B const & b) // :446-450
{
    using T = decltype(std::declval<typename A::value_type>()*std::declval<typename B::value_type>());
    Big<T, 1> c({a.len(0)}, T());
    gemv(a, b, c);
    return c;
}
template <class A, class B> auto gevm(A const & a, B const & b) // :452-456
{
    using T = decltype(std::declval<typename A::value_type>()*std::declval<typename B::value_type>());
    Big<T, 1> c({b.len(1)}, T());
    gevm(a, b, c);
    return c;
}

} // namespace ra::cuda
