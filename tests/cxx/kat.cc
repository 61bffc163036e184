// kat.cc - the reference's known-answer tests for the traversal path, written against ra/cuda.hh so that they read like
// the reference's own test files. Linked twice by tests/test_cxx_header.py: against oracle/libracu_oracle.so (the racu_* ABI
// over host memory, run by the CPU oracle: checks this header without a GPU) and against ra_ra_b200/libracu.so (-m gpu).
// Values and shapes are the reference's (file:line cited per section); no reference code is used.
#include <numeric>

#include "harness.hh"

namespace rc = ra::cuda;
using real = double;
template <class T> using vec = std::vector<T>;

template <class T, rc::rank_t R> rc::Big<T, R> iota_filled(std::initializer_list<rc::dim_t> s, T first)
{
    rc::Big<T, R> a(s, rc::none);
    vec<T> h(a.size());
    std::iota(h.begin(), h.end(), first);
    a.from_host(h.data(), rc::dim_t(h.size()));
    return a;
}

int main()
{
    TestRecorder tr;
    tr.section("traversal - Map with scalar: test/ply.cc:31-44");
    {
        auto a = iota_filled<real, 2>({3, 2}, 0);
        rc::Big<real, 2> c({3, 2}, 0.);
        c = a - 3.0;
        tr.test_eq(vec<real> {-3, -2, -1, 0, 1, 2}, c);
    }
    tr.section("traversal - int neg / mul: test/ply.cc:51-61");
    {
        rc::Big<int, 1> A {1, 2, 3}, B {+1, -2, +3}, C({3}, 0);
        C = -A;
        tr.test_eq(vec<int> {-1, -2, -3}, C);
        C = A*B;
        tr.test_eq(vec<int> {1, -4, 9}, C);
    }
    tr.section("rank matching - rank 3 vs rank 2 prefix: test/ply.cc:262-288");
    {
        auto a = iota_filled<real, 3>({3, 2, 4}, 1);
        auto b = iota_filled<real, 2>({3, 2}, 1);
        vec<real> check { 0, 1, 2, 3, 3, 4, 5, 6, 6, 7, 8, 9, 9, 10, 11, 12, 12, 13, 14, 15, 15, 16, 17, 18 };
        rc::Big<real, 3> c0(a-b);
        tr.test_eq(check, c0);
        rc::Big<real, 3> c({3, 2, 4}, 0.);
        c = -(b-a);
        tr.test_eq(check, c);
        tr.test(rc::shape(a-b)==vec<rc::dim_t> {3, 2, 4}, "shape of expression");
    }
    tr.section("mismatched steps with transpose: test/ra-1.cc:183-203");
    {
        auto a = iota_filled<int, 2>({2, 3}, 1);
        auto b = iota_filled<int, 2>({3, 2}, 1);
        rc::Big<int, 2> c({2, 3}, 0);
        c = a - rc::transpose<1, 0>(b);
        tr.test_eq(vec<int> {0, -1, -2, 2, 1, 0}, c);
        c = 0;
        rc::transpose<1, 0>(c) = rc::transpose<1, 0>(a) - b;
        tr.test_eq(vec<int> {0, -1, -2, 2, 1, 0}, c);
    }
    tr.section("operators: test/operators.cc:122-137");
    {
        rc::Big<int, 2> a({3, 2}, { 1, 2, 3, 20, 5, 6 });
        rc::Big<int, 1> b { 10, 20, 30 };
        tr.test_eq(vec<int> {11, 12, 23, 40, 35, 36}, a+b);
        tr.test_eq(vec<bool> {false, false, false, true, false, false}, a==b);
        tr.test_eq(vec<bool> {false, false, false, true, false, false}, !(a!=b));
    }
    tr.section("README first example: examples/read-me.cc:26-37");
    {
        rc::Big<float> A({2, 4}, {1, 2, 3, 4, 5, 6, 7, 8});
        rc::Big<float, 2> B({2, 4}, {1, 2, 3, 4, 5, 6, 7, 8});
        rc::Big<float, 2> C({2, 4}, {1, 2, 3, 4, 5, 6, 7, 8});
        B += A + C + std::vector {100., 200.};
        B(rc::all, rc::iota(rc::len/2, rc::len/2)) *= -1;
        tr.test_eq(vec<float> {103, 106, -109, -112, 215, 218, -221, -224}, B);
    }
    tr.section("rank extension with std::vector<float>: examples/read-me.cc:162-165");
    {
        rc::Big<float, 2> B({2, 2}, {1, 2, 3, 4});
        B += std::vector<float> {10, 20};
        tr.test_eq(vec<float> {11, 12, 23, 24}, B);
    }
    tr.section("C = A*B, D += A*B: test/ra-12.cc:30-55, examples/read-me.cc:82-91");
    {
        rc::Big<float, 2> A({2, 3}, {1, 2, 3, 1, 2, 3});
        rc::Big<float, 1> B {-1, +1};
        rc::Big<float, 2> C({2, 3}, 99.f);
        C = A * B;
        tr.test_eq(vec<float> {-1, -2, -3, 1, 2, 3}, C);
        rc::Big<float, 1> D({2}, 0.f);
        D += A * B;
        tr.test_eq(vec<float> {-6, 6}, D);
        rc::Big<float> Ad({2, 3}, {1, 2, 3, 1, 2, 3}), Bd({2}, {-1, 1}), Dd({2}, 0.f); // run time rank (case III)
        Dd += Ad * Bd;
        tr.test_eq(vec<float> {-6, 6}, Dd);
    }
    tr.section("index operands: test/frame-old.cc:28-58");
    {
        rc::Big<int, 2> c({3, 2}, 0);
        c = rc::_0 + rc::_1;
        tr.test_eq(vec<int> {0, 1, 1, 2, 2, 3}, c);
        c = rc::_0 * 2;
        tr.test_eq(vec<int> {0, 0, 2, 2, 4, 4}, c);
    }
    tr.section("periodic axes and outer product with insert: test/frame-new.cc:146-190");
    {
        rc::Big<int, 3> a({2, 3, 4}, (rc::_0+1)*100 + (rc::_1+1)*10 + (rc::_2+1));
        auto b = a(rc::all, rc::insert<1>, rc::iota(4, 0, 0));
        tr.test(4==b.rank() && 2==b.len(0) && rc::UNB==b.len(1) && 4==b.len(2) && 4==b.len(3), "shape of a(all, insert<1>, iota(4, 0, 0))");
        tr.test(rc::shape(a+b)==vec<rc::dim_t> {2, 3, 4, 4}, "rank 4 match");
        rc::Big<int, 4> c({2, 3, 4, 4}, 0);
        for (int j=0; j<3; ++j) c(rc::all, j) = a(rc::all, rc::iota(4, 0, 0));
        rc::Big<int, 4> e(a+c), f(a+b), g(b+a);
        tr.test_eq(e.to_host(), f);
        tr.test_eq(e.to_host(), g);
        rc::Big<int, 2> p({4, 3}, 10*rc::_1+100*rc::_0);
        rc::Big<int, 1> q({5}, rc::_0);
        rc::Big<int, 3> o(p(rc::dots<2>, rc::insert<1>) - q(rc::insert<2>, rc::dots<1>));
        vec<int> check;
        for (int i=0; i<4; ++i) for (int j=0; j<3; ++j) for (int k=0; k<5; ++k) check.push_back(10*j+100*i-k);
        tr.test_eq(check, o);
    }
    tr.section("reverse: test/view-ops.cc:19-52");
    {
        auto a = iota_filled<real, 3>({3, 2, 4}, 1);
        vec<real> check0 { 17, 18, 19, 20, 21, 22, 23, 24, 9, 10, 11, 12, 13, 14, 15, 16, 1, 2, 3, 4, 5, 6, 7, 8 };
        vec<real> check1 { 5, 6, 7, 8, 1, 2, 3, 4, 13, 14, 15, 16, 9, 10, 11, 12, 21, 22, 23, 24, 17, 18, 19, 20 };
        vec<real> check2 { 4, 3, 2, 1, 8, 7, 6, 5, 12, 11, 10, 9, 16, 15, 14, 13, 20, 19, 18, 17, 24, 23, 22, 21 };
        tr.test_eq(check0, rc::reverse(a.view(), 0));
        tr.test_eq(check0, a(rc::iota(rc::len, rc::len-1, -1)));
        tr.test_eq(check1, rc::reverse(a.view(), 1));
        tr.test_eq(check1, a(rc::dots<1>, rc::iota(rc::len, rc::len-1, -1)));
        tr.test_eq(check2, rc::reverse(a.view(), 2));
        tr.test_eq(check2, a(rc::dots<2>, rc::iota(rc::len, rc::len-1, -1)));
    }
    tr.section("transpose / diag: test/view-ops.cc:54-75,121-138");
    {
        auto a = iota_filled<real, 2>({3, 2}, 1);
        tr.test_eq(vec<real> {1, 3, 5, 2, 4, 6}, rc::transpose<1, 0>(a.view()));
        tr.test_eq(vec<real> {1, 3, 5, 2, 4, 6}, rc::transpose(a.view()));
        tr.test_eq(vec<real> {1, 3, 5, 2, 4, 6}, rc::transpose(a.view(), {1, 0}));
        auto s = iota_filled<real, 2>({3, 3}, 0);
        tr.test_eq(vec<real> {0, 4, 8}, rc::diag(s.view()));
    }
    tr.section("beaten subscripts, base pointer offsets: test/fromb.cc:46-81");
    {
        rc::Big<int, 2> a({10, 10}, rc::_1 + 10*rc::_0);
        auto off = [&](auto const & b) { return b.data()-a.data(); };
        { auto b = a(3, rc::all); tr.test(30==off(b)); tr.test_eq(vec<int> {30, 31, 32, 33, 34, 35, 36, 37, 38, 39}, b); }
        { auto b = a(rc::iota(4)); tr.test(0==off(b) && 4==b.len(0) && 10==b.len(1)); }
        { auto b = a(rc::iota(4, 4)); tr.test(40==off(b)); }
        { auto b = a(3, rc::iota(5, 4)); tr.test(34==off(b)); tr.test_eq(vec<int> {34, 35, 36, 37, 38}, b); }
        { auto b = a(rc::all, rc::iota(4, 2, 2)); tr.test(2==off(b) && 10==b.len(0) && 4==b.len(1)); }
        { auto b = a(rc::iota(3, 1, 2), rc::iota(2, 2, 3)); tr.test(12==off(b)); tr.test_eq(vec<int> {12, 15, 32, 35, 52, 55}, b); }
        { auto b = a(rc::iota(3, 9, -2), rc::iota(2, 2, 3)); tr.test(92==off(b)); tr.test_eq(vec<int> {92, 95, 72, 75, 52, 55}, b); }
        tr.test_eq(vec<int> {90, 91, 92, 93, 94, 95, 96, 97, 98, 99}, a(rc::len-1, rc::all)); // test/fromb.cc:131-162
        tr.test(int(a(2, 3))==23, "rank 0 subscript converts to a value");
    }
    tr.section("beatable multi-axis selectors, every form twice on one object: test/fromb.cc:166-193");
    {
        auto test = [&tr](auto const & a) {
            for (int rep=0; rep<2; ++rep) { // (the second pass on the same object is the regression: dots<> sizing kept state between calls)
                tr.test_eq(tr.host(a(0)), a(rc::dots<0>, 0), "a(dots<0>, 0)");
                tr.test_eq(tr.host(a(1)), a(rc::dots<0>, 1), "a(dots<0>, 1)");
                tr.test_eq(tr.host(a(rc::all, 0)), a(rc::dots<1>, 0), "a(dots<1>, 0)");
                tr.test_eq(tr.host(a(rc::all, 1)), a(rc::dots<1>, 1), "a(dots<1>, 1)");
                tr.test_eq(tr.host(a(rc::all, rc::all, 0)), a(rc::dots<2>, 0), "a(dots<2>, 0)");
                tr.test_eq(tr.host(a(rc::all, rc::all, 1)), a(rc::dots<2>, 1), "a(dots<2>, 1)");
                tr.test_eq(tr.host(a(rc::all, rc::all, 3)), a(rc::dots<2>, rc::len-1), "a(dots<2>, len-1)");
                tr.test_eq(tr.host(a(rc::all, rc::all, 1)), a(rc::dots<>, 1), "a(dots<>, 1)");
                tr.test_eq(tr.host(a(0, rc::all, rc::all)), a(0), "a(0)");
                tr.test_eq(tr.host(a(1, rc::all, rc::all)), a(1), "a(1)");
                tr.test_eq(tr.host(a(0, rc::all, rc::all)), a(0, rc::dots<2>), "a(0, dots<2>)");
                tr.test_eq(tr.host(a(1, rc::all, rc::all)), a(1, rc::dots<2>), "a(1, dots<2>)");
                tr.test_eq(tr.host(a(1, rc::all, rc::all)), a(rc::len-1, rc::dots<2>), "a(len-1, dots<2>)");
                tr.test_eq(tr.host(a(1, rc::all, rc::all)), a(1, rc::dots<>), "a(1, dots<>)");
                tr.test_eq(tr.host(a(0, rc::all, 1)), a(0, rc::dots<>, 1), "a(0, dots<>, 1)");
                tr.test_eq(tr.host(a(1, rc::all, 0)), a(1, rc::dots<>, 0), "a(1, dots<>, 0)");
            }
            tr.test_eq(vec<int> {100, 101, 102, 103, 110, 111, 112, 113, 120, 121, 122, 123}, a(1, rc::dots<>), "values of a(1, dots<>)");
            tr.test_eq(vec<int> {1, 11, 21}, a(0, rc::dots<>, 1), "values of a(0, dots<>, 1)");
        };
        test(rc::Big<int, 3>({2, 3, 4}, rc::_0*100 + rc::_1*10 + rc::_2)); // fixed rank
        test(rc::Big<int>({2, 3, 4}, rc::_0*100 + rc::_1*10 + rc::_2));    // var rank
    }
    tr.section("insert, and insert mixed with dots: test/fromb.cc:195-255");
    {
        auto test = [&tr](auto const & a) {
            tr.test_eq(tr.host(a(0)), a(rc::insert<0>, 0), "a(insert<0>, 0)");
            rc::Big<int, 4> a1({1, 2, 3, 4}, rc::_1*100 + rc::_2*10 + rc::_3);
            rc::Big<int, 4> a2({2, 1, 3, 4}, rc::_0*100 + rc::_2*10 + rc::_3);
            rc::Big<int, 5> a3({2, 1, 1, 3, 4}, rc::_0*100 + rc::_3*10 + rc::_4);
            // an inserted axis is unbounded: give it its length by matching against an array of that shape
            rc::Big<int, 4> z1({1, 2, 3, 4}, 0), z2({2, 1, 3, 4}, 0);
            rc::Big<int, 5> z3({2, 1, 1, 3, 4}, 0);
            tr.test_eq(a1.to_host(), z1 + a(rc::insert<1>), "a(insert<1>)");
            tr.test_eq(a2.to_host(), z2 + a(rc::all, rc::insert<1>), "a(all, insert<1>)");
            tr.test_eq(a3.to_host(), z3 + a(rc::all, rc::insert<2>), "a(all, insert<2>)");
            rc::Big<int, 3> z10({1, 3, 4}, 0);
            tr.test_eq(a1(rc::all, 0).to_host(), z10 + a(0, rc::insert<1>), "a(0, insert<1>)");
            tr.test_eq(a1(rc::all, 0).to_host(), z10 + a(rc::insert<1>, 0), "a(insert<1>, 0)");
            rc::Big<int, 4> aa1({2, 2, 3, 4}, a(rc::insert<1>));
            tr.test_eq(a.to_host(), aa1(0), "insert with undefined len 0");
            tr.test_eq(a.to_host(), aa1(1), "insert with undefined len 1");
            // mix insert + dots, twice on the same object
            for (int rep=0; rep<2; ++rep) {
                tr.test_eq(tr.host(a(rc::insert<0>, rc::dots<3>)), a(rc::insert<0>, rc::dots<>), "a(insert<0>, dots<>)");
                tr.test_eq(tr.host(a(rc::insert<0>, rc::all, rc::all, rc::all)), a(rc::insert<0>, rc::dots<>), "a(insert<0>, all, all, all)");
                tr.test_eq(tr.host(a(rc::insert<1>, rc::dots<3>) + rc::iota(2)), a(rc::insert<1>, rc::dots<>) + rc::iota(2), "a(insert<1>, dots<>)");
                tr.test_eq(tr.host(a(rc::insert<1>) + rc::iota(2)), a(rc::insert<1>, rc::dots<>) + rc::iota(2), "a(insert<1>) vs a(insert<1>, dots<>)");
                tr.test_eq(tr.host(a(rc::dots<3>, rc::insert<1>) + rc::Big<int, 4>({2, 3, 4, 2}, 0)), a(rc::dots<>, rc::insert<1>) + rc::Big<int, 4>({2, 3, 4, 2}, 0), "a(dots<>, insert<1>)");
            }
        };
        test(rc::Big<int, 3>({2, 3, 4}, rc::_0*100 + rc::_1*10 + rc::_2));
        test(rc::Big<int>({2, 3, 4}, rc::_0*100 + rc::_1*10 + rc::_2));
    }
    tr.section("bounds and agreement are checked before launch: ra/ply.hh:395,407; test/checks.cc:159-233");
    {
        rc::Big<int, 2> a({10, 10}, 0);
        tr.test_throws([&]{ a(10, rc::all); }, RACU_ERR_BOUNDS);
        tr.test_throws([&]{ a(rc::iota(4, 8)); }, RACU_ERR_BOUNDS);
        rc::Big<int, 1> b({3}, 0);
        tr.test_throws([&]{ a += b; }, RACU_ERR_SHAPE);
        rc::Big<real, 1> c({3}, 0.);
        tr.test_throws([&]{ c = rc::_1; }, RACU_ERR_SHAPE, "unbounded expression");
    }
    tr.section("iota: test/iota.cc:21-67");
    {
        rc::Big<int, 1> a({4}, 0);
        a = rc::iota(4, 3);
        tr.test_eq(vec<int> {3, 4, 5, 6}, a);
        a = rc::iota(4, 3, 2);
        tr.test_eq(vec<int> {3, 5, 7, 9}, a);
        rc::Big<int, 3> b({3, 2, 4}, 0);
        rc::transpose<0, 2, 1>(b.view()) = rc::iota(3, 1);
        vec<int> check;
        for (int i=0; i<3; ++i) for (int j=0; j<8; ++j) check.push_back(i+1);
        tr.test_eq(check, b);
    }
    tr.section("where / pick: test/where.cc:22-38,98-122");
    {
        rc::Big<real, 1> a0 { 1, 2, 3 }, a1 { 10, 20, 30 }, a({3}, 0.);
        rc::Big<int, 1> p { 0, 1, 0 };
        a = rc::pick(p, a0, a1);
        tr.test_eq(vec<real> {1, 20, 3}, a);
        a = rc::where(p, a0, a1);
        tr.test_eq(vec<real> {10, 2, 30}, a);
        rc::Big<bool, 1> w({4}, {true, false, false, true});
        tr.test_eq(vec<int> {1, 2, 2, 1}, rc::where(w, 1, 2));
        rc::Big<int, 1> v { 1, 2, 3, 4 };
        tr.test_eq(vec<int> {17, 2, 3, 17}, rc::where(rc::_0>0 && rc::_0<3, v, 17));
        rc::Big<int, 1> s {-1, 0, 1};
        tr.test_eq(vec<int> {44, 77, 99}, rc::where(s>=0, rc::where(s<1, 77, 99), 44));
    }
    tr.section("where as lvalue: test/where.cc:148-157");
    {
        rc::Big<int, 1> a { 0, 0, 0, 0 }, b { 0, 0, 0, 0 };
        rc::where(rc::_0>0 && rc::_0<3, a, b) = 7;
        tr.test_eq(vec<int> {0, 7, 7, 0}, a);
        tr.test_eq(vec<int> {7, 0, 0, 7}, b);
        rc::where(rc::_0<=0 || rc::_0>=3, a, b) += 2;
        tr.test_eq(vec<int> {2, 7, 7, 2}, a);
        tr.test_eq(vec<int> {7, 2, 2, 7}, b);
        rc::Big<real, 1> x { 1, 2, 3 }, y { 10, 20, 30 };                    // pick with an array selector and an array right hand side
        rc::Big<int, 1> p { 0, 1, 0 };
        rc::pick(p, x, y) *= rc::Big<real, 1> { 2, 3, 4 };
        tr.test_eq(vec<real> {2, 2, 12}, x);
        tr.test_eq(vec<real> {10, 60, 30}, y);
    }
    tr.section("named functions and <=>: ra/ra.hh:190,210,221-222; test/operators.cc:58,65-68; test/test.cc:86-88; test/reexported.cc:25-26");
    {
        tr.test_eq(vec<bool> {true, false, true, true}, rc::odd(rc::Big<int, 1> {1, 2, 3, -1}));
        rc::Big<int, 1> i3 {3, 4, 5};
        rc::Big<real, 1> d3 {2., 4., 6.};
        auto c3 = TestRecorder::host(i3 <=> d3);
        std::string sign;
        for (int z: c3) sign += z>0 ? '+' : z<0 ? '-' : '0';
        tr.test(sign=="+0-", "<=> maps to +0-");
        tr.test_eq(vec<bool> {true, false, false}, (i3 <=> d3) > 0);
        real const inf = std::numeric_limits<real>::infinity(), nan = std::numeric_limits<real>::quiet_NaN();
        rc::Big<real, 1> x {3., inf, nan, -inf};
        tr.test_eq(vec<bool> {true, false, false, false}, rc::isfinite(x));
        tr.test_eq(vec<bool> {false, false, true, false}, rc::isnan(x));
        tr.test_eq(vec<bool> {false, true, false, true}, rc::isinf(x));
        tr.test_eq(vec<real> {2.5, 4., 1.}, rc::lerp(rc::Big<real, 1> {4., 4., 4.}, 1., rc::Big<real, 1> {0.5, 0., 1.}));
        tr.test_eq(vec<int> {-1, 0, 3, 4}, rc::clamp(rc::Big<int, 1> {-5, 0, 3, 9}, -1, 4));
        tr.test_eq(vec<real> {0., 0., 2./3.}, rc::rel_error(rc::Big<real, 1> {1., 0., 2.}, rc::Big<real, 1> {1., 0., 1.}));
        auto close = [&](auto const & e, vec<real> const & want, char const * what) {
            auto got = TestRecorder::host(e);
            bool ok = got.size()==want.size();
            for (size_t k=0; ok && k<got.size(); ++k) ok = std::abs(got[k]-want[k])<=1e-14*std::max(1., std::abs(want[k]));
            tr.test(ok, what);
        };
        rc::Big<real, 1> v {-0.75, 0., 0.3};
        close(rc::tan(v), {std::tan(-0.75), 0., std::tan(0.3)}, "tan");
        close(rc::sinh(v), {std::sinh(-0.75), 0., std::sinh(0.3)}, "sinh");
        close(rc::cosh(v), {std::cosh(-0.75), 1., std::cosh(0.3)}, "cosh");
        close(rc::tanh(v), {std::tanh(-0.75), 0., std::tanh(0.3)}, "tanh");
        close(rc::asin(v), {std::asin(-0.75), 0., std::asin(0.3)}, "asin");
        close(rc::acos(v), {std::acos(-0.75), std::acos(0.), std::acos(0.3)}, "acos");
        close(rc::atan(v), {std::atan(-0.75), 0., std::atan(0.3)}, "atan");
        close(rc::expm1(v), {std::expm1(-0.75), 0., std::expm1(0.3)}, "expm1");
        close(rc::log1p(v), {std::log1p(-0.75), 0., std::log1p(0.3)}, "log1p");
        close(rc::log10(v+2.), {std::log10(1.25), std::log10(2.), std::log10(2.3)}, "log10");
    }
    tr.section("reductions: test/reduction.cc:85-138,199-201");
    {
        rc::Big<real, 1> a {1, 2}, b {3, 4};
        tr.test(11.==rc::dot(a, b));
        tr.test(8.==rc::reduce_sqrm(a-b));
        tr.test(std::sqrt(8.)==rc::norm2(a-b));
        rc::Big<real, 1> c {4, 6, -3, 2};
        tr.test(9.==rc::sum(c) && -144.==rc::prod(c) && 6.==rc::amax(c) && -3.==rc::amin(c));
        rc::Big<real, 1> q({3}, std::numeric_limits<real>::quiet_NaN());
        tr.test(-std::numeric_limits<real>::infinity()==rc::amax(q) && std::numeric_limits<real>::infinity()==rc::amin(q), "amax/amin ignore NaN");
        rc::Big<real, 2> e({0, 3}, 0.);
        tr.test(0.==rc::sum(e) && 1.==rc::prod(e), "empty");
    }
    tr.section("sum columns / rows by rank extension: test/reduction.cc:141-222");
    {
        rc::Big<real, 2> A({100, 111}, rc::_0 - rc::_1);
        vec<real> B(100, 0.), Br(111, 0.);
        for (int i=0; i<100; ++i) for (int j=0; j<111; ++j) { B[i] += i-j; Br[j] += i-j; }
        rc::Big<real, 1> C({100}, 0.);
        C += A;
        tr.test_eq(B, C);
        C = 0.;
        C = C + A;
        tr.test_eq(B, C, "C = C + A is a running sum");
        rc::Big<real, 1> D({111}, 0.);
        D += rc::transpose(A.view());
        tr.test_eq(Br, D);
        D = 0.;
        D(rc::insert<1>) += A;
        tr.test_eq(Br, D);
    }
    tr.section("sum rows, seven spellings: test/reduction.cc:182-222, bench/bench-sum-rows.cc:62-81");
    {
        rc::Big<real, 2> A({100, 111}, rc::_0 - rc::_1);
        vec<real> B(111);
        for (int j=0; j<111; ++j) { real s = 0; for (int i=0; i<100; ++i) s += i-j; B[j] = s; }
        std::string kernels;
        auto check = [&](auto const & C, char const * what) { tr.test_eq(B, C, what); kernels += std::string(racu_last_kernel()) + " "; };
        tr.test_eq(B, rc::sum(rc::iter<1>(A)), "sum(iter<1>(A))");
        { rc::Big<real, 1> C({111}, 0.); rc::iter<1>(C) += rc::iter<1>(A); check(C, "iter<1>(C) += iter<1>(A)"); }
        { rc::Big<real, 1> C({111}, 0.); rc::scalar(C) += rc::iter<1>(A); check(C, "scalar(C) += iter<1>(A)"); }
        { rc::Big<real, 1> C({111}, 0.); rc::for_each(rc::wrank<1, 1>(rc::add_assign), C, A); check(C, "rank conjunction"); }
        { rc::Big<real, 1> C({111}, 0.); rc::for_each(rc::wrank<1, 1>(rc::wrank<0, 0>(rc::add_assign)), C, A); check(C, "double rank conjunction"); }
        { rc::Big<real, 1> C({111}, 0.); C += rc::transpose(A.view()); check(C, "C += transpose(A)"); }
        { rc::Big<real, 1> C({111}, 0.); C(rc::insert<1>) += A; check(C, "C(insert<1>) += A"); }
        rc::Big<real, 2> E({0, 111}, 0.);
        tr.test_eq(vec<real>(111, 0.), rc::sum(rc::iter<1>(E)), "no cells: sum = 0");
        tr.test_eq(vec<real>(111, 1.), rc::prod(rc::iter<1>(E)), "no cells: prod = 1");
        std::printf("    kernels: %s\n", kernels.c_str());
        // every spelling is the same descriptor: the same kernel family ran each time
        bool same = true;
        { size_t p0 = kernels.find(' '); std::string first = kernels.substr(0, p0); for (size_t q=0; q<kernels.size();) { size_t e = kernels.find(' ', q); same = same && kernels.substr(q, e-q)==first; q = e+1; } }
        tr.test(same, "all spellings dispatch to one kernel");
        // gemm1 of bench/bench-gemm.cc:25-30 through a double rank conjunction: the contraction descriptor
        rc::Big<real, 2> a({3, 2}, 2.), b({2, 4}, 3.), c({3, 4}, 0.);
        rc::for_each(rc::wrank<1, 2, 1>(rc::wrank<0, 1, 1>(rc::maybe_fma)), a, b, c);
        tr.test_eq(vec<real>(12, 12.), c, "gemm1: wrank<1, 2, 1>(wrank<0, 1, 1>(maybe_fma))");
    }
    tr.section("gemv / gevm / gemm: test/reduction.cc:241-333, bench/bench-gemm.cc:229-237");
    {
        rc::Big<real, 2> a({3, 4}, rc::_0 - rc::_1);
        rc::Big<real, 1> b {1, 2, 3, 4};
        tr.test_eq(vec<real> {-20, -10, 0}, rc::gemv(a, b));
        rc::Big<real, 2> at({4, 3}, rc::_1 - rc::_0);
        tr.test_eq(vec<real> {-20, -10, 0}, rc::gevm(b, at));
        rc::Big<real, 2> x({3, 2}, 2.), y({2, 4}, 3.);
        tr.test_eq(vec<real>(12, 12.), rc::gemm(x, y));
        rc::Big<real, 2> c({3, 4}, 1.);
        rc::gemm(x, y, c);
        tr.test_eq(vec<real>(12, 13.), c);
        int const M = 13, K = 17, N = 11;
        rc::Big<real, 2> g({M, K}, rc::_0 - rc::_1), h({K, N}, rc::_1 - 2*rc::_0);
        vec<real> check(M*N, 0.);
        for (int i=0; i<M; ++i) for (int k=0; k<K; ++k) for (int j=0; j<N; ++j) check[i*N+j] += real(i-k)*real(j-2*k);
        tr.test_eq(check, rc::gemm(g, h));
    }
    tr.section("3-D 7-point stencil as slices: bench/bench-stencil3.cc:55-64");
    {
        int const nx = 7, ny = 6, nz = 9;
        rc::Big<float, 3> A({nx, ny, nz}, rc::cast<float>((rc::_0*7 + rc::_1*3 + rc::_2*5) % 11) - 5.f), Anext({nx, ny, nz}, 0.f);
        vec<float> ref = A.to_host(), refn(ref.size(), 0.f);
        auto I = rc::iota(nx-2, 1), J = rc::iota(ny-2, 1), K = rc::iota(nz-2, 1);
        auto at = [&](vec<float> const & v, int i, int j, int k) { return v[(i*ny+j)*nz+k]; };
        for (int t=0; t<2; ++t) {
            Anext(I, J, K) = -6*A(I, J, K) + A(I+1, J, K) + A(I, J+1, K) + A(I, J, K+1) + A(I-1, J, K) + A(I, J-1, K) + A(I, J, K-1);
            std::swap(A.cp, Anext.cp);
            for (int i=1; i+1<nx; ++i) for (int j=1; j+1<ny; ++j) for (int k=1; k+1<nz; ++k) {
                refn[(i*ny+j)*nz+k] = -6*at(ref, i, j, k) + at(ref, i+1, j, k) + at(ref, i, j+1, k) + at(ref, i, j, k+1) + at(ref, i-1, j, k) + at(ref, i, j-1, k) + at(ref, i, j, k-1);
            }
            std::swap(ref, refn);
        }
        tr.test_eq(ref, A);
        tr.test_eq(refn, Anext);
    }
    tr.section("the benchmark's time loop as one call (stencil_steps): bench/bench-stencil3.cc:55-64 with its swap :117-121, T = 7 and 10 on 20 x 18 x 136");
    {
        int const nx = 20, ny = 18, nz = 136;
        for (int T: {7, 10}) {
            rc::Big<float, 3> A({nx, ny, nz}, (rc::cast<float>((rc::_0*7 + rc::_1*3 + rc::_2*5) % 11) - 5.f) * 0.125f), Anext({nx, ny, nz}, 0.25f);
            rc::Big<float, 3> B({nx, ny, nz}, (rc::cast<float>((rc::_0*7 + rc::_1*3 + rc::_2*5) % 11) - 5.f) * 0.125f), Bnext({nx, ny, nz}, 0.25f);
            auto I = rc::iota(nx-2, 1), J = rc::iota(ny-2, 1), K = rc::iota(nz-2, 1);
            for (int t=0; t<T; ++t) {
                Anext(I, J, K) = -6*A(I, J, K) + A(I+1, J, K) + A(I, J+1, K) + A(I, J, K+1) + A(I-1, J, K) + A(I, J-1, K) + A(I, J, K-1);
                std::swap(A.cp, Anext.cp);
            }
            if (rc::stencil_steps(B, Bnext, T)) std::swap(B.cp, Bnext.cp);
            tr.test_eq(A.to_host(), B);          // state T, bit for bit (different boundary values in the two buffers, like the benchmark's value / 0)
            tr.test_eq(Anext.to_host(), Bnext);  // state T-1
        }
    }
    tr.section("any / every / index / lexical_compare: test/early.cc:17-82, ra/ra.hh:279-304");
    {
        rc::Big<int, 0> a0({}, 99);
        tr.test(rc::any(a0==99)); tr.test(rc::every(a0==99));
        rc::Big<int, 2> a({2, 3}, rc::_0*2 + rc::_1*10);
        tr.test(!rc::any(a%2!=0)); tr.test(rc::every(a%2==0));
        rc::Big<int, 2> b({2, 3}, 1);
        b(1, 1) = 0;
        tr.test(!rc::every(1==b)); tr.test(rc::any(0==b));
        auto bt = rc::transpose<1, 0>(b);
        tr.test(!rc::every(1==bt)); tr.test(rc::any(0==bt));
        rc::Big<int, 1> d {1, 3, -1, 9};
        tr.test(2==rc::index(d<0)); tr.test(-1==rc::index(d>10));
        rc::Big<int, 2> e({2, 2}, 0);
        e(0, 0) = 1; e(0, 1) = 3; e(1, 0) = -1; e(1, 1) = 9;
        tr.test(1==rc::index(e<0), "index with rank 2 returns the first axis-0 index"); tr.test(-1==rc::index(e>10));
        rc::Big<int, 1> z({0}, 0);
        tr.test(!rc::any(z==0)); tr.test(rc::every(z==1)); tr.test(-1==rc::index(z==0));
        rc::Big<int, 2> x({2, 3}, rc::_0*3 + rc::_1), y({2, 3}, rc::_0*3 + rc::_1);
        tr.test(!rc::lexical_compare(x, y));
        y(1, 1) = 7; y(1, 2) = 0;
        tr.test(rc::lexical_compare(x, y)); tr.test(!rc::lexical_compare(y, x));
    }
    tr.section("unbeaten subscripts a(i ...): test/fromu.cc:46-67,70-83,107-121");
    {
        rc::Big<double, 1> v {7, 9, 3, 4};
        std::vector<int> i3201 {3, 2, 0, 1};
        tr.test_eq(vec<double> {4, 3, 7, 9}, v(i3201));
        rc::Big<int, 2> ii({2, 2}, 0);
        ii(0, 0) = 3; ii(0, 1) = 2; ii(1, 0) = 0; ii(1, 1) = 1;
        tr.test_eq(vec<double> {4, 3, 7, 9}, v(ii));                                   // rank-2 index array: result 2x2
        rc::Big<double, 2> m({2, 2}, rc::_0*2 + rc::_1 + 1);                          // {{1, 2}, {3, 4}}
        std::vector<int> i01 {0, 1}, i10 {1, 0};
        tr.test_eq(vec<double> {1, 2, 3, 4}, m(i01, i01));
        tr.test_eq(vec<double> {2, 1, 4, 3}, m(i01, i10));
        tr.test_eq(vec<double> {3, 4, 1, 2}, m(i10, i01));
        tr.test_eq(vec<double> {4, 3, 2, 1}, m(i10, i10));
        tr.test_eq(vec<double> {4, 2}, m(i10, 1));                                     // mixed unbeaten / scalar
        tr.test_eq(vec<double> {4, 3}, m(1, i10));
        rc::Big<int, 2> b({4, 3}, rc::_0*10 + rc::_1);
        rc::Big<int, 1> rows {3, 0, 3};
        tr.test_eq(vec<int> {30, 31, 32, 0, 1, 2, 30, 31, 32}, b(rows));               // other axis kept whole
        std::vector<int> cols {2, 0};
        tr.test_eq(vec<int> {2, 0, 12, 10, 22, 20, 32, 30}, b(rc::all, cols));
        tr.test_eq(vec<int> {12*2+1, 10*2+1, 22*2+1, 20*2+1}, b(rc::iota(2, 1), cols)*2 + 1);
        rc::Big<int, 2> ij({3, 2}, 0);                                                 // at(a, i): index tuples (ra/ra.hh:228-239)
        ij(0, 1) = 1; ij(1, 0) = 3; ij(1, 1) = 2; ij(2, 0) = 1;
        tr.test_eq(vec<int> {1, 32, 10}, rc::at(b, ij));
        tr.test_eq(vec<int> {9, 8, 7, 6}, rc::Big<int, 1>({10}, rc::_0)(9 - rc::Big<int, 1> {0, 1, 2, 3})); // index EXPRESSION (test/fromu.cc:26)
    }
    tr.section("a(i ...) with index arrays as an lvalue: test/fromu.cc:28-31,53-61,84-92");
    {
        rc::Big<double, 1> a({4}, 0.);
        std::vector<int> i3201 {3, 2, 0, 1};
        a(i3201) = std::vector<double> {9, 7, 1, 4};
        tr.test_eq(vec<double> {1, 4, 7, 9}, a);
        a(i3201) = 77.;
        tr.test_eq(vec<double> {77, 77, 77, 77}, a);
        rc::Big<double, 2> m({2, 2}, 0.), x({2, 2}, 0.);
        x(0, 0) = 9; x(0, 1) = 7; x(1, 0) = 1; x(1, 1) = 4;
        std::vector<int> i10 {1, 0};
        m(i10, i10) = x;
        tr.test_eq(vec<double> {4, 1, 7, 9}, m);
        m(i10, i10) *= x;
        tr.test_eq(vec<double> {16, 1, 49, 81}, m);
        m(i10, i10) = std::vector<double> {9, 7};                                      // prefix matching of the right hand side
        tr.test_eq(vec<double> {7, 7, 9, 9}, m);
        rc::Big<int, 1> idx {0, 3, 3, 1, 3, 0, 2, 3}, h({5}, 0);
        h(idx) += 1;                                                                   // repeated indices accumulate
        tr.test_eq(vec<int> {2, 1, 1, 4, 0}, h);
        rc::Big<double, 1> src {10, 20, 30, 40}, out({6}, 0.);
        out(std::vector<int> {5, 0, 2}) = src(std::vector<int> {3, 3, 1})*2 + 1;
        tr.test_eq(vec<double> {81, 0, 41, 0, 0, 81}, out);
    }
    tr.section("stencil() view, f_sumprod form: ra/arrays.hh:615-630, bench/bench-stencil2.cc");
    {
        rc::Big<double, 2> B({6, 8}, rc::cast<double>((rc::_0*5 + rc::_1*3) % 7)), Bn({6, 8}, 0.), Br({6, 8}, 0.);
        rc::Big<double, 2> mask({3, 3}, 0.);
        mask(1, 1) = -4; mask(0, 1) = 1; mask(2, 1) = 1; mask(1, 0) = 1; mask(1, 2) = 1;
        auto I = rc::iota(4, 1), J = rc::iota(6, 1);
        Br(I, J) = -4*B(I, J) + B(I+1, J) + B(I, J+1) + B(I-1, J) + B(I, J-1);           // bench-stencil2.cc:53-55
        auto S = rc::stencil(B, 1, 1);
        tr.test(4==S.rank() && 4==S.len(0) && 6==S.len(1) && 3==S.len(2) && 3==S.len(3));
        Bn(I, J) += S*mask(rc::insert<2>);
        tr.test_eq(Br.to_host(), Bn);
    }
    tr.section("reference reductions refmin / refmax: test/reduction.cc:404-424");
    {
        rc::Big<double, 2> A({2, 3}, rc::_1 - rc::_0), B({2, 3}, rc::_1 - rc::_0);
        auto mn = rc::refmin(A);
        tr.test(-1==double(mn));
        mn = -99.;
        B(1, 0) = -99.;
        tr.test_eq(B.to_host(), A);
        auto mx = rc::refmax(A);
        tr.test(2==double(mx));
        mx = 0.;
        B(0, 2) = 0.;
        tr.test_eq(B.to_host(), A);
        auto my = rc::refmax(A);                                                      // the first of the two 1s: A(0, 1)
        tr.test(1==double(my));
        my = 77.;
        B(0, 1) = 77.;
        tr.test_eq(B.to_host(), A);
    }
    tr.section("prepare: plan + specialise without launching (racu_prepare)");
    {
        rc::Big<float, 2> a({8, 16}, 1.f), c({8, 16}, 2.f);
        tr.test(!rc::prepare(rc::where(a<c, a*c, rc::sqrt(c))).empty());
        tr.test_eq(vec<float>(128, 1.f), a);                                           // nothing ran
    }
    tr.section("cghs with device resident scalars: examples/laplace2d.cc:60-91, examples/cghs.hh:15-45");
    {
        int const N = 50;
        double const h = 1./N, PI = 3.14159265358979323846;
        std::vector<double> gv((N+1)*(N+1), 0.), fv((N+1)*(N+1), 0.);
        for (int i=1; i<N; ++i) for (int j=1; j<N; ++j) { gv[i*(N+1)+j] = std::sin(2*PI*i*h)*std::sin(2*PI*j*h); fv[i*(N+1)+j] = 8*PI*PI*gv[i*(N+1)+j]; }
        rc::Big<double, 2> v({N+1, N+1}, 0.), w({N+1, N+1}, 0.), u({N+1, N+1}, 0.), b({N+1, N+1}, 0.);
        rc::ck_status(racu_memcpy_h2d(v.data(), gv.data(), gv.size()*8, rc::stream()));
        rc::ck_status(racu_memcpy_h2d(w.data(), fv.data(), fv.size()*8, rc::stream()));
        auto i = rc::iota(N-1, 1), j = rc::iota(N-1, 1);
        auto mult_stiff = [&](auto const & vv, auto const & ww) { ww(i, j) = 4*vv(i, j) - vv(i-1, j) - vv(i, j-1) - vv(i+1, j) - vv(i, j+1); };
        b(i, j) = (h*h)*(w(i, j)/2. + (w(i-1, j) + w(i, j-1) + w(i+1, j) + w(i, j+1) + w(i+1, j+1) + w(i-1, j-1))/12.);
        rc::Big<double, 3> work({3, N+1, N+1}, 0.);
        int const its = rc::cghs(mult_stiff, b, u, work, 1e-5, 200);
        double const mx = rc::amax(rc::abs(u-v));
        tr.test(mx<=0.00463, "max error " + std::to_string(mx));
        tr.test(its<=67, "iterations " + std::to_string(its));
    }
    return tr.summary();
}
