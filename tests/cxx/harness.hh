// harness.hh - a minimal TestRecorder in the spirit of ra/test.hh:33-234 (test, test_eq, summary = number of failures).
#pragma once
#include <cmath>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "ra/cuda.hh"

struct TestRecorder
{
    int total = 0, failed = 0;
    std::string sect;
    void section(std::string const & s) { sect = s; std::printf("--- %s\n", s.c_str()); }
    void test(bool c, std::string const & what = "")
    {
        ++total;
        if (!c) { ++failed; std::printf("FAIL [%s] %s\n", sect.c_str(), what.c_str()); }
    }
    // row-major values of a device view / expression against host values, exact
    template <class T, class X> void test_eq(std::vector<T> const & expected, X const & x, std::string const & what = "")
    {
        auto got = host(x);
        bool ok = got.size()==expected.size();
        for (size_t i=0; ok && i<got.size(); ++i) ok = (got[i]==static_cast<typename decltype(got)::value_type>(expected[i]));
        if (!ok) {
            std::printf("  expected:"); for (auto v: expected) std::cout << " " << v;
            std::printf("\n  got:     "); for (auto v: got) std::cout << " " << v;
            std::printf("\n");
        }
        test(ok, what);
    }
    template <class X> static auto host(X const & x)
    {
        if constexpr (requires { x.to_host(); }) return x.to_host();
        else return ra::cuda::concrete(x).to_host();
    }
    template <class F> void test_throws(F && f, int status, std::string const & what = "")
    {
        bool ok = false;
        try { f(); } catch (ra::cuda::error const & e) { ok = e.status==status; if (!ok) std::printf("  status %d: %s\n", e.status, e.what()); }
        test(ok, what);
    }
    int summary() const { std::printf("%d tests, %d failed\n", total, failed); return failed; }
};
