// bench_header.cc - the C++ host header (include/ra/cuda.hh) timed end to end and per call, linked against ra_ra_b200/libracu.so.
// Run by bench.py (one JSON object on stdout); built by __graft_entry__.build() so that it travels to the GPU box.
//   (1) per-call overhead of ra::cuda::sum (synchronous, result on the host) against the same reduction issued asynchronously
//       back to back (reduce_into: the kernel-bound rate), 2^20 ... 2^28 floats;
//   (2) e2e: the reduce workload of bench.py (sum, reduce_sqrm, dot, sum-rows, sum-cols: ra/ra.hh:348-401, bench-sum-rows/cols.cc)
//       from HOST std::vector inputs: upload, five traversals, results read back, all inside the timed region.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ra/cuda.hh"

namespace rc = ra::cuda;
using clk = std::chrono::steady_clock;
static double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now()-t0).count(); }

int main(int argc, char ** argv)
{
    int const log2n = argc>1 ? std::atoi(argv[1]) : 28;
    int const e2e_steps = argc>2 ? std::atoi(argv[2]) : 5;
    std::printf("{\"per_call\": [");
    bool first = true;
    for (int l2 = 20; l2<=28; l2 += 2) {
        rc::dim_t const n = rc::dim_t(1)<<l2;
        rc::Big<float, 1> a({n}, 1.f);
        rc::Big<float, 1> out({4}, 0.f);
        auto o0 = out(0);
        float s = 0;
        for (int w=0; w<5; ++w) s = rc::sum(a);
        int const reps = l2<=24 ? 200 : 40;
        auto t0 = clk::now();
        for (int r=0; r<reps; ++r) s = rc::sum(a);
        double const sync_us = 1e3*ms_since(t0)/reps;
        for (int w=0; w<5; ++w) rc::reduce_into(o0, a.view());
        rc::sync();
        t0 = clk::now();
        for (int r=0; r<reps; ++r) rc::reduce_into(o0, a.view());
        rc::sync();
        double const async_us = 1e3*ms_since(t0)/reps;
        std::printf("%s{\"log2n\": %d, \"sum_sync_us\": %.2f, \"reduce_into_async_us\": %.2f, \"overhead_us\": %.2f, \"GB/s_sync\": %.1f, \"ok\": %s}",
                    first ? "" : ", ", l2, sync_us, async_us, sync_us-async_us, 4.0*double(n)/sync_us/1e3, s==float(n) || l2>24 ? "true" : "false");
        first = false;
    }
    std::printf("], ");
    {
        rc::dim_t const N = rc::dim_t(1)<<log2n, m = rc::dim_t(1)<<(log2n/2), n = N/m;
        std::vector<float> ha(N), hb(N);
        uint32_t x = 0x5EED0011u;
        for (rc::dim_t i=0; i<N; ++i) { x = x*1664525u+1013904223u; ha[i] = float(x>>8)*(1.f/16777216.f); x = x*1664525u+1013904223u; hb[i] = float(x>>8)*(1.f/16777216.f); }
        rc::Big<float, 1> a({N}, rc::none), b({N}, rc::none);
        rc::Big<float, 1> rows({n}, 0.f), cols({m}, 0.f);
        std::vector<float> hrows, hcols;
        float s0 = 0, s1 = 0, s2 = 0;
        auto step = [&] {
            a.from_host(ha.data(), N);
            b.from_host(hb.data(), N);
            rc::View<float, 2> A2(std::array<rc::Dim, 2> {rc::Dim {m, n}, rc::Dim {n, 1}}, a.data());
            s0 = rc::sum(a);
            s1 = rc::reduce_sqrm(a);
            s2 = rc::dot(a, b);
            rows = 0.f; cols = 0.f;
            rows(rc::insert<1>) += A2;   // c[j] = sum_i a[i, j]   bench-sum-rows.cc
            cols += A2;                  // c[i] = sum_j a[i, j]   bench-sum-cols.cc
            hrows = rows.to_host();
            hcols = cols.to_host();
        };
        step();
        auto t0 = clk::now();
        for (int r=0; r<e2e_steps; ++r) step();
        double const ms = ms_since(t0)/e2e_steps;
        double const bytes = 4.0*double(N)*6 + 4.0*double(n+m);
        double ref = 0; for (rc::dim_t i=0; i<N; ++i) ref += ha[i];
        std::printf("\"e2e\": {\"value\": %.2f, \"unit\": \"GB/s\", \"ms_per_step\": %.3f, \"steps\": %d, \"log2n\": %d, \"h2d_bytes_per_step\": %.0f, \"d2h_bytes_per_step\": %.0f, "
                    "\"sum_rel_err\": %.3g, \"inputs\": \"pageable std::vector<float> (the header's from_host)\"}}\n",
                    bytes/ms/1e6, ms, e2e_steps, log2n, 8.0*double(N), 4.0*double(n+m)+12, std::abs(double(s0)-ref)/ref);
        (void)s1; (void)s2;
    }
    return 0;
}
