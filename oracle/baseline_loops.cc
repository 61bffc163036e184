// baseline_loops.cc - the reference's hot loops for the benchmark workloads, written the way ra::ply_ravel compiles
// them (one pointer per leaf bumped by its inner step, flat inner loop, odometer outside: ra/ply.hh:58-84), single
// threaded like the reference (README.md:81). TEST / BENCH INFRASTRUCTURE ONLY: timed by bench.py as the CPU baseline
// ("port": the reference needs gcc >= 14 and cannot be built in this image). Never linked into the product.
// Build: g++ -O3 (the reference's benchmarks are built -O3; docs/ra-ra.texi:2216).
#include <cstdint>

extern "C" {

// ra::sum(a): ra/ra.hh:348-354. float accumulator, sequential.
float base_sum_f32(float const * a, int64_t n) { float c = 0; for (int64_t s=n; --s>=0; ++a) { c += *a; } return c; }
// ra::reduce_sqrm(a): ra/ra.hh:395-401.
float base_sqrm_f32(float const * a, int64_t n) { float c = 0; for (int64_t s=n; --s>=0; ++a) { c += (*a)*(*a); } return c; }
// ra::dot(a, b): ra/ra.hh:379-385.
float base_dot_f32(float const * a, float const * b, int64_t n) { float c = 0; for (int64_t s=n; --s>=0; ++a, ++b) { c += (*a)*(*b); } return c; }

// bench/bench-sum-cols.cc:62-65  c += a  : c[i] += a[i,j], j inner, c has step 0 on axis 1.
void base_sum_cols_f32(float * c, float const * a, int64_t m, int64_t n)
{
    for (int64_t i=0; i<m; ++i, ++c) { for (int64_t s=n; --s>=0; ++a) { *c += *a; } }
}
// bench/bench-sum-rows.cc:74-77  iter<1>(c) += iter<1>(a) : for each row i, c[:] += a[i,:].
void base_sum_rows_f32(float * c, float const * a, int64_t m, int64_t n)
{
    for (int64_t i=0; i<m; ++i) { float * cj = c; for (int64_t s=n; --s>=0; ++a, ++cj) { *cj += *a; } }
}

// examples/read-me.cc:32  B += A*C + v, v a std::vector<double> matched on axis 0 (inner step 0).
void base_map_f32(float * B, float const * A, float const * C, double const * v, int64_t m, int64_t n)
{
    for (int64_t i=0; i<m; ++i, ++v) { for (int64_t s=n; --s>=0; ++B, ++A, ++C) { *B += (*A)*(*C) + *v; } }
}
void base_map_f32v(float * B, float const * A, float const * C, float const * v, int64_t m, int64_t n)
{
    for (int64_t i=0; i<m; ++i, ++v) { for (int64_t s=n; --s>=0; ++B, ++A, ++C) { *B += (*A)*(*C) + *v; } }
}

// bench/bench-stencil3.cc:55-64 f_slices: 8 leaves with identical dims; innermost axis not collapsible -> 2-level odometer.
void base_stencil3_f32(float * An, float const * A, int64_t n0, int64_t n1, int64_t n2)
{
    int64_t s0 = n1*n2, s1 = n2;
    for (int64_t i=1; i+1<n0; ++i) {
        for (int64_t j=1; j+1<n1; ++j) {
            float * y = An + i*s0 + j*s1 + 1;
            float const * c = A + i*s0 + j*s1 + 1;
            float const * ip = c+s0, * jp = c+s1, * kp = c+1, * im = c-s0, * jm = c-s1, * km = c-1;
            for (int64_t s=n2-2; --s>=0; ++y, ++c, ++ip, ++jp, ++kp, ++im, ++jm, ++km) {
                *y = -6*(*c) + *ip + *jp + *kp + *im + *jm + *km;
            }
        }
    }
}

// B = transpose(A, {2,1,0}) on rank 3 (ra/arrays.hh:431-461): destination order, source read with a large inner step.
void base_transpose210_f64(double * B, double const * A, int64_t n)
{
    for (int64_t i=0; i<n; ++i) for (int64_t j=0; j<n; ++j) {
        double const * a = A + j*n + i;
        for (int64_t s=n; --s>=0; ++B, a+=n*n) { *B = *a; }
    }
}

// gemm(a, b, c): ra/ra.hh:407  c(i, ., j) += a(i, k) * b(., k, j): loops i, k, j with j innermost.
void base_gemm_f64(double * c, double const * a, double const * b, int64_t M, int64_t N, int64_t K)
{
    for (int64_t i=0; i<M; ++i) for (int64_t k=0; k<K; ++k) {
        double const aik = a[i*K+k];
        double * cj = c + i*N;
        double const * bj = b + k*N;
        for (int64_t s=N; --s>=0; ++cj, ++bj) { *cj += aik*(*bj); }
    }
}
// gemv(a, b, c): ra/ra.hh:418  c[i] += a[i,j]*b[j], j inner.
void base_gemv_f64(double * c, double const * a, double const * b, int64_t M, int64_t N)
{
    for (int64_t i=0; i<M; ++i, ++c) { double const * bj = b; for (int64_t s=N; --s>=0; ++a, ++bj) { *c += (*a)*(*bj); } }
}

} // extern "C"
