// special.h - entry points of the specialised kernels (stencil.cu, gemm.cu, transpose.cu) and shared error helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include "../../include/racu.h"

namespace racu {
int fail(int status, std::string const & msg);
int cuda_fail(cudaError_t e, char const * where);
void set_kernel(const char * k);
// Pattern recognition in racu_map: stencils over shifted views of one array, gemm/gemv written with insert<1>,
// permuted copies. Sets *handled when a specialised kernel ran the whole descriptor.
int try_special_map(racu_desc const * d, cudaStream_t stream, int * handled);
int try_special_stencil(racu_stencil_desc const * s, cudaStream_t stream, int * handled);
int stencil_push(racu_stencil_desc const * s, racu_halo_push const * h, cudaStream_t stream);
int stencil_steps(racu_stencil_desc const * s, int steps, void * scratch, cudaStream_t stream);
int racu_stencil_c(const racu_stencil_desc * s, void * stream); // racu_stencil (special.cc), callable from the other translation units
int try_tcgen05_sgemm(racu_gemm_desc const * g, cudaStream_t stream, int * handled);
int try_special_gemm(racu_gemm_desc const * g, cudaStream_t stream, int * handled);
int build_stencil_desc(racu_stencil_desc const * s, racu_desc & d);

// Do two strided views touch a common byte range? (Bounding ranges, like the planner's alias test in plan.cc: the specialised
// kernels read and write in an order of their own, so a destination that overlaps a source must not reach them.)
inline bool
views_overlap(void const * a, int ranka, int64_t const * lena, int64_t const * stra, void const * b, int rankb, int64_t const * lenb, int64_t const * strb, size_t es)
{
    auto ext = [es](void const * p, int rank, int64_t const * len, int64_t const * st, uintptr_t & lo, uintptr_t & hi) {
        int64_t mn = 0, mx = 0;
        for (int k=0; k<rank; ++k) { if (len[k]<=0) continue; int64_t d = (len[k]-1)*st[k]; if (d<0) mn += d; else mx += d; }
        lo = uintptr_t(p) + uintptr_t(mn*int64_t(es));
        hi = uintptr_t(p) + uintptr_t(mx*int64_t(es)) + es;
    };
    uintptr_t alo, ahi, blo, bhi;
    ext(a, ranka, lena, stra, alo, ahi);
    ext(b, rankb, lenb, strb, blo, bhi);
    return alo<bhi && blo<ahi;
}
} // namespace racu
