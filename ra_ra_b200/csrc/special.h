// special.h - entry points of the specialised kernels (stencil.cu, gemm.cu, transpose.cu) and shared error helpers.
#pragma once
#include <cuda_runtime.h>
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
int try_tcgen05_sgemm(racu_gemm_desc const * g, cudaStream_t stream, int * handled);
int try_special_gemm(racu_gemm_desc const * g, cudaStream_t stream, int * handled);
int build_stencil_desc(racu_stencil_desc const * s, racu_desc & d);
} // namespace racu
