// special_kernels.cu - specialised kernels behind racu_map / racu_stencil / racu_gemm. (Placeholders: nothing matches yet.)
#include "special.h"
namespace racu {
int try_special_map(racu_desc const *, cudaStream_t, int * handled) { *handled = 0; return RACU_OK; }
int try_special_stencil(racu_stencil_desc const *, cudaStream_t, int * handled) { *handled = 0; return RACU_OK; }
int try_special_gemm(racu_gemm_desc const *, cudaStream_t, int * handled) { *handled = 0; return RACU_OK; }
}
