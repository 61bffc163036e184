// transpose.cu - tiled permuted copy: `B = transpose(A, perm)` and any plain copy whose source is contiguous along a
// different axis than the destination (ra/arrays.hh:431-461 views traversed by ra/ply.hh:21-86).
// The reference reads such a source with a large inner step (TODO at ra/ply.hh:11,34: no tiling). Here a 32x32 tile goes
// through shared memory so that BOTH sides move in full 128-byte lines: rows of the tile are read along the source's unit
// axis, columns written along the destination's. Remaining axes are batch axes. Pure data movement: bit exact.
#include <cuda_runtime.h>

#include "launch.h"
#include "special.h"

namespace racu {

struct TParams
{
    void * dst; void const * src;
    int64_t n0, nS;          // extents of the destination-unit axis and the source-unit axis
    int64_t d0, dS, s0, sS;  // element strides of dst / src over those two axes (d0 == 1, |sS| == 1)
    int32_t nbatch;          // number of batch axes (0..2)
    uint32_t lb0;            // extent of batch axis 0 (batch index = b0 + lb0*b1)
    int64_t db[2], sb[2];    // strides over the batch axes
};

template <class T> __global__ void __launch_bounds__(256)
transpose_kernel(const __grid_constant__ TParams p)
{
    __shared__ T tile[32][33];
    int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int64_t const t0 = int64_t(blockIdx.x)*32, tS = int64_t(blockIdx.y)*32;
    uint32_t const b = blockIdx.z;
    uint32_t const b1 = p.nbatch>1 ? b/p.lb0 : 0, b0 = b - b1*p.lb0;
    T const * src = static_cast<T const *>(p.src) + int64_t(b0)*p.sb[0] + int64_t(b1)*p.sb[1];
    T * dst = static_cast<T *>(p.dst) + int64_t(b0)*p.db[0] + int64_t(b1)*p.db[1];
#pragma unroll
    for (int r=0; r<32; r+=8) {
        int64_t i0 = t0 + ty + r, iS = tS + tx;
        if (i0<p.n0 && iS<p.nS) tile[ty+r][tx] = src[i0*p.s0 + iS*p.sS];
    }
    __syncthreads();
#pragma unroll
    for (int r=0; r<32; r+=8) {
        int64_t iS = tS + ty + r, i0 = t0 + tx;
        if (i0<p.n0 && iS<p.nS) dst[iS*p.dS + i0*p.d0] = tile[tx][ty+r];
    }
}

// Is this plan a single-source copy with mismatched unit axes? If so run it here.
bool
try_transpose(Plan const & plan, cudaStream_t stream, cudaError_t & err)
{
    err = cudaSuccess;
    KParams const & kp = plan.kp;
    if (plan.is_static || kp.mode!=K_MAP || kp.nins!=1 || kp.ins[0].op!=RACU_OP_PUSH || kp.rankA<2 || kp.rankA>4) return false;
    if (std::getenv("RACU_NO_TRANSPOSE")) return false;
    KOpnd const & y = kp.opnd[kp.dst];
    KOpnd const & x = kp.opnd[kp.ins[0].arg];
    if (x.kind!=RACU_MEM || x.dtype!=y.dtype || (x.esize!=4 && x.esize!=8) || 1!=y.sA[0]) return false;
    if (1==x.sA[0] || -1==x.sA[0]) return false; // same unit axis: the fused map kernel is already coalesced
    int js = -1;
    for (int j=1; j<kp.rankA; ++j) { if (1==x.sA[j] || -1==x.sA[j]) { js = j; break; } }
    if (js<0) return false;
    int64_t len[KGR];
    len[0] = kp.len0;
    for (int j=1; j<kp.rankA; ++j) len[j] = kp.dA[j].len;
    if (len[0]<8 || len[js]<8) return false; // tiny axes: tiles would be mostly empty
    TParams p {};
    p.dst = reinterpret_cast<void *>(uintptr_t(y.base)); p.src = reinterpret_cast<void const *>(uintptr_t(x.base));
    p.n0 = len[0]; p.nS = len[js];
    p.d0 = y.sA[0]; p.dS = y.sA[js]; p.s0 = x.sA[0]; p.sS = x.sA[js];
    int nb = 0;
    int64_t nbatch_total = 1, lb[2] = {1, 1};
    for (int j=1; j<kp.rankA; ++j) {
        if (j==js) continue;
        if (nb>=2) return false;
        lb[nb] = len[j]; p.db[nb] = y.sA[j]; p.sb[nb] = x.sA[j]; nbatch_total *= len[j]; ++nb;
    }
    p.nbatch = nb; p.lb0 = uint32_t(lb[0]);
    int64_t gx = (p.n0+31)/32, gy = (p.nS+31)/32;
    if (gx>=(int64_t(1)<<31) || gy>65535 || nbatch_total>65535) return false;
    dim3 grid {unsigned(gx), unsigned(gy), unsigned(nbatch_total)};
    if (4==x.esize) transpose_kernel<uint32_t><<<grid, 256, 0, stream>>>(p); else transpose_kernel<uint64_t><<<grid, 256, 0, stream>>>(p);
    launch_count_add(1);
    err = cudaGetLastError();
    set_kernel("transpose_tiled");
    return true;
}

} // namespace racu
