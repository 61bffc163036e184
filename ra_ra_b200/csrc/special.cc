// special.cc - entry points that are expressed through the generic traversal: racu_stencil / racu_gemm / racu_gemv build
// the reference's own expression (same association order) and hand it to racu_map, which routes to a specialised kernel
// when one matches (try_special_map) and to the fused ply kernels otherwise.
#include <cuda_runtime.h>

#include <cstring>

#include "special.h"

namespace racu {

static void
mem_operand(racu_operand & o, int dtype, void const * ptr, int rank, int64_t const * len, int64_t const * stride)
{
    std::memset(&o, 0, sizeof(o));
    o.kind = RACU_MEM; o.dtype = dtype; o.rank = rank; o.base.ptr = const_cast<void *>(ptr);
    for (int k=0; k<rank; ++k) { o.len[k] = len[k]; o.stride[k] = stride[k]; }
}
static void
scalar_operand(racu_operand & o, int dtype, double f, int64_t i)
{
    std::memset(&o, 0, sizeof(o));
    o.kind = RACU_SCALAR; o.dtype = dtype; o.rank = 0;
    if (dtype==RACU_F32 || dtype==RACU_F64) o.base.f = f; else o.base.i = i;
}
static void ins(racu_desc & d, int op, int type, int arg=0, int aux=0)
{
    racu_ins & i = d.ins[d.nins++];
    i.op = uint8_t(op); i.type = uint8_t(type); i.arg = uint8_t(arg); i.aux = uint8_t(aux);
}

// The reference's slice expressions, flattened exactly as the host header would:
//   3D7  Anext(I,J,K) = -6*A(I,J,K) + A(I+1,J,K) + A(I,J+1,K) + A(I,J,K+1) + A(I-1,J,K) + A(I,J-1,K) + A(I,J,K-1)
//   2D5  Anext(I,J) = -4*A(I,J) + A(I+1,J) + A(I,J+1) + A(I-1,J) + A(I,J-1)
//   1D3  Anext(I) = -2*A(I) + A(I+1) + A(I-1)
//   2D5_STIFF  w(i,j) = 4*v(i,j) - v(i-1,j) - v(i,j-1) - v(i+1,j) - v(i,j+1)
int
build_stencil_desc(racu_stencil_desc const * s, racu_desc & d)
{
    std::memset(&d, 0, sizeof(d));
    int const rank = s->rank, T = s->dtype;
    if (T!=RACU_F32 && T!=RACU_F64) return fail(RACU_ERR_UNSUPPORTED, "stencil dtype must be f32 or f64");
    int want = s->kind==RACU_STENCIL_3D7 ? 3 : s->kind==RACU_STENCIL_1D3 ? 1 : (s->kind==RACU_STENCIL_2D5 || s->kind==RACU_STENCIL_2D5_STIFF) ? 2 : -1;
    if (want!=rank) return fail(RACU_ERR_BAD_DESC, "stencil kind does not match rank");
    int64_t len[3], lo[3];
    for (int k=0; k<rank; ++k) { len[k] = s->len[k]-2; lo[k] = 1; if (len[k]<0) len[k] = 0; }
    int64_t lo0 = s->lo0, hi0 = s->hi0;
    if (0==lo0 && 0==hi0) hi0 = s->len[0];
    if (lo0<1) lo0 = 1;
    if (hi0>s->len[0]-1) hi0 = s->len[0]-1;
    lo[0] = lo0; len[0] = hi0>lo0 ? hi0-lo0 : 0;
    size_t const es = T==RACU_F32 ? 4 : 8;
    auto at = [&](void const * base, int64_t const * st, int axis, int shift) {
        int64_t off = 0;
        for (int k=0; k<rank; ++k) off += (lo[k] + (k==axis ? shift : 0))*st[k];
        return static_cast<char const *>(base) + off*int64_t(es);
    };
    // operand order = order of appearance in the expression
    struct Sh { int axis, shift; };
    Sh sh[8]; int nsh = 0;
    double coef = 0; bool stiff = false;
    switch (s->kind) {
    case RACU_STENCIL_3D7: coef = -6; sh[nsh++] = {0, +1}; sh[nsh++] = {1, +1}; sh[nsh++] = {2, +1}; sh[nsh++] = {0, -1}; sh[nsh++] = {1, -1}; sh[nsh++] = {2, -1}; break;
    case RACU_STENCIL_2D5: coef = -4; sh[nsh++] = {0, +1}; sh[nsh++] = {1, +1}; sh[nsh++] = {0, -1}; sh[nsh++] = {1, -1}; break;
    case RACU_STENCIL_1D3: coef = -2; sh[nsh++] = {0, +1}; sh[nsh++] = {0, -1}; break;
    default: coef = 4; stiff = true; sh[nsh++] = {0, -1}; sh[nsh++] = {1, -1}; sh[nsh++] = {0, +1}; sh[nsh++] = {1, +1}; break;
    }
    mem_operand(d.opnd[0], T, at(s->dst, s->dst_stride, -1, 0), rank, len, s->dst_stride);
    scalar_operand(d.opnd[1], RACU_I32, 0, int64_t(coef));   // the literal -6 is an int: -6*A -> T(-6)*A
    mem_operand(d.opnd[2], T, at(s->src, s->src_stride, -1, 0), rank, len, s->src_stride);
    for (int q=0; q<nsh; ++q) mem_operand(d.opnd[3+q], T, at(s->src, s->src_stride, sh[q].axis, sh[q].shift), rank, len, s->src_stride);
    d.nopnd = 3+nsh; d.dst = 0; d.assign = RACU_ASSIGN_SET;
    ins(d, RACU_OP_PUSH, RACU_I32, 1); ins(d, RACU_OP_CAST, T, 0, RACU_I32); ins(d, RACU_OP_PUSH, T, 2); ins(d, RACU_OP_MUL, T);
    for (int q=0; q<nsh; ++q) { ins(d, RACU_OP_PUSH, T, 3+q); ins(d, stiff ? RACU_OP_SUB : RACU_OP_ADD, T); }
    return RACU_OK;
}

int racu_stencil_c(const racu_stencil_desc * s, void * stream) { return racu_stencil(s, stream); }

} // namespace racu

using namespace racu;

extern "C" {

int racu_map(const racu_desc * d, void * stream);

int racu_stencil(const racu_stencil_desc * s, void * stream)
{
    if (!s || !s->src || !s->dst || s->rank<1 || s->rank>3) return fail(RACU_ERR_BAD_DESC, "bad stencil descriptor");
    int handled = 0;
    if (int e = try_special_stencil(s, cudaStream_t(stream), &handled)) return e;
    if (handled) return RACU_OK;
    racu_desc d;
    if (int e = build_stencil_desc(s, d)) return e;
    return racu_map(&d, stream);
}

int racu_stencil_steps(const racu_stencil_desc * s, int32_t steps, void * scratch, void * stream)
{
    if (!s || !s->src || !s->dst || s->rank<1 || s->rank>3) return fail(RACU_ERR_BAD_DESC, "bad stencil descriptor");
    return stencil_steps(s, steps, scratch, cudaStream_t(stream));
}

int racu_stencil_push(const racu_stencil_desc * s, const racu_halo_push * h, void * stream)
{
    if (!s || !h || !s->src || !s->dst) return fail(RACU_ERR_BAD_DESC, "bad stencil descriptor");
    return stencil_push(s, h, cudaStream_t(stream));
}

// c(all, insert<1>) += a * b(insert<1>): expression axes (i, k, j). ra/ra.hh:407.
int racu_gemm(const racu_gemm_desc * g, void * stream)
{
    if (!g || !g->a || !g->b || !g->c) return fail(RACU_ERR_BAD_DESC, "bad gemm descriptor");
    int const T = g->dtype;
    if (T!=RACU_F32 && T!=RACU_F64) return fail(RACU_ERR_UNSUPPORTED, "gemm dtype must be f32 or f64");
    int handled = 0;
    if (int e = try_special_gemm(g, cudaStream_t(stream), &handled)) return e;
    if (handled) return RACU_OK;
    racu_desc d;
    std::memset(&d, 0, sizeof(d));
    int64_t lc[3] = {g->M, RACU_UNB, g->N}, sc[3] = {g->ldc, 0, 1};
    int64_t la[2] = {g->M, g->K}, sa[2] = {g->lda, 1};
    int64_t lb[3] = {RACU_UNB, g->K, g->N}, sb[3] = {0, g->ldb, 1};
    mem_operand(d.opnd[0], T, g->c, 3, lc, sc);
    mem_operand(d.opnd[1], T, g->a, 2, la, sa);
    mem_operand(d.opnd[2], T, g->b, 3, lb, sb);
    d.nopnd = 3; d.dst = 0; d.assign = RACU_ASSIGN_ADD;
    ins(d, RACU_OP_PUSH, T, 1); ins(d, RACU_OP_PUSH, T, 2); ins(d, RACU_OP_MUL, T);
    return racu_map(&d, stream);
}

// trans=0: c += a * b(insert<1>) (gemv ra/ra.hh:418). trans=1: c(insert<1>) += a * b (gevm ra/ra.hh:431), a = x vector.
int racu_gemv(int dtype, int trans, int64_t M, int64_t N, const void * a, int64_t lda, const void * x, int64_t incx,
              void * y, int64_t incy, void * stream)
{
    if (!a || !x || !y) return fail(RACU_ERR_BAD_DESC, "null argument");
    if (dtype!=RACU_F32 && dtype!=RACU_F64) return fail(RACU_ERR_UNSUPPORTED, "gemv dtype must be f32 or f64");
    racu_desc d;
    std::memset(&d, 0, sizeof(d));
    int64_t la[2] = {M, N}, sa[2] = {lda, 1};
    if (!trans) {
        int64_t ly[1] = {M}, sy[1] = {incy};
        int64_t lx[2] = {RACU_UNB, N}, sx[2] = {0, incx};
        mem_operand(d.opnd[0], dtype, y, 1, ly, sy);
        mem_operand(d.opnd[1], dtype, a, 2, la, sa);
        mem_operand(d.opnd[2], dtype, x, 2, lx, sx);
        ins(d, RACU_OP_PUSH, dtype, 1); ins(d, RACU_OP_PUSH, dtype, 2); ins(d, RACU_OP_MUL, dtype);
    } else {
        int64_t ly[2] = {RACU_UNB, N}, sy[2] = {0, incy};
        int64_t lx[1] = {M}, sx[1] = {incx};
        mem_operand(d.opnd[0], dtype, y, 2, ly, sy);
        mem_operand(d.opnd[1], dtype, x, 1, lx, sx);
        mem_operand(d.opnd[2], dtype, a, 2, la, sa);
        ins(d, RACU_OP_PUSH, dtype, 1); ins(d, RACU_OP_PUSH, dtype, 2); ins(d, RACU_OP_MUL, dtype);
    }
    d.nopnd = 3; d.dst = 0; d.assign = RACU_ASSIGN_ADD;
    return racu_map(&d, stream);
}

} // extern "C"
