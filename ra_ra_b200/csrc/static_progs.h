// static_progs.h - the table of program shapes that have compile-time specialised kernels (static_kernels.cu), and the
// parameters those kernels take. Plain C++: included by the planner (g++) and by the kernels (nvcc).
//
// A static program is the planner's canonical postfix program with operand indices replaced by ROLES: role r = the r-th
// distinct operand in order of first appearance; `yrole` says which role (if any) is the destination's own old value
// (`dst OP= rhs` puts it first). Instructions and roles are TYPED, so mixed-type expressions (float arrays with a double
// vector, comparisons, casts) can be specialised exactly like the single-type ones.
#pragma once
#include "kcore.h"

namespace racu {

constexpr int SP_MAXINS = 32, SP_MAXIN = 12; // (12 = RACU_MAX_OPND: the 2-D mass stencil of examples/laplace2d.cc:48 has 7 slices + 3 scalars)
struct SIns { uint8_t op, arg, type, aux; };
struct SProg
{
    int n;
    SIns ins[SP_MAXINS];
    int nin;
    int yrole;
    bool fold;
    uint8_t rtype[SP_MAXIN]; // element type of each role
    uint8_t otype;           // maps: destination dtype (== type of the program result). folds: accumulator type
};

// shapes; each is instantiated for the element types listed in static_prog()
enum SShape {
    SH_COPY = 0,      // y = a
    SH_NEG,           // y = -a
    SH_YADD, SH_YSUB, SH_YMUL, SH_YDIV,   // y OP= a
    SH_ADD, SH_SUB, SH_MUL, SH_DIV,       // y = a OP b
    SH_YFMA, SH_YFMS,                     // y += a*b, y -= a*b      (B += A*C; cghs: x += t*r, g -= t*p)
    SH_MULADD, SH_YMULADD,                // y = a*b + c ; y = y*b + c (cghs: r = r*gam + g)
    SH_YFMA_ADD,                          // y += a*b + c             (examples/read-me.cc:32 shape: B += A*C + v)
    SH_F_ID, SH_F_SQR, SH_F_MUL, SH_F_SQRDIFF, // folds of a, a*a, a*b, (a-b)^2  (sum, reduce_sqrm, dot, reduce_sqrm(a-b))
    SH_NSHAPE
};
constexpr int SP_NUNIFORM = SH_NSHAPE*3; // every shape x {f32, f64, i32}
// mixed-type programs, after the uniform ones
enum {
    SPM_YFMA_ADD_F32_VD = SP_NUNIFORM, // float y += float a * float b + double v      (examples/read-me.cc:32 with std::vector<double>)
    SPM_YADD_F32_SD_MUL,               // float y += double s * float a                 (B += 2.5*A with a double literal)
    SPM_YSUB_F32_SD_MUL,               // float y -= double s * float a
    SPM_F_SUM_F32_D,                   // double acc += float a                         (double c += float array)
    SP_COUNT,
    SP_JIT = SP_COUNT                  // a program specialised at run time (jit.cc): static_prog(SP_JIT) is RACU_JIT_PROG
};

constexpr SIns sP(int r, int t) { return SIns {RACU_OP_PUSH, uint8_t(r), uint8_t(t), 0}; }
constexpr SIns sO(int op, int t) { return SIns {uint8_t(op), 0, uint8_t(t), 0}; }
constexpr SIns sC(int from, int to) { return SIns {RACU_OP_CAST, 0, uint8_t(to), uint8_t(from)}; }

constexpr SProg
uniform_prog(int shape, int T)
{
    uint8_t const t = uint8_t(T);
    switch (shape) {
    case SH_COPY: return SProg {1, {sP(0, T)}, 1, -1, false, {t}, t};
    case SH_NEG: return SProg {2, {sP(0, T), sO(RACU_OP_NEG, T)}, 1, -1, false, {t}, t};
    case SH_YADD: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_ADD, T)}, 2, 0, false, {t, t}, t};
    case SH_YSUB: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_SUB, T)}, 2, 0, false, {t, t}, t};
    case SH_YMUL: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_MUL, T)}, 2, 0, false, {t, t}, t};
    case SH_YDIV: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_DIV, T)}, 2, 0, false, {t, t}, t};
    case SH_ADD: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_ADD, T)}, 2, -1, false, {t, t}, t};
    case SH_SUB: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_SUB, T)}, 2, -1, false, {t, t}, t};
    case SH_MUL: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_MUL, T)}, 2, -1, false, {t, t}, t};
    case SH_DIV: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_DIV, T)}, 2, -1, false, {t, t}, t};
    case SH_YFMA: return SProg {5, {sP(0, T), sP(1, T), sP(2, T), sO(RACU_OP_MUL, T), sO(RACU_OP_ADD, T)}, 3, 0, false, {t, t, t}, t};
    case SH_YFMS: return SProg {5, {sP(0, T), sP(1, T), sP(2, T), sO(RACU_OP_MUL, T), sO(RACU_OP_SUB, T)}, 3, 0, false, {t, t, t}, t};
    case SH_MULADD: return SProg {5, {sP(0, T), sP(1, T), sO(RACU_OP_MUL, T), sP(2, T), sO(RACU_OP_ADD, T)}, 3, -1, false, {t, t, t}, t};
    case SH_YMULADD: return SProg {5, {sP(0, T), sP(1, T), sO(RACU_OP_MUL, T), sP(2, T), sO(RACU_OP_ADD, T)}, 3, 0, false, {t, t, t}, t};
    case SH_YFMA_ADD: return SProg {7, {sP(0, T), sP(1, T), sP(2, T), sO(RACU_OP_MUL, T), sP(3, T), sO(RACU_OP_ADD, T), sO(RACU_OP_ADD, T)}, 4, 0, false, {t, t, t, t}, t};
    case SH_F_ID: return SProg {1, {sP(0, T)}, 1, -1, true, {t}, t};
    case SH_F_SQR: return SProg {2, {sP(0, T), sO(RACU_OP_SQR, T)}, 1, -1, true, {t}, t};
    case SH_F_MUL: return SProg {3, {sP(0, T), sP(1, T), sO(RACU_OP_MUL, T)}, 2, -1, true, {t, t}, t};
    case SH_F_SQRDIFF: return SProg {4, {sP(0, T), sP(1, T), sO(RACU_OP_SUB, T), sO(RACU_OP_SQR, T)}, 2, -1, true, {t, t}, t};
    default: return SProg {0, {}, 0, -1, false, {}, 0};
    }
}

constexpr SProg
static_prog(int id)
{
    constexpr int F = RACU_F32, D = RACU_F64;
#ifdef RACU_JIT_PROG
    if (id==SP_JIT) return RACU_JIT_PROG;
#endif
    if (id<SP_NUNIFORM) {
        constexpr int types[3] = {RACU_F32, RACU_F64, RACU_I32};
        return uniform_prog(id/3, types[id%3]);
    }
    switch (id) {
    case SPM_YFMA_ADD_F32_VD: // y(f32) -> f64 ; a*b in f32 -> f64 ; + v(f64) ; + ; -> f32
        return SProg {10, {sP(0, F), sC(F, D), sP(1, F), sP(2, F), sO(RACU_OP_MUL, F), sC(F, D), sP(3, D), sO(RACU_OP_ADD, D), sO(RACU_OP_ADD, D), sC(D, F)},
                      4, 0, false, {F, F, F, D}, F};
    case SPM_YADD_F32_SD_MUL: // y(f32) -> f64 ; s(f64) * f64(a) ; + ; -> f32
        return SProg {8, {sP(0, F), sC(F, D), sP(1, D), sP(2, F), sC(F, D), sO(RACU_OP_MUL, D), sO(RACU_OP_ADD, D), sC(D, F)}, 3, 0, false, {F, D, F}, F};
    case SPM_YSUB_F32_SD_MUL:
        return SProg {8, {sP(0, F), sC(F, D), sP(1, D), sP(2, F), sC(F, D), sO(RACU_OP_MUL, D), sO(RACU_OP_SUB, D), sC(D, F)}, 3, 0, false, {F, D, F}, F};
    case SPM_F_SUM_F32_D: // fold: a(f32) -> f64 accumulator
        return SProg {2, {sP(0, F), sC(F, D)}, 1, -1, true, {F}, D};
    default: return SProg {0, {}, 0, -1, false, {}, 0};
    }
}

constexpr bool
prog_wide(SProg const & P)
{
    bool w = P.otype==RACU_I64 || P.otype==RACU_F64;
    for (int i=0; i<P.n; ++i) w = w || P.ins[i].type==RACU_I64 || P.ins[i].type==RACU_F64 || ((P.ins[i].op==RACU_OP_CAST || P.ins[i].op==RACU_OP_LOAD) && (P.ins[i].aux==RACU_I64 || P.ins[i].aux==RACU_F64));
    for (int r=0; r<P.nin; ++r) w = w || P.rtype[r]==RACU_I64 || P.rtype[r]==RACU_F64;
    return w;
}

constexpr int SV = 4; // elements per static vector (16 bytes of 4-byte elements, 32 bytes of 8-byte ones)
constexpr int SLANES_CT = 8; // lane-wise map kernel: warp tiles (of 32*SV elements) a warp walks per work unit

enum { SK_SCALAR = 0, SK_VEC = 1, SK_VECREV = 2 /* unit step backwards: aligned loads, lanes reversed */, SK_ROW = 3 /* step 0 on axis 0: one value per row */,
       SK_SEQ = 4 /* iota / tindex (ra/expr.hh:79-100): value = origin + element offset, no memory */,
       SK_PTR = 5 /* gathered array (RACU_OP_LOAD): in[] = base address, sA[][0..1] = valid element offsets; never fetched */,
       SK_VECU = 6 /* unit step forwards, rows NOT 16-byte aligned (slices shifted by one element: A(I, J+1) of a user-written stencil,
                      examples/laplace2d.cc:48, bench/bench-stencil1.cc:45): lane-wise kernel only (smap_lanes_body) */,
       SK_VECREVU = 7 /* unit step backwards, not aligned: lane-wise kernel only */ };

constexpr uint64_t SK_RUNTIME = ~uint64_t(0); // KINDS template argument: operand kinds read from SParams::kind at run time
constexpr uint64_t sk_pack(int k0, int k1=0, int k2=0, int k3=0) { return uint64_t(k0) | (uint64_t(k1)<<4) | (uint64_t(k2)<<8) | (uint64_t(k3)<<12); }

struct SParams
{
    int32_t id, mode, fold_op, nchunks, rankB, rankA;
    int32_t wide_partials;        // folds: partial / finish slots are 64-bit
    int32_t general;              // maps: some input is SK_VECREV / SK_ROW
    uint64_t identity;            // accumulator start, slot bits
    int64_t n;                    // elements along the flat axis (row length)
    int64_t nB;                   // fold-outer: folded rows
    int64_t chunk_len, nk;        // fold-inner: vectors per chunk; kept elements
    void * dst;
    void * partial;
    uint64_t in[SP_MAXIN];        // role -> base address (vector inputs) or value bits in the role's type (scalars)
    uint8_t kind[SP_MAXIN];       // SK_*
    uint8_t reused[SP_MAXIN];     // folds: the operand does not move along the kept rows (every row re-reads it): cache in L1
    int32_t inner_rows;           // fold-inner: one warp per kept row (sfold_inner_rows_body) instead of one CTA per (row, chunk)
    int32_t zero;                 // always 0, and the compiler cannot know: see loads_issued() in static_device.h
    int32_t lanes;                // maps: lane-wise kernel (smap_lanes_body): rows that are not 16-byte aligned or not whole vectors
    int32_t pad1;
    int64_t sB[SP_MAXIN][KGR];    // role -> element strides over group B
    KDim dB[KGR];
    // maps over up to KGR uncollapsible axes (axis 0 vectorised)
    int64_t nvec, vpr;            // vectors in total, vectors per row
    KDim vprdiv;
    KDim dA[KGR];
    int64_t sA[SP_MAXIN][KGR];    // role -> element strides over group A (sA[.][0] is 1, -1 or 0)
    int64_t dsA[KGR];             // destination strides (dsA[0] == 1)
};

} // namespace racu
