// static_progs.h - the table of program shapes that have compile-time specialised kernels (static_kernels.cu), and the
// parameters those kernels take. Plain C++: included by the planner (g++) and by the kernels (nvcc).
// A program is a postfix sequence over ROLES: role r = the r-th distinct operand in order of first appearance;
// `yrole` says which role (if any) is the destination's own old value (`dst OP= rhs` puts it first).
#pragma once
#include "kcore.h"

namespace racu {

constexpr int SP_MAXINS = 8, SP_MAXIN = 4;
struct SIns { uint8_t op, arg; };
struct SProg { int n; SIns ins[SP_MAXINS]; int nin; int yrole; bool fold; };

enum {
    SP_COPY = 0,      // y = a
    SP_NEG,           // y = -a
    SP_YADD, SP_YSUB, SP_YMUL, SP_YDIV,   // y OP= a
    SP_ADD, SP_SUB, SP_MUL, SP_DIV,       // y = a OP b
    SP_YFMA, SP_YFMS,                     // y += a*b, y -= a*b      (B += A*C; cghs: x += t*r, g -= t*p)
    SP_MULADD, SP_YMULADD,                // y = a*b + c ; y = y*b + c (cghs: r = r*gam + g)
    SP_YFMA_ADD,                          // y += a*b + c             (examples/read-me.cc:32 shape: B += A*C + v)
    SF_ID, SF_SQR, SF_MUL, SF_SQRDIFF,    // folds of a, a*a, a*b, (a-b)^2  (sum, reduce_sqrm, dot, reduce_sqrm(a-b))
    SP_COUNT
};

constexpr SIns sP(int r) { return SIns {RACU_OP_PUSH, uint8_t(r)}; }
constexpr SIns sO(int op) { return SIns {uint8_t(op), 0}; }

constexpr SProg
static_prog(int id)
{
    switch (id) {
    case SP_COPY: return SProg {1, {sP(0)}, 1, -1, false};
    case SP_NEG: return SProg {2, {sP(0), sO(RACU_OP_NEG)}, 1, -1, false};
    case SP_YADD: return SProg {3, {sP(0), sP(1), sO(RACU_OP_ADD)}, 2, 0, false};
    case SP_YSUB: return SProg {3, {sP(0), sP(1), sO(RACU_OP_SUB)}, 2, 0, false};
    case SP_YMUL: return SProg {3, {sP(0), sP(1), sO(RACU_OP_MUL)}, 2, 0, false};
    case SP_YDIV: return SProg {3, {sP(0), sP(1), sO(RACU_OP_DIV)}, 2, 0, false};
    case SP_ADD: return SProg {3, {sP(0), sP(1), sO(RACU_OP_ADD)}, 2, -1, false};
    case SP_SUB: return SProg {3, {sP(0), sP(1), sO(RACU_OP_SUB)}, 2, -1, false};
    case SP_MUL: return SProg {3, {sP(0), sP(1), sO(RACU_OP_MUL)}, 2, -1, false};
    case SP_DIV: return SProg {3, {sP(0), sP(1), sO(RACU_OP_DIV)}, 2, -1, false};
    case SP_YFMA: return SProg {5, {sP(0), sP(1), sP(2), sO(RACU_OP_MUL), sO(RACU_OP_ADD)}, 3, 0, false};
    case SP_YFMS: return SProg {5, {sP(0), sP(1), sP(2), sO(RACU_OP_MUL), sO(RACU_OP_SUB)}, 3, 0, false};
    case SP_MULADD: return SProg {5, {sP(0), sP(1), sO(RACU_OP_MUL), sP(2), sO(RACU_OP_ADD)}, 3, -1, false};
    case SP_YMULADD: return SProg {5, {sP(0), sP(1), sO(RACU_OP_MUL), sP(2), sO(RACU_OP_ADD)}, 3, 0, false};
    case SP_YFMA_ADD: return SProg {7, {sP(0), sP(1), sP(2), sO(RACU_OP_MUL), sP(3), sO(RACU_OP_ADD), sO(RACU_OP_ADD)}, 4, 0, false};
    case SF_ID: return SProg {1, {sP(0)}, 1, -1, true};
    case SF_SQR: return SProg {2, {sP(0), sO(RACU_OP_SQR)}, 1, -1, true};
    case SF_MUL: return SProg {3, {sP(0), sP(1), sO(RACU_OP_MUL)}, 2, -1, true};
    case SF_SQRDIFF: return SProg {4, {sP(0), sP(1), sO(RACU_OP_SUB), sO(RACU_OP_SQR)}, 2, -1, true};
    default: return SProg {0, {}, 0, -1, false};
    }
}

struct SParams
{
    int32_t id, dtype, mode, fold_op, nchunks, rankB;
    uint64_t identity;
    int64_t n;            // elements along the flat axis (row length)
    int64_t nB;           // fold-outer: folded rows
    int64_t chunk_len, nk;
    void * dst;
    void * partial;
    uint64_t in[SP_MAXIN];        // role -> base address (vector inputs) or value bits (scalars)
    uint8_t scalar[SP_MAXIN];
    uint8_t kind[SP_MAXIN];       // maps: SK_* below
    uint8_t reused[SP_MAXIN];     // folds: the operand does not move along the kept rows (every row re-reads it): cache in L1
    int64_t sB[SP_MAXIN][KGR];    // role -> element strides over group B
    KDim dB[KGR];
    // maps over up to KGR uncollapsible axes (axis 0 vectorised)
    int32_t rankA, rb;            // rb: rows blocked per CTA in fold-inner (1 or 4)
    int32_t chunk_major, pad2;    // fold-inner CTA order: blockIdx = chunk*nk + row instead of row*nchunks + chunk
    int64_t nvec, vpr;            // vectors in total, vectors per row
    KDim vprdiv;
    KDim dA[KGR];
    int64_t sA[SP_MAXIN][KGR];    // role -> element strides over group A (sA[.][0] is 1, -1 or 0)
    int64_t dsA[KGR];             // destination strides (dsA[0] == 1)
};

enum { SK_SCALAR = 0, SK_VEC = 1, SK_VECREV = 2 /* unit step backwards: aligned 16-byte load, lanes reversed */, SK_ROW = 3 /* step 0 on axis 0: one value per row */ };

} // namespace racu
