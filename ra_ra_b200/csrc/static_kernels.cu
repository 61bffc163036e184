// static_kernels.cu - compile-time specialisations of the ply kernels for the hottest program shapes.
//
// The generic kernels (ply_kernels.cu) interpret the postfix program; for short programs the interpretive overhead, not
// HBM, is the limit. Here the SAME typed postfix program is a constexpr array (static_progs.h): the evaluator loop is fully
// unrolled by the compiler, every switch on opcode / type folds away, the value stack lives in registers, and operands
// are read with 128-bit loads issued several deep before the first use. Same ops, same order, same types, no contraction
// (-fmad=false): results are bit-identical to the interpreter's and to the oracle's.
//
// The planner (plan.cc: match_static) picks a specialisation when the canonical program equals a table entry (opcodes,
// types, roles) and every operand is a 16-byte aligned unit-step vector (forward or reversed), one value per row, or a
// scalar; maps may have up to 4 uncollapsible axes, fold-inner any kept rows, fold-outer one folded axis.
//
// NOTE (measured): any RUN-TIME choice between load instruction forms inside the unrolled load loop makes the compiler load
// into temporaries and select right after each load, which serialises the loads on their latency (dot 7.0 -> 3.0 TB/s).
// Operand kinds that change the load form are therefore compile-time template switches (GENERAL, REUSE).
#include <cuda_runtime.h>
#include "launch.h"
#include "static_device.h"

namespace racu {

// kernels: thin __global__ wrappers of the bodies in static_device.h
template <int ID, bool GENERAL> __global__ void __launch_bounds__(KBLOCK)
smap_kernel(const __grid_constant__ SParams p) { smap_body<ID, GENERAL>(p); }
template <int ID, int FOLD, bool REUSE> __global__ void __launch_bounds__(KBLOCK)
sfold_inner_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f) { sfold_inner_body<ID, FOLD, REUSE>(p, f); }
template <int ID, int FOLD> __global__ void __launch_bounds__(KBLOCK)
sfold_outer_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f) { sfold_outer_body<ID, FOLD>(p, f); }

// ---------------- dispatch ----------------

template <int ID> static cudaError_t
launch_one(Plan const & plan, dim3 grid, cudaStream_t stream)
{
    SParams const & p = plan.sp;
    constexpr SProg P = SPT<ID>::P;
    if (p.mode==K_MAP) {
        if constexpr (!P.fold) {
            if (p.general) smap_kernel<ID, true><<<grid, KBLOCK, 0, stream>>>(p); else smap_kernel<ID, false><<<grid, KBLOCK, 0, stream>>>(p);
        }
    } else if constexpr (P.fold) {
        bool any = false;
        for (int k=0; k<SP_MAXIN; ++k) any = any || p.reused[k];
#define RACU_SF(KERNEL, ...) \
        switch (p.fold_op) { \
        case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: KERNEL<ID, RACU_ASSIGN_ADD __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        case RACU_ASSIGN_AMIN: KERNEL<ID, RACU_ASSIGN_AMIN __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        case RACU_ASSIGN_AMAX: KERNEL<ID, RACU_ASSIGN_AMAX __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        default: KERNEL<ID, RACU_ASSIGN_MUL __VA_ARGS__><<<grid, KBLOCK, 0, stream>>>(p, plan.fp); break; \
        }
        if (p.mode==K_FOLD_INNER) {
            if (any) { RACU_SF(sfold_inner_kernel, , true) } else { RACU_SF(sfold_inner_kernel, , false) }
        } else { RACU_SF(sfold_outer_kernel) }
#undef RACU_SF
    }
    return cudaGetLastError();
}

// The table is compiled in SP_NPART translation units (ids congruent to SP_PART), in parallel: build.py.
#ifndef SP_NPART
#define SP_NPART 1
#define SP_PART 0
#endif

template <int ID=SP_PART> static cudaError_t
launch_id(Plan const & plan, dim3 grid, cudaStream_t stream)
{
    if constexpr (ID<SP_COUNT) {
        if (plan.sp.id==ID) return launch_one<ID>(plan, grid, stream);
        return launch_id<ID+SP_NPART>(plan, grid, stream);
    } else {
        return cudaErrorInvalidValue;
    }
}

#define RACU_CAT_(a, b) a##b
#define RACU_CAT(a, b) RACU_CAT_(a, b)
cudaError_t
RACU_CAT(launch_static_part, SP_PART)(Plan const & plan, cudaStream_t stream)
{
    dim3 grid(plan.grid_x, plan.grid_y);
    return launch_id<>(plan, grid, stream);
}

#if SP_PART==0
cudaError_t launch_static_part1(Plan const & plan, cudaStream_t stream);
cudaError_t launch_static_part2(Plan const & plan, cudaStream_t stream);
cudaError_t launch_static_part3(Plan const & plan, cudaStream_t stream);
cudaError_t
launch_static(Plan const & plan, cudaStream_t stream)
{
    if (SP_JIT==plan.sp.id) return launch_jit(plan, stream);
    switch (plan.sp.id % SP_NPART) {
    case 0: return launch_static_part0(plan, stream);
#if SP_NPART>1
    case 1: return launch_static_part1(plan, stream);
#endif
#if SP_NPART>2
    case 2: return launch_static_part2(plan, stream);
#endif
#if SP_NPART>3
    case 3: return launch_static_part3(plan, stream);
#endif
    default: return cudaErrorInvalidValue;
    }
}
#endif

} // namespace racu
