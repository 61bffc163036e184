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
#include <map>
#include <mutex>
#include "launch.h"
#include "static_device.h"

namespace racu {

// kernels: thin __global__ wrappers of the bodies in static_device.h
template <int ID, bool GENERAL, uint64_t KINDS=SK_RUNTIME> __global__ void __launch_bounds__(KBLOCK)
smap_kernel(const __grid_constant__ SParams p) { smap_body<ID, GENERAL, KINDS>(p); }
template <int ID, uint64_t KINDS> __global__ void __launch_bounds__(KBLOCK)
smap_lanes_kernel(const __grid_constant__ SParams p) { smap_lanes_body<ID, KINDS>(p); }
template <int ID, int FOLD, int RMASK, uint64_t KINDS> __global__ void __launch_bounds__(KBLOCK)
sfold_inner_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f) { sfold_inner_body<ID, FOLD, RMASK, KINDS>(p, f); }
template <int ID, int FOLD, int RMASK, uint64_t KINDS> __global__ void __launch_bounds__(KBLOCK)
sfold_inner_rows_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f) { sfold_inner_rows_body<ID, FOLD, RMASK, KINDS>(p, f); }
template <int ID, int FOLD, uint64_t KINDS> __global__ void __launch_bounds__(KBLOCK)
sfold_outer_kernel(const __grid_constant__ SParams p, const __grid_constant__ FinishParams f) { sfold_outer_body<ID, FOLD, KINDS>(p, f); }

// ---------------- dispatch ----------------
// Pre-compiled fold variants per program. Sums (the configs' folds: sum, dot, reduce_sqrm, row / column sums, gemv, gevm) get the
// operand kinds and the reuse mask as compile-time constants for the layouts that occur there: every role a unit-step vector, and
// for two-role programs one value per folded row against a vector (gevm) or one role re-read by every kept row (gemv). Any other
// layout, and the other combiners (prod, amin, amax), run the same bodies with the kinds read at run time.

template <int ID> constexpr uint64_t all_vec_kinds() { uint64_t k = 0; for (int r=0; r<SPT<ID>::P.nin; ++r) k |= uint64_t(SK_VEC)<<(4*r); return k; }

// The kernel a plan runs on (a __global__ function pointer), or null.
template <int ID> static void const *
select_kernel(SParams const & p)
{
    constexpr SProg P = SPT<ID>::P;
#define RACU_GO(KERNEL, ...) return reinterpret_cast<void const *>(&KERNEL<ID, __VA_ARGS__>)
    if (p.mode==K_MAP) {
        if constexpr (!P.fold) {
            if constexpr (4==P.nin && 0==P.yrole) { // B += A*C + v with v matched on the rows (examples/read-me.cc:32): kinds known at build time
                uint64_t kinds = 0;
                for (int k=0; k<P.nin; ++k) kinds |= uint64_t(p.kind[k])<<(4*k);
                if (!p.lanes && kinds==sk_pack(SK_VEC, SK_VEC, SK_VEC, SK_ROW)) RACU_GO(smap_kernel, true, sk_pack(SK_VEC, SK_VEC, SK_VEC, SK_ROW));
            }
            if (p.lanes) RACU_GO(smap_lanes_kernel, SK_RUNTIME);
            if (p.general) RACU_GO(smap_kernel, true); else RACU_GO(smap_kernel, false);
        }
    } else if constexpr (P.fold) {
        uint64_t kinds = 0; int rmask = 0;
        for (int k=0; k<P.nin; ++k) { kinds |= uint64_t(p.kind[k])<<(4*k); rmask |= (p.reused[k] ? 1 : 0)<<k; }
        constexpr uint64_t VV = all_vec_kinds<ID>();
        bool const sum = p.fold_op==RACU_ASSIGN_ADD || p.fold_op==RACU_ASSIGN_SUB;
        if (p.mode==K_FOLD_INNER) {
            if (sum && kinds==VV) { // dot, sum, sqrm, sum-cols, gemv
                if (p.inner_rows) {
                    if (0==rmask) RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_ADD, 0, VV);
                    if constexpr (2==P.nin) { if (1==rmask) RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_ADD, 1, VV); if (2==rmask) RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_ADD, 2, VV); }
                } else {
                    if (0==rmask) RACU_GO(sfold_inner_kernel, RACU_ASSIGN_ADD, 0, VV);
                    if constexpr (2==P.nin) { if (1==rmask) RACU_GO(sfold_inner_kernel, RACU_ASSIGN_ADD, 1, VV); if (2==rmask) RACU_GO(sfold_inner_kernel, RACU_ASSIGN_ADD, 2, VV); }
                }
            }
            constexpr int ALL = (1<<P.nin)-1; // run-time kinds: loads of a kernel with any re-read role all go through L1
            if (p.inner_rows) {
                switch (p.fold_op) {
                case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: if (rmask) RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_ADD, ALL, SK_RUNTIME); else RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_ADD, 0, SK_RUNTIME);
                case RACU_ASSIGN_AMIN: RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_AMIN, 0, SK_RUNTIME);
                case RACU_ASSIGN_AMAX: RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_AMAX, 0, SK_RUNTIME);
                default: RACU_GO(sfold_inner_rows_kernel, RACU_ASSIGN_MUL, 0, SK_RUNTIME);
                }
            }
            switch (p.fold_op) {
            case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: if (rmask) RACU_GO(sfold_inner_kernel, RACU_ASSIGN_ADD, ALL, SK_RUNTIME); else RACU_GO(sfold_inner_kernel, RACU_ASSIGN_ADD, 0, SK_RUNTIME);
            case RACU_ASSIGN_AMIN: RACU_GO(sfold_inner_kernel, RACU_ASSIGN_AMIN, 0, SK_RUNTIME);
            case RACU_ASSIGN_AMAX: RACU_GO(sfold_inner_kernel, RACU_ASSIGN_AMAX, 0, SK_RUNTIME);
            default: RACU_GO(sfold_inner_kernel, RACU_ASSIGN_MUL, 0, SK_RUNTIME);
            }
        } else {
            if (sum) {
                if (kinds==VV) RACU_GO(sfold_outer_kernel, RACU_ASSIGN_ADD, VV); // sum-rows
                if constexpr (2==P.nin) {
                    if (kinds==sk_pack(SK_ROW, SK_VEC)) RACU_GO(sfold_outer_kernel, RACU_ASSIGN_ADD, sk_pack(SK_ROW, SK_VEC)); // gevm
                    if (kinds==sk_pack(SK_VEC, SK_ROW)) RACU_GO(sfold_outer_kernel, RACU_ASSIGN_ADD, sk_pack(SK_VEC, SK_ROW));
                }
            }
            switch (p.fold_op) {
            case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: RACU_GO(sfold_outer_kernel, RACU_ASSIGN_ADD, SK_RUNTIME);
            case RACU_ASSIGN_AMIN: RACU_GO(sfold_outer_kernel, RACU_ASSIGN_AMIN, SK_RUNTIME);
            case RACU_ASSIGN_AMAX: RACU_GO(sfold_outer_kernel, RACU_ASSIGN_AMAX, SK_RUNTIME);
            default: RACU_GO(sfold_outer_kernel, RACU_ASSIGN_MUL, SK_RUNTIME);
            }
        }
    }
#undef RACU_GO
    return nullptr;
}

// The table is compiled in SP_NPART translation units (ids congruent to SP_PART), in parallel: build.py.
#ifndef SP_NPART
#define SP_NPART 1
#define SP_PART 0
#endif

template <int ID=SP_PART> static void const *
kernel_of(SParams const & p)
{
    if constexpr (ID<SP_COUNT) {
        if (p.id==ID) return select_kernel<ID>(p);
        return kernel_of<ID+SP_NPART>(p);
    } else {
        return nullptr;
    }
}

#define RACU_CAT_(a, b) a##b
#define RACU_CAT(a, b) RACU_CAT_(a, b)
void const * RACU_CAT(static_kernel_part, SP_PART)(SParams const & p) { return kernel_of<>(p); }

#if SP_PART==0
void const * static_kernel_part1(SParams const & p);
void const * static_kernel_part2(SParams const & p);
void const * static_kernel_part3(SParams const & p);

static void const *
static_kernel(SParams const & p)
{
    switch (p.id % SP_NPART) {
    case 0: return static_kernel_part0(p);
#if SP_NPART>1
    case 1: return static_kernel_part1(p);
#endif
#if SP_NPART>2
    case 2: return static_kernel_part2(p);
#endif
#if SP_NPART>3
    case 3: return static_kernel_part3(p);
#endif
    default: return nullptr;
    }
}

cudaError_t
launch_static(Plan const & plan, cudaStream_t stream)
{
    if (SP_JIT==plan.sp.id) return launch_jit(plan, stream);
    void const * fn = static_kernel(plan.sp);
    if (!fn) return cudaErrorInvalidValue;
    void * args[2] = {const_cast<SParams *>(&plan.sp), const_cast<FinishParams *>(&plan.fp)};
    return cudaLaunchKernel(fn, dim3(plan.grid_x, plan.grid_y), dim3(KBLOCK), args, 0, stream);
}

// CTAs of the plan's kernel that fit on one SM (0: unknown - no device). The planner sizes fold grids in whole waves with it.
int
static_occupancy(Plan const & plan)
{
    if (SP_JIT==plan.sp.id) return jit_occupancy(plan);
    void const * fn = static_kernel(plan.sp);
    if (!fn) return 0;
    static std::mutex mu;
    static std::map<void const *, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(fn);
    if (it!=cache.end()) return it->second;
    int n = 0;
    if (cudaSuccess!=cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, KBLOCK, 0)) { cudaGetLastError(); return 0; } // (not cached: no device yet)
    cache.emplace(fn, n);
    return n;
}

static struct OccInstall { OccInstall() { g_static_occupancy = &static_occupancy; } } occ_install;
#endif

} // namespace racu
