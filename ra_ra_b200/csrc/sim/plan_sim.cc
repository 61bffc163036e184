// plan_sim.cc - CPU simulator of the planner + kernel bodies. TEST INFRASTRUCTURE ONLY (built as
// ra_ra_b200/csrc/sim/libracu_sim.so by tests; never linked into libracu.so, never used by the product path).
// It runs make_plan() and then walks the kernel's thread mapping sequentially with the SAME stage/interp/store/finish
// code the device runs (kbody.h), so the planner (axis sorting, merging, flips, staging modes, stack levels, chunking)
// can be checked against the oracle in this GPU-less container. Folds combine in a different (sequential) order than the
// device tree, so float folds are compared by tolerance; integer-valued ones exactly.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../kbody.h"
#include "../plan.h"

namespace racu {

struct HostCopy
{
    void copy16(void * s, void const * g) const { std::memcpy(s, g, 16); }
    void copy8(void * s, void const * g) const { std::memcpy(s, g, 8); }
    void copy4(void * s, void const * g) const { std::memcpy(s, g, 4); }
};

template <class W> void
sim_run(Plan & plan)
{
    constexpr int V = Lanes<W>::V;
    KParams & p = plan.kp;
    FinishParams & f = plan.fp;
    std::vector<W> partial(plan.partial_bytes/sizeof(W)+1);
    p.partial = partial.data(); f.partial = partial.data(); f.identity = p.identity;
    alignas(16) unsigned char stage[4*RACU_MAX_OPND*KSLOT];
    alignas(16) unsigned char stack[RACU_MAX_INS*KSLOT];
    int const ss = KSLOT; // one thread: slots are contiguous
    KIdx a, b;
    std::memset(&a, 0, sizeof(a)); std::memset(&b, 0, sizeof(b));
    if (p.mode==K_MAP) {
        for (int64_t g=0; g<p.nvecA; ++g) {
            decompA(p, g, a);
            { int nv = int(std::min<int64_t>(V, p.len0-a.e0)); stage_group<W, 1, 1, 1>(p, &a, &b, &nv, stage, ss, HostCopy {}); }
            W t[1][V];
            interp<W, 1>(p, stage, stack, ss, t);
            if (p.scatter) scatter_vector<W>(p, reinterpret_cast<W const *>(stack), t[0], int(std::min<int64_t>(V, p.len0-a.e0)));
            else store_vector<W>(p, a, t[0]);
        }
    } else if (p.mode==K_FOLD_INNER) {
        for (int64_t kb=0; kb<p.nk; ++kb) {
            if (p.rankB>0) decompB(p, kb, b);
            for (int c=0; c<p.nchunks; ++c) {
                int64_t start = int64_t(c)*p.chunk_len, end = std::min(p.nvecA, start+p.chunk_len);
                W r = W(p.identity);
                for (int64_t g=start; g<end; ++g) {
                    decompA(p, g, a);
                    { int nv = int(std::min<int64_t>(V, p.len0-a.e0)); stage_group<W, 1, 1, 1>(p, &a, &b, &nv, stage, ss, HostCopy {}); }
                    W t[1][V];
                    interp<W, 1>(p, stage, stack, ss, t);
                    for (int j=0; j<V; ++j) { if (j<p.len0-a.e0) r = combine<W>(p.fold_op, p.acc_type, r, t[0][j]); }
                }
                if (p.nchunks>1) partial[int64_t(c)*p.nk+kb] = r; else finish_apply<W>(f, kb, r);
            }
        }
    } else {
        for (int64_t g=0; g<p.nvecA; ++g) {
            decompA(p, g, a);
            int64_t row = p.rankA>1 ? fastdiv(uint32_t(g), p.vprdiv) : 0;
            int64_t n0 = row*p.len0+a.e0;
            for (int c=0; c<p.nchunks; ++c) {
                int64_t r0 = int64_t(c)*p.chunk_len, r1 = std::min(p.nB, r0+p.chunk_len);
                W acc[V];
                for (int j=0; j<V; ++j) acc[j] = W(p.identity);
                for (int64_t r=r0; r<r1; ++r) {
                    decompB(p, r, b);
                    { int nv = int(std::min<int64_t>(V, p.len0-a.e0)); stage_group<W, 1, 1, 1>(p, &a, &b, &nv, stage, ss, HostCopy {}); }
                    W t[1][V];
                    interp<W, 1>(p, stage, stack, ss, t);
                    group_combine<W, 1>(p.fold_op, p.acc_type, acc, t, [](int, int) { return true; });
                }
                for (int j=0; j<V; ++j) {
                    if (j<p.len0-a.e0) { if (p.nchunks>1) partial[int64_t(c)*p.nk+n0+j] = acc[j]; else finish_apply<W>(f, n0+j, acc[j]); }
                }
            }
        }
    }
    if (plan.need_finish) {
        for (int64_t n=0; n<f.nk; ++n) {
            W r = f.per_warp ? W(f.identity) : partial[n];
            for (int c=(f.per_warp ? 0 : 1); c<f.nchunks; ++c) r = combine<W>(f.fold_op, f.acc_type, r, partial[int64_t(c)*f.nk+n]);
            finish_apply<W>(f, n, r);
        }
    }
}

thread_local std::string g_sim_err;
thread_local Plan g_sim_plan;

} // namespace racu

using namespace racu;

extern "C" {

const char * racu_sim_last_error(void) { return g_sim_err.c_str(); }

// Runs d on HOST memory through make_plan + the kernel bodies. reduce_out (host) non-null: whole-array reduction.
int
racu_sim_run(racu_desc const * d, int redop, void * reduce_out)
{
    static int const once = setenv("RACU_NO_STATIC", "1", 1); // the simulator walks the generic kernels' bodies only
    (void)once;
    Plan & plan = g_sim_plan;
    plan = Plan {};
    if (int e = make_plan(d, redop, reduce_out, plan, g_sim_err)) {
        if (e==RACU_ERR_SEQ) { // as api.cc run_plan: the folded axes one index at a time
            Shape sh;
            if (int e2 = agree(d, sh, g_sim_err)) return e2;
            int const k = first_folded_axis(d, sh);
            if (k<0) return RACU_ERR_UNSUPPORTED;
            racu_desc sub;
            for (int64_t i=0; i<sh.len[k]; ++i) {
                peel_axis(*d, k, i, sub);
                if (int e2 = racu_sim_run(&sub, 0, nullptr)) return e2;
            }
            return RACU_OK;
        }
        if (e!=RACU_ERR_PEEL) return e;
        if (reduce_out) { // as api.cc run_plan: `out OP= expr` onto a rank-0 destination filled with the identity
            racu_desc d2;
            int t; uint64_t bits;
            if (!peel_reduction(*d, redop, reduce_out, d2, t, bits)) return RACU_ERR_UNSUPPORTED;
            std::memcpy(reduce_out, &bits, dtype_size(t));
            return racu_sim_run(&d2, 0, nullptr);
        }
        if (!can_peel(d, reduce_out!=nullptr)) return RACU_ERR_UNSUPPORTED;
        Shape sh;
        if (int e2 = agree(d, sh, g_sim_err)) return e2;
        racu_desc sub;
        for (int64_t i = peel_last_only(d) ? sh.len[0]-1 : 0; i<sh.len[0]; ++i) {
            peel_axis0(*d, i, sub);
            if (int e2 = racu_sim_run(&sub, redop, reduce_out)) return e2;
        }
        return RACU_OK;
    }
    if (plan.empty) {
        if (plan.fill_identity) std::memcpy(reduce_out, &plan.kp.identity, dtype_size(plan.kp.acc_type));
        return RACU_OK;
    }
    if (plan.kp.wide) sim_run<uint64_t>(plan); else sim_run<uint32_t>(plan);
    return RACU_OK;
}

// Introspection of the last plan, for tests of the planner's decisions.
int
racu_sim_plan_info(int32_t * out /* 16 ints */, int64_t * out64 /* 8 */)
{
    Plan const & p = g_sim_plan;
    out[0] = p.kp.mode; out[1] = p.kp.wide; out[2] = p.kp.rankA; out[3] = p.kp.rankB; out[4] = p.kp.nstage; out[5] = p.kp.depth;
    out[6] = p.kp.U; out[7] = p.kp.nchunks; out[8] = int32_t(p.grid_x); out[9] = int32_t(p.grid_y); out[10] = int32_t(p.smem);
    out[11] = p.need_finish; out[12] = p.kp.nins; out[13] = p.empty;
    int nvec = 0; for (int o=0; o<p.kp.nopnd; ++o) nvec += p.kp.opnd[o].stage==S_VEC;
    out[14] = nvec; out[15] = p.kp.nopnd;
    out64[0] = p.kp.len0; out64[1] = p.kp.nvecA; out64[2] = p.kp.nB; out64[3] = p.kp.nk; out64[4] = p.kp.chunk_len;
    return RACU_OK;
}

} // extern "C"
