// kbody.h - the per-vector bodies of the ply kernels: operand staging, destination store, fold finish.
// Shared by ply_kernels.cu (device; BE = cp.async) and sim/plan_sim.cc (host, tests only; BE = memcpy).
#pragma once
#include "kcore.h"

namespace racu {

// Stage every used operand of the U vectors a thread evaluates together. Vector u sits at group-A position a[u*(UA>1)]
// and group-B position b[u*(UB>1)] (maps: UB = 0; fold-inner: one shared b; fold-outer: one shared a). nvalid[u] is the
// number of live lanes of vector u (0: skip it). Slot s of vector u lives at stage + (u*nstage + s)*sstride.
// Each operand is decoded once for all U vectors.
template <class W, int U, int UA, int UB, class BE> RACU_HD void
stage_group(KParams const & p, KIdx const * a, KIdx const * b, int const * nvalid, unsigned char * stage, int sstride, BE be)
{
    constexpr int V = Lanes<W>::V;
    int const ustride = p.nstage*sstride;
    for (int o=0; o<p.nopnd; ++o) {
        KOpnd const & k = p.opnd[o];
        if (S_NONE==k.stage || 255==k.slot) continue;
        unsigned char * const sl0 = stage + int(k.slot)*sstride;
        unsigned char const * const base = reinterpret_cast<unsigned char const *>(k.base);
        int const es = k.esize;
        int64_t offb = 0;
        if constexpr (1==UB) offb = offsetB(p, k, b[0]);
        if (S_VEC==k.stage && 4==es) { // the common case first: 16-byte (V=4) or 8-byte (V=2) vectors of 4-byte elements
#pragma unroll
            for (int u=0; u<U; ++u) {
                if (nvalid[u]>0) {
                    int64_t off = offsetA(p, k, a[UA>1 ? u : 0]) + offb;
                    if constexpr (UB>1) off += offsetB(p, k, b[u]);
                    unsigned char * sl = sl0 + u*ustride;
                    unsigned char const * src = base + off*4;
                    if (nvalid[u]==V) { if constexpr (4==V) be.copy16(sl, src); else be.copy8(sl, src); }
                    else {
#pragma unroll
                        for (int j=0; j<V; ++j) { if (j<nvalid[u]) be.copy4(sl+4*j, src+4*j); else *reinterpret_cast<uint32_t *>(sl+4*j) = 0; }
                    }
                }
            }
            continue;
        }
#pragma unroll
        for (int u=0; u<U; ++u) {
            if (nvalid[u]<=0) continue;
            int64_t off = offsetA(p, k, a[UA>1 ? u : 0]) + offb;
            if constexpr (UB>1) off += offsetB(p, k, b[u]);
            unsigned char * sl = sl0 + u*ustride;
            if (S_VEC==k.stage && nvalid[u]==V) {
                unsigned char const * src = base + off*es;
                if (8==es) be.copy16(sl, src);                                                            // V = 2
                else if (4==V) *reinterpret_cast<uint32_t *>(sl) = *reinterpret_cast<uint32_t const *>(src); // 4 bools
                else *reinterpret_cast<uint16_t *>(sl) = *reinterpret_cast<uint16_t const *>(src);          // 2 bools
            } else if (S_VEC==k.stage || S_GATHER==k.stage) {
                int64_t const s0 = k.sA[0];
#pragma unroll
                for (int j=0; j<V; ++j) {
                    if (j<nvalid[u]) {
                        unsigned char const * src = base + (off+j*s0)*es;
                        if (4==es) be.copy4(sl+4*j, src); else if (8==es) be.copy8(sl+8*j, src); else sl[j] = *src;
                    } else {
                        if (4==es) *reinterpret_cast<uint32_t *>(sl+4*j) = 0; else if (8==es) *reinterpret_cast<uint64_t *>(sl+8*j) = 0; else sl[j] = 0;
                    }
                }
            } else if (S_BCAST==k.stage) {
                W v = load_elem<W>(base + off*es, k.dtype);
#pragma unroll
                for (int j=0; j<V; ++j) store_elem<W>(sl+j*es, k.dtype, v);
            } else { // S_SEQ
                int64_t const s0 = k.sA[0];
#pragma unroll
                for (int j=0; j<V; ++j) store_elem<W>(sl+j*es, k.dtype, seq_value<W>(k, off+j*s0));
            }
        }
    }
}

// Store the V result lanes of a map to the destination.
template <class W> RACU_HD void
store_vector(KParams const & p, KIdx const & a, W const * t)
{
    constexpr int V = Lanes<W>::V;
    KOpnd const & k = p.opnd[p.dst];
    int64_t const rem = p.len0-a.e0;
    int const nvalid = rem<V ? int(rem) : V;
    int const es = k.esize;
    int64_t const off = offsetA(p, k, a);
    unsigned char * dst = reinterpret_cast<unsigned char *>(k.base) + off*es;
    if (S_VEC==k.stage && nvalid==V) {
        if constexpr (sizeof(W)==4) {
            if (4==es) { uint4 q; q.x = t[0]; q.y = t[1]; q.z = t[2]; q.w = t[3]; *reinterpret_cast<uint4 *>(dst) = q; }
            else { *reinterpret_cast<uint32_t *>(dst) = (t[0]&1u) | ((t[1]&1u)<<8) | ((t[2]&1u)<<16) | ((t[3]&1u)<<24); }
        } else {
            if (8==es) { ulonglong2 q; q.x = t[0]; q.y = t[1]; *reinterpret_cast<ulonglong2 *>(dst) = q; }
            else if (4==es) { uint2 q; q.x = uint32_t(t[0]); q.y = uint32_t(t[1]); *reinterpret_cast<uint2 *>(dst) = q; }
            else { *reinterpret_cast<uint16_t *>(dst) = uint16_t((uint32_t(t[0])&1u) | ((uint32_t(t[1])&1u)<<8)); }
        }
    } else {
        int64_t const s0 = k.sA[0];
#pragma unroll
        for (int j=0; j<V; ++j) { if (j<nvalid) store_elem<W>(dst+j*s0*es, k.dtype, t[j]); }
    }
}

// Scatter: a[off] OP= x for the live lanes of one vector (a(i ...) OP= x with index arrays, ra/ply.hh:461-499 as an lvalue).
// Offsets may repeat: on the device the update is an atomic compare-and-swap loop on the element's own 4 or 8 bytes, computing
// exactly what the sequential reference computes for one element - old -> value type, OP, -> element type - in an unspecified
// order; `=` leaves any one of the values. Offsets are clamped into the array like RACU_OP_LOAD.
template <class W> RACU_HD void
scatter_vector(KParams const & p, W const * off, W const * x, int nvalid)
{
    constexpr int V = Lanes<W>::V;
    KOpnd const & k = p.opnd[p.dst];
    int const dt = k.dtype, vt = p.acc_type, es = k.esize;
#pragma unroll
    for (int j=0; j<V; ++j) {
        if (j>=nvalid) continue;
        int64_t o = int64_t(off[j]);
        o = o<k.sA[0] ? k.sA[0] : o>k.sA[1] ? k.sA[1] : o;
        unsigned char * const ptr = reinterpret_cast<unsigned char *>(k.base) + o*int64_t(es);
        if (RACU_ASSIGN_SET==p.fold_op || 1==es) { store_elem<W>(ptr, dt, cast_any<W>(vt, dt, x[j])); continue; }
        auto rmw = [&](W old) { return cast_any<W>(vt, dt, finish_op<W>(p.fold_op, vt, cast_any<W>(dt, vt, old), x[j])); };
#if defined(__CUDA_ARCH__)
        // a(i) += x / -= x in the array's own type: ONE reduction instruction at the L2 (red.global.add) instead of a read + compare-and-swap
        // round trip per attempt. The same single IEEE addition per contribution, in the same unspecified order.
        if ((RACU_ASSIGN_ADD==p.fold_op || RACU_ASSIGN_SUB==p.fold_op) && dt==vt && RACU_BOOL!=dt) {
            bool const sub = RACU_ASSIGN_SUB==p.fold_op;
            if (RACU_F32==dt) { float const v = bits_f32(uint32_t(x[j])); atomicAdd(reinterpret_cast<float *>(ptr), sub ? -v : v); }
            else if (RACU_I32==dt) { uint32_t const v = uint32_t(x[j]); atomicAdd(reinterpret_cast<unsigned int *>(ptr), sub ? 0u-v : v); }
            else if constexpr (sizeof(W)==8) {
                if (RACU_F64==dt) { double const v = bits_f64(uint64_t(x[j])); atomicAdd(reinterpret_cast<double *>(ptr), sub ? -v : v); }
                else { uint64_t const v = uint64_t(x[j]); atomicAdd(reinterpret_cast<unsigned long long *>(ptr), (unsigned long long)(sub ? 0ull-v : v)); }
            }
            continue;
        }
        if (4==es) {
            unsigned int * const w = reinterpret_cast<unsigned int *>(ptr);
            unsigned int old = *w, assumed;
            do { assumed = old; old = atomicCAS(w, assumed, (unsigned int)(rmw(W(assumed)))); } while (old!=assumed);
        } else {
            if constexpr (sizeof(W)==8) {
                unsigned long long * const w = reinterpret_cast<unsigned long long *>(ptr);
                unsigned long long old = *w, assumed;
                do { assumed = old; old = atomicCAS(w, assumed, (unsigned long long)(rmw(W(assumed)))); } while (old!=assumed);
            }
        }
#else
        store_elem<W>(ptr, dt, rmw(load_elem<W>(ptr, dt)));
#endif
    }
}

// dst[kept n] = overwrite ? total : dst[kept n] OP total, with the accumulator <-> destination conversions.
template <class W> RACU_HD void
finish_apply(FinishParams const & f, int64_t n, W total)
{
    int64_t off = 0;
    if (1==f.rank) {
        off = n*f.stride[0];
    } else if (f.rank>1) {
        uint32_t m = uint32_t(n), q;
        for (int j=0; j<f.rank; ++j) {
            if (j+1<f.rank) { q = fastdiv(m, f.d[j]); off += int64_t(m-q*f.d[j].len)*f.stride[j]; m = q; }
            else { off += int64_t(m)*f.stride[j]; }
        }
    }
    unsigned char * dst = static_cast<unsigned char *>(f.dst) + off*int64_t(f.dst_dtype==RACU_BOOL ? 1 : (f.dst_dtype==RACU_I32 || f.dst_dtype==RACU_F32) ? 4 : 8);
    W val = total;
    if (!f.overwrite) {
        W old = cast_any<W>(f.dst_dtype, f.acc_type, load_elem<W>(dst, f.dst_dtype));
        val = finish_op<W>(f.fold_op, f.acc_type, old, total);
    }
    store_elem<W>(dst, f.dst_dtype, cast_any<W>(f.acc_type, f.dst_dtype, val));
}

} // namespace racu
