// stencil.cu - 5- and 7-point Laplace stencils with TMA-staged shared-memory tiles (sm_100a).
//
// Reference forms (exact association order is kept, no contraction, so results are bit-identical to the CPU ply):
//   3D7   Anext(I,J,K) = -6*A(I,J,K) + A(I+1,J,K) + A(I,J+1,K) + A(I,J,K+1) + A(I-1,J,K) + A(I,J-1,K) + A(I,J,K-1)   bench/bench-stencil3.cc:59-61
//   2D5   Anext(I,J)   = -4*A(I,J) + A(I+1,J) + A(I,J+1) + A(I-1,J) + A(I,J-1)                                         bench/bench-stencil2.cc:53-55
//   STIFF w(i,j)       =  4*v(i,j) - v(i-1,j) - v(i,j-1) - v(i+1,j) - v(i,j+1)                                         examples/laplace2d.cc:36
//   MASS  w(i,j)       =  sqrm(h) * (v(i,j)/2. + (v(i-1,j) + v(i,j-1) + v(i+1,j) + v(i,j+1) + v(i+1,j+1) + v(i-1,j-1))/12.)    examples/laplace2d.cc:48
//         (reached from racu_map only: the three scalars are operands of the expression; for float arrays the six are summed in float and
//          the rest is double arithmetic, as C++ promotes it)
//   1D3   Anext(I)     = -2*A(I) + A(I+1) + A(I-1)                                                                          bench/bench-stencil1.cc:45
// Boundary cells of the destination are never written (bench/bench-stencil3.cc:125-126).
//
// 2.5-D streaming: a CTA owns a (TJ x TK) tile of the two inner axes and marches along axis 0. Every plane tile
// (TJ+2) x (TK+2*HO), halo included, is brought in ONCE by a TMA tensor copy (cp.async.bulk.tensor, zero fill outside the
// array) into a ring of shared-memory stages signalled by mbarriers; the i-1 / i / i+1 centre values of a thread's
// column live in registers, the j and k neighbours come from the current stage. Each input cell is therefore read from
// L2/HBM once per CTA (plus halo), each output written once: 8 B / cell algorithmic for f32.
// 2-D stencils are the same kernel with a single plane (no marching).
//
// Slab-sharded sweeps (one process per GPU, racu_stencil_push): the FUSED instantiation of the same kernel also stores the
// first / last owned plane it produces into the neighbour's ghost plane (peer memory mapped with CUDA IPC, plain 16-byte
// stores that travel over NVLink) and keeps the two neighbours in lock step with one flag word per side, so a step is ONE
// kernel launch with no separate exchange. See the protocol comment at racu_stencil_push.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstring>
#include <type_traits>
#include <mutex>

#include "kcore.h"
#include "launch.h"
#include "special.h"

namespace racu {

constexpr int ST_KIND_MASS = 100; // internal: the 2-D 7-point mass stencil (no public racu_stencil_kind: it carries three scalars)
constexpr int ST_TR = 8;        // thread rows per tile
constexpr int ST_RPT = 2;       // rows per thread (rows tj and tj+ST_TR): halves the j halo and the per-plane fixed costs
constexpr int ST_TJ = ST_TR*ST_RPT; // tile rows
constexpr int ST_THREADS = 256;
constexpr int ST_STAGES = 4;

template <class T> struct StTile
{
    static constexpr int VT = 16/sizeof(T);         // outputs per thread: one 16-byte vector along k
    static constexpr int TK = (ST_THREADS/ST_TR)*VT; // 128 (f32) / 64 (f64)
    static constexpr int HO = VT;                   // halo offset along k: keeps interior vectors 16-byte aligned in smem
    static constexpr int TKB = TK+2*HO;             // box extent along k
    static constexpr int PLANE = (((ST_TJ+2)*TKB*int(sizeof(T))+127)/128*128)/int(sizeof(T)); // elements per stage, 128-byte multiple (TMA destination alignment)
};

struct StParams
{
    void * dst;
    int64_t ds0, ds1, ds2;   // destination strides (elements)
    int32_t n0, n1, n2;      // full array extents (marching axis first). 2-D: n0 = 1
    int32_t i_lo, i_hi;      // output planes [i_lo, i_hi) (3-D); ignored for 2-D
    int32_t ci;              // planes per CTA chunk
    int32_t gz;              // chunks along the marching axis; gridDim.z = gz * bands
    int32_t gyb;             // tile rows per band (gridDim.y)
    int32_t vec_store;       // destination allows aligned 16-byte stores
    // FUSED only (racu_stencil_push)
    void * push_lo, * push_hi;           // neighbour's ghost plane (element [.,0,0]) mirroring my plane 1 / plane n0-2, or null
    uint32_t * flags;                    // local: [0] lower neighbour's completed steps, [1] higher's, [2]/[3] CTA counters, [4] timeout
    uint32_t * signal_lo, * signal_hi;   // the neighbours' flag words this rank sets (peer memory)
    uint32_t step;                       // steps completed before this launch
    int32_t edge_ctas;                   // CTAs per edge chunk (gridDim.x*gridDim.y)
    int32_t dbg;                         // measurement knobs (RACU_PUSH_DEBUG): 1 = no remote stores, 2 = no waits, 4 = no fence / signal. Wrong results.
    double c0, c1, c2;                   // MASS only: sqrm(h), the divisor of the centre, the divisor of the six
};

// ---- system-scope flag helpers for the fused halo push
__device__ __forceinline__ uint32_t ld_acquire_sys(uint32_t const * p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(uint32_t * p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint64_t globaltimer_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
constexpr uint64_t ST_WAIT_TIMEOUT_NS = 20ull*1000*1000*1000; // a neighbour that never arrives sets flags[4] instead of hanging the GPU

__device__ __forceinline__ void
wait_neighbour(uint32_t * flags, int side, uint32_t want)
{
    if (int32_t(ld_acquire_sys(flags+side)-want)>=0) return;
    uint64_t const t0 = globaltimer_ns();
    while (int32_t(ld_acquire_sys(flags+side)-want)<0) {
        __nanosleep(64);
        if (globaltimer_ns()-t0>ST_WAIT_TIMEOUT_NS) { atomicExch(flags+4, 1u+side); break; }
    }
}

__device__ __forceinline__ uint32_t smem_u32(void const * p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITA_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAITA_DONE;\n"
        "bra WAITA_LOOP;\n"
        "WAITA_DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void * smem, CUtensorMap const * tmap, int c0, int c1, int c2, uint64_t * bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

template <class T, int KIND> __device__ __forceinline__ T
stencil_value(T c, T ip, T jp, T kp, T im, T jm, T km)
{
    if constexpr (KIND==RACU_STENCIL_3D7) return ((((((T(-6)*c) + ip) + jp) + kp) + im) + jm) + km;
    else if constexpr (KIND==RACU_STENCIL_2D5) return ((((T(-4)*c) + jp) + kp) + jm) + km;            // (i,j) of the 2-D form are (j,k) here
    else return ((((T(4)*c) - jm) - km) - jp) - kp;                                                     // STIFF
}

template <class T, int KIND, bool FUSED> __global__ void __launch_bounds__(ST_THREADS)
stencil_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StParams p)
{
    using TL = StTile<T>;
    constexpr int VT = TL::VT, TK = TL::TK, HO = TL::HO, TKB = TL::TKB, PLANE = TL::PLANE;
    constexpr bool is3d = KIND==RACU_STENCIL_3D7;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T * const stages = reinterpret_cast<T *>(smem_raw);
    __shared__ __align__(8) uint64_t full[ST_STAGES];

    int const tid = threadIdx.x;
    int const tk = tid % (TK/VT), tj = tid / (TK/VT);          // tj in [0, ST_TR)
    // Bands: blockIdx.z = band*gz + chunk. A band (all of k, gyb tile rows of j) is marched through completely before the next
    // one starts, so that the CTAs resident together - consecutive chunks of one band - find each other's halo planes in L2
    // whatever the plane size (2048^2 planes: a whole plane's worth of tiles per chunk is 4x what the GPU holds at once).
    int const band = is3d ? int(blockIdx.z)/p.gz : 0;
    // 2-D: a CTA takes p.ci consecutive tile rows of its column strip, one after the other through the same TMA ring (p.gyb = tile rows in all)
    int const k0 = blockIdx.x*TK, j0 = 1 + (is3d ? band*p.gyb + int(blockIdx.y) : int(blockIdx.y)*p.ci)*ST_TJ; // first output k of the tile (vector aligned), first output j
    if (j0>=p.n1-1) return;                                    // last band may be short (whole CTA leaves before any barrier)
    int const k = k0 + tk*VT, j = j0 + tj;
    // FUSED: the two edge chunks (0 and gz-1: they read a ghost plane and feed the neighbours) are SPREAD over the grid: tile q of a band
    // runs its two edge chunks at grid positions z = 2q, 2q+1 (mod the first three quarters of the band) and its interior chunks, in
    // order, at the others. Every slab of the grid then carries a few percent of edge CTAs - their flag waits, remote stores and
    // fences hide behind the interior CTAs resident with them - instead of one wave made of nothing else (round 1: the two edge chunks
    // of all tiles ran together in mid-grid and cost 87 % -> the rest of the strong-scaling loss at 8 GPUs). Nothing is scheduled in
    // the last quarter, so the pushed planes have landed and the neighbour's flag is up long before the next step needs them.
    int chunk = is3d ? int(blockIdx.z) - band*p.gz : 0;
    if constexpr (FUSED) {
        int const gz = p.gz;
        if (gz>=4) {
            int const zmax = max(2, (3*gz/4)&~1), q = int(blockIdx.y)*int(gridDim.x) + int(blockIdx.x);
            int const zlo = (2*q)%zmax, zhi = zlo+1, z = chunk;
            chunk = z==zlo ? 0 : z==zhi ? gz-1 : 1 + z - (z>zlo) - (z>zhi);
        } else if (gz>2) {
            int const e = (gz-2)/2;
            chunk = chunk<e ? chunk+1 : (chunk==e ? 0 : (chunk==e+1 ? gz-1 : chunk-1));
        }
    }
    int const ilo = is3d ? p.i_lo + chunk*p.ci : 0;
    int const ihi = is3d ? min(p.i_hi, ilo + p.ci) : 1;
    bool const touch_lo = FUSED && p.push_lo && 1==ilo, touch_hi = FUSED && p.push_hi && ihi==p.n0-1;
    int const nplanes = is3d ? (ihi-ilo+2) : min(p.ci, p.gyb - int(blockIdx.y)*p.ci); // input planes ilo-1 .. ihi; 2-D: tiles of this CTA
    constexpr uint32_t plane_bytes = (ST_TJ+2)*TKB*sizeof(T); // bytes one TMA box delivers

    if (0==tid) {
#pragma unroll
        for (int s=0; s<ST_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if constexpr (FUSED) {
            // The neighbour has finished step p.step-1: its pushes into my ghost plane of the source have landed, and it no
            // longer reads the ghost plane of ITS destination that this CTA is about to overwrite.
            if (touch_lo && !(p.dbg&2)) wait_neighbour(p.flags, 0, p.step);
            if (touch_hi && !(p.dbg&2)) wait_neighbour(p.flags, 1, p.step);
            if (touch_lo || touch_hi) asm volatile("fence.proxy.async;" ::: "memory"); // remote generic-proxy writes -> my TMA reads
        }
    }
    __syncthreads();
    auto issue = [&](int q) { // input plane q of this chunk -> stage q % STAGES
        int s = q % ST_STAGES;
        mbar_expect_tx(&full[s], plane_bytes);
        tma_load_3d(stages + size_t(s)*PLANE, &tmap, k0-HO, is3d ? j0-1 : j0-1+q*ST_TJ, is3d ? ilo-1+q : 0, &full[s]);
    };
    if (0==tid) { for (int q=0; q<ST_STAGES && q<nplanes; ++q) issue(q); }

    // lanes of this thread's vectors that are interior (row r of this thread is j + r*ST_TR)
    bool lane_ok[ST_RPT][VT], all_ok[ST_RPT];
#pragma unroll
    for (int r=0; r<ST_RPT; ++r) {
        bool const row_ok = j+r*ST_TR<p.n1-1;
        all_ok[r] = row_ok;
#pragma unroll
        for (int l=0; l<VT; ++l) { lane_ok[r][l] = row_ok && (k+l>=1) && (k+l<p.n2-1); all_ok[r] = all_ok[r] && lane_ok[r][l]; }
    }
    int const so = (tj+1)*TKB + HO + tk*VT; // this thread's first centre vector inside a stage; row r adds r*ST_TR*TKB

    auto load_vec = [&](T const * st, int off, T * v) {
        uint4 q = *reinterpret_cast<uint4 const *>(st + off);
        memcpy(v, &q, 16);
    };
    // PUSH is a compile-time tag: only the CTAs of the two edge chunks run the instantiation that also stores into the neighbours'
    // ghost planes; measured, the extra (never taken) store path in the inner loop cost EVERY CTA 18 % (0.198 vs 0.168 ms on
    // a 130 x 1024^2 slab).
    auto store_out = [&](auto push_tag, int i, int r, T const * o) {
        T * d = static_cast<T *>(p.dst) + int64_t(i)*p.ds0 + int64_t(j+r*ST_TR)*p.ds1 + int64_t(k)*p.ds2;
        if (all_ok[r] && p.vec_store) { uint4 q; memcpy(&q, o, 16); *reinterpret_cast<uint4 *>(d) = q; }
        else {
#pragma unroll
            for (int l=0; l<VT; ++l) { if (lane_ok[r][l]) d[int64_t(l)*p.ds2] = o[l]; }
        }
        if constexpr (FUSED && decltype(push_tag)::value) { // the same values into the neighbours' ghost planes (same inner strides)
#pragma unroll
            for (int side=0; side<2; ++side) {
                void * const peer = side ? p.push_hi : p.push_lo;
                if (!peer || i!=(side ? p.n0-2 : 1) || (p.dbg&1)) continue;
                T * g = static_cast<T *>(peer) + int64_t(j+r*ST_TR)*p.ds1 + int64_t(k)*p.ds2;
                if (all_ok[r] && p.vec_store) { uint4 q; memcpy(&q, o, 16); *reinterpret_cast<uint4 *>(g) = q; }
                else {
#pragma unroll
                    for (int l=0; l<VT; ++l) { if (lane_ok[r][l]) g[int64_t(l)*p.ds2] = o[l]; }
                }
            }
        }
    };

    if constexpr (!is3d) {
        // One tile per CTA (round 1) left the TMA latency and the CTA start-up exposed: 0.57-0.78 of the HBM peak on 4100^2 / 8196^2. Now the
        // tiles of a strip stream through the ring like the planes of the 3-D sweep.
        bool kok[VT], kall = true;
#pragma unroll
        for (int l=0; l<VT; ++l) { kok[l] = (k+l>=1) && (k+l<p.n2-1); kall = kall && kok[l]; }
        for (int q=0; q<nplanes; ++q) {
            mbar_wait(&full[q % ST_STAGES], uint32_t((q/ST_STAGES)&1));
            T const * st = stages + size_t(q % ST_STAGES)*PLANE;
#pragma unroll
            for (int r=0; r<ST_RPT; ++r) {
                int const sr = so + r*ST_TR*TKB, jr = j + q*ST_TJ + r*ST_TR;
                T c[VT], jp[VT], jm[VT], o[VT];
                load_vec(st, sr, c); load_vec(st, sr+TKB, jp); load_vec(st, sr-TKB, jm);
                T const left = st[sr-1], right = st[sr+VT];
                if constexpr (KIND==ST_KIND_MASS) {
                    T const jpr = st[sr+TKB+VT], jml = st[sr-TKB-1];   // the two diagonal neighbours outside this thread's vectors
#pragma unroll
                    for (int l=0; l<VT; ++l) {
                        T const km = l==0 ? left : c[l-1], kp = l==VT-1 ? right : c[l+1];
                        T const dpp = l==VT-1 ? jpr : jp[l+1], dmm = l==0 ? jml : jm[l-1];
                        T const six = ((((jm[l] + km) + jp[l]) + kp) + dpp) + dmm;          // in T, in the order written (rows are axis 0 of the expression)
                        o[l] = T(p.c0 * (double(c[l])/p.c1 + double(six)/p.c2));
                    }
                } else {
#pragma unroll
                    for (int l=0; l<VT; ++l) {
                        T km = l==0 ? left : c[l-1], kp = l==VT-1 ? right : c[l+1];
                        o[l] = stencil_value<T, KIND>(c[l], T(0), jp[l], kp, T(0), jm[l], km);
                    }
                }
                if (jr<p.n1-1) {
                    T * d = static_cast<T *>(p.dst) + int64_t(jr)*p.ds1 + int64_t(k)*p.ds2;
                    if (kall && p.vec_store) { uint4 v; memcpy(&v, o, 16); *reinterpret_cast<uint4 *>(d) = v; }
                    else {
#pragma unroll
                        for (int l=0; l<VT; ++l) { if (kok[l]) d[int64_t(l)*p.ds2] = o[l]; }
                    }
                }
            }
            __syncthreads();                                   // tile q is no longer read by anyone
            if (0==tid && q+ST_STAGES<nplanes) issue(q+ST_STAGES);
        }
    } else {
      auto march = [&](auto push_tag) {
        T prev[ST_RPT][VT], cur[ST_RPT][VT], next[ST_RPT][VT];
        mbar_wait(&full[0], 0);
#pragma unroll
        for (int r=0; r<ST_RPT; ++r) load_vec(stages, so + r*ST_TR*TKB, prev[r]);
        if (nplanes>1) {
            mbar_wait(&full[1 % ST_STAGES], 0);
#pragma unroll
            for (int r=0; r<ST_RPT; ++r) load_vec(stages + size_t(1 % ST_STAGES)*PLANE, so + r*ST_TR*TKB, cur[r]);
        }
        __syncthreads();                                   // everyone is done with plane 0
        if (0==tid && ST_STAGES<nplanes) issue(ST_STAGES);  // refill its stage
        for (int q=1; q+1<nplanes; ++q) {                   // output plane i = ilo-1+q
            int const sn = (q+1) % ST_STAGES;
            mbar_wait(&full[sn], uint32_t(((q+1)/ST_STAGES)&1));
            T const * st = stages + size_t(q % ST_STAGES)*PLANE;
#pragma unroll
            for (int r=0; r<ST_RPT; ++r) {
                int const sr = so + r*ST_TR*TKB;
                load_vec(stages + size_t(sn)*PLANE, sr, next[r]);
                T jp[VT], jm[VT], o[VT];
                load_vec(st, sr+TKB, jp); load_vec(st, sr-TKB, jm);
                T const left = st[sr-1], right = st[sr+VT];
#pragma unroll
                for (int l=0; l<VT; ++l) {
                    T km = l==0 ? left : cur[r][l-1], kp = l==VT-1 ? right : cur[r][l+1];
                    o[l] = stencil_value<T, KIND>(cur[r][l], next[r][l], jp[l], kp, prev[r][l], jm[l], km);
                }
                store_out(push_tag, ilo-1+q, r, o);
#pragma unroll
                for (int l=0; l<VT; ++l) { prev[r][l] = cur[r][l]; cur[r][l] = next[r][l]; }
            }
            __syncthreads();                                  // plane q is no longer read by anyone
            if (0==tid && q+ST_STAGES<nplanes) issue(q+ST_STAGES);
        }
      };
        if constexpr (FUSED) { if (touch_lo || touch_hi) march(std::true_type {}); else march(std::false_type {}); }
        else march(std::false_type {});
        if constexpr (FUSED) {
            // Edge CTAs count themselves; the last one of an edge chunk tells that neighbour "step p.step+1 done on my side":
            // every push has been made visible system wide (fence after the CTA barrier, cumulative) and every read of my
            // ghost plane has completed (each TMA stage was waited for).
            if ((touch_lo || touch_hi) && 0==tid && !(p.dbg&4)) {
                asm volatile("fence.acq_rel.sys;" ::: "memory"); // release is enough (no sequential consistency needed): lighter than __threadfence_system()
#pragma unroll
                for (int side=0; side<2; ++side) {
                    if (!(side ? touch_hi : touch_lo)) continue;
                    if (atomicAdd(p.flags+2+side, 1u)==uint32_t(p.edge_ctas-1)) {
                        p.flags[2+side] = 0;
                        asm volatile("fence.acq_rel.sys;" ::: "memory");
                        st_release_sys(side ? p.signal_hi : p.signal_lo, p.step+1);
                    }
                }
            }
        }
    }
}

// ---------------- two sweeps per pass (temporal blocking of the 7-point sweep) ----------------
// T ping-pong steps read and write the whole grid T times. This kernel makes TWO steps per pass: state t is streamed in with the same TMA
// ring, state t+1 exists only as three planes of shared memory, state t+2 is written out - 8 B per cell for two steps instead of 16.
// Every cell of t+1 and t+2 is produced by the same expression in the same association order as a single step, so the results are
// bit-identical to two single steps (asserted by the tests on both buffers).
//
// Geometry: stage 1 (t -> t+1) of a CTA covers ST_TJ x TK cells of a plane, exactly what the single-step kernel computes from the same
// rows; stage 2 (t+1 -> t+2) covers the cells whose neighbours stage 1 produced: rows 1 .. ST_TJ-2 and the vectors 1 .. TK/VT-2 of the
// tile. Tiles therefore overlap by two rows and two vectors (14 x 120 outputs per 18 x 128 box for f32) and a chunk of P output planes
// reads P+4 input planes; the overlaps are L2 hits like the halos of the single-step kernel. The k-1 / k+1 neighbours of a thread's
// vector come from the neighbouring LANES (two shuffles per row): the first and last lane of a warp then compute garbage in the outermost
// cell of the tile row, which nothing uses (stage 2 needs t+1 from the last cell of vector 0 and the first of vector 31 only), so the box
// needs no halo along k - and the 4-byte shared-memory reads at a stride of 16 bytes that the single-step kernel makes for these two
// neighbours (a 4-way bank conflict each; 39 % of the first version's shared-memory wavefronts, which ran the shared-memory pipe at
// 70 %) are gone. A thread owns two ADJACENT rows (each row's j+1 / j-1 neighbour of the other comes from registers: 6 instead of 10
// 16-byte shared-memory loads per thread and plane) and marches along axis 0 with the centre values of t (i-1, i, i+1) and of t+1
// (i-2, i-1, i) in registers; t+1 goes through a ring of THREE shared planes so that one CTA barrier per plane is enough.
//
// Boundary cells keep the values of the buffer a state lives in (bench/bench-stencil3.cc:125-126), and in the reference's ping-pong the
// odd states live in the OTHER buffer: the boundary cells of t+1 are read from `bsrc` (that buffer; through pointers set up once and
// prefetched one plane ahead, see the loop), those of t+2 are never written. CTAs that touch no boundary run a copy of the loop without
// any of this. Tuning history and ncu evidence: profiles/r02_stencil2_full_summary.md, DESIGN.md 3.7.

// One vector of 7-point updates, -6*c + ip + jp + kp + im + jm + km per lane in exactly this order. PACKED (float): the six additions of two
// lanes at a time as add.rn.f32x2 (Blackwell's packed FP32 pipe: one instruction, two IEEE round-to-nearest additions, bit-identical to two
// scalar ones); the products stay scalar (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false: not bit-exact). The
// k-1 / k+1 operands straddle the lane pairs and are re-packed (three pairs per vector).
__device__ __forceinline__ uint64_t f2_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t x, float & a, float & b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <class T, bool PACKED, int VT> __device__ __forceinline__ void
row7(T (&o)[VT], T const (&c)[VT], T const (&ip)[VT], T const (&jp)[VT], T const (&im)[VT], T const (&jm)[VT], T left, T right)
{
    if constexpr (PACKED && std::is_same_v<T, float> && 4==VT) {
        uint64_t const k12 = f2_pack(c[1], c[2]);
        uint64_t const kp[2] = {k12, f2_pack(c[3], right)}, km[2] = {f2_pack(left, c[0]), k12};
#pragma unroll
        for (int h=0; h<2; ++h) {
            uint64_t t = f2_pack(T(-6)*c[2*h], T(-6)*c[2*h+1]);
            t = f2_add(t, f2_pack(ip[2*h], ip[2*h+1]));
            t = f2_add(t, f2_pack(jp[2*h], jp[2*h+1]));
            t = f2_add(t, kp[h]);
            t = f2_add(t, f2_pack(im[2*h], im[2*h+1]));
            t = f2_add(t, f2_pack(jm[2*h], jm[2*h+1]));
            t = f2_add(t, km[h]);
            f2_unpack(t, o[2*h], o[2*h+1]);
        }
    } else {
#pragma unroll
        for (int l=0; l<VT; ++l) {
            T const km = l==0 ? left : c[l-1], kp = l==VT-1 ? right : c[l+1];
            o[l] = stencil_value<T, RACU_STENCIL_3D7>(c[l], ip[l], jp[l], kp, im[l], jm[l], km);
        }
    }
}

template <class T> struct St2Tile
{
    static constexpr int VT = StTile<T>::VT, TK = StTile<T>::TK;
    static constexpr int PLANE = (ST_TJ+2)*TK;                // the TMA box: ST_TJ+2 rows of TK cells, NO halo along k (see the kernel)
    static constexpr int OUT_J = ST_TJ-2, OUT_K = TK-2*VT;    // outputs per tile and plane
    static constexpr int IPLANE = ST_TJ*TK;                   // one plane of t+1
    static_assert(TK/VT==32 && ST_RPT==2 && 0==(PLANE*sizeof(T))%128);
};

struct St2Params
{
    void * dst; int64_t ds0, ds1, ds2;
    void const * bsrc; int64_t bs0, bs1, bs2;   // the buffer whose boundary cells state t+1 carries
    int32_t n0, n1, n2;
    int32_t i_lo, i_hi;                          // output planes [i_lo, i_hi)
    int32_t ci, gz, gyb;
    int32_t vec_store;
};

constexpr int S2_STAGES_DEFAULT = 5;   // TMA ring of the two-step kernel: plane m+5 is requested when plane m is released, four planes ahead of its use (the loads are DRAM
                               // latency under load; with a ring of three the CTAs waited for data most of the time: 37 % issue-slot utilisation in the lean loop)

template <class T, int S2_STAGES, bool PACKED> __global__ void __launch_bounds__(ST_THREADS, (3==S2_STAGES ? 4 : 6==S2_STAGES ? 2 : 3))
stencil2_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ St2Params p)
{
    using TL = St2Tile<T>;
    constexpr int VT = TL::VT, TK = TL::TK, TKB = TK, PLANE = TL::PLANE, IPLANE = TL::IPLANE, KIND = RACU_STENCIL_3D7;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T * const stages = reinterpret_cast<T *>(smem_raw);
    T * const inter = stages + size_t(S2_STAGES)*PLANE;       // 3 planes of t+1 (+ one row of slack behind them)
    __shared__ __align__(8) uint64_t full[S2_STAGES];

    int const tid = threadIdx.x, tk = tid&31, tj = tid>>5;
    int const band = int(blockIdx.z)/p.gz, chunk = int(blockIdx.z) - band*p.gz;
    int const kb = int(blockIdx.x)*TL::OUT_K - VT;                                   // first stage-1 column of the tile (may be -VT)
    int const jb = (band*p.gyb + int(blockIdx.y))*TL::OUT_J - 1;                     // first stage-1 row (may be -1)
    if (jb+1>=p.n1-1) return;                                                        // no output row (whole CTA, before any barrier)
    int const ilo = p.i_lo + chunk*p.ci, ihi = min(p.i_hi, ilo+p.ci);
    int const nin = (ihi-ilo+2+2)/3*3 + 2;                                           // input planes ilo-2 .. ihi+1, rounded up so that the planes of t+1 are a multiple of three
    constexpr uint32_t plane_bytes = PLANE*sizeof(T);

    if (0==tid) {
#pragma unroll
        for (int s=0; s<S2_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int q, int s) { // input plane q of this chunk -> stage s = q % S2_STAGES
        mbar_expect_tx(&full[s], plane_bytes);
        tma_load_3d(stages + size_t(s)*PLANE, &tmap, kb, jb-1, ilo-2+q, &full[s]); // planes -1 and >= n0 are zero filled (never used)
    };
    if (0==tid) { for (int q=0; q<S2_STAGES && q<nin; ++q) issue(q, q); }

    int const k = kb + tk*VT;
    // per row of this thread (tile rows 2*tj, 2*tj+1): which lanes of t+1 come from bsrc, which lanes of t+2 are stored
    uint32_t fix[2], dom[2]; bool lane_ok[2][VT], all_ok[2], any_ok[2];
#pragma unroll
    for (int r=0; r<2; ++r) {
        int const R = 2*tj+r, j = jb+R;
        bool const row_in = j>=0 && j<p.n1, row_bd = 0==j || j==p.n1-1, row_out = R>=1 && R<=ST_TJ-2 && j>=1 && j<p.n1-1 && tk>=1 && tk<=30;
        fix[r] = 0; dom[r] = 0; all_ok[r] = true; any_ok[r] = false;
#pragma unroll
        for (int l=0; l<VT; ++l) {
            bool const in = row_in && k+l>=0 && k+l<p.n2, bd = row_bd || 0==k+l || k+l==p.n2-1;
            if (in) dom[r] |= 1u<<l;
            if (in && bd) fix[r] |= 1u<<l;
            lane_ok[r][l] = row_out && k+l>=1 && k+l<p.n2-1;
            all_ok[r] = all_ok[r] && lane_ok[r][l]; any_ok[r] = any_ok[r] || lane_ok[r][l];
        }
    }
    auto load_vec = [&](T const * st, int off, T * v) { uint4 q = *reinterpret_cast<uint4 const *>(st + off); memcpy(v, &q, 16); };
    auto store_vec_s = [&](T * st, int off, T const * v) { uint4 q; memcpy(&q, v, 16); *reinterpret_cast<uint4 *>(st + off) = q; };

    // Everything that does not change from plane to plane is computed here, every shared-memory access of the loop is one base pointer
    // plus an immediate, and CTAs whose tile and chunk touch no boundary run a loop without the boundary code. The first version spent
    // 27 instructions per cell update (7 of them arithmetic) and was issue bound: 72 % issue-slot utilisation, DRAM at 31 %
    // (profiles/r02_stencil2_*).
    T const * const sp = stages + (2*tj+1)*TKB + tk*VT;        // this thread's first centre vector in stage 0; its second row: + TKB
    T * const ip = inter + (2*tj)*TK + tk*VT;                  // the same in plane 0 of t+1; second row: + TK
    T const * const ipj0 = tj>0 ? ip-TK : ip, * const ipj1 = tj<ST_TR-1 ? ip+2*TK : ip+TK; // rows above / below for stage 2 (tile rows 0 and 15 store nothing)
    T * dptr[2];                                        // where this thread's vector of t+2 at plane ilo goes; + ds0 per plane
    int smode[2];                                       // 0: stores nothing, 1: one 16-byte store, 2: lane by lane
#pragma unroll
    for (int r=0; r<2; ++r) {
        dptr[r] = static_cast<T *>(p.dst) + int64_t(ilo)*p.ds0 + int64_t(jb+2*tj+r)*p.ds1 + int64_t(k)*p.ds2;
        smode[r] = !any_ok[r] ? 0 : (all_ok[r] && p.vec_store) ? 1 : 2;
    }
    int64_t const ds0 = p.ds0;
    int const n0m1 = p.n0-1;
    bool const t0 = 0==tid;
    // boundary path only: where the boundary cells of t+1 at plane ilo-1 (m = 1) are read from, + bs0 per plane; the lanes of t+2 to store as a mask
    // (a third of the CTAs of a 1024^3 sweep touch a boundary and ran 347 instructions per plane against 165 for the others: predicate logic
    // and a 64-bit address from scratch for every boundary read)
    T const * bptr[2];
    uint32_t smask[2], bneed[2];
#pragma unroll
    for (int r=0; r<2; ++r) {
        bptr[r] = static_cast<T const *>(p.bsrc) + int64_t(ilo-1)*p.bs0 + int64_t(jb+2*tj+r)*p.bs1 + int64_t(k)*p.bs2;
        bneed[r] = ilo-1>p.n0-1 ? 0u : (0==ilo-1 || ilo-1==p.n0-1) ? dom[r] : fix[r];
        smask[r] = 0;
#pragma unroll
        for (int l=0; l<VT; ++l) smask[r] |= lane_ok[r][l] ? 1u<<l : 0u;
    }
    int64_t const bs0 = p.bs0, bs2 = p.bs2, ds2 = p.ds2;
    // no boundary cell of t+1 in this CTA's tile and planes, whole vectors, no plane past the end of the chunk
    bool const interior = jb>=1 && jb+ST_TJ<=p.n1-1 && kb>=1 && kb+TK<=p.n2-1 && ilo>=2 && ilo-2+(nin-2)<=p.n0-2 && nin==ihi-ilo+4 && p.vec_store;

    struct Regs { T v[2][VT]; };
    Regs p1, c1, n1v;                     // state t at planes m-1, m, m+1
    Regs p2, c2, n2v;                     // state t+1 at planes m-2, m-1, m
#pragma unroll
    for (int r=0; r<2; ++r)
#pragma unroll
        for (int l=0; l<VT; ++l) { p2.v[r][l] = T(0); c2.v[r][l] = T(0); n2v.v[r][l] = T(0); }
    mbar_wait(&full[0], 0);
#pragma unroll
    for (int r=0; r<2; ++r) load_vec(sp, r*TKB, p1.v[r]);
    mbar_wait(&full[1], 0);
#pragma unroll
    for (int r=0; r<2; ++r) load_vec(sp, PLANE + r*TKB, c1.v[r]);
    __syncthreads();                                    // everyone is done with plane 0
    if (t0 && S2_STAGES<nin) issue(S2_STAGES, 0);
    // the ring: plane m of t sits in stage si (pointer pst), plane m+1 in stage sni (psn, waited for with parity npar)
    T const * pst = sp + PLANE, * psn = sp + 2*PLANE;
    int si = 1, sni = 2;
    uint32_t npar = 0;
    uint32_t const bar0 = smem_u32(&full[0]);

    // One plane m = 3g+1+PH. PH is a compile-time value (the loop below is unrolled by three): t+1 at plane m goes to plane (1+PH)%3 of
    // `inter` and its neighbours at plane m-1 are read from plane PH; the six register sets rotate by renaming, without copies.
    auto plane = [&]<int PH, bool INTERIOR>(std::integral_constant<int, PH>, std::integral_constant<bool, INTERIOR>, int m,
                                            Regs & rP1, Regs & rC1, Regs & rN1, Regs & rP2, Regs & rC2, Regs & rN2) {
        auto & P1 = rP1.v; auto & C1 = rC1.v; auto & N1 = rN1.v; auto & P2 = rP2.v; auto & C2 = rC2.v; auto & N2 = rN2.v;
        constexpr int IW = (1+PH)%3, IB = PH;
        int const gi = ilo-2+m;
        mbar_wait_a(bar0 + 8u*uint32_t(sni), npar);
#pragma unroll
        for (int r=0; r<2; ++r) load_vec(psn, r*TKB, N1[r]);
        // ---- stage 1: t+1 at plane m
        {
            T jm0[VT], jp1[VT];
            load_vec(pst, -TKB, jm0); load_vec(pst, 2*TKB, jp1);
#pragma unroll
            for (int r=0; r<2; ++r) {
                T const left = __shfl_up_sync(0xffffffffu, C1[r][VT-1], 1), right = __shfl_down_sync(0xffffffffu, C1[r][0], 1);
                if (0==r) row7<T, PACKED, VT>(N2[0], C1[0], N1[0], C1[1], P1[0], jm0, left, right);
                else row7<T, PACKED, VT>(N2[1], C1[1], N1[1], jp1, P1[1], C1[0], left, right);
                if constexpr (!INTERIOR) {
                    uint32_t const need = bneed[r];               // boundary cells of t+1: the values of the buffer t+1 lives in (which lanes: worked out one plane ahead)
                    if (need) {
#pragma unroll
                        for (int l=0; l<VT; ++l) { if ((need>>l)&1u) N2[r][l] = bptr[r][int64_t(l)*bs2]; }
                    }
                    bptr[r] += bs0;
                    // the boundary cells of the NEXT plane, one plane ahead: read on demand, every plane of a boundary CTA waited a DRAM round trip
                    // here before it could store t+1 and reach the barrier (ncu: the long-scoreboard stalls of the kernel sat on these stores)
                    uint32_t nneed = fix[r];
                    if (gi+1>=n0m1) nneed = gi+1==n0m1 ? dom[r] : 0u;
                    bneed[r] = nneed;
                    if (nneed) asm volatile("prefetch.global.L1 [%0];" :: "l"(bptr[r] + ((nneed&1u) ? 0 : (VT-1)*bs2)));
                }
                store_vec_s(ip, IW*IPLANE + r*TK, N2[r]);
            }
        }
        __syncthreads();                                // t+1 at plane m is visible; plane m of t is no longer read by anyone
        if (t0 && m+S2_STAGES<nin) issue(m+S2_STAGES, si);
        si = sni; pst = psn;
        if (++sni==S2_STAGES) { sni = 0; psn = sp; npar ^= 1u; } else psn += PLANE;
        // ---- stage 2: t+2 at plane m-1 from t+1 at planes m-2, m-1 (neighbours in shared memory), m
        if (m>=3) {
            T jm0[VT], jp1[VT];
            load_vec(ipj0, IB*IPLANE, jm0); load_vec(ipj1, IB*IPLANE, jp1);
            bool const live = INTERIOR || gi-1<ihi;     // (the last chunk may run planes past its end: the loop count is a multiple of three)
#pragma unroll
            for (int r=0; r<2; ++r) {
                T const left = __shfl_up_sync(0xffffffffu, C2[r][VT-1], 1), right = __shfl_down_sync(0xffffffffu, C2[r][0], 1);
                T o[VT];
                if (0==r) row7<T, PACKED, VT>(o, C2[0], N2[0], C2[1], P2[0], jm0, left, right);
                else row7<T, PACKED, VT>(o, C2[1], N2[1], jp1, P2[1], C2[0], left, right);
                if (1==smode[r] && live) { uint4 q; memcpy(&q, o, 16); *reinterpret_cast<uint4 *>(dptr[r]) = q; }
                if constexpr (!INTERIOR) {
                    if (2==smode[r] && live) {
#pragma unroll
                        for (int l=0; l<VT; ++l) { if ((smask[r]>>l)&1u) dptr[r][int64_t(l)*ds2] = o[l]; }
                    }
                }
                dptr[r] += ds0;
            }
        }
    };
    // nin-2 planes of t+1, a multiple of three (planes past n0 are zero filled by TMA and never stored)
    auto run = [&](auto interior_tag) {
        for (int m=1; m+1<nin; m+=3) {
            plane(std::integral_constant<int, 0> {}, interior_tag, m, p1, c1, n1v, p2, c2, n2v);
            plane(std::integral_constant<int, 1> {}, interior_tag, m+1, c1, n1v, p1, c2, n2v, p2);
            plane(std::integral_constant<int, 2> {}, interior_tag, m+2, n1v, p1, c1, n2v, p2, c2);
        }
    };
    if (interior) run(std::true_type {}); else run(std::false_type {});
}

// copies the boundary cells (the six faces) of one buffer into another of the same shape
template <class T> __global__ void
copy_faces_kernel(T const * src, int64_t ss0, int64_t ss1, int64_t ss2, T * dst, int64_t ds0, int64_t ds1, int64_t ds2, int n0, int n1, int n2)
{
    int const face = blockIdx.y; // 0,1: i = 0 / n0-1; 2,3: j; 4,5: k
    int64_t const na = face<2 ? n1 : n0, nb = face<4 ? n2 : n1, total = na*nb;
    for (int64_t e = int64_t(blockIdx.x)*blockDim.x + threadIdx.x; e<total; e += int64_t(gridDim.x)*blockDim.x) {
        int64_t const a = e/nb, b = e - a*nb;
        int64_t i, j, k;
        if (face<2) { i = (face&1) ? n0-1 : 0; j = a; k = b; }
        else if (face<4) { i = a; j = (face&1) ? n1-1 : 0; k = b; }
        else { i = a; j = b; k = (face&1) ? n2-1 : 0; }
        dst[i*ds0 + j*ds1 + k*ds2] = src[i*ss0 + j*ss1 + k*ss2];
    }
}

// ---------------- 1-D 3-point sweep: Anext(I) = -2*A(I) + A(I+1) + A(I-1) (bench/bench-stencil1.cc:45) ----------------
// No staging at all: a lane loads whole 16-byte vectors of the row, its k-1 / k+1 neighbours are the neighbouring lanes' registers (two
// shuffles per vector); only the first and the last lane of a warp read one more element each (a sector their neighbours load anyway).
// The slice expression shifts every operand by one element, so the vector kernels of the planner cannot take it (rows not 16-byte
// aligned) and the lane-wise kernel pays three 4-byte loads per element: 0.73 of the HBM peak; this one reads and writes every byte once.
constexpr int S1_UNR = 4;

template <class T> __global__ void __launch_bounds__(ST_THREADS, 3)
stencil1d3_kernel(T const * __restrict__ src, T * __restrict__ dst, int64_t ds, int64_t n, int vec_store)
{
    constexpr int VT = 16/sizeof(T);
    int const lane = threadIdx.x&31;
    int64_t const nvec = (n+VT-1)/VT, warps = int64_t(gridDim.x)*(ST_THREADS/32), w = int64_t(blockIdx.x)*(ST_THREADS/32) + (threadIdx.x>>5);
    for (int64_t v0 = w*32*S1_UNR; v0<nvec; v0 += warps*32*S1_UNR) {   // (v0 is the same for the whole warp: every lane reaches the shuffles)
        T c[S1_UNR][VT], left[S1_UNR], right[S1_UNR];
#pragma unroll
        for (int u=0; u<S1_UNR; ++u) {
            int64_t const e = (v0 + u*32 + lane)*VT;
            if (e+VT<=n) { uint4 q = *reinterpret_cast<uint4 const *>(src + e); memcpy(c[u], &q, 16); }
            else {
#pragma unroll
                for (int l=0; l<VT; ++l) c[u][l] = e+l<n ? src[e+l] : T(0);
            }
            left[u] = T(0); right[u] = T(0);
            if (0==lane && e>0 && e<=n) left[u] = src[e-1];
            if (31==lane && e+VT<n) right[u] = src[e+VT];
        }
#pragma unroll
        for (int u=0; u<S1_UNR; ++u) {
            int64_t const e = (v0 + u*32 + lane)*VT;
            T const lf = __shfl_up_sync(0xffffffffu, c[u][VT-1], 1), rt = __shfl_down_sync(0xffffffffu, c[u][0], 1);
            if (0!=lane) left[u] = lf;
            if (31!=lane) right[u] = rt;
            T o[VT];
#pragma unroll
            for (int l=0; l<VT; ++l) {
                T const km = l==0 ? left[u] : c[u][l-1], kp = l==VT-1 ? right[u] : c[u][l+1];
                o[l] = ((T(-2)*c[u][l]) + kp) + km;
            }
            if (vec_store && e>=1 && e+VT<=n-1) { uint4 q; memcpy(&q, o, 16); *reinterpret_cast<uint4 *>(dst + e) = q; }
            else {
#pragma unroll
                for (int l=0; l<VT; ++l) { if (e+l>=1 && e+l<n-1) dst[(e+l)*ds] = o[l]; }
            }
        }
    }
}

// ---------------- host ----------------

using EncodeTiled = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled
encode_fn()
{
    static EncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []{
        void * p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaSuccess==cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) && q==cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiled>(p);
        else cudaGetLastError();
    });
    return fn;
}

template <class T, int KIND> static int
launch_stencil(racu_stencil_desc const * s, cudaStream_t stream, int * handled, racu_halo_push const * h = nullptr, double const * coef = nullptr)
{
    using TL = StTile<T>;
    constexpr bool is3d = KIND==RACU_STENCIL_3D7;
    int const r = s->rank;
    // view the array as (n0, n1, n2) with n0 the marching axis; 2-D arrays get n0 = 1
    int64_t n0 = is3d ? s->len[0] : 1, n1 = s->len[r-2], n2 = s->len[r-1];
    int64_t ss0 = is3d ? s->src_stride[0] : 0, ss1 = s->src_stride[r-2], ss2 = s->src_stride[r-1];
    int64_t ds0 = is3d ? s->dst_stride[0] : 0, ds1 = s->dst_stride[r-2], ds2 = s->dst_stride[r-1];
    size_t const es = sizeof(T);
    // TMA constraints on the source; anything else falls back to the fused ply kernels (still exact)
    if (1!=ss2 || 0!=(uintptr_t(s->src)%16) || 0!=((ss1*es)%16) || (is3d && 0!=((ss0*es)%16)) || ss1<=0 || (is3d && ss0<=0)) return RACU_OK;
    if (n0>=(int64_t(1)<<31) || n1>=(int64_t(1)<<31) || n2>=(int64_t(1)<<31) || n1<3 || n2<3 || (is3d && n0<3)) return RACU_OK;
    EncodeTiled enc = encode_fn();
    if (!enc) return RACU_OK;
    int64_t lo = s->lo0, hi = s->hi0;
    if (0==lo && 0==hi) hi = s->len[0];
    if (lo<1) lo = 1;
    if (hi>s->len[0]-1) hi = s->len[0]-1;
    if (is3d && hi<=lo) { *handled = 1; return RACU_OK; }

    CUtensorMap tmap;
    cuuint64_t gdim[3] = {cuuint64_t(n2), cuuint64_t(n1), cuuint64_t(n0)};
    cuuint64_t gstr[2] = {cuuint64_t(ss1*es), cuuint64_t((is3d ? ss0 : ss1*n1)*es)};
    cuuint32_t box[3] = {cuuint32_t(TL::TKB), cuuint32_t(ST_TJ+2), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = enc(&tmap, sizeof(T)==4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(s->src),
                      gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (CUDA_SUCCESS!=cr) return RACU_OK; // not expressible as a tensor map: fall back

    StParams p {};
    p.dst = s->dst; p.ds0 = ds0; p.ds1 = ds1; p.ds2 = ds2;
    p.n0 = int32_t(n0); p.n1 = int32_t(n1); p.n2 = int32_t(n2);
    p.i_lo = int32_t(lo); p.i_hi = int32_t(hi);
    p.vec_store = 1==ds2 && 0==(uintptr_t(s->dst)%16) && 0==((ds1*es)%16) && (!is3d || 0==((ds0*es)%16));
    if (coef) { p.c0 = coef[0]; p.c1 = coef[1]; p.c2 = coef[2]; }
    unsigned gx = unsigned((n2+TL::TK-1)/TL::TK), gy = unsigned((n1-2+ST_TJ-1)/ST_TJ), gz = 1;
    if (is3d) {
        // Short chunks of the marching axis: measured on B200 (1024^3 f32), 4-6 planes per CTA give 826-829 Gcells/s against
        // 654 for 256 and 535 for 1024. The two extra halo planes per chunk are L2 hits (neighbouring chunks are resident
        // together), while the 30x larger grid removes the tail and keeps every SM's TMA queue full.
        int64_t planes = hi-lo;
        p.ci = int32_t(std::min<int64_t>(planes, 4));
        if (char const * e = std::getenv("RACU_ST_CI")) p.ci = std::max(1, std::atoi(e)); // tuning knob
        gz = unsigned((planes+p.ci-1)/p.ci);
        if (gz>65535) { p.ci = int32_t((planes+65534)/65535); gz = unsigned((planes+p.ci-1)/p.ci); }
    }
    p.gz = int32_t(gz); p.gyb = int32_t(gy);
    unsigned const tiles_y = gy;                 // valid tile rows (edge CTA count of the fused step)
    if (!is3d) {                                 // 2-D: 2 or 4 tile rows per CTA (8196^2 on B200, fraction of the HBM peak for 2 / 4 / 8 / 16: stiffness f64 0.99 / 0.99 / 0.96 / 0.92,
                                                 // f32 0.87 / 0.89 / 0.91 / 0.82; one tile per CTA, round 1: 0.75 / 0.78)
        p.ci = int64_t(gx)*gy>=20000 ? 4 : int64_t(gx)*gy>=1480 ? 2 : 1;
        if (char const * e = std::getenv("RACU_ST_CI2")) p.ci = std::max(1, std::atoi(e)); // tuning knob
        gy = (gy+p.ci-1)/p.ci;
    }
    if (is3d) {
        // bands of about 512 tiles (the 1024^2 plane this was tuned on: 8 x 64 tiles per chunk)
        int64_t band_tiles = 512;
        if (char const * e = std::getenv("RACU_ST_BAND")) band_tiles = std::max(1, std::atoi(e)); // tuning knob
        unsigned gyb = unsigned(std::max<int64_t>(1, band_tiles/gx));
        unsigned nb = (gy+gyb-1)/gyb;
        if (nb>1 && uint64_t(nb)*gz<=65535) { p.gyb = int32_t(gyb); gy = gyb; gz *= nb; }
    }
    size_t smem = size_t(ST_STAGES)*TL::PLANE*sizeof(T);
    auto kernel = stencil_tma_kernel<T, KIND, false>;
    if constexpr (is3d) {
        if (h) {
            kernel = stencil_tma_kernel<T, KIND, true>;
            p.push_lo = h->push_lo; p.push_hi = h->push_hi; p.flags = h->flags; p.signal_lo = h->signal_lo; p.signal_hi = h->signal_hi;
            p.step = h->step; p.edge_ctas = int32_t(gx*tiles_y);
            if (char const * e = std::getenv("RACU_PUSH_DEBUG")) p.dbg = std::atoi(e);
        }
    }
    if (smem>32*1024) { if (cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) return cuda_fail(e, "stencil smem"); }
    kernel<<<dim3(gx, gy, gz), ST_THREADS, smem, stream>>>(tmap, p);
    launch_count_add(1);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "stencil_tma");
    set_kernel(is3d ? (h ? "stencil_tma<3D7, halo push>" : "stencil_tma<3D7>") : (KIND==RACU_STENCIL_2D5 ? "stencil_tma<2D5>" : KIND==ST_KIND_MASS ? "stencil_tma<2D7_MASS>" : "stencil_tma<2D5_STIFF>"));
    *handled = 1;
    return RACU_OK;
}

int
try_special_stencil(racu_stencil_desc const * s, cudaStream_t stream, int * handled)
{
    *handled = 0;
    if (std::getenv("RACU_NO_TMA_STENCIL")) return RACU_OK;
    if (s->dtype!=RACU_F32 && s->dtype!=RACU_F64) return RACU_OK;
    if (s->rank<1 || s->rank>3) return RACU_OK;
    // An in-place sweep (dst overlapping src) would be a race here; the planner rejects it ("traversal order would be observable").
    if (views_overlap(s->src, s->rank, s->len, s->src_stride, s->dst, s->rank, s->len, s->dst_stride, s->dtype==RACU_F32 ? 4 : 8)) return RACU_OK;
#define RACU_ST(KIND) (s->dtype==RACU_F32 ? launch_stencil<float, KIND>(s, stream, handled) : launch_stencil<double, KIND>(s, stream, handled))
    if (RACU_STENCIL_1D3==s->kind) {
        if (1!=s->rank || 1!=s->src_stride[0] || 0!=(uintptr_t(s->src)%16) || s->len[0]<3 || 0!=s->lo0 || 0!=s->hi0 || s->dst_stride[0]<=0) return RACU_OK;
        int64_t const n = s->len[0], nvec = (n+(s->dtype==RACU_F32 ? 3 : 1))/(s->dtype==RACU_F32 ? 4 : 2);
        unsigned const grid = unsigned(std::min<int64_t>((nvec+32*S1_UNR*(ST_THREADS/32)-1)/(32*S1_UNR*(ST_THREADS/32)), 148*8));
        int const vs = 1==s->dst_stride[0] && 0==(uintptr_t(s->dst)%16);
        if (s->dtype==RACU_F32) stencil1d3_kernel<float><<<grid, ST_THREADS, 0, stream>>>(static_cast<float const *>(s->src), static_cast<float *>(s->dst), s->dst_stride[0], n, vs);
        else stencil1d3_kernel<double><<<grid, ST_THREADS, 0, stream>>>(static_cast<double const *>(s->src), static_cast<double *>(s->dst), s->dst_stride[0], n, vs);
        launch_count_add(1);
        if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "stencil1d3");
        set_kernel("stencil_1d3");
        *handled = 1;
        return RACU_OK;
    }
    switch (s->kind) {
    case RACU_STENCIL_3D7: return 3==s->rank ? RACU_ST(RACU_STENCIL_3D7) : RACU_OK;
    case RACU_STENCIL_2D5: return 2==s->rank ? RACU_ST(RACU_STENCIL_2D5) : RACU_OK;
    case RACU_STENCIL_2D5_STIFF: return 2==s->rank ? RACU_ST(RACU_STENCIL_2D5_STIFF) : RACU_OK;
    default: return RACU_OK;
    }
#undef RACU_ST
}

// ---- racu_stencil_steps: T ping-pong steps src -> dst -> src ..., two per pass where the arrays allow it
struct Buf3 { void * p; int64_t st[3]; };

template <class T> static bool
tma_ok(Buf3 const & b, int64_t const * n)
{
    size_t const es = sizeof(T);
    return 1==b.st[2] && 0==(uintptr_t(b.p)%16) && 0==((b.st[1]*es)%16) && 0==((b.st[0]*es)%16) && b.st[1]>0 && b.st[0]>0
        && n[0]<(int64_t(1)<<31) && n[1]<(int64_t(1)<<31) && n[2]<(int64_t(1)<<31) && n[0]>=3 && n[1]>=3 && n[2]>=3;
}

// one pass: dst interior = two steps from src; the boundary cells of the intermediate state are those of bsrc
template <class T> static int
launch_stencil2(Buf3 const & src, Buf3 const & dst, Buf3 const & bsrc, int64_t const * n, cudaStream_t stream)
{
    using TL = St2Tile<T>;
    size_t const es = sizeof(T);
    EncodeTiled enc = encode_fn();
    if (!enc) return fail(RACU_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
    CUtensorMap tmap;
    cuuint64_t gdim[3] = {cuuint64_t(n[2]), cuuint64_t(n[1]), cuuint64_t(n[0])};
    cuuint64_t gstr[2] = {cuuint64_t(src.st[1]*es), cuuint64_t(src.st[0]*es)};
    cuuint32_t box[3] = {cuuint32_t(TL::TK), cuuint32_t(ST_TJ+2), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = enc(&tmap, sizeof(T)==4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, src.p,
                      gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (CUDA_SUCCESS!=cr) return fail(RACU_ERR_UNSUPPORTED, "racu_stencil_steps: the array is not expressible as a TMA tensor map");
    St2Params p {};
    p.dst = dst.p; p.ds0 = dst.st[0]; p.ds1 = dst.st[1]; p.ds2 = dst.st[2];
    p.bsrc = bsrc.p; p.bs0 = bsrc.st[0]; p.bs1 = bsrc.st[1]; p.bs2 = bsrc.st[2];
    p.n0 = int32_t(n[0]); p.n1 = int32_t(n[1]); p.n2 = int32_t(n[2]);
    p.i_lo = 1; p.i_hi = int32_t(n[0]-1);
    p.vec_store = 1==dst.st[2] && 0==(uintptr_t(dst.p)%16) && 0==((dst.st[1]*es)%16) && 0==((dst.st[0]*es)%16);
    int64_t const planes = n[0]-2;
    // P output planes per CTA read P+4 input planes and compute P+2 planes of t+1: long chunks (measured on 1024^3 f32: 975 Gcells/s per step
    // at P = 16, 1018 at 61, 1006 at 94), of equal length, with P+2 a multiple of three (the kernel's loop is unrolled by three)
    int64_t const nchunks = (planes+60)/61;
    p.ci = int32_t((planes+nchunks-1)/nchunks);
    while (0!=(p.ci+2)%3) ++p.ci;
    if (char const * e = std::getenv("RACU_ST2_CI")) p.ci = std::max(1, std::atoi(e)); // tuning knob
    unsigned gx = unsigned((n[2]-1+TL::OUT_K-1)/TL::OUT_K), gy = unsigned((n[1]-1+TL::OUT_J-1)/TL::OUT_J), gz = unsigned((planes+p.ci-1)/p.ci);
    if (gz>65535) { p.ci = int32_t((planes+65534)/65535); gz = unsigned((planes+p.ci-1)/p.ci); }
    p.gz = int32_t(gz); p.gyb = int32_t(gy);
    {   // bands of tiles, as in the single-step kernel
        int64_t band_tiles = 512;
        if (char const * e = std::getenv("RACU_ST_BAND")) band_tiles = std::max(1, std::atoi(e));
        unsigned gyb = unsigned(std::max<int64_t>(1, band_tiles/gx));
        unsigned nb = (gy+gyb-1)/gyb;
        if (nb>1 && uint64_t(nb)*gz<=65535) { p.gyb = int32_t(gyb); gy = gyb; gz *= nb; }
    }
    int stages = S2_STAGES_DEFAULT;
    if (char const * e = std::getenv("RACU_ST2_STAGES")) stages = std::atoi(e); // tuning knob: 3 .. 6
    if (stages<3 || stages>6) stages = S2_STAGES_DEFAULT;
    size_t const smem = (size_t(stages)*TL::PLANE + 3*size_t(TL::IPLANE) + TL::TK)*es;
    bool packed = false; // add.rn.f32x2 for the additions of the float sweep (knob: RACU_ST2_PACKED)
    if (char const * e = std::getenv("RACU_ST2_PACKED")) packed = 0!=std::atoi(e);
    auto kernel = 3==stages ? stencil2_tma_kernel<T, 3, false> : 4==stages ? stencil2_tma_kernel<T, 4, false> : 6==stages ? stencil2_tma_kernel<T, 6, false>
                  : (packed && sizeof(T)==4) ? stencil2_tma_kernel<T, 5, true> : stencil2_tma_kernel<T, 5, false>;
    if (cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) return cuda_fail(e, "stencil2 smem");
    kernel<<<dim3(gx, gy, gz), ST_THREADS, smem, stream>>>(tmap, p);
    launch_count_add(1);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "stencil2_tma");
    set_kernel("stencil_tma<3D7, two steps>");
    return RACU_OK;
}

template <class T> static int
copy_faces(Buf3 const & src, Buf3 const & dst, int64_t const * n, cudaStream_t stream)
{
    int64_t const big = std::max(n[0]*n[1], std::max(n[0]*n[2], n[1]*n[2]));
    copy_faces_kernel<T><<<dim3(unsigned(std::min<int64_t>((big+255)/256, 1184)), 6), 256, 0, stream>>>(
        static_cast<T const *>(src.p), src.st[0], src.st[1], src.st[2], static_cast<T *>(dst.p), dst.st[0], dst.st[1], dst.st[2], int(n[0]), int(n[1]), int(n[2]));
    launch_count_add(1);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "copy_faces");
    return RACU_OK;
}

int racu_stencil_c(const racu_stencil_desc * s, void * stream); // (special.cc: racu_stencil)

template <class T> static int
stencil_steps_t(racu_stencil_desc const * s, int steps, void * scratch, cudaStream_t stream)
{
    int64_t const * n = s->len;
    Buf3 A {const_cast<void *>(s->src), {s->src_stride[0], s->src_stride[1], s->src_stride[2]}};
    Buf3 B {s->dst, {s->dst_stride[0], s->dst_stride[1], s->dst_stride[2]}};
    auto single = [&](Buf3 const & x, Buf3 const & y) {
        racu_stencil_desc d = *s;
        d.src = x.p; d.dst = y.p; d.lo0 = 0; d.hi0 = 0;
        for (int k=0; k<3; ++k) { d.src_stride[k] = x.st[k]; d.dst_stride[k] = y.st[k]; }
        return racu_stencil_c(&d, stream);
    };
    bool const fused = steps>=4 && tma_ok<T>(A, n) && tma_ok<T>(B, n) && !std::getenv("RACU_NO_TMA_STENCIL") && !std::getenv("RACU_NO_STENCIL2") && encode_fn();
    if (!fused) {
        for (int t=0; t<steps; ++t) { if (int e = (t&1) ? single(B, A) : single(A, B)) return e; }
        return RACU_OK;
    }
    // States t, t+2, ... live with the boundary of E, the odd ones with the boundary of O. The pairs ping-pong between E and a scratch buffer
    // C that carries E's boundary; the last two steps are single ones, so that on return BOTH buffers hold what T single steps leave:
    // state T in (T odd ? dst : src) and state T-1 in the other.
    Buf3 E = (steps&1) ? B : A, O = (steps&1) ? A : B;
    int64_t const bytes = ((n[0]-1)*E.st[0] + (n[1]-1)*E.st[1] + n[2])*int64_t(sizeof(T));
    if (scratch) { // a caller's scratch buffer is checked before anything is launched: it is a TMA source too, and must be a third buffer
        Buf3 const Cc {scratch, {E.st[0], E.st[1], E.st[2]}};
        if (!tma_ok<T>(Cc, n)) return fail(RACU_ERR_BAD_DESC, "racu_stencil_steps: the scratch buffer must be 16-byte aligned");
        auto overlaps = [&](Buf3 const & x) {
            int64_t const xb = ((n[0]-1)*x.st[0] + (n[1]-1)*x.st[1] + n[2])*int64_t(sizeof(T));
            uintptr_t const a0 = uintptr_t(scratch), a1 = a0+uintptr_t(bytes), b0 = uintptr_t(x.p), b1 = b0+uintptr_t(xb);
            return a0<b1 && b0<a1;
        };
        if (overlaps(A) || overlaps(B)) return fail(RACU_ERR_BAD_DESC, "racu_stencil_steps: the scratch buffer overlaps one of the arrays");
    }
    int left = steps;
    if (steps&1) { if (int e = single(A, B)) return e; --left; }
    void * own = nullptr;
    if (!scratch) {
        if (cudaError_t e = cudaMallocAsync(&own, size_t(bytes), stream)) return cuda_fail(e, "racu_stencil_steps scratch");
        scratch = own;
    }
    Buf3 C {scratch, {E.st[0], E.st[1], E.st[2]}};
    int rc = copy_faces<T>(E, C, n, stream);
    Buf3 X = E;
    for (int pair = 0; !rc && pair<left/2-1; ++pair) {
        Buf3 const & Y = X.p==E.p ? C : E;
        rc = launch_stencil2<T>(X, Y, O, n, stream);
        X = Y;
    }
    if (!rc) rc = single(X, O);
    if (!rc) rc = single(O, E);
    if (own) cudaFreeAsync(own, stream);
    return rc;
}

int
stencil_steps(racu_stencil_desc const * s, int steps, void * scratch, cudaStream_t stream)
{
    if (steps<0) return fail(RACU_ERR_BAD_DESC, "racu_stencil_steps: negative step count");
    if (s->kind!=RACU_STENCIL_3D7 || 3!=s->rank || (s->dtype!=RACU_F32 && s->dtype!=RACU_F64) || 0!=s->lo0 || 0!=s->hi0) {
        // any other sweep: single steps (racu_stencil validates)
        for (int t=0; t<steps; ++t) {
            racu_stencil_desc d = *s;
            if (t&1) { d.src = s->dst; d.dst = const_cast<void *>(s->src); for (int k=0; k<3; ++k) { d.src_stride[k] = s->dst_stride[k]; d.dst_stride[k] = s->src_stride[k]; } }
            if (int e = racu_stencil_c(&d, stream)) return e;
        }
        return RACU_OK;
    }
    size_t const es = s->dtype==RACU_F32 ? 4 : 8;
    if (views_overlap(s->src, 3, s->len, s->src_stride, s->dst, 3, s->len, s->dst_stride, es))
        return fail(RACU_ERR_UNSUPPORTED, "racu_stencil_steps: destination overlaps the source: traversal order would be observable");
    return s->dtype==RACU_F32 ? stencil_steps_t<float>(s, steps, scratch, stream) : stencil_steps_t<double>(s, steps, scratch, stream);
}

// One step of a slab-sharded 7-point sweep as ONE launch: sweep + halo push + neighbour lock step (see include/racu.h).
int
stencil_push(racu_stencil_desc const * s, racu_halo_push const * h, cudaStream_t stream)
{
    if (s->kind!=RACU_STENCIL_3D7 || 3!=s->rank) return fail(RACU_ERR_UNSUPPORTED, "racu_stencil_push: 3-D 7-point stencil only");
    if (s->dtype!=RACU_F32 && s->dtype!=RACU_F64) return fail(RACU_ERR_UNSUPPORTED, "stencil dtype must be f32 or f64");
    if (0!=s->lo0 || 0!=s->hi0) return fail(RACU_ERR_BAD_DESC, "racu_stencil_push sweeps the whole local interior");
    if (!h->flags || (h->push_lo && !h->signal_lo) || (h->push_hi && !h->signal_hi)) return fail(RACU_ERR_BAD_DESC, "racu_stencil_push: missing flag words");
    if (s->len[0]<3) return fail(RACU_ERR_BAD_DESC, "racu_stencil_push: a slab needs at least one owned plane between its ghosts");
    if (views_overlap(s->src, 3, s->len, s->src_stride, s->dst, 3, s->len, s->dst_stride, s->dtype==RACU_F32 ? 4 : 8))
        return fail(RACU_ERR_UNSUPPORTED, "racu_stencil_push: destination overlaps the source: traversal order would be observable");
    int handled = 0;
    int e = s->dtype==RACU_F32 ? launch_stencil<float, RACU_STENCIL_3D7>(s, stream, &handled, h) : launch_stencil<double, RACU_STENCIL_3D7>(s, stream, &handled, h);
    if (e) return e;
    if (!handled) return fail(RACU_ERR_UNSUPPORTED, "racu_stencil_push: the slab is not expressible as a TMA tensor map (innermost stride 1, 16-byte aligned rows)");
    return RACU_OK;
}

// examples/laplace2d.cc:48 as the host headers flatten it:
//   PUSH h [SQR] | PUSH c [CAST f64] PUSH s1 DIV | PUSH x0 (PUSH xk ADD) x 5 [CAST f64] PUSH s2 DIV | ADD MUL
// with the scalar arithmetic in f64, the six summed in the element type, and the six operands the views shifted by
// (-1,0) (0,-1) (+1,0) (0,+1) (+1,+1) (-1,-1), in this order.
static int
try_mass_stencil(racu_desc const * d, cudaStream_t stream, int * handled)
{
    if (std::getenv("RACU_NO_TMA_STENCIL")) return RACU_OK;
    racu_operand const & y = d->opnd[d->dst];
    int const T = y.dtype;
    if ((T!=RACU_F32 && T!=RACU_F64) || 2!=y.rank || y.kind!=RACU_MEM) return RACU_OK;
    racu_ins const * in = d->ins;
    int q = 0;
    int const n = d->nins;
    auto is = [&](int op, int type) { return q<n && in[q].op==op && in[q].type==type; };
    auto scalar = [&](double & v) {
        if (!is(RACU_OP_PUSH, RACU_F64)) return false;
        racu_operand const & o = d->opnd[in[q].arg];
        if (o.kind!=RACU_SCALAR || o.dtype!=RACU_F64) return false;
        v = o.base.f; ++q;
        return true;
    };
    auto view = [&](racu_operand const * & o) {
        if (!is(RACU_OP_PUSH, T)) return false;
        o = &d->opnd[in[q].arg];
        if (o->kind!=RACU_MEM || o->dtype!=T || 2!=o->rank) return false;
        ++q;
        return true;
    };
    double coef[3];
    racu_operand const * c = nullptr, * x[6];
    if (!scalar(coef[0])) return RACU_OK;
    if (is(RACU_OP_SQR, RACU_F64)) { coef[0] = coef[0]*coef[0]; ++q; }
    if (!view(c)) return RACU_OK;
    if (T==RACU_F32) { if (!is(RACU_OP_CAST, RACU_F64)) return RACU_OK; ++q; }
    if (!scalar(coef[1]) || !is(RACU_OP_DIV, RACU_F64)) return RACU_OK;
    ++q;
    for (int k=0; k<6; ++k) {
        if (!view(x[k])) return RACU_OK;
        if (k>0) { if (!is(RACU_OP_ADD, T)) return RACU_OK; ++q; }
    }
    if (T==RACU_F32) { if (!is(RACU_OP_CAST, RACU_F64)) return RACU_OK; ++q; }
    if (!scalar(coef[2]) || !is(RACU_OP_DIV, RACU_F64)) return RACU_OK;
    ++q;
    if (!is(RACU_OP_ADD, RACU_F64)) return RACU_OK;
    ++q;
    if (!is(RACU_OP_MUL, RACU_F64)) return RACU_OK;
    ++q;
    if (q!=n) return RACU_OK;
    static constexpr int sh[6][2] = {{-1, 0}, {0, -1}, {1, 0}, {0, 1}, {1, 1}, {-1, -1}};
    int64_t const es = T==RACU_F32 ? 4 : 8;
    for (int k=0; k<2; ++k) { if (y.len[k]!=c->len[k] || c->len[k]<1) return RACU_OK; }
    for (int k=0; k<6; ++k) {
        for (int a=0; a<2; ++a) { if (x[k]->len[a]!=c->len[a] || x[k]->stride[a]!=c->stride[a]) return RACU_OK; }
        int64_t const delta = static_cast<char *>(x[k]->base.ptr)-static_cast<char *>(c->base.ptr);
        if (delta!=(sh[k][0]*c->stride[0] + sh[k][1]*c->stride[1])*es) return RACU_OK;
    }
    racu_stencil_desc s;
    std::memset(&s, 0, sizeof(s));
    s.kind = ST_KIND_MASS; s.dtype = T; s.rank = 2;
    for (int k=0; k<2; ++k) { s.len[k] = c->len[k]+2; s.src_stride[k] = c->stride[k]; s.dst_stride[k] = y.stride[k]; }
    s.src = static_cast<char *>(c->base.ptr) - (c->stride[0]+c->stride[1])*es;
    s.dst = static_cast<char *>(y.base.ptr) - (y.stride[0]+y.stride[1])*es;
    if (views_overlap(s.src, 2, s.len, s.src_stride, s.dst, 2, s.len, s.dst_stride, size_t(es))) return RACU_OK;
    return T==RACU_F32 ? launch_stencil<float, ST_KIND_MASS>(&s, stream, handled, nullptr, coef) : launch_stencil<double, ST_KIND_MASS>(&s, stream, handled, nullptr, coef);
}

// racu_map pattern recognition: is this descriptor one of the reference's stencil slice expressions? (special.cc builds the
// canonical form; a descriptor flattened by a host header may order / share operands differently, so compare by meaning.)
int
try_special_map(racu_desc const * d, cudaStream_t stream, int * handled)
{
    *handled = 0;
    if (d->dst<0) return RACU_OK;
    // gemm as the reference writes it: c(all, insert<1>) += a * b(insert<1>)  (ra/ra.hh:407): axes (i, k, j)
    if (d->assign==RACU_ASSIGN_ADD && 3==d->nins && d->ins[0].op==RACU_OP_PUSH && d->ins[1].op==RACU_OP_PUSH && d->ins[2].op==RACU_OP_MUL) {
        racu_operand const & c = d->opnd[d->dst];
        racu_operand const & a = d->opnd[d->ins[0].arg];
        racu_operand const & b = d->opnd[d->ins[1].arg];
        int const T = c.dtype;
        if ((T==RACU_F32 || T==RACU_F64) && d->ins[2].type==T && a.dtype==T && b.dtype==T && a.kind==RACU_MEM && b.kind==RACU_MEM
            && 3==c.rank && 2==a.rank && 3==b.rank && RACU_UNB==c.len[1] && 0==c.stride[1] && RACU_UNB==b.len[0] && 0==b.stride[0]
            && 1==c.stride[2] && 1==a.stride[1] && 1==b.stride[2] && c.len[0]==a.len[0] && a.len[1]==b.len[1] && c.len[2]==b.len[2]
            && c.stride[0]>0 && a.stride[0]>0 && b.stride[1]>0) {
            racu_gemm_desc g;
            std::memset(&g, 0, sizeof(g));
            g.dtype = T; g.M = c.len[0]; g.N = c.len[2]; g.K = a.len[1];
            g.a = a.base.ptr; g.lda = a.stride[0]; g.b = b.base.ptr; g.ldb = b.stride[1]; g.c = c.base.ptr; g.ldc = c.stride[0];
            return try_special_gemm(&g, stream, handled); // (c overlapping a or b is not handled there: the planner then rejects it)
        }
    }
    if (d->assign!=RACU_ASSIGN_SET || d->nins<8) return RACU_OK;
    if (int e = try_mass_stencil(d, stream, handled)) return e;
    if (*handled) return RACU_OK;
    racu_operand const & y = d->opnd[d->dst];
    int const T = y.dtype, r = y.rank;
    if ((T!=RACU_F32 && T!=RACU_F64) || r<1 || r>3) return RACU_OK;
    // program: PUSH s(i32) CAST PUSH c MUL (PUSH x OP)*
    racu_ins const * in = d->ins;
    if (in[0].op!=RACU_OP_PUSH || in[1].op!=RACU_OP_CAST || in[1].type!=T || in[2].op!=RACU_OP_PUSH || in[3].op!=RACU_OP_MUL || in[3].type!=T) return RACU_OK;
    racu_operand const & sc = d->opnd[in[0].arg];
    racu_operand const & c = d->opnd[in[2].arg];
    if (sc.kind!=RACU_SCALAR || sc.dtype!=RACU_I32 || c.kind!=RACU_MEM || c.dtype!=T || c.rank!=r) return RACU_OK;
    int const nsh = (d->nins-4)/2;
    if (d->nins!=4+2*nsh || nsh!=2*r) return RACU_OK;
    int op = in[5].op;
    if (op!=RACU_OP_ADD && op!=RACU_OP_SUB) return RACU_OK;
    struct Sh { int axis, shift; };
    Sh want[6];
    int kind;
    if (3==r && op==RACU_OP_ADD && sc.base.i==-6) { kind = RACU_STENCIL_3D7; Sh w[6] = {{0, 1}, {1, 1}, {2, 1}, {0, -1}, {1, -1}, {2, -1}}; std::memcpy(want, w, sizeof(w)); }
    else if (2==r && op==RACU_OP_ADD && sc.base.i==-4) { kind = RACU_STENCIL_2D5; Sh w[4] = {{0, 1}, {1, 1}, {0, -1}, {1, -1}}; std::memcpy(want, w, sizeof(w)); }
    else if (2==r && op==RACU_OP_SUB && sc.base.i==4) { kind = RACU_STENCIL_2D5_STIFF; Sh w[4] = {{0, -1}, {1, -1}, {0, 1}, {1, 1}}; std::memcpy(want, w, sizeof(w)); }
    else if (1==r && op==RACU_OP_ADD && sc.base.i==-2) { kind = RACU_STENCIL_1D3; Sh w[2] = {{0, 1}, {0, -1}}; std::memcpy(want, w, sizeof(w)); }
    else return RACU_OK;
    size_t const es = T==RACU_F32 ? 4 : 8;
    for (int k=0; k<r; ++k) { if (y.len[k]!=c.len[k] || c.len[k]<1) return RACU_OK; }
    for (int q=0; q<nsh; ++q) {
        racu_ins const & pu = in[4+2*q], & o = in[5+2*q];
        if (pu.op!=RACU_OP_PUSH || o.op!=op || o.type!=T) return RACU_OK;
        racu_operand const & x = d->opnd[pu.arg];
        if (x.kind!=RACU_MEM || x.dtype!=T || x.rank!=r) return RACU_OK;
        for (int k=0; k<r; ++k) { if (x.len[k]!=c.len[k] || x.stride[k]!=c.stride[k]) return RACU_OK; }
        int64_t delta = (static_cast<char *>(x.base.ptr)-static_cast<char *>(c.base.ptr));
        if (delta!=int64_t(want[q].shift)*c.stride[want[q].axis]*int64_t(es)) return RACU_OK;
    }
    racu_stencil_desc s;
    std::memset(&s, 0, sizeof(s));
    s.kind = kind; s.dtype = T; s.rank = r;
    int64_t so = 0, dof = 0;
    for (int k=0; k<r; ++k) { s.len[k] = c.len[k]+2; s.src_stride[k] = c.stride[k]; s.dst_stride[k] = y.stride[k]; so += c.stride[k]; dof += y.stride[k]; }
    s.src = static_cast<char *>(c.base.ptr) - so*int64_t(es);
    s.dst = static_cast<char *>(y.base.ptr) - dof*int64_t(es);
    return try_special_stencil(&s, stream, handled);
}

} // namespace racu
