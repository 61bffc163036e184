// gemm.cu - dense contraction c(M,N) += a(M,K) * b(K,N), the only tensor-core path of the back end.
//
// Reference: gemm(a, b, c) is `c(all, insert<1>) += a * b(insert<1>)` (ra/ra.hh:403-412): a rank-3 traversal (i, k, j)
// whose destination has step 0 on k, i.e. per c[i,j] a sequential sum over ascending k of a[i,k]*b[k,j] with a plain
// `c += a*b` (no fma macro on this path). The order of a fold is unspecified (docs/ra-ra.texi:3102), so this kernel sums
// in tensor-core order; inputs whose products and sums are exactly representable (bench/bench-gemm.cc:229-237 uses
// integer-valued a = _0-_1, b = _1-2*_0) give bit-identical results, anything else is compared by tolerance.
//
//   f64: FP64 tensor-core MMA (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind, so DMMA is the only tensor path on sm_100).
//   f32: 3xTF32 - a = hi + lo with hi = tf32(a), lo = tf32(a - hi); acc += lo_a*hi_b + hi_a*lo_b + hi_a*hi_b on
//        mma.sync.m16n8k8.tf32 with fp32 accumulation. Dropped term lo*lo ~ 2^-22 |a||b|: error comparable to an fp32 FMA
//        chain (bound stated in DESIGN.md; tests compare against an fp64 reference).
//
// Tiling: CTA 128x128, 8 MMA warps as 2 (M) x 4 (N), warp tile 64x32, BK = 32, 3-stage cp.async pipeline, shared-memory pitches
// padded so that fragment reads are bank-conflict free. f64 (default form "w"): two more warps are the producers and the stages are
// handed over with mbarriers, so the MMA warps carry no copies, no address arithmetic and no CTA barrier (8192^3 on B200: 35.2 TFLOP/s
// against 32.2 for the form where every warp copies and multiplies, and 35.5 for cuBLAS; profiles/r02_dgemm_forms_8192.log).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "launch.h"
#include "special.h"

namespace racu {

constexpr int GM = 128, GN = 128, GSTAGES = 3, GTHREADS = 256;

__device__ __forceinline__ void cp16(void * s, void const * g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(uint32_t(__cvta_generic_to_shared(s))), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void
dmma884(double & d0, double & d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void
mma_tf32(float (&d)[4], uint32_t const (&a)[4], uint32_t const (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

struct GParams { void const * a; void const * b; void * c; int64_t lda, ldb, ldc; int32_t M, N, K; };

// ---------------- f64 ----------------
// (sm_100a has ONE FP64 tensor instruction, DMMA.8x8x4: mma.sync.m16n8k16.f64 compiles to eight of them, so the larger PTX shapes buy
// nothing. At one DMMA per 16 cycles and SM sub-partition the pipe peaks at 37.2 TFLOP/s at 1965 MHz; cuBLAS reaches 36.1.)
template <int BK> struct DTile
{
    static constexpr int PA = BK+4 /* A pitch (doubles) */, PB = GN+4 /* B pitch */;
    static constexpr int STAGE = GM*PA + BK*PB; // doubles per stage
};
constexpr int DK = 16;

template <int BK, int STAGES, bool SPREAD> __global__ void __launch_bounds__(GTHREADS)
dgemm_kernel(const __grid_constant__ GParams p)
{
    constexpr int DPA = DTile<BK>::PA, DPB = DTile<BK>::PB, DSTAGE = DTile<BK>::STAGE;
    constexpr int NA = GM*BK/2/GTHREADS, NB = BK*GN/2/GTHREADS; // 16-byte copies per thread and stage (A, B)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double * const sm = reinterpret_cast<double *>(smem_raw);
    int const tid = threadIdx.x, lane = tid&31, warp = tid>>5;
    int const wm = warp>>2, wn = warp&3;   // 2 x 4 warps
    int const g = lane>>2, t = lane&3;
    int64_t const m0 = int64_t(blockIdx.y)*GM, n0 = int64_t(blockIdx.x)*GN;
    double const * A = static_cast<double const *>(p.a) + m0*p.lda;
    double const * B = static_cast<double const *>(p.b) + n0;
    auto load_a = [&](int s, int kt, int i) { // A tile 128 x BK: BK/2 16-byte copies per row
        int c = tid + i*GTHREADS, r = c/(BK/2), q = c%(BK/2);
        cp16(sm + s*DSTAGE + r*DPA + q*2, A + r*p.lda + int64_t(kt)*BK + q*2);
    };
    auto load_b = [&](int s, int kt, int i) { // B tile BK x 128: 64 copies per row
        int c = tid + i*GTHREADS, r = c>>6, q = c&63;
        cp16(sm + s*DSTAGE + GM*DPA + r*DPB + q*2, B + (int64_t(kt)*BK+r)*p.ldb + q*2);
    };
    auto load_stage = [&](int s, int kt) {
#pragma unroll
        for (int i=0; i<NA; ++i) load_a(s, kt, i);
#pragma unroll
        for (int i=0; i<NB; ++i) load_b(s, kt, i);
    };
    double acc[8][4][2];
#pragma unroll
    for (int i=0; i<8; ++i)
#pragma unroll
        for (int j=0; j<4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    int const nk = p.K/BK;
#pragma unroll
    for (int s=0; s<STAGES-1; ++s) { if (s<nk) load_stage(s, s); cp_commit(); }
    for (int kt=0; kt<nk; ++kt) {
        cp_wait<STAGES-2>();
        __syncthreads();
        // SPREAD: the copies of the stage after next are issued a few at a time between the MMAs of this k-tile. Issued in one block
        // right after the barrier (round 1), the 8 warps of the CTA queue 64 KiB of LDGSTS on the load/store unit at the same moment,
        // and since a warp issues in order none of them reaches its first DMMA until its own copies have been accepted.
        bool const more = kt+STAGES-1<nk;
        int const ls = (kt+STAGES-1)%STAGES, lk = kt+STAGES-1;
        if constexpr (!SPREAD) { if (more) load_stage(ls, lk); cp_commit(); }
        double const * As = sm + (kt%STAGES)*DSTAGE + (wm*64)*DPA, * Bs = sm + (kt%STAGES)*DSTAGE + GM*DPA + wn*32;
#pragma unroll
        for (int kk=0; kk<BK; kk+=4) {
            double af[8], bf[4];
#pragma unroll
            for (int i=0; i<8; ++i) af[i] = As[(i*8+g)*DPA + kk + t];
#pragma unroll
            for (int j=0; j<4; ++j) bf[j] = Bs[(kk+t)*DPB + j*8 + g];
#pragma unroll
            for (int i=0; i<8; ++i) {
                if constexpr (SPREAD) { // NA + NB copies over BK/4 * 8 slots (a slot = 4 MMAs), evenly
                    constexpr int SLOTS = BK/4*8, EVERY = SLOTS/(NA+NB);
                    static_assert(EVERY>=1 && EVERY*(NA+NB)==SLOTS);
                    int const slot = (kk/4)*8 + i;
                    if (more && 0==slot%EVERY) { int const c = slot/EVERY; if (c<NA) load_a(ls, lk, c); else load_b(ls, lk, c-NA); }
                }
#pragma unroll
                for (int j=0; j<4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        if constexpr (SPREAD) cp_commit();
    }
    cp_wait<0>();
    double * Cp = static_cast<double *>(p.c) + (m0+wm*64)*p.ldc + n0 + wn*32;
#pragma unroll
    for (int i=0; i<8; ++i)
#pragma unroll
        for (int j=0; j<4; ++j) {
            double2 * dst = reinterpret_cast<double2 *>(Cp + (i*8+g)*p.ldc + j*8 + t*2);
            double2 c = *dst;
            c.x += acc[i][j][0]; c.y += acc[i][j][1];
            *dst = c;
        }
}

// Warp-specialised form: warps 0-7 only multiply (no barrier, no address arithmetic, no global loads in their instruction stream), warps 8
// and 9 are the producers (A and B tiles, cp.async into the same padded layout) and the stages are handed over with mbarriers:
// full[s] completes when the 64 producer lanes' copies of stage s have landed (cp.async.mbarrier.arrive.noinc), empty[s] when the 8
// consumer warps have read it.
constexpr int GWS_THREADS = GTHREADS+64;

__device__ __forceinline__ uint32_t g_smem_u32(void const * p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void g_mbar_init(uint64_t * bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(g_smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void g_mbar_arrive(uint64_t * bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(g_smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void g_mbar_arrive_cp(uint64_t * bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(g_smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void g_mbar_wait(uint64_t * bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GW_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GW_DONE;\n"
        "bra GW_LOOP;\n"
        "GW_DONE:\n"
        "}\n" :: "r"(g_smem_u32(bar)), "r"(parity) : "memory");
}

template <int BK, int STAGES> __global__ void __launch_bounds__(GWS_THREADS)
dgemm_ws_kernel(const __grid_constant__ GParams p)
{
    constexpr int DPA = DTile<BK>::PA, DPB = DTile<BK>::PB, DSTAGE = DTile<BK>::STAGE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double * const sm = reinterpret_cast<double *>(smem_raw);
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    int const tid = threadIdx.x, lane = tid&31, warp = tid>>5;
    int64_t const m0 = int64_t(blockIdx.y)*GM, n0 = int64_t(blockIdx.x)*GN;
    int const nk = p.K/BK;
    if (0==tid) {
#pragma unroll
        for (int s=0; s<STAGES; ++s) { g_mbar_init(&full[s], 64); g_mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp>=8) {
        bool const isb = 9==warp;
        double const * A = static_cast<double const *>(p.a) + m0*p.lda;
        double const * B = static_cast<double const *>(p.b) + n0;
        for (int kt=0; kt<nk; ++kt) {
            int const s = kt%STAGES;
            if (kt>=STAGES) g_mbar_wait(&empty[s], uint32_t((kt/STAGES-1)&1));
            double * As = sm + s*DSTAGE, * Bs = As + GM*DPA;
            int64_t const k0 = int64_t(kt)*BK;
            if (!isb) {
                // A tile 128 x BK: BK/2 copies per row; a warp instruction covers 32/(BK/2) rows
                constexpr int CPR = BK/2, RPI = 32/CPR;
                int const r0 = lane/CPR, q = lane%CPR;
                double const * src = A + int64_t(r0)*p.lda + k0 + q*2;
                double * dst = As + r0*DPA + q*2;
#pragma unroll 8
                for (int i=0; i<GM/RPI; ++i) { cp16(dst, src); src += int64_t(RPI)*p.lda; dst += RPI*DPA; }
            } else {
                // B tile BK x 128: 64 copies per row, two warp instructions per row
                double const * src = B + k0*p.ldb + lane*2;
                double * dst = Bs + lane*2;
#pragma unroll 8
                for (int r=0; r<BK; ++r) { cp16(dst, src); cp16(dst+64, src+64); src += p.ldb; dst += DPB; }
            }
            g_mbar_arrive_cp(&full[s]);
        }
        cp_wait_all();
        return;
    }
    int const wm = warp>>2, wn = warp&3;   // 2 x 4 consumer warps
    int const g = lane>>2, t = lane&3;
    double acc[8][4][2];
#pragma unroll
    for (int i=0; i<8; ++i)
#pragma unroll
        for (int j=0; j<4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    // (Double buffering the fragments by hand across the stage wait needs 24 more registers than the 168 a 10-warp CTA can have - three
    // warps share a sub-partition's 16 K registers - and spills; the k-tile boundary costs 0.2 % as it is: BK = 16 and 32 time the same.)
    for (int kt=0; kt<nk; ++kt) {
        int const s = kt%STAGES;
        g_mbar_wait(&full[s], uint32_t((kt/STAGES)&1));
        double const * As = sm + s*DSTAGE + (wm*64)*DPA, * Bs = sm + s*DSTAGE + GM*DPA + wn*32;
#pragma unroll
        for (int kk=0; kk<BK; kk+=4) {
            double af[8], bf[4];
#pragma unroll
            for (int i=0; i<8; ++i) af[i] = As[(i*8+g)*DPA + kk + t];
#pragma unroll
            for (int j=0; j<4; ++j) bf[j] = Bs[(kk+t)*DPB + j*8 + g];
#pragma unroll
            for (int i=0; i<8; ++i)
#pragma unroll
                for (int j=0; j<4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (0==lane) g_mbar_arrive(&empty[s]); // every lane's fragment reads of this stage fed MMAs that have been issued
    }
    double * Cp = static_cast<double *>(p.c) + (m0+wm*64)*p.ldc + n0 + wn*32;
#pragma unroll
    for (int i=0; i<8; ++i)
#pragma unroll
        for (int j=0; j<4; ++j) {
            double2 * dst = reinterpret_cast<double2 *>(Cp + (i*8+g)*p.ldc + j*8 + t*2);
            double2 c = *dst;
            c.x += acc[i][j][0]; c.y += acc[i][j][1];
            *dst = c;
        }
}

// ---------------- f32 via 3xTF32 ----------------
constexpr int SK = 32, SPA = SK+4 /* A pitch (floats) */, SPB = GN+8 /* B pitch */;
constexpr int SSTAGE = GM*SPA + SK*SPB;

__global__ void __launch_bounds__(GTHREADS)
sgemm3x_kernel(const __grid_constant__ GParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float * const sm = reinterpret_cast<float *>(smem_raw);
    int const tid = threadIdx.x, lane = tid&31, warp = tid>>5;
    int const wm = warp>>2, wn = warp&3;
    int const g = lane>>2, t = lane&3;
    int64_t const m0 = int64_t(blockIdx.y)*GM, n0 = int64_t(blockIdx.x)*GN;
    float const * A = static_cast<float const *>(p.a) + m0*p.lda;
    float const * B = static_cast<float const *>(p.b) + n0;
    auto load_stage = [&](int s, int kt) {
        float * As = sm + s*SSTAGE, * Bs = As + GM*SPA;
        int64_t const k0 = int64_t(kt)*SK;
#pragma unroll
        for (int i=0; i<4; ++i) { // A tile 128 x 32 floats: 8 copies per row
            int c = tid + i*GTHREADS, r = c>>3, q = c&7;
            cp16(As + r*SPA + q*4, A + r*p.lda + k0 + q*4);
        }
#pragma unroll
        for (int i=0; i<4; ++i) { // B tile 32 x 128 floats: 32 copies per row
            int c = tid + i*GTHREADS, r = c>>5, q = c&31;
            cp16(Bs + r*SPB + q*4, B + (k0+r)*p.ldb + q*4);
        }
    };
    float acc[4][4][4];
#pragma unroll
    for (int i=0; i<4; ++i)
#pragma unroll
        for (int j=0; j<4; ++j)
#pragma unroll
            for (int r=0; r<4; ++r) acc[i][j][r] = 0.f;
    int const nk = p.K/SK;
#pragma unroll
    for (int s=0; s<GSTAGES-1; ++s) { if (s<nk) load_stage(s, s); cp_commit(); }
    for (int kt=0; kt<nk; ++kt) {
        cp_wait<GSTAGES-2>();
        __syncthreads();
        if (kt+GSTAGES-1<nk) load_stage((kt+GSTAGES-1)%GSTAGES, kt+GSTAGES-1);
        cp_commit();
        float const * As = sm + (kt%GSTAGES)*SSTAGE + (wm*64)*SPA, * Bs = sm + (kt%GSTAGES)*SSTAGE + GM*SPA + wn*32;
#pragma unroll
        for (int kk=0; kk<SK; kk+=8) {
            uint32_t ah[4][4], al[4][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i=0; i<4; ++i) {
                float v[4] = { As[(i*16+g)*SPA + kk + t], As[(i*16+g+8)*SPA + kk + t], As[(i*16+g)*SPA + kk + t + 4], As[(i*16+g+8)*SPA + kk + t + 4] };
#pragma unroll
                for (int r=0; r<4; ++r) { ah[i][r] = to_tf32(v[r]); al[i][r] = to_tf32(v[r]-__uint_as_float(ah[i][r])); }
            }
#pragma unroll
            for (int j=0; j<4; ++j) {
                float v[2] = { Bs[(kk+t)*SPB + j*8 + g], Bs[(kk+t+4)*SPB + j*8 + g] };
#pragma unroll
                for (int r=0; r<2; ++r) { bh[j][r] = to_tf32(v[r]); bl[j][r] = to_tf32(v[r]-__uint_as_float(bh[j][r])); }
            }
#pragma unroll
            for (int i=0; i<4; ++i)
#pragma unroll
                for (int j=0; j<4; ++j) {
                    mma_tf32(acc[i][j], al[i], bh[j]); // small terms first
                    mma_tf32(acc[i][j], ah[i], bl[j]);
                    mma_tf32(acc[i][j], ah[i], bh[j]);
                }
        }
    }
    cp_wait<0>();
    float * Cp = static_cast<float *>(p.c) + (m0+wm*64)*p.ldc + n0 + wn*32;
#pragma unroll
    for (int i=0; i<4; ++i)
#pragma unroll
        for (int j=0; j<4; ++j) {
            float2 * d0 = reinterpret_cast<float2 *>(Cp + (i*16+g)*p.ldc + j*8 + t*2);
            float2 * d1 = reinterpret_cast<float2 *>(Cp + (i*16+g+8)*p.ldc + j*8 + t*2);
            float2 c0 = *d0, c1 = *d1;
            c0.x += acc[i][j][0]; c0.y += acc[i][j][1]; c1.x += acc[i][j][2]; c1.y += acc[i][j][3];
            *d0 = c0; *d1 = c1;
        }
}

int
try_special_gemm(racu_gemm_desc const * gd, cudaStream_t stream, int * handled)
{
    *handled = 0;
    if (std::getenv("RACU_NO_MMA_GEMM")) return RACU_OK;
    bool const f64 = gd->dtype==RACU_F64;
    int const bk = f64 ? DK : SK;
    size_t const es = f64 ? 8 : 4;
    {   // c overlapping a or b: not for these kernels (they read tiles of a and b after other CTAs have written c); the planner rejects it
        int64_t const lc[2] = {gd->M, gd->N}, sc[2] = {gd->ldc, 1}, la[2] = {gd->M, gd->K}, sa[2] = {gd->lda, 1}, lb[2] = {gd->K, gd->N}, sb[2] = {gd->ldb, 1};
        if (gd->M>0 && gd->N>0 && gd->K>0 && (views_overlap(gd->c, 2, lc, sc, gd->a, 2, la, sa, es) || views_overlap(gd->c, 2, lc, sc, gd->b, 2, lb, sb, es))) return RACU_OK;
    }
    if (gd->M<=0 || gd->N<=0 || gd->K<=0) { *handled = 1; return RACU_OK; } // nothing to add (test/reduction.cc:300-310: 0x0 corner)
    if (!f64) { // f32: tcgen05 / TMEM kernel first (gemm_tc5.cu, any shape); the mma.sync 3xTF32 kernel below is the second choice
        if (int e = try_tcgen05_sgemm(gd, stream, handled)) return e;
        if (*handled) return RACU_OK;
    }
    // tile multiples and 16-byte rows only; anything else runs as a fold on the ply kernels
    if (0!=gd->M%GM || 0!=gd->N%GN || 0!=gd->K%bk) return RACU_OK;
    if (gd->M>=(int64_t(1)<<31) || gd->N>=(int64_t(1)<<31) || gd->K>=(int64_t(1)<<31)) return RACU_OK;
    if (0!=(uintptr_t(gd->a)%16) || 0!=(uintptr_t(gd->b)%16) || 0!=(uintptr_t(gd->c)%16)) return RACU_OK;
    if (0!=(gd->lda*es)%16 || 0!=(gd->ldb*es)%16 || 0!=(gd->ldc*es)%16) return RACU_OK;
    GParams p {gd->a, gd->b, gd->c, gd->lda, gd->ldb, gd->ldc, int32_t(gd->M), int32_t(gd->N), int32_t(gd->K)};
    dim3 grid {unsigned(gd->N/GN), unsigned(gd->M/GM), 1};
    if (grid.y>65535) return RACU_OK;
    size_t smem = size_t(GSTAGES)*SSTAGE*4;
    cudaError_t e;
    if (f64) {
        // RACU_DGEMM = <form><BK>x<stages>: form "b" = every warp copies and multiplies, copies in one block after the barrier (round 1),
        // "s" = the same with the copies spread between the MMAs, "w" = warp specialised (two producer warps, mbarriers). BK 16 or 32, 3 or 4 stages.
        char const * which = std::getenv("RACU_DGEMM");
        char form = 'w'; int bkv = 32, st = 3;
        if (which && (which[0]=='b' || which[0]=='s' || which[0]=='w')) { form = which[0]; ++which; }
        if (which && 0==std::strcmp(which, "16x3")) { bkv = 16; st = 3; } else if (which && 0==std::strcmp(which, "16x4")) { bkv = 16; st = 4; }
        if (32==bkv && 0!=gd->K%32) { bkv = 16; st = 4; }
        using Kern = void (*)(GParams);
        Kern kernel; unsigned threads = GTHREADS;
        if ('w'==form) { kernel = 32==bkv ? dgemm_ws_kernel<32, 3> : 3==st ? dgemm_ws_kernel<16, 3> : dgemm_ws_kernel<16, 4>; threads = GWS_THREADS; }
        else if ('s'==form) kernel = 32==bkv ? dgemm_kernel<32, 3, true> : 3==st ? dgemm_kernel<16, 3, true> : dgemm_kernel<16, 4, true>;
        else kernel = 32==bkv ? dgemm_kernel<32, 3, false> : 3==st ? dgemm_kernel<16, 3, false> : dgemm_kernel<16, 4, false>;
        smem = 32==bkv ? size_t(3)*DTile<32>::STAGE*8 : size_t(st)*DTile<16>::STAGE*8;
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)))) return cuda_fail(e, "dgemm smem");
        kernel<<<grid, threads, smem, stream>>>(p);
        set_kernel("dgemm_dmma884");
    } else {
        if ((e = cudaFuncSetAttribute(sgemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)))) return cuda_fail(e, "sgemm smem");
        sgemm3x_kernel<<<grid, GTHREADS, smem, stream>>>(p);
        set_kernel("sgemm_3xtf32");
    }
    launch_count_add(1);
    if ((e = cudaGetLastError())) return cuda_fail(e, "gemm");
    *handled = 1;
    return RACU_OK;
}

} // namespace racu
