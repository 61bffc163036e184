// gemm_tc5.cu - f32 dense contraction c(M,N) += a(M,K) * b(K,N) on the 5th-generation tensor cores (sm_100a):
// 3xTF32 split on tcgen05.mma.kind::tf32 with the accumulator in TMEM, operands staged by TMA.
//
// Reference: gemm (ra/ra.hh:403-412) is `c(all, insert<1>) += a * b(insert<1>)`, a fold over k whose order is unspecified
// (docs/ra-ra.texi:3102). The float path must stay within fp32-level tolerance, so every product is done as
//     a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,   x_hi = x with the low 13 mantissa bits cleared (a TF32 value), x_lo = x - x_hi (exact)
// accumulated in fp32 in this order (small terms first). Dropped: a_lo*b_lo (~2^-22 |a||b|) and the TF32 truncation of x_lo.
//
// One CTA per 128x128 tile of c, 384 threads in three roles (one warp per role, not a warpgroup: tcgen05.mma is issued by ONE
// thread on behalf of the CTA):
//   warp 0    TMA producer: per k-block (32 k = one 128-byte swizzle row) the a tile [128 m][32 k] lands 128B-swizzled
//             (K-major, ready for the MMA) and the b tile [32 k][128 n] lands as it lies in memory (N contiguous).
//   warps 4-11 split: a -> a_hi (masked in place) and a_lo (same swizzled positions); b is TRANSPOSED on the way to K-major
//             b_hi / b_lo tiles [128 n][32 k] (16-byte chunk c of row n goes to chunk c ^ (n & 7): the 128B swizzle), then
//             fence.proxy.async + arrive. The same eight warps run the epilogue (warp w reads TMEM lanes 32*(w%4)..+31, half of the columns).
//   warp 1    MMA issuer: 4 k-steps (UMMA_K = 8) x 3 MMAs M128 N128 per k-block into 128 TMEM columns; tcgen05.commit hands the
//             stage back to the producer and, after the last k-block, the accumulator to the epilogue.
// Three 64 KiB operand stages (a_hi a_lo b_hi b_lo) + a two-slot ring for the raw b tile: the TMA runs up to three k-blocks ahead
// of the MMA, the split of k-block kb+1 overlaps the MMAs of kb. 1 CTA / SM. Tiles are rasterised in groups of 16 tile rows so that co-resident CTAs share a / b panels in L2.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "launch.h"
#include "special.h"

namespace racu {
namespace tc5 {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int BSTAGES = 2 /* raw b ring */, THREADS = 384, SPLIT_THREADS = 256, GROUP_M = 16;
constexpr uint32_t TILE = BM*BK*4;  // 16 KiB, every tile
// Two forms of the a operand:
//   TS = false  a_hi / a_lo are shared-memory tiles (tcgen05.mma SS form). Stage: a_hi (TMA target) | a_lo | b_hi | b_lo, 3 stages.
//   TS = true   a_hi / a_lo live in TENSOR MEMORY (TS form): the split warps read the raw a tile once and tcgen05.st both halves
//               into 2 x 32 TMEM columns per stage; the MMA then reads only b from shared memory, which halves the operand traffic
//               on the shared-memory port - the limiter of this kernel. Stage: a raw (TMA target) | b_hi | b_lo, 4 stages.
template <bool TS> struct Cfg
{
    static constexpr int STAGES = TS ? 4 : 3;
    static constexpr uint32_t OFF_AHI = 0, OFF_ALO = TILE /* SS only */, OFF_BHI = TS ? TILE : 2*TILE, OFF_BLO = OFF_BHI+TILE, STAGE_BYTES = OFF_BLO+TILE;
    static constexpr uint32_t OFF_BRAW = STAGES*STAGE_BYTES;
    static constexpr uint32_t SMEM_BYTES = STAGES*STAGE_BYTES + BSTAGES*TILE + 1024; // + slack for the manual 1024-byte alignment (swizzle atoms)
    static constexpr uint32_t TMEM_COLS = TS ? 512 : 128;   // accumulator: columns 0..127; TS: a_hi / a_lo of stage s at 128 + 64 s (+32)
};
// instruction descriptor (upper 32 bits of the 64-bit idesc): D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), both K-major, N>>3, M>>4
constexpr uint32_t IDESC = (1u<<4) | (2u<<7) | (2u<<10) | (uint32_t(BN>>3)<<17) | (uint32_t(BM>>4)<<24);

__device__ __forceinline__ uint32_t s32(void const * p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t * bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok!=0;
}
// A wait of more than ~2 s is a protocol bug: trap (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    long long const t0 = clock64();
    while (!mbar_try_wait(bar, parity)) { if (clock64()-t0>4000000000ll) __trap(); }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem, CUtensorMap const * tmap, int c0, int c1, uint64_t * bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem), "l"(tmap), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: start >> 4 | LBO (unused for swizzled K-major: 1) | SBO = 8 rows * 128 B | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return uint64_t((saddr & 0x3FFFFu)>>4) | (uint64_t(1)<<16) | (uint64_t(1024>>4)<<32) | (uint64_t(1)<<46) | (uint64_t(2)<<61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
// TS form: a from tensor memory (lane = row m, one tf32 per 32-bit column = k), b from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, uint32_t const (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                    "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split4(float4 v, float4 & hi, float4 & lo)
{
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    lo.x = v.x-hi.x; lo.y = v.y-hi.y; lo.z = v.z-hi.z; lo.w = v.w-hi.w;
}

template <bool TS> __global__ void __launch_bounds__(THREADS, 1)
sgemm3x_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float * C, int64_t ldc, int M, int N, int K, int vec_c, int mask_hi)
{
    using CF = Cfg<TS>;
    constexpr int STAGES = CF::STAGES;
    constexpr uint32_t OFF_AHI = CF::OFF_AHI, OFF_ALO = CF::OFF_ALO, OFF_BHI = CF::OFF_BHI, OFF_BLO = CF::OFF_BLO, STAGE_BYTES = CF::STAGE_BYTES,
                       OFF_BRAW = CF::OFF_BRAW, TMEM_COLS = CF::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    // a_full: TMA -> split, MMA.  b_full / b_empty: raw b ring, TMA <-> split.  ready: split -> MMA.  empty: MMA (commit) -> TMA.
    __shared__ __align__(8) uint64_t a_full[STAGES], ready[STAGES], empty[STAGES], b_full[BSTAGES], b_empty[BSTAGES], tmem_full;
    __shared__ uint32_t tmem_holder;
    uint32_t const base = (s32(smem_raw)+1023u) & ~1023u;          // shared-space address of stage 0
    unsigned char * const gen = smem_raw + (base-s32(smem_raw));   // the same place, generic pointer
    int const warp = threadIdx.x>>5, lane = threadIdx.x&31;
    int const nkb = (K+BK-1)/BK;
    // tile rasterisation: groups of GROUP_M tile rows, column-major inside a group
    int const tiles_m = (M+BM-1)/BM, tiles_n = (N+BN-1)/BN;
    int const width = GROUP_M*tiles_n, group = int(blockIdx.x)/width, first_m = group*GROUP_M;
    int const gsz = min(tiles_m-first_m, GROUP_M);
    int const m0 = (first_m + (int(blockIdx.x)%width)%gsz)*BM, n0 = ((int(blockIdx.x)%width)/gsz)*BN;

    if (0==threadIdx.x) {
#pragma unroll
        for (int s=0; s<STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&ready[s], SPLIT_THREADS); mbar_init(&empty[s], 1); }
#pragma unroll
        for (int r=0; r<BSTAGES; ++r) { mbar_init(&b_full[r], 1); mbar_init(&b_empty[r], SPLIT_THREADS); }
        mbar_init(&tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (1==warp) { // one warp allocates the accumulator columns and owns the deallocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(&tmem_holder)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t const tmem = tmem_holder;

    if (0==warp) {
        if (0==lane) { // ---- TMA producer
            for (int kb=0; kb<nkb; ++kb) {
                int const s = kb%STAGES, r = kb%BSTAGES;
                mbar_wait(&b_empty[r], (uint32_t(kb/BSTAGES)&1u)^1u); // the split warps have read this raw b slot
                mbar_expect_tx(&b_full[r], TILE);
                tma_load_2d(base + OFF_BRAW + r*TILE, &tmB, n0, kb*BK, &b_full[r]);
                mbar_wait(&empty[s], (uint32_t(kb/STAGES)&1u)^1u);    // the MMAs that read this operand stage have completed
                mbar_expect_tx(&a_full[s], TILE);
                tma_load_2d(base + s*STAGE_BYTES + OFF_AHI, &tmA, kb*BK, m0, &a_full[s]);
            }
        }
    } else if (1==warp) {
        if (0==lane) { // ---- MMA issuer
            for (int kb=0; kb<nkb; ++kb) {
                int const s = kb%STAGES;
                uint32_t const ph = uint32_t(kb/STAGES)&1u;
                mbar_wait(&a_full[s], ph);
                mbar_wait(&ready[s], ph);                             // hi / lo tiles written and fenced by the split warps
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t const st = base + s*STAGE_BYTES;
                uint64_t const bhi = umma_desc(st+OFF_BHI), blo = umma_desc(st+OFF_BLO);
                if constexpr (TS) {
                    uint32_t const ahi = tmem + 128u + uint32_t(s)*64u, alo = ahi + 32u;
#pragma unroll
                    for (int k4=0; k4<BK/8; ++k4) {                   // a: +8 TMEM columns per k-step; b: +32 bytes inside the swizzle row
                        uint64_t const o = uint64_t(k4*2);
                        umma_tf32_ts(tmem, alo+8u*k4, bhi+o, (kb|k4)>0);
                        umma_tf32_ts(tmem, ahi+8u*k4, blo+o, 1);
                        umma_tf32_ts(tmem, ahi+8u*k4, bhi+o, 1);
                    }
                } else {
                    uint64_t const ahi = umma_desc(st+OFF_AHI), alo = umma_desc(st+OFF_ALO);
#pragma unroll
                    for (int k4=0; k4<BK/8; ++k4) {                   // UMMA_K = 8 tf32 = 32 bytes inside the swizzle row: +2 in the (>>4) address field
                        uint64_t const o = uint64_t(k4*2);
                        umma_tf32(tmem, alo+o, bhi+o, (kb|k4)>0);
                        umma_tf32(tmem, ahi+o, blo+o, 1);
                        umma_tf32(tmem, ahi+o, bhi+o, 1);
                    }
                }
                umma_commit(&empty[s]);                               // (implies tcgen05.fence::before_thread_sync)
            }
            umma_commit(&tmem_full);
        }
    } else if (warp>=4) {
        int const t = int(threadIdx.x)-(THREADS-SPLIT_THREADS);      // 0..255
        int const bn = t&(BN-1), bh = t>>7;                          // b: row n of the K-major tiles, and which half of its 8 chunks
        for (int kb=0; kb<nkb; ++kb) { // ---- split
            int const s = kb%STAGES, r = kb%BSTAGES;
            mbar_wait(&a_full[s], uint32_t(kb/STAGES)&1u);           // (a_full follows empty[s]: a_lo / b_hi / b_lo of this stage are free too)
            unsigned char * const st = gen + s*STAGE_BYTES;
            if constexpr (TS) {
                // a: this thread owns row m = 32 (warp % 4) + lane (its TMEM lane) and 16 of the 32 k: four 16-byte chunks out of the
                // swizzled tile (chunk c of row m sits at c ^ (m & 7): conflict free across a quarter warp), hi and lo to TMEM
                int const m = (warp&3)*32 + lane, h = (warp-4)>>2;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c=0; c<4; ++c) {
                    float4 v = *reinterpret_cast<float4 const *>(st+OFF_AHI + m*128 + (((4*h+c) ^ (m&7))<<4)), vh, vl;
                    split4(v, vh, vl);
                    hi[4*c] = __float_as_uint(v.x); hi[4*c+1] = __float_as_uint(v.y); hi[4*c+2] = __float_as_uint(v.z); hi[4*c+3] = __float_as_uint(v.w);
                    lo[4*c] = __float_as_uint(vl.x); lo[4*c+1] = __float_as_uint(vl.y); lo[4*c+2] = __float_as_uint(vl.z); lo[4*c+3] = __float_as_uint(vl.w);
                }
                uint32_t const ta = tmem + (uint32_t((warp&3)*32)<<16) + 128u + uint32_t(s)*64u + uint32_t(16*h);
                tmem_st16(ta, hi);
                tmem_st16(ta+32u, lo);
            } else
            // a: hi in place, lo at the same (swizzled) position of its own tile
#pragma unroll
            for (int j=0; j<int(TILE/16)/SPLIT_THREADS; ++j) {
                int const idx = (t + SPLIT_THREADS*j)*16;
                float4 v = *reinterpret_cast<float4 const *>(st+OFF_AHI+idx), hi, lo;
                split4(v, hi, lo);
                if (mask_hi) *reinterpret_cast<float4 *>(st+OFF_AHI+idx) = hi; // default off: kind::tf32 ignores the low 13 mantissa bits itself (measured: identical results)
                *reinterpret_cast<float4 *>(st+OFF_ALO+idx) = lo;
            }
            // b: staging [BK k][BN n] -> K-major [BN n][BK k] with the 128-byte swizzle; this thread owns half of row n = t % 128
            mbar_wait(&b_full[r], uint32_t(kb/BSTAGES)&1u);
            float const * const bs = reinterpret_cast<float const *>(gen+OFF_BRAW+r*TILE) + bn;
#pragma unroll
            for (int kc=bh*(BK/8); kc<(bh+1)*(BK/8); ++kc) {
                float4 v {bs[(4*kc+0)*BN], bs[(4*kc+1)*BN], bs[(4*kc+2)*BN], bs[(4*kc+3)*BN]}, hi, lo;
                split4(v, hi, lo);
                int const off = bn*128 + ((kc ^ (bn&7))<<4);
                *reinterpret_cast<float4 *>(st+OFF_BHI+off) = mask_hi ? hi : v;
                *reinterpret_cast<float4 *>(st+OFF_BLO+off) = lo;
            }
            mbar_arrive(&b_empty[r]);
            if constexpr (TS) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> tensor-core (async proxy) reads
            mbar_arrive(&ready[s]);
        }
        // ---- epilogue: TMEM -> registers -> c += acc
        mbar_wait(&tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int const q = warp&3, half = (warp-4)>>2;                    // TMEM lane quarter of this warp (warp % 4), column half
        int const row = m0 + q*32 + lane;
#pragma unroll 1
        for (int c0=half*(BN/2); c0<(half+1)*(BN/2); c0+=32) {
            uint32_t v[32];
            tmem_ld32(tmem + (uint32_t(q*32)<<16) + uint32_t(c0), v);
            if (row<M) {
                float * crow = C + int64_t(row)*ldc + n0 + c0;
                if (vec_c && n0+c0+32<=N) {
#pragma unroll
                    for (int j=0; j<32; j+=4) {
                        float4 c = *reinterpret_cast<float4 *>(crow+j);
                        c.x += __uint_as_float(v[j]); c.y += __uint_as_float(v[j+1]); c.z += __uint_as_float(v[j+2]); c.w += __uint_as_float(v[j+3]);
                        *reinterpret_cast<float4 *>(crow+j) = c;
                    }
                } else {
#pragma unroll
                    for (int j=0; j<32; ++j) { if (n0+c0+j<N) crow[j] += __uint_as_float(v[j]); }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (1==warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
}

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

} // namespace tc5

// f32 contraction on tcgen05; sets *handled when it ran. Needs 16-byte aligned a / b with row pitches that are multiples of
// 16 bytes (TMA); any M, N, K (tails are zero-filled on load and masked on store).
int
try_tcgen05_sgemm(racu_gemm_desc const * gd, cudaStream_t stream, int * handled)
{
    using namespace tc5;
    *handled = 0;
    if (std::getenv("RACU_NO_TCGEN05")) return RACU_OK;
    if (gd->dtype!=RACU_F32) return RACU_OK;
    if (gd->M>=(int64_t(1)<<31) || gd->N>=(int64_t(1)<<31) || gd->K>=(int64_t(1)<<31)) return RACU_OK;
    if (0!=(uintptr_t(gd->a)%16) || 0!=(uintptr_t(gd->b)%16) || 0!=(gd->lda*4)%16 || 0!=(gd->ldb*4)%16 || gd->lda<gd->K || gd->ldb<gd->N) return RACU_OK;
    EncodeTiled enc = encode_fn();
    if (!enc) return RACU_OK;
    CUtensorMap tmA, tmB;
    {
        cuuint64_t gdim[2] = {cuuint64_t(gd->K), cuuint64_t(gd->M)}, gstr[1] = {cuuint64_t(gd->lda*4)};
        cuuint32_t box[2] = {BK, BM}, estr[2] = {1, 1};
        if (CUDA_SUCCESS!=enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(gd->a), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return RACU_OK;
    }
    {
        cuuint64_t gdim[2] = {cuuint64_t(gd->N), cuuint64_t(gd->K)}, gstr[1] = {cuuint64_t(gd->ldb*4)};
        cuuint32_t box[2] = {BN, BK}, estr[2] = {1, 1};
        if (CUDA_SUCCESS!=enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(gd->b), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return RACU_OK;
    }
    int64_t const tiles = ((gd->M+BM-1)/BM)*((gd->N+BN-1)/BN);
    if (tiles>=(int64_t(1)<<31)) return RACU_OK;
    int const vec_c = 0==(uintptr_t(gd->c)%16) && 0==(gd->ldc*4)%16;
    cudaError_t e;
    int const mask = std::getenv("RACU_TC5_MASK") ? 1 : 0;
    if (std::getenv("RACU_TC5_SS")) { // a operand from shared memory (the first version of this kernel; kept for comparison)
        if ((e = cudaFuncSetAttribute(sgemm3x_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<false>::SMEM_BYTES)))) return cuda_fail(e, "sgemm tcgen05 smem");
        sgemm3x_tc5_kernel<false><<<unsigned(tiles), THREADS, Cfg<false>::SMEM_BYTES, stream>>>(tmA, tmB, static_cast<float *>(gd->c), gd->ldc, int(gd->M), int(gd->N), int(gd->K), vec_c, mask);
    } else {
        if ((e = cudaFuncSetAttribute(sgemm3x_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<true>::SMEM_BYTES)))) return cuda_fail(e, "sgemm tcgen05 smem");
        sgemm3x_tc5_kernel<true><<<unsigned(tiles), THREADS, Cfg<true>::SMEM_BYTES, stream>>>(tmA, tmB, static_cast<float *>(gd->c), gd->ldc, int(gd->M), int(gd->N), int(gd->K), vec_c, mask);
    }
    launch_count_add(1);
    if ((e = cudaGetLastError())) return cuda_fail(e, "sgemm tcgen05");
    set_kernel("sgemm_3xtf32_tcgen05");
    *handled = 1;
    return RACU_OK;
}

} // namespace racu
