/* racu.h - C ABI of the B200 execution back end for ra-ra's traversal hot path.
 *
 * The reference (lloda/ra-ra, header-only C++23) has no FFI. Its de-facto operator interface for
 * this path is the single traversal entry `ra::ply(Iterator&&)` (ra/ply.hh:157-166) that every
 * assignment (ra/expr.hh:50-60), reduction (ra/ra.hh:279-401) and gemm/gemv (ra/ra.hh:403-456)
 * funnels into, plus the ten-method Iterator concept (ra/base.hh:298-311) that expression nodes
 * implement. This header is the seam introduced AT `ply`: a flattened expression tree
 * (racu_desc) crosses it and is executed by hand-written sm_100a kernels.
 *
 * Plain C: pointers, sizes, POD structs, int status returns. No torch / C++ types.
 *
 *   reference interface replaced                          entry point here
 *   ----------------------------------------------------  ---------------------------------
 *   ply(map(op, dst, src...)) / dst OP= expr               racu_map
 *     ra/ply.hh:21-86,157-168; ra/expr.hh:50-60
 *   Match::check / agree (prefix shape agreement)          racu_agree
 *     ra/expr.hh:529-608,639
 *   sum prod amin amax dot reduce_sqrm                     racu_reduce / racu_reduce_dev
 *     ra/ra.hh:306-322,348-401
 *   Anext(I,J,K) = -6*A(I,J,K) + A(I+1,J,K)... ; swap      racu_stencil (also reached from racu_map
 *     bench/bench-stencil3.cc:55-64, bench-stencil2.cc:49-58,    by pattern recognition); racu_stencil_steps
 *                                                                = the benchmark's loop of T such steps
 *     examples/laplace2d.cc:36
 *   gemm / gemv / gevm                                     racu_gemm / racu_gemv
 *     ra/ra.hh:403-456
 *   Array(shape, none|value) storage, default_init         racu_alloc / racu_free / racu_memcpy_* / racu_fill
 *     ra/arrays.hh:224-327
 *   a(i ...) / at(a, i) with index arrays, both as         racu_map with a RACU_PTR operand + RACU_OP_LOAD (gather)
 *     rvalue and lvalue: ra/ply.hh:461-499,531-549;          or a RACU_PTR destination (scatter)
 *     ra/ra.hh:228-250
 *   the compile-time instantiation of ply_ravel for an     racu_prepare (plan + specialise the kernel of a descriptor
 *     expression type (ra/ply.hh:157-166)                    ahead of its first use; needs no GPU)
 *   (none: the reference is single process)                racu_comm_* (NCCL over NVLink), racu_allreduce,
 *                                                          racu_halo_exchange, racu_ipc_*, racu_stencil_push
 */
#ifndef RACU_H
#define RACU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RACU_VERSION 1
#define RACU_MAX_RANK 8    /* the reference's own examples reach expression rank 6 (examples/maxwell.cc:34) */
#define RACU_MAX_OPND 12
#define RACU_MAX_INS 48

/* ra/base.hh:232-234. UNB marks a dead / unbounded axis (insert<n>, tindex, unbounded iota). */
#define RACU_ANY ((int64_t)-1944444444)
#define RACU_UNB ((int64_t)-1988888888)
#define RACU_MIS ((int64_t)-1922222222)

typedef enum {
    RACU_OK = 0,
    RACU_ERR_BAD_DESC = 1,     /* malformed descriptor / program */
    RACU_ERR_SHAPE = 2,        /* operands do not agree: RA_CK "Bad shapes" ra/expr.hh:562; or unbounded */
    RACU_ERR_UNSUPPORTED = 3,  /* well formed, outside what the kernels implement */
    RACU_ERR_CUDA = 4,
    RACU_ERR_NCCL = 5,
    RACU_ERR_ALLOC = 6,
    RACU_ERR_BOUNDS = 7
} racu_status;

typedef enum {
    RACU_BOOL = 0, /* 1 byte, 0/1 (C++ bool) */
    RACU_I32 = 1,
    RACU_I64 = 2,
    RACU_F32 = 3,
    RACU_F64 = 4,
    RACU_NDTYPE = 5
} racu_dtype;

typedef enum {
    RACU_MEM = 0,    /* Cell over memory: base.ptr + sum idx[k]*stride[k] (elements). ra/expr.hh:266-312 */
    RACU_SEQ = 1,    /* Cell over Seq (iota, tindex): value = base + T(sum idx[k]*stride[k]). ra/expr.hh:79-100 */
    RACU_SCALAR = 2, /* Scalar<C>: rank 0, step 0. ra/expr.hh:104-120 */
    RACU_PTR = 3     /* An array addressed by COMPUTED element offsets (RACU_OP_LOAD): the unbeaten subscripts A(I, J) of
                        ra/ply.hh:461-499 and at(A, I) (ra/ra.hh:228-250). rank 0, takes no part in shape agreement.
                        base.ptr = element at offset 0; len[0] / len[1] = smallest / largest valid element offset. Offsets
                        are clamped into that range (the reference asserts on a bad index, RA_CK ra/ply.hh:395). */
} racu_kind;

typedef union {
    void * ptr;
    int64_t i;
    double f;
} racu_base;

/* One leaf of the expression tree. Axes k >= rank have step 0 (Cell::step, ra/expr.hh:294-295) so a
 * lower rank operand repeats along TRAILING axes (prefix matching, docs/ra-ra.texi:633-681).
 * len[k] may be RACU_UNB (dead axis: insert<n> ra/ply.hh:423-427, Reframe ra/expr.hh:425-436). */
typedef struct {
    int32_t kind;   /* racu_kind */
    int32_t dtype;  /* racu_dtype of the element */
    int32_t rank;
    int32_t reserved;
    racu_base base; /* MEM: device pointer to element [0,..,0]. SEQ: origin (i for integer dtypes, f for float). SCALAR: value */
    int64_t len[RACU_MAX_RANK];
    int64_t stride[RACU_MAX_RANK]; /* in elements; 0 = broadcast, <0 = reversed */
} racu_operand;

/* Postfix op program. Every arithmetic instruction works in ONE type: the flattener inserts explicit
 * CASTs where C++'s usual arithmetic conversions would apply (float+double -> double, int*float ->
 * float, ...), so `type` is both the operand and the result type, except comparisons and NOT, whose
 * result is RACU_BOOL. */
typedef enum {
    RACU_OP_PUSH = 0,  /* arg = operand index; type = its dtype */
    RACU_OP_CAST = 1,  /* aux = source type; type = destination type. ra::cast ra/expr.hh:680 */
    /* unary T -> T */
    RACU_OP_NEG = 2,
    RACU_OP_ABS = 3,
    RACU_OP_SQRT = 4,
    RACU_OP_SQR = 5,   /* sqr / sqrm(x) ra/ra.hh:65,71 */
    RACU_OP_EXP = 6,
    RACU_OP_LOG = 7,
    RACU_OP_SIN = 8,
    RACU_OP_COS = 9,
    RACU_OP_NOT = 10,  /* T -> BOOL */
    /* binary T,T -> T */
    RACU_OP_ADD = 16,
    RACU_OP_SUB = 17,
    RACU_OP_MUL = 18,
    RACU_OP_DIV = 19,
    RACU_OP_MOD = 20,  /* integer % ; fmod for floats */
    RACU_OP_MIN = 21,  /* std::min(a,b) = b<a ? b : a   (ra/ra.hh:216) */
    RACU_OP_MAX = 22,  /* std::max(a,b) = a<b ? b : a */
    RACU_OP_POW = 23,
    RACU_OP_ATAN2 = 24,
    RACU_OP_BAND = 25,
    RACU_OP_BOR = 26,
    RACU_OP_BXOR = 27,
    /* binary T,T -> BOOL */
    RACU_OP_EQ = 32,
    RACU_OP_NE = 33,
    RACU_OP_LT = 34,
    RACU_OP_LE = 35,
    RACU_OP_GT = 36,
    RACU_OP_GE = 37,
    /* other */
    RACU_OP_FMA = 40,  /* a b c -> fma(a,b,c) */
    RACU_OP_PICK = 41, /* sel x0 .. x(n-1) -> x[sel]; arg = n, aux = selector type, type = T of branches.
                          ra::pick ra/expr.hh:686-728; where(w,t,f) = pick(bool(w), f, t) ra/ra.hh:255-256.
                          Only the chosen branch is observable: the opcode set is side-effect free. */
    RACU_OP_LOAD = 42, /* off -> opnd[arg][off]: gather. arg = index of a RACU_PTR operand, type = its element dtype,
                          aux = integer type of the offset (I32 or I64). The host writes A(I, J) as
                          LOAD(A, I*step0 + J*step1) with I and J placed on their own expression axes (ra/ply.hh:461-499). */
    RACU_OP_CLAMP = 43, /* v lo hi -> std::clamp(v, lo, hi) = v<lo ? lo : hi<v ? hi : v   (ra/ra.hh:222), any arithmetic T */
    RACU_OP_LERP = 44,  /* a b t -> std::lerp(a, b, t) (ra/ra.hh:222), float T: exact at t = 0 and t = 1, monotonic */
    /* the rest of the named functions of ra/ra.hh:221-222. unary float T -> T */
    RACU_OP_TAN = 48, RACU_OP_SINH = 49, RACU_OP_COSH = 50, RACU_OP_TANH = 51, RACU_OP_ASIN = 52, RACU_OP_ACOS = 53, RACU_OP_ATAN = 54,
    RACU_OP_EXPM1 = 55, RACU_OP_LOG1P = 56, RACU_OP_LOG10 = 57,
    /* unary float T -> BOOL */
    RACU_OP_ISFINITE = 58, RACU_OP_ISNAN = 59, RACU_OP_ISINF = 60
    /* odd(x), rel_error(a, b) (ra/ra.hh:61,94-95,210) and `<=>` (ra/ra.hh:190) need no opcode: the host headers write them with
       the ones above ((x & 1) != 0; D = |a|+|b|, D==0 ? 0. : 2.*|a-b|/D; a<b ? -1 : a>b ? 1 : a==b ? 0 : 2 (unordered)). */
} racu_opcode;

typedef struct {
    uint8_t op;   /* racu_opcode */
    uint8_t type; /* racu_dtype */
    uint8_t arg;
    uint8_t aux;
} racu_ins;

/* `dst OP= program(srcs)`. ra/expr.hh:50-53: every assignment is for_each([](y, x){ y OP= x; }, dst, rhs)
 * and the destination takes part in shape agreement like any operand. When dst has step 0 along an
 * expression axis longer than 1, the sequential reference turns the assignment into a fold
 * (c(m) += a(m,n) sums over n: docs/ra-ra.texi:684-690, test/reduction.cc:155-158); the device does it
 * with a deterministic tree. AMIN/AMAX are the combiners of ra::amin/amax (ra/ra.hh:306-322). */
typedef enum {
    RACU_ASSIGN_SET = 0,
    RACU_ASSIGN_ADD = 1,
    RACU_ASSIGN_SUB = 2,
    RACU_ASSIGN_MUL = 3,
    RACU_ASSIGN_DIV = 4,
    RACU_ASSIGN_AMIN = 5, /* if (x<y) y=x */
    RACU_ASSIGN_AMAX = 6  /* if (y<x) y=x */
} racu_assign;

typedef struct {
    int32_t nopnd;
    int32_t nins;
    int32_t dst;    /* index of the destination operand (kind MEM), or -1 for racu_reduce. A RACU_PTR destination is a SCATTER
                       (a(i ...) OP= x with index arrays, ra/ply.hh:461-499 used as an lvalue): the program then leaves TWO values,
                       [element offset, value in the common type of dst and value]; repeated offsets combine atomically for
                       += -= *= /= (in an unspecified order) and leave any one of the values for = */
    int32_t assign; /* racu_assign */
    racu_operand opnd[RACU_MAX_OPND];
    racu_ins ins[RACU_MAX_INS];
} racu_desc;

/* ---- library ---- */
int racu_version(void);
/* Thread-local text for the last non-OK status returned on this thread. */
const char * racu_last_error_string(void);
int racu_device_count(void);
int racu_set_device(int dev);

/* ---- storage ---- */
int racu_alloc(void ** ptr, size_t bytes);
int racu_free(void * ptr);
int racu_memcpy_h2d(void * dst_dev, const void * src_host, size_t bytes, void * stream);
int racu_memcpy_d2h(void * dst_host, const void * src_dev, size_t bytes, void * stream);
int racu_memcpy_d2d(void * dst_dev, const void * src_dev, size_t bytes, void * stream);
int racu_fill(void * dst_dev, int dtype, racu_base value, size_t count, void * stream);
int racu_stream_sync(void * stream);
/* Pinned host staging used by the host header for std::vector operands. */
int racu_host_alloc(void ** ptr, size_t bytes);
int racu_host_free(void * ptr);

/* ---- traversal ---- */
/* Shape agreement only (Match::len / check). rank_out and len_out (RACU_MAX_RANK entries) may be NULL. */
int racu_agree(const racu_desc * d, int32_t * rank_out, int64_t * len_out);
/* dst OP= program. Asynchronous on `stream` (a cudaStream_t, NULL = default stream). */
int racu_map(const racu_desc * d, void * stream);

/* ---- whole-array reductions of a fused map (d->dst must be -1) ---- */
typedef enum {
    RACU_RED_SUM = 1,  /* sum, dot, reduce_sqrm: start T(0), c += x */
    RACU_RED_PROD = 3, /* start T(1), c *= x */
    RACU_RED_AMIN = 5, /* start +inf / max, if (x<c) c=x : NaN never enters */
    RACU_RED_AMAX = 6
} racu_redop;
/* Result has the program's result type. out_host receives it after a stream synchronise. */
int racu_reduce(const racu_desc * d, int redop, void * out_host, void * stream);
/* Same, result left in device memory (asynchronous). */
int racu_reduce_dev(const racu_desc * d, int redop, void * out_dev, void * stream);
/* Same, asynchronous, result written by the device into page-locked HOST memory from racu_host_alloc (valid after a
 * racu_stream_sync of `stream`): a chain of reductions costs one synchronise, not one per scalar. */
int racu_reduce_async(const racu_desc * d, int redop, void * out_pinned, void * stream);
/* Plan d exactly as racu_map would (redop == 0) or racu_reduce (redop != 0) and make sure its kernel exists, WITHOUT
 * launching anything. A program that is not in the pre-compiled table is specialised now (NVRTC) and cached in memory and
 * on disk ($RACU_JIT_CACHE, default: jit_cache/ next to the library), so the first real call does not pay for it. Needs
 * no GPU: on a build host only the disk cache is filled. *kernel_out (optional) = the kernel family that will run.
 * The device analogue of the reference instantiating ply_ravel for an expression type at compile time. */
int racu_prepare(const racu_desc * d, int redop, const char ** kernel_out);

/* Kernel configuration introspection for tests/bench: how many of this library's kernels were
 * launched by this thread since the last reset. */
int64_t racu_launch_count(void);
void racu_launch_count_reset(void);
/* Name of the kernel family the last racu_map / racu_reduce on this thread dispatched to. */
const char * racu_last_kernel(void);
/* Plans are cached per thread (key: descriptor + destination + device); counters of this thread's lookups. */
void racu_plan_cache_stats(int64_t * hits, int64_t * misses);

/* ---- stencils ---- */
typedef enum {
    RACU_STENCIL_3D7 = 1, /* bench/bench-stencil3.cc:59-61  -6*c + (i+1) + (j+1) + (k+1) + (i-1) + (j-1) + (k-1) */
    RACU_STENCIL_2D5 = 2, /* bench/bench-stencil2.cc:53-55  -4*c + (i+1) + (j+1) + (i-1) + (j-1) */
    RACU_STENCIL_1D3 = 3, /* bench/bench-stencil1.cc:45     -2*c + (i+1) + (i-1) */
    RACU_STENCIL_2D5_STIFF = 4 /* examples/laplace2d.cc:36  4*c - (i-1) - (j-1) - (i+1) - (j+1) */
} racu_stencil_kind;

/* One sweep over the interior [1, len-1) of every axis; boundary cells of dst are NOT written
 * (bench/bench-stencil3.cc:125-126). lo0/hi0 restrict the sweep on axis 0 to planes [lo0, hi0)
 * (clipped to the interior) so a slab-sharded caller can do boundary planes first. */
typedef struct {
    int32_t kind;  /* racu_stencil_kind */
    int32_t dtype; /* RACU_F32 or RACU_F64 */
    int32_t rank;
    int32_t reserved;
    const void * src;
    void * dst;
    int64_t len[3];
    int64_t src_stride[3]; /* elements; innermost must be 1 */
    int64_t dst_stride[3];
    int64_t lo0, hi0;      /* 0,0 = whole interior */
} racu_stencil_desc;
int racu_stencil(const racu_stencil_desc * s, void * stream);

/* `steps` ping-pong sweeps, as the reference's benchmark loop writes them (bench/bench-stencil3.cc:55-64 + std::swap, :117-121):
 * src -> dst, dst -> src, ... On return state `steps` is in (steps odd ? dst : src) and state steps-1 in the other buffer, boundary cells
 * of both untouched: exactly what `steps` calls of racu_stencil with the roles swapped leave, bit for bit. For the 3-D 7-point sweep over
 * arrays a TMA tensor map can describe (innermost stride 1, 16-byte aligned rows) the steps are taken TWO PER PASS (the intermediate state
 * lives in shared memory only: half the HBM traffic per step) through a scratch buffer of the extent of the array; `scratch` = NULL lets the
 * library allocate it on the stream for the duration of the call. Anything else runs as single steps. lo0/hi0 must be 0. */
int racu_stencil_steps(const racu_stencil_desc * s, int32_t steps, void * scratch, void * stream);

/* ---- dense contraction ---- */
/* c(M,N) += a(M,K) * b(K,N), row-major with leading dimensions in elements. ra/ra.hh:403-412.
 * F64 uses FP64 tensor-core MMA; F32 uses the 3xTF32 split (error bound in DESIGN.md). */
typedef struct {
    int32_t dtype;
    int32_t reserved;
    int64_t M, N, K;
    const void * a; int64_t lda;
    const void * b; int64_t ldb;
    void * c; int64_t ldc;
} racu_gemm_desc;
int racu_gemm(const racu_gemm_desc * g, void * stream);
/* c(M) += a(M,N) * b(N)   (ra/ra.hh:414-425), trans=0;  c(N) += a(M) * b(M,N) (gevm ra/ra.hh:427-438), trans=1. */
int racu_gemv(int dtype, int trans, int64_t M, int64_t N, const void * a, int64_t lda,
              const void * x, int64_t incx, void * y, int64_t incy, void * stream);

/* ---- multi GPU (one process per GPU; NCCL over NVLink/NVSwitch) ---- */
#define RACU_UNIQUE_ID_BYTES 128
typedef struct racu_comm racu_comm;
int racu_comm_unique_id(void * id_out /* RACU_UNIQUE_ID_BYTES */);
int racu_comm_init_rank(racu_comm ** comm, const void * id, int nranks, int rank);
int racu_comm_destroy(racu_comm * comm);
/* In-place allreduce of count elements of dtype with redop SUM/PROD/AMIN/AMAX. */
int racu_allreduce(racu_comm * comm, void * buf_dev, size_t count, int dtype, int redop, void * stream);
/* Slab halo exchange along axis 0: send `count` elements at send_lo to rank-1 and send_hi to rank+1, receive
 * into recv_lo (from rank-1) and recv_hi (from rank+1). Ends are non periodic: missing neighbours are skipped. */
int racu_halo_exchange(racu_comm * comm, const void * send_lo, void * recv_lo, const void * send_hi, void * recv_hi,
                       size_t count, int dtype, void * stream);


/* ---- slab-sharded stencil step as ONE kernel: sweep + halo push over NVLink peer memory ----
 * One process per GPU. Each rank keeps its two slab buffers (A / Anext, ghost planes included) and one flag block in
 * memory from racu_alloc, exports them with racu_ipc_export, and maps its two neighbours' with racu_ipc_open (the
 * handles travel over any control plane: torch.distributed, MPI, a file). Replaces, for the slab owner, the sequence
 * "racu_halo_exchange; racu_stencil" and has no counterpart in the reference (single process, README.md:81).
 *
 * racu_stencil_push sweeps the whole local interior src -> dst like racu_stencil and, in the same kernel, stores the
 * first / last owned plane it produces into the neighbour's ghost plane of ITS dst buffer (push_lo / push_hi). Lock step:
 * flags[0] / flags[1] hold the number of steps the lower / higher neighbour has completed on the shared edge; the CTAs that
 * read a ghost plane (or overwrite the neighbour's) first wait for flag >= step, and the last of them to finish sets the
 * neighbour's flag to step+1 (system-scope release after the pushes). Only those edge CTAs wait, and they are scheduled in
 * the middle of the grid, so the exchange hides behind the interior. flags[4] != 0 after a sync means a neighbour did
 * not arrive within 20 s (the kernel gives up instead of hanging).
 * Call with step = 0, 1, 2, ... and swapped buffers; ghost planes of the first src must hold the neighbours' data. */
#define RACU_IPC_HANDLE_BYTES 64
#define RACU_HALO_FLAG_WORDS 8
int racu_ipc_export(void * dev_ptr, void * handle_out /* RACU_IPC_HANDLE_BYTES */);
int racu_ipc_open(const void * handle, void ** dev_ptr_out);
int racu_ipc_close(void * dev_ptr);
typedef struct {
    void * push_lo;        /* lower neighbour's HIGH ghost plane in its dst buffer (peer mapping), NULL at the global end */
    void * push_hi;        /* higher neighbour's LOW ghost plane in its dst buffer, NULL at the global end */
    uint32_t * flags;      /* local, RACU_HALO_FLAG_WORDS words, zeroed once before step 0 */
    uint32_t * signal_lo;  /* &flags[1] of the lower neighbour (peer mapping) */
    uint32_t * signal_hi;  /* &flags[0] of the higher neighbour (peer mapping) */
    uint32_t step;         /* steps already completed */
    uint32_t reserved;
} racu_halo_push;
int racu_stencil_push(const racu_stencil_desc * s, const racu_halo_push * h, void * stream);

#ifdef __cplusplus
}
#endif
#endif /* RACU_H */
