"""ctypes mirror of include/racu.h (the C ABI). Keep field order identical to the header."""
import ctypes as C

VERSION = 1
MAX_RANK = 8
MAX_OPND = 12
MAX_INS = 48

# ra/base.hh:232-234
ANY = -1944444444
UNB = -1988888888
MIS = -1922222222

OK, ERR_BAD_DESC, ERR_SHAPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_ALLOC, ERR_BOUNDS = range(8)
STATUS_NAMES = ["OK", "ERR_BAD_DESC", "ERR_SHAPE", "ERR_UNSUPPORTED", "ERR_CUDA", "ERR_NCCL", "ERR_ALLOC", "ERR_BOUNDS"]

BOOL, I32, I64, F32, F64 = range(5)
DTYPE_NAMES = ["bool", "i32", "i64", "f32", "f64"]
DTYPE_SIZE = [1, 4, 8, 4, 8]

MEM, SEQ, SCALAR, PTR = range(4)

OP_PUSH, OP_CAST = 0, 1
OP_NEG, OP_ABS, OP_SQRT, OP_SQR, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_NOT = range(2, 11)
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD, OP_MIN, OP_MAX, OP_POW, OP_ATAN2, OP_BAND, OP_BOR, OP_BXOR = range(16, 28)
OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE = range(32, 38)
OP_FMA, OP_PICK, OP_LOAD, OP_CLAMP, OP_LERP = 40, 41, 42, 43, 44
(OP_TAN, OP_SINH, OP_COSH, OP_TANH, OP_ASIN, OP_ACOS, OP_ATAN, OP_EXPM1, OP_LOG1P, OP_LOG10,
 OP_ISFINITE, OP_ISNAN, OP_ISINF) = range(48, 61)

ASSIGN_SET, ASSIGN_ADD, ASSIGN_SUB, ASSIGN_MUL, ASSIGN_DIV, ASSIGN_AMIN, ASSIGN_AMAX = range(7)
RED_SUM, RED_PROD, RED_AMIN, RED_AMAX = 1, 3, 5, 6

STENCIL_3D7, STENCIL_2D5, STENCIL_1D3, STENCIL_2D5_STIFF = 1, 2, 3, 4

UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
HALO_FLAG_WORDS = 8


class Base(C.Union):
    _fields_ = [("ptr", C.c_void_p), ("i", C.c_int64), ("f", C.c_double)]


class Operand(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dtype", C.c_int32),
        ("rank", C.c_int32),
        ("reserved", C.c_int32),
        ("base", Base),
        ("len", C.c_int64 * MAX_RANK),
        ("stride", C.c_int64 * MAX_RANK),
    ]


class Ins(C.Structure):
    _fields_ = [("op", C.c_uint8), ("type", C.c_uint8), ("arg", C.c_uint8), ("aux", C.c_uint8)]


class Desc(C.Structure):
    _fields_ = [
        ("nopnd", C.c_int32),
        ("nins", C.c_int32),
        ("dst", C.c_int32),
        ("assign", C.c_int32),
        ("opnd", Operand * MAX_OPND),
        ("ins", Ins * MAX_INS),
    ]


class StencilDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dtype", C.c_int32),
        ("rank", C.c_int32),
        ("reserved", C.c_int32),
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("len", C.c_int64 * 3),
        ("src_stride", C.c_int64 * 3),
        ("dst_stride", C.c_int64 * 3),
        ("lo0", C.c_int64),
        ("hi0", C.c_int64),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("reserved", C.c_int32),
        ("M", C.c_int64),
        ("N", C.c_int64),
        ("K", C.c_int64),
        ("a", C.c_void_p),
        ("lda", C.c_int64),
        ("b", C.c_void_p),
        ("ldb", C.c_int64),
        ("c", C.c_void_p),
        ("ldc", C.c_int64),
    ]


class HaloPush(C.Structure):
    _fields_ = [
        ("push_lo", C.c_void_p),
        ("push_hi", C.c_void_p),
        ("flags", C.c_void_p),
        ("signal_lo", C.c_void_p),
        ("signal_hi", C.c_void_p),
        ("step", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


# Every symbol include/racu.h declares: (name, restype, argtypes). tests/test_abi.py checks the built
# library exports each of them.
_P = C.c_void_p
SYMBOLS = [
    ("racu_version", C.c_int, []),
    ("racu_last_error_string", C.c_char_p, []),
    ("racu_device_count", C.c_int, []),
    ("racu_set_device", C.c_int, [C.c_int]),
    ("racu_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("racu_free", C.c_int, [_P]),
    ("racu_memcpy_h2d", C.c_int, [_P, _P, C.c_size_t, _P]),
    ("racu_memcpy_d2h", C.c_int, [_P, _P, C.c_size_t, _P]),
    ("racu_memcpy_d2d", C.c_int, [_P, _P, C.c_size_t, _P]),
    ("racu_fill", C.c_int, [_P, C.c_int, Base, C.c_size_t, _P]),
    ("racu_stream_sync", C.c_int, [_P]),
    ("racu_host_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("racu_host_free", C.c_int, [_P]),
    ("racu_agree", C.c_int, [C.POINTER(Desc), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    ("racu_map", C.c_int, [C.POINTER(Desc), _P]),
    ("racu_reduce", C.c_int, [C.POINTER(Desc), C.c_int, _P, _P]),
    ("racu_reduce_dev", C.c_int, [C.POINTER(Desc), C.c_int, _P, _P]),
    ("racu_reduce_async", C.c_int, [C.POINTER(Desc), C.c_int, _P, _P]),
    ("racu_plan_cache_stats", None, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("racu_prepare", C.c_int, [C.POINTER(Desc), C.c_int, C.POINTER(C.c_char_p)]),
    ("racu_launch_count", C.c_int64, []),
    ("racu_launch_count_reset", None, []),
    ("racu_last_kernel", C.c_char_p, []),
    ("racu_stencil", C.c_int, [C.POINTER(StencilDesc), _P]),
    ("racu_stencil_steps", C.c_int, [C.POINTER(StencilDesc), C.c_int32, _P, _P]),
    ("racu_gemm", C.c_int, [C.POINTER(GemmDesc), _P]),
    ("racu_gemv", C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P]),
    ("racu_comm_unique_id", C.c_int, [_P]),
    ("racu_comm_init_rank", C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int]),
    ("racu_comm_destroy", C.c_int, [_P]),
    ("racu_allreduce", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, _P]),
    ("racu_halo_exchange", C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, C.c_int, _P]),
    ("racu_ipc_export", C.c_int, [_P, _P]),
    ("racu_ipc_open", C.c_int, [_P, C.POINTER(_P)]),
    ("racu_ipc_close", C.c_int, [_P]),
    ("racu_stencil_push", C.c_int, [C.POINTER(StencilDesc), C.POINTER(HaloPush), _P]),
]


def bind(lib):
    """Set restype/argtypes for every ABI symbol on a loaded CDLL. Raises AttributeError if one is missing."""
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
