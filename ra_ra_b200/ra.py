"""Host-side mirror of ra-ra's array surface for the traversal hot path, in Python.

The reference is a C++ header library; the product host side is the C++ header include/ra/cuda.hh. This module
is the same surface for Python callers (tests, bench.py): views with prefix-matching rank extension, beaten
slicing (iota / all / dots / insert / len), transpose / reverse, operator expression trees, reductions and
gemm/gemv. It does no arithmetic on array contents: every operation is flattened into a racu_desc
(include/racu.h) and handed to an executor, which for the product is the CUDA library (ra_ra_b200.cuda).

Reference lines mirrored are cited per function (relative to the reference root).
"""
import builtins as _builtins

import numpy as np

from . import abi
from .abi import ANY, BOOL, F32, F64, I32, I64, MEM, SCALAR, SEQ, UNB  # noqa: F401

builtins_len = _builtins.len

NP2RACU = {np.dtype(np.bool_): BOOL, np.dtype(np.int32): I32, np.dtype(np.int64): I64,
           np.dtype(np.float32): F32, np.dtype(np.float64): F64}
RACU2NP = {v: k for k, v in NP2RACU.items()}


class RaError(RuntimeError):
    """RA_CK failure (ra/expr.hh:17-43): raised before any kernel is launched."""

    def __init__(self, status, msg):
        super().__init__(f"{abi.STATUS_NAMES[status] if 0 <= status < builtins_len(abi.STATUS_NAMES) else status}: {msg}")
        self.status = status


def _ck(cond, msg, status=abi.ERR_BOUNDS):
    if not cond:
        raise RaError(status, msg)


# ---------------- C++ usual arithmetic conversions over the five element types ----------------

def promote(t):
    """Integer promotion: bool -> int."""
    return I32 if t == BOOL else t


def common_type(a, b):
    a, b = promote(a), promote(b)
    if F64 in (a, b):
        return F64
    if F32 in (a, b):
        return F32
    if I64 in (a, b):
        return I64
    return I32


def is_float(t):
    return t in (F32, F64)


# ---------------- subscripts: ra/ply.hh:314-330 ----------------

class LenExpr:
    """ra::len inside subscripts (ra/base.hh:278-291), resolved by wlen (ra/ply.hh:243-295) against the axis length."""

    def __init__(self, f=lambda n: n):
        self.f = f

    def resolve(self, n):
        v = self.f(n)
        return v if isinstance(v, np.ndarray) else int(v)  # (len - std::array {...}: an unbeatable subscript, test/fromu.cc:26)

    def _bin(self, other, op):
        g = other.f if isinstance(other, LenExpr) else (lambda n, o=other: o)
        return LenExpr(lambda n, f=self.f, g=g: op(f(n), g(n)))

    def _rbin(self, other, op):
        return LenExpr(lambda n, f=self.f, o=other: op(o, f(n)))

    @staticmethod
    def _cdiv(a, b):  # C++ integer division truncates toward zero
        q = _builtins.abs(a) // _builtins.abs(b)
        return q if (a >= 0) == (b >= 0) else -q

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._rbin(o, lambda a, b: a + b)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._rbin(o, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._rbin(o, lambda a, b: a * b)
    def __floordiv__(self, o): return self._bin(o, LenExpr._cdiv)
    def __truediv__(self, o): return self._bin(o, LenExpr._cdiv)
    def __neg__(self): return LenExpr(lambda n, f=self.f: -f(n))


len = LenExpr()  # noqa: A001  (ra::len)


def _wlen(x, n):
    return x.resolve(n) if isinstance(x, LenExpr) else x


class Dots:
    def __init__(self, n):
        self.n = n


class Insert:
    def __init__(self, n):
        self.n = n


def dots(n=UNB):
    return Dots(n)


def insert(n=1):
    return Insert(n)


all = Dots(1)  # noqa: A001  (ra::all = dots<1>, ra/ply.hh:318)


class Iota:
    """ra::iota(n, o, s): rank-1 view over Seq{o} with {len n, step s} (ra/ply.hh:320-321). Usable as a subscript
    (beaten, ra/ply.hh:399-415) and as an expression operand (kind SEQ)."""

    __array_ufunc__ = None

    def __init__(self, n=UNB, o=None, s=1, dtype=None):
        self.n, self.s = n, s
        if dtype is None:
            dtype = I64 if o is None else (F64 if isinstance(o, float) else I32)
        self.o = 0 if o is None else o
        self.dtype = dtype

    # iota arithmetic stays iota: ra/ra.hh:129-143 (opt). Needed so A(I+1, J, K) is a view (bench/bench-stencil3.cc:59-61).
    def _isint(self, k):
        return isinstance(k, (int, np.integer)) and not isinstance(k, bool)

    def __add__(self, k):
        if isinstance(k, Iota):
            return Iota(_colen_py(self.n, k.n), self.o + k.o, self.s + k.s, common_type(self.dtype, k.dtype))
        if self._isint(k) or isinstance(k, LenExpr):
            return Iota(self.n, self.o + k, self.s, self.dtype)
        return as_expr(self) + k

    __radd__ = __add__

    def __sub__(self, k):
        if isinstance(k, Iota):
            return Iota(_colen_py(self.n, k.n), self.o - k.o, self.s - k.s, common_type(self.dtype, k.dtype))
        if self._isint(k) or isinstance(k, LenExpr):
            return Iota(self.n, self.o - k, self.s, self.dtype)
        return as_expr(self) - k

    def __rsub__(self, k):
        if self._isint(k) or isinstance(k, LenExpr):
            return Iota(self.n, k - self.o, -self.s, self.dtype)
        return k - as_expr(self)

    def __mul__(self, k):
        if self._isint(k):
            return Iota(self.n, self.o * k, self.s * k, self.dtype)
        return as_expr(self) * k

    __rmul__ = __mul__

    def __neg__(self):
        return Iota(self.n, -self.o, -self.s, self.dtype)


def _colen_py(a, b):
    if a == UNB:
        return b
    if b == UNB:
        return a
    _ck(a == b, f"Bad shapes {a} {b}.", abi.ERR_SHAPE)
    return a


def iota(n=UNB, o=None, s=1, dtype=None):
    return Iota(n, o, s, dtype)


# ---------------- expression tree ----------------

class Expr:
    """A node of the expression template tree (Map / Pick / Cell / Scalar, ra/expr.hh:104-120,266-312,652-722)."""
    dtype = None
    __array_ufunc__ = None  # numpy operands defer to our reflected operators

    def _b(self, other, op, rev=False):
        o = as_expr(other)
        a, b = (o, self) if rev else (self, o)
        return binop(op, a, b)

    def __add__(self, o): return self._b(o, abi.OP_ADD)
    def __radd__(self, o): return self._b(o, abi.OP_ADD, True)
    def __sub__(self, o): return self._b(o, abi.OP_SUB)
    def __rsub__(self, o): return self._b(o, abi.OP_SUB, True)
    def __mul__(self, o): return self._b(o, abi.OP_MUL)
    def __rmul__(self, o): return self._b(o, abi.OP_MUL, True)
    def __truediv__(self, o): return self._b(o, abi.OP_DIV)
    def __rtruediv__(self, o): return self._b(o, abi.OP_DIV, True)
    def __mod__(self, o): return self._b(o, abi.OP_MOD)
    def __rmod__(self, o): return self._b(o, abi.OP_MOD, True)
    def __and__(self, o): return self._b(o, abi.OP_BAND)
    def __or__(self, o): return self._b(o, abi.OP_BOR)
    def __xor__(self, o): return self._b(o, abi.OP_BXOR)
    def __lt__(self, o): return self._b(o, abi.OP_LT)
    def __le__(self, o): return self._b(o, abi.OP_LE)
    def __gt__(self, o): return self._b(o, abi.OP_GT)
    def __ge__(self, o): return self._b(o, abi.OP_GE)
    def eq(self, o): return self._b(o, abi.OP_EQ)
    def ne(self, o): return self._b(o, abi.OP_NE)

    def __neg__(self):
        t = promote(self.dtype)
        return Op(abi.OP_NEG, [cast(t, self)], t, t)

    def __pos__(self):
        return cast(promote(self.dtype), self)

    def __invert__(self):
        """operator! of the reference (Python has no overloadable `not`)."""
        return Op(abi.OP_NOT, [self], self.dtype, BOOL)


class Leaf(Expr):
    def __init__(self, view):
        self.view = view
        self.dtype = view.dtype


class SeqLeaf(Expr):
    """Cell over Seq: iota operands and tindex (ra/arrays.hh:361-364)."""

    def __init__(self, dtype, org, lens, steps):
        self.dtype, self.org, self.lens, self.steps = dtype, org, list(lens), list(steps)


class ScalarLeaf(Expr):
    def __init__(self, value, dtype):
        self.value, self.dtype = value, dtype


class HostVec(Expr):
    """A foreign rank-1 range (std::vector, ra/expr.hh:385): uploaded when the expression executes."""

    def __init__(self, arr):
        self.arr = np.ascontiguousarray(arr)
        _ck(self.arr.ndim == 1, "foreign vectors are rank 1")
        _ck(self.arr.dtype in NP2RACU, f"unsupported dtype {self.arr.dtype}", abi.ERR_UNSUPPORTED)
        self.dtype = NP2RACU[self.arr.dtype]


class Op(Expr):
    def __init__(self, op, args, optype, rtype, arg=0, aux=0):
        self.op, self.args, self.optype, self.dtype, self.arg, self.aux = op, args, optype, rtype, arg, aux

    # pick / where as an lvalue (test/where.cc:148-157, ra/expr.hh:686-728): `pick(sel, a0, a1, ...) OP= x` updates, per element, the
    # branch the selector names. On the device: one masked assignment per branch, a_k = pick(sel, ..., a_k OP x at k, ...).
    def _assign(self, assign_op, rhs):
        _ck(self.op == abi.OP_PICK and _builtins.all(isinstance(b, Leaf) for b in self.args[1:]),
            "only pick / where over arrays can be assigned to", abi.ERR_UNSUPPORTED)
        sel, branches = self.args[0], self.args[1:]
        binop_of = {abi.ASSIGN_ADD: abi.OP_ADD, abi.ASSIGN_SUB: abi.OP_SUB, abi.ASSIGN_MUL: abi.OP_MUL, abi.ASSIGN_DIV: abi.OP_DIV}
        for k, b in enumerate(branches):
            new = as_expr(rhs) if assign_op == abi.ASSIGN_SET else binop(binop_of[assign_op], b, rhs)
            new = cast(b.dtype, new)
            b.view.assign(Op(abi.OP_PICK, [sel] + [new if j == k else b for j in range(builtins_len(branches))], b.dtype, b.dtype,
                             arg=builtins_len(branches), aux=sel.dtype))
        return self

    def assign(self, rhs): return self._assign(abi.ASSIGN_SET, rhs)
    def __iadd__(self, rhs): return self._assign(abi.ASSIGN_ADD, rhs)
    def __isub__(self, rhs): return self._assign(abi.ASSIGN_SUB, rhs)
    def __imul__(self, rhs): return self._assign(abi.ASSIGN_MUL, rhs)
    def __itruediv__(self, rhs): return self._assign(abi.ASSIGN_DIV, rhs)


class Gather(Expr):
    """a(i ...) with unbeaten subscripts (index arrays / index expressions): ra/ply.hh:461-499,531-549. `view` is the source after
    its beaten subscripts, `off` the I64 element offset expression sum(index_k * step_k) with every index placed on its own
    expression axes, [lo, hi] the valid offsets of `view` (the device clamps: the reference asserts, ra/ply.hh:395)."""

    def __init__(self, view, off, lo, hi):
        self.view, self.off, self.lo, self.hi, self.dtype = view, off, lo, hi, view.dtype

    # as an lvalue (test/fromu.cc:53-61,84-92): a[off] OP= x over the frame of the subscripts and x. Repeated indices combine in an
    # unspecified order (the reference: sequentially), `=` then leaves any one of the values.
    def _assign(self, op, rhs):
        d, keep = flatten(as_expr(rhs), dst=self, assign=op, ex=self.view.ex)
        self.view.ex.map(d)
        del keep
        return self

    def assign(self, rhs): return self._assign(abi.ASSIGN_SET, rhs)
    def __iadd__(self, rhs): return self._assign(abi.ASSIGN_ADD, rhs)
    def __isub__(self, rhs): return self._assign(abi.ASSIGN_SUB, rhs)
    def __imul__(self, rhs): return self._assign(abi.ASSIGN_MUL, rhs)
    def __itruediv__(self, rhs): return self._assign(abi.ASSIGN_DIV, rhs)


def _expr_rank(e):
    if isinstance(e, Leaf):
        return e.view.rank
    if isinstance(e, SeqLeaf):
        return builtins_len(e.lens)
    if isinstance(e, HostVec):
        return 1
    if isinstance(e, Gather):
        return _expr_rank(e.off)
    if isinstance(e, Op):
        return _builtins.max([_expr_rank(a) for a in e.args] + [0])
    return 0


def _shift_axes(e, n, ex):
    """The same expression placed n axes further in: its first n axes are dead (len UNB, step 0), like insert<n>."""
    if n == 0 or isinstance(e, ScalarLeaf):
        return e
    if isinstance(e, Leaf):
        return Leaf(e.view(insert(n)))
    if isinstance(e, HostVec):
        return Leaf(ex.from_numpy(e.arr)(insert(n)))
    if isinstance(e, SeqLeaf):
        return SeqLeaf(e.dtype, e.org, [UNB] * n + e.lens, [0] * n + e.steps)
    if isinstance(e, Gather):
        return Gather(e.view, _shift_axes(e.off, n, ex), e.lo, e.hi)
    return Op(e.op, [_shift_axes(a, n, ex) for a in e.args], e.optype, e.dtype, e.arg, e.aux)


def tindex(w, dtype=I64):
    """ra::_0 .. : index along axis w, dead on the axes before it (Reframe, ra/expr.hh:402-455)."""
    return SeqLeaf(dtype, 0, [UNB] * (w + 1), [0] * w + [1])


_0, _1, _2, _3, _4 = (tindex(w) for w in range(5))


def as_expr(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, View):
        return Leaf(x)
    if isinstance(x, Iota):
        n = x.n
        _ck(not isinstance(n, LenExpr) and not isinstance(x.o, LenExpr), "Stray ra::len.")  # ra/expr.hh:513
        return SeqLeaf(x.dtype, x.o, [n], [x.s])
    if isinstance(x, (bool, np.bool_)):
        return ScalarLeaf(bool(x), BOOL)
    if isinstance(x, (np.float32,)):
        return ScalarLeaf(float(x), F32)
    if isinstance(x, (np.int64,)):
        return ScalarLeaf(int(x), I64)
    if isinstance(x, (int, np.integer)):
        return ScalarLeaf(int(x), I32)  # C++ int literal
    if isinstance(x, (float, np.floating)):
        return ScalarLeaf(float(x), F64)  # C++ double literal
    if isinstance(x, np.ndarray):
        return HostVec(x)
    if isinstance(x, (list, tuple)):
        return HostVec(np.asarray(x))
    raise TypeError(f"cannot use {type(x)} in an array expression")


def cast(t, e):
    """ra::cast<T> (ra/expr.hh:680-681)."""
    e = as_expr(e)
    if e.dtype == t:
        return e
    return Op(abi.OP_CAST, [e], t, t, aux=e.dtype)


_CMP = (abi.OP_EQ, abi.OP_NE, abi.OP_LT, abi.OP_LE, abi.OP_GT, abi.OP_GE)
_BIT = (abi.OP_BAND, abi.OP_BOR, abi.OP_BXOR)


def binop(op, a, b):
    """RA_BINOP (ra/ra.hh:181-190): the node's type is the C++ type of `a OP b`."""
    a, b = as_expr(a), as_expr(b)
    if op in _BIT and a.dtype == BOOL and b.dtype == BOOL:
        t = BOOL
    else:
        t = common_type(a.dtype, b.dtype)
        _ck(not (op in _BIT and is_float(t)), "bit op on floating point", abi.ERR_BAD_DESC)
    return Op(op, [cast(t, a), cast(t, b)], t, BOOL if op in _CMP else t)


def _unary_float(op, a):
    a = as_expr(a)
    t = a.dtype if is_float(a.dtype) else F64  # std::sqrt(int) -> double
    return Op(op, [cast(t, a)], t, t)


def sqrt(a): return _unary_float(abi.OP_SQRT, a)
def exp(a): return _unary_float(abi.OP_EXP, a)
def log(a): return _unary_float(abi.OP_LOG, a)
def sin(a): return _unary_float(abi.OP_SIN, a)
def cos(a): return _unary_float(abi.OP_COS, a)
# the rest of the using-declared <cmath> functions of ra/ra.hh:221-222
def tan(a): return _unary_float(abi.OP_TAN, a)
def sinh(a): return _unary_float(abi.OP_SINH, a)
def cosh(a): return _unary_float(abi.OP_COSH, a)
def tanh(a): return _unary_float(abi.OP_TANH, a)
def asin(a): return _unary_float(abi.OP_ASIN, a)
def acos(a): return _unary_float(abi.OP_ACOS, a)
def atan(a): return _unary_float(abi.OP_ATAN, a)
def expm1(a): return _unary_float(abi.OP_EXPM1, a)
def log1p(a): return _unary_float(abi.OP_LOG1P, a)
def log10(a): return _unary_float(abi.OP_LOG10, a)


def _classify(op, a):
    a = as_expr(a)
    t = a.dtype if is_float(a.dtype) else F64
    return Op(op, [cast(t, a)], t, BOOL)


def isfinite(a): return _classify(abi.OP_ISFINITE, a)
def isnan(a): return _classify(abi.OP_ISNAN, a)
def isinf(a): return _classify(abi.OP_ISINF, a)


def clamp(v, lo, hi):
    """std::clamp (ra/ra.hh:222): v < lo ? lo : hi < v ? hi : v, in the common type of the three."""
    v, lo, hi = as_expr(v), as_expr(lo), as_expr(hi)
    t = common_type(common_type(v.dtype, lo.dtype), hi.dtype)
    return Op(abi.OP_CLAMP, [cast(t, v), cast(t, lo), cast(t, hi)], t, t)


def lerp(a, b, t_):
    """std::lerp (ra/ra.hh:222)."""
    a, b, t_ = as_expr(a), as_expr(b), as_expr(t_)
    t = common_type(common_type(a.dtype, b.dtype), t_.dtype)
    t = t if is_float(t) else F64
    return Op(abi.OP_LERP, [cast(t, a), cast(t, b), cast(t, t_)], t, t)


def odd(a):
    """ra/ra.hh:61: N & 1 on the value converted to unsigned int."""
    a = as_expr(a)
    _ck(not is_float(a.dtype), "odd() of a floating point value", abi.ERR_BAD_DESC)
    return (cast(promote(a.dtype), a) & 1).ne(0)


def rel_error(a, b):
    """ra/ra.hh:94: D = |a| + |b|; D == 0 ? 0. : 2.*|a - b| / D, computed with the double literals and returned in the real type."""
    a, b = as_expr(a), as_expr(b)
    R = common_type(a.dtype, b.dtype)
    R = R if is_float(R) else F64
    a, b = cast(R, a), cast(R, b)
    D = abs(a) + abs(b)
    return cast(R, where(D.eq(0), 0., 2. * abs(a - b) / D))


def cmp3(a, b):
    """operator<=> (ra/ra.hh:190) as an int: -1, 0, +1, and 2 where the operands are unordered (a NaN). The reference yields
    std::partial_ordering objects, which are not array elements on the device; comparing the result with 0 reads the same."""
    a, b = as_expr(a), as_expr(b)
    return where(a < b, -1, where(a > b, 1, where(a.eq(b), 0, 2)))


def abs(a):  # noqa: A001
    a = as_expr(a)
    t = promote(a.dtype)
    return Op(abi.OP_ABS, [cast(t, a)], t, t)


def sqr(a):
    """ra/ra.hh:65."""
    a = as_expr(a)
    t = promote(a.dtype)
    return Op(abi.OP_SQR, [cast(t, a)], t, t)


def sqrm(a, b=None):
    """ra/ra.hh:71-72: sqrm(x) = x*x, sqrm(x, y) = sqr(x-y)."""
    return sqr(a) if b is None else sqr(as_expr(a) - b)


def max(a, b): return binop(abi.OP_MAX, a, b)  # noqa: A001
def min(a, b): return binop(abi.OP_MIN, a, b)  # noqa: A001


def pow(a, b):  # noqa: A001
    a, b = as_expr(a), as_expr(b)
    t = F32 if (a.dtype == F32 and b.dtype == F32) else F64
    return Op(abi.OP_POW, [cast(t, a), cast(t, b)], t, t)


def atan2(a, b):
    a, b = as_expr(a), as_expr(b)
    t = F32 if (a.dtype == F32 and b.dtype == F32) else F64
    return Op(abi.OP_ATAN2, [cast(t, a), cast(t, b)], t, t)


def fma(a, b, c):
    a, b, c = as_expr(a), as_expr(b), as_expr(c)
    t = common_type(common_type(a.dtype, b.dtype), c.dtype)
    t = t if is_float(t) else F64
    return Op(abi.OP_FMA, [cast(t, a), cast(t, b), cast(t, c)], t, t)


def pick(sel, *branches):
    """ra::pick(sel, a0, a1, ...) (ra/expr.hh:686-728)."""
    sel = as_expr(sel)
    _ck(not is_float(sel.dtype) and builtins_len(branches) >= 1, "bad pick", abi.ERR_BAD_DESC)
    bs = [as_expr(b) for b in branches]
    t = bs[0].dtype
    for b in bs[1:]:
        t = b.dtype if b.dtype == t else common_type(t, b.dtype)
    return Op(abi.OP_PICK, [sel] + [cast(t, b) for b in bs], t, t, arg=builtins_len(bs), aux=sel.dtype)


def where(w, t, f):
    """ra/ra.hh:255-256: pick(cast<bool>(w), f, t)."""
    return pick(cast(BOOL, w), f, t)


def logical_and(a, b):
    """operator&& ra/ra.hh:262-263: where(a, cast<bool>(b), false)."""
    return where(a, cast(BOOL, b), False)


def logical_or(a, b):
    """operator|| ra/ra.hh:265-266: where(a, true, cast<bool>(b))."""
    return where(a, True, cast(BOOL, b))




# ---------------- views: ra/arrays.hh:84-149 ----------------

class View:
    """Non-owning pointer + dims (ViewBig, ra/expr.hh:233). `owner` keeps the memory alive; `ptr` is the byte address
    of element [0,...,0]; dims is a list of [len, step] in elements. `ex` is the executor that runs traversals."""

    __array_ufunc__ = None

    def __init__(self, ex, owner, ptr, dtype, dims):
        self.ex, self.owner, self.ptr, self.dtype = ex, owner, int(ptr), dtype
        self.dims = [[int(l), int(s)] for l, s in dims]

    # -- queries
    @property
    def rank(self):
        return builtins_len(self.dims)

    def len(self, k):
        return self.dims[k][0]

    def step(self, k):
        return self.dims[k][1] if k < self.rank else 0  # ra/expr.hh:294-295

    @property
    def shape(self):
        return tuple(d[0] for d in self.dims)

    @property
    def size(self):
        s = 1
        for l, _ in self.dims:
            if l < 0:
                return l
            s *= l
        return s

    def _at(self, off, dims):
        return View(self.ex, self.owner, self.ptr + off * abi.DTYPE_SIZE[self.dtype], self.dtype, dims)

    # -- beaten slicing: from() / fromb(), ra/ply.hh:367-435,503-551
    def __call__(self, *subs):
        a, ak, pl, bv = self.dims, 0, 0, []
        subs = list(subs)
        gathers = {}  # position in bv -> integer index expression (unbeaten subscript)

        def fsrc(i):  # ra/ply.hh:344-348
            return i.n if isinstance(i, Dots) else 0 if isinstance(i, Insert) else 1

        for q, i0 in enumerate(subs):
            if isinstance(i0, Dots):
                n = i0.n
                if n == UNB:  # :417
                    n = self.rank - ak - _builtins.sum(fsrc(i) for i in subs[q + 1:] if not (isinstance(i, Dots) and i.n == UNB))
                _ck(n >= 0 and ak + n <= self.rank, "Too many subscripts.")
                bv += [list(d) for d in a[ak:ak + n]]
                ak += n
            elif isinstance(i0, Insert):  # :423-427
                bv += [[UNB, 0] for _ in range(i0.n)]
            elif isinstance(i0, Iota):  # :399-415
                _ck(ak < self.rank, "Too many subscripts.")
                la, st = a[ak]
                n, o, s = _wlen(i0.n, la), _wlen(i0.o, la), _wlen(i0.s, la)
                _ck(n != UNB, "unbounded iota subscripts are not beatable (ra/ply.hh:341)", abi.ERR_UNSUPPORTED)
                _ck(n == 0 or (0 <= o < la and 0 <= o + (n - 1) * s < la),
                    f"Bad iota len {n} step {s} in len[{ak}]={la}.")  # :407
                bv.append([n, s * st])
                pl += o * st
                ak += 1
            elif isinstance(i0, (View, Expr, list, tuple, np.ndarray)) or (isinstance(i0, LenExpr) and ak < self.rank
                                                                          and isinstance(_wlen(i0, a[ak][0]), np.ndarray)):
                # unbeaten: the axis is kept whole in the beaten view and addressed through the index expression (:461-499)
                _ck(ak < self.rank, "Too many subscripts.")
                ie = as_expr(_wlen(i0, a[ak][0]) if isinstance(i0, LenExpr) else i0)
                _ck(ie.dtype in (I32, I64), "subscripts must be integers", abi.ERR_BAD_DESC)
                gathers[builtins_len(bv)] = ie
                bv.append(list(a[ak]))
                ak += 1
            elif isinstance(i0, (int, np.integer, LenExpr)):  # :390-398
                _ck(ak < self.rank, "Too many subscripts.")
                la, st = a[ak]
                i = _wlen(i0, la)
                _ck(0 <= i < la or (la == UNB and st == 0), f"Bad index {i} in len[{ak}]={la}.")  # :395
                pl += i * st
                ak += 1
            else:
                raise RaError(abi.ERR_UNSUPPORTED, "unbeaten (gather) subscripts are not on the device path yet")
        bv += [list(d) for d in a[ak:]]  # trailing axes implied, :374-381
        _ck(builtins_len(bv) <= abi.MAX_RANK, "Rank too large.")
        if not gathers:
            return self._at(pl, bv)
        # fromu (:480-499): map(a-as-function, i ..., iota(len) for the other axes): every index gets its own expression axes
        pos, off, lo, hi = 0, None, 0, 0
        for q, (ln, st) in enumerate(bv):
            if q in gathers:
                ie = gathers[q]
                term = cast(I64, _shift_axes(ie, pos, self.ex)) * np.int64(st)
                pos += _expr_rank(ie)
            else:
                _ck(ln != UNB, "insert<> next to an index array subscript", abi.ERR_UNSUPPORTED)
                term = SeqLeaf(I64, 0, [UNB] * pos + [ln], [0] * pos + [1]) * np.int64(st)
                pos += 1
            if ln > 0:
                lo, hi = lo + _builtins.min(0, (ln - 1) * st), hi + _builtins.max(0, (ln - 1) * st)
            off = term if off is None else off + term
        _ck(pos <= abi.MAX_RANK, "Rank too large.")
        return Gather(self._at(pl, bv), off if off is not None else ScalarLeaf(0, I64), lo, hi)

    def __getitem__(self, subs):
        if subs is Ellipsis:
            return self
        return self(*subs) if isinstance(subs, tuple) else self(subs)

    def __setitem__(self, subs, rhs):
        dst = self[subs]
        if isinstance(rhs, View) and rhs.ptr == dst.ptr and rhs.dims == dst.dims and rhs.owner is dst.owner:
            return  # result of `A[I] += x`
        if isinstance(rhs, Gather) and isinstance(dst, Gather) and rhs.view.ptr == dst.view.ptr and rhs.view.dims == dst.view.dims:
            return  # result of `A[I] += x` with index arrays: the scatter already happened
        dst.assign(rhs)

    # -- assignment ops: ra/expr.hh:50-60, ra/arrays.hh:131-135
    def assign(self, rhs): return self._assign(abi.ASSIGN_SET, rhs)
    def __iadd__(self, rhs): return self._assign(abi.ASSIGN_ADD, rhs)
    def __isub__(self, rhs): return self._assign(abi.ASSIGN_SUB, rhs)
    def __imul__(self, rhs): return self._assign(abi.ASSIGN_MUL, rhs)
    def __itruediv__(self, rhs): return self._assign(abi.ASSIGN_DIV, rhs)

    def _assign(self, op, rhs):
        d, keep = flatten(as_expr(rhs), dst=self, assign=op, ex=self.ex)
        self.ex.map(d)
        del keep
        return self

    # -- operators build expressions
    def _e(self):
        return Leaf(self)

    def __add__(self, o): return self._e() + o
    def __radd__(self, o): return o + self._e()
    def __sub__(self, o): return self._e() - o
    def __rsub__(self, o): return o - self._e()
    def __mul__(self, o): return self._e() * o
    def __rmul__(self, o): return o * self._e()
    def __truediv__(self, o): return self._e() / o
    def __rtruediv__(self, o): return o / self._e()
    def __mod__(self, o): return self._e() % o
    def __and__(self, o): return self._e() & o
    def __or__(self, o): return self._e() | o
    def __xor__(self, o): return self._e() ^ o
    def __lt__(self, o): return self._e() < o
    def __le__(self, o): return self._e() <= o
    def __gt__(self, o): return self._e() > o
    def __ge__(self, o): return self._e() >= o
    def eq(self, o): return self._e().eq(o)
    def ne(self, o): return self._e().ne(o)
    def __neg__(self): return -self._e()
    def __invert__(self): return ~self._e()

    def numpy(self):
        """Copy the viewed elements to a host numpy array of the view's shape (dead axes not allowed)."""
        return self.ex.to_numpy(self)


def reverse(a, k=0):
    """ra/arrays.hh:408-428."""
    _ck(0 <= k < a.rank, f"Bad axis {k} for rank {a.rank}.")
    _ck(a.len(k) != UNB, f"Cannot reverse unbounded axis {k}.")
    dv = [list(d) for d in a.dims]
    off = 0 if dv[k][0] == 0 else dv[k][1] * (dv[k][0] - 1)
    dv[k][1] *= -1
    return a._at(off, dv)


def transpose(a, s=(1, 0)):
    """ra/arrays.hh:431-461: axis k of a goes to axis s[k] of the result; repeated targets give diagonals."""
    s = list(s)
    _ck(builtins_len(s) == a.rank, f"Bad rank {a.rank} for {builtins_len(s)} transposed axes.")
    r = 0 if not s else 1 + _builtins.max(s)
    dst = [[UNB, 0] for _ in range(r)]
    for k, sk in enumerate(s):
        dst[sk][1] += a.dims[k][1]
        dst[sk][0] = _builtins.min(dst[sk][0], a.dims[k][0]) if dst[sk][0] >= 0 else a.dims[k][0]
    return a._at(0, dst)


def diag(a):
    return transpose(a, (0, 0))


# ---------------- flattening: expression tree -> racu_desc ----------------

def flatten(e, dst=None, assign=abi.ASSIGN_SET, ex=None):
    """What ra::cuda::ply does in the C++ header: leaves -> operands, nodes -> postfix program.
    Returns (desc, keepalive)."""
    d = abi.Desc()
    keep = []
    nopnd = [0]
    nins = [0]

    def add_operand(kind, dtype, rank, base, lens, steps):
        _ck(nopnd[0] < abi.MAX_OPND, "too many operands", abi.ERR_UNSUPPORTED)
        _ck(rank <= abi.MAX_RANK, "Rank too large.")
        p = d.opnd[nopnd[0]]
        p.kind, p.dtype, p.rank = kind, dtype, rank
        if kind == MEM:
            p.base.ptr = base
        elif kind == abi.PTR:
            p.base.ptr = base
            p.len[0], p.len[1] = lens  # smallest / largest valid element offset
            lens = steps = []
        elif is_float(dtype):
            p.base.f = float(base)
        else:
            p.base.i = int(base)
        for k in range(rank):
            p.len[k], p.stride[k] = lens[k], steps[k]
        nopnd[0] += 1
        return nopnd[0] - 1

    def emit(op, t, arg=0, aux=0):
        _ck(nins[0] < abi.MAX_INS, "program too long", abi.ERR_UNSUPPORTED)
        i = d.ins[nins[0]]
        i.op, i.type, i.arg, i.aux = op, t, arg, aux
        nins[0] += 1

    seen = {}

    def once(key, make):
        """The same leaf used twice in the tree is one operand (one load)."""
        if key not in seen:
            seen[key] = make()
        return seen[key]

    def view_operand(v):
        keep.append(v.owner)
        return add_operand(MEM, v.dtype, v.rank, v.ptr, [x[0] for x in v.dims], [x[1] for x in v.dims])

    def view_key(v):
        return ("mem", v.ptr, v.dtype, tuple(map(tuple, v.dims)))

    def walk(n):
        if isinstance(n, Leaf):
            emit(abi.OP_PUSH, n.dtype, once(view_key(n.view), lambda: view_operand(n.view)))
        elif isinstance(n, HostVec):
            emit(abi.OP_PUSH, n.dtype, once(("vec", id(n)), lambda: view_operand(ex.from_numpy(n.arr))))
        elif isinstance(n, SeqLeaf):
            emit(abi.OP_PUSH, n.dtype, once(("seq", n.dtype, n.org, tuple(n.lens), tuple(n.steps)),
                                            lambda: add_operand(SEQ, n.dtype, builtins_len(n.lens), n.org, n.lens, n.steps)))
        elif isinstance(n, ScalarLeaf):
            emit(abi.OP_PUSH, n.dtype, once(("scalar", n.dtype, repr(n.value)), lambda: add_operand(SCALAR, n.dtype, 0, n.value, [], [])))
        elif isinstance(n, Gather):
            walk(n.off)
            keep.append(n.view.owner)
            emit(abi.OP_LOAD, n.dtype, once(("ptr", n.view.ptr, n.dtype, n.lo, n.hi),
                                            lambda: add_operand(abi.PTR, n.dtype, 0, n.view.ptr, [n.lo, n.hi], [])), I64)
        else:
            for a in n.args:
                walk(a)
            emit(n.op, n.optype, n.arg, n.aux)

    if isinstance(dst, Gather):  # scatter: the program leaves [element offset, value]; the destination is the RACU_PTR operand
        walk(dst.off)
        keep.append(dst.view.owner)
        d.dst = once(("ptr", dst.view.ptr, dst.dtype, dst.lo, dst.hi),
                     lambda: add_operand(abi.PTR, dst.dtype, 0, dst.view.ptr, [dst.lo, dst.hi], []))
    elif dst is not None:
        d.dst = once(view_key(dst), lambda: view_operand(dst))
    else:
        d.dst = -1
    d.assign = assign
    walk(e)
    d.nopnd, d.nins = nopnd[0], nins[0]
    return d, keep


# ---------------- whole-array reductions: ra/ra.hh:306-322,348-401 ----------------

def _reduce(e, redop):
    e = as_expr(e)
    ex = _find_ex(e)
    _ck(ex is not None, "Bad shapes: the expression has no array operand, so no shape (examples/agreement.cc:56-61).", abi.ERR_SHAPE)
    d, keep = flatten(e, ex=ex)
    out = ex.reduce(d, redop, e.dtype)
    del keep
    return out


def reduce_into(dst, e, redop=abi.RED_SUM):
    """Whole-expression reduction whose result STAYS where the arrays live: dst is a rank-0 view of the expression's type.
    The reference returns the scalar by value (ra/ra.hh:348-401); chains of reductions and maps (examples/cghs.hh) then need no
    host round trip per scalar."""
    e = as_expr(e)
    _ck(isinstance(dst, View) and dst.rank == 0 and dst.dtype == e.dtype, "reduce_into needs a rank-0 destination of the expression's type", abi.ERR_BAD_DESC)
    d, keep = flatten(e, ex=dst.ex)
    dst.ex.reduce_into(d, redop, dst.ptr)
    del keep
    return dst


def _find_ex(e):
    if isinstance(e, Leaf):
        return e.view.ex
    if isinstance(e, Gather):
        return e.view.ex
    if isinstance(e, Op):
        for a in e.args:
            x = _find_ex(a)
            if x is not None:
                return x
    return None


def sum(a): return _reduce(a, abi.RED_SUM)  # noqa: A001
def prod(a): return _reduce(a, abi.RED_PROD)
def amin(a): return _reduce(a, abi.RED_AMIN)
def amax(a): return _reduce(a, abi.RED_AMAX)


def stencil(a, lo, hi):
    """ra::stencil(a, lo, hi) (ra/arrays.hh:615-630): a view of rank 2*rank(a) over the same memory. Axes [0, r) run over the
    positions whose whole neighbourhood [-lo, +hi] is inside a, axes [r, 2r) over the neighbourhood: s[i..., d...] = a[i+d ...].
    `Anext(I,J,K) += stencil(A,1,1) * mask(insert<3>)` is the f_sumprod form of bench/bench-stencil3.cc:94-105."""
    r = a.rank
    lo = [lo] * r if isinstance(lo, (int, np.integer)) else list(lo)
    hi = [hi] * r if isinstance(hi, (int, np.integer)) else list(hi)
    _ck(_builtins.all(x >= 0 for x in lo + hi), f"Bad stencil bounds lo {lo} hi {hi}.")
    _ck(_builtins.all(a.len(k) >= lo[k] + hi[k] for k in range(r)), "Stencil is too large for array.")
    dims = [[a.len(k) - lo[k] - hi[k], a.step(k)] for k in range(r)] + [[lo[k] + hi[k] + 1, a.step(k)] for k in range(r)]
    _ck(builtins_len(dims) <= abi.MAX_RANK, "Rank too large.")
    return a._at(0, dims)


def at(a, i):
    """ra::at(a, i) (ra/ra.hh:228-239): i holds index TUPLES along its last axis (rank(a) coordinates each); the result has the
    shape of i without that axis: at(a, i)[...] = a[i[..., 0], i[..., 1], ...]."""
    _ck(isinstance(a, View) and isinstance(i, View), "at(a, i) takes views", abi.ERR_BAD_DESC)
    _ck(i.rank >= 1 and i.len(i.rank - 1) == a.rank, "at(): the last axis of i must hold rank(a) coordinates")
    off, lo, hi = None, 0, 0
    for k, (ln, st) in enumerate(a.dims):
        term = cast(I64, Leaf(i(dots(i.rank - 1), k))) * np.int64(st)
        off = term if off is None else off + term
        if ln > 0:
            lo, hi = lo + _builtins.min(0, (ln - 1) * st), hi + _builtins.max(0, (ln - 1) * st)
    return Gather(a, off if off is not None else ScalarLeaf(0, I64), lo, hi)


# any / every / index / lexical_compare (ra/ra.hh:279-304). The reference short-circuits a sequential traversal (early(),
# ra/ply.hh:169); the order is unspecified (docs/ra-ra.texi:3143-3148), so here they are max / min folds over the truth values
# and over the position of the first hit. Empty: False / True / -1 / False, like the reference's defaults.
_NONE = (1 << 63) - 1


def any(a):  # noqa: A001
    return bool(_reduce(cast(I32, cast(BOOL, a)), abi.RED_AMAX) > 0)


def every(a):
    return bool(_reduce(cast(I32, cast(BOOL, a)), abi.RED_AMIN) > 0)


def index(a):
    """First true element along axis 0 (a rank-1 iota() matched against axis 0: for rank > 1 the reference "checks elements
    but only returns the first index", test/early.cc:77-82): the smallest axis-0 index holding a true element, or -1."""
    i = int(_reduce(where(a, tindex(0, I64), np.int64(_NONE)), abi.RED_AMIN))
    return -1 if i == _NONE else i


def lexical_compare(a, b):
    """a < b at the first position, in row-major order, where they differ; False if they do not."""
    a, b = as_expr(a), as_expr(b)
    ne = a.ne(b)
    ex = _find_ex(ne)
    d, keep = flatten(ne, ex=ex)
    shape = ex.agree(d)
    del keep
    lt01 = where(a < b, np.int64(0), np.int64(1))
    if _builtins.len(shape) == 0:
        return int(_reduce(where(ne, lt01, np.int64(_NONE)), abi.RED_AMIN)) == 0
    pos, acc = None, 1
    for k in range(_builtins.len(shape) - 1, -1, -1):
        term = tindex(k, I64) * np.int64(acc)
        pos = term if pos is None else pos + term
        acc *= shape[k]
    key = int(_reduce(where(ne, pos * np.int64(2) + lt01, np.int64(_NONE)), abi.RED_AMIN))
    return key != _NONE and key % 2 == 0


def _ref_extreme(a, redop):
    """refmin / refmax (ra/ra.hh:326-346): a REFERENCE to the first smallest / largest element in row-major order, here a rank-0 view
    of it (assignable like the reference's double &). Two folds: the extreme value, then the smallest row-major position holding it."""
    _ck(isinstance(a, View), "refmin / refmax need an array (an lvalue), not an expression", abi.ERR_BAD_DESC)
    _ck(a.size > 0, "refmin requires nonempty argument.")
    m = _reduce(a, redop)
    shape = a.shape
    if builtins_len(shape) == 0:
        return a
    pos, acc = None, 1
    for k in range(builtins_len(shape) - 1, -1, -1):
        term = tindex(k, I64) * np.int64(acc)
        pos = term if pos is None else pos + term
        acc *= shape[k]
    first = int(_reduce(where(as_expr(a).eq(RACU2NP[a.dtype].type(m)), pos, np.int64(_NONE)), abi.RED_AMIN))
    idx = []
    for k in range(builtins_len(shape) - 1, -1, -1):
        idx.append(first % shape[k])
        first //= shape[k]
    return a(*reversed(idx))


def refmin(a): return _ref_extreme(a, abi.RED_AMIN)
def refmax(a): return _ref_extreme(a, abi.RED_AMAX)


def dot(a, b):
    """ra/ra.hh:379-385 with RA_FMA=0: c += a*b."""
    return _reduce(as_expr(a) * b, abi.RED_SUM)


def reduce_sqrm(a):
    """ra/ra.hh:395-401."""
    return _reduce(sqrm(a), abi.RED_SUM)


def norm2(a):
    """ra/ra.hh:376."""
    return float(np.sqrt(reduce_sqrm(a)))


# ---------------- BLAS-like: ra/ra.hh:403-456 ----------------

def gemm(a, b, c):
    """c(all, insert<1>) += a * b(insert<1>)   (ra/ra.hh:407)."""
    cc = c(all, insert(1))
    cc += a * b(insert(1))
    return c


def gemv(a, b, c):
    """c += a * b(insert<1>)   (ra/ra.hh:418)."""
    c += a * b(insert(1))
    return c


def gevm(a, b, c):
    """c(insert<1>) += a * b   (ra/ra.hh:431)."""
    cc = c(insert(1))
    cc += a * b
    return c
