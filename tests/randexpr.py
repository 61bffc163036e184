"""Random strided views and random expression trees, for oracle-vs-device parity tests."""
import numpy as np

from ra_ra_b200 import abi, ra

DT = [np.float32, np.float64, np.int32, np.int64]


def random_view(rng, ex, shape, dtype, intvals=False):
    """A view of `shape` over a larger random buffer, with random steps: sliced, reversed, transposed or broadcast axes."""
    rank = len(shape)
    perm = list(rng.permutation(rank))
    pad = [int(rng.integers(0, 3)) for _ in range(rank)]
    step = [int(rng.integers(1, 3)) for _ in range(rank)]
    big = [0] * rank
    for k in range(rank):
        big[perm[k]] = shape[k] * step[k] + pad[k]
    if dtype in (np.float32, np.float64):
        buf = rng.integers(-8, 9, big).astype(dtype) / dtype(1 if intvals else 4)
    else:
        buf = rng.integers(-8, 9, big).astype(dtype)
    v = ex.from_numpy(buf)
    nv = buf
    v = ra.transpose(v, [perm.index(k) for k in range(rank)]) if rank else v
    nv = np.transpose(nv, perm) if rank else nv
    subs, nsl = [], []
    for k in range(rank):
        o = int(rng.integers(0, pad[k] + 1))
        if rng.integers(0, 4) == 0 and shape[k] > 0:
            subs.append(ra.iota(shape[k], o + (shape[k] - 1) * step[k], -step[k]))
            nsl.append(slice(o + (shape[k] - 1) * step[k], None if o - step[k] < 0 else o - step[k], -step[k]))
        else:
            subs.append(ra.iota(shape[k], o, step[k]))
            nsl.append(slice(o, o + shape[k] * step[k], step[k]))
    v = v(*subs) if rank else v
    nv = nv[tuple(nsl)] if rank else nv
    assert nv.shape == tuple(shape)
    return v, nv


def random_expr(rng, ex, shape, depth):
    """Random expression whose operands have a random prefix of `shape` as their own shape."""
    if depth == 0 or rng.integers(0, 4) == 0:
        kind = rng.integers(0, 7)
        if kind == 6:  # gather: a(i) with an index EXPRESSION of the frame's prefix shape (unbeaten subscript, ra/ply.hh:461-499)
            r = int(rng.integers(0, len(shape) + 1))
            iv, _ = random_view(rng, ex, shape[:r], DT[rng.integers(2, 4)])
            src, _ = random_view(rng, ex, (7,), DT[rng.integers(0, 4)])
            return src(((ra.as_expr(iv) % 7) + 7) % 7)
        if kind == 0:
            return ra.as_expr(int(rng.integers(-3, 4)))
        if kind == 1:
            return ra.as_expr(float(rng.integers(-6, 7)) / 2)
        if kind == 2 and len(shape) > 0:
            return ra.tindex(int(rng.integers(0, len(shape))), abi.I32)
        r = int(rng.integers(0, len(shape) + 1))
        v, _ = random_view(rng, ex, shape[:r], DT[rng.integers(0, 4)])
        return ra.as_expr(v)
    op = rng.integers(0, 9)
    a = random_expr(rng, ex, shape, depth - 1)
    if op == 0:
        return -a
    if op == 1:
        return ra.abs(a)
    b = random_expr(rng, ex, shape, depth - 1)
    if op == 2:
        return a + b
    if op == 3:
        return a - b
    if op == 4:
        return a * b
    if op == 5:
        return ra.max(a, b)
    if op == 6:
        return ra.where(a < b, a, b)
    if op == 7:
        return ra.where(ra.logical_or(a.eq(b), a > 1), b, a + 1)
    return ra.min(a, b)


SHAPES = [(), (1,), (7,), (4, 5), (3, 1, 6), (2, 3, 4), (5, 2, 3, 2), (9, 8), (2, 2, 2, 2, 3), (33, 5), (6, 130)]


def random_scatter(rng, ex):
    """dst(idx) OP= expr through index arrays (ra/ply.hh:461-499 as an lvalue). Distinct indices for every op (any order gives the
    same result); repeated indices only for integer += / -= (exact in any order). Returns the destination buffer after the op."""
    n = int(rng.integers(1, 40))
    m = n + int(rng.integers(0, 9))
    dt = DT[rng.integers(0, 4)]
    dst, _ = random_view(rng, ex, (m,), dt, True)
    rep = bool(rng.integers(0, 2)) and dt in (np.int32, np.int64)
    idx = rng.integers(0, m, n) if rep else rng.permutation(m)[:n]
    iv = ex.from_numpy(idx.astype([np.int32, np.int64][int(rng.integers(0, 2))]))
    src, _ = random_view(rng, ex, (n,), np.int32 if rep else DT[rng.integers(0, 4)], True)
    e = ra.as_expr(src) * 2 + ra.tindex(0, abi.I32)
    op = ["__iadd__", "__isub__"][int(rng.integers(0, 2))] if rep else ["assign", "__iadd__", "__isub__", "__imul__"][int(rng.integers(0, 4))]
    g = dst(iv)
    getattr(g, op)(e)
    return dst
