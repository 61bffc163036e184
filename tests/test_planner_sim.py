"""The planner (ra_ra_b200/csrc/plan.cc) and the kernels' per-vector bodies, run on the CPU simulator and compared
with the oracle: known-answer tests first, then random expression trees over random strided views."""
import numpy as np
import pytest

import kat
from oracle_exec import OracleExecutor
from sim_exec import SimExecutor, plan_info
from ra_ra_b200 import abi, ra


@pytest.mark.parametrize("case", kat.CASES, ids=lambda f: f.__name__)
def test_kat_on_planner_sim(case):
    case(SimExecutor())


from randexpr import DT, SHAPES, random_expr, random_scatter, random_view


@pytest.mark.parametrize("seed", range(40))
def test_random_scatter_sim_vs_oracle(seed):
    outs = []
    for ex in (OracleExecutor(), SimExecutor()):
        rng = np.random.default_rng(9000 + seed)
        outs.append(np.array(random_scatter(rng, ex).owner, copy=True))
    assert outs[0].dtype == outs[1].dtype and np.array_equal(outs[0], outs[1]), seed


@pytest.mark.parametrize("seed", range(60))
def test_random_maps_sim_vs_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    shape = SHAPES[seed % len(SHAPES)]
    exo, exs = OracleExecutor(), SimExecutor()
    for assign in ("assign", "__iadd__", "__isub__", "__imul__"):
        state = rng.bit_generator.state
        outs = []
        for ex in (exo, exs):
            rng.bit_generator.state = state
            dst, _ = random_view(rng, ex, shape, DT[rng.integers(0, 4)])
            e = random_expr(rng, ex, shape, 2)
            try:
                getattr(dst, assign)(e)
            except ra.RaError as err:  # descriptor capacity (MAX_INS / MAX_OPND): same on both sides
                assert err.status == abi.ERR_UNSUPPORTED and ("too long" in str(err) or "too many" in str(err)), err
                outs = None
                break
            outs.append(np.array(dst.owner, copy=True))
        if outs is None:
            continue
        assert outs[0].dtype == outs[1].dtype
        assert np.array_equal(outs[0], outs[1], equal_nan=True), (seed, assign)


@pytest.mark.parametrize("seed", range(40))
def test_random_folds_sim_vs_oracle(seed):
    """dst with a lower rank (prefix) or inserted axes: += folds, integer-valued so any order is exact."""
    rng = np.random.default_rng(5000 + seed)
    shape = [s for s in SHAPES if len(s) >= 2][seed % 8]
    exo, exs = OracleExecutor(), SimExecutor()
    for variant in range(3):
        state = rng.bit_generator.state
        outs = []
        for ex in (exo, exs):
            rng.bit_generator.state = state
            rank = len(shape)
            if variant == 0:  # trailing axes folded: dst is a prefix
                r = int(rng.integers(0, rank))
                dst, _ = random_view(rng, ex, shape[:r], DT[rng.integers(0, 4)], True)
            elif variant == 1:  # leading axes folded: insert<n> in front
                r = int(rng.integers(1, rank))
                dst, _ = random_view(rng, ex, shape[r:], DT[rng.integers(0, 4)], True)
                dst = dst(ra.insert(r))
            else:  # middle axis folded (the gemm pattern)
                k = int(rng.integers(0, rank))
                dst, _ = random_view(rng, ex, shape[:k] + shape[k + 1:], DT[rng.integers(0, 4)], True)
                dst = dst(ra.dots(k), ra.insert(1))
            a, _ = random_view(rng, ex, shape, DT[rng.integers(0, 4)], True)  # integer valued: exact in any order, and under narrowing
            b, _ = random_view(rng, ex, shape[:int(rng.integers(0, rank + 1))], DT[rng.integers(0, 4)], True)
            op = ["__iadd__", "__isub__", "assign"][int(rng.integers(0, 3))]
            getattr(dst, op)(a * b + 1)
            outs.append(np.array(dst.owner, copy=True))
        assert np.array_equal(outs[0], outs[1]), (seed, variant)


def test_reductions_sim_vs_oracle():
    rng = np.random.default_rng(9)
    exo, exs = OracleExecutor(), SimExecutor()
    for shape in [(1,), (5,), (1000,), (40, 50), (3, 4, 300), (2049,), (70000,)]:
        for dt in DT:
            x = rng.integers(-4, 5, shape).astype(dt)
            y = rng.integers(-4, 5, shape).astype(dt)
            for f in (ra.sum, ra.amax, ra.amin, ra.reduce_sqrm):
                assert f(exo.wrap(x)) == f(exs.wrap(x)), (shape, dt, f.__name__)
            assert ra.dot(exo.wrap(x), exo.wrap(y)) == ra.dot(exs.wrap(x), exs.wrap(y))
            s = (x % 2 + 1).astype(dt)
            if np.prod(shape) <= 50:
                assert ra.prod(exo.wrap(s)) == ra.prod(exs.wrap(s))


def test_planner_decisions():
    """The contiguous fast path is found: B += A*C collapses to one axis with 16-byte staging (ra/ply.hh:42-46 analogue)."""
    ex = SimExecutor()
    A, B, Cc = (ex.from_numpy(np.ones((64, 48), dtype=np.float32)) for _ in range(3))
    B += A * Cc
    pi = plan_info()
    assert pi["mode"] == 0 and pi["rankA"] == 1 and pi["len0"] == 64 * 48 and pi["wide"] == 0 and pi["nvec"] == 3
    v = np.arange(64, dtype=np.float32)
    B += A * Cc + v  # prefix broadcast: v has step 0 on the inner axis, so rows do not collapse
    pi = plan_info()
    assert pi["rankA"] == 2 and pi["len0"] == 48 and pi["nvec"] == 3
    # transposed destination: axes are sorted by destination step, so dst stays the coalesced one
    T = ex.from_numpy(np.zeros((48, 64), dtype=np.float32))
    ra.transpose(T).assign(A)
    pi = plan_info()
    assert pi["rankA"] == 2 and pi["len0"] == 64
    # full reduction -> fold-inner over one collapsed axis; row sums -> fold-inner per kept row when rows are long
    ra.sum(A)
    pi = plan_info()
    assert pi["mode"] == 1 and pi["rankA"] == 1 and pi["rankB"] == 0 and pi["nk"] == 1
    big = ex.from_numpy(np.ones((300, 1024), dtype=np.float32))
    c = ex.from_numpy(np.zeros(300, dtype=np.float32))
    c += big
    pi = plan_info()
    assert pi["mode"] == 1 and pi["nk"] == 300 and pi["len0"] == 1024
    c2 = ex.from_numpy(np.zeros(1024, dtype=np.float32))
    cc = c2(ra.insert(1))
    cc += big
    pi = plan_info()
    assert pi["mode"] == 2 and pi["nk"] == 1024 and pi["nB"] == 300
    assert np.array_equal(c2.numpy(), np.full(1024, 300, dtype=np.float32))


@pytest.mark.parametrize("shape", [(2, 2, 2, 2, 3), (2, 3, 2, 2, 3, 2), (2,) * 8])
@pytest.mark.parametrize("red", ["sum", "amax", "amin"])
def test_whole_reduction_over_more_axes_than_one_launch(shape, red):
    """Five to eight uncollapsible axes (random strides): a whole-expression reduction is run as `out OP= expr` onto a rank-0
    destination, peeled along the outermost axis (api.cc run_plan / plan.cc peel_reduction); integer valued, so exact."""
    vals = []
    for ex in (OracleExecutor(), SimExecutor()):
        rng = np.random.default_rng(77)
        a, _ = random_view(rng, ex, shape, np.int64, True)
        b, _ = random_view(rng, ex, shape[:3], np.int32, True)
        vals.append(int(getattr(ra, red)(a * b - 3)))
    assert vals[0] == vals[1]


@pytest.mark.parametrize("seed", range(12))
def test_broadcast_destination_with_many_axes(seed):
    """Six to eight uncollapsible axes with a destination that is dead on some of them: `=` keeps the last slice of the peeled axis
    ("last one wins", docs/ra-ra.texi:1986-2007), `+=` and the running forms `dst = dst OP expr` accumulate over every slice."""
    shape = [(2, 3, 2, 2, 3, 2), (2,) * 8, (3, 2, 2, 2, 2, 2)][seed % 3]
    outs = []
    for ex in (OracleExecutor(), SimExecutor()):
        rng = np.random.default_rng(4400 + seed)
        k = int(rng.integers(0, len(shape)))
        dst, _ = random_view(rng, ex, shape[:k] + shape[k + 1:], DT[rng.integers(0, 4)], True)
        dst = dst(ra.dots(k), ra.insert(1))
        a, _ = random_view(rng, ex, shape, DT[rng.integers(0, 4)], True)
        form = seed % 4
        if form == 0:
            dst.assign(a * 2 + 1)
        elif form == 1:
            dst += a
        elif form == 2:
            dst.assign(dst + a)
        else:
            dst.assign(ra.max(dst, ra.cast(dst.dtype, a)))
        outs.append(np.array(dst.owner, copy=True))
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("seed", range(12))
def test_narrowing_fold_sim_vs_oracle(seed):
    """for_each(y OP= x) narrows to the destination at every step (ra/expr.hh:50-53): an integer destination folded over a
    floating point right hand side is NOT the tree sum (int c += 0.5, four times, is 0). Non-integer values on purpose."""
    outs = []
    for ex in (OracleExecutor(), SimExecutor()):
        rng = np.random.default_rng(31000 + seed)
        m, n, k = (int(x) for x in rng.integers(2, 6, 3))
        a = ex.from_numpy(rng.uniform(-3, 3, (m, n, k)).astype([np.float32, np.float64][seed % 2]))
        c = ex.from_numpy(rng.integers(-5, 5, (m,) if seed % 3 == 0 else (m, n)).astype([np.int32, np.int64][(seed // 2) % 2]))
        op = ["__iadd__", "__isub__", "__imul__"][seed % 3]
        if seed % 4 == 3:  # leading axis folded
            c = ex.from_numpy(rng.integers(-5, 5, (n, k)).astype(np.int32))
            getattr(c(ra.insert(1)), op)(a)
        else:
            getattr(c, op)(a)
        outs.append(np.array(c.owner, copy=True))
    assert outs[0].dtype == outs[1].dtype and np.array_equal(outs[0], outs[1]), seed


def test_narrowing_fold_known_answer():
    for ex in (OracleExecutor(), SimExecutor()):
        c = ex.from_numpy(np.zeros(3, np.int32))
        c += ex.from_numpy(np.full((3, 4), 0.5, np.float32))
        assert c.owner.tolist() == [0, 0, 0], ex.name


def test_interleaved_disjoint_views_are_not_aliases():
    """Even / odd and red / black updates: destination and source interleave in memory but never share an element. Legal and order
    independent in the reference; only true overlap is rejected."""
    for ex in (OracleExecutor(), SimExecutor()):
        a = ex.from_numpy(np.arange(20, dtype=np.float64))
        a(ra.iota(10, 0, 2)).assign(a(ra.iota(10, 1, 2)) * 2)
        assert a.owner.tolist() == [2 * (i + 1) if i % 2 == 0 else i for i in range(20)], ex.name
        b = ex.from_numpy(np.arange(48, dtype=np.int32).reshape(6, 8))
        b(ra.iota(3, 0, 2), ra.all).__iadd__(b(ra.iota(3, 1, 2), ra.all))          # even rows += odd rows
        ref = np.arange(48, dtype=np.int32).reshape(6, 8)
        ref[0::2] += ref[1::2]
        assert np.array_equal(b.owner, ref), ex.name
        c = ex.from_numpy(np.arange(64, dtype=np.float32).reshape(8, 8))
        c(ra.iota(4, 0, 2), ra.iota(4, 0, 2)).assign(c(ra.iota(4, 1, 2), ra.iota(4, 1, 2)))  # red <- black, both strided axes
        ref = np.arange(64, dtype=np.float32).reshape(8, 8)
        ref[0::2, 0::2] = ref[1::2, 1::2]
        assert np.array_equal(c.owner, ref), ex.name
    sim = SimExecutor()
    a = sim.from_numpy(np.arange(20, dtype=np.float64))
    with pytest.raises(ra.RaError) as e:  # shifted by a whole step: elements ARE shared
        a(ra.iota(8, 0, 2)).assign(a(ra.iota(8, 2, 2)))
    assert e.value.status == abi.ERR_UNSUPPORTED
    with pytest.raises(ra.RaError):
        a(ra.iota(19, 0)).assign(a(ra.iota(19, 1)))
    d = sim.from_numpy(np.arange(64, dtype=np.float32).reshape(8, 8))
    with pytest.raises(ra.RaError):  # rows 0,2,4 <- rows 2,4,6
        d(ra.iota(3, 0, 2), ra.all).assign(d(ra.iota(3, 2, 2), ra.all))
