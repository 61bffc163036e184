"""Expressions that are NOT in the pre-compiled program table (csrc/static_progs.h) and are laid out so that the run-time
specialiser (csrc/jit.cc) takes them: aligned unit-step rows, per-row values, reversed rows, scalars. Shared by the CPU test
(racu_prepare: plan + NVRTC compile, no GPU) and the GPU parity test (same cases against the oracle)."""
import numpy as np

from ra_ra_b200 import abi, ra

M, N = 48, 256  # N a multiple of 4: whole 16-byte vectors per row


def make_inputs(ex, seed=11):
    rng = np.random.default_rng(seed)
    f = lambda *s: (rng.integers(-40, 41, s) / 8).astype(np.float32)  # noqa: E731  (exactly representable eighths)
    return {
        "A": ex.from_numpy(f(M, N)), "B": ex.from_numpy(f(M, N)), "C": ex.from_numpy(f(M, N)),
        "D": ex.from_numpy((rng.integers(-40, 41, (M, N)) / 8).astype(np.float64)),
        "I": ex.from_numpy(rng.integers(-5, 6, (M, N)).astype(np.int32)),
        "L": ex.from_numpy(rng.integers(-5, 6, (M, N)).astype(np.int64)),
        "v": ex.from_numpy((rng.integers(-40, 41, M) / 8).astype(np.float64)),   # one value per row (prefix matching)
        "w": ex.from_numpy(f(M)),
        "x": ex.from_numpy(f(M * N)), "y": ex.from_numpy(f(M * N)),              # flat
        "rows": ex.from_numpy(np.zeros(M, np.float32)), "cols": ex.from_numpy(np.zeros(N, np.float32)),
        "mask": ex.from_numpy(np.zeros((M, N), bool)),
        "ridx": ex.from_numpy(rng.integers(0, M, M).astype(np.int32)),        # row indices (with repeats)
        "perm": ex.from_numpy(rng.permutation(M * N).astype(np.int64)),       # flat permutation
        "cidx": ex.from_numpy(rng.integers(0, N, N).astype(np.int32)),
    }


def _rev(v):
    return v(ra.all, ra.iota(N, N - 1, -1))


# name -> (kind, build(inputs) -> (expr, dst or None, assign, redop)); kind: map | fold
CASES = {
    "where_sqrt": ("map", lambda q: (ra.where(q["A"] < q["C"], q["A"] * q["C"], ra.sqrt(ra.abs(q["C"]))), q["B"], abi.ASSIGN_SET, 0)),
    "mixed_f32_f64_rowvec": ("map", lambda q: (q["A"] * q["D"] + q["v"] - 2.5, q["B"], abi.ASSIGN_ADD, 0)),
    "pick3": ("map", lambda q: (ra.pick(ra.cast(abi.I32, q["A"] > 0) + ra.cast(abi.I32, q["C"] > 0), q["A"], q["C"], q["A"] + q["C"]),
                                q["B"], abi.ASSIGN_SET, 0)),
    "int_ops": ("map", lambda q: ((q["I"] * 3 - q["I"] % 4) & 7, q["I"], abi.ASSIGN_SUB, 0)),
    "int64_from_int32": ("map", lambda q: (q["I"] * q["L"] + q["I"], q["L"], abi.ASSIGN_SET, 0)),
    "compare_to_bool": ("map", lambda q: (ra.logical_and(q["A"] > q["C"], q["D"] < 0.5), q["mask"], abi.ASSIGN_SET, 0)),
    "reversed_rows": ("map", lambda q: (ra.max(_rev(q["A"]), q["C"]) - q["w"], q["B"], abi.ASSIGN_MUL, 0)),
    "five_operands": ("map", lambda q: ((q["A"] + q["C"]) * (q["B"] - q["D"]) / (ra.abs(q["I"]) + 1), q["B"], abi.ASSIGN_SET, 0)),
    "flat_axpby": ("map", lambda q: (q["x"] * 0.5 - q["y"] * 0.25 + ra.min(q["x"], q["y"]), q["y"], abi.ASSIGN_SET, 0)),
    "index_init_i64": ("map", lambda q: (ra.tindex(0, abi.I64) * N + ra.tindex(1, abi.I64), q["L"], abi.ASSIGN_SET, 0)),       # A = _0*n + _1
    "index_mixed_f32": ("map", lambda q: (q["A"] * ra.tindex(1, abi.I32) - ra.tindex(0, abi.I32), q["B"], abi.ASSIGN_SET, 0)),
    "index_flat_iota": ("map", lambda q: (q["x"] + ra.iota(M * N, 3, 2), q["y"], abi.ASSIGN_ADD, 0)),
    "fold_index_weighted": ("fold", lambda q: (q["x"] * (ra.as_expr(ra.iota(M * N, 1, 1, abi.I32)) % 8), None, 0, abi.RED_SUM)),
    "fold_rows_index": ("foldto", lambda q: (q["A"] * ra.tindex(1, abi.I32) + ra.tindex(0, abi.I32), q["rows"], abi.ASSIGN_ADD, 0)),
    "fold_cols_index": ("foldto", lambda q: (q["C"] * ra.tindex(0, abi.I32) + ra.tindex(1, abi.I32), q["cols"](ra.insert(1)), abi.ASSIGN_ADD, 0)),
    "gather_rows": ("map", lambda q: (q["A"](q["ridx"]) + q["C"], q["B"], abi.ASSIGN_SET, 0)),                  # B = A(I) + C: whole rows by index (ra/ply.hh:461-499)
    "gather_flat_perm": ("map", lambda q: (q["x"](q["perm"]) * 2, q["y"], abi.ASSIGN_ADD, 0)),
    "gather_outer_product": ("map", lambda q: (q["D"](q["ridx"], q["cidx"]), q["B"], abi.ASSIGN_SET, 0)),         # B(i,j) = D(I(i), J(j))
    "gather_fold": ("fold", lambda q: (q["x"](q["perm"]) * q["y"], None, 0, abi.RED_SUM)),
    "fold_sum_abs_diff_mul": ("fold", lambda q: (ra.abs(q["A"] - q["C"]) * q["C"], None, 0, abi.RED_SUM)),
    "fold_amax_where": ("fold", lambda q: (ra.where(q["A"] > 0, q["A"], -q["C"]), None, 0, abi.RED_AMAX)),
    "fold_rows_inner": ("foldto", lambda q: (q["A"] * q["C"] - q["D"], q["rows"], abi.ASSIGN_ADD, 0)),          # c(i) += f(i,j): fold-inner per row
    "fold_cols_outer": ("foldto", lambda q: (q["A"] * q["w"] + q["C"], q["cols"](ra.insert(1)), abi.ASSIGN_ADD, 0)),  # c(j) += f(i,j): fold-outer
}
