"""CPU-only: the host mirror's argument logic that decides WHICH C call is made -- descriptor names, mask flags, dtype
unification, operator strings -- pinned to the reference's rules (no GPU, no compute calls; the .so is only dlopen'ed for its
data symbols)."""
import itertools

import numpy as np
import pytest

import graphblas_b200 as gb
from graphblas_b200 import base, dtypes, operator


def test_descriptor_names_follow_reference_table():
    """reference core/descriptor.py:51-84: the 32 (replace, structure, complement, T0, T1) combinations map to GrB_DESC_<R><S><C><T0><T1>
    and the all-default one to NULL."""
    assert base.descriptor_lookup() is None
    seen = set()
    for r, s, c, t0, t1 in itertools.product([False, True], repeat=5):
        d = base.descriptor_lookup(output_replace=r, mask_structure=s, mask_complement=c, transpose_first=t0, transpose_second=t1)
        if not any((r, s, c, t0, t1)):
            assert d is None
            continue
        want = "GrB_DESC_" + ("R" if r else "") + ("S" if s else "") + ("C" if c else "") + ("T0" if t0 else "") + ("T1" if t1 else "")
        assert d.gb_name == want and d.gb_obj is not None
        seen.add(d.gb_name)
    assert len(seen) == 31 and {"GrB_DESC_RSC", "GrB_DESC_ST0", "GrB_DESC_T0T1", "GrB_DESC_RSCT0T1"} <= seen
    # reference core/descriptor.py:122-125 / tests/test_matrix.py:4329-4331: unknown options are an error on a non-suitesparse backend
    with pytest.raises(ValueError, match="Extra descriptor options not possible"):
        base.descriptor_lookup(nthreads=4)


def test_mask_flag_algebra():
    """reference core/mask.py:133-203: .S / .V and their complements carry (complement, structure, value) flags; ~~m is m."""
    class P:   # stands in for a Matrix / Vector: masks only hold a reference
        name, _carg = "p", None
    s, v = base.StructuralMask(P), base.ValueMask(P)
    assert (s.structure, s.value, s.complement) == (True, False, False)
    assert (v.structure, v.value, v.complement) == (False, True, False)
    assert ((~s).structure, (~s).complement) == (True, True) and ((~v).value, (~v).complement) == (True, True)
    assert type(~~s) is base.StructuralMask and type(~~v) is base.ValueMask and (~~s).parent is P


def test_dtype_unify_is_numpy_promotion():
    """reference core/dtypes.py:552-568: unify(a, b) = numpy promote_types on the 11 builtin types (BOOL absorbs into the other)."""
    all_types = [dtypes.BOOL, dtypes.INT8, dtypes.INT16, dtypes.INT32, dtypes.INT64, dtypes.UINT8, dtypes.UINT16, dtypes.UINT32,
                 dtypes.UINT64, dtypes.FP32, dtypes.FP64]
    for a, b in itertools.product(all_types, repeat=2):
        want = np.promote_types(a.np_type, b.np_type)
        assert dtypes.unify(a, b) == dtypes.lookup_dtype(want), (a, b)
    assert dtypes.unify(dtypes.INT64, dtypes.UINT64) == dtypes.FP64 and dtypes.unify(dtypes.BOOL, dtypes.INT8) == dtypes.INT8
    for key, want in ((int, dtypes.INT64), (float, dtypes.FP64), (bool, dtypes.BOOL), ("FP32", dtypes.FP32), (np.int16, dtypes.INT16),
                      ("GrB_UINT8", dtypes.UINT8), (np.dtype("float64"), dtypes.FP64)):
        assert dtypes.lookup_dtype(key) is want
    assert dtypes._INDEX is dtypes.UINT64   # GrB_Index = uint64_t (reference core/dtypes.py:389-396)


def test_operator_strings_and_typed_lookup():
    """reference core/operator/utils.py:60-157 + base.py string parsing: "+" / "plus" / "min.+" / "plus_times[FP32]"."""
    operator.initialize()
    assert operator.get_typed_op("+", dtypes.INT64, kind="binary").gb_name == "GrB_PLUS_INT64"
    assert operator.get_typed_op("plus", dtypes.FP32, kind="binary").gb_name == "GrB_PLUS_FP32"
    assert operator.get_typed_op("min.+", dtypes.INT32, dtypes.INT32, kind="semiring").gb_name == "GrB_MIN_PLUS_SEMIRING_INT32"
    assert operator.get_typed_op("plus_times[FP32]", dtypes.INT8, kind="semiring").gb_name == "GrB_PLUS_TIMES_SEMIRING_FP32"
    assert operator.get_typed_op(gb.monoid.plus, dtypes.INT16, kind="monoid").gb_name == "GrB_PLUS_MONOID_INT16"
    # two operand types unify before the lookup
    assert operator.get_typed_op(gb.semiring.plus_times, dtypes.INT32, dtypes.FP32, kind="semiring").type == dtypes.FP64
    assert operator.get_typed_op(gb.binary.minus, dtypes.UINT8, dtypes.INT8, kind="binary").type == dtypes.INT16
    with pytest.raises(ValueError, match="Unknown"):
        operator.get_typed_op("no_such_op", dtypes.INT64, kind="binary")
    with pytest.raises(TypeError):
        operator.get_typed_op(3.5, dtypes.INT64, kind="binary")
    # every hot semiring exists for every type it is defined on
    for name in ("plus_times", "min_plus", "plus_second", "any_pair", "max_plus", "plus_first", "min_first", "max_first"):
        sr = getattr(gb.semiring, name)
        for t in (dtypes.INT8, dtypes.INT64, dtypes.UINT32, dtypes.FP32, dtypes.FP64):
            assert t in sr, (name, t)
    assert dtypes.BOOL in gb.semiring.lor_land and gb.semiring.lor_land[dtypes.BOOL].return_type == dtypes.BOOL
    # the multiply's monoid identity table behind Matrix.power(0)
    from graphblas_b200.matrix import _MULT_IDENTITY

    assert _MULT_IDENTITY["times"](np.int64) == 1 and _MULT_IDENTITY["plus"](np.float32) == 0
    assert _MULT_IDENTITY["min"](np.int8) == 127 and _MULT_IDENTITY["max"](np.float64) == -np.inf and "pair" not in _MULT_IDENTITY


def test_mxm_row_costs_and_partition():
    """graphblas_b200/distributed.py: rows the split kernel hashes on several CTAs weigh more than their flops, rows that fit one table
    do not; the cost-balanced prefix split is monotone, covers every row once and balances the weights."""
    import numpy as np

    from graphblas_b200 import distributed as D

    f = np.array([0, 10, 10440, 10441, 40000, 100000], dtype=np.int64)
    c = D.mxm_row_costs(f)
    assert c[0] == 0 and c[1] == 10 and c[2] == 10440            # not split: cost = flops
    parts = np.ceil(f * 1.125 / 16320)
    assert np.allclose(c[3:], f[3:] * (1.0 + 1.0 * (parts[3:] - 1)))
    assert np.array_equal(D.mxm_row_costs(f, reread=0.0), f.astype(np.float64))
    import torch

    assert np.allclose(D.mxm_row_costs(torch.tensor(f)).numpy(), c)
    rng = np.random.default_rng(0)
    flops = rng.integers(0, 3000, 5000)
    flops[:20] = rng.integers(20000, 200000, 20)                 # heavy rows at the low indices, like an R-MAT matrix
    cost = D.mxm_row_costs(flops)
    prefix = np.concatenate([[0.0], np.cumsum(cost)])
    for world in (2, 4, 8):
        b = D.row_blocks_by_prefix(prefix, world)
        assert b[0] == 0 and b[-1] == flops.size and all(x <= y for x, y in zip(b[:-1], b[1:]))
        shares = np.array([cost[b[g]:b[g + 1]].sum() for g in range(world)])
        assert shares.max() <= cost.sum() / world + cost.max() + 1


def test_aggregator_recipes_name_ops_the_library_exports():
    """graphblas_b200/agg.py (reference graphblas/core/operator/agg.py:347-534): every aggregator the reference defines outside its
    `ss` namespace exists here, and every recipe step names a monoid / semiring / unary op that libgrb_cuda.so exports as a data
    symbol for the types it is used with (symbol discovery needs no GPU)."""
    from graphblas_b200 import agg
    from graphblas_b200 import ffi as F

    names = set(dir(F.__getattr__("lib")))
    reference_aggs = ["sum", "prod", "all", "any", "min", "max", "any_value", "count", "count_nonzero", "count_zero", "sum_of_squares",
                      "sum_of_inverses", "exists", "hypot", "logaddexp", "logaddexp2", "L0norm", "L1norm", "L2norm", "Linfnorm", "mean",
                      "peak_to_peak", "varp", "vars", "stdp", "stds", "geometric_mean", "harmonic_mean", "root_mean_square"]
    for n in reference_aggs:
        assert isinstance(getattr(agg, n), agg.Aggregator), n
    assert agg.L0norm is agg.count_nonzero and agg.L2norm is agg.hypot

    def semiring_symbol(name, t):
        add, mul = name.upper().split("_", 1)
        mul = {"PAIR": "PAIR"}.get(mul, mul)
        return [f"GrB_{add}_{mul}_SEMIRING_{t}", f"GxB_{add}_{mul}_{t}"]

    for a in agg._ALL:
        for sr in (a._semiring, a._semiring2):
            if sr is None:
                continue
            for t in ("INT64", "FP64", "FP32"):
                assert any(s in names for s in semiring_symbol(sr, t)), (a.name, sr, t)
        if a._monoid is not None:
            m = a._monoid.upper()
            cands = [f"GrB_{m}_MONOID_FP64", f"GxB_{m}_FP64_MONOID", f"GrB_{m}_MONOID_BOOL", f"GxB_{m}_BOOL_MONOID"]
            assert any(s in names for s in cands), (a.name, a._monoid)
        for u in (a._applybegin, a._finalize):
            if u is not None:
                assert f"GrB_{u.upper()}_FP64" in names or f"GxB_{u.upper()}_FP64" in names, (a.name, u)
        if a._composite is not None:
            assert all(isinstance(p, agg.Aggregator) for p in a._composite) and a._combine is not None
    # the ops the composite finalizers use
    for s in ("GrB_DIV_FP64", "GrB_MINUS_FP64", "GxB_POW_FP64", "GxB_SQRT_FP64", "GrB_MINV_FP64", "GrB_IDENTITY_FP64", "GxB_PAIR_INT64"):
        assert s in names, s
    # comparison-multiply semirings and GrB_Matrix_diag (round 2)
    for s in ("GxB_LOR_GT_INT32", "GxB_LAND_LE_FP64", "GxB_ANY_EQ_UINT8", "GxB_LOR_EQ_BOOL", "GxB_PLUS_ISGT_INT64", "GrB_Matrix_diag"):
        assert s in names, s
