"""The reference's known-answer tests for GrB_mxm / GrB_mxv / GrB_vxm as data.

Numbers come from tests/golden/reference_vectors.json (extracted from the reference's
graphblas/tests/test_matrix.py and test_vector.py by tests/golden/make_golden.py);
this module only says which operation each expected value pins, citing the reference line.

A case is a dict:
  kind        'mxm' | 'mxv' | 'vxm'
  semiring    e.g. 'plus_times'
  a, b        operand refs (file, function, index-in-function) ; for mxv: a=matrix b=vector; vxm: a=vector b=matrix
  ta, tb      operand transposed (.T)  -> GrB_DESC_T0 / T1
  out         ref of the initial output (dup of it) or None for a fresh empty output
  out_shape   shape of a fresh output
  mask, mask_kind ('S'|'V'), complement, replace, accum
  expect      ref of the expected result
"""
import json
import pathlib

import numpy as np

_J = json.loads((pathlib.Path(__file__).parent / "golden" / "reference_vectors.json").read_text())

M, V = "test_matrix.py", "test_vector.py"


def load(ref):
    """-> dict(kind='Matrix'|'Vector', idx arrays, vals, shape kwargs, dtype)"""
    f, fn, k = ref
    c = _J[f][fn]["from_coo"][k]
    args, kw = c["args"], c["kwargs"]
    vals = np.array(args[-1])
    if vals.dtype == np.dtype(int):
        vals = vals.astype(np.int64)  # python ints -> INT64 (reference core/utils.py:78-114)
    out = {"kind": c["kind"], "vals": vals, "line": c["line"]}
    if c["kind"] == "Matrix":
        out["rows"], out["cols"] = np.array(args[0], dtype=np.int64), np.array(args[1], dtype=np.int64)
        out["nrows"] = kw.get("nrows", int(out["rows"].max()) + 1)
        out["ncols"] = kw.get("ncols", int(out["cols"].max()) + 1)
    else:
        out["idx"] = np.array(args[0], dtype=np.int64)
        out["size"] = kw.get("size", int(out["idx"].max()) + 1)
    return out


def _c(**kw):
    base = dict(ta=False, tb=False, out=None, mask=None, mask_kind=None, complement=False, replace=False,
                accum=None, out_shape=None)
    base.update(kw)
    return base


A_M, v_M = (M, "A", 0), (M, "v", 0)
A_V, v_V = (V, "A", 0), (V, "v", 0)

CASES = [
    # ---- graphblas/tests/test_matrix.py
    _c(id="mxm:307", kind="mxm", semiring="plus_times", a=A_M, b=A_M, expect=(M, "test_mxm", 0), out_shape=(7, 7)),
    _c(id="mxm_T1:317", kind="mxm", semiring="plus_times", a=A_M, b=A_M, tb=True, out=A_M,
       expect=(M, "test_mxm_transpose", 0)),
    _c(id="mxm_T0:325", kind="mxm", semiring="plus_times", a=A_M, b=A_M, ta=True, out=A_M,
       expect=(M, "test_mxm_transpose", 1)),
    _c(id="mxm_mask_V:351", kind="mxm", semiring="plus_times", a=A_M, b=A_M, out=A_M,
       mask=(M, "test_mxm_mask", 0), mask_kind="V", expect=(M, "test_mxm_mask", 2)),
    _c(id="mxm_mask_CV:359", kind="mxm", semiring="plus_times", a=A_M, b=A_M, out=A_M,
       mask=(M, "test_mxm_mask", 0), mask_kind="V", complement=True, expect=(M, "test_mxm_mask", 3)),
    _c(id="mxm_mask_S_replace:367", kind="mxm", semiring="plus_times", a=A_M, b=A_M, out=A_M,
       mask=(M, "test_mxm_mask", 1), mask_kind="S", replace=True, expect=(M, "test_mxm_mask", 4)),
    _c(id="mxm_new_mask_S:371", kind="mxm", semiring="plus_times", a=A_M, b=A_M, out_shape=(7, 7),
       mask=(M, "test_mxm_mask", 1), mask_kind="S", expect=(M, "test_mxm_mask", 4)),
    _c(id="mxm_accum:377", kind="mxm", semiring="plus_times", a=A_M, b=A_M, out=A_M, accum="plus",
       expect=(M, "test_mxm_accum", 0)),
    _c(id="mxv:389", kind="mxv", semiring="plus_times", a=A_M, b=v_M, out_shape=(7,), expect=(M, "test_mxv", 0)),
    # ---- graphblas/tests/test_vector.py
    _c(id="vxm:299", kind="vxm", semiring="plus_times", a=v_V, b=A_V, out_shape=(7,), expect=(V, "test_vxm", 0)),
    _c(id="vxm_T1:305", kind="vxm", semiring="plus_times", a=v_V, b=A_V, tb=True, out_shape=(7,),
       expect=(V, "test_vxm_transpose", 0)),
    _c(id="vxm_nonsquare_min_plus:311", kind="vxm", semiring="min_plus", a=v_V, b=(V, "test_vxm_nonsquare", 0),
       out_shape=(2,), expect=(V, "test_vxm_nonsquare", 1)),
    _c(id="vxm_mask_S:329", kind="vxm", semiring="plus_times", a=v_V, b=A_V, out=v_V,
       mask=(V, "test_vxm_mask", 1), mask_kind="S", expect=(V, "test_vxm_mask", 2)),
    _c(id="vxm_mask_CS:335", kind="vxm", semiring="plus_times", a=v_V, b=A_V, out=v_V,
       mask=(V, "test_vxm_mask", 1), mask_kind="S", complement=True, expect=(V, "test_vxm_mask", 3)),
    _c(id="vxm_mask_V_replace:339", kind="vxm", semiring="plus_times", a=v_V, b=A_V, out=v_V,
       mask=(V, "test_vxm_mask", 0), mask_kind="V", replace=True, expect=(V, "test_vxm_mask", 4)),
    _c(id="vxm_new_mask_V:345", kind="vxm", semiring="plus_times", a=v_V, b=A_V, out_shape=(7,),
       mask=(V, "test_vxm_mask", 0), mask_kind="V", expect=(V, "test_vxm_mask", 4)),
    _c(id="vxm_accum:350", kind="vxm", semiring="plus_times", a=v_V, b=A_V, out=v_V, accum="plus",
       expect=(V, "test_vxm_accum", 0)),
]

# 1x5 . 5x1 max_plus = 33 (graphblas/tests/test_matrix.py:335-341): expected value is a scalar in the
# reference test (C[0, 0].new() == 33), so it is listed here rather than in the JSON.
NONSQUARE = dict(a=(M, "test_mxm_nonsquare", 0), b=(M, "test_mxm_nonsquare", 1), semiring="max_plus", expect=33)

# docs/user_guide/operations.rst:77-153 worked examples (values quoted in SURVEY.md section 8c).
