#!/usr/bin/env python
"""Extract the reference's known-answer vectors for GrB_mxm / GrB_mxv / GrB_vxm.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It parses -- with ``ast``, nothing is executed or imported -- the literal
``Matrix.from_coo(...)`` / ``Vector.from_coo(...)`` calls in the reference's own
tests and writes them, in source order, per test function, to
``tests/golden/reference_vectors.json``.  The operation each expected value pins
(which semiring, mask flavour, accum ...) is described in ``CASES`` below with
the reference file:line it was read from; the numbers themselves come from the
reference files.
"""
import ast
import json
import pathlib
import sys

REF = pathlib.Path("/root/reference/graphblas/tests")
OUT = pathlib.Path(__file__).with_name("reference_vectors.json")

WANTED = {
    "test_matrix.py": [
        "A", "v", "test_mxm", "test_mxm_transpose", "test_mxm_nonsquare", "test_mxm_mask",
        "test_mxm_accum", "test_mxv",
    ],
    "test_vector.py": [
        "A", "v", "test_vxm", "test_vxm_transpose", "test_vxm_nonsquare", "test_vxm_mask",
        "test_vxm_accum",
    ],
}


def _literal(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        return None


def extract(path, names):
    tree = ast.parse(path.read_text())
    out = {}
    for fn in tree.body:
        if not isinstance(fn, ast.FunctionDef) or fn.name not in names:
            continue
        calls = []
        local = {}
        for node in ast.walk(fn):
            # `data = [[...], [...], [...]]` used by the fixtures
            if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
                val = _literal(node.value)
                if val is not None:
                    local[node.targets[0].id] = val
        for node in ast.walk(fn):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)):
                continue
            if node.func.attr != "from_coo" or not isinstance(node.func.value, ast.Name):
                continue
            kind = node.func.value.id  # Matrix | Vector
            args = []
            for a in node.args:
                if isinstance(a, ast.Starred) and isinstance(a.value, ast.Name):
                    args.extend(local[a.value.id])
                else:
                    args.append(_literal(a))
            kwargs = {k.arg: _literal(k.value) for k in node.keywords}
            calls.append({"kind": kind, "args": args, "kwargs": kwargs, "line": node.lineno})
        calls.sort(key=lambda c: c["line"])
        out[fn.name] = {"line": fn.lineno, "from_coo": calls}
    return out


def main():
    if not REF.exists():
        sys.exit("reference tree not present; the committed JSON is the fixture")
    result = {}
    for fname, names in WANTED.items():
        result[fname] = extract(REF / fname, set(names))
    OUT.write_text(json.dumps(result, indent=1, sort_keys=True) + "\n")
    n = sum(len(v["from_coo"]) for f in result.values() for v in f.values())
    print(f"wrote {OUT} ({n} literal from_coo calls)")


if __name__ == "__main__":
    main()
