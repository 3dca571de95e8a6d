"""
Extract the dealiasing / low-pass mask known-answer arrays that the reference
keeps in its own test-suite (tests/test_filter_masks.py) into a JSON fixture,
WITHOUT importing the reference (it needs jax/equinox, absent here): the test
file is parsed with `ast` and each `np.testing.assert_equal(
ex.spectral.low_pass_filter_mask(D, N, cutoff=..., axis_separate=...),
np.array([...]))` call is turned into {args, expected}.

Run (in the build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_filter_mask_golden.py
"""
import ast
import json
import os

SRC = "/root/reference/tests/test_filter_masks.py"
OUT = os.path.join(os.path.dirname(__file__), "filter_masks.json")


def main():
    tree = ast.parse(open(SRC).read())
    cases = []
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and ast.unparse(node.func) == "np.testing.assert_equal"):
            continue
        call, expected = node.args
        fn = ast.unparse(call.func)
        if not fn.endswith("low_pass_filter_mask"):
            continue  # oddball masks are not on the hot path
        D, N = (ast.literal_eval(a) for a in call.args)
        kw = {k.arg: ast.literal_eval(k.value) for k in call.keywords}
        arr = ast.literal_eval(expected.args[0])
        cases.append({"num_spatial_dims": D, "num_points": N, "cutoff": kw["cutoff"],
                      "axis_separate": kw.get("axis_separate", True), "expected": arr,
                      "source": f"tests/test_filter_masks.py:{node.lineno}"})
    json.dump(cases, open(OUT, "w"))
    print(f"wrote {len(cases)} cases to {OUT}")


if __name__ == "__main__":
    main()
