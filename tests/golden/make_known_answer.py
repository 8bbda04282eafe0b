"""Extract the hand-typed PyTorch known-answer case from the reference's own test.

Source: test/model/model.jl:80-283 ("Testing Against Pytorch"): 3 tables 5x4, batch 4,
bottom MLP 5->4, top MLP 10->5->1, values typed to 5 decimals.  The `py_*` literals are
parsed (not re-typed) into tests/golden/known_answer_small.json, PyTorch (row-major)
orientation, sparse indices kept 0-based as PyTorch has them.

    python tests/golden/make_known_answer.py [/root/reference]
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _matrix(text: str):
    rows = [r.strip() for r in text.strip().splitlines() if r.strip()]
    out = [[float(tok) for tok in r.replace(",", " ").split()] for r in rows]
    return out[0] if len(out) == 1 else out


def main(ref_root: str) -> None:
    src = open(os.path.join(ref_root, "test", "model", "model.jl")).read()
    src = src[src.index('@testset "Testing Against Pytorch"'):]
    out = {}
    # plain `py_x = Float32.([ ... ])` / `Float32.(\n [ ... ],\n )` literals
    for m in re.finditer(r"(py_\w+)\s*=\s*Float32\.\(\s*\[(.*?)\]('?)\s*,?\s*\)", src, re.S):
        name, body, _t = m.groups()
        if "Float32" in body:  # the vector-of-matrices literal, handled below
            continue
        out[name] = _matrix(body)
    # py_embedding_outputs = [ Float32.([..]), Float32.([..]), Float32.([..]) ]
    seg = src[src.index("py_embedding_outputs = ["):src.index("# note: [:, 1:4]")]
    out["py_embedding_outputs"] = [
        _matrix(b) for b in re.findall(r"Float32\.\(\s*\[(.*?)\]\s*,?\s*\)", seg, re.S)
    ]
    # py_sparse_input = [[3, 1, 4, 2] .+ 1, ...]  -> keep PyTorch's 0-based values
    seg = re.search(r"py_sparse_input\s*=\s*\[(.*)\]\n", src).group(1)
    out["py_sparse_input"] = [
        [int(t) for t in grp.split(",")] for grp in re.findall(r"\[([\d,\s]+)\]", seg)
    ]
    expected = {
        "py_dense_input": (4, 5), "py_dense1_weights": (4, 5), "py_dense1_bias": (4,),
        "py_embedding1_weights": (5, 4), "py_embedding2_weights": (5, 4),
        "py_embedding3_weights": (5, 4), "py_dense2_weights": (5, 10),
        "py_dense2_bias": (5,), "py_dense3_weights": (5,), "py_dense3_bias": (1,),
        "py_bottom_mlp_output": (4, 4), "py_interaction_output": (4, 10),
        "py_top_mlp_output": (4,), "py_expected_result": (4,),
    }
    for k, shape in expected.items():
        v = out[k]
        got = (len(v),) if not isinstance(v[0], list) else (len(v), len(v[0]))
        assert got == shape, (k, got, shape)
    assert len(out["py_embedding_outputs"]) == 3 and len(out["py_sparse_input"]) == 3
    dst = os.path.join(HERE, "known_answer_small.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print(dst, sorted(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
