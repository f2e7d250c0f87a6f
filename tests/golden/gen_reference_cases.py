"""Transcribe the reference's own parametrised pytest cases into JSON fixtures.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/gen_reference_cases.py

Sources: /root/reference/MADN/test.py:7-945 (64 classic + 64 deterministic env_step cases) and
/root/reference/DOG/test.py:6-832 (52 normal, 17 neg, 14 swap, 29 hot-7 cases).  The parametrize
lists are literal Python; they are evaluated with a stand-in `jnp.array` so no JAX is needed.
Also records how each test function builds its env (the keyword defaults it passes to env_reset),
because those defaults differ from the env's own defaults.
"""
import ast
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


class _Jnp:
    int8 = np.int8
    int32 = np.int32

    @staticmethod
    def array(x, dtype=None):
        return np.array(x)


def _cases(path):
    """yield (test function name, argnames, [case tuples]) for each @pytest.mark.parametrize."""
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if not isinstance(node, ast.FunctionDef):
            continue
        for dec in node.decorator_list:
            if isinstance(dec, ast.Call) and ast.unparse(dec.func).endswith("parametrize"):
                names = [s.strip() for s in ast.literal_eval(dec.args[0]).split(",")]
                expr = ast.Expression(dec.args[1])
                ast.fix_missing_locations(expr)
                vals = eval(compile(expr, path, "eval"), {"jnp": _Jnp, "np": np})
                yield node.name, names, vals


def _plain(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def main():
    out = {}
    for rel in ("MADN/test.py", "DOG/test.py"):
        for fn, names, vals in _cases(os.path.join(REF, rel)):
            rows = []
            for case in vals:
                rows.append({k: _plain(v) for k, v in zip(names, case)})
            out[f"{rel}::{fn}"] = rows
    counts = {k: len(v) for k, v in out.items()}
    with open(os.path.join(OUT, "reference_cases.json"), "w") as f:
        json.dump({"source": "transcribed from /root/reference/{MADN,DOG}/test.py", "counts": counts, "cases": out}, f)
    print(counts)


if __name__ == "__main__":
    main()
