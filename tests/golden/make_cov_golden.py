#!/usr/bin/env python3
"""Golden vectors for row f3 (covariance blocks of the VAMP-2 scan) produced by THE REFERENCE'S OWN CODE.

    python tests/golden/make_cov_golden.py            -> tests/golden/cov_blocks.npz

`_cov_blocks` (CAVIAR-1.0.py:396-402) and `_stack_even_odd` (:409-421) are nested functions of `run_pooled`, so they
cannot be imported; their source is cut out of the reference file with `ast` (no edits), compiled as it stands and run
on fixed inputs: two centred "ensembles" of distance-like features (float32, as the reference builds XA / XB at
:379-382), one lag, both parities.  Stored: the inputs and the reference's C00, Ctt, C0t for (train parity 0, validation
parity 1) exactly as `vamp2_cv_score` requests them (:426-429).  Runs only where /root/reference exists.
"""
import ast
import os
import sys

import numpy as np

REF = os.environ.get("CAVIAR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def nested_functions(path, outer, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == outer:
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef) and sub.name in names and sub.name not in out:
                    out[sub.name] = ast.get_source_segment(src, sub)
            break
    ns = {"np": np}
    for name in names:
        body = out[name]
        indent = len(body.splitlines()[1]) - len(body.splitlines()[1].lstrip()) - 4      # nested one level deep
        lines = body.splitlines()
        dedented = [lines[0].lstrip()] + [ln[indent:] if ln.strip() else ln for ln in lines[1:]]
        exec(compile("\n".join(dedented), f"{path}:{name}", "exec"), ns)
    return [ns[n] for n in names], out


def features(rng, n, K):
    t = np.linspace(0, 1, n)[:, None]
    base = rng.uniform(0.5, 4.0, size=(1, K))                               # nm, like mdtraj distances
    drift = np.sin(2 * np.pi * (t * rng.uniform(0.5, 3.0, size=(1, K)) + rng.random((1, K)))) * rng.uniform(0.02, 0.2, size=(1, K))
    return (base + drift + rng.normal(0, 0.03, size=(n, K))).astype(np.float32)


def main():
    (cov_blocks, stack_even_odd), src = nested_functions(os.path.join(REF, "CAVIAR-1.0.py"), "run_pooled",
                                                          ["_cov_blocks", "_stack_even_odd"])
    rng = np.random.default_rng(20261020)
    K, tau = 48, 4
    dist_a, dist_b = features(rng, 333, K), features(rng, 210, K)
    mu = (dist_a.mean(axis=0) + dist_b.mean(axis=0)) / 2.0                      # :379 mu_pool_sel
    XA, XB = dist_a - mu, dist_b - mu                                          # :381-382
    out = {"dist_a": dist_a, "dist_b": dist_b, "mu": mu.astype(np.float32), "tau": np.int64(tau)}
    for parity in (0, 1):
        X0, Xt = stack_even_odd([XA, XB], tau, parity)
        C00, Ctt, C0t = cov_blocks(X0, Xt)
        out[f"p{parity}_C00"], out[f"p{parity}_Ctt"], out[f"p{parity}_C0t"] = C00, Ctt, C0t
        out[f"p{parity}_n"] = np.int64(X0.shape[0])
        assert C00.dtype == np.float32
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "cov_blocks.npz"), **out)
    print("wrote cov_blocks.npz;", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})
    print("--- reference source used ---")
    print(src["_cov_blocks"])


if __name__ == "__main__":
    main()
