"""TEST INFRASTRUCTURE ONLY -- a stand-in for the `opt_einsum` package.

The reference (`/root/reference/psgd.py:42`) imports `opt_einsum`, which is neither
installed nor vendored, and there is no network.  psgd.py uses exactly two entry points,
`contract_expression(subscripts, *shapes)` (psgd.py:192-195,223,226,243,246,249,252) and
`get_symbol(i)` (psgd.py:212-245); misc/psgd_kron_verification.py:220 also uses `contract`.

This file contains NO PSGD arithmetic: it only orders pairwise contractions (greedy,
smallest-intermediate first, never an outer product) and replays them with `torch.einsum`.
Contraction order changes rounding only, never the math.

Usage (torch must be imported BEFORE the shim is installed, otherwise torch.einsum would
try to use it as its path optimiser):

    import torch
    from oracle.opt_einsum_shim import install; install()
    import psgd   # the unmodified reference
"""
import sys
import types

import torch

_BASE = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i):
    # same mapping as the real package: 0..51 -> letters, beyond -> chr(i + 140) (805 -> 'α' ...)
    return _BASE[i] if i < 52 else chr(i + 140)


def _plan(terms, out, sizes):
    """Greedy pairwise contraction path over index sets. Returns list of (i, j, result_term)."""
    terms = list(terms)
    steps = []
    while len(terms) > 1:
        best = None
        for i in range(len(terms)):
            for j in range(i + 1, len(terms)):
                si, sj = set(terms[i]), set(terms[j])
                if not (si & sj) and (terms[i] != "" and terms[j] != ""):
                    continue  # would be an outer product
                others = set(out)
                for k, t in enumerate(terms):
                    if k != i and k != j:
                        others |= set(t)
                keep = [c for c in dict.fromkeys(terms[i] + terms[j]) if c in others]
                size = 1
                for c in keep:
                    size *= sizes[c]
                cost = 1
                for c in si | sj:
                    cost *= sizes[c]
                key = (size, cost)
                if best is None or key < best[0]:
                    best = (key, i, j, "".join(keep))
        if best is None:  # only outer products remain
            i, j = 0, 1
            others = set(out)
            for k, t in enumerate(terms):
                if k > 1:
                    others |= set(t)
            keep = [c for c in dict.fromkeys(terms[0] + terms[1]) if c in others]
            best = (None, i, j, "".join(keep))
        _, i, j, res = best
        steps.append((i, j, res))
        terms = [t for k, t in enumerate(terms) if k != i and k != j] + [res]
    return steps


def contract_expression(subscripts, *shapes, **_kw):
    subscripts = subscripts.replace(" ", "")
    lhs, out = subscripts.split("->")
    terms = lhs.split(",")
    assert len(terms) == len(shapes), (subscripts, shapes)
    # remap every distinct symbol to [a-zA-Z]: torch.einsum accepts letters only
    symbols = list(dict.fromkeys("".join(terms) + out))
    assert len(symbols) <= 52
    remap = {s: _BASE[k] for k, s in enumerate(symbols)}
    terms = ["".join(remap[c] for c in t) for t in terms]
    out = "".join(remap[c] for c in out)
    sizes = {}
    for t, shp in zip(terms, shapes):
        assert len(t) == len(shp), (subscripts, shapes)
        for c, n in zip(t, shp):
            sizes[c] = int(n)
    steps = _plan(terms, out, sizes)

    def expr(*ops):
        ops = list(ops)
        cur = list(terms)
        if len(ops) == 1:
            return torch.einsum(cur[0] + "->" + out, ops[0])
        for (i, j, res) in steps:
            a, b = ops[i], ops[j]
            ta, tb = cur[i], cur[j]
            r = torch.einsum(f"{ta},{tb}->{res}", a, b)
            ops = [o for k, o in enumerate(ops) if k != i and k != j] + [r]
            cur = [t for k, t in enumerate(cur) if k != i and k != j] + [res]
        final = ops[0]
        if cur[0] != out:
            final = torch.einsum(cur[0] + "->" + out, final)
        return final

    return expr


def contract(subscripts, *ops, **kw):
    return contract_expression(subscripts, *[o.shape for o in ops])(*ops)


def install():
    m = types.ModuleType("opt_einsum")
    m.get_symbol = get_symbol
    m.contract_expression = contract_expression
    m.contract = contract
    sys.modules["opt_einsum"] = m
    return m


def load_reference(path="/root/reference"):
    """Import the unmodified reference psgd.py (only possible where /root/reference exists)."""
    import importlib
    import os

    if not os.path.isdir(path):
        raise FileNotFoundError(f"{path} is not present (it does not travel to the GPU box)")
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    # the product package also has a module named psgd; keep the reference under its own name
    spec = importlib.util.spec_from_file_location("psgd_reference", os.path.join(path, "psgd.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
