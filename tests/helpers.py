"""Shared builders for the tests (CPU side: oracle only)."""
import os

import numpy as np

import oracle
from oracle import legacy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def ex_graph(group):
    z = _cache.setdefault("ex", np.load(os.path.join(GOLDEN, "ex_graph.npz")))
    return z["rowptr%d" % group], z["col%d" % group], z["w%d" % group]


def ex_windows(reps=2, L=80, seed=0):
    """Oracle pipeline on the ex_* graphs: walks -> sets -> integrate -> CSR windows."""
    key = ("win", reps, L, seed)
    if key not in _cache:
        sets = []
        for g in (0, 1):
            rp, col, w = ex_graph(g)
            nodes, lens = oracle.walks(rp, col, oracle.quantise_weights(w), L, seed, g, 0, reps * (len(rp) - 1))
            sets.append(oracle.path_set(nodes, lens))
        rows = legacy.integrate_pathSet(sets)
        _cache[key] = (legacy.windows_from_rows(rows), rows)
    return _cache[key]


def random_graph(V, deg, seed, dead_frac=0.1):
    rs = np.random.RandomState(seed)
    src, dst = [], []
    for v in range(V):
        if rs.rand() < dead_frac:
            continue
        k = min(V - 1, max(1, rs.poisson(deg)))
        nb = rs.choice(V - 1, size=k, replace=False)
        nb = np.sort(np.where(nb >= v, nb + 1, nb))
        src += [v] * k; dst += list(nb)
    src = np.array(src, dtype=np.int32); dst = np.array(dst, dtype=np.int32)
    rowptr = np.zeros(V + 1, dtype=np.int32)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    w = (0.5 + 0.5 * rs.rand(len(dst))).astype(np.float32) + np.float32(1e-4)
    return rowptr, dst, w


def random_windows(N, V, lmin, lmax, seed):
    rs = np.random.RandomState(seed)
    lens = rs.randint(lmin, lmax + 1, size=N)
    rowptr = np.zeros(N + 1, dtype=np.int32); rowptr[1:] = np.cumsum(lens)
    gene = np.concatenate([np.sort(rs.choice(V, size=l, replace=False)) for l in lens]).astype(np.int32) \
        if N else np.zeros(0, np.int32)
    label = (rs.rand(N) < 0.5).astype(np.uint8)
    return rowptr, gene, label


def init_weights(V, D, seed):
    rs = np.random.RandomState(seed)
    s = 1.0 / np.sqrt(D)
    return (np.clip(rs.randn(V, D), -2, 2) * s).astype(np.float32), (np.clip(rs.randn(D), -2, 2) * s).astype(np.float32)
