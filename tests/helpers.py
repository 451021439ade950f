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


def write_ex_tsv(tmp_path):
    """The ex_* inputs as the three TSV files the command line reads, rebuilt from the golden fixtures
    (restricted expression + edge list), plus one gene and one edge that step 2 must drop."""
    e = np.load(os.path.join(GOLDEN, "ex_expr.npz"))
    gr = np.load(os.path.join(GOLDEN, "ex_graph.npz"))
    genes = e["gene"]
    samples = ["TCGA-%04d" % i for i in range(135)]
    ef, cf, nf = (os.path.join(str(tmp_path), n) for n in ("E.txt", "C.txt", "N.txt"))
    with open(ef, "w") as f:
        f.write("PATIENT\t" + "\t".join(samples) + "\n")
        for g, col in zip(genes, e["expr"].T):
            f.write(g + "\t" + "\t".join(repr(float(x)) for x in col) + "\n")
        f.write("NOT_IN_NETWORK\t" + "\t".join("0.5" for _ in samples) + "\n")
    with open(cf, "w") as f:
        f.write("PATIENT\tLABEL\n")
        f.writelines("%s\t%d\n" % (s, l) for s, l in zip(samples, gr["label"]))
    with open(nf, "w") as f:
        f.write("src\tdest\n")
        f.writelines("%s\t%s\n" % (genes[a], genes[b]) for a, b in zip(e["src"], e["dst"]))
        # genes whose only partners are outside the expression data: in the network's gene set (so they stay
        # in the common gene list, 7523) while the edge itself is dropped by step 2
        seen = np.zeros(len(genes), bool); seen[e["src"]] = True; seen[e["dst"]] = True
        f.writelines("%s\tNOT_IN_EXPRESSION\n" % g for g in genes[~seen])
        f.write("%s\tNOT_IN_EXPRESSION\n" % genes[0])
    return ef, cf, nf, genes
