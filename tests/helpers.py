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


def pcg_init(V, D, seed):
    """The initial tensors of the CBOW goldens: what oracle/tf1_shim.truncated_normal draws for
    tf.truncated_normal([V, D], stddev=1/sqrt(D)) then ([D, 1]) (G2Vec.py:234-235) after seed_initialisers(seed):
    PCG64(seed) standard normals, re-drawn while |x| > 2, times stddev, float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for shape in ((V, D), (D,)):
        x = rng.standard_normal(size=shape)
        bad = np.abs(x) > 2.0
        while bad.any():
            x[bad] = rng.standard_normal(size=int(bad.sum()))
            bad = np.abs(x) > 2.0
        out.append((x * (1.0 / np.sqrt(D))).astype(np.float32))
    return out[0], out[1]


def cbow_golden(name):
    """tests/golden/cbow_*.npz -> dict with the windows, split, init and the reference's outputs."""
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    V, D, seed = (int(x) for x in z["meta"])
    W0, Wo0 = pcg_init(V, D, seed)
    assert (Wo0 == z["W_ho0"]).all()
    chk = np.array([W0.astype(np.float64).sum(), np.abs(W0.astype(np.float64)).sum()])
    assert np.allclose(chk, z["W_ih0_check"], rtol=0, atol=1e-9)
    want = W0.copy()
    want[z["W_touched_rows"]] = z["W_touched"]
    N = len(z["rowptr"]) - 1
    tr, va = oracle.split_indices(N, seed)
    return {"rowptr": z["rowptr"], "gene": z["gene"].astype(np.int32), "label": z["label"], "V": V, "D": D, "seed": seed,
            "lr": float(z["lr"]), "W0": W0, "Wo0": Wo0, "W_ref": want, "acc_val": z["acc_val"], "acc_tr": z["acc_tr"],
            "stop_step": int(z["stop_step"]), "tr": tr, "va": va, "log": str(z["log"])}
