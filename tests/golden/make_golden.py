"""Generate golden fixtures by RUNNING THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/G2Vec.py unmodified (oracle/ref_import.py, TensorFlow stubbed) and
records outputs of its own functions:

  walk_small.npz   generate_pathSet (G2Vec.py:324-352) under np.random.seed on small dense
                   graphs (isolated nodes, dead ends, full-length walks, L=1).
  glue_small.npz   integrate_pathSet / count_geneFreq (G2Vec.py:288-322) on the sets above.
  pcc_small.npz    construct_adjMat (G2Vec.py:370-391) on a small expression matrix.
  ex_expr.npz      ex_* restricted expression matrix [135, 7523] + the 216 540 restricted edges as gene
                   indices (the inputs of construct_adjMat after the reference's steps 1-2).
  ex_graph.npz     ex_* data through the reference's steps 1-3a: n_samples/n_genes/n_edges
                   (README.md:26-28), per-group CSR of construct_adjMat, and
                   generate_pathSet(adjMat, 80, 1) under np.random.seed(0) for group 0.
  cbow_small.npz   the UNMODIFIED compute_genetovec (G2Vec.py:217-286) run on oracle/tf1_shim.py (the TF 1.x
  cbow_ex.npz      ops it calls, restated on torch-CPU; TensorFlow itself is not installable here) under
                   np.random.seed(seed): the dense pathList it was given (as CSR windows), the initial tensors
                   tf.truncated_normal drew, the returned W_ih, every per-step ACC[val]/ACC[tr] the loop
                   computed, the stop step and the printed log.  `python tests/golden/make_golden.py cbow`
                   regenerates only these two.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle.legacy import csr_from_dense  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = ref_import.load()


def run_reference_cbow(rowptr, gene, label, V, D, lr, seed):
    """compute_genetovec(pathList, n_genes, hidden_size, learning_rate) exactly as main() calls it (G2Vec.py:74)
    on the dense int32 pathList of integrate_pathSet (:310-322); np.random.seed(seed) fixes its shuffle (:219),
    tf1_shim.seed_initialisers(seed) its two truncated_normal draws (:234-235)."""
    import contextlib
    import io
    from oracle import tf1_shim
    N = len(rowptr) - 1
    P = np.zeros((N, V + 1), dtype=np.int32)
    for n in range(N):
        P[n, gene[rowptr[n]:rowptr[n + 1]]] = 1
    P[:, -1] = label
    tf1_shim.reset()
    tf1_shim.seed_initialisers(seed)
    np.random.seed(seed)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        W = ref.compute_genetovec(P, V, D, lr)
    accs = [float(e[1]) for e in tf1_shim.trace() if e[0] == "eval"]       # val, tr, val, tr, ... (:266-267)
    W0, Wo0 = tf1_shim.initial_values()
    log = re.sub(r"\([0-9.]+ sec\)", "(T sec)", buf.getvalue())
    return W, np.array(accs[0::2], dtype=np.float32), np.array(accs[1::2], dtype=np.float32), W0, Wo0.reshape(-1), log


def learnable_windows(N, V, lmin, lmax, seed, noise):
    """Small windows whose label is predictable from the genes (so the run trains for a while before it stops)."""
    rs = np.random.RandomState(seed)
    label = (rs.rand(N) < 0.5).astype(np.uint8)
    lens = rs.randint(lmin, lmax + 1, size=N)
    lens[:3] = 0                                  # empty windows (all-zero rows of the dense pathList)
    rowptr = np.zeros(N + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(lens)
    genes, half = [], V // 2
    for y, l in zip(label, lens):
        p = np.full(V, noise / V)
        lo = 0 if y == 0 else half
        p[lo:lo + half] += (1 - noise) / half
        genes.append(np.sort(rs.choice(V, size=l, replace=False, p=p / p.sum())))
    return rowptr, np.concatenate(genes).astype(np.int32), label


def save_cbow(name, rowptr, gene, label, V, D, lr, seed):
    W, acc_val, acc_tr, W0, Wo0, log = run_reference_cbow(rowptr, gene, label, V, D, lr, seed)
    touched = np.flatnonzero(np.abs(W - W0).max(axis=1) > 0)
    stopped = "Epoch(stop)" in log
    print(name, "windows", len(rowptr) - 1, "evaluated steps", len(acc_val), "stopped", stopped, "touched rows", len(touched))
    print(log)
    np.savez_compressed(os.path.join(OUT, name), rowptr=rowptr.astype(np.int32), gene=gene.astype(np.int16),
                        label=label.astype(np.uint8), meta=np.array([V, D, seed], dtype=np.int64), lr=np.float64(lr),
                        W_touched_rows=touched.astype(np.int32), W_touched=W[touched], acc_val=acc_val, acc_tr=acc_tr,
                        # step whose validation accuracy dropped (loop index at the break, G2Vec.py:276-279); -1 = none
                        stop_step=np.int64(len(acc_val) - 1 if stopped else -1), log=np.array(log),
                        # the initial tensors are PCG64(seed) truncated normals (tests/helpers.pcg_init restates the
                        # rule); W_ho and two float64 checksums of W_ih pin them without storing V*D floats
                        W_ho0=Wo0, W_ih0_check=np.array([W0.astype(np.float64).sum(), np.abs(W0.astype(np.float64)).sum()]))


def cbow_goldens():
    rowptr, gene, label = learnable_windows(1500, 120, 1, 6, 5, 0.7)
    save_cbow("cbow_small.npz", rowptr, gene, label, 120, 32, 0.005, 7)
    from tests import helpers
    (rowptr, gene, label), _ = helpers.ex_windows(reps=2)
    save_cbow("cbow_ex.npz", rowptr, gene, label, 7523, 128, 0.005, 0)


def pack_paths(ps):
    rows = sorted(ps)
    lens = np.array([len(p) for p in rows], dtype=np.int32)
    flat = np.array([g for p in rows for g in p], dtype=np.int32)
    return flat, lens


def small_graphs():
    rs = np.random.RandomState(1234)
    cases = []
    # (V, density, L, iters, seed)
    for V, dens, L, iters, seed in [(12, 0.25, 5, 3, 1), (30, 0.10, 8, 2, 2), (60, 0.08, 80, 2, 3),
                                    (25, 0.5, 25, 2, 4), (10, 0.3, 1, 2, 5), (40, 0.05, 6, 3, 6)]:
        A = np.zeros((V, V), dtype=np.float32)
        mask = rs.rand(V, V) < dens
        np.fill_diagonal(mask, False)
        A[mask] = (0.5 + 0.5 * rs.rand(int(mask.sum()))).astype(np.float32) + np.float32(1e-4)
        A[V // 3, :] = 0.      # a node with no out-edges (dead end / singleton path)
        A[:, V // 2] = 0.      # a node nobody reaches
        cases.append((A, L, iters, seed))
    return cases


def main():
    # ---- walks on small graphs
    d = {}
    sets = []
    for i, (A, L, iters, seed) in enumerate(small_graphs()):
        np.random.seed(seed)
        ps = ref.generate_pathSet(A, L, iters)
        sets.append(ps)
        flat, lens = pack_paths(ps)
        d["A%d" % i] = A
        d["meta%d" % i] = np.array([L, iters, seed], dtype=np.int64)
        d["flat%d" % i] = flat
        d["lens%d" % i] = lens
    d["n_cases"] = np.array(len(sets))
    np.savez_compressed(os.path.join(OUT, "walk_small.npz"), **d)

    # ---- glue: pair the sets of cases (0,1)... need same V: build two sets on the same graph
    A, L, iters, _ = small_graphs()[2]
    np.random.seed(11); s0 = ref.generate_pathSet(A, L, iters)
    B = A.copy(); B[::2] = 0.
    np.random.seed(12); s1 = ref.generate_pathSet(B, L, iters)
    V = A.shape[0]
    pl = ref.integrate_pathSet([s0, s1], V)
    genes = np.array(["G%03d" % i for i in range(V)])
    gf = ref.count_geneFreq(pl, genes)
    f0, l0 = pack_paths(s0); f1, l1 = pack_paths(s1)
    order = np.lexsort(pl.T[::-1])
    np.savez_compressed(os.path.join(OUT, "glue_small.npz"), V=np.array(V), f0=f0, l0=l0, f1=f1, l1=l1,
                        pathList_sorted=pl[order].astype(np.int8),
                        gf_gene=np.array(sorted(gf.keys())),
                        gf_val=np.array([gf[k] for k in sorted(gf.keys())], dtype=np.int8))

    # ---- PCC adjacency on a small expression matrix
    rs = np.random.RandomState(77)
    S, G = 40, 30
    base = rs.randn(S, 6).astype(np.float32)
    expr = (base[:, rs.randint(0, 6, G)] + 0.6 * rs.randn(S, G)).astype(np.float32)
    expr[:, 5] = 1.25   # zero-variance gene -> weight 0
    label = (rs.rand(S) < 0.45).astype(np.int64)
    genes = np.array(["G%03d" % i for i in range(G)])
    edges = [[genes[a], genes[b]] for a, b in rs.randint(0, G, (300, 2)) if a != b]
    data = {"gene": genes, "expr": expr, "label": label}
    adj = [ref.construct_adjMat(edges, data, g) for g in (0, 1)]
    g2i = {g: i for i, g in enumerate(genes)}
    np.savez_compressed(os.path.join(OUT, "pcc_small.npz"), expr=expr, label=label,
                        src=np.array([g2i[e[0]] for e in edges], dtype=np.int32),
                        dst=np.array([g2i[e[1]] for e in edges], dtype=np.int32),
                        adj0=adj[0], adj1=adj[1])

    # ---- ex_* through the reference
    data = ref.load_data(os.path.join(ref_import.REF_DIR, "ex_EXPRESSION.txt"))
    clinical = ref.load_clinical(os.path.join(ref_import.REF_DIR, "ex_CLINICAL.txt"))
    network = ref.load_network(os.path.join(ref_import.REF_DIR, "ex_NETWORK.txt"))
    data["label"] = ref.match_labels(clinical, data["sample"])
    common = ref.find_commonGeneList(network["gene"], data["gene"])
    network = ref.restrict_network(network, common)
    data = ref.restrict_data(data, common)
    n_samples, n_genes = data["expr"].shape
    n_edges = len(network["edge"])
    print("n_samples/n_genes/n_edges", n_samples, n_genes, n_edges)
    out = {"shape": np.array([n_samples, n_genes, n_edges], dtype=np.int64),
           "label": data["label"].astype(np.int8)}
    for g in (0, 1):
        A = ref.construct_adjMat(network["edge"], data, g)
        rp, col, w = csr_from_dense(A)
        out["rowptr%d" % g] = rp; out["col%d" % g] = col; out["w%d" % g] = w
        print("group", g, "nnz", len(col))
        if g == 0:
            np.random.seed(0)
            ps = ref.generate_pathSet(A, 80, 1)
            flat, lens = pack_paths(ps)
            out["ps0_flat"] = flat; out["ps0_lens"] = lens
            print("group 0 paths (1 repetition, seed 0):", len(ps))
    np.savez_compressed(os.path.join(OUT, "ex_graph.npz"), **out)
    # the inputs of construct_adjMat after the reference's steps 1-2 (restricted expression + edge list as
    # gene indices), so that the GPU box can run BASELINE configs[0] from the same data without /root/reference
    g2i = {g: i for i, g in enumerate(data["gene"])}
    np.savez_compressed(os.path.join(OUT, "ex_expr.npz"), expr=data["expr"].astype(np.float32),
                        src=np.array([g2i[e[0]] for e in network["edge"]], dtype=np.int16),
                        dst=np.array([g2i[e[1]] for e in network["edge"]], dtype=np.int16),
                        gene=np.array(data["gene"]))


if __name__ == "__main__":
    if sys.argv[1:] == ["cbow"]:
        cbow_goldens()
    else:
        main()
        cbow_goldens()
