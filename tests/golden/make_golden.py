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
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle.legacy import csr_from_dense  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = ref_import.load()


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
    main()
