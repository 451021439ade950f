"""GPU glue between the hot paths (canonicalise / dedup / cross-group removal / CSR windows /
gene-frequency vote) against the oracle's restatement of G2Vec.py:288-322,345,351."""
import pytest

import oracle
from oracle import legacy
from tests import helpers

pytestmark = pytest.mark.gpu


def test_pipeline_equals_oracle_on_ex():
    import g2vec_b200 as g2v
    from g2vec_b200 import paths
    reps, L, seed = 2, 80, 0
    (o_rowptr, o_gene, o_label), o_rows = helpers.ex_windows(reps=reps, L=L, seed=seed)
    rows = []
    for g in (0, 1):
        rp, col, w = helpers.ex_graph(g)
        wg = g2v.WalkGraph(rp, col, weights=w)
        nodes, lens = g2v.generate_paths(wg, L, reps, seed=seed, group=g)
        rows.append(paths.canonical_rows(nodes, lens))
    prow, plab = paths.integrate(rows[0], rows[1])
    got = {(int(l), tuple(int(x) for x in r[r != paths.PAD])) for r, l in zip(prow.cpu().numpy(), plab.cpu().numpy())}
    assert got == set(o_rows)
    rowptr, gene, label = paths.windows_csr(prow, plab)
    assert int(rowptr[-1]) == len(o_gene) and rowptr.shape[0] == len(o_rowptr)
    code = paths.gene_freq_codes(rowptr, gene, label, 7523).cpu().numpy()
    want = legacy.count_geneFreq(o_rows)
    assert {i: int(c) for i, c in enumerate(code) if c >= 0} == want
    # the reference-shaped adapter
    rp, col, w = helpers.ex_graph(0)
    ps = g2v.generate_pathSet(g2v.WalkGraph(rp, col, weights=w), 80, 1, seed=5, group=0)
    nodes, lens = oracle.walks(rp, col, oracle.quantise_weights(w), 80, 5, 0, 0, len(rp) - 1)
    assert ps == oracle.path_set(nodes, lens)
