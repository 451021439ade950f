"""Statistical parity of the Philox integer draw with the reference's np.random.choice draw
(G2Vec.py:341): the two cannot be bit-compared (the reference consumes one global MT19937 stream
sequentially), so transition frequencies and path-length histograms are compared by chi-square.
The legacy side is oracle.legacy.walks_generic + LegacyDraw, which test_oracle_pin.py pins bit-exact to
the reference's own generate_pathSet."""
import numpy as np
from scipy import stats

import oracle
from oracle import legacy


def graph():
    rs = np.random.RandomState(42)
    V = 24
    A = np.zeros((V, V), dtype=np.float32)
    mask = rs.rand(V, V) < 0.3
    np.fill_diagonal(mask, False)
    A[mask] = (0.5 + 0.5 * rs.rand(int(mask.sum()))).astype(np.float32) + np.float32(1e-4)
    A[0, 1:9] = np.linspace(0.51, 1.0, 8).astype(np.float32)     # node 0: 8+ out-edges, spread weights
    return A


def test_first_transition_and_length_distributions_agree():
    A = graph()
    V = A.shape[0]
    rp, col, w = legacy.csr_from_dense(A)
    q = oracle.quantise_weights(w)
    L, reps = 6, 1500
    ids = [r * V for r in range(reps)]                            # walker ids whose start node is 0
    leg = legacy.walks_generic(rp, col, w, L, ids, legacy.LegacyDraw(V, 7), consume_last=True)
    nodes, lens = oracle.walks(rp, col, q, L, 7, 0, 0, reps * V, V)
    phi = [list(map(int, r[:n])) for r, n in zip(nodes, lens)]
    nb = col[rp[0]:rp[1]]
    p = w[rp[0]:rp[1]].astype(np.float64); p /= p.sum()
    for name, walks in (("legacy", leg), ("philox", phi)):
        first = np.array([x[1] for x in walks])
        obs = np.array([(first == c).sum() for c in nb])
        chi2 = ((obs - reps * p) ** 2 / (reps * p)).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-4, len(nb) - 1), (name, chi2)
    # second transitions pooled over the first node: two-sample chi-square legacy vs philox
    a = np.bincount([x[2] for x in leg if len(x) > 2], minlength=V).astype(np.float64)
    b = np.bincount([x[2] for x in phi if len(x) > 2], minlength=V).astype(np.float64)
    m = (a + b) > 10
    chi2 = (((a[m] - b[m]) ** 2) / (a[m] + b[m])).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-4, int(m.sum()) - 1)
    la = np.bincount([len(x) for x in leg], minlength=L + 1)[1:]
    lb = np.bincount([len(x) for x in phi], minlength=L + 1)[1:]
    m = (la + lb) > 10
    chi2 = (((la[m] - lb[m]) ** 2) / (la[m] + lb[m])).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-4, max(int(m.sum()) - 1, 1))


def test_quantisation_error_is_far_below_sampling_noise():
    w = (0.5 + 0.5 * np.random.RandomState(0).rand(1000)).astype(np.float32) + np.float32(1e-6)
    q = oracle.quantise_weights(w).astype(np.float64)
    p, pq = w.astype(np.float64) / w.sum(dtype=np.float64), q / q.sum()
    assert np.abs(p - pq).max() / p.min() < 2e-5


def test_distribution_on_a_dense_graph_where_most_neighbours_are_visited():
    """A small dense graph: late in a walk most neighbours are already visited, so the masked renormalised
    draw matters.  Node-at-position distributions must match the reference's draw."""
    rs = np.random.RandomState(8)
    V = 10
    A = ((rs.rand(V, V) < 0.85) * (0.5 + 0.5 * rs.rand(V, V))).astype(np.float32)
    np.fill_diagonal(A, 0)
    rp, col, w = legacy.csr_from_dense(A)
    q = oracle.quantise_weights(w)
    L, reps = 8, 3000
    ids = [r * V + 3 for r in range(reps)]                       # all walkers start at node 3
    leg = legacy.walks_generic(rp, col, w, L, ids, legacy.LegacyDraw(V, 11), consume_last=True)
    nodes, lens = oracle.walks(rp, col, q, L, 11, 1, 3, reps * V, V)
    phi = [list(map(int, r[:n])) for r, n in zip(nodes, lens)]
    for pos in (1, 4, 6, 7):
        a = np.bincount([x[pos] for x in leg if len(x) > pos], minlength=V).astype(np.float64)
        b = np.bincount([x[pos] for x in phi if len(x) > pos], minlength=V).astype(np.float64)
        m = (a + b) > 10
        chi2 = (((a[m] - b[m]) ** 2) / (a[m] + b[m])).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-4, int(m.sum()) - 1), (pos, chi2)
