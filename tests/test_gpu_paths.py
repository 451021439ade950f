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


def test_mark_kernel_is_exact_under_key_collisions():
    """g2v_paths_mark compares full rows inside a key run, so set semantics hold even if different rows share
    a key: feed it deliberately colliding keys (everything in one run, and 7 runs) and compare with Python sets."""
    import numpy as np
    import torch
    from g2vec_b200 import _capi, paths
    lib = _capi.load()
    rs = np.random.RandomState(5)
    base = [tuple(sorted(rs.choice(40, size=rs.randint(1, 9), replace=False))) for _ in range(60)]
    picks = [base[i] for i in rs.randint(0, 60, size=400)]
    L = 8
    rows = np.full((400, L), paths.PAD, np.int32)
    for i, p in enumerate(picks):
        rows[i, :len(p)] = p
    grp = (rs.rand(400) < 0.5).astype(np.uint8)
    for nkeys in (1, 7):
        key = np.array([hash(p) % nkeys for p in picks], dtype=np.int64)        # equal rows -> equal keys
        r_d = torch.from_numpy(rows).cuda(); k_d = torch.from_numpy(key).cuda(); g_d = torch.from_numpy(grp).cuda()
        ks, perm = torch.sort(k_d, stable=True)
        flag = torch.empty(400, dtype=torch.uint8, device="cuda")
        _capi.check(lib.g2v_paths_mark(r_d.data_ptr(), ks.data_ptr(), perm.data_ptr(), 0, 400, L, flag.data_ptr(), 0), "mark")
        torch.cuda.synchronize()
        kept = [picks[i] for i in perm[flag.bool()].cpu().numpy()]
        assert len(kept) == len(set(kept)) and set(kept) == set(picks)          # first occurrences only, all of them
        _capi.check(lib.g2v_paths_mark(r_d.data_ptr(), ks.data_ptr(), perm.data_ptr(), g_d.data_ptr(), 400, L,
                                       flag.data_ptr(), 0), "mark")
        torch.cuda.synchronize()
        s0 = {p for p, g in zip(picks, grp) if g == 0}; s1 = {p for p, g in zip(picks, grp) if g == 1}
        surv = perm[flag.bool()].cpu().numpy()
        assert {(int(grp[i]), picks[i]) for i in surv} == {(0, p) for p in s0 - s1} | {(1, p) for p in s1 - s0}


def test_canonicalise_sorts_rows_of_any_length():
    import numpy as np
    import torch
    from g2vec_b200 import paths
    rs = np.random.RandomState(1)
    for L in (1, 2, 31, 32, 33, 80, 160, 1000):
        n = 50
        nodes = np.full((n, L), -1, np.int32)
        for i in range(n):
            k = rs.randint(1, L + 1)
            nodes[i, :k] = rs.choice(5000, size=k, replace=False)
        nodes[7] = nodes[3]                                   # a duplicate in a different visit order
        nodes[7, :(nodes[3] >= 0).sum()] = rs.permutation(nodes[3][nodes[3] >= 0])
        rows, key = paths._canon(torch.from_numpy(nodes).cuda())
        rows = rows.cpu().numpy(); key = key.cpu().numpy()
        want = np.sort(np.where(nodes < 0, paths.PAD, nodes), axis=1)
        assert (rows == want).all() and (key >= 0).all() and key[7] == key[3]
        uniq = paths.canonical_rows(torch.from_numpy(nodes).cuda())
        assert paths.rows_to_set(uniq) == {tuple(int(x) for x in r[r != paths.PAD]) for r in want}


def _walk_both_groups(g2v, reps, L, seed):
    """Canonical rows of both ex_* groups in one buffer: (rows, lens, key, group) device tensors."""
    import torch
    V = 7523
    n = reps * V
    rows = torch.empty((2 * n, L), dtype=torch.int32, device="cuda")
    lens = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    key = torch.empty(2 * n, dtype=torch.int64, device="cuda")
    for g in (0, 1):
        rp, col, w = helpers.ex_graph(g)
        wg = g2v.WalkGraph(rp, col, weights=w)
        g2v.generate_paths(wg, L, reps, seed=seed, group=g, canonical=True,
                           out=(rows[g * n:(g + 1) * n], lens[g * n:(g + 1) * n], key[g * n:(g + 1) * n]))
    group = torch.cat([torch.zeros(n, dtype=torch.uint8, device="cuda"), torch.ones(n, dtype=torch.uint8, device="cuda")])
    return rows, lens, key, group


def test_sort_free_pipeline_equals_oracle_on_ex():
    """g2v_paths_set_select / _emit (hash table on the row keys, prefix sums, one emit kernel): the windows are the
    oracle's set of (label, path), in first-occurrence order, and the gene-frequency codes are count_geneFreq's."""
    import numpy as np
    import g2vec_b200 as g2v
    from g2vec_b200 import paths
    reps, L, seed = 2, 80, 0
    (o_rowptr, o_gene, o_label), o_rows = helpers.ex_windows(reps=reps, L=L, seed=seed)
    rows, lens, key, group = _walk_both_groups(g2v, reps, L, seed)
    rowptr, gene, label, code = paths.build_windows(rows, lens, key, group, 7523)
    rp, ge, la = rowptr.cpu().numpy(), gene.cpu().numpy(), label.cpu().numpy()
    got = [(int(la[i]), tuple(int(x) for x in ge[rp[i]:rp[i + 1]])) for i in range(len(la))]
    assert len(got) == len(set(got)) == len(o_rows) and set(got) == set(o_rows)
    assert rp[0] == 0 and rp[-1] == len(o_gene) == len(ge)
    # input order = first occurrence in (group, repetition, start gene) order: group 0 windows first, and within a
    # group the walker index of the first walk that produced each path is increasing
    assert (np.diff(la.astype(np.int8)) >= 0).all()
    r_h, l_h, g_h = rows.cpu().numpy(), lens.cpu().numpy(), group.cpu().numpy()
    first = {}
    for i in range(len(l_h)):
        first.setdefault((int(g_h[i]), tuple(int(x) for x in r_h[i, :l_h[i]])), i)
    order = [first[w] for w in got]
    assert order == sorted(order)
    want = legacy.count_geneFreq(o_rows)
    c = code.cpu().numpy()
    assert {i: int(x) for i, x in enumerate(c) if x >= 0} == want


def test_sort_free_pipeline_small_cases_and_collision_fallback(monkeypatch):
    """Hand-made rows: duplicates inside a group, a path common to both groups (removed from both), empty input; and
    keys that collide on purpose -- different rows, same key -- which must be detected and routed to the exact path."""
    import numpy as np
    import torch
    from g2vec_b200 import paths
    P = paths.PAD
    rows = np.array([[1, 2, 3, P], [1, 2, 3, P], [4, P, P, P], [2, 5, P, P],      # group 0: a duplicate
                     [4, P, P, P], [2, 5, 6, P], [7, 8, 9, 10], [7, 8, 9, 10]], dtype=np.int32)   # group 1: (4,) is common
    lens = np.array([3, 3, 1, 2, 1, 3, 4, 4], dtype=np.int32)
    group = np.array([0, 0, 0, 0, 1, 1, 1, 1], dtype=np.uint8)
    want = [(0, (1, 2, 3)), (0, (2, 5)), (1, (2, 5, 6)), (1, (7, 8, 9, 10))]
    r_d, l_d, g_d = (torch.from_numpy(a).cuda() for a in (rows, lens, group))
    _, key = paths._canon(r_d)
    for k_d in (key, torch.full_like(key, 12345)):                  # real keys; then everything on ONE key
        rowptr, gene, label, code = paths.build_windows(r_d, l_d, k_d, g_d, 12)
        rp, ge, la = rowptr.cpu().numpy(), gene.cpu().numpy(), label.cpu().numpy()
        got = [(int(la[i]), tuple(int(x) for x in ge[rp[i]:rp[i + 1]])) for i in range(len(la))]
        assert sorted(got) == sorted(want)
        c = code.cpu().numpy()
        assert list(c) == [-1, 0, 0, 0, -1, 2, 1, 1, 1, 1, 1, -1]     # gene 5: one good + one poor path = tie; 4: only in the common path
    e = torch.empty((0, 4), dtype=torch.int32, device="cuda")
    rowptr, gene, label, code = paths.build_windows(e, l_d[:0], key[:0], g_d[:0], 12)
    assert rowptr.cpu().tolist() == [0] and gene.numel() == 0 and label.numel() == 0 and (code.cpu().numpy() == -1).all()
