"""Pin the oracle against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference/G2Vec.py) and against published Philox
known-answer vectors.  CPU only."""
import os

import numpy as np
import pytest

import oracle
from oracle import legacy


def unpack(flat, lens):
    out, o = set(), 0
    for n in lens:
        out.add(tuple(int(x) for x in flat[o:o + n]))
        o += n
    return out


# Random123 kat_vectors, philox4x32-10
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,exp", PHILOX_KAT)
def test_philox_known_answers(ctr, key, exp):
    assert tuple(oracle.philox4x32_10(ctr, key)) == exp
    assert tuple(oracle.philox4x32_10_c(ctr, key)) == exp


def test_draw64_c_equals_python():
    rs = np.random.RandomState(0)
    for _ in range(200):
        seed = int(rs.randint(0, 2**62)); sub = int(rs.randint(0, 2**45)); s = int(rs.randint(0, 400))
        assert oracle.draw64(seed, sub, s) == oracle.draw64_py(seed, sub, s)


def test_legacy_walks_equal_reference_small(golden_dir):
    z = np.load(os.path.join(golden_dir, "walk_small.npz"))
    for i in range(int(z["n_cases"])):
        A = z["A%d" % i]; L, iters, seed = (int(x) for x in z["meta%d" % i])
        want = unpack(z["flat%d" % i], z["lens%d" % i])
        got_dense = legacy.generate_pathSet_dense(A, L, iters, np.random.RandomState(seed))
        assert got_dense == want, "dense port, case %d" % i
        rp, col, w = legacy.csr_from_dense(A)
        got_csr = legacy.generate_pathSet_csr(rp, col, w, L, iters, seed)
        assert got_csr == want, "CSR walk logic + legacy draw, case %d" % i


def test_legacy_walks_equal_reference_ex(golden_dir):
    """ex_* group 0, lenPath 80, 1 repetition, np.random.seed(0): the CSR walk logic with the
    legacy draw reproduces the reference's 7391 paths exactly."""
    z = np.load(os.path.join(golden_dir, "ex_graph.npz"))
    assert tuple(z["shape"]) == (135, 7523, 216540)          # README.md:26-28
    want = unpack(z["ps0_flat"], z["ps0_lens"])
    got = legacy.generate_pathSet_csr(z["rowptr0"], z["col0"], z["w0"], 80, 1, 0)
    assert len(got) == len(want) == 7391
    assert got == want


def test_philox_walk_logic_is_the_pinned_walk_logic(golden_dir):
    """Same walk code (legacy.walks_generic), Philox integer draw  ==  C oracle == walks_py."""
    z = np.load(os.path.join(golden_dir, "walk_small.npz"))
    for i in range(int(z["n_cases"])):
        A = z["A%d" % i]; L, iters, seed = (int(x) for x in z["meta%d" % i])
        rp, col, w = legacy.csr_from_dense(A)
        q = oracle.quantise_weights(w)
        V = A.shape[0]
        for group in (0, 1):
            ids = list(range(iters * V))
            gen = legacy.walks_generic(rp, col, q, L, ids, legacy.PhiloxIntDraw(seed, group, oracle.draw64_py))
            py = oracle.walks_py(rp, col, q, L, seed, group, ids)
            nodes, lens = oracle.walks(rp, col, q, L, seed, group, 0, iters * V)
            c = [list(map(int, r[:n])) for r, n in zip(nodes, lens)]
            assert gen == py == c
            assert (nodes[np.arange(L)[None, :] >= lens[:, None]] == -1).all()


def test_c_oracle_equals_python_on_ex_subset(golden_dir):
    z = np.load(os.path.join(golden_dir, "ex_graph.npz"))
    rp, col = z["rowptr1"], z["col1"]; q = oracle.quantise_weights(z["w1"])
    deg = np.diff(rp)
    ids = [int(x) for x in np.argsort(-deg)[:40]] + [7523 + int(x) for x in np.argsort(-deg)[:20]]
    py = oracle.walks_py(rp, col, q, 80, 99, 1, ids)
    for w, want in zip(ids, py):
        nodes, lens = oracle.walks(rp, col, q, 80, 99, 1, w, w + 1)
        assert list(map(int, nodes[0, :lens[0]])) == want
    assert max(len(p) for p in py) > 20


def test_walk_sharding_invariance(golden_dir):
    z = np.load(os.path.join(golden_dir, "ex_graph.npz"))
    rp, col = z["rowptr0"], z["col0"]; q = oracle.quantise_weights(z["w0"])
    full, fl = oracle.walks(rp, col, q, 80, 5, 0, 0, 3000)
    for r in range(3):
        part, pl = oracle.walks(rp, col, q, 80, 5, 0, r, 3000, 3)
        assert (part == full[r::3]).all() and (pl == fl[r::3]).all()


def test_glue_equals_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "glue_small.npz"))
    V = int(z["V"])
    s0, s1 = unpack(z["f0"], z["l0"]), unpack(z["f1"], z["l1"])
    rows = legacy.integrate_pathSet([s0, s1])
    dense = np.zeros((len(rows), V + 1), dtype=np.int8)
    for i, (lab, p) in enumerate(rows):
        dense[i, list(p)] = 1; dense[i, -1] = lab
    order = np.lexsort(dense.T[::-1])
    assert (dense[order] == z["pathList_sorted"]).all()
    gf = legacy.count_geneFreq(rows)
    want = {int(k[1:]): int(v) for k, v in zip(z["gf_gene"], z["gf_val"])}
    assert gf == want


def test_pcc_adjacency_equals_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "pcc_small.npz"))
    for g in (0, 1):
        grp = z["expr"][z["label"] == g]
        w = legacy.edge_weights(z["src"], z["dst"], grp)
        A = legacy.dense_from_edges(z["src"], z["dst"], w, z["expr"].shape[1])
        assert (A == z["adj%d" % g]).all()
        assert (A > 0).sum() > 10
