"""GPU edge weighting against the reference's construct_adjMat goldens, and the drop-in
command line end to end on the ex_* data (BASELINE configs[0]: lenPath 80, hidden 128, epoch 5)."""
import os

import numpy as np
import pytest


pytestmark = pytest.mark.gpu
TOL_W = 2e-6          # |PCC| agreement: float32 NumPy pairwise sums (reference) vs double accumulation (GPU)


def _compare(rp, col, w, ref_dense=None, ref_csr=None):
    if ref_dense is None:
        ref_dense = None
    got = {(int(s), int(d)): float(x) for s, d, x in
           zip(np.repeat(np.arange(len(rp) - 1), np.diff(rp)), col, w)}
    if ref_csr is not None:
        rrp, rcol, rw = ref_csr
        ref = {(int(s), int(d)): float(x) for s, d, x in
               zip(np.repeat(np.arange(len(rrp) - 1), np.diff(rrp)), rcol, rw)}
    else:
        r, c = np.nonzero(ref_dense)
        ref = {(int(a), int(b)): float(ref_dense[a, b]) for a, b in zip(r, c)}
    both = set(got) & set(ref)
    assert max(abs(got[k] - ref[k]) for k in both) < TOL_W
    for k in set(got) ^ set(ref):                      # only edges sitting on the 0.5 threshold may differ
        assert abs((got.get(k) or ref.get(k)) - 0.5) < 1e-5, k
    return len(both), len(set(got) ^ set(ref))


def test_gpu_adjacency_equals_reference_small(golden_dir):
    from g2vec_b200 import graph
    z = np.load(os.path.join(golden_dir, "pcc_small.npz"))
    for g in (0, 1):
        rp, col, w = graph.group_csr_gpu(z["expr"], z["label"], g, z["src"], z["dst"])
        n, diff = _compare(rp.cpu().numpy(), col.cpu().numpy(), w.cpu().numpy(), ref_dense=z["adj%d" % g])
        assert n > 10 and diff == 0
        rp2, col2, w2 = graph.group_csr(z["expr"], z["label"], g, z["src"], z["dst"])     # host path, same rule
        assert (rp.cpu().numpy() == rp2).all() and (col.cpu().numpy() == col2).all()


def test_gpu_adjacency_equals_reference_ex(golden_dir):
    """All 216 540 ex_* edges, both groups, against the CSR the reference's construct_adjMat produced."""
    from g2vec_b200 import graph
    e = np.load(os.path.join(golden_dir, "ex_expr.npz"))
    gr = np.load(os.path.join(golden_dir, "ex_graph.npz"))
    assert e["expr"].shape == (135, 7523) and len(e["src"]) == 216540
    for g in (0, 1):
        rp, col, w = graph.group_csr_gpu(e["expr"], gr["label"], g, e["src"].astype(np.int32), e["dst"].astype(np.int32))
        n, diff = _compare(rp.cpu().numpy(), col.cpu().numpy(), w.cpu().numpy(),
                           ref_csr=(gr["rowptr%d" % g], gr["col%d" % g], gr["w%d" % g]))
        assert n > 25000 and diff <= 2
        # weights already on the device quantise exactly as on the host
        import g2vec_b200 as g2v
        wg = g2v.WalkGraph(rp, col, weights=w)
        assert (wg.qw.cpu().numpy().view(np.uint32) == graph.quantise_weights(w.cpu().numpy())).all()


def test_command_line_end_to_end(tmp_path, golden_dir, capsys):
    from g2vec_b200 import cli
    from tests import helpers
    ef, cf, nf, genes = helpers.write_ex_tsv(tmp_path)
    outs = []
    for run in range(2):
        prefix = str(tmp_path / ("res%d" % run))
        cli.main([ef, cf, nf, prefix, "-r", "2", "-e", "5", "-n", "20", "--seed", "3"])
        log = capsys.readouterr().out
        assert "    n_samples: 135\n" in log and "    n_genes  : 7523\t(common genes" in log
        assert "    n_edges  : 216540\t(edges with the common genes)" in log          # README.md:26-28
        assert ">>> 4. Compute distributed representations using modified CBOW" in log
        assert "    - Epoch: 000\tACC[val]=" in log and "    Optimization Finish" in log
        vec = open(prefix + "_vectors.txt").read().splitlines()
        assert vec[0] == "GeneSymbol\t" + "\t".join("V%d" % i for i in range(128)) and len(vec) == 7524
        assert vec[1].split("\t")[0] == genes[0] and len(vec[1].split("\t")) == 129
        lg = open(prefix + "_lgroups.txt").read().splitlines()
        assert lg[0] == "GeneSymbol\tLgroup(0:good,1:poor,2:other)" and len(lg) == 7524
        assert {l.split("\t")[1] for l in lg[1:]} == {"0", "1", "2"}
        bm = open(prefix + "_biomarkers.txt").read().splitlines()
        assert bm[0] == "GeneSymbol" and 20 <= len(bm) - 1 <= 40 and bm[1:] == sorted(bm[1:])
        outs.append((np.array([[float(x) for x in l.split("\t")[1:]] for l in vec[1:]]), lg))
    # same seed -> same walks, same split, same init: vectors agree to the atomics' reassociation noise
    assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-4
