"""Host side of the training loop (g2vec_b200.cbow._LoopLog: what the reference does after its three session runs,
G2Vec.py:268-283) fed with the per-step counters of the reference's own run: it must print the reference's log.  CPU."""
import numpy as np

from g2vec_b200 import cbow
from tests import helpers


def test_loop_log_reproduces_the_reference_log_from_the_counters():
    g = helpers.cbow_golden("cbow_ex.npz")
    n_tr, n_va = len(g["tr"]), len(g["va"])
    lines = []
    log = cbow._LoopLog(n_tr, n_va, lines.append)
    val = np.rint(g["acc_val"].astype(np.float64) * n_va).astype(np.int64)
    trc = np.rint(g["acc_tr"].astype(np.float64) * n_tr).astype(np.int64)
    stop = g["stop_step"]
    over = False
    for s in range(len(val)):
        shown = s % 5 == 0
        # acc = [loss bits, pre-update train correct (= ACC[tr] of step s-1), validation correct, train correct if evaluated]
        acc = [0, trc[s - 1] if s else 0, val[s], trc[s] if shown else 0]
        over = log.step(s, acc, shown, s == stop)
        assert over == (s == stop)
    strip = lambda l: l.split(" (")[0]
    ref = [strip(l) for l in g["log"].splitlines()[1:-1]]           # without the Start / Finish banners
    assert [strip(l) for l in lines] == ref
    # history: every step's ACC[tr] is known one step later at the latest (pipelined training accuracy)
    assert [h[0] for h in log.hist] == list(range(stop + 1))
    assert all(abs(h[1] - float(a)) < 1e-7 for h, a in zip(log.hist, g["acc_val"]))
    assert all(h[2] is not None and abs(h[2] - float(a)) < 1e-7 for h, a in zip(log.hist[:-1], g["acc_tr"]))


def test_split_and_init_are_the_reference_ones():
    """The product's split / init equal what the reference drew in the golden run (np.random.seed(seed) shuffle of the
    rows, PCG64(seed) truncated normals): the GPU test compares against the returned matrix with nothing in between."""
    g = helpers.cbow_golden("cbow_small.npz")
    tr, va = cbow.split_indices(len(g["rowptr"]) - 1, g["seed"])
    assert (tr == g["tr"]).all() and (va == g["va"]).all()
    W0, Wo0 = cbow.init_weights(g["V"], g["D"], g["seed"])
    assert (W0 == g["W0"]).all() and (Wo0 == g["Wo0"]).all()
