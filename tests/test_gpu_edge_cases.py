"""Edge cases of the C ABI on the GPU: empty / degenerate inputs, maximum sizes, and error returns."""
import numpy as np
import pytest

import oracle
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g2v():
    import torch
    assert torch.cuda.is_available()
    import g2vec_b200
    return g2vec_b200


def test_graph_without_edges_and_single_node(g2v):
    for V in (1, 7):
        rp = np.zeros(V + 1, np.int32)
        nodes, lens = g2v.generate_paths_host(rp, np.zeros(0, np.int32), np.zeros(0, np.uint32), 5, 3)
        assert (lens == 1).all() and (nodes[:, 0] == np.arange(3 * V) % V).all() and (nodes[:, 1:] == -1).all()
    # zero walkers in the requested range
    nodes, lens = g2v.generate_paths_host(np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32),
                                          np.array([40000, 40000], np.uint32), 4, 2, walker_begin=4, walker_end=4)
    assert nodes.shape == (0, 4) and lens.shape == (0,)


def test_two_node_cycle_self_loop_and_max_len(g2v):
    # 0 <-> 1 plus a self loop on 0: the self loop is never taken (cur is in the path), walks have 2 nodes
    rp = np.array([0, 2, 3], np.int32); col = np.array([0, 1, 0], np.int32); q = np.array([65536, 40000, 50000], np.uint32)
    want, wl = oracle.walks(rp, col, q, 4096, 1, 0, 0, 4)
    nodes, lens = g2v.generate_paths_host(rp, col, q, 4096, 2, seed=1)            # L = 4096 is the maximum
    assert (nodes == want).all() and (lens == wl).all() and (lens == 2).all()
    # a long path graph 0 -> 1 -> ... -> 2999 walks its whole length when L allows
    V = 3000
    rp = np.minimum(np.arange(V + 1), V - 1).astype(np.int32); col = np.arange(1, V, dtype=np.int32)
    q = np.full(V - 1, 50000, np.uint32)
    nodes, lens = g2v.generate_paths_host(rp, col, q, 4096, 1, walker_end=3)
    assert list(lens) == [3000, 2999, 2998] and (nodes[0, :3000] == np.arange(3000)).all()


def test_argument_errors_are_reported_not_crashed(g2v):
    from g2vec_b200 import _capi
    lib = _capi.load()
    a = np.zeros(8, np.int32)
    assert lib.g2v_walk_host(a.ctypes.data, a.ctypes.data, a.ctypes.data, 3, 0, 0, 0, 0, 0, 3, 1, a.ctypes.data, a.ctypes.data) != 0
    assert b"bad arguments" in lib.g2v_last_error()
    assert lib.g2v_walk_launch(a.ctypes.data, 0, 0, 3, 0, 5000, 0, 0, 0, 3, 1, a.ctypes.data, a.ctypes.data, a.ctypes.data, 0) != 0
    assert b"lenPath" in lib.g2v_last_error()
    assert lib.g2v_cbow_fwdbwd(0, 0, 0, 0, 0, 5, 1.0, 0, 0, 0, 0, 0, 0, 10, 128, 0, 0) != 0
    assert b"null pointer" in lib.g2v_last_error()
    assert lib.g2v_cbow_update(1, 1, 1, 1, 1, 1, 1, 1, 10, 128, 7, 0.1, 0.9, 0.999, 1e-8, 1, 0, 0) != 0
    assert b"unknown optimizer" in lib.g2v_last_error()
    with pytest.raises(RuntimeError):
        g2v.CbowModel(np.array([0, 1]), np.array([0]), np.array([0]), 4, 128, np.zeros((4, 128), np.float32),
                      np.zeros(128, np.float32)).fwdbwd(None, 1, win_begin=-1, n_win=1)
    with pytest.raises(ValueError):
        g2v.WalkGraph(np.array([0, 1]), np.array([0]), weights=np.array([-0.5], np.float32))
    with pytest.raises(ValueError):
        g2v.train_cbow(np.array([0, 1]), np.array([0]), np.array([0]), 4, 128, 0.005, log=None)   # < 2 windows


@pytest.mark.parametrize("algo", ["rows", "rank1"])
def test_empty_windows_and_untouched_rows(g2v, algo):
    """Windows with no gene (never produced by the pipeline, but legal CSR) contribute log 2 to the loss and
    nothing to the gradient; rows no window touches keep their initial value under dense Adam."""
    import torch
    V, D = 50, 128
    rowptr = np.array([0, 0, 3, 3, 5, 5], np.int32); gene = np.array([1, 4, 9, 4, 30], np.int32)
    label = np.array([0, 1, 1, 0, 1], np.uint8)
    W0, Wo0 = helpers.init_weights(V, D, 3)
    m = g2v.CbowModel(rowptr, gene, label, V, D, W0, Wo0, algo=algo)
    m.fwdbwd(None, 5, win_begin=0, n_win=5)
    torch.cuda.synchronize()
    acc = m.acc.cpu()
    o_gih, o_gho, o_loss, o_nc = oracle.cbow_grad(rowptr, gene, label, np.arange(5), 5, W0, Wo0)
    assert abs(m.loss_sum(acc) / 5 - o_loss) < 1e-6 and int(acc[1]) == o_nc
    m.update(); m.update()                       # second step: zero gradient, Adam still moves touched rows
    torch.cuda.synchronize()
    W = m.W_ih.cpu().numpy()
    touched = np.zeros(V, bool); touched[[1, 4, 9, 30]] = True
    assert (W[~touched] == W0[~touched]).all() and (W[touched] != W0[touched]).any()
    m.acc.zero_(); m.evaluate(None, 2, win_begin=0, n_win=0); m.evaluate(None, 2, win_begin=1, n_win=1)
    torch.cuda.synchronize()
    assert int(m.acc.cpu()[2]) in (0, 1)
