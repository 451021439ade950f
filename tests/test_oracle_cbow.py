"""The CBOW oracle has no TF to be pinned against (TensorFlow 1.x is absent: 'parity unpinned' for the TF
kernels themselves).  What can be checked on CPU: (a) the closed-form gradients of g2v_oracle.c equal
torch autograd of the reference's graph (dense X, two matmuls, sigmoid BCE mean); (b) the sparse oracle
equals the dense port used as the CPU baseline; (c) TF1-form Adam differs from torch.optim.Adam exactly
where the survey says (epsilon placement); (d) the README accuracy envelope on ex_* windows."""
import numpy as np
import torch

import oracle
from oracle import dense_cbow
from tests import helpers


def test_gradients_equal_torch_autograd_of_the_reference_graph():
    V, N, D = 60, 200, 16
    rowptr, gene, label = helpers.random_windows(N, V, 1, 12, seed=1)
    W0, Wo0 = helpers.init_weights(V, D, 2)
    idx = np.arange(N)
    X, y = dense_cbow.densify(rowptr, gene, label, idx, V)
    W = torch.tensor(W0, dtype=torch.float64, requires_grad=True)
    Wo = torch.tensor(Wo0.reshape(-1, 1), dtype=torch.float64, requires_grad=True)
    O = (X.double() @ W) @ Wo
    cost = torch.nn.functional.binary_cross_entropy_with_logits(O, y.double(), reduction="mean")
    cost.backward()
    g_ih, g_ho, loss, nc = oracle.cbow_grad(rowptr, gene, label, idx, N, W0, Wo0)
    assert np.abs(g_ih - W.grad.numpy()).max() < 1e-7
    assert np.abs(g_ho - Wo.grad.numpy().ravel()).max() < 1e-6
    assert abs(loss - float(cost.detach())) < 1e-6
    assert nc == int(((O.detach() > 0).double() == y.double()).sum())
    # rank-1 structure the survey notes: dW_ih = (X^T dO) (x) W_ho
    dO = (torch.sigmoid(O.detach()) - y.double()) / N
    c = (X.double().t() @ dO).numpy().ravel()
    assert np.abs(g_ih - np.outer(c, Wo0)).max() < 1e-7


def test_sparse_oracle_equals_dense_port_over_steps():
    V, N, D = 80, 400, 32
    rowptr, gene, label = helpers.random_windows(N, V, 1, 20, seed=3)
    W0, Wo0 = helpers.init_weights(V, D, 4)
    tr, va = oracle.split_indices(N, 1)
    want, hist, stop, _ = oracle.cbow_train(rowptr, gene, label, tr, va, W0, Wo0, 0.005, max_steps=8, early_stop=False)
    Xtr, ytr = dense_cbow.densify(rowptr, gene, label, tr, V)
    Xva, yva = dense_cbow.densify(rowptr, gene, label, va, V)
    m = dense_cbow.DenseCbow(W0, Wo0, 0.005)
    for s in range(8):
        av, at = m.epoch(Xtr, ytr, Xva, yva)
        assert abs(av - hist[s][1]) <= 1.0 / len(va) + 1e-6 and abs(at - hist[s][2]) <= 1.0 / len(tr) + 1e-6
    assert np.abs(m.W.numpy() - want).max() < 1e-4 * np.abs(want).max()


def test_tf1_adam_is_not_torch_adam():
    rs = np.random.RandomState(0)
    w0 = rs.randn(1000).astype(np.float32); g = (rs.randn(1000) * 1e-7).astype(np.float32)
    w = w0.copy(); m = np.zeros_like(w); v = np.zeros_like(w)
    oracle.adam_(w, m, v, g, 0.005, 1)
    # TF1: step = lr * sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + eps)
    a = np.float32(0.005) * np.sqrt(np.float32(1) - np.float32(0.999)) / (np.float32(1) - np.float32(0.9))
    want = w0 - (np.float32(0.1) * g * a) / (np.sqrt(np.float32(0.001) * g * g) + np.float32(1e-8))
    assert np.abs(w - want).max() < 1e-6
    p = torch.tensor(w0.copy(), requires_grad=True)
    opt = torch.optim.Adam([p], lr=0.005, eps=1e-8); p.grad = torch.tensor(g); opt.step()
    assert np.abs(p.detach().numpy() - w).max() > 1e-4           # eps placement matters at |g| ~ 1e-7


def test_readme_accuracy_envelope_on_ex_windows():
    """README.md:35-41 (unseeded run): ACC[val] 0.63 -> 0.80 -> 0.84 within ~10 steps.  Oracle on oracle-made
    ex_* windows (2 repetitions to keep the test short) must land in the same envelope."""
    (rowptr, gene, label), _ = helpers.ex_windows(reps=2)
    N = len(rowptr) - 1
    tr, va = oracle.split_indices(N, 0)
    W0, Wo0 = helpers.init_weights(7523, 128, 0)
    _, hist, _, _ = oracle.cbow_train(rowptr, gene, label, tr, va, W0, Wo0, 0.005, max_steps=11, early_stop=False)
    assert 0.45 < hist[0][1] < 0.80
    assert hist[10][1] > 0.72 and hist[10][2] >= hist[10][1] - 0.02
