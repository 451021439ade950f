"""Pin the CBOW oracle to the reference's own ``compute_genetovec`` (G2Vec.py:217-286).

tests/golden/cbow_small.npz and cbow_ex.npz were produced by tests/golden/make_golden.py running the
UNMODIFIED reference function on oracle/tf1_shim.py (the TF 1.x ops it calls, restated from their published
definitions on torch-CPU; TensorFlow itself cannot be installed here).  The reference's own control flow --
shuffle and 80/20 split (:219-226), one full-batch Adam step then validation then training accuracy with the
updated weights (:264-267), strict-drop early stop (:276), snapshot after the check (:283), return of the
previous snapshot (:286) -- therefore fixes what ``oracle.cbow_train`` must reproduce: the same stop step, the
same per-step accuracies and the same vectors up to float32 reassociation (BLAS matmul vs the oracle's
sequential sums).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import ref_import
from tests import helpers

# max |W_oracle - W_reference| / max |W_reference|; measured 6.0e-7 (small) and 2.1e-6 (ex) -- float32
# summation order of the dense matmuls (X^T.dH over up to 8.6k rows) against the oracle's sequential loops
VEC_TOL = 5e-6


@pytest.mark.parametrize("name", ["cbow_small.npz", "cbow_ex.npz"])
def test_oracle_reproduces_the_reference_run(name):
    g = helpers.cbow_golden(name)
    assert g["stop_step"] > 5, "the golden run must train for a while and then stop early"
    got, hist, stop, _ = oracle.cbow_train(g["rowptr"], g["gene"], g["label"], g["tr"], g["va"], g["W0"], g["Wo0"],
                                           g["lr"], max_steps=500)
    assert stop == g["stop_step"]                                  # same early-stop step (:276-279)
    assert len(hist) == len(g["acc_val"])
    n_va, n_tr = len(g["va"]), len(g["tr"])
    flips_val = [round(abs(h[1] - a) * n_va) for h, a in zip(hist, g["acc_val"])]
    flips_tr = [round(abs(h[2] - a) * n_tr) for h, a in zip(hist, g["acc_tr"])]
    assert max(flips_val) == 0 and max(flips_tr) == 0, (flips_val, flips_tr)      # every step's accuracies, exactly
    err = float(np.abs(got - g["W_ref"]).max() / np.abs(g["W_ref"]).max())
    assert err < VEC_TOL, err
    # rows no training window touches keep their initial value in the reference (dense Adam with g = m = v = 0)
    untouched = np.ones(g["V"], bool)
    untouched[np.unique(g["gene"])] = False
    assert (got[untouched] == g["W0"][untouched]).all() and (g["W_ref"][untouched] == g["W0"][untouched]).all()


def test_returned_matrix_is_the_snapshot_before_the_drop():
    """:283/:286 -- the result is W_ih after the last step whose validation accuracy did not drop, not the
    weights of the step that triggered the stop."""
    g = helpers.cbow_golden("cbow_small.npz")
    s = g["stop_step"]
    upto, _, _, _ = oracle.cbow_train(g["rowptr"], g["gene"], g["label"], g["tr"], g["va"], g["W0"], g["Wo0"], g["lr"],
                                      max_steps=s, early_stop=False)          # weights after steps 0..s-1
    one_more, _, _, _ = oracle.cbow_train(g["rowptr"], g["gene"], g["label"], g["tr"], g["va"], g["W0"], g["Wo0"],
                                          g["lr"], max_steps=s + 1, early_stop=False)
    e_prev = np.abs(upto - g["W_ref"]).max()
    e_next = np.abs(one_more - g["W_ref"]).max()
    assert e_prev < VEC_TOL * np.abs(g["W_ref"]).max() < e_next


def test_log_lines_of_the_reference_run():
    g = helpers.cbow_golden("cbow_ex.npz")
    lines = g["log"].splitlines()
    assert lines[0] == "     Start training the modified CBOW with early stopping"
    assert lines[1].startswith("    - Epoch: 000\tACC[val]=%.4f\tACC[tr]=%.4f" % (g["acc_val"][0], g["acc_tr"][0]))
    s = g["stop_step"]
    assert lines[-2].startswith("    - Epoch(stop): %03d\tACC[val]=%.4f\tACC[tr]=%.4f"
                                % (s - 1, g["acc_val"][s - 1], g["acc_tr"][s - 1]))
    assert lines[-1] == "    Optimization Finish"


@pytest.mark.skipif(not ref_import.available(), reason="reference script not present (neither /root/reference nor oracle/_ref)")
def test_golden_is_what_the_reference_produces_now():
    """Re-run the unmodified reference function under the shim and compare with the committed fixture."""
    import contextlib
    import io
    from oracle import tf1_shim
    ref = ref_import.load()
    g = helpers.cbow_golden("cbow_small.npz")
    N = len(g["rowptr"]) - 1
    P = np.zeros((N, g["V"] + 1), dtype=np.int32)
    for n in range(N):
        P[n, g["gene"][g["rowptr"][n]:g["rowptr"][n + 1]]] = 1
    P[:, -1] = g["label"]
    tf1_shim.reset(); tf1_shim.seed_initialisers(g["seed"]); np.random.seed(g["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        W = ref.compute_genetovec(P, g["V"], g["D"], g["lr"])
    assert np.abs(W - g["W_ref"]).max() < 1e-6 * np.abs(g["W_ref"]).max()     # BLAS thread count may differ
    accs = [float(e[1]) for e in tf1_shim.trace() if e[0] == "eval"]
    assert np.allclose(accs[0::2], g["acc_val"], atol=1e-7) and len(accs) // 2 - 1 == g["stop_step"]
    W0, Wo0 = tf1_shim.initial_values()
    assert (W0 == g["W0"]).all() and (Wo0.reshape(-1) == g["Wo0"]).all()


def test_shim_adam_is_tf1_apply_adam():
    """The shim's optimizer against the closed form of training_ops.cc ApplyAdam for two steps on a scalar
    problem: cost = mean((w*x - 0)^2)-like graph is not available, so use the BCE graph with one window."""
    from oracle import tf1_shim as tf
    tf.reset()
    X = tf.placeholder(tf.float32, [None, 2]); Y = tf.placeholder(tf.float32, [None, 1])
    W = tf.Variable(np.array([[0.3], [-0.2]], dtype=np.float32))
    cost = tf.reduce_mean(tf.nn.sigmoid_cross_entropy_with_logits(logits=tf.matmul(X, W), labels=Y))
    op = tf.train.AdamOptimizer(0.005).minimize(cost)
    x = np.array([[1.0, 2.0]], dtype=np.float32); y = np.array([[1.0]], dtype=np.float32)
    f32 = np.float32
    w = np.array([0.3, -0.2], dtype=np.float32); m = np.zeros(2, f32); v = np.zeros(2, f32)
    b1p, b2p = f32(0.9), f32(0.999)
    with tf.Session() as sess:
        tf.global_variables_initializer().run()
        for _ in range(2):
            sess.run(op, feed_dict={X: x, Y: y})
            o = f32(w[0] * x[0, 0] + w[1] * x[0, 1])
            g = (f32(1) / (f32(1) + np.exp(-o, dtype=f32)) - f32(1)) * x[0]
            alpha = f32(0.005) * np.sqrt(f32(1) - b2p) / (f32(1) - b1p)
            m = m + (g - m) * (f32(1) - f32(0.9)); v = v + (g * g - v) * (f32(1) - f32(0.999))
            w = w - (m * alpha) / (np.sqrt(v) + f32(1e-8))
            b1p, b2p = f32(b1p * f32(0.9)), f32(b2p * f32(0.999))
            assert np.abs(sess.run(W).ravel() - w).max() < 2e-7
