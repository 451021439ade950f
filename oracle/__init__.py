"""CPU ORACLE for the G2Vec hot paths -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``g2vec_b200`` never does.

Contents
--------
* ``g2v_oracle.c`` (via ctypes): scalar-C restatement of the walk sampler
  (/root/reference/G2Vec.py:324-352) with the Philox integer draw, and of one CBOW
  optimizer step (/root/reference/G2Vec.py:239-246) on CSR windows.
* ``philox4x32_10`` / ``walks_py``: a second, pure-Python restatement used to
  cross-check the C on small cases.
* ``cbow_train``: the epoch loop / early stop of G2Vec.py:259-286.
* ``oracle.legacy``: the reference's legacy-MT-stream algorithm (dense adjacency,
  ``np.random.choice``), pinned bit-exact to the reference's own functions.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libg2v_oracle.so")
_lib = None


def build(force=False):
    """Compile g2v_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "g2v_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f32p = ctypes.POINTER(ctypes.c_float)
        L.g2v_oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.g2v_oracle_philox4x32_10.restype = None
        L.g2v_oracle_draw64.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.g2v_oracle_draw64.restype = ctypes.c_uint64
        L.g2v_oracle_walks.argtypes = [i32p, i32p, u32p, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_int64, i32p, i32p]
        L.g2v_oracle_walks.restype = ctypes.c_int
        L.g2v_oracle_cbow_grad.argtypes = [i32p, i32p, u8p, i64p, ctypes.c_int64, ctypes.c_int64,
                                           f32p, f32p, ctypes.c_int32, ctypes.c_int32,
                                           f32p, f32p, ctypes.POINTER(ctypes.c_double), i64p]
        L.g2v_oracle_cbow_grad.restype = ctypes.c_int
        L.g2v_oracle_adam.argtypes = [f32p, f32p, f32p, f32p, ctypes.c_int64, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int32]
        L.g2v_oracle_adam.restype = None
        L.g2v_oracle_sgd.argtypes = [f32p, f32p, ctypes.c_int64, ctypes.c_float]
        L.g2v_oracle_sgd.restype = None
        L.g2v_oracle_cbow_eval.argtypes = [i32p, i32p, u8p, i64p, ctypes.c_int64,
                                           f32p, f32p, ctypes.c_int32, f32p]
        L.g2v_oracle_cbow_eval.restype = ctypes.c_int64
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------- Philox
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Pure-Python Philox4x32-10 (curand_philox4x32_x.h constants). ctr: 4 words, key: 2."""
    c = [int(x) & _MASK for x in ctr]
    k = [int(x) & _MASK for x in key]
    for r in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & _MASK, p1 & _MASK,
             ((p0 >> 32) ^ c[3] ^ k[1]) & _MASK, p0 & _MASK]
        if r < 9:
            k = [(k[0] + _W0) & _MASK, (k[1] + _W1) & _MASK]
    return c


def philox4x32_10_c(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.zeros(4, dtype=np.uint32)
    lib().g2v_oracle_philox4x32_10(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(out, ctypes.c_uint32))
    return [int(x) for x in out]


def draw64_py(seed, subseq, s):
    w = philox4x32_10([s >> 1, 0, subseq & _MASK, (subseq >> 32) & _MASK],
                      [seed & _MASK, (seed >> 32) & _MASK])
    return (w[2 * (s & 1) + 1] << 32) | w[2 * (s & 1)]


def draw64(seed, subseq, s):
    return int(lib().g2v_oracle_draw64(seed, subseq, s))


# ------------------------------------------------------------------------------ walks
def quantise_weights(w):
    """q = rint(w * 2^16) as uint32, at least 1 for a positive weight (oracle's own copy
    of the rule in g2vec_b200.graph.quantise_weights; the test compares the two)."""
    w = np.asarray(w, dtype=np.float32)
    q = np.rint(w.astype(np.float64) * 65536.0)
    q = np.where((w > 0) & (q < 1), 1, q)
    return q.astype(np.uint32)


def walks(rowptr, col, qw, L, seed, group, walker_begin, walker_end, walker_stride=1):
    """C oracle. Returns (nodes int32 [n, L] padded with -1, lengths int32 [n])."""
    rowptr = _c(rowptr, np.int32); col = _c(col, np.int32); qw = _c(qw, np.uint32)
    V = rowptr.shape[0] - 1
    n = max(0, (walker_end - walker_begin + walker_stride - 1) // walker_stride)
    nodes = np.empty((n, L), dtype=np.int32)
    lens = np.empty(n, dtype=np.int32)
    rc = lib().g2v_oracle_walks(_p(rowptr, ctypes.c_int32), _p(col, ctypes.c_int32),
                                _p(qw, ctypes.c_uint32), V, L, seed, group,
                                walker_begin, walker_end, walker_stride,
                                _p(nodes, ctypes.c_int32), _p(lens, ctypes.c_int32))
    if rc != 0:
        raise ValueError("g2v_oracle_walks: bad arguments")
    return nodes, lens


def walks_py(rowptr, col, qw, L, seed, group, walker_ids):
    """Pure-Python restatement of G2Vec.py:328-346 with the Philox integer draw (small cases)."""
    V = len(rowptr) - 1
    out = []
    for w in walker_ids:
        cur = int(w % V)
        subseq = (group << 40) + int(w)
        path = []
        for s in range(L):
            path.append(cur)                                        # :332
            if s == L - 1:
                break
            seen = set(path)
            nb = [(int(col[j]), int(qw[j])) for j in range(rowptr[cur], rowptr[cur + 1])
                  if int(col[j]) not in seen]                       # :334-336
            T = sum(q for _, q in nb)                               # :338
            if T == 0:
                break                                               # :342-344
            r = (draw64_py(seed, subseq, s) * T) >> 64
            acc = 0
            for c, q in nb:                                         # :341 inverse CDF
                acc += q
                if acc > r:
                    cur = c
                    break
        out.append(path)
    return out


def path_set(nodes, lens):
    """tuple(sorted(path)) into a set -- G2Vec.py:345,351."""
    return {tuple(sorted(int(x) for x in row[:n])) for row, n in zip(nodes, lens)}


# ------------------------------------------------------------------------------- CBOW
def cbow_grad(rowptr, gene, label, win, n_total, W_ih, W_ho):
    rowptr = _c(rowptr, np.int32); gene = _c(gene, np.int32); label = _c(label, np.uint8)
    win = _c(win, np.int64); W_ih = _c(W_ih, np.float32); W_ho = _c(W_ho, np.float32).reshape(-1)
    V, D = W_ih.shape
    g_ih = np.empty((V, D), dtype=np.float32); g_ho = np.empty(D, dtype=np.float32)
    loss = ctypes.c_double(0); nc = ctypes.c_int64(0)
    rc = lib().g2v_oracle_cbow_grad(_p(rowptr, ctypes.c_int32), _p(gene, ctypes.c_int32),
                                    _p(label, ctypes.c_uint8), _p(win, ctypes.c_int64), len(win),
                                    n_total, _p(W_ih, ctypes.c_float), _p(W_ho, ctypes.c_float), V, D,
                                    _p(g_ih, ctypes.c_float), _p(g_ho, ctypes.c_float),
                                    ctypes.byref(loss), ctypes.byref(nc))
    assert rc == 0
    return g_ih, g_ho, loss.value, nc.value


def adam_(var, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    for a in (var, m, v, g):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().g2v_oracle_adam(_p(var, ctypes.c_float), _p(m, ctypes.c_float), _p(v, ctypes.c_float),
                          _p(g, ctypes.c_float), var.size, lr, beta1, beta2, eps, t)


def sgd_(var, g, lr):
    lib().g2v_oracle_sgd(_p(var, ctypes.c_float), _p(g, ctypes.c_float), var.size, lr)


def cbow_eval(rowptr, gene, label, win, W_ih, W_ho, return_logits=False):
    rowptr = _c(rowptr, np.int32); gene = _c(gene, np.int32); label = _c(label, np.uint8)
    win = _c(win, np.int64); W_ih = _c(W_ih, np.float32); W_ho = _c(W_ho, np.float32).reshape(-1)
    o = np.empty(len(win), dtype=np.float32)
    nc = lib().g2v_oracle_cbow_eval(_p(rowptr, ctypes.c_int32), _p(gene, ctypes.c_int32),
                                    _p(label, ctypes.c_uint8), _p(win, ctypes.c_int64), len(win),
                                    _p(W_ih, ctypes.c_float), _p(W_ho, ctypes.c_float),
                                    W_ih.shape[1], _p(o, ctypes.c_float))
    return (int(nc), o) if return_logits else int(nc)


def split_indices(n, seed):
    """np.random.shuffle(pathList) + 80/20 split (G2Vec.py:219-222) on an index vector.
    Legacy RandomState so that, given the same MT state, the permutation is the one the
    reference's in-place row shuffle produces."""
    perm = np.arange(n, dtype=np.int64)
    np.random.RandomState(seed).shuffle(perm)
    pivot = int(n * 0.8)
    return perm[:pivot], perm[pivot:]


def cbow_train(rowptr, gene, label, tr, va, W_ih0, W_ho0, lr, max_steps=500, optimizer="adam",
               early_stop=True, log=None):
    """Epoch loop of G2Vec.py:259-286. Returns (W_ih result, history list of
    (step, acc_val, acc_tr), stop_step or None, final W_ho)."""
    W_ih = np.array(W_ih0, dtype=np.float32, copy=True)
    W_ho = np.array(W_ho0, dtype=np.float32, copy=True).reshape(-1)
    m_ih = np.zeros_like(W_ih); v_ih = np.zeros_like(W_ih)
    m_ho = np.zeros_like(W_ho); v_ho = np.zeros_like(W_ho)
    hist = []
    before_val, before_tr = -1.0, None
    result = W_ih.copy()
    stop = None
    for step in range(max_steps):                                               # :262
        g_ih, g_ho, _, _ = cbow_grad(rowptr, gene, label, tr, len(tr), W_ih, W_ho)
        if optimizer == "adam":                                                  # :246,264
            adam_(W_ih, m_ih, v_ih, g_ih, lr, step + 1)
            adam_(W_ho, m_ho, v_ho, g_ho, lr, step + 1)
        else:
            sgd_(W_ih, g_ih, lr); sgd_(W_ho, g_ho, lr)
        # float32 mean of a 0/1 vector, as tf.reduce_mean(tf.cast(correction, float32)) :251
        acc_val = np.float32(cbow_eval(rowptr, gene, label, va, W_ih, W_ho)) / np.float32(max(len(va), 1))
        acc_tr = np.float32(cbow_eval(rowptr, gene, label, tr, W_ih, W_ho)) / np.float32(max(len(tr), 1))
        hist.append((step, float(acc_val), float(acc_tr)))
        if log is not None and step % 5 == 0:
            log("    - Epoch: %03d\tACC[val]=%.4f\tACC[tr]=%.4f" % (step, acc_val, acc_tr))
        if early_stop and acc_val < before_val:                                  # :276
            stop = step
            break
        before_val, before_tr = acc_val, acc_tr                                  # :280-281
        result = W_ih.copy()                                                     # :283
    return result, hist, stop, W_ho
