"""Legacy-stream restatement of the reference -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/G2Vec.py step 3 (and the small glue around it) with the
reference's own RNG (NumPy legacy MT19937 ``RandomState``) and float arithmetic, so
that it can be pinned BIT-EXACT against outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/make_golden.py, checked by
tests/test_oracle_pin.py).  Two forms:

* ``walks_generic`` -- the walk logic on CSR with a pluggable draw.  With
  ``LegacyDraw`` it reproduces the reference's path sets exactly; with the Philox
  integer draw it is what ``g2v_oracle.c`` and the CUDA kernel compute.  This is the
  link that carries the pin from the reference to the Philox oracle: same walk code,
  only `np.random.choice` (G2Vec.py:341) swapped.
* ``generate_pathSet_dense`` -- a port doing the same dense-row work per step as the
  reference; used as the timed CPU baseline (bench.py ``cpu_baseline`` / ``--impl reference``).
"""
import numpy as np


# ----------------------------------------------------------------- graph construction
def pcc_f32(x, y):
    """G2Vec.py:354-368 compute_PCC, same NumPy calls in the same order (float32 in,
    population std, 0 if either std is 0)."""
    sx, sy = x.std(), y.std()
    if sx > 0. and sy > 0.:
        zx = (x - x.mean()) / sx
        zy = (y - y.mean()) / sy
        return (zx * zy).mean()
    return 0.


def edge_weights(src_idx, dst_idx, expr_group):
    """G2Vec.py:379-390 per edge: weight = abs(PCC) over the group's samples."""
    w = np.zeros(len(src_idx), dtype=np.float32)
    for e, (s, d) in enumerate(zip(src_idx, dst_idx)):
        w[e] = abs(pcc_f32(expr_group[:, s], expr_group[:, d]))
    return w


def dense_from_edges(src_idx, dst_idx, w, V, threshold=0.5):
    """adjMat[src][dest] = weight if weight > 0.5 (G2Vec.py:389-390); later duplicates overwrite."""
    A = np.zeros((V, V), dtype=np.float32)
    for s, d, x in zip(src_idx, dst_idx, w):
        if x > threshold:
            A[s, d] = x
    return A


def csr_from_dense(A):
    """Rows = out-edges, neighbours ascending by dest index (dense row order)."""
    V = A.shape[0]
    r, c = np.nonzero(A)
    rowptr = np.zeros(V + 1, dtype=np.int32)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, c.astype(np.int32), A[r, c].astype(np.float32)


def dense_from_csr(rowptr, col, w):
    V = len(rowptr) - 1
    A = np.zeros((V, V), dtype=np.float32)
    rows = np.repeat(np.arange(V), np.diff(rowptr))
    A[rows, col] = w
    return A


# ------------------------------------------------------------------------------ draws
class LegacyDraw:
    """np.random.choice(n_genes, size=1, p=prob) of G2Vec.py:336-341 on the unvisited
    neighbour list: scatter into a dense float32 row so that prob.sum() (pairwise
    float32) and the f64 cdf are the reference's, bit for bit."""

    def __init__(self, V, seed):
        self.V = V
        self.rng = np.random.RandomState(seed)

    def __call__(self, walker, step, cols, weights):
        prob = np.zeros(self.V, dtype=np.float32)
        prob[cols] = weights
        z = prob.sum()
        prob /= z
        return int(self.rng.choice(self.V, size=1, p=prob)[0])


class PhiloxIntDraw:
    """The oracle/GPU rule: r = floor(x*T/2^64), first inclusive prefix > r."""

    def __init__(self, seed, group, draw64):
        self.seed, self.group, self.draw64 = seed, group, draw64

    def __call__(self, walker, step, cols, q):
        T = int(sum(int(x) for x in q))
        r = (self.draw64(self.seed, (self.group << 40) + int(walker), step) * T) >> 64
        acc = 0
        for c, x in zip(cols, q):
            acc += int(x)
            if acc > r:
                return int(c)
        raise AssertionError("unreachable: r < T")


def walks_generic(rowptr, col, w, L, walker_ids, draw, consume_last=False):
    """The walk of G2Vec.py:328-346 on CSR; returns ordered paths (lists).
    consume_last=True makes the draw at the last iteration (s = L-1) that the reference
    performs and discards (needed only to keep a sequential stream in step)."""
    V = len(rowptr) - 1
    out = []
    for wid in walker_ids:
        cur = int(wid % V)
        path = []
        for s in range(L):
            path.append(cur)
            seen = set(path)
            lo, hi = rowptr[cur], rowptr[cur + 1]
            keep = [j for j in range(lo, hi) if int(col[j]) not in seen]
            if not keep or sum(float(w[j]) for j in keep) <= 0:
                break
            if s == L - 1 and not consume_last:
                break
            cur = draw(wid, s, [int(col[j]) for j in keep], [w[j] for j in keep])
        out.append(path)
    return out


def generate_pathSet_csr(rowptr, col, w, L, iters, seed):
    """Reference semantics through walks_generic + LegacyDraw."""
    V = len(rowptr) - 1
    draw = LegacyDraw(V, seed)
    paths = walks_generic(rowptr, col, w, L, range(iters * V), draw, consume_last=True)
    return {tuple(sorted(p)) for p in paths}


def generate_pathSet_dense(A, L, iters, rng, start_nodes=None, counter=None):
    """Port of G2Vec.py:324-352 doing the reference's dense-row work per step
    (row copy, mask, float32 sum, normalise, legacy choice).  `rng` is a RandomState
    (the reference uses the global one).  `counter` (list of one int) accumulates node
    visits (path.append events) for the steps/s metric."""
    V = A.shape[0]
    starts = range(V) if start_nodes is None else start_nodes
    found = set()
    visits = 0
    for _ in range(iters):
        for src in starts:
            walk, node = [], src
            for _s in range(L):
                walk.append(node)
                p = A[node].copy()
                p[walk] = 0.
                z = p.sum()
                if not z > 0.:
                    break
                p /= z
                node = rng.choice(V, size=1, p=p)[0]
            visits += len(walk)
            found.add(tuple(sorted(walk)))
    if counter is not None:
        counter[0] += visits
    return found


# -------------------------------------------------------------------- path integration
def integrate_pathSet(pathSetList):
    """G2Vec.py:310-322 as (sorted list of (label, path tuple)); the dense multi-hot row of
    the reference is `row[list(path)] = 1; row[-1] = label`."""
    common = pathSetList[0] & pathSetList[1]
    rows = []
    for label, ps in enumerate(pathSetList):
        rows += [(label, p) for p in ps - common]
    return sorted(rows)


def count_geneFreq(rows):
    """G2Vec.py:288-308 on (label, path) rows -> {gene index: 0|1|2}."""
    f = [dict(), dict()]
    for label, p in rows:
        for g in set(p):
            f[label][g] = f[label].get(g, 0) + 1
    out = {}
    for g in set(f[0]) | set(f[1]):
        a, b = f[0].get(g, 0), f[1].get(g, 0)
        out[g] = 0 if a > b else (1 if a < b else 2)
    return out


def windows_from_rows(rows):
    """(label, path) rows -> CSR windows (rowptr int32, gene int32, label uint8)."""
    lens = np.array([len(p) for _, p in rows], dtype=np.int64)
    rowptr = np.zeros(len(rows) + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(lens)
    gene = np.fromiter((g for _, p in rows for g in p), dtype=np.int32, count=int(lens.sum()))
    label = np.array([l for l, _ in rows], dtype=np.uint8)
    return rowptr, gene, label
