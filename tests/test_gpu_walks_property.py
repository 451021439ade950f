"""Property-based parity (hypothesis): random small directed graphs, weights, lenPath, seeds, walker ranges --
the GPU sampler must equal the oracle bit for bit on every example."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

pytestmark = pytest.mark.gpu


@st.composite
def graphs(draw):
    V = draw(st.integers(1, 48))
    rs = np.random.RandomState(draw(st.integers(0, 2**31 - 1)))
    dens = draw(st.sampled_from([0.0, 0.03, 0.15, 0.5, 1.0]))
    A = (rs.rand(V, V) < dens)
    if not draw(st.booleans()):
        np.fill_diagonal(A, False)                         # self loops allowed in half of the examples
    r, c = np.nonzero(A)
    rowptr = np.zeros(V + 1, np.int32); np.add.at(rowptr, r + 1, 1); rowptr = np.cumsum(rowptr).astype(np.int32)
    kind = draw(st.sampled_from(["pcc", "tiny", "wide"]))
    if kind == "pcc":
        q = np.rint((0.5 + 0.5 * rs.rand(len(c))) * 65536).astype(np.uint32) + 1
    elif kind == "tiny":
        q = rs.randint(1, 4, size=len(c)).astype(np.uint32)
    else:
        q = rs.randint(1, 1 << 24, size=len(c)).astype(np.uint32)
    L = draw(st.sampled_from([1, 2, 5, 33, 80, 200]))
    reps = draw(st.integers(1, 3))
    seed = draw(st.integers(0, 2**64 - 1))
    group = draw(st.integers(0, 3))
    begin = draw(st.integers(0, V)); stride = draw(st.integers(1, 4))
    return rowptr, c.astype(np.int32), q, L, reps, seed, group, begin, stride


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(graphs())
def test_gpu_walks_equal_oracle_on_random_graphs(g):
    import torch
    import g2vec_b200 as g2v
    rowptr, col, q, L, reps, seed, group, begin, stride = g
    V = len(rowptr) - 1
    want, wl = oracle.walks(rowptr, col, q, L, seed, group, begin, reps * V, stride)
    wg = g2v.WalkGraph(rowptr, col, qw=q)
    nodes, lens = g2v.generate_paths(wg, L, reps, seed=seed, group=group, walker_begin=begin, walker_stride=stride)
    torch.cuda.synchronize()
    assert (lens.cpu().numpy() == wl).all() and (nodes.cpu().numpy() == want).all()
