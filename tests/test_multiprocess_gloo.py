"""world_size-2 gloo (CPU) checks of the N>1 host logic: window/walker sharding is a partition, the
per-step exchange (sum all-reduce of the dense gradient + accuracy counters) reproduces the single-process
result, walker shards reassemble the full walk set.  Compute on each rank is the oracle (no GPU here);
the product's own sharding functions are what is under test."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from g2vec_b200 import cbow, walks
    from tests import helpers
    V, N, D = 120, 600, 16
    rowptr, gene, label = helpers.random_windows(N, V, 1, 25, seed=8)
    W0, Wo0 = helpers.init_weights(V, D, 1)
    tr, va = cbow.split_indices(N, 5)
    lens = np.diff(rowptr).astype(np.int64)
    tr_loc = cbow.shard_by_nnz(tr, lens, world, rank)
    va_loc = cbow.shard_by_nnz(va, lens, world, rank)
    g_ih, g_ho, _, _ = oracle.cbow_grad(rowptr, gene, label, tr_loc, len(tr), W0, Wo0)   # 1/N of the GLOBAL N
    g = torch.from_numpy(g_ih); go = torch.from_numpy(g_ho)
    dist.all_reduce(g); dist.all_reduce(go)
    cnt = torch.tensor([oracle.cbow_eval(rowptr, gene, label, va_loc, W0, Wo0),
                        oracle.cbow_eval(rowptr, gene, label, tr_loc, W0, Wo0)], dtype=torch.int64)
    dist.all_reduce(cnt)
    # walker shards
    rp, col, w = helpers.random_graph(300, 5, seed=2)
    qw = oracle.quantise_weights(w)
    n_mine = walks.num_walkers(300, 2, rank, None, world)
    nodes, wl = oracle.walks(rp, col, qw, 20, 3, 1, rank, 600, world)
    assert nodes.shape[0] == n_mine
    gathered = [None] * world
    dist.all_gather_object(gathered, (nodes, wl, sorted(map(int, tr_loc))))
    # the product's own gather (what the command line does under torchrun): every rank ends up with the 1-GPU arrays
    from g2vec_b200 import paths
    full_n, full_l = paths.gather_walker_shards(dist, world, 600, torch.from_numpy(nodes), torch.from_numpy(wl))
    want_n, want_l = oracle.walks(rp, col, qw, 20, 3, 1, 0, 600)
    assert (full_n.numpy() == want_n).all() and (full_l.numpy() == want_l).all()
    # mini-batches: the shuffled list is dealt in order, so batch b is the same set of windows for any world size
    mb = cbow.shard_by_nnz(tr, lens, world, rank, keep_order=True)
    assert (mb == tr[rank::world]).all()
    if rank == 0:
        q.put((g.numpy(), go.numpy(), cnt.numpy(), gathered))
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    g, go, cnt, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    import oracle
    from g2vec_b200 import cbow
    from tests import helpers
    V, N, D = 120, 600, 16
    rowptr, gene, label = helpers.random_windows(N, V, 1, 25, seed=8)
    W0, Wo0 = helpers.init_weights(V, D, 1)
    tr, va = cbow.split_indices(N, 5)
    g1, go1, _, _ = oracle.cbow_grad(rowptr, gene, label, tr, len(tr), W0, Wo0)
    assert np.abs(g - g1).max() <= 1e-6 * np.abs(g1).max() + 1e-12
    assert np.abs(go - go1).max() <= 1e-5 * np.abs(go1).max()
    assert cnt[0] == oracle.cbow_eval(rowptr, gene, label, va, W0, Wo0)
    assert cnt[1] == oracle.cbow_eval(rowptr, gene, label, tr, W0, Wo0)
    # shards partition the training windows
    assert sorted(gathered[0][2] + gathered[1][2]) == sorted(map(int, tr))
    # walker shards interleave back into the full result
    rp, col, w = helpers.random_graph(300, 5, seed=2)
    full, fl = oracle.walks(rp, col, oracle.quantise_weights(w), 20, 3, 1, 0, 600)
    for r in range(world):
        assert (gathered[r][0] == full[r::world]).all() and (gathered[r][1] == fl[r::world]).all()
