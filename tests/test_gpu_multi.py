"""N-GPU == 1-GPU (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`).
Walks: bit-exact (counter-based RNG).  CBOW: vectors within the fp32-reassociation tolerance."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle
from tests import helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpus_equal_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "mgpu.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    z = np.load(out)
    rp, col, w = helpers.ex_graph(1)
    want, wl = oracle.walks(rp, col, oracle.quantise_weights(w), 80, 7, 1, 0, 2 * 7523)
    assert (z["nodes"] == want).all() and (z["lens"] == wl).all()
    (rowptr, gene, label), _ = helpers.ex_windows(reps=2)
    N = len(rowptr) - 1
    tr, va = oracle.split_indices(N, 0)
    W0, Wo0 = helpers.init_weights(7523, 128, 0)
    ref, hist, _, _ = oracle.cbow_train(rowptr, gene, label, tr, va, W0, Wo0, 0.005, max_steps=11, early_stop=False)
    assert np.abs(z["W"] - ref).max() < 1e-4 * np.abs(ref).max()
    print("multi-GPU chunks ran as CUDA graphs with the exchange inside:", bool(z["graph"]), "| exchanges:", list(z["exchange"]))
    assert str(z["exchange"][2]).startswith("nccl")
    for k in ("W_p2p", "W_nccl", "W_rank1"):                       # every exchange path / algorithm gives the same vectors
        assert np.abs(z[k] - ref).max() < 1e-4 * np.abs(ref).max(), k
    assert np.abs(z["W"] - z["W_nccl"]).max() < 2e-5 * np.abs(ref).max()
    assert np.abs(z["hist"] - z["hist_nccl"]).max() <= 2.0 / len(va) + 1e-7
    assert len(z["hist"]) == 11
    for (s_, av, at), row in zip(hist, z["hist"]):
        assert abs(av - row[1]) <= 2.0 / len(va) + 1e-7 and abs(at - row[2]) <= 2.0 / len(tr) + 1e-7


def test_command_line_under_torchrun_equals_one_gpu(tmp_path, capsys):
    """`torchrun --nproc-per-node N G2Vec.py ...`: walkers and windows sharded, rank 0 writes the files."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from g2vec_b200 import cli
    ef, cf, nf, genes = helpers.write_ex_tsv(tmp_path)
    opts = ["-r", "2", "-e", "5", "-n", "20", "--seed", "3"]
    one = str(tmp_path / "one")
    cli.main([ef, cf, nf, one] + opts)
    capsys.readouterr()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    two = str(tmp_path / "two")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "G2Vec.py"), ef, cf, nf, two] + opts
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count(">>> 7. Save results") == 1                       # only rank 0 talks
    n1 = [l for l in open(one + "_vectors.txt")][1:]; n2 = [l for l in open(two + "_vectors.txt")][1:]
    a = np.array([[float(x) for x in l.split("\t")[1:]] for l in n1])
    b = np.array([[float(x) for x in l.split("\t")[1:]] for l in n2])
    assert a.shape == b.shape == (7523, 128) and np.abs(a - b).max() < 1e-4 * np.abs(a).max() + 2e-6
    l1 = [l.split("\t")[1] for l in open(one + "_lgroups.txt")][1:]
    l2 = [l.split("\t")[1] for l in open(two + "_lgroups.txt")][1:]
    assert np.mean([x == y for x, y in zip(l1, l2)]) > 0.995                # KMeans on vectors equal to ~1e-6
