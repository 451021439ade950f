"""torchrun worker for tests/test_gpu_multi.py: N-GPU run of both hot paths, rank 0 saves the results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import g2vec_b200 as g2v
    from tests import helpers
    # walks: interleaved shards, gathered on rank 0
    rp, col, w = helpers.ex_graph(1)
    g = g2v.WalkGraph(rp, col, weights=w)
    nodes, lens = g2v.generate_paths(g, 80, 2, seed=7, group=1, walker_begin=rank, walker_stride=world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (nodes.cpu().numpy(), lens.cpu().numpy()))
    # CBOW: 5 steps on the oracle-made ex_* windows
    (rowptr, gene, label), _ = helpers.ex_windows(reps=2)
    W0, Wo0 = helpers.init_weights(7523, 128, 0)
    # 11 steps: step 0 eagerly, then two 5-step chunks -- on N GPUs each chunk is ONE CUDA graph that contains the
    # gradient exchange (g2vec_b200.cbow.DeviceLoop).  Three exchanges: the fused reduce-scatter + Adam + all-gather
    # kernel over NVLink with NVLS multicast (default), the same with plain peer loads/stores, and NCCL all_reduce.
    runs = {}
    for name, env in (("nvl", {}), ("nvl_p2p", {"G2V_CBOW_NVL_MULTICAST": "0"}), ("nccl", {"G2V_CBOW_NVL": "0"})):
        os.environ.update(env)
        got, info = g2v.train_cbow(rowptr, gene, label, 7523, 128, 0.005, max_epoch=11, seed=0, W_ih0=W0, W_ho0=Wo0,
                                   early_stop=False, log=None, return_info=True)
        for k in env:
            os.environ.pop(k)
        runs[name] = (got, info)
    got, info = runs["nvl"]
    # the collapsed trainer exchanges c (4*V bytes) with NCCL
    r1, r1_info = g2v.train_cbow(rowptr, gene, label, 7523, 128, 0.005, max_epoch=11, seed=0, W_ih0=W0, W_ho0=Wo0,
                                 early_stop=False, log=None, return_info=True, algo="rank1")
    if rank == 0:
        full = np.empty((2 * 7523, 80), np.int32); fl = np.empty(2 * 7523, np.int32)
        for r in range(world):
            full[r::world], fl[r::world] = gathered[r]
        np.savez(out, W=got, hist=np.array(info["history"], dtype=np.float64), nodes=full, lens=fl,
                 graph=np.array(info["graph"]), exchange=np.array([runs[k][1]["exchange"] for k in ("nvl", "nvl_p2p", "nccl")]),
                 W_p2p=runs["nvl_p2p"][0], W_nccl=runs["nccl"][0], W_rank1=r1,
                 hist_nccl=np.array(runs["nccl"][1]["history"], dtype=np.float64))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
