#!/usr/bin/env python
"""Round-2 tuning harness (run on the GPU box): times one kernel family under a sweep of its tuning hooks.

    python profiles/r2/tune.py slabs  [--reps 2]      CBOW fwd+bwd on the stress table (V=200k, D=512), fused vs slabs
    python profiles/r2/tune.py walk                    walk sampler on syn10k / syn20k / stress200k, per layout
Prints one JSON line per setting (CUDA events on the launching stream, 256 MiB L2 flush before every timed run).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import g2vec_b200 as g2v                    # noqa: E402
from g2vec_b200 import graph, cbow          # noqa: E402


def synthetic_windows(N, V, L, seed=777, device="cuda"):
    """N windows of L distinct genes, labels Bernoulli(0.5).  Distinct + sorted by construction:
    L draws from [0, V-L] sorted, plus 0..L-1."""
    g = torch.Generator(device=device); g.manual_seed(seed)
    x = torch.randint(0, V - L + 1, (N, L), generator=g, device=device, dtype=torch.int32)
    x, _ = torch.sort(x, dim=1)
    x += torch.arange(L, device=device, dtype=torch.int32)[None, :]
    label = (torch.rand(N, generator=g, device=device) < 0.5).to(torch.uint8)
    rowptr = torch.arange(0, (N + 1) * L, L, device=device, dtype=torch.int32)
    return rowptr, x.reshape(-1).contiguous(), label


def timeit(fn, n=5, warm=2):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for i in range(n):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def slabs(args):
    V, D, L = args.V, args.D, 80
    N = 2 * args.reps * V
    rowptr, gene, label = synthetic_windows(N, V, L)
    n_tr = int(N * 0.8)
    W0 = (torch.randn(V, D, device="cuda") / np.sqrt(D)).clamp_(-2 / np.sqrt(D), 2 / np.sqrt(D))
    Wo0 = torch.randn(D, device="cuda") / np.sqrt(D)
    tr = torch.randperm(N, device="cuda")[:n_tr].to(torch.int32)
    va = torch.arange(n_tr, N, device="cuda", dtype=torch.int32)
    alg_bytes = n_tr * (L * (8 * D + 4) + 5)
    settings = [("fused", None, None)] + [("slabs", mb, grp) for mb in args.mb for grp in args.group]
    ref = None
    for kind, mb, grp in settings:
        for k in ("G2V_CBOW_SLABS", "G2V_CBOW_SLAB_MB", "G2V_CBOW_SLAB_FWD_GROUP"):
            os.environ.pop(k, None)
        if kind == "fused":
            os.environ["G2V_CBOW_SLABS"] = "1"
        else:
            os.environ["G2V_CBOW_SLAB_MB"] = str(mb); os.environ["G2V_CBOW_SLAB_FWD_GROUP"] = str(grp)
        m = g2v.CbowModel(rowptr, gene, label, V, D, W0, Wo0, lr=0.005)
        m.prepare_slabs(tr); m.prepare_slabs(va)
        def step():
            m.fwdbwd(tr, n_tr)
        ms, mn = timeit(step)
        def ev():
            m.evaluate(va, 2)
        ems, _ = timeit(ev, n=3, warm=1)
        m.g_ih.zero_(); m.g_ho.zero_(); m.acc.zero_()
        m.fwdbwd(tr, n_tr); torch.cuda.synchronize()
        chk = (float(m.g_ih.double().abs().sum()), float(m.g_ho.double().abs().sum()), int(m.acc[1]))
        if ref is None:
            ref = chk
        print(json.dumps({"kind": kind, "slab_mb": mb, "fwd_group": grp, "n_slabs": getattr(m, "_n_slabs", 1),
                          "fwdbwd_ms": ms, "fwdbwd_ms_min": mn, "alg_GBps": alg_bytes / ms / 1e6, "eval_ms": ems,
                          "n_tr": n_tr, "check_rel": [abs(chk[0] - ref[0]) / ref[0], abs(chk[1] - ref[1]) / ref[1], chk[2] - ref[2]]}),
              flush=True)
        del m
        torch.cuda.empty_cache()


def walk(args):
    for name in args.workloads:
        V, E, D, L = graph.BENCH_CONFIGS[name]
        reps = 10 if V <= 50_000 else 2
        for layout in ("csr", "e8", "auto"):
            os.environ.pop("G2V_WALK_LAYOUT", None)
            if layout == "e8":
                os.environ["G2V_WALK_LAYOUT"] = "e8"
            gs = [graph.synthetic_graph(V, E, g) for g in (0, 1)] if layout == "csr" else gs
            graphs = [g2v.WalkGraph(rp, col, weights=w) for rp, col, w in gs]
            n = V * reps
            for canon in (False, True):
                if canon and layout == "csr":
                    continue
                outs = [(torch.empty((n, L), dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"),
                         torch.empty(n, dtype=torch.int64, device="cuda")) for _ in (0, 1)]
                def run():
                    for g in (0, 1):
                        g2v.generate_paths(graphs[g], L, reps, seed=12345, group=g, out=outs[g], canonical=canon,
                                           plain_csr=(layout == "csr"))
                ms, mn = timeit(run)
                visits = int(outs[0][1].sum()) + int(outs[1][1].sum())
                print(json.dumps({"workload": name, "layout": layout, "packed_layout": graphs[0].layout, "canonical": canon,
                                  "pass_ms": ms, "pass_ms_min": mn, "visits": visits, "steps_per_s": visits / ms * 1e3}), flush=True)


def fused(args):
    """The fused fwd+bwd kernel on an L2-resident table (V=10k, D=128: BASELINE configs[1]'s shape) with the default
    LDG.128 / RED.128 memory path and with the two TMA forms (G2V_CBOW_SCATTER=tma -> UBLKRED, G2V_CBOW_GATHER=tma ->
    UBLKCP); under `ncu -k regex:cbow_rows_kernel` every variant contributes 5 launches in this order."""
    V, D, L = args.V, args.D, 80
    N = 200_000
    rowptr, gene, label = synthetic_windows(N, V, L)
    n_tr = int(N * 0.8)
    W0 = (torch.randn(V, D, device="cuda") / np.sqrt(D)).clamp_(-2 / np.sqrt(D), 2 / np.sqrt(D))
    Wo0 = torch.randn(D, device="cuda") / np.sqrt(D)
    tr = torch.randperm(N, device="cuda")[:n_tr].to(torch.int32)
    alg = n_tr * (L * (8 * D + 4) + 5)
    os.environ["G2V_CBOW_SLABS"] = "1"
    for name, env in (("ldg+red (default)", {}), ("tma scatter (UBLKRED)", {"G2V_CBOW_SCATTER": "tma"}),
                      ("tma gather (UBLKCP)", {"G2V_CBOW_GATHER": "tma"}),
                      ("tma gather + scatter", {"G2V_CBOW_GATHER": "tma", "G2V_CBOW_SCATTER": "tma"})):
        for k in ("G2V_CBOW_SCATTER", "G2V_CBOW_GATHER"):
            os.environ.pop(k, None)
        os.environ.update(env)
        m = g2v.CbowModel(rowptr, gene, label, V, D, W0, Wo0, lr=0.005)
        ms, mn = timeit(lambda: m.fwdbwd(tr, n_tr), n=3, warm=2)
        print(json.dumps({"variant": name, "V": V, "D": D, "fwdbwd_ms": ms, "fwdbwd_ms_min": mn, "alg_GBps": alg / ms / 1e6}), flush=True)
        del m


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("what", choices=["slabs", "walk", "fused"])
    p.add_argument("--reps", type=int, default=2)
    p.add_argument("--V", type=int, default=200_000)
    p.add_argument("--D", type=int, default=512)
    p.add_argument("--mb", type=float, nargs="+", default=[16, 24, 32, 48])
    p.add_argument("--group", type=int, nargs="+", default=[1, 2, 3])
    p.add_argument("--workloads", nargs="+", default=["syn10k", "syn20k", "stress200k"])
    a = p.parse_args()
    {"slabs": slabs, "walk": walk, "fused": fused}[a.what](a)
