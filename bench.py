#!/usr/bin/env python
"""bench.py -- throughput of the two G2Vec hot paths on B200 (BASELINE.json metric:
"CBOW context-windows/sec and random-walk steps/sec at 1/2/4/8 B200 vs CPU ref").

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference's CPU path (oracle port)

One "step":  CBOW  = one iteration of the reference's training loop (G2Vec.py:262-267): a full-batch
                     optimizer step over all training windows (fwd+bwd+[all-reduce]+update) plus the
                     validation and training accuracy passes and the host read of the accuracies;
             walks = one pass of the sampler over every walker of both patient groups.
The headline `value` is CBOW context windows/s (training windows x steps / time, eval passes inside the
timed region, as the reference runs them); the walk sampler's steps/s is reported in the `walk` object of
the same line.  Workload at N=1: BASELINE configs[1] (synthetic 10k genes / 500k edges per group,
128-dim, lenPath 80, 10 repetitions -> 200k walkers / ~200k windows of 80 genes).  N>1: weak scaling --
every rank keeps that per-GPU work (numRepetition = 10*N), parameters replicated, dense gradient
NCCL-all-reduced once per step.

Timing: CUDA events on the launching stream, W warm-up steps, L2 flushed (256 MiB write) before every
timed step, max over ranks.  CPU baseline: oracle port of the reference's own dense formulation, timed in
the same run on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cbow_context_windows_per_sec"
UNIT = "windows/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="syn10k", choices=["syn10k", "syn20k", "syn50k", "stress200k", "ex"])
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    p.add_argument("--reps", type=int, default=10, help="numRepetition per GPU (weak) or in total (strong)")
    p.add_argument("--optimizer", default="adam", choices=["adam", "sgd"])
    p.add_argument("--algo", default="rows", choices=["rows", "rank1"],
                   help="CBOW formulation the headline value is measured on (rows = north_star's gather/scatter kernel)")
    p.add_argument("--no-alt-algo", action="store_true", help="do not also time the other formulation")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-sample-windows", type=int, default=32768)
    p.add_argument("--cpu-sample-starts", type=int, default=2000)
    return p.parse_args()


def workload(name):
    from g2vec_b200 import graph
    if name == "ex":
        z = np.load(os.path.join(ROOT, "tests", "golden", "ex_graph.npz"))
        gs = [(z["rowptr%d" % g], z["col%d" % g], z["w%d" % g]) for g in (0, 1)]
        return gs, 7523, 128, 80, "ex_* graphs (tests/golden/ex_graph.npz, made by the reference's construct_adjMat)"
    V, E, D, L = graph.BENCH_CONFIGS[name]
    gs = [graph.synthetic_graph(V, E, g) for g in (0, 1)]
    return gs, V, D, L, "synthetic directed ER, %d genes / %d edges per group, weights U(0.5,1)" % (V, E)


_RUN = {"reps": None, "world": 1}


def traffic_lookup(kernel, workload_name):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernel from the
    committed ncu --set full capture of this same command (profiles/traffic.json); None when no capture
    exists for this workload / numRepetition / GPU count."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[kernel][workload_name]
        return e["dram_bytes"] if (e.get("reps") == _RUN["reps"] and _RUN["world"] == 1) else None
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        load = [x for x in sm if x > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# =============================================================================== this repo's arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created; keep stdout for the one JSON
        # line by pointing fd 1 at stderr while the group comes up (first collective included)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    assert world == args.gpus, "--gpus must equal WORLD_SIZE"
    import g2vec_b200 as g2v
    from g2vec_b200 import _capi, paths, cbow

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t)
        return float(t[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    _RUN.update(reps=args.reps, world=world)
    gs, V, D, L, desc = workload(args.workload)
    reps_total = args.reps * (world if args.scaling == "weak" else 1)
    graphs = [g2v.WalkGraph(rp, col, weights=w) for rp, col, w in gs]
    n_walk = g2v.walks.num_walkers(V, reps_total, rank, None, world)
    outs = [(torch.empty((n_walk, L), dtype=torch.int32, device=dev), torch.empty((n_walk,), dtype=torch.int32, device=dev))
            for _ in (0, 1)]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def walk_pass():
        for g in (0, 1):
            g2v.generate_paths(graphs[g], L, reps_total, seed=12345, group=g, walker_begin=rank, walker_stride=world,
                               out=outs[g])

    def timed(fn, n, marks=0):
        """n steps, each: L2 flush, start event, fn(marks...), end event.  Returns per-step ms (+ inner marks)."""
        pairs = []
        for i in range(n):
            flush_buf.fill_(i & 0xFF)
            a, b = ev(), ev()
            inner = [ev() for _ in range(marks)]
            a.record(); fn(*inner) if marks else fn(); b.record()
            pairs.append((a, b, inner))
        torch.cuda.synchronize()
        tot = [a.elapsed_time(b) for a, b, _ in pairs]
        inn = [[a.elapsed_time(m) for m in inner] for a, _, inner in pairs]
        return tot, inn

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _capi.launch_count()

    # ------------------------------------------------------------------ walks
    timed(walk_pass, W)
    barrier()
    l0 = _capi.launch_count()
    wt, _ = timed(walk_pass, K)
    barrier()
    walk_launches = _capi.launch_count() - l0
    walk_ms = allmax(float(np.mean(wt)))
    visits_loc = int(sum(int(o[1].sum()) for o in outs))
    visits = allsum(visits_loc)
    # algorithmic bytes of one pass: per visit 4 B (node id written); per visit that scans its row
    # (every visit but the L-th of a full-length walk) 8 B rowptr + 8 B per neighbour (col + weight)
    wbytes = 0
    for g in (0, 1):
        nodes, lens = outs[g]
        deg = (graphs[g].rowptr[1:] - graphs[g].rowptr[:-1]).to(torch.int64)
        scan = nodes[:, :L - 1] if L > 1 else nodes[:, :0]
        m = scan >= 0
        wbytes += int(lens.sum()) * 4 + int(m.sum()) * 8 + 8 * int(deg[scan[m].to(torch.int64)].sum())
    walk_gbs = wbytes / (float(np.mean(wt)) * 1e-3) / 1e9

    walk_e2e = None
    if not args.no_e2e:
        qws = [g2v.graph.quantise_weights(w) for _, _, w in gs]
        def walk_host_pass():
            for g in (0, 1):
                g2v.generate_paths_host(gs[g][0], gs[g][1], qws[g], L, reps_total, seed=12345, group=g,
                                        walker_begin=rank, walker_stride=world)
        walk_host_pass()
        barrier(); t0 = time.perf_counter()
        for _ in range(max(1, min(K, 3))):
            walk_host_pass()
        barrier(); dt = (time.perf_counter() - t0) / max(1, min(K, 3))
        dt = allmax(dt)
        csr_b = sum(4 * (len(rp)) + 8 * len(col) for rp, col, _ in gs)
        walk_e2e = {"value": visits / dt, "unit": "steps/s", "h2d_bytes_per_step": int(csr_b),
                    "d2h_bytes_per_step": int(2 * n_walk * (L + 1) * 4), "api": "g2v_walk_host (C ABI, host buffers)"}

    # ------------------------------------------------------------------ windows from the walks
    rows = [paths.canonical_rows(*outs[g]) for g in (0, 1)]
    prow, plab = paths.integrate(rows[0], rows[1])
    rowptr, gene, label = paths.windows_csr(prow, plab)
    del rows, prow, plab
    N_loc = int(rowptr.shape[0]) - 1
    lens_np = np.diff(rowptr.cpu().numpy()).astype(np.int64)
    tr, va = cbow.split_indices(N_loc, 1000 + rank)
    n_tr_tot, n_va_tot = int(allsum(len(tr))), int(allsum(len(va)))
    W0, Wo0 = cbow.init_weights(V, D, 0)
    tr_d = torch.from_numpy(tr.astype(np.int32)).to(dev); va_d = torch.from_numpy(va.astype(np.int32)).to(dev)
    acc_pin = torch.zeros(4, dtype=torch.int64).pin_memory()
    ltr = lens_np[tr]
    peak, peak_src = peaks()
    l2_bytes = 126e6

    def make_step(model):
        def cbow_step(m_fb=None, m_upd=None, m_val=None):
            model.acc.zero_()
            model.fwdbwd(tr_d, n_tr_tot)
            if m_fb is not None:
                m_fb.record()
            if world > 1:
                for g in model.grad_tensors():
                    dist.all_reduce(g)
            model.update()
            if m_upd is not None:
                m_upd.record()
            model.evaluate(va_d, 2)
            if m_val is not None:
                m_val.record()
            model.evaluate(tr_d, 3)
            if world > 1:
                dist.all_reduce(model.acc[2:4])
            acc_pin.copy_(model.acc, non_blocking=True)
        return cbow_step

    def measure(algo):
        model = g2v.CbowModel(rowptr, gene, label, V, D, W0, Wo0, optimizer=args.optimizer, lr=0.005, algo=algo)
        model.prepare_csc(tr_d)                    # rank1: transposed incidence of the static training list
        step = make_step(model)
        timed(step, W)
        barrier()
        l0 = _capi.launch_count()
        ct, marks = timed(step, K, marks=3)
        barrier()
        r = {"launches": _capi.launch_count() - l0, "step_ms": allmax(float(np.mean(ct))),
             "fb_ms": float(np.mean([m[0] for m in marks])),
             "upd_ms": allmax(float(np.mean([m[1] for m in marks]))),
             "val_ms": allmax(float(np.mean([m[2] for m in marks]))),
             "acc_val": int(acc_pin[2]) / max(n_va_tot, 1), "model": model, "step": step}
        r["eager_ms"] = r["step_ms"]
        if world == 1:
            # what g2vec_b200.train_cbow runs on one GPU: the same launches replayed as one CUDA graph
            gstep = model.make_step(tr_d, n_tr_tot, va_d, acc_pin, True)
            timed(gstep, W)
            barrier()
            gt, _ = timed(gstep, K)
            barrier()
            r["step_ms"] = float(np.mean(gt))
            r["launches"] += 0                 # replays launch the same kernels; counted once below
            r["graph"] = True
        r["value"] = n_tr_tot / (r["step_ms"] * 1e-3)
        return r

    def roofline_of(algo, r):
        opt_b = (32 if args.optimizer == "adam" else 16) * V * D
        if algo == "rows":
            b = int((ltr * (8 * D + 4) + 5).sum())          # l*(8D+4)+5 per window
            gbs = b / (r["fb_ms"] * 1e-3) / 1e9
            resident = V * D * 4 < l2_bytes
            return {"kernel": "cbow_rows_kernel<%d,true> (fused gather/sum/logit/BCE/scatter-add)" % max(D // 128, 0),
                    "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                    "traffic": traffic_lookup("cbow_rows_fwdbwd", args.workload), "peak_source": peak_src,
                    "kernel_ms": r["fb_ms"], "algorithmic_bytes_per_launch": b,
                    "bytes_model": "sum over this rank's training windows of l*(8D+4)+5; the optimizer epilogue "
                                   "(%d B) is a separate kernel" % opt_b,
                    "note": ("W_ih + gradient (%.0f MB) are L2-resident at this config: the algorithmic rate can exceed "
                             "DRAM traffic and the HBM peak" % (2 * V * D * 4 / 1e6)) if resident else
                            ("W_ih + gradient (%.0f MB) exceed L2: the gather and the scatter-add go to HBM"
                             % (2 * V * D * 4 / 1e6))}
        b = (28 if args.optimizer == "adam" else 12) * V * D      # update rows R/W W,m,v + prepare re-read of W
        ms = r["upd_ms"] - r["fb_ms"]
        gbs = b / (ms * 1e-3) / 1e9
        return {"kernel": "r1_update_kernel + r1_update_ho_kernel + r1_prepare_kernel (dense optimizer pass)",
                "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": traffic_lookup("r1_update", args.workload), "peak_source": peak_src, "kernel_ms": ms,
                "algorithmic_bytes_per_launch": b,
                "bytes_model": "28*V*D (Adam: read W,m,v + write W,m,v, then re-read W for s); the window kernels "
                               "(forward + CSC segmented sum) move only %d B (16*l+9 per window) in %.3f ms"
                               % (int((ltr * 16 + 9).sum()), r["fb_ms"])}

    res = {args.algo: measure(args.algo)}
    alt = "rank1" if args.algo == "rows" else "rows"
    if not args.no_alt_algo:
        res[alt] = measure(alt)
    main = res[args.algo]
    model, cbow_step = main["model"], main["step"]
    step_ms, upd_ms, value, acc_val, cbow_launches = main["step_ms"], main["upd_ms"], main["value"], main["acc_val"], main["launches"]

    e2e = None
    if not args.no_e2e:
        # Every step's windows come from pinned host memory.  The upload of step k+1 runs on a copy stream
        # into the other of two device buffer sets while step k computes (input double-buffering); the timed
        # region contains every copy and the per-step device->host read of the accuracies.
        orig = (model.rowptr, model.gene, model.label)
        feeder = g2v.WindowFeeder(model, *orig)

        def e2e_run(n):
            feeder.upload(0)
            for i in range(n):
                k = i & 1
                if i + 1 < n:
                    feeder.upload(k ^ 1)
                feeder.use(k)
                cbow_step()
                feeder.release(k)
                torch.cuda.current_stream().synchronize()       # the accuracies are on the host
        e2e_run(2)
        barrier(); t0 = time.perf_counter()
        e2e_run(K)
        barrier(); dt = allmax((time.perf_counter() - t0) / K)
        model.rowptr, model.gene, model.label = orig
        e2e = {"value": n_tr_tot / dt, "unit": UNIT,
               "h2d_bytes_per_step": feeder.h2d_bytes,
               "d2h_bytes_per_step": 32,
               "api": "g2vec_b200.CbowModel step (C ABI kernels) fed by g2vec_b200.WindowFeeder: every step's windows "
                      "are uploaded from pinned host memory (double-buffered on a copy stream, gene ids as int16 when "
                      "n_genes <= 32768) and the accuracies are read back every step"}

    clocks = sampler.stop() if sampler else None
    total_launches = _capi.launch_count() - launches0

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu = None
    walk_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, walk_cpu = cpu_baseline(args, gs, V, D, L, rowptr.cpu().numpy(), gene.cpu().numpy(), label.cpu().numpy())

    walkers_total = int(allsum(2 * n_walk))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic" if args.workload != "ex" else "ex_* graphs (fixture)",
            "config": {"workload": "%s: %s, hidden %d, lenPath %d, numRepetition %d%s" % (
                           args.workload, desc, D, L, reps_total,
                           " (10 per GPU)" if args.scaling == "weak" and world > 1 else ""),
                       "windows_train": n_tr_tot, "windows_val": n_va_tot, "mean_window_len": float(lens_np.mean()),
                       "optimizer": args.optimizer, "step": "fwd+bwd+update + val acc + train acc (G2Vec.py:262-267)",
                       "launch": ("one CUDA graph replay per step (eager launches: %.3f ms per step)" % main["eager_ms"])
                                 if main.get("graph") else "eager launches",
                       "parallelism": "dp%d (windows/walkers sharded, W replicated, dense grad all-reduce per step)" % world,
                       "l2": "256 MiB flush write before every timed step"},
            "train_only": {"value": n_tr_tot / (upd_ms * 1e-3), "unit": UNIT, "ms_per_step": upd_ms,
                           "note": "fwd+bwd+all-reduce+update, without the two accuracy passes"},
            "production_loop": {
                "value": n_tr_tot / ((step_ms - 0.8 * (main["eager_ms"] - main["val_ms"])) * 1e-3), "unit": UNIT,
                "note": "what g2vec_b200.train_cbow runs: the training-accuracy pass of G2Vec.py:267 equals the next "
                        "step's training forward, so the separate pass is executed only on the steps that print it "
                        "(every 5th); derived from the same timed steps as `value`"},
            "acc_val_last": acc_val,
            "e2e": e2e,
            "gpu_launches": int(cbow_launches + walk_launches),
            "gpu_launches_total_process": int(total_launches),
            "clocks": clocks,
            "roofline": roofline_of(args.algo, main),
            "algo": args.algo,
            "alt_algo": None if alt not in res else {
                "algo": alt, "value": res[alt]["value"], "unit": UNIT, "ms_per_step": res[alt]["step_ms"],
                "train_only_ms": res[alt]["upd_ms"], "acc_val_last": res[alt]["acc_val"],
                "roofline": roofline_of(alt, res[alt]),
                "note": "rank1 = collapsed trainer (s = W_ih.W_ho, c = X^T.dO; SURVEY 8f-3), same results up to fp32 "
                        "reassociation; rows = north_star's embedding-row gather/scatter kernel"},
            "cpu_baseline": cpu,
            "walk": {"metric": "random_walk_steps_per_sec", "value": visits / (walk_ms * 1e-3), "unit": "steps/s",
                     "ms_per_pass": walk_ms, "walkers": walkers_total,
                     "visits_per_pass": visits,
                     "roofline": {"kernel": "walk_kernel", "bound": "hbm", "achieved": walk_gbs, "peak": peak,
                                  "unit": "GB/s", "frac": walk_gbs / peak,
                                  "traffic": traffic_lookup("walk", args.workload),
                                  "algorithmic_bytes_per_pass": wbytes,
                                  "bytes_model": "4 B per visit + (8 + 8*deg) B per visit that scans its row"},
                     "e2e": walk_e2e, "cpu_baseline": walk_cpu},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# =============================================================================== CPU side
def _cpu_windows(gs, V, L, n_windows):
    """Windows for the CPU arm, made by the oracle's walker (no GPU on this path)."""
    import oracle
    from oracle import legacy
    per = max(1, n_windows // 2)
    sets = []
    for g, (rp, col, w) in enumerate(gs):
        nodes, lens = oracle.walks(rp, col, oracle.quantise_weights(w), L, 12345, g, 0, min(per, 10 * V))
        sets.append(oracle.path_set(nodes, lens))
    rows = legacy.integrate_pathSet(sets)
    return legacy.windows_from_rows(rows)


class LazyDense:
    """adjMat[node] for graphs whose dense [V, V] float32 form does not fit: builds the dense row on demand
    (the per-step cost of the reference -- a V-long row copy -- is preserved)."""

    def __init__(self, rp, col, w):
        self.rp, self.col, self.w = rp, col, w
        self.shape = (len(rp) - 1, len(rp) - 1)

    def __getitem__(self, i):
        row = np.zeros(self.shape[0], dtype=np.float32)
        row[self.col[self.rp[i]:self.rp[i + 1]]] = self.w[self.rp[i]:self.rp[i + 1]]
        return row


def _dense_adj(rp, col, w):
    from oracle import legacy
    V = len(rp) - 1
    return legacy.dense_from_csr(rp, col, w) if V * V * 4 <= (4 << 30) else LazyDense(rp, col, w)


_WALK_A = None


def _walk_worker(job):
    from oracle import legacy
    starts, L, seed = job
    cnt = [0]
    legacy.generate_pathSet_dense(_WALK_A, L, 1, np.random.RandomState(seed), start_nodes=starts, counter=cnt)
    return cnt[0]


def cpu_walk_rate(gs, L, n_starts, procs):
    """Reference walk (dense-row port of G2Vec.py:324-352) on a sample of start nodes; `procs` processes,
    start nodes partitioned (1 = how the reference runs)."""
    global _WALK_A
    rp, col, w = gs[0]
    V = len(rp) - 1
    _WALK_A = _dense_adj(rp, col, w)
    starts = np.random.RandomState(0).choice(V, size=min(n_starts, V), replace=False).tolist()
    t0 = time.perf_counter()
    if procs <= 1:
        visits = _walk_worker((starts, L, 0))
    else:
        import multiprocessing as mp
        chunks = [starts[i::procs] for i in range(procs)]
        with mp.get_context("fork").Pool(procs) as pool:
            t0 = time.perf_counter()
            visits = sum(pool.map(_walk_worker, [(c, L, i) for i, c in enumerate(chunks) if c]))
    dt = time.perf_counter() - t0
    _WALK_A = None
    return visits / dt, visits, dt


def cpu_cbow_setup(rowptr, gene, label, V, D, n_sample):
    import torch
    from oracle import dense_cbow
    N = len(rowptr) - 1
    idx = np.random.RandomState(0).permutation(N)[:min(N, n_sample)]
    pivot = int(len(idx) * 0.8)
    tr, va = idx[:pivot], idx[pivot:]
    Xtr, ytr = dense_cbow.densify(rowptr, gene, label, tr, V)
    Xva, yva = dense_cbow.densify(rowptr, gene, label, va, V)
    rs = np.random.RandomState(0)
    W0 = (np.clip(rs.randn(V, D), -2, 2) / np.sqrt(D)).astype(np.float32)
    Wo0 = (np.clip(rs.randn(D), -2, 2) / np.sqrt(D)).astype(np.float32)
    return dense_cbow.DenseCbow(W0, Wo0, 0.005), (Xtr, ytr, Xva, yva), len(tr), len(va), torch.get_num_threads()


def cpu_baseline(args, gs, V, D, L, rowptr, gene, label):
    model, data, n_tr, n_va, threads = cpu_cbow_setup(rowptr, gene, label, V, D, args.cpu_sample_windows)
    model.epoch(*data)
    t0 = time.perf_counter(); n = 5
    for _ in range(n):
        model.epoch(*data)
    dt = (time.perf_counter() - t0) / n
    cpu = {"value": n_tr / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%d training + %d validation windows (dense X [%d,%d] f32), 5 epochs of the reference's dense "
                     "formulation (oracle/dense_cbow.py, torch-CPU matmul); TensorFlow 1.x is not installable here"
                     % (n_tr, n_va, n_tr, V), "ms_per_step": dt * 1e3}
    rate, visits, wdt = cpu_walk_rate(gs, L, args.cpu_sample_starts, 1)
    walk_cpu = {"value": rate, "unit": "steps/s", "cores": 1, "kind": "port",
                "sample": "%d start nodes x 1 repetition of group 0 (%d node visits, %.1f s), dense-row port of "
                          "G2Vec.py:324-352 (oracle/legacy.py), single thread as the reference runs"
                          % (min(args.cpu_sample_starts, V), visits, wdt)}
    return cpu, walk_cpu


def run_reference(args):
    """The reference's own CPU implementation of the path, timed on this box's host cores.  The reference
    is a Python script that needs TensorFlow 1.x for step 4 and is not present on the GPU box, so this arm
    runs the oracle PORT of its algorithm (kind "port"): dense X matmuls on all torch threads for CBOW,
    dense-row NumPy walks on all cores (one process per core, start nodes partitioned) for the walks."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    K, W = args.steps, args.warmup
    gs, V, D, L, desc = workload(args.workload)
    rowptr, gene, label = _cpu_windows(gs, V, L, args.cpu_sample_windows)
    model, data, n_tr, n_va, threads = cpu_cbow_setup(rowptr, gene, label, V, D, args.cpu_sample_windows)
    for _ in range(W):
        model.epoch(*data)
    t0 = time.perf_counter()
    for _ in range(K):
        av, at = model.epoch(*data)
    dt = (time.perf_counter() - t0) / K
    value = n_tr / dt
    cores = os.cpu_count() or 1
    rate, visits, wdt = cpu_walk_rate(gs, L, max(args.cpu_sample_starts, 100 * cores), cores)
    sample = ("each step = one epoch (G2Vec.py:262-267) of the dense formulation on %d training + %d validation "
              "windows of the workload (dense X [%d,%d] f32)" % (n_tr, n_va, n_tr, V))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic" if args.workload != "ex" else "ex_* graphs (fixture)",
        "config": {"workload": "%s: %s, hidden %d, lenPath %d" % (args.workload, desc, D, L), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "walk": {"metric": "random_walk_steps_per_sec", "value": rate, "unit": "steps/s", "cores": cores, "kind": "port",
                 "sample": "%d node visits in %.1f s: dense-row port of G2Vec.py:324-352, %d processes"
                           % (visits, wdt, cores)},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
