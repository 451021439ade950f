"""Glue between the two hot paths: canonicalise walks, de-duplicate, drop paths
common to both groups, emit CSR context windows, and the gene-frequency vote.

Reference: ``tuple(sorted(path))`` into a set (G2Vec.py:345,351), ``integrate_pathSet``
(:310-322, dense int32 [n_paths, n_genes+1]) and ``count_geneFreq`` (:288-308).  Here paths stay on
the device as padded sorted rows and the dense multi-hot matrix (1.37 GB at ex_* scale) is never
built: the trainer consumes CSR windows.  Row sorting and the exact duplicate / cross-group tests are
csrc/g2v_paths.cu; the 8-byte key sort, the compaction and the prefix sums are torch plumbing.
"""
import numpy as np
import torch

PAD = 2**31 - 1


def _canon(nodes):
    """csrc/g2v_paths.cu: every row sorted ascending (PAD at the end) + a 64-bit key per row."""
    from . import _capi
    lib = _capi.load()
    nodes = nodes.contiguous()
    n, L = nodes.shape
    rows = torch.empty_like(nodes)
    key = torch.empty((n,), dtype=torch.int64, device=nodes.device)
    st = torch.cuda.current_stream(nodes.device).cuda_stream
    _capi.check(lib.g2v_paths_canonicalise(nodes.data_ptr(), n, L, rows.data_ptr(), key.data_ptr(), st),
                "g2v_paths_canonicalise")
    return rows, key


def _mark(rows, key, group=None):
    """Key-order visit of the rows; returns (perm, flag) with flag[i] about row perm[i] (see g2v_paths_mark)."""
    from . import _capi
    lib = _capi.load()
    n, L = rows.shape
    ks, perm = torch.sort(key)                         # radix sort of 8-byte keys (plumbing)
    flag = torch.empty((n,), dtype=torch.uint8, device=rows.device)
    st = torch.cuda.current_stream(rows.device).cuda_stream
    _capi.check(lib.g2v_paths_mark(rows.data_ptr(), ks.data_ptr(), perm.data_ptr(),
                                   0 if group is None else group.data_ptr(), n, L, flag.data_ptr(), st),
                "g2v_paths_mark")
    return perm, flag.bool()


def canonical_rows(nodes, lens=None):
    """Walk rows (visit order, -1 padded) -> the set of G2Vec.py:345,351: unique rows, each sorted ascending
    and PAD-padded.  Row order: ascending key (arbitrary but deterministic)."""
    if nodes.shape[0] == 0:
        return torch.empty_like(nodes)
    rows, key = _canon(nodes)
    perm, first = _mark(rows, key)
    return rows[perm[first]]


def gather_walker_shards(dist, world, n_total, *shards):
    """Multi-GPU: rank r ran walkers r, r+world, ... (< n_total) and holds ``shards`` (tensors whose first dimension is
    its walker count).  all_gather them and put every row back at its walker index, so that each rank ends up with
    the arrays a single GPU would have produced -- same order, hence the same windows order and the same ``--seed``
    train/validation split whatever the number of GPUs.  Works on any backend (CPU tensors + gloo in the tests)."""
    per = -(-n_total // world)
    out = []
    for t in shards:
        pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        full = torch.empty((n_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        for r in range(world):
            full[r::world] = parts[r][:len(range(r, n_total, world))]
        out.append(full)
    return out


def unique_rows(rows, key):
    """The set of G2Vec.py:351 from rows that are ALREADY canonical (sorted, PAD-padded) with their 64-bit keys --
    what ``walks.generate_paths(..., canonical=True)`` returns: first occurrences in ascending key order."""
    if rows.shape[0] == 0:
        return rows
    perm, first = _mark(rows, key)
    return rows[perm[first]]


def integrate(rows_good, rows_poor):
    """integrate_pathSet (G2Vec.py:310-322): remove paths present in both groups, label the rest
    (0 = good, 1 = poor).  Returns (rows [N, L] PAD-padded, labels uint8 [N]); good rows first."""
    L = max(rows_good.shape[1], rows_poor.shape[1])

    def widen(r):
        if r.shape[1] == L:
            return r
        pad = torch.full((r.shape[0], L - r.shape[1]), PAD, dtype=r.dtype, device=r.device)
        return torch.cat([r, pad], dim=1)

    a, b = widen(rows_good), widen(rows_poor)
    both = torch.cat([a, b], dim=0).contiguous()
    lab = torch.cat([torch.zeros(a.shape[0], dtype=torch.uint8, device=a.device),
                     torch.ones(b.shape[0], dtype=torch.uint8, device=b.device)])
    if both.shape[0] == 0:
        return both, lab
    _, key = _canon(both)                              # rows are already sorted: this only makes the keys
    perm, survives = _mark(both, key, group=lab)
    idx, _ = torch.sort(perm[survives])                # back to input order: good rows first
    return both[idx], lab[idx]


def windows_csr(rows, labels):
    """Padded sorted rows -> CSR windows (rowptr int32 [N+1], gene int32 [nnz], label uint8 [N])."""
    valid = rows != PAD
    lens = valid.sum(dim=1)
    rowptr = torch.zeros(rows.shape[0] + 1, dtype=torch.int64, device=rows.device)
    rowptr[1:] = torch.cumsum(lens, dim=0)
    if int(rowptr[-1]) >= 2**31:
        raise ValueError("too many window entries for int32 CSR")
    gene = rows[valid].to(torch.int32)
    return rowptr.to(torch.int32), gene, labels.to(torch.uint8)


def gene_freq_codes(rowptr, gene, labels, n_genes):
    """count_geneFreq (G2Vec.py:288-308) as a vector: code[g] = 0 (more good paths), 1 (more poor),
    2 (tie), -1 (gene in no path).  The reference returns a dict over the genes that occur."""
    lens = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    lab = torch.repeat_interleave(labels.to(torch.int64), lens)
    g = gene.to(torch.int64)
    fp = torch.bincount(g, weights=lab.to(torch.float64), minlength=n_genes)
    tot = torch.bincount(g, minlength=n_genes).to(torch.float64)
    fg = tot - fp
    code = torch.full((n_genes,), -1, dtype=torch.int64, device=gene.device)
    code[(tot > 0) & (fg > fp)] = 0
    code[(tot > 0) & (fg < fp)] = 1
    code[(tot > 0) & (fg == fp)] = 2
    return code


def gene_freq_dict(code, gene_names):
    code = code.cpu().numpy().astype(np.int64)
    return {gene_names[i]: int(c) for i, c in enumerate(code) if c >= 0}


def build_windows(rows, lens, key, group, n_genes):
    """From the sampler's canonical output of BOTH groups to the trainer's input, without a sort
    (csrc/g2v_paths.cu, g2v_paths_set_*): ``rows`` int32 [n, L] sorted + PAD-padded, ``lens`` int32 [n], ``key``
    int64 [n] (walks.generate_paths(..., canonical=True)), ``group`` uint8 [n] (0 good / 1 poor).

    Returns (rowptr int32 [N+1], gene int32 [nnz], label uint8 [N], code int8 [n_genes]): the CSR windows of
    integrate_pathSet (G2Vec.py:310-322) -- duplicates inside a group and paths common to both groups removed, kept
    rows in input order, i.e. first occurrence in (group, repetition, start gene) order -- and count_geneFreq's vote
    per gene (:288-308; -1 = gene in no path).  A 64-bit key shared by two different rows (never observed) makes the
    function fall back to the exact sort-based path."""
    from . import _capi
    lib = _capi.load()
    rows, lens, key, group = rows.contiguous(), lens.contiguous(), key.contiguous(), group.contiguous()
    n, L = rows.shape
    dev = rows.device
    st = torch.cuda.current_stream(dev).cuda_stream
    ws = torch.empty(max(int(lib.g2v_paths_set_workspace_bytes(n)), 8), dtype=torch.uint8, device=dev)
    keep = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    totals = torch.zeros(3, dtype=torch.int64, device=dev)
    _capi.check(lib.g2v_paths_set_select(rows.data_ptr(), key.data_ptr(), group.data_ptr(), lens.data_ptr(), n, L,
                                         ws.data_ptr(), keep.data_ptr(), totals.data_ptr(), st), "g2v_paths_set_select")
    kept, nnz, clashes = (int(x) for x in totals.cpu())          # the pipeline's one host sync: sizes of the outputs
    if clashes:                                                   # exact path: sort by key + full row comparisons
        g0, g1 = group == 0, group == 1
        prow, plab = integrate(unique_rows(rows[g0], key[g0]), unique_rows(rows[g1], key[g1]))
        rowptr, gene, label = windows_csr(prow, plab)
        return rowptr, gene, label, gene_freq_codes(rowptr, gene, label, n_genes).to(torch.int8)
    if nnz >= 2**31:
        raise ValueError("too many window entries for int32 CSR")
    rowptr = torch.empty(kept + 1, dtype=torch.int32, device=dev)
    gene = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    label = torch.empty(max(kept, 1), dtype=torch.uint8, device=dev)[:kept]
    freq = torch.empty(2 * n_genes, dtype=torch.int32, device=dev)
    code = torch.empty(n_genes, dtype=torch.int8, device=dev)
    _capi.check(lib.g2v_paths_set_emit(rows.data_ptr(), group.data_ptr(), lens.data_ptr(), keep.data_ptr(), n, L,
                                       int(n_genes), ws.data_ptr(), kept, nnz, rowptr.data_ptr(), gene.data_ptr(),
                                       label.data_ptr(), freq.data_ptr(), code.data_ptr(), st), "g2v_paths_set_emit")
    return rowptr, gene, label, code


def rows_to_set(rows):
    """Padded rows -> python set of tuples (for the reference-shaped adapters and the tests)."""
    r = rows.cpu().numpy()
    return {tuple(int(x) for x in row[row != PAD]) for row in r}


def dense_pathlist_to_csr(pathList):
    """The reference's dense pathList [N, n_genes+1] (last column = label) -> CSR windows (NumPy)."""
    P = np.asarray(pathList)
    X = P[:, :-1]
    r, c = np.nonzero(X)
    rowptr = np.zeros(P.shape[0] + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), c.astype(np.int32), P[:, -1].astype(np.uint8)
