#!/usr/bin/env python
"""SASS opcode histogram of the shipped kernels: python profiles/sass/make_histogram.py > profiles/sass/opcode_histogram_r2.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(ROOT, "g2vec_b200", "libg2vec_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
funcs, cur = {}, None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", ln)
    if m and cur:
        funcs[cur][m.group(1)] += 1
names = dict(zip(funcs, subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()))
SHIP = ["walk_pair_kernel<true>", "walk_pair_kernel<false>", "walk_kernel<true, 2, true>", "walk_kernel<true, 2, false>", "walk_kernel<false, 1, true>", "walk_kernel<true, 0, false>",
        "cbow_rows_kernel<1, true, false, false>", "cbow_rows_kernel<1, false, false, false>", "cbow_rows_kernel<4, true, false, false>",
        "cbow_rows_kernel<1, true, true, false>", "cbow_rows_kernel<1, true, false, true>",
        "cbow_slab_fwd_kernel<4, 0, false, false>", "cbow_slab_bwd_kernel<4, false>", "cbow_slab_bwd_kernel<4, true>",
        "cbow_update_kernel<0>", "r1_windows_kernel<2>", "r1_update_kernel<1, 0>", "paths_insert_kernel", "paths_flag_kernel",
        "paths_emit_kernel", "loop_begin_kernel", "loop_decide_kernel", "pcc_edge_kernel"]
print("SASS opcode histograms of libg2vec_b200.so (cuobjdump -sass): the kernels the default paths launch + the TMA variants.")
print("ELF images:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
print("Markers: LDG.E.128 / LDG.E.64 = vector loads; REDG.E.ADD.F32x4 = red.global.add.v4.f32; UBLKCP = cp.async.bulk (TMA gather);")
print("UBLKRED = cp.reduce.async.bulk (TMA scatter); REDUX = warp integer reduction; SHFL = warp shuffles; ATOMG = global atomics.\n")
for mangled, c in funcs.items():
    short = re.sub(r"^void g2v::|^g2v::", "", names[mangled]).split("(")[0]
    if short not in SHIP:
        continue
    key = {k: v for k, v in c.items() if re.match(r"LDG|STG|REDG|UBLK|REDUX|SHFL|ATOM|LDS|STS|BAR|SYNCS|UTMA|VOTE|MUFU|FFMA$|IMAD$|LOP3|POPC|FLO", k)}
    print("%s  (%d instructions)" % (short, sum(c.values())))
    print("    " + "  ".join("%s=%d" % kv for kv in sorted(key.items(), key=lambda kv: -kv[1])))
