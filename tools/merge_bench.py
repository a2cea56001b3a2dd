#!/usr/bin/env python
"""Throughput of the paired-end merge stage (csrc/merge.cu) on synthetic Illumina-like pairs of the
BASELINE configs[4] shape (2 x 250 bp off 330-441 bp fragments), beside the CPU oracle on the box's host cores.
Prints one JSON line.   usage: python tools/merge_bench.py [--pairs N] [--steps K] [--warmup W] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import synth
    from itsxpress_b200 import _lib
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(4 * 1_000_003, a.pairs, frag_len=(330, 441), read_len=250)
    ctx = _lib.Context(0)
    prm = _lib.merge_params()
    ms_k, ms_e2e = [], []
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        ml, why, idx, ooff, oseq, oqual = ctx.merge_pairs(fs, fq, fo, rs, rq, ro, prm)
        t1 = time.perf_counter()
        if it >= a.warmup:
            ms_k.append(ctx.merge_stats().ms_kernel)
            ms_e2e.append((t1 - t0) * 1e3)
    st = ctx.merge_stats()
    k = float(np.mean(ms_k))
    out = {
        "stage": "paired-end merge (itsx_merge_pairs)", "pairs": a.pairs, "merged": int(st.n_merged),
        "by_reason": {_lib.MERGE_REASONS[r]: int(st.by_reason[r]) for r in range(10) if st.by_reason[r]},
        "kernel_ms": k, "kernel_pairs_per_s": a.pairs / k * 1e3,
        "kernel_hbm_gbs": (st.bytes_in + st.bytes_out) / k / 1e6,
        "diagonal_cells_per_s": float(np.sum(np.diff(fo).astype(np.float64) * np.diff(ro))) / k * 1e3,
        "e2e_ms_host_buffers": float(np.mean(ms_e2e)), "e2e_pairs_per_s": a.pairs / float(np.mean(ms_e2e)) * 1e3,
        "h2d_bytes": int(st.bytes_in + 16 * a.pairs), "d2h_bytes": int(st.bytes_out + 17 * a.pairs),
    }
    if not a.no_cpu:
        from oracle import oracle as O
        n = min(a.pairs, 40_000)
        t0 = time.perf_counter()
        O.merge_pairs(fs[:fo[n]], fq[:fo[n]], fo[:n + 1], rs[:ro[n]], rq[:ro[n]], ro[:n + 1])
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"pairs_per_s": n / dt, "cores": os.cpu_count(), "sample": "first %d pairs" % n,
                             "kind": "port"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
