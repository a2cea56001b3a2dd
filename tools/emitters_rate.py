#!/usr/bin/env python
"""Native against Python emitters of the inter-stage files (domtbl.txt, uc.txt, rep.fa, record ids) on a synthetic
BASELINE configs[1] sample; host only.  python tools/emitters_rate.py [--reads N]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
from cli_e2e import write_fastq  # noqa: E402
from itsxpress_b200 import _lib, fastq as fq, host  # noqa: E402


def timed(f, *a, **k):
    t0 = time.perf_counter()
    r = f(*a, **k)
    return time.perf_counter() - t0, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200000)
    ap.add_argument("--rows-per-unique", type=int, default=40)
    a = ap.parse_args()
    seq, off, _, _ = synth.make_config("c2", seed=5, scale=a.reads / 1e6)
    tmp = tempfile.mkdtemp(prefix="itsx_emit_")
    path = os.path.join(tmp, "in.fastq")
    write_fastq(path, seq, off, synth.make_quals(77, off))
    b = fq.read_fastq(path)
    s, o = b.seq_concat()
    # exact classes of the forward strand (a measuring tool: no derep engine needed for synthetic reads)
    first_of = {}
    rep = np.empty(b.n, np.int32)
    for i in range(b.n):
        rep[i] = first_of.setdefault(s[o[i]:o[i + 1]].tobytes(), i)
    strand, nu = np.zeros(b.n, np.uint8), len(first_of)
    out = {"reads": b.n, "uniques": nu, "host_cores": fq.host_share(), "seconds": {}}
    t_ids, ids = timed(b.ids)
    order = host.cluster_order(rep, ids)
    t1, x = timed(host.write_uc, rep, strand, ids, b.s_len, order, batch=b)
    t2, y = timed(host.write_uc_py, rep, strand, ids, b.s_len, order)
    assert x == y
    out["seconds"]["uc.txt"] = {"native": t1, "python": t2, "bytes": len(x)}
    t1, x = timed(host.write_rep_fasta, b, order, ids)
    t2, y = timed(host.write_rep_fasta_py, b, order, ids)
    assert x == y
    out["seconds"]["rep.fa"] = {"native": t1, "python": t2, "bytes": len(x)}
    out["seconds"]["ids"] = {"native": t_ids}
    # a domain table of the size a real search leaves for this sample (rows per unique as measured on configs[1])
    rng = np.random.default_rng(1)
    nprof = 98
    n = nu * a.rows_per_unique
    rows = np.zeros(n, dtype=_lib.ROW_DTYPE)
    rows["prof"] = np.sort(rng.integers(0, nprof, n))
    rows["seq"] = rng.integers(0, nu, n)
    rows["ienv"] = rng.integers(1, 200, n)
    rows["jenv"] = rows["ienv"] + 40
    rows["tlen"] = 250
    rows["bitscore"] = rng.normal(30, 10, n)
    rows["seq_score"] = rows["bitscore"]
    rows["lnP"] = -rng.exponential(20, n)
    rows["seq_lnP"] = rows["lnP"]
    seq_ids = [ids[i] for i in np.flatnonzero(rep == np.arange(b.n)).tolist()]
    names = ["3_profile_%d" % i for i in range(nprof)]
    M, nrep = np.full(nprof, 45, np.int32), np.full(nprof, 1000, np.int32)
    t1, x = timed(host.write_domtbl, rows, seq_ids, names, M, nu, nrep)
    k = min(n, 300000)
    t2, y = timed(host.write_domtbl_py, rows[:k], seq_ids, names, M, nu, nrep)
    assert host.write_domtbl(rows[:k], seq_ids, names, M, nu, nrep) == y
    out["seconds"]["domtbl.txt"] = {"rows": n, "native": t1, "python_extrapolated_from_%d_rows" % k: t2 * n / k, "bytes": len(x)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
