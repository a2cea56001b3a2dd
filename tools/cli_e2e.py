#!/usr/bin/env python
"""End-to-end wall clock of the `itsxpress` command line (FASTQ file in -> trimmed FASTQ file out) on a synthetic
BASELINE configs[1] sample written to disk.  Run on the GPU box:  python tools/cli_e2e.py [--scale S] [--gz]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402


def write_fastq(path, seq, off, qual):
    n = len(off) - 1
    lens = np.diff(off)
    titles = [b"@SYN:2:%d 1:N:0:1\n" % i for i in range(n)]
    tl = np.array([len(t) for t in titles], np.int64)
    rec = tl + lens + 1 + 2 + lens + 1
    ro = np.zeros(n + 1, np.int64)
    np.cumsum(rec, out=ro[1:])
    out = np.empty(int(ro[-1]), np.uint8)
    tcat = np.frombuffer(b"".join(titles), np.uint8)
    to = np.zeros(n + 1, np.int64)
    np.cumsum(tl, out=to[1:])

    def scatter(dst_off, src, src_off, length):
        total = int(length.sum())
        ar = np.arange(total, dtype=np.int64)
        cs = np.zeros(len(length) + 1, np.int64)
        np.cumsum(length, out=cs[1:])
        within = ar - np.repeat(cs[:-1], length)
        out[np.repeat(dst_off, length) + within] = src[np.repeat(src_off, length) + within]
    scatter(ro[:-1], tcat, to[:-1], tl)
    p = ro[:-1] + tl
    scatter(p, seq, off[:-1], lens)
    p = p + lens
    out[p] = 10; out[p + 1] = ord("+"); out[p + 2] = 10
    scatter(p + 3, qual, off[:-1], lens)
    out[p + 3 + lens] = 10
    with open(path, "wb") as f:
        f.write(out.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--gz", action="store_true", help="gzipped output")
    ap.add_argument("--keeptemp", action="store_true")
    a = ap.parse_args()
    cfg = dict(synth.CONFIGS["c2"])
    for k in ("region", "taxa"):
        cfg.pop(k)
    cfg["n_reads"] = int(cfg["n_reads"] * a.scale)
    cfg["n_unique"] = int(cfg["n_unique"] * a.scale)
    seq, off, qual, which = synth.make_reads(2 * 1_000_003, with_qual=True, **cfg)
    tmp = tempfile.mkdtemp(prefix="itsx_e2e_")
    fq_in = os.path.join(tmp, "in.fastq")
    write_fastq(fq_in, seq, off, qual)
    from itsxpress_b200 import main as cli
    out = os.path.join(tmp, "out.fastq" + (".gz" if a.gz else ""))
    argv = ["--fastq", fq_in, "--single_end", "--outfile", out, "--region", "ITS1", "--taxa", "Metazoa",
            "--log", os.path.join(tmp, "log.txt"), "--tempdir", tmp]
    if a.keeptemp:
        argv.append("--keeptemp")
    for rep in range(2):                      # second run: page cache and CUDA context warm
        t0 = time.perf_counter()
        cli.main(args=cli.myparser().parse_args(argv))
        dt = time.perf_counter() - t0
    n = len(off) - 1
    print(json.dumps({"cli_e2e_reads_per_s": n / dt, "seconds": dt, "reads": n, "input_bytes": os.path.getsize(fq_in),
                      "output_bytes": os.path.getsize(out), "gz": a.gz, "keeptemp": a.keeptemp}))
    print(open(os.path.join(tmp, "log.txt")).read()[-1500:])


if __name__ == "__main__":
    main()
