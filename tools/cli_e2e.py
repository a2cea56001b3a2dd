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


def write_fastq(path, seq, off, qual, mate=1):
    n = len(off) - 1
    lens = np.diff(off)
    titles = [b"@SYN:2:%d %d:N:0:1\n" % (i, mate) for i in range(n)]
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


def gzip_one_member(data, level=1, piece=8 << 20):
    """``data`` as ONE gzip member (what a sequencer writes: a single deflate stream), compressed on every core the way pigz
    does it: independent raw-deflate pieces ending in a sync flush, concatenated, closed by an empty final block."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def part(i):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        return co.compress(data[i:i + piece]) + co.flush(zlib.Z_SYNC_FLUSH)
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        body = b"".join(ex.map(part, range(0, len(data), piece)))
    return (bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3]) + body + b"\x03\x00" +
            zlib.crc32(data).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little"))


def paired(a):
    n = int(781_250 * a.scale)
    frag, foff, _, _ = synth.make_reads(5 * 1_000_003, n, max(300, n // 20), (330, 441), "M.hmm", "3_", "4_", zipf_s=1.2,
                                        spacer=(150, 230))
    fs, fq_, fo, rs, rq, ro = synth.make_pairs(5 * 1_000_003 + 1, frag, foff, read_len=250, err_scale=0.0, n_rate=0.0)
    tmp = tempfile.mkdtemp(prefix="itsx_e2e_")
    r1, r2 = os.path.join(tmp, "in_R1.fastq"), os.path.join(tmp, "in_R2.fastq")
    write_fastq(r1, fs, fo, fq_, 1)
    write_fastq(r2, rs, ro, rq, 2)
    from itsxpress_b200 import main as cli
    ext = ".gz" if a.gz else ""
    o1, o2 = os.path.join(tmp, "out_R1.fastq" + ext), os.path.join(tmp, "out_R2.fastq" + ext)
    argv = ["--fastq", r1, "--fastq2", r2, "--outfile", o1, "--region", "ITS2", "--taxa", "Metazoa",
            "--log", os.path.join(tmp, "log.txt"), "--tempdir", tmp]
    if a.paired == "unmerged":
        argv += ["--outfile2", o2]
    for rep in range(2):
        t0 = time.perf_counter()
        cli.main(args=cli.myparser().parse_args(argv))
        dt = time.perf_counter() - t0
    print(json.dumps({"cli_e2e_pairs_per_s": n / dt, "seconds": dt, "pairs": n, "mode": a.paired,
                      "input_bytes": os.path.getsize(r1) + os.path.getsize(r2),
                      "output_bytes": os.path.getsize(o1) + (os.path.getsize(o2) if a.paired == "unmerged" else 0),
                      "gz": a.gz}))
    print(open(os.path.join(tmp, "log.txt")).read()[-1800:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--gz", action="store_true", help="gzipped output")
    ap.add_argument("--keeptemp", action="store_true")
    ap.add_argument("--paired", choices=["merged", "unmerged"], default=None,
                    help="one BASELINE configs[4] sample: 781 250 read pairs (2 x 250 bp off 330-441 bp ITS2 amplicons, 5 %% "
                         "unique), R1 + R2 files in, merged reads or trimmed R1 + R2 files out")
    a = ap.parse_args()
    if a.paired:
        return paired(a)
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
