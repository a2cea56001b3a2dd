#!/usr/bin/env python
"""Rate of the native gzip reader (csrc/inflate_host.cpp) against zlib on a synthetic BASELINE configs[1] FASTQ file:
one gzip member written by zlib level 6 (a sequencer's file), inflated by zlib, by the reader on one core and on
1..N cores.  Host only (no GPU):  python tools/gunzip_rate.py [--scale S] [--threads 1,2,4,8]"""
import argparse
import gzip
import json
import os
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
from cli_e2e import write_fastq  # noqa: E402
from itsxpress_b200 import fastq as fq  # noqa: E402


def best(f, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = f()
        t.append(time.perf_counter() - t0)
    return min(t), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.2)
    ap.add_argument("--threads", default="")
    a = ap.parse_args()
    seq, off, _, _ = synth.make_config("c2", seed=5, scale=a.scale)
    qual = synth.make_quals(77, off)
    tmp = tempfile.mkdtemp(prefix="itsx_gunzip_")
    path = os.path.join(tmp, "a.fastq")
    write_fastq(path, seq, off, qual)
    data = open(path, "rb").read()
    os.remove(path)
    os.rmdir(tmp)
    comp = gzip.compress(data, 6)
    cores = os.cpu_count() or 1
    threads = [int(t) for t in a.threads.split(",")] if a.threads else sorted({1, 2, 4, 8, 16, cores} & set(range(1, cores + 1)))
    mb = len(data) / 1e6
    t_z, ref = best(lambda: zlib.decompress(comp, 31))
    assert ref == data
    out = {"workload": "configs[1] FASTQ x %g: %d reads, %.0f MB of text, %.0f MB as one gzip member (zlib -6)" %
                       (a.scale, len(off) - 1, mb, len(comp) / 1e6),
           "host_cores": cores, "zlib_MB_per_s": mb / t_z, "native_MB_per_s": {}}
    for th in threads:
        t, got = best(lambda: fq.gunzip(comp, th))
        assert got.tobytes() == data
        out["native_MB_per_s"][str(th)] = mb / t
    out["speedup_vs_zlib"] = {k: v / out["zlib_MB_per_s"] for k, v in out["native_MB_per_s"].items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
