#!/usr/bin/env python
"""Throughput and ratio of the GPU gzip writer (itsx_gzip_compress) on FASTQ text, host buffer in -> host buffer out,
next to zlib on the host cores.  Prints one JSON line.   python tools/gzip_rate.py [--mb 256]"""
import argparse
import json
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
from itsxpress_b200 import _lib, fastq as fq  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=128)
    a = ap.parse_args()
    seq, off, which, cfg = synth.make_config("c2", scale=min(1.0, a.mb / 500.0))
    qual = synth.make_quals(9, off)
    from cli_e2e import write_fastq
    path = "/tmp/gzip_rate.fastq"
    write_fastq(path, seq, off, qual, 1)
    text = np.fromfile(path, np.uint8)
    ctx = _lib.Context(0)
    out = ctx.gzip_compress(text[:1 << 20])
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = ctx.gzip_compress(text)
        times.append(time.perf_counter() - t0)
    import gzip
    head = ctx.gzip_compress(text[:32 << 20])            # (the whole stream is checked in tests/test_gpu_gzip.py)
    assert gzip.decompress(head.tobytes()) == text[:32 << 20].tobytes()
    cores = os.cpu_count() or 1

    def member(chunk):
        co = zlib.compressobj(6, zlib.DEFLATED, 31)
        return co.compress(chunk) + co.flush()
    sample = text[:min(text.size, 32 << 20)].tobytes()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        parts = list(ex.map(member, [sample[i:i + (4 << 20)] for i in range(0, len(sample), 4 << 20)]))
    t_host = time.perf_counter() - t0
    print(json.dumps({"text_bytes": int(text.size), "gpu_gzip_bytes": int(out.size), "gpu_ratio": out.size / text.size,
                      "gpu_seconds_best": min(times), "gpu_GBps_host_to_host": text.size / min(times) / 1e9,
                      "zlib6_ratio": sum(len(p) for p in parts) / len(sample), "zlib6_host_cores": cores,
                      "zlib6_GBps_all_cores": len(sample) / t_host / 1e9,
                      "launches": ctx.launch_count()}))


if __name__ == "__main__":
    main()
