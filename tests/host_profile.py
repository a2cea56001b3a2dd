#!/usr/bin/env python
"""Where the HOST time of a run goes, measured without a GPU: the command line (or the QIIME 2 per-sample loop) runs with
tests/oracle_context.py standing in for the device and with the expensive stand-in stages (profile search, pair merge)
replaced by constant answers, under cProfile.  What is left is the product's host code: FASTQ inflate / scan / gather,
the inter-stage files, formatting, compression (ITSX_GZIP_LEVEL=0 takes zlib out of the picture, as the GPU writer does on
the box), file output.  This is how the Python row loop of domtbl.txt (25 us per row) was found.

  python tests/host_profile.py cli  --reads 200000 [--gz]
  python tests/host_profile.py q2   --samples 6 --pairs 50000

(Lives under tests/ because it executes the oracle: only tests, smoke() and the bench's CPU legs may.)
"""
import argparse
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
os.environ.setdefault("ITSX_GZIP", "host")

import numpy as np  # noqa: E402
import synth  # noqa: E402
from cli_e2e import gzip_one_member, write_fastq  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle_context import OracleContext  # noqa: E402
from itsxpress_b200 import SeqSample, fastq as fq, main as cli, q2_itsxpress as q2  # noqa: E402

POS = ["start", "stop", "tlen", "left_score10", "left_from", "left_to", "right_score10", "right_from", "right_to"]


class Fake(OracleContext):
    """constant positions instead of a search, R1 instead of a merged read; the time of the other stand-in stages is kept"""
    standin = 0.0

    def _run_search(self, seq, off, params):
        self._seqlen = np.diff(off).astype(np.int32)
        self._pos = {k: np.zeros(len(self._seqlen), np.int32) for k in POS}
        self._pos["start"][:], self._pos["stop"][:], self._pos["tlen"][:] = 30, 200, self._seqlen
        self._orows, self._nrep = np.zeros(0, O.DOM_DTYPE), np.zeros(len(self.names), np.int32)

    def search_stage2(self):
        pass

    def merge_pairs(self, fseq, fqual, foff, rseq, rqual, roff, params=None, fetch=True):
        n = len(foff) - 1
        self._merge = (n, n, np.bincount(np.zeros(1, np.int64), minlength=16) * n)
        return (np.diff(foff).astype(np.int32), np.zeros(n, np.uint8), np.arange(n, dtype=np.int32),
                np.array(foff, np.int64), np.array(fseq), np.array(fqual))

    def _timed(self, f, *a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        Fake.standin += time.perf_counter() - t0
        return r

    def derep(self, *a):
        return self._timed(super().derep, *a)

    def trim_gather(self, *a, **k):
        return self._timed(super().trim_gather, *a, **k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["cli", "q2"])
    ap.add_argument("--reads", type=int, default=200000)
    ap.add_argument("--samples", type=int, default=6)
    ap.add_argument("--pairs", type=int, default=50000)
    ap.add_argument("--gz", action="store_true", help="cli: .gz input and output")
    ap.add_argument("--zst", action="store_true", help="cli: .zst input and output")
    ap.add_argument("--batched", action="store_true", help="q2: small samples share device passes (the default of the driver)")
    ap.add_argument("--top", type=int, default=14)
    a = ap.parse_args()
    O.lib()
    ctx = Fake(O)
    SeqSample.get_context = lambda: ctx
    tmp = tempfile.mkdtemp(prefix="itsx_hostprof_")
    if a.mode == "cli":
        seq, off, _, _ = synth.make_config("c2", seed=5, scale=a.reads / 1e6)
        src = os.path.join(tmp, "in.fastq")
        write_fastq(src, seq, off, synth.make_quals(77, off))
        if a.gz:
            with open(src, "rb") as f, open(src + ".gz", "wb") as g:
                g.write(gzip_one_member(f.read()))
            src += ".gz"
        if a.zst:
            from itsxpress_b200 import _zstd
            with open(src, "rb") as f, open(src + ".zst", "wb") as g:
                g.write(_zstd.compress(f.read()))
            src += ".zst"
        argv = ["--fastq", src, "--single_end", "--outfile", os.path.join(tmp, "out.fastq" + (".gz" if a.gz else ".zst" if a.zst else "")), "--region", "ITS1",
                "--taxa", "Metazoa", "--log", os.path.join(tmp, "log.txt"), "--tempdir", tmp]
        units, what = len(off) - 1, "reads"

        def run():
            SeqSample.reset_sessions()
            cli.main(args=cli.myparser().parse_args(argv))
    else:
        art = os.path.join(tmp, "in")
        os.makedirs(art)
        lines = ["sample-id,filename,direction"]
        for k in range(a.samples):
            frag, foff, _, _ = synth.make_reads(5 * 1_000_003 + k, a.pairs, max(300, a.pairs // 20), (330, 441), "M.hmm", "3_", "4_",
                                                zipf_s=1.2, spacer=(150, 230))
            fs, fq_, fo, rs, rq, ro = synth.make_pairs(7 + k, frag, foff, read_len=250, err_scale=0.0, n_rate=0.0)
            for tag, d, (x, y, z) in (("R1", "forward", (fs, fo, fq_)), ("R2", "reverse", (rs, ro, rq))):
                fn = "S%d_%d_L001_%s_001.fastq.gz" % (k, k, tag)
                p = os.path.join(tmp, "x.fastq")
                write_fastq(p, x, y, z, 1 if tag == "R1" else 2)
                with open(p, "rb") as f, open(os.path.join(art, fn), "wb") as g:
                    g.write(gzip_one_member(f.read()))
                lines.append("S%d,%s,%s" % (k, fn, d))
        with open(os.path.join(art, "MANIFEST"), "w") as f:
            f.write("\n".join(lines) + "\n")
        if not a.batched:
            q2.BATCH_READS = 0
        units, what = a.samples * a.pairs, "pairs"

        def run():
            SeqSample.reset_sessions()
            q2.trim_pair_output_unmerged(q2.PerSampleDir(art), region="ITS2", taxa="M")
    run()
    Fake.standin = 0.0
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    print("%d %s: wall %.2f s, stand-in compute %.2f s -> host %.2f s (%.2f us per %s); host cores %d" %
          (units, what, dt, Fake.standin, dt - Fake.standin, (dt - Fake.standin) / units * 1e6, what[:-1], fq.host_share()))
    prof = os.path.join(tmp, "prof.out")
    cProfile.runctx("run()", {"run": run}, {}, prof)
    pstats.Stats(prof).sort_stats("tottime").print_stats(a.top)


if __name__ == "__main__":
    main()
