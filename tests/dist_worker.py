"""Worker for the world_size-2 gloo test of itsxpress_b200.distributed.run_sharded (CPU, no GPU):
the orchestration (hash-partitioned derep exchange, domZ all-reduce, position all-gather) is the product's;
the per-rank compute engine here is the CPU ORACLE -- test infrastructure standing in for libitsx_b200."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_COMP = bytes.maketrans(b"ACGTRYMKHDBV", b"TGCAYRKMDHVB")


def canon_key(s):
    s = s.upper().replace(b"U", b"T")
    rc = s.translate(_COMP)[::-1]
    return int.from_bytes(hashlib.blake2b(min(s, rc), digest_size=8).digest(), "little")


class OracleEngine:
    def __init__(self, O, db, side):
        self.O, self.db, self.side = O, db, side

    def derep(self, seq, off):
        rep, strand, nu = self.O.derep(seq, off)
        first = np.flatnonzero(rep == np.arange(len(rep))).astype(np.int32)
        keys = np.array([canon_key(seq[off[i]:off[i + 1]].tobytes()) for i in first], np.uint64)
        return rep, strand, first, keys

    def search_stage1(self, seq, off):
        prm = self.O.default_params()
        prm.domE = 1e300                       # keep every domain of a reported hit; domE is applied in stage 2
        self.seqlen = np.diff(off).astype(np.int32)
        if len(off) > 1:
            self.rows, nrep, _ = self.db.search(self.O.digitize(seq.tobytes()), off, prm)
        else:
            self.rows, nrep = np.zeros(0, self.O.DOM_DTYPE), np.zeros(self.db.n, np.int32)
        return nrep.astype(np.int64)

    def search_stage2(self, nrep_global, nseq):
        r = self.rows
        ok = np.exp(r["lnP"]) * nrep_global[r["prof"]] <= 10.0
        pos = self.O.itspos(r[ok], self.side, self.seqlen)
        return pos["start"], pos["stop"], pos["tlen"]

    def trim_bounds(self, uid, n_unique, start, stop, tlen, off, mode=0):
        return self.O.trim_bounds(off, uid, start, stop, tlen, mode=mode)


def dataset():
    from itsxpress_b200.fastq import read_fastq
    b = read_fastq(os.path.join(ROOT, "tests", "test_data", "ex_tmpdir", "seq.fq.gz"))
    seq, off = b.seq_concat()
    reads = [seq[off[i]:off[i + 1]].tobytes() for i in range(b.n)]
    # cross-block duplicates on both strands + case variants, so that classes span ranks
    extra = [reads[5].translate(_COMP)[::-1], reads[200].lower(), reads[3], reads[150].translate(_COMP)[::-1]]
    reads = reads[:120] + extra[:2] + reads[120:] + extra[2:] + [reads[0]]
    o = np.zeros(len(reads) + 1, np.int64)
    o[1:] = np.cumsum([len(r) for r in reads])
    return np.frombuffer(b"".join(reads), np.uint8).copy(), o


def engine():
    from oracle import oracle as O
    O.lib()
    hmm = os.path.join(ROOT, "itsxpress_b200", "ITSx_db", "HMMs", "M.hmm")
    db = O.ProfileDB([hmm], ["3_", "4_"])
    side = np.array([0 if n.startswith("3_") else 1 for n in db.names], np.int8)
    return O, OracleEngine(O, db, side)


def main():
    import torch.distributed as dist
    from itsxpress_b200.distributed import Comm, block_range, run_sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq, off = dataset()
    n = len(off) - 1
    lo, hi = block_range(n, rank, world)
    O, eng = engine()
    out = run_sharded(eng, Comm(), seq[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], lo)
    np.savez(os.path.join(sys.argv[1], "rank%d.npz" % rank), blo=lo, bhi=hi, **{k: np.asarray(v) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
