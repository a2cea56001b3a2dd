"""Worker for the world_size-2 gloo test of itsxpress_b200.distributed.run_sharded (CPU, no GPU):
the orchestration (hash-partitioned derep exchange, domZ all-reduce, answers through the inverse exchange) is the
product's;
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
    """numpy stand-in for itsxpress_b200.distributed.GpuEngine (same methods, same buffer contents)."""

    def __init__(self, O, db, side):
        self.O, self.db, self.side = O, db, side

    # -- block side --
    def local_derep(self, seq, off):
        self.seq, self.off = np.ascontiguousarray(seq, np.uint8), np.ascontiguousarray(off, np.int64)
        self.rep_l, self.strand_l, nu = self.O.derep(self.seq, self.off)
        self.first = np.flatnonzero(self.rep_l == np.arange(len(self.rep_l))).astype(np.int64)
        self.uid = np.searchsorted(self.first, self.rep_l).astype(np.int64)
        self.keys = np.array([canon_key(self.seq[self.off[i]:self.off[i + 1]].tobytes()) for i in self.first], np.uint64)
        return len(self.first)

    def plan(self, G):
        owner = (self.keys % np.uint64(G)).astype(np.int64) if len(self.keys) else np.zeros(0, np.int64)
        self.order = np.argsort(owner, kind="stable")
        lens = (self.off[1:] - self.off[:-1])[self.first]
        bc = np.zeros(G, np.int64)
        np.add.at(bc, owner, lens)
        return np.bincount(owner, minlength=G).astype(np.int64), bc

    def pack(self, gidx0, nbytes):
        r = self.first[self.order]
        lens = (self.off[1:] - self.off[:-1])[r]
        rec = ((r + gidx0).astype(np.uint64) | (lens.astype(np.uint64) << np.uint64(32))).view(np.int64)
        bases = np.concatenate([self.seq[self.off[i]:self.off[i + 1]] for i in r]) if len(r) else np.zeros(0, np.uint8)
        assert len(bases) == nbytes
        return rec, bases

    # -- owner side --
    def owner_derep(self, rec, bases):
        rec = rec.view(np.uint64)
        self.gidx = (rec & np.uint64(0xffffffff)).astype(np.int64)
        assert np.all(np.diff(self.gidx) > 0)        # arrival order = ascending global read index
        lens = (rec >> np.uint64(32)).astype(np.int64)
        self.o_off = np.zeros(len(rec) + 1, np.int64)
        np.cumsum(lens, out=self.o_off[1:])
        self.o_seq = np.ascontiguousarray(bases, np.uint8)
        self.o_rep, self.o_strand, nu = self.O.derep(self.o_seq, self.o_off)
        self.o_first = np.flatnonzero(self.o_rep == np.arange(len(self.o_rep))).astype(np.int64)
        self.o_uid = np.searchsorted(self.o_first, self.o_rep)
        return len(self.o_first)

    def search_stage1(self):
        prm = self.O.default_params()
        prm.domE = 1e300                       # keep every domain of a reported hit; domE is applied in stage 2
        parts = [self.o_seq[self.o_off[i]:self.o_off[i + 1]] for i in self.o_first]
        uoff = np.zeros(len(parts) + 1, np.int64)
        uoff[1:] = np.cumsum([len(p) for p in parts])
        self.seqlen = np.diff(uoff).astype(np.int32)
        if len(parts):
            self.rows, nrep, _ = self.db.search(self.O.digitize(np.concatenate(parts).tobytes()), uoff, prm)
        else:
            self.rows, nrep = np.zeros(0, self.O.DOM_DTYPE), np.zeros(self.db.n, np.int32)
        return nrep.astype(np.int64)

    def search_stage2(self, nrep_global):
        r = self.rows
        ok = np.exp(r["lnP"]) * nrep_global[r["prof"]] <= 10.0
        self.pos = self.O.itspos(r[ok], self.side, self.seqlen)

    def answers(self):
        m = len(self.o_rep)
        ans = np.full((m, 4), -1, np.int32)
        if m:
            g = self.gidx[self.o_rep].astype(np.uint32) | (self.o_strand.astype(np.uint32) << np.uint32(31))
            ans[:, 0] = g.view(np.int32)
            for k, name in ((1, "start"), (2, "stop"), (3, "tlen")):
                ans[:, k] = self.pos[name][self.o_uid]
        return ans

    # -- block side again --
    def apply(self, ans, want_rep=True):
        nu = len(self.first)
        tab = np.full((nu, 4), -1, np.int32)
        tab[self.order] = ans
        self.tab = tab
        g = tab[:, 0].view(np.uint32)[self.uid] if len(self.uid) else np.zeros(0, np.uint32)
        rep = (g & np.uint32(0x7fffffff)).astype(np.int64)
        strand = ((g >> np.uint32(31)).astype(np.uint8) ^ self.strand_l.astype(np.uint8)) & 1
        return rep, strand

    def trim_bounds(self, mode=0):
        t = self.tab
        return self.O.trim_bounds(self.off, self.uid.astype(np.int32), t[:, 1].copy(), t[:, 2].copy(), t[:, 3].copy(),
                                  mode=mode)


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
