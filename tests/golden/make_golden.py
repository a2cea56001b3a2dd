"""Recover the (start, stop, tlen) table that the reference's missing domtbl.txt fixture implied.

Inputs are the reference's own fixtures (copied under tests/test_data/):
  4774-1-MSITS3_R1.fastq / _R2.fastq  paired input
  ex_tmpdir/seq.fq.gz                 merged reads (vsearch output, 227 reads)
  ex_tmpdir/uc.txt                    vsearch derep map
  t2_r1.fq / t2_r2.fq                 byte goldens of create_paired_trimmed_seqs
                                      (reference tests/test_main_pytest.py:378-397)
For a kept read: R1 slice = [start:stop] and R2 slice = [tlen-stop : tlen-start]
(SeqSample.py:639-655), so start = offset of the t2_r1 record inside its R1 read and
stop = tlen - offset of the t2_r2 record inside its R2 read; tlen = merged length.
Writes tests/golden/c1_positions.tsv (rep_id, start, stop, tlen), one row per representative.
Run: python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from itsxpress_b200.fastq import read_fastq  # noqa: E402

TD = os.path.join(ROOT, "tests", "test_data")


def load(path):
    b = read_fastq(path)
    return {i: (b.seq(k), b.qual(k)) for k, i in enumerate(b.ids())}, b


def main():
    r1, _ = load(os.path.join(TD, "4774-1-MSITS3_R1.fastq"))
    r2, _ = load(os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    mg, _ = load(os.path.join(TD, "ex_tmpdir", "seq.fq.gz"))
    t1, b1 = load(os.path.join(TD, "t2_r1.fq"))
    t2, _ = load(os.path.join(TD, "t2_r2.fq"))
    rep = {}
    with open(os.path.join(TD, "ex_tmpdir", "uc.txt")) as f:
        for line in f:
            ll = line.split()
            if ll[0] == "S":
                rep[ll[8]] = ll[8]
            elif ll[0] == "H":
                rep[ll[8]] = ll[9]
    table = {}
    for rid in b1.ids():
        s1, q1 = t1[rid]
        s2, q2 = t2[rid]
        R1s, R1q = r1[rid]
        R2s, R2q = r2[rid]
        tlen = len(mg[rid][0])
        a = R1s.find(s1)
        assert a >= 0 and R1s.find(s1, a + 1) < 0 and R1q[a:a + len(s1)] == q1, rid
        b = R2s.find(s2)
        assert b >= 0 and R2s.find(s2, b + 1) < 0 and R2q[b:b + len(s2)] == q2, rid
        start, stop = a, tlen - b
        # consistency of the two slices with one (start, stop, tlen)
        assert R1s[start:stop] == s1 and R2s[tlen - stop:tlen - start] == s2, rid
        key = rep[rid]
        if key in table:
            assert table[key] == (start, stop, tlen), (rid, key)
        table[key] = (start, stop, tlen)
    out = os.path.join(HERE, "c1_positions.tsv")
    with open(out, "w") as f:
        f.write("#rep_id\tstart\tstop\ttlen\n")
        for k in sorted(table):
            f.write("%s\t%d\t%d\t%d\n" % ((k,) + table[k]))
    nb = sum(v[1] - v[0] for k, v in table.items() for r in rep if rep[r] == k and r in t1)
    print("representatives:", len(table), "reads:", len(t1), "bases:", nb)


if __name__ == "__main__":
    main()
