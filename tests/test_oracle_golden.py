"""CPU tests: pin the oracle against every golden vector the reference's own tests hold for the path.

 - derep  : tests/test_data/ex_tmpdir/{seq.fq.gz,uc.txt,rep.fa} are a REAL `vsearch --fastx_uniques` output
            (reference tests/test_main_pytest.py:49-65)
 - trim   : byte goldens t2_r1.fq / t2_r2.fq (reference tests/test_main_pytest.py:378-397), the QIIME2
            single-end output singleOut/.../*.fastq.gz, and the 226-record / 42 637-base counts
            (reference tests/test_main_pytest.py:68-161)
 - HMM    : the STATS LOCAL calibration lines inside the profiles (HMMER-free known-answer test) live in
            test_oracle_calibration.py; real hmmsearch rows are unavailable ("parity unpinned").
"""
import gzip
import os

import numpy as np

from conftest import TD


def _uc(path):
    rows = []
    with open(path) as f:
        for line in f:
            rows.append(line.rstrip("\n").split("\t"))
    return rows


def test_derep_matches_vsearch_uc(oracle, fixture_reads):
    b, seq, off, _ = fixture_reads
    ids = b.ids()
    rep, strand, nu = oracle.derep(seq, off)
    rows = _uc(os.path.join(TD, "ex_tmpdir", "uc.txt"))
    s_rows = [r for r in rows if r[0] == "S"]
    h_rows = [r for r in rows if r[0] == "H"]
    c_rows = [r for r in rows if r[0] == "C"]
    assert b.n == 227 and nu == 137 == len(s_rows) == len(c_rows)
    assert len(h_rows) == 90
    idx = {k: i for i, k in enumerate(ids)}
    # reference test_dedup (tests/test_main_pytest.py:49-65): 227 entries, H -> its S, S -> itself
    match = {}
    for r in s_rows:
        match[r[8]] = r[8]
    for r in h_rows:
        match[r[8]] = r[9]
    assert len(match) == 227
    for rid, target in match.items():
        assert ids[rep[idx[rid]]] == target
    assert match["M02696:28:000000000-ATWK5:1:1101:11740:1800"] == "M02696:28:000000000-ATWK5:1:1101:10899:1561"
    assert strand.sum() == 0          # every H row of the fixture is '+'
    # abundance and vsearch cluster order (abundance desc, label asc)
    ab = np.bincount(rep, minlength=b.n)
    for r in c_rows:
        assert ab[idx[r[8]]] == int(r[2])
    first = np.flatnonzero(rep == np.arange(b.n))
    order = sorted(first.tolist(), key=lambda i: (-ab[i], ids[i]))
    assert [ids[i] for i in order] == [r[8] for r in s_rows]
    # representative = first occurrence in input order
    for i in range(b.n):
        assert rep[i] <= i


def test_rep_fa_matches_fixture(oracle, fixture_reads):
    """rep.fa of the fixture = the representatives' sequences in uc order, 80-column FASTA."""
    from itsxpress_b200.host import write_rep_fasta, cluster_order
    b, seq, off, _ = fixture_reads
    rep, strand, nu = oracle.derep(seq, off)
    ids = b.ids()
    order = cluster_order(rep, ids)
    text = write_rep_fasta(b, order, ids)
    assert text == open(os.path.join(TD, "ex_tmpdir", "rep.fa"), "rb").read()


def test_uc_text_matches_fixture(oracle, fixture_reads):
    from itsxpress_b200.host import write_uc, cluster_order
    b, seq, off, _ = fixture_reads
    rep, strand, nu = oracle.derep(seq, off)
    ids = b.ids()
    text = write_uc(rep, strand, ids, b.s_len, cluster_order(rep, ids))
    assert text == open(os.path.join(TD, "ex_tmpdir", "uc.txt"), "rb").read()


def _golden_table():
    tab = {}
    with open(os.path.join(os.path.dirname(__file__), "golden", "c1_positions.tsv")) as f:
        for line in f:
            if not line.startswith("#"):
                k, a, b, c = line.split("\t")
                tab[k] = (int(a), int(b), int(c))
    return tab


def _positions_by_read(b, rep):
    ids = b.ids()
    tab = _golden_table()
    s = np.full(b.n, -1, np.int32)
    e = s.copy()
    t = s.copy()
    for i in np.flatnonzero(rep == np.arange(b.n)):
        if ids[i] in tab:
            s[i], e[i], t[i] = tab[ids[i]]
    return s, e, t


def test_golden_table_contains_reference_assertion():
    # reference tests/test_main_pytest.py:36-41: left to_pos 128, right from_pos 282 -> (128, 281, 341)
    tab = _golden_table()
    assert tab["M02696:28:000000000-ATWK5:1:1101:19331:3209"] == (128, 281, 341)
    assert "M02696:28:000000000-ATWK5:1:1101:23011:4341" not in tab      # right boundary only -> dropped
    assert len(tab) == 136


def test_trim_single_bytes(oracle, fixture_reads):
    from itsxpress_b200.fastq import format_records
    b, seq, off, qual = fixture_reads
    rep, _, _ = oracle.derep(seq, off)
    s, e, t = _positions_by_read(b, rep)
    keep, lo, hi = oracle.trim_bounds(off, rep, s, e, t, mode=0)
    assert int(keep.sum()) == 226                                   # reference :93
    assert int((hi - lo)[keep == 1].sum()) == 42637                 # reference :98
    ki = np.flatnonzero(keep)
    text = format_records(b, ki, lo[ki], hi[ki])
    gold = os.path.join(TD, "singleOut", "75aea4f5-f10e-421e-91d2-feda9fe7b2e1", "data",
                        "4774-1-MSITS3_0_L001_R1_001.fastq.gz")
    assert text == gzip.open(gold, "rb").read()


def test_trim_paired_bytes(oracle, fixture_reads):
    from itsxpress_b200.fastq import format_records, read_fastq
    b, seq, off, qual = fixture_reads
    rep, _, _ = oracle.derep(seq, off)
    s, e, t = _positions_by_read(b, rep)
    ids = b.ids()
    r1 = read_fastq(os.path.join(TD, "4774-1-MSITS3_R1.fastq"))
    r2 = read_fastq(os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    id1 = {k: i for i, k in enumerate(r1.ids())}
    order = np.array([id1[k] for k in ids])
    for mode, rb, goldf in ((2, r1, "t2_r1.fq"), (1, r2, "t2_r2.fq")):
        s_off = np.zeros(b.n + 1, np.int64)
        s_off[1:] = np.cumsum(rb.s_len[order])
        keep, lo, hi = oracle.trim_bounds(s_off if mode == 2 else off, rep, s, e, t, mode=mode, off_r2=s_off)
        sel = np.flatnonzero(keep)
        assert len(sel) == 226                                        # reference :350-375
        text = format_records(rb, order[sel], lo[sel], hi[sel])
        assert text == open(os.path.join(TD, goldf), "rb").read()     # reference :395-396


def test_trim_edge_cases(oracle):
    """start = 0 is a valid coordinate (reference :412-439); start >= stop drops; None drops;
    python slice clipping when stop exceeds the read."""
    off = np.array([0, 10, 20, 30, 40, 50], np.int64)
    rep = np.arange(5, dtype=np.int32)
    start = np.array([0, 5, -1, 3, 2], np.int32)
    stop = np.array([4, 5, 8, -1, 50], np.int32)
    tlen = np.array([10, 10, 10, 10, 10], np.int32)
    keep, lo, hi = oracle.trim_bounds(off, rep, start, stop, tlen, mode=0)
    assert keep.tolist() == [1, 0, 0, 0, 1]
    assert (lo[0], hi[0]) == (0, 4)
    assert (lo[4], hi[4]) == (2, 10)


def test_derep_both_strands_and_case(oracle):
    reads = [b"ACGTTGCA", b"acgttgca", b"TGCAACGT", b"ACGUTGCA", b"ACGTNGCA", b"TGCNACGT", b"ACGTTGC"]
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    seq = np.frombuffer(b"".join(reads), np.uint8)
    rep, strand, nu = oracle.derep(seq, off)
    # 0,1,3 identical (case, U==T); 2 = revcomp(0); 4 has N; 5 = revcomp(4); 6 shorter
    assert rep.tolist() == [0, 0, 0, 0, 4, 4, 6]
    assert strand.tolist() == [0, 0, 1, 0, 0, 1, 0]
    assert nu == 3


def test_itspos_first_row_wins_ties(oracle):
    """ItsPosition keeps the highest printed score with strict '>' in row order (SeqSample.py:420)."""
    from oracle.oracle import DOM_DTYPE
    rows = np.zeros(4, DOM_DTYPE)
    rows["seq"] = [0, 0, 0, 0]
    rows["prof"] = [0, 1, 2, 3]
    rows["ienv"] = [10, 11, 200, 210]
    rows["jenv"] = [54, 55, 244, 254]
    rows["bitscore"] = [30.04, 29.96, 41.2, 41.26]     # 30.0 vs 30.0 (tie: first wins); 41.2 vs 41.3
    rows["is_reported"] = 1
    side = np.array([0, 0, 1, 1], np.int8)
    pos = oracle.itspos(rows, side, np.array([300], np.int32))
    assert pos["start"][0] == 54 and pos["left_from"][0] == 10
    assert pos["stop"][0] == 209 and pos["right_score10"][0] == 413
    assert pos["tlen"][0] == 300


def test_score10_printf_rounding(oracle):
    for bits in (52.2, 59.1, 34.0, 10.05, 10.15, -3.25, 0.04999, 99.95):
        f = np.float32(bits)
        assert oracle.score10(f) == int(round(float("%.1f" % float(f)) * 10))
