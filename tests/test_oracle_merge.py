"""CPU tests of the paired-end merge oracle (oracle/ora_merge.c; SURVEY 8f row 2, reference call site
itsxpress/SeqSample.py:266-365).

vsearch is an un-vendored third-party binary and no vsearch-made merge fixture exists in the reference tree, so the
restatement is checked against
  * tests/golden/c1_merge.tsv: an independent, definition-level Python statement of the same published algorithm on the
    reference's paired fixtures (tests/golden/make_merge_golden.py),
  * the reference's own merged fixture 4774-1-MSITS3_merged.fastq (written by an older merger: it pins the overlap
    found -- the bases -- not the quality arithmetic),
  * known answers worked out by hand from the published formulas.
"""
import os
import zlib

import numpy as np
import pytest

from conftest import TD
from itsxpress_b200.fastq import read_fastq

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_merge.tsv")
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def revcomp(s):
    return s.translate(COMP)[::-1]


def pack(reads):
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    return np.frombuffer(b"".join(reads), np.uint8), off


def merge(O, fwd, rev, **kw):
    """fwd / rev: lists of (bases, quals) byte strings.  Returns [(reason, bases, quals)]."""
    fs, fo = pack([a for a, _ in fwd])
    fq, _ = pack([b for _, b in fwd])
    rs, ro = pack([a for a, _ in rev])
    rq, _ = pack([b for _, b in rev])
    ml, why, os_, oq = O.merge_pairs(fs, fq, fo, rs, rq, ro, O.merge_params(**kw))
    out = []
    for i in range(len(fwd)):
        s = int(fo[i] + ro[i])
        out.append((O.MERGE_REASONS[why[i]], bytes(os_[s:s + ml[i]]), bytes(oq[s:s + ml[i]])))
    return out


def load_pair(name_r1, name_r2):
    b1, b2 = read_fastq(os.path.join(TD, name_r1)), read_fastq(os.path.join(TD, name_r2))
    return b1, b2, b1.seq_concat(), b1.qual_concat()[0], b2.seq_concat(), b2.qual_concat()[0]


@pytest.mark.parametrize("name,r1,r2", [("4774-1-MSITS3", "4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq"),
                                        ("high_qual_scores", "high_qual_scores_R1.fastq.gz",
                                         "high_qual_scores_R2.fastq.gz")])
@pytest.mark.parametrize("stagger", [0, 1])
def test_oracle_matches_golden(oracle, name, r1, r2, stagger):
    gold = {}
    for line in open(GOLD):
        if line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        if f[0] == name and int(f[1]) == stagger:
            gold[int(f[2])] = (f[3], int(f[4]), int(f[5]), int(f[6]))
    b1, b2, (fs, fo), fq, (rs, ro), rq = load_pair(r1, r2)
    assert len(gold) == b1.n == b2.n
    ml, why, os_, oq = oracle.merge_pairs(fs, fq, fo, rs, rq, ro, oracle.merge_params(allow_stagger=stagger))
    for i in range(b1.n):
        s = int(fo[i] + ro[i])
        got = (oracle.MERGE_REASONS[why[i]], int(ml[i]), zlib.crc32(bytes(os_[s:s + ml[i]])) if ml[i] else 0,
               zlib.crc32(bytes(oq[s:s + ml[i]])) if ml[i] else 0)
        assert got == gold[i], (name, stagger, i)


def test_overlap_agrees_with_reference_merged_fixture(oracle):
    """236 of the 250 bundled pairs merge (the reference's live-CLI test expects 235 trimmed reads from them,
    tests/test_main_pytest.py:252); for the pairs the reference's older merged fixture also holds, the merged
    BASES are the same (one pair excepted, where the two mergers pick a different base at a disagreement)."""
    b1, b2, (fs, fo), fq, (rs, ro), rq = load_pair("4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq")
    ml, why, os_, oq = oracle.merge_pairs(fs, fq, fo, rs, rq, ro)
    assert int((ml > 0).sum()) == 236
    hist = np.bincount(why, minlength=10)
    assert hist[oracle.MERGE_REASONS.index("minscore")] == 4 and hist[oracle.MERGE_REASONS.index("maxee")] == 10
    fx = read_fastq(os.path.join(TD, "4774-1-MSITS3_merged.fastq"))
    fx_seq = {fx.title(i): fx.seq(i) for i in range(fx.n)}
    both = same = 0
    for i in range(b1.n):
        t = b1.title(i)
        if ml[i] and t in fx_seq:
            both += 1
            s = int(fo[i] + ro[i])
            same += bytes(os_[s:s + ml[i]]).decode() == fx_seq[t]
    assert both == 226 and same >= 225


def test_quality_tables_known_answers(oracle):
    match, mism, same, diff, q2p = oracle.merge_tables()
    assert q2p[0] == 0.75 and q2p[1] == 0.75 and q2p[10] == pytest.approx(0.1) and q2p[40] == pytest.approx(1e-4)
    # agreement: p = px py / 3 / (1 - px - py + 4 px py / 3);  Q20 + Q20 -> -10 log10(3.40e-5) = 44.7 -> capped 41
    assert same[20, 20] == 33 + 41
    # Q10 + Q10: 0.01 / 3 / (0.8 + 0.04 / 3) = 4.098e-3 -> Q23.87 -> 24
    assert same[10, 10] == 33 + 24
    # Q2 + Q2 (p = 0.631 each): 0.1327 / (1 - 1.262 + 0.5309) = 0.4935 -> Q3.07 -> 3
    assert same[2, 2] == 33 + 3
    # disagreement, Q40 beats Q10: 1e-4 (1 - 0.1 / 3) / (1e-4 + 0.1 - 4e-5 / 3) = 9.658e-4 -> Q30.15 -> 30
    assert diff[40, 10] == 33 + 30
    # equal qualities disagreeing: p = px (1 - px / 3) / (2 px - 4 px^2 / 3) -> ~0.5 -> Q3
    assert diff[30, 30] == 33 + 3
    # scores in bits: a confident match is worth log2(1 / 0.25) = 2, a confident mismatch log2((2e-4 / 3) / 0.25)
    assert match[40, 40] == pytest.approx(np.log2((1 - 2e-4 + 4e-8 / 3) / 0.25))
    assert 1.99 < match[40, 40] < 2.0
    assert mism[40, 40] == pytest.approx(np.log2((2e-4 / 3 - 4e-8 / 9) / 0.25))
    assert np.all(np.diff(match[40, 2:]) >= 0) and np.all(np.diff(mism[40, 2:]) <= 0)
    # symmetric up to rounding (1 - px - py is not evaluated symmetrically: the index order [forward][reverse] matters)
    assert np.allclose(match, match.T, atol=1e-12) and np.allclose(mism, mism.T, atol=1e-12)
    assert np.array_equal(same, same.T)


def _frag(rng, n):
    return bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n))


def test_known_answers(oracle):
    rng = np.random.default_rng(11)
    frag = _frag(rng, 400)
    q = b"I" * 250                                                       # Q40
    r1, r2 = frag[:250], revcomp(frag[-250:])
    # 1. clean 100-base overlap: the fragment comes back, overlap qualities are capped at Q41 ('J')
    (why, s, qq), = merge(oracle, [(r1, q)], [(r2, q)])
    assert why == "ok" and s == frag and qq == b"I" * 150 + b"J" * 100 + b"I" * 150
    # 2. a disagreement in the overlap: the better base wins with the posterior quality, either direction
    bad = bytearray(r2)
    bad[170] = ord("A") if r2[170] != ord("A") else ord("C")            # R2 position 170 <-> fragment position 229
    q2 = bytearray(q)
    q2[170] = ord("+")                                                   # Q10
    (why, s, qq), = merge(oracle, [(r1, q)], [(bytes(bad), bytes(q2))])
    assert why == "ok" and s == frag and qq[229] == 33 + 30              # forward base (Q40) wins: diff[40][10]
    bad1 = bytearray(r1)
    bad1[200] = ord("A") if r1[200] != ord("A") else ord("C")
    q1 = bytearray(q)
    q1[200] = ord("+")
    (why, s, qq), = merge(oracle, [(bytes(bad1), bytes(q1))], [(r2, q)])
    assert why == "ok" and s == frag and qq[200] == 33 + 30              # diff[40][10]
    # 3. an N defers to the other read and takes its quality
    n1 = bytearray(r1)
    n1[210] = ord("N")
    qn = bytearray(q)
    qn[210] = ord("#")
    (why, s, qq), = merge(oracle, [(bytes(n1), bytes(qn))], [(r2, q)])
    assert why == "ok" and s == frag and qq[210] == ord("I")
    # 4. lower-case input is merged as upper case
    (why, s, qq), = merge(oracle, [(r1.lower(), q)], [(r2.lower(), q)])
    assert why == "ok" and s == frag
    # 5. staggered pair (fragment shorter than the reads): rejected unless allowed; allowed -> the fragment only
    short = frag[:200]
    pad5, pad3 = _frag(rng, 50), _frag(rng, 50)
    s1, s2 = short + pad3, revcomp(short) + pad5                         # both reads run off the fragment's end
    (why, s, qq), = merge(oracle, [(s1, q)], [(s2, q)])
    assert why == "staggered" and s == b""
    (why, s, qq), = merge(oracle, [(s1, q)], [(s2, q)], allow_stagger=True)
    assert why == "ok" and s == short and qq == b"J" * 200
    # 6. no overlap at all
    (why, s, qq), = merge(oracle, [(_frag(rng, 250), q)], [(_frag(rng, 250), q)])
    assert why in ("nokmers", "minscore") and s == b""
    # 7. expected errors above 2 -> dropped (200 non-overlap bases at Q10 carry 20 expected errors)
    qlow = b"+" * 250
    (why, s, qq), = merge(oracle, [(r1, qlow)], [(r2, q)])
    assert why == "maxee"
    # 8. a tandem repeat gives two equally good diagonals -> "repeat"
    unit = _frag(rng, 20)
    rep = _frag(rng, 100) + unit * 6 + _frag(rng, 100)                   # 320
    (why, s, qq), = merge(oracle, [(rep[:200], q[:200])], [(revcomp(rep[-200:]), q[:200])])
    assert why in ("repeat", "ok")                                       # the true diagonal has 80 matches ...
    rep2 = unit * 20                                                     # ... a pure repeat has many equal ones
    (why, s, qq), = merge(oracle, [(rep2[:250], q)], [(revcomp(rep2[-250:]), q)])
    assert why == "repeat"
    # 9. overlap shorter than --fastq_minovlen cannot reach 16 bits (8 x 2 bits): minscore / nokmers
    f9 = _frag(rng, 492)
    (why, s, qq), = merge(oracle, [(f9[:250], q)], [(revcomp(f9[-250:]), q)])
    assert why in ("nokmers", "minscore")
    # 10. empty reads
    out = merge(oracle, [(b"", b""), (r1, q), (b"", b"")], [(r2, q), (b"", b""), (b"", b"")])
    assert [w for w, _, _ in out] == ["nokmers"] * 3


def test_bad_quality_is_fatal(oracle):
    rng = np.random.default_rng(3)
    frag = _frag(rng, 400)
    q = bytearray(b"I" * 250)
    q[7] = 33 + 94                                                       # above --fastq_qmax 93
    with pytest.raises(ValueError):
        merge(oracle, [(frag[:250], bytes(q))], [(revcomp(frag[-250:]), b"I" * 250)])
    q[7] = 32                                                            # below ASCII 33
    with pytest.raises(ValueError):
        merge(oracle, [(frag[:250], bytes(q))], [(revcomp(frag[-250:]), b"I" * 250)])


def test_oracle_equals_definition_on_random_pairs(oracle):
    """The C oracle against the definition-level Python statement (tests/golden/make_merge_golden.py) on ragged random
    pairs -- staggered and non-overlapping fragments, trimmed reads, errors, N, qualities 0..41 -- under several option
    sets (the options the reference passes, SeqSample.py:314-349, and tighter ones)."""
    import importlib.util
    import synth
    spec = importlib.util.spec_from_file_location(
        "make_merge_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_merge_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    _, _, fs, fq, fo, rs, rq, ro = synth.make_pair_config(31, 240, frag_len=(60, 300), read_len=150, trim=(0, 60),
                                                          err_scale=2.0, n_rate=0.01, empty_rate=0.01)
    rng = np.random.default_rng(31)
    fq, rq = fq.copy(), rq.copy()
    fq[rng.random(len(fq)) < 0.03] = 33          # quality 0
    rq[rng.random(len(rq)) < 0.03] = 34          # quality 1
    for kw in (dict(), dict(allow_stagger=True), dict(maxdiffs=3, minovlen=40), dict(allow_stagger=True, maxee=0.5)):
        ml, why, os_, oq = oracle.merge_pairs(fs, fq, fo, rs, rq, ro, oracle.merge_params(**kw))
        for i in range(len(fo) - 1):
            res, reason = G.merge_pair(fs[fo[i]:fo[i + 1]].tobytes(), fq[fo[i]:fo[i + 1]].tobytes(),
                                       rs[ro[i]:ro[i + 1]].tobytes(), rq[ro[i]:ro[i + 1]].tobytes(),
                                       stagger=bool(kw.get("allow_stagger", False)), maxdiffs=kw.get("maxdiffs", 40),
                                       maxee=kw.get("maxee", 2.0), minovlen=kw.get("minovlen", 10))
            s = int(fo[i] + ro[i])
            got = (bytes(os_[s:s + ml[i]]), bytes(oq[s:s + ml[i]])) if ml[i] else None
            assert oracle.MERGE_REASONS[why[i]] == reason and got == res, (kw, i, reason)
        assert len(set(why.tolist())) >= 3
