"""Golden vectors for the paired-end merge: an independent, definition-level pure-Python statement
of `vsearch --fastq_mergepairs --fastq_maxdiffs 40 --fastq_maxee 2 [--fastq_allowmergestagger] --fastq_qmax 93`
(SeqSample.py:314-349), run on the reference's own paired fixtures, written to tests/golden/c1_merge.tsv:

    file  stagger  pair  reason  merged_len  crc32(bases)  crc32(qualities)

It shares no code with oracle/ora_merge.c (byte loops over every diagonal, k-mers counted by their definition) and
exists so that the C oracle and the CUDA kernel are checked against something that is not themselves.  vsearch
itself is not installed anywhere this project builds, so these are vectors of the RESTATED algorithm ("parity
unpinned" against a real vsearch run; see oracle/ora_merge.c for what the reference's fixtures do pin).

Usage (repo root):  python tests/golden/make_merge_golden.py
"""
import gzip
import math
import os
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
TD = os.path.join(HERE, "..", "test_data")
REASONS = ("ok", "repeat", "staggered", "maxdiffs", "maxdiffpct", "nokmers", "minscore", "minovlen", "maxee")


def q_to_p(x):
    return 0.75 if x < 2 else 10.0 ** (-x / 10.0)


def q_from_p(p):
    return 33 + max(0, min(41, int(round(-10.0 * math.log10(p)))))


P = [q_to_p(x) for x in range(94)]
SAME = [[q_from_p(px * py / 3.0 / (1.0 - px - py + 4.0 * px * py / 3.0)) for py in P] for px in P]
DIFF = [[q_from_p(px * (1.0 - py / 3.0) / (px + py - 4.0 * px * py / 3.0)) for py in P] for px in P]
MATCH = [[math.log2((1.0 - px - py + px * py * 4.0 / 3.0) / 0.25) for py in P] for px in P]
MISM = [[math.log2(((px + py) / 3.0 - px * py * 4.0 / 9.0) / 0.25) for py in P] for px in P]
COMP = {ord(a): ord(b) for a, b in zip("ACGTURYSWKMBDHVN", "TGCAAYRSWMKVHDBN")}


def read_fq(path):
    op = gzip.open if path.endswith(".gz") else open
    out = []
    with op(path, "rb") as f:
        while True:
            h = f.readline().rstrip(b"\r\n")
            if not h:
                break
            s = f.readline().rstrip(b"\r\n")
            f.readline()
            q = f.readline().rstrip(b"\r\n")
            out.append((h[1:], s, q))
    return out


def merge_pair(fs, fq, rs, rq, stagger=False, maxdiffs=40, maxee=2.0, minovlen=10):
    fs = fs.upper()
    F, R = len(fs), len(rs)
    rc = bytes(COMP.get(c, ord("N")) for c in rs.upper()[::-1])
    rcq = rq[::-1]
    best_score, best_i, best_diffs, hits, kmers = 0.0, 0, 0, 0, 0
    for i in range(1, F + R):
        sh = F - i
        p0, p1 = max(0, sh), min(F, sh + R)
        cnt = 0
        for p in range(p0, p1 - 4):                      # 5-mers starting at p, by definition
            a, b = fs[p:p + 5], rc[p - sh:p - sh + 5]
            if a == b and all(c in b"ACGTU" for c in a):
                cnt += 1
        if cnt < 4:
            continue
        kmers = 1
        score = high = dropmax = 0.0
        diffs = 0
        for p in range(p1 - 1, p0 - 1, -1):
            qa, qb = fq[p] - 33, rcq[p - sh] - 33
            if fs[p] == rc[p - sh]:
                score += MATCH[qa][qb]
                high = max(high, score)
            else:
                score += MISM[qa][qb]
                diffs += 1
                if score < high - dropmax:
                    dropmax = high - score
        if dropmax >= 16.0:
            score = 0.0
        if score >= 16.0:
            hits += 1
        if score > best_score:
            best_score, best_i, best_diffs = score, i, diffs
    if hits > 1:
        return None, "repeat"
    if not stagger and best_i > F:
        return None, "staggered"
    if best_diffs > maxdiffs:
        return None, "maxdiffs"
    if best_i and 100.0 * best_diffs / best_i > 100.0:
        return None, "maxdiffpct"
    if not kmers:
        return None, "nokmers"
    if best_score < 16.0:
        return None, "minscore"
    if best_i < minovlen:
        return None, "minovlen"
    sh = F - best_i
    ms, mq, ee = bytearray(), bytearray(), 0.0
    for m in range(max(sh, 0) + min(F - max(sh, 0), R - max(-sh, 0)) + (R - max(-sh, 0) - min(F - max(sh, 0), R - max(-sh, 0)))):
        r = m - sh
        if m < sh:
            s, q = fs[m], fq[m]
        elif m < F and r < R:
            a, b, qa, qb = fs[m], rc[r], fq[m] - 33, rcq[r] - 33
            if b == ord("N"):
                s, q = a, fq[m]
            elif a == ord("N"):
                s, q = b, rcq[r]
            elif a == b:
                s, q = a, SAME[qa][qb]
            elif qa > qb:
                s, q = a, DIFF[qa][qb]
            else:
                s, q = b, DIFF[qb][qa]
        else:
            s, q = rc[r], rcq[r]
        ms.append(s)
        mq.append(q)
        ee += P[q - 33]
    if ee <= maxee:
        return (bytes(ms), bytes(mq)), "ok"
    return None, "maxee"


def main():
    rows = []
    for name, r1, r2 in (("4774-1-MSITS3", "4774-1-MSITS3_R1.fastq", "4774-1-MSITS3_R2.fastq"),
                         ("high_qual_scores", "high_qual_scores_R1.fastq.gz", "high_qual_scores_R2.fastq.gz")):
        a, b = read_fq(os.path.join(TD, r1)), read_fq(os.path.join(TD, r2))
        assert len(a) == len(b)
        for stagger in (0, 1):
            for i, ((_, s1, q1), (_, s2, q2)) in enumerate(zip(a, b)):
                res, why = merge_pair(s1, q1, s2, q2, stagger=bool(stagger))
                if res:
                    rows.append((name, stagger, i, why, len(res[0]), zlib.crc32(res[0]), zlib.crc32(res[1])))
                else:
                    rows.append((name, stagger, i, why, 0, 0, 0))
    with open(os.path.join(HERE, "c1_merge.tsv"), "w") as f:
        f.write("# file\tstagger\tpair\treason\tmerged_len\tcrc32_bases\tcrc32_quals\n")
        for r in rows:
            f.write("\t".join(str(x) for x in r) + "\n")
    print(len(rows), "rows")


if __name__ == "__main__":
    main()
