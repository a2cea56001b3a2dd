/*
 * ora_merge.c -- CPU oracle: restatement of
 *     vsearch --fastq_mergepairs R1 --reverse R2 --fastqout seq.fq --fastq_maxdiffs 40 --fastq_maxee 2
 *             [--fastq_allowmergestagger] --fastq_qmax 93
 * as the reference invokes it (itsxpress/SeqSample.py:266-365, argv :314-349; constants
 * itsxpress/definitions.py:79,82).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * vsearch (>= 2.21.1, recipes/itsxpress/meta.yaml:37; verified upstream with 2.22.1) is an un-vendored third-party
 * binary that is neither in /root/reference nor installed, so this follows its published algorithm (Rognes et al.
 * 2016; quality arithmetic of Edgar & Flyvbjerg 2015) with vsearch's documented defaults for every option the
 * reference leaves alone (--fastq_minovlen 10, --fastq_maxdiffpct 100, --fastq_qmaxout 41, --fastq_qminout 0,
 * --fastq_ascii 33, no length / N / truncation filters):
 *   1. candidate diagonals: an offset i (1 .. F+R-1) of the reverse-complemented R2 against R1 is examined only if the
 *      two reads share at least 4 exactly matching 5-mers (no ambiguous symbol) on that diagonal;
 *   2. every candidate is scored from the 3' end of the forward read with log2-odds (bits) of "the two observed
 *      bases are the same true base" given both error probabilities; a diagonal whose running score drops 16 bits
 *      or more below its running maximum scores 0; >= 16 bits is a hit; the best strictly-greater score wins;
 *   3. more than one hit -> "repeat"; staggered (i > F) unless allowed; more than maxdiffs mismatches; best < 16
 *      bits; overlap < minovlen -> not merged;
 *   4. merged read = R1's 5' overhang, consensus of the overlap (agreement / disagreement posterior qualities,
 *      an N defers to the other read), R2's 5' overhang; dropped if its expected errors exceed maxee.
 * PARITY PIN STATUS: "parity unpinned" against a real vsearch run.  The only merged fixture of the reference,
 * tests/test_data/4774-1-MSITS3_merged.fastq, was written by an older merger (additive qualities, 227 of 250 pairs);
 * it pins the OVERLAP this code finds (same bases for 225 of the 226 pairs both merge, tests/test_oracle_merge.py)
 * but not the quality arithmetic.  The live-CLI expectation of the reference's tests (235 trimmed reads from these
 * 250 pairs, tests/test_main_pytest.py:252) is consistent with the 236 pairs merged here.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { KMER = 5, MINDIAG = 4 };
static const double MINSCORE = 16.0, DROPMAX = 16.0;

typedef struct {
    double  match[94][94], mism[94][94], q2p[94];
    uint8_t same[94][94], diff[94][94];
} merge_tabs;

static double q_to_p(int x) { return x < 2 ? 0.75 : exp10(-x / 10.0); }

static uint8_t q_from_p(double p, const ora_merge_params *prm)
{
    int q = (int)round(-10.0 * log10(p));
    if (q > prm->qmaxout) q = prm->qmaxout;
    if (q < prm->qminout) q = prm->qminout;
    return (uint8_t)(prm->ascii + q);
}

static void make_tabs(merge_tabs *t, const ora_merge_params *prm)
{
    for (int x = 0; x < 94; x++) {
        const double px = q_to_p(x);
        t->q2p[x] = px;
        for (int y = 0; y < 94; y++) {
            const double py = q_to_p(y);
            t->same[x][y]  = q_from_p(px * py / 3.0 / (1.0 - px - py + 4.0 * px * py / 3.0), prm);
            t->diff[x][y]  = q_from_p(px * (1.0 - py / 3.0) / (px + py - 4.0 * px * py / 3.0), prm);   /* x = the better base */
            t->match[x][y] = log2((1.0 - px - py + px * py * 4.0 / 3.0) / 0.25);
            t->mism[x][y]  = log2(((px + py) / 3.0 - px * py * 4.0 / 9.0) / 0.25);
        }
    }
}

void ora_merge_default_params(ora_merge_params *p)
{
    p->maxdiffs = 40; p->maxee = 2.0; p->allow_stagger = 0; p->qmax = 93;
    p->minovlen = 10; p->qmaxout = 41; p->qminout = 0; p->ascii = 33; p->maxdiffpct = 100.0;
}

static unsigned char up(unsigned char c) { return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c; }
static unsigned char comp(unsigned char c)
{
    static const char *a = "ACGTURYSWKMBDHVN", *b = "TGCAAYRSWMKVHDBN";
    const char *f = strchr(a, c);
    return f && c ? (unsigned char)b[f - a] : (unsigned char)'N';
}
static int plain(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'U'; }

/* one pair; fs/fq forward read, rs/rq R2 as read from the file.  Returns the merged length (0 = not merged). */
static int merge_one(const merge_tabs *t, const ora_merge_params *prm, const uint8_t *fs0, const uint8_t *fq, int F,
                     const uint8_t *rs0, const uint8_t *rq, int R, uint8_t *reason, uint8_t *oseq, uint8_t *oqual,
                     uint8_t *work)
{
    uint8_t *fs = work, *rc = work + F, *rcq = rc + R;    /* rc / rcq: reverse complement of R2 and its qualities */
    const int a0 = prm->ascii;
    for (int p = 0; p < F; p++) {
        fs[p] = up(fs0[p]);
        const int q = fq[p] - a0;
        if (q < 0 || q > prm->qmax) { *reason = ORA_MERGE_BADQUAL; return 0; }
    }
    for (int p = 0; p < R; p++) {
        rc[p] = comp(up(rs0[R - 1 - p]));
        rcq[p] = rq[R - 1 - p];
        const int q = rcq[p] - a0;
        if (q < 0 || q > prm->qmax) { *reason = ORA_MERGE_BADQUAL; return 0; }
    }
    double best_score = 0.0;
    int best_i = 0, best_diffs = 0, hits = 0, kmers = 0;
    for (int i = 1; i <= F + R - 1; i++) {
        const int sh = F - i;                             /* forward position p pairs with rc position p - sh */
        const int p0 = sh > 0 ? sh : 0, p1 = F < sh + R ? F : sh + R;
        int run = 0, cnt = 0;
        for (int p = p0; p < p1; p++) {
            if (fs[p] == rc[p - sh] && plain(fs[p])) { if (++run >= KMER) cnt++; }
            else run = 0;
        }
        if (cnt < MINDIAG) continue;
        kmers = 1;
        double score = 0.0, high = 0.0, dropmax = 0.0;
        int diffs = 0;
        for (int p = p1 - 1; p >= p0; p--) {
            const int qa = fq[p] - a0, qb = rcq[p - sh] - a0;
            if (fs[p] == rc[p - sh]) {
                score += t->match[qa][qb];
                if (score > high) high = score;
            } else {
                score += t->mism[qa][qb];
                diffs++;
                if (score < high - dropmax) dropmax = high - score;
            }
        }
        if (dropmax >= DROPMAX) score = 0.0;
        if (score >= MINSCORE) hits++;
        if (score > best_score) { best_score = score; best_i = i; best_diffs = diffs; }
    }
    if (hits > 1) { *reason = ORA_MERGE_REPEAT; return 0; }
    if (!prm->allow_stagger && best_i > F) { *reason = ORA_MERGE_STAGGERED; return 0; }
    if (best_diffs > prm->maxdiffs) { *reason = ORA_MERGE_MAXDIFFS; return 0; }
    if (best_i > 0 && 100.0 * best_diffs / best_i > prm->maxdiffpct) { *reason = ORA_MERGE_MAXDIFFPCT; return 0; }
    if (!kmers) { *reason = ORA_MERGE_NOKMERS; return 0; }
    if (best_score < MINSCORE) { *reason = ORA_MERGE_MINSCORE; return 0; }
    if (best_i < prm->minovlen) { *reason = ORA_MERGE_MINOVLEN; return 0; }

    /* the merged read, in forward orientation: rc position of forward position p is p - (F - best_i) */
    const int sh = F - best_i;
    int n = 0;
    double ee = 0.0;
    int p = 0;
    for (; p < sh; p++) { oseq[n] = fs[p]; oqual[n] = fq[p]; ee += t->q2p[fq[p] - a0]; n++; }
    int r = p - sh;                                       /* >= 0; > 0 when staggered: R2's 3' overhang is dropped */
    for (; p < F && r < R; p++, r++) {
        const uint8_t a = fs[p], b = rc[r];
        const int qa = fq[p] - a0, qb = rcq[r] - a0;
        uint8_t s, q;
        if (b == 'N') { s = a; q = fq[p]; }
        else if (a == 'N') { s = b; q = rcq[r]; }
        else if (a == b) { s = a; q = t->same[qa][qb]; }
        else if (qa > qb) { s = a; q = t->diff[qa][qb]; }
        else { s = b; q = t->diff[qb][qa]; }
        oseq[n] = s; oqual[n] = q; ee += t->q2p[q - a0]; n++;
    }
    for (; r < R; r++) { oseq[n] = rc[r]; oqual[n] = rcq[r]; ee += t->q2p[rcq[r] - a0]; n++; }
    if (ee <= prm->maxee) { *reason = ORA_MERGE_OK; return n; }
    *reason = ORA_MERGE_MAXEE;
    return 0;
}

/* Pair i is written to the slot starting at foff[i] + roff[i] of out_seq / out_qual (capacity F + R >= merged length).
 * merged_len[i] = 0 and reason[i] != ORA_MERGE_OK for a pair that is not merged.  Returns the number merged, or -1
 * when a quality value lies outside [0, qmax] (vsearch stops with a fatal error there). */
int64_t ora_merge_pairs(const uint8_t *fseq, const uint8_t *fqual, const int64_t *foff, const uint8_t *rseq,
                        const uint8_t *rqual, const int64_t *roff, int64_t npairs, const ora_merge_params *prm,
                        int32_t *merged_len, uint8_t *reason, uint8_t *out_seq, uint8_t *out_qual, int nthreads)
{
    merge_tabs *t = (merge_tabs *)malloc(sizeof(merge_tabs));
    make_tabs(t, prm);
    int64_t nmerged = 0;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : nmerged) reduction(| : bad)
    {
        uint8_t *work = NULL;
        size_t cap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < npairs; i++) {
            const int F = (int)(foff[i + 1] - foff[i]), R = (int)(roff[i + 1] - roff[i]);
            const size_t need = (size_t)F + 2 * (size_t)R + 16;
            if (need > cap) { cap = need * 2; work = (uint8_t *)realloc(work, cap); }
            const int64_t slot = foff[i] + roff[i];
            merged_len[i] = merge_one(t, prm, fseq + foff[i], fqual + foff[i], F, rseq + roff[i], rqual + roff[i], R,
                                      &reason[i], out_seq + slot, out_qual + slot, work);
            if (merged_len[i] > 0) nmerged++;
            if (reason[i] == ORA_MERGE_BADQUAL) bad = 1;
        }
        free(work);
    }
    free(t);
    return bad ? -1 : nmerged;
}

/* the tables themselves, for known-answer tests: out[94*94] each (q index = quality value) */
void ora_merge_tables(const ora_merge_params *prm, double *match, double *mism, uint8_t *same, uint8_t *diff, double *q2p)
{
    merge_tabs *t = (merge_tabs *)malloc(sizeof(merge_tabs));
    make_tabs(t, prm);
    memcpy(match, t->match, sizeof t->match);
    memcpy(mism, t->mism, sizeof t->mism);
    memcpy(same, t->same, sizeof t->same);
    memcpy(diff, t->diff, sizeof t->diff);
    memcpy(q2p, t->q2p, sizeof t->q2p);
    free(t);
}
