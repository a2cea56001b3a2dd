/*
 * ora_derep.c -- CPU oracle: restatement of `vsearch --fastx_uniques <fq> --fastaout rep.fa
 * --uc uc.txt --strand both` as invoked by the reference (itsxpress/SeqSample.py:106-116),
 * and of the trim / re-expansion rules (SeqSample.py:564-711, 792-884).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  vsearch (>=2.21.1) is an un-vendored third-party
 * binary; semantics follow SURVEY.md Appendix B and are pinned against the reference's own
 * vsearch fixture tests/test_data/ex_tmpdir/{seq.fq.gz,uc.txt,rep.fa}.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static unsigned char norm_tab[256], comp_tab[256];
static int           tabs_ready = 0;
static void          init_tabs(void)
{
    if (tabs_ready) return;
    for (int c = 0; c < 256; c++) {
        int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
        if (u == 'U') u = 'T';
        norm_tab[c] = (unsigned char)u;
        comp_tab[c] = (unsigned char)u;
    }
    const char *a = "ACGTRYMKSWHBVDN", *b = "TGCAYRKMSWDVBHN";
    for (int i = 0; a[i]; i++) {
        comp_tab[(unsigned char)a[i]]        = (unsigned char)b[i];
        comp_tab[(unsigned char)(a[i] + 32)] = (unsigned char)b[i];
    }
    comp_tab['U'] = comp_tab['u'] = 'A';
    tabs_ready = 1;
}

static uint64_t fnv1a(const unsigned char *s, int64_t n)
{
    uint64_t h = 1469598103934665603ULL;
    for (int64_t i = 0; i < n; i++) { h ^= s[i]; h *= 1099511628211ULL; }
    return h;
}

/* Exact full-length match on the normalised (upper-case, U->T) string.  Each read is looked up
 * forward first, then as its reverse complement; the first read of a class is its representative. */
int64_t ora_derep(const char *seq, const int64_t *off, int64_t nreads, int32_t *rep_index, uint8_t *strand)
{
    init_tabs();
    int64_t cap = 16;
    while (cap < nreads * 2 + 16) cap <<= 1;
    int32_t *table = malloc((size_t)cap * sizeof(int32_t)); /* read index of representative */
    for (int64_t i = 0; i < cap; i++) table[i] = -1;
    int64_t maxL = 0;
    for (int64_t i = 0; i < nreads; i++)
        if (off[i + 1] - off[i] > maxL) maxL = off[i + 1] - off[i];
    unsigned char *fw = malloc((size_t)maxL + 1), *rc = malloc((size_t)maxL + 1), *tmp = malloc((size_t)maxL + 1);
    int64_t nclust = 0;
    for (int64_t i = 0; i < nreads; i++) {
        const int64_t L = off[i + 1] - off[i];
        const unsigned char *s = (const unsigned char *)seq + off[i];
        for (int64_t j = 0; j < L; j++) { fw[j] = norm_tab[s[j]]; rc[L - 1 - j] = comp_tab[s[j]]; }
        int found = 0;
        for (int pass = 0; pass < 2 && !found; pass++) {
            const unsigned char *q = pass == 0 ? fw : rc;
            uint64_t h = fnv1a(q, L) & (uint64_t)(cap - 1);
            while (table[h] >= 0) {
                int64_t r = table[h];
                if (off[r + 1] - off[r] == L) {
                    const unsigned char *rs = (const unsigned char *)seq + off[r];
                    for (int64_t j = 0; j < L; j++) tmp[j] = norm_tab[rs[j]];
                    if (memcmp(tmp, q, (size_t)L) == 0) {
                        rep_index[i] = (int32_t)r;
                        strand[i]    = (uint8_t)pass;
                        found        = 1;
                        break;
                    }
                }
                h = (h + 1) & (uint64_t)(cap - 1);
            }
        }
        if (!found) {
            uint64_t h = fnv1a(fw, L) & (uint64_t)(cap - 1);
            while (table[h] >= 0) h = (h + 1) & (uint64_t)(cap - 1);
            table[h]     = (int32_t)i;
            rep_index[i] = (int32_t)i;
            strand[i]    = 0;
            nclust++;
        }
    }
    free(table); free(fw); free(rc); free(tmp);
    return nclust;
}

/* Python slice clipping for non-negative-or-negative integer bounds on a string of length n */
static void py_slice(int64_t a, int64_t b, int64_t n, int32_t *lo, int32_t *hi)
{
    if (a < 0) { a += n; if (a < 0) a = 0; }
    if (b < 0) { b += n; if (b < 0) b = 0; }
    if (a > n) a = n;
    if (b > n) b = n;
    if (b < a) b = a;
    *lo = (int32_t)a;
    *hi = (int32_t)b;
}

/* SeqSample.py:814-825 (filter) and :862 / :639-655 (slices). */
int64_t ora_trim_bounds(const int64_t *off, int64_t nreads, const int32_t *rep_index,
                        const int32_t *start, const int32_t *stop, const int32_t *tlen,
                        int mode, const int64_t *off_r2,
                        uint8_t *keep, int32_t *out_lo, int32_t *out_hi)
{
    int64_t nkept = 0;
    for (int64_t i = 0; i < nreads; i++) {
        int32_t r = rep_index[i];
        keep[i] = 0; out_lo[i] = out_hi[i] = 0;
        if (r < 0) continue;
        if (start[r] < 0 || stop[r] < 0) continue;   /* None */
        if (!(start[r] < stop[r])) continue;
        keep[i] = 1;
        nkept++;
        if (mode == 0 || mode == 2) {
            int64_t n = off[i + 1] - off[i];
            /* single-end (mode 0): record[start:stop] (SeqSample.py:862);
             * paired R1 (mode 2): stop > tlen -> record[start:], else [start:stop] (:642-645) */
            if (mode == 2 && stop[r] > tlen[r]) py_slice(start[r], n, n, &out_lo[i], &out_hi[i]);
            else py_slice(start[r], stop[r], n, &out_lo[i], &out_hi[i]);
        } else {
            int64_t n = off_r2[i + 1] - off_r2[i];
            int64_t r2start = (int64_t)tlen[r] - stop[r];
            int64_t r2end   = (int64_t)tlen[r] - start[r];
            if (r2end > tlen[r]) {
                /* record2[r2start:] */
                int32_t lo, hi;
                py_slice(r2start, n, n, &lo, &hi);
                out_lo[i] = lo; out_hi[i] = (int32_t)n;
                if (out_lo[i] > out_hi[i]) out_lo[i] = out_hi[i];
            } else {
                py_slice(r2start, r2end, n, &out_lo[i], &out_hi[i]);
            }
        }
    }
    return nkept;
}
