/*
 * oracle.h -- CPU oracle for the ITSxpress hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of what the reference pipeline computes on the
 * path  deduplicate -> _search -> ItsPosition -> Dedup -> create_*trimmed_seqs
 * (reference: itsxpress/SeqSample.py:93-131, 178-225, 368-498, 542-562, 564-949;
 * itsxpress/main.py:176-231).  The arithmetic of that path lives in two
 * third-party binaries that are NOT in /root/reference and NOT installed:
 *   - vsearch >= 2.21.1  (`--fastx_uniques --strand both`), restated in ora_derep.c
 *   - HMMER  >= 3.1b2    (`hmmsearch --domtblout -T 10 --F1 1e-6 --F2 1e-6 --F3 1e-6`),
 *     restated in ora_hmm.c from the published algorithm (Eddy 2011, "Accelerated
 *     profile HMM searches"; HMMER 3.x p7_pipeline / p7_domaindef behaviour).
 *
 * PARITY PIN STATUS:
 *   derep   : pinned against the reference's own vsearch fixture
 *             tests/test_data/ex_tmpdir/{seq.fq.gz,uc.txt,rep.fa}.
 *   trim    : pinned byte-for-byte against tests/test_data/t2_r1.fq, t2_r2.fq and
 *             singleOut/.../4774-1-MSITS3_0_L001_R1_001.fastq.gz.
 *   merge   : (ora_merge.c, `vsearch --fastq_mergepairs` as SeqSample.py:314-349 calls it) "parity unpinned"
 *             against a real vsearch run: the reference holds no vsearch-made merge output.  Checked against an
 *             independent definition-level Python statement (tests/golden/make_merge_golden.py), the merged bases
 *             of the reference's older merged fixture, and constructed known answers.
 *   hmm     : "parity unpinned" against real hmmsearch output (no HMMER binary, the
 *             domtbl.txt fixture and F.hmm are missing from the mount).  What IS
 *             pinned: the STATS LOCAL calibration lines of every profile (lambda exactly,
 *             mu/tau statistically) -- a HMMER-free known-answer test of parsing,
 *             profile configuration, MSV/Forward arithmetic (tests/test_oracle_calibration.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this library.  The product (itsxpress_b200/) never does.
 */
#ifndef ITSX_ORACLE_H
#define ITSX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_NCODE 16          /* residue codes: 0..3 ACGT, 4..14 RYMKSWHBVDN, 15 illegal */
#define ORA_CODE_N 14

/* ---- profile database ------------------------------------------------- */
typedef struct ora_db ora_db;

ora_db *ora_db_load(const char *path, const char *const *prefixes, int nprefix);
int     ora_db_append(ora_db *db, const char *path, const char *const *prefixes, int nprefix);
void    ora_db_free(ora_db *db);
int     ora_db_count(const ora_db *db);
const char *ora_db_name(const ora_db *db, int p);
int     ora_db_M(const ora_db *db, int p);
void    ora_db_evparam(const ora_db *db, int p, float out6[6]);
/* raw probabilities as parsed: mat[(M+1)*4], t[(M+1)*7] */
void    ora_db_raw(const ora_db *db, int p, float *mat, float *t, float *compo);
/* configured MSV byte profile: cost[(M+1)*16], scalars {bias, base, tbm, tec} */
void    ora_db_msv(const ora_db *db, int p, uint8_t *cost, int *scalars4);

/* ---- sequence helpers -------------------------------------------------- */
/* ascii -> residue codes; returns number of illegal characters */
int ora_digitize(const char *seq, int64_t L, uint8_t *dsq);

/* ---- single-pair stage functions (for stage-by-stage parity tests) ------ */
typedef struct {
    int32_t msv_xJ;       /* final xJ byte, or -1 on overflow */
    int32_t msv_overflow; /* 1 if overflow (score = +inf, passes) */
    float   usc;          /* MSV score (nats), +inf on overflow */
    float   nullsc;       /* null1 score (nats) */
    float   filtersc;     /* bias-filter null score (nats); valid if pass_msv */
    float   fwdsc;        /* Forward parser score (nats); valid if pass_bias */
    double  P_msv, P_bias, P_fwd;
    int32_t pass_msv, pass_bias, pass_fwd;
    float   bcksc;        /* Backward parser score (nats); valid if pass_fwd */
    int32_t nregions;
    int32_t nmultidomain; /* regions flagged multidomain */
    int32_t ndom;
    int32_t reported;     /* per-sequence score >= T */
    float   seq_score;    /* bits */
    float   pre_score;    /* bits */
    double  lnP;
} ora_pair;

typedef struct {
    int32_t seq, prof;        /* indices into the searched set */
    int32_t ienv, jenv;       /* 1-based envelope coords */
    float   envsc;            /* nats */
    float   domcorrection;    /* nats */
    float   bitscore;         /* bits */
    float   dombias;          /* nats */
    double  lnP;
    int32_t dom_idx;          /* 0-based index of the domain within its (seq,prof) hit */
    int32_t is_multidomain;   /* envelope comes from a region flagged multidomain */
    int32_t is_reported;      /* filled by ora_search (needs domZ) */
    int32_t pad;
} ora_dom;

typedef struct {
    float  T;          /* -T    (10)   */
    double F1, F2, F3; /* 1e-6 each    */
    double domE;       /* 10.0         */
    int    nthreads;   /* OpenMP threads (<=0: all) */
    int    resolve_multidomain; /* 1 (default): multidomain regions go through the stochastic-traceback clustering */
} ora_params;

void ora_default_params(ora_params *prm);
/* HMMER's "fast" generator as the multidomain resolver uses it: state after seeding, state after k more draws */
uint32_t ora_rng_state0(uint32_t seed);
uint32_t ora_rng_jump(uint32_t x, uint64_t k);

/* run the cascade for one (sequence, profile).  doms: caller buffer of capacity domcap.
 * Returns number of domains written (also in pr->ndom). */
int ora_pair_run(const ora_db *db, int p, const uint8_t *dsq, int L,
                 const ora_params *prm, ora_pair *pr, ora_dom *doms, int domcap);

/* parser specials for parity tests: arrays of (L+1) floats each */
int ora_forward_parser(const ora_db *db, int p, const uint8_t *dsq, int L,
                       float *xE, float *xN, float *xJ, float *xB, float *xC, float *scale, float *fwdsc);
int ora_domain_decoding(const ora_db *db, int p, const uint8_t *dsq, int L,
                        float *btot, float *etot, float *mocc);
/* raw filter scores for calibration tests */
float ora_msv_score(const ora_db *db, int p, const uint8_t *dsq, int L, int *overflow);
float ora_forward_score(const ora_db *db, int p, const uint8_t *dsq, int L);
/* 16-bit Viterbi filter score in nats (p7_ViterbiFilter; never run on the reference's path, F1 == F2: see ora_hmm.c) */
float ora_viterbi_filter(const ora_db *db, int p, const uint8_t *dsq, int L, int *overflow);
float ora_backward_score(const ora_db *db, int p, const uint8_t *dsq, int L);
float ora_nullsc(int L);
float ora_bias_filtersc(const ora_db *db, int p, const uint8_t *dsq, int L);
float ora_flogsum(float a, float b);

/* ---- whole search (= one hmmsearch run over nseq targets) ---------------- */
typedef struct {
    int64_t  npairs_total, n_past_msv, n_past_bias, n_past_fwd, n_reported_pairs;
    int64_t  n_multidomain_regions;
    int64_t  ndom_total, ndom_reported;
    double   msv_cells, fwd_cells, bck_cells, env_cells;
} ora_stats;

/* seqs: concatenated residue codes, off[nseq+1].  Output rows (domtbl-equivalent, in
 * hmmsearch row order: profile file order; within profile by lnP asc then seq index;
 * within a hit by position) are malloc'd into *rows_out (free with ora_free). */
int64_t ora_search(const ora_db *db, const uint8_t *codes, const int64_t *off, int64_t nseq,
                   const ora_params *prm, ora_dom **rows_out, int32_t *nreported_per_profile,
                   ora_stats *stats);
void ora_free(void *p);

/* ---- ItsPosition restatement (SeqSample.py:380-498) ----------------------- */
/* side_of_profile[p]: 0 = left prefix, 1 = right prefix, -1 = neither.
 * Outputs per sequence: start/stop/tlen (-1 == None), plus the winning rows' printed
 * score*10 and from/to for both sides (score10 = INT32_MIN when absent). */
void ora_itspos(const ora_dom *rows, int64_t nrows, const int8_t *side_of_profile,
                const int32_t *seqlen, int64_t nseq,
                int32_t *start, int32_t *stop, int32_t *tlen,
                int32_t *left_score10, int32_t *left_from, int32_t *left_to,
                int32_t *right_score10, int32_t *right_from, int32_t *right_to);
int32_t ora_score10(float bits);   /* printf("%6.1f") rounding, as integer tenths */

/* ---- derep restatement (vsearch --fastx_uniques --strand both) ------------ */
/* seq: concatenated ASCII reads, off[nreads+1]. rep_index[i] = index of the first read
 * of i's cluster; strand[i] = 0 '+', 1 '-'.  Returns number of clusters. */
int64_t ora_derep(const char *seq, const int64_t *off, int64_t nreads,
                  int32_t *rep_index, uint8_t *strand);

/* ---- trim / re-expansion restatement (SeqSample.py:792-884, 564-711) ------- */
/* For every read i (input order): keep iff start[rep]!=None && stop[rep]!=None && start<stop.
 * single: slice [start:stop] of seq and qual (python slice clipping).
 * Outputs: keep[i], out_lo[i], out_hi[i] (clipped slice bounds into read i). */
int64_t ora_trim_bounds(const int64_t *off, int64_t nreads, const int32_t *rep_index,
                        const int32_t *start, const int32_t *stop, const int32_t *tlen,
                        int mode /*0 single-end, 2 paired R1, 1 paired R2*/,
                        const int64_t *off_r2,
                        uint8_t *keep, int32_t *out_lo, int32_t *out_hi);

/* ---- paired-end merge: `vsearch --fastq_mergepairs` as SeqSample.py:314-349 calls it (ora_merge.c) ---- */
enum { ORA_MERGE_OK = 0, ORA_MERGE_REPEAT, ORA_MERGE_STAGGERED, ORA_MERGE_MAXDIFFS, ORA_MERGE_MAXDIFFPCT,
       ORA_MERGE_NOKMERS, ORA_MERGE_MINSCORE, ORA_MERGE_MINOVLEN, ORA_MERGE_MAXEE, ORA_MERGE_BADQUAL };
typedef struct {
    int32_t maxdiffs;        /* --fastq_maxdiffs 40  (definitions.py:79) */
    int32_t allow_stagger;   /* --fastq_allowmergestagger */
    int32_t qmax;            /* --fastq_qmax 93      (definitions.py:82) */
    int32_t minovlen, qmaxout, qminout, ascii;   /* vsearch defaults 10, 41, 0, 33 */
    int32_t pad;
    double  maxee;           /* --fastq_maxee 2 */
    double  maxdiffpct;      /* vsearch default 100 */
} ora_merge_params;
void    ora_merge_default_params(ora_merge_params *p);
int64_t ora_merge_pairs(const uint8_t *fseq, const uint8_t *fqual, const int64_t *foff, const uint8_t *rseq,
                        const uint8_t *rqual, const int64_t *roff, int64_t npairs, const ora_merge_params *prm,
                        int32_t *merged_len, uint8_t *reason, uint8_t *out_seq, uint8_t *out_qual, int nthreads);
void    ora_merge_tables(const ora_merge_params *prm, double *match, double *mism, uint8_t *same, uint8_t *diff,
                         double *q2p);

#ifdef __cplusplus
}
#endif
#endif
