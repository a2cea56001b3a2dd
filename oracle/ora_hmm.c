/*
 * ora_hmm.c -- CPU oracle: restatement of `hmmsearch --domtblout -T 10 --F1 1e-6 --F2 1e-6
 * --F3 1e-6` as invoked by the reference (itsxpress/SeqSample.py:191-209).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  HMMER (>=3.1b2, verified by the reference
 * with 3.4) is an un-vendored third-party dependency that is absent from /root/reference;
 * the algorithm below is restated from the published HMMER3 acceleration pipeline
 * (MSV filter -> bias filter -> [Viterbi filter, never run when F1==F2] -> Forward ->
 * Backward -> posterior-heuristic domain definition -> envelope rescoring with null2)
 * and anchored on SURVEY.md Appendix A.  "parity unpinned" vs a real hmmsearch.
 *
 * Numerics: probability (odds-ratio) space fp32 with sparse rescaling (xE > 1e4), like
 * HMMER's vector implementation; summation over model nodes is in plain ascending-k order
 * (HMMER's 4-lane striped order differs by fp32 rounding only).  E-value maths in double.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG2 0.69314718055994529
#define NTRANS 7
enum { T_MM = 0, T_MI, T_MD, T_IM, T_II, T_DM, T_DD };
enum { EV_MMU = 0, EV_MLAMBDA, EV_VMU, EV_VLAMBDA, EV_FTAU, EV_FLAMBDA };

typedef struct {
    char   name[192];
    int    M;
    float *mat;  /* (M+1)*4 probabilities, row 0 unused */
    float *t;    /* (M+1)*7 probabilities, row 0 = node 0 (begin) */
    float  compo[4];
    int    has_compo;
    float  ev[6];
    /* configured search profile (local, multihit), A.3 */
    float *msc;  /* (M+1)*16 match log-odds (nats) */
    float *e;    /* (M+1)*16 match odds = expf(msc) */
    float *bm;   /* (M+2) B->M_k probability, index k */
    float *tp;   /* (M+2)*7 transition probabilities out of node k (0 for k=0 and k>=M) */
    /* MSV byte profile, A.4 step 1 */
    uint8_t *cost; /* (M+1)*16 biased costs */
    int     bias_b, base_b, tbm_b, tec_b;
    float   scale_b;
    /* bias filter 2-state HMM emission odds [code][state] */
    float   eo[ORA_NCODE][2];
} prof_t;

struct ora_db {
    int     n, cap;
    prof_t *p;
};

static float viterbi_filter(const prof_t *pf, const uint8_t *dsq, int L, int *overflow);

static const int degen_mask[ORA_NCODE] = {1, 2, 4, 8, 5, 10, 3, 12, 6, 9, 11, 14, 7, 13, 15, 0};

/* ------------------------------------------------------------------------ */
/* p7_FLogsum: table-driven log(e^a + e^b), as HMMER (16000 entries, scale 1000) */
#define LOGSUM_TBL 16000
static float flogsum_lookup[LOGSUM_TBL];
static int   flogsum_ready = 0;
static void flogsum_init(void)
{
    if (flogsum_ready) return;
    for (int i = 0; i < LOGSUM_TBL; i++)
        flogsum_lookup[i] = (float)log(1. + exp((double)-i / 1000.0));
    flogsum_ready = 1;
}
float ora_flogsum(float a, float b)
{
    flogsum_init();
    const float max = a > b ? a : b;
    const float min = a > b ? b : a;
    return (min == -INFINITY || (max - min) >= 15.7f) ? max : max + flogsum_lookup[(int)((max - min) * 1000.0f)];
}

/* ------------------------------------------------------------------------ */
int ora_digitize(const char *seq, int64_t L, uint8_t *dsq)
{
    static int8_t map[256];
    static int    ready = 0;
    if (!ready) {
        memset(map, 15, sizeof(map));
        const char *sym = "ACGTRYMKSWHBVDN";
        for (int i = 0; sym[i]; i++) {
            map[(unsigned char)sym[i]]          = (int8_t)i;
            map[(unsigned char)tolower(sym[i])] = (int8_t)i;
        }
        map['U'] = map['u'] = 3;
        map['X'] = map['x'] = 14;
        ready = 1;
    }
    int bad = 0;
    for (int64_t i = 0; i < L; i++) {
        dsq[i] = (uint8_t)map[(unsigned char)seq[i]];
        if (dsq[i] == 15) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------------ */
/* HMMER3/f ASCII parser.  Values are -ln(p); '*' means p = 0.    */
static float parse_prob(const char *tok)
{
    if (tok[0] == '*') return 0.0f;
    return expf(-1.0f * (float)atof(tok));
}

static uint8_t unbiased_byteify(const prof_t *pf, float sc)
{
    sc = -1.0f * roundf(pf->scale_b * sc);
    return (sc > 255.f) ? 255 : (uint8_t)sc;
}
static uint8_t biased_byteify(const prof_t *pf, float sc)
{
    sc = -1.0f * roundf(pf->scale_b * sc);
    return (sc > 255.f - (float)pf->bias_b) ? 255 : (uint8_t)((uint8_t)sc + pf->bias_b);
}

static void configure_profile(prof_t *pf)
{
    const int M = pf->M;
    pf->msc  = calloc((size_t)(M + 1) * 16, sizeof(float));
    pf->e    = calloc((size_t)(M + 1) * 16, sizeof(float));
    pf->bm   = calloc((size_t)(M + 2), sizeof(float));
    pf->tp   = calloc((size_t)(M + 2) * 7, sizeof(float));
    pf->cost = calloc((size_t)(M + 1) * 16, 1);

    /* local entry distribution: occ[k] / sum_i occ[i]*(M-i+1) */
    float *occ = calloc((size_t)M + 1, sizeof(float));
    occ[0] = 0.f;
    occ[1] = pf->t[0 * 7 + T_MI] + pf->t[0 * 7 + T_MM];
    for (int k = 2; k <= M; k++)
        occ[k] = occ[k - 1] * (pf->t[(k - 1) * 7 + T_MM] + pf->t[(k - 1) * 7 + T_MI]) +
                 (1.0f - occ[k - 1]) * pf->t[(k - 1) * 7 + T_DM];
    float Z = 0.f;
    for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
    for (int k = 1; k <= M; k++) {
        float sc  = (float)log(occ[k] / Z);
        pf->bm[k] = expf(sc);
    }
    free(occ);

    /* transitions out of nodes 1..M-1 */
    for (int k = 1; k < M; k++)
        for (int s = 0; s < 7; s++) {
            float sc          = (float)log(pf->t[k * 7 + s]);
            pf->tp[k * 7 + s] = expf(sc);
        }

    /* match scores, degenerate codes = f-weighted mean of member scores (f uniform) */
    for (int k = 1; k <= M; k++) {
        float sc[16];
        for (int x = 0; x < 4; x++) sc[x] = (float)log((double)pf->mat[k * 4 + x] / 0.25f);
        for (int x = 4; x < 15; x++) {
            float result = 0.f, denom = 0.f;
            for (int y = 0; y < 4; y++)
                if (degen_mask[x] & (1 << y)) {
                    result += sc[y] * 0.25f;
                    denom += 0.25f;
                }
            sc[x] = result / denom;
        }
        sc[15] = -INFINITY;
        for (int x = 0; x < 16; x++) {
            pf->msc[k * 16 + x] = sc[x];
            pf->e[k * 16 + x]   = expf(sc[x]);
        }
    }

    /* MSV byte costs */
    float max = 0.0f;
    for (int k = 1; k <= M; k++)
        for (int x = 0; x < 4; x++)
            if (pf->msc[k * 16 + x] > max) max = pf->msc[k * 16 + x];
    pf->scale_b = (float)(3.0 / LOG2);
    pf->base_b  = 190;
    pf->bias_b  = 0;
    pf->bias_b  = unbiased_byteify(pf, -1.0f * max);
    for (int k = 1; k <= M; k++)
        for (int x = 0; x < 16; x++) pf->cost[k * 16 + x] = biased_byteify(pf, pf->msc[k * 16 + x]);
    pf->tbm_b = unbiased_byteify(pf, logf(2.0f / ((float)M * (float)(M + 1))));
    pf->tec_b = unbiased_byteify(pf, logf(0.5f));

    /* bias-filter HMM emission odds: state 0 = background, state 1 = model composition */
    for (int x = 0; x < 4; x++) {
        pf->eo[x][0] = 0.25f / 0.25f;
        pf->eo[x][1] = pf->compo[x] / 0.25f;
    }
    for (int x = 4; x < 15; x++)
        for (int s = 0; s < 2; s++) {
            float num = 0.f, denom = 0.f;
            for (int y = 0; y < 4; y++)
                if (degen_mask[x] & (1 << y)) {
                    num += (s == 0) ? 0.25f : pf->compo[y];
                    denom += 0.25f;
                }
            pf->eo[x][s] = (denom > 0.f) ? num / denom : 0.f;
        }
    pf->eo[15][0] = pf->eo[15][1] = 1.0f;
}

static void prof_free(prof_t *pf)
{
    free(pf->mat); free(pf->t); free(pf->msc); free(pf->e);
    free(pf->bm); free(pf->tp); free(pf->cost);
}

static int name_selected(const char *name, const char *const *prefixes, int nprefix)
{
    if (nprefix <= 0 || prefixes == NULL) return 1;
    for (int i = 0; i < nprefix; i++)
        if (strncmp(name, prefixes[i], strlen(prefixes[i])) == 0) return 1;
    return 0;
}

static int read_floats(char *line, int skip_first, float *out, int n)
{
    char *save = NULL;
    char *tok  = strtok_r(line, " \t\r\n", &save);
    if (skip_first) tok = strtok_r(NULL, " \t\r\n", &save);
    for (int i = 0; i < n; i++) {
        if (!tok) return -1;
        out[i] = parse_prob(tok);
        tok    = strtok_r(NULL, " \t\r\n", &save);
    }
    return 0;
}

int ora_db_append(ora_db *db, const char *path, const char *const *prefixes, int nprefix)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    char   *line = NULL;
    size_t  cap  = 0;
    prof_t  cur;
    int     in_rec = 0, added = 0;
    memset(&cur, 0, sizeof(cur));
    while (getline(&line, &cap, fp) > 0) {
        if (strncmp(line, "HMMER3/", 7) == 0) {
            memset(&cur, 0, sizeof(cur));
            in_rec = 1;
            continue;
        }
        if (!in_rec) continue;
        if (strncmp(line, "NAME ", 5) == 0) {
            char *s = line + 5;
            while (*s == ' ') s++;
            size_t n = strcspn(s, "\r\n");
            while (n > 0 && isspace((unsigned char)s[n - 1])) n--;
            if (n >= sizeof(cur.name)) n = sizeof(cur.name) - 1;
            memcpy(cur.name, s, n);
            cur.name[n] = 0;
        } else if (strncmp(line, "LENG ", 5) == 0) {
            cur.M = atoi(line + 5);
        } else if (strncmp(line, "STATS LOCAL", 11) == 0) {
            char  kind[32];
            float a, b;
            if (sscanf(line + 11, "%31s %f %f", kind, &a, &b) == 3) {
                if (!strcmp(kind, "MSV")) { cur.ev[EV_MMU] = a; cur.ev[EV_MLAMBDA] = b; }
                else if (!strcmp(kind, "VITERBI")) { cur.ev[EV_VMU] = a; cur.ev[EV_VLAMBDA] = b; }
                else if (!strcmp(kind, "FORWARD")) { cur.ev[EV_FTAU] = a; cur.ev[EV_FLAMBDA] = b; }
            }
        } else if (strncmp(line, "HMM ", 4) == 0) {
            /* main model section */
            const int M = cur.M;
            cur.mat = calloc((size_t)(M + 1) * 4, sizeof(float));
            cur.t   = calloc((size_t)(M + 1) * 7, sizeof(float));
            if (getline(&line, &cap, fp) <= 0) break; /* transition header line */
            if (getline(&line, &cap, fp) <= 0) break;
            char *s = line;
            while (*s == ' ') s++;
            if (strncmp(s, "COMPO", 5) == 0) {
                read_floats(line, 1, cur.compo, 4);
                cur.has_compo = 1;
                if (getline(&line, &cap, fp) <= 0) break;
            }
            /* node 0 insert emissions (ignored: inserts are hard-wired to background) */
            if (getline(&line, &cap, fp) <= 0) break;
            read_floats(line, 0, cur.t + 0, 7); /* node 0 transitions */
            for (int k = 1; k <= M; k++) {
                if (getline(&line, &cap, fp) <= 0) break;
                read_floats(line, 1, cur.mat + k * 4, 4);
                if (getline(&line, &cap, fp) <= 0) break; /* insert emissions */
                if (getline(&line, &cap, fp) <= 0) break;
                read_floats(line, 0, cur.t + k * 7, 7);
            }
        } else if (strncmp(line, "//", 2) == 0) {
            if (cur.mat && name_selected(cur.name, prefixes, nprefix)) {
                if (db->n == db->cap) {
                    db->cap = db->cap ? db->cap * 2 : 64;
                    db->p   = realloc(db->p, (size_t)db->cap * sizeof(prof_t));
                }
                configure_profile(&cur);
                db->p[db->n++] = cur;
                added++;
            } else {
                free(cur.mat);
                free(cur.t);
            }
            memset(&cur, 0, sizeof(cur));
            in_rec = 0;
        }
    }
    free(line);
    fclose(fp);
    return added;
}

ora_db *ora_db_load(const char *path, const char *const *prefixes, int nprefix)
{
    ora_db *db = calloc(1, sizeof(*db));
    if (path && ora_db_append(db, path, prefixes, nprefix) < 0) {
        free(db);
        return NULL;
    }
    flogsum_init();
    return db;
}
void ora_db_free(ora_db *db)
{
    if (!db) return;
    for (int i = 0; i < db->n; i++) prof_free(&db->p[i]);
    free(db->p);
    free(db);
}
int         ora_db_count(const ora_db *db) { return db->n; }
const char *ora_db_name(const ora_db *db, int p) { return db->p[p].name; }
int         ora_db_M(const ora_db *db, int p) { return db->p[p].M; }
void        ora_db_evparam(const ora_db *db, int p, float out6[6]) { memcpy(out6, db->p[p].ev, 6 * sizeof(float)); }
void ora_db_raw(const ora_db *db, int p, float *mat, float *t, float *compo)
{
    const prof_t *pf = &db->p[p];
    memcpy(mat, pf->mat, (size_t)(pf->M + 1) * 4 * sizeof(float));
    memcpy(t, pf->t, (size_t)(pf->M + 1) * 7 * sizeof(float));
    memcpy(compo, pf->compo, 4 * sizeof(float));
}
void ora_db_msv(const ora_db *db, int p, uint8_t *cost, int *scalars4)
{
    const prof_t *pf = &db->p[p];
    memcpy(cost, pf->cost, (size_t)(pf->M + 1) * 16);
    scalars4[0] = pf->bias_b; scalars4[1] = pf->base_b; scalars4[2] = pf->tbm_b; scalars4[3] = pf->tec_b;
}

void ora_default_params(ora_params *prm)
{
    prm->T = 10.0f;
    prm->F1 = prm->F2 = prm->F3 = 1e-6;
    prm->domE     = 10.0;
    prm->nthreads = 0;
    prm->resolve_multidomain = 1;
}

/* ------------------------------------------------------------------------ */
/* statistics (Easel esl_gumbel_surv / esl_exp_surv / esl_exp_logsurv)           */
static double gumbel_surv(double x, double mu, double lambda)
{
    double y  = lambda * (x - mu);
    double ey = -exp(-y);
    if (fabs(ey) < 5e-9) return -ey;
    return 1 - exp(ey);
}
static double exp_surv(double x, double mu, double lambda)
{
    if (x < mu) return 1.0;
    return exp(-lambda * (x - mu));
}
static double exp_logsurv(double x, double mu, double lambda)
{
    if (x < mu) return 0.0;
    return -lambda * (x - mu);
}

float ora_nullsc(int L)
{
    float p1 = (float)L / (float)(L + 1);
    return (float)((float)L * log((double)p1) + log(1. - (double)p1));
}

/* ------------------------------------------------------------------------ */
/* A.4 step 1: MSV filter, unsigned-byte saturating arithmetic.  dsq is 0-based here. */
static float msv_filter(const prof_t *pf, const uint8_t *dsq, int L, int *overflow, int *xJ_out)
{
    const int M = pf->M;
    uint8_t   dp[512];
    uint8_t  *mp = dp;
    uint8_t  *heap = NULL;
    if (M + 1 > 512) mp = heap = malloc((size_t)M + 1);
    const int tjb  = unbiased_byteify(pf, logf(3.0f / (float)(L + 3)));
    const int tjbm = (uint8_t)(tjb + pf->tbm_b);
    const int bias = pf->bias_b, base = pf->base_b, tec = pf->tec_b;
    for (int k = 0; k <= M; k++) mp[k] = 0;
    int xJ = 0;
    int xB = base - tjbm; if (xB < 0) xB = 0;
    *overflow = 0;
    for (int i = 0; i < L; i++) {
        const uint8_t *cost = pf->cost + dsq[i];
        int xE   = 0;
        int prev = 0; /* M_{k-1}(i-1); M_0 = -inf = 0 */
        for (int k = 1; k <= M; k++) {
            int sv = prev > xB ? prev : xB;
            sv += bias; if (sv > 255) sv = 255;
            sv -= cost[k * 16]; if (sv < 0) sv = 0;
            if (sv > xE) xE = sv;
            prev  = mp[k];
            mp[k] = (uint8_t)sv;
        }
        if (xE + bias >= 255) {
            *overflow = 1;
            if (xJ_out) *xJ_out = -1;
            free(heap);
            return INFINITY;
        }
        xE -= tec; if (xE < 0) xE = 0;
        if (xE > xJ) xJ = xE;
        xB = (base > xJ ? base : xJ) - tjbm; if (xB < 0) xB = 0;
    }
    free(heap);
    if (xJ_out) *xJ_out = xJ;
    float sc = ((float)(xJ - tjb) - (float)base);
    sc /= pf->scale_b;
    sc -= 3.0f;
    return sc;
}

/* A.4 step 2: bias filter = 2-state HMM Forward (Easel esl_hmm_Forward semantics) */
static float bias_filtersc(const prof_t *pf, const uint8_t *dsq, int L)
{
    const float L0 = 400.0f, L1 = (float)pf->M / 8.0f;
    const float t00 = L0 / (L0 + 1.0f), t01 = 1.0f / (L0 + 1.0f);
    const float t10 = 1.0f / (L1 + 1.0f), t11 = L1 / (L1 + 1.0f);
    const float pi0 = 0.999f, pi1 = 0.001f;
    float       logsc = 0.f, max, d0, d1;
    if (L == 0) return 0.f;
    d0  = pf->eo[dsq[0]][0] * pi0;
    d1  = pf->eo[dsq[0]][1] * pi1;
    max = 0.f;
    if (d0 > max) max = d0;
    if (d1 > max) max = d1;
    d0 /= max; d1 /= max;
    logsc += (float)log(max);
    for (int i = 1; i < L; i++) {
        float n0 = 0.f, n1 = 0.f;
        n0 += d0 * t00; n0 += d1 * t10;
        n1 += d0 * t01; n1 += d1 * t11;
        n0 *= pf->eo[dsq[i]][0];
        n1 *= pf->eo[dsq[i]][1];
        max = 0.f;
        if (n0 > max) max = n0;
        if (n1 > max) max = n1;
        d0 = n0 / max; d1 = n1 / max;
        logsc += (float)log(max);
    }
    float end = 0.f;
    end += d0 * 1.0f;
    end += d1 * 1.0f;
    logsc += (float)log(end);
    float p1 = (float)L / (float)(L + 1);
    return logsc + (float)L * logf(p1) + logf(1.f - p1);
}

/* ------------------------------------------------------------------------ */
/* Forward / Backward in odds space.  Row storage is optional (full != NULL).    */
typedef struct {
    float E_move, E_loop, N_loop, N_move; /* N, C, J share loop/move */
} xf_t;

static xf_t xf_multihit(int L)
{
    xf_t  x;
    float pmove = (2.0f + 1.0f) / ((float)L + 2.0f + 1.0f);
    x.N_move    = pmove;
    x.N_loop    = 1.0f - pmove;
    x.E_move    = expf(-(float)LOG2);
    x.E_loop    = expf(-(float)LOG2);
    return x;
}
static xf_t xf_unihit(int L)
{
    xf_t  x;
    float pmove = (2.0f + 0.0f) / ((float)L + 2.0f + 0.0f);
    x.N_move    = pmove;
    x.N_loop    = 1.0f - pmove;
    x.E_move    = 1.0f;
    x.E_loop    = 0.0f;
    return x;
}

typedef struct {
    float *E, *N, *J, *B, *C, *S; /* (L+1) each */
} specials_t;

static specials_t specials_alloc(int L)
{
    specials_t s;
    float     *buf = malloc((size_t)(L + 1) * 6 * sizeof(float));
    s.E = buf; s.N = buf + (L + 1); s.J = buf + 2 * (L + 1);
    s.B = buf + 3 * (L + 1); s.C = buf + 4 * (L + 1); s.S = buf + 5 * (L + 1);
    return s;
}
static void specials_free(specials_t *s) { free(s->E); }

/* dsq points at the first residue of the (sub)sequence; rows are 1..L.
 * fullM/fullI (optional): (L+1)*(M+1) floats, row-major, for posterior decoding. */
static float forward_engine(const prof_t *pf, const xf_t *xf, const uint8_t *dsq, int L,
                            specials_t *sp, float *fullM, float *fullI, float *fullD)
{
    const int    M  = pf->M;
    float       *Mx = calloc((size_t)(M + 2) * 3, sizeof(float));
    float       *Ix = Mx + (M + 2), *Dx = Ix + (M + 2);
    const float *tp = pf->tp, *bm = pf->bm;
    float        xN = 1.f, xJ = 0.f, xC = 0.f, xE = 0.f, xB = xf->N_move;
    float        totscale = 0.f;
    sp->E[0] = 0.f; sp->N[0] = 1.f; sp->J[0] = 0.f; sp->B[0] = xB; sp->C[0] = 0.f; sp->S[0] = 1.f;
    for (int i = 1; i <= L; i++) {
        const float *er    = pf->e + dsq[i - 1];
        float        mprev = 0.f, iprev = 0.f, dprev = 0.f; /* row i-1, node k-1 */
        float        mcur = 0.f, dcur = 0.f;                /* row i,   node k-1 */
        float        xEm = 0.f, xEd = 0.f;
        for (int k = 1; k <= M; k++) {
            const float *tk1 = tp + (k - 1) * 7; /* transitions out of node k-1 */
            const float *tk  = tp + k * 7;
            float        sv  = xB * bm[k];
            sv               = fmaf(mprev, tk1[T_MM], sv);
            sv               = fmaf(iprev, tk1[T_IM], sv);
            sv               = fmaf(dprev, tk1[T_DM], sv);
            sv               = sv * er[k * 16];
            float dc         = fmaf(dcur, tk1[T_DD], mcur * tk1[T_MD]);
            float mp = Mx[k], ip = Ix[k], dp = Dx[k];
            float ic = fmaf(ip, tk[T_II], mp * tk[T_MI]);
            Mx[k] = sv; Ix[k] = ic; Dx[k] = dc;
            xEm += sv; xEd += dc;
            mprev = mp; iprev = ip; dprev = dp;
            mcur = sv; dcur = dc;
        }
        xE = xEm + xEd;
        xN = xN * xf->N_loop;
        xC = fmaf(xC, xf->N_loop, xE * xf->E_move);
        xJ = fmaf(xJ, xf->N_loop, xE * xf->E_loop);
        xB = fmaf(xJ, xf->N_move, xN * xf->N_move);
        if (xE > 1.0e4f) {
            xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
            float inv = 1.0f / xE;
            for (int k = 1; k <= M; k++) { Mx[k] *= inv; Dx[k] *= inv; Ix[k] *= inv; }
            sp->S[i] = xE;
            totscale += (float)log((double)xE);
            xE = 1.0f;
        } else sp->S[i] = 1.0f;
        sp->E[i] = xE; sp->N[i] = xN; sp->J[i] = xJ; sp->B[i] = xB; sp->C[i] = xC;
        if (fullM) {
            memcpy(fullM + (size_t)i * (M + 1), Mx, (size_t)(M + 1) * sizeof(float));
            memcpy(fullI + (size_t)i * (M + 1), Ix, (size_t)(M + 1) * sizeof(float));
            if (fullD) memcpy(fullD + (size_t)i * (M + 1), Dx, (size_t)(M + 1) * sizeof(float));
        }
    }
    free(Mx);
    return totscale + (float)log((double)(xC * xf->N_move));
}

/* Backward.  Uses the Forward scale factors fsp->S.  Writes backward specials to bsp.
 * If fullM is given (Forward match rows) it accumulates the unnormalised expected usage of
 * every match state, acc[1..M], which is all null2-by-expectation needs (see rescore_envelope). */
static float backward_engine(const prof_t *pf, const xf_t *xf, const uint8_t *dsq, int L,
                             const specials_t *fsp, specials_t *bsp,
                             const float *fullM, const float *fullI, float *acc /*[M+1]*/)
{
    const int    M   = pf->M;
    float       *Mx  = calloc((size_t)(M + 3) * 4, sizeof(float));
    float       *Ix  = Mx + (M + 3), *Dx = Ix + (M + 3), *mpe = Dx + (M + 3);
    const float *tp  = pf->tp, *bm = pf->bm;
    float        xJ = 0.f, xB = 0.f, xN = 0.f;
    float        xC = xf->N_move;
    float        xE = xC * xf->E_move;
    float        totscale;
    if (acc) for (int a = 0; a <= M; a++) acc[a] = 0.f;

    /* row L */
    Dx[M + 1] = 0.f;
    for (int k = M; k >= 1; k--) {
        const float *tk = tp + k * 7;
        Dx[k] = fmaf(tk[T_DD], Dx[k + 1], xE);
        Mx[k] = fmaf(tk[T_MD], Dx[k + 1], xE);
        Ix[k] = 0.f;
    }
    {
        float s = fsp->S[L];
        if (s > 1.0f) {
            xE = xE / s; xN = xN / s; xC = xC / s; xJ = xJ / s; xB = xB / s;
            float inv = 1.0f / s;
            for (int k = 1; k <= M; k++) { Mx[k] *= inv; Dx[k] *= inv; Ix[k] *= inv; }
        }
        bsp->S[L] = s;
        totscale  = (float)log((double)s);
    }
    bsp->E[L] = xE; bsp->N[L] = xN; bsp->J[L] = xJ; bsp->B[L] = xB; bsp->C[L] = xC;

    for (int i = L; i >= 1; i--) {
        if (acc) {
            /* expected match-state usage: nk[k] += fM_k(i) * bM_k(i) * scale(i)   (acc = nk[M+1]) */
            const float *fM = fullM + (size_t)i * (M + 1);
            const float  w  = fsp->S[i];
            for (int k = 1; k <= M; k++) acc[k] = fmaf(fM[k] * Mx[k], w, acc[k]);
        }
        if (i == 1) break;
        /* compute row i-1 from row i */
        const int    r  = i - 1;
        const float *er = pf->e + dsq[i - 1]; /* residue x_i (row i) */
        xB = 0.f;
        for (int k = 1; k <= M; k++) {
            mpe[k] = Mx[k] * er[k * 16];
            xB     = fmaf(mpe[k], bm[k], xB);
        }
        mpe[M + 1] = 0.f;
        /* specials (need xB) */
        xC = xC * xf->N_loop;
        xJ = fmaf(xB, xf->N_move, xJ * xf->N_loop);
        xN = fmaf(xB, xf->N_move, xN * xf->N_loop);
        xE = fmaf(xC, xf->E_move, xJ * xf->E_loop);
        Dx[M + 1] = 0.f;
        for (int k = M; k >= 1; k--) {
            const float *tk    = tp + k * 7;
            float        mnext = mpe[k + 1];
            float        ic    = fmaf(mnext, tk[T_IM], Ix[k] * tk[T_II]);
            /* every M and D state also exits to E: xE is the addend the FMA chains start from */
            float        dc    = fmaf(mnext, tk[T_DM], fmaf(Dx[k + 1], tk[T_DD], xE));
            float        mc    = fmaf(mnext, tk[T_MM], fmaf(Ix[k], tk[T_MI], fmaf(Dx[k + 1], tk[T_MD], xE)));
            Mx[k] = mc; Ix[k] = ic; Dx[k] = dc;
        }
        float s = fsp->S[r];
        if (s > 1.0f) {
            xE = xE / s; xN = xN / s; xC = xC / s; xJ = xJ / s; xB = xB / s;
            float inv = 1.0f / s;
            for (int k = 1; k <= M; k++) { Mx[k] *= inv; Dx[k] *= inv; Ix[k] *= inv; }
        }
        bsp->S[r] = s;
        totscale += (float)log((double)s);
        bsp->E[r] = xE; bsp->N[r] = xN; bsp->J[r] = xJ; bsp->B[r] = xB; bsp->C[r] = xC;
    }
    /* row 0 */
    {
        const float *er = pf->e + dsq[0];
        xB = 0.f;
        for (int k = 1; k <= M; k++) xB = fmaf(Mx[k] * er[k * 16], bm[k], xB);
        xN = fmaf(xB, xf->N_move, xN * xf->N_loop);
        bsp->B[0] = xB; bsp->C[0] = 0.f; bsp->J[0] = 0.f; bsp->N[0] = xN; bsp->E[0] = 0.f; bsp->S[0] = 1.0f;
    }
    free(Mx);
    return totscale + (float)log((double)xN);
}

/* p7_DomainDecoding (A.5): btot, etot, mocc from parser specials */
static void domain_decoding(const xf_t *xf, int L, const specials_t *f, const specials_t *b,
                            float *btot, float *etot, float *mocc)
{
    float scaleproduct = 1.0f / b->N[0];
    btot[0] = etot[0] = mocc[0] = 0.f;
    for (int i = 1; i <= L; i++) {
        btot[i] = btot[i - 1] + (f->B[i - 1] * b->B[i - 1] * f->S[i - 1] * scaleproduct);
        etot[i] = etot[i - 1] + (f->E[i] * b->E[i] * f->S[i] * scaleproduct);
        float njcp = f->N[i - 1] * b->N[i] * xf->N_loop * scaleproduct;
        njcp += f->J[i - 1] * b->J[i] * xf->N_loop * scaleproduct;
        njcp += f->C[i - 1] * b->C[i] * xf->N_loop * scaleproduct;
        mocc[i] = 1.f - njcp;
    }
}

/* rescore one envelope i..j (1-based on the full sequence) in unihit mode */
static void rescore_envelope(const prof_t *pf, const uint8_t *dsq, int L, int i, int j,
                             float *n2sc, ora_dom *dom)
{
    const int  M  = pf->M;
    const int  Ld = j - i + 1;
    xf_t       xf = xf_unihit(L);
    specials_t fs = specials_alloc(Ld), bs = specials_alloc(Ld);
    float     *fM = malloc((size_t)(Ld + 1) * (M + 1) * 2 * sizeof(float));
    float     *fI = fM + (size_t)(Ld + 1) * (M + 1);
    float      acc[64];
    float envsc = forward_engine(pf, &xf, dsq + (i - 1), Ld, &fs, fM, fI, NULL);
    backward_engine(pf, &xf, dsq + (i - 1), Ld, &fs, &bs, fM, fI, acc);
    /* null2 by expectation (p7_Null2_ByExpectation): null2[x] = sum_k pbar(M_k) odds_k[x] + sum_k pbar(I_k)
     * + pbar(N) + pbar(C) + pbar(J), pbar = posterior usage averaged over the Ld envelope positions.  Every
     * envelope residue is emitted by exactly one of these states and all non-match states emit with odds 1,
     * so the non-match mass is 1 - sum_k pbar(M_k) and
     *     null2[x] = 1 + sum_k pbar(M_k) * (odds_k[x] - 1),
     * which needs only the Forward MATCH rows (half the scratch traffic of the direct form on the GPU). */
    float scaleproduct = 1.0f / bs.N[0];
    float norm         = 1.0f / (float)Ld;
    float null2[16];
    for (int x = 0; x < 4; x++) {
        float w = 0.f;
        for (int k = 1; k <= M; k++) w = fmaf(acc[k], pf->e[k * 16 + x] - 1.0f, w);
        null2[x] = 1.0f + w * scaleproduct * norm;
    }
    for (int x = 4; x < 15; x++) {
        float s = 0.f;
        int   n = 0;
        for (int y = 0; y < 4; y++)
            if (degen_mask[x] & (1 << y)) { s += null2[y]; n++; }
        null2[x] = s / (float)n;
    }
    null2[15] = 1.0f;
    float domcorrection = 0.f;
    for (int pos = i; pos <= j; pos++) {
        float v = logf(null2[dsq[pos - 1]]);
        if (n2sc) n2sc[pos] = v;
        domcorrection += v;
    }
    dom->ienv = i; dom->jenv = j;
    dom->envsc = envsc;
    dom->domcorrection = domcorrection;
    free(fM);
    specials_free(&fs);
    specials_free(&bs);
}

/* ------------------------------------------------------------------------ */
/* Multidomain regions (p7_domaindef.c: is_multidomain_region -> region_trace_ensemble ->
 * p7_spensemble_Cluster -> rescore_isolated_domain(null2_is_done = TRUE)), restated from HMMER 3.1b2+ sources
 * (SURVEY Appendix A.5).  HMMER is absent from the reference mount, so this is "parity unpinned" like the rest
 * of the HMM stage; the CUDA path (mdom_kernel) follows this restatement operation for operation.
 *
 *   1. multihit Forward (length model of the full target) over the region, full M/I/D matrix;
 *   2. the domain definition's RNG is HMMER's "fast" generator (esl_randomness_CreateFast(42): x0 = jenkins_mix3(42,
 *      87654321, 12345678), x <- 69069 x + 1, u = x / 2^32), re-seeded for every region.  DEVIATION, on purpose: HMMER
 *      draws the 200 tracebacks of a region from ONE sequential stream; here trace t draws from the same generator
 *      leap-frogged to x_(t * 2^20), i.e. 200 non-overlapping substreams, so that the traces are independent work
 *      items (the CUDA path runs one thread per trace).  The ensemble has the same distribution; it is not
 *      draw-for-draw the one hmmsearch samples (which a different fp32 summation order would not reproduce either);
 *   3. 200 stochastic tracebacks (p7_StochasticTrace, impl_sse/stotrace.c): one draw per state choice, paths
 *      normalised (esl_vec_FNorm) and chosen by cumulative sum (esl_rnd_FChoose); the E state scans
 *      M_k, D_k in the striped order of the SSE matrix (q = 0..Q-1, r = 0..3, k = r Q + q + 1);
 *   4. every domain of every trace is one segment (i, j, k, m); null2 by trace (p7_Null2_ByTrace) is averaged
 *      over the 200 traces into n2sc[] for every position of the region: n2sc[pos] = log((sum of the null2 odds
 *      of the domains that cover pos, in trace order, + number of traces that do not cover it) / 200);
 *   5. single-linkage clustering of the segments (min_overlap 0.8 of the smaller, max_diagdiff 4), clusters with
 *      posterior >= 0.25, envelope = leftmost start / rightmost end whose endpoint count reaches
 *      ceil(0.02 * cluster size); clusters ordered by start;
 *   6. each cluster envelope is rescored by a unihit Forward; its domcorrection is the sum of the trace n2sc. */
#define MD_NSAMPLES 200
#define MD_MAXSEG   1024  /* segments kept per region (200 traces x domains per trace) */
typedef struct { int idx, i, j, k, m; } spseg_t;

static uint32_t jenkins_mix3(uint32_t a, uint32_t b, uint32_t c)
{
    a -= b; a -= c; a ^= (c >> 13);
    b -= c; b -= a; b ^= (a << 8);
    c -= a; c -= b; c ^= (b >> 13);
    a -= b; a -= c; a ^= (c >> 12);
    b -= c; b -= a; b ^= (a << 16);
    c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 3);
    b -= c; b -= a; b ^= (a << 10);
    c -= a; c -= b; c ^= (b >> 15);
    return c;
}
uint32_t ora_rng_state0(uint32_t seed)
{
    uint32_t x = jenkins_mix3(seed, 87654321u, 12345678u);
    return x ? x : 42u;
}
static inline double rng_next(uint32_t *x)
{
    *x = *x * 69069u + 1u;
    return (double)*x / 4294967296.0;
}
/* state after k more draws: x_(n+k) = A x_n + C (mod 2^32), by doubling */
uint32_t ora_rng_jump(uint32_t x, uint64_t k)
{
    uint32_t A = 1u, C = 0u, a = 69069u, c = 1u;
    while (k) {
        if (k & 1u) { A = A * a; C = C * a + c; }
        c = c * (a + 1u);
        a = a * a;
        k >>= 1;
    }
    return A * x + C;
}
#define MD_STREAM_LOG2 20
#define MD_MAXTDOM 4 /* domains kept per trace */
/* esl_vec_FNorm + esl_rnd_FChoose over n <= 4 paths */
static int fchoose(uint32_t *rng, float *p, int n)
{
    float sum = 0.f;
    for (int a = 0; a < n; a++) sum += p[a];
    if (sum != 0.f) { const float inv = 1.0f / sum; for (int a = 0; a < n; a++) p[a] = p[a] * inv; } /* esl_vec_FScale(1/sum) */
    else for (int a = 0; a < n; a++) p[a] = 1.0f / (float)n;
    for (;;) {
        double roll = rng_next(rng);
        float  c    = 0.f;
        for (int a = 0; a < n; a++) { c += p[a]; if (roll < (double)c) return a; }
    }
}
enum { ST_M = 0, ST_D, ST_I, ST_N, ST_C, ST_J, ST_E, ST_B, ST_S };

static int seg_link(const spseg_t *a, const spseg_t *b)
{
    int nov = (a->j < b->j ? a->j : b->j) - (a->i > b->i ? a->i : b->i) + 1;
    int la = a->j - a->i + 1, lb = b->j - b->i + 1;
    int n  = la < lb ? la : lb;
    if ((float)nov / (float)n < 0.8f) return 0;
    nov = (a->m < b->m ? a->m : b->m) - (a->k > b->k ? a->k : b->k) + 1;
    la = a->m - a->k + 1; lb = b->m - b->k + 1;
    n  = la < lb ? la : lb;
    if ((float)nov / (float)n < 0.8f) return 0;
    if (abs((a->i - a->k) - (b->i - b->k)) > 4) return 0;
    if (abs((a->j - a->m) - (b->j - b->m)) > 4) return 0;
    return 1;
}

/* Resolves region ireg..jreg (1-based, full-sequence coordinates).  Fills n2sc[ireg..jreg] and returns the number of
 * cluster envelopes written to env_i/env_j (ascending start), at most cap. */
static int resolve_multidomain(const prof_t *pf, const uint8_t *dsq, int L, int ireg, int jreg, float *n2sc,
                               int *env_i, int *env_j, int cap)
{
    const int      M  = pf->M, Ld = jreg - ireg + 1, W = M + 1;
    const uint8_t *x  = dsq + (ireg - 1); /* x[r-1] = residue of region row r */
    xf_t           xf = xf_multihit(L);
    specials_t     fs = specials_alloc(Ld);
    float         *fM = calloc((size_t)(Ld + 1) * W * 3, sizeof(float));
    float         *fI = fM + (size_t)(Ld + 1) * W, *fD = fI + (size_t)(Ld + 1) * W;
    forward_engine(pf, &xf, x, Ld, &fs, fM, fI, fD);
    const float *tp = pf->tp, *bm = pf->bm;
    const int    Q  = (((M - 1) / 4) + 1) > 2 ? (((M - 1) / 4) + 1) : 2;

    float   *acc = calloc((size_t)Ld + 2, sizeof(float)); /* sum of the null2 odds of the domains that cover a position */
    int     *cov = calloc((size_t)Ld + 2, sizeof(int));   /* number of traces that cover it */
    spseg_t *seg = malloc(sizeof(spseg_t) * MD_MAXSEG);
    int      nseg = 0;
    const uint32_t rng0 = ora_rng_state0(42u);

    for (int t = 0; t < MD_NSAMPLES; t++) {
        uint32_t rng = ora_rng_jump(rng0, (uint64_t)t << MD_STREAM_LOG2);
        /* domains of this trace are discovered right to left */
        int dfrom[MD_MAXTDOM], dto[MD_MAXTDOM], dk[MD_MAXTDOM], dm[MD_MAXTDOM], nd = 0;   /* as the kernel */
        float dnull[MD_MAXTDOM][16];
        int   i = Ld, k = 0, sprv = ST_C, open = 0, nI = 0;
        float sx[4] = {0.f, 0.f, 0.f, 0.f};
        while (sprv != ST_S) {
            int   scur = ST_S;
            float path[4];
            switch (sprv) {
            case ST_M: {
                path[0] = fs.B[i - 1] * bm[k];
                path[1] = fM[(size_t)(i - 1) * W + k - 1] * tp[(k - 1) * 7 + T_MM];
                path[2] = fI[(size_t)(i - 1) * W + k - 1] * tp[(k - 1) * 7 + T_IM];
                path[3] = fD[(size_t)(i - 1) * W + k - 1] * tp[(k - 1) * 7 + T_DM];
                static const int st[4] = {ST_B, ST_M, ST_I, ST_D};
                scur = st[fchoose(&rng, path, 4)];
                k--; i--;
            } break;
            case ST_D: {
                path[0] = fM[(size_t)i * W + k - 1] * tp[(k - 1) * 7 + T_MD];
                path[1] = fD[(size_t)i * W + k - 1] * tp[(k - 1) * 7 + T_DD];
                scur = fchoose(&rng, path, 2) == 0 ? ST_M : ST_D;
                k--;
            } break;
            case ST_I: {
                path[0] = fM[(size_t)(i - 1) * W + k] * tp[k * 7 + T_MI];
                path[1] = fI[(size_t)(i - 1) * W + k] * tp[k * 7 + T_II];
                scur = fchoose(&rng, path, 2) == 0 ? ST_M : ST_I;
                i--;
            } break;
            case ST_N: scur = (i == 0) ? ST_S : ST_N; break;
            case ST_C: {
                path[0] = fs.C[i - 1] * xf.N_loop;
                path[1] = fs.E[i] * xf.E_move * fs.S[i];
                scur = fchoose(&rng, path, 2) == 0 ? ST_C : ST_E;
            } break;
            case ST_J: {
                path[0] = fs.J[i - 1] * xf.N_loop;
                path[1] = fs.E[i] * xf.E_loop * fs.S[i];
                scur = fchoose(&rng, path, 2) == 0 ? ST_J : ST_E;
            } break;
            case ST_E: {
                /* M_k / D_k of row i in striped order, all scaled by 1 / xE(i) */
                const double roll = rng_next(&rng);
                const float  nrm  = (float)(1.0 / (double)fs.E[i]);
                double       sum  = 0.0;
                int          done = 0;
                while (!done) {
                    for (int q = 0; q < Q && !done; q++) {
                        for (int r = 0; r < 4 && !done; r++) {
                            int kk = r * Q + q + 1;
                            sum += (double)((kk <= M ? fM[(size_t)i * W + kk] : 0.f) * nrm);
                            if (roll < sum) { k = kk; scur = ST_M; done = 1; }
                        }
                        for (int r = 0; r < 4 && !done; r++) {
                            int kk = r * Q + q + 1;
                            sum += (double)((kk <= M ? fD[(size_t)i * W + kk] : 0.f) * nrm);
                            if (roll < sum) { k = kk; scur = ST_D; done = 1; }
                        }
                    }
                    if (!done && sum < 0.99) { k = 1; scur = ST_M; done = 1; } /* HMMER raises an exception here */
                }
            } break;
            case ST_B: {
                path[0] = fs.N[i] * xf.N_move;
                path[1] = fs.J[i] * xf.N_move;
                scur = fchoose(&rng, path, 2) == 0 ? ST_N : ST_J;
            } break;
            }
            /* bookkeeping of the appended state (scur, k, i) */
            if (scur == ST_E) {
                if (nd < MD_MAXTDOM) { open = 1; dto[nd] = 0; dm[nd] = 0; dfrom[nd] = 0; dk[nd] = 0; nI = 0; sx[0] = sx[1] = sx[2] = sx[3] = 0.f; }
            } else if (scur == ST_M && open) {
                if (dto[nd] == 0) { dto[nd] = i; dm[nd] = k; }
                dfrom[nd] = i; dk[nd] = k;
                for (int a = 0; a < 4; a++) sx[a] += pf->e[k * 16 + a];
            } else if (scur == ST_I && open) {
                nI++;
            } else if (scur == ST_B && open) {
                /* p7_Null2_ByTrace over the domain's emitting states: null2[x] = (sum of the match odds of the M states
                 * used + number of insert emissions) / Ld, summed in trace-walk order */
                const int   ld   = dto[nd] - dfrom[nd] + 1;
                const float norm = 1.0f / (float)ld;
                float      *n2   = dnull[nd];
                for (int a = 0; a < 4; a++) n2[a] = (sx[a] + (float)nI) * norm;
                for (int a = 4; a < 15; a++) {
                    float sa = 0.f;
                    int   na = 0;
                    for (int y = 0; y < 4; y++)
                        if (degen_mask[a] & (1 << y)) { sa += n2[y]; na++; }
                    n2[a] = sa / (float)na;
                }
                n2[15] = 1.0f;
                open = 0;
                nd++;
            }
            if ((scur == ST_N || scur == ST_J || scur == ST_C) && scur == sprv) i--;
            sprv = scur;
        }
        /* left to right: segments for the ensemble, null2 odds per covered position */
        for (int d = nd - 1; d >= 0; d--) {
            if (nseg < MD_MAXSEG) {
                seg[nseg].idx = t; seg[nseg].i = dfrom[d] + ireg - 1; seg[nseg].j = dto[d] + ireg - 1;
                seg[nseg].k = dk[d]; seg[nseg].m = dm[d];
                nseg++;
            }
            for (int pos = dfrom[d]; pos <= dto[d]; pos++) { acc[pos] += dnull[d][x[pos - 1]]; cov[pos]++; }
        }
    }
    for (int pos = 1; pos <= Ld; pos++)
        n2sc[ireg + pos - 1] = (float)log((double)((acc[pos] + (float)(MD_NSAMPLES - cov[pos])) / (float)MD_NSAMPLES));

    /* single linkage clustering (connected components of seg_link) */
    int *asg = malloc(sizeof(int) * (size_t)(nseg + 1));
    for (int a = 0; a < nseg; a++) asg[a] = -1;
    int  nc = 0;
    int *stack = malloc(sizeof(int) * (size_t)(nseg + 1));
    for (int a = 0; a < nseg; a++) {
        if (asg[a] >= 0) continue;
        int top = 0;
        stack[top++] = a; asg[a] = nc;
        while (top) {
            int v = stack[--top];
            for (int b = 0; b < nseg; b++)
                if (asg[b] < 0 && seg_link(&seg[v], &seg[b])) { asg[b] = nc; stack[top++] = b; }
        }
        nc++;
    }
    int nenv = 0;
    for (int c = 0; c < nc; c++) {
        int ninc = 0, ntr = 0, last = -1;
        int imin = 1 << 30, imax = 0, jmin = 1 << 30, jmax = 0;
        for (int a = 0; a < nseg; a++) {
            if (asg[a] != c) continue;
            ninc++;
            if (seg[a].idx != last) { ntr++; last = seg[a].idx; }
            if (seg[a].i < imin) imin = seg[a].i;
            if (seg[a].i > imax) imax = seg[a].i;
            if (seg[a].j < jmin) jmin = seg[a].j;
            if (seg[a].j > jmax) jmax = seg[a].j;
        }
        if ((float)ntr / (float)MD_NSAMPLES < 0.25f) continue;
        const int thr = (int)ceilf((float)ninc * 0.02f);
        int best_i = imin, best_j = jmax;
        for (best_i = imin; best_i <= imax; best_i++) {
            int cnt = 0;
            for (int a = 0; a < nseg; a++) cnt += (asg[a] == c && seg[a].i == best_i);
            if (cnt >= thr) break;
        }
        for (best_j = jmax; best_j >= jmin; best_j--) {
            int cnt = 0;
            for (int a = 0; a < nseg; a++) cnt += (asg[a] == c && seg[a].j == best_j);
            if (cnt >= thr) break;
        }
        /* insert by (start, end) ascending */
        if (nenv < cap) {
            int at = nenv;
            while (at > 0 && (env_i[at - 1] > best_i || (env_i[at - 1] == best_i && env_j[at - 1] > best_j))) {
                env_i[at] = env_i[at - 1]; env_j[at] = env_j[at - 1]; at--;
            }
            env_i[at] = best_i; env_j[at] = best_j;
            nenv++;
        }
    }
    free(asg); free(stack); free(seg); free(acc); free(cov); free(fM);
    specials_free(&fs);
    return nenv;
}

/* rescore_isolated_domain with null2_is_done: unihit Forward score of the envelope, domcorrection from n2sc[] */
static void rescore_envelope_n2done(const prof_t *pf, const uint8_t *dsq, int L, int i, int j, const float *n2sc,
                                    ora_dom *dom)
{
    const int  Ld = j - i + 1;
    xf_t       xf = xf_unihit(L);
    specials_t fs = specials_alloc(Ld);
    dom->envsc    = forward_engine(pf, &xf, dsq + (i - 1), Ld, &fs, NULL, NULL, NULL);
    specials_free(&fs);
    float dc = 0.f;
    for (int pos = i; pos <= j; pos++) dc += n2sc[pos];
    dom->ienv = i; dom->jenv = j;
    dom->domcorrection = dc;
}

/* ------------------------------------------------------------------------ */
int ora_pair_run(const ora_db *db, int p, const uint8_t *dsq, int L,
                 const ora_params *prm, ora_pair *pr, ora_dom *doms, int domcap)
{
    const prof_t *pf = &db->p[p];
    memset(pr, 0, sizeof(*pr));
    /* p7_Pipeline returns at once for a zero-length target ("silently skip length 0 seqs") */
    if (L <= 0) return 0;
    pr->nullsc = ora_nullsc(L);
    int xJ;
    pr->usc          = msv_filter(pf, dsq, L, &pr->msv_overflow, &xJ);
    pr->msv_xJ       = xJ;
    float seq_score  = (float)((pr->usc - pr->nullsc) / LOG2);
    pr->P_msv        = gumbel_surv(seq_score, pf->ev[EV_MMU], pf->ev[EV_MLAMBDA]);
    if (pr->P_msv > prm->F1) return 0;
    pr->pass_msv = 1;

    pr->filtersc = bias_filtersc(pf, dsq, L);
    seq_score    = (float)((pr->usc - pr->filtersc) / LOG2);
    pr->P_bias   = gumbel_surv(seq_score, pf->ev[EV_MMU], pf->ev[EV_MLAMBDA]);
    if (pr->P_bias > prm->F1) return 0;
    pr->pass_bias = 1;
    /* Viterbi filter runs only if P > F2; with F1 == F2 -- what the reference passes -- it is never executed
     * (A.4 step 3).  An overflowing filter counts as a pass. */
    if (pr->P_bias > prm->F2) {
        int   ovf = 0;
        float vsc = viterbi_filter(pf, dsq, L, &ovf);
        if (!ovf) {
            seq_score = (float)((vsc - pr->filtersc) / LOG2);
            if (gumbel_surv(seq_score, pf->ev[EV_VMU], pf->ev[EV_VLAMBDA]) > prm->F2) return 0;
        }
    }

    xf_t       xf = xf_multihit(L);
    specials_t fs = specials_alloc(L), bs = specials_alloc(L);
    pr->fwdsc     = forward_engine(pf, &xf, dsq, L, &fs, NULL, NULL, NULL);
    seq_score     = (float)((pr->fwdsc - pr->filtersc) / LOG2);
    pr->P_fwd     = exp_surv(seq_score, pf->ev[EV_FTAU], pf->ev[EV_FLAMBDA]);
    if (pr->P_fwd > prm->F3) { specials_free(&fs); specials_free(&bs); return 0; }
    pr->pass_fwd = 1;

    pr->bcksc = backward_engine(pf, &xf, dsq, L, &fs, &bs, NULL, NULL, NULL);

    float *btot = malloc((size_t)(L + 1) * 4 * sizeof(float));
    float *etot = btot + (L + 1), *mocc = etot + (L + 1), *n2sc = mocc + (L + 1);
    domain_decoding(&xf, L, &fs, &bs, btot, etot, mocc);
    for (int q = 0; q <= L; q++) n2sc[q] = 0.f;

    const float rt1 = 0.25f, rt2 = 0.10f, rt3 = 0.20f;
    int i = -1, triggered = 0, ndom = 0;
    for (int j = 1; j <= L; j++) {
        if (!triggered) {
            if (mocc[j] - (btot[j] - btot[j - 1]) < rt2) i = j;
            else if (i == -1) i = j;
            if (mocc[j] >= rt1) triggered = 1;
        } else if (mocc[j] - (etot[j] - etot[j - 1]) < rt2) {
            pr->nregions++;
            float max = -1.0f;
            for (int z = i; z <= j; z++) {
                float a = etot[z] - etot[i - 1], b = btot[j] - btot[z - 1];
                float expected_n = a < b ? a : b;
                if (expected_n > max) max = expected_n;
            }
            int multi = (max >= rt3);
            if (multi) pr->nmultidomain++;
            if (multi && prm->resolve_multidomain) {
                int ei[16], ej[16];
                int nenv = resolve_multidomain(pf, dsq, L, i, j, n2sc, ei, ej, 16);
                for (int c = 0; c < nenv && ndom < domcap; c++) {
                    ora_dom *d = &doms[ndom];
                    memset(d, 0, sizeof(*d));
                    rescore_envelope_n2done(pf, dsq, L, ei[c], ej[c], n2sc, d);
                    d->prof = p;
                    d->is_multidomain = 1;
                    d->dom_idx = ndom;
                    ndom++;
                }
            } else if (ndom < domcap) {
                ora_dom *d = &doms[ndom];
                memset(d, 0, sizeof(*d));
                /* resolve_multidomain == 0: a multidomain region is rescored as one envelope and flagged */
                rescore_envelope(pf, dsq, L, i, j, n2sc, d);
                d->prof = p;
                d->is_multidomain = multi;
                d->dom_idx = ndom;
                ndom++;
            }
            i = -1;
            triggered = 0;
        }
    }
    pr->ndom = ndom;
    specials_free(&fs);
    specials_free(&bs);
    if (ndom == 0) { free(btot); return 0; }

    /* null2-corrected per-sequence score (A.4 step 6) */
    float seqbias = 0.f;
    for (int q = 0; q <= L; q++) seqbias += n2sc[q];
    free(btot);
    const float omega = 1.0f / 256.0f;
    seqbias           = ora_flogsum(0.0f, (float)log((double)omega) + seqbias);
    float pre_score   = (float)((pr->fwdsc - pr->nullsc) / LOG2);
    float sscore      = (float)((pr->fwdsc - (pr->nullsc + seqbias)) / LOG2);
    float sum_score = 0.0f, sbias = 0.0f;
    int   Ld = 0;
    for (int d = 0; d < ndom; d++)
        if (doms[d].envsc - doms[d].domcorrection > 0.0f) {
            sum_score += doms[d].envsc;
            Ld += doms[d].jenv - doms[d].ienv + 1;
            sbias += doms[d].domcorrection;
        }
    sbias = ora_flogsum(0.0f, (float)log((double)omega) + sbias);
    sum_score += (float)((L - Ld) * log((double)((float)L / (float)(L + 3))));
    float pre2_score = (float)((sum_score - pr->nullsc) / LOG2);
    sum_score        = (float)((sum_score - (pr->nullsc + sbias)) / LOG2);
    if (Ld > 0 && sum_score > sscore) { sscore = sum_score; pre_score = pre2_score; }
    pr->seq_score = sscore;
    pr->pre_score = pre_score;
    pr->lnP       = exp_logsurv(sscore, pf->ev[EV_FTAU], pf->ev[EV_FLAMBDA]);
    pr->reported  = (sscore >= prm->T);

    for (int d = 0; d < ndom; d++) {
        int   ld = doms[d].jenv - doms[d].ienv + 1;
        float bs_ = doms[d].envsc + (float)((L - ld) * log((double)((float)L / (float)(L + 3))));
        doms[d].dombias  = ora_flogsum(0.0f, (float)log((double)omega) + doms[d].domcorrection);
        doms[d].bitscore = (float)((bs_ - (pr->nullsc + doms[d].dombias)) / LOG2);
        doms[d].lnP      = exp_logsurv(doms[d].bitscore, pf->ev[EV_FTAU], pf->ev[EV_FLAMBDA]);
    }
    return ndom;
}

/* ------------------------------------------------------------------------ */
/* stage accessors for parity tests */
float ora_msv_score(const ora_db *db, int p, const uint8_t *dsq, int L, int *overflow)
{
    int ov, xJ;
    float sc = msv_filter(&db->p[p], dsq, L, &ov, &xJ);
    if (overflow) *overflow = ov;
    return sc;
}
float ora_bias_filtersc(const ora_db *db, int p, const uint8_t *dsq, int L)
{
    return bias_filtersc(&db->p[p], dsq, L);
}
/* ---- Viterbi filter, 16-bit (HMMER p7_ViterbiFilter semantics) ----
 * NOT on the reference's path: hmmsearch runs it only `if (P > F2)` for a pair that already has P <= F1, and ITSxpress
 * passes F1 == F2 (SeqSample.py:191-209).  It is restated here so that the `STATS LOCAL VITERBI` line of every
 * profile -- hmmbuild's own calibration of exactly this routine -- can serve as a known-answer test of the word
 * profile (tests/test_oracle_calibration.py), and as the checker for a future kernel when a caller asks for F2 < F1.
 * Scores are int16 in 1/500 bit units around base 12000 with saturating adds (-32768 acts as -infinity), insert
 * emissions are 0, the N/C/J self loops cost 0 and a flat 3 nats is charged at the end, E is reached from M only.  The
 * D path of a row is resolved in full (what the lazy-F loop of the striped code converges to). */
static int wordify(float sc)
{
    const float scale_w = 500.0f / (float)LOG2;
    if (!(sc > -INFINITY)) return -32768;
    sc = roundf(scale_w * sc);
    if (sc >= 32767.0f) return 32767;
    if (sc <= -32768.0f) return -32768;
    return (int)sc;
}
static inline int adds16(int a, int b)
{
    const int v = a + b;
    return v > 32767 ? 32767 : v < -32768 ? -32768 : v;
}
static inline int max2(int a, int b) { return a > b ? a : b; }

/* Returns the filter score in nats (the -3 nat correction included); *overflow = 1 (and +inf returned) when a row's
 * best match cell saturates, which the pipeline treats as a pass. */
static float viterbi_filter(const prof_t *pf, const uint8_t *dsq, int L, int *overflow)
{
    const int     M       = pf->M, W = M + 2;
    const float   scale_w = 500.0f / (float)LOG2;
    const int     base_w  = 12000;
    int          *buf     = malloc((size_t)W * 14 * sizeof(int));
    int *tBM = buf, *tMM = tBM + W, *tIM = tMM + W, *tDM = tIM + W, *tMD = tDM + W, *tDD = tMD + W, *tMI = tDD + W,
        *tII = tMI + W;
    int *Mp = tII + W, *Ip = Mp + W, *Dp = Ip + W, *Mn = Dp + W, *In = Mn + W, *Dn = In + W;
    for (int k = 0; k < W; k++) {
        const float *t = pf->tp + (size_t)(k <= M ? k : 0) * 7;     /* transitions out of node k (0 for k = 0, k = M) */
        const int    ok = k >= 1 && k <= M;
        tBM[k] = ok ? wordify(logf(pf->bm[k])) : -32768;
        tMM[k] = ok ? wordify(logf(t[T_MM])) : -32768;
        tIM[k] = ok ? wordify(logf(t[T_IM])) : -32768;
        tDM[k] = ok ? wordify(logf(t[T_DM])) : -32768;
        tMD[k] = ok ? wordify(logf(t[T_MD])) : -32768;
        tDD[k] = ok ? wordify(logf(t[T_DD])) : -32768;
        tMI[k] = ok ? wordify(logf(t[T_MI])) : -32768;
        tII[k] = ok ? wordify(logf(t[T_II])) : -32768;
        if (tII[k] == 0) tII[k] = -1;                               /* an II cost of 0 is not allowed in the filter */
        Mp[k] = Ip[k] = Dp[k] = -32768;
    }
    const int xw_move = wordify(logf(3.0f / ((float)L + 3.0f)));     /* N, C, J -> move; multihit length model */
    const int xw_E    = wordify(-(float)LOG2);                       /* E -> C and E -> J */
    int       xN = base_w, xB = xN + xw_move, xJ = -32768, xC = -32768;
    float     sc = -INFINITY;
    if (overflow) *overflow = 0;
    for (int i = 1; i <= L; i++) {
        const float *rs = pf->msc + dsq[i - 1];
        int          xE = -32768;
        Mn[0] = In[0] = Dn[0] = -32768;
        for (int k = 1; k <= M; k++) {
            int sv = adds16(xB, tBM[k]);
            sv     = max2(sv, adds16(Mp[k - 1], tMM[k - 1]));
            sv     = max2(sv, adds16(Ip[k - 1], tIM[k - 1]));
            sv     = max2(sv, adds16(Dp[k - 1], tDM[k - 1]));
            sv     = adds16(sv, wordify(rs[k * 16]));
            Mn[k]  = sv;
            xE     = max2(xE, sv);
            Dn[k]  = max2(adds16(Mn[k - 1], tMD[k - 1]), adds16(Dn[k - 1], tDD[k - 1]));
            In[k]  = max2(adds16(Mp[k], tMI[k]), adds16(Ip[k], tII[k]));
        }
        if (xE >= 32767) {
            if (overflow) *overflow = 1;
            free(buf);
            return INFINITY;
        }
        xC = max2(xC, xE + xw_E);                                    /* C, J, N loops cost 0 */
        xJ = max2(xJ, xE + xw_E);
        xB = max2(xJ + xw_move, xN + xw_move);
        int *sw;
        sw = Mp; Mp = Mn; Mn = sw;
        sw = Ip; Ip = In; In = sw;
        sw = Dp; Dp = Dn; Dn = sw;
    }
    if (xC > -32768) sc = ((float)xC + (float)xw_move - (float)base_w) / scale_w - 3.0f;
    free(buf);
    return sc;
}
float ora_viterbi_filter(const ora_db *db, int p, const uint8_t *dsq, int L, int *overflow)
{
    return viterbi_filter(&db->p[p], dsq, L, overflow);
}

float ora_forward_score(const ora_db *db, int p, const uint8_t *dsq, int L)
{
    xf_t       xf = xf_multihit(L);
    specials_t fs = specials_alloc(L);
    float      sc = forward_engine(&db->p[p], &xf, dsq, L, &fs, NULL, NULL, NULL);
    specials_free(&fs);
    return sc;
}
float ora_backward_score(const ora_db *db, int p, const uint8_t *dsq, int L)
{
    xf_t       xf = xf_multihit(L);
    specials_t fs = specials_alloc(L), bs = specials_alloc(L);
    forward_engine(&db->p[p], &xf, dsq, L, &fs, NULL, NULL, NULL);
    float sc = backward_engine(&db->p[p], &xf, dsq, L, &fs, &bs, NULL, NULL, NULL);
    specials_free(&fs);
    specials_free(&bs);
    return sc;
}
int ora_forward_parser(const ora_db *db, int p, const uint8_t *dsq, int L,
                       float *xE, float *xN, float *xJ, float *xB, float *xC, float *scale, float *fwdsc)
{
    xf_t       xf = xf_multihit(L);
    specials_t fs = specials_alloc(L);
    float      sc = forward_engine(&db->p[p], &xf, dsq, L, &fs, NULL, NULL, NULL);
    size_t     n  = (size_t)(L + 1) * sizeof(float);
    memcpy(xE, fs.E, n); memcpy(xN, fs.N, n); memcpy(xJ, fs.J, n);
    memcpy(xB, fs.B, n); memcpy(xC, fs.C, n); memcpy(scale, fs.S, n);
    if (fwdsc) *fwdsc = sc;
    specials_free(&fs);
    return 0;
}
int ora_domain_decoding(const ora_db *db, int p, const uint8_t *dsq, int L,
                        float *btot, float *etot, float *mocc)
{
    xf_t       xf = xf_multihit(L);
    specials_t fs = specials_alloc(L), bs = specials_alloc(L);
    forward_engine(&db->p[p], &xf, dsq, L, &fs, NULL, NULL, NULL);
    backward_engine(&db->p[p], &xf, dsq, L, &fs, &bs, NULL, NULL, NULL);
    domain_decoding(&xf, L, &fs, &bs, btot, etot, mocc);
    specials_free(&fs);
    specials_free(&bs);
    return 0;
}

/* ------------------------------------------------------------------------ */
typedef struct {
    ora_dom *d;
    int64_t  n, cap;
} domvec_t;

typedef struct {
    int32_t seq;
    int32_t first, ndom;
    double  lnP;
} hit_t;

static int hit_cmp(const void *a, const void *b)
{
    const hit_t *x = a, *y = b;
    if (x->lnP < y->lnP) return -1;
    if (x->lnP > y->lnP) return 1;
    return (x->seq < y->seq) ? -1 : (x->seq > y->seq);
}

int64_t ora_search(const ora_db *db, const uint8_t *codes, const int64_t *off, int64_t nseq,
                   const ora_params *prm, ora_dom **rows_out, int32_t *nreported_per_profile,
                   ora_stats *stats)
{
    const int P = db->n;
    ora_stats st;
    memset(&st, 0, sizeof(st));
    st.npairs_total = (int64_t)P * nseq;
    int nth = 1;
#ifdef _OPENMP
    nth = prm->nthreads > 0 ? prm->nthreads : omp_get_max_threads();
#endif
    /* per-profile hit lists, built per thread then merged */
    domvec_t *tdoms = calloc((size_t)nth, sizeof(domvec_t));
    typedef struct { hit_t *h; int64_t n, cap; int32_t prof; } hitrec_t;
    /* hits carry their profile; we bucket afterwards */
    typedef struct { hit_t h; int32_t prof; int32_t tid; } phit_t;
    phit_t **thits = calloc((size_t)nth, sizeof(phit_t *));
    int64_t *thn = calloc((size_t)nth, sizeof(int64_t)), *thcap = calloc((size_t)nth, sizeof(int64_t));
    int64_t  n_msv = 0, n_bias = 0, n_fwd = 0, n_multi = 0;
    double   msv_cells = 0, fwd_cells = 0, env_cells = 0;
    (void)sizeof(hitrec_t);

#pragma omp parallel for schedule(dynamic, 16) num_threads(nth) \
    reduction(+ : n_msv, n_bias, n_fwd, n_multi, msv_cells, fwd_cells, env_cells)
    for (int64_t s = 0; s < nseq; s++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        const uint8_t *dsq = codes + off[s];
        const int      L   = (int)(off[s + 1] - off[s]);
        ora_dom        buf[64];
        for (int p = 0; p < P; p++) {
            ora_pair pr;
            int      nd = ora_pair_run(db, p, dsq, L, prm, &pr, buf, 64);
            msv_cells += (double)L * db->p[p].M;
            if (pr.pass_msv) n_msv++;
            if (pr.pass_bias) { n_bias++; fwd_cells += (double)L * db->p[p].M; }
            if (pr.pass_fwd) n_fwd++;
            n_multi += pr.nmultidomain;
            if (nd > 0 && pr.reported) {
                domvec_t *dv = &tdoms[tid];
                if (dv->n + nd > dv->cap) {
                    dv->cap = (dv->cap ? dv->cap * 2 : 1024) + nd;
                    dv->d   = realloc(dv->d, (size_t)dv->cap * sizeof(ora_dom));
                }
                if (thn[tid] == thcap[tid]) {
                    thcap[tid] = thcap[tid] ? thcap[tid] * 2 : 1024;
                    thits[tid] = realloc(thits[tid], (size_t)thcap[tid] * sizeof(phit_t));
                }
                phit_t *ph = &thits[tid][thn[tid]++];
                ph->h.seq = (int32_t)s; ph->h.first = (int32_t)dv->n; ph->h.ndom = nd; ph->h.lnP = pr.lnP;
                ph->prof = p; ph->tid = tid;
                for (int d = 0; d < nd; d++) {
                    buf[d].seq = (int32_t)s;
                    env_cells += (double)(buf[d].jenv - buf[d].ienv + 1) * db->p[p].M;
                    dv->d[dv->n++] = buf[d];
                }
            }
        }
    }
    st.n_past_msv = n_msv; st.n_past_bias = n_bias; st.n_past_fwd = n_fwd;
    st.n_multidomain_regions = n_multi;
    st.msv_cells = msv_cells; st.fwd_cells = fwd_cells; st.bck_cells = 0; st.env_cells = env_cells;

    /* bucket hits per profile */
    int64_t *cnt = calloc((size_t)P + 1, sizeof(int64_t));
    int64_t  total_hits = 0;
    for (int t = 0; t < nth; t++)
        for (int64_t h = 0; h < thn[t]; h++) { cnt[thits[t][h].prof + 1]++; total_hits++; }
    for (int p = 0; p < P; p++) cnt[p + 1] += cnt[p];
    phit_t  *all  = malloc((size_t)(total_hits ? total_hits : 1) * sizeof(phit_t));
    int64_t *fill = calloc((size_t)P, sizeof(int64_t));
    for (int t = 0; t < nth; t++)
        for (int64_t h = 0; h < thn[t]; h++) {
            int p = thits[t][h].prof;
            all[cnt[p] + fill[p]++] = thits[t][h];
        }
    /* rows */
    int64_t  ndom_total = 0;
    for (int t = 0; t < nth; t++) ndom_total += tdoms[t].n;
    ora_dom *rows = malloc((size_t)(ndom_total ? ndom_total : 1) * sizeof(ora_dom));
    int64_t  nrows = 0;
    st.n_reported_pairs = total_hits;
    st.ndom_total = ndom_total;
    for (int p = 0; p < P; p++) {
        int64_t nh   = cnt[p + 1] - cnt[p];
        double  domZ = (double)nh; /* = nreported for this profile (A.4 step 7) */
        if (nreported_per_profile) nreported_per_profile[p] = (int32_t)nh;
        /* sort this profile's hits by lnP ascending, ties by sequence index */
        hit_t *hs = malloc((size_t)(nh ? nh : 1) * sizeof(hit_t));
        int32_t *tidv = malloc((size_t)(nh ? nh : 1) * sizeof(int32_t));
        /* need tid to find domains; encode in a parallel sort via index */
        int64_t *order = malloc((size_t)(nh ? nh : 1) * sizeof(int64_t));
        for (int64_t h = 0; h < nh; h++) { hs[h] = all[cnt[p] + h].h; hs[h].first = (int32_t)h; }
        qsort(hs, (size_t)nh, sizeof(hit_t), hit_cmp);
        for (int64_t h = 0; h < nh; h++) {
            const phit_t *ph = &all[cnt[p] + hs[h].first];
            const ora_dom *src = tdoms[ph->tid].d + ph->h.first;
            for (int d = 0; d < ph->h.ndom; d++) {
                ora_dom r = src[d];
                r.is_reported = (exp(r.lnP) * domZ <= prm->domE);
                if (r.is_reported) { rows[nrows++] = r; st.ndom_reported++; }
            }
        }
        free(hs); free(tidv); free(order);
    }
    for (int t = 0; t < nth; t++) { free(tdoms[t].d); free(thits[t]); }
    free(tdoms); free(thits); free(thn); free(thcap); free(cnt); free(fill); free(all);
    if (stats) *stats = st;
    *rows_out = rows;
    return nrows;
}

void ora_free(void *p) { free(p); }

/* ------------------------------------------------------------------------ */
/* printf("%6.1f") of an fp32 bit score, as integer tenths.  float*10 is exact in double;
 * glibc rounds exact decimal ties to even, which is what rint() does on the tenths.   */
int32_t ora_score10(float bits)
{
    double t = (double)bits * 10.0;
    return (int32_t)rint(t);
}

/* ItsPosition (SeqSample.py:400-498): strict '>' keeps the first row on ties. */
void ora_itspos(const ora_dom *rows, int64_t nrows, const int8_t *side_of_profile,
                const int32_t *seqlen, int64_t nseq,
                int32_t *start, int32_t *stop, int32_t *tlen,
                int32_t *left_score10, int32_t *left_from, int32_t *left_to,
                int32_t *right_score10, int32_t *right_from, int32_t *right_to)
{
    for (int64_t s = 0; s < nseq; s++) {
        start[s] = stop[s] = tlen[s] = -1;
        left_score10[s] = right_score10[s] = INT32_MIN;
        left_from[s] = left_to[s] = right_from[s] = right_to[s] = -1;
    }
    for (int64_t r = 0; r < nrows; r++) {
        const ora_dom *d = &rows[r];
        int side = side_of_profile[d->prof];
        if (side < 0) continue;
        int32_t sc = ora_score10(d->bitscore);
        int64_t s  = d->seq;
        if (side == 0) {
            if (left_score10[s] == INT32_MIN || sc > left_score10[s]) {
                if (left_score10[s] == INT32_MIN) tlen[s] = seqlen[s];
                left_score10[s] = sc; left_from[s] = d->ienv; left_to[s] = d->jenv;
            }
        } else {
            if (right_score10[s] == INT32_MIN || sc > right_score10[s]) {
                if (right_score10[s] == INT32_MIN) tlen[s] = seqlen[s];
                right_score10[s] = sc; right_from[s] = d->ienv; right_to[s] = d->jenv;
            }
        }
    }
    for (int64_t s = 0; s < nseq; s++) {
        if (left_score10[s] != INT32_MIN) start[s] = left_to[s];
        if (right_score10[s] != INT32_MIN) stop[s] = right_from[s] - 1;
    }
}
