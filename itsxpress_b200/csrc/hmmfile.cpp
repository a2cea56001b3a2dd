// hmmfile.cpp -- HMMER3/f ASCII profile reader and search-profile configuration (host side).
//
// Replaces what `hmmsearch` does with the runtime HMM file the reference hands it
// (itsxpress/SeqSample.py:191-209) and the prefix filter of create_runtime_hmm
// (itsxpress/main.py:200-229): profiles whose NAME starts with one of the requested prefixes are
// kept, in file order.  Configuration follows HMMER3's local multihit search profile
// (SURVEY.md Appendix A.1-A.4): match log-odds against a uniform DNA background, degenerate
// residues scored by the mean of their members, occupancy-weighted local entry, MSV byte costs
// in third-bits, bias-filter emission odds from COMPO.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include "itsx_internal.h"

namespace {

const double kLn2 = 0.69314718055994529;
// IUPAC membership masks over ACGT for codes 0..15 (ACGT RYMKSWHBVDN, 15 = none)
const int kDegen[ITSX_NCODE] = {1, 2, 4, 8, 5, 10, 3, 12, 6, 9, 11, 14, 7, 13, 15, 0};

float prob_from_token(const std::string &tok)
{
    if (!tok.empty() && tok[0] == '*') return 0.0f;
    return expf(-1.0f * (float)atof(tok.c_str()));
}

bool read_probs(const std::string &line, bool skip_first, float *out, int n)
{
    std::istringstream is(line);
    std::string tok;
    if (skip_first && !(is >> tok)) return false;
    for (int i = 0; i < n; i++) {
        if (!(is >> tok)) return false;
        out[i] = prob_from_token(tok);
    }
    return true;
}

uint8_t biased_byteify(const HostProfile &pf, float sc)
{
    sc = -1.0f * roundf(pf.scale_b * sc);
    return (sc > 255.f - (float)pf.bias_b) ? 255 : (uint8_t)((uint8_t)sc + pf.bias_b);
}

void configure(HostProfile &pf)
{
    const int M = pf.M;
    pf.msc.assign((size_t)(M + 1) * 16, 0.f);
    pf.e.assign((size_t)(M + 1) * 16, 0.f);
    pf.bm.assign((size_t)M + 2, 0.f);
    pf.tp.assign((size_t)(M + 2) * 7, 0.f);
    pf.cost.assign((size_t)(M + 1) * 16, 0);

    // local entry: occ[k] / sum_i occ[i] * (M - i + 1)
    std::vector<float> occ((size_t)M + 1, 0.f);
    occ[1] = pf.t[T_MI] + pf.t[T_MM];
    for (int k = 2; k <= M; k++)
        occ[k] = occ[k - 1] * (pf.t[(k - 1) * 7 + T_MM] + pf.t[(k - 1) * 7 + T_MI]) +
                 (1.0f - occ[k - 1]) * pf.t[(k - 1) * 7 + T_DM];
    float Z = 0.f;
    for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
    for (int k = 1; k <= M; k++) pf.bm[k] = expf((float)log(occ[k] / Z));

    for (int k = 1; k < M; k++)
        for (int s = 0; s < 7; s++) pf.tp[k * 7 + s] = expf((float)log(pf.t[k * 7 + s]));

    for (int k = 1; k <= M; k++) {
        float sc[16];
        for (int x = 0; x < 4; x++) sc[x] = (float)log((double)pf.mat[k * 4 + x] / 0.25f);
        for (int x = 4; x < 15; x++) {
            float num = 0.f, den = 0.f;
            for (int y = 0; y < 4; y++)
                if (kDegen[x] & (1 << y)) {
                    num += sc[y] * 0.25f;
                    den += 0.25f;
                }
            sc[x] = num / den;
        }
        sc[15] = -INFINITY;
        for (int x = 0; x < 16; x++) {
            pf.msc[k * 16 + x] = sc[x];
            pf.e[k * 16 + x]   = expf(sc[x]);
        }
    }

    float maxsc = 0.0f;
    for (int k = 1; k <= M; k++)
        for (int x = 0; x < 4; x++) maxsc = std::max(maxsc, pf.msc[k * 16 + x]);
    pf.scale_b = (float)(3.0 / kLn2);
    pf.base_b  = 190;
    pf.bias_b  = msv_unbiased_byteify(pf.scale_b, -1.0f * maxsc);
    for (int k = 1; k <= M; k++)
        for (int x = 0; x < 16; x++) pf.cost[k * 16 + x] = biased_byteify(pf, pf.msc[k * 16 + x]);
    pf.tbm_b = msv_unbiased_byteify(pf.scale_b, logf(2.0f / ((float)M * (float)(M + 1))));
    pf.tec_b = msv_unbiased_byteify(pf.scale_b, logf(0.5f));

    for (int x = 0; x < 4; x++) {
        pf.eo[x][0] = 0.25f / 0.25f;
        pf.eo[x][1] = pf.compo[x] / 0.25f;
    }
    for (int x = 4; x < 15; x++)
        for (int s = 0; s < 2; s++) {
            float num = 0.f, den = 0.f;
            for (int y = 0; y < 4; y++)
                if (kDegen[x] & (1 << y)) {
                    num += (s == 0) ? 0.25f : pf.compo[y];
                    den += 0.25f;
                }
            pf.eo[x][s] = den > 0.f ? num / den : 0.f;
        }
    pf.eo[15][0] = pf.eo[15][1] = 1.0f;
}

bool selected(const std::string &name, const char *const *prefixes, int nprefix)
{
    if (nprefix <= 0 || !prefixes) return true;
    for (int i = 0; i < nprefix; i++)
        if (name.compare(0, strlen(prefixes[i]), prefixes[i]) == 0) return true;
    return false;
}

}  // namespace

uint8_t msv_unbiased_byteify(float scale_b, float sc)
{
    sc = -1.0f * roundf(scale_b * sc);
    return (sc > 255.f) ? 255 : (uint8_t)sc;
}

int hmmfile_append(const char *path, const char *const *prefixes, int nprefix,
                   std::vector<HostProfile> &out, std::string &err)
{
    std::ifstream in(path);
    if (!in) {
        err = std::string("cannot open profile file ") + path;
        return ITSX_EIO;
    }
    int added = 0;
    std::string line;
    HostProfile cur;
    bool in_rec = false, have_model = false;
    while (std::getline(in, line)) {
        if (line.compare(0, 7, "HMMER3/") == 0) {
            cur = HostProfile();
            in_rec = true;
            have_model = false;
            continue;
        }
        if (!in_rec) continue;
        if (line.compare(0, 5, "NAME ") == 0) {
            size_t a = line.find_first_not_of(" \t", 5);
            size_t b = line.find_last_not_of(" \t\r\n");
            cur.name = (a == std::string::npos) ? "" : line.substr(a, b - a + 1);
        } else if (line.compare(0, 5, "LENG ") == 0) {
            cur.M = atoi(line.c_str() + 5);
        } else if (line.compare(0, 11, "STATS LOCAL") == 0) {
            char kind[32];
            float a, b;
            if (sscanf(line.c_str() + 11, "%31s %f %f", kind, &a, &b) == 3) {
                if (!strcmp(kind, "MSV")) { cur.ev[EV_MMU] = a; cur.ev[EV_MLAMBDA] = b; }
                else if (!strcmp(kind, "VITERBI")) { cur.ev[EV_VMU] = a; cur.ev[EV_VLAMBDA] = b; }
                else if (!strcmp(kind, "FORWARD")) { cur.ev[EV_FTAU] = a; cur.ev[EV_FLAMBDA] = b; }
            }
        } else if (line.compare(0, 4, "HMM ") == 0) {
            const int M = cur.M;
            if (M <= 0) { err = std::string("profile without LENG in ") + path; return ITSX_EIO; }
            cur.mat.assign((size_t)(M + 1) * 4, 0.f);
            cur.t.assign((size_t)(M + 1) * 7, 0.f);
            bool ok = (bool)std::getline(in, line);            // column header of the transitions
            ok = ok && (bool)std::getline(in, line);
            if (ok && line.find("COMPO") != std::string::npos) {
                ok = read_probs(line, true, cur.compo, 4);
                ok = ok && (bool)std::getline(in, line);       // node-0 insert emissions
            }
            ok = ok && (bool)std::getline(in, line);           // node-0 transitions
            ok = ok && read_probs(line, false, &cur.t[0], 7);
            for (int k = 1; ok && k <= M; k++) {
                ok = (bool)std::getline(in, line) && read_probs(line, true, &cur.mat[k * 4], 4);
                ok = ok && (bool)std::getline(in, line);       // insert emissions: scored 0 in search profiles
                ok = ok && (bool)std::getline(in, line) && read_probs(line, false, &cur.t[k * 7], 7);
            }
            if (!ok) { err = std::string("truncated profile '") + cur.name + "' in " + path; return ITSX_EIO; }
            have_model = true;
        } else if (line.compare(0, 2, "//") == 0) {
            if (have_model && selected(cur.name, prefixes, nprefix)) {
                configure(cur);
                out.push_back(cur);
                added++;
            }
            in_rec = false;
            have_model = false;
        }
    }
    return added;
}
