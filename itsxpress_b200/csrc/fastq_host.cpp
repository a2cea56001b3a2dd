// fastq_host.cpp -- host-side FASTQ scanner / packer / formatter of libitsx_b200 (multi-threaded C++).
//
// Replaces the per-record Biopython work on the reference's hot path: SeqIO.parse(handle, "fastq") feeding
// Dedup's generators (itsxpress/SeqSample.py:746-757, 926-949) and SeqIO.write(seqs, g, "fastq")
// (SeqSample.py:727-733, 912-945).  The scanner turns a decompressed 4-line FASTQ file into offset arrays
// (structure of arrays: what the device path wants), the packer lays sequences / qualities back to back
// for itsx_derep / itsx_trim_*, and the formatter writes '@title\nseq\n+\nqual\n' records from the slices that
// itsx_trim_gather returned.  Semantics checked: SURVEY.md Appendix C (title right-stripped, '+' line may repeat
// the title, len(seq) == len(qual), qualities in ASCII 33..126, otherwise an error).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstring>
#include <new>
#include <mutex>
#include <exception>
#include <unordered_map>
#include <cstdio>
#include <cmath>
#include <string>
#include <cstdlib>
#include <thread>
#if defined(__linux__)
#include <sched.h>
#endif
#include <vector>
#include "itsx_internal.h"

namespace {

thread_local std::string g_host_err;

// line ends of the buffer the counting call (cap == 0) has just scanned, kept for the filling call that follows it on the
// same thread with the same buffer: the file is scanned once, not twice
struct LineIndexCache {
    const uint8_t *buf = nullptr;
    int64_t nbytes = 0, nlines = 0;
    std::vector<int64_t> lend;
};
thread_local LineIndexCache g_lines;

// host threads of this process: the cores it may run on, divided by the ranks that share the box (torchrun sets
// LOCAL_WORLD_SIZE; eight ranks that each start 32 threads on a 32-core host only get in each other's way), at most 32;
// ITSX_HOST_THREADS overrides
int host_threads()
{
    static const int n = [] {
        if (const char *e = getenv("ITSX_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return std::min(v, 64); }
        int hw = 0;
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) hw = CPU_COUNT(&set);
#endif
        if (hw <= 0) hw = (int)std::thread::hardware_concurrency();
        if (hw <= 0) hw = 4;
        const char *w = getenv("LOCAL_WORLD_SIZE");
        if (!w) w = getenv("WORLD_SIZE");
        const int ranks = w ? std::max(1, atoi(w)) : 1;
        return std::max(1, std::min(hw / ranks, 32));
    }();
    return n;
}

int nthreads_for(int64_t work, int64_t grain)
{
    return (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), work / std::max<int64_t>(grain, 1)));
}

// f(a, b, t) over [0, n) cut into one range per thread.  What a worker throws (bad_alloc) is rethrown in the caller after
// all threads have been joined; a thread that cannot be started has its range run by the caller.
template <typename F> void parallel_for(int64_t n, int64_t grain, F f)
{
    const int nt = nthreads_for(n, grain);
    if (nt <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    auto guarded = [&](int64_t a, int64_t b, int t) {
        try { f(a, b, t); }
        catch (...) { std::lock_guard<std::mutex> g(mu); if (!err) err = std::current_exception(); }
    };
    const int64_t step = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int64_t a = t * step, b = std::min(n, a + step);
        if (a >= b) break;
        try { th.emplace_back([&guarded, a, b, t] { guarded(a, b, t); }); }
        catch (...) { guarded(a, b, t); }
    }
    for (auto &x : th) x.join();
    if (err) std::rethrow_exception(err);
}

}  // namespace

extern "C" {

const char *itsx_host_last_error(void) { return g_host_err.c_str(); }

// Index the records of a 4-line FASTQ buffer.  With cap == 0 only the record count is returned (arrays may be
// NULL); the filling call that follows it on the same thread for the same (unchanged) buffer reuses its line scan.  Returns the number of records, or ITSX_EFORMAT with the reason in itsx_host_last_error().
int64_t itsx_fastq_index(const uint8_t *buf, int64_t nbytes, int64_t cap, int64_t *t_off, int32_t *t_len,
                         int64_t *s_off, int32_t *s_len, int64_t *q_off)
{
    g_host_err.clear();
    if (nbytes < 0 || (nbytes && !buf)) { g_host_err = "fastq_index: null buffer"; return ITSX_EINVAL; }
    if (nbytes == 0) return 0;
    std::vector<int64_t> lend;
    int64_t nlines = 0;
    if (cap != 0 && g_lines.buf == buf && g_lines.nbytes == nbytes && !g_lines.lend.empty()) {
        lend.swap(g_lines.lend);
        nlines = g_lines.nlines;
        g_lines = LineIndexCache();
    } else {
        // pass 1: newline counts per chunk
        const int nt = nthreads_for(nbytes, 8 << 20);
        const int64_t step = (nbytes + nt - 1) / nt;
        std::vector<int64_t> cnt((size_t)nt + 1, 0);
        parallel_for(nt, 1, [&](int64_t a, int64_t b, int) {
            for (int64_t t = a; t < b; t++) {
                const uint8_t *p = buf + t * step, *e = buf + std::min(nbytes, (t + 1) * step);
                int64_t c = 0;
                while (p < e) {
                    const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
                    if (!q) break;
                    c++;
                    p = q + 1;
                }
                cnt[(size_t)t + 1] = c;
            }
        });
        for (int t = 0; t < nt; t++) cnt[(size_t)t + 1] += cnt[(size_t)t];
        nlines = cnt[(size_t)nt] + (buf[nbytes - 1] != '\n' ? 1 : 0);
        // pass 2: line ends
        lend.resize((size_t)nlines);
        parallel_for(nt, 1, [&](int64_t a, int64_t b, int) {
            for (int64_t t = a; t < b; t++) {
                const uint8_t *p = buf + t * step, *e = buf + std::min(nbytes, (t + 1) * step);
                int64_t k = cnt[(size_t)t];
                while (p < e) {
                    const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
                    if (!q) break;
                    lend[(size_t)k++] = q - buf;
                    p = q + 1;
                }
            }
        });
        if (buf[nbytes - 1] != '\n') lend[(size_t)nlines - 1] = nbytes;
    }
    // drop trailing blank lines
    auto lstart = [&](int64_t i) { return i == 0 ? (int64_t)0 : lend[(size_t)i - 1] + 1; };
    auto lstop = [&](int64_t i) {                     // exclusive end without '\r'
        int64_t e = lend[(size_t)i];
        if (e > lstart(i) && buf[e - 1] == '\r') e--;
        return e;
    };
    while (nlines > 0 && lstop(nlines - 1) == lstart(nlines - 1)) nlines--;
    if (nlines % 4 != 0) { g_host_err = "FASTQ is truncated or not in 4-line format"; return ITSX_EFORMAT; }
    const int64_t n = nlines / 4;
    if (cap == 0) {
        g_lines.buf = buf; g_lines.nbytes = nbytes; g_lines.nlines = nlines;
        g_lines.lend.swap(lend);
        return n;
    }
    if (cap < n) { g_host_err = "fastq_index: output arrays too small"; return ITSX_EINVAL; }
    std::atomic<int> bad(0);
    parallel_for(n, 1 << 14, [&](int64_t a, int64_t b, int) {
        for (int64_t r = a; r < b && !bad.load(std::memory_order_relaxed); r++) {
            const int64_t t0 = lstart(4 * r), t1 = lstop(4 * r);
            const int64_t s0 = lstart(4 * r + 1), s1 = lstop(4 * r + 1);
            const int64_t p0 = lstart(4 * r + 2), p1 = lstop(4 * r + 2);
            const int64_t q0 = lstart(4 * r + 3), q1 = lstop(4 * r + 3);
            if (t1 == t0 || buf[t0] != '@') { bad = 1; break; }
            if (p1 == p0 || buf[p0] != '+') { bad = 2; break; }
            if (q1 - q0 != s1 - s0) { bad = 3; break; }
            int64_t tl = t1 - (t0 + 1);
            while (tl > 0 && (buf[t0 + tl] == ' ' || buf[t0 + tl] == '\t')) tl--;
            if (p1 - p0 > 1) {                       // '+' line repeats the title: must be identical
                int64_t pl = p1 - (p0 + 1);
                while (pl > 0 && (buf[p0 + pl] == ' ' || buf[p0 + pl] == '\t')) pl--;
                if (pl != tl || memcmp(buf + p0 + 1, buf + t0 + 1, (size_t)tl) != 0) { bad = 4; break; }
            }
            for (int64_t i = q0; i < q1; i++)
                if (buf[i] < 33 || buf[i] > 126) { bad = 5; break; }
            if (bad.load(std::memory_order_relaxed)) break;
            t_off[r] = t0 + 1; t_len[r] = (int32_t)tl;
            s_off[r] = s0; s_len[r] = (int32_t)(s1 - s0);
            q_off[r] = q0;
        }
    });
    switch (bad.load()) {
        case 0: break;
        case 1: g_host_err = "Records in Fastq files should start with '@' character"; return ITSX_EFORMAT;
        case 2: g_host_err = "Expected '+' line in FASTQ record"; return ITSX_EFORMAT;
        case 3: g_host_err = "Lengths of sequence and quality values differs"; return ITSX_EFORMAT;
        case 4: g_host_err = "Sequence and quality captions differ."; return ITSX_EFORMAT;
        default: g_host_err = "Invalid character in quality string"; return ITSX_EFORMAT;
    }
    return n;
}

// Offset behind the last line of buf whose number is a multiple of four (0: fewer than four lines): newlines are counted on
// all threads, then the 0-3 lines of the unfinished record are stepped over from the end.
static int64_t fastq_cut_impl(const uint8_t *buf, int64_t nbytes)
{
    if (nbytes <= 0 || !buf) return 0;
    const int nt = nthreads_for(nbytes, 8 << 20);
    std::vector<int64_t> part((size_t)nt + 1, 0);
    parallel_for(nbytes, 8 << 20, [&](int64_t a, int64_t b, int t) {
        int64_t c = 0;
        for (int64_t i = a; i < b; i++) c += buf[i] == '\n';
        part[(size_t)t] = c;
    });
    int64_t lines = 0;
    for (int64_t c : part) lines += c;
    if (lines < 4) return 0;
    int64_t p = nbytes;                                   // behind the newline that ends line number `lines`
    while (p > 0 && buf[p - 1] != '\n') p--;
    for (int64_t k = lines % 4; k > 0; k--) {
        p--;                                              // onto that newline, then back to the one before it
        while (p > 0 && buf[p - 1] != '\n') p--;
    }
    return p;
}

// out_off[n+1] = prefix sums of len; out (if non-NULL) receives the segments packed back to back.
// Returns the total number of bytes.
int64_t itsx_bytes_gather(const uint8_t *buf, const int64_t *off, const int32_t *len, int64_t n, uint8_t *out,
                          int64_t *out_off)
{
    int64_t tot = 0;
    for (int64_t i = 0; i < n; i++) { out_off[i] = tot; tot += len[i]; }
    out_off[n] = tot;
    if (out)
        parallel_for(n, 1 << 14, [&](int64_t a, int64_t b, int) {
            for (int64_t i = a; i < b; i++) memcpy(out + out_off[i], buf + off[i], (size_t)len[i]);
        });
    return tot;
}

// FASTQ text of nkeep records: title of record keep_idx[t] of the indexed buffer, bases / qualities
// out_seq / out_qual [out_off[t], out_off[t+1]), optional constant prefix / suffix on both (--trim-ccs).
// dst == NULL: size query.  Returns the number of bytes.
int64_t itsx_fastq_format(const uint8_t *buf, const int64_t *t_off, const int32_t *t_len, const int32_t *keep_idx,
                          int64_t nkeep, const int64_t *out_off, const uint8_t *out_seq, const uint8_t *out_qual,
                          const uint8_t *pre_s, const uint8_t *pre_q, int32_t lp, const uint8_t *suf_s,
                          const uint8_t *suf_q, int32_t ls, uint8_t *dst)
{
    std::vector<int64_t> ro((size_t)nkeep + 1);
    int64_t tot = 0;
    for (int64_t t = 0; t < nkeep; t++) {
        ro[(size_t)t] = tot;
        const int64_t sl = out_off[t + 1] - out_off[t] + lp + ls;
        tot += 1 + t_len[keep_idx[t]] + 1 + sl + 1 + 2 + sl + 1;
    }
    ro[(size_t)nkeep] = tot;
    if (!dst) return tot;
    parallel_for(nkeep, 1 << 13, [&](int64_t a, int64_t b, int) {
        for (int64_t t = a; t < b; t++) {
            uint8_t *p = dst + ro[(size_t)t];
            const int32_t r = keep_idx[t];
            const int64_t sl = out_off[t + 1] - out_off[t];
            *p++ = '@';
            memcpy(p, buf + t_off[r], (size_t)t_len[r]); p += t_len[r];
            *p++ = '\n';
            if (lp) { memcpy(p, pre_s, (size_t)lp); p += lp; }
            memcpy(p, out_seq + out_off[t], (size_t)sl); p += sl;
            if (ls) { memcpy(p, suf_s, (size_t)ls); p += ls; }
            *p++ = '\n'; *p++ = '+'; *p++ = '\n';
            if (lp) { memcpy(p, pre_q, (size_t)lp); p += lp; }
            memcpy(p, out_qual + out_off[t], (size_t)sl); p += sl;
            if (ls) { memcpy(p, suf_q, (size_t)ls); p += ls; }
            *p++ = '\n';
        }
    });
    return tot;
}

// domtbl.txt rows in hmmsearch's --domtblout layout from the rows itsx_hits returns (what host.write_domtbl did row by row
// in Python: 25 us per row, minutes for a sample of 10^5 reads).  Labels come as byte strings with offsets (sequence i:
// seq_lab[seq_off[i] .. seq_off[i + 1]), no terminator).  E-values as HMMER computes them: Z = searched sequences, domZ =
// reported hits of the profile; alignment-derived columns are constants (include/itsx_b200.h).  Rows are formatted on the
// host threads into private strings and copied to dst in order.  Returns the bytes written, or ITSX_ELIMIT if cap is too
// small (nothing written then).
static int64_t domtbl_format_impl(const itsx_dom_row *rows, int64_t nrows, const uint8_t *seq_lab, const int64_t *seq_off,
                           const uint8_t *prof_lab, const int64_t *prof_off, const int32_t *prof_M, const int32_t *nreported,
                           double Z, uint8_t *dst, int64_t cap)
{
    if (nrows < 0 || (nrows && (!rows || !seq_lab || !seq_off || !prof_lab || !prof_off || !prof_M || !nreported)) || !dst)
        return ITSX_EINVAL;
    // domains per (profile, sequence) hit and the running index of every row within its hit, in row order
    std::vector<int32_t> k_in_hit((size_t)nrows), ndom((size_t)nrows);
    {
        // open addressing over (profile, sequence): slot = first row of the hit, count kept at that row's ndom
        size_t cap2 = 16;
        while (cap2 < (size_t)nrows * 2) cap2 <<= 1;
        std::vector<int64_t> slot(cap2, -1);
        std::vector<int64_t> head((size_t)nrows);
        auto key_of = [&](int64_t t) { return (uint64_t)(uint32_t)rows[t].prof << 32 | (uint32_t)rows[t].seq; };
        for (int64_t t = 0; t < nrows; t++) {
            const uint64_t k = key_of(t);
            if (t > 0 && key_of(t - 1) == k) {                    // (the rows of a hit normally follow each other)
                head[(size_t)t] = head[(size_t)t - 1];
            } else {
                size_t h = (size_t)((k * 0x9e3779b97f4a7c15ull) >> 20) & (cap2 - 1);
                while (slot[h] >= 0 && key_of(slot[h]) != k) h = (h + 1) & (cap2 - 1);
                if (slot[h] < 0) { slot[h] = t; ndom[(size_t)t] = 0; }
                head[(size_t)t] = slot[h];
            }
            k_in_hit[(size_t)t] = ++ndom[(size_t)head[(size_t)t]];
        }
        for (int64_t t = 0; t < nrows; t++) ndom[(size_t)t] = ndom[(size_t)head[(size_t)t]];
    }
    const int nt = nthreads_for(nrows, 1 << 14);
    std::vector<std::string> part((size_t)nt);
    // glibc's printf spends ~1 us per floating-point conversion; the three "%9.2g" go through std::to_chars (same digits:
    // the shortest-round-trip machinery rounds the exact value to 2 significant digits as printf does), "%6.1f" of a score
    // (a float, so 10 x it is exact in a double) through rint (ties to even, like printf), integers by hand
    auto pad_int = [](std::string &o, int64_t v, int width) {
        char t[24];
        int k = 0;
        const bool neg = v < 0;
        uint64_t u = neg ? (uint64_t)(-v) : (uint64_t)v;
        do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (neg) t[k++] = '-';
        for (int i = k; i < width; i++) o.push_back(' ');
        while (k) o.push_back(t[--k]);
    };
    auto put_g2 = [](std::string &o, double v) {                    // "%9.2g"
        char t[64];
        int n;
        if (std::isfinite(v)) {
            const auto r = std::to_chars(t, t + sizeof t, v, std::chars_format::general, 2);
            n = (int)(r.ptr - t);
        } else {
            n = snprintf(t, sizeof t, "%.2g", v);
        }
        for (int i = n; i < 9; i++) o.push_back(' ');
        o.append(t, (size_t)n);
    };
    auto put_f1 = [&](std::string &o, double v, int width) {        // "%<width>.1f" of a value whose tenfold is exact
        if (!std::isfinite(v) || fabs(v) > 1e15) { char t[400]; const int n = snprintf(t, sizeof t, "%*.1f", width, v); o.append(t, (size_t)n); return; }
        const double r = nearbyint(v * 10.0);                        // (round-to-nearest-even is the default mode)
        const bool neg = std::signbit(v);
        const uint64_t u = (uint64_t)fabs(r);
        char t[32];
        int k = 0;
        t[k++] = (char)('0' + u % 10);
        t[k++] = '.';
        uint64_t w = u / 10;
        do { t[k++] = (char)('0' + w % 10); w /= 10; } while (w);
        if (neg) t[k++] = '-';
        for (int i = k; i < width; i++) o.push_back(' ');
        while (k) o.push_back(t[--k]);
    };
    parallel_for(nrows, 1 << 14, [&](int64_t a, int64_t b, int t) {
        std::string &o = part[(size_t)t];
        o.reserve((size_t)(b - a) * 200);
        for (int64_t r = a; r < b; r++) {
            const itsx_dom_row &d = rows[r];
            const int p = d.prof;
            const double ev = exp(d.seq_lnP) * Z, cev = exp(d.lnP) * (double)nreported[p], iev = exp(d.lnP) * Z;
            const int ls = (int)(seq_off[d.seq + 1] - seq_off[d.seq]), lp = (int)(prof_off[p + 1] - prof_off[p]);
            o.append((const char *)seq_lab + seq_off[d.seq], (size_t)ls);
            if (ls < 20) o.append((size_t)(20 - ls), ' ');
            o += " -          ";                                   // " %-10s " of "-"
            pad_int(o, d.tlen, 5);
            o.push_back(' ');
            o.append((const char *)prof_lab + prof_off[p], (size_t)lp);
            if (lp < 20) o.append((size_t)(20 - lp), ' ');
            o += " -          ";
            pad_int(o, prof_M[p], 5); o.push_back(' ');
            put_g2(o, ev); o.push_back(' ');
            put_f1(o, (double)d.seq_score, 6);
            o += "   0.0 ";                                        // " %5.1f " of the bias column
            pad_int(o, k_in_hit[(size_t)r], 3); o.push_back(' ');
            pad_int(o, ndom[(size_t)r], 3); o.push_back(' ');
            put_g2(o, cev); o.push_back(' ');
            put_g2(o, iev); o.push_back(' ');
            put_f1(o, (double)d.bitscore, 6);
            o += "   0.0 ";
            pad_int(o, 1, 5); o.push_back(' ');
            pad_int(o, prof_M[p], 5); o.push_back(' ');
            pad_int(o, d.ienv, 5); o.push_back(' ');
            pad_int(o, d.jenv, 5); o.push_back(' ');
            pad_int(o, d.ienv, 5); o.push_back(' ');
            pad_int(o, d.jenv, 5);
            o += " 0.00 -\n";
        }
    });
    int64_t total = 0;
    for (const std::string &o : part) total += (int64_t)o.size();
    if (total > cap) return ITSX_ELIMIT;
    int64_t at = 0;
    for (const std::string &o : part) { memcpy(dst + at, o.data(), o.size()); at += (int64_t)o.size(); }
    return total;
}

// First whitespace-delimited token of every title (Biopython's record.id = title.split(None, 1)[0]): where it starts in buf
// and how long it is (0 for a title without one).
static int64_t fastq_labels_impl(const uint8_t *buf, const int64_t *t_off, const int64_t *t_len, int64_t n, int64_t *lab_off,
                          int32_t *lab_len)
{
    if (n < 0 || (n && (!buf || !t_off || !t_len || !lab_off || !lab_len))) return ITSX_EINVAL;
    auto ws = [](uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); };
    parallel_for(n, 1 << 16, [&](int64_t a, int64_t b, int) {
        for (int64_t i = a; i < b; i++) {
            const uint8_t *p = buf + t_off[i], *e = p + t_len[i];
            while (p < e && ws(*p)) p++;
            const uint8_t *q = p;
            while (q < e && !ws(*q)) q++;
            lab_off[i] = (int64_t)(p - buf);
            lab_len[i] = (int32_t)(q - p);
        }
    });
    return n;
}

namespace {
inline void put_int(std::string &o, int64_t v)
{
    char t[24];
    int k = 0;
    if (v < 0) { o.push_back('-'); v = -v; }
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) o.push_back(t[--k]);
}
}   // namespace

// uc.txt as `vsearch --uc` writes it for --fastx_uniques (SeqSample.py:104-119): per cluster, in
// `order`, an S row and the H rows of its other members in input order, then one C row per cluster.  rep[i] = read index of
// read i's representative, strand[i] != 0 -> '-' (may be NULL), len[i] = sequence length, order[c] = representative of
// cluster c, labels as itsx_fastq_labels returns them.  Returns the bytes written, ITSX_ELIMIT if cap is too small.
static int64_t uc_format_impl(const int32_t *rep, const uint8_t *strand, const int64_t *len, int64_t n, const int64_t *order, int64_t nc,
                       const uint8_t *buf, const int64_t *lab_off, const int32_t *lab_len, uint8_t *dst, int64_t cap)
{
    if (n < 0 || nc < 0 || (n && (!rep || !len || !buf || !lab_off || !lab_len)) || (nc && !order) || !dst) return ITSX_EINVAL;
    std::vector<int64_t> cl((size_t)n, -1), cnt((size_t)nc + 1, 0);
    for (int64_t c = 0; c < nc; c++) { if (order[c] < 0 || order[c] >= n) return ITSX_EINVAL; cl[(size_t)order[c]] = c; }
    for (int64_t i = 0; i < n; i++) {
        if (rep[i] < 0 || rep[i] >= n || cl[(size_t)rep[i]] < 0) return ITSX_EINVAL;
        if (rep[i] != i) cnt[(size_t)cl[(size_t)rep[i]] + 1]++;
    }
    for (int64_t c = 0; c < nc; c++) cnt[(size_t)c + 1] += cnt[(size_t)c];
    std::vector<int64_t> mem((size_t)cnt[(size_t)nc]), at(cnt.begin(), cnt.end() - 1);
    for (int64_t i = 0; i < n; i++) if (rep[i] != i) mem[(size_t)at[(size_t)cl[(size_t)rep[i]]]++] = i;
    std::string o;
    o.reserve((size_t)n * 64);
    auto label = [&](int64_t i) { o.append((const char *)buf + lab_off[i], (size_t)lab_len[i]); };
    for (int64_t c = 0; c < nc; c++) {
        const int64_t r = order[c];
        o += "S\t"; put_int(o, c); o.push_back('\t'); put_int(o, len[r]); o += "\t*\t*\t*\t*\t*\t"; label(r); o += "\t*\n";
        for (int64_t k = cnt[(size_t)c]; k < cnt[(size_t)c + 1]; k++) {
            const int64_t i = mem[(size_t)k];
            o += "H\t"; put_int(o, c); o.push_back('\t'); put_int(o, len[i]); o += "\t100.0\t";
            o.push_back(strand && strand[i] ? '-' : '+');
            o += "\t0\t0\t*\t"; label(i); o.push_back('\t'); label(r); o.push_back('\n');
        }
    }
    for (int64_t c = 0; c < nc; c++) {
        o += "C\t"; put_int(o, c); o.push_back('\t'); put_int(o, 1 + cnt[(size_t)c + 1] - cnt[(size_t)c]);
        o += "\t*\t*\t*\t*\t*\t"; label(order[c]); o += "\t*\n";
    }
    if ((int64_t)o.size() > cap) return ITSX_ELIMIT;
    memcpy(dst, o.data(), o.size());
    return (int64_t)o.size();
}

// rep.fa as `vsearch --fastaout` writes it: '>label', then the sequence as stored, wrapped at `width` columns.
static int64_t repfa_format_impl(const uint8_t *buf, const int64_t *s_off, const int64_t *s_len, const int64_t *lab_off,
                          const int32_t *lab_len, const int64_t *order, int64_t nc, int32_t width, uint8_t *dst, int64_t cap)
{
    if (nc < 0 || width <= 0 || (nc && (!buf || !s_off || !s_len || !lab_off || !lab_len || !order)) || !dst) return ITSX_EINVAL;
    std::vector<int64_t> at((size_t)nc + 1, 0);
    for (int64_t c = 0; c < nc; c++) {
        const int64_t r = order[c], L = s_len[r];
        at[(size_t)c + 1] = at[(size_t)c] + 1 + lab_len[r] + 1 + L + (L + width - 1) / width;
    }
    if (at[(size_t)nc] > cap) return ITSX_ELIMIT;
    parallel_for(nc, 1 << 12, [&](int64_t a, int64_t b, int) {
        for (int64_t c = a; c < b; c++) {
            const int64_t r = order[c];
            uint8_t *o = dst + at[(size_t)c];
            *o++ = '>';
            memcpy(o, buf + lab_off[r], (size_t)lab_len[r]); o += lab_len[r];
            *o++ = '\n';
            const uint8_t *q = buf + s_off[r];
            for (int64_t j = 0; j < s_len[r]; j += width) {
                const int64_t k = std::min<int64_t>(width, s_len[r] - j);
                memcpy(o, q + j, (size_t)k); o += k;
                *o++ = '\n';
            }
        }
    });
    return at[(size_t)nc];
}

// ---- no exception crosses the C ABI: the formatters above allocate (strings, vectors) and start threads ----
int64_t itsx_fastq_cut(const uint8_t *buf, int64_t nbytes)
{
    try { return fastq_cut_impl(buf, nbytes); }
    catch (const std::bad_alloc &) { return ITSX_ELIMIT; }
    catch (...) { return ITSX_EINVAL; }
}

int64_t itsx_domtbl_format(const itsx_dom_row *rows, int64_t nrows, const uint8_t *seq_lab, const int64_t *seq_off,
                           const uint8_t *prof_lab, const int64_t *prof_off, const int32_t *prof_M, const int32_t *nreported,
                           double Z, uint8_t *dst, int64_t cap)
{
    try { return domtbl_format_impl(rows, nrows, seq_lab, seq_off, prof_lab, prof_off, prof_M, nreported, Z, dst, cap); }
    catch (const std::bad_alloc &) { return ITSX_ELIMIT; }
    catch (...) { return ITSX_EINVAL; }
}

int64_t itsx_fastq_labels(const uint8_t *buf, const int64_t *t_off, const int64_t *t_len, int64_t n, int64_t *lab_off,
                          int32_t *lab_len)
{
    try { return fastq_labels_impl(buf, t_off, t_len, n, lab_off, lab_len); }
    catch (const std::bad_alloc &) { return ITSX_ELIMIT; }
    catch (...) { return ITSX_EINVAL; }
}

int64_t itsx_uc_format(const int32_t *rep, const uint8_t *strand, const int64_t *len, int64_t n, const int64_t *order, int64_t nc,
                       const uint8_t *buf, const int64_t *lab_off, const int32_t *lab_len, uint8_t *dst, int64_t cap)
{
    try { return uc_format_impl(rep, strand, len, n, order, nc, buf, lab_off, lab_len, dst, cap); }
    catch (const std::bad_alloc &) { return ITSX_ELIMIT; }
    catch (...) { return ITSX_EINVAL; }
}

int64_t itsx_repfa_format(const uint8_t *buf, const int64_t *s_off, const int64_t *s_len, const int64_t *lab_off,
                          const int32_t *lab_len, const int64_t *order, int64_t nc, int32_t width, uint8_t *dst, int64_t cap)
{
    try { return repfa_format_impl(buf, s_off, s_len, lab_off, lab_len, order, nc, width, dst, cap); }
    catch (const std::bad_alloc &) { return ITSX_ELIMIT; }
    catch (...) { return ITSX_EINVAL; }
}

}  // extern "C"
