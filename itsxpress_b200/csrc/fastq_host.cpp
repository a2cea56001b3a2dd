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
#include <cstring>
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

template <typename F> void parallel_for(int64_t n, int64_t grain, F f)
{
    const int nt = nthreads_for(n, grain);
    if (nt <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    const int64_t step = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int64_t a = t * step, b = std::min(n, a + step);
        if (a >= b) break;
        th.emplace_back([=] { f(a, b, t); });
    }
    for (auto &x : th) x.join();
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
int64_t itsx_fastq_cut(const uint8_t *buf, int64_t nbytes)
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

}  // extern "C"
