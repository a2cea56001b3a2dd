// inflate_host.cpp -- gzip reader of libitsx_b200 (host C++): the input side of the FASTQ path.
//
// Replaces gzip.open(path, "rt") under SeqIO.parse (itsxpress/SeqSample.py:742-752, 767-788; main.py:295-330): a
// Casava / Illumina .fastq.gz is ONE deflate stream, so it inflates on one core, and with the outputs compressed on the
// GPU (deflate.cu) that core is what bounds a multi-sample artifact.  zlib's inflate runs at 110-250 MB/s of FASTQ per
// core; this decoder keeps a 64-bit bit buffer that is refilled without branches, decodes through one 11-bit table
// look-up per literal / length symbol (8-bit for distances; longer codes through second-level tables), copies matches in
// 8-byte words and checks CRC-32 and ISIZE of every member (slice-by-8; for a large member on a second thread that
// follows the decoder through the output).
// Every stream feature of RFC 1951 / 1952 is handled (stored, fixed and dynamic blocks, multi-member files, header
// extras); anything malformed returns ITSX_EFORMAT and the Python layer hands the file to the gzip module, which raises
// what the reference would have raised.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "deflate_core.h"
#include "itsx_internal.h"

namespace {

constexpr int LIT_TB = 11, DIST_TB = 8;
constexpr int LIT_TAB = 1 << 13, DIST_TAB = 1 << 11;          // primary + second-level tables (generous)
enum : uint32_t { T_LIT = 0, T_LEN = 1, T_EOB = 2, T_SUB = 3, T_BAD = 4 };
// entry: value << 16 | type << 12 | extra << 8 | bits
inline uint32_t mk(uint32_t value, uint32_t type, uint32_t extra, uint32_t bits) { return value << 16 | type << 12 | extra << 8 | bits; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// Huffman decoding table from code lengths (canonical codes, looked up LSB first).  kind 0: literal / length alphabet,
// kind 1: distances, kind 2: code-length alphabet (plain values).  false = over-subscribed or (not allowed) incomplete.
bool build_table(const uint8_t *lens, int n, int tb, int kind, uint32_t *tab, int tabsize)
{
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1, used = 0;
    for (int l = 1; l < 16; l++) { left = (left << 1) - count[l]; used += count[l]; if (left < 0) return false; }
    if (used == 0) {                       // no codes at all: every look-up is an error (legal for distances if unused)
        for (int i = 0; i < (1 << tb); i++) tab[i] = mk(0, T_BAD, 0, 1);
        return kind == 1;
    }
    if (left > 0 && !(kind == 1 && used == 1)) return false;      // incomplete: only a single distance code may be
    int next[16];
    next[0] = 0;
    { int c = 0; for (int l = 1; l < 16; l++) { c = (c + count[l - 1]) << 1; next[l] = c; } }
    const int prim = 1 << tb;
    for (int i = 0; i < prim; i++) tab[i] = mk(0, T_BAD, 0, 1);
    uint8_t sublen[1 << LIT_TB] = {0};
    uint16_t rev[320];
    auto entry = [&](int sym, int bits) -> uint32_t {
        if (kind == 2) return mk((uint32_t)sym, T_LIT, 0, (uint32_t)bits);
        if (kind == 1) return sym < 30 ? mk(DIST_BASE[sym], T_LEN, DIST_EXTRA[sym], (uint32_t)bits) : mk(0, T_BAD, 0, (uint32_t)bits);
        if (sym < 256) return mk((uint32_t)sym, T_LIT, 0, (uint32_t)bits);
        if (sym == 256) return mk(0, T_EOB, 0, (uint32_t)bits);
        return sym < 286 ? mk(LEN_BASE[sym - 257], T_LEN, LEN_EXTRA[sym - 257], (uint32_t)bits) : mk(0, T_BAD, 0, (uint32_t)bits);
    };
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t c = (uint32_t)next[l]++, r = 0;
        for (int b = 0; b < l; b++) { r = r << 1 | (c & 1u); c >>= 1; }
        rev[s] = (uint16_t)r;
        if (l <= tb) {
            const uint32_t e = entry(s, l);
            for (uint32_t i = r; i < (uint32_t)prim; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t low = r & (uint32_t)(prim - 1);
            sublen[low] = std::max<uint8_t>(sublen[low], (uint8_t)(l - tb));
        }
    }
    int free_at = prim;
    for (int low = 0; low < prim; low++) {
        if (!sublen[low]) continue;
        if (free_at + (1 << sublen[low]) > tabsize) return false;
        tab[low] = mk((uint32_t)free_at, T_SUB, sublen[low], (uint32_t)tb);
        for (int i = 0; i < (1 << sublen[low]); i++) tab[free_at + i] = mk(0, T_BAD, 0, 1);
        free_at += 1 << sublen[low];
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (l <= tb) continue;
        const uint32_t r = rev[s], low = r & (uint32_t)(prim - 1), start = tab[low] >> 16, sb = (tab[low] >> 8) & 15u;
        const uint32_t e = entry(s, l - tb);
        for (uint32_t i = r >> tb; i < (1u << sb); i += 1u << (l - tb)) tab[start + i] = e;
    }
    return true;
}

struct Inflater {
    const uint8_t *in, *in_end;
    uint8_t *out, *out_beg, *out_end;
    uint64_t bb = 0;          // bit buffer, LSB first
    int bc = 0;               // valid bits
    uint32_t lit[LIT_TAB], dist[DIST_TAB];
    bool overrun = false;     // input ended inside the stream
    std::atomic<int64_t> *progress = nullptr;   // output produced so far (published at block ends for the CRC follower)

    // byte-wise refill for the ends of the input: behind the last byte zeros are fed and COUNTED (in moves on), so that
    // the caller can tell how many of them were really used; more than a buffer's worth means the stream is truncated
    inline void refill_safe()
    {
        while (bc <= 56) {
            if (in < in_end) bb |= (uint64_t)*in << bc;
            else if (in >= in_end + 8) overrun = true;
            in++;
            bc += 8;
        }
    }
    inline uint32_t take(int n) { const uint32_t v = (uint32_t)(bb & ((1ull << n) - 1)); bb >>= n; bc -= n; return v; }
};

// one deflate stream at s.in -> s.out.  0 = ok (s.in behind the last byte of the stream), 1 = output full, < 0 = error
int inflate_stream(Inflater &s)
{
    for (;;) {
        s.refill_safe();
        const uint32_t last = s.take(1), type = s.take(2);
        if (type == 0) {
            s.take(s.bc & 7);                            // to the byte boundary
            // bytes still in the bit buffer belong to the input
            const uint8_t *p = s.in - s.bc / 8;
            s.bb = 0; s.bc = 0;
            if (p + 4 > s.in_end) return ITSX_EFORMAT;
            const uint32_t len = p[0] | p[1] << 8, nlen = p[2] | p[3] << 8;
            if ((len ^ nlen) != 0xffffu) return ITSX_EFORMAT;
            p += 4;
            if (p + len > s.in_end) return ITSX_EFORMAT;
            if (s.out + len > s.out_end) return 1;
            memcpy(s.out, p, len);
            s.out += len;
            s.in = p + len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[320];
            int hlit, hdist;
            if (type == 1) {
                hlit = 288; hdist = 32;                  // (codes 286 / 287 and 30 / 31 complete the fixed codes; invalid in data)
                for (int i = 0; i < 144; i++) lens[i] = 8;
                for (int i = 144; i < 256; i++) lens[i] = 9;
                for (int i = 256; i < 280; i++) lens[i] = 7;
                for (int i = 280; i < 288; i++) lens[i] = 8;
                for (int i = 0; i < 32; i++) lens[288 + i] = 5;
            } else {
                hlit = (int)s.take(5) + 257; hdist = (int)s.take(5) + 1;
                const int hclen = (int)s.take(4) + 4;
                if (hlit > 286 || hdist > 30) return ITSX_EFORMAT;
                static const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (int i = 0; i < hclen; i++) { s.refill_safe(); cl[perm[i]] = (uint8_t)s.take(3); }
                uint32_t cltab[1 << 7];
                if (!build_table(cl, 19, 7, 2, cltab, 1 << 7)) return ITSX_EFORMAT;
                int i = 0;
                while (i < hlit + hdist) {
                    s.refill_safe();
                    if (s.overrun) return ITSX_EFORMAT;
                    const uint32_t e = cltab[s.bb & 127u];
                    if (((e >> 12) & 15u) == T_BAD) return ITSX_EFORMAT;
                    s.take((int)(e & 255u));
                    const int sym = (int)(e >> 16);
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    int rep, val = 0;
                    if (sym == 16) { if (i == 0) return ITSX_EFORMAT; val = lens[i - 1]; rep = 3 + (int)s.take(2); }
                    else if (sym == 17) rep = 3 + (int)s.take(3);
                    else rep = 11 + (int)s.take(7);
                    if (i + rep > hlit + hdist) return ITSX_EFORMAT;
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (lens[256] == 0) return ITSX_EFORMAT;
                memmove(lens + 288, lens + hlit, (size_t)hdist);       // distances behind a fixed offset
                for (int k = hlit; k < 288; k++) lens[k] = 0;
            }
            if (!build_table(lens, type == 1 ? 288 : hlit, LIT_TB, 0, s.lit, LIT_TAB)) return ITSX_EFORMAT;
            if (!build_table(lens + 288, hdist, DIST_TB, 1, s.dist, DIST_TAB)) return ITSX_EFORMAT;
            // ---- symbols ----
            for (;;) {
                // fast loop: at least 8 readable input bytes beyond the refill and room for the longest match.  The
                // state lives in locals (the byte stores into the output could alias the struct otherwise).
                {
                    const uint8_t *in = s.in, *const in_fast = s.in_end - 16;
                    uint8_t *out = s.out, *const out_fast = s.out_end - (258 + 16);
                    uint8_t *const out_beg = s.out_beg;
                    uint64_t bb = s.bb;
                    int bc = s.bc;
                    const uint32_t *const lit = s.lit, *const dist = s.dist;
                    int stop = 0;                        // 1: end of block, 2: bad stream
                    while (in <= in_fast && out <= out_fast) {
                        uint64_t w;
                        memcpy(&w, in, 8);
                        bb |= w << bc;
                        in += (63 - bc) >> 3;
                        bc |= 56;
                        uint32_t e = lit[bb & ((1u << LIT_TB) - 1)];
                        if (__builtin_expect(((e >> 12) & 15u) == T_SUB, 0)) {
                            bb >>= LIT_TB; bc -= LIT_TB;
                            e = lit[(e >> 16) + (bb & ((1u << ((e >> 8) & 15u)) - 1))];
                        }
                        bb >>= (e & 255u); bc -= (int)(e & 255u);
                        const uint32_t ty = (e >> 12) & 15u;
                        if (ty == T_LIT) {
                            *out++ = (uint8_t)(e >> 16);
                            // a second and third literal usually fit what is left in the buffer
                            e = lit[bb & ((1u << LIT_TB) - 1)];
                            if (((e >> 12) & 15u) != T_LIT) continue;
                            bb >>= (e & 255u); bc -= (int)(e & 255u);
                            *out++ = (uint8_t)(e >> 16);
                            e = lit[bb & ((1u << LIT_TB) - 1)];
                            if (((e >> 12) & 15u) != T_LIT) continue;
                            bb >>= (e & 255u); bc -= (int)(e & 255u);
                            *out++ = (uint8_t)(e >> 16);
                            continue;
                        }
                        if (__builtin_expect(ty != T_LEN, 0)) { stop = ty == T_EOB ? 1 : 2; break; }
                        const uint32_t xl = (e >> 8) & 15u;
                        const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << xl) - 1));
                        bb >>= xl; bc -= (int)xl;
                        uint32_t d = dist[bb & ((1u << DIST_TB) - 1)];
                        if (__builtin_expect(((d >> 12) & 15u) == T_SUB, 0)) {
                            bb >>= DIST_TB; bc -= DIST_TB;
                            d = dist[(d >> 16) + (bb & ((1u << ((d >> 8) & 15u)) - 1))];
                        }
                        if (__builtin_expect(((d >> 12) & 15u) != T_LEN, 0)) { stop = 2; break; }
                        bb >>= (d & 255u); bc -= (int)(d & 255u);
                        const uint32_t xd = (d >> 8) & 15u;
                        const uint32_t distv = (d >> 16) + (uint32_t)(bb & ((1u << xd) - 1));
                        bb >>= xd; bc -= (int)xd;
                        if (__builtin_expect(distv > (uint32_t)(out - out_beg), 0)) { stop = 2; break; }
                        uint8_t *o = out;
                        const uint8_t *m = o - distv;
                        out += len;
                        if (distv >= 8) {
                            uint64_t v;
                            memcpy(&v, m, 8); memcpy(o, &v, 8);
                            memcpy(&v, m + 8, 8); memcpy(o + 8, &v, 8);
                            for (uint32_t k = 16; k < len; k += 8) { memcpy(&v, m + k, 8); memcpy(o + k, &v, 8); }
                        } else if (distv == 1) {
                            memset(o, *m, len);
                        } else {
                            for (uint32_t k = 0; k < len; k++) o[k] = m[k];
                        }
                    }
                    s.in = in; s.out = out; s.bb = bb; s.bc = bc;
                    if (stop == 1) goto block_done;
                    if (stop == 2) return ITSX_EFORMAT;
                }
                // careful path near the ends of input / output: one symbol
                s.refill_safe();
                if (s.overrun) return ITSX_EFORMAT;
                uint32_t e = s.lit[s.bb & ((1u << LIT_TB) - 1)];
                if (((e >> 12) & 15u) == T_SUB) { s.take(LIT_TB); e = s.lit[(e >> 16) + (s.bb & ((1u << ((e >> 8) & 15u)) - 1))]; }
                s.take((int)(e & 255u));
                const uint32_t ty = (e >> 12) & 15u;
                if (ty == T_LIT) {
                    if (s.out >= s.out_end) return 1;
                    *s.out++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (ty == T_EOB) break;
                if (ty != T_LEN) return ITSX_EFORMAT;
                const uint32_t len = (e >> 16) + s.take((int)((e >> 8) & 15u));
                s.refill_safe();
                uint32_t d = s.dist[s.bb & ((1u << DIST_TB) - 1)];
                if (((d >> 12) & 15u) == T_SUB) { s.take(DIST_TB); d = s.dist[(d >> 16) + (s.bb & ((1u << ((d >> 8) & 15u)) - 1))]; }
                if (((d >> 12) & 15u) != T_LEN) return ITSX_EFORMAT;
                s.take((int)(d & 255u));
                const uint32_t distv = (d >> 16) + s.take((int)((d >> 8) & 15u));
                if (distv > (uint32_t)(s.out - s.out_beg)) return ITSX_EFORMAT;
                if (s.out + len > s.out_end) return 1;
                for (uint32_t k = 0; k < len; k++) { *s.out = *(s.out - distv); s.out++; }
            }
        block_done:;
        } else {
            return ITSX_EFORMAT;
        }
        if (s.overrun) return ITSX_EFORMAT;
        if (s.progress) s.progress->store((int64_t)(s.out - s.out_beg), std::memory_order_release);
        if (last) break;
    }
    // give back the whole bytes still in the bit buffer
    s.in -= s.bc / 8;
    s.bb = 0; s.bc = 0;
    if (s.in > s.in_end) return ITSX_EFORMAT;
    return 0;
}

// ---- CRC-32, slice-by-8, over the host threads for large buffers ----
uint32_t g_crc[8][256];
bool crc_fill()
{
    for (uint32_t i = 0; i < 256; i++) g_crc[0][i] = dfl_crc_table_entry(i);
    for (uint32_t i = 0; i < 256; i++)
        for (int k = 1; k < 8; k++) g_crc[k][i] = g_crc[0][g_crc[k - 1][i] & 0xffu] ^ (g_crc[k - 1][i] >> 8);
    return true;
}
void crc_init()
{
    static const bool ready = crc_fill();       // thread-safe one-time initialisation
    (void)ready;
}
uint32_t crc_update(uint32_t c, const uint8_t *p, size_t n)      // raw register (no pre / post inversion)
{
    while (n && ((uintptr_t)p & 7)) { c = g_crc[0][(c ^ *p++) & 0xffu] ^ (c >> 8); n--; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        const uint32_t lo = (uint32_t)w ^ c, hi = (uint32_t)(w >> 32);
        c = g_crc[7][lo & 0xffu] ^ g_crc[6][(lo >> 8) & 0xffu] ^ g_crc[5][(lo >> 16) & 0xffu] ^ g_crc[4][lo >> 24] ^
            g_crc[3][hi & 0xffu] ^ g_crc[2][(hi >> 8) & 0xffu] ^ g_crc[1][(hi >> 16) & 0xffu] ^ g_crc[0][hi >> 24];
        p += 8; n -= 8;
    }
    while (n--) c = g_crc[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
    return c;
}
uint32_t crc32_buf(const uint8_t *p, size_t n)
{
    crc_init();
    return crc_update(0xffffffffu, p, n) ^ 0xffffffffu;
}

// CRC-32 of a member WHILE it is being inflated: a second thread follows the decoder through the output (the decoder
// publishes its position at every deflate block end), so the check costs no time of the inflating core.
struct CrcFollower {
    const uint8_t *base;
    std::atomic<int64_t> produced{0};
    std::atomic<bool> done{false};
    uint32_t reg = 0xffffffffu;
    std::thread th;
    void start(const uint8_t *b)
    {
        base = b;
        crc_init();
        th = std::thread([this] {
            int64_t at = 0;
            for (;;) {
                const bool fin = done.load(std::memory_order_acquire);
                const int64_t have = produced.load(std::memory_order_acquire);
                if (have - at >= (1 << 20) || (fin && have > at)) {
                    reg = crc_update(reg, base + at, (size_t)(have - at));
                    at = have;
                } else if (fin) {
                    break;
                } else {
                    std::this_thread::yield();
                }
            }
        });
    }
    uint32_t finish(int64_t total)
    {
        produced.store(total, std::memory_order_release);
        done.store(true, std::memory_order_release);
        th.join();
        return reg ^ 0xffffffffu;
    }
};

}   // namespace

extern "C" {

// A gzip file (one or more members) -> dst.  Whole members are decoded while they fit into cap bytes.
//   *in_used / *out_used: input consumed / output produced (always at a member boundary)
//   returns 0: the whole input was consumed; 1: the next member does not fit (grow dst and call again with
//   src + *in_used; if *in_used == 0 the FIRST member alone is larger than cap); ITSX_EFORMAT: not a valid gzip stream
int itsx_gunzip(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap, int64_t *in_used, int64_t *out_used)
{
    if (!src || n < 0 || (!dst && cap > 0) || !in_used || !out_used) return ITSX_EINVAL;
    *in_used = 0; *out_used = 0;
    static thread_local Inflater *s = nullptr;
    if (!s) s = new Inflater();
    const uint8_t *p = src, *end = src + n;
    uint8_t *o = dst;
    while (p < end) {
        // trailing zero padding after the last member is tolerated like gzip does
        if (*p == 0) { const uint8_t *q = p; while (q < end && *q == 0) q++; if (q == end && p != src) { p = end; break; } }
        if (end - p < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return ITSX_EFORMAT;
        const uint8_t flg = p[3];
        const uint8_t *q = p + 10;
        if (flg & 4) { if (end - q < 2) return ITSX_EFORMAT; const int xl = q[0] | q[1] << 8; q += 2; if (end - q < xl) return ITSX_EFORMAT; q += xl; }
        if (flg & 8) { while (q < end && *q) q++; if (q >= end) return ITSX_EFORMAT; q++; }
        if (flg & 16) { while (q < end && *q) q++; if (q >= end) return ITSX_EFORMAT; q++; }
        if (flg & 2) { if (end - q < 2) return ITSX_EFORMAT; q += 2; }
        s->in = q; s->in_end = end; s->out = o; s->out_beg = o; s->out_end = dst + cap;
        s->bb = 0; s->bc = 0; s->overrun = false;
        // members of more than a few MB of input get a CRC follower thread; small ones are checked afterwards
        const bool follow = end - q > (4 << 20);
        CrcFollower fol;
        s->progress = follow ? &fol.produced : nullptr;
        if (follow) fol.start(o);
        const int rc = inflate_stream(*s);
        s->progress = nullptr;
        const size_t produced = (size_t)(s->out - o);
        const uint32_t crc_have = follow ? fol.finish(rc == 0 ? (int64_t)produced : 0) : 0u;
        if (rc == 1) { *in_used = p - src; *out_used = o - dst; return 1; }
        if (rc < 0) return rc;
        if (end - s->in < 8) return ITSX_EFORMAT;
        const uint8_t *t = s->in;
        const uint32_t crc = t[0] | t[1] << 8 | t[2] << 16 | (uint32_t)t[3] << 24;
        const uint32_t isize = t[4] | t[5] << 8 | t[6] << 16 | (uint32_t)t[7] << 24;
        if ((uint32_t)produced != isize || (follow ? crc_have : crc32_buf(o, produced)) != crc) return ITSX_EFORMAT;
        o = s->out;
        p = t + 8;
    }
    *in_used = p - src; *out_used = o - dst;
    return 0;
}

}   // extern "C"
