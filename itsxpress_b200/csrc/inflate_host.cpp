// inflate_host.cpp -- gzip reader of libitsx_b200 (host C++): the input side of the FASTQ path.
//
// Replaces gzip.open(path, "rt") under SeqIO.parse (itsxpress/SeqSample.py:742-752, 767-788; main.py:295-330): a
// Casava / Illumina .fastq.gz is ONE deflate stream, so the reference inflates it on one core, and with the outputs
// compressed on the GPU (deflate.cu) that core is what bounds a sample.  Two things are done about it:
//
//  * a decoder that is faster per core than zlib's: 64-bit bit buffer refilled without branches, one 11-bit table look-up
//    per literal / length symbol (8-bit for distances; longer codes through second-level tables), table entries that
//    carry TWO literals when both codes fit into the look-up (FASTQ is literal-dominated), matches copied in 8-byte
//    words, BMI2 shifts when the CPU has them;
//  * one deflate stream inflated on several cores (two passes, as pugz / rapidgzip do): the input is cut into chunks, every
//    chunk but the first looks for the next deflate block header at bit granularity (strict header validation) or the
//    next gzip member header, decodes
//    from there into 16-bit symbols in which a back-reference that reaches in front of the chunk becomes a MARKER
//    (0x8000 + position in the unknown 32 KB window), and stops at the block boundary where the next chunk started.  The
//    chunks are then stitched in order (the end bit of one must be the start bit of the next; anything else is decoded
//    again from where the predecessor really ended), the 32 KB windows are resolved front to back, and every chunk is
//    translated to bytes in parallel.  CRC-32 and ISIZE of every member are checked (pieces combined with x^n mod P).
//
// Every stream feature of RFC 1951 / 1952 is handled (stored, fixed and dynamic blocks, multi-member files, header
// extras, zero padding behind the last member); anything malformed is ITSX_EFORMAT and the Python layer hands the file to
// the gzip module, which raises what the reference would have raised.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "deflate_core.h"
#include "itsx_internal.h"

namespace {

constexpr int LIT_TB = 11, DIST_TB = 8;
constexpr int LIT_TAB = 1 << 13, DIST_TAB = 1 << 11;          // primary + second-level tables (generous)
constexpr uint32_t LMASK = (1u << LIT_TB) - 1, DMASK = (1u << DIST_TB) - 1;
constexpr int WSIZE = 32768;
enum : uint32_t { T_LIT = 0, T_LEN = 1, T_EOB = 2, T_SUB = 3, T_BAD = 4 };
enum { R_END = 0, R_FULL = 1, R_STOP = 2 };                   // results of run(); negative = ITSX_E*
// entry: value << 16 | type << 12 | extra << 8 | bits.  T_LIT: extra = number of literals (1 or 2), value = lit0 | lit1 << 8.
// T_LEN (lengths and distances): value = base, bits = code length + extra bits, so that ONE shift consumes both and the
// extra bits are cut out of a copy of the bit buffer off the critical path
inline uint32_t mk(uint32_t value, uint32_t type, uint32_t extra, uint32_t bits) { return value << 16 | type << 12 | extra << 8 | bits; }
#define E_TYPE(e) (((e) >> 12) & 15u)
#define E_XTRA(e) (((e) >> 8) & 15u)
#define E_BITS(e) ((e) & 255u)

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CL_PERM[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

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
        if (kind == 2) return mk((uint32_t)sym, T_LIT, 1, (uint32_t)bits);
        if (kind == 1) return sym < 30 ? mk(DIST_BASE[sym], T_LEN, DIST_EXTRA[sym], (uint32_t)bits + DIST_EXTRA[sym]) : mk(0, T_BAD, 0, (uint32_t)bits);
        if (sym < 256) return mk((uint32_t)sym, T_LIT, 1, (uint32_t)bits);
        if (sym == 256) return mk(0, T_EOB, 0, (uint32_t)bits);
        return sym < 286 ? mk(LEN_BASE[sym - 257], T_LEN, LEN_EXTRA[sym - 257], (uint32_t)bits + LEN_EXTRA[sym - 257]) : mk(0, T_BAD, 0, (uint32_t)bits);
    };
    bool any_long = false;
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
            any_long = true;
        }
    }
    if (kind == 0) {
        // two literals per look-up where both codes fit into the index: the second one is read off the (still
        // single-literal) entry of the index shifted down by the first code's length
        uint32_t one[1 << LIT_TB];
        memcpy(one, tab, sizeof(uint32_t) * (size_t)prim);
        for (int i = 0; i < prim; i++) {
            const uint32_t e = one[i];
            if (E_TYPE(e) != T_LIT) continue;
            const uint32_t b1 = E_BITS(e);
            if ((int)b1 >= tb) continue;
            const uint32_t e2 = one[(uint32_t)i >> b1];
            if (E_TYPE(e2) != T_LIT || b1 + E_BITS(e2) > (uint32_t)tb) continue;
            if (sublen[(uint32_t)i >> b1]) continue;                // (that index leads to a second-level table)
            tab[i] = mk((e >> 16) | (e2 >> 16) << 8, T_LIT, 2, b1 + E_BITS(e2));
        }
    }
    if (!any_long) return true;
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
        const uint32_t r = rev[s], low = r & (uint32_t)(prim - 1), start = tab[low] >> 16, sb = E_XTRA(tab[low]);
        const uint32_t e = entry(s, l - tb);
        for (uint32_t i = r >> tb; i < (1u << sb); i += 1u << (l - tb)) tab[start + i] = e;
    }
    return true;
}

struct FixedTables {
    uint32_t lit[LIT_TAB], dist[DIST_TAB];
    FixedTables()
    {
        uint8_t lens[320];
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;          // (codes 286 / 287 and 30 / 31 complete the fixed codes; invalid in data)
        for (int i = 0; i < 32; i++) lens[288 + i] = 5;
        build_table(lens, 288, LIT_TB, 0, lit, LIT_TAB);
        build_table(lens + 288, 32, DIST_TB, 1, dist, DIST_TAB);
    }
};
const FixedTables &fixed_tables()
{
    static const FixedTables t;          // thread-safe one-time initialisation
    return t;
}

// ---- one deflate decoder: resumable at every output-full condition, stoppable at block boundaries ----
struct Dec {
    const uint8_t *in_beg = nullptr, *in = nullptr, *in_end = nullptr;
    uint64_t bb = 0;          // bit buffer, LSB first
    int bc = 0;               // valid bits
    int phase = 0;            // 0: a block header comes next, 1: inside a stored block, 2: inside a Huffman block
    bool last = false;        // the current block is the final one of its stream
    bool overrun = false;     // input ended inside the stream
    uint32_t stored_left = 0;
    const uint32_t *litp = nullptr, *distp = nullptr;
    std::atomic<int64_t> *progress = nullptr;   // output position (elements behind prog_base) published at block ends for
    const void *prog_base = nullptr;            // the CRC follower
    uint32_t lit[LIT_TAB], dist[DIST_TAB];

    // byte-wise refill for the ends of the input: behind the last byte zeros are fed and COUNTED (in moves on), so that
    // the caller can tell how many of them were really used; more than a buffer's worth means the stream is truncated
    inline void refill_safe()
    {
        while (bc < 56) {                 // (never to 64: the word-wise refill of the fast loop shifts by bc)
            if (in < in_end) bb |= (uint64_t)*in << bc;
            else if (in >= in_end + 8) overrun = true;
            in++;
            bc += 8;
        }
    }
    inline uint32_t take(int n) { const uint32_t v = (uint32_t)(bb & ((1ull << n) - 1)); bb >>= n; bc -= n; return v; }
    inline int64_t bitpos() const { return (int64_t)(in - in_beg) * 8 - bc; }
    void start_at_bit(const uint8_t *beg, const uint8_t *end, int64_t bit)
    {
        in_beg = beg; in_end = end; in = beg + (bit >> 3);
        bb = 0; bc = 0; phase = 0; last = false; overrun = false; stored_left = 0;
        if (bit & 7) { refill_safe(); take((int)(bit & 7)); }
    }
    // the first whole byte that has not been used (call at a byte-aligned point of the stream)
    const uint8_t *byte_ptr() const { return in - bc / 8; }
};

template <class T> inline void put2(T *out, uint32_t v);
template <> inline void put2<uint8_t>(uint8_t *out, uint32_t v) { const uint16_t w = (uint16_t)v; memcpy(out, &w, 2); }
template <> inline void put2<uint16_t>(uint16_t *out, uint32_t v) { const uint32_t w = (v & 0xffu) | (v & 0xff00u) << 8; memcpy(out, &w, 4); }

// Block header at the decoder's position -> phase 1 / 2.  0 or ITSX_EFORMAT
int read_block_header(Dec &s)
{
    s.refill_safe();
    s.last = s.take(1) != 0;
    const uint32_t type = s.take(2);
    if (type == 0) {
        s.take(s.bc & 7);                            // to the byte boundary
        const uint8_t *p = s.byte_ptr();             // bytes still in the bit buffer belong to the input
        s.bb = 0; s.bc = 0;
        if (p + 4 > s.in_end) return ITSX_EFORMAT;
        const uint32_t len = p[0] | p[1] << 8, nlen = p[2] | p[3] << 8;
        if ((len ^ nlen) != 0xffffu) return ITSX_EFORMAT;
        s.in = p + 4;
        if (s.in + len > s.in_end) return ITSX_EFORMAT;
        s.stored_left = len;
        s.phase = 1;
        return 0;
    }
    if (type == 1) {
        const FixedTables &f = fixed_tables();
        s.litp = f.lit; s.distp = f.dist;
        s.phase = 2;
        return 0;
    }
    if (type != 2) return ITSX_EFORMAT;
    uint8_t lens[320];
    const int hlit = (int)s.take(5) + 257, hdist = (int)s.take(5) + 1, hclen = (int)s.take(4) + 4;
    if (hlit > 286 || hdist > 30) return ITSX_EFORMAT;
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; i++) { s.refill_safe(); cl[CL_PERM[i]] = (uint8_t)s.take(3); }
    uint32_t cltab[1 << 7];
    if (!build_table(cl, 19, 7, 2, cltab, 1 << 7)) return ITSX_EFORMAT;
    int i = 0;
    while (i < hlit + hdist) {
        s.refill_safe();
        if (s.overrun) return ITSX_EFORMAT;
        const uint32_t e = cltab[s.bb & 127u];
        if (E_TYPE(e) == T_BAD) return ITSX_EFORMAT;
        s.take((int)E_BITS(e));
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
    if (!build_table(lens, hlit, LIT_TB, 0, s.lit, LIT_TAB)) return ITSX_EFORMAT;
    if (!build_table(lens + 288, hdist, DIST_TB, 1, s.dist, DIST_TAB)) return ITSX_EFORMAT;
    s.litp = s.lit; s.distp = s.dist;
    s.phase = 2;
    return 0;
}

// Decode from the decoder's state into out (elements of T: bytes, or 16-bit symbols with markers).
//   hist_beg: the oldest element a back-reference may reach;  stop_bit: return R_STOP at the first block boundary at or
//   behind this bit position.  R_END: the final block of the stream has ended (decoder at the bit behind it);
//   R_FULL: the next symbol does not fit in front of out_end (nothing of it has been consumed).
template <class T>
static inline __attribute__((always_inline)) int run_impl(Dec &s, T *&outp, T *hist_beg, T *out_end, int64_t stop_bit)
{
    constexpr uint32_t W = 8 / sizeof(T);            // elements per 8-byte word
    T *out = outp;
    for (;;) {
        if (s.phase == 0) {
            if (s.progress) s.progress->store((int64_t)(out - (const T *)s.prog_base), std::memory_order_release);
            if (s.bitpos() >= stop_bit) { outp = out; return R_STOP; }
            const int rc = read_block_header(s);
            if (rc < 0) { outp = out; return rc; }
        }
        if (s.phase == 1) {
            const size_t room = (size_t)(out_end - out);
            const uint32_t n = (uint32_t)std::min<size_t>(s.stored_left, room);
            if (sizeof(T) == 1) memcpy(out, s.in, n);
            else for (uint32_t k = 0; k < n; k++) out[k] = (T)s.in[k];
            out += n; s.in += n; s.stored_left -= n;
            if (s.stored_left) { outp = out; return R_FULL; }
        } else {
            // ---- symbols ----
            for (;;) {
                // fast loop: at least 16 readable input bytes and room for the longest match plus the copy's overshoot.
                // The state lives in locals (the stores into the output could alias the struct otherwise).
                {
                    const uint8_t *in = s.in, *const in_fast = s.in_end - 16;
                    T *const out_fast = out_end - (258 + 2 * W + 8);
                    uint64_t bb = s.bb;
                    int bc = s.bc;
                    const uint32_t *const lit = s.litp, *const dist = s.distp;
                    int stop = 0;                        // 1: end of block, 2: bad stream
#define REFILL() do { uint64_t w_; memcpy(&w_, in, 8); bb |= w_ << bc; in += (63 - bc) >> 3; bc |= 56; } while (0)
#define EAT(e_) do { bb >>= E_BITS(e_); bc -= (int)E_BITS(e_); } while (0)
#define VAL(e_, saved_) (((e_) >> 16) + (uint32_t)(((saved_) >> (E_BITS(e_) - E_XTRA(e_))) & ((1u << E_XTRA(e_)) - 1)))
                    if (in <= in_fast && out <= out_fast) {
                        REFILL();
                        uint32_t e = lit[bb & LMASK];        // always the entry of the bits in front: looked up ahead of its use
                        for (;;) {
                            if (E_TYPE(e) == T_LIT) {
                                // up to three look-ups (six literals) on one refill: 3 x 11 bits at most
                                put2<T>(out, e >> 16); out += E_XTRA(e); EAT(e);
                                e = lit[bb & LMASK];
                                if (E_TYPE(e) == T_LIT) {
                                    put2<T>(out, e >> 16); out += E_XTRA(e); EAT(e);
                                    e = lit[bb & LMASK];
                                    if (E_TYPE(e) == T_LIT) {
                                        put2<T>(out, e >> 16); out += E_XTRA(e); EAT(e);
                                        e = lit[bb & LMASK];
                                    }
                                }
                            } else {
                                if (__builtin_expect(E_TYPE(e) == T_SUB, 0)) {
                                    bb >>= LIT_TB; bc -= LIT_TB;
                                    e = lit[(e >> 16) + (bb & ((1u << E_XTRA(e)) - 1))];
                                    if (E_TYPE(e) == T_LIT) { *out++ = (T)(e >> 16); EAT(e); e = lit[bb & LMASK]; goto next; }
                                }
                                if (__builtin_expect(E_TYPE(e) != T_LEN, 0)) { EAT(e); stop = E_TYPE(e) == T_EOB ? 1 : 2; break; }
                                const uint32_t len = VAL(e, bb);
                                EAT(e);
                                uint32_t d = dist[bb & DMASK];
                                if (__builtin_expect(E_TYPE(d) == T_SUB, 0)) {
                                    bb >>= DIST_TB; bc -= DIST_TB;
                                    d = dist[(d >> 16) + (bb & ((1u << E_XTRA(d)) - 1))];
                                }
                                if (__builtin_expect(E_TYPE(d) != T_LEN, 0)) { stop = 2; break; }
                                const uint32_t distv = VAL(d, bb);
                                EAT(d);
                                e = lit[bb & LMASK];             // the next symbol's entry loads while the match is copied
                                if (__builtin_expect((size_t)distv > (size_t)(out - hist_beg), 0)) { stop = 2; break; }
                                T *o = out;
                                const T *m = o - distv;
                                out += len;
                                if (distv >= W) {
                                    uint64_t v;
                                    memcpy(&v, m, 8); memcpy(o, &v, 8);
                                    memcpy(&v, m + W, 8); memcpy(o + W, &v, 8);
                                    for (uint32_t k = 2 * W; k < len; k += W) { memcpy(&v, m + k, 8); memcpy(o + k, &v, 8); }
                                } else if (distv == 1) {
                                    const T c = *m;
                                    for (uint32_t k = 0; k < len; k++) o[k] = c;
                                } else {
                                    for (uint32_t k = 0; k < len; k++) o[k] = m[k];
                                }
                            }
                        next:
                            if (!(in <= in_fast && out <= out_fast)) break;      // (e has not been consumed)
                            REFILL();
                        }
                    }
#undef REFILL
#undef EAT
                    s.in = in; s.bb = bb; s.bc = bc;
                    if (stop == 1) break;
                    if (stop == 2) { outp = out; return ITSX_EFORMAT; }
                }
                // careful path near the ends of input / output: one symbol, looked at before it is consumed
                s.refill_safe();
                if (s.overrun) { outp = out; return ITSX_EFORMAT; }
                uint64_t bb = s.bb;
                int used = 0;
                uint32_t e = s.litp[bb & LMASK];
                if (E_TYPE(e) == T_SUB) { bb >>= LIT_TB; used += LIT_TB; e = s.litp[(e >> 16) + (bb & ((1u << E_XTRA(e)) - 1))]; }
                const uint64_t at_e = bb;                // (length extra bits are cut out of this)
                bb >>= E_BITS(e); used += (int)E_BITS(e);
                const uint32_t ty = E_TYPE(e);
                if (ty == T_LIT) {
                    const uint32_t cnt = E_XTRA(e);
                    if ((size_t)(out_end - out) < cnt) { outp = out; return R_FULL; }
                    *out++ = (T)((e >> 16) & 0xffu);
                    if (cnt == 2) *out++ = (T)(e >> 24);
                    s.bb = bb; s.bc -= used;
                    continue;
                }
                if (ty == T_EOB) { s.bb = bb; s.bc -= used; break; }
                if (ty != T_LEN) { outp = out; return ITSX_EFORMAT; }
                const uint32_t len = VAL(e, at_e);
                uint32_t d = s.distp[bb & DMASK];
                if (E_TYPE(d) == T_SUB) { bb >>= DIST_TB; used += DIST_TB; d = s.distp[(d >> 16) + (bb & ((1u << E_XTRA(d)) - 1))]; }
                if (E_TYPE(d) != T_LEN) { outp = out; return ITSX_EFORMAT; }
                const uint32_t distv = VAL(d, bb);
                bb >>= E_BITS(d); used += (int)E_BITS(d);
                if ((size_t)distv > (size_t)(out - hist_beg)) { outp = out; return ITSX_EFORMAT; }
                if ((size_t)(out_end - out) < len) { outp = out; return R_FULL; }
                for (uint32_t k = 0; k < len; k++) { *out = *(out - distv); out++; }
                s.bb = bb; s.bc -= used;
            }
        }
#undef VAL
        // end of a block
        s.phase = 0;
        if (s.overrun) { outp = out; return ITSX_EFORMAT; }
        if (s.last) {
            if (s.progress) s.progress->store((int64_t)(out - (const T *)s.prog_base), std::memory_order_release);
            outp = out;
            return R_END;
        }
    }
}

int run8_generic(Dec &s, uint8_t *&o, uint8_t *h, uint8_t *e, int64_t stop) { return run_impl<uint8_t>(s, o, h, e, stop); }
int run16_generic(Dec &s, uint16_t *&o, uint16_t *h, uint16_t *e, int64_t stop) { return run_impl<uint16_t>(s, o, h, e, stop); }
#if defined(__x86_64__)
__attribute__((target("bmi2"))) int run8_bmi2(Dec &s, uint8_t *&o, uint8_t *h, uint8_t *e, int64_t stop) { return run_impl<uint8_t>(s, o, h, e, stop); }
__attribute__((target("bmi2"))) int run16_bmi2(Dec &s, uint16_t *&o, uint16_t *h, uint16_t *e, int64_t stop) { return run_impl<uint16_t>(s, o, h, e, stop); }
bool have_bmi2()
{
    static const bool b = __builtin_cpu_supports("bmi2");
    return b;
}
#else
#define run8_bmi2 run8_generic
#define run16_bmi2 run16_generic
bool have_bmi2() { return false; }
#endif
inline int run(Dec &s, uint8_t *&o, uint8_t *h, uint8_t *e, int64_t stop) { return have_bmi2() ? run8_bmi2(s, o, h, e, stop) : run8_generic(s, o, h, e, stop); }
inline int run(Dec &s, uint16_t *&o, uint16_t *h, uint16_t *e, int64_t stop) { return have_bmi2() ? run16_bmi2(s, o, h, e, stop) : run16_generic(s, o, h, e, stop); }

// ---- CRC-32, slice-by-8 ----
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
#if defined(__x86_64__)
// The same register advance by carry-less multiplication (Gopal et al., "Fast CRC computation for generic polynomials
// using PCLMULQDQ"): four 128-bit lanes folded by x^512, then by x^128, then 128 -> 64 -> 32 bits with a Barrett
// reduction; constants for the reflected polynomial 0xEDB88320.  n >= 64 and a multiple of 16.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc_update_clmul(uint32_t c, const uint8_t *p, size_t n)
{
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
    const __m128i k5k0 = _mm_set_epi64x(0x0000000000ll, 0x0163cd6124ll);
    const __m128i poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i *)(p + 0x00));
    x2 = _mm_loadu_si128((const __m128i *)(p + 0x10));
    x3 = _mm_loadu_si128((const __m128i *)(p + 0x20));
    x4 = _mm_loadu_si128((const __m128i *)(p + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)c));
    x0 = k1k2;
    p += 64; n -= 64;
    while (n >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i *)(p + 0x00)); y6 = _mm_loadu_si128((const __m128i *)(p + 0x10));
        y7 = _mm_loadu_si128((const __m128i *)(p + 0x20)); y8 = _mm_loadu_si128((const __m128i *)(p + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        p += 64; n -= 64;
    }
    x0 = k3k4;
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (n >= 16) {
        x2 = _mm_loadu_si128((const __m128i *)p);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        p += 16; n -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = k5k0;
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = poly;
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
bool have_clmul()
{
    // known answer first: the fast path is used only if it reproduces the table-driven register on this CPU
    static const bool ok = [] {
        if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1") || getenv("ITSX_NO_CLMUL")) return false;
        crc_init();
        uint8_t buf[256 + 48];
        for (int i = 0; i < (int)sizeof buf; i++) buf[i] = (uint8_t)(i * 131 + 7);
        for (size_t n : {(size_t)64, (size_t)80, (size_t)128, (size_t)304})
            if (crc_update_clmul(0x12345678u, buf, n) != crc_update(0x12345678u, buf, n)) return false;
        return true;
    }();
    return ok;
}
#else
bool have_clmul() { return false; }
uint32_t crc_update_clmul(uint32_t c, const uint8_t *, size_t) { return c; }
#endif
uint32_t crc_update_fast(uint32_t c, const uint8_t *p, size_t n)
{
    if (n >= 256 && have_clmul()) {
        const size_t body = n & ~(size_t)15;
        c = crc_update_clmul(c, p, body);
        p += body; n -= body;
    }
    return crc_update(c, p, n);
}
// zlib's crc32(crc, buf, len): the finalised CRC of A || buf from the finalised CRC of A
uint32_t crc32_cont(uint32_t crc, const uint8_t *p, size_t n)
{
    crc_init();
    return crc_update_fast(crc ^ 0xffffffffu, p, n) ^ 0xffffffffu;
}
// finalised CRC of A || B from those of A and B and B's length (any length): x^(8 len) by square and multiply
uint32_t crc32_join(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    if (len_b == 0) return crc_a;
    uint32_t r = 1u << 31, sq = 1u << 23;                 // x^0, x^8
    for (uint64_t m = len_b; m; m >>= 1) {
        if (m & 1u) r = dfl_multmodp(sq, r);
        sq = dfl_multmodp(sq, sq);
    }
    return dfl_multmodp(r, crc_a) ^ crc_b;
}

// CRC-32 of a member WHILE it is being inflated: a second thread follows the decoder through the output (the decoder
// publishes its position at every deflate block end), so the check costs no time of the inflating core.
struct CrcFollower {
    const uint8_t *base = nullptr;
    std::atomic<int64_t> produced{0};
    std::atomic<bool> done{false};
    uint32_t reg = 0xffffffffu;
    std::thread th;
    void start(const uint8_t *b, uint32_t crc_so_far)
    {
        base = b;
        reg = crc_so_far ^ 0xffffffffu;
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
    uint32_t finish(int64_t total)          // total: bytes at base that belong to the checksum
    {
        produced.store(total, std::memory_order_release);
        done.store(true, std::memory_order_release);
        th.join();
        return reg ^ 0xffffffffu;
    }
};

// ---- gzip member header at p: pointer behind it, or nullptr ----
const uint8_t *skip_member_header(const uint8_t *p, const uint8_t *end)
{
    if (end - p < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xe0)) return nullptr;
    const uint8_t flg = p[3];
    const uint8_t *q = p + 10;
    if (flg & 4) { if (end - q < 2) return nullptr; const int xl = q[0] | q[1] << 8; q += 2; if (end - q < xl) return nullptr; q += xl; }
    if (flg & 8) { while (q < end && *q) q++; if (q >= end) return nullptr; q++; }
    if (flg & 16) { while (q < end && *q) q++; if (q >= end) return nullptr; q++; }
    if (flg & 2) { if (end - q < 2) return nullptr; q += 2; }
    return q;
}

// ---- search for the start of a (non-final, dynamic) deflate block at bit granularity ----
inline uint64_t peek_at(const uint8_t *beg, const uint8_t *end, int64_t bit)     // >= 57 bits of the stream from `bit` on
{
    const uint8_t *p = beg + (bit >> 3);
    uint64_t w = 0;
    if (p + 8 <= end) memcpy(&w, p, 8);
    else if (p < end) memcpy(&w, p, (size_t)(end - p));
    return w >> (bit & 7);
}
// a dynamic block header (BTYPE = 2, either BFINAL) at bit p that a decoder would accept: HLIT / HDIST in range, a
// COMPLETE code-length code, valid run-lengths, end-of-block present, complete literal and distance codes
bool dyn_header_ok(const uint8_t *beg, const uint8_t *end, int64_t p)
{
    const uint64_t w = peek_at(beg, end, p);
    if ((w & 6u) != 4u) return false;
    const int hlit = (int)((w >> 3) & 31u) + 257, hdist = (int)((w >> 8) & 31u) + 1, hclen = (int)((w >> 13) & 15u) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t cl[19] = {0};
    const uint64_t w2 = peek_at(beg, end, p + 17);    // 19 x 3 = 57 bits
    int kraft = 0;
    for (int i = 0; i < hclen; i++) { const int l = (int)((w2 >> (3 * i)) & 7u); cl[CL_PERM[i]] = (uint8_t)l; if (l) kraft += 128 >> l; }
    if (kraft != 128) return false;
    uint32_t cltab[1 << 7];
    if (!build_table(cl, 19, 7, 2, cltab, 1 << 7)) return false;
    uint8_t lens[320];
    int64_t q = p + 17 + 3 * hclen;
    int i = 0;
    bool ok = true;
    while (i < hlit + hdist) {
        uint64_t v = peek_at(beg, end, q);
        const uint32_t e = cltab[v & 127u];
        if (E_TYPE(e) == T_BAD) { ok = false; break; }
        q += E_BITS(e); v >>= E_BITS(e);
        const int sym = (int)(e >> 16);
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) { if (i == 0) { ok = false; break; } val = lens[i - 1]; rep = 3 + (int)(v & 3u); q += 2; }
        else if (sym == 17) { rep = 3 + (int)(v & 7u); q += 3; }
        else { rep = 11 + (int)(v & 127u); q += 7; }
        if (i + rep > hlit + hdist) { ok = false; break; }
        while (rep--) lens[i++] = (uint8_t)val;
    }
    if (!ok || lens[256] == 0 || q >= (int64_t)(end - beg) * 8) return false;
    // both codes complete (a single distance code or none at all is legal too)
    int cnt[16] = {0}, left = 1, used = 0;
    for (int k = 0; k < hlit; k++) cnt[lens[k]]++;
    cnt[0] = 0;
    for (int l = 1; l < 16 && left >= 0; l++) { left = (left << 1) - cnt[l]; used += cnt[l]; }
    if (left != 0 || used < 2) return false;
    memset(cnt, 0, sizeof cnt); left = 1; used = 0;
    for (int k = 0; k < hdist; k++) cnt[lens[hlit + k]]++;
    cnt[0] = 0;
    for (int l = 1; l < 16 && left >= 0; l++) { left = (left << 1) - cnt[l]; used += cnt[l]; }
    if (left < 0 || (left > 0 && used > 1)) return false;
    return true;
}

// The first place in [from_bit, to_bit) where decoding can start: a non-final dynamic block header (any bit), or a gzip
// member header (a byte boundary: magic, CM = 8, reserved flags clear, and a first block that is acceptable) -- BGZF and
// other many-member files have one FINAL block per member.  -1 if none.
int64_t find_start(const uint8_t *beg, const uint8_t *end, int64_t from_bit, int64_t to_bit, bool *is_header)
{
    const int64_t last_bit = (int64_t)(end - beg) * 8 - 64 * 8;        // this close to the end everything is left to the predecessor
    to_bit = std::min(to_bit, last_bit);
    *is_header = false;
    // byte by byte: one load serves the eight bit offsets (3 header bits + HLIT + HDIST = 13 bits, 20 with the offset)
    for (int64_t byte = from_bit >> 3; byte * 8 < to_bit; byte++) {
        const uint8_t *q = beg + byte;
        if (byte * 8 >= from_bit && q[0] == 0x1f && q[1] == 0x8b && q[2] == 8 && !(q[3] & 0xe0)) {
            const uint8_t *d = skip_member_header(q, end);
            if (d && end - d > 16) {
                const int type = (d[0] >> 1) & 3;
                const bool ok = type == 1 || (type == 0 && ((d[1] | d[2] << 8) ^ (d[3] | d[4] << 8)) == 0xffff) ||
                                (type == 2 && dyn_header_ok(beg, end, (int64_t)(d - beg) * 8));
                if (ok) { *is_header = true; return byte * 8; }
            }
        }
        uint32_t w4;
        memcpy(&w4, q, 4);
        for (int o = 0; o < 8; o++) {
            const uint32_t w = w4 >> o;
            if ((w & 7u) != 4u) continue;                 // BFINAL = 0, BTYPE = 2
            if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) continue;       // HLIT, HDIST
            const int64_t p = byte * 8 + o;
            if (p < from_bit || p >= to_bit) continue;
            if (dyn_header_ok(beg, end, p)) return p;
        }
    }
    return -1;
}

// ---- a piece of one gzip file decoded into 16-bit symbols ----
struct MemberEnd { int64_t out_pos; uint32_t crc, isize; };      // a member ended in front of symbol out_pos of the segment
struct Segment {
    std::vector<uint16_t> sym;        // [WSIZE markers][symbols]
    int64_t nout = 0;
    int64_t start_bit = -1, end_bit = -1;
    bool start_header = false;        // starts at a gzip member header (byte start_bit / 8) instead of a block header
    bool end_header = false;          // ended behind a member trailer, i.e. where the next member header (or the end) is
    bool eof = false;                 // ended at the end of the input (behind the last member and its padding)
    int err = 0;
    std::vector<MemberEnd> ends;
    std::vector<uint32_t> piece_crc;  // CRC-32 of the resolved bytes between member ends (ends.size() + 1 pieces)
    uint8_t window[WSIZE];            // the 32 KB in front of the segment, resolved
};

// decode [start .. first block boundary / member header at or behind stop_bit) of the file into seg
void decode_segment(Segment &g, Dec &d, const uint8_t *beg, const uint8_t *end, int64_t start_bit, bool at_header, int64_t stop_bit,
                    size_t guess)
{
    g.nout = 0; g.ends.clear(); g.err = 0; g.eof = false; g.end_header = false;
    g.start_bit = start_bit; g.start_header = at_header; g.end_bit = -1;
    if (g.sym.size() < WSIZE + guess + 1024) g.sym.resize(WSIZE + guess + 1024);
    for (int j = 0; j < WSIZE; j++) g.sym[j] = (uint16_t)(0x8000 + j);
    size_t out_at = WSIZE, hist_at = 0;
    if (at_header) {
        const uint8_t *q = skip_member_header(beg + (start_bit >> 3), end);
        if (!q) { g.err = ITSX_EFORMAT; return; }
        d.start_at_bit(beg, end, (int64_t)(q - beg) * 8);
        hist_at = out_at;
    } else {
        d.start_at_bit(beg, end, start_bit);
    }
    d.progress = nullptr;
    for (;;) {
        uint16_t *base = g.sym.data(), *out = base + out_at;
        const int rc = run(d, out, base + hist_at, base + g.sym.size(), stop_bit);
        out_at = (size_t)(out - base);
        if (rc == R_FULL) { g.sym.resize(g.sym.size() + g.sym.size() / 2 + (1 << 16)); continue; }
        if (rc < 0) { g.err = rc; return; }
        if (rc == R_STOP) { g.end_bit = d.bitpos(); break; }
        // R_END: trailer, padding, next member
        d.take(d.bc & 7);
        const uint8_t *p = d.byte_ptr();
        if (p > end || end - p < 8) { g.err = ITSX_EFORMAT; return; }
        MemberEnd me;
        me.out_pos = (int64_t)(out_at - WSIZE);
        me.crc = p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24;
        me.isize = p[4] | p[5] << 8 | p[6] << 16 | (uint32_t)p[7] << 24;
        g.ends.push_back(me);
        p += 8;
        if (p < end && *p == 0) { const uint8_t *z = p; while (z < end && *z == 0) z++; if (z == end) p = end; }
        g.end_bit = (int64_t)(p - beg) * 8;
        g.end_header = true;
        if (p == end) { g.eof = true; break; }
        if (g.end_bit >= stop_bit) break;
        const uint8_t *q = skip_member_header(p, end);
        if (!q) { g.err = ITSX_EFORMAT; return; }
        d.start_at_bit(beg, end, (int64_t)(q - beg) * 8);
        hist_at = out_at;
        g.end_header = false;
    }
    g.nout = (int64_t)(out_at - WSIZE);
}

inline int hw_threads()
{
    int hw = (int)std::thread::hardware_concurrency();
    return hw <= 0 ? 4 : std::min(hw, 64);
}

}   // namespace

// ---- the reader ----
struct itsx_gz {
    const uint8_t *src = nullptr, *end = nullptr;
    int threads = 1;
    // position: a member header at byte (pos_bit / 8) comes next (at_header), or a block header at bit pos_bit
    int64_t pos_bit = 0;
    bool at_header = true, eof = false;
    int err = 0;
    // the member being decoded: finalised CRC and length of what has been produced of it
    uint32_t mcrc = 0;
    uint64_t mlen = 0;
    // output produced but not handed out yet
    std::vector<uint8_t> pending;
    size_t pend_at = 0;
    // last 32 KB of the output (right-aligned; handed out or pending)
    uint8_t window[WSIZE];
    // during a call: how many of the most recent output bytes lie right in front of the next byte to write (the
    // caller's promise `hist` plus what the call has written): decoding goes on right there when that covers the window
    uint64_t contig = 0;
    // sequential decoder
    std::unique_ptr<Dec> dec;
    bool dec_live = false;                          // dec holds the state at pos (inside a member, between calls)
    std::vector<uint8_t> ibuf;                      // [WSIZE window][space] for decoding away from the caller's buffer
    // parallel decoder
    std::vector<std::unique_ptr<Segment>> segs, bridges;
    std::vector<std::unique_ptr<Dec>> decs;
    int64_t n_batches = 0, n_bridges = 0;
    int64_t chunk_min = 1 << 20, chunk_max = 2 << 20;   // compressed bytes per chunk of a batch
    int64_t par_min = 3 << 20;                          // less compressed input than this left: one thread
};

namespace {

constexpr size_t IBUF_SPACE = 1 << 17;
constexpr int64_t DIRECT_MIN_ROOM = 1 << 16;

void push_window(itsx_gz *h, const uint8_t *p, size_t n)
{
    if (n >= WSIZE) memcpy(h->window, p + n - WSIZE, WSIZE);
    else if (n) { memmove(h->window, h->window + n, WSIZE - n); memcpy(h->window + WSIZE - n, p, n); }
}

// trailer behind a finished member (decoder at the bit behind the final block): checks it, moves to the next header
int finish_member(itsx_gz *h, Dec &d)
{
    d.take(d.bc & 7);
    const uint8_t *p = d.byte_ptr();
    if (p > h->end || h->end - p < 8) return ITSX_EFORMAT;
    const uint32_t crc = p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24;
    const uint32_t isize = p[4] | p[5] << 8 | p[6] << 16 | (uint32_t)p[7] << 24;
    if (crc != h->mcrc || isize != (uint32_t)h->mlen) return ITSX_EFORMAT;
    p += 8;
    if (p < h->end && *p == 0) { const uint8_t *z = p; while (z < h->end && *z == 0) z++; if (z == h->end) p = h->end; }
    h->pos_bit = (int64_t)(p - h->src) * 8;
    h->at_header = true;
    h->dec_live = false;
    h->mcrc = 0; h->mlen = 0;
    if (p == h->end) h->eof = true;
    return 0;
}

// One stretch of sequential decoding into dst[0 .. room).  Returns the bytes written to dst (0 is possible: an empty
// block, or everything went to `pending`), or an error.
int64_t step_sequential(itsx_gz *h, uint8_t *dst, int64_t room)
{
    if (!h->dec) h->dec.reset(new Dec());
    Dec &d = *h->dec;
    if (h->at_header) {
        const uint8_t *q = skip_member_header(h->src + (h->pos_bit >> 3), h->end);
        if (!q) return ITSX_EFORMAT;
        d.start_at_bit(h->src, h->end, (int64_t)(q - h->src) * 8);
        h->at_header = false; h->dec_live = true;
        h->mcrc = 0; h->mlen = 0;
    } else if (!h->dec_live) {
        d.start_at_bit(h->src, h->end, h->pos_bit);          // behind a parallel batch that stopped inside a member
        h->dec_live = true;
    }
    const uint64_t need_hist = std::min<uint64_t>(h->mlen, WSIZE);
    const bool hist_in_place = h->contig >= need_hist;
    int rc;
    int64_t wrote = 0;
    if (hist_in_place && room >= DIRECT_MIN_ROOM) {
        uint8_t *out = dst;
        CrcFollower fol;
        const bool follow = room >= (8 << 20) && h->end - d.in > (2 << 20) && !have_clmul();      // (13 GB/s inline needs no helper)
        if (follow) { d.progress = &fol.produced; d.prog_base = dst; fol.start(dst, h->mcrc); }
        rc = run(d, out, dst - need_hist, dst + room, INT64_MAX);
        d.progress = nullptr;
        const size_t n = (size_t)(out - dst);
        h->mcrc = follow ? fol.finish((int64_t)n) : crc32_cont(h->mcrc, dst, n);
        h->mlen += n;
        push_window(h, dst, n);
        h->contig += n;
        wrote = (int64_t)n;
        if (rc == R_FULL && n == 0) rc = -12345;             // not one symbol fits: go through ibuf
    } else {
        rc = -12345;
    }
    if (rc == -12345) {
        if (h->ibuf.size() < WSIZE + IBUF_SPACE) h->ibuf.resize(WSIZE + IBUF_SPACE);
        memcpy(h->ibuf.data(), h->window, WSIZE);
        uint8_t *const out0 = h->ibuf.data() + WSIZE;
        uint8_t *out = out0;
        rc = run(d, out, out0 - need_hist, out0 + IBUF_SPACE, INT64_MAX);
        const size_t n = (size_t)(out - out0);
        h->mcrc = crc32_cont(h->mcrc, out0, n);
        h->mlen += n;
        push_window(h, out0, n);
        const size_t take = (size_t)std::min<int64_t>((int64_t)n, room);
        memcpy(dst, out0, take);
        if (take < n) { h->pending.assign(out0 + take, out0 + n); h->pend_at = 0; }
        h->contig += take;
        wrote = (int64_t)take;
        if (rc == R_FULL && n == 0) return ITSX_EFORMAT;     // (cannot happen: the space holds any symbol)
    }
    if (rc < 0) return rc;
    if (rc == R_END) { const int e = finish_member(h, d); if (e < 0) return e; }
    return wrote;
}

// f(0 .. n-1) on n threads (the caller's is one of them).  All or nothing: the workers wait behind a gate until every
// thread exists (chunks wait for each other's block starts, so a missing thread would stall the rest); if one cannot be
// started nothing is run and false comes back.  What a worker throws (bad_alloc under memory pressure) is rethrown here
// after every thread has been joined.
template <typename F> bool on_threads(int n, F f)
{
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    std::atomic<int> gate{0};
    auto guarded = [&](int t) {
        try { f(t); }
        catch (...) { std::lock_guard<std::mutex> g(mu); if (!err) err = std::current_exception(); }
    };
    for (int t = 1; t < n; t++) {
        try {
            th.emplace_back([&guarded, &gate, t] {
                int g;
                while ((g = gate.load(std::memory_order_acquire)) == 0) std::this_thread::yield();
                if (g > 0) guarded(t);
            });
        } catch (...) {
            gate.store(-1, std::memory_order_release);
            for (auto &x : th) x.join();
            return false;
        }
    }
    gate.store(1, std::memory_order_release);
    guarded(0);
    for (auto &x : th) x.join();
    if (err) std::rethrow_exception(err);
    return true;
}

// symbols -> bytes through the window in front of the segment; CRC-32 of the pieces between member ends
void translate_segment(Segment &g, uint8_t *dst)
{
    const uint16_t *sym = g.sym.data() + WSIZE;
    const uint8_t *win = g.window;
    const int64_t n = g.nout;
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {                  // markers are rare: 64 symbols without one are narrowed in one sweep
        uint16_t any = 0;
        for (int k = 0; k < 64; k++) any |= sym[i + k];
        if (!(any & 0x8000)) {
            for (int k = 0; k < 64; k++) dst[i + k] = (uint8_t)sym[i + k];
        } else {
            for (int k = 0; k < 64; k++) { const uint16_t v = sym[i + k]; dst[i + k] = v < 0x8000 ? (uint8_t)v : win[v - 0x8000]; }
        }
    }
    for (; i < n; i++) {
        const uint16_t v = sym[i];
        dst[i] = v < 0x8000 ? (uint8_t)v : win[v - 0x8000];
    }
    g.piece_crc.clear();
    int64_t at = 0;
    for (const MemberEnd &me : g.ends) { g.piece_crc.push_back(crc32_cont(0, dst + at, (size_t)(me.out_pos - at))); at = me.out_pos; }
    g.piece_crc.push_back(crc32_cont(0, dst + at, (size_t)(n - at)));
}

// One batch of parallel decoding.  Returns bytes written to dst (the surplus goes to `pending`), < 0 on an error (the
// reader's state is untouched then and the caller goes on sequentially).
int64_t batch_parallel(itsx_gz *h, uint8_t *dst, int64_t room)
{
    const int64_t pos_byte = h->pos_bit >> 3, n_in = (int64_t)(h->end - h->src);
    const int64_t rem = n_in - pos_byte;
    const int64_t C = std::max(h->chunk_min, std::min(h->chunk_max, rem / h->threads));
    int nch = (int)std::min<int64_t>(h->threads, (rem + C - 1) / C);
    if (nch < 2) return -1;
    const bool to_end = pos_byte + (int64_t)nch * C + C / 4 >= n_in;
    const int64_t batch_stop = to_end ? INT64_MAX : (pos_byte + (int64_t)nch * C) * 8;
    while ((int)h->segs.size() < nch) { h->segs.emplace_back(new Segment()); h->decs.emplace_back(new Dec()); }
    std::vector<std::atomic<int64_t>> starts(nch);
    std::vector<char> start_hdr((size_t)nch, 0);         // (written before the release-store of starts[k])
    for (int k = 0; k < nch; k++) starts[k].store(-2);
    start_hdr[0] = h->at_header;
    starts[0].store(h->pos_bit);
    const uint8_t *src = h->src, *end = h->end;
    const int64_t pos_bit = h->pos_bit;
    const bool at_header = h->at_header;
    auto stop_for = [&](int k) -> int64_t {             // the next chunk that found a start (waits for the finders)
        for (int j = k + 1; j < nch; j++) {
            int64_t s;
            while ((s = starts[j].load(std::memory_order_acquire)) == -2) std::this_thread::yield();
            if (s >= 0) return s;
        }
        return batch_stop;
    };
    const bool started = on_threads(nch, [&](int k) {
        Segment &g = *h->segs[k];
        if (k > 0) {
            const int64_t a = (pos_byte + (int64_t)k * C) * 8;
            const int64_t b = (k + 1 < nch || !to_end) ? a + C * 8 : n_in * 8;
            bool hdr = false;
            const int64_t st = find_start(src, end, a, b, &hdr);
            start_hdr[(size_t)k] = hdr;
            starts[k].store(st, std::memory_order_release);
            if (st < 0) { g.start_bit = -1; return; }
        }
        const int64_t stop = stop_for(k);
        decode_segment(g, *h->decs[k], src, end, k ? starts[k].load() : pos_bit, k ? (bool)start_hdr[(size_t)k] : at_header, stop, (size_t)(C * 7 / 2));
    });
    if (!started) return -1;                                 // (the single-core decoder goes on from here)
    // ---- stitch ----
    std::vector<Segment *> acc;
    Segment *cur = h->segs[0].get();
    if (cur->err) return cur->err;
    acc.push_back(cur);
    size_t nbridge = 0;
    auto bridge_to = [&](int64_t stop) -> int {
        if (nbridge == h->bridges.size()) h->bridges.emplace_back(new Segment());
        Segment *b = h->bridges[nbridge++].get();
        decode_segment(*b, *h->decs[0], src, end, cur->end_bit, cur->end_header, stop, (size_t)(C * 7 / 2));
        h->n_bridges++;
        if (b->err) return b->err;
        acc.push_back(b);
        cur = b;
        return 0;
    };
    int j = 1;
    while (j < nch && !cur->eof) {
        const int64_t s = starts[j].load();
        if (s < 0) { j++; continue; }
        if (cur->end_bit == s && cur->end_header == (bool)start_hdr[(size_t)j]) {
            Segment *g = h->segs[j].get();
            if (g->err) return g->err;
            acc.push_back(g);
            cur = g;
            j++;
        } else if (cur->end_bit >= s) {
            j++;                                              // not a block start after all
        } else {
            const int e = bridge_to(s);                       // the predecessor was stopped early by a start that was none
            if (e < 0) return e;
        }
    }
    if (!cur->eof && cur->end_bit < batch_stop) { const int e = bridge_to(batch_stop); if (e < 0) return e; }
    // ---- windows, front to back ----
    int64_t total = 0;
    uint8_t win[WSIZE];
    memcpy(win, h->window, WSIZE);
    for (Segment *g : acc) {
        memcpy(g->window, win, WSIZE);
        total += g->nout;
        const uint16_t *sym = g->sym.data() + WSIZE;
        const int64_t n = g->nout, keep = std::min<int64_t>(n, WSIZE);
        if (keep < WSIZE) memmove(win, win + keep, (size_t)(WSIZE - keep));
        for (int64_t i = n - keep, o = WSIZE - keep; i < n; i++, o++) {
            const uint16_t v = sym[i];
            win[o] = v < 0x8000 ? (uint8_t)v : g->window[v - 0x8000];
        }
    }
    // ---- bytes ----
    uint8_t *base = dst;
    const bool direct = total <= room;
    if (!direct) { h->pending.resize((size_t)total); h->pend_at = 0; base = h->pending.data(); }
    {
        std::vector<int64_t> off(acc.size());
        int64_t o = 0;
        for (size_t i = 0; i < acc.size(); i++) { off[i] = o; o += acc[i]->nout; }
        std::atomic<size_t> next{0};
        auto work = [&](int) { for (size_t i; (i = next.fetch_add(1)) < acc.size();) translate_segment(*acc[i], base + off[i]); };
        if (!on_threads(std::min<int>(nch, (int)acc.size()), work)) work(0);
    }
    // ---- members: CRC-32 and ISIZE ----
    uint32_t mcrc = h->mcrc;
    uint64_t mlen = h->mlen;
    for (Segment *g : acc) {
        int64_t at = 0;
        for (size_t i = 0; i < g->ends.size(); i++) {
            const MemberEnd &me = g->ends[i];
            mcrc = crc32_join(mcrc, g->piece_crc[i], (uint64_t)(me.out_pos - at));
            mlen += (uint64_t)(me.out_pos - at);
            at = me.out_pos;
            if (mcrc != me.crc || (uint32_t)mlen != me.isize) {
                if (!direct) { h->pending.clear(); h->pend_at = 0; }
                return ITSX_EFORMAT;
            }
            mcrc = 0; mlen = 0;
        }
        mcrc = crc32_join(mcrc, g->piece_crc[g->ends.size()], (uint64_t)(g->nout - at));
        mlen += (uint64_t)(g->nout - at);
    }
    // ---- commit ----
    h->mcrc = mcrc; h->mlen = mlen;
    memcpy(h->window, win, WSIZE);
    h->pos_bit = cur->end_bit; h->at_header = cur->end_header; h->eof = cur->eof;
    h->dec_live = false;
    h->n_batches++;
    if (direct) { h->contig += (uint64_t)total; return total; }
    const int64_t take = std::min(total, room);
    memcpy(dst, base, (size_t)take);
    h->pend_at = (size_t)take;
    h->contig += (uint64_t)take;
    return take;
}

}   // namespace

extern "C" {

/* see include/itsx_b200.h */
itsx_gz *itsx_gz_open(const uint8_t *src, int64_t n, int threads)
{
    if ((!src && n) || n < 0) return nullptr;
    itsx_gz *h = new (std::nothrow) itsx_gz();
    if (!h) return nullptr;
    h->src = src; h->end = src + n;
    h->threads = threads > 0 ? std::min(threads, 64) : hw_threads();
    memset(h->window, 0, WSIZE);
    if (n == 0) h->eof = true;
    return h;
}

static int64_t gz_read_impl(itsx_gz *h, uint8_t *dst, int64_t cap, int64_t hist)
{
    int64_t done = 0;
    h->contig = (uint64_t)hist;
    if (h->pend_at < h->pending.size()) {
        const size_t take = std::min<size_t>(h->pending.size() - h->pend_at, (size_t)cap);
        memcpy(dst, h->pending.data() + h->pend_at, take);
        h->pend_at += take;
        done = (int64_t)take;
        h->contig += take;
        if (h->pend_at == h->pending.size()) { h->pending.clear(); h->pend_at = 0; }
    }
    while (done < cap && !h->eof && h->pend_at == h->pending.size()) {
        int64_t r = -1;
        const bool par = h->threads > 2 &&          // (two passes cost two cores: worth it from three on)
                         (h->end - h->src) - (h->pos_bit >> 3) >= h->par_min && !h->dec_live;
        if (par) {
            r = batch_parallel(h, dst + done, cap - done);
            if (r < 0) h->threads = 1;                           // whatever it was, the sequential decoder reports it
        }
        if (r < 0) r = step_sequential(h, dst + done, cap - done);
        if (r < 0) { h->err = (int)r; return r; }
        done += r;
    }
    return done;
}

int64_t itsx_gz_read(itsx_gz *h, uint8_t *dst, int64_t cap, int64_t hist)
{
    if (!h || (!dst && cap > 0) || cap < 0 || hist < 0) return ITSX_EINVAL;
    if (h->err) return h->err;
    try {
        return gz_read_impl(h, dst, cap, hist);
    } catch (...) {                                   // out of memory in a decoder thread: no exception crosses the C ABI
        h->err = ITSX_ELIMIT;
        return ITSX_ELIMIT;
    }
}

int itsx_gz_tune(itsx_gz *h, int64_t chunk_min, int64_t chunk_max, int64_t par_min)
{
    if (!h || chunk_min < 256 || chunk_max < chunk_min || par_min < 0) return ITSX_EINVAL;
    h->chunk_min = chunk_min; h->chunk_max = chunk_max; h->par_min = par_min;
    return ITSX_OK;
}

int64_t itsx_gz_stat(const itsx_gz *h, int what) { return !h ? -1 : what == 0 ? h->n_batches : what == 1 ? h->n_bridges : -1; }

int itsx_gz_eof(const itsx_gz *h) { return h && h->eof && h->pend_at == h->pending.size() ? 1 : 0; }

void itsx_gz_close(itsx_gz *h) { delete h; }

}   // extern "C"
