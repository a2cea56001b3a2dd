// deflate_core.h -- per-thread bodies of the GPU gzip writer (deflate.cu), written so that the same code also
// compiles for the host: tools/deflate_emul.cpp runs the phases thread by thread on the CPU and inflates the result
// with zlib (this container has no GPU; the kernel itself is exercised by tests/test_gpu_gzip.py on the GPU box).
//
// One CTA (DFL_THREADS threads) turns one chunk of DFL_CHUNK input bytes into ONE deflate block with its own dynamic
// Huffman codes (or a stored block when that is smaller).  DFL_GROUP consecutive blocks form a gzip member: a block may
// reach back into the previous block of its member (the deflate window), ends on a byte boundary (an empty stored block
// behind it, zlib's Z_SYNC_FLUSH) so that the blocks are concatenated as bytes, and the last one carries BFINAL.  Phases,
// separated by CTA barriers:
//   1. candidates   the previous block's positions enter the 4-byte hash table; then positions in time slices of
//                   DFL_THREADS: look the hash up (largest earlier position of the slices before), then enter the
//                   slice's own positions (atomicMax => deterministic)
//   2. parse        a thread parses its DFL_SUB bytes (hash chain of up to DFL_MAX_CHAIN candidates vs. distance 1, one
//                   step of lazy evaluation; its last match may run into the next threads' bytes), tokens to HBM;
//      stitch       prefix maximum of the threads' end positions: a thread drops / cuts the tokens the threads before it
//                   already cover, then counts its symbols by shared-memory atomics
//   3. codes        one thread: length-limited Huffman code lengths (two-queue merge, weights halved until the depth
//                   fits), canonical codes, the block header (code lengths with zero runs as symbols 17 / 18)
//   4. sizes        bits of every thread's tokens, exclusive scan
//   5. emit         every thread writes its tokens at its bit offset (atomicOr on 32-bit words)
// plus the CRC-32 of the chunk: every thread runs the byte-wise register over its bytes, the registers are advanced
// to the end of the chunk by multiplication with x^(8 m) modulo the CRC polynomial and XOR-ed together.
#ifndef ITSX_DEFLATE_CORE_H
#define ITSX_DEFLATE_CORE_H
#include <stdint.h>

#ifdef __CUDACC__
#define DFL_HD __host__ __device__ __forceinline__
#else
#define DFL_HD inline
#endif

constexpr int DFL_THREADS = 256;
constexpr int DFL_SUB = 128;                        // bytes parsed by one thread
constexpr int DFL_CHUNK = DFL_THREADS * DFL_SUB;    // 32 768 input bytes per deflate block (= the deflate window)
constexpr int DFL_GROUP = 32;                       // blocks per gzip member (1 MiB of input)
constexpr int DFL_HIST = DFL_CHUNK;                 // bytes of the previous block a block may reach back into
constexpr int DFL_HASH_BITS = 13;
constexpr int DFL_HASH_SIZE = 1 << DFL_HASH_BITS;
constexpr int DFL_NLL = 286, DFL_ND = 30, DFL_NCL = 19;
constexpr int DFL_MIN_HASH_MATCH = 4, DFL_MIN_RUN = 3, DFL_MAX_MATCH = 258;
#ifndef DFL_MAX_CHAIN
#define DFL_MAX_CHAIN 16                            // candidates tried per position (zlib level 3: 32, level 1: 4)
#endif
#ifndef DFL_NICE_MATCH
#define DFL_NICE_MATCH 64                           // a match this long ends the search
#endif
#ifndef DFL_LAZY_MAX
#define DFL_LAZY_MAX 32                             // matches shorter than this are compared with the one a byte later
#endif
constexpr int DFL_TOKS = DFL_SUB + 2;                // token slots of a thread: two spare ones in front (see dfl_stitch)
constexpr int DFL_OUT_WORDS = (DFL_CHUNK + 64) / 4; // output words per chunk: a stored block needs CHUNK + 5 bytes
constexpr int DFL_HDR_WORDS = 96;                   // the dynamic block header is at most ~ 2 600 bits
constexpr uint32_t DFL_TOK_MATCH = 0x80000000u;     // token: literal byte | MATCH | (len - 3) << 16 | (dist - 1)
constexpr uint32_t DFL_CRC_POLY = 0xedb88320u;

// ---- atomics: real ones on the device, plain ones in the one-thread-at-a-time host emulation ----
#ifdef __CUDA_ARCH__
#define DFL_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define DFL_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define DFL_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
#define DFL_ATOMIC_ADD(p, v) (*(p) += (v))
#define DFL_ATOMIC_MAX(p, v) (*(p) = *(p) > (v) ? *(p) : (v))
#define DFL_ATOMIC_OR(p, v) (*(p) |= (v))
#endif

DFL_HD int dfl_top_bit(uint32_t v)          // position of the highest set bit (v > 0)
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    int n = 0;
    while (v >>= 1) n++;
    return n;
#endif
}

DFL_HD uint32_t dfl_hash4(const uint8_t *p)
{
    const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    return (v * 2654435761u) >> (32 - DFL_HASH_BITS);
}

// RFC 1951 3.2.5: length 3..258 -> symbol 257..285, extra bits
DFL_HD int dfl_len_sym(int len, int &ebits, int &eval)
{
    const int l = len - 3;
    if (l < 8) { ebits = 0; eval = 0; return 257 + l; }
    if (l == 255) { ebits = 0; eval = 0; return 285; }
    const int nb = dfl_top_bit((uint32_t)l);
    ebits = nb - 2;
    eval = l & ((1 << ebits) - 1);
    return 257 + 4 * ebits + 4 + ((l >> ebits) & 3);
}
// distance 1..32768 -> symbol 0..29, extra bits
DFL_HD int dfl_dist_sym(int dist, int &ebits, int &eval)
{
    const int d = dist - 1;
    if (d < 4) { ebits = 0; eval = 0; return d; }
    const int nb = dfl_top_bit((uint32_t)d);
    ebits = nb - 1;
    eval = d & ((1 << ebits) - 1);
    return 2 * nb + ((d >> ebits) & 1);
}

DFL_HD int dfl_match_len(const uint8_t *a, const uint8_t *b, int maxlen)
{
    int l = 0;
    while (l < maxlen && a[l] == b[l]) l++;
    return l;
}

// ---- LSB-first bit writer into zero-initialised 32-bit words shared with other writers ----
struct DflBits {
    uint32_t *w;
    uint64_t  acc;
    int       n;
};
DFL_HD void dfl_bits_start(DflBits &b, uint32_t *words, uint32_t bit_offset)
{
    b.w = words + (bit_offset >> 5);
    b.acc = 0;
    b.n = (int)(bit_offset & 31u);
}
DFL_HD void dfl_bits_put(DflBits &b, uint32_t value, int nbits)      // nbits <= 16 per call
{
    b.acc |= (uint64_t)value << b.n;
    b.n += nbits;
    if (b.n >= 32) {
        DFL_ATOMIC_OR(b.w, (uint32_t)b.acc);
        b.w++;
        b.acc >>= 32;
        b.n -= 32;
    }
}
DFL_HD void dfl_bits_finish(DflBits &b)
{
    if (b.n > 0) DFL_ATOMIC_OR(b.w, (uint32_t)b.acc);
}

// ---- Huffman: code lengths of n symbols with frequencies f[], at most `limit` bits (n <= 286) ----
// At least two symbols get a code (zlib does the same), so every tree is complete.  Scratch: the caller's arrays.
struct DflHuffScratch {
    uint16_t order[DFL_NLL];            // used symbols, by (weight, symbol)
    uint32_t weight[2 * DFL_NLL];       // leaves 0..m-1 in order, then the merged nodes
    int16_t  parent[2 * DFL_NLL];
    uint8_t  depth[2 * DFL_NLL];
    uint32_t fw[DFL_NLL];               // working copy of the frequencies
};
DFL_HD void dfl_huff_lengths(const uint32_t *f, int n, int limit, uint8_t *len, DflHuffScratch &s)
{
    int used = 0;
    for (int i = 0; i < n; i++) { s.fw[i] = f[i]; len[i] = 0; used += f[i] != 0; }
    if (used == 0) { s.fw[0] = 1; s.fw[1] = 1; }
    else if (used == 1) { if (s.fw[0] == 0) s.fw[0] = 1; else s.fw[1] = 1; }
    for (;;) {
        int m = 0;
        for (int i = 0; i < n; i++) {
            if (s.fw[i] == 0) continue;
            int at = m++;                                  // insertion sort by (weight, symbol); symbols arrive ascending
            while (at > 0 && s.fw[s.order[at - 1]] > s.fw[i]) { s.order[at] = s.order[at - 1]; at--; }
            s.order[at] = (uint16_t)i;
        }
        for (int i = 0; i < m; i++) s.weight[i] = s.fw[s.order[i]];
        // two queues: leaves [lq, m) and merged nodes [iq, next)
        int lq = 0, iq = m, next = m;
        while (next < 2 * m - 1) {
            int pick[2];
            for (int k = 0; k < 2; k++) {
                if (lq < m && (iq >= next || s.weight[lq] <= s.weight[iq])) pick[k] = lq++;
                else pick[k] = iq++;
            }
            s.weight[next] = s.weight[pick[0]] + s.weight[pick[1]];
            s.parent[pick[0]] = (int16_t)next;
            s.parent[pick[1]] = (int16_t)next;
            next++;
        }
        const int root = 2 * m - 2;
        s.depth[root] = 0;
        int maxd = 0;
        for (int i = root - 1; i >= 0; i--) {
            s.depth[i] = (uint8_t)(s.depth[s.parent[i]] + 1);
            if (i < m && s.depth[i] > maxd) maxd = s.depth[i];
        }
        if (maxd <= limit) {
            for (int i = 0; i < m; i++) len[s.order[i]] = s.depth[i];
            return;
        }
        for (int i = 0; i < n; i++)
            if (s.fw[i]) s.fw[i] = (s.fw[i] + 1) >> 1;     // flatten the distribution and try again
    }
}
// canonical codes, stored bit-reversed so that they can be written LSB first
DFL_HD void dfl_huff_codes(const uint8_t *len, int n, uint16_t *code)
{
    int count[16], nextc[16];
    for (int b = 0; b < 16; b++) count[b] = 0;
    for (int i = 0; i < n; i++) count[len[i]]++;
    count[0] = 0;
    int c = 0;
    nextc[0] = 0;
    for (int b = 1; b < 16; b++) { c = (c + count[b - 1]) << 1; nextc[b] = c; }
    for (int i = 0; i < n; i++) {
        const int l = len[i];
        if (l == 0) { code[i] = 0; continue; }
        uint32_t v = (uint32_t)nextc[l]++, r = 0;
        for (int b = 0; b < l; b++) { r = (r << 1) | (v & 1u); v >>= 1; }
        code[i] = (uint16_t)r;
    }
}

// ---- the block's shared state (shared memory on the device, plain arrays in the emulation) ----
struct DflShared {
    const uint8_t *buf;                 // the chunk (+ 8 readable bytes behind it); `hist` bytes of history in front of it
    int            len;
    int            hist;                // 0 (first block of a member) or DFL_HIST
    int            final;               // last block of its member: BFINAL = 1, no byte-alignment marker behind it
    uint16_t      *cand;                // [DFL_CHUNK]: candidate position (counted from the start of the history) + 1, 0 = none
    uint32_t      *table;               // [DFL_HASH_SIZE]: largest position (from the start of the history) + 1 entered so far
    uint32_t      *freq_ll, *freq_d;    // [288], [32]
    uint8_t       *len_ll, *len_d;      // [288], [32]
    uint16_t      *code_ll, *code_d;    // [288], [32]
    uint32_t      *ntok;                // [DFL_THREADS]: tokens the thread emits ...
    uint32_t      *tbeg;                // [DFL_THREADS]: ... from this slot of its DFL_TOKS
    uint32_t      *tend;                // [DFL_THREADS]: position behind the thread's last token (>= its own end)
    uint32_t      *bits;                // [DFL_THREADS + 1]: token bits per thread, then their exclusive scan
    uint32_t      *hdr;                 // [DFL_HDR_WORDS]: the block header
    uint32_t      *hdr_bits;            // [1]
    uint32_t      *tokens;              // HBM: [DFL_THREADS * DFL_TOKS]
    uint32_t      *out;                 // HBM: [DFL_OUT_WORDS], zero on entry
};

// phase 1.  History first: every position of the previous block enters the table (any order: the maximum wins).  Then
// one time slice after the other: p = slice * DFL_THREADS + t, look-up first (barrier), then enter (barrier).
// Positions count from the start of the history: a = hist + p.
DFL_HD void dfl_hist_enter(const DflShared &S, int a)
{
    if (a < S.hist) DFL_ATOMIC_MAX(&S.table[dfl_hash4(S.buf - S.hist + a)], (uint32_t)(a + 1));
}
DFL_HD void dfl_cand_lookup(const DflShared &S, int p)
{
    if (p < S.len) S.cand[p] = (p + DFL_MIN_HASH_MATCH <= S.len) ? (uint16_t)S.table[dfl_hash4(S.buf + p)] : (uint16_t)0;
}
DFL_HD void dfl_cand_enter(const DflShared &S, int p)
{
    if (p + DFL_MIN_HASH_MATCH <= S.len) DFL_ATOMIC_MAX(&S.table[dfl_hash4(S.buf + p)], (uint32_t)(S.hist + p + 1));
}

// the best match at p (it may run past the thread's own bytes: dfl_stitch trims what follows): longest of the hash chain (at most DFL_MAX_CHAIN candidates, nearest
// first) and of distance 1 (runs; preferred on ties: the cheapest distance)
DFL_HD int dfl_find_match(const DflShared &S, int p, int &bdist)
{
    const int maxlen = S.len - p < DFL_MAX_MATCH ? S.len - p : DFL_MAX_MATCH;
    int best = 0;
    bdist = 0;
    // cand[] links every position to the head of its hash before its own time slice: a chain of earlier positions
    // with the same hash, nearest first (deterministic: it does not depend on the order threads ran in)
    if (maxlen >= DFL_MIN_HASH_MATCH) {
        const uint8_t *ext = S.buf - S.hist;
        int c = S.cand[p];
        for (int tries = 0; c != 0 && tries < DFL_MAX_CHAIN; tries++) {
            const int a = c - 1, dist = S.hist + p - a;
            if (dist > 32768) break;                                    // older ones are farther still
            const uint8_t *q = ext + a;
            if (q[best] == S.buf[p + best]) {                           // can only win if it matches one byte further
                const int l = dfl_match_len(q, S.buf + p, maxlen);
                if (l >= DFL_MIN_HASH_MATCH && l > best) { best = l; bdist = dist; }
                if (best >= DFL_NICE_MATCH || best == maxlen) break;
            }
            if (a < S.hist) break;                                      // the history has no links of its own
            c = S.cand[a - S.hist];
        }
    }
    if (S.hist + p > 0 && maxlen >= DFL_MIN_RUN) {
        const int l = dfl_match_len(S.buf + p - 1, S.buf + p, maxlen);
        if (l >= DFL_MIN_RUN && l >= best) { best = l; bdist = 1; }
    }
    return best;
}

// phase 2a: parse of thread t's bytes, greedy with one step of lazy evaluation (a longer match one byte later wins).
// The last match may run into the following threads' bytes; tend[t] says how far.
DFL_HD void dfl_parse(const DflShared &S, int t)
{
    int p = t * DFL_SUB;
    const int end = p + DFL_SUB < S.len ? p + DFL_SUB : S.len;
    uint32_t *tok = S.tokens + t * DFL_TOKS + 2;
    uint32_t nt = 0;
    int bdist = 0, best = p < end ? dfl_find_match(S, p, bdist) : 0;
    while (p < end) {
        bool literal = best == 0;
        int ndist = 0, nbest = 0;
        if (!literal && best < DFL_LAZY_MAX && p + 1 < end) {
            nbest = dfl_find_match(S, p + 1, ndist);
            if (nbest > best) literal = true;
        }
        if (!literal) {
            tok[nt++] = DFL_TOK_MATCH | ((uint32_t)(best - 3) << 16) | (uint32_t)(bdist - 1);
            p += best;
            best = p < end ? dfl_find_match(S, p, bdist) : 0;
        } else {
            tok[nt++] = S.buf[p];
            p++;
            if (nbest > 0) { best = nbest; bdist = ndist; }              // the match found one byte later
            else best = p < end ? dfl_find_match(S, p, bdist) : 0;
        }
    }
    S.ntok[t] = nt;
    S.tbeg[t] = 2;
    S.tend[t] = (uint32_t)p;
}

// phase 2b: `covered` = how far the threads before t got (maximum of their tend).  Tokens that lie below it are dropped,
// the one that crosses it is cut at the front (a match keeps its distance; fewer than 3 bytes left become literals in the
// two spare slots); then the symbols of what the thread will emit are counted.
DFL_HD void dfl_stitch(const DflShared &S, int t, uint32_t covered)
{
    uint32_t *tok = S.tokens + t * DFL_TOKS;
    uint32_t beg = 2, nt = S.ntok[t];
    uint32_t p = (uint32_t)(t * DFL_SUB);
    while (nt > 0 && p < covered) {
        const uint32_t k = tok[beg];
        const uint32_t l = (k & DFL_TOK_MATCH) ? ((k >> 16) & 0xffu) + 3u : 1u;
        if (p + l <= covered) { p += l; beg++; nt--; continue; }
        // a match crosses the line: [p, p + l) with p < covered < p + l
        const uint32_t left = p + l - covered;
        if (left >= 3u) {
            tok[beg] = DFL_TOK_MATCH | ((left - 3u) << 16) | (k & 0xffffu);
        } else {
            beg++; nt--;
            for (uint32_t j = left; j > 0; j--) { tok[--beg] = S.buf[covered + j - 1]; nt++; }
        }
        break;
    }
    S.tbeg[t] = beg;
    S.ntok[t] = nt;
    for (uint32_t i = 0; i < nt; i++) {
        const uint32_t k = tok[beg + i];
        if (k & DFL_TOK_MATCH) {
            int eb, ev;
            DFL_ATOMIC_ADD(&S.freq_ll[dfl_len_sym((int)((k >> 16) & 0xffu) + 3, eb, ev)], 1u);
            DFL_ATOMIC_ADD(&S.freq_d[dfl_dist_sym((int)(k & 0xffffu) + 1, eb, ev)], 1u);
        } else {
            DFL_ATOMIC_ADD(&S.freq_ll[k], 1u);
        }
    }
}

// phase 3 (one thread): code lengths, codes, block header
DFL_HD void dfl_build_codes(const DflShared &S, DflHuffScratch &hs)
{
    S.freq_ll[256] = 1;                                   // end of block
    dfl_huff_lengths(S.freq_ll, DFL_NLL, 15, S.len_ll, hs);
    dfl_huff_lengths(S.freq_d, DFL_ND, 15, S.len_d, hs);
    dfl_huff_codes(S.len_ll, DFL_NLL, S.code_ll);
    dfl_huff_codes(S.len_d, DFL_ND, S.code_d);
    int hlit = DFL_NLL, hdist = DFL_ND;
    while (hlit > 257 && S.len_ll[hlit - 1] == 0) hlit--;
    while (hdist > 1 && S.len_d[hdist - 1] == 0) hdist--;
    // the code length sequence as (symbol, extra) pairs: lengths 0..15, 17 = 3..10 zeros, 18 = 11..138 zeros
    uint8_t seq_sym[DFL_NLL + DFL_ND], seq_ext[DFL_NLL + DFL_ND];
    uint32_t fcl[DFL_NCL];
    for (int i = 0; i < DFL_NCL; i++) fcl[i] = 0;
    int nseq = 0;
    const int total = hlit + hdist;
    for (int i = 0; i < total;) {
        const int l = i < hlit ? S.len_ll[i] : S.len_d[i - hlit];
        if (l == 0) {
            int run = 1;
            while (i + run < total && run < 138 && (i + run < hlit ? S.len_ll[i + run] : S.len_d[i + run - hlit]) == 0) run++;
            if (run >= 11) { seq_sym[nseq] = 18; seq_ext[nseq] = (uint8_t)(run - 11); nseq++; fcl[18]++; i += run; continue; }
            if (run >= 3) { seq_sym[nseq] = 17; seq_ext[nseq] = (uint8_t)(run - 3); nseq++; fcl[17]++; i += run; continue; }
        }
        seq_sym[nseq] = (uint8_t)l; seq_ext[nseq] = 0; nseq++; fcl[l]++; i++;
    }
    uint8_t len_cl[DFL_NCL];
    uint16_t code_cl[DFL_NCL];
    dfl_huff_lengths(fcl, DFL_NCL, 7, len_cl, hs);
    dfl_huff_codes(len_cl, DFL_NCL, code_cl);
    const int perm[DFL_NCL] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = DFL_NCL;
    while (hclen > 4 && len_cl[perm[hclen - 1]] == 0) hclen--;
    for (int i = 0; i < DFL_HDR_WORDS; i++) S.hdr[i] = 0;
    DflBits b;
    dfl_bits_start(b, S.hdr, 0);
    dfl_bits_put(b, S.final ? 1u : 0u, 1);                // BFINAL
    dfl_bits_put(b, 2u, 2);                               // BTYPE = dynamic Huffman
    dfl_bits_put(b, (uint32_t)(hlit - 257), 5);
    dfl_bits_put(b, (uint32_t)(hdist - 1), 5);
    dfl_bits_put(b, (uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) dfl_bits_put(b, len_cl[perm[i]], 3);
    for (int i = 0; i < nseq; i++) {
        dfl_bits_put(b, code_cl[seq_sym[i]], len_cl[seq_sym[i]]);
        if (seq_sym[i] == 17) dfl_bits_put(b, seq_ext[i], 3);
        else if (seq_sym[i] == 18) dfl_bits_put(b, seq_ext[i], 7);
    }
    const uint32_t nbits = (uint32_t)((b.w - S.hdr) * 32 + b.n);
    dfl_bits_finish(b);
    *S.hdr_bits = nbits;
}

// phase 4: bits of thread t's tokens
DFL_HD void dfl_count_bits(const DflShared &S, int t)
{
    const uint32_t *tok = S.tokens + t * DFL_TOKS + S.tbeg[t];
    const uint32_t nt = S.ntok[t];
    uint32_t nb = 0;
    for (uint32_t i = 0; i < nt; i++) {
        const uint32_t k = tok[i];
        if (k & DFL_TOK_MATCH) {
            int eb, ev;
            nb += S.len_ll[dfl_len_sym((int)((k >> 16) & 0xffu) + 3, eb, ev)] + eb;
            nb += S.len_d[dfl_dist_sym((int)(k & 0xffffu) + 1, eb, ev)] + eb;
        } else {
            nb += S.len_ll[k];
        }
    }
    S.bits[t] = nb;
}

// phase 5: thread t writes its tokens at bit `start` of the output
DFL_HD void dfl_emit(const DflShared &S, int t, uint32_t start)
{
    const uint32_t *tok = S.tokens + t * DFL_TOKS + S.tbeg[t];
    const uint32_t nt = S.ntok[t];
    DflBits b;
    dfl_bits_start(b, S.out, start);
    for (uint32_t i = 0; i < nt; i++) {
        const uint32_t k = tok[i];
        if (k & DFL_TOK_MATCH) {
            int eb, ev;
            const int ls = dfl_len_sym((int)((k >> 16) & 0xffu) + 3, eb, ev);
            dfl_bits_put(b, S.code_ll[ls], S.len_ll[ls]);
            if (eb) dfl_bits_put(b, (uint32_t)ev, eb);
            const int ds = dfl_dist_sym((int)(k & 0xffffu) + 1, eb, ev);
            dfl_bits_put(b, S.code_d[ds], S.len_d[ds]);
            if (eb) dfl_bits_put(b, (uint32_t)ev, eb);
        } else {
            dfl_bits_put(b, S.code_ll[k], S.len_ll[k]);
        }
    }
    dfl_bits_finish(b);
}

// A block that is not the last of its member ends on a byte boundary: an empty stored block (BFINAL = 0, BTYPE = 00,
// padding, LEN = 0, NLEN = 0xffff -- zlib's Z_SYNC_FLUSH) follows its end-of-block code.  The words are zero already, so
// only the two 0xff bytes are written.  Returns the block's length in bytes; `total_bits` = bits up to the end-of-block code.
DFL_HD uint32_t dfl_block_bytes(uint32_t total_bits, int final)
{
    if (final) return (total_bits + 7u) >> 3;
    return ((total_bits + 3u + 7u) >> 3) + 4u;
}
DFL_HD void dfl_sync_marker(uint32_t *out, uint32_t total_bits)
{
    const uint32_t at = ((total_bits + 3u + 7u) >> 3) + 2u;          // byte offset of NLEN
    for (uint32_t b = at; b < at + 2u; b++) DFL_ATOMIC_OR(&out[b >> 2], 0xffu << (8u * (b & 3u)));
}

// ---- CRC-32 (gzip): byte-wise register, and the advance of a register over m more bytes ----
DFL_HD uint32_t dfl_crc_table_entry(uint32_t i)
{
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ DFL_CRC_POLY : c >> 1;
    return c;
}
// a(x) b(x) mod P(x) in the reflected representation (bit 31 = x^0), as zlib's multmodp
DFL_HD uint32_t dfl_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ DFL_CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 m) mod P by square and multiply (m < 2^24)
DFL_HD uint32_t dfl_xpow8(uint32_t m)
{
    uint32_t r = 1u << 31, sq = 1u << 23;                 // x^0, x^8
    while (m) {
        if (m & 1u) r = dfl_multmodp(sq, r);
        sq = dfl_multmodp(sq, sq);
        m >>= 1;
    }
    return r;
}
// register after thread t's bytes (thread 0 starts from the gzip preset), advanced to the end of the chunk
DFL_HD uint32_t dfl_crc_part(const uint8_t *buf, int len, int t, const uint32_t *table)
{
    const int p0 = t * DFL_SUB;
    if (p0 >= len && t > 0) return 0u;
    const int end = p0 + DFL_SUB < len ? p0 + DFL_SUB : len;
    uint32_t c = t == 0 ? 0xffffffffu : 0u;
    for (int p = p0; p < end; p++) c = table[(c ^ buf[p]) & 0xffu] ^ (c >> 8);
    const int after = len - end;
    if (after > 0 && c != 0u) c = dfl_multmodp(dfl_xpow8((uint32_t)after), c);
    return c;
}

// CRC-32 of A || B from the (finalised) CRC-32s of A and B and B's length: zlib's crc32_combine
DFL_HD uint32_t dfl_crc_combine(uint32_t crc_a, uint32_t crc_b, uint32_t len_b)
{
    if (len_b == 0) return crc_a;
    return dfl_multmodp(dfl_xpow8(len_b), crc_a) ^ crc_b;
}

#endif
