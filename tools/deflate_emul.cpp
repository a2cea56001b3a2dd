// Host emulation of deflate_kernel (itsxpress_b200/csrc/deflate.cu): the same per-thread bodies (deflate_core.h), run
// one "thread" after the other with the CTA barriers as loop boundaries, framed like gz_frame_kernel.  Development
// tool for a container without a GPU: the output must inflate (zlib / gzip) to the input.
//   g++ -O2 -o /tmp/deflate_emul tools/deflate_emul.cpp && /tmp/deflate_emul in.txt out.gz
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include "../itsxpress_b200/csrc/deflate_core.h"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: deflate_emul IN OUT.gz\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    std::vector<uint8_t> in;
    uint8_t tmp[1 << 16];
    size_t r;
    while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) in.insert(in.end(), tmp, tmp + r);
    fclose(f);
    const int64_t n = (int64_t)in.size();
    FILE *o = fopen(argv[2], "wb");
    uint32_t crc_table[256];
    for (uint32_t i = 0; i < 256; i++) crc_table[i] = dfl_crc_table_entry(i);
    long long stored_blocks = 0, total_out = 0;
    const int64_t chunks = n ? (n + DFL_CHUNK - 1) / DFL_CHUNK : 1;
    uint32_t member_crc = 0, member_len = 0;
    for (int64_t j = 0; j < chunks; j++) {
        const int len = (int)std::min<int64_t>(DFL_CHUNK, n - j * DFL_CHUNK);
        const bool first = j % DFL_GROUP == 0, final = j % DFL_GROUP == DFL_GROUP - 1 || j == chunks - 1;
        const int hist = first ? 0 : DFL_HIST;
        std::vector<uint8_t> ext(DFL_HIST + DFL_CHUNK + 16, 0);
        if (hist) memcpy(ext.data() + DFL_HIST - hist, in.data() + j * DFL_CHUNK - hist, (size_t)hist);
        if (len > 0) memcpy(ext.data() + DFL_HIST, in.data() + j * DFL_CHUNK, (size_t)len);
        const uint8_t *buf = ext.data() + DFL_HIST;
        std::vector<uint16_t> cand(DFL_CHUNK, 0), code_ll(288, 0), code_d(32, 0);
        std::vector<uint32_t> table(DFL_HASH_SIZE, 0), freq_ll(288, 0), freq_d(32, 0), ntok(DFL_THREADS, 0),
            bits(DFL_THREADS + 1, 0), hdr(DFL_HDR_WORDS, 0), tokens(DFL_THREADS * DFL_TOKS, 0), out(DFL_OUT_WORDS, 0),
            tbeg(DFL_THREADS, 0), tend(DFL_THREADS, 0);
        std::vector<uint8_t> len_ll(288, 0), len_d(32, 0);
        uint32_t hdr_bits = 0;
        DflShared S;
        S.buf = buf; S.len = len; S.hist = hist; S.final = final ? 1 : 0; S.cand = cand.data(); S.table = table.data();
        S.freq_ll = freq_ll.data(); S.freq_d = freq_d.data(); S.len_ll = len_ll.data(); S.len_d = len_d.data();
        S.code_ll = code_ll.data(); S.code_d = code_d.data(); S.ntok = ntok.data(); S.bits = bits.data();
        S.hdr = hdr.data(); S.hdr_bits = &hdr_bits; S.tokens = tokens.data(); S.out = out.data();
        S.tbeg = tbeg.data(); S.tend = tend.data();
        for (int a0 = 0; a0 < hist; a0 += DFL_THREADS)
            for (int t = 0; t < DFL_THREADS; t++) dfl_hist_enter(S, a0 + t);
        for (int p0 = 0; p0 < len; p0 += DFL_THREADS) {
            for (int t = 0; t < DFL_THREADS; t++) dfl_cand_lookup(S, p0 + t);
            for (int t = 0; t < DFL_THREADS; t++) dfl_cand_enter(S, p0 + t);
        }
        for (int t = 0; t < DFL_THREADS; t++) dfl_parse(S, t);
        {
            std::vector<uint32_t> covered(DFL_THREADS, 0);          // exclusive prefix maximum of tend[]
            uint32_t mx = 0;
            for (int t = 0; t < DFL_THREADS; t++) { covered[t] = mx; mx = std::max(mx, tend[t]); }
            for (int t = 0; t < DFL_THREADS; t++) dfl_stitch(S, t, covered[t]);
        }
        static DflHuffScratch hs;
        dfl_build_codes(S, hs);
        for (int t = 0; t < DFL_THREADS; t++) dfl_count_bits(S, t);
        uint32_t run = 0;
        for (int t = 0; t < DFL_THREADS; t++) { const uint32_t b = bits[t]; bits[t] = run; run += b; }
        bits[DFL_THREADS] = run;
        const uint32_t total_bits = hdr_bits + run + len_ll[256];
        const uint32_t dyn_bytes = dfl_block_bytes(total_bits, S.final);
        const bool stored = dyn_bytes >= (uint32_t)len + 5u;
        uint32_t nb;
        if (!stored) {
            for (uint32_t w = 0; w < (hdr_bits + 31) >> 5; w++) out[w] |= hdr[w];
            for (int t = 0; t < DFL_THREADS; t++) dfl_emit(S, t, hdr_bits + bits[t]);
            DflBits b;
            dfl_bits_start(b, S.out, hdr_bits + run);
            dfl_bits_put(b, code_ll[256], len_ll[256]);
            dfl_bits_finish(b);
            if (!S.final) dfl_sync_marker(S.out, total_bits);
            nb = dyn_bytes;
        } else {
            uint8_t *ob = (uint8_t *)out.data();
            ob[0] = (uint8_t)S.final; ob[1] = (uint8_t)(len & 0xff); ob[2] = (uint8_t)(len >> 8);
            ob[3] = (uint8_t)(~len & 0xff); ob[4] = (uint8_t)((~len >> 8) & 0xff);
            memcpy(ob + 5, buf, (size_t)len);
            nb = (uint32_t)len + 5;
            stored_blocks++;
        }
        uint32_t crc = 0;
        for (int t = 0; t < DFL_THREADS; t++) crc ^= dfl_crc_part(buf, len, t, crc_table);
        crc ^= 0xffffffffu;
        if (first) {
            const uint8_t gh[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
            fwrite(gh, 1, 10, o);
            total_out += 10;
            member_crc = crc; member_len = (uint32_t)len;
        } else {
            member_crc = dfl_crc_combine(member_crc, crc, (uint32_t)len);
            member_len += (uint32_t)len;
        }
        fwrite(out.data(), 1, nb, o);
        total_out += nb;
        if (final) {
            uint8_t tr[8];
            for (int k = 0; k < 4; k++) { tr[k] = (uint8_t)(member_crc >> (8 * k)); tr[4 + k] = (uint8_t)(member_len >> (8 * k)); }
            fwrite(tr, 1, 8, o);
            total_out += 8;
        }
    }
    fclose(o);
    fprintf(stderr, "%lld bytes in, %lld out (%.3f), %lld blocks in %lld members, %lld stored\n", (long long)n, total_out,
            n ? (double)total_out / (double)n : 0.0, (long long)chunks, (long long)((chunks + DFL_GROUP - 1) / DFL_GROUP),
            stored_blocks);
    return 0;
}
