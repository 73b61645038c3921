// vfkio_inflate.cpp -- raw DEFLATE (RFC 1951) decoder for whole BGZF blocks.
//
// The BAM decoder is bound by inflate (tools/bench_decode.py: zlib runs at ~180 MB/s per core on the build
// host, 60 % of a decode).  A BGZF member is at most 64 KB, is always inflated in one go into a buffer of known
// size, and never needs a sliding window or streaming state -- so this decoder drops all of zlib's generality:
// a 64-bit bit buffer refilled with one unaligned load, an 11-bit first-level literal/length table (one lookup
// for nearly every symbol, second level for the longer codes), an 8-bit distance table, word-wise match copies.
// Written from the RFC; every refusal (malformed or unusual stream, any bound that would be crossed) returns
// non-zero WITHOUT having written outside [out, out + out_n), and the caller falls back to zlib, which then
// gives the authoritative verdict.  tests/test_formats.py checks it against zlib on every block of the test BAMs,
// on all compression levels and strategies, and on damaged streams.
#include <cstdint>
#include <cstring>
#include "vfk_io.h"

namespace {
constexpr int LIT_ROOT = 11, DIST_ROOT = 8;
constexpr uint32_t F_INVALID = 1u << 7, F_LITERAL = 1u << 13, F_EOB = 1u << 14, F_SUB = 1u << 15;
// entry: bits 0-4 bits to consume | F_INVALID | bits 8-12 extra-bit count (sub-table: its index width) | flags |
//        bits 16-31 literal / base length / base distance / first entry of the sub-table
constexpr int LIT_CAP = (1 << LIT_ROOT) + 288 * 16, DIST_CAP = (1 << DIST_ROOT) + 32 * 128;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t litlen_entry(int sym) {
    if (sym < 256) return F_LITERAL | ((uint32_t)sym << 16);
    if (sym == 256) return F_EOB;
    if (sym <= 285) return ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_EXTRA[sym - 257] << 8);
    return F_INVALID;
}
inline uint32_t dist_entry(int sym) {
    if (sym < 30) return ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_EXTRA[sym] << 8);
    return F_INVALID;
}
inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// Canonical Huffman code (RFC 1951 3.2.2) -> lookup table indexed by the next bits of the stream (LSB first).
// Refuses over-subscribed sets and incomplete ones (zlib accepts an incomplete set only for a single code; the
// fallback settles those).  An empty set (no distance codes: a block of literals) yields a table of invalid entries.
template <int ROOT, int CAP, bool LITLEN>
bool build_table(const uint8_t *lens, int n, uint32_t *table) {
    int count[16] = {0};
    for (int s = 0; s < n; ++s) ++count[lens[s]];
    const int used = n - count[0];
    if (used == 0) {
        if (LITLEN) return false;
        for (int i = 0; i < (1 << ROOT); ++i) table[i] = F_INVALID | 1u;
        return true;
    }
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left = left * 2 - count[l];
        if (left < 0) return false;
    }
    if (left > 0) return false;
    uint32_t next[16];
    {
        uint32_t code = 0;
        count[0] = 0;
        for (int l = 1; l <= 15; ++l) {
            code = (code + (uint32_t)count[l - 1]) << 1;
            next[l] = code;
        }
    }
    uint8_t sub_max[1 << ROOT];
    uint16_t sub_at[1 << ROOT];
    bool any_long = false;
    uint32_t rev_of[288];
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next[l]++, l);
        rev_of[s] = rev;
        const uint32_t e = (LITLEN ? litlen_entry(s) : dist_entry(s));
        if (l <= ROOT) {
            for (uint32_t i = rev; i < (1u << ROOT); i += 1u << l) table[i] = e | (uint32_t)l;
        } else {
            if (!any_long) {
                memset(sub_max, 0, sizeof(sub_max));
                any_long = true;
            }
            const uint32_t prefix = rev & ((1u << ROOT) - 1u);
            if (l > sub_max[prefix]) sub_max[prefix] = (uint8_t)l;
        }
    }
    if (!any_long) return true;
    int free_at = 1 << ROOT;
    for (uint32_t p = 0; p < (1u << ROOT); ++p) {
        if (!sub_max[p]) continue;
        const int bits = sub_max[p] - ROOT;
        if (free_at + (1 << bits) > CAP) return false;
        sub_at[p] = (uint16_t)free_at;
        table[p] = F_SUB | ((uint32_t)free_at << 16) | ((uint32_t)bits << 8) | (uint32_t)ROOT;
        free_at += 1 << bits;
    }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l <= ROOT) continue;
        const uint32_t rev = rev_of[s], p = rev & ((1u << ROOT) - 1u);
        const int bits = sub_max[p] - ROOT;
        const uint32_t e = (LITLEN ? litlen_entry(s) : dist_entry(s)) | (uint32_t)(l - ROOT);
        for (uint32_t i = rev >> ROOT; i < (1u << bits); i += 1u << (l - ROOT)) table[sub_at[p] + i] = e;
    }
    return true;
}

struct Bits {
    const uint8_t *next, *end;
    uint64_t buf = 0;
    int cnt = 0;  // valid bits in buf
    inline void refill() {
        if (end - next >= 8) {
            uint64_t w;
            memcpy(&w, next, 8);  // little endian host (x86-64 / aarch64)
            buf |= w << cnt;
            next += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && next < end) {
                buf |= (uint64_t)*next++ << cnt;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
};
}  // namespace

extern "C" int vfkio_inflate_raw(const uint8_t *in, size_t in_n, uint8_t *out, size_t out_n) {
    if ((!in && in_n) || (!out && out_n)) return -1;
    Bits b;
    b.next = in;
    b.end = in + in_n;
    uint8_t *o = out, *const o_end = out + out_n;
    uint32_t lit[LIT_CAP], dist[DIST_CAP];
    uint8_t lens[288 + 32];
    for (int final_block = 0; !final_block;) {
        b.refill();
        if (b.cnt < 3) return -1;
        final_block = (int)b.peek(1);
        const uint32_t type = (b.peek(3) >> 1);
        b.drop(3);
        if (type == 0) {  // stored: back to the byte boundary, LEN / NLEN, raw bytes
            b.drop(b.cnt & 7);
            const uint8_t *p = b.next - (b.cnt >> 3);  // bytes still in the bit buffer are handed back
            b.buf = 0;
            b.cnt = 0;
            if (b.end - p < 4) return -1;
            const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8), nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
            if ((len ^ nlen) != 0xFFFFu) return -1;
            p += 4;
            if ((size_t)(b.end - p) < len || (size_t)(o_end - o) < len) return -1;
            memcpy(o, p, len);
            o += len;
            b.next = p + len;
            continue;
        }
        if (type == 3) return -1;
        if (type == 1) {  // fixed code
            for (int s = 0; s < 144; ++s) lens[s] = 8;
            for (int s = 144; s < 256; ++s) lens[s] = 9;
            for (int s = 256; s < 280; ++s) lens[s] = 7;
            for (int s = 280; s < 288; ++s) lens[s] = 8;
            for (int s = 0; s < 32; ++s) lens[288 + s] = 5;
            if (!build_table<LIT_ROOT, LIT_CAP, true>(lens, 288, lit)) return -1;
            if (!build_table<DIST_ROOT, DIST_CAP, false>(lens + 288, 32, dist)) return -1;
        } else {  // dynamic code: the code lengths are themselves Huffman coded (3.2.7)
            if (b.cnt < 14) return -1;
            const int hlit = (int)b.peek(5) + 257, hdist = (int)(b.peek(10) >> 5) + 1, hclen = (int)(b.peek(14) >> 10) + 4;
            b.drop(14);
            if (hlit > 286 || hdist > 30) return -1;
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; ++i) {
                if (b.cnt < 3) b.refill();
                if (b.cnt < 3) return -1;
                cl[CL_ORDER[i]] = (uint8_t)b.peek(3);
                b.drop(3);
            }
            // code-length code: at most 7 bits, one flat table of 128
            uint8_t cl_sym[128], cl_len[128];
            {
                int count[8] = {0};
                for (int s = 0; s < 19; ++s) ++count[cl[s]];
                if (count[0] == 19) return -1;
                int left = 1;
                for (int l = 1; l <= 7; ++l) {
                    left = left * 2 - count[l];
                    if (left < 0) return -1;
                }
                if (left > 0) return -1;
                uint32_t next[8], code = 0;
                count[0] = 0;
                for (int l = 1; l <= 7; ++l) {
                    code = (code + (uint32_t)count[l - 1]) << 1;
                    next[l] = code;
                }
                memset(cl_len, 0, sizeof(cl_len));
                for (int s = 0; s < 19; ++s) {
                    const int l = cl[s];
                    if (!l) continue;
                    const uint32_t rev = reverse_bits(next[l]++, l);
                    for (uint32_t i = rev; i < 128u; i += 1u << l) {
                        cl_sym[i] = (uint8_t)s;
                        cl_len[i] = (uint8_t)l;
                    }
                }
            }
            const int total = hlit + hdist;
            int i = 0;
            uint8_t all[286 + 30 + 138];
            while (i < total) {
                if (b.cnt < 14) b.refill();
                const uint32_t idx = b.peek(7);
                const int l = cl_len[idx], s = cl_sym[idx];
                if (!l || l > b.cnt) return -1;
                b.drop(l);
                if (s < 16) {
                    all[i++] = (uint8_t)s;
                    continue;
                }
                int rep, extra;
                uint8_t v = 0;
                if (s == 16) {
                    if (i == 0) return -1;
                    v = all[i - 1];
                    extra = 2;
                    rep = 3;
                } else if (s == 17) {
                    extra = 3;
                    rep = 3;
                } else {
                    extra = 7;
                    rep = 11;
                }
                if (b.cnt < extra) return -1;
                rep += (int)b.peek(extra);
                b.drop(extra);
                if (i + rep > total) return -1;
                memset(all + i, v, (size_t)rep);
                i += rep;
            }
            if (all[256] == 0) return -1;  // no end-of-block code
            memset(lens, 0, sizeof(lens));
            memcpy(lens, all, (size_t)hlit);
            memcpy(lens + 288, all + hlit, (size_t)hdist);
            if (!build_table<LIT_ROOT, LIT_CAP, true>(lens, 288, lit)) return -1;
            if (!build_table<DIST_ROOT, DIST_CAP, false>(lens + 288, 32, dist)) return -1;
        }
        // ---- symbols of the block.  One refill guarantees 56 bits (while 8 input bytes remain): a length + distance
        // pair needs at most 15 + 5 + 15 + 13 = 48.
        bool eob = false;
        while (!eob) {
            // fast loop: 8 input bytes and a maximal match plus a word of slack are available, so no bit count and no
            // output bound needs checking inside
            while (b.end - b.next >= 8 && o_end - o >= 274) {
                {
                    uint64_t w;
                    memcpy(&w, b.next, 8);
                    b.buf |= w << b.cnt;
                    b.next += (63 - b.cnt) >> 3;
                    b.cnt |= 56;
                }
                uint32_t e = lit[b.buf & ((1u << LIT_ROOT) - 1u)];
                if (e & F_LITERAL) {  // up to three first-level literals (<= 11 bits each) per refill
                    b.drop((int)(e & 31u));
                    *o++ = (uint8_t)(e >> 16);
                    e = lit[b.buf & ((1u << LIT_ROOT) - 1u)];
                    if (e & F_LITERAL) {
                        b.drop((int)(e & 31u));
                        *o++ = (uint8_t)(e >> 16);
                        e = lit[b.buf & ((1u << LIT_ROOT) - 1u)];
                        if (e & F_LITERAL) {
                            b.drop((int)(e & 31u));
                            *o++ = (uint8_t)(e >> 16);
                        }
                    }
                    continue;
                }
                if (e & F_SUB) {
                    b.drop(LIT_ROOT);
                    e = lit[(e >> 16) + b.peek((int)((e >> 8) & 31u))];
                }
                if ((e & F_INVALID) || (e & 31u) == 0) return -1;
                b.drop((int)(e & 31u));
                if (e & F_LITERAL) {
                    *o++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (e & F_EOB) {
                    eob = true;
                    break;
                }
                int x = (int)((e >> 8) & 31u);
                const uint32_t length = (e >> 16) + b.peek(x);
                b.drop(x);
                uint32_t d = dist[b.buf & ((1u << DIST_ROOT) - 1u)];
                if (d & F_SUB) {
                    b.drop(DIST_ROOT);
                    d = dist[(d >> 16) + b.peek((int)((d >> 8) & 31u))];
                }
                if ((d & F_INVALID) || (d & 31u) == 0) return -1;
                b.drop((int)(d & 31u));
                x = (int)((d >> 8) & 31u);
                const uint32_t distance = (d >> 16) + b.peek(x);
                b.drop(x);
                if (distance > (size_t)(o - out)) return -1;
                const uint8_t *src = o - distance;
                uint8_t *dst = o;
                o += length;
                if (distance >= 8) {
                    do {
                        uint64_t w;
                        memcpy(&w, src, 8);
                        memcpy(dst, &w, 8);
                        src += 8;
                        dst += 8;
                    } while (dst < o);
                } else if (distance == 1) {
                    const uint64_t w = 0x0101010101010101ull * (uint64_t)*src;
                    do {
                        memcpy(dst, &w, 8);
                        dst += 8;
                    } while (dst < o);
                } else {
                    do {
                        *dst++ = *src++;
                    } while (dst < o);
                }
            }
            if (eob) break;
            // careful step: one symbol with every bound checked (the last bytes of the input or of the output)
            if (b.cnt < 48) b.refill();
            uint32_t e = lit[b.peek(LIT_ROOT)];
            if (e & F_SUB) {
                if (b.cnt < LIT_ROOT) return -1;
                b.drop(LIT_ROOT);
                e = lit[(e >> 16) + b.peek((int)((e >> 8) & 31u))];
            }
            int l = (int)(e & 31u);
            if ((e & F_INVALID) || l == 0 || l > b.cnt) return -1;
            b.drop(l);
            if (e & F_LITERAL) {
                if (o == o_end) return -1;
                *o++ = (uint8_t)(e >> 16);
                continue;
            }
            if (e & F_EOB) break;
            int x = (int)((e >> 8) & 31u);
            if (x > b.cnt) return -1;
            const uint32_t length = (e >> 16) + b.peek(x);
            b.drop(x);
            uint32_t d = dist[b.peek(DIST_ROOT)];
            if (d & F_SUB) {
                if (b.cnt < DIST_ROOT) return -1;
                b.drop(DIST_ROOT);
                d = dist[(d >> 16) + b.peek((int)((d >> 8) & 31u))];
            }
            l = (int)(d & 31u);
            if ((d & F_INVALID) || l == 0 || l > b.cnt) return -1;
            b.drop(l);
            x = (int)((d >> 8) & 31u);
            if (x > b.cnt) return -1;
            const uint32_t distance = (d >> 16) + b.peek(x);
            b.drop(x);
            if (distance > (size_t)(o - out) || length > (size_t)(o_end - o)) return -1;
            const uint8_t *src = o - distance;
            for (uint32_t k = 0; k < length; ++k) o[k] = src[k];
            o += length;
        }
    }
    return o == o_end ? 0 : -1;  // the BGZF trailer states the size: anything else is an error
}
