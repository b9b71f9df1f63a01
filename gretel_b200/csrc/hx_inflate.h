// Whole-buffer raw DEFLATE (RFC 1951) decoder for BGZF blocks: the compressed and the inflated size are both known
// (BAM's BGZF frames carry them), so there is no streaming state, no window and no allocation.  Replaces zlib's
// inflate() in the BAM packer (bampack.cpp), where inflating was 65 % of util.load_from_bam; zlib stays as the
// verdict on anything this decoder declines (hx_inflate returns false: corrupt or unusual stream -> the caller asks
// zlib and reports zlib's answer).
//
// How it goes fast: a 64-bit bit buffer refilled without branches from unaligned 8-byte loads; one 2048-entry
// table lookup per literal / length symbol (codes longer than 11 bits go through a second-level table), 256 entries
// for distances; up to three literals per refill; matches copied eight bytes at a time while the output has room
// for a whole maximum-length match, byte by byte in the last few hundred bytes of the block.
#pragma once
#include <stdint.h>
#include <string.h>

namespace hxz {

constexpr int LTB = 11;                 // litlen primary table bits
constexpr int DTB = 8;                  // distance primary table bits
constexpr int PTB = 7;                  // precode (code-length code) table bits
constexpr int L_CAP = (1 << LTB) + 286 * 16;
constexpr int D_CAP = (1 << DTB) + 30 * 128;

// table entry: bits 0-5 input bits to consume (a shift count as it stands), bits 8-12 extra bits (or second-level bits
// for T_SUB), bits 13-15 type, bits 16-31 literal / base length / base distance / second-level table offset;
// T_LIT2 = two literals decoded by one lookup (first in bits 16-23, second in bits 24-31; litlen primary table only)
enum : uint32_t { T_LIT = 0u << 13, T_LIT2 = 1u << 13, T_LEN = 2u << 13, T_EOB = 3u << 13, T_SUB = 4u << 13, T_BAD = 5u << 13,
                  T_MASK = 7u << 13 };

struct Tables {
    uint32_t lit[L_CAP];
    uint32_t dist[D_CAP];
};

static inline uint32_t rev_bits(uint32_t c, int len) {
    c = ((c & 0x5555u) << 1) | ((c >> 1) & 0x5555u);
    c = ((c & 0x3333u) << 2) | ((c >> 2) & 0x3333u);
    c = ((c & 0x0f0fu) << 4) | ((c >> 4) & 0x0f0fu);
    c = ((c & 0x00ffu) << 8) | ((c >> 8) & 0x00ffu);
    return c >> (16 - len);
}

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

enum Kind { K_PRE, K_LIT, K_DIST };

static inline uint32_t sym_entry(Kind kind, int sym) {
    if (kind == K_PRE) return T_LIT | ((uint32_t)sym << 16);
    if (kind == K_LIT) {
        if (sym < 256) return T_LIT | ((uint32_t)sym << 16);
        if (sym == 256) return T_EOB;
        if (sym < 286) return T_LEN | ((uint32_t)kLenExtra[sym - 257] << 8) | ((uint32_t)kLenBase[sym - 257] << 16);
        return T_BAD;
    }
    if (sym < 30) return T_LEN | ((uint32_t)kDistExtra[sym] << 8) | ((uint32_t)kDistBase[sym] << 16);
    return T_BAD;
}

// Canonical Huffman code lengths -> decode table indexed by the next `tb` input bits (LSB first).  Holes of an
// incomplete code decode to T_BAD.  false: over-subscribed code or not enough room.
static inline bool build_table(const uint8_t *lens, int nsym, Kind kind, int tb, uint32_t *table, int cap) {
    int count[16] = {0};
    for (int s = 0; s < nsym; ++s) count[lens[s]]++;
    count[0] = 0;
    int left = 1;
    uint32_t next[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        code = (code + (uint32_t)count[l - 1]) << 1;
        next[l] = code;
    }
    const int psize = 1 << tb;
    for (int i = 0; i < psize; ++i) table[i] = T_BAD | 1u;
    uint8_t sub_max[1 << LTB];
    bool any_long = false;
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = rev_bits(next[l]++, l);
        if (l <= tb) {
            const uint32_t e = sym_entry(kind, s) | (uint32_t)l;
            for (uint32_t i = r; i < (uint32_t)psize; i += 1u << l) table[i] = e;
        } else {
            if (!any_long) { memset(sub_max, 0, (size_t)psize); any_long = true; }
            const uint32_t p = r & (uint32_t)(psize - 1);
            if (sub_max[p] < l) sub_max[p] = (uint8_t)l;
        }
    }
    if (kind == K_LIT) {
        // pairs of literals whose codes fit the primary index together: one lookup, one shift, two bytes
        for (int i = psize - 1; i >= 0; --i) {                       // downwards: table[i >> l1] is still a single literal
            const uint32_t e1 = table[i];
            if ((e1 & T_MASK) != T_LIT) continue;
            const int l1 = (int)(e1 & 63);
            const uint32_t e2 = table[(uint32_t)i >> l1];          // the bits behind the first code, zero-extended ...
            if ((e2 & T_MASK) != T_LIT) continue;
            const int l2 = (int)(e2 & 63);
            if ((e2 >> 16) > 255 || l1 + l2 > tb) continue;         // ... which decide the second code only if it fits
            table[i] = T_LIT2 | (uint32_t)(l1 + l2) | (e1 & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8);
        }
    }
    if (!any_long) return true;
    // second level: one table per primary prefix that long codes share, as wide as its longest code needs
    code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    int used = psize;
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = rev_bits(next[l]++, l);
        if (l <= tb) continue;
        const uint32_t p = r & (uint32_t)(psize - 1);
        const int sb = sub_max[p] - tb;
        if ((table[p] & T_MASK) != T_SUB) {
            if (used + (1 << sb) > cap) return false;
            table[p] = T_SUB | (uint32_t)tb | ((uint32_t)sb << 8) | ((uint32_t)used << 16);
            for (int i = 0; i < (1 << sb); ++i) table[used + i] = T_BAD | 1u;
            used += 1 << sb;
        }
        const uint32_t base = table[p] >> 16;
        const uint32_t e = sym_entry(kind, s) | (uint32_t)(l - tb);
        for (uint32_t i = r >> tb; i < (1u << sb); i += 1u << (l - tb)) table[base + i] = e;
    }
    return true;
}

static inline const Tables *fixed_tables() {
    static const Tables *t = [] {
        Tables *x = new Tables();
        uint8_t l[288];
        for (int i = 0; i < 144; ++i) l[i] = 8;
        for (int i = 144; i < 256; ++i) l[i] = 9;
        for (int i = 256; i < 280; ++i) l[i] = 7;
        for (int i = 280; i < 288; ++i) l[i] = 8;
        build_table(l, 288, K_LIT, LTB, x->lit, L_CAP);
        uint8_t d[32];
        for (int i = 0; i < 32; ++i) d[i] = 5;
        build_table(d, 32, K_DIST, DTB, x->dist, D_CAP);
        return x;
    }();
    return t;
}

static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void st64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

// src[0..n) raw DEFLATE, n_readable >= n bytes may be READ from src (BGZF: the 8-byte frame trailer follows);
// dst[0..m) receives exactly m bytes.  true iff the stream is well formed, ends with its final block and inflates
// to exactly m bytes having consumed at most n bytes.
static inline bool inflate_raw(const uint8_t *src, size_t n, size_t n_readable, uint8_t *dst, size_t m) {
    const uint8_t *in = src;
    const uint8_t *const in_end = src + n;
    const uint8_t *const in_fast = n_readable >= 8 ? src + (n_readable - 8) : src;   // 8-byte loads allowed up to here
    const bool can_fast = n_readable >= 8;
    uint8_t *out = dst;
    uint8_t *const out_end = dst + m;
    uint64_t bb = 0;
    int bc = 0;
    size_t over = 0;                                   // zero bytes fed past the end of the input

#define HXZ_REFILL()                                                          \
    do {                                                                      \
        if (can_fast && in <= in_fast) {                                      \
            bb |= ld64(in) << bc;                                             \
            in += (63 - bc) >> 3;                                             \
            bc |= 56;                                                         \
        } else {                                                              \
            while (bc <= 56) {                                                \
                if (in < in_end) bb |= (uint64_t)(*in++) << bc; else ++over;  \
                bc += 8;                                                      \
            }                                                                 \
        }                                                                     \
    } while (0)
#define HXZ_DROP(k) do { bb >>= (k); bc -= (int)(k); } while (0)

    Tables dyn;
    bool last = false;
    while (!last) {
        HXZ_REFILL();
        last = bb & 1;
        const int type = (int)((bb >> 1) & 3);
        HXZ_DROP(3);
        if (type == 0) {                                // stored: back to a byte boundary, LEN, NLEN, bytes
            HXZ_DROP(bc & 7);
            HXZ_REFILL();
            const uint32_t len = (uint32_t)(bb & 0xffff), nlen = (uint32_t)((bb >> 16) & 0xffff);
            if ((len ^ nlen) != 0xffff) return false;
            HXZ_DROP(32);
            // the bit buffer holds whole bytes: hand them back
            const size_t pos = (size_t)(in - src) + over - (size_t)(bc >> 3);
            if (pos > n) return false;
            in = src + pos;
            over = 0;
            bb = 0; bc = 0;
            if ((size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, in, len);
            in += len; out += len;
            continue;
        }
        if (type == 3) return false;
        const uint32_t *lt, *dt;
        if (type == 1) {
            const Tables *f = fixed_tables();
            lt = f->lit; dt = f->dist;
        } else {
            const int hlit = (int)(bb & 31) + 257, hdist = (int)((bb >> 5) & 31) + 1, hclen = (int)((bb >> 10) & 15) + 4;
            HXZ_DROP(14);
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pl[19] = {0};
            HXZ_REFILL();                               // 19 x 3 = 57 bits: two refills
            for (int i = 0; i < hclen; ++i) {
                if (bc < 3) HXZ_REFILL();
                pl[order[i]] = (uint8_t)(bb & 7);
                HXZ_DROP(3);
            }
            uint32_t pt[1 << PTB];
            if (!build_table(pl, 19, K_PRE, PTB, pt, 1 << PTB)) return false;
            uint8_t lens[286 + 30 + 138] = {0};
            int i = 0;
            const int total = hlit + hdist;
            while (i < total) {
                HXZ_REFILL();
                const uint32_t e = pt[bb & ((1u << PTB) - 1)];
                if ((e & T_MASK) != T_LIT) return false;
                HXZ_DROP(e & 63);
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                int rep; uint8_t v = 0;
                if (sym == 16) {
                    if (i == 0) return false;
                    v = lens[i - 1]; rep = 3 + (int)(bb & 3); HXZ_DROP(2);
                } else if (sym == 17) { rep = 3 + (int)(bb & 7); HXZ_DROP(3); }
                else { rep = 11 + (int)(bb & 127); HXZ_DROP(7); }
                if (i + rep > total) return false;
                memset(lens + i, v, (size_t)rep);
                i += rep;
            }
            if (lens[256] == 0) return false;           // no end-of-block code
            if (!build_table(lens, hlit, K_LIT, LTB, dyn.lit, L_CAP)) return false;
            if (!build_table(lens + hlit, hdist, K_DIST, DTB, dyn.dist, D_CAP)) return false;
            lt = dyn.lit; dt = dyn.dist;
        }
        // ---- symbols of the block
        for (;;) {
            HXZ_REFILL();
            uint32_t e = lt[bb & ((1u << LTB) - 1)];
            if ((e & T_MASK) == T_SUB) {
                HXZ_DROP(LTB);
                e = lt[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 31)) - 1))];
            }
            // literals: up to four lookups (one or two bytes each) out of one refill: at most 15 + 3 x 11 bits
            int k = 0;
            while ((e & T_MASK) <= T_LIT2) {
                if (out_end - out >= 2) {                            // both bytes stored, the pointer moves by one or two
                    out[0] = (uint8_t)(e >> 16);
                    out[1] = (uint8_t)(e >> 24);
                    out += 1 + ((e >> 13) & 1);
                } else {
                    if (out >= out_end || (e & T_MASK) == T_LIT2) return false;
                    *out++ = (uint8_t)(e >> 16);
                }
                HXZ_DROP(e & 63);
                if (++k == 4) break;
                e = lt[bb & ((1u << LTB) - 1)];
            }
            if (k == 4 || (k && (e & T_MASK) == T_SUB)) continue;   // (a peeked entry has consumed nothing)
            if ((e & T_MASK) == T_EOB) { HXZ_DROP(e & 63); break; }
            if ((e & T_MASK) != T_LEN) return false;
            HXZ_DROP(e & 63);
            const uint32_t xl = (e >> 8) & 31;
            const size_t len = (size_t)(e >> 16) + (size_t)(bb & ((1u << xl) - 1));
            HXZ_DROP(xl);
            HXZ_REFILL();
            uint32_t d = dt[bb & ((1u << DTB) - 1)];
            if ((d & T_MASK) == T_SUB) {
                HXZ_DROP(DTB);
                d = dt[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 31)) - 1))];
            }
            if ((d & T_MASK) != T_LEN) return false;
            HXZ_DROP(d & 63);
            const uint32_t xd = (d >> 8) & 31;
            const size_t dist = (size_t)(d >> 16) + (size_t)(bb & ((1u << xd) - 1));
            HXZ_DROP(xd);
            if (dist > (size_t)(out - dst) || len > (size_t)(out_end - out)) return false;
            const uint8_t *s = out - dist;
            uint8_t *const stop = out + len;
            if ((size_t)(out_end - out) >= len + 16) {               // room to overshoot by whole words
                if (dist >= 8) {
                    do {
                        st64(out, ld64(s));
                        st64(out + 8, ld64(s + 8));
                        out += 16; s += 16;
                    } while (out < stop);
                } else if (dist == 1) {
                    const uint64_t v = 0x0101010101010101ull * (uint64_t)*s;
                    do { st64(out, v); out += 8; } while (out < stop);
                } else {
                    do { *out++ = *s++; } while (out < stop);
                }
                out = stop;
            } else {
                do { *out++ = *s++; } while (out < stop);
            }
        }
    }
#undef HXZ_REFILL
#undef HXZ_DROP
    // bytes fetched but not used stay in the bit buffer: the stream must not have needed more than n bytes
    const size_t consumed = (size_t)(in - src) + over - (size_t)(bc >> 3);
    return out == out_end && consumed <= n;
}

}  // namespace hxz
