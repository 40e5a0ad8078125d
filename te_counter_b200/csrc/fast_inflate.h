// Raw-DEFLATE (RFC 1951) decoder for BGZF blocks: whole input and whole output in memory, output
// size known up front (the block's ISIZE), at most 64 KiB.  Written for the BAM decoder
// (bamdecode.cpp), where inflate is > 90 % of the time per record:
//   * 64-bit bit buffer, refilled with one unaligned 8-byte load (at most one refill per symbol);
//   * one table lookup per literal/length symbol for codes up to 11 bits (a 16-entry second level
//     for longer ones), same for distances with 8 bits (128-entry second level);
//   * up to three literals decoded per refill; matches copied in 8-byte words when the distance
//     allows it;
//   * a careful tail (byte-wise refill and copies) for the last bytes of input and output, so
//     nothing is read outside [in, in + in_n) or written outside [out, out + out_n).
// It returns false on ANY anomaly -- malformed or truncated stream, output size not equal to
// out_n, a code set it does not handle (an incomplete literal/length code) -- and the caller then
// runs zlib on the same block, which either decodes it or reports the error.  zlib therefore
// stays the arbiter of what is valid; this decoder only has to be right when it says true
// (tests/test_fast_inflate.py compares it with zlib on generated and corrupted streams).
#pragma once
#include <cstdint>
#include <cstring>

namespace fast_inflate {

constexpr int LIT_BITS = 11, LIT_SUB_BITS = 4, DIST_BITS = 8, DIST_SUB_BITS = 7, PRE_BITS = 7;
constexpr uint32_t F_LIT = 0x8000, F_EOB = 0x4000, F_SUB = 0x2000;
constexpr uint32_t NO_SYMBOL = 0xFFFFFFFFu;             // in a value table: the symbol may be coded but never used
// entry: [7:0] bits to consume, [12:8] extra bits, [15:13] flags, [31:16] literal / base / subtable start.  0 = no such code.

struct Tables {
    uint32_t lit[(1 << LIT_BITS) + 288 * (1 << LIT_SUB_BITS)];
    uint32_t dist[(1 << DIST_BITS) + 32 * (1 << DIST_SUB_BITS)];
    uint32_t pre[1 << PRE_BITS];
};

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// Builds a decode table from code lengths.  `value[s]` = entry without the length field for symbol s.
// Returns false for an over-subscribed code, and for an incomplete one unless `allow_incomplete`
// (then unused codes stay 0 and decode as errors).
inline bool build(const uint8_t *lens, int n_syms, const uint32_t *value, int table_bits, int sub_bits, uint32_t *table,
                  int table_cap, bool allow_incomplete) {
    int count[16] = {0};
    for (int s = 0; s < n_syms; s++) count[lens[s]]++;
    count[0] = 0;
    uint32_t next_code[16], code = 0;
    int64_t kraft = 0;
    for (int l = 1; l <= 15; l++) {
        code = (code + uint32_t(count[l - 1])) << 1;
        next_code[l] = code;
        kraft += int64_t(count[l]) << (15 - l);
    }
    if (kraft > (1 << 15)) return false;
    if (kraft < (1 << 15) && !allow_incomplete) return false;
    const int primary = 1 << table_bits;
    memset(table, 0, sizeof(uint32_t) * size_t(primary));
    int used = primary;
    for (int s = 0; s < n_syms; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        if (value[s] == NO_SYMBOL) continue;            // its entries stay 0: decoding it is an error
        if (l <= table_bits) {
            const uint32_t e = value[s] | uint32_t(l);
            for (uint32_t i = rev; i < uint32_t(primary); i += 1u << l) table[i] = e;
        } else {
            if (l - table_bits > sub_bits) return false;
            const uint32_t lo = rev & uint32_t(primary - 1);
            if (!(table[lo] & F_SUB)) {
                if (used + (1 << sub_bits) > table_cap) return false;
                memset(table + used, 0, sizeof(uint32_t) << sub_bits);
                table[lo] = F_SUB | (uint32_t(used) << 16) | uint32_t(table_bits);
                used += 1 << sub_bits;
            }
            uint32_t *sub = table + (table[lo] >> 16);
            const uint32_t e = value[s] | uint32_t(l - table_bits);
            for (uint32_t i = rev >> table_bits; i < (1u << sub_bits); i += 1u << (l - table_bits)) sub[i] = e;
        }
    }
    return true;
}

struct Values {
    uint32_t lit[288], dist[32], pre[19];
    Values() {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                           4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (uint32_t s = 0; s < 256; s++) lit[s] = F_LIT | (s << 16);
        lit[256] = F_EOB;
        for (int s = 257; s < 286; s++) lit[s] = (uint32_t(lbase[s - 257]) << 16) | (uint32_t(lext[s - 257]) << 8);
        lit[286] = lit[287] = NO_SYMBOL;                // part of the fixed code, never valid in a stream
        for (int s = 0; s < 30; s++) dist[s] = (uint32_t(dbase[s]) << 16) | (uint32_t(dext[s]) << 8);
        dist[30] = dist[31] = NO_SYMBOL;
        for (uint32_t s = 0; s < 19; s++) pre[s] = s << 16;
    }
};

class Inflater {
public:
    // Decodes one complete raw-deflate stream.  true iff it ended cleanly with exactly out_n bytes.
    bool run(const uint8_t *in, size_t in_n, uint8_t *out, size_t out_n) {
        p_ = in;
        in_end_ = in + in_n;
        buf_ = 0;
        cnt_ = 0;
        uint8_t *o = out, *const out_end = out + out_n;
        for (;;) {
            if (!need(3)) return false;
            const uint32_t hdr = take(3);
            const bool final = hdr & 1;
            switch (hdr >> 1) {
            case 0: {                                   // stored
                take(cnt_ & 7);
                p_ -= cnt_ >> 3;                        // whole bytes still in the buffer go back
                buf_ = 0;
                cnt_ = 0;
                if (in_end_ - p_ < 4) return false;
                const uint32_t len = p_[0] | (uint32_t(p_[1]) << 8), nlen = p_[2] | (uint32_t(p_[3]) << 8);
                p_ += 4;
                if ((len ^ nlen) != 0xFFFF || size_t(in_end_ - p_) < len || size_t(out_end - o) < len) return false;
                memcpy(o, p_, len);
                o += len;
                p_ += len;
                break;
            }
            case 1:
                if (!fixed_tables()) return false;
                if (!codes(out, o, out_end)) return false;
                break;
            case 2:
                if (!dynamic_tables()) return false;
                if (!codes(out, o, out_end)) return false;
                break;
            default:
                return false;
            }
            if (final) return o == out_end;
        }
    }

private:
    static uint64_t load64(const uint8_t *p) {
        uint64_t v;
        memcpy(&v, p, 8);
        return v;                                       // little-endian hosts only (x86-64, aarch64)
    }
    // careful refill: never reads at or beyond in_end_.  The buffer holds exactly cnt_ valid bits.
    bool need(int n) {
        while (cnt_ < n) {
            if (p_ >= in_end_) return false;
            buf_ |= uint64_t(*p_++) << cnt_;
            cnt_ += 8;
        }
        return true;
    }
    uint32_t take(int n) {
        const uint32_t v = uint32_t(buf_ & ((uint64_t(1) << n) - 1));
        buf_ >>= n;
        cnt_ -= n;
        return v;
    }

    bool fixed_tables() {
        uint8_t lens[288 + 32];
        int s = 0;
        for (; s < 144; s++) lens[s] = 8;
        for (; s < 256; s++) lens[s] = 9;
        for (; s < 280; s++) lens[s] = 7;
        for (; s < 288; s++) lens[s] = 8;
        for (s = 0; s < 32; s++) lens[288 + s] = 5;
        return build(lens, 288, v_.lit, LIT_BITS, LIT_SUB_BITS, t_.lit, int(sizeof(t_.lit) / 4), false) &&
               build(lens + 288, 32, v_.dist, DIST_BITS, DIST_SUB_BITS, t_.dist, int(sizeof(t_.dist) / 4), false);
    }

    bool dynamic_tables() {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        if (!need(14)) return false;
        const int hlit = int(take(5)) + 257, hdist = int(take(5)) + 1, hclen = int(take(4)) + 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t pre_lens[19] = {0};
        for (int i = 0; i < hclen; i++) {
            if (!need(3)) return false;
            pre_lens[order[i]] = uint8_t(take(3));
        }
        if (!build(pre_lens, 19, v_.pre, PRE_BITS, 0, t_.pre, 1 << PRE_BITS, false)) return false;
        uint8_t lens[288 + 32];
        memset(lens, 0, sizeof(lens));
        int i = 0;
        const int total = hlit + hdist;
        while (i < total) {
            need(PRE_BITS + 7);                         // fewer bits near the end are fine: checked through e and cnt_
            const uint32_t e = t_.pre[buf_ & ((1u << PRE_BITS) - 1)];
            const int l = int(e & 0xFF);
            if (l == 0 || l > cnt_) return false;
            take(l);
            const int sym = int(e >> 16);
            if (sym < 16) {
                lens[i++] = uint8_t(sym);
                continue;
            }
            int rep, xbits, base;
            uint8_t v = 0;
            if (sym == 16) {
                if (!i) return false;
                v = lens[i - 1];
                xbits = 2;
                base = 3;
            } else if (sym == 17) {
                xbits = 3;
                base = 3;
            } else {
                xbits = 7;
                base = 11;
            }
            if (cnt_ < xbits) return false;
            rep = base + int(take(xbits));
            if (i + rep > total) return false;
            while (rep--) lens[i++] = v;
        }
        if (!lens[256]) return false;                   // no end-of-block code
        uint8_t dl[32] = {0};
        memcpy(dl, lens + hlit, size_t(hdist));
        memset(lens + hlit, 0, size_t(288 - hlit));
        // distance codes: complete, or the degenerate sets compressors emit for blocks with at most
        // one distance in use (one code of one bit, or none at all) -- what zlib accepts as well
        int n_used = 0, max_len = 0;
        for (int s = 0; s < 30; s++) {
            n_used += dl[s] != 0;
            if (dl[s] > max_len) max_len = dl[s];
        }
        return build(lens, 288, v_.lit, LIT_BITS, LIT_SUB_BITS, t_.lit, int(sizeof(t_.lit) / 4), false) &&
               build(dl, 32, v_.dist, DIST_BITS, DIST_SUB_BITS, t_.dist, int(sizeof(t_.dist) / 4), n_used == 0 || (n_used == 1 && max_len == 1));
    }

    // Decodes symbols up to and including the end-of-block code.  The loop is shift-and-mask bound:
    // on x86-64 a second copy is compiled for BMI2 (shrx / bzhi) and picked at run time.
    bool codes(uint8_t *const out, uint8_t *&o_ref, uint8_t *const out_end) {
#if defined(__x86_64__) && defined(__GNUC__)
        static const bool bmi2 = __builtin_cpu_supports("bmi2");
        if (bmi2) return codes_bmi2(out, o_ref, out_end);
#endif
        return codes_body(out, o_ref, out_end);
    }
#if defined(__x86_64__) && defined(__GNUC__)
    __attribute__((target("bmi,bmi2"))) bool codes_bmi2(uint8_t *const out, uint8_t *&o_ref, uint8_t *const out_end) {
        return codes_body(out, o_ref, out_end);
    }
#endif
    __attribute__((always_inline)) inline bool codes_body(uint8_t *const out, uint8_t *&o_ref, uint8_t *const out_end) {
        uint8_t *o = o_ref;
        const uint32_t *const lit = t_.lit, *const dist = t_.dist;
        const uint8_t *p = p_;
        uint64_t buf = buf_;
        int cnt = cnt_;
        // ---- fast loop: 8 readable input bytes, room for the longest match plus one overshooting word
        while (in_end_ - p >= 16 && out_end - o >= 258 + 3 + 8) {
            buf |= load64(p) << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
            uint32_t e = lit[buf & ((1u << LIT_BITS) - 1)];
            if (e & F_LIT) {                            // up to three literals on one refill (3 x 15 bits <= 56)
                buf >>= e & 0xFF; cnt -= int(e & 0xFF);
                *o++ = uint8_t(e >> 16);
                e = lit[buf & ((1u << LIT_BITS) - 1)];
                if (e & F_LIT) {
                    buf >>= e & 0xFF; cnt -= int(e & 0xFF);
                    *o++ = uint8_t(e >> 16);
                    e = lit[buf & ((1u << LIT_BITS) - 1)];
                    if (e & F_LIT) {
                        buf >>= e & 0xFF; cnt -= int(e & 0xFF);
                        *o++ = uint8_t(e >> 16);
                        continue;
                    }
                }
                buf |= load64(p) << cnt;                // a non-literal follows: top the buffer up again
                p += (63 - cnt) >> 3;
                cnt |= 56;
            }
            if (e & F_SUB) {
                buf >>= LIT_BITS; cnt -= LIT_BITS;
                e = lit[(e >> 16) + (buf & ((1u << LIT_SUB_BITS) - 1))];
            }
            if (!e) return false;
            buf >>= e & 0xFF; cnt -= int(e & 0xFF);
            if (e & F_LIT) {
                *o++ = uint8_t(e >> 16);
                continue;
            }
            if (e & F_EOB) {
                sync(p, buf, cnt);
                o_ref = o;
                return true;
            }
            int x = int(e >> 8) & 31;
            const uint32_t len = (e >> 16) + uint32_t(buf & ((uint64_t(1) << x) - 1));
            buf >>= x; cnt -= x;                        // <= 15 + 5 bits gone since the refill, >= 36 left; distance needs <= 28
            uint32_t d = dist[buf & ((1u << DIST_BITS) - 1)];
            if (d & F_SUB) {
                buf >>= DIST_BITS; cnt -= DIST_BITS;
                d = dist[(d >> 16) + (buf & ((1u << DIST_SUB_BITS) - 1))];
            }
            if (!d) return false;
            buf >>= d & 0xFF; cnt -= int(d & 0xFF);
            x = int(d >> 8) & 31;
            const uint32_t off = (d >> 16) + uint32_t(buf & ((uint64_t(1) << x) - 1));
            buf >>= x; cnt -= x;
            if (off > uint32_t(o - out)) return false;
            const uint8_t *src = o - off;
            uint8_t *const stop = o + len;
            if (off >= 8) {
                do {
                    memcpy(o, src, 8);
                    o += 8; src += 8;
                } while (o < stop);
            } else if (off == 1) {
                memset(o, *src, len);
            } else {
                do { *o++ = *src++; } while (o < stop);
            }
            o = stop;
        }
        // ---- careful tail
        sync(p, buf, cnt);
        for (;;) {
            need(15 + 5);                               // as many as the input still has
            uint32_t e = lit[buf_ & ((1u << LIT_BITS) - 1)];
            int used = 0;
            if (e & F_SUB) {
                used = LIT_BITS;
                e = lit[(e >> 16) + ((buf_ >> LIT_BITS) & ((1u << LIT_SUB_BITS) - 1))];
            }
            if (!e) return false;
            used += int(e & 0xFF);
            if (used > cnt_) return false;
            take(used);
            if (e & F_LIT) {
                if (o >= out_end) return false;
                *o++ = uint8_t(e >> 16);
                continue;
            }
            if (e & F_EOB) {
                o_ref = o;
                return true;
            }
            int x = int(e >> 8) & 31;
            if (!need(x)) return false;
            const uint32_t len = (e >> 16) + take(x);
            need(15 + 13);
            uint32_t d = dist[buf_ & ((1u << DIST_BITS) - 1)];
            used = 0;
            if (d & F_SUB) {
                used = DIST_BITS;
                d = dist[(d >> 16) + ((buf_ >> DIST_BITS) & ((1u << DIST_SUB_BITS) - 1))];
            }
            if (!d) return false;
            used += int(d & 0xFF);
            if (used > cnt_) return false;
            take(used);
            x = int(d >> 8) & 31;
            if (!need(x)) return false;
            const uint32_t off = (d >> 16) + take(x);
            if (off > uint32_t(o - out) || len > uint32_t(out_end - o)) return false;
            const uint8_t *src = o - off;
            for (uint32_t k = 0; k < len; k++) o[k] = src[k];
            o += len;
        }
    }

    // The fast loop counts the bits of whole loaded words (the buffer may hold more valid bits than
    // cnt says); hand the careful code an exact state: drop whole unread bytes back to the input.
    void sync(const uint8_t *p, uint64_t buf, int cnt) {
        p -= cnt >> 3;
        cnt &= 7;
        p_ = p;
        buf_ = buf & ((uint64_t(1) << cnt) - 1);
        cnt_ = cnt;
    }

    Tables t_;
    Values v_;
    const uint8_t *p_ = nullptr, *in_end_ = nullptr;
    uint64_t buf_ = 0;
    int cnt_ = 0;
};

}  // namespace fast_inflate
