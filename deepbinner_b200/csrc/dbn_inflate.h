// Fast DEFLATE (RFC 1951) / zlib (RFC 1950) decompressor for the fast5 reader.
//
// Why: after the HDF5 structures are parsed, ~85 % of the time of reading a fast5 file is spent inflating
// the raw signal (one deflate-level-1 chunk of noisy int16 samples - literal-heavy Huffman data that
// zlib's inflate decodes at 150-250 MB/s per core; this decoder is ~11 % faster on the GPU boxes' Xeons
// and stops as soon as the wanted prefix of a read has been produced).  This decoder follows the well-known recipe of
// word-at-a-time decoders: a 64-bit bit buffer refilled with one unaligned load, two-level decode
// tables whose entries carry symbol, base value, extra-bit count and code length in one 32-bit word,
// several literals decoded per refill, and word-wise match copies.  Written from the RFC; checked
// against zlib on every fixture and on randomised streams (tests/test_fast5_readers.py).
//
// Contract: `src` must be readable for 8 bytes beyond `src_len` (the reader pads its file buffer);
// output is truncated at `dst_len` (HDF5 edge chunks are declared larger than the data set).  Returns
// false on a malformed stream (the caller then lets zlib give the verdict).
#pragma once

#include <cstdint>
#include <cstring>

namespace dbn_inflate {

constexpr int kLitlenBits = 11;     // primary table index width for literal/length codes
constexpr int kOffsetBits = 8;      // primary table index width for distance codes
constexpr int kPrecodeBits = 7;
constexpr int kLitlenSize = (1 << kLitlenBits) + 1024;   // + room for all second-level tables
constexpr int kOffsetSize = (1 << kOffsetBits) + 512;

// entry: bits 0-7 = bits to consume (code length, or for a subtable pointer the primary width),
//        bits 8-9 = kind, bits 10-14 = extra bits (or subtable width), bits 16-31 = payload
enum Kind : uint32_t { kLiteral = 0, kBase = 1, kEndOfBlock = 2, kSubtable = 3 };
inline uint32_t make_entry(uint32_t kind, uint32_t payload, uint32_t extra, uint32_t len) {
    return len | (kind << 8) | (extra << 10) | (payload << 16);
}

struct Tables {
    uint32_t litlen[kLitlenSize];
    uint32_t offset[kOffsetSize];
    uint32_t precode[1 << kPrecodeBits];
};

static const uint16_t kLengthBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                         35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLengthExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                         3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kOffsetBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                         257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kOffsetExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                         7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// Builds a two-level decode table for a canonical Huffman code.  `symbol_entry(sym)` gives the entry of
// a symbol without its length field.  Returns false for an over-subscribed code; an incomplete code is
// accepted only in the forms deflate allows (a single code of length 1 / no codes at all for distances).
template <typename EntryOf>
inline bool build_table(const uint8_t* lens, int nsyms, int table_bits, int max_len, uint32_t* table, int table_size,
                        EntryOf symbol_entry) {
    int count[16] = {0};
    for (int s = 0; s < nsyms; ++s) ++count[lens[s]];
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= max_len; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;                        // over-subscribed
    }
    // codes of the same length are consecutive, in symbol order
    uint16_t offs[17];
    offs[1] = 0;
    for (int l = 1; l < 16; ++l) offs[l + 1] = static_cast<uint16_t>(offs[l] + count[l]);
    uint16_t sorted[320];
    for (int s = 0; s < nsyms; ++s)
        if (lens[s]) sorted[offs[lens[s]]++] = static_cast<uint16_t>(s);
    const int primary = 1 << table_bits;
    // an incomplete code: fill with an entry that makes the decoder fail (kind kSubtable with width 0 and
    // payload 0xFFFF is never produced otherwise)
    const uint32_t invalid = make_entry(kEndOfBlock, 0xFFFF, 31, 1);
    for (int i = 0; i < primary; ++i) table[i] = invalid;
    uint32_t code = 0;            // canonical code, MSB first
    int idx = 0;                  // index into sorted
    int next_sub = primary;       // next free second-level slot
    uint32_t cur_prefix = ~0u;    // primary index of the subtable being filled
    int cur_sub_bits = 0, cur_sub_start = 0;
    for (int len = 1; len <= max_len; ++len) {
        for (int k = 0; k < count[len]; ++k, ++idx, ++code) {
            const int sym = sorted[idx];
            // bit-reverse the code: deflate packs Huffman codes starting from the most significant bit
            uint32_t rev = 0;
            for (int b = 0; b < len; ++b) rev |= ((code >> b) & 1u) << (len - 1 - b);
            if (len <= table_bits) {
                const uint32_t e = symbol_entry(sym) | static_cast<uint32_t>(len);
                for (uint32_t i = rev; i < static_cast<uint32_t>(primary); i += 1u << len) table[i] = e;
            } else {
                const uint32_t prefix = rev & (primary - 1);
                if (prefix != cur_prefix) {
                    // new subtable: wide enough for the longest code sharing this prefix; codes are visited
                    // in increasing length, so size it from the remaining space of the code tree
                    int sub_bits = len - table_bits;
                    int space = (1 << sub_bits) - (count[len] - k);
                    for (int l2 = len; space > 0 && l2 < max_len;) {
                        ++l2;
                        ++sub_bits;
                        space = (space << 1) - count[l2];
                    }
                    cur_prefix = prefix;
                    cur_sub_bits = sub_bits;
                    cur_sub_start = next_sub;
                    next_sub += 1 << sub_bits;
                    if (next_sub > table_size) return false;
                    for (int i = cur_sub_start; i < next_sub; ++i) table[i] = invalid;
                    table[prefix] = make_entry(kSubtable, static_cast<uint32_t>(cur_sub_start), static_cast<uint32_t>(sub_bits),
                                               static_cast<uint32_t>(table_bits));
                }
                const uint32_t e = symbol_entry(sym) | static_cast<uint32_t>(len - table_bits);
                const uint32_t step = 1u << (len - table_bits);
                for (uint32_t i = rev >> table_bits; i < (1u << cur_sub_bits); i += step) table[cur_sub_start + i] = e;
            }
        }
        code <<= 1;
    }
    return true;
}

inline uint64_t load64(const uint8_t* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;   // little-endian hosts only (x86-64 / aarch64)
}

struct Bits {
    uint64_t buf = 0;
    int cnt = 0;
    const uint8_t* in;
    const uint8_t* in_end;   // refills never move `in` beyond in_end + 8 (padding contract)
    inline void refill() {
        buf |= load64(in) << cnt;
        in += (63 - cnt) >> 3;
        cnt |= 56;
    }
    inline uint32_t peek(int n) const { return static_cast<uint32_t>(buf) & ((1u << n) - 1); }
    inline void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t take(int n) {
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    // bytes of real input consumed so far must not exceed the stream: checked at block boundaries
    inline bool overrun() const { return in - ((cnt + 7) >> 3) > in_end; }
};

inline uint32_t litlen_entry(int sym) {
    if (sym < 256) return make_entry(kLiteral, static_cast<uint32_t>(sym), 0, 0);
    if (sym == 256) return make_entry(kEndOfBlock, 0, 0, 0);
    if (sym > 285) return make_entry(kEndOfBlock, 0xFFFF, 31, 0);     // invalid symbol
    return make_entry(kBase, kLengthBase[sym - 257], kLengthExtra[sym - 257], 0);
}
inline uint32_t offset_entry(int sym) {
    if (sym > 29) return make_entry(kEndOfBlock, 0xFFFF, 31, 0);
    return make_entry(kBase, kOffsetBase[sym], kOffsetExtra[sym], 0);
}
inline uint32_t precode_entry(int sym) { return make_entry(kLiteral, static_cast<uint32_t>(sym), 0, 0); }

// Raw deflate stream -> dst (at most dst_len bytes; the rest of the stream is still parsed for validity
// only as far as needed to stop cleanly).  *produced receives the bytes written.
inline bool inflate_raw(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len, size_t* produced) {
    static thread_local Tables T;
    Bits b;
    b.in = src;
    b.in_end = src + src_len;
    uint8_t* out = dst;
    uint8_t* const out_end = dst + dst_len;
    bool last = false;
    while (!last) {
        b.refill();
        last = b.take(1) != 0;
        const uint32_t type = b.take(2);
        if (type == 0) {                                  // stored block
            b.drop(b.cnt & 7);                             // to a byte boundary
            // un-read the whole bytes still in the bit buffer
            b.in -= b.cnt >> 3;
            b.buf = 0;
            b.cnt = 0;
            if (b.in + 4 > b.in_end) return false;
            const uint32_t len = b.in[0] | (b.in[1] << 8), nlen = b.in[2] | (b.in[3] << 8);
            if ((len ^ 0xFFFF) != nlen) return false;
            b.in += 4;
            if (b.in + len > b.in_end) return false;
            const size_t n = len < static_cast<size_t>(out_end - out) ? len : static_cast<size_t>(out_end - out);
            std::memcpy(out, b.in, n);
            out += n;
            b.in += len;
            if (out == out_end && n < len) break;          // output full: truncate
            continue;
        }
        if (type == 3) return false;
        uint8_t lens[288 + 32];
        int nlit, ndist;
        if (type == 1) {                                  // fixed Huffman codes
            nlit = 288;
            ndist = 32;
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
        } else {                                          // dynamic Huffman codes
            nlit = static_cast<int>(b.take(5)) + 257;
            ndist = static_cast<int>(b.take(5)) + 1;
            const int nprec = static_cast<int>(b.take(4)) + 4;
            if (nlit > 286 || ndist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t plens[19] = {0};
            b.refill();
            for (int i = 0; i < nprec; ++i) {
                if (b.cnt < 3) b.refill();
                plens[order[i]] = static_cast<uint8_t>(b.take(3));
            }
            if (!build_table(plens, 19, kPrecodeBits, 7, T.precode, 1 << kPrecodeBits, precode_entry)) return false;
            int i = 0;
            while (i < nlit + ndist) {
                b.refill();
                const uint32_t e = T.precode[b.peek(kPrecodeBits)];
                if (((e >> 8) & 3) != kLiteral) return false;
                b.drop(static_cast<int>(e & 0xFF));
                const uint32_t sym = e >> 16;
                if (sym < 16) {
                    lens[i++] = static_cast<uint8_t>(sym);
                    continue;
                }
                int rep;
                uint8_t val = 0;
                if (sym == 16) {
                    if (i == 0) return false;
                    val = lens[i - 1];
                    rep = 3 + static_cast<int>(b.take(2));
                } else if (sym == 17) {
                    rep = 3 + static_cast<int>(b.take(3));
                } else {
                    rep = 11 + static_cast<int>(b.take(7));
                }
                if (i + rep > nlit + ndist) return false;
                while (rep--) lens[i++] = val;
            }
            if (lens[256] == 0) return false;
            if (nlit < 288) std::memmove(lens + 288, lens + nlit, static_cast<size_t>(ndist));
            for (int k = nlit; k < 288; ++k) lens[k] = 0;
        }
        if (!build_table(lens, type == 1 ? 288 : nlit, kLitlenBits, 15, T.litlen, kLitlenSize, litlen_entry)) return false;
        if (!build_table(lens + 288, ndist, kOffsetBits, 15, T.offset, kOffsetSize, offset_entry)) return false;
        if (b.overrun()) return false;

        // ---- the decode loop ----
        bool full = false;
        for (;;) {
            b.refill();                                    // >= 56 bits: enough for a whole length/distance pair
            uint32_t e = T.litlen[b.peek(kLitlenBits)];
            if (((e >> 8) & 3) == kSubtable) {
                b.drop(static_cast<int>(e & 0xFF));
                e = T.litlen[(e >> 16) + b.peek(static_cast<int>((e >> 10) & 31))];
            }
            // up to three literals per refill (15 bits each at most; 56 are available).  (Entries that decode
            // two literals at once - low byte + high byte of a sample - were measured: no gain, 226 vs 219 us
            // per fixture file; the table has to be rebuilt for every ~16 KB block.)
            int budget = 2;
            while (((e >> 8) & 3) == kLiteral) {
                if (out == out_end) {
                    full = true;
                    break;
                }
                b.drop(static_cast<int>(e & 0xFF));
                *out++ = static_cast<uint8_t>(e >> 16);
                if (budget-- == 0) break;
                e = T.litlen[b.peek(kLitlenBits)];
                if (((e >> 8) & 3) == kSubtable) {
                    b.drop(static_cast<int>(e & 0xFF));
                    e = T.litlen[(e >> 16) + b.peek(static_cast<int>((e >> 10) & 31))];
                }
            }
            if (full) break;
            const uint32_t kind = (e >> 8) & 3;
            if (kind == kLiteral) {
                if (b.in > b.in_end + 8) return false;
                continue;                                  // literal budget used up: refill and go on
            }
            if (kind != kBase) {
                if ((e >> 16) == 0xFFFF) return false;      // invalid code
                b.drop(static_cast<int>(e & 0xFF));         // end of block
                break;
            }
            if (b.cnt < 48) b.refill();                     // (after three literals fewer than 48 bits may be left)
            b.drop(static_cast<int>(e & 0xFF));
            uint32_t length = (e >> 16) + b.take(static_cast<int>((e >> 10) & 31));
            uint32_t o = T.offset[b.peek(kOffsetBits)];
            if (((o >> 8) & 3) == kSubtable) {
                b.drop(static_cast<int>(o & 0xFF));
                o = T.offset[(o >> 16) + b.peek(static_cast<int>((o >> 10) & 31))];
            }
            if (((o >> 8) & 3) != kBase) return false;
            b.drop(static_cast<int>(o & 0xFF));
            const uint32_t dist = (o >> 16) + b.take(static_cast<int>((o >> 10) & 31));
            if (dist > static_cast<size_t>(out - dst)) return false;
            if (length > static_cast<size_t>(out_end - out)) {
                length = static_cast<uint32_t>(out_end - out);
                full = true;
            }
            const uint8_t* from = out - dist;
            if (dist >= 8 && static_cast<size_t>(out_end - out) >= length + 8) {
                uint8_t* const stop = out + length;        // word copies may overshoot by up to 7 bytes (room checked)
                do {
                    std::memcpy(out, from, 8);
                    out += 8;
                    from += 8;
                } while (out < stop);
                out = stop;
            } else {
                for (uint32_t k = 0; k < length; ++k) out[k] = from[k];
                out += length;
            }
            if (full) break;
            if (b.in > b.in_end + 8) return false;
        }
        if (full) break;
        if (b.overrun()) return false;
    }
    *produced = static_cast<size_t>(out - dst);
    return true;
}

// zlib container: 2-byte header, deflate stream, Adler-32 (not verified: HDF5 chunks carry their own
// optional fletcher32 filter, and a damaged stream fails Huffman decoding in practice).
inline bool inflate_zlib(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len, size_t* produced) {
    if (src_len < 6) return false;
    const uint32_t cmf = src[0], flg = src[1];
    if ((cmf & 0x0F) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return false;
    return inflate_raw(src + 2, src_len - 6, dst, dst_len, produced);
}

}  // namespace dbn_inflate
