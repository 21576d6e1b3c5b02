// gcn10_b200/csrc/inflate_core.h -- the sequential half of the GPU tile inflater.
//
// The reference reads its land-cover window through GDAL (load_raster, /root/reference/src/raster.c:
// 106-189; GDALRasterIO at :177-179); for the ESA WorldCover GeoTIFFs that means zlib-inflating
// TIFF Compression=8 tiles on the CPU (1024 x 1024 tiles, landcover/esa_worldcover_2021.vrt).  The
// inflater here decodes those zlib streams (RFC 1950 container, RFC 1951 DEFLATE: stored, fixed and
// dynamic Huffman blocks) on the device, one warp per tile, so that only compressed bytes cross PCIe.
//
// A DEFLATE stream is a serial bit stream, so one lane of the warp ("the decode lane") owns the bit
// reader, parses block headers, builds the Huffman lookup tables and turns code words into a queue of
// up to 32 LZ77 symbols; the whole warp then executes the queue (inflate_tiles.cuh).  Everything the
// decode lane runs lives in this header as plain C++ with no CUDA intrinsics, so that the very same
// code is compiled for the host by tests/harness/inflate_host.cpp and checked against zlib on the CPU.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GCN10_HD __host__ __device__ __forceinline__
#else
#define GCN10_HD inline
#endif

namespace gcn10 {
namespace inflate {

// History.  DEFLATE reaches back 32768 bytes (RFC 1951), but 96 % of the matches of a land-cover tile stay within
// 4 KB (run, pixel above, a few rows up).  The ring in shared memory therefore holds only the last kWindow bytes;
// older history is read back from the tile's place in the destination plane, where the flusher warp has already put
// it (L2 resident).  With an 8 KB ring a tile needs 18 KB of shared memory instead of 43 KB and eleven tiles fit an SM:
// all 1296 tiles of a 36000 x 36000 block decode at once instead of in 1.75 waves.
constexpr int kWindow = 8192;           // shared-memory history ring (power of two)
constexpr int kPart = 1024;             // a batch is executed in parts: symbols whose last byte falls into the same
                                        // kPart-aligned span of the batch's output (a part is < kPart + 258 bytes)
constexpr int kPiece = 2048;            // the ring is flushed to the plane in pieces of this size, two in flight
constexpr int kLlBits = 10;             // literal/length lookup width; longer codes take the canonical walk
constexpr int kDBits = 9;               // distance lookup width
constexpr int kRingWords = 512;         // compressed-input ring: 2 KB, refilled 512 B at a time by the warp
constexpr int kQueue = 32;              // symbols per batch (one per lane)
constexpr int kMaxBatchOut = kQueue * 258;

// lookup entry: bits 0-3 code length (0 = not a direct hit), 4-7 extra bits, 8-9 kind, 16-31 base value
enum { kLit = 0, kLen = 1, kEob = 2, kSlow = 3 };

enum {
    kErrNone = 0,
    kErrZlibHeader = 1,     // CMF/FLG check failed, preset dictionary, method != 8
    kErrBlockType = 2,      // BTYPE 3
    kErrStoredLen = 3,      // LEN != ~NLEN
    kErrCodeLengths = 4,    // over-subscribed code, bad repeat, too many lengths, no end-of-block code
    kErrBadCode = 5,        // bit pattern that is not a code word / symbols 286, 287, distance 30, 31
    kErrDistance = 6,       // distance reaches before the start of the tile
    kErrOverflow = 7,       // stream holds more bytes than the tile
    kErrInput = 8,          // ran past the end of the compressed tile
    kErrShort = 9,          // stream ended before the tile was full
    kErrChecksum = 10       // the Adler-32 trailer does not match the decoded bytes (zlib: Z_DATA_ERROR, which GDAL
                            // turns into the read error of raster.c:182-186)
};

// event reported by one step of the decode lane
enum { kEvMore = 0, kEvStored = 1, kEvEnd = 2, kEvError = 3 };

struct Tables {
    uint32_t ll_lut[1 << kLlBits];
    uint32_t d_lut[1 << kDBits];
    uint16_t sorted[288 + 32];      // symbols ordered by (code length, symbol): literal/length, then distance
    uint16_t codes[288 + 32];       // bit-reversed code word of every symbol
    uint16_t ll_count[16], d_count[16];
    uint8_t lens[288 + 32];
};

struct DecodeLane {
    uint64_t buf;           // bit buffer, LSB first
    int cnt;                // valid bits in buf
    uint32_t cons;          // byte offset (from the 16-byte aligned stream base) of the next ring word to load
    uint32_t in_end;        // byte offset one past the compressed tile
    int in_block;           // inside a Huffman block
    int bfinal;
    int fixed_ready;        // tables currently hold the fixed code
    int err;
    uint32_t stored_src;    // kEvStored: byte offset of the raw bytes, and how many
    uint32_t stored_len;
};

GCN10_HD uint32_t bit_reverse(uint32_t v, int n)
{
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++)
        r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// cnt >= 32 afterwards
GCN10_HD void need32(DecodeLane &s, const uint32_t *ring)
{
    if (s.cnt <= 32) {
        const uint32_t w = ring[(s.cons >> 2) & (kRingWords - 1)];
        s.buf |= (uint64_t)w << s.cnt;
        s.cnt += 32;
        s.cons += 4;
    }
}

GCN10_HD uint32_t get_bits(DecodeLane &s, const uint32_t *ring, int n)       // n <= 16
{
    need32(s, ring);
    const uint32_t v = (uint32_t)s.buf & ((1u << n) - 1u);
    s.buf >>= n;
    s.cnt -= n;
    return v;
}

// byte offset of the next unread bit (rounded down)
GCN10_HD uint32_t byte_pos(const DecodeLane &s) { return s.cons - (uint32_t)(s.cnt >> 3); }

// restart the bit reader at byte offset q (ring must hold the word containing q)
GCN10_HD void seek(DecodeLane &s, const uint32_t *ring, uint32_t q)
{
    s.cons = q & ~3u;
    s.buf = 0;
    s.cnt = 0;
    need32(s, ring);
    const int skip = 8 * (int)(q & 3u);
    s.buf >>= skip;
    s.cnt -= skip;
}

GCN10_HD uint32_t ll_entry(int sym, int nbits)
{
    if (sym < 256)
        return (uint32_t)nbits | (kLit << 8) | ((uint32_t)sym << 16);
    if (sym == 256)
        return (uint32_t)nbits | (kEob << 8);
    const int c = sym - 257;            // 0..28 (286, 287 never get here)
    int eb = 0, base;
    if (c < 8)
        base = 3 + c;
    else if (c == 28)
        base = 258;
    else {
        eb = (c >> 2) - 1;
        base = 3 + ((4 + (c & 3)) << eb);
    }
    return (uint32_t)nbits | ((uint32_t)eb << 4) | (kLen << 8) | ((uint32_t)base << 16);
}

GCN10_HD uint32_t d_entry(int sym, int nbits)
{
    int eb = 0, base;
    if (sym < 4)
        base = 1 + sym;
    else {
        eb = (sym >> 1) - 1;
        base = 1 + ((2 + (sym & 1)) << eb);
    }
    return (uint32_t)nbits | ((uint32_t)eb << 4) | ((uint32_t)base << 16);
}

// Canonical Huffman code assignment from code lengths (RFC 1951 3.2.2), sequential part: per-length
// counts, the symbols ordered by (length, symbol) for the canonical walk, and every symbol's code word
// (bit-reversed, ready to index a lookup table).  Returns 0, or kErrCodeLengths for an over-subscribed
// set.  Incomplete sets are accepted (unused bit patterns decode to kErrBadCode).
GCN10_HD int assign_codes(const uint8_t *lens, int nsyms, uint16_t *sorted, uint16_t *count, uint16_t *codes)
{
    uint16_t offs[16], next[16];
    for (int l = 0; l < 16; l++)
        count[l] = 0;
    for (int s = 0; s < nsyms; s++)
        count[lens[s]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0)
            return kErrCodeLengths;
    }
    offs[1] = 0;
    next[0] = 0;
    uint32_t code = 0;
    for (int l = 1; l < 16; l++) {
        if (l > 1)
            offs[l] = (uint16_t)(offs[l - 1] + count[l - 1]);
        code = (code + count[l - 1]) << 1;
        next[l] = (uint16_t)code;
    }
    for (int s = 0; s < nsyms; s++) {
        const int l = lens[s];
        if (!l)
            continue;
        sorted[offs[l]++] = (uint16_t)s;
        codes[s] = (uint16_t)bit_reverse(next[l]++, l);
    }
    return 0;
}

// Lookup-table fill, written so that `nlanes` workers can share it: worker `lane` clears and fills a
// strided part.  The caller separates clear_lut and fill_lut by a barrier (a warp barrier on the device).
GCN10_HD void clear_lut(uint32_t *lut, int tbits, int lane, int nlanes)
{
    const uint32_t slow = (uint32_t)(kSlow << 8);
    for (int i = lane; i < (1 << tbits); i += nlanes)
        lut[i] = slow;
}

GCN10_HD void fill_lut(const uint8_t *lens, const uint16_t *codes, int nsyms, int max_valid, bool dist, uint32_t *lut,
                       int tbits, int lane, int nlanes)
{
    for (int s = lane; s < nsyms && s < max_valid; s += nlanes) {
        const int l = lens[s];
        if (!l || l > tbits)
            continue;
        const uint32_t e = dist ? d_entry(s, l) : ll_entry(s, l);
        for (uint32_t k = codes[s]; k < (1u << tbits); k += 1u << l)
            lut[k] = e;
    }
}

// Bit-by-bit canonical decode for code words longer than the lookup width (or invalid patterns).
// Returns the symbol and the number of bits it used, or -1.
GCN10_HD int slow_symbol(uint64_t buf, const uint16_t *sorted, const uint16_t *count, int *used)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)(buf & 1u);
        buf >>= 1;
        const int c = count[l];
        if (code - c < first) {
            *used = l;
            return sorted[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

GCN10_HD void fixed_lengths(uint8_t *lens)
{
    for (int s = 0; s < 144; s++) lens[s] = 8;
    for (int s = 144; s < 256; s++) lens[s] = 9;
    for (int s = 256; s < 280; s++) lens[s] = 7;
    for (int s = 280; s < 288; s++) lens[s] = 8;
    for (int s = 0; s < 32; s++) lens[288 + s] = 5;
}

// Dynamic block header (RFC 1951 3.2.7): code-length code, then HLIT + HDIST code lengths.
GCN10_HD int read_dynamic_header(DecodeLane &s, const uint32_t *ring, Tables &t)
{
    const int hlit = 257 + (int)get_bits(s, ring, 5);
    const int hdist = 1 + (int)get_bits(s, ring, 5);
    const int hclen = 4 + (int)get_bits(s, ring, 4);
    if (hlit > 286 || hdist > 30)
        return kErrCodeLengths;
    const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    uint8_t cl[19];
    for (int i = 0; i < 19; i++)
        cl[i] = 0;
    for (int i = 0; i < hclen; i++)
        cl[order[i]] = (uint8_t)get_bits(s, ring, 3);
    // the code-length code is decoded through the (not yet needed) distance lookup storage
    uint16_t cl_count[16];
    uint32_t *cl_lut = t.d_lut;
    {
        // 7-bit lookup: entry = nbits | sym << 16
        uint16_t next[16];
        for (int l = 0; l < 16; l++)
            cl_count[l] = 0;
        for (int i = 0; i < 19; i++)
            cl_count[cl[i]]++;
        cl_count[0] = 0;
        int left = 1;
        for (int l = 1; l < 8; l++) {
            left <<= 1;
            left -= cl_count[l];
            if (left < 0)
                return kErrCodeLengths;
        }
        uint32_t code = 0;
        for (int l = 1; l < 8; l++) {
            code = (code + cl_count[l - 1]) << 1;
            next[l] = (uint16_t)code;
        }
        for (int i = 0; i < 128; i++)
            cl_lut[i] = 0;
        for (int i = 0; i < 19; i++) {
            const int l = cl[i];
            if (!l)
                continue;
            const uint32_t c = next[l]++;
            for (uint32_t k = bit_reverse(c, l); k < 128u; k += 1u << l)
                cl_lut[k] = (uint32_t)l | ((uint32_t)i << 16);
        }
    }
    const int total = hlit + hdist;
    int n = 0, prev = 0;
    uint8_t *lens = t.lens;
    while (n < total) {
        need32(s, ring);
        const uint32_t e = cl_lut[(uint32_t)s.buf & 127u];
        const int l = (int)(e & 15u);
        if (!l)
            return kErrCodeLengths;
        s.buf >>= l;
        s.cnt -= l;
        const int sym = (int)(e >> 16);
        // lengths are stored literal/length at [0, hlit), distance at [288, 288 + hdist)
        if (sym < 16) {
            lens[n < hlit ? n : 288 + (n - hlit)] = (uint8_t)sym;
            prev = sym;
            n++;
        }
        else {
            int rep, val = 0;
            if (sym == 16) {
                if (n == 0)
                    return kErrCodeLengths;
                val = prev;
                rep = 3 + (int)get_bits(s, ring, 2);
            }
            else if (sym == 17)
                rep = 3 + (int)get_bits(s, ring, 3);
            else
                rep = 11 + (int)get_bits(s, ring, 7);
            if (n + rep > total)
                return kErrCodeLengths;
            for (int k = 0; k < rep; k++, n++)
                lens[n < hlit ? n : 288 + (n - hlit)] = (uint8_t)val;
            if (sym != 16)
                prev = 0;
        }
    }
    for (int i = hlit; i < 288; i++)
        lens[i] = 0;
    for (int i = hdist; i < 32; i++)
        lens[288 + i] = 0;
    if (lens[256] == 0)
        return kErrCodeLengths;
    return 0;
}

// code assignment for both alphabets of a block (sequential; the lookup fill follows, see fill_block_luts)
GCN10_HD int assign_block_codes(Tables &t)
{
    const int rc = assign_codes(t.lens, 288, t.sorted, t.ll_count, t.codes);
    if (rc)
        return rc;
    return assign_codes(t.lens + 288, 32, t.sorted + 288, t.d_count, t.codes + 288);
}

GCN10_HD void clear_block_luts(Tables &t, int lane, int nlanes)
{
    clear_lut(t.ll_lut, kLlBits, lane, nlanes);
    clear_lut(t.d_lut, kDBits, lane, nlanes);
}

GCN10_HD void fill_block_luts(Tables &t, int lane, int nlanes)
{
    fill_lut(t.lens, t.codes, 288, 286, false, t.ll_lut, kLlBits, lane, nlanes);
    fill_lut(t.lens + 288, t.codes + 288, 32, 30, true, t.d_lut, kDBits, lane, nlanes);
}

GCN10_HD void lane_init(DecodeLane &s, uint32_t first_byte, uint32_t in_end)
{
    s.buf = 0;
    s.cnt = 0;
    s.cons = first_byte & ~3u;
    s.in_end = in_end;
    s.in_block = 0;
    s.bfinal = 0;
    s.fixed_ready = 0;
    s.err = 0;
    s.stored_src = 0;
    s.stored_len = 0;
}

// zlib container header (RFC 1950 2.2)
GCN10_HD int read_zlib_header(DecodeLane &s, const uint32_t *ring, uint32_t first_byte)
{
    seek(s, ring, first_byte);
    const uint32_t cmf = get_bits(s, ring, 8), flg = get_bits(s, ring, 8);
    if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 0x20u))
        return kErrZlibHeader;
    return 0;
}

// Adler-32 trailer (RFC 1950 2.2): the four bytes behind the final block, most significant first.  The
// reader is byte-aligned first.  Returns false when the stream ends before the trailer does.
GCN10_HD bool read_adler_trailer(DecodeLane &s, const uint32_t *ring, uint32_t *adler)
{
    const int drop = s.cnt & 7;
    s.buf >>= drop;
    s.cnt -= drop;
    if (byte_pos(s) + 4u > s.in_end)
        return false;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++)
        v = (v << 8) | get_bits(s, ring, 8);
    *adler = v;
    return true;
}

// Adler-32 of a stream continued over a piece of n bytes whose byte sum is a and whose position-weighted sum
// sum((n - i) * d[i]) is b:  s1' = s1 + a, s2' = s2 + n * s1 + b  (mod 65521)
GCN10_HD void adler_advance(uint32_t &s1, uint32_t &s2, uint32_t n, uint64_t a, uint64_t b)
{
    s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)(n % 65521u) * s1 + b % 65521u) % 65521u);
    s1 = (uint32_t)(((uint64_t)s1 + a) % 65521u);
}

// what read_block_header() asks its caller to do
enum { kHdrError = 0, kHdrStored = 1, kHdrReady = 2, kHdrBuild = 3 };

// Block header (RFC 1951 3.2.3).  kHdrStored: s.stored_src / s.stored_len describe the raw bytes, the caller
// copies them and re-seeks the reader behind them.  kHdrBuild: t.lens holds the block's code lengths and
// t.codes / t.sorted / counts are assigned; the caller runs clear_block_luts + fill_block_luts.  kHdrReady:
// the tables already hold this (fixed) code.
GCN10_HD int read_block_header(DecodeLane &s, const uint32_t *ring, Tables &t)
{
    if (byte_pos(s) > s.in_end + 8u) {
        s.err = kErrInput;
        return kHdrError;
    }
    s.bfinal = (int)get_bits(s, ring, 1);
    const int btype = (int)get_bits(s, ring, 2);
    if (btype == 0) {
        const int drop = s.cnt & 7;
        s.buf >>= drop;
        s.cnt -= drop;
        const uint32_t len = get_bits(s, ring, 16), nlen = get_bits(s, ring, 16);
        if ((len ^ nlen) != 0xFFFFu) {
            s.err = kErrStoredLen;
            return kHdrError;
        }
        s.stored_src = byte_pos(s);
        s.stored_len = len;
        if (s.stored_src + len > s.in_end) {
            s.err = kErrInput;
            return kHdrError;
        }
        return kHdrStored;
    }
    if (btype == 3) {
        s.err = kErrBlockType;
        return kHdrError;
    }
    s.in_block = 1;
    if (btype == 1) {
        if (s.fixed_ready)
            return kHdrReady;
        fixed_lengths(t.lens);
        s.fixed_ready = 1;
    }
    else {
        s.fixed_ready = 0;
        s.err = read_dynamic_header(s, ring, t);
        if (s.err)
            return kHdrError;
    }
    s.err = assign_block_codes(t);
    return s.err ? kHdrError : kHdrBuild;
}

// Up to kQueue symbols of the current Huffman block.
//   queue entry: literal = byte; match = 1<<31 | (dist-1) << 16 | len
// The decode lane does not know output positions: range checks (distance before the start of the tile,
// more bytes than the tile holds) are made by whoever executes the queue (check_batch below / the writer
// warp).  Returns the number of symbols queued; *event = kEvMore, kEvEnd (final block closed) or kEvError.
GCN10_HD int decode_symbols(DecodeLane &s, const uint32_t *ring, const Tables &t, uint32_t *queue, int *event)
{
    *event = kEvMore;
    if (byte_pos(s) > s.in_end + 8u) {
        s.err = kErrInput;
        *event = kEvError;
        return 0;
    }
    int n = 0;
    while (n < kQueue) {
        need32(s, ring);                                    // 33..64 valid bits
        const uint32_t lo = (uint32_t)s.buf;
        uint32_t e = t.ll_lut[lo & ((1u << kLlBits) - 1u)];
        // ---- fast paths: code words inside the lookup tables, all bits of the symbol inside the buffer.  One
        // branch for a literal; a match is decoded from the low word with a single 64-bit shift at the end.
        if ((e & 0x300u) == 0u) {                           // kLit
            const int nb = (int)(e & 15u);
            s.buf >>= nb;
            s.cnt -= nb;
            queue[n++] = e >> 16;
            continue;
        }
        if ((e & 0x300u) == (uint32_t)(kLen << 8)) {
            const uint32_t nb = e & 15u, eb = (e >> 4) & 15u;           // nb <= 10, eb <= 5
            uint32_t w = lo >> nb;
            const uint32_t len = (e >> 16) + (w & ~(~0u << eb));
            w >>= eb;                                                    // >= 17 valid bits left in w
            const uint32_t d = t.d_lut[w & ((1u << kDBits) - 1u)];
            const uint32_t dn = d & 15u, deb = (d >> 4) & 15u;
            const uint32_t used = nb + eb + dn, need = used + deb;      // <= 24, <= 37
            if (dn != 0u && (int)need <= s.cnt) {
                const uint32_t dist = (d >> 16) + ((uint32_t)(s.buf >> used) & ~(~0u << deb));
                s.buf >>= need;
                s.cnt -= (int)need;
                queue[n++] = 0x80000000u | ((dist - 1u) << 16) | len;
                continue;
            }
        }
        // ---- general path: long code words (canonical walk), end of block, a symbol that straddles the refill
        if ((e & 15u) == 0u) {
            int used = 0;
            const int sym = slow_symbol(s.buf, t.sorted, t.ll_count, &used);
            if (sym < 0 || sym > 285) {
                s.err = kErrBadCode;
                break;
            }
            e = ll_entry(sym, used);
        }
        const int nb = (int)(e & 15u);
        s.buf >>= nb;
        s.cnt -= nb;
        const uint32_t kind = (e >> 8) & 3u;
        if (kind == kLit) {
            queue[n++] = e >> 16;
            continue;
        }
        if (kind == kEob) {
            s.in_block = 0;
            if (s.bfinal)
                *event = kEvEnd;
            break;
        }
        const int eb = (int)((e >> 4) & 15u);
        const uint32_t len = (e >> 16) + ((uint32_t)s.buf & ~(~0u << eb));
        s.buf >>= eb;
        s.cnt -= eb;
        need32(s, ring);
        uint32_t d = t.d_lut[(uint32_t)s.buf & ((1u << kDBits) - 1u)];
        if ((d & 15u) == 0u) {
            int used = 0;
            const int sym = slow_symbol(s.buf, t.sorted + 288, t.d_count, &used);
            if (sym < 0 || sym > 29) {
                s.err = kErrBadCode;
                break;
            }
            d = d_entry(sym, used);
        }
        const int dn = (int)(d & 15u), deb = (int)((d >> 4) & 15u);
        s.buf >>= dn;
        s.cnt -= dn;
        const uint32_t dist = (d >> 16) + ((uint32_t)s.buf & ~(~0u << deb));
        s.buf >>= deb;
        s.cnt -= deb;
        queue[n++] = 0x80000000u | ((dist - 1u) << 16) | len;
    }
    if (s.err) {
        *event = kEvError;
        return 0;
    }
    return n;
}

GCN10_HD uint32_t sym_is_match(uint32_t sym) { return sym >> 31; }
GCN10_HD uint32_t sym_len(uint32_t sym) { return (sym >> 31) ? (sym & 0x1FFu) : 1u; }
GCN10_HD uint32_t sym_dist(uint32_t sym) { return ((sym >> 16) & 0x7FFFu) + 1u; }

// Execution rule of a part that covers output [p0, p1) (p1 - p0 < kPart + 258), shared by the writer warp and
// the CPU harness.  When the part starts, the ring holds [p1 - kWindow, p0) and the plane holds every byte below
// flushed_done, with flushed_done >= p0 - (3 * kPiece - 1) >= p1 - kWindow (pieces are handed over after every part,
// at most two in flight).  So a source byte at position s comes from the ring when s >= p1 - kWindow and from the
// plane otherwise.  Order inside a part: (1) all literals, (2) the matches that start below p1 - kWindow ("far":
// their source ends below p0, so they depend on nothing in the part), (3) the other matches, in stream order.
GCN10_HD bool byte_from_plane(uint32_t s, uint32_t p1) { return p1 > (uint32_t)kWindow && s < p1 - (uint32_t)kWindow; }

}  // namespace inflate
}  // namespace gcn10
