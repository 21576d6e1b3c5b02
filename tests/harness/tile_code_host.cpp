// tests/harness/tile_code_host.cpp -- TEST HARNESS, not product code.
//
// Compiles gcn10_b200/csrc/tile_code.h (the host-side design of the tuned Huffman code that the fused Curve
// Number + DEFLATE kernel writes its tiles with) and a scalar reference encoder that applies the kernel's token
// rules (per row: match at distance 256 = pixel above, match at distance 1 = run, literal) with that code.
// tests/test_tile_code.py inflates the streams with zlib: the header must parse, the code must be complete, and
// the bytes must come back.  Built by the test itself (g++ -shared); nothing under gcn10_b200/ links it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../gcn10_b200/csrc/tile_code.h"

using namespace gcn10;

namespace {

struct Out {
    std::vector<uint8_t> bytes;
    uint64_t pos = 0;
    void put(uint32_t v, int n)
    {
        for (int i = 0; i < n; i++, pos++) {
            if ((pos >> 3) >= bytes.size())
                bytes.push_back(0);
            if ((v >> i) & 1u)
                bytes[pos >> 3] |= (uint8_t)(1u << (pos & 7));
        }
    }
};

void put_match(Out &o, const TileCode &tc, int len, bool above)
{
    int idx, e = 0, extra = 0;
    if (len == 258)
        idx = 28;
    else {
        const int l = len - 3;
        if (l < 8)
            idx = l;
        else {
            e = 29 - __builtin_clz((unsigned)l);
            idx = 4 + 4 * e + ((l - (4 << e)) >> e);
            extra = (l - (4 << e)) & ((1 << e) - 1);
        }
    }
    o.put(tc.len_code[idx], tc.len_bits[idx]);
    o.put((uint32_t)extra, e);
    if (above) {
        o.put(tc.dist_code[1], tc.dist_bits[1]);
        o.put(63u, 6);                  // 193 + 63 = 256
    }
    else
        o.put(tc.dist_code[0], tc.dist_bits[0]);
}

}  // namespace

extern "C" int gcn10_test_tile_code(const uint8_t *present, int *lit_bits, int *header_bits, int *len284_bits, int *eob_bits)
{
    bool p[256];
    for (int v = 0; v < 256; v++)
        p[v] = present[v] != 0;
    TileCode tc;
    if (!build_tile_code(p, tc))
        return 1;
    *lit_bits = tc.lit_bits;
    *header_bits = tc.header_bits;
    *len284_bits = tc.len_bits[27];
    *eob_bits = tc.eob_bits;
    return 0;
}

// tile: 256 x 256 bytes whose values are all `present`; returns the stream length (0 = no code / overflow)
extern "C" uint32_t gcn10_test_tile_encode(const uint8_t *present, const uint8_t *tile, uint8_t *out, uint32_t cap)
{
    bool p[256];
    for (int v = 0; v < 256; v++)
        p[v] = present[v] != 0;
    TileCode tc;
    if (!build_tile_code(p, tc))
        return 0;
    Out o;
    for (int i = 0; i < tc.header_bits; i++)
        o.put((tc.header_words[i >> 5] >> (i & 31)) & 1u, 1);
    uint32_t s1 = 1, s2 = 0;
    for (int r = 0; r < 256; r++) {
        const uint8_t *row = tile + 256 * r;
        int x = 0;
        while (x < 256) {
            int la = 0, lr = 0;
            if (r > 0)
                while (x + la < 256 && row[x + la] == row[x + la - 256])
                    la++;
            if (x > 0)
                while (x + lr < 256 && row[x + lr] == row[x - 1])
                    lr++;
            const int len = la >= lr ? la : lr;
            if (len >= 3) {
                put_match(o, tc, len, la >= lr);
                x += len;
            }
            else {
                const uint32_t code = tc.lit_first + tc.lit_rank[row[x]];
                o.put(tile_code_detail::reverse_bits(code, tc.lit_bits), tc.lit_bits);
                x += 1;
            }
        }
    }
    o.put(tc.eob_code, tc.eob_bits);
    for (int i = 0; i < 65536; i++) {
        s1 = (s1 + tile[i]) % 65521u;
        s2 = (s2 + s1) % 65521u;
    }
    std::vector<uint8_t> &b = o.bytes;
    b.push_back((uint8_t)(s2 >> 8));
    b.push_back((uint8_t)s2);
    b.push_back((uint8_t)(s1 >> 8));
    b.push_back((uint8_t)s1);
    if (b.size() > cap)
        return 0;
    memcpy(out, b.data(), b.size());
    return (uint32_t)b.size();
}

// Token statistics of the fused kernel's parse of one 256 x 256 tile of record ids (cn_deflate_fused.cuh: rows that
// repeat the row above, two or more in a row, become matches of length 258 that run across the tile rows; every other
// row is parsed greedily -- pixel above, run, literal).  hist: [0] literals, [1] end-of-block, [2..30] length symbols
// 257..285, [31] extra length bits, [32] matches at distance 256, [33] at distance 1.  tools/token_stats.py turns these
// into the model table of build_tile_code().
extern "C" void gcn10_test_tile_tokens(const uint8_t *tile, uint64_t *hist)
{
    auto add_match = [&](int len, bool above) {
        int idx, e = 0;
        if (len == 258)
            idx = 28;
        else {
            const int l = len - 3;
            if (l < 8)
                idx = l;
            else {
                e = 29 - __builtin_clz((unsigned)l);
                idx = 4 + 4 * e + ((l - (4 << e)) >> e);
            }
        }
        hist[2 + idx]++;
        hist[31] += (uint64_t)e;
        hist[above ? 32 : 33]++;
    };
    int r = 0;
    while (r < 256) {
        const uint8_t *row = tile + 256 * r;
        if (r > 0 && memcmp(row, row - 256, 256) == 0) {
            int n = 1;
            while (r + n < 256 && memcmp(row + 256 * n, row + 256 * (n - 1), 256) == 0)
                n++;
            if (n >= 2) {
                const unsigned span = 256u * (unsigned)n;
                unsigned q = span / 258u;
                const unsigned rest = span - 258u * q;
                int tail[2] = { (int)rest, 0 };
                if (rest == 1u || rest == 2u) {
                    q--;
                    tail[0] = 129;
                    tail[1] = 129 + (int)rest;
                }
                for (unsigned i = 0; i < q; i++)
                    add_match(258, true);
                for (int t = 0; t < 2; t++)
                    if (tail[t])
                        add_match(tail[t], true);
                r += n;
                continue;
            }
        }
        int x = 0;
        while (x < 256) {
            int la = 0, lr = 0;
            if (r > 0)
                while (x + la < 256 && row[x + la] == row[x + la - 256])
                    la++;
            if (x > 0)
                while (x + lr < 256 && row[x + lr] == row[x - 1])
                    lr++;
            const int len = la >= lr ? la : lr;
            if (len >= 3) {
                add_match(len, la >= lr);
                x += len;
            }
            else {
                hist[0]++;
                x++;
            }
        }
        r++;
    }
    hist[1]++;
}

// the code build_tile_code() designs for `present`: bits of the 29 length symbols, end-of-block, literals, header
extern "C" int gcn10_test_tile_code_lengths(const uint8_t *present, int *len_bits /*[29]*/, int *eob_bits, int *lit_bits,
                                            int *header_bits)
{
    bool p[256];
    for (int v = 0; v < 256; v++)
        p[v] = present[v] != 0;
    TileCode tc;
    if (!build_tile_code(p, tc))
        return 1;
    for (int i = 0; i < 29; i++)
        len_bits[i] = tc.len_bits[i];
    *eob_bits = tc.eob_bits;
    *lit_bits = tc.lit_bits;
    *header_bits = tc.header_bits;
    return 0;
}

// What-if model (tools/token_stats.py, DESIGN.md "Next"): the same two distances, but the greedy parse runs over the
// tile as ONE stream of 65536 bytes, so a match may run on from the end of a row into the next one (the kernel's
// matches stop at row ends; only whole repeated rows are chained).  Same histogram layout as gcn10_test_tile_tokens.
extern "C" void gcn10_test_tile_tokens_stream(const uint8_t *t, uint64_t *hist, int run_on_tie)
{
    const int n = 65536;
    int p = 0;
    while (p < n) {
        int la = 0, lr = 0;
        if (p >= 256)
            while (p + la < n && la < 258 && t[p + la] == t[p + la - 256])
                la++;
        if (p >= 1)
            while (p + lr < n && lr < 258 && t[p + lr] == t[p + lr - 1])
                lr++;
        const bool above = run_on_tie ? la > lr : la >= lr;
        const int len = above ? la : lr;
        if (len >= 3) {
            int idx, e = 0;
            if (len == 258)
                idx = 28;
            else {
                const int l = len - 3;
                if (l < 8)
                    idx = l;
                else {
                    e = 29 - __builtin_clz((unsigned)l);
                    idx = 4 + 4 * e + ((l - (4 << e)) >> e);
                }
            }
            hist[2 + idx]++;
            hist[31] += (uint64_t)e;
            hist[above ? 32 : 33]++;
            p += len;
        }
        else {
            hist[0]++;
            p++;
        }
    }
    hist[1]++;
}
