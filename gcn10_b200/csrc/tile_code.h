// gcn10_b200/csrc/tile_code.h -- the Huffman code of the Curve Number tile streams (host side, plain C++).
//
// save_raster() of the reference writes DEFLATE tiles through zlib level 6 (/root/reference/src/raster.c:
// 206-207), which fits a Huffman code to every tile.  The GPU encoder cannot afford a code per tile, but it does
// not need the fixed code of RFC 1951 3.2.6 either: all tiles of a run have the same statistics, known in
// advance -- the only distances are 1 (run) and 256 (pixel above), the longest match (258: runs of rows that repeat the row
// above, coded across the tile rows) carries a third of the tokens, and the literals are the few dozen Curve Number values the
// lookup tables can produce.  build_tile_code() designs ONE code for them and serialises it as the header of a
// dynamic-Huffman block (RFC 1951 3.2.7); every tile stream starts with that header (about 60 bytes) and then
// spends 15 bits instead of 24 on a repeated row, 1 bit instead of 5 on a distance, 7 instead of 8 on a literal.
//
// All literals that can occur get the SAME code length L, so that a token has the same bit length in every
// plane and the fused kernel needs one bit position for all streams.  The code is complete by construction
// (zlib rejects incomplete literal/length sets): either the literals (padded with values that never occur) take
// 2^L - 2^(L-m) code words of length L and the length symbols + end-of-block hang as one optimal Huffman tree
// under the remaining m-bit prefix, or the literals are one leaf of that tree split into 2^k code words -- see
// build_tile_code().
#pragma once

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace gcn10 {

struct TileCode {
    int lit_bits;                       // L: code length of every literal that can occur
    uint32_t lit_first;                 // canonical first code word of length L (MSB-first value)
    uint8_t lit_rank[256];              // literal value -> index among the length-L literal symbols
    uint16_t len_code[29];              // length symbols 257..285: bit-reversed code word
    uint8_t len_bits[29];
    uint16_t eob_code;                  // symbol 256
    uint8_t eob_bits;
    uint16_t dist_code[2];              // distance symbol 0 (distance 1) and 15 (193..256): bit-reversed
    uint8_t dist_bits[2];
    int header_bits;                    // zlib header (16) + BFINAL/BTYPE (3) + dynamic block header
    uint32_t header_words[96];          // those bits, LSB first, ready to be OR-ed in at bit 0 of a stream
};

namespace tile_code_detail {

// optimal prefix code lengths by repeated merging (n is tiny); zero frequencies get length 0
inline void huffman_lengths(const std::vector<uint64_t> &freq, std::vector<int> &len)
{
    const int n = (int)freq.size();
    len.assign(n, 0);
    struct Node { uint64_t w; int left, right; };
    std::vector<Node> nodes;
    std::vector<int> live;
    for (int i = 0; i < n; i++)
        if (freq[i]) {
            nodes.push_back({ freq[i], -1 - i, -1 - i });
            live.push_back((int)nodes.size() - 1);
        }
    if (live.size() == 1) {
        len[-1 - nodes[live[0]].left] = 1;
        return;
    }
    while (live.size() > 1) {
        std::sort(live.begin(), live.end(), [&](int a, int b) { return nodes[a].w != nodes[b].w ? nodes[a].w > nodes[b].w : a > b; });
        const int a = live.back();
        live.pop_back();
        const int b = live.back();
        live.pop_back();
        nodes.push_back({ nodes[a].w + nodes[b].w, a, b });
        live.push_back((int)nodes.size() - 1);
    }
    // depth of every leaf
    std::vector<std::pair<int, int>> stack = { { live[0], 0 } };
    while (!stack.empty()) {
        const auto [id, d] = stack.back();
        stack.pop_back();
        if (nodes[id].left < 0 && nodes[id].left == nodes[id].right) {
            len[-1 - nodes[id].left] = d;
            continue;
        }
        stack.push_back({ nodes[id].left, d + 1 });
        stack.push_back({ nodes[id].right, d + 1 });
    }
}

// Huffman lengths limited to maxbits: flatten the frequencies until the tree is shallow enough
inline void limited_lengths(std::vector<uint64_t> freq, int maxbits, std::vector<int> &len)
{
    for (int round = 0; round < 64; round++) {
        huffman_lengths(freq, len);
        if (*std::max_element(len.begin(), len.end()) <= maxbits)
            return;
        uint64_t total = 0;
        for (uint64_t f : freq)
            total += f;
        const uint64_t floor_w = std::max<uint64_t>(1, total >> (maxbits - 1 - std::min(round, maxbits - 2)));
        for (uint64_t &f : freq)
            if (f)
                f = std::max(f, floor_w);
    }
}

inline uint32_t reverse_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++)
        r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// canonical code words (RFC 1951 3.2.2) from lengths
inline void canonical_codes(const std::vector<int> &len, std::vector<uint32_t> &code)
{
    int count[16] = { 0 };
    for (int l : len)
        count[l]++;
    count[0] = 0;
    uint32_t next[16] = { 0 };
    uint32_t c = 0;
    for (int b = 1; b < 16; b++) {
        c = (c + (uint32_t)count[b - 1]) << 1;
        next[b] = c;
    }
    code.assign(len.size(), 0);
    for (size_t s = 0; s < len.size(); s++)
        if (len[s])
            code[s] = next[len[s]]++;
}

struct BitWriter {
    uint32_t *w;
    int cap_bits;
    int pos = 0;
    bool overflow = false;
    void put(uint32_t v, int n)
    {
        if (pos + n > cap_bits) {
            overflow = true;
            return;
        }
        for (int i = 0; i < n; i++, pos++)
            if ((v >> i) & 1u)
                w[pos >> 5] |= 1u << (pos & 31);
    }
};

}  // namespace tile_code_detail

// present[v] = true for every byte value the planes can hold (the lookup-table values, nodata 255, padding 0).
// Returns false when no such code exists (more than 240 values): the caller then keeps the fixed code.
inline bool build_tile_code(const bool present[256], TileCode &tc)
{
    using namespace tile_code_detail;
    memset(&tc, 0, sizeof(tc));
    int n_lit = 0;
    for (int v = 0; v < 256; v++)
        n_lit += present[v] ? 1 : 0;
    if (n_lit == 0 || n_lit > 240)
        return false;

    // model of a Curve Number tile's tokens, per 1000 tokens: the length symbols and end-of-block (index = symbol - 256).
    // Measured, not guessed: tools/token_stats.py runs the kernel's parse (tests/harness/tile_code_host.cpp) over the
    // record-id tiles of the synthetic WorldCover-like and coastal blocks of BASELINE.json (10 m land-cover patches over
    // 25-pixel soil cells) and prints this table.  Two things carry it: the row-run tokens (258: the inside of a soil
    // cell row, a fifth of all tokens) and the lengths 11..226, which together are a third of the tokens -- a soil cell
    // is 25 pixels wide, a land-cover patch a few hundred.  The round-1 guess gave those 8 bits each and cost 4 % more
    // bytes per block than this table, which is within 0.1 % of a code fitted to each block's own statistics.
    static const uint64_t model[30] = { 2,                                  // end of block
                                        12, 10, 109, 21, 4, 7, 3, 7,        // lengths 3..10
                                        22, 9, 24, 16,                      // 11..18
                                        13, 21, 18, 15,                     // 19..34
                                        27, 22, 15, 19,                     // 35..66
                                        23, 18, 15, 11,                     // 67..130
                                        18, 12, 8,                          // 131..226
                                        79,                                 // 227..257: a single row that repeats the row above
                                        195 };                              // 258: runs of such rows, coded across the tile rows
    std::vector<uint64_t> w(model, model + 30);
    const uint64_t lit_weight = 230;                    // literals per 1000 tokens (same measurement)
    uint64_t nonlit_weight = 0;
    for (uint64_t f : w)
        nonlit_weight += f;
    std::vector<int> depth;
    limited_lengths(w, 11, depth);

    // Two shapes of literal/length code keep every literal equally long.  Shape A: the literals (L bits) fill all but
    // one m-bit prefix and the length symbols hang under that prefix -- right when literals dominate.  Shape B: the
    // literals are ONE leaf of the length symbols' Huffman tree, d bits deep, that is split into 2^k code words
    // (L = d + k) -- right for Curve Number tiles, where three tokens in four are matches: with the shipped tables
    // (64 values) that is 8-bit literals under a 2-bit prefix and three quarters of the code space for the length
    // symbols, 1.6-2.3 % fewer bytes than shape A's 7-bit literals (tools/token_stats.py).  The model picks.
    int best_l = 0, best_m = 0;
    double best_cost = 1e300;
    for (int l = 1; l <= 8; l++)
        for (int m = 1; m <= 4 && m <= l; m++) {
            const int cap = (1 << l) - (1 << (l - m));
            if (cap < n_lit || cap > 256)
                continue;
            double cost = (double)lit_weight * l;
            for (int i = 0; i < 30; i++)
                cost += (double)w[i] * (m + depth[i]);
            if (cost < best_cost) {
                best_cost = cost;
                best_l = l;
                best_m = m;
            }
        }
    (void)nonlit_weight;
    int k = 0;
    while ((1 << k) < n_lit)
        k++;
    std::vector<uint64_t> wj(w);
    wj.push_back(lit_weight);                           // the leaf that becomes the literals
    std::vector<int> depth_j;
    limited_lengths(wj, 11, depth_j);
    double cost_j = (double)lit_weight * (depth_j[30] + k);
    for (int i = 0; i < 30; i++)
        cost_j += (double)w[i] * depth_j[i];
    const bool shape_b = depth_j[30] + k <= 8 && cost_j < best_cost;   // (literals stay within 8 bits, as in shape A)
    if (!best_l && !shape_b)
        return false;
    const int L = shape_b ? depth_j[30] + k : best_l, m = shape_b ? 0 : best_m;
    const int cap = shape_b ? 1 << k : (1 << L) - (1 << (L - m));

    // literal/length code lengths: present literals, then never-used values as padding up to `cap`
    std::vector<int> ll(286, 0);
    int given = 0;
    for (int v = 0; v < 256; v++)
        if (present[v]) {
            ll[v] = L;
            given++;
        }
    for (int v = 255; v >= 0 && given < cap; v--)
        if (!present[v]) {
            ll[v] = L;
            given++;
        }
    if (given != cap)
        return false;
    for (int i = 0; i < 30; i++)
        ll[256 + i] = shape_b ? depth_j[i] : m + depth[i];
    // distance code: symbols 0 and 15, one bit each
    std::vector<int> dl(16, 0);
    dl[0] = 1;
    dl[15] = 1;

    std::vector<uint32_t> lcode, dcode;
    canonical_codes(ll, lcode);
    canonical_codes(dl, dcode);
    tc.lit_bits = L;
    bool first_set = false;
    int rank = 0;
    for (int v = 0; v < 256; v++)
        if (ll[v] == L) {
            if (!first_set) {
                tc.lit_first = lcode[v];
                first_set = true;
            }
            tc.lit_rank[v] = (uint8_t)rank++;
        }
    // (length symbols may share the length L; they follow the literals in symbol order, so literal code words are
    // lit_first + rank)
    for (int i = 0; i < 29; i++) {
        tc.len_bits[i] = (uint8_t)ll[257 + i];
        tc.len_code[i] = (uint16_t)reverse_bits(lcode[257 + i], ll[257 + i]);
    }
    tc.eob_bits = (uint8_t)ll[256];
    tc.eob_code = (uint16_t)reverse_bits(lcode[256], ll[256]);
    tc.dist_bits[0] = 1;
    tc.dist_code[0] = (uint16_t)reverse_bits(dcode[0], 1);
    tc.dist_bits[1] = 1;
    tc.dist_code[1] = (uint16_t)reverse_bits(dcode[15], 1);

    // ---- dynamic block header: the code lengths, run-length coded with the code-length alphabet
    std::vector<int> seq(ll.begin(), ll.end());
    seq.insert(seq.end(), dl.begin(), dl.end());
    struct Item { int sym, extra, nbits; };
    std::vector<Item> items;
    for (size_t i = 0; i < seq.size();) {
        size_t j = i;
        while (j < seq.size() && seq[j] == seq[i])
            j++;
        size_t run = j - i;
        if (seq[i] == 0) {
            while (run >= 11) {
                const size_t r = std::min<size_t>(run, 138);
                items.push_back({ 18, (int)r - 11, 7 });
                run -= r;
            }
            if (run >= 3) {
                items.push_back({ 17, (int)run - 3, 3 });
                run = 0;
            }
            while (run--)
                items.push_back({ 0, 0, 0 });
        }
        else {
            items.push_back({ seq[i], 0, 0 });
            run--;
            while (run >= 3) {
                const size_t r = std::min<size_t>(run, 6);
                items.push_back({ 16, (int)r - 3, 2 });
                run -= r;
            }
            while (run--)
                items.push_back({ seq[i], 0, 0 });
        }
        i = j;
    }
    std::vector<uint64_t> clf(19, 0);
    for (const Item &it : items)
        clf[it.sym]++;
    std::vector<int> cll;
    limited_lengths(clf, 7, cll);
    std::vector<uint32_t> clcode;
    canonical_codes(cll, clcode);
    static const int order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    int hclen = 19;
    while (hclen > 4 && cll[order[hclen - 1]] == 0)
        hclen--;

    BitWriter bw{ tc.header_words, (int)sizeof(tc.header_words) * 8 };
    bw.put(0x9C78u, 16);                // CMF = 0x78 (deflate, 32K window), FLG = 0x9C
    bw.put(1u, 1);                      // BFINAL
    bw.put(2u, 2);                      // BTYPE = 10: dynamic Huffman
    bw.put(286 - 257, 5);               // HLIT
    bw.put(16 - 1, 5);                  // HDIST
    bw.put((uint32_t)hclen - 4, 4);     // HCLEN
    for (int i = 0; i < hclen; i++)
        bw.put((uint32_t)cll[order[i]], 3);
    for (const Item &it : items) {
        bw.put(reverse_bits(clcode[it.sym], cll[it.sym]), cll[it.sym]);
        if (it.nbits)
            bw.put((uint32_t)it.extra, it.nbits);
    }
    if (bw.overflow)
        return false;
    tc.header_bits = bw.pos;
    return true;
}

}  // namespace gcn10
