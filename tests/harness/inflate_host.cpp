// tests/harness/inflate_host.cpp -- TEST HARNESS, not product code.
//
// Compiles gcn10_b200/csrc/inflate_core.h (the decode-lane half of the GPU tile inflater) for the host
// and drives it with a scalar emulation of the warp loop of inflate_tiles.cuh: same 2 KB input ring and
// refill rule, same batch order (all literals of a batch first, then the matches one after the other in
// 32- or 128-byte read-then-write steps against a 32 KB history ring, the batch cut behind every far-reaching match).  tests/test_inflate_core.py compares the
// result with zlib on the CPU, so the bit reader, the Huffman table builder, the symbol decoder and the
// batch-hazard rule are checked without a GPU.  Built by the test itself (g++ -shared); nothing under
// gcn10_b200/ links or loads it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../gcn10_b200/csrc/inflate_core.h"

using namespace gcn10::inflate;

extern "C" int gcn10_test_inflate(const uint8_t *stream, uint32_t size, uint32_t misalign, uint8_t *out,
                                  uint32_t out_len, uint64_t *stats /* [symbols, matches, steps, blocks] */)
{
    // the device reads 16-byte aligned chunks and may run up to ~2 KB past the stream: give it that room
    std::vector<uint8_t> padded((size_t)misalign + size + 4096 + 64, 0xA5);
    const uint32_t first = misalign & 15u;
    memcpy(padded.data() + first, stream, size);
    const uint8_t *base = padded.data();

    static Tables t;
    uint32_t ring[kRingWords];
    uint32_t queue[kQueue];
    std::vector<uint8_t> window(kWindow, 0);
    uint32_t filled = 0;
    auto top_up = [&](uint32_t cons) {
        while (filled < cons + 1024u) {
            for (uint32_t k = 0; k < 512; k++)
                ((uint8_t *)ring)[(filled + k) & 2047u] = base[filled + k];
            filled += 512u;
        }
    };
    auto emit = [&](uint32_t pos, uint8_t b) {
        window[pos & (kWindow - 1)] = b;
        if (pos < out_len)
            out[pos] = b;
    };
    // the flusher warp's checksum: (byte sum, position-weighted sum) of every piece it stores
    uint32_t s1 = 1, s2 = 0;
    auto sum_piece = [&](uint32_t from, uint32_t n) {
        uint64_t a = 0, b = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t d = window[(from + i) & (kWindow - 1)];
            a += d;
            b += (uint64_t)(n - i) * d;
        }
        adler_advance(s1, s2, n, a, b);
    };
    // the plane as the flusher warp fills it: pieces of kPiece bytes, at most two in flight -- a piece counts as
    // stored (readable by a far match) only once the writer has posted two younger ones, the guarantee the device has
    std::vector<uint8_t> plane((size_t)out_len + kWindow + 64, 0xCD);
    uint32_t posted = 0, plane_done = 0;
    auto flush_upto = [&](uint32_t upto) {
        while (posted + (uint32_t)kPiece <= upto) {
            for (uint32_t i = 0; i < (uint32_t)kPiece; i++)
                plane[posted + i] = window[(posted + i) & (kWindow - 1)];
            sum_piece(posted, (uint32_t)kPiece);
            posted += (uint32_t)kPiece;
            plane_done = posted >= 2u * (uint32_t)kPiece ? posted - 2u * (uint32_t)kPiece : 0u;
        }
    };
    DecodeLane s;
    lane_init(s, first, first + size);
    top_up(first & ~3u);
    s.err = read_zlib_header(s, ring, first);
    uint32_t out_base = 0;
    int werr = 0;                       // what the writer warp finds (range checks)
    uint64_t nsym = 0, nmatch = 0, nsteps = 0, nblocks = 0;
    for (;;) {
        top_up(s.cons);
        if (!s.err && werr)
            s.err = werr;
        int n = 0, ev = kEvMore, fin = 0;
        uint32_t so = 0, sl = 0;
        bool post = false;
        nsteps++;
        if (s.err) {
            ev = kEvError;
            post = true;
        }
        else if (!s.in_block) {
            const int action = read_block_header(s, ring, t);
            if (action != kHdrError)
                nblocks++;
            if (action == kHdrBuild) {
                // the device runs these two with 32 lanes; emulate the lanes one after the other
                for (int lane = 0; lane < 32; lane++)
                    clear_block_luts(t, lane, 32);
                for (int lane = 0; lane < 32; lane++)
                    fill_block_luts(t, lane, 32);
            }
            else if (action == kHdrStored) {
                so = s.stored_src;
                sl = s.stored_len;
                fin = s.bfinal;
                const uint32_t q = so + sl;
                filled = q & ~511u;
                top_up(q & ~3u);
                seek(s, ring, q);
                ev = kEvStored;
                post = true;
            }
            else if (action == kHdrError) {
                ev = kEvError;
                post = true;
            }
        }
        else {
            n = decode_symbols(s, ring, t, queue, &ev);
            post = true;
        }
        if (!post)
            continue;

        // ---- writer warp
        if (n > 0 && !werr) {
            uint32_t start[kQueue], end[kQueue], pos = out_base;
            bool bad = false;
            for (int k = 0; k < n; k++) {
                start[k] = pos;
                pos += sym_len(queue[k]);
                end[k] = pos;
                if (sym_is_match(queue[k]) && sym_dist(queue[k]) > start[k])
                    bad = true;
            }
            if (pos > out_len)
                werr = kErrOverflow;
            else if (bad)
                werr = kErrDistance;
            else {
                // as the writer warp does it (inflate_core.h): the batch runs in parts; per part all literals first,
                // then the matches that start below the ring (read from the plane), then the others in order
                const uint32_t nparts = (pos - out_base - 1u) / (uint32_t)kPart + 1u;
                for (uint32_t pid = 0; pid < nparts && !werr; pid++) {
                    int lo = -1, hi = -1;
                    for (int k = 0; k < n; k++)
                        if ((end[k] - out_base - 1u) / (uint32_t)kPart == pid) {
                            if (lo < 0)
                                lo = k;
                            hi = k;
                        }
                    if (lo < 0)
                        continue;
                    const uint32_t p1 = end[hi];
                    for (int k = lo; k <= hi; k++)
                        if (!sym_is_match(queue[k]))
                            emit(start[k], (uint8_t)(queue[k] & 255u));
                    for (int pass = 0; pass < 2 && !werr; pass++)
                        for (int k = lo; k <= hi && !werr; k++) {
                            if (!sym_is_match(queue[k]))
                                continue;
                            const uint32_t len = queue[k] & 0x1FFu, dist = sym_dist(queue[k]), mp = start[k];
                            const bool far = byte_from_plane(mp - dist, p1);
                            if (far != (pass == 0))
                                continue;
                            nmatch++;
                            if (far) {
                                std::vector<uint8_t> tmp(len);
                                for (uint32_t j = 0; j < len; j++) {
                                    const uint32_t sp = mp - dist + j;
                                    if (byte_from_plane(sp, p1)) {
                                        if (sp >= plane_done)
                                            werr = 99;              // would read a byte the flusher has not stored yet
                                        tmp[j] = plane[sp];
                                    }
                                    else
                                        tmp[j] = window[sp & (kWindow - 1)];
                                }
                                for (uint32_t j = 0; j < len; j++)
                                    emit(mp + j, tmp[j]);
                            }
                            else if (dist >= len) {
                                std::vector<uint8_t> tmp(len);
                                for (uint32_t j = 0; j < len; j++)
                                    tmp[j] = window[(mp - dist + j) & (kWindow - 1)];
                                for (uint32_t j = 0; j < len; j++)
                                    emit(mp + j, tmp[j]);
                            }
                            else if (dist >= 32u) {
                                const uint32_t stepw = dist >= 128u ? 128u : 32u;
                                for (uint32_t b = 0; b < len; b += stepw) {
                                    uint8_t tmp[128];
                                    const uint32_t m = len - b < stepw ? len - b : stepw;
                                    for (uint32_t j = 0; j < m; j++)
                                        tmp[j] = window[(mp - dist + b + j) & (kWindow - 1)];
                                    for (uint32_t j = 0; j < m; j++)
                                        emit(mp + b + j, tmp[j]);
                                }
                            }
                            else {
                                uint8_t pat[32];
                                for (uint32_t j = 0; j < dist; j++)
                                    pat[j] = window[(mp - dist + j) & (kWindow - 1)];
                                for (uint32_t i = 0; i < len; i++)
                                    emit(mp + i, pat[i % dist]);
                            }
                        }
                    flush_upto(p1);
                }
                nsym += (uint64_t)n;
                out_base = pos;
            }
        }
        if (ev == kEvStored && !werr) {
            if (out_base + sl > out_len)
                werr = kErrOverflow;
            else {
                for (uint32_t i = 0; i < sl; i++) {
                    emit(out_base + i, base[so + i]);
                    if (((out_base + i + 1) % (uint32_t)kPart) == 0 || i + 1 == sl)
                        flush_upto(out_base + i + 1);
                }
                out_base += sl;
            }
        }
        if (ev == kEvEnd || (ev == kEvStored && fin)) {
            if (!werr && out_base != out_len)
                werr = kErrShort;
            if (!werr) {
                // decoder lane: the trailer behind the final block; flusher: the sum over everything decoded
                if (posted < out_base)
                    sum_piece(posted, out_base - posted);         // the last, partial piece
                uint32_t want = 0;
                top_up(s.cons);
                if (!read_adler_trailer(s, ring, &want))
                    werr = kErrInput;
                else if (want != ((s2 << 16) | s1))
                    werr = kErrChecksum;
            }
            break;
        }
        if (ev == kEvError) {
            if (!werr)
                werr = s.err;
            break;
        }
    }
    if (stats) {
        stats[0] = nsym;
        stats[1] = nmatch;
        stats[2] = nsteps;
        stats[3] = nblocks;
    }
    return werr;
}
