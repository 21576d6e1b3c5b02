// gcn10_b200/csrc/deflate_tiles.cuh -- GPU-side DEFLATE of 256 x 256 GeoTIFF tiles (sm_100a).
//
// The reference ends process_block() by handing each plane to save_raster(), which writes a tiled
// DEFLATE GeoTIFF through GDAL/zlib on the CPU (/root/reference/src/raster.c:192-227, options at
// :206-207).  Once the Curve Number kernel runs at HBM speed that encode -- and the PCIe transfer of
// nine to eighteen raw 1.3 GB planes in front of it -- is >99 % of a block's wall time (SURVEY.md
// 7.4-3, 8f-1).  This kernel compresses every 256 x 256 tile of every output plane on the GPU into a
// complete zlib stream (RFC 1950 header, one RFC 1951 fixed-Huffman block, Adler-32), so that only the
// compressed tiles cross PCIe and the host merely lays them into the TIFF files.
//
// CN rasters are piecewise constant (25 x 25 pixel soil cells modulated by land-cover patches), so the
// matcher only looks at two distances: 256 (the pixel above, i.e. the previous tile row) and 1 (run of
// the previous pixel).  One thread parses one tile row greedily; a block-wide scan of the per-row bit
// counts gives every row its position in the bit stream; a second parse writes the bits.  Tiles that
// would not shrink are emitted as stored blocks.  Any inflate implementation decodes the result;
// decoded tiles are compared bit-for-bit with the raw planes in tests/test_gpu_deflate.py.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gcn10 {

constexpr int kTile = 256;                  // GDAL's default block size for TILED=YES
constexpr int kTileStride = 272;            // smem row stride: 17 x 16 B -> one-row-per-thread LDS.128 is conflict free
constexpr int kTileBytes = kTile * kTile;
constexpr int kEncCap = 40 * 1024;          // compressed tiles larger than this are emitted as stored blocks
constexpr int kStoredBytes = 2 + 2 * 5 + kTileBytes + 4;     // zlib header + two stored blocks + Adler-32
constexpr int kEncSmem = kTile * kTileStride + kEncCap + 64;

struct TileEncParams {
    const uint8_t *plane[18];   // device planes (first row of this strip)
    size_t pitch;
    int w, rows;                // valid pixels per row / rows in the strip
    int tiles_x, tile_rows;
    uint8_t *blob;              // output arena for the strip
    unsigned long long *cursor; // bump allocator (bytes, kept 16-byte aligned)
    unsigned long long *offsets;   // [plane][tile_row][tile_x]
    uint32_t *sizes;               // [plane][tile_row][tile_x]
};

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// fixed-Huffman literal: (reversed code, bit count)
__device__ __forceinline__ void lit_code(uint32_t v, uint32_t &bits, int &n)
{
    if (v < 144) { bits = rev_bits(0x30 + v, 8); n = 8; }
    else         { bits = rev_bits(0x190 + (v - 144), 9); n = 9; }
}

// match of length len (3..258) at distance 1 or 256: all bits of the token, LSB first
__device__ __forceinline__ void match_code(int len, bool above, uint32_t &bits, int &n)
{
    int code, e = 0, extra = 0;
    if (len == 258) {
        code = 285;
    }
    else {
        int l = len - 3;
        if (l < 8) {
            code = 257 + l;
        }
        else {
            e = 29 - __clz(l);
            code = 261 + 4 * e + ((l - (4 << e)) >> e);
            extra = (l - (4 << e)) & ((1 << e) - 1);
        }
    }
    int nb;
    uint32_t cb;
    if (code <= 279) { cb = rev_bits(code - 256, 7); nb = 7; }
    else             { cb = rev_bits(0xC0 + (code - 280), 8); nb = 8; }
    uint32_t v = cb | ((uint32_t)extra << nb);
    nb += e;
    if (above) {
        // distance 256: code 15 (193..256), 6 extra bits = 63
        v |= rev_bits(15, 5) << nb;
        v |= 63u << (nb + 5);
        nb += 11;
    }
    else {
        // distance 1: code 0, no extra bits
        nb += 5;
    }
    bits = v;
    n = nb;
}

// Per-row word masks, built once and reused by both parses:
//   above bit j : word j of the row equals word j of the row above
//   left  bit j : word j consists of four copies of the last byte of word j-1
struct RowMasks {
    unsigned long long above, left;
};

__device__ __forceinline__ uint32_t bcast_byte(uint32_t v) { return (v & 255u) * 0x01010101u; }

// number of consecutive set bits of m starting at bit j (j < 64)
__device__ __forceinline__ int ones_from(unsigned long long m, int j)
{
    const unsigned long long inv = ~(m >> j);       // bits shifted in from the top are 0 -> 1 after the inversion
    return __ffsll((long long)inv) - 1;             // inv != 0 whenever j > 0; for j == 0 a full mask gives ffs = 0
}

// length of the match starting at column x where the row equals `pat` word by word:
// pat_word(j) is the reference word for word j (row above, or four copies of the run byte)
template <bool ABOVE>
__device__ __forceinline__ int run_len(const uint8_t *row, unsigned long long mask, uint32_t runv, int x)
{
    const uint32_t *rw = reinterpret_cast<const uint32_t *>(row);
    const uint32_t *uw = reinterpret_cast<const uint32_t *>(row - kTileStride);
    int j = x >> 2, len = 0;
    const int o = x & 3;
    if (o) {
        const uint32_t ref = ABOVE ? uw[j] : runv;
        const uint32_t d = (rw[j] ^ ref) >> (8 * o);
        if (d)
            return (__ffs(d) - 1) >> 3;
        len = 4 - o;
        j++;
        if (j == kTile / 4)
            return len;
    }
    int nw = (j == 0 && mask == ~0ull) ? 64 : ones_from(mask, j);
    if (nw > kTile / 4 - j)
        nw = kTile / 4 - j;
    len += 4 * nw;
    j += nw;
    if (j < kTile / 4) {
        const uint32_t ref = ABOVE ? uw[j] : runv;
        const uint32_t d = rw[j] ^ ref;             // non-zero: the mask bit is clear
        len += d ? (__ffs(d) - 1) >> 3 : 4;
    }
    return len;
}

__device__ __forceinline__ void put_bits(uint32_t *out, unsigned long long pos, uint32_t v, int n)
{
    const uint32_t word = (uint32_t)(pos >> 5), sh = (uint32_t)pos & 31u;
    atomicOr(out + word, v << sh);
    if (sh + n > 32)
        atomicOr(out + word + 1, v >> (32 - sh));
}

// Greedy parse of one tile row.  WRITE = false: returns the bit count.  WRITE = true: emits the bits
// at position pos and returns the end position.
template <bool WRITE>
__device__ __forceinline__ unsigned long long parse_row(const uint8_t *tile, int r, const RowMasks &m, uint32_t *out,
                                                        unsigned long long pos)
{
    const uint8_t *row = tile + r * kTileStride;
    int x = 0;
    while (x < kTile) {
        const int la = r > 0 ? run_len<true>(row, m.above, 0, x) : 0;
        int lr = 0;
        if (x > 0) {
            // the left mask says "word == copies of the previous word's last byte"; starting inside a run of
            // row[x-1] that chain is exactly "equals row[x-1]"
            lr = run_len<false>(row, m.left, bcast_byte(row[x - 1]), x);
        }
        const int len = la >= lr ? la : lr;
        uint32_t bits;
        int n;
        if (len >= 3) {
            match_code(len, la >= lr, bits, n);     // len <= 256 < 258
            x += len;
        }
        else {
            lit_code(row[x], bits, n);
            x += 1;
        }
        if (WRITE)
            put_bits(out, pos, bits, n);
        pos += n;
    }
    return pos;
}

// One CTA = one tile of one plane.  grid = (tiles_x, tile_rows, planes), 256 threads.
__global__ void __launch_bounds__(kTile, 2)
deflate_tiles_kernel(const __grid_constant__ TileEncParams p)
{
    extern __shared__ __align__(16) uint8_t smem_enc[];
    uint8_t *tile = smem_enc;
    uint32_t *out = reinterpret_cast<uint32_t *>(smem_enc + kTile * kTileStride);
    __shared__ unsigned long long s_scan[kTile / 32];
    __shared__ unsigned long long s_a[kTile / 32], s_b[kTile / 32];
    __shared__ unsigned long long s_total_bits, s_dst;
    __shared__ uint32_t s_adler;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y, pl = blockIdx.z;
    const uint8_t *src = p.plane[pl];
    const int x0 = tx * kTile, y0 = ty * kTile;

    // ---- tile -> shared memory, zero padded at the right / bottom edge (like the CPU writer)
    const bool interior = x0 + kTile <= p.w && y0 + kTile <= p.rows &&
                          ((reinterpret_cast<uintptr_t>(src) | p.pitch) & 15) == 0;
    if (interior) {
        // all 16 independent 16-byte loads of a thread are in flight before the first store
        const uint8_t *g = src + (size_t)(y0 + (tid >> 4)) * p.pitch + x0 + (tid & 15) * 16;
        uint4 v[16];
#pragma unroll
        for (int k = 0; k < 16; k++)
            v[k] = __ldcs(reinterpret_cast<const uint4 *>(g + (size_t)k * 16 * p.pitch));
#pragma unroll
        for (int k = 0; k < 16; k++)
            *reinterpret_cast<uint4 *>(tile + ((tid >> 4) + 16 * k) * kTileStride + (tid & 15) * 16) = v[k];
    }
    else {
        for (int i = tid; i < kTile * (kTile / 16); i += kTile) {
            const int r = i >> 4, c = (i & 15) * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            const int gy = y0 + r, gx = x0 + c;
            if (gy < p.rows && gx < p.w) {
                const uint8_t *g = src + (size_t)gy * p.pitch + gx;
                if (gx + 16 <= p.w && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
                    v = *reinterpret_cast<const uint4 *>(g);
                }
                else {
                    uint32_t w4[4] = { 0, 0, 0, 0 };
                    for (int k = 0; k < 16 && gx + k < p.w; k++)
                        w4[k >> 2] |= (uint32_t)g[k] << (8 * (k & 3));
                    v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
            }
            *reinterpret_cast<uint4 *>(tile + r * kTileStride + c) = v;
        }
    }
    for (int i = tid; i < kEncCap / 4 + 8; i += kTile)
        out[i] = 0;
    __syncthreads();

    // ---- pass 1 (thread r owns tile row r): word masks, Adler-32 partial sums, bits per row
    RowMasks m;
    m.above = 0;
    m.left = 0;
    unsigned long long sa = 0, sb = 0;
    {
        const uint4 *rowv = reinterpret_cast<const uint4 *>(tile + tid * kTileStride);
        const uint4 *upv = rowv - kTileStride / 16;
        uint32_t s1 = 0, s2 = 0, last = 0;
#pragma unroll 4
        for (int i = 0; i < kTile / 16; i++) {
            const uint4 cv = rowv[i];
            const uint4 uv = tid > 0 ? upv[i] : make_uint4(~cv.x, ~cv.y, ~cv.z, ~cv.w);
            const uint32_t cw[4] = { cv.x, cv.y, cv.z, cv.w };
            const uint32_t uw[4] = { uv.x, uv.y, uv.z, uv.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = 4 * i + k;
                if (cw[k] == uw[k])
                    m.above |= 1ull << j;
                if (j > 0 && cw[k] == last * 0x01010101u)
                    m.left |= 1ull << j;
                last = cw[k] >> 24;
                const uint32_t s = __dp4a(cw[k], 0x01010101u, 0u);           // sum of the four bytes
                s1 += s;
                s2 += 4u * j * s + __dp4a(cw[k], 0x03020100u, 0u);           // sum of x * byte
            }
        }
        sa = s1;
        // contribution of this row to s2 = N + sum_i (N - i) d_i with i = 256 r + x
        sb = (unsigned long long)(kTileBytes - kTile * tid) * s1 - s2;
    }
    const unsigned long long row_bits = parse_row<false>(tile, tid, m, nullptr, 0);
    // inclusive scan of row_bits over the block + block sums of sa / sb
    unsigned long long inc = row_bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_down_sync(0xffffffffu, sa, o);
        sb += __shfl_down_sync(0xffffffffu, sb, o);
    }
    if (lane == 31)
        s_scan[warp] = inc;
    if (lane == 0) {
        s_a[warp] = sa;
        s_b[warp] = sb;
    }
    __syncthreads();
    unsigned long long base = 16 + 3;           // zlib header (16 bits) + block header (3 bits)
    for (int wv = 0; wv < warp; wv++)
        base += s_scan[wv];
    const unsigned long long row_pos = base + inc - row_bits;
    if (tid == kTile - 1) {
        s_total_bits = row_pos + row_bits + 7;  // + end-of-block code (7 zero bits)
        unsigned long long A = 0, B = 0;
        for (int wv = 0; wv < kTile / 32; wv++) {
            A += s_a[wv];
            B += s_b[wv];
        }
        const uint32_t s1 = (uint32_t)((1 + A) % 65521ull);
        const uint32_t s2 = (uint32_t)((kTileBytes + B) % 65521ull);
        s_adler = (s2 << 16) | s1;
    }
    __syncthreads();
    const unsigned long long total_bits = s_total_bits;
    const uint32_t deflate_end = (uint32_t)((total_bits + 7) >> 3);     // bytes incl. the 2-byte zlib header
    const bool stored = deflate_end + 4 > (uint32_t)kEncCap;
    const uint32_t nbytes = stored ? (uint32_t)kStoredBytes : deflate_end + 4;

    if (tid == 0) {
        const unsigned long long need = ((unsigned long long)nbytes + 15ull) & ~15ull;
        const unsigned long long off = atomicAdd(p.cursor, need);
        s_dst = off;
        const size_t ti = ((size_t)pl * p.tile_rows + ty) * p.tiles_x + tx;
        p.offsets[ti] = off;
        p.sizes[ti] = nbytes;
    }

    if (!stored) {
        // ---- pass 2: write the bit stream into shared memory
        if (tid == 0) {
            put_bits(out, 0, 0x9C78u, 16);      // CMF = 0x78 (deflate, 32K window), FLG = 0x9C
            put_bits(out, 16, 0x3u, 3);         // BFINAL = 1, BTYPE = 01 (fixed Huffman), LSB first
        }
        parse_row<true>(tile, tid, m, out, row_pos);
        __syncthreads();
        if (tid == 0) {
            // end-of-block is seven zero bits (already zero); Adler-32 big endian after the padding
            uint8_t *ob = reinterpret_cast<uint8_t *>(out);
            const uint32_t a = s_adler;
            ob[deflate_end + 0] = (uint8_t)(a >> 24);
            ob[deflate_end + 1] = (uint8_t)(a >> 16);
            ob[deflate_end + 2] = (uint8_t)(a >> 8);
            ob[deflate_end + 3] = (uint8_t)a;
        }
        __syncthreads();
        uint4 *dst = reinterpret_cast<uint4 *>(p.blob + s_dst);
        const uint4 *so = reinterpret_cast<const uint4 *>(out);
        for (uint32_t i = tid; i < (nbytes + 15) / 16; i += kTile)
            dst[i] = so[i];
    }
    else {
        // ---- incompressible tile: zlib header, two stored blocks of 32768 bytes, Adler-32
        __syncthreads();
        uint8_t *dst = p.blob + s_dst;
        if (tid == 0) {
            dst[0] = 0x78; dst[1] = 0x9C;
            dst[2] = 0x00; dst[3] = 0x00; dst[4] = 0x80; dst[5] = 0xFF; dst[6] = 0x7F;                     // BFINAL 0, LEN 32768
            uint8_t *b2 = dst + 7 + 32768;
            b2[0] = 0x01; b2[1] = 0x00; b2[2] = 0x80; b2[3] = 0xFF; b2[4] = 0x7F;                           // BFINAL 1
            const uint32_t a = s_adler;
            uint8_t *ad = dst + kStoredBytes - 4;
            ad[0] = (uint8_t)(a >> 24); ad[1] = (uint8_t)(a >> 16); ad[2] = (uint8_t)(a >> 8); ad[3] = (uint8_t)a;
        }
        for (int i = tid; i < kTileBytes; i += kTile) {
            const int r = i >> 8, c = i & 255;
            const int o = i < 32768 ? 7 + i : 7 + 5 + i;
            dst[o] = tile[r * kTileStride + c];
        }
    }
}

}  // namespace gcn10
