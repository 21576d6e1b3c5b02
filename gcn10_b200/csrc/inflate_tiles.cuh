// gcn10_b200/csrc/inflate_tiles.cuh -- GPU-side inflate of DEFLATE-compressed GeoTIFF tiles (sm_100a).
//
// Input side of the block pipeline: the reference's load_raster() (/root/reference/src/raster.c:106-189)
// has GDAL decode the land-cover window on the CPU; here the compressed tiles of the window go to the
// device as they lie in the file and one warp per tile inflates them straight into the land-cover plane
// the Curve Number kernel reads (window clipping included), so PCIe carries ~1/20 of the raster.
//
// One CTA = one warp = one tile (zlib stream).  Shared memory per warp (41 KB, five warps per SM):
//   window   32 KB   the DEFLATE history as a ring indexed by output position (every distance <= 32768)
//   tables    7 KB   10-bit literal/length and 9-bit distance lookup + canonical-walk arrays (inflate_core.h)
//   ring      2 KB   compressed input, refilled 512 B at a time by all lanes (16-byte loads)
//   queue   128 B    one batch of LZ77 symbols
// Loop: the warp tops up the input ring; lane 0 runs one decode step (a block header, or up to 32
// symbols); a warp scan of the symbol lengths gives every symbol its output position; all literals are
// written at once; matches are copied one after the other, 32 bytes per step, from the history ring.
// Every byte goes to the ring and -- if it falls inside the requested window -- to the plane in HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.h"

namespace gcn10 {

struct InflateParams {
    const uint8_t *blob;                // device copy of the compressed tiles; readable 4 KB past the last one
    const unsigned long long *offsets;  // [tiles_y][tiles_x] byte offset of a tile's zlib stream in blob
    const uint32_t *sizes;              // [tiles_y][tiles_x] bytes; 0 = sparse tile (all zero, as GDAL reads it)
    int tiles_x, tiles_y;               // tile grid handed over
    int tile_w, tile_h;                 // TIFF TileWidth / TileLength
    int tw_shift;                       // log2(tile_w) when it is a power of two, else -1
    int x_off, y_off;                   // position of destination pixel (0, 0) inside the tile grid
    uint8_t *dst;                       // destination plane (row 0 of the window)
    size_t pitch;
    int w, h;                           // window size: pixels outside are decoded but not stored
    int *status;                        // [tiles_y][tiles_x] 0 or an inflate::kErr* code
};

struct InflateSmem {
    inflate::Tables t;
    uint32_t ring[inflate::kRingWords];
    uint32_t queue[inflate::kQueue];
    uint8_t window[inflate::kWindow];
};

constexpr int kInflateSmem = (int)sizeof(InflateSmem);

struct TileDst {
    uint8_t *dst;
    size_t pitch;
    int dx0, dy0, w, h, tile_w, tw_shift;
};

__device__ __forceinline__ void inflate_emit(uint8_t *window, const TileDst &d, uint32_t pos, uint32_t byte)
{
    window[pos & (inflate::kWindow - 1)] = (uint8_t)byte;
    const uint32_t r = d.tw_shift >= 0 ? pos >> d.tw_shift : pos / (uint32_t)d.tile_w;
    const uint32_t c = pos - r * (uint32_t)d.tile_w;
    const int gy = d.dy0 + (int)r, gx = d.dx0 + (int)c;
    if ((unsigned)gy < (unsigned)d.h && (unsigned)gx < (unsigned)d.w)
        d.dst[(size_t)gy * d.pitch + gx] = (uint8_t)byte;
}

__global__ void __launch_bounds__(32)
inflate_tiles_kernel(const __grid_constant__ InflateParams p)
{
    using namespace inflate;
    extern __shared__ __align__(16) uint8_t smem_inf[];
    InflateSmem &sm = *reinterpret_cast<InflateSmem *>(smem_inf);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;

    TileDst d;
    d.dst = p.dst;
    d.pitch = p.pitch;
    d.dx0 = tx * p.tile_w - p.x_off;
    d.dy0 = ty * p.tile_h - p.y_off;
    d.w = p.w;
    d.h = p.h;
    d.tile_w = p.tile_w;
    d.tw_shift = p.tw_shift;

    const uint32_t size = p.sizes[tile];
    if (size == 0) {
        // sparse tile: GDAL returns zeros for a tile without data
        const int x0 = max(d.dx0, 0), x1 = min(d.dx0 + p.tile_w, p.w);
        const int y0 = max(d.dy0, 0), y1 = min(d.dy0 + p.tile_h, p.h);
        for (int y = y0; y < y1; y++)
            for (int x = x0 + lane; x < x1; x += 32)
                p.dst[(size_t)y * p.pitch + x] = 0;
        if (lane == 0)
            p.status[tile] = 0;
        return;
    }

    const uint8_t *src = p.blob + p.offsets[tile];
    const uint8_t *base = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
    const uint32_t first = (uint32_t)(src - base);
    uint32_t filled = 0;                // [base, base + filled) has been staged; the ring holds its last 2 KB

    auto top_up = [&](uint32_t cons) {
        while (filled < cons + 1024u) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + filled) + lane);
            *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(sm.ring) + ((filled + 16u * lane) & 2047u)) = v;
            filled += 512u;
        }
        __syncwarp();
    };

    DecodeLane s;
    lane_init(s, first, first + size, (uint32_t)p.tile_w * (uint32_t)p.tile_h);
    top_up(first & ~3u);
    if (lane == 0)
        s.err = read_zlib_header(s, sm.ring, first);
    int ev = __shfl_sync(full, s.err, 0) ? kEvError : kEvMore;
    uint32_t out_base = 0;

    while (ev != kEvError && ev != kEvEnd) {
        top_up(__shfl_sync(full, s.cons, 0));
        int n = 0;
        if (lane == 0)
            n = decode_step(s, sm.ring, sm.t, sm.queue, &ev);
        n = __shfl_sync(full, n, 0);
        ev = __shfl_sync(full, ev, 0);
        __syncwarp();

        if (n > 0) {
            const uint32_t sym = lane < n ? sm.queue[lane] : 0u;
            const bool is_match = (sym >> 31) != 0u;
            const uint32_t l = lane < n ? (is_match ? (sym & 0x1FFu) : 1u) : 0u;
            uint32_t inc = l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(full, inc, o);
                if (lane >= o)
                    inc += t;
            }
            const uint32_t start = out_base + inc - l;
            out_base += __shfl_sync(full, inc, 31);
            if (lane < n && !is_match)
                inflate_emit(sm.window, d, start, sym & 255u);
            __syncwarp();
            unsigned mm = __ballot_sync(full, is_match);
            while (mm) {
                const int owner = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t ms = __shfl_sync(full, sym, owner);
                const uint32_t mp = __shfl_sync(full, start, owner);
                const uint32_t len = ms & 0x1FFu, dist = ((ms >> 16) & 0x7FFFu) + 1u;
                if (dist >= 32u) {
                    // bytes of one 32-byte step never read what the same step writes
                    for (uint32_t b = 0; b < len; b += 32u) {
                        const uint32_t i = b + lane;
                        if (i < len)
                            inflate_emit(sm.window, d, mp + i, sm.window[(mp - dist + i) & (kWindow - 1)]);
                        __syncwarp();
                    }
                }
                else {
                    // overlapping copy: the pattern of the last `dist` bytes repeats
                    for (uint32_t i = lane; i < len; i += 32u)
                        inflate_emit(sm.window, d, mp + i, sm.window[(mp - dist + i % dist) & (kWindow - 1)]);
                    __syncwarp();
                }
            }
        }

        if (ev == kEvStored) {
            // raw bytes: global -> history ring + plane, then restart the bit reader behind them
            const uint32_t so = __shfl_sync(full, s.stored_src, 0), sl = __shfl_sync(full, s.stored_len, 0);
            for (uint32_t i = lane; i < sl; i += 32u)
                inflate_emit(sm.window, d, out_base + i, base[so + i]);
            out_base += sl;
            const uint32_t q = so + sl;
            filled = q & ~511u;
            top_up(q & ~3u);
            int done = 0;
            if (lane == 0) {
                s.out_pos += sl;
                seek(s, sm.ring, q);
                if (s.bfinal) {
                    done = 1;
                    if (s.out_pos != s.out_end)
                        s.err = kErrShort;
                }
                else if (q > s.in_end)
                    s.err = kErrInput;
            }
            done = __shfl_sync(full, done, 0);
            ev = __shfl_sync(full, s.err, 0) ? kEvError : (done ? kEvEnd : kEvMore);
        }
    }
    if (lane == 0)
        p.status[tile] = s.err;
}

}  // namespace gcn10
