// gcn10_b200/csrc/inflate_tiles.cuh -- GPU-side inflate of DEFLATE-compressed GeoTIFF tiles (sm_100a).
//
// Input side of the block pipeline: the reference's load_raster() (/root/reference/src/raster.c:106-189)
// has GDAL decode the land-cover window on the CPU; here the compressed tiles of the window go to the
// device as they lie in the file and one warp per tile inflates them straight into the land-cover plane
// the Curve Number kernel reads (window clipping included), so PCIe carries ~1/20 of the raster.
//
// One CTA = one tile (zlib stream) = three warps in a pipeline:
//   decoder warp  tops up the 2 KB compressed-input ring (16-byte loads by all lanes); lane 0 owns the bit
//                 reader and turns code words into batches of up to 32 LZ77 symbols (inflate_core.h); block
//                 headers are parsed by lane 0, the Huffman lookup tables are filled by all 32 lanes.
//   writer warp   executes the batches against an 8 KB history ring in shared memory: a warp scan of the
//                 symbol lengths gives every symbol its output position; per ~1 KB part of a batch all literals are
//                 written at once, then the matches that reach back further than the ring (4 % of them: their
//                 source is read back from the plane, where the flusher has put it), then the others one after the
//                 other (32 or 128 bytes per step).
//   flusher warp  writes finished 2 KB pieces of the ring to the land-cover plane in HBM with 16-byte stores,
//                 clipped to the requested window, while the writer goes on (at most two pieces in flight), and
//                 accumulates the stream's Adler-32.  Tiles that the window clips also go to a scratch copy, so
//                 that their older history can be read back like that of any other tile.
// The warps hand over through a double-buffered symbol queue and a two-deep piece queue, both guarded by
// mbarriers in shared memory, so decoding batch k+1, executing batch k and storing the bytes of earlier batches
// overlap.  Shared memory per CTA: 18 KB (history 8 KB, tables 7.6 KB, input ring 2 KB, queues 0.5 KB) -> eleven
// tiles in flight per SM: the 1296 tiles of a 36000 x 36000 block are all resident at once.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.h"

namespace gcn10 {

constexpr int kInflateMaxParts = 9;     // a block window over a mosaic touches at most 3 x 3 sources of its own size

// One source raster of a mosaic (a single GeoTIFF = one part covering the whole window): its tile grid and the
// rectangle of the destination plane it fills.
struct InflatePart {
    int first_tile;                     // index of the part's first tile in offsets / sizes / status
    int tiles_x, tiles_y;               // tile grid handed over
    int tile_w, tile_h;                 // TIFF TileWidth / TileLength
    int tw_shift;                       // log2(tile_w) when it is a power of two, else -1
    int x_off, y_off;                   // position of the rectangle's pixel (0, 0) inside the tile grid
    uint8_t *dst;                       // the rectangle's pixel (0, 0) in the destination plane
    int w, h;                           // rectangle size: pixels outside are decoded but not stored
};

struct InflateParams {
    const uint8_t *blob;                // device copy of the compressed tiles; readable 4 KB past the last one
    const unsigned long long *offsets;  // per tile: byte offset of its zlib stream in blob
    const uint32_t *sizes;              // per tile: bytes; 0 = sparse tile (all zero, as GDAL reads it)
    size_t pitch;                       // of the destination plane
    int *status;                        // per tile: 0 or an inflate::kErr* code
    const int *order;                   // launch order: CTA i takes tile order[i] (longest streams first), or NULL
    int probe;                          // measurement aid: 1 = the writer warp drops the batches (decoder speed alone),
                                        // 2 = the writer runs but does not flush the ring to the plane
    int nparts;
    uint8_t *scratch;                   // private copies of the tiles the window clips (their older history)
    const int *scratch_index;           // per tile: its slot in `scratch`, or -1 = the plane holds the whole tile
    size_t scratch_stride;              // bytes per slot
    InflatePart part[kInflateMaxParts];
};

struct InflateSmem {
    inflate::Tables t;
    uint32_t ring[inflate::kRingWords];
    uint32_t queue[2][inflate::kQueue];
    int meta[2][8];                     // n, event, stored_src, stored_len, decoder error, final flag
    volatile int writer_err;
    volatile uint32_t fl_pos[2], fl_len[2];     // ring pieces handed to the flusher warp (len 0 = quit)
    uint32_t adler_want, adler_got;             // trailer of the stream (decoder warp) / sum over the decoded bytes (flusher warp)
    int trailer_state;                          // 0 = the decoder never reached the trailer, 1 = read, 2 = stream ends before it
    int final_err;                              // the writer warp's verdict
    int pad[3];
    // hand-over between the warps: mbarriers (count 1: the elected lane of the producing warp arrives, the whole
    // consuming warp waits).  Named barriers would do, but an SM has 16 per resident CTA slot at 4 CTAs, and eight
    // of them per tile would cap the SM at seven tiles.
    unsigned long long bar_full[2], bar_empty[2];       // decoder -> writer: batch b queued / writer -> decoder: consumed
    unsigned long long bar_ready[2], bar_done[2];       // writer -> flusher: piece k posted / flusher -> writer: stored
    uint8_t window[inflate::kWindow];
};

constexpr int kInflateSmem = (int)sizeof(InflateSmem);
constexpr int kInflateThreads = 96;      // decoder warp, writer warp, flusher warp
constexpr uint32_t kFlushChunk = inflate::kPiece;

struct TileDst {
    uint8_t *dst;
    size_t pitch;
    int dx0, dy0, w, h, tile_w, tw_shift;
    bool rows16;                        // tile_w % 16 == 0: a 16-byte group never straddles tile rows
    // where the tile's older history is read back from: its own rectangle of the plane, or -- for a tile that the
    // window clips (its bytes are not all stored in the plane) -- a private scratch copy of the whole tile
    uint8_t *hist;                      // address of tile pixel (0, 0)
    size_t hist_pitch;
    bool scratch;                       // hist is a scratch copy: the flusher fills it as well
};

// one-producer / one-consumer-warp mbarrier hand-over
__device__ __forceinline__ void warp_signal(unsigned long long *bar, int lane)
{
    __syncwarp();
    if (lane == 0)
        mbar_arrive(smem_u32(bar));
}
__device__ __forceinline__ void warp_wait(unsigned long long *bar, uint32_t parity) { mbar_wait(smem_u32(bar), parity); }

// The writer warp addresses the history ring through its 32-bit shared-memory address with explicit ld.shared /
// st.shared: through a generic pointer the compiler rebuilds the shared window base (S2R SR_CgaCtaId + LEA) at
// every predicated access of its copy loops.
struct Ring {
    uint32_t base;              // shared-space address of window[0]
    __device__ __forceinline__ uint32_t at(uint32_t pos) const { return base + (pos & (uint32_t)(inflate::kWindow - 1)); }
    __device__ __forceinline__ uint32_t ld8(uint32_t pos) const
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(at(pos)) : "memory");
        return v;
    }
    __device__ __forceinline__ void st8(uint32_t pos, uint32_t v) const
    {
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(at(pos)), "r"(v) : "memory");
    }
};

// byte `pos` of the tile as the flusher stored it (read around L1: it was written by another warp of this CTA)
__device__ __forceinline__ uint32_t hist_byte(const TileDst &d, uint32_t pos)
{
    const uint32_t r = d.tw_shift >= 0 ? pos >> d.tw_shift : pos / (uint32_t)d.tile_w;
    const uint32_t c = pos - r * (uint32_t)d.tile_w;
    return (uint32_t)__ldcg(d.hist + (size_t)r * d.hist_pitch + c);
}

// history ring piece [p0, p0 + nbytes) -> plane (p0 is a multiple of 16); whole warp
__device__ __forceinline__ void inflate_flush(const uint8_t *window, const TileDst &d, uint32_t p0, uint32_t nbytes,
                                              int lane)
{
    using inflate::kWindow;
    if (d.scratch) {
        // the private copy takes every byte of the piece, clipped or not
        for (uint32_t i = lane; i < nbytes; i += 32u) {
            const uint32_t p = p0 + i;
            const uint32_t r = d.tw_shift >= 0 ? p >> d.tw_shift : p / (uint32_t)d.tile_w;
            d.hist[(size_t)r * d.hist_pitch + (p - r * (uint32_t)d.tile_w)] = window[p & (kWindow - 1)];
        }
    }
    if (d.rows16) {
        for (uint32_t g = lane; g * 16u < nbytes; g += 32u) {
            const uint32_t p = p0 + 16u * g;
            const uint32_t r = d.tw_shift >= 0 ? p >> d.tw_shift : p / (uint32_t)d.tile_w;
            const uint32_t c = p - r * (uint32_t)d.tile_w;
            const int gy = d.dy0 + (int)r, gx = d.dx0 + (int)c;
            if ((unsigned)gy >= (unsigned)d.h || gx + 16 <= 0 || gx >= d.w)
                continue;
            uint8_t *o = d.dst + (size_t)gy * d.pitch + gx;
            const uint32_t left = nbytes - 16u * g;
            if (gx >= 0 && gx + 16 <= d.w && left >= 16u && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
                *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<const uint4 *>(window + (p & (kWindow - 1)));
            }
            else {
                const uint32_t m = left < 16u ? left : 16u;
                for (uint32_t k = 0; k < m; k++)
                    if ((unsigned)(gx + (int)k) < (unsigned)d.w)
                        o[k] = window[(p + k) & (kWindow - 1)];
            }
        }
    }
    else {
        for (uint32_t i = lane; i < nbytes; i += 32u) {
            const uint32_t p = p0 + i;
            const uint32_t r = p / (uint32_t)d.tile_w, c = p - r * (uint32_t)d.tile_w;
            const int gy = d.dy0 + (int)r, gx = d.dx0 + (int)c;
            if ((unsigned)gy < (unsigned)d.h && (unsigned)gx < (unsigned)d.w)
                d.dst[(size_t)gy * d.pitch + gx] = window[p & (kWindow - 1)];
        }
    }
}

// Adler-32 over ring piece [p0, p0 + n) (p0 a multiple of 16), whole warp: per 16-byte group the byte sum and the
// sum of k * d[k] come from dp4a; s1 / s2 stay uniform across the lanes.  zlib verifies this checksum when GDAL reads
// a tile (raster.c:177-186 reports the failure); so does the flusher warp, on everything the stream decodes to.
__device__ __forceinline__ void inflate_adler_piece(const uint8_t *window, uint32_t p0, uint32_t n, int lane,
                                                    uint32_t &s1, uint32_t &s2)
{
    using inflate::kWindow;
    uint32_t a = 0;
    unsigned long long b = 0;
    for (uint32_t g = lane; g * 16u < n; g += 32u) {
        const uint32_t o = 16u * g, left = n - o;
        const uint4 v = *reinterpret_cast<const uint4 *>(window + ((p0 + o) & (kWindow - 1)));
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
        if (left < 16u) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int valid = (int)left - 4 * j;
                w[j] = valid >= 4 ? w[j] : (valid <= 0 ? 0u : w[j] & ((1u << (8 * valid)) - 1u));
            }
        }
        uint32_t sum = 0, ksum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sum = __dp4a(w[j], 0x01010101u, sum);
            ksum = __dp4a(w[j], 0x03020100u + 0x04040404u * (uint32_t)j, ksum);
        }
        a += sum;
        b += (unsigned long long)left * sum - ksum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    inflate::adler_advance(s1, s2, n, a, b);
}

// copy of one match by the whole warp: bytes [mp, mp + len) := bytes [mp - dist, ...) of the history ring
__device__ __forceinline__ void inflate_copy(const Ring window, uint32_t mp, uint32_t len, uint32_t dist, int lane)
{
    if (dist >= len) {
        // source and destination do not overlap (nearly every match): a plain 32-lane loop, one barrier
        for (uint32_t i = lane; i < len; i += 32u)
            window.st8(mp + i, window.ld8(mp - dist + i));
        __syncwarp();
    }
    else if (dist >= 128u) {
        // overlapping (a 258-byte match at distance 256: the row above in a 256-wide tile), but a 128-byte step
        // never reads what the same step writes
        for (uint32_t b = 0; b < len; b += 128u) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = b + lane + 32u * k;
                v[k] = i < len ? window.ld8(mp - dist + i) : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = b + lane + 32u * k;
                if (i < len)
                    window.st8(mp + i, v[k]);
            }
            __syncwarp();
        }
    }
    else if (dist >= 32u) {
        // overlapping, but a 32-byte step never reads what the same step writes
        for (uint32_t b = 0; b < len; b += 32u) {
            const uint32_t i = b + lane;
            if (i < len)
                window.st8(mp + i, window.ld8(mp - dist + i));
            __syncwarp();
        }
    }
    else if (dist == 1u) {
        const uint32_t v = window.ld8(mp - 1u);
        for (uint32_t i = lane; i < len; i += 32u)
            window.st8(mp + i, v);
        __syncwarp();
    }
    else {
        // the pattern of the last `dist` bytes repeats; sources all lie before mp
        uint32_t q = (uint32_t)lane % dist;
        const uint32_t step = 32u % dist;
        for (uint32_t i = lane; i < len; i += 32u) {
            window.st8(mp + i, window.ld8(mp - dist + q));
            q += step;
            if (q >= dist)
                q -= dist;
        }
        __syncwarp();
    }
}

// a match whose source starts below the ring (p1 = end of the part, see inflate::byte_from_plane): the bytes come
// from the plane / scratch copy where the flusher put them, or -- the tail of a source that straddles -- from the ring;
// source and destination never overlap here (the source ends more than a ring length below the destination)
__device__ __forceinline__ void inflate_copy_far(const Ring window, const TileDst &d, uint32_t mp, uint32_t len,
                                                 uint32_t dist, uint32_t p1, int lane)
{
    for (uint32_t i = lane; i < len; i += 32u) {
        const uint32_t s = mp - dist + i;
        window.st8(mp + i, inflate::byte_from_plane(s, p1) ? hist_byte(d, s) : window.ld8(s));
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kInflateThreads)
inflate_tiles_kernel(const __grid_constant__ InflateParams p)
{
    using namespace inflate;
    extern __shared__ __align__(16) uint8_t smem_inf[];
    InflateSmem &sm = *reinterpret_cast<InflateSmem *>(smem_inf);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = p.order ? p.order[blockIdx.x] : (int)blockIdx.x;
    int pi = 0;
    while (pi + 1 < p.nparts && tile >= p.part[pi + 1].first_tile)
        pi++;
    const InflatePart &pt = p.part[pi];
    const int local = tile - pt.first_tile;
    const int tx = local % pt.tiles_x, ty = local / pt.tiles_x;

    TileDst d;
    d.dst = pt.dst;
    d.pitch = p.pitch;
    d.dx0 = tx * pt.tile_w - pt.x_off;
    d.dy0 = ty * pt.tile_h - pt.y_off;
    d.w = pt.w;
    d.h = pt.h;
    d.tile_w = pt.tile_w;
    d.tw_shift = pt.tw_shift;
    d.rows16 = (pt.tile_w & 15) == 0;
    // older history: the tile's own rectangle of the plane when all of it is stored there, else a scratch copy
    const int sidx = p.scratch_index ? p.scratch_index[tile] : -1;
    d.scratch = sidx >= 0;
    if (d.scratch) {
        d.hist = p.scratch + (size_t)sidx * (size_t)p.scratch_stride;
        d.hist_pitch = (size_t)pt.tile_w;
    }
    else {
        d.hist = pt.dst + (ptrdiff_t)d.dy0 * (ptrdiff_t)p.pitch + d.dx0;
        d.hist_pitch = p.pitch;
    }

    const uint32_t size = p.sizes[tile];
    if (size == 0) {
        // sparse tile: GDAL returns zeros for a tile without data
        const int x0 = max(d.dx0, 0), x1 = min(d.dx0 + pt.tile_w, pt.w);
        const int y0 = max(d.dy0, 0), y1 = min(d.dy0 + pt.tile_h, pt.h);
        for (int y = y0 + warp; y < y1; y += kInflateThreads / 32)
            for (int x = x0 + lane; x < x1; x += 32)
                pt.dst[(size_t)y * p.pitch + x] = 0;
        if (threadIdx.x == 0)
            p.status[tile] = 0;
        return;
    }

    const uint8_t *src = p.blob + p.offsets[tile];
    const uint8_t *base = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
    const uint32_t first = (uint32_t)(src - base);
    const uint32_t out_end = (uint32_t)pt.tile_w * (uint32_t)pt.tile_h;
    if (threadIdx.x == 0) {
        sm.writer_err = 0;
        for (int k = 0; k < 2; k++) {
            mbar_init(smem_u32(&sm.bar_full[k]), 1);
            mbar_init(smem_u32(&sm.bar_empty[k]), 1);
            mbar_init(smem_u32(&sm.bar_ready[k]), 1);
            mbar_init(smem_u32(&sm.bar_done[k]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ------------------------------------------------------------------ decoder warp
        uint32_t filled = 0;            // [base, base + filled) has been staged; the ring holds its last 2 KB
        auto top_up = [&](uint32_t cons) {
            while (filled < cons + 1024u) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + filled) + lane);
                *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(sm.ring) + ((filled + 16u * lane) & 2047u)) = v;
                filled += 512u;
            }
            __syncwarp();
        };
        DecodeLane s;
        lane_init(s, first, first + size);
        top_up(first & ~3u);
        if (lane == 0)
            s.err = read_zlib_header(s, sm.ring, first);
        int b = 0;
        uint32_t use = 0;               // hand-overs so far: buffer b = use & 1 is on its (use >> 1)-th use
        // waiting for the phase BEFORE a barrier's first one returns at once: the first use of a buffer does not block
        auto wait_empty = [&]() { warp_wait(&sm.bar_empty[b], ((use >> 1) + 1u) & 1u); };
        for (;;) {
            top_up(__shfl_sync(full, s.cons, 0));
            if (lane == 0 && !s.err && sm.writer_err)
                s.err = sm.writer_err;
            const int err = __shfl_sync(full, s.err, 0);
            const int in_block = __shfl_sync(full, s.in_block, 0);
            int n = 0, ev = kEvMore, fin = 0;
            uint32_t so = 0, sl = 0;
            bool post = false, acquired = false;
            if (err) {
                ev = kEvError;
                post = true;
            }
            else if (!in_block) {
                int action = 0;
                if (lane == 0)
                    action = read_block_header(s, sm.ring, sm.t);
                action = __shfl_sync(full, action, 0);
                if (action == kHdrBuild) {
                    __syncwarp();
                    clear_block_luts(sm.t, lane, 32);
                    __syncwarp();
                    fill_block_luts(sm.t, lane, 32);
                    __syncwarp();
                }
                else if (action == kHdrStored) {
                    so = __shfl_sync(full, s.stored_src, 0);
                    sl = __shfl_sync(full, s.stored_len, 0);
                    fin = __shfl_sync(full, s.bfinal, 0);
                    const uint32_t q = so + sl;
                    filled = q & ~511u;
                    top_up(q & ~3u);
                    if (lane == 0)
                        seek(s, sm.ring, q);
                    ev = kEvStored;
                    post = true;
                }
                else if (action == kHdrError) {
                    ev = kEvError;
                    post = true;
                }
            }
            else {
                wait_empty();
                acquired = true;
                if (lane == 0)
                    n = decode_symbols(s, sm.ring, sm.t, sm.queue[b], &ev);
                ev = __shfl_sync(full, ev, 0);
                post = true;
            }
            if (post) {
                if (!acquired)
                    wait_empty();
                if (lane == 0) {
                    sm.meta[b][0] = n;
                    sm.meta[b][1] = ev;
                    sm.meta[b][2] = (int)so;
                    sm.meta[b][3] = (int)sl;
                    sm.meta[b][4] = s.err;
                    sm.meta[b][5] = fin;
                }
                warp_signal(&sm.bar_full[b], lane);
                b ^= 1;
                use++;
                if (ev == kEvEnd || ev == kEvError || (ev == kEvStored && fin))
                    break;
            }
        }
        // behind the final block: the Adler-32 trailer (compared with the flusher warp's sum at the end)
        top_up(__shfl_sync(full, s.cons, 0));
        if (lane == 0) {
            uint32_t want = 0;
            sm.trailer_state = s.err ? 0 : (read_adler_trailer(s, sm.ring, &want) ? 1 : 2);
            sm.adler_want = want;
        }
    }
    else if (warp == 2) {
        // ------------------------------------------------------------------ flusher warp: ring pieces -> plane
        uint32_t s1 = 1u, s2 = 0u;
        for (uint32_t n = 0;; n++) {
            const int k = (int)(n & 1u);
            warp_wait(&sm.bar_ready[k], (n >> 1) & 1u);
            const uint32_t pos = sm.fl_pos[k], len = sm.fl_len[k];
            if (len == 0)
                break;
            inflate_flush(sm.window, d, pos, len, lane);
            inflate_adler_piece(sm.window, pos, len, lane, s1, s2);
            warp_signal(&sm.bar_done[k], lane);
        }
        if (lane == 0)
            sm.adler_got = (s2 << 16) | s1;
    }
    else {
        // ------------------------------------------------------------------ writer warp
        Ring window;
        window.base = (uint32_t)__cvta_generic_to_shared(sm.window);
        uint32_t out_base = 0, flushed = 0;
        uint32_t npost = 0;
        // hands ring piece [pos, pos + len) to the flusher warp (len 0: no more pieces).  At most two are in flight:
        // slot k is reused only after its previous piece has been stored, which is also what keeps the ring positions
        // the writer is about to overwrite -- and the plane bytes a far match reads -- valid (inflate_core.h)
        auto post = [&](uint32_t pos, uint32_t len) {
            const int k = (int)(npost & 1u);
            warp_wait(&sm.bar_done[k], ((npost >> 1) + 1u) & 1u);
            if (lane == 0) {
                sm.fl_pos[k] = pos;
                sm.fl_len[k] = len;
            }
            warp_signal(&sm.bar_ready[k], lane);
            npost++;
        };
        auto post_upto = [&](uint32_t pos) {
            while (flushed + kFlushChunk <= pos) {
                if (p.probe != 2)
                    post(flushed, kFlushChunk);
                flushed += kFlushChunk;
            }
        };
        int werr = 0, b = 0;
        for (uint32_t use = 0;; use++) {
            warp_wait(&sm.bar_full[b], (use >> 1) & 1u);
            const int n = sm.meta[b][0], ev = sm.meta[b][1], derr = sm.meta[b][4], fin = sm.meta[b][5];
            const uint32_t so = (uint32_t)sm.meta[b][2], sl = (uint32_t)sm.meta[b][3];
            const uint32_t sym = lane < n ? sm.queue[b][lane] : 0u;
            warp_signal(&sm.bar_empty[b], lane);
            b ^= 1;

            if (n > 0 && !werr && p.probe == 1) {
                out_base = out_end;             // pretend the tile is complete; nothing is written
            }
            else if (n > 0 && !werr) {
                const bool is_match = sym_is_match(sym) != 0u;
                const uint32_t l = lane < n ? sym_len(sym) : 0u;
                uint32_t inc = l;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(full, inc, o);
                    if (lane >= o)
                        inc += t;
                }
                const uint32_t start = out_base + inc - l;
                const uint32_t total = __shfl_sync(full, inc, 31);
                const unsigned bad = __ballot_sync(full, is_match && sym_dist(sym) > start);
                if (out_base + total > out_end)
                    werr = kErrOverflow;
                else if (bad)
                    werr = kErrDistance;
                else {
                    // The batch runs in parts (inflate_core.h): per part all literals at once, then the matches that
                    // start below the ring (read back from the plane, independent of the part), then the others in
                    // stream order; the finished pieces of the ring go to the flusher after every part.
                    const uint32_t my_part = lane < n ? (inc - 1u) / (uint32_t)kPart : 0xFFFFFFFFu;
                    const uint32_t nparts = (total - 1u) / (uint32_t)kPart + 1u;
                    for (uint32_t pid = 0; pid < nparts; pid++) {
                        const unsigned members = __ballot_sync(full, my_part == pid);
                        if (!members)
                            continue;
                        const uint32_t p1 = out_base + __shfl_sync(full, inc, 31 - __clz(members));
                        const bool mine = ((members >> lane) & 1u) != 0u;
                        if (mine && !is_match)
                            window.st8(start, sym);
                        __syncwarp();
                        const unsigned mm_all = __ballot_sync(full, mine && is_match);
                        unsigned far = __ballot_sync(full, mine && is_match && byte_from_plane(start - sym_dist(sym), p1));
                        unsigned mm = mm_all & ~far;
                        while (far) {
                            const int owner = __ffs(far) - 1;
                            far &= far - 1;
                            const uint32_t ms = __shfl_sync(full, sym, owner);
                            const uint32_t mp = __shfl_sync(full, start, owner);
                            inflate_copy_far(window, d, mp, ms & 0x1FFu, sym_dist(ms), p1, lane);
                        }
                        while (mm) {
                            const int owner = __ffs(mm) - 1;
                            mm &= mm - 1;
                            const uint32_t ms = __shfl_sync(full, sym, owner);
                            const uint32_t mp = __shfl_sync(full, start, owner);
                            inflate_copy(window, mp, ms & 0x1FFu, sym_dist(ms), lane);
                        }
                        post_upto(p1);
                    }
                    out_base += total;
                }
            }
            if (ev == kEvStored && !werr) {
                if (out_base + sl > out_end)
                    werr = kErrOverflow;
                else {
                    // raw bytes: global -> history ring, flushed piecewise (a stored block can exceed the ring)
                    uint32_t done = 0;
                    while (done < sl) {
                        const uint32_t m = min(sl - done, (uint32_t)kPart);
                        for (uint32_t i = lane; i < m; i += 32u)
                            window.st8(out_base + i, base[so + done + i]);
                        __syncwarp();
                        out_base += m;
                        done += m;
                        post_upto(out_base);
                    }
                }
            }
            if (werr && lane == 0)
                sm.writer_err = werr;
            if (ev == kEvEnd || (ev == kEvStored && fin)) {
                if (!werr && out_base != out_end)
                    werr = kErrShort;
                break;
            }
            if (ev == kEvError) {
                if (!werr)
                    werr = derr;
                break;
            }
        }
        if (!werr && flushed < out_end && p.probe != 2) {
            __syncwarp();
            post(flushed, out_end - flushed);
        }
        post(0, 0);
        if (lane == 0)
            sm.final_err = werr;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int e = sm.final_err;
        if (!e && !p.probe) {
            if (sm.trailer_state == 2)
                e = kErrInput;
            else if (sm.trailer_state == 1 && sm.adler_want != sm.adler_got)
                e = kErrChecksum;
        }
        p.status[tile] = e;
    }
}

}  // namespace gcn10
