// gcn10_b200/csrc/cn_kernels.cuh -- sm_100a device code for GCN10's Curve Number hot path.
//
// What the reference does in five CPU passes per output plane (/root/reference/src/cn.c):
//   resample (cn.c:218-232) -> memcpy (274) -> dual-HSG remap (275, 88-111) -> memset 255 (289)
//   -> table pass (290, 114-131), 18 times per block,
// is done here in ONE pass over the land-cover raster that writes all requested planes:
//
//   index_map_kernel   O(W+H) threads, fp64: the separable pixel->HSG-cell maps of cn.c:219-229,
//                      evaluated with explicitly rounded, never-fused IEEE operations.
//   cn_block_kernel    the streaming kernel (HBM bound, ~1 B read + NP*G B written per pixel):
//                      one thread owns a 16-pixel column group and walks down the rows of its
//                      CTA's row chunk.  Per row: one 16-byte load of land cover, 16 shared-memory
//                      record fetches (one 16-byte record = the 9 CN values of a (class, soil
//                      group) pair), a register byte-transpose (PRMT), and one 16-byte streaming
//                      store per output plane.  The HSG cells a CTA needs (a few rows x <=256
//                      columns of the 250 m grid) are staged into shared memory with one 2-D TMA
//                      tensor load; a thread re-gathers its 16 soil codes only when the HSG row
//                      changes (every ~25 raster rows).
//   cn_bytes_kernel    byte-wise version for the <16-pixel right edge and for misaligned buffers.
//
// No tensor cores: the path is a byte gather with no contraction (BASELINE.json north_star).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gcn10 {

constexpr int kThreads = 256;          // threads per CTA
constexpr int kVecPx = 16;             // pixels per thread per row (one 16-byte vector)
constexpr int kStripPx = kThreads * kVecPx;   // 4096 pixels of a row per CTA
constexpr int kBoxCols = 256;          // TMA box: HSG columns staged per CTA (max box extent)
constexpr int kBoxRows = 16;           // TMA box: HSG rows staged per CTA
constexpr int kLutRecords = 256 * 8;   // (land cover, slot) -> 16-byte record
constexpr int kLutBytes = kLutRecords * 16;
// Launches of at most four planes per drainage condition (BASELINE configs[0]: g_ii alone) use 4-byte records:
// with 16-byte records every pixel costs a 16-byte shared-memory read for one useful byte, and the 128 B/clk of an
// SM's shared memory (8 px/clk) caps a one-plane launch at 0.58 ms per 36000^2 tile -- below the HBM rate.  Narrow
// records are read with LDS.32 (32 px/clk).  Their row stride is 9 words so that the land-cover classes (multiples
// of 10) do not all fall onto the same eight banks; no swizzle is needed.
constexpr int kLut4Stride = 9;         // words per land-cover row of the narrow layout (slots 0..7 used, 8 = padding)
constexpr int kLut4Bytes = 256 * kLut4Stride * 4;
constexpr int kNarrowPlanes = 4;       // NP <= kNarrowPlanes -> narrow records
__host__ __device__ constexpr int lut_bytes_for(int np) { return np <= kNarrowPlanes ? kLut4Bytes : kLutBytes; }
constexpr int kSgInvalid = 5;          // soil-group slot whose record is all 255
#ifndef GCN10_PREFETCH
#define GCN10_PREFETCH 2
#endif
#ifndef GCN10_STORE_POLICY
#define GCN10_STORE_POLICY 0
#endif
constexpr int kPrefetch = GCN10_PREFETCH;   // rows of land cover in flight per thread
#ifndef GCN10_NARROW_PF
#define GCN10_NARROW_PF 2
#endif
#ifndef GCN10_NARROW_CTAS
#define GCN10_NARROW_CTAS 5
#endif
#ifndef GCN10_NARROW_DEPTH
#define GCN10_NARROW_DEPTH 6    // land-cover rows in flight per thread in the direct-store narrow kernels (cp.async ring)
#endif

struct BlockParams {
    const uint8_t *esa;         // first row of this launch
    size_t esa_pitch;
    int w, h;                   // pixels per row, rows in this launch
    int y_base;                 // row of the block that esa row 0 corresponds to (row_idx offset)
    const int32_t *col_idx;     // [roundup16(w)]  HSG column of each pixel column
    const int32_t *row_idx;     // [block rows]    HSG row of each pixel row
    const uint8_t *hsg;         // coarse window
    size_t hsg_pitch;
    int hsx, hsy;
    const uint4 *lut;           // kLutRecords 16-byte records or the narrow layout, see pack_lut_records()
    int rec_bytes;              // 16 or 4 (narrow layout)
    int swz_shift;              // bank swizzle of the 16-byte layout: slot = sg ^ ((lc >> swz_shift) & 7)
    int rows_per_cta;
    int use_tma;
    int group_drained[2];       // per output group: 1 = "drained" remap, 0 = "undrained"
    uint8_t *out[18];           // group g, plane k at out[g*NP + k]
    size_t out_pitch;
};

// ---------------------------------------------------------------------------------------------
// fp64 index maps (cn.c:219-229).  Every operation is an explicitly rounded intrinsic so that
// nvcc cannot contract a*b+c into an FMA: the reference build has no FMA (no -march,
// src/CMakeLists.txt:68) and a fused evaluation moves up to 62 tie columns per block.

__device__ __forceinline__ double c99_round(double v)
{
    // round half away from zero without an inexact v+0.5: v - trunc(v) is exact in fp64
    double t = trunc(v);
    if (fabs(__dsub_rn(v, t)) >= 0.5)
        t = __dadd_rn(t, copysign(1.0, v));
    return t;
}

__device__ __forceinline__ int int_from_double_x86(double v)
{
    // (int) as the reference's x86-64 object code performs it (cvttsd2si): INT_MIN for NaN and
    // for values outside int range.  v is already integral here.
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT_MIN;
    return (int)v;
}

__device__ __forceinline__ int clamp_index(int i, int n)
{
    return i < 0 ? 0 : (i >= n ? n - 1 : i);       // cn.c:228-229
}

__global__ void index_map_kernel(int w, int w_pad, int h,
                                 double g0, double g1, double g3, double g5,
                                 double s0, double s1, double s3, double s5,
                                 int hsx, int hsy, int32_t *__restrict__ col_idx, int32_t *__restrict__ row_idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w_pad) {
        int x = i < w ? i : w - 1;                  // padding repeats the last column
        double px = __dadd_rn(g0, __dmul_rn(__dadd_rn((double)x, 0.5), g1));      // cn.c:222
        double dc = __ddiv_rn(__dsub_rn(px, s0), s1);                             // cn.c:223
        col_idx[i] = clamp_index(int_from_double_x86(c99_round(dc)), hsx);        // cn.c:225,228
    }
    int j = i - w_pad;
    if (j >= 0 && j < h) {
        double py = __dadd_rn(g3, __dmul_rn(__dadd_rn((double)j, 0.5), g5));      // cn.c:219
        double dr = __ddiv_rn(__dsub_rn(s3, py), fabs(s5));                       // cn.c:224
        row_idx[j] = clamp_index(int_from_double_x86(c99_round(dr)), hsy);        // cn.c:226,229
    }
}

// ---------------------------------------------------------------------------------------------
// small PTX helpers

__device__ __forceinline__ uint4 ldg_stream16(const void *p)
{
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void stg_stream16(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
#if GCN10_STORE_POLICY == 0
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#elif GCN10_STORE_POLICY == 1
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#elif GCN10_STORE_POLICY == 2
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#else
    asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#endif
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// soil code -> LUT slot, both drainage conditions (cn.c:88-111 + the sg<5 test of cn.c:123-124)
__device__ __forceinline__ uint32_t soil_slot(uint32_t hv, int drained)
{
    uint32_t d = hv - 11u;
    uint32_t s = (d <= 3u) ? (drained ? 4u : d + 1u) : hv;
    return s < (uint32_t)kSgInvalid ? s : (uint32_t)kSgInvalid;
}

// ---------------------------------------------------------------------------------------------
// The streaming kernel.  NP = planes per group (1..9), G = groups (1 or 2 drainage conditions).
//
// Shared memory: [0, 32 KB) LUT records, then the kBoxRows x kBoxCols HSG tile, then one mbarrier.

__host__ __device__ constexpr int smem_hsg_off(int np) { return lut_bytes_for(np); }
__host__ __device__ constexpr int smem_bar_off(int np) { return smem_hsg_off(np) + kBoxRows * kBoxCols; }
// store path: with GCN10_BULK_STORE=1 (default) results are staged per row in shared memory and written with
// cp.async.bulk (TMA engine, SASS UBLKCP: +3 % over per-thread STG.128, profiles/r01_kernel_sweeps.md);
// GCN10_BULK_STORE=0 keeps the direct-store kernel
#ifndef GCN10_BULK_STORE
#define GCN10_BULK_STORE 1
#endif
#ifndef GCN10_STORE_HINT
#define GCN10_STORE_HINT 1      // L2 evict-first cache hint on the bulk stores (+0.5 %: the planes are never re-read)
#endif
__host__ __device__ constexpr int smem_stage_off(int np) { return smem_bar_off(np) + 128; }
// Launches that write at most kNarrowPlanes planes keep the direct per-thread stores: with one to four 4 KB rows per
// CTA and row, the per-row CTA barrier of the staged path costs more than the TMA stores save (one plane:
// 0.93 ms staged vs 0.63 ms direct on B200), and without the stage eight CTAs fit an SM.
__host__ __device__ constexpr bool bulk_store_for(int np, int groups) { return GCN10_BULK_STORE && np * groups > kNarrowPlanes; }
// The direct-store narrow kernels are latency bound on the land-cover stream (ncu: long-scoreboard stalls with two
// rows prefetched into registers), so they keep GCN10_NARROW_DEPTH rows per thread in flight through a cp.async ring in
// shared memory instead: every thread copies, waits for and reads back only its own 16 bytes -- no CTA barrier.
__host__ __device__ constexpr bool deep_ring_for(int np, int groups) { return GCN10_NARROW_DEPTH > 0 && !bulk_store_for(np, groups) && np <= kNarrowPlanes; }
// np = planes per drainage condition, groups = conditions in the launch
constexpr int smem_bytes_for(int np, int groups)
{
    return bulk_store_for(np, groups) ? smem_stage_off(np) + 2 * np * groups * kStripPx
           : deep_ring_for(np, groups) ? smem_stage_off(np) + GCN10_NARROW_DEPTH * kStripPx
                                       : smem_bar_off(np) + 16;
}

template <int NP>
__device__ __forceinline__ void transpose_store_word(const uint4 (&r)[4], uint32_t (&ow)[NP][4], int j)
{
    const uint32_t a[3] = { r[0].x, r[0].y, r[0].z };
    const uint32_t b[3] = { r[1].x, r[1].y, r[1].z };
    const uint32_t c[3] = { r[2].x, r[2].y, r[2].z };
    const uint32_t d[3] = { r[3].x, r[3].y, r[3].z };
#pragma unroll
    for (int wd = 0; wd < (NP + 3) / 4; wd++) {
        // 4x4 byte transpose: rows = pixels a,b,c,d; columns = planes 4wd..4wd+3
        uint32_t t0 = __byte_perm(a[wd], b[wd], 0x5140);    // a0 b0 a1 b1
        uint32_t t1 = __byte_perm(c[wd], d[wd], 0x5140);    // c0 d0 c1 d1
        if (4 * wd + 0 < NP) ow[4 * wd + 0][j] = __byte_perm(t0, t1, 0x5410);
        if (4 * wd + 1 < NP) ow[4 * wd + 1][j] = __byte_perm(t0, t1, 0x7632);
        if (4 * wd + 2 < NP) {
            uint32_t t2 = __byte_perm(a[wd], b[wd], 0x7362);    // a2 b2 a3 b3
            uint32_t t3 = __byte_perm(c[wd], d[wd], 0x7362);
            ow[4 * wd + 2][j] = __byte_perm(t2, t3, 0x5410);
            if (4 * wd + 3 < NP) ow[4 * wd + 3][j] = __byte_perm(t2, t3, 0x7632);
        }
    }
}

// the same 4 x 4 byte transpose for narrow records: one word per pixel holds its (up to four) planes
template <int NP>
__device__ __forceinline__ void transpose_store_word4(const uint32_t (&r)[4], uint32_t (&ow)[NP][4], int j)
{
    const uint32_t t0 = __byte_perm(r[0], r[1], 0x5140), t1 = __byte_perm(r[2], r[3], 0x5140);
    ow[0][j] = __byte_perm(t0, t1, 0x5410);
    if (NP > 1) ow[NP > 1 ? 1 : 0][j] = __byte_perm(t0, t1, 0x7632);
    if (NP > 2) {
        const uint32_t t2 = __byte_perm(r[0], r[1], 0x7362), t3 = __byte_perm(r[2], r[3], 0x7362);
        ow[NP > 2 ? 2 : 0][j] = __byte_perm(t2, t3, 0x5410);
        if (NP > 3) ow[NP > 3 ? 3 : 0][j] = __byte_perm(t2, t3, 0x7632);
    }
}

#ifndef GCN10_MIN_CTAS
#define GCN10_MIN_CTAS 1
#endif
template <int NP, int G>
__global__ void __launch_bounds__(kThreads, (NP * G <= kNarrowPlanes ? GCN10_NARROW_CTAS : GCN10_MIN_CTAS))
cn_block_kernel(const __grid_constant__ BlockParams p, const __grid_constant__ CUtensorMap hsg_map)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool kNarrow = NP <= kNarrowPlanes;
    constexpr int kPF = (NP * G <= kNarrowPlanes) ? GCN10_NARROW_PF : kPrefetch;     // rows of land cover in flight
    constexpr bool kBulk = bulk_store_for(NP, G);
    constexpr bool kDeep = deep_ring_for(NP, G);
    constexpr int kDepth = kDeep ? GCN10_NARROW_DEPTH : 1;
    constexpr int kLutSize = lut_bytes_for(NP);
    constexpr int kSmemHsgOff = smem_hsg_off(NP), kSmemBarOff = smem_bar_off(NP), kSmemStageOff = smem_stage_off(NP);
    constexpr int kSlotShift = kNarrow ? 2 : 4;     // per-pixel slot byte = slot * record bytes
    const uint32_t s_hsg = smem_u32(smem + kSmemHsgOff);
    const uint32_t s_bar = smem_u32(smem + kSmemBarOff);

    const int tid = threadIdx.x;
    const int x_first = blockIdx.x * kStripPx;
    const int w16 = p.w & ~(kVecPx - 1);            // the right edge (< 16 px) is cn_bytes_kernel's
    // staged stores: every thread stays for the per-row barrier; threads right of the raster recompute column group 0
    const bool active = x_first + tid * kVecPx < w16;
    const int x0 = (kBulk && !active) ? x_first : x_first + tid * kVecPx;
    const int y_begin = blockIdx.y * p.rows_per_cta;
    const int y_end = min(p.h, y_begin + p.rows_per_cta);

    // HSG footprint of this CTA.  Both index maps are monotone (each fp64 step of cn.c:219-229
    // is monotone in x / y), so the end points bound the range.
    const int x_last = min(w16, x_first + kStripPx) - 1;
    const int ca = __ldg(p.col_idx + x_first), cb = __ldg(p.col_idx + x_last);
    const int ra = __ldg(p.row_idx + p.y_base + y_begin), rb = __ldg(p.row_idx + p.y_base + y_end - 1);
    // the box start is rounded down to a 16-element boundary: TMA moves 16-byte granules and
    // faults on an inner coordinate whose byte address is not 16-byte aligned
    const int ci_min = min(ca, cb) & ~15, cj_min = min(ra, rb);
    const bool staged = p.use_tma && (max(ca, cb) - ci_min < kBoxCols) && (max(ra, rb) - cj_min < kBoxRows);

    // One thread arms the mbarrier and launches the two asynchronous fills of shared memory: the 32 KB of
    // LUT records (1-D bulk copy, L2 resident after the first CTA) and the HSG box (2-D tensor load).
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(s_bar, kLutSize + (staged ? kBoxRows * kBoxCols : 0));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(smem)), "l"(p.lut), "r"(kLutSize), "r"(s_bar) : "memory");
        if (staged)
            tma_load_2d(s_hsg, &hsg_map, ci_min, cj_min, s_bar);
    }

    const uint32_t row_bytes = (uint32_t)(min(w16, x_first + kStripPx) - x_first);      // multiple of 16
    size_t row_off = (size_t)y_begin * p.out_pitch + x_first;

    // software pipeline: the land-cover vectors (and HSG row indices) of the next kPF rows are in
    // flight while the current row is looked up and stored; the first ones are issued before the wait
    // on the shared-memory fills
    const uint8_t *esa_ptr = p.esa + (size_t)y_begin * p.esa_pitch + x0;
    size_t out_off = (size_t)y_begin * p.out_pitch + x0;
    const int32_t *rowp = p.row_idx + p.y_base + y_begin;
    uint4 eq[kPF];
    int cq[kPF];
#pragma unroll
    for (int s = 0; s < kPF; s++) {
        const bool in = y_begin + s < y_end && (kBulk || active);
        if (!kDeep)
            eq[s] = in ? ldg_stream16(esa_ptr + (size_t)s * p.esa_pitch) : make_uint4(0, 0, 0, 0);
        cq[s] = in ? __ldg(rowp + s) : 0;
    }
    // deep ring: slot s of this thread at ring + s * 4 KB; one commit group per row, empty when the row does not exist
    const uint32_t ring = smem_u32(smem) + (uint32_t)smem_stage_off(NP) + (uint32_t)tid * kVecPx;
    uint32_t ring_slot = 0;
    if constexpr (kDeep) {
#pragma unroll
        for (int s = 0; s < kDepth; s++) {
            if (active && y_begin + s < y_end)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(ring + (uint32_t)s * kStripPx),
                             "l"(esa_ptr + (size_t)s * p.esa_pitch) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }

    // this thread's 16 HSG columns, relative to the staged box (one byte each): loaded once, not per HSG row
    uint32_t box_col[4] = { 0, 0, 0, 0 };
    if (kNarrow && staged && (kBulk || active)) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int4 c4 = __ldg(reinterpret_cast<const int4 *>(p.col_idx + x0) + j);
            box_col[j] = (uint32_t)(c4.x - ci_min) | ((uint32_t)(c4.y - ci_min) << 8) | ((uint32_t)(c4.z - ci_min) << 16) |
                         ((uint32_t)(c4.w - ci_min) << 24);
        }
    }

    __syncthreads();            // the mbarrier is initialised
    mbar_wait(s_bar, 0);

    if (!kBulk && !active)
        return;

    const uint32_t swz_mask = 0x07070707u;
    uint32_t slot[G][4];            // per pixel: (slot << 4) in one byte, 4 pixels per word
    // one-condition narrow launches keep, per pixel, the shared-memory address of its soil group's column of the
    // narrow table (16 registers): the lookup is then PRMT (class byte) + IMAD (class * 36 + column) + LDS.32
    constexpr bool kColumns = kNarrow && G == 1;
    uint32_t column[kColumns ? 16 : 1];
    const uint32_t s_lut = smem_u32(smem);
    int cj_cur = INT_MIN;           // row_idx is clamped to >= 0, so this never matches
#pragma unroll
    for (int g = 0; g < G; g++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            slot[g][j] = 0;

    for (int y = y_begin; y < y_end; y++) {
        uint4 e;
        if constexpr (kDeep) {
            // the oldest of the kDepth rows in flight has landed; read it, then reuse its slot for row y + kDepth
            asm volatile("cp.async.wait_group %0;" :: "n"(kDepth - 1) : "memory");
            const uint32_t at = ring + ring_slot * kStripPx;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "r"(at) : "memory");
            if (y + kDepth < y_end)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(at),
                             "l"(esa_ptr + (size_t)kDepth * p.esa_pitch) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            ring_slot = ring_slot + 1 == (uint32_t)kDepth ? 0u : ring_slot + 1;
        }
        else {
            e = eq[0];
        }
        const int cj = cq[0];
#pragma unroll
        for (int s = 0; s + 1 < kPF; s++) {
            if (!kDeep)
                eq[s] = eq[s + 1];
            cq[s] = cq[s + 1];
        }
        if (y + kPF < y_end) {
            if (!kDeep)
                eq[kPF - 1] = ldg_stream16(esa_ptr + (size_t)kPF * p.esa_pitch);
            cq[kPF - 1] = __ldg(rowp + kPF);
        }
        rowp++;

        if (cj != cj_cur) {
            // new HSG row: gather this thread's 16 soil codes (cn.c:230) and turn them into
            // LUT slots for each drainage condition (cn.c:88-111)
            cj_cur = cj;
            const uint8_t *row_g = p.hsg + (size_t)cj * p.hsg_pitch;
            const uint32_t row_s = s_hsg + (uint32_t)((cj - cj_min) * kBoxCols);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t acc[G];
#pragma unroll
                for (int g = 0; g < G; g++)
                    acc[g] = 0;
                if (kNarrow && staged) {
                    // the columns inside the staged box were packed into four registers before the row loop:
                    // no global load is left on this path
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint32_t hv;
                        asm("ld.shared.u8 %0, [%1];" : "=r"(hv) : "r"(row_s + ((box_col[j] >> (8 * q)) & 0xFFu)));
#pragma unroll
                        for (int g = 0; g < G; g++)
                            acc[g] |= (soil_slot(hv, p.group_drained[g]) << kSlotShift) << (8 * q);
                    }
                }
                else {
                    const int4 c4 = __ldg(reinterpret_cast<const int4 *>(p.col_idx + x0) + j);
                    const int cc[4] = { c4.x, c4.y, c4.z, c4.w };
                    const int row_off_s = kSmemHsgOff + (cj - cj_min) * kBoxCols - ci_min;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t hv = staged ? (uint32_t)smem[row_off_s + cc[q]] : (uint32_t)__ldg(row_g + cc[q]);
#pragma unroll
                        for (int g = 0; g < G; g++)
                            acc[g] |= (soil_slot(hv, p.group_drained[g]) << kSlotShift) << (8 * q);
                    }
                }
#pragma unroll
                for (int g = 0; g < G; g++)
                    slot[g][j] = acc[g];
                if constexpr (kColumns) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        column[4 * j + q] = s_lut + ((acc[0] >> (8 * q)) & 0xFFu);
                }
            }
        }

        const uint32_t ew[4] = { e.x, e.y, e.z, e.w };
        // bank swizzle term per pixel (16-byte records): ((lc >> swz_shift) & 7) << 4, byte-parallel
        uint32_t fz[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            fz[j] = kNarrow ? 0u : ((ew[j] >> p.swz_shift) & swz_mask) << 4;

#pragma unroll
        for (int g = 0; g < G; g++) {
            uint32_t ow[NP][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t sx = slot[g][j] ^ fz[j];
                if constexpr (kColumns) {
                    uint32_t r[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t lc = __byte_perm(ew[j], 0, 0x4440 | q);
                        asm("ld.shared.u32 %0, [%1];" : "=r"(r[q]) : "r"(lc * (4u * kLut4Stride) + column[4 * j + q]));
                    }
                    transpose_store_word4<NP>(r, ow, j);
                }
                else if constexpr (kNarrow) {
                    uint32_t r[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        // record offset = lc * 36 + slot * 4
                        const uint32_t lc = __byte_perm(ew[j], 0, 0x4440 | q);
                        const uint32_t sb = __byte_perm(sx, 0, 0x4440 | q);
                        r[q] = *reinterpret_cast<const uint32_t *>(smem + (lc * (4u * kLut4Stride) + sb));
                    }
                    transpose_store_word4<NP>(r, ow, j);
                }
                else {
                    uint4 r[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        // record offset = lc*128 + (slot ^ swizzle)*16
                        const uint32_t lc = __byte_perm(ew[j], 0, 0x4440 | q);
                        const uint32_t sb = __byte_perm(sx, 0, 0x4440 | q);
                        r[q] = *reinterpret_cast<const uint4 *>(smem + (lc * 128u + sb));
                    }
                    transpose_store_word<NP>(r, ow, j);
                }
            }
            if constexpr (kBulk) {
                uint8_t *stage = smem + kSmemStageOff + ((y - y_begin) & 1) * (NP * G * kStripPx) + g * NP * kStripPx;
#pragma unroll
                for (int k = 0; k < NP; k++)
                    *reinterpret_cast<uint4 *>(stage + k * kStripPx + tid * kVecPx) =
                        make_uint4(ow[k][0], ow[k][1], ow[k][2], ow[k][3]);
            }
            else {
#pragma unroll
                for (int k = 0; k < NP; k++)
                    stg_stream16(p.out[g * NP + k] + out_off, ow[k][0], ow[k][1], ow[k][2], ow[k][3]);
            }
        }

        if constexpr (kBulk) {
        // generic-proxy writes -> visible to the async proxy; the row before last has left its stage
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid == 0)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t stage = smem_u32(smem + kSmemStageOff + ((y - y_begin) & 1) * (NP * G * kStripPx));
#if GCN10_STORE_HINT
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
            for (int k = 0; k < NP * G; k++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                             :: "l"(p.out[k] + row_off), "r"(stage + k * kStripPx), "r"(row_bytes), "l"(pol) : "memory");
#else
#pragma unroll
            for (int k = 0; k < NP * G; k++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(p.out[k] + row_off), "r"(stage + k * kStripPx), "r"(row_bytes) : "memory");
#endif
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        row_off += p.out_pitch;
        }
        esa_ptr += p.esa_pitch;
        out_off += p.out_pitch;
    }
    if (kBulk && tid == 0)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all bulk stores of this CTA have completed
}

// ---------------------------------------------------------------------------------------------
// Byte-wise kernel for columns [x_begin, w) of every row: the right edge the vector kernel leaves
// (w % 16 pixels) or, with x_begin = 0, whole blocks whose buffers are not 16-byte aligned.
// Planes are addressed through the same compacted out[] list; NP and G are runtime values here.

__global__ void cn_bytes_kernel(const __grid_constant__ BlockParams p, int x_begin, int np, int groups)
{
    const int cols = p.w - x_begin;
    const long long n = (long long)cols * p.h;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / cols);
        const int x = x_begin + (int)(i - (long long)y * cols);
        const uint32_t lc = p.esa[(size_t)y * p.esa_pitch + x];
        const int ci = __ldg(p.col_idx + x);
        const int cj = __ldg(p.row_idx + p.y_base + y);
        const uint32_t hv = __ldg(p.hsg + (size_t)cj * p.hsg_pitch + ci);
        for (int g = 0; g < groups; g++) {
            const uint32_t sg = soil_slot(hv, p.group_drained[g]);
            const uint8_t *rec = p.rec_bytes == 16
                                     ? reinterpret_cast<const uint8_t *>(p.lut + (lc * 8u + (sg ^ ((lc >> p.swz_shift) & 7u))))
                                     : reinterpret_cast<const uint8_t *>(p.lut) + 4u * (lc * kLut4Stride + sg);
            for (int k = 0; k < np; k++)
                p.out[g * np + k][(size_t)y * p.out_pitch + x] = __ldg(rec + k);
        }
    }
}

}  // namespace gcn10
