// K1 (fast path) — fused `into_luma8` + `adaptive_threshold(radius 7)` over a batch of frames, sm_100a.
//
// Replaces, bit for bit, /root/reference/src/aruco.rs:60-61 (image 0.25 `into_luma8`, imageproc 0.25
// `adaptive_threshold`; semantics in SURVEY.md A.1-A.2).  Same results as the generic kernel in k1_threshold.cu, which
// stays the path for unaligned inputs and other radii.
//
// Shape: the kernel was issue-bound, not HBM-bound (ncu, profiles/r01a_*: alu pipe 71 %, dram 20 %), so this version
// is organised around instructions per pixel and has no block-wide barrier at all:
//   * a WARP owns a 256-column strip (240 output columns + an 8-column halo on each side) of one row segment of one
//     frame and marches down it; lane l owns 8 adjacent columns.  Warps never talk to each other.
//   * RGB rows arrive through TMA: one `cp.async.bulk.tensor.3d` box of {256 px, kRows rows, 1 frame} per step into a
//     per-warp ring of stages guarded by mbarriers (SASS UTMALDG).  The tensor map is over the byte plane viewed as
//     u32 [n][h][pitch/4]; out-of-image rows and columns are zero-filled by the TMA unit, which is exactly the
//     clipped-window semantics of the reference (a clipped window sums only in-image pixels), so the inner loop has
//     no bounds checks on its loads.
//   * luma: v = 2126 R + 7152 G + 722 B by two dp4a (byte-split weights), grey = umulhi(v, ceil(2^40/10^4)) >> 8
//     (exact for every v <= 2 550 000, tests/test_k1_identities.py); no per-channel byte extraction.
//   * vertical 15-row column sums live in registers as u16 pairs (one IADD3 per pair per row: + new - old); the 15
//     previous grey rows of the lane's 8 columns are a lane-private shared-memory ring (8 B per lane per row).
//   * horizontal 15-column sums: the 8 neighbouring column sums on each side come from lanes l-1 / l+1 by warp
//     shuffles; pair sums by dp2a; a sliding 7-pair window T; S(even) = T + hi(pair before), S(odd) = T + lo(pair after).
//   * `pix >= floor(S / cnt)`  <=>  `S < (pix + 1) * cnt`  <=>  sign of  S - 256 cnt + (255 - pix) * cnt, evaluated as
//     dp4a(~pix4, cnt << 8j, dp2a(neighbour pair, T - 256 cnt)): one dp4a picks the pixel's byte and multiplies it,
//     the sign bits are collected with funnel shifts.  cnt = nx * ny is the clipped window area (<= 225, one byte).
//   * outputs: grey 8 B / lane / row, optional byte mask 8 B / lane / row, optional 1-bit mask 1 B / lane / row.
// Algorithmic traffic: 3 B read + 1 B grey + 1 B mask = 5 B / pixel; with the 1-bit mask the kernel
// moves 3 + 1 + 1/8 B / pixel (+ the row / column halos, served mostly by L2).
#include <cuda.h>

#include <type_traits>

#include "a3_internal.h"

namespace a3 {
namespace {

constexpr int kWarpsPerCta = 4;
#ifndef A3_K1_MIN_CTAS
#define A3_K1_MIN_CTAS 7
#endif
constexpr int kMinCtasPerSm = A3_K1_MIN_CTAS;   // register cap: 65536 / (128 * 7) -> 72 per thread, matches the 7 CTAs the shared memory allows
constexpr int kRing = 16;          // >= 2 * 7 + 1 grey rows; a power of two so ring slots are `row & 15`
constexpr int kCore = 240;         // output columns per warp
constexpr int kHalo = 8;           // >= radius, keeps every lane's 8 columns 8-px aligned
constexpr uint32_t kMagic = 109951163u;  // ceil(2^40 / 10000)

struct StripArgs {
    uint8_t *grey, *mask, *bits;   // bits addressed by byte: byte (x >> 3) of row, bit x & 7
    uint32_t n, w, h;
    uint32_t nstrips, nsegs, seg_rows, njobs;
    uint32_t bits_row_bytes;       // bytes between mask rows (4 * ceil(w / 32) when tightly packed)
    uint32_t bits_col_bytes;       // bytes between 32-pixel word columns (4 when row-major)
    size_t bits_frame_bytes;       // bytes between frames of the 1-bit mask
    uint32_t stages;               // TMA ring depth per warp (kStages)
    int wide_stores;               // w % 8 == 0 and grey / mask bases 8-byte aligned
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "K1S_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra K1S_DONE;\n"
        "bra K1S_WAIT;\n"
        "K1S_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA: box {box_x u32, kRows, 1} of the [n][h][pitch/4] u32 view at element coordinates (cx, cy, cz); out-of-range
// elements are written as zero.
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *map, int cx, int cy, int cz, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(bar))
        : "memory");
}

// The TMA unit wants the first byte of a box 16-byte aligned in global memory, so a box starts at the 16-px boundary
// at or before the warp's first column (x0 = 240 s - 8  ->  240 s - 16) and is 16 px wider than the 256 columns when
// 8 px are not 16 bytes (RGB8, LUMA8).
template <int FMT>
struct Fmt {
    static constexpr int bpp = fmt_bpp(FMT);
    static constexpr int lead_px = bpp == 4 ? 0 : 8;                      // columns staged before x0
    static constexpr int lead_bytes = lead_px * bpp;                      // 24 / 0 / 8: keeps 8-byte alignment of lane loads
    static constexpr int row_bytes = (256 + 2 * lead_px) * bpp;           // 816 / 1024 / 272 (multiples of 16)
    static constexpr int box_x = row_bytes / 4;                           // box width in u32 elements (<= 256)
};

// Luma by two dp2a (u16 weights x pixel bytes): v = 2126 R + 7152 G + 722 B needs no byte-split weights, no shift between
// the two dot products and, for 3-byte pixels, no byte permutes either: a pixel that straddles two words takes one dp2a from
// each.  A weight pair names the bytes (lo: bytes 0,1; hi: bytes 2,3 of the word) it multiplies.
//   floor(v / 10000) << 8 | fraction byte = umulhi(v, ceil(2^40 / 10^4))   (exact for every v <= 2 550 000)
template <bool BGR>
struct W {
    static constexpr uint32_t r = BGR ? 722u : 2126u, g = 7152u, b = BGR ? 2126u : 722u;  // weights of pixel bytes 0, 1, 2
    static constexpr uint32_t w01 = r | (g << 16);   // bytes (0,1) or (2,3) of a word = pixel bytes 0,1
    static constexpr uint32_t w2_ = b;               // ... = pixel byte 2, then a byte of another pixel (weight 0)
    static constexpr uint32_t w_0 = r << 16;         // ... = a byte of another pixel, then pixel byte 0
    static constexpr uint32_t w12 = g | (b << 16);   // ... = pixel bytes 1,2
};
// grey of the pixel in bytes 0..2 of `px` (byte 3 ignored), left in byte 1 of the result (bytes 2 and 3 are zero)
template <bool BGR>
__device__ __forceinline__ uint32_t luma_h(uint32_t px) {
    return __umulhi(__dp2a_hi(W<BGR>::w2_, px, __dp2a_lo(W<BGR>::w01, px, 0u)), kMagic);
}

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// the lane's 8 greys of one staged row (shared address `row` = first byte of the lane's pixels):
// p01..p67 = u16 pairs (g_even | g_odd << 16), g = packed bytes
template <int FMT>
__device__ __forceinline__ void load_grey8(uint32_t row, uint32_t &p01, uint32_t &p23, uint32_t &p45, uint32_t &p67, uint2 &g) {
    if constexpr (FMT == A3_FMT_LUMA8) {
        g = lds64(row);
        p01 = __byte_perm(g.x, 0u, 0x4140); p23 = __byte_perm(g.x, 0u, 0x4342);
        p45 = __byte_perm(g.y, 0u, 0x4140); p67 = __byte_perm(g.y, 0u, 0x4342);
    } else {
        uint32_t h[8];
        constexpr bool BGR = fmt_bgr(FMT);
        if constexpr (fmt_bpp(FMT) == 3) {
            const uint2 a = lds64(row), b = lds64(row + 8), c = lds64(row + 16);
            // 12 bytes = 4 pixels: [p0 p0 p0 p1] [p1 p1 p2 p2] [p2 p3 p3 p3]
            using K = W<BGR>;
            auto quad = [](uint32_t x, uint32_t y, uint32_t z, uint32_t *h) {
                h[0] = __umulhi(__dp2a_hi(K::w2_, x, __dp2a_lo(K::w01, x, 0u)), kMagic);
                h[1] = __umulhi(__dp2a_lo(K::w12, y, __dp2a_hi(K::w_0, x, 0u)), kMagic);
                h[2] = __umulhi(__dp2a_lo(K::w2_, z, __dp2a_hi(K::w01, y, 0u)), kMagic);
                h[3] = __umulhi(__dp2a_hi(K::w12, z, __dp2a_lo(K::w_0, z, 0u)), kMagic);
            };
            quad(a.x, a.y, b.x, h);
            quad(b.y, c.x, c.y, h + 4);
        } else {
            const uint4 a = lds128(row), b = lds128(row + 16);
            h[0] = luma_h<BGR>(a.x); h[1] = luma_h<BGR>(a.y); h[2] = luma_h<BGR>(a.z); h[3] = luma_h<BGR>(a.w);
            h[4] = luma_h<BGR>(b.x); h[5] = luma_h<BGR>(b.y); h[6] = luma_h<BGR>(b.z); h[7] = luma_h<BGR>(b.w);
        }
        p01 = __byte_perm(h[0], h[1], 0x6521); p23 = __byte_perm(h[2], h[3], 0x6521);
        p45 = __byte_perm(h[4], h[5], 0x6521); p67 = __byte_perm(h[6], h[7], 0x6521);
        g.x = __byte_perm(p01, p23, 0x6420); g.y = __byte_perm(p45, p67, 0x6420);
    }
}

// 4 mask bits -> 4 bytes of 0 / 255
__device__ __forceinline__ uint32_t expand4(uint32_t nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xffu; }

// Per-lane marching state (registers).  Everything the steady-state rows do not need lives elsewhere: the kernel runs at the
// 72-register cap of 7 CTAs per SM, and every value kept live across the loop was a value rematerialised inside it.
struct Lane {
    uint32_t cs0, cs1, cs2, cs3;   // running 15-row column sums of the lane's 8 columns, u16 pairs
    uint32_t tab;                  // shared address of the window-area constants of the lane's class (see set_ny)
    uint32_t valid8;               // which of the lane's 8 pixels are output pixels
    int store_mode;                // 0 nothing, 1 one 8-byte store per array, 2 4-byte stores
    uint8_t *grey, *mask, *bits;   // output addresses of the lane's pixels in the next output row
};

// Warp-uniform marching context (registers): shared addresses and the loop bounds.
struct March {
    uint32_t base;       // shared address of the warp's carve: stages, then the grey ring, then the constants table, the mbarriers, the scratch
    uint32_t ring;       // this lane's slot 0 of the grey ring (slot s at ring + 256 s)
    uint32_t lane_src;   // shared address of the lane's pixels in row 0 of stage 0
    int nboxes;          // boxes of 2 input rows
    int fast_lo, fast_hi;  // boxes fast_lo .. fast_hi: both rows exist, emit output and have the full 15 window rows
};

// The march consumes TMA boxes of 2 input rows in a ROLLED loop over a 2-stage ring (an earlier version unrolled 8 rows with
// compile-time ring slots, 8600 instructions per kernel, and ncu showed it waiting for the instruction cache:
// stall_no_instruction 2.4 cycles per issue, icc hit rate 77 %).
constexpr int kBoxRows = 2, kStages = 2;
// Window areas cnt = nx * ny (clipped window width x height).  ny is warp-uniform and changes only in the top / bottom 7 rows
// of the frame; nx differs from 15 only for the few lanes within 7 columns of the left / right image edge.  The per-pixel
// constants (-256 cnt, cnt << 8 (j & 3)) therefore live in a small per-warp shared table of kClasses lane classes - class 0:
// nx = 15 everywhere, classes 1..3: one clipped lane each - instead of 16 registers per lane (which spilled at the 72-register
// cap).  x-interior warps do not read it at all in the steady state (template INT: cnt = 225 is an immediate).
constexpr int kClasses = 4;
constexpr int kTabBytes = kClasses * 64;
// per-warp scratch behind the mbarriers: what only lane 0's TMA requests and the few border rows need
constexpr int kScrX0 = 0, kScrYs = 4, kScrCx = 8, kScrFrame = 12, kScrNy = 16, kScrRows = 20, kScrBytes = 32;
template <int FMT>
struct Stage {
    static constexpr int tx_bytes = kBoxRows * Fmt<FMT>::row_bytes;  // bytes one box delivers
    static constexpr int bytes = (tx_bytes + 127) & ~127;            // stage stride: TMA destinations are 128-byte aligned
    static constexpr int ring_off = kStages * bytes;
    static constexpr int tab_off = ring_off + kRing * 256;
    static constexpr int bar_off = tab_off + kTabBytes;              // mbarrier of stage s at bar_off + 8 s
    static constexpr int scr_off = bar_off + 8 * kStages;
    static constexpr int per_warp = (scr_off + kScrBytes + 127) & ~127;
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ uint32_t clipped_nx(int w, int xx) {
    return (xx >= 0 && xx < w) ? (uint32_t)(min(w - 1, xx + 7) - max(0, xx - 7) + 1) : 0u;
}
// (re)write the constants table for output rows with ny window rows: entry j of a class = {-256 cnt_j, cnt_j << 8 (j & 3)}.
// Lanes clipped by an image edge write their own class, the first unclipped lane writes class 0 (nx = 15).
template <int FMT>
__device__ __noinline__ void set_ny(uint32_t base, uint32_t tab, uint32_t ny, int w) {
    const int lane = threadIdx.x & 31;
    const uint32_t scr = base + Stage<FMT>::scr_off;
    __syncwarp();  // every lane has finished reading the previous values
    const int x = (int)lds32(scr + kScrX0) + 8 * lane;
    const uint32_t cls = (tab - (base + Stage<FMT>::tab_off)) >> 6;
    const uint32_t first0 = (uint32_t)__ffs(__ballot_sync(0xffffffffu, cls == 0u)) - 1u;
    if (cls != 0u || (uint32_t)lane == first0) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t nx = cls == 0u ? 15u : clipped_nx(w, x + j);
            const uint32_t c = nx * ny;
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(tab + 8 * j), "r"(0u - 256u * c), "r"(c << (8 * (j & 3))) : "memory");
        }
    }
    if (lane == 0) sts32(scr + kScrNy, ny);
    __syncwarp();
}

// Measured and left off: boxes further ahead than the two stages hold can be asked into L2 (cp.async.bulk.prefetch.tensor), so
// that the stage's own request finds its rows there — ncu's samples have 24 % of the kernel's stalls on the two mbarrier waits, and a
// third stage would cost the 7th CTA of an SM.  256 x 1080p: 0.421 ms without, 0.435 ms with a prefetch 2 boxes ahead, 0.442 ms 4
// boxes ahead: lane 0's extra request per box costs more issue slots than the shorter waits give back.
#ifndef A3_K1_PREFETCH
#define A3_K1_PREFETCH 0
#endif
constexpr int kPrefetchAhead = A3_K1_PREFETCH;  // boxes beyond the one being requested; 0 = off
template <int FMT>
__device__ __forceinline__ void arm_box(const CUtensorMap *tmap, uint32_t base, int box, int nboxes) {  // lane 0: request rows ys - 7 + 2 box .. into stage box & 1
    const uint32_t st = (uint32_t)box & 1u, bar = base + Stage<FMT>::bar_off + 8 * st, scr = base + Stage<FMT>::scr_off;
    const int cx = (int)lds32(scr + kScrCx), cy = (int)lds32(scr + kScrYs) - 7 + box * kBoxRows, cz = (int)lds32(scr + kScrFrame);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(Stage<FMT>::tx_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            base + st * Stage<FMT>::bytes),
        "l"(tmap), "r"(cx), "r"(cy), "r"(cz), "r"(bar)
        : "memory");
    if (kPrefetchAhead > 0 && box + kPrefetchAhead < nboxes)
        asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(cx), "r"(cy + kPrefetchAhead * kBoxRows), "r"(cz)
                     : "memory");
}
__device__ __forceinline__ void wait_box(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "K1S_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra K1S_DONE;\n"
        "bra K1S_WAIT;\n"
        "K1S_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Consume one staged input row: update the column sums and, when OUT, emit the output row 7 rows behind it.
// src: the lane's pixels of the row; ring_new / ring_old / ring_pix: the lane's ring slots of this row, of the row
// 15 behind (leaving the window) and of the row 7 behind (the output row).
// INT (only with OUT, steady-state rows of x-interior warps): every pixel of the warp has the full 15 x 15 window, so cnt = 225
// is a compile-time constant: -256 cnt rides in on the accumulator of the pair sums and cnt << 8 (j & 3) is an immediate.
template <int FMT, bool MASK, bool BITS, bool OUT, bool INT = false>
__device__ __forceinline__ void row_step(Lane &L, const StripArgs &a, uint32_t src, uint32_t ring_new, uint32_t ring_old, uint32_t ring_pix) {
    uint32_t p01, p23, p45, p67;
    uint2 g;
    load_grey8<FMT>(src, p01, p23, p45, p67, g);
    const uint2 old = lds64(ring_old);
    sts64(ring_new, g);
    L.cs0 = L.cs0 + p01 - __byte_perm(old.x, 0u, 0x4140);
    L.cs1 = L.cs1 + p23 - __byte_perm(old.x, 0u, 0x4342);
    L.cs2 = L.cs2 + p45 - __byte_perm(old.y, 0u, 0x4140);
    L.cs3 = L.cs3 + p67 - __byte_perm(old.y, 0u, 0x4342);
    if constexpr (OUT) {
        const uint2 pix = lds64(ring_pix);
        // pair words of columns -8..15 relative to the lane's first column
        uint32_t w[12];
        w[0] = __shfl_up_sync(0xffffffffu, L.cs0, 1); w[1] = __shfl_up_sync(0xffffffffu, L.cs1, 1);
        w[2] = __shfl_up_sync(0xffffffffu, L.cs2, 1); w[3] = __shfl_up_sync(0xffffffffu, L.cs3, 1);
        w[4] = L.cs0; w[5] = L.cs1; w[6] = L.cs2; w[7] = L.cs3;
        w[8] = __shfl_down_sync(0xffffffffu, L.cs0, 1); w[9] = __shfl_down_sync(0xffffffffu, L.cs1, 1);
        w[10] = __shfl_down_sync(0xffffffffu, L.cs2, 1); w[11] = __shfl_down_sync(0xffffffffu, L.cs3, 1);
        // sliding sums of 7 pair words, still packed (even columns | odd columns << 16; a half is at most 7 * 15 * 255 < 2^16),
        // then T[i] = both halves of window i: the 14 columns that all of pixel 2i's and pixel 2i+1's windows share
        uint32_t A[4], T[4];
        A[0] = (w[1] + w[2] + w[3]) + (w[4] + w[5] + w[6]) + w[7];
        A[1] = A[0] + w[8] - w[1];
        A[2] = A[1] + w[9] - w[2];
        A[3] = A[2] + w[10] - w[3];
#pragma unroll
        for (int i = 0; i < 4; i++) T[i] = __dp2a_lo(A[i], 0x0101u, INT ? 0u - 256u * 225u : 0u);
        const uint32_t pc0 = ~pix.x, pc1 = ~pix.y;
        uint32_t bits8 = 0;
#pragma unroll
        for (int jj = 3; jj >= 0; jj--) {
            uint4 c = make_uint4(0u, 0u, 0u, 0u);   // {-256 cnt, cnt << 8 (j & 3)} of pixels 2 jj and 2 jj + 1
            if constexpr (!INT) c = lds128(L.tab + 16 * jj);
#pragma unroll
            for (int j = 2 * jj + 1; j >= 2 * jj; j--) {
                // S - 256 cnt:  even column j: T[j/2] + hi(w[j/2]);  odd: T[j/2] + lo(w[(j+15)/2])
                const uint32_t t = INT ? T[jj] : T[jj] + ((j & 1) ? c.z : c.x);
                const uint32_t acc = (j & 1) ? __dp2a_lo(w[(j + 15) / 2], 0x0001u, t) : __dp2a_lo(w[j / 2], 0x0100u, t);
                const uint32_t u = __dp4a(j < 4 ? pc0 : pc1, INT ? 225u << (8 * (j & 3)) : ((j & 1) ? c.w : c.y), acc);  // + (255 - pix) * cnt
                bits8 = __funnelshift_l(u, bits8, 1);                          // sign bit: S < (pix + 1) * cnt
            }
        }
        if constexpr (INT) {
            // an x-interior warp: its core lanes have all 8 pixels and take the wide stores, its halo lanes store nothing
            if (L.store_mode) {
                *reinterpret_cast<uint2 *>(L.grey) = pix;
                if constexpr (MASK) *reinterpret_cast<uint2 *>(L.mask) = make_uint2(expand4(bits8 & 15u), expand4(bits8 >> 4));
                if constexpr (BITS) *L.bits = (uint8_t)bits8;
            }
        } else {
            bits8 &= L.valid8;
            if (L.store_mode == 1) {
                *reinterpret_cast<uint2 *>(L.grey) = pix;
                if constexpr (MASK) *reinterpret_cast<uint2 *>(L.mask) = make_uint2(expand4(bits8 & 15u), expand4(bits8 >> 4));
                if constexpr (BITS) *L.bits = (uint8_t)bits8;
            } else if (L.store_mode == 2) {
                if (L.valid8 & 0x0fu) {
                    *reinterpret_cast<uint32_t *>(L.grey) = pix.x;
                    if constexpr (MASK) *reinterpret_cast<uint32_t *>(L.mask) = expand4(bits8 & 15u);
                }
                if (L.valid8 & 0xf0u) {
                    *reinterpret_cast<uint32_t *>(L.grey + 4) = pix.y;
                    if constexpr (MASK) *reinterpret_cast<uint32_t *>(L.mask + 4) = expand4(bits8 >> 4);
                }
                if constexpr (BITS) *L.bits = (uint8_t)bits8;
            }
        }
        L.grey += a.w;
        if constexpr (MASK) L.mask += a.w;
        if constexpr (BITS) L.bits += a.bits_row_bytes;
    }
}

template <int FMT, bool MASK, bool BITS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kMinCtasPerSm) k1_strips_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StripArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t job = blockIdx.x * kWarpsPerCta + warp;
    if (job >= a.njobs) return;  // whole warp; warps are independent (no block-wide barrier anywhere)
    const uint32_t strip = job % a.nstrips;
    const uint32_t seg = (job / a.nstrips) % a.nsegs;
    const uint32_t frame = job / (a.nstrips * a.nsegs);
    const int x0 = (int)strip * kCore - kHalo;        // first column of the warp's 256
    const int x = x0 + 8 * lane;                      // first of this lane's 8 columns
    const int ys = (int)(seg * a.seg_rows);
    const int ye = min((int)a.h, ys + (int)a.seg_rows);
    const int total_rows = (ye - ys) + 14;            // input rows ys-7 .. ye+6

    March m;
    m.base = smem_u32(smem) + (uint32_t)warp * Stage<FMT>::per_warp;
    m.ring = m.base + Stage<FMT>::ring_off + 8 * lane;
    m.lane_src = m.base + Fmt<FMT>::lead_bytes + lane * 8 * Fmt<FMT>::bpp;
    m.nboxes = (total_rows + kBoxRows - 1) / kBoxRows;
    // box b holds input rows k = 2b, 2b + 1, whose output rows are yo = ys + k - 14 and yo + 1.  Fast: k >= 14, k + 1 < total_rows,
    // yo >= 7 and yo + 1 <= h - 8 (window rows all inside the frame)
    {
        const int k_lo = max(14, 21 - ys), k_hi = min(total_rows - 2, (int)a.h - 9 - ys + 14);
        m.fast_lo = (k_lo + 1) >> 1;
        m.fast_hi = k_hi >= 0 ? k_hi >> 1 : -1;
    }
    const uint32_t scr = m.base + Stage<FMT>::scr_off, bar = m.base + Stage<FMT>::bar_off;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sts32(scr + kScrX0, (uint32_t)x0);
        sts32(scr + kScrYs, (uint32_t)ys);
        sts32(scr + kScrCx, (uint32_t)((x0 - Fmt<FMT>::lead_px) * Fmt<FMT>::bpp / 4));  // u32 element coordinate of the box, a multiple of 4
        sts32(scr + kScrFrame, frame);
        sts32(scr + kScrNy, 0u);                      // window rows the constants table is written for
        sts32(scr + kScrRows, (uint32_t)total_rows);
    }
#pragma unroll
    for (int s = 0; s < kRing; s++) sts64(m.ring + 256 * s, make_uint2(0u, 0u));
    __syncwarp();
    if (lane == 0) {
        arm_box<FMT>(&tmap, m.base, 0, m.nboxes);
        if (m.nboxes > 1) arm_box<FMT>(&tmap, m.base, 1, m.nboxes);
    }

    Lane L;
    L.cs0 = L.cs1 = L.cs2 = L.cs3 = 0;
    const bool lane_core = lane >= 1 && lane <= 30 && x < (int)a.w;
    L.valid8 = 0;
    bool clipped = false;  // an output pixel of this lane has a window narrower than 15 columns
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (lane_core && x + j < (int)a.w) {
            L.valid8 |= 1u << j;
            clipped |= clipped_nx((int)a.w, x + j) != 15u;
        }
    // lane classes of the constants table: a warp has at most kClasses - 1 clipped lanes (the lane at x = 0 and the one or two
    // lanes that hold the columns w - 7 .. w - 1), each gets its own class
    const uint32_t clipped_lanes = __ballot_sync(0xffffffffu, clipped);
    const uint32_t cls = clipped ? 1u + (uint32_t)__popc(clipped_lanes & ((1u << lane) - 1u)) : 0u;
    L.tab = m.base + Stage<FMT>::tab_off + 64u * cls;
    L.store_mode = !lane_core ? 0 : ((a.wide_stores && L.valid8 == 0xffu) ? 1 : 2);
    const size_t o_px = ((size_t)frame * a.h + ys) * a.w + x;
    L.grey = a.grey + o_px;
    L.mask = a.mask + o_px;
    L.bits = a.bits + (size_t)frame * a.bits_frame_bytes + (size_t)ys * a.bits_row_bytes + (size_t)(x >> 5) * a.bits_col_bytes + ((x >> 3) & 3);

    // x-interior warp: none of its core columns is within 7 px of the left or right image edge (6 of the 8 strips of a 1080p
    // row, 14 of 16 at 4K) and its core lanes take the 8-byte stores; warp-uniform
    const bool interior = strip > 0 && (uint32_t)kCore * strip + kCore + 7 <= a.w && a.wide_stores;
    // One box (2 input rows, first row k = 2 box, stage ST) with the ring slots of its rows: new rows at n0 / n1, the rows
    // leaving the window at o0 / o1, the output rows at p0 / p1.
    auto do_box = [&](auto st_tag, int box, uint32_t parity, uint32_t n0, uint32_t n1, uint32_t o0, uint32_t o1, uint32_t p0, uint32_t p1) {
        constexpr uint32_t ST = decltype(st_tag)::value;
        wait_box(m.base + Stage<FMT>::bar_off + 8 * ST, parity);
        const uint32_t src = m.lane_src + ST * Stage<FMT>::bytes;
        const bool fast = box >= m.fast_lo && box <= m.fast_hi;
        if (fast && interior) {
            row_step<FMT, MASK, BITS, true, true>(L, a, src, n0, o0, p0);
            row_step<FMT, MASK, BITS, true, true>(L, a, src + Fmt<FMT>::row_bytes, n1, o1, p1);
        } else if (fast) {
            if (lds32(m.base + Stage<FMT>::scr_off + kScrNy) != 15u) set_ny<FMT>(m.base, L.tab, 15u, (int)a.w);
            row_step<FMT, MASK, BITS, true>(L, a, src, n0, o0, p0);
            row_step<FMT, MASK, BITS, true>(L, a, src + Fmt<FMT>::row_bytes, n1, o1, p1);
        } else {
            const uint32_t scr = m.base + Stage<FMT>::scr_off;
            const int k = box * kBoxRows, rows = (int)lds32(scr + kScrRows), yo = (int)lds32(scr + kScrYs) + k - 14;
#pragma unroll 1
            for (int r = 0; r < kBoxRows; r++) {
                if (k + r >= rows) break;
                const uint32_t rn = r ? n1 : n0, ro = r ? o1 : o0, rp = r ? p1 : p0;
                if (k + r >= 14) {
                    const int y = yo + r;
                    const uint32_t ny = (uint32_t)(min((int)a.h - 1, y + 7) - max(0, y - 7) + 1);
                    if (ny != lds32(scr + kScrNy)) set_ny<FMT>(m.base, L.tab, ny, (int)a.w);  // only in the top / bottom 7 rows of the frame
                    row_step<FMT, MASK, BITS, true>(L, a, src + r * Fmt<FMT>::row_bytes, rn, ro, rp);
                } else {
                    row_step<FMT, MASK, BITS, false>(L, a, src + r * Fmt<FMT>::row_bytes, rn, ro, rp);
                }
            }
        }
        // every lane has consumed the box (its values are in registers or in the ring): refill the stage with the box 2 ahead
        __syncwarp();
        if (lane == 0 && box + kStages < m.nboxes) arm_box<FMT>(&tmap, m.base, box + kStages, m.nboxes);
    };
    // Two boxes (4 input rows) per iteration, so the stage of a box and its mbarrier are compile-time constants.  Ring slots:
    // row r lives in slot r & 15, 256 bytes apart; with k = 4 it the rows k .. k + 3 are at A .. A + 768, the rows leaving the
    // window (k - 15 .. k - 12) at A + 256 .. A + 768 and B, the output rows (k - 7 .. k - 4) at C + 256 .. C + 768 and D.
#pragma unroll 1
    for (int it = 0; 2 * it < m.nboxes; it++) {
        const uint32_t oa = ((uint32_t)it & 3u) << 10, ob = ((uint32_t)(it + 1) & 3u) << 10;
        const uint32_t A = m.ring + oa, B = m.ring + ob, C = m.ring + (oa ^ 0x800u), D = m.ring + (ob ^ 0x800u);
        const uint32_t parity = (uint32_t)it & 1u;              // each stage completes once per iteration
        do_box(std::integral_constant<uint32_t, 0>{}, 2 * it, parity, A, A + 256, A + 256, A + 512, C + 256, C + 512);
        if (2 * it + 1 < m.nboxes)
            do_box(std::integral_constant<uint32_t, 1>{}, 2 * it + 1, parity, A + 512, A + 768, A + 768, B, C + 768, D);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <int FMT>
cudaError_t launch_strips(const CUtensorMap &map, const StripArgs &a, bool mask, bool bits, uint32_t grid, size_t smem, cudaStream_t stream) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kWarpsPerCta * 32, smem, stream>>>(map, a);
        return cudaGetLastError();
    };
    if (mask) return bits ? go(k1_strips_kernel<FMT, true, true>) : go(k1_strips_kernel<FMT, true, false>);
    return bits ? go(k1_strips_kernel<FMT, false, true>) : go(k1_strips_kernel<FMT, false, false>);
}

}  // namespace

bool k1_strips_eligible(const K1Params &p) {
    const uint32_t bpp = fmt_bpp(p.format);
    return p.radius == 7 && p.grey != nullptr && p.w % 4 == 0 && p.w >= 4 && p.pitch % 16 == 0 && p.frame_stride % 16 == 0 &&
           p.frame_stride % p.pitch == 0 && (uintptr_t)p.src % 16 == 0 && (uintptr_t)p.grey % 4 == 0 && (uintptr_t)p.mask % 4 == 0 &&
           (uint64_t)p.w * bpp <= p.pitch && p.n <= 0x7fffffffu && encode_tiled_fn() != nullptr;
}

cudaError_t k1_strips(const K1Params &p, const K1Tuning *tuning, cudaStream_t stream, K1LaunchInfo *info) {
    const uint32_t bpp = fmt_bpp(p.format);

    // ---- tensor map over the frames viewed as u32 [n][h][pitch / 4]; dim 0 stops at the last pixel's bytes ----
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.w * bpp / 4, p.h, p.n};
    const cuuint64_t gstride[2] = {p.pitch, p.frame_stride};
    const uint32_t row_bytes = (256u + (bpp == 4 ? 0u : 16u)) * bpp;  // Fmt<>::row_bytes
    const cuuint32_t box[3] = {row_bytes / 4, (cuuint32_t)kBoxRows, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    const CUresult cr = encode_tiled_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t *>(p.src), gdim, gstride, box, estride,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t per_warp = bpp == 3 ? Stage<A3_FMT_RGB8>::per_warp : (bpp == 4 ? Stage<A3_FMT_RGBA8>::per_warp : Stage<A3_FMT_LUMA8>::per_warp);
    const size_t smem = (size_t)per_warp * kWarpsPerCta;
    // resident CTAs per SM: bounded by shared memory (~30 KB per CTA) and by the register cap of __launch_bounds__
    uint32_t ctas_per_sm = (uint32_t)((227 * 1024) / (smem + 1024));
    if (ctas_per_sm > (uint32_t)kMinCtasPerSm) ctas_per_sm = kMinCtasPerSm;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const uint32_t slots = (uint32_t)sms * ctas_per_sm * kWarpsPerCta;  // warps resident at once

    StripArgs a;
    a.grey = p.grey; a.mask = p.mask; a.bits = reinterpret_cast<uint8_t *>(p.bits);
    a.n = p.n; a.w = p.w; a.h = p.h;
    a.nstrips = (p.w + kCore - 1) / kCore;
    // row segments: at least as many as fill the resident warp slots once, and short enough (about 180 rows) that the grid
    // is several waves of CTAs: the block scheduler then evens out the slower edge strips and the warps drift out of phase.
    // Measured on B200, 256 x 1080p: one wave of 540-row segments 0.465 ms, 360 rows 0.454, 270 rows 0.426, 180 rows 0.420,
    // 135 rows 0.422, 108 rows 0.438 (every segment re-reads a 14-row halo, so a batch keeps them long; a single frame, whose time is
    // the march of one warp, is cut down to 64-row segments)
    uint32_t nsegs = tuning && tuning->seg_rows ? (p.h + tuning->seg_rows - 1) / tuning->seg_rows : 0;
    if (nsegs == 0) {
        const uint64_t per_seg = (uint64_t)p.n * a.nstrips;
        nsegs = per_seg >= slots ? 1 : (uint32_t)(slots / per_seg);
        const uint32_t pref = (p.h + 90) / 180;
        if (nsegs < pref) nsegs = pref;
        const uint32_t max_segs = p.h / 64 ? p.h / 64 : 1;  // calls of a few frames: down to 64 rows, so that more warps share the frame
        if (nsegs > max_segs) nsegs = max_segs;
    }
    a.seg_rows = (p.h + nsegs - 1) / nsegs;
    a.nsegs = (p.h + a.seg_rows - 1) / a.seg_rows;
    const uint64_t njobs = (uint64_t)p.n * a.nsegs * a.nstrips;
    if (njobs > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    a.njobs = (uint32_t)njobs;
    a.bits_row_bytes = (uint32_t)(4 * (p.bits_row_words ? p.bits_row_words : (p.w + 31) / 32));
    a.bits_col_bytes = (uint32_t)(4 * (p.bits_col_words ? p.bits_col_words : 1));
    a.bits_frame_bytes = p.bits_frame_words ? 4 * p.bits_frame_words : (size_t)p.h * 4 * ((p.w + 31) / 32);
    a.stages = kStages;
    a.wide_stores = (p.w % 8 == 0) && ((uintptr_t)p.grey % 8 == 0) && ((uintptr_t)p.mask % 8 == 0);
    const uint32_t grid = (a.njobs + kWarpsPerCta - 1) / kWarpsPerCta;

    // the kernel writes the 1-bit mask by bytes; when w is not a multiple of 32 the tail bytes of each row's last
    // word are never touched by it and must read as zero
    if (p.bits && p.w % 32 != 0) {
        // the words of the last 32-pixel column: the kernel only writes their first ceil((w % 32) / 8) bytes
        uint8_t *last = reinterpret_cast<uint8_t *>(p.bits) + (size_t)((p.w + 31) / 32 - 1) * a.bits_col_bytes;
        cudaError_t e;
        if (a.bits_row_bytes == 4)  // column-major: the column is h consecutive words per frame
            e = cudaMemset2DAsync(last, a.bits_frame_bytes, 0, (size_t)p.h * 4, p.n, stream);
        else if (a.bits_frame_bytes == (size_t)p.h * a.bits_row_bytes)  // row-major, frames back to back
            e = cudaMemset2DAsync(last, a.bits_row_bytes, 0, 4, (size_t)p.n * p.h, stream);
        else {
            e = cudaSuccess;
            for (uint32_t f = 0; f < p.n && e == cudaSuccess; f++)
                e = cudaMemset2DAsync(last + f * a.bits_frame_bytes, a.bits_row_bytes, 0, 4, p.h, stream);
        }
        if (e != cudaSuccess) return e;
    }
    if (info) {
        info->grid = grid; info->block = kWarpsPerCta * 32; info->smem_bytes = (uint32_t)smem; info->strips = a.nstrips; info->segs = a.nsegs;
        info->strip_cols = kCore; info->seg_rows = a.seg_rows; info->tma = 2; info->specialised_radius = 1;
    }
    const bool mask = p.mask != nullptr, bits = p.bits != nullptr;
    switch (p.format) {
        case A3_FMT_RGB8: return launch_strips<A3_FMT_RGB8>(map, a, mask, bits, grid, smem, stream);
        case A3_FMT_RGBA8: return launch_strips<A3_FMT_RGBA8>(map, a, mask, bits, grid, smem, stream);
        case A3_FMT_LUMA8: return launch_strips<A3_FMT_LUMA8>(map, a, mask, bits, grid, smem, stream);
        case A3_FMT_BGR8: return launch_strips<A3_FMT_BGR8>(map, a, mask, bits, grid, smem, stream);
        case A3_FMT_BGRA8: return launch_strips<A3_FMT_BGRA8>(map, a, mask, bits, grid, smem, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace a3
