// K1 — fused `into_luma8` + `adaptive_threshold` over a batch of frames (sm_100a).
//
// Replaces, bit for bit, the two calls at /root/reference/src/aruco.rs:60-61:
//   grey = image.into_luma8()                        (image 0.25:  (2126 R + 7152 G + 722 B) / 10000, u32, truncating)
//   thresholded = adaptive_threshold(&grey, radius)  (imageproc 0.25: clipped (2r+1)^2 box mean, integer floor,
//                                                     out = 255 iff pix >= mean)
//
// Shape of the computation (HBM-bound integer streaming; no tensor cores — nothing here is a contraction):
//   * a CTA owns a column strip x a row segment of one frame and marches down the rows once;
//   * each RGB row of the strip is brought into shared memory by one TMA bulk copy (cp.async.bulk, SASS UBLKCP)
//     into a NS-deep ring guarded by mbarriers, so HBM reads run NS-1 rows ahead of the arithmetic;
//   * a thread owns 4 adjacent columns: it converts its 12 RGB bytes to 4 grey bytes, keeps the running
//     (2r+1)-row column sums of those 4 columns in registers (add the new row, subtract the row that left the
//     window — the 2r+1 previous grey rows live in a thread-private shared-memory ring), publishes the 4 column
//     sums as u16 in a double-buffered exchange row, and after one __syncthreads sums the 2r+1 neighbouring
//     column sums for each of its pixels;
//   * `pix >= floor(S / cnt)`  <=>  `S < (pix + 1) * cnt` (cnt = clipped window area), so there is no division;
//   * every grey / mask byte is written exactly once with 32-bit coalesced stores; every RGB byte of the strip is
//     read exactly once (plus a 2r-row halo per segment and a (r..r+31)-column halo per interior strip edge).
// Algorithmic traffic: 3 B read + 1 B grey + 1 B mask = 5 B / pixel.
#include "a3_internal.h"

namespace a3 {
namespace {

constexpr int kStages = 4;        // TMA ring depth (rows in flight)
constexpr int kMaxThreads = 512;  // 4 columns per thread -> at most 2048 columns per strip
constexpr int kMaxRadius = 127;   // column sums of 2 r + 1 rows are exchanged as u16: 255 * 255 < 2^16

// Exact fixed-point form of (2126 R + 7152 G + 722 B) / 10000 for all 2^24 (R,G,B):
//   grey = (kWr*R + kWg*G + kWb*B + kBias) >> 24        (weights sum to 2^24; the maximum is < 2^32)
// verified exhaustively in tests/test_k1_identities.py.
constexpr uint32_t kWr = 3566836u, kWg = 11999065u, kWb = 1211315u, kBias = 1678u;

struct K1Args {
    const uint8_t *src;
    uint8_t *grey;
    uint8_t *mask;
    uint32_t *bits;
    size_t pitch, frame_stride;
    uint32_t n, w, h;
    uint32_t radius;
    uint32_t strips, strip_cols, segs, seg_rows;
    uint32_t words_per_row;
    size_t bits_row_words, bits_col_words, bits_frame_words;
    uint32_t stage_bytes;  // bytes of one RGB row slot (multiple of 16)
    int out_aligned;       // w % 4 == 0 and output bases 4-byte aligned: 32-bit stores allowed
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
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both addresses 16-B aligned).
__device__ __forceinline__ void tma_load_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t luma(uint32_t r, uint32_t g, uint32_t b) {
    return (kWr * r + kWg * g + kWb * b + kBias) >> 24;
}

template <int FMT>
struct Bpp { static constexpr int v = fmt_bpp(FMT); };
// luma of the pixel bytes b0 b1 b2 as they lie in memory (R,G,B or B,G,R)
template <int FMT>
__device__ __forceinline__ uint32_t luma_px(uint32_t b0, uint32_t b1, uint32_t b2) { return fmt_bgr(FMT) ? luma(b2, b1, b0) : luma(b0, b1, b2); }

// 4 grey bytes (little endian: pixel j in byte j) of the thread's 4 columns from the staged row.
template <int FMT>
__device__ __forceinline__ uint32_t grey4_from_stage(const uint8_t *stage, int t) {
    if constexpr (Bpp<FMT>::v == 3) {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(stage) + 3 * t;
        uint32_t w0 = s[0], w1 = s[1], w2 = s[2];
        uint32_t g0 = luma_px<FMT>(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff);
        uint32_t g1 = luma_px<FMT>(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff);
        uint32_t g2 = luma_px<FMT>((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff);
        uint32_t g3 = luma_px<FMT>((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24);
        return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else if constexpr (Bpp<FMT>::v == 4) {
        uint4 v = reinterpret_cast<const uint4 *>(stage)[t];
        uint32_t g0 = luma_px<FMT>(v.x & 0xff, (v.x >> 8) & 0xff, (v.x >> 16) & 0xff);
        uint32_t g1 = luma_px<FMT>(v.y & 0xff, (v.y >> 8) & 0xff, (v.y >> 16) & 0xff);
        uint32_t g2 = luma_px<FMT>(v.z & 0xff, (v.z >> 8) & 0xff, (v.z >> 16) & 0xff);
        uint32_t g3 = luma_px<FMT>(v.w & 0xff, (v.w >> 8) & 0xff, (v.w >> 16) & 0xff);
        return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        return reinterpret_cast<const uint32_t *>(stage)[t];
    }
}

// Same, straight from global memory with byte loads (rows that do not meet the bulk-copy alignment rules).
template <int FMT>
__device__ __forceinline__ uint32_t grey4_from_global(const uint8_t *row, uint32_t x, uint32_t xend) {
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (x + j < xend) {
            const uint8_t *p = row + (size_t)(x + j) * Bpp<FMT>::v;
            uint32_t g;
            if constexpr (FMT == A3_FMT_LUMA8) g = p[0];
            else g = luma_px<FMT>(p[0], p[1], p[2]);
            out |= g << (8 * j);
        }
    }
    return out;
}

// R = compile-time radius (0: use a.radius). TMA = rows staged by cp.async.bulk.
template <int FMT, int R, bool TMA>
__global__ void __launch_bounds__(kMaxThreads) k1_kernel(const K1Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int r = R ? R : (int)a.radius;
    const int win = 2 * r + 1;
    const int al = 4 * ((r + 3) / 4);          // exchange-row pad (multiple of 4, >= r)
    const int nt = blockDim.x, t = threadIdx.x;
    const int xrow = 4 * nt + 2 * al;          // u16 entries of one exchange row

    // ---- carve shared memory ----
    uint8_t *stages = smem;                                                         // kStages * stage_bytes (TMA only)
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem + (TMA ? kStages * a.stage_bytes : 0));  // win * nt
    uint16_t *xbuf = reinterpret_cast<uint16_t *>(ring + win * nt);                 // 2 * xrow (8-byte aligned: nt % 32 == 0)
    uint64_t *full = reinterpret_cast<uint64_t *>(xbuf + 2 * xrow + ((2 * xrow) & 3 ? 4 - ((2 * xrow) & 3) : 0));

    // ---- which tile ----
    uint32_t bid = blockIdx.x;
    const uint32_t strip = bid % a.strips; bid /= a.strips;
    const uint32_t seg = bid % a.segs;
    const uint32_t frame = bid / a.segs;
    const int cx0 = strip * a.strip_cols;
    const int cx1 = min((int)a.w, cx0 + (int)a.strip_cols);
    const int lx0 = max(0, cx0 - r) & ~31;
    const int lx1 = min((int)a.w, cx1 + r);
    const int y0 = seg * a.seg_rows;
    const int y1 = min((int)a.h, y0 + (int)a.seg_rows);
    const int x = lx0 + 4 * t;                   // first of this thread's 4 columns
    const uint8_t *fsrc = a.src + (size_t)frame * a.frame_stride + (size_t)lx0 * Bpp<FMT>::v;
    const uint32_t row_bytes = (uint32_t)(lx1 - lx0) * Bpp<FMT>::v;  // TMA path: launcher guarantees % 16 == 0
    // pixels of this thread that exist in the loaded span (others read as 0)
    uint32_t valid_mask = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (x + j < lx1) valid_mask |= 0xffu << (8 * j);

    // ---- init ----
    for (int i = t; i < win * nt; i += nt) ring[i] = 0;  // thread-private columns, but zero everything once
    for (int i = t; i < 2 * xrow; i += nt) xbuf[i] = 0;
    const int yl0 = max(0, y0 - r), yl1 = min((int)a.h, y1 + r);  // rows that are actually loaded
    if constexpr (TMA) {
        if (t == 0) {
            for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    if constexpr (TMA) {
        if (t == 0) {
            for (int j = 0; j < kStages && yl0 + j < yl1; j++) {
                mbar_expect_tx(&full[j], row_bytes);
                tma_load_row(stages + (size_t)j * a.stage_bytes, fsrc + (size_t)(yl0 + j) * a.pitch, row_bytes, &full[j]);
            }
        }
    }

    uint32_t cs0 = 0, cs1 = 0, cs2 = 0, cs3 = 0;  // column sums over the last `win` rows
    int slot_new = 0;                             // ring slot of row yi (== slot of row yi - win)
    int slot_pix = win - r;                       // ring slot of row yi - r  ((k - r) mod win at k = 0)
    if (slot_pix >= win) slot_pix -= win;
    int jload = 0;                                // index of the next loaded row (TMA ring position)

    // x-clipped window widths of the 4 pixels (do not change with the row)
    uint32_t nx[4];
#pragma unroll
    for (int j = 0; j < 4; j++) nx[j] = (uint32_t)(min((int)a.w - 1, x + j + r) - max(0, x + j - r) + 1);

    const int nrows = (y1 - y0) + 2 * r;
    for (int k = 0; k < nrows; k++) {
        const int yi = y0 - r + k;
        uint32_t g4 = 0;
        if (yi >= 0 && yi < (int)a.h) {
            if constexpr (TMA) {
                const int s = jload % kStages;
                mbar_wait(&full[s], (jload / kStages) & 1);
                if (x < lx1) g4 = grey4_from_stage<FMT>(stages + (size_t)s * a.stage_bytes, t) & valid_mask;
            } else {
                if (x < lx1) g4 = grey4_from_global<FMT>(fsrc + (size_t)yi * a.pitch - (size_t)lx0 * Bpp<FMT>::v, x, lx1);
            }
        }
        // vertical running sums
        const uint32_t old = ring[slot_new * nt + t];
        ring[slot_new * nt + t] = g4;
        cs0 += (g4 & 0xff) - (old & 0xff);
        cs1 += ((g4 >> 8) & 0xff) - ((old >> 8) & 0xff);
        cs2 += ((g4 >> 16) & 0xff) - ((old >> 16) & 0xff);
        cs3 += (g4 >> 24) - (old >> 24);
        uint16_t *xb = xbuf + (k & 1) * xrow;
        *reinterpret_cast<uint2 *>(xb + al + 4 * t) = make_uint2(cs0 | (cs1 << 16), cs2 | (cs3 << 16));
        __syncthreads();
        if constexpr (TMA) {
            // every thread has consumed stage jload % kStages: refill it with the row kStages ahead
            if (yi >= 0 && yi < (int)a.h) {
                if (t == 0 && yl0 + jload + kStages < yl1) {
                    const int s = jload % kStages;
                    mbar_expect_tx(&full[s], row_bytes);
                    tma_load_row(stages + (size_t)s * a.stage_bytes, fsrc + (size_t)(yl0 + jload + kStages) * a.pitch, row_bytes,
                                 &full[s]);
                }
                jload++;
            }
        }
        const int yo = yi - r;
        if (yo >= y0) {
            const uint32_t pix4 = ring[slot_pix * nt + t];
            uint32_t s[4];
            if constexpr (R > 0) {
                // window of 4 + 2*al u16 starting at entry 4t (8-byte aligned)
                constexpr int AL = 4 * ((R + 3) / 4);
                constexpr int NV = 4 + 2 * AL;
                uint32_t v[NV];
                const uint2 *wp = reinterpret_cast<const uint2 *>(xb + 4 * t);
#pragma unroll
                for (int i = 0; i < NV / 4; i++) {
                    uint2 q = wp[i];
                    v[4 * i] = q.x & 0xffff; v[4 * i + 1] = q.x >> 16;
                    v[4 * i + 2] = q.y & 0xffff; v[4 * i + 3] = q.y >> 16;
                }
                uint32_t acc = 0;
#pragma unroll
                for (int i = AL - R; i <= AL + R; i++) acc += v[i];
                s[0] = acc;
#pragma unroll
                for (int j = 1; j < 4; j++) {
                    acc += v[AL + j + R] - v[AL + j - 1 - R];
                    s[j] = acc;
                }
            } else {
                const uint16_t *c = xb + al + 4 * t;
                uint32_t acc = 0;
                for (int i = -r; i <= r; i++) acc += c[i];
                s[0] = acc;
                for (int j = 1; j < 4; j++) {
                    acc += (uint32_t)c[j + r] - (uint32_t)c[j - 1 - r];
                    s[j] = acc;
                }
            }
            const uint32_t ny = (uint32_t)(min((int)a.h - 1, yo + r) - max(0, yo - r) + 1);
            uint32_t m4 = 0, nib = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t pix = (pix4 >> (8 * j)) & 0xff;
                if (s[j] < (pix + 1) * (nx[j] * ny)) { m4 |= 0xffu << (8 * j); nib |= 1u << j; }
            }
            const size_t obase = ((size_t)frame * a.h + yo) * a.w;
            const bool core = x >= cx0 && x < cx1;
            if (core) {
                if (a.out_aligned && x + 4 <= cx1) {
                    if (a.grey) *reinterpret_cast<uint32_t *>(a.grey + obase + x) = pix4;
                    if (a.mask) *reinterpret_cast<uint32_t *>(a.mask + obase + x) = m4;
                } else {
                    for (int j = 0; j < 4 && x + j < cx1; j++) {
                        if (a.grey) a.grey[obase + x + j] = (uint8_t)(pix4 >> (8 * j));
                        if (a.mask) a.mask[obase + x + j] = (uint8_t)(m4 >> (8 * j));
                    }
                }
            }
            if (a.bits) {  // uniform branch: all lanes take part in the shuffles
                uint32_t valid = 0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j >= cx0 && x + j < cx1) valid |= 1u << j;
                uint32_t wv = (nib & valid) << (4 * (t & 7));
                wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
                wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
                wv |= __shfl_xor_sync(0xffffffffu, wv, 4);
                const int xw = x & ~31;  // first column of this lane group's word
                if ((t & 7) == 0 && xw >= cx0 && xw < cx1)
                    a.bits[(size_t)frame * a.bits_frame_words + (size_t)yo * a.bits_row_words + (size_t)(xw >> 5) * a.bits_col_words] = wv;
            }
        }
        if (++slot_new == win) slot_new = 0;
        if (++slot_pix == win) slot_pix = 0;
    }
}

template <int FMT, int R, bool TMA>
cudaError_t launch(const K1Args &a, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    auto kern = k1_kernel<FMT, R, TMA>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, block, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int FMT>
cudaError_t dispatch(const K1Args &a, bool r7, bool tma, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    if (r7) return tma ? launch<FMT, 7, true>(a, grid, block, smem, stream) : launch<FMT, 7, false>(a, grid, block, smem, stream);
    return tma ? launch<FMT, 0, true>(a, grid, block, smem, stream) : launch<FMT, 0, false>(a, grid, block, smem, stream);
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// K0 — into_luma8 of the DynamicImage variants K1 does not read directly (semantics [RECALLED] from image
// 0.25: `FromColor<LumaA<S>> for Luma<T>` keeps the luma channel, `FromColor<Rgb<S>> for Luma<T>` is
// T::from_primitive(rgb_to_luma(rgb)) with rgb_to_luma in the subpixel's `Larger` type — u32 for u16 — and
// `FromPrimitive<u16> for u8` is (c + 128) / 257).  One thread per pixel; a streaming pass, nothing to tile.
template <int FMT>
__global__ void __launch_bounds__(256) k0_kernel(const uint8_t *src, uint32_t w, uint32_t h, size_t pitch, size_t frame_stride, uint8_t *dst,
                                                 size_t dst_pitch, size_t dst_frame_stride) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= w) return;
    const uint8_t *p = src + (size_t)f * frame_stride + (size_t)y * pitch + (size_t)x * fmt_bpp(FMT);
    uint32_t g;
    if constexpr (FMT == A3_FMT_LUMAA8) {
        g = p[0];
    } else {
        auto u16at = [&](int i) { return (uint32_t)p[2 * i] | ((uint32_t)p[2 * i + 1] << 8); };  // native (little-endian) u16, any alignment
        uint32_t l;
        if constexpr (FMT == A3_FMT_LUMA16 || FMT == A3_FMT_LUMAA16) l = u16at(0);
        else l = (2126u * u16at(0) + 7152u * u16at(1) + 722u * u16at(2)) / 10000u;
        g = (l + 128u) / 257u;
    }
    dst[(size_t)f * dst_frame_stride + (size_t)y * dst_pitch + x] = (uint8_t)g;
}

}  // namespace

cudaError_t k0_to_luma8(const uint8_t *src, int format, uint32_t n, uint32_t w, uint32_t h, size_t pitch, size_t frame_stride, uint8_t *dst,
                        size_t dst_pitch, size_t dst_frame_stride, cudaStream_t stream) {
    if (n == 0 || w == 0 || h == 0) return cudaSuccess;
    if (h > 65535 || n > 65535) return cudaErrorInvalidConfiguration;
    const dim3 grid((w + 255) / 256, h, n), block(256);
    switch (format) {
        case A3_FMT_LUMAA8: k0_kernel<A3_FMT_LUMAA8><<<grid, block, 0, stream>>>(src, w, h, pitch, frame_stride, dst, dst_pitch, dst_frame_stride); break;
        case A3_FMT_LUMA16: k0_kernel<A3_FMT_LUMA16><<<grid, block, 0, stream>>>(src, w, h, pitch, frame_stride, dst, dst_pitch, dst_frame_stride); break;
        case A3_FMT_LUMAA16: k0_kernel<A3_FMT_LUMAA16><<<grid, block, 0, stream>>>(src, w, h, pitch, frame_stride, dst, dst_pitch, dst_frame_stride); break;
        case A3_FMT_RGB16: k0_kernel<A3_FMT_RGB16><<<grid, block, 0, stream>>>(src, w, h, pitch, frame_stride, dst, dst_pitch, dst_frame_stride); break;
        case A3_FMT_RGBA16: k0_kernel<A3_FMT_RGBA16><<<grid, block, 0, stream>>>(src, w, h, pitch, frame_stride, dst, dst_pitch, dst_frame_stride); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t k1_gray_threshold(const K1Params &p, const K1Tuning *tuning, cudaStream_t stream, K1LaunchInfo *info) {
    if (p.n == 0 || p.w == 0 || p.h == 0) return cudaSuccess;
    if (p.radius == 0 || p.radius > (uint32_t)kMaxRadius) return cudaErrorInvalidValue;
    if (fmt_wide(p.format)) return cudaErrorInvalidValue;  // callers convert those with k0_to_luma8 first
    if (!(tuning && tuning->force_generic) && k1_strips_eligible(p)) return k1_strips(p, tuning, stream, info);
    const uint32_t r = p.radius;
    const uint32_t bpp = fmt_bpp(p.format);
    // threads per CTA: bounded by kMaxThreads and by shared memory (a ring of 2 r + 1 grey rows, 4 bytes per thread and row, the
    // TMA stages and the exchange rows); a strip's loaded span is its core plus r columns on either side, its start rounded down to 32
    const uint32_t per_thread = 4 * (2 * r + 1) + 4 * bpp * (uint32_t)kStages + 16;
    uint32_t nt_max = ((200u * 1024u) / per_thread) & ~31u;
    if (nt_max > (uint32_t)kMaxThreads) nt_max = kMaxThreads;
    if (4 * nt_max < 2 * r + 32 + 32) return cudaErrorInvalidConfiguration;
    const uint32_t max_core = (4 * nt_max - (2 * r + 32)) & ~31u;

    // ---- tiling ----
    uint32_t strip_cols = tuning && tuning->strip_cols ? round_up(tuning->strip_cols, 32) : 0;
    uint32_t seg_rows = tuning && tuning->seg_rows ? tuning->seg_rows : 0;
    if (strip_cols == 0) {
        uint32_t ns = (p.w + max_core - 1) / max_core;
        strip_cols = round_up((p.w + ns - 1) / ns, 32);
    }
    if (strip_cols > max_core) strip_cols = max_core;
    if (strip_cols < 32) strip_cols = 32;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (seg_rows == 0) {
        // enough CTAs for ~8 per SM, but keep the 2r-row halo below ~25 % of a segment
        const uint32_t target = 8u * (uint32_t)sms;
        uint32_t strips0 = (p.w + strip_cols - 1) / strip_cols;
        uint32_t per_frame = (target + p.n * strips0 - 1) / (p.n * strips0);
        if (per_frame < 1) per_frame = 1;
        seg_rows = (p.h + per_frame - 1) / per_frame;
        if (seg_rows < 8 * r) seg_rows = 8 * r;
        if (seg_rows > p.h) seg_rows = p.h;
        if (!(tuning && tuning->strip_cols)) {
            // few frames: narrow the strips until the grid can cover the machine twice
            while ((uint64_t)p.n * ((p.h + seg_rows - 1) / seg_rows) * ((p.w + strip_cols - 1) / strip_cols) < 2u * (uint32_t)sms &&
                   strip_cols > 224)
                strip_cols = round_up(strip_cols / 2, 32);
        }
    }
    const uint32_t strips = (p.w + strip_cols - 1) / strip_cols;
    const uint32_t segs = (p.h + seg_rows - 1) / seg_rows;

    // ---- widest loaded span over the strips -> threads per CTA ----
    uint32_t max_cols = 0;
    bool tma = !(tuning && tuning->force_no_tma) && (p.pitch % 16 == 0) && (p.frame_stride % 16 == 0) &&
               ((uintptr_t)p.src % 16 == 0);
    for (uint32_t s = 0; s < strips; s++) {
        uint32_t cx0 = s * strip_cols, cx1 = cx0 + strip_cols < p.w ? cx0 + strip_cols : p.w;
        uint32_t lx0 = (cx0 > r ? cx0 - r : 0u) & ~31u;
        uint32_t lx1 = cx1 + r < p.w ? cx1 + r : p.w;
        if (lx1 - lx0 > max_cols) max_cols = lx1 - lx0;
        if (((lx1 - lx0) * bpp) % 16 != 0 || (lx0 * bpp) % 16 != 0) tma = false;
    }
    uint32_t nt = round_up((max_cols + 3) / 4, 32);
    if (nt > (uint32_t)kMaxThreads) return cudaErrorInvalidConfiguration;

    K1Args a;
    a.src = p.src; a.grey = p.grey; a.mask = p.mask; a.bits = p.bits;
    a.pitch = p.pitch; a.frame_stride = p.frame_stride;
    a.n = p.n; a.w = p.w; a.h = p.h; a.radius = r;
    a.strips = strips; a.strip_cols = strip_cols; a.segs = segs; a.seg_rows = seg_rows;
    a.words_per_row = (p.w + 31) / 32;
    a.bits_row_words = p.bits_row_words ? p.bits_row_words : a.words_per_row;
    a.bits_col_words = p.bits_col_words ? p.bits_col_words : 1;
    a.bits_frame_words = p.bits_frame_words ? p.bits_frame_words : (size_t)p.h * a.words_per_row;
    a.stage_bytes = round_up(4 * nt * bpp, 16);
    a.out_aligned = (p.w % 4 == 0) && ((uintptr_t)p.grey % 4 == 0) && ((uintptr_t)p.mask % 4 == 0);

    const uint32_t win = 2 * r + 1, al = 4 * ((r + 3) / 4), xrow = 4 * nt + 2 * al;
    size_t smem = (tma ? (size_t)kStages * a.stage_bytes : 0) + (size_t)win * nt * 4 + (size_t)2 * xrow * 2 + 8 + kStages * 8;
    const uint64_t grid64 = (uint64_t)p.n * segs * strips;
    if (grid64 > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    dim3 grid((uint32_t)grid64), block(nt);
    if (info) {
        info->grid = grid.x; info->block = nt; info->smem_bytes = (uint32_t)smem; info->strips = strips; info->segs = segs;
        info->strip_cols = strip_cols; info->seg_rows = seg_rows; info->tma = tma; info->specialised_radius = r == 7;
    }
    switch (p.format) {
        case A3_FMT_RGB8: return dispatch<A3_FMT_RGB8>(a, r == 7, tma, grid, block, smem, stream);
        case A3_FMT_RGBA8: return dispatch<A3_FMT_RGBA8>(a, r == 7, tma, grid, block, smem, stream);
        case A3_FMT_LUMA8: return dispatch<A3_FMT_LUMA8>(a, r == 7, tma, grid, block, smem, stream);
        case A3_FMT_BGR8: return dispatch<A3_FMT_BGR8>(a, r == 7, tma, grid, block, smem, stream);
        case A3_FMT_BGRA8: return dispatch<A3_FMT_BGRA8>(a, r == 7, tma, grid, block, smem, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace a3
