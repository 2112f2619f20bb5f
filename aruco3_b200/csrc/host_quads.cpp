// Host stage of the detection path: border following on the thresholded mask, polygon simplification and the
// quad filters.  BASELINE.json's north_star keeps this stage on the host ("the existing contour tracing and
// approxPolyDP quad filtering produce identical candidates"); this is a from-scratch C++ implementation of it
// that works on the 1-bit mask K1 emits, not a port of imageproc's i32 label image.
//
// Behaviour it reproduces (checked against oracle/a3ref.c on every test frame):
//   imageproc::contours::find_contours::<u32>                 call site /root/reference/src/aruco.rs:64
//   contours_to_candidates (RDP eps = 0.05*len, 4 vertices,   /root/reference/src/aruco.rs:124-166
//     convex_hull keeps 4, squared-edge vs unsquared limit)
//   enforce_clockwise_corners                                 /root/reference/src/aruco.rs:168-185
//   discard_too_near + perimeter                              /root/reference/src/aruco.rs:187-232, 328-338
//
// Suzuki-Abe labels collapse to two bit planes: `seen` (label != 1: the pixel lies on a border already followed)
// and `redge` (label < 0: followed with the zero pixel to its east examined).  An outer border starts at a
// foreground pixel with a zero west neighbour that is not `seen`; a hole border at a foreground pixel with a zero
// east neighbour that is not `redge`.  Start candidates are found 32 pixels at a time from the mask words.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "a3_internal.h"

namespace a3 {
namespace {

struct Pt { int32_t x, y; };

// direction ring, screen-clockwise starting west (imageproc's VecDeque order)
constexpr int kDx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
constexpr int kDy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
// index of (dx,dy) in the ring, dx,dy in {-1,0,1}: table[(dy+1)*3 + dx+1]
constexpr int kDirOf[9] = {1, 2, 3, 0, -1, 4, 7, 6, 5};

// Bit planes with a guard word on each side of a row and a guard row above and below, so a 3x3 neighbourhood can be
// read without bounds checks: pixel (x, y) is bit (x & 31) of word (y + 1) * stride + 1 + (x >> 5).
struct Planes {
    std::vector<uint32_t> fg, seen, redge;
    uint32_t stride = 0, w = 0, h = 0;
    void reset(const uint32_t *bits, uint32_t wpr, uint32_t w_, uint32_t h_, size_t row_stride) {
        w = w_; h = h_; stride = wpr + 2;
        const size_t n = (size_t)stride * (h + 2);
        fg.assign(n, 0u);
        seen.assign(n, 0u);
        redge.assign(n, 0u);
        for (uint32_t y = 0; y < h; y++) memcpy(&fg[(size_t)(y + 1) * stride + 1], bits + (size_t)y * row_stride, (size_t)wpr * 4);
    }
    inline size_t word(int x, int y) const { return (size_t)(y + 1) * stride + 1 + (x >> 5); }
    // 3 bits of row y at columns x-1, x, x+1 (bit 0 = x-1); x in [0, w), y in [-1, h]
    inline uint32_t three(int x, int y) const {
        const int o = x + 31;  // bit offset of column x-1 within the padded row
        uint64_t v;
        memcpy(&v, &fg[(size_t)(y + 1) * stride + (o >> 5)], 8);
        return (uint32_t)(v >> (o & 31)) & 7u;
    }
    // 3x3 neighbourhood of (x, y): bits 0-2 row y-1, bits 3-5 row y, bits 6-8 row y+1 (bit 0 of each = column x-1)
    inline uint32_t hood(int x, int y) const {
        const int o = x + 31;
        const uint32_t *p = &fg[(size_t)y * stride + (o >> 5)];  // row y-1 of the padded plane
        const int sh = o & 31;
        uint64_t t, m, b;
        memcpy(&t, p, 8);
        memcpy(&m, p + stride, 8);
        memcpy(&b, p + 2 * stride, 8);
        return ((uint32_t)(t >> sh) & 7u) | (((uint32_t)(m >> sh) & 7u) << 3) | (((uint32_t)(b >> sh) & 7u) << 6);
    }
    // the same as ring bits: bit d set iff the neighbour in ring direction d (w nw n ne e se s sw) is foreground
    static inline uint32_t ring_of(uint32_t hood9) {
        const uint32_t t = hood9 & 7u, m = (hood9 >> 3) & 7u, b = hood9 >> 6;
        return (m & 1u) | ((t & 1u) << 1) | ((t & 2u) << 1) | ((t & 4u) << 1) | ((m & 4u) << 2) | ((b & 4u) << 3) | ((b & 2u) << 5) | ((b & 1u) << 7);
    }
    inline void mark(int x, int y, uint32_t right_edge) {
        const size_t i = word(x, y);
        const uint32_t b = 1u << (x & 31);
        seen[i] |= b;
        redge[i] |= b & (0u - right_edge);
    }
};

// Next step of the border following, tabulated: for the ring direction `front` of the previous border pixel and the
// 3x3 neighbourhood, the direction d4 of the next border pixel (first foreground neighbour counter-clockwise, starting
// just after `front`, `front` itself last) and whether the east neighbour was examined before it was found
// (imageproc's is_right_edge).  entry = d4 | right_edge << 3 | (dx + 1) << 4 | (dy + 1) << 6 | next front << 8.
struct StepTable {
    uint16_t e[8][512];
    StepTable() {
        for (int front = 0; front < 8; front++)
            for (int hood = 0; hood < 512; hood++) {
                const uint32_t nb = Planes::ring_of((uint32_t)hood);
                int d4 = front;
                for (int k = 1; k <= 7; k++) {
                    const int d = (front - k) & 7;
                    if ((nb >> d) & 1) { d4 = d; break; }
                }
                const int ord_e = (front - 1 - 4) & 7, ord_4 = (front - 1 - d4) & 7;
                e[front][hood] = (uint16_t)(d4 | ((ord_e < ord_4) << 3) | ((kDx[d4] + 1) << 4) | ((kDy[d4] + 1) << 6) | (((d4 + 4) & 7) << 8));
            }
    }
};
const StepTable kStep;

struct PtBuf {  // grow-only point buffer, structure of arrays (the simplification pass vectorises over it)
    std::vector<int32_t> xs, ys;
    size_t n = 0;
    inline void clear() { n = 0; }
    inline void push(int x, int y) {
        if (n == xs.size()) {
            const size_t cap = xs.size() ? xs.size() * 2 : 4096;
            xs.resize(cap);
            ys.resize(cap);
        }
        xs[n] = x; ys[n] = y;
        n++;
    }
};

// Follow one border from (sx,sy); `adj_dir` = ring index of the zero neighbour the scan came from.
void follow(Planes &pl, int sx, int sy, int adj_dir, PtBuf &out) {
    out.clear();
    const uint32_t nb0 = Planes::ring_of(pl.hood(sx, sy));
    int first = -1;
    for (int k = 0; k < 8; k++) {  // clockwise from the zero neighbour
        const int d = (adj_dir + k) & 7;
        if ((nb0 >> d) & 1) { first = d; break; }
    }
    if (first < 0) {  // isolated pixel
        out.push(sx, sy);
        pl.mark(sx, sy, 1u);
        return;
    }
    const int p1x = sx + kDx[first], p1y = sy + kDy[first];
    int p3x = sx, p3y = sy;
    uint32_t front = (uint32_t)first;  // ring direction from p3 to p2 (the previous border pixel)
    const int last_col = (int)pl.w - 1;
    for (;;) {
        out.push(p3x, p3y);
        const uint32_t e = kStep.e[front][pl.hood(p3x, p3y)];
        pl.mark(p3x, p3y, ((e >> 3) & 1u) | (uint32_t)(p3x == last_col));
        const int p4x = p3x + (int)((e >> 4) & 3u) - 1, p4y = p3y + (int)((e >> 6) & 3u) - 1;
        if (p4x == sx && p4y == sy && p3x == p1x && p3y == p1y) break;
        front = e >> 8;  // from the new pixel back to this one
        p3x = p4x; p3y = p4y;
    }
}

// max_i |a x_i + b y_i + c| over i in [lo, hi] — exact in int32 (the caller checks the coordinate range)
__attribute__((target_clones("avx2", "default")))
int32_t max_abs_line32(const int32_t *xs, const int32_t *ys, uint32_t lo, uint32_t hi, int32_t a, int32_t b, int32_t c) {
    int32_t m = 0;
    for (uint32_t i = lo; i <= hi; i++) {
        int32_t v = a * xs[i] + b * ys[i] + c;
        v = v < 0 ? -v : v;
        m = v > m ? v : m;
    }
    return m;
}
// first i in [lo, hi] with |a x_i + b y_i + c| >= t (one exists)
__attribute__((target_clones("avx2", "default")))
uint32_t first_at_least32(const int32_t *xs, const int32_t *ys, uint32_t lo, uint32_t hi, int32_t a, int32_t b, int32_t c, int32_t t) {
    uint32_t i = lo;
    for (; i + 31 <= hi; i += 32) {
        int any = 0;
        for (uint32_t j = i; j < i + 32; j++) {
            int32_t v = a * xs[j] + b * ys[j] + c;
            v = v < 0 ? -v : v;
            any |= v >= t;
        }
        if (any) break;
    }
    for (; i < hi; i++) {
        int32_t v = a * xs[i] + b * ys[i] + c;
        v = v < 0 ? -v : v;
        if (v >= t) return i;
    }
    return hi;
}

// approximate_polygon_dp(curve, eps, closed = true) — iterative Ramer-Douglas-Peucker with the reference's
// "first strict maximum of |a x + b y + c| / sqrt(a^2+b^2)" rule.  The numerator is an exact integer and
// the quotient is monotone in it, so a span is two passes: the maximum numerator M (vectorised), then the first point
// whose quotient equals M's — i.e. whose numerator reaches the smallest integer t with t / den == M / den (t is M unless
// two quotients round to the same double).  `small`: coordinates <= 16384, int32 arithmetic is exact.
// Returns the vertex count, writing at most `cap` vertices; counts above cap mean "not a quad".
size_t simplify_closed(const int32_t *xs, const int32_t *ys, size_t n, double eps, bool small, Pt *out, size_t cap) {
    struct Span { uint32_t lo, hi; };
    static thread_local std::vector<Span> stack;
    stack.clear();
    stack.push_back({0, (uint32_t)(n - 1)});
    size_t nout = 0;
    // vertices are emitted in curve order: for a span we emit its first point when it is a leaf; the very last
    // point of the whole curve would follow, and closed=true pops exactly that one.
    while (!stack.empty()) {
        const Span s = stack.back();
        stack.pop_back();
        const int64_t sx = xs[s.lo], sy = ys[s.lo], ex = xs[s.hi], ey = ys[s.hi];
        const int64_t a = sy - ey, b = ex - sx, cc = sx * ey - ex * sy;
        const double den = sqrt((double)a * (double)a + (double)b * (double)b);
        double dmax = 0.0;
        uint32_t index = 0;
        if (s.hi > s.lo) {
            if (small) {
                const int32_t M = max_abs_line32(xs, ys, s.lo + 1, s.hi, (int32_t)a, (int32_t)b, (int32_t)cc);
                if (M > 0) {
                    const double q = (double)M / den;
                    if (q > 0.0) {  // not NaN: den == 0 implies M == 0
                        int32_t t = M;
                        while (t > 1 && (double)(t - 1) / den == q) t--;
                        dmax = q;
                        index = first_at_least32(xs, ys, s.lo + 1, s.hi, (int32_t)a, (int32_t)b, (int32_t)cc, t);
                    }
                }
            } else {
                int64_t best_num = 0;
                for (uint32_t i = s.lo + 1; i <= s.hi; i++) {
                    int64_t num = a * xs[i] + b * ys[i] + cc;
                    if (num < 0) num = -num;
                    if (num > best_num) {
                        const double d = (double)num / den;
                        if (d > dmax) { dmax = d; index = i; best_num = num; }
                    }
                }
            }
        }
        if (dmax > eps) {
            stack.push_back({index, s.hi});
            stack.push_back({s.lo, index});
        } else {
            if (nout < cap) out[nout] = Pt{xs[s.lo], ys[s.lo]};
            nout++;
            if (nout > cap) return nout;
        }
    }
    return nout;
}

inline int orient(Pt p, Pt q, Pt r) {
    const int64_t v = (int64_t)(q.y - p.y) * (r.x - q.x) - (int64_t)(q.x - p.x) * (r.y - q.y);  // 64-bit: coordinates go up to 65535
    return v == 0 ? 0 : (v > 0 ? 1 : -1);  // 1 clockwise, -1 counter-clockwise (imageproc's naming)
}
inline double dist(Pt p, Pt q) {
    const double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y;
    return sqrt(dx * dx + dy * dy);
}

// imageproc::geometry::convex_hull on exactly four points; true when the hull keeps all four (written to q).
bool hull4(Pt q[4]) {
    int sp = 0;
    for (int i = 1; i < 4; i++)
        if (q[i].y < q[sp].y || (q[i].y == q[sp].y && q[i].x < q[sp].x)) sp = i;
    const Pt start = q[sp];
    Pt rest[3];
    {  // swap(0, sp); remove(0)
        Pt tmp[4] = {q[0], q[1], q[2], q[3]};
        std::swap(tmp[0], tmp[sp]);
        rest[0] = tmp[1]; rest[1] = tmp[2]; rest[2] = tmp[3];
    }
    auto less = [&](Pt a, Pt b) {  // sort_by closure: never Equal
        const int o = orient(start, a, b);
        if (o == 0) return dist(start, a) < dist(start, b);
        return o < 0;
    };
    for (int i = 1; i < 3; i++) {  // insertion sort (std's small-slice path)
        const Pt key = rest[i];
        int j = i;
        while (j > 0 && less(key, rest[j - 1])) { rest[j] = rest[j - 1]; j--; }
        rest[j] = key;
    }
    Pt rem[3];
    int nr = 0;
    for (int i = 0; i < 3;) {
        Pt p = rest[i++];
        while (i < 3 && orient(start, p, rest[i]) == 0) p = rest[i++];
        rem[nr++] = p;
    }
    Pt st[4];
    int ns = 0;
    st[ns++] = start;
    for (int k = 0; k < nr; k++) {
        while (ns > 1 && orient(st[ns - 2], st[ns - 1], rem[k]) != -1) ns--;
        st[ns++] = rem[k];
    }
    if (ns != 4) return false;
    for (int i = 0; i < 4; i++) q[i] = st[i];
    return true;
}

inline float perimeter(const uint32_t *q) {
    float p = 0.0f;
    for (int i = 0; i < 4; i++) {
        const int j = (i + 1) & 3;
        const float dx = (float)q[2 * i] - (float)q[2 * j], dy = (float)q[2 * i + 1] - (float)q[2 * j + 1];
        p += sqrtf((dx * dx) + (dy * dy));
    }
    return p;
}

}  // namespace

void bits_from_mask(const uint8_t *mask, uint32_t w, uint32_t h, std::vector<uint32_t> &bits, uint32_t *words_per_row) {
    const uint32_t wpr = (w + 31) / 32;
    *words_per_row = wpr;
    bits.assign((size_t)wpr * h, 0);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++)
            if (mask[(size_t)y * w + x]) bits[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
}

void quads_from_bits(const uint32_t *bits, uint32_t wpr, uint32_t w, uint32_t h, const a3_config &cfg,
                     std::vector<uint32_t> &quads_out, QuadStats *stats, size_t row_stride_words) {
    const uint32_t mn = w < h ? w : h;
    const uint32_t min_edge_length = (uint32_t)((float)mn * cfg.min_side_length_factor);  // src/aruco.rs:55
    const float min_corner_separation = (float)mn * cfg.min_corner_separation_factor;     // src/aruco.rs:56
    const double eps_factor = cfg.contour_simplification_epsilon;

    static thread_local Planes pl;  // per host thread scratch, reused across frames (no page faults after the first)
    static thread_local PtBuf contour;
    pl.reset(bits, wpr, w, h, row_stride_words ? row_stride_words : wpr);
    const uint32_t S = pl.stride;
    const bool small = w <= 16384 && h <= 16384;

    // A surviving quad has two vertices at least sqrt(min_edge_length) apart; an 8-connected closed border that
    // reaches that far and comes back has at least sqrt(2 * min_edge_length) points.  Shorter borders cannot
    // pass the edge test (src/aruco.rs:149-159) whatever RDP does with them, so they are followed (their marks
    // matter to later starts) but not simplified.
    const size_t min_points = (size_t)floor(sqrt(2.0 * (double)min_edge_length));

    std::vector<uint32_t> quads;  // before discard
    QuadStats st;
    const uint32_t last_q = (w - 1) >> 6;                     // 64-column group that holds the last column
    const uint64_t last_bit = 1ull << ((w - 1) & 63);
    for (uint32_t y = 0; y < h; y++) {
        const size_t r0 = (size_t)(y + 1) * S + 1;            // word of column 0; r0 - 1 and r0 + wpr are zero guard words
        const uint32_t *row = &pl.fg[r0];
        for (uint32_t k = 0; k < wpr; k += 2) {               // 64 columns at a time; an odd last word pairs with the guard
            uint64_t f;
            memcpy(&f, row + k, 8);
            if (!f) continue;
            const uint32_t prev = row[(int)k - 1], next = row[k + 2];  // k + 2 <= wpr + 1: at worst the next row's guard
            if (f == ~0ull && (prev >> 31) && (next & 1u)) continue;  // interior of a white area: no border starts here
            const uint64_t west = (f << 1) | (prev >> 31);
            const uint64_t east = (f >> 1) | ((uint64_t)next << 63);
            uint64_t outer_geom = f & ~west, hole_geom = f & ~east;
            if (k == 0) outer_geom &= ~1ull;                   // `x > 0`
            if ((k >> 1) == last_q) hole_geom &= ~last_bit;    // `x + 1 < w`: the last column never starts a hole border
            if (!(outer_geom | hole_geom)) continue;
            const size_t wi = r0 + k;
            uint64_t below = 0;  // bits at or below the last start taken in this group
            for (;;) {
                // start candidates left in this group: the marks change while borders are followed, so re-read them
                uint64_t seen, redge;
                memcpy(&seen, &pl.seen[wi], 8);
                memcpy(&redge, &pl.redge[wi], 8);
                const uint64_t o = outer_geom & ~seen & ~below;    // zero pixel to the west, not on any followed border
                const uint64_t hh = hole_geom & ~redge & ~below;   // zero pixel to the east, not examined yet
                const uint64_t pending = o | hh;
                if (!pending) break;
                const uint64_t b = pending & (0ull - pending);
                below |= b | (b - 1);
                const int x = (int)(k * 32 + __builtin_ctzll(b));
                const int adj_dir = (o & b) ? 0 : 4;  // the outer test comes first (if / else if in the reference)
                follow(pl, x, (int)y, adj_dir, contour);
                st.n_contours++;
                st.n_contour_points += contour.n;
                const size_t n = contour.n;
                if (n < min_points || n < 4) continue;
                Pt q[4];
                if (simplify_closed(contour.xs.data(), contour.ys.data(), n, (double)n * eps_factor, small, q, 4) != 4) continue;
                if (!hull4(q)) continue;
                uint32_t cmin = min_edge_length + 1;
                for (int i = 0; i < 4; i++) {
                    const int j = (i + 1) & 3;
                    const int32_t dx = q[i].x - q[j].x, dy = q[i].y - q[j].y;
                    cmin = std::min(cmin, (uint32_t)(dx * dx + dy * dy));
                }
                if (cmin < min_edge_length) continue;  // squared vs unsquared on purpose
                for (int i = 0; i < 4; i++) { quads.push_back((uint32_t)q[i].x); quads.push_back((uint32_t)q[i].y); }
            }
        }
    }
    size_t nq = quads.size() / 8;
    st.n_before_discard = nq;
    // enforce_clockwise_corners
    for (size_t i = 0; i < nq; i++) {
        uint32_t *p = &quads[i * 8];
        const int32_t dx1 = (int32_t)p[2] - (int32_t)p[0], dy1 = (int32_t)p[3] - (int32_t)p[1];
        const int32_t dx2 = (int32_t)p[4] - (int32_t)p[0], dy2 = (int32_t)p[5] - (int32_t)p[1];
        if (dx1 * dy2 - dy1 * dx2 < 0) { std::swap(p[2], p[6]); std::swap(p[3], p[7]); }
    }
    // discard_too_near
    std::vector<uint8_t> dead(nq, 0);
    for (size_t i = 0; i + 1 < nq; i++) {
        if (dead[i]) continue;
        const float per_i = perimeter(&quads[i * 8]);
        for (size_t j = i + 1; j < nq; j++) {
            if (dead[j]) continue;
            float d = 0.0f;
            for (int c = 0; c < 4; c++) {
                const float dx = (float)quads[i * 8 + 2 * c] - (float)quads[j * 8 + 2 * c];
                const float dy = (float)quads[i * 8 + 2 * c + 1] - (float)quads[j * 8 + 2 * c + 1];
                d += sqrtf((dx * dx) + (dy * dy));
            }
            if ((d / 4.0f) < min_corner_separation) {
                if (dead[i] || dead[j]) continue;
                if (per_i >= perimeter(&quads[j * 8])) dead[j] = 1;
                else dead[i] = 1;  // the reference keeps scanning with this i
            }
        }
    }
    for (size_t i = 0; i < nq; i++)
        if (!dead[i]) quads_out.insert(quads_out.end(), quads.begin() + i * 8, quads.begin() + i * 8 + 8);
    if (stats) *stats = st;
}

}  // namespace a3
