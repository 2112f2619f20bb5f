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

struct Planes {
    const uint32_t *fg;
    std::vector<uint32_t> seen, redge;
    uint32_t wpr, w, h;
    inline bool on(int x, int y) const {
        return (unsigned)x < w && (unsigned)y < h && ((fg[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u);
    }
    inline void mark(int x, int y, bool right_edge) {
        const size_t i = (size_t)y * wpr + (x >> 5);
        const uint32_t b = 1u << (x & 31);
        seen[i] |= b;
        if (right_edge) redge[i] |= b;
    }
};

// Follow one border from (sx,sy); `adj_dir` = ring index of the zero neighbour the scan came from.
void follow(Planes &pl, int sx, int sy, int adj_dir, std::vector<Pt> &out) {
    out.clear();
    int first = -1;
    for (int k = 0; k < 8; k++) {  // clockwise from the zero neighbour
        const int d = (adj_dir + k) & 7;
        if (pl.on(sx + kDx[d], sy + kDy[d])) { first = d; break; }
    }
    if (first < 0) {  // isolated pixel
        out.push_back({sx, sy});
        pl.mark(sx, sy, true);
        return;
    }
    const int p1x = sx + kDx[first], p1y = sy + kDy[first];
    int p2x = p1x, p2y = p1y, p3x = sx, p3y = sy;
    for (;;) {
        out.push_back({p3x, p3y});
        const int front = kDirOf[(p2y - p3y + 1) * 3 + (p2x - p3x + 1)];
        // counter-clockwise, starting just before `front`, `front` itself last
        int d4 = front;
        for (int k = 1; k <= 7; k++) {
            const int d = (front - k) & 7;
            if (pl.on(p3x + kDx[d], p3y + kDy[d])) { d4 = d; break; }
        }
        // east (ring index 4) examined before the pixel that was found?
        const int ord_e = (front - 1 - 4) & 7, ord_4 = (front - 1 - d4) & 7;
        pl.mark(p3x, p3y, p3x + 1 == (int)pl.w || ord_e < ord_4);
        const int p4x = p3x + kDx[d4], p4y = p3y + kDy[d4];
        if (p4x == sx && p4y == sy && p3x == p1x && p3y == p1y) break;
        p2x = p3x; p2y = p3y; p3x = p4x; p3y = p4y;
    }
}

// approximate_polygon_dp(curve, eps, closed = true) — iterative Ramer-Douglas-Peucker with the reference's
// "first strict maximum of |a x + b y + c| / sqrt(a^2+b^2)" rule.  The numerator is an exact integer, so the
// division is only evaluated for points whose numerator beats the best so far (the quotient is monotone in it).
// Returns the vertex count, writing at most `cap` vertices; counts above cap mean "not a quad".
size_t simplify_closed(const Pt *c, size_t n, double eps, Pt *out, size_t cap) {
    struct Span { uint32_t lo, hi; };
    std::vector<Span> stack;
    stack.push_back({0, (uint32_t)(n - 1)});
    size_t nout = 0;
    // vertices are emitted in curve order: for a span we emit its first point when it is a leaf; the very last
    // point of the whole curve would follow, and closed=true pops exactly that one.
    while (!stack.empty()) {
        const Span s = stack.back();
        stack.pop_back();
        const int64_t sx = c[s.lo].x, sy = c[s.lo].y, ex = c[s.hi].x, ey = c[s.hi].y;
        const int64_t a = sy - ey, b = ex - sx, cc = sx * ey - ex * sy;
        const double den = sqrt((double)a * (double)a + (double)b * (double)b);
        int64_t best_num = 0;
        double dmax = 0.0;
        uint32_t index = 0;
        for (uint32_t i = s.lo + 1; i <= s.hi; i++) {
            int64_t num = a * c[i].x + b * c[i].y + cc;
            if (num < 0) num = -num;
            if (num > best_num) {
                const double d = (double)num / den;
                if (d > dmax) { dmax = d; index = i; best_num = num; }
            }
        }
        if (dmax > eps) {
            stack.push_back({index, s.hi});
            stack.push_back({s.lo, index});
        } else {
            if (nout < cap) out[nout] = c[s.lo];
            nout++;
            if (nout > cap) return nout;
        }
    }
    return nout;
}

inline int orient(Pt p, Pt q, Pt r) {
    const int32_t v = (q.y - p.y) * (r.x - q.x) - (q.x - p.x) * (r.y - q.y);
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
                     std::vector<uint32_t> &quads_out, QuadStats *stats) {
    const uint32_t mn = w < h ? w : h;
    const uint32_t min_edge_length = (uint32_t)((float)mn * cfg.min_side_length_factor);  // src/aruco.rs:55
    const float min_corner_separation = (float)mn * cfg.min_corner_separation_factor;     // src/aruco.rs:56
    const double eps_factor = cfg.contour_simplification_epsilon;

    Planes pl;
    pl.fg = bits; pl.wpr = wpr; pl.w = w; pl.h = h;
    pl.seen.assign((size_t)wpr * h, 0);
    pl.redge.assign((size_t)wpr * h, 0);

    // A surviving quad has two vertices at least sqrt(min_edge_length) apart; an 8-connected closed border that
    // reaches that far and comes back has at least sqrt(2 * min_edge_length) points.  Shorter borders cannot
    // pass the edge test (src/aruco.rs:149-159) whatever RDP does with them, so they are followed (their marks
    // matter to later starts) but not simplified.
    const size_t min_points = (size_t)floor(sqrt(2.0 * (double)min_edge_length));

    std::vector<Pt> contour;
    std::vector<uint32_t> quads;  // before discard
    QuadStats st;
    const uint32_t last_word_bits = w & 31;
    for (uint32_t y = 0; y < h; y++) {
        const uint32_t *row = bits + (size_t)y * wpr;
        for (uint32_t k = 0; k < wpr; k++) {
            uint32_t f = row[k];
            if (!f) continue;
            const uint32_t west = (f << 1) | (k ? row[k - 1] >> 31 : 0u);
            const uint32_t east = (f >> 1) | (k + 1 < wpr ? row[k + 1] << 31 : 0u);
            uint32_t outer_geom = f & ~west, hole_geom = f & ~east;
            if (k == 0) outer_geom &= ~1u;  // `x > 0`
            // `x + 1 < w`: the last column never starts a hole border
            if (k == wpr - 1) hole_geom &= ~(1u << ((last_word_bits ? last_word_bits : 32) - 1));
            uint32_t pending = outer_geom | hole_geom;
            while (pending) {
                const uint32_t b = pending & (0u - pending);
                pending ^= b;
                const int x = (int)(k * 32 + __builtin_ctz(b));
                const size_t wi = (size_t)y * wpr + k;
                int adj_dir;
                if ((outer_geom & b) && !(pl.seen[wi] & b)) adj_dir = 0;        // zero pixel to the west
                else if ((hole_geom & b) && !(pl.redge[wi] & b)) adj_dir = 4;   // zero pixel to the east
                else continue;
                follow(pl, x, (int)y, adj_dir, contour);
                st.n_contours++;
                st.n_contour_points += contour.size();
                const size_t n = contour.size();
                if (n < min_points || n < 4) continue;
                Pt q[4];
                if (simplify_closed(contour.data(), n, (double)n * eps_factor, q, 4) != 4) continue;
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
