/*
 * a3ref.c — CPU ORACLE (test infrastructure, see a3ref.h).  PARITY UNPINNED for the pixel path.
 *
 * Written to be read next to the reference: every function names the reference lines (or the
 * third-party routine called from those lines) it restates.  Straight-line and single-threaded on
 * purpose; speed is irrelevant here, order of floating-point operations is not.
 * Compile with -ffp-contract=off (see Makefile): Rust never fuses a*b+c.
 */
#define _POSIX_C_SOURCE 200809L
#include "a3ref.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* dictionary blob (aruco3_b200/data/dictionaries.bin, produced by tools/extract_dictionaries.py) */
#ifndef A3_DICT_BIN_PATH
#error "define A3_DICT_BIN_PATH"
#endif
__asm__(".section .rodata\n"
        ".balign 16\n"
        "a3ref_dict_blob:\n"
        ".incbin \"" A3_DICT_BIN_PATH "\"\n"
        "a3ref_dict_blob_end:\n"
        ".byte 0\n"
        ".previous\n");
extern const uint8_t a3ref_dict_blob[];

typedef struct {
    char name[24];
    uint8_t num_bits, tau_table;
    uint16_t reserved;
    uint32_t n_codes, first_code, reserved2;
} blob_entry;

static uint32_t blob_n_entries(void) { uint32_t v; memcpy(&v, a3ref_dict_blob + 8, 4); return v; }
static const blob_entry *blob_entries(void) { return (const blob_entry *)(a3ref_dict_blob + 16); }
static const uint64_t *blob_codes(void) {
    return (const uint64_t *)(a3ref_dict_blob + 16 + sizeof(blob_entry) * blob_n_entries());
}

int a3ref_dictionary_count(void) { return (int)blob_n_entries(); }
const char *a3ref_dictionary_name(int i) {
    if (i < 0 || (uint32_t)i >= blob_n_entries()) return NULL;
    return blob_entries()[i].name;
}

/* src/lib.rs:11-21 — bit-serial on purpose. */
uint8_t a3ref_hamming_distance(uint64_t a, uint64_t b) {
    uint64_t flipped = a ^ b;
    uint8_t count = 0;
    while (flipped > 0) {
        if (flipped % 2 == 1) count += 1;
        flipped >>= 1;
    }
    return count;
}

/* src/dictionaries.rs:129-138 */
uint8_t a3ref_calculate_tau(const uint64_t *codes, uint32_t n) {
    uint8_t tau = 255;
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = i + 1; j < n; j++) {
            uint8_t d = (uint8_t)__builtin_popcountll(codes[i] ^ codes[j]); /* == hamming_distance, KAT-checked */
            if (d < tau) tau = d;
        }
    return tau;
}

/* src/dictionaries.rs:140-145 (+116-127). Unknown name: the reference panics; we return -1. */
int a3ref_dictionary_by_name(const char *name, a3ref_dictionary *out) {
    char up[24];
    size_t n = strlen(name);
    if (n >= sizeof(up)) return -1;
    for (size_t i = 0; i <= n; i++) up[i] = (name[i] >= 'a' && name[i] <= 'z') ? (char)(name[i] - 32) : name[i];
    const blob_entry *e = blob_entries();
    for (uint32_t i = 0; i < blob_n_entries(); i++) {
        if (strncmp(e[i].name, up, sizeof(e[i].name)) == 0) {
            out->num_bits = e[i].num_bits;
            out->n_codes = e[i].n_codes;
            out->codes = blob_codes() + e[i].first_code;
            out->tau = e[i].tau_table ? e[i].tau_table : a3ref_calculate_tau(out->codes, out->n_codes);
            return 0;
        }
    }
    return -1;
}

/* src/dictionaries.rs:154-156: (num_bits as f32).sqrt().ceil() as u8 + 2 */
uint8_t a3ref_mark_size(const a3ref_dictionary *d) { return (uint8_t)((uint8_t)ceilf(sqrtf((float)d->num_bits)) + 2); }

/* src/dictionaries.rs:160-196 */
void a3ref_find_nearest(const a3ref_dictionary *d, uint64_t bits, uint64_t *index, uint8_t *dist) {
    uint64_t min_index = 0;
    uint8_t min_distance = 0xFF;
    for (uint32_t idx = 0; idx < d->n_codes; idx++) {
        uint8_t dd = a3ref_hamming_distance(d->codes[idx], bits);
        if (dd < min_distance) {
            min_distance = dd;
            min_index = idx;
        }
    }
    *index = min_index;
    *dist = min_distance;
}

/* src/dictionaries.rs:200-207 */
int a3ref_try_find_nearest(const a3ref_dictionary *d, uint64_t bits, uint64_t *index, uint8_t *dist) {
    a3ref_find_nearest(d, bits, index, dist);
    return *dist < d->tau;
}

/* src/dictionaries.rs:212-232. Returns width; bits_out gets the bools (len = what the reference pushes). */
uint32_t a3ref_make_binary_image(const a3ref_dictionary *d, uint64_t marker_id, uint8_t *bits_out, uint32_t cap) {
    uint64_t code = d->codes[marker_id];
    uint8_t width = a3ref_mark_size(d);
    uint32_t len = 0;
#define PUSH(v) do { if (len < cap) bits_out[len] = (v); len++; } while (0)
    for (uint8_t i = 0; i < width; i++) PUSH(0);
    for (uint8_t i = 0; i < d->num_bits; i++) {
        if ((uint8_t)len % width == 0) PUSH(0);
        /* `code & (1 << i)`: the literal 1 is inferred u64 in Rust */
        PUSH((code & ((uint64_t)1 << i)) != 0);
        if ((uint8_t)len % width == width - 1) PUSH(0);
    }
    for (uint8_t i = 0; i < width; i++) PUSH(0);
#undef PUSH
    return width;
}

/* ------------------------------------------------------------------------------------------ */
/* A.1  image::DynamicImage::into_luma8 (call site src/aruco.rs:60) */
void a3ref_to_luma8(const uint8_t *src, int format, uint32_t w, uint32_t h, size_t pitch, uint8_t *grey) {
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *row = src + (size_t)y * pitch;
        uint8_t *o = grey + (size_t)y * w;
        if (format == A3REF_FMT_LUMA8) {
            memcpy(o, row, w);
            continue;
        }
        if (format >= A3REF_FMT_LUMAA8) {
            /* [RECALLED, image 0.25 color.rs / traits.rs; nothing in the reference pins it] `to_luma8` = ImageBuffer::convert:
             *   Luma<u8> from LumaA<u8>: the luma channel, alpha dropped;
             *   Luma<u8> from Luma<u16> / LumaA<u16>: u8::from_primitive(l) = ((l as u32 + 128) / 257) as u8;
             *   Luma<u8> from Rgb<u16> / Rgba<u16>: u8::from_primitive(rgb_to_luma(rgb)), rgb_to_luma in u32 (u16's Larger):
             *     (2126 R + 7152 G + 722 B) / 10000, alpha ignored. */
            const uint32_t ch = format == A3REF_FMT_LUMAA8 ? 2 : format == A3REF_FMT_LUMA16 ? 1 : format == A3REF_FMT_LUMAA16 ? 2
                                : format == A3REF_FMT_RGB16 ? 3 : 4;
            for (uint32_t x = 0; x < w; x++) {
                if (format == A3REF_FMT_LUMAA8) { o[x] = row[x * 2]; continue; }
                uint16_t s[4] = {0, 0, 0, 0};
                memcpy(s, row + (size_t)x * ch * 2, (size_t)ch * 2);
                uint32_t l = (format == A3REF_FMT_LUMA16 || format == A3REF_FMT_LUMAA16)
                                 ? s[0] : (2126u * s[0] + 7152u * s[1] + 722u * s[2]) / 10000u;
                o[x] = (uint8_t)((l + 128u) / 257u);
            }
            continue;
        }
        uint32_t bpp = (format == A3REF_FMT_RGBA8 || format == A3REF_FMT_BGRA8) ? 4 : 3;
        int bgr = format == A3REF_FMT_BGR8 || format == A3REF_FMT_BGRA8; /* examples/webcam_kamera.rs:38-52: r = buffer[idx+2], b = buffer[idx+0] */
        for (uint32_t x = 0; x < w; x++) {
            uint32_t r = row[x * bpp + (bgr ? 2 : 0)], g = row[x * bpp + 1], b = row[x * bpp + (bgr ? 0 : 2)];
            uint32_t l = 2126u * r + 7152u * g + 722u * b; /* SRGB_LUMA */
            o[x] = (uint8_t)(l / 10000u);                  /* SRGB_LUMA_DIV, truncating */
        }
    }
}

/* A.2  imageproc::contrast::adaptive_threshold (call site src/aruco.rs:61) */
void a3ref_adaptive_threshold(const uint8_t *grey, uint32_t w, uint32_t h, uint32_t r, uint8_t *out) {
    /* integral_image::<_, u32>: (w+1) x (h+1), first row/column zero */
    size_t iw = (size_t)w + 1;
    uint32_t *integral = (uint32_t *)calloc(iw * ((size_t)h + 1), sizeof(uint32_t));
    for (uint32_t y = 0; y < h; y++) {
        uint32_t rowsum = 0;
        for (uint32_t x = 0; x < w; x++) {
            rowsum += grey[(size_t)y * w + x];
            integral[(size_t)(y + 1) * iw + (x + 1)] = integral[(size_t)y * iw + (x + 1)] + rowsum;
        }
    }
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            uint32_t y_low = (int32_t)y - (int32_t)r > 0 ? y - r : 0;
            uint32_t y_high = y + r < h - 1 ? y + r : h - 1;
            uint32_t x_low = (int32_t)x - (int32_t)r > 0 ? x - r : 0;
            uint32_t x_high = x + r < w - 1 ? x + r : w - 1;
            uint32_t cnt = (y_high - y_low + 1) * (x_high - x_low + 1);
            /* sum_image_pixels(integral, x_low, y_low, x_high, y_high) */
            uint32_t sum = integral[(size_t)(y_high + 1) * iw + (x_high + 1)] - integral[(size_t)y_low * iw + (x_high + 1)]
                         - integral[(size_t)(y_high + 1) * iw + x_low] + integral[(size_t)y_low * iw + x_low];
            uint32_t mean = sum / cnt;
            out[(size_t)y * w + x] = ((uint32_t)grey[(size_t)y * w + x] >= mean) ? 255 : 0;
        }
    }
    free(integral);
}

/* ------------------------------------------------------------------------------------------ */
/* A.3  imageproc::contours::find_contours::<u32> (call site src/aruco.rs:64)                  */

static const int DIFFS[8][2] = { /* VecDeque initial order: w nw n ne e se s sw (screen clockwise) */
    {-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}};

static int diff_index(int dx, int dy) {
    for (int i = 0; i < 8; i++)
        if (DIFFS[i][0] == dx && DIFFS[i][1] == dy) return i;
    return -1;
}

typedef struct {
    a3ref_point *p;
    size_t n, cap;
} pvec;
static void pvec_push(pvec *v, uint32_t x, uint32_t y) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->p = (a3ref_point *)realloc(v->p, v->cap * sizeof(a3ref_point));
    }
    v->p[v->n].x = x;
    v->p[v->n].y = y;
    v->n++;
}

static int nonzero_at(const int32_t *img, int w, int h, int x, int y) {
    return x > -1 && x < w && y > -1 && y < h && img[(size_t)y * w + x] != 0;
}

a3ref_contours *a3ref_find_contours(const uint8_t *mask, uint32_t uw, uint32_t uh) {
    const int w = (int)uw, h = (int)uh;
    int32_t *img = (int32_t *)calloc((size_t)w * h, sizeof(int32_t));
    for (size_t i = 0; i < (size_t)w * h; i++) img[i] = mask[i] > 0 ? 1 : 0; /* threshold 0 */

    pvec pts = {0};
    uint32_t *offsets = NULL;
    uint8_t *outer = NULL;
    size_t ncont = 0, ccap = 0;
    int front = 0; /* index into DIFFS of the deque's front element; the deque persists across contours */
    int curr_border_num = 1;

    for (int y = 0; y < h; y++) {
        int parent_border_num = 1;
        for (int x = 0; x < w; x++) {
            int32_t v = img[(size_t)y * w + x];
            if (v == 0) continue;
            int start = 0, adjx = 0, adjy = y, is_outer = 0;
            if (v == 1 && x > 0 && img[(size_t)y * w + x - 1] == 0) {
                start = 1; adjx = x - 1; is_outer = 1;
            } else if (v > 0 && x + 1 < w && img[(size_t)y * w + x + 1] == 0) {
                if (v > 1) parent_border_num = v;
                start = 1; adjx = x + 1; is_outer = 0;
            }
            if (start) {
                curr_border_num += 1;
                /* parent bookkeeping of the reference is unused by aruco3 (only .points is read) */
                if (ncont + 2 > ccap) {
                    ccap = ccap ? ccap * 2 : 256;
                    offsets = (uint32_t *)realloc(offsets, (ccap + 1) * sizeof(uint32_t));
                    outer = (uint8_t *)realloc(outer, ccap);
                }
                offsets[ncont] = (uint32_t)pts.n;
                outer[ncont] = (uint8_t)is_outer;

                front = diff_index(adjx - x, adjy - y); /* rotate_to_value(diffs, adj - curr) */
                int found1 = 0, p1x = 0, p1y = 0;
                for (int k = 0; k < 8; k++) { /* forward = clockwise */
                    const int *d = DIFFS[(front + k) & 7];
                    if (nonzero_at(img, w, h, x + d[0], y + d[1])) { found1 = 1; p1x = x + d[0]; p1y = y + d[1]; break; }
                }
                if (found1) {
                    int p2x = p1x, p2y = p1y, p3x = x, p3y = y;
                    for (;;) {
                        pvec_push(&pts, (uint32_t)p3x, (uint32_t)p3y);
                        front = diff_index(p2x - p3x, p2y - p3y);
                        int p4x = 0, p4y = 0;
                        for (int k = 7; k >= 0; k--) { /* .rev() = counter-clockwise, front element last */
                            const int *d = DIFFS[(front + k) & 7];
                            if (nonzero_at(img, w, h, p3x + d[0], p3y + d[1])) { p4x = p3x + d[0]; p4y = p3y + d[1]; break; }
                        }
                        int is_right_edge = 0;
                        for (int k = 7; k >= 0; k--) {
                            const int *d = DIFFS[(front + k) & 7];
                            if (d[0] == p4x - p3x && d[1] == p4y - p3y) break;
                            if (d[0] == 1 && d[1] == 0) { is_right_edge = 1; break; }
                        }
                        int32_t *cell = &img[(size_t)p3y * w + p3x];
                        if (p3x + 1 == w || is_right_edge) *cell = -curr_border_num;
                        else if (*cell == 1) *cell = curr_border_num;
                        if (p4x == x && p4y == y && p3x == p1x && p3y == p1y) break;
                        p2x = p3x; p2y = p3y; p3x = p4x; p3y = p4y;
                    }
                } else {
                    pvec_push(&pts, (uint32_t)x, (uint32_t)y);
                    img[(size_t)y * w + x] = -curr_border_num;
                }
                ncont++;
            }
            v = img[(size_t)y * w + x];
            if (v != 1) parent_border_num = v < 0 ? -v : v;
        }
        (void)parent_border_num;
    }
    free(img);
    a3ref_contours *c = (a3ref_contours *)calloc(1, sizeof(*c));
    if (!offsets) offsets = (uint32_t *)malloc(sizeof(uint32_t));
    offsets[ncont] = (uint32_t)pts.n;
    c->n_contours = (uint32_t)ncont;
    c->n_points = (uint32_t)pts.n;
    c->offsets = offsets;
    c->points = pts.p;
    c->is_outer = outer;
    return c;
}

void a3ref_contours_free(a3ref_contours *c) {
    if (!c) return;
    free(c->offsets);
    free(c->points);
    free(c->is_outer);
    free(c);
}

/* A.4  imageproc::geometry::approximate_polygon_dp (call site src/aruco.rs:133) */
static void rdp_rec(const a3ref_point *curve, size_t n, double epsilon, int closed, pvec *res) {
    double dmax = 0.0;
    size_t index = 0, end = n - 1;
    /* Line::from_points(curve[0], curve[end]) */
    double sx = (double)curve[0].x, sy = (double)curve[0].y, ex = (double)curve[end].x, ey = (double)curve[end].y;
    double a = sy - ey, b = ex - sx, c = sx * ey - ex * sy;
    for (size_t i = 1; i <= end; i++) {
        double px = (double)curve[i].x, py = (double)curve[i].y;
        double d = fabs(a * px + b * py + c) / sqrt(a * a + b * b);
        if (d > dmax) { index = i; dmax = d; }
    }
    if (dmax > epsilon) {
        rdp_rec(curve, index + 1, epsilon, 0, res);
        res->n--; /* partial1.pop() */
        rdp_rec(curve + index, end - index + 1, epsilon, 0, res);
    } else {
        pvec_push(res, curve[0].x, curve[0].y);
        pvec_push(res, curve[end].x, curve[end].y);
    }
    if (closed) res->n--;
}

size_t a3ref_approximate_polygon_dp(const a3ref_point *curve, size_t n, double epsilon, int closed, a3ref_point *out) {
    if (!(epsilon > 0.0) || n == 0) return 0; /* the reference panics on epsilon <= 0 */
    pvec res = {0};
    rdp_rec(curve, n, epsilon, closed, &res);
    memcpy(out, res.p, res.n * sizeof(a3ref_point));
    size_t k = res.n;
    free(res.p);
    return k;
}

/* A.5  imageproc::geometry::convex_hull (call site src/aruco.rs:143) */
static int orientation(int32_t px, int32_t py, int32_t qx, int32_t qy, int32_t rx, int32_t ry) {
    /* 64-bit so that frames beyond 46340 px a side are not signed-overflow UB here; identical to i32 below that */
    int64_t val = (int64_t)(qy - py) * (rx - qx) - (int64_t)(qx - px) * (ry - qy);
    return val == 0 ? 0 : (val > 0 ? 1 /* Clockwise */ : -1 /* CounterClockwise */);
}
static double pdist(a3ref_point p, a3ref_point q) {
    double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y;
    return sqrt(dx * dx + dy * dy);
}
/* comparator of the sort_by closure: returns <0 Less, >0 Greater (never Equal) */
static int hull_cmp(a3ref_point s, a3ref_point a, a3ref_point b) {
    int o = orientation((int32_t)s.x, (int32_t)s.y, (int32_t)a.x, (int32_t)a.y, (int32_t)b.x, (int32_t)b.y);
    if (o == 0) return pdist(s, a) < pdist(s, b) ? -1 : 1;
    return o > 0 ? 1 : -1;
}
size_t a3ref_convex_hull(const a3ref_point *in, size_t n, a3ref_point *out) {
    if (n == 0) return 0;
    a3ref_point *pts = (a3ref_point *)malloc(n * sizeof(*pts));
    memcpy(pts, in, n * sizeof(*pts));
    size_t sp = 0;
    a3ref_point start = pts[0];
    for (size_t i = 1; i < n; i++)
        if (pts[i].y < start.y || (pts[i].y == start.y && pts[i].x < start.x)) { sp = i; start = pts[i]; }
    a3ref_point t = pts[0]; pts[0] = pts[sp]; pts[sp] = t; /* swap(0, pos) */
    memmove(pts, pts + 1, (n - 1) * sizeof(*pts));         /* remove(0) */
    size_t m = n - 1;
    /* slice::sort_by — insertion sort, which is what std uses for len <= 20 (we only ever sort 3 points) */
    for (size_t i = 1; i < m; i++) {
        a3ref_point key = pts[i];
        size_t j = i;
        while (j > 0 && hull_cmp(start, key, pts[j - 1]) < 0) { pts[j] = pts[j - 1]; j--; }
        pts[j] = key;
    }
    /* keep the farthest of each run collinear with start */
    a3ref_point *rem = (a3ref_point *)malloc((m + 1) * sizeof(*rem));
    size_t nr = 0, i = 0;
    while (i < m) {
        a3ref_point p = pts[i++];
        while (i < m && orientation((int32_t)start.x, (int32_t)start.y, (int32_t)p.x, (int32_t)p.y, (int32_t)pts[i].x, (int32_t)pts[i].y) == 0)
            p = pts[i++];
        rem[nr++] = p;
    }
    size_t ns = 0;
    out[ns++] = start;
    for (size_t k = 0; k < nr; k++) {
        while (ns > 1 && orientation((int32_t)out[ns - 2].x, (int32_t)out[ns - 2].y, (int32_t)out[ns - 1].x, (int32_t)out[ns - 1].y,
                                     (int32_t)rem[k].x, (int32_t)rem[k].y) != -1)
            ns--;
        out[ns++] = rem[k];
    }
    free(pts);
    free(rem);
    return ns;
}

/* src/aruco.rs:124-166 */
uint32_t a3ref_contours_to_candidates(const a3ref_contours *c, uint32_t min_edge_length, double eps,
                                      uint32_t **quads_out, a3ref_stats *stats) {
    uint32_t *quads = NULL;
    size_t nq = 0, cap = 0;
    size_t maxlen = 1;
    for (uint32_t i = 0; i < c->n_contours; i++)
        if (c->offsets[i + 1] - c->offsets[i] > maxlen) maxlen = c->offsets[i + 1] - c->offsets[i];
    a3ref_point *edges = (a3ref_point *)malloc((maxlen + 2) * sizeof(*edges));
    for (uint32_t i = 0; i < c->n_contours; i++) {
        size_t len = c->offsets[i + 1] - c->offsets[i];
        size_t ne = a3ref_approximate_polygon_dp(c->points + c->offsets[i], len, (double)len * eps, 1, edges);
        if (ne != 4) { if (stats) stats->reject_point_count++; continue; }
        a3ref_point hull[4];
        size_t nh = a3ref_convex_hull(edges, 4, hull);
        if (nh != 4) { if (stats) stats->reject_convexity++; continue; }
        uint32_t cmin = min_edge_length + 1;
        for (int k = 0; k < 4; k++) {
            int j = (k + 1) % 4;
            int32_t dx = (int32_t)hull[k].x - (int32_t)hull[j].x;
            int32_t dy = (int32_t)hull[k].y - (int32_t)hull[j].y;
            uint32_t e = (uint32_t)(dx * dx + dy * dy);
            if (e < cmin) cmin = e;
        }
        if (cmin < min_edge_length) { if (stats) stats->reject_edge_length++; continue; } /* squared vs unsquared: Q1 */
        if (nq == cap) { cap = cap ? cap * 2 : 64; quads = (uint32_t *)realloc(quads, cap * 8 * sizeof(uint32_t)); }
        for (int k = 0; k < 4; k++) { quads[nq * 8 + 2 * k] = hull[k].x; quads[nq * 8 + 2 * k + 1] = hull[k].y; }
        nq++;
    }
    free(edges);
    if (!quads) quads = (uint32_t *)malloc(8 * sizeof(uint32_t));
    *quads_out = quads;
    return (uint32_t)nq;
}

/* src/aruco.rs:168-185 */
void a3ref_enforce_clockwise_corners(uint32_t *q, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        uint32_t *p = q + i * 8;
        int32_t dx1 = (int32_t)p[2] - (int32_t)p[0], dy1 = (int32_t)p[3] - (int32_t)p[1];
        int32_t dx2 = (int32_t)p[4] - (int32_t)p[0], dy2 = (int32_t)p[5] - (int32_t)p[1];
        if (dx1 * dy2 - dy1 * dx2 < 0) {
            uint32_t sx = p[2], sy = p[3];
            p[2] = p[6]; p[3] = p[7];
            p[6] = sx; p[7] = sy;
        }
    }
}

/* src/aruco.rs:328-338 */
float a3ref_perimeter(const uint32_t *q) {
    float p = 0.0f;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        float dx = (float)q[2 * i] - (float)q[2 * j];
        float dy = (float)q[2 * i + 1] - (float)q[2 * j + 1];
        p += sqrtf((dx * dx) + (dy * dy));
    }
    return p;
}

/* src/aruco.rs:187-232 */
uint32_t a3ref_discard_too_near(uint32_t *q, uint32_t n, float min_distance) {
    if (n == 0) return 0;
    uint8_t *dead = (uint8_t *)calloc(n, 1);
    for (uint32_t i = 0; i + 1 < n; i++) {
        if (dead[i]) continue;
        float perimeter_i = a3ref_perimeter(q + i * 8);
        for (uint32_t j = i + 1; j < n; j++) {
            if (dead[j]) continue;
            float distance = 0.0f;
            for (int k = 0; k < 4; k++) {
                float dx = (float)q[i * 8 + 2 * k] - (float)q[j * 8 + 2 * k];
                float dy = (float)q[i * 8 + 2 * k + 1] - (float)q[j * 8 + 2 * k + 1];
                distance += sqrtf((dx * dx) + (dy * dy));
            }
            if ((distance / 4.0f) < min_distance) {
                float perimeter_j = a3ref_perimeter(q + j * 8);
                if (dead[i] || dead[j]) {
                    /* nothing */
                } else if (perimeter_i >= perimeter_j) {
                    dead[j] = 1;
                } else {
                    dead[i] = 1; /* note: the i-loop keeps running for this i, as in the reference */
                }
            }
        }
    }
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++)
        if (!dead[i]) { if (k != i) memcpy(q + k * 8, q + i * 8, 8 * sizeof(uint32_t)); k++; }
    free(dead);
    return k;
}

/* ------------------------------------------------------------------------------------------ */
/* A.6  imageproc Projection::from_control_points (call site src/aruco.rs:244-247)
 * The 8x8 system is the reference's; it solves it with nalgebra's f64 SVD, we use f64 Gaussian
 * elimination with partial pivoting (results agree to ~1e-12 relative before the f32 cast).
 * The CUDA decode kernel performs the very same sequence of f64 operations. */
static int solve8(double a[8][9]) {
    for (int col = 0; col < 8; col++) {
        int piv = col;
        double best = fabs(a[col][col]);
        for (int r = col + 1; r < 8; r++) {
            double v = fabs(a[r][col]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0) return 0;
        if (piv != col)
            for (int k = 0; k < 9; k++) { double t = a[col][k]; a[col][k] = a[piv][k]; a[piv][k] = t; }
        for (int r = col + 1; r < 8; r++) {
            double f = a[r][col] / a[col][col];
            for (int k = col; k < 9; k++) a[r][k] = a[r][k] - f * a[col][k];
        }
    }
    for (int r = 7; r >= 0; r--) {
        double s = a[r][8];
        for (int k = r + 1; k < 8; k++) s = s - a[r][k] * a[k][8];
        a[r][8] = s / a[r][r];
    }
    return 1;
}

static int try_inverse(const float t[9], float inv[9]) {
    float t00 = t[0], t01 = t[1], t02 = t[2], t10 = t[3], t11 = t[4], t12 = t[5], t20 = t[6], t21 = t[7], t22 = t[8];
    float m00 = t11 * t22 - t12 * t21;
    float m01 = t10 * t22 - t12 * t20;
    float m02 = t10 * t21 - t11 * t20;
    float det = t00 * m00 - t01 * m01 + t02 * m02;
    if (fabsf(det) < 1e-10f) return 0;
    float m10 = t01 * t22 - t02 * t21;
    float m11 = t00 * t22 - t02 * t20;
    float m12 = t00 * t21 - t01 * t20;
    float m20 = t01 * t12 - t02 * t11;
    float m21 = t00 * t12 - t02 * t10;
    float m22 = t00 * t11 - t01 * t10;
    float r[9] = {m00 / det, -m10 / det, m20 / det, -m01 / det, m11 / det, -m21 / det, m02 / det, -m12 / det, m22 / det};
    /* normalize(inv): every entry divided by inv[8] (upstream detail recalled, not verifiable here) */
    float s = r[8];
    for (int i = 0; i < 8; i++) inv[i] = r[i] / s;
    inv[8] = 1.0f;
    return 1;
}

/* cls: 0 Translation, 1 Affine, 2 Projection (class_from_matrix on the forward matrix) */
int a3ref_projection_from_control_points(const float from[8], const float to[8], float transform[9],
                                         float inverse[9], int *cls) {
    double a[8][9];
    for (int k = 0; k < 4; k++) {
        double xf = (double)from[2 * k], yf = (double)from[2 * k + 1];
        double x = (double)to[2 * k], y = (double)to[2 * k + 1];
        double r0[9] = {0.0, 0.0, 0.0, -xf, -yf, -1.0, y * xf, y * yf, -y};
        double r1[9] = {xf, yf, 1.0, 0.0, 0.0, 0.0, -x * xf, -x * yf, x};
        memcpy(a[2 * k], r0, sizeof(r0));
        memcpy(a[2 * k + 1], r1, sizeof(r1));
    }
    if (!solve8(a)) return 0;
    for (int i = 0; i < 8; i++) transform[i] = (float)a[i][8];
    transform[8] = 1.0f;
    for (int i = 0; i < 8; i++)
        if (!isfinite(transform[i])) return 0;
    int c = 2;
    if (fabsf(transform[6]) < 1e-10f && fabsf(transform[7]) < 1e-10f && fabsf(transform[8] - 1.0f) < 1e-10f) {
        if (fabsf(transform[0] - 1.0f) < 1e-10f && fabsf(transform[1]) < 1e-10f && fabsf(transform[3]) < 1e-10f &&
            fabsf(transform[4] - 1.0f) < 1e-10f)
            c = 0;
        else
            c = 1;
    }
    *cls = c;
    return try_inverse(transform, inverse);
}

static uint8_t clamp_u8_trunc(float x) { /* <u8 as Clamp<f32>>::clamp */
    if (x < 255.0f) {
        if (x > 0.0f) return (uint8_t)x;
        return 0;
    }
    return 255;
}

/* The second half of from_control_points (f32 transform -> class + inverse) and warp_into, for a GIVEN forward transform:
 * what a3ref_extract_homography does after its own solve.  Exposed so that tests/test_oracle_risk.py can feed it the
 * coefficients of an independent solver and switch the bilinear variant (R4):
 *   variant 0  the restated imageproc 0.25 blend_bilinear: the two horizontal blends are clamped and truncated to u8 before the
 *              vertical blend (what the product implements)
 *   variant 1  one stage: the three blends in f32, one clamp + truncation at the end
 *   variant 2  one stage, rounded to nearest at the end
 * Returns 1 when the projection exists. */
int a3ref_warp_with_transform(const uint8_t *grey, uint32_t w, uint32_t h, const float transform[9], uint32_t size, int variant,
                              uint8_t *patch) {
    memset(patch, 0, (size_t)size * size);
    for (int i = 0; i < 8; i++)
        if (!isfinite(transform[i])) return 0;
    int cls = 2;
    if (fabsf(transform[6]) < 1e-10f && fabsf(transform[7]) < 1e-10f && fabsf(transform[8] - 1.0f) < 1e-10f) {
        if (fabsf(transform[0] - 1.0f) < 1e-10f && fabsf(transform[1]) < 1e-10f && fabsf(transform[3]) < 1e-10f &&
            fabsf(transform[4] - 1.0f) < 1e-10f)
            cls = 0;
        else
            cls = 1;
    }
    float inv[9];
    if (!try_inverse(transform, inv)) return 0;
    const float *t = inv; /* projection.invert(): maps output pixels back into the image */
    for (uint32_t oy = 0; oy < size; oy++) {
        for (uint32_t ox = 0; ox < size; ox++) {
            float x = (float)ox, y = (float)oy, px, py;
            if (cls == 2) {
                float d = t[6] * x + t[7] * y + t[8];
                px = (t[0] * x + t[1] * y + t[2]) / d;
                py = (t[3] * x + t[4] * y + t[5]) / d;
            } else if (cls == 1) {
                px = t[0] * x + t[1] * y + t[2];
                py = t[3] * x + t[4] * y + t[5];
            } else {
                px = x + t[2];
                py = y + t[5];
            }
            /* interpolate_bilinear */
            float left = floorf(px), right = left + 1.0f, top = floorf(py), bottom = top + 1.0f;
            float rw = px - left, bw = py - top;
            uint8_t o = 0;
            if (!(left < 0.0f || right >= (float)w || top < 0.0f || bottom >= (float)h)) {
                /* NaN coordinates: every comparison above is false in Rust too, and `NaN as u32` is 0 */
                uint32_t l = isnan(left) ? 0 : (uint32_t)left, r = isnan(right) ? 0 : (uint32_t)right;
                uint32_t tp = isnan(top) ? 0 : (uint32_t)top, b = isnan(bottom) ? 0 : (uint32_t)bottom;
                float tl = grey[(size_t)tp * w + l], tr = grey[(size_t)tp * w + r];
                float bl = grey[(size_t)b * w + l], br = grey[(size_t)b * w + r];
                if (variant == 0) {
                    uint8_t topv = clamp_u8_trunc((1.0f - rw) * tl + rw * tr);
                    uint8_t botv = clamp_u8_trunc((1.0f - rw) * bl + rw * br);
                    o = clamp_u8_trunc((1.0f - bw) * (float)topv + bw * (float)botv);
                } else {
                    float topf = (1.0f - rw) * tl + rw * tr, botf = (1.0f - rw) * bl + rw * br;
                    float v = (1.0f - bw) * topf + bw * botf;
                    o = variant == 1 ? clamp_u8_trunc(v) : clamp_u8_trunc(v + 0.5f);
                }
            }
            patch[(size_t)oy * size + ox] = o;
        }
    }
    return 1;
}

/* A.7  warp_into(grey, projection, Bilinear, Luma([0]), out) (call site src/aruco.rs:253)
 * Returns 1 when the projection exists, 0 when the reference pushes GrayImage::new(1,1) (Q5). */
int a3ref_extract_homography(const uint8_t *grey, uint32_t w, uint32_t h, const uint32_t quad[8],
                             uint32_t size, uint8_t *patch) {
    float hs = (float)size;
    float from[8], to[8] = {0.0f, 0.0f, hs, 0.0f, hs, hs, 0.0f, hs};
    for (int i = 0; i < 8; i++) from[i] = (float)quad[i];
    float fwd[9], inv[9];
    int cls;
    memset(patch, 0, (size_t)size * size);
    if (!a3ref_projection_from_control_points(from, to, fwd, inv, &cls)) return 0;
    return a3ref_warp_with_transform(grey, w, h, fwd, size, 0, patch);
}

/* A.8  imageproc::contrast::otsu_level (call site src/aruco.rs:264) */
uint8_t a3ref_otsu_level(const uint8_t *img, uint32_t w, uint32_t h) {
    uint32_t hist[256] = {0};
    for (size_t i = 0; i < (size_t)w * h; i++) hist[img[i]]++;
    uint32_t total_weight = w * h;
    double total_pixel_sum = 0.0;
    for (uint32_t t = 0; t < 256; t++) total_pixel_sum = total_pixel_sum + (double)(t * hist[t]);
    double background_pixel_sum = 0.0;
    uint32_t background_weight = 0, foreground_weight;
    double largest_variance = 0.0;
    uint8_t best_threshold = 0;
    for (uint32_t t = 0; t < 256; t++) {
        background_weight += hist[t];
        if (background_weight == 0) continue;
        foreground_weight = total_weight - background_weight;
        if (foreground_weight == 0) break;
        background_pixel_sum += (double)(t * hist[t]);
        double foreground_pixel_sum = total_pixel_sum - background_pixel_sum;
        double background_mean = background_pixel_sum / (double)background_weight;
        double foreground_mean = foreground_pixel_sum / (double)foreground_weight;
        double diff = background_mean - foreground_mean;
        double mean_diff_squared = diff * diff; /* powi(2) */
        double variance = (double)background_weight * (double)foreground_weight * mean_diff_squared;
        if (variance > largest_variance) { largest_variance = variance; best_threshold = (uint8_t)t; }
    }
    return best_threshold;
}

/* A.10  image::imageops::resize(_, dw, dh, FilterType::Triangle) (call site src/aruco.rs:273) */
static float triangle_kernel(float x) { return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f; }

typedef struct { uint32_t left, count; float *w; } taps;
static taps make_taps(uint32_t n_in, uint32_t n_out, uint32_t o) {
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float src_support = 1.0f * sratio;
    float input = ((float)o + 0.5f) * ratio;
    int64_t left = (int64_t)floorf(input - src_support);
    if (left < 0) left = 0;
    if (left > (int64_t)n_in - 1) left = (int64_t)n_in - 1;
    int64_t right = (int64_t)ceilf(input + src_support);
    if (right < left + 1) right = left + 1;
    if (right > (int64_t)n_in) right = (int64_t)n_in;
    input = input - 0.5f;
    taps t;
    t.left = (uint32_t)left;
    t.count = (uint32_t)(right - left);
    t.w = (float *)malloc(t.count * sizeof(float));
    float sum = 0.0f;
    for (int64_t i = left; i < right; i++) {
        float wv = triangle_kernel(((float)i - input) / sratio);
        t.w[i - left] = wv;
        sum += wv;
    }
    for (uint32_t i = 0; i < t.count; i++) t.w[i] /= sum;
    return t;
}

void a3ref_resize_triangle(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh, uint8_t *dst) {
    if (sw == dw && sh == dh) { memcpy(dst, src, (size_t)sw * sh); return; }
    float *tmp = (float *)malloc((size_t)sw * dh * sizeof(float)); /* vertical_sample -> f32 image, sw x dh */
    for (uint32_t oy = 0; oy < dh; oy++) {
        taps t = make_taps(sh, dh, oy);
        for (uint32_t x = 0; x < sw; x++) {
            float acc = 0.0f;
            for (uint32_t i = 0; i < t.count; i++) acc += (float)src[(size_t)(t.left + i) * sw + x] * t.w[i];
            tmp[(size_t)oy * sw + x] = acc;
        }
        free(t.w);
    }
    for (uint32_t ox = 0; ox < dw; ox++) { /* horizontal_sample -> u8 */
        taps t = make_taps(sw, dw, ox);
        for (uint32_t y = 0; y < dh; y++) {
            float acc = 0.0f;
            for (uint32_t i = 0; i < t.count; i++) acc += tmp[(size_t)y * sw + t.left + i] * t.w[i];
            float c = acc < 0.0f ? 0.0f : (acc > 255.0f ? 255.0f : acc); /* clamp(t, min, max) */
            dst[(size_t)y * dw + ox] = (uint8_t)roundf(c);               /* FloatNearest: round half away */
        }
        free(t.w);
    }
    free(tmp);
}

/* src/aruco.rs:315-326 — 90 degrees counter-clockwise: new[i][j] = old[j][W-1-i] */
void a3ref_rotate_bit_matrix(const uint8_t *in, uint32_t rows, uint32_t cols, uint8_t *out) {
    uint32_t r = 0;
    for (int32_t x = (int32_t)cols - 1; x >= 0; x--, r++)
        for (uint32_t y = 0; y < rows; y++) out[r * rows + y] = in[y * cols + (uint32_t)x];
}

static uint64_t rotl64(uint64_t v, unsigned s) { return (v << (s & 63)) | (v >> ((64 - s) & 63)); }
static uint64_t rotr64(uint64_t v, unsigned s) { return (v >> (s & 63)) | (v << ((64 - s) & 63)); }

/* src/aruco.rs:263-313.  Returns 1 for Some(codes), 0 for None. */
int a3ref_homography_to_code_permutations(const uint8_t *patch, uint32_t pw, uint32_t ph, uint8_t mark_size,
                                          uint64_t codes[4], uint8_t *otsu_out, uint8_t *reduced_out) {
    uint32_t ms = mark_size;
    uint8_t level = a3ref_otsu_level(patch, pw, ph);
    if (otsu_out) *otsu_out = level;
    uint8_t *bin = (uint8_t *)malloc((size_t)pw * ph);
    for (size_t i = 0; i < (size_t)pw * ph; i++) bin[i] = patch[i] > level ? 255 : 0; /* A.9 ThresholdType::Binary */
    uint8_t *reduced = (uint8_t *)malloc((size_t)ms * ms);
    a3ref_resize_triangle(bin, pw, ph, ms, ms, reduced);
    if (reduced_out) memcpy(reduced_out, reduced, (size_t)ms * ms);
    uint8_t *bits = (uint8_t *)malloc((size_t)ms * ms), *rot = (uint8_t *)malloc((size_t)ms * ms);
    for (uint32_t i = 0; i < ms * ms; i++) bits[i] = reduced[i] > 127; /* row-major, bits[y*ms+x] */
    int ok = 1;
    uint32_t end = ms ? ms - 1 : 0;
    for (uint32_t i = 0; i < ms && ok; i++) {
        if (bits[i * ms + 0] || bits[i * ms + end]) ok = 0;
        else if (bits[0 * ms + i] || bits[end * ms + i]) ok = 0;
    }
    if (ok) {
        for (int r = 0; r < 4; r++) {
            uint64_t b = 0;
            for (uint32_t y = 1; y + 1 < ms; y++)
                for (uint32_t x = 1; x + 1 < ms; x++) {
                    if (bits[y * ms + x]) b |= 1;
                    b = rotl64(b, 1);
                }
            b = rotr64(b, 1);
            codes[r] = b;
            a3ref_rotate_bit_matrix(bits, ms, ms, rot);
            memcpy(bits, rot, (size_t)ms * ms);
        }
    }
    free(bin); free(reduced); free(bits); free(rot);
    return ok;
}

/* ------------------------------------------------------------------------------------------ */
void a3ref_default_config(a3ref_config *c) { /* src/aruco.rs:32-43 */
    c->threshold_window = 7;
    c->contour_simplification_epsilon = 0.05;
    c->min_side_length_factor = 0.2f;
    c->min_corner_separation_factor = 0.1f;
    c->homography_sample_size = 49;
    c->filter_high_bit_errors = 1;
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* src/aruco.rs:52-121 */
a3ref_detection *a3ref_detect(const a3ref_config *cfg, const a3ref_dictionary *dict, const uint8_t *image,
                              int format, uint32_t w, uint32_t h, size_t pitch) {
    a3ref_detection *det = (a3ref_detection *)calloc(1, sizeof(*det));
    double t0 = now_ms(), t1;
    det->width = w;
    det->height = h;
    uint32_t mn = w < h ? w : h;
    uint32_t min_edge_length = (uint32_t)((float)mn * cfg->min_side_length_factor);
    float min_corner_separation = (float)mn * cfg->min_corner_separation_factor;

    det->grey = (uint8_t *)malloc((size_t)w * h);
    a3ref_to_luma8(image, format, w, h, pitch, det->grey);
    t1 = now_ms(); det->stats.ms_gray = t1 - t0;
    det->mask = (uint8_t *)malloc((size_t)w * h);
    a3ref_adaptive_threshold(det->grey, w, h, cfg->threshold_window, det->mask);
    double t2 = now_ms(); det->stats.ms_threshold = t2 - t1;
    a3ref_contours *contours = a3ref_find_contours(det->mask, w, h);
    double t3 = now_ms(); det->stats.ms_contours = t3 - t2;
    det->stats.n_contours = contours->n_contours;
    det->stats.n_contour_points = contours->n_points;

    uint32_t *quads = NULL;
    uint32_t nq = a3ref_contours_to_candidates(contours, min_edge_length, cfg->contour_simplification_epsilon, &quads, &det->stats);
    a3ref_contours_free(contours);
    det->stats.n_candidates_before_discard = nq;
    a3ref_enforce_clockwise_corners(quads, nq);
    nq = a3ref_discard_too_near(quads, nq, min_corner_separation);
    double t4 = now_ms(); det->stats.ms_quads = t4 - t3;
    det->n_candidates = nq;
    det->candidates = quads;
    det->stats.n_candidates = nq;

    uint32_t hs = cfg->homography_sample_size;
    det->patch_size = hs;
    size_t psz = (size_t)hs * hs;
    det->homographies = (uint8_t *)calloc(nq ? nq * psz : 1, 1);
    det->homography_ok = (uint8_t *)calloc(nq ? nq : 1, 1);
    det->otsu = (uint8_t *)calloc(nq ? nq : 1, 1);
    det->has_codes = (uint8_t *)calloc(nq ? nq : 1, 1);
    det->codes = (uint64_t *)calloc(nq ? nq * 4 : 1, sizeof(uint64_t));
    det->markers = (a3ref_marker *)calloc(nq ? nq : 1, sizeof(a3ref_marker));
    for (uint32_t i = 0; i < nq; i++)
        det->homography_ok[i] = (uint8_t)a3ref_extract_homography(det->grey, w, h, quads + i * 8, hs, det->homographies + i * psz);
    double t5 = now_ms(); det->stats.ms_warp = t5 - t4;

    uint8_t mark_size = a3ref_mark_size(dict);
    for (uint32_t i = 0; i < nq; i++) {
        uint64_t codes[4] = {0, 0, 0, 0};
        int some;
        if (det->homography_ok[i]) {
            some = a3ref_homography_to_code_permutations(det->homographies + i * psz, hs, hs, mark_size, codes, &det->otsu[i], NULL);
        } else {
            uint8_t zero = 0; /* GrayImage::new(1, 1), src/aruco.rs:256 */
            some = a3ref_homography_to_code_permutations(&zero, 1, 1, mark_size, codes, &det->otsu[i], NULL);
        }
        det->has_codes[i] = (uint8_t)some;
        memcpy(det->codes + i * 4, codes, sizeof(codes));
        int found_any = 0;
        uint32_t min_code_distance = 0x7FFFFFFF;
        uint64_t min_code = 0x7FFFFFFF, min_code_id = 0x7FFFFFFF;
        uint32_t min_rotation = 0;
        if (some) {
            det->stats.n_border_pass++;
            for (uint32_t r = 0; r < 4; r++) {
                uint64_t nearest_id; uint8_t nearest_dist;
                a3ref_find_nearest(dict, codes[r], &nearest_id, &nearest_dist);
                if ((uint32_t)nearest_dist < min_code_distance) {
                    min_code = codes[r];
                    min_code_distance = nearest_dist;
                    min_code_id = nearest_id;
                    min_rotation = r;
                    found_any = 1;
                }
            }
        }
        if (found_any && (!cfg->filter_high_bit_errors || min_code_distance < (uint32_t)dict->tau)) {
            a3ref_marker *m = &det->markers[det->n_markers++];
            m->id = min_code_id;
            m->code = min_code;
            m->hamming_distance = (uint8_t)min_code_distance;
            m->rotation = (uint8_t)min_rotation;
            m->candidate = i;
            for (uint32_t k = 0; k < 4; k++) { /* corners.rotate_left(min_rotation) */
                uint32_t s = (k + min_rotation) % 4;
                m->corners[2 * k] = quads[i * 8 + 2 * s];
                m->corners[2 * k + 1] = quads[i * 8 + 2 * s + 1];
            }
        }
    }
    double t6 = now_ms();
    det->stats.ms_decode = t6 - t5;
    det->stats.ms_total = t6 - t0;
    det->stats.n_markers = det->n_markers;
    return det;
}

void a3ref_detection_free(a3ref_detection *d) {
    if (!d) return;
    free(d->grey); free(d->mask); free(d->candidates); free(d->homographies); free(d->homography_ok);
    free(d->otsu); free(d->has_codes); free(d->codes); free(d->markers);
    free(d);
}

/* ---- frame-parallel driver for the CPU baseline (not part of the restatement) ---- */
typedef struct {
    const a3ref_config *cfg; const a3ref_dictionary *dict; const uint8_t *frames;
    int format; uint32_t n, w, h; size_t pitch, frame_stride;
    uint32_t *next; pthread_mutex_t *mu; uint64_t markers; a3ref_stats sum;
} many_job;

static void *many_worker(void *arg) {
    many_job *j = (many_job *)arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        uint32_t i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n) break;
        a3ref_detection *d = a3ref_detect(j->cfg, j->dict, j->frames + (size_t)i * j->frame_stride, j->format, j->w, j->h, j->pitch);
        j->markers += d->n_markers;
        j->sum.n_contours += d->stats.n_contours; j->sum.n_contour_points += d->stats.n_contour_points;
        j->sum.n_candidates += d->stats.n_candidates; j->sum.n_markers += d->stats.n_markers;
        j->sum.ms_gray += d->stats.ms_gray; j->sum.ms_threshold += d->stats.ms_threshold;
        j->sum.ms_contours += d->stats.ms_contours; j->sum.ms_quads += d->stats.ms_quads;
        j->sum.ms_warp += d->stats.ms_warp; j->sum.ms_decode += d->stats.ms_decode; j->sum.ms_total += d->stats.ms_total;
        a3ref_detection_free(d);
    }
    return NULL;
}

uint64_t a3ref_detect_many(const a3ref_config *cfg, const a3ref_dictionary *dict, const uint8_t *frames,
                           int format, uint32_t n, uint32_t w, uint32_t h, size_t pitch, size_t frame_stride,
                           uint32_t threads, a3ref_stats *sum_stats) {
    if (threads == 0) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    many_job jobs[256];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    uint32_t next = 0;
    for (uint32_t t = 0; t < threads; t++) {
        many_job j = {cfg, dict, frames, format, n, w, h, pitch, frame_stride, &next, &mu, 0, {0}};
        jobs[t] = j;
        pthread_create(&th[t], NULL, many_worker, &jobs[t]);
    }
    uint64_t total = 0;
    a3ref_stats s;
    memset(&s, 0, sizeof(s));
    for (uint32_t t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].markers;
        s.n_contours += jobs[t].sum.n_contours; s.n_contour_points += jobs[t].sum.n_contour_points;
        s.n_candidates += jobs[t].sum.n_candidates; s.n_markers += jobs[t].sum.n_markers;
        s.ms_gray += jobs[t].sum.ms_gray; s.ms_threshold += jobs[t].sum.ms_threshold;
        s.ms_contours += jobs[t].sum.ms_contours; s.ms_quads += jobs[t].sum.ms_quads;
        s.ms_warp += jobs[t].sum.ms_warp; s.ms_decode += jobs[t].sum.ms_decode; s.ms_total += jobs[t].sum.ms_total;
    }
    if (sum_stats) *sum_stats = s;
    return total;
}
