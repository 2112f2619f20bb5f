/*
 * a3ref — CPU ORACLE for the aruco3 detection path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's `Detector::detect`
 * (/root/reference/src/aruco.rs:52-121) and of every function it calls, including the
 * third-party `image 0.25` / `imageproc 0.25` routines whose source is NOT in the reference
 * tree (each function below cites the call site it restates).
 *
 * It exists to CHECK the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Nothing under aruco3_b200/ links,
 * imports or calls it, and the product never falls back to it.
 *
 * PARITY STATUS: the reference's own unit tests pin hamming_distance, find_nearest,
 * try_find_nearest, tau, rotate_bit_matrix, enforce_clockwise_corners and discard_too_near
 * (restated in tests/test_oracle_kat.py and green).  The pixel path (luma, adaptive threshold,
 * contours, RDP, hull, projection, warp, Otsu, resize) is pinned by NO test or fixture of the
 * reference and the reference cannot be built here (no Rust toolchain, crates not vendored):
 * for those stages this oracle is "PARITY UNPINNED" against upstream and is cross-checked only
 * against independent numpy / OpenCV models where the semantics provably coincide.
 */
#ifndef A3REF_H
#define A3REF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* BGR8 / BGRA8: the host swizzle of examples/webcam_kamera.rs:38-52 followed by into_luma8 */
enum { A3REF_FMT_RGB8 = 0, A3REF_FMT_RGBA8 = 1, A3REF_FMT_LUMA8 = 2, A3REF_FMT_BGR8 = 3, A3REF_FMT_BGRA8 = 4,
       /* the other integer DynamicImage variants; 16-bit subpixels are native-endian u16 */
       A3REF_FMT_LUMAA8 = 5, A3REF_FMT_LUMA16 = 6, A3REF_FMT_LUMAA16 = 7, A3REF_FMT_RGB16 = 8, A3REF_FMT_RGBA16 = 9 };

/* src/aruco.rs:23-30 (DetectorConfig), same field order. */
typedef struct a3ref_config {
    uint32_t threshold_window;
    double contour_simplification_epsilon;
    float min_side_length_factor;
    float min_corner_separation_factor;
    uint32_t homography_sample_size;
    uint8_t filter_high_bit_errors;
} a3ref_config;

/* src/dictionaries.rs:22-28 (ARDictionary). tau is the *effective* tau (dictionaries.rs:124). */
typedef struct a3ref_dictionary {
    uint8_t num_bits;
    uint8_t tau;
    uint32_t n_codes;
    const uint64_t *codes;
} a3ref_dictionary;

typedef struct a3ref_point {
    uint32_t x, y;
} a3ref_point;

/* src/aruco.rs:8-13 (Marker) + bookkeeping (candidate index, winning rotation). */
typedef struct a3ref_marker {
    uint64_t id;
    uint64_t code;
    uint32_t corners[8]; /* x0,y0,...,x3,y3 after rotate_left(rotation) */
    uint32_t candidate;
    uint8_t hamming_distance;
    uint8_t rotation;
    uint8_t pad[2];
} a3ref_marker;

typedef struct a3ref_contours {
    uint32_t n_contours;
    uint32_t n_points;
    uint32_t *offsets;   /* n_contours + 1 */
    a3ref_point *points; /* n_points */
    uint8_t *is_outer;   /* n_contours: 1 = BorderType::Outer, 0 = Hole */
} a3ref_contours;

typedef struct a3ref_stats {
    uint32_t n_contours, n_contour_points;
    uint32_t reject_point_count, reject_convexity, reject_edge_length;
    uint32_t n_candidates_before_discard, n_candidates, n_border_pass, n_markers;
    double ms_gray, ms_threshold, ms_contours, ms_quads, ms_warp, ms_decode, ms_total;
} a3ref_stats;

/* src/aruco.rs:16-21 (Detection) with every intermediate exposed for stage-by-stage diffs. */
typedef struct a3ref_detection {
    uint32_t width, height;
    uint8_t *grey;          /* width*height */
    uint8_t *mask;          /* width*height, the thresholded image (not kept by the reference) */
    uint32_t n_candidates;
    uint32_t *candidates;   /* n_candidates * 8 (x0,y0..x3,y3) */
    uint32_t patch_size;    /* homography_sample_size */
    uint8_t *homographies;  /* n_candidates * patch_size^2; a failed projection leaves zeros */
    uint8_t *homography_ok; /* n_candidates: 0 => the reference pushed a 1x1 image (Q5) */
    uint8_t *otsu;          /* n_candidates */
    uint8_t *has_codes;     /* n_candidates: Some/None of homography_to_code_permutations */
    uint64_t *codes;        /* n_candidates * 4 */
    uint32_t n_markers;
    a3ref_marker *markers;
    a3ref_stats stats;
} a3ref_detection;

/* ---- dictionaries (src/dictionaries.rs) ---- */
int a3ref_dictionary_count(void);
const char *a3ref_dictionary_name(int index);
int a3ref_dictionary_by_name(const char *name, a3ref_dictionary *out);   /* :140-145; 0 ok, -1 unknown */
uint8_t a3ref_hamming_distance(uint64_t a, uint64_t b);                   /* src/lib.rs:11-21 */
uint8_t a3ref_calculate_tau(const uint64_t *codes, uint32_t n);           /* :129-138 */
uint8_t a3ref_mark_size(const a3ref_dictionary *d);                       /* :154-156 */
void a3ref_find_nearest(const a3ref_dictionary *d, uint64_t bits, uint64_t *index, uint8_t *dist); /* :160-196 */
int a3ref_try_find_nearest(const a3ref_dictionary *d, uint64_t bits, uint64_t *index, uint8_t *dist); /* :200-207 */
uint32_t a3ref_make_binary_image(const a3ref_dictionary *d, uint64_t marker_id, uint8_t *bits_out, uint32_t cap); /* :212-232 */

/* ---- pixel front end ---- */
void a3ref_to_luma8(const uint8_t *src, int format, uint32_t w, uint32_t h, size_t pitch, uint8_t *grey);
void a3ref_adaptive_threshold(const uint8_t *grey, uint32_t w, uint32_t h, uint32_t block_radius, uint8_t *out);

/* ---- contours / polygons ---- */
a3ref_contours *a3ref_find_contours(const uint8_t *mask, uint32_t w, uint32_t h);
void a3ref_contours_free(a3ref_contours *c);
size_t a3ref_approximate_polygon_dp(const a3ref_point *curve, size_t n, double epsilon, int closed, a3ref_point *out);
size_t a3ref_convex_hull(const a3ref_point *pts, size_t n, a3ref_point *out);
uint32_t a3ref_contours_to_candidates(const a3ref_contours *c, uint32_t min_edge_length, double eps,
                                      uint32_t **quads_out, a3ref_stats *stats);
void a3ref_enforce_clockwise_corners(uint32_t *quads, uint32_t n);
uint32_t a3ref_discard_too_near(uint32_t *quads, uint32_t n, float min_distance);
float a3ref_perimeter(const uint32_t *quad);

/* ---- homography / decode ---- */
int a3ref_projection_from_control_points(const float from[8], const float to[8], float transform[9],
                                         float inverse[9], int *cls);
int a3ref_extract_homography(const uint8_t *grey, uint32_t w, uint32_t h, const uint32_t quad[8],
                             uint32_t size, uint8_t *patch);
/* the warp for a given forward transform (f32[9]) and bilinear variant (0 = the restated two-stage blend): risk probes */
int a3ref_warp_with_transform(const uint8_t *grey, uint32_t w, uint32_t h, const float transform[9], uint32_t size, int variant,
                              uint8_t *patch);
uint8_t a3ref_otsu_level(const uint8_t *img, uint32_t w, uint32_t h);
void a3ref_resize_triangle(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh, uint8_t *dst);
int a3ref_homography_to_code_permutations(const uint8_t *patch, uint32_t pw, uint32_t ph, uint8_t mark_size,
                                          uint64_t codes[4], uint8_t *otsu_out, uint8_t *reduced_out);
void a3ref_rotate_bit_matrix(const uint8_t *in, uint32_t rows, uint32_t cols, uint8_t *out);

/* ---- the whole path ---- */
void a3ref_default_config(a3ref_config *cfg);
a3ref_detection *a3ref_detect(const a3ref_config *cfg, const a3ref_dictionary *dict, const uint8_t *image,
                              int format, uint32_t w, uint32_t h, size_t pitch);
void a3ref_detection_free(a3ref_detection *d);

/* frame-parallel helper for the CPU baseline: runs a3ref_detect over n frames on `threads` pthreads and
 * returns only marker counts (results are discarded). Returns total markers. */
uint64_t a3ref_detect_many(const a3ref_config *cfg, const a3ref_dictionary *dict, const uint8_t *frames,
                           int format, uint32_t n, uint32_t w, uint32_t h, size_t pitch, size_t frame_stride,
                           uint32_t threads, a3ref_stats *sum_stats);

#ifdef __cplusplus
}
#endif
#endif
