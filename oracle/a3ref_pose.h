/*
 * a3ref_pose — CPU ORACLE for the pose step that follows `detect`.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of /root/reference/src/pose.rs:52-348 (IPPE-style closed form for a square marker, all f32) and
 * of the pinhole helpers it uses (/root/reference/src/pinhole.rs:25-94).  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load it; nothing under aruco3_b200/ does.
 *
 * PARITY STATUS: PINNED.  Unlike the pixel path, the reference's own tests hold golden vectors for this step
 * (src/pose.rs:379-392 transforms, 441-455 marker square, 457-474 homography, 476-512 canonical solve,
 * 514-552 and 554-598 end to end); tests/test_oracle_pose.py restates every one of them against this file.
 * Matrix products follow nalgebra 0.33's evaluation order (column-by-column axpy, so each entry is
 * ((m0*x0 + m1*x1) + m2*x2)); `normalize` is x / sqrt((x*x + y*y) + z*z).  Compile with -ffp-contract=off.
 */
#ifndef A3REF_POSE_H
#define A3REF_POSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* MarkerPose (src/pose.rs:8-12); rotation row-major (m11 m12 m13 m21 ...). */
typedef struct a3ref_pose {
    float error;
    float rotation[9];
    float translation[3];
} a3ref_pose;

/* CameraIntrinsics (src/pinhole.rs:11-18). */
typedef struct a3ref_intrinsics {
    uint32_t image_width, image_height;
    float focal_x, focal_y, principal_x, principal_y;
} a3ref_intrinsics;

void a3ref_pose_default(a3ref_pose *p);                                             /* pose.rs:42-50   */
void a3ref_make_marker_square(float marker_size_mm, float square[12]);              /* pose.rs:85-93   */
void a3ref_homography_from_marker_square(float marker_size_mm, const float pts[8], float h[9]); /* :96-123 */
void a3ref_find_rotation_to_z(const float v[3], float rot[9]);                      /* pose.rs:238-267 */
void a3ref_compute_rotations(const float jacobian[4], float tx, float ty, float r1[9], float r2[9]); /* :158-235 */
void a3ref_compute_translation(const float square[12], const float pts[8], const float rot[9], float t[3]); /* :269-335 */
float a3ref_reprojection_error(const a3ref_pose *p, const float square[12], const float pts[8]); /* :337-348 */
void a3ref_solve_canonical_form(const float square[12], const float pts[8], const float h[9], a3ref_pose *p1,
                                a3ref_pose *p2);                                     /* pose.rs:130-156 */
void a3ref_solve_with_normalized_points(const float pts[8], float marker_size_mm, a3ref_pose *best,
                                        a3ref_pose *alt);                            /* pose.rs:64-81   */
void a3ref_solve_with_undistorted_points(const uint32_t corners[8], float marker_size_mm, uint32_t image_w,
                                         uint32_t image_h, a3ref_pose *best, a3ref_pose *alt); /* :59-62 */
void a3ref_solve_with_intrinsics(const uint32_t corners[8], float marker_size_mm, const a3ref_intrinsics *k,
                                 a3ref_pose *best, a3ref_pose *alt);                 /* pose.rs:52-55   */
/* apply_transform_to_vectors / apply_inverse_transform_to_vectors (pose.rs:24-28, 35-39); pts n*3 */
void a3ref_pose_apply(const a3ref_pose *p, const float *pts, size_t n, int inverse, float *out);

/* pinhole.rs:26-35, 37-60, 65-71, 76-84, 88-93 */
void a3ref_intrinsics_new(uint32_t w, uint32_t h, float fx, float fy, const float *px, const float *py,
                          a3ref_intrinsics *out);
void a3ref_intrinsics_from_fov_horizontal(float hfov_rad, float sensor_width_mm, uint32_t res_x, uint32_t res_y,
                                          a3ref_intrinsics *out);
void a3ref_project(const a3ref_intrinsics *k, float x, float y, float z, float out[3]);
int a3ref_project_culled(const a3ref_intrinsics *k, float x, float y, float z, float out[2]);
void a3ref_unproject(const a3ref_intrinsics *k, float x, float y, float out[2]);

#ifdef __cplusplus
}
#endif
#endif
