/*
 * a3ref_pose.c — CPU ORACLE for the pose step (test infrastructure, see a3ref_pose.h).  PARITY PINNED by the
 * reference's own golden vectors (tests/test_oracle_pose.py).
 *
 * Every function names the lines of /root/reference/src/pose.rs or src/pinhole.rs it restates.  All arithmetic is
 * f32, evaluated left to right as Rust does; build with -ffp-contract=off.
 */
#include "a3ref_pose.h"

#include <math.h>
#include <string.h>

/* 3x3 row-major helpers. M(r,c) with r,c in 1..3 reads like nalgebra's m11..m33. */
#define M(m, r, c) ((m)[((r) - 1) * 3 + ((c) - 1)])

/* nalgebra Matrix3 * Vector3 (gemv: y = col0*x0; y += col1*x1; y += col2*x2) */
static void mat_vec(const float m[9], const float v[3], float out[3]) {
    for (int r = 0; r < 3; ++r) {
        float acc = m[r * 3 + 0] * v[0];
        acc = m[r * 3 + 1] * v[1] + acc;
        acc = m[r * 3 + 2] * v[2] + acc;
        out[r] = acc;
    }
}

/* pose.rs:42-50 */
void a3ref_pose_default(a3ref_pose *p) {
    static const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    p->error = 1e31f;
    memcpy(p->rotation, eye, sizeof eye);
    p->translation[0] = p->translation[1] = p->translation[2] = 0.0f;
}

/* pose.rs:24-28 and 35-39 */
void a3ref_pose_apply(const a3ref_pose *p, const float *pts, size_t n, int inverse, float *out) {
    for (size_t i = 0; i < n; ++i) {
        const float *v = pts + 3 * i;
        float r[3];
        if (!inverse) {
            mat_vec(p->rotation, v, r);
            for (int k = 0; k < 3; ++k) out[3 * i + k] = r[k] + p->translation[k];
        } else {
            float rt[9], d[3];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) rt[a * 3 + b] = p->rotation[b * 3 + a];
            for (int k = 0; k < 3; ++k) d[k] = v[k] - p->translation[k];
            mat_vec(rt, d, r);
            for (int k = 0; k < 3; ++k) out[3 * i + k] = r[k];
        }
    }
}

/* pose.rs:85-93 — clockwise from top-left, +Y up, z = 0 */
void a3ref_make_marker_square(float marker_size_mm, float sq[12]) {
    float hw = 0.5f * marker_size_mm;
    const float sx[4] = {-hw, hw, hw, -hw}, sy[4] = {hw, hw, -hw, -hw};
    for (int i = 0; i < 4; ++i) {
        sq[3 * i] = sx[i];
        sq[3 * i + 1] = sy[i];
        sq[3 * i + 2] = 0.0f;
    }
}

/* pose.rs:96-123 — closed-form homography from the centred square to four image points (all signs flipped first) */
void a3ref_homography_from_marker_square(float marker_size_mm, const float pts[8], float h[9]) {
    float x1 = -pts[0], y1 = -pts[1], x2 = -pts[2], y2 = -pts[3];
    float x3 = -pts[4], y3 = -pts[5], x4 = -pts[6], y4 = -pts[7];
    float hw = marker_size_mm / 2.0f;
    float det_inv = -1.0f / (hw * (x1 * y2 - x2 * y1 - x1 * y4 + x2 * y3 - x3 * y2 + x4 * y1 + x3 * y4 - x4 * y3));

    h[0] = det_inv * (x1 * x3 * y2 - x2 * x3 * y1 - x1 * x4 * y2 + x2 * x4 * y1 - x1 * x3 * y4 + x1 * x4 * y3 +
                      x2 * x3 * y4 - x2 * x4 * y3);
    h[1] = det_inv * (x1 * x2 * y3 - x1 * x3 * y2 - x1 * x2 * y4 + x2 * x4 * y1 + x1 * x3 * y4 - x3 * x4 * y1 -
                      x2 * x4 * y3 + x3 * x4 * y2);
    h[2] = det_inv * hw * (x1 * x2 * y3 - x2 * x3 * y1 - x1 * x2 * y4 + x1 * x4 * y2 - x1 * x4 * y3 + x3 * x4 * y1 +
                           x2 * x3 * y4 - x3 * x4 * y2);
    h[3] = det_inv * (x1 * y2 * y3 - x2 * y1 * y3 - x1 * y2 * y4 + x2 * y1 * y4 - x3 * y1 * y4 + x4 * y1 * y3 +
                      x3 * y2 * y4 - x4 * y2 * y3);
    h[4] = det_inv * (x2 * y1 * y3 - x3 * y1 * y2 - x1 * y2 * y4 + x4 * y1 * y2 + x1 * y3 * y4 - x4 * y1 * y3 -
                      x2 * y3 * y4 + x3 * y2 * y4);
    h[5] = det_inv * hw * (x1 * y2 * y3 - x3 * y1 * y2 - x2 * y1 * y4 + x4 * y1 * y2 - x1 * y3 * y4 + x3 * y1 * y4 +
                           x2 * y3 * y4 - x4 * y2 * y3);
    h[6] = -det_inv * (x1 * y3 - x3 * y1 - x1 * y4 - x2 * y3 + x3 * y2 + x4 * y1 + x2 * y4 - x4 * y2);
    h[7] = det_inv * (x1 * y2 - x2 * y1 - x1 * y3 + x3 * y1 + x2 * y4 - x4 * y2 - x3 * y4 + x4 * y3);
    h[8] = 1.0f;
}

/* pose.rs:238-267 */
void a3ref_find_rotation_to_z(const float v[3], float rot[9]) {
    memset(rot, 0, 9 * sizeof(float));
    float norm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float ax = v[0] / norm, ay = v[1] / norm, az = v[2] / norm;
    if (fabsf(1.0f + az) < 1e-6f) {
        M(rot, 1, 1) = 1.0f;
        M(rot, 2, 2) = 1.0f;
        M(rot, 3, 3) = -1.0f;
    } else {
        float d = 1.0f / (1.0f + az);
        float ax2 = ax * ax, ay2 = ay * ay, axay = ax * ay;
        M(rot, 1, 1) = -ax2 * d + 1.0f;
        M(rot, 1, 2) = -axay * d;
        M(rot, 1, 3) = -ax;
        M(rot, 2, 1) = -axay * d;
        M(rot, 2, 2) = -ay2 * d + 1.0f;
        M(rot, 2, 3) = -ay;
        M(rot, 3, 1) = ax;
        M(rot, 3, 2) = ay;
        M(rot, 3, 3) = 1.0f - (ax2 + ay2) * d;
    }
}

/* pose.rs:158-235 */
void a3ref_compute_rotations(const float j[4], float tx, float ty, float r1[9], float r2[9]) {
    float t[3] = {tx, ty, 1.0f}, rz[9], rv[9];
    a3ref_find_rotation_to_z(t, rz);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) rv[a * 3 + b] = rz[b * 3 + a]; /* .transpose() :166 */

    float b00 = M(rv, 1, 1) - tx * M(rv, 3, 1);
    float b01 = M(rv, 1, 2) - tx * M(rv, 3, 2);
    float b10 = M(rv, 2, 1) - ty * M(rv, 3, 1);
    float b11 = M(rv, 2, 2) - ty * M(rv, 3, 2);

    float inv_det = 1.0f / (b00 * b11 - b01 * b10);
    float binv00 = inv_det * b11, binv01 = -inv_det * b01, binv10 = -inv_det * b10, binv11 = inv_det * b00;

    /* jacobian row-major: j[0]=m11 j[1]=m12 j[2]=m21 j[3]=m22 */
    float a00 = binv00 * j[0] + binv01 * j[2];
    float a01 = binv00 * j[1] + binv01 * j[3];
    float a10 = binv10 * j[0] + binv11 * j[2];
    float a11 = binv10 * j[1] + binv11 * j[3];

    float ata00 = a00 * a00 + a01 * a01;
    float ata01 = a00 * a10 + a01 * a11;
    float ata11 = a10 * a10 + a11 * a11;

    float gamma = sqrtf(0.5f * (ata00 + ata11 + sqrtf((ata00 - ata11) * (ata00 - ata11) + 4.0f * ata01 * ata01)));

    float q00 = a00 / gamma, q01 = a01 / gamma, q10 = a10 / gamma, q11 = a11 / gamma;
    float q00s = q00 * q00, q01s = q01 * q01, q10s = q10 * q10, q11s = q11 * q11;

    float c0 = sqrtf(-q00s - q10s + 1.0f);
    float c1 = sqrtf(-q01s - q11s + 1.0f);
    float sp = -q00 * q01 - q10 * q11;
    if (sp < 0.0f) c1 = -c1;

    for (int r = 1; r <= 3; ++r) {
        float v1 = M(rv, r, 1), v2 = M(rv, r, 2), v3 = M(rv, r, 3);
        M(r1, r, 1) = q00 * v1 + q10 * v2 + c0 * v3;
        M(r1, r, 2) = q01 * v1 + q11 * v2 + c1 * v3;
        M(r1, r, 3) = (c1 * q10 - c0 * q11) * v1 + (c0 * q01 - c1 * q00) * v2 + (q00 * q11 - q01 * q10) * v3;
        M(r2, r, 1) = q00 * v1 + q10 * v2 + (-c0) * v3;
        M(r2, r, 2) = q01 * v1 + q11 * v2 + (-c1) * v3;
        M(r2, r, 3) = (c0 * q11 - c1 * q10) * v1 + (c1 * q00 - c0 * q01) * v2 + (q00 * q11 - q01 * q10) * v3;
    }
}

/* pose.rs:269-335 — normal equations of A t = b accumulated over the four corners */
void a3ref_compute_translation(const float sq[12], const float pts[8], const float rot[9], float t[3]) {
    float m11 = 4.0f, m22 = 4.0f, m13 = 0.0f, m23 = 0.0f, m31 = 0.0f, m32 = 0.0f, m33 = 0.0f;
    float atb0 = 0.0f, atb1 = 0.0f, atb2 = 0.0f;
    for (int i = 0; i < 4; ++i) {
        float ox = sq[3 * i], oy = sq[3 * i + 1];
        float rx = M(rot, 1, 1) * ox + M(rot, 1, 2) * oy;
        float ry = M(rot, 2, 1) * ox + M(rot, 2, 2) * oy;
        float rz = M(rot, 3, 1) * ox + M(rot, 3, 2) * oy;
        float a2 = -pts[2 * i], b2 = -pts[2 * i + 1];
        m13 += a2;
        m23 += b2;
        m31 += a2;
        m32 += b2;
        m33 += a2 * a2 + b2 * b2;
        float bx = -a2 * rz - rx;
        float by = -b2 * rz - ry;
        atb0 += bx;
        atb1 += by;
        atb2 += a2 * bx + b2 * by;
    }
    float det_inv = 1.0f / (m11 * m22 * m33 - m11 * m23 * m32 - m13 * m22 * m31);
    float s11 = m22 * m33 - m23 * m32, s12 = m13 * m32, s13 = -m13 * m22;
    float s21 = m23 * m31, s22 = m11 * m33 - m13 * m31, s23 = -m11 * m23;
    float s31 = -m22 * m31, s32 = -m11 * m32, s33 = m11 * m22;
    t[0] = det_inv * (s11 * atb0 + s12 * atb1 + s13 * atb2);
    t[1] = det_inv * (s21 * atb0 + s22 * atb1 + s23 * atb2);
    t[2] = det_inv * (s31 * atb0 + s32 * atb1 + s33 * atb2);
}

/* pose.rs:337-348 */
float a3ref_reprojection_error(const a3ref_pose *p, const float sq[12], const float pts[8]) {
    float proj[12];
    a3ref_pose_apply(p, sq, 4, 0, proj);
    float error = 0.0f;
    for (int i = 0; i < 4; ++i) {
        float z = fmaxf(proj[3 * i + 2], 1e-5f);
        float dx = proj[3 * i] / z - pts[2 * i];
        float dy = proj[3 * i + 1] / z - pts[2 * i + 1];
        error += sqrtf(dx * dx + dy * dy);
    }
    return error;
}

/* pose.rs:130-156 */
void a3ref_solve_canonical_form(const float sq[12], const float pts[8], const float h[9], a3ref_pose *p1,
                                a3ref_pose *p2) {
    float j[4] = {M(h, 1, 1) - M(h, 3, 1) * M(h, 1, 3), M(h, 1, 2) - M(h, 3, 2) * M(h, 1, 3),
                  M(h, 2, 1) - M(h, 3, 1) * M(h, 2, 3), M(h, 2, 2) - M(h, 3, 2) * M(h, 2, 3)};
    a3ref_pose_default(p1);
    a3ref_pose_default(p2);
    a3ref_compute_rotations(j, M(h, 1, 3), M(h, 2, 3), p1->rotation, p2->rotation);
    a3ref_compute_translation(sq, pts, p1->rotation, p1->translation);
    a3ref_compute_translation(sq, pts, p2->rotation, p2->translation);
    p1->error = a3ref_reprojection_error(p1, sq, pts);
    p2->error = a3ref_reprojection_error(p2, sq, pts);
}

/* pose.rs:64-81 */
void a3ref_solve_with_normalized_points(const float pts[8], float marker_size_mm, a3ref_pose *best, a3ref_pose *alt) {
    float sq[12], h[9];
    a3ref_pose p1, p2;
    a3ref_make_marker_square(marker_size_mm, sq);
    a3ref_homography_from_marker_square(marker_size_mm, pts, h);
    a3ref_solve_canonical_form(sq, pts, h, &p1, &p2);
    if (p1.error < p2.error) {
        *best = p1;
        *alt = p2;
    } else {
        *best = p2;
        *alt = p1;
    }
}

/* pose.rs:59-62 */
void a3ref_solve_with_undistorted_points(const uint32_t c[8], float marker_size_mm, uint32_t image_w,
                                         uint32_t image_h, a3ref_pose *best, a3ref_pose *alt) {
    float pts[8];
    for (int i = 0; i < 4; ++i) {
        pts[2 * i] = (float)c[2 * i] / (float)image_w;
        pts[2 * i + 1] = (float)c[2 * i + 1] / (float)image_h;
    }
    a3ref_solve_with_normalized_points(pts, marker_size_mm, best, alt);
}

/* pose.rs:52-55 */
void a3ref_solve_with_intrinsics(const uint32_t c[8], float marker_size_mm, const a3ref_intrinsics *k,
                                 a3ref_pose *best, a3ref_pose *alt) {
    float pts[8];
    for (int i = 0; i < 4; ++i) a3ref_unproject(k, (float)c[2 * i], (float)c[2 * i + 1], pts + 2 * i);
    a3ref_solve_with_normalized_points(pts, marker_size_mm, best, alt);
}

/* pinhole.rs:26-35 — NULL principal point = image centre */
void a3ref_intrinsics_new(uint32_t w, uint32_t h, float fx, float fy, const float *px, const float *py,
                          a3ref_intrinsics *out) {
    out->image_width = w;
    out->image_height = h;
    out->focal_x = fx;
    out->focal_y = fy;
    out->principal_x = px ? *px : (float)w / 2.0f;
    out->principal_y = py ? *py : (float)h / 2.0f;
}

/* pinhole.rs:37-60 */
void a3ref_intrinsics_from_fov_horizontal(float hfov, float sensor_w, uint32_t rx, uint32_t ry, a3ref_intrinsics *out) {
    float aspect = (float)rx / (float)ry;
    float vfov = hfov / aspect;
    float sensor_h = sensor_w / aspect;
    out->image_width = rx;
    out->image_height = ry;
    out->focal_x = (sensor_w * 0.5f) / tanf(hfov * 0.5f);
    out->focal_y = (sensor_h * 0.5f) / tanf(vfov * 0.5f);
    out->principal_x = (float)rx * 0.5f;
    out->principal_y = (float)ry * 0.5f;
}

/* pinhole.rs:65-71 */
void a3ref_project(const a3ref_intrinsics *k, float x, float y, float z, float out[3]) {
    out[0] = (x * k->focal_x) + (z * k->principal_x);
    out[1] = (y * k->focal_y) + (z * k->principal_y);
    out[2] = z;
}

/* pinhole.rs:76-84 */
int a3ref_project_culled(const a3ref_intrinsics *k, float x, float y, float z, float out[2]) {
    if (z <= 0.0f) return 0;
    out[0] = (x * k->focal_x) / z + k->principal_x;
    out[1] = (y * k->focal_y) / z + k->principal_y;
    return 1;
}

/* pinhole.rs:88-93 */
void a3ref_unproject(const a3ref_intrinsics *k, float x, float y, float out[2]) {
    out[0] = (x - k->principal_x) / k->focal_x;
    out[1] = (y - k->principal_y) / k->focal_y;
}
