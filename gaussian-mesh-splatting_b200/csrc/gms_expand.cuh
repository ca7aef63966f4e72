// gms_expand.cuh -- per-face maths of the mesh -> Gaussian expansion, forward and hand-derived backward.
// Host+device (see gms_common.cuh).  One thread owns one face: the face frame and its quaternion are
// computed ONCE per face (the reference recomputes them for each of the K splats of a face).
//
// Replaces (reference, pure PyTorch, ~45 ATen kernels per call + their autograd):
//   GaussianMeshModel.update_alpha / _calc_xyz      games/mesh_splatting/scene/gaussian_mesh_model.py:153-169, 86-101
//   GaussianMeshModel.prepare_scaling_rot           games/mesh_splatting/scene/gaussian_mesh_model.py:103-151
//   rot_to_quat_batch / _sqrt_positive_part / standardize_quaternion      utils/general_utils.py:19-96
//   get_scaling = exp, get_rotation = normalize (optional fused outputs)  scene/gaussian_model.py:95-101
#pragma once
#include "gms_common.cuh"
#include "../../include/gms_b200.h"

struct GmsFrame {
    float v0[3], v1[3], v2[3];   // rotation COLUMNS
    float n[3], nn;              // un-normalised normal and its norm
    float a1[3], na1;            // t1 - mean, |a1|
    float a2[3];                 // t2 - mean
    float u[3], nu;              // Gram-Schmidt residual and its norm
    float d0, d1;                // a2.v0, a2.v1
    float s[3];                  // (eps, s1, s2)
};

GMS_HD float gms_norm3(const float* v) { return GMS_SQRTN(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
GMS_HD float gms_dotv(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// t = [t0 | t1 | t2] (9 floats)
GMS_HD void gms_face_frame(const float* t, float eps, GmsFrame& f) {
    const float* t0 = t; const float* t1 = t + 3; const float* t2 = t + 6;
    float e1[3], e2[3], m[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { e1[i] = t1[i] - t0[i]; e2[i] = t2[i] - t0[i]; m[i] = GMS_DIVN(t0[i] + t1[i] + t2[i], 3.0f); }
    f.n[0] = e1[1] * e2[2] - e1[2] * e2[1];
    f.n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    f.n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    f.nn = gms_norm3(f.n);
    const float inn = GMS_DIVN(1.0f, f.nn + eps);
#pragma unroll
    for (int i = 0; i < 3; i++) { f.v0[i] = f.n[i] * inn; f.a1[i] = t1[i] - m[i]; f.a2[i] = t2[i] - m[i]; }
    f.na1 = gms_norm3(f.a1);
    const float l1 = f.na1 + eps;
#pragma unroll
    for (int i = 0; i < 3; i++) f.v1[i] = GMS_DIVN(f.a1[i], l1);
    f.d0 = gms_dotv(f.a2, f.v0);
    f.d1 = gms_dotv(f.a2, f.v1);
#pragma unroll
    for (int i = 0; i < 3; i++) f.u[i] = f.a2[i] - f.d0 * f.v0[i] - f.d1 * f.v1[i];
    f.nu = gms_norm3(f.u);
    const float lu = f.nu + eps;
#pragma unroll
    for (int i = 0; i < 3; i++) f.v2[i] = GMS_DIVN(f.u[i], lu);
    f.s[0] = eps;
    f.s[1] = l1 / 2.0f;
    f.s[2] = gms_dotv(f.a2, f.v2) / 2.0f;
}

struct GmsQuatAux { int sel; float sgn; float qa; float D; float cand[4]; };

// pytorch3d matrix_to_quaternion on R = [v0|v1|v2] (columns); q = (w,x,y,z) with w >= 0
GMS_HD void gms_frame_quat(const GmsFrame& f, float* q, GmsQuatAux& ax) {
    const float m00 = f.v0[0], m01 = f.v1[0], m02 = f.v2[0];
    const float m10 = f.v0[1], m11 = f.v1[1], m12 = f.v2[1];
    const float m20 = f.v0[2], m21 = f.v1[2], m22 = f.v2[2];
    float x[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    float qa[4];
#pragma unroll
    for (int i = 0; i < 4; i++) qa[i] = x[i] > 0.f ? GMS_SQRTN(x[i]) : 0.f;
    int sel = 0;
#pragma unroll
    for (int i = 1; i < 4; i++) if (qa[i] > qa[sel]) sel = i;   // first maximum, like torch.argmax
    float c[4];
    const float qq = qa[sel] * qa[sel];
    if (sel == 0)      { c[0] = qq;        c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (sel == 1) { c[0] = m21 - m12; c[1] = qq;        c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (sel == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qq;        c[3] = m12 + m21; }
    else               { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qq; }
    const float D = 2.0f * fmaxf(qa[sel], 0.1f);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = GMS_DIVN(c[i], D);        // D >= 0.2
    const float sgn = o[0] < 0.f ? -1.f : 1.f;
#pragma unroll
    for (int i = 0; i < 4; i++) { q[i] = sgn * o[i]; ax.cand[i] = c[i]; }
    ax.sel = sel; ax.sgn = sgn; ax.qa = qa[sel]; ax.D = D;
}

// gradient of the quaternion w.r.t. the three frame columns
GMS_HD void gms_frame_quat_backward(const GmsQuatAux& ax, const float* dq, float* dv0, float* dv1, float* dv2) {
    float dc[4]; float dD = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) { dc[i] = GMS_DIVN(ax.sgn * dq[i], ax.D); dD -= GMS_DIVN(ax.sgn * dq[i] * ax.cand[i], ax.D * ax.D); }
    // d x_sel : cand[sel] = qa^2 = x (x > 0) ; D = 2*max(qa, 0.1)
    float dx = 0.f;
    if (ax.qa > 0.f) {
        dx = dc[ax.sel];
        if (ax.qa > 0.1f) dx += GMS_DIVN(dD, ax.qa);        // dD * dD/dqa * dqa/dx = dD * 2 * 1/(2 qa)
    }
    float dm[9];
#pragma unroll
    for (int i = 0; i < 9; i++) dm[i] = 0.f;
    // x_sel = 1 + s0*m00 + s1*m11 + s2*m22
    const float s00 = (ax.sel == 0 || ax.sel == 1) ? 1.f : -1.f;
    const float s11 = (ax.sel == 0 || ax.sel == 2) ? 1.f : -1.f;
    const float s22 = (ax.sel == 0 || ax.sel == 3) ? 1.f : -1.f;
    dm[0] += s00 * dx; dm[4] += s11 * dx; dm[8] += s22 * dx;
    // off-diagonal candidates; index dm[3*a+b] = m_ab
    if (ax.sel == 0) {
        dm[7] += dc[1]; dm[5] -= dc[1];   // m21 - m12
        dm[2] += dc[2]; dm[6] -= dc[2];   // m02 - m20
        dm[3] += dc[3]; dm[1] -= dc[3];   // m10 - m01
    } else if (ax.sel == 1) {
        dm[7] += dc[0]; dm[5] -= dc[0];   // m21 - m12
        dm[3] += dc[2]; dm[1] += dc[2];   // m10 + m01
        dm[2] += dc[3]; dm[6] += dc[3];   // m02 + m20
    } else if (ax.sel == 2) {
        dm[2] += dc[0]; dm[6] -= dc[0];   // m02 - m20
        dm[3] += dc[1]; dm[1] += dc[1];   // m10 + m01
        dm[5] += dc[3]; dm[7] += dc[3];   // m12 + m21
    } else {
        dm[3] += dc[0]; dm[1] -= dc[0];   // m10 - m01
        dm[6] += dc[1]; dm[2] += dc[1];   // m20 + m02
        dm[7] += dc[2]; dm[5] += dc[2];   // m21 + m12
    }
    // m_ab = v_b[a]
#pragma unroll
    for (int a = 0; a < 3; a++) { dv0[a] += dm[3 * a + 0]; dv1[a] += dm[3 * a + 1]; dv2[a] += dm[3 * a + 2]; }
}

// Back-propagate (dv0, dv1, dv2, ds1, ds2) through the face frame to the triangle corners; dt += ...
GMS_HD void gms_face_frame_backward(const float* t, float eps, const GmsFrame& f, float* dv0, float* dv1, float* dv2,
                                    float ds1, float ds2, float* dt) {
    const float* t0 = t; const float* t1 = t + 3; const float* t2 = t + 6;
    float da1[3] = {0.f, 0.f, 0.f}, da2[3] = {0.f, 0.f, 0.f};
    // s2 = (a2 . v2) / 2
#pragma unroll
    for (int i = 0; i < 3; i++) { da2[i] += 0.5f * ds2 * f.v2[i]; dv2[i] += 0.5f * ds2 * f.a2[i]; }
    // v2 = u / (|u| + eps)
    const float lu = f.nu + eps;
    float du[3];
    {
        const float k = f.nu > 0.f ? gms_dotv(dv2, f.u) / (f.nu * lu * lu) : 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) du[i] = GMS_DIVN(dv2[i], lu) - k * f.u[i];
    }
    // u = a2 - (a2.v0) v0 - (a2.v1) v1
    {
        const float p0 = gms_dotv(du, f.v0), p1 = gms_dotv(du, f.v1);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            da2[i] += du[i] - p0 * f.v0[i] - p1 * f.v1[i];
            dv0[i] += -p0 * f.a2[i] - f.d0 * du[i];
            dv1[i] += -p1 * f.a2[i] - f.d1 * du[i];
        }
    }
    // v1 = a1 / l1 ; l1 = |a1| + eps ; s1 = l1 / 2
    {
        const float l1 = f.na1 + eps;
        const float dl1 = 0.5f * ds1 - GMS_DIVN(gms_dotv(dv1, f.a1), l1 * l1);      // l1 >= eps: l1^2 >= 1e-16
        const float k = f.na1 > 0.f ? dl1 / f.na1 : 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) da1[i] += GMS_DIVN(dv1[i], l1) + k * f.a1[i];
    }
    // a1 = t1 - m ; a2 = t2 - m ; m = (t0+t1+t2)/3
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float dm = GMS_DIVN(-(da1[i] + da2[i]), 3.0f);
        dt[i] += dm; dt[3 + i] += da1[i] + dm; dt[6 + i] += da2[i] + dm;
    }
    // v0 = n / (|n| + eps) ; n = e1 x e2
    {
        const float ln = f.nn + eps;
        const float k = f.nn > 0.f ? gms_dotv(dv0, f.n) / (f.nn * ln * ln) : 0.f;
        float dn[3], e1[3], e2[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { dn[i] = GMS_DIVN(dv0[i], ln) - k * f.n[i]; e1[i] = t1[i] - t0[i]; e2[i] = t2[i] - t0[i]; }
        const float de1[3] = {e2[1] * dn[2] - e2[2] * dn[1], e2[2] * dn[0] - e2[0] * dn[2], e2[0] * dn[1] - e2[1] * dn[0]};
        const float de2[3] = {dn[1] * e1[2] - dn[2] * e1[1], dn[2] * e1[0] - dn[0] * e1[2], dn[0] * e1[1] - dn[1] * e1[0]};
#pragma unroll
        for (int i = 0; i < 3; i++) { dt[3 + i] += de1[i]; dt[6 + i] += de2[i]; dt[i] -= de1[i] + de2[i]; }
    }
}

// ---- whole-face forward / backward (called by k_expand_fwd / k_expand_bwd, and by tests/hostshim on the CPU)
#if defined(__CUDA_ARCH__)
#define GMS_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define GMS_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

// `f` indexes the per-FACE arrays (faces / triangles_in / triangles); `fl` indexes the per-GAUSSIAN streams (alpha_raw,
// scale_raw and every output row): fl == f when they are the caller's arrays, fl == the face's slot in the block when the
// kernel has redirected those pointers to its shared-memory staging buffers (k_expand_fwd / k_expand_bwd).
GMS_HD void gms_expand_face_fwd(const gms_expand_args& a, int f, int fl) {
    float t[9];
    if (a.triangles_in) {
#pragma unroll
        for (int k = 0; k < 9; k++) t[k] = a.triangles_in[9 * (size_t)f + k];
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int64_t vi = a.faces[3 * (size_t)f + c];
            t[3 * c] = a.vertices[3 * vi]; t[3 * c + 1] = a.vertices[3 * vi + 1]; t[3 * c + 2] = a.vertices[3 * vi + 2];
        }
    }
    if (a.triangles) {
#pragma unroll
        for (int k = 0; k < 9; k++) a.triangles[9 * (size_t)f + k] = t[k];
    }
    GmsFrame fr;
    gms_face_frame(t, a.eps, fr);
    float q[4];
    GmsQuatAux ax;
    gms_frame_quat(fr, q, ax);
    const float qn = fmaxf(GMS_SQRTN(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
    for (int k = 0; k < a.K; k++) {
        const size_t p = (size_t)fl * a.K + k;
        const float r0 = fmaxf(a.alpha_raw[3 * p], 0.f) + 1e-8f, r1 = fmaxf(a.alpha_raw[3 * p + 1], 0.f) + 1e-8f,
                    r2 = fmaxf(a.alpha_raw[3 * p + 2], 0.f) + 1e-8f;
        const float S = r0 + r1 + r2;
        const float al0 = GMS_DIVN(r0, S), al1 = GMS_DIVN(r1, S), al2 = GMS_DIVN(r2, S);        // S >= 3e-8
        if (a.alpha) { a.alpha[3 * p] = al0; a.alpha[3 * p + 1] = al1; a.alpha[3 * p + 2] = al2; }
        if (a.xyz) {
#pragma unroll
            for (int c = 0; c < 3; c++) a.xyz[3 * p + c] = al0 * t[c] + al1 * t[3 + c] + al2 * t[6 + c];
        }
        const float cs = a.scale_raw[p];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float inner = fmaxf(cs * fr.s[c], 0.f) + a.eps;
            if (a.scaling_log) a.scaling_log[3 * p + c] = logf(inner);
            if (a.scaling_act) a.scaling_act[3 * p + c] = expf(logf(inner));
        }
        if (a.rotation_raw) { float* o = a.rotation_raw + 4 * p; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; }
        if (a.rotation_act) { float* o = a.rotation_act + 4 * p; o[0] = GMS_DIVN(q[0], qn); o[1] = GMS_DIVN(q[1], qn); o[2] = GMS_DIVN(q[2], qn); o[3] = GMS_DIVN(q[3], qn); }
    }
}

GMS_HD void gms_expand_face_bwd(const gms_expand_args& a, const gms_expand_grads& g, int f, int fl) {
    float t[9];
    int64_t vi[3] = {0, 0, 0};
    if (a.triangles_in) {
#pragma unroll
        for (int k = 0; k < 9; k++) t[k] = a.triangles_in[9 * (size_t)f + k];
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            vi[c] = a.faces[3 * (size_t)f + c];
            t[3 * c] = a.vertices[3 * vi[c]]; t[3 * c + 1] = a.vertices[3 * vi[c] + 1]; t[3 * c + 2] = a.vertices[3 * vi[c] + 2];
        }
    }
    GmsFrame fr;
    gms_face_frame(t, a.eps, fr);
    float q[4];
    GmsQuatAux ax;
    gms_frame_quat(fr, q, ax);
    const float qnorm = GMS_SQRTN(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float qn = fmaxf(qnorm, 1e-12f);
    float dt[9];
#pragma unroll
    for (int k = 0; k < 9; k++) dt[k] = 0.f;
    float dq[4] = {0.f, 0.f, 0.f, 0.f};
    float ds1 = 0.f, ds2 = 0.f;
    for (int k = 0; k < a.K; k++) {
        const size_t p = (size_t)fl * a.K + k;
        // --- xyz = alpha @ triangle
        float dx[3] = {0.f, 0.f, 0.f};
        if (g.dL_dxyz) { dx[0] = g.dL_dxyz[3 * p]; dx[1] = g.dL_dxyz[3 * p + 1]; dx[2] = g.dL_dxyz[3 * p + 2]; }
        const float ar[3] = {a.alpha_raw[3 * p], a.alpha_raw[3 * p + 1], a.alpha_raw[3 * p + 2]};
        const float r[3] = {fmaxf(ar[0], 0.f) + 1e-8f, fmaxf(ar[1], 0.f) + 1e-8f, fmaxf(ar[2], 0.f) + 1e-8f};
        const float S = r[0] + r[1] + r[2];
        const float al[3] = {GMS_DIVN(r[0], S), GMS_DIVN(r[1], S), GMS_DIVN(r[2], S)};
        float dal[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            dal[j] = dx[0] * t[3 * j] + dx[1] * t[3 * j + 1] + dx[2] * t[3 * j + 2];
#pragma unroll
            for (int c = 0; c < 3; c++) dt[3 * j + c] += al[j] * dx[c];
        }
        const float dsum = dal[0] * al[0] + dal[1] * al[1] + dal[2] * al[2];
        if (g.dL_dalpha_raw) {
#pragma unroll
            for (int j = 0; j < 3; j++) g.dL_dalpha_raw[3 * p + j] = ar[j] > 0.f ? GMS_DIVN(dal[j] - dsum, S) : 0.f;
        }
        // --- scaling
        const float cs = a.scale_raw[p];
        float dcs = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float gl = g.dL_dscaling_log ? g.dL_dscaling_log[3 * p + c] : 0.f;
            const float prod = cs * fr.s[c];
            const float inner = fmaxf(prod, 0.f) + a.eps;
            if (g.dL_dscaling_act) gl += g.dL_dscaling_act[3 * p + c] * expf(logf(inner));
            const float dprod = prod > 0.f ? GMS_DIVN(gl, inner) : 0.f;       // inner >= eps
            dcs += dprod * fr.s[c];
            if (c == 1) ds1 += dprod * cs;
            if (c == 2) ds2 += dprod * cs;
        }
        if (g.dL_dscale_raw) g.dL_dscale_raw[p] = dcs;
        // --- rotation (same quaternion for the K splats of the face: sum the incoming rows)
        if (g.dL_drotation_raw) {
            const float* d = g.dL_drotation_raw + 4 * p;
            dq[0] += d[0]; dq[1] += d[1]; dq[2] += d[2]; dq[3] += d[3];
        }
        if (g.dL_drotation_act) {
            const float* dp = g.dL_drotation_act + 4 * p;
            const float d_x = dp[0], d_y = dp[1], d_z = dp[2], d_w = dp[3];
            if (qnorm >= 1e-12f) {
                const float u[4] = {GMS_DIVN(q[0], qn), GMS_DIVN(q[1], qn), GMS_DIVN(q[2], qn), GMS_DIVN(q[3], qn)};       // qn >= 1e-12
                const float dd = d_x * u[0] + d_y * u[1] + d_z * u[2] + d_w * u[3];
                dq[0] += GMS_DIVN(d_x - u[0] * dd, qn); dq[1] += GMS_DIVN(d_y - u[1] * dd, qn);
                dq[2] += GMS_DIVN(d_z - u[2] * dd, qn); dq[3] += GMS_DIVN(d_w - u[3] * dd, qn);
            } else {
                dq[0] += GMS_DIVN(d_x, qn); dq[1] += GMS_DIVN(d_y, qn); dq[2] += GMS_DIVN(d_z, qn); dq[3] += GMS_DIVN(d_w, qn);
            }
        }
    }
    float dv0[3] = {0.f, 0.f, 0.f}, dv1[3] = {0.f, 0.f, 0.f}, dv2[3] = {0.f, 0.f, 0.f};
    gms_frame_quat_backward(ax, dq, dv0, dv1, dv2);
    gms_face_frame_backward(t, a.eps, fr, dv0, dv1, dv2, ds1, ds2, dt);
    if (g.dL_dtriangles) {
#pragma unroll
        for (int k = 0; k < 9; k++) g.dL_dtriangles[9 * (size_t)f + k] = dt[k];
    }
    if (g.dL_dvertices && !a.triangles_in) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            GMS_ATOMIC_ADD(&g.dL_dvertices[3 * vi[c]], dt[3 * c]);
            GMS_ATOMIC_ADD(&g.dL_dvertices[3 * vi[c] + 1], dt[3 * c + 1]);
            GMS_ATOMIC_ADD(&g.dL_dvertices[3 * vi[c] + 2], dt[3 * c + 2]);
        }
    }
}


// ---- gs_points pseudo-mesh path: every Gaussian carries its own triangle (v1, v2, v3);
// PointsGaussianModel.prepare_scaling_rot  games/flat_splatting/scene/points_gaussian_model.py:61-104 and the per-frame
// call in renderer/gaussian_points_animated_renderer/__init__.py:61-66 (_xyz = triangles[:, 0]).  Forward only: the
// reference uses it under torch.no_grad() in scripts/render_points_time_animated.py.
GMS_HD void gms_points_face_fwd(const gms_points_args& a, int i) {
    const float* t = a.triangles + 9 * (size_t)i;
    const float* v1 = t; const float* v2 = t + 3; const float* v3 = t + 6;
    float e2[3], e3[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { e2[k] = v2[k] - v1[k]; e3[k] = v3[k] - v1[k]; }
    GmsFrame f;                                   // v0 <- r1 (normal), v1 <- r2, v2 <- r3: rotation COLUMNS
    f.n[0] = e2[1] * e3[2] - e2[2] * e3[1];
    f.n[1] = e2[2] * e3[0] - e2[0] * e3[2];
    f.n[2] = e2[0] * e3[1] - e2[1] * e3[0];
    f.nn = gms_norm3(f.n);
    const float s2 = gms_norm3(e2) + a.eps;
    const float inn = 1.0f / (f.nn + a.eps);
#pragma unroll
    for (int k = 0; k < 3; k++) { f.v0[k] = f.n[k] * inn; f.v1[k] = e2[k] / s2; }
    const float d0 = gms_dotv(e3, f.v0), d1 = gms_dotv(e3, f.v1);
    float u[3];
#pragma unroll
    for (int k = 0; k < 3; k++) u[k] = e3[k] - d0 * f.v0[k] - d1 * f.v1[k];
    const float lu = gms_norm3(u) + a.eps;
#pragma unroll
    for (int k = 0; k < 3; k++) f.v2[k] = u[k] / lu;
    const float s3 = gms_dotv(e3, f.v2);
    float q[4];
    GmsQuatAux ax;
    gms_frame_quat(f, q, ax);
    if (a.xyz) { a.xyz[3 * (size_t)i] = v1[0]; a.xyz[3 * (size_t)i + 1] = v1[1]; a.xyz[3 * (size_t)i + 2] = v1[2]; }
    if (a.scaling_log) { a.scaling_log[2 * (size_t)i] = logf(fabsf(s2)); a.scaling_log[2 * (size_t)i + 1] = logf(fabsf(s3)); }
    if (a.scaling_act) {      // get_scaling = cat([eps], exp(_scaling))  (points_gaussian_model.py:106-109)
        a.scaling_act[3 * (size_t)i] = a.eps;
        a.scaling_act[3 * (size_t)i + 1] = expf(logf(fabsf(s2)));
        a.scaling_act[3 * (size_t)i + 2] = expf(logf(fabsf(s3)));
    }
    if (a.rotation_raw) { float* o = a.rotation_raw + 4 * (size_t)i; o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; }
    if (a.rotation_act) {
        const float qn = fmaxf(GMS_SQRTN(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
        float* o = a.rotation_act + 4 * (size_t)i; o[0] = GMS_DIVN(q[0], qn); o[1] = GMS_DIVN(q[1], qn); o[2] = GMS_DIVN(q[2], qn); o[3] = GMS_DIVN(q[3], qn);
    }
}

// PointsGaussianModel.prepare_vertices (games/flat_splatting/scene/points_gaussian_model.py:28-59): pseudo-mesh triangle of
// one flat Gaussian.  scales = (eps, exp(_scaling[-2]), exp(_scaling[-1])) (:106-109); R = build_rotation(_rotation)
// (utils/general_utils.py:158-179, normalises q); arms along R's 2nd and 3rd COLUMNS (R.transpose(-2,-1)[:, 1|2]);
// the longer arm becomes v2 (mask = s_2 > s_3, :43-52).
GMS_HD void gms_points_vertices_fwd(const gms_points_vertices_args& a, int i) {
    const float* c = a.xyz + 3 * (size_t)i;
    const float* sl = a.scaling_log + (size_t)a.scaling_cols * i + (a.scaling_cols - 2);
    const float* qr = a.rotation_raw + 4 * (size_t)i;
    const float s2 = expf(sl[0]), s3 = expf(sl[1]);
    const float nrm = sqrtf(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
    const float r = qr[0] / nrm, x = qr[1] / nrm, y = qr[2] / nrm, z = qr[3] / nrm;
    const float ax2[3] = {2.0f * (x * y - r * z), 1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z + r * x)};
    const float ax3[3] = {2.0f * (x * z + r * y), 2.0f * (y * z - r * x), 1.0f - 2.0f * (x * x + y * y)};
    float p2[3], p3[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { p2[k] = c[k] + s2 * ax2[k]; p3[k] = c[k] + s3 * ax3[k]; }
    const bool keep = s2 > s3;
    float* t = a.triangles + 9 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; k++) { t[k] = c[k]; t[3 + k] = keep ? p2[k] : p3[k]; t[6 + k] = keep ? p3[k] : p2[k]; }
}
