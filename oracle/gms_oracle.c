/*
 * gms_oracle.c -- CPU ORACLE for the mesh-Gaussian rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gaussian-mesh-splatting_b200/ may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker / the CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" for the rasterizer.  The algorithm lives in
 * the third-party submodule graphdeco-inria/diff-gaussian-rasterization
 * (pinned 59f5f77e3ddbac3ed9db93ec2cfe99ed6c5d121d in /root/reference/.gitmodules:4-6,
 * but the reference's call sites -- renderer/gaussian_renderer/__init__.py:43-57,
 * 94-102 -- need the later antialiasing / inverse-depth API).  Its source is NOT
 * vendored under /root/reference (the directory is empty) and the reference
 * ships no tests or golden vectors for it.  This file restates the published
 * algorithm as specified in SURVEY.md Appendix A.  It is pinned instead against
 *   (1) the pieces the reference DOES ship in Python: SH evaluation
 *       (utils/sh_utils.py:57-112), cov3D = (R S)(R S)^T packing
 *       (utils/general_utils.py:144-190, scene/gaussian_model.py:27-31) and the
 *       projection conventions (utils/graphics_utils.py:22-29,51-71) -- golden
 *       vectors generated from those live under tests/golden/;
 *   (2) an independent dense PyTorch autograd model (oracle/torch_dense.py)
 *       for the compositing forward and every gradient;
 *   (3) closed-form known-answer tests (tests/test_oracle_kat.py).
 *
 * Floating-point convention: every expression that decides an INTEGER output
 * (radii, tile rectangles, depth key bits, hence tiles_touched and the sorted
 * point list) is written as an explicit sequence of IEEE fp32 mul/add/fma/div/
 * sqrt so that the CUDA kernels can replay exactly the same sequence; compile
 * with -ffp-contract=off so gcc adds no fusions of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GMSO_BLOCK 16
#define GMSO_NCH 3

typedef struct {
    int32_t P;            /* number of Gaussians */
    int32_t D;            /* active SH degree 0..3 */
    int32_t M;            /* SH coefficients stored per Gaussian (16) */
    int32_t W, H;         /* image size */
    float tanfovx, tanfovy;
    float scale_modifier;
    int32_t prefiltered;
    int32_t antialiasing;
    float viewmatrix[16]; /* flat, transposed convention: m[4*c+r] = true (r,c) */
    float projmatrix[16];
    float campos[3];
    float bg[3];
} gmso_settings;

/* SH constants: utils/sh_utils.py:26-43 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* Work statistics of the last composite forward (for DESIGN.md / kernel design, not used by any test):
 * [0] (pixel,splat) pairs visited before the pixel terminated, [1] pairs that blended (alpha >= 1/255),
 * [2] (8x8 quad, splat) pairs with at least one blending pixel, [3] (tile, splat) pairs with >= 1 blending pixel. */
static long long g_stats[4];
void gmso_get_stats(long long* out) { for (int i = 0; i < 4; i++) out[i] = g_stats[i]; }

/* Bounded-sample support for the CPU baseline (bench.py): composite only tiles with tile % stride == 0. */
static int g_tile_stride = 1;
void gmso_set_tile_stride(int s) { g_tile_stride = s > 0 ? s : 1; }

int gmso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void gmso_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- canonical fp32 sequences (replayed bit-for-bit by the CUDA kernels) ---- */
static inline float dot3_fma(float a0, float b0, float a1, float b1, float a2, float b2) {
    /* ((a0*b0 + a1*b1) + a2*b2) with the two adds fused */
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

static inline void xform4x3(const float* m, float x, float y, float z, float* o) {
    o[0] = dot3_fma(m[0], x, m[4], y, m[8], z) + m[12];
    o[1] = dot3_fma(m[1], x, m[5], y, m[9], z) + m[13];
    o[2] = dot3_fma(m[2], x, m[6], y, m[10], z) + m[14];
}

static inline void xform4x4(const float* m, float x, float y, float z, float* o) {
    o[0] = dot3_fma(m[0], x, m[4], y, m[8], z) + m[12];
    o[1] = dot3_fma(m[1], x, m[5], y, m[9], z) + m[13];
    o[2] = dot3_fma(m[2], x, m[6], y, m[10], z) + m[14];
    o[3] = dot3_fma(m[3], x, m[7], y, m[11], z) + m[15];
}

/* rotation matrix of an (unnormalised) quaternion q=(r,x,y,z), row-major;
 * utils/general_utils.py:170-178, Appendix A.1 step 3 */
static inline void quat_to_R(const float* q, float* R) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.f - 2.f * fmaf(y, y, z * z);
    R[1] = 2.f * fmaf(x, y, -(r * z));
    R[2] = 2.f * fmaf(x, z, r * y);
    R[3] = 2.f * fmaf(x, y, r * z);
    R[4] = 1.f - 2.f * fmaf(x, x, z * z);
    R[5] = 2.f * fmaf(y, z, -(r * x));
    R[6] = 2.f * fmaf(x, z, -(r * y));
    R[7] = 2.f * fmaf(y, z, r * x);
    R[8] = 1.f - 2.f * fmaf(x, x, y * y);
}

/* Sigma = (R S)(R S)^T packed [00,01,02,11,12,22]; utils/general_utils.py:144-153 */
static inline void cov3d_from_scale_rot(const float* scale, float mod, const float* q, float* cov6) {
    float R[9], Mx[9];
    quat_to_R(q, R);
    float s0 = mod * scale[0], s1 = mod * scale[1], s2 = mod * scale[2];
    for (int i = 0; i < 3; i++) {
        Mx[3 * i + 0] = R[3 * i + 0] * s0;
        Mx[3 * i + 1] = R[3 * i + 1] * s1;
        Mx[3 * i + 2] = R[3 * i + 2] * s2;
    }
    cov6[0] = dot3_fma(Mx[0], Mx[0], Mx[1], Mx[1], Mx[2], Mx[2]);
    cov6[1] = dot3_fma(Mx[0], Mx[3], Mx[1], Mx[4], Mx[2], Mx[5]);
    cov6[2] = dot3_fma(Mx[0], Mx[6], Mx[1], Mx[7], Mx[2], Mx[8]);
    cov6[3] = dot3_fma(Mx[3], Mx[3], Mx[4], Mx[4], Mx[5], Mx[5]);
    cov6[4] = dot3_fma(Mx[3], Mx[6], Mx[4], Mx[7], Mx[5], Mx[8]);
    cov6[5] = dot3_fma(Mx[6], Mx[6], Mx[7], Mx[7], Mx[8], Mx[8]);
}

typedef struct {
    float M0[3], M1[3]; /* rows of the 2x3 matrix J*W */
    float tx, ty, tz;   /* clamped view-space point used in J */
    float xmul, ymul;   /* 0 where the guard-band clamp was active */
    float a, b, c;      /* raw cov2D (no dilation) */
} cov2d_aux;

/* EWA projection of Sigma; Appendix A.1 step 4 */
static inline void cov2d_project(const float* pview, const float* cov6, const float* view,
                                 float focal_x, float focal_y, float tanfovx, float tanfovy,
                                 cov2d_aux* o) {
    float tz = pview[2];
    float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    float txtz = pview[0] / tz, tytz = pview[1] / tz;
    o->xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    o->ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    o->tx = tx; o->ty = ty; o->tz = tz;
    float tz2 = tz * tz;
    float J00 = focal_x / tz;
    float J02 = -(focal_x * tx) / tz2;
    float J11 = focal_y / tz;
    float J12 = -(focal_y * ty) / tz2;
    /* W[i][j] = view[4*j+i] */
    for (int j = 0; j < 3; j++) {
        float W0j = view[4 * j + 0], W1j = view[4 * j + 1], W2j = view[4 * j + 2];
        o->M0[j] = fmaf(J02, W2j, J00 * W0j);
        o->M1[j] = fmaf(J12, W2j, J11 * W1j);
    }
    const float* M0 = o->M0; const float* M1 = o->M1;
    float S00 = cov6[0], S01 = cov6[1], S02 = cov6[2], S11 = cov6[3], S12 = cov6[4], S22 = cov6[5];
    float v0 = dot3_fma(S00, M0[0], S01, M0[1], S02, M0[2]);
    float v1 = dot3_fma(S01, M0[0], S11, M0[1], S12, M0[2]);
    float v2 = dot3_fma(S02, M0[0], S12, M0[1], S22, M0[2]);
    float w0 = dot3_fma(S00, M1[0], S01, M1[1], S02, M1[2]);
    float w1 = dot3_fma(S01, M1[0], S11, M1[1], S12, M1[2]);
    float w2 = dot3_fma(S02, M1[0], S12, M1[1], S22, M1[2]);
    o->a = dot3_fma(M0[0], v0, M0[1], v1, M0[2], v2);
    o->b = dot3_fma(M1[0], v0, M1[1], v1, M1[2], v2);
    o->c = dot3_fma(M1[0], w0, M1[1], w1, M1[2], w2);
}

/* Canonical evaluation of  -0.5*(cx*dx*dx + cz*dy*dy) - cy*dx*dy  (Appendix A.2): the FMA-contracted
 * form of the upstream expression; edge-on flat Gaussians make the three terms cancel, so the op order
 * is part of the contract (the CUDA kernels replay it bit-for-bit; only exp() may differ by ulps). */
static inline float quad_power(const float* co, float dx, float dy) {
    float m1 = co[0] * dx;
    float m2 = m1 * dx;
    float m3 = co[2] * dy;
    float t = fmaf(m3, dy, m2);
    float h = -0.5f * t;
    float m4 = co[1] * dx;
    return fmaf(-m4, dy, h);
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static inline void get_rect(float px, float py, int radius, int gx, int gy, int* rmin, int* rmax) {
    float r = (float)radius;
    rmin[0] = imin(gx, imax(0, (int)((px - r) / (float)GMSO_BLOCK)));
    rmin[1] = imin(gy, imax(0, (int)((py - r) / (float)GMSO_BLOCK)));
    rmax[0] = imin(gx, imax(0, (int)((px + r + (float)(GMSO_BLOCK - 1)) / (float)GMSO_BLOCK)));
    rmax[1] = imin(gy, imax(0, (int)((py + r + (float)(GMSO_BLOCK - 1)) / (float)GMSO_BLOCK)));
}

/* SH basis, utils/sh_utils.py:74-100 (degree <= 3) */
static inline void sh_basis(int deg, float x, float y, float z, float* B) {
    B[0] = SH_C0;
    if (deg > 0) {
        B[1] = -SH_C1 * y; B[2] = SH_C1 * z; B[3] = -SH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            B[4] = SH_C2[0] * xy; B[5] = SH_C2[1] * yz;
            B[6] = SH_C2[2] * (2.0f * zz - xx - yy);
            B[7] = SH_C2[3] * xz; B[8] = SH_C2[4] * (xx - yy);
            if (deg > 2) {
                B[9] = SH_C3[0] * y * (3.0f * xx - yy);
                B[10] = SH_C3[1] * xy * z;
                B[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
                B[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                B[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
                B[14] = SH_C3[5] * z * (xx - yy);
                B[15] = SH_C3[6] * x * (xx - 3.0f * yy);
            }
        }
    }
}

/* d(basis)/d(dir) */
static inline void sh_basis_grad(int deg, float x, float y, float z, float* Bx, float* By, float* Bz) {
    for (int k = 0; k < 16; k++) { Bx[k] = 0.f; By[k] = 0.f; Bz[k] = 0.f; }
    if (deg > 0) {
        By[1] = -SH_C1; Bz[2] = SH_C1; Bx[3] = -SH_C1;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            Bx[4] = SH_C2[0] * y; By[4] = SH_C2[0] * x;
            By[5] = SH_C2[1] * z; Bz[5] = SH_C2[1] * y;
            Bx[6] = SH_C2[2] * (-2.f * x); By[6] = SH_C2[2] * (-2.f * y); Bz[6] = SH_C2[2] * (4.f * z);
            Bx[7] = SH_C2[3] * z; Bz[7] = SH_C2[3] * x;
            Bx[8] = SH_C2[4] * (2.f * x); By[8] = SH_C2[4] * (-2.f * y);
            if (deg > 2) {
                Bx[9] = SH_C3[0] * 6.f * xy; By[9] = SH_C3[0] * (3.f * xx - 3.f * yy);
                Bx[10] = SH_C3[1] * yz; By[10] = SH_C3[1] * xz; Bz[10] = SH_C3[1] * xy;
                Bx[11] = SH_C3[2] * (-2.f * xy); By[11] = SH_C3[2] * (4.f * zz - xx - 3.f * yy);
                Bz[11] = SH_C3[2] * 8.f * yz;
                Bx[12] = SH_C3[3] * (-6.f * xz); By[12] = SH_C3[3] * (-6.f * yz);
                Bz[12] = SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
                Bx[13] = SH_C3[4] * (4.f * zz - 3.f * xx - yy); By[13] = SH_C3[4] * (-2.f * xy);
                Bz[13] = SH_C3[4] * 8.f * xz;
                Bx[14] = SH_C3[5] * 2.f * xz; By[14] = SH_C3[5] * (-2.f * yz); Bz[14] = SH_C3[5] * (xx - yy);
                Bx[15] = SH_C3[6] * (3.f * xx - 3.f * yy); By[15] = SH_C3[6] * (-6.f * xy);
            }
        }
    }
}

/* =====================================================================
 * A.1  preprocess forward.  Any output pointer may be NULL except radii.
 * ===================================================================== */
int gmso_preprocess_forward(const gmso_settings* s,
                            const float* means3D,      /* [P,3] */
                            const float* scales,       /* [P,3] or NULL */
                            const float* rotations,    /* [P,4] or NULL */
                            const float* cov3D_precomp,/* [P,6] or NULL */
                            const float* opacities,    /* [P] */
                            const float* shs,          /* [P,M,3] or NULL */
                            const float* colors_precomp,/* [P,3] or NULL */
                            int32_t* radii, float* means2D /*[P,2]*/, float* depths,
                            float* cov3Ds /*[P,6]*/, float* conic_opacity /*[P,4]*/,
                            float* rgb /*[P,3]*/, uint8_t* clamped /*[P,3]*/,
                            uint32_t* tiles_touched, int32_t* rects /*[P,4] xmin,ymin,xmax,ymax*/) {
    const int P = s->P, W = s->W, H = s->H;
    const int gx = (W + GMSO_BLOCK - 1) / GMSO_BLOCK, gy = (H + GMSO_BLOCK - 1) / GMSO_BLOCK;
    const float focal_x = (float)W / (2.0f * s->tanfovx);
    const float focal_y = (float)H / (2.0f * s->tanfovy);
    if (s->D < 0 || s->D > 3) return -1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        if (tiles_touched) tiles_touched[i] = 0;
        if (rects) { rects[4 * i] = rects[4 * i + 1] = rects[4 * i + 2] = rects[4 * i + 3] = 0; }
        if (means2D) { means2D[2 * i] = 0.f; means2D[2 * i + 1] = 0.f; }
        if (depths) depths[i] = 0.f;
        if (conic_opacity) { for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f; }
        if (rgb) { for (int k = 0; k < 3; k++) rgb[3 * i + k] = 0.f; }
        if (clamped) { for (int k = 0; k < 3; k++) clamped[3 * i + k] = 0; }
        if (cov3Ds) { for (int k = 0; k < 6; k++) cov3Ds[6 * i + k] = 0.f; }

        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        float pview[3];
        xform4x3(s->viewmatrix, x, y, z, pview);
        if (pview[2] <= 0.2f) continue; /* near cull (in_frustum) */
        float phom[4];
        xform4x4(s->projmatrix, x, y, z, phom);
        float pw = 1.0f / (phom[3] + 0.0000001f);
        float pprojx = phom[0] * pw, pprojy = phom[1] * pw;

        float cov6[6];
        if (cov3D_precomp) {
            for (int k = 0; k < 6; k++) cov6[k] = cov3D_precomp[6 * i + k];
        } else {
            cov3d_from_scale_rot(scales + 3 * i, s->scale_modifier, rotations + 4 * i, cov6);
        }
        if (cov3Ds) { for (int k = 0; k < 6; k++) cov3Ds[6 * i + k] = cov6[k]; }

        cov2d_aux ca;
        cov2d_project(pview, cov6, s->viewmatrix, focal_x, focal_y, s->tanfovx, s->tanfovy, &ca);
        float a = ca.a, b = ca.b, c = ca.c;
        const float h_var = 0.3f;
        float det_cov = fmaf(a, c, -(b * b));
        a = a + h_var; c = c + h_var;
        float det = fmaf(a, c, -(b * b));
        if (det == 0.0f) continue;
        float h_scale = 1.0f;
        if (s->antialiasing) h_scale = sqrtf(fmaxf(0.000025f, det_cov / det));
        float det_inv = 1.f / det;
        float conx = c * det_inv, cony = -b * det_inv, conz = a * det_inv;

        float mid = 0.5f * (a + c);
        float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float lambda1 = mid + disc, lambda2 = mid - disc;
        int my_radius = (int)ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float pixx = fmaf(pprojx + 1.0f, (float)W, -1.0f) * 0.5f;
        float pixy = fmaf(pprojy + 1.0f, (float)H, -1.0f) * 0.5f;
        int rmin[2], rmax[2];
        get_rect(pixx, pixy, my_radius, gx, gy, rmin, rmax);
        int area = (rmax[0] - rmin[0]) * (rmax[1] - rmin[1]);
        if (area == 0) continue;

        if (rgb) {
            if (colors_precomp) {
                for (int k = 0; k < 3; k++) rgb[3 * i + k] = colors_precomp[3 * i + k];
            } else {
                float dx = x - s->campos[0], dy = y - s->campos[1], dz = z - s->campos[2];
                float len = sqrtf(dx * dx + dy * dy + dz * dz);
                dx /= len; dy /= len; dz /= len;
                float B[16];
                sh_basis(s->D, dx, dy, dz, B);
                int nc = (s->D + 1) * (s->D + 1);
                const float* sh = shs + (size_t)i * s->M * 3;
                for (int ch = 0; ch < 3; ch++) {
                    float acc = 0.f;
                    for (int k = 0; k < nc; k++) acc += B[k] * sh[3 * k + ch];
                    acc += 0.5f;
                    if (clamped) clamped[3 * i + ch] = (acc < 0.f) ? 1 : 0;
                    rgb[3 * i + ch] = fmaxf(acc, 0.f);
                }
            }
        }
        if (depths) depths[i] = pview[2];
        radii[i] = my_radius;
        if (means2D) { means2D[2 * i] = pixx; means2D[2 * i + 1] = pixy; }
        if (conic_opacity) {
            conic_opacity[4 * i] = conx; conic_opacity[4 * i + 1] = cony; conic_opacity[4 * i + 2] = conz;
            conic_opacity[4 * i + 3] = opacities[i] * h_scale;
        }
        if (tiles_touched) tiles_touched[i] = (uint32_t)area;
        if (rects) { rects[4 * i] = rmin[0]; rects[4 * i + 1] = rmin[1]; rects[4 * i + 2] = rmax[0]; rects[4 * i + 3] = rmax[1]; }
    }
    return 0;
}

/* number of key bits for the tile id (upstream getHigherMsb) */
int gmso_tile_bits(int ntiles) {
    uint32_t n = (uint32_t)ntiles;
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}

/* =====================================================================
 * Binning: inclusive scan, duplicate-with-keys, stable sort, tile ranges.
 * offsets[P] is written; returns N.  If keys_sorted/point_list are NULL only
 * the count is produced.  Capacity of the key arrays must be >= N.
 * ===================================================================== */
int64_t gmso_bin(const gmso_settings* s, const int32_t* radii, const float* depths,
                 const int32_t* rects, const uint32_t* tiles_touched, uint32_t* offsets,
                 uint64_t* keys_sorted, uint32_t* point_list, int32_t* ranges /*[T,2]*/,
                 int64_t capacity) {
    const int P = s->P;
    const int gx = (s->W + GMSO_BLOCK - 1) / GMSO_BLOCK, gy = (s->H + GMSO_BLOCK - 1) / GMSO_BLOCK;
    uint64_t run = 0;
    for (int i = 0; i < P; i++) { run += tiles_touched[i]; offsets[i] = (uint32_t)run; }
    int64_t N = (int64_t)run;
    if (!keys_sorted || !point_list) return N;
    if (N > capacity) return -N;
    uint64_t* k0 = keys_sorted; uint32_t* v0 = point_list;
    uint64_t* k1 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
    uint32_t* v1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N > 0 ? N : 1));
    /* duplicateWithKeys */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint32_t off = (i == 0) ? 0u : offsets[i - 1];
        uint32_t dbits; memcpy(&dbits, &depths[i], 4);
        for (int y = rects[4 * i + 1]; y < rects[4 * i + 3]; y++)
            for (int x = rects[4 * i]; x < rects[4 * i + 2]; x++) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key = (key << 32) | dbits;
                k0[off] = key; v0[off] = (uint32_t)i; off++;
            }
    }
    /* stable LSD radix sort, 16 bits per pass, over bits [0, 32+tile_bits) */
    int nbits = 32 + gmso_tile_bits(gx * gy);
    int npass = (nbits + 15) / 16;
    uint32_t* hist = (uint32_t*)malloc(sizeof(uint32_t) * 65536);
    uint64_t* ka = k0; uint32_t* va = v0; uint64_t* kb = k1; uint32_t* vb = v1;
    for (int p = 0; p < npass; p++) {
        int shift = 16 * p;
        memset(hist, 0, sizeof(uint32_t) * 65536);
        for (int64_t j = 0; j < N; j++) hist[(ka[j] >> shift) & 0xFFFF]++;
        uint32_t sum = 0;
        for (int d = 0; d < 65536; d++) { uint32_t c = hist[d]; hist[d] = sum; sum += c; }
        for (int64_t j = 0; j < N; j++) {
            uint32_t d = (uint32_t)((ka[j] >> shift) & 0xFFFF);
            uint32_t pos = hist[d]++;
            kb[pos] = ka[j]; vb[pos] = va[j];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != k0) { memcpy(k0, ka, sizeof(uint64_t) * (size_t)N); memcpy(v0, va, sizeof(uint32_t) * (size_t)N); }
    free(hist); free(k1); free(v1);
    /* identifyTileRanges */
    if (ranges) {
        memset(ranges, 0, sizeof(int32_t) * 2 * (size_t)(gx * gy));
        for (int64_t j = 0; j < N; j++) {
            uint32_t t = (uint32_t)(k0[j] >> 32);
            if (j == 0) ranges[2 * t] = 0;
            else {
                uint32_t tp = (uint32_t)(k0[j - 1] >> 32);
                if (tp != t) { ranges[2 * tp + 1] = (int32_t)j; ranges[2 * t] = (int32_t)j; }
            }
            if (j == N - 1) ranges[2 * t + 1] = (int32_t)N;
        }
    }
    return N;
}

/* =====================================================================
 * A.2 composite forward (one pixel at a time, the tile's list front to back).
 * ambiguous[px] (optional) is set when a skip/stop decision was within a few
 * ulp of its threshold, i.e. a different-but-valid exp() could flip it.
 * ===================================================================== */
int gmso_composite_forward(const gmso_settings* s, const int32_t* ranges, const uint32_t* point_list,
                           const float* means2D, const float* rgb, const float* conic_opacity,
                           const float* depths,
                           float* out_color /*[3,H,W]*/, float* final_T /*[H,W]*/,
                           int32_t* n_contrib /*[H,W]*/, float* out_invdepth /*[H,W] or NULL*/,
                           uint8_t* ambiguous /*[H,W] or NULL*/) {
    const int W = s->W, H = s->H;
    const int gx = (W + GMSO_BLOCK - 1) / GMSO_BLOCK, gy = (H + GMSO_BLOCK - 1) / GMSO_BLOCK;
    long long st0 = 0, st1 = 0, st2 = 0, st3 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : st0, st1, st2, st3)
    for (int tile = 0; tile < gx * gy; tile++) {
        if (tile % g_tile_stride) continue;
        int tx0 = (tile % gx) * GMSO_BLOCK, ty0 = (tile / gx) * GMSO_BLOCK;
        int r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        uint8_t* qhit = (uint8_t*)calloc((size_t)(r1 - r0) + 1, 1);
        for (int py = ty0; py < ty0 + GMSO_BLOCK && py < H; py++)
            for (int px = tx0; px < tx0 + GMSO_BLOCK && px < W; px++) {
                const int quad = ((py - ty0) >> 3) * 2 + ((px - tx0) >> 3);
                float T = 1.0f, C[3] = {0.f, 0.f, 0.f}, Dacc = 0.f;
                int contributor = 0, last = 0; uint8_t amb = 0;
                float pfx = (float)px, pfy = (float)py;
                for (int j = r0; j < r1; j++) {
                    contributor++; st0++;
                    uint32_t g = point_list[j];
                    float dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
                    const float* co = conic_opacity + 4 * g;
                    float power = quad_power(co, dx, dy);
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * expf(power));
                    if (fabsf(alpha - 1.0f / 255.0f) < 4e-6f * (1.0f / 255.0f) + 1e-9f) amb = 1;
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1.f - alpha);
                    if (fabsf(test_T - 0.0001f) < 1e-9f) amb = 1;
                    if (test_T < 0.0001f) break;
                    float w = alpha * T;
                    for (int ch = 0; ch < 3; ch++) C[ch] += rgb[3 * g + ch] * w;
                    Dacc += (1.f / depths[g]) * w;
                    T = test_T;
                    last = contributor;
                    st1++; qhit[j - r0] |= (uint8_t)(1u << quad);
                }
                size_t pix = (size_t)py * W + px;
                final_T[pix] = T; n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pix] = C[ch] + T * s->bg[ch];
                if (out_invdepth) out_invdepth[pix] = Dacc;
                if (ambiguous) ambiguous[pix] = amb;
            }
        for (int j = 0; j < r1 - r0; j++) {
            if (qhit[j]) st3++;
            st2 += ((qhit[j] >> 0) & 1) + ((qhit[j] >> 1) & 1) + ((qhit[j] >> 2) & 1) + ((qhit[j] >> 3) & 1);
        }
        free(qhit);
    }
    g_stats[0] = st0; g_stats[1] = st1; g_stats[2] = st2; g_stats[3] = st3;
    return 0;
}

/* =====================================================================
 * A.3 composite backward.  Accumulates in double (this is the checker).
 * Outputs are per-Gaussian: dL_dmean2D[P,2] (NDC-scaled: *0.5W / *0.5H),
 * dL_dconic[P,3] = (xx, xy-HALF-convention, yy), dL_dopacity[P] (w.r.t.
 * conic_opacity.w), dL_dcolor[P,3], dL_dinvdepth[P].
 * ===================================================================== */
int gmso_composite_backward(const gmso_settings* s, const int32_t* ranges, const uint32_t* point_list,
                            const float* means2D, const float* rgb, const float* conic_opacity,
                            const float* depths, const float* final_T, const int32_t* n_contrib,
                            const float* dL_dpix /*[3,H,W]*/, const float* dL_dinvdepth_pix /*[H,W] or NULL*/,
                            double* dL_dmean2D, double* dL_dconic, double* dL_dopacity,
                            double* dL_dcolor, double* dL_dinvdepth) {
    const int W = s->W, H = s->H, P = s->P;
    const int gx = (W + GMSO_BLOCK - 1) / GMSO_BLOCK, gy = (H + GMSO_BLOCK - 1) / GMSO_BLOCK;
    memset(dL_dmean2D, 0, sizeof(double) * 2 * (size_t)P);
    memset(dL_dconic, 0, sizeof(double) * 3 * (size_t)P);
    memset(dL_dopacity, 0, sizeof(double) * (size_t)P);
    memset(dL_dcolor, 0, sizeof(double) * 3 * (size_t)P);
    memset(dL_dinvdepth, 0, sizeof(double) * (size_t)P);
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    /* serial over tiles: per-Gaussian accumulators are shared between tiles */
    int nth = gmso_num_threads();
    (void)nth;
#pragma omp parallel
    {
        double* l_m = (double*)calloc((size_t)P * 10, sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int tile = 0; tile < gx * gy; tile++) {
            if (tile % g_tile_stride) continue;
            int tx0 = (tile % gx) * GMSO_BLOCK, ty0 = (tile / gx) * GMSO_BLOCK;
            int r0 = ranges[2 * tile];
            for (int py = ty0; py < ty0 + GMSO_BLOCK && py < H; py++)
                for (int px = tx0; px < tx0 + GMSO_BLOCK && px < W; px++) {
                    size_t pix = (size_t)py * W + px;
                    float Tfin = final_T[pix];
                    float T = Tfin;
                    int last = n_contrib[pix];
                    float dpix[3];
                    for (int ch = 0; ch < 3; ch++) dpix[ch] = dL_dpix[(size_t)ch * H * W + pix];
                    float dinv = dL_dinvdepth_pix ? dL_dinvdepth_pix[pix] : 0.f;
                    float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
                    float accum_inv = 0.f, last_inv = 0.f, last_alpha = 0.f;
                    float bg_dot = s->bg[0] * dpix[0] + s->bg[1] * dpix[1] + s->bg[2] * dpix[2];
                    float pfx = (float)px, pfy = (float)py;
                    for (int j = r0 + last - 1; j >= r0; j--) {
                        uint32_t g = point_list[j];
                        float dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
                        const float* co = conic_opacity + 4 * g;
                        float power = quad_power(co, dx, dy);
                        if (power > 0.0f) continue;
                        float G = expf(power);
                        float alpha = fminf(0.99f, co[3] * G);
                        if (alpha < 1.0f / 255.0f) continue;
                        T = T / (1.f - alpha);
                        float dch = alpha * T;
                        float dL_dalpha = 0.f;
                        double* acc = l_m + (size_t)g * 10;
                        for (int ch = 0; ch < 3; ch++) {
                            float c = rgb[3 * g + ch];
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                            last_color[ch] = c;
                            dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                            acc[6 + ch] += (double)(dch * dpix[ch]);
                        }
                        {
                            float invd = 1.f / depths[g];
                            accum_inv = last_alpha * last_inv + (1.f - last_alpha) * accum_inv;
                            last_inv = invd;
                            dL_dalpha += (invd - accum_inv) * dinv;
                            acc[9] += (double)(dch * dinv);
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-Tfin / (1.f - alpha)) * bg_dot;
                        float dL_dG = co[3] * dL_dalpha;
                        float gdx = G * dx, gdy = G * dy;
                        float dG_ddelx = -gdx * co[0] - gdy * co[1];
                        float dG_ddely = -gdy * co[2] - gdx * co[1];
                        acc[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                        acc[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                        acc[2] += (double)(-0.5f * gdx * dx * dL_dG);
                        acc[3] += (double)(-0.5f * gdx * dy * dL_dG);
                        acc[4] += (double)(-0.5f * gdy * dy * dL_dG);
                        acc[5] += (double)(G * dL_dalpha);
                    }
                }
        }
#pragma omp critical
        {
            for (int g = 0; g < P; g++) {
                const double* a = l_m + (size_t)g * 10;
                dL_dmean2D[2 * g] += a[0]; dL_dmean2D[2 * g + 1] += a[1];
                dL_dconic[3 * g] += a[2]; dL_dconic[3 * g + 1] += a[3]; dL_dconic[3 * g + 2] += a[4];
                dL_dopacity[g] += a[5];
                dL_dcolor[3 * g] += a[6]; dL_dcolor[3 * g + 1] += a[7]; dL_dcolor[3 * g + 2] += a[8];
                dL_dinvdepth[g] += a[9];
            }
        }
        free(l_m);
    }
    return 0;
}

/* =====================================================================
 * A.4 preprocess backward (computeCov2D backward + projection + SH + cov3D).
 * Inputs are the per-Gaussian gradients from the composite backward (float).
 * ===================================================================== */
int gmso_preprocess_backward(const gmso_settings* s, const int32_t* radii,
                             const float* means3D, const float* scales, const float* rotations,
                             const float* cov3D_precomp, const float* opacities, const float* shs,
                             const float* colors_precomp, const float* cov3Ds, const uint8_t* clamped,
                             const float* dL_dmean2D /*[P,2]*/, const float* dL_dconic /*[P,3]*/,
                             const float* dL_dopacity_in /*[P] wrt conic_opacity.w*/,
                             const float* dL_dcolor /*[P,3]*/, const float* dL_dinvdepth /*[P] or NULL*/,
                             float* dL_dmeans3D /*[P,3]*/, float* dL_dcov3D /*[P,6]*/,
                             float* dL_dsh /*[P,M,3] or NULL*/, float* dL_dcolors_precomp /*[P,3] or NULL*/,
                             float* dL_dscale /*[P,3] or NULL*/, float* dL_drot /*[P,4] or NULL*/,
                             float* dL_dopacity_out /*[P]*/) {
    const int P = s->P, W = s->W, H = s->H;
    const float focal_x = (float)W / (2.0f * s->tanfovx);
    const float focal_y = (float)H / (2.0f * s->tanfovy);
    const float* view = s->viewmatrix; const float* proj = s->projmatrix;
    (void)colors_precomp; (void)cov3D_precomp;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] = 0.f;
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
        if (dL_dsh) memset(dL_dsh + (size_t)i * s->M * 3, 0, sizeof(float) * 3 * (size_t)s->M);
        if (dL_dcolors_precomp) for (int k = 0; k < 3; k++) dL_dcolors_precomp[3 * i + k] = 0.f;
        if (dL_dscale) for (int k = 0; k < 3; k++) dL_dscale[3 * i + k] = 0.f;
        if (dL_drot) for (int k = 0; k < 4; k++) dL_drot[4 * i + k] = 0.f;
        dL_dopacity_out[i] = 0.f;
        if (radii[i] <= 0) continue;

        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        const float* cov6 = cov3Ds + 6 * i;
        float pview[3];
        xform4x3(view, x, y, z, pview);
        cov2d_aux ca;
        cov2d_project(pview, cov6, view, focal_x, focal_y, s->tanfovx, s->tanfovy, &ca);
        const float h_var = 0.3f;
        float a0 = ca.a, b = ca.b, c0 = ca.c;       /* raw */
        float a = a0 + h_var, c = c0 + h_var;        /* dilated */
        float det_cov = a0 * c0 - b * b;
        float det = a * c - b * b;

        float dcx = dL_dconic[3 * i], dcy = dL_dconic[3 * i + 1], dcz = dL_dconic[3 * i + 2];
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;

        /* opacity (and antialiasing compensation) */
        float dop = dL_dopacity_in[i];
        if (s->antialiasing) {
            float ratio = det_cov / det;
            float h_scale = sqrtf(fmaxf(0.000025f, ratio));
            dL_dopacity_out[i] = dop * h_scale;
            if (ratio > 0.000025f) {
                /* d L / d ratio, ratio = det_cov/det; both depend on raw a0,b,c0 */
                float dL_dratio = dop * opacities[i] / (2.f * h_scale);
                float inv_det = 1.f / det;
                /* d ratio / d a0 = (c0*det - det_cov*c)/det^2 etc. */
                dL_da += dL_dratio * (c0 * det - det_cov * c) * inv_det * inv_det;
                dL_dc += dL_dratio * (a0 * det - det_cov * a) * inv_det * inv_det;
                dL_db += dL_dratio * (-2.f * b * det + 2.f * b * det_cov) * inv_det * inv_det;
            }
        } else {
            dL_dopacity_out[i] = dop;
        }

        float denom = det;
        float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0.f) {
            dL_da += denom2inv * (-c * c * dcx + 2.f * b * c * dcy + (denom - a * c) * dcz);
            dL_dc += denom2inv * (-a * a * dcz + 2.f * a * b * dcy + (denom - a * c) * dcx);
            dL_db += denom2inv * 2.f * (b * c * dcx - (denom + 2.f * b * b) * dcy + a * b * dcz);
        }
        const float* M0 = ca.M0; const float* M1 = ca.M1;
        float g6[6];
        g6[0] = M0[0] * M0[0] * dL_da + M0[0] * M1[0] * dL_db + M1[0] * M1[0] * dL_dc;
        g6[3] = M0[1] * M0[1] * dL_da + M0[1] * M1[1] * dL_db + M1[1] * M1[1] * dL_dc;
        g6[5] = M0[2] * M0[2] * dL_da + M0[2] * M1[2] * dL_db + M1[2] * M1[2] * dL_dc;
        g6[1] = 2.f * M0[0] * M0[1] * dL_da + (M0[0] * M1[1] + M0[1] * M1[0]) * dL_db + 2.f * M1[0] * M1[1] * dL_dc;
        g6[2] = 2.f * M0[0] * M0[2] * dL_da + (M0[0] * M1[2] + M0[2] * M1[0]) * dL_db + 2.f * M1[0] * M1[2] * dL_dc;
        g6[4] = 2.f * M0[2] * M0[1] * dL_da + (M0[1] * M1[2] + M0[2] * M1[1]) * dL_db + 2.f * M1[1] * M1[2] * dL_dc;
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = g6[k];

        /* dL/dM = [2 da M0 + db M1 ; 2 dc M1 + db M0] * Sigma */
        float S[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
        float u0[3], u1[3], dM0[3], dM1[3];
        for (int j = 0; j < 3; j++) {
            u0[j] = 2.f * dL_da * M0[j] + dL_db * M1[j];
            u1[j] = 2.f * dL_dc * M1[j] + dL_db * M0[j];
        }
        for (int j = 0; j < 3; j++) {
            dM0[j] = u0[0] * S[0 * 3 + j] + u0[1] * S[1 * 3 + j] + u0[2] * S[2 * 3 + j];
            dM1[j] = u1[0] * S[0 * 3 + j] + u1[1] * S[1 * 3 + j] + u1[2] * S[2 * 3 + j];
        }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
        for (int j = 0; j < 3; j++) {
            float W0j = view[4 * j + 0], W1j = view[4 * j + 1], W2j = view[4 * j + 2];
            dJ00 += dM0[j] * W0j; dJ02 += dM0[j] * W2j;
            dJ11 += dM1[j] * W1j; dJ12 += dM1[j] * W2j;
        }
        float tz = 1.f / ca.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        float dL_dtx = ca.xmul * -focal_x * tz2 * dJ02;
        float dL_dty = ca.ymul * -focal_y * tz2 * dJ12;
        float dL_dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 +
                       (2.f * focal_x * ca.tx) * tz3 * dJ02 + (2.f * focal_y * ca.ty) * tz3 * dJ12;
        if (dL_dinvdepth) dL_dtz -= dL_dinvdepth[i] / (pview[2] * pview[2]);
        /* transformVec4x3Transpose: dmean = W^T dt */
        float dmx = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        float dmy = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        float dmz = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        /* mean2D -> mean3D through the projection */
        float phom[4];
        xform4x4(proj, x, y, z, phom);
        float m_w = 1.0f / (phom[3] + 0.0000001f);
        float mul1 = phom[0] * m_w * m_w;
        float mul2 = phom[1] * m_w * m_w;
        float g2x = dL_dmean2D[2 * i], g2y = dL_dmean2D[2 * i + 1];
        dmx += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dmy += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dmz += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

        /* colour */
        if (shs && dL_dsh) {
            float dx = x - s->campos[0], dy = y - s->campos[1], dz = z - s->campos[2];
            float len = sqrtf(dx * dx + dy * dy + dz * dz);
            float ux = dx / len, uy = dy / len, uz = dz / len;
            float B[16], Bx[16], By[16], Bz[16];
            sh_basis(s->D, ux, uy, uz, B);
            sh_basis_grad(s->D, ux, uy, uz, Bx, By, Bz);
            int nc = (s->D + 1) * (s->D + 1);
            const float* sh = shs + (size_t)i * s->M * 3;
            float* dsh = dL_dsh + (size_t)i * s->M * 3;
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
            for (int ch = 0; ch < 3; ch++) {
                float g = dL_dcolor[3 * i + ch] * (clamped[3 * i + ch] ? 0.f : 1.f);
                for (int k = 0; k < nc; k++) {
                    dsh[3 * k + ch] = B[k] * g;
                    ddx += Bx[k] * sh[3 * k + ch] * g;
                    ddy += By[k] * sh[3 * k + ch] * g;
                    ddz += Bz[k] * sh[3 * k + ch] * g;
                }
            }
            /* through v/|v| */
            float dotp = ux * ddx + uy * ddy + uz * ddz;
            dmx += (ddx - ux * dotp) / len;
            dmy += (ddy - uy * dotp) / len;
            dmz += (ddz - uz * dotp) / len;
        } else if (dL_dcolors_precomp) {
            for (int k = 0; k < 3; k++) dL_dcolors_precomp[3 * i + k] = dL_dcolor[3 * i + k];
        }
        dL_dmeans3D[3 * i] = dmx; dL_dmeans3D[3 * i + 1] = dmy; dL_dmeans3D[3 * i + 2] = dmz;

        /* cov3D -> scale, rotation */
        if (scales && rotations && dL_dscale && dL_drot) {
            float R[9];
            const float* q = rotations + 4 * i;
            quat_to_R(q, R);
            float mod = s->scale_modifier;
            float sv[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
            float Mx[9];
            for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) Mx[3 * r + cc] = R[3 * r + cc] * sv[cc];
            float dS[9] = {g6[0], 0.5f * g6[1], 0.5f * g6[2], 0.5f * g6[1], g6[3], 0.5f * g6[4],
                           0.5f * g6[2], 0.5f * g6[4], g6[5]};
            float dMx[9];
            for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++)
                dMx[3 * r + cc] = 2.f * (dS[3 * r] * Mx[cc] + dS[3 * r + 1] * Mx[3 + cc] + dS[3 * r + 2] * Mx[6 + cc]);
            float dR[9];
            for (int cc = 0; cc < 3; cc++) {
                float acc = 0.f;
                for (int r = 0; r < 3; r++) { acc += R[3 * r + cc] * dMx[3 * r + cc]; dR[3 * r + cc] = dMx[3 * r + cc] * sv[cc]; }
                dL_dscale[3 * i + cc] = acc * mod;
            }
            float qr = q[0], qx = q[1], qy = q[2], qz = q[3];
            dL_drot[4 * i + 0] = 2.f * (-qz * dR[1] + qy * dR[2] + qz * dR[3] - qx * dR[5] - qy * dR[6] + qx * dR[7]);
            dL_drot[4 * i + 1] = 2.f * (-2.f * qx * (dR[4] + dR[8]) + qy * (dR[1] + dR[3]) + qz * (dR[2] + dR[6]) + qr * (dR[7] - dR[5]));
            dL_drot[4 * i + 2] = 2.f * (-2.f * qy * (dR[0] + dR[8]) + qx * (dR[1] + dR[3]) + qr * (dR[2] - dR[6]) + qz * (dR[5] + dR[7]));
            dL_drot[4 * i + 3] = 2.f * (-2.f * qz * (dR[0] + dR[4]) + qr * (dR[3] - dR[1]) + qx * (dR[2] + dR[6]) + qy * (dR[5] + dR[7]));
        }
    }
    return 0;
}

/* markVisible / checkFrustum (R10) */
int gmso_mark_visible(const gmso_settings* s, const float* means3D, uint8_t* present) {
    for (int i = 0; i < s->P; i++) {
        float pview[3];
        xform4x3(s->viewmatrix, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2], pview);
        present[i] = pview[2] > 0.2f ? 1 : 0;
    }
    return 0;
}
