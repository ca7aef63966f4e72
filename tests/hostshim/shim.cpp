// tests/hostshim/shim.cpp -- compiles the PRODUCT's host+device maths headers (csrc/*.cuh) with g++ so the
// per-element expansion / preprocess maths can be unit-tested against the oracle on a machine with no GPU.
// TEST-ONLY: never loaded by the product; the product path is the CUDA kernels in gms_kernels.cu.
#include <stdint.h>
#include <string.h>
#include "../../gaussian-mesh-splatting_b200/csrc/gms_common.cuh"
#include "../../gaussian-mesh-splatting_b200/csrc/gms_preprocess.cuh"
#include "../../gaussian-mesh-splatting_b200/csrc/gms_expand.cuh"

extern "C" {

int shim_expand_forward(const gms_expand_args* a) {
    for (int f = 0; f < a->F; f++) gms_expand_face_fwd(*a, f, f);
    return 0;
}

int shim_expand_backward(const gms_expand_args* a, const gms_expand_grads* g) {
    for (int f = 0; f < a->F; f++) gms_expand_face_bwd(*a, *g, f, f);
    return 0;
}

int shim_points_expand_forward(const gms_points_args* a) {
    for (int i = 0; i < a->P; i++) gms_points_face_fwd(*a, i);
    return 0;
}

int shim_points_prepare_vertices(const gms_points_vertices_args* a) {
    for (int i = 0; i < a->P; i++) gms_points_vertices_fwd(*a, i);
    return 0;
}

// stock-layout outputs; mirrors the glue of k_preprocess_fwd
int shim_preprocess_forward(int P, int D, int M, int W, int H, float tanfovx, float tanfovy, float mod, int aa,
                            const float* means, const float* scales, const float* rots, const float* cov_pre,
                            const float* opac, const float* shs, const float* colors_pre, const float* view,
                            const float* proj, const float* campos, int32_t* radii, float* means2D, float* depths,
                            float* cov3D, float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles) {
    const int gx = (W + GMS_TILE - 1) / GMS_TILE, gy = (H + GMS_TILE - 1) / GMS_TILE;
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
    for (int i = 0; i < P; i++) {
        GmsPre o;
        const bool vis = gms_preprocess_geom(means + 3 * i, scales ? scales + 3 * i : nullptr, rots ? rots + 4 * i : nullptr,
                                             cov_pre ? cov_pre + 6 * i : nullptr, opac[i], view, proj, W, H, tanfovx,
                                             tanfovy, fx, fy, mod, aa, gx, gy, o);
        radii[i] = 0; tiles[i] = 0;
        means2D[2 * i] = means2D[2 * i + 1] = 0.f; depths[i] = 0.f;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0.f;
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f;
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }
        if (!vis) continue;
        radii[i] = o.radius; tiles[i] = o.tiles;
        means2D[2 * i] = o.px; means2D[2 * i + 1] = o.py; depths[i] = o.depth;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = o.cov6[k];
        conic_opacity[4 * i] = o.conx; conic_opacity[4 * i + 1] = o.cony; conic_opacity[4 * i + 2] = o.conz;
        conic_opacity[4 * i + 3] = o.opac;
        if (shs) gms_sh_color(D, means + 3 * i, campos, shs + (size_t)i * M * 3, rgb + 3 * i, clamped + 3 * i);
        else for (int k = 0; k < 3; k++) rgb[3 * i + k] = colors_pre[3 * i + k];
    }
    return 0;
}

int shim_preprocess_backward(int P, int D, int M, int W, int H, float tanfovx, float tanfovy, float mod, int aa,
                             const int32_t* radii, const float* means, const float* scales, const float* rots,
                             const float* opac, const float* shs, const float* view, const float* proj,
                             const float* campos, const float* cov3D, const uint8_t* clamped,
                             const float* dmean2D, const float* dconic, const float* dopac_in, const float* dcolor,
                             const float* dinvdepth, float* dmeans3D, float* dcov3D, float* dsh, float* dscale,
                             float* drot, float* dopac_out) {
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) dmeans3D[3 * i + k] = 0.f;
        for (int k = 0; k < 6; k++) dcov3D[6 * i + k] = 0.f;
        if (dsh) memset(dsh + (size_t)i * M * 3, 0, sizeof(float) * 3 * M);
        if (dscale) for (int k = 0; k < 3; k++) dscale[3 * i + k] = 0.f;
        if (drot) for (int k = 0; k < 4; k++) drot[4 * i + k] = 0.f;
        dopac_out[i] = 0.f;
        if (radii[i] <= 0) continue;
        GmsPreGradIn gi;
        gi.dmean2D[0] = dmean2D[2 * i]; gi.dmean2D[1] = dmean2D[2 * i + 1];
        for (int k = 0; k < 3; k++) { gi.dconic[k] = dconic[3 * i + k]; gi.dcolor[k] = dcolor[3 * i + k]; }
        gi.dopac = dopac_in[i]; gi.dinvdepth = dinvdepth[i];
        GmsPreGradOut go;
        gms_preprocess_backward_geom(means + 3 * i, scales ? scales + 3 * i : nullptr, rots ? rots + 4 * i : nullptr,
                                     cov3D + 6 * i, opac[i], view, proj, tanfovx, tanfovy, fx, fy, mod, aa, gi, go);
        if (shs && dsh) {
            float tmp[48];
            gms_sh_backward(D, M < 16 ? M : 16, means + 3 * i, campos, shs + (size_t)i * M * 3, gi.dcolor,
                            clamped + 3 * i, tmp, go.dmean3D);
            for (int k = 0; k < 3 * (M < 16 ? M : 16); k++) dsh[(size_t)i * M * 3 + k] = tmp[k];
        }
        for (int k = 0; k < 3; k++) dmeans3D[3 * i + k] = go.dmean3D[k];
        for (int k = 0; k < 6; k++) dcov3D[6 * i + k] = go.dcov6[k];
        if (dscale) for (int k = 0; k < 3; k++) dscale[3 * i + k] = go.dscale[k];
        if (drot) for (int k = 0; k < 4; k++) drot[4 * i + k] = go.drot[k];
        dopac_out[i] = go.dopacity;
    }
    return 0;
}
}
