// refstyle.cu -- "ref-style" GPU comparator.  NOT the product, NOT the stock binary.
//
// The reference's rasterizer (graphdeco-inria/diff-gaussian-rasterization) is an empty, un-vendored submodule of
// /root/reference, so it cannot be built or timed.  SURVEY.md section 8(d) asks for a clearly labelled stand-in:
// the algorithm of SURVEY.md Appendix A with the STOCK work decomposition, compiled for sm_100a without fast-math:
//   * preprocess: one thread per Gaussian, per-lane scalar SH row reads, separate SoA outputs
//   * inclusive scan of tiles_touched, BLOCKING device->host copy of N, binning buffers resized from it
//   * duplicateWithKeys: one thread per Gaussian loops over its tile rectangle, 64-bit (tile << 32 | depth) keys
//   * ONE cub::DeviceRadixSort over all N 64-bit keys, bits [0, 32 + bits(T))
//   * identifyTileRanges
//   * composite forward: one 16x16 CTA per tile, one thread per pixel, 256-splat batches through shared memory,
//     __syncthreads per batch, block leaves when every pixel is done
//   * composite backward: same decomposition back to front, PER-PIXEL atomicAdd of every gradient
//   * gradient buffers zero-filled per call, cov2D backward and preprocess backward as two kernels
// It lives in its own shared library (baseline/refstyle/librefstyle.so), is never linked into the product, and is
// parity-checked against the oracle (tests/test_gpu_refstyle.py) so that the timings compare equal results.
// Per-Gaussian maths comes from the product's host+device headers (the formulas are Appendix A's either way); the data
// flow, kernels and launch shapes here are the stock ones.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gms_b200.h"
#include "../../gaussian-mesh-splatting_b200/csrc/gms_common.cuh"
#include "../../gaussian-mesh-splatting_b200/csrc/gms_preprocess.cuh"

#define RS_BLOCK_X 16
#define RS_BLOCK_Y 16
#define RS_BLOCK (RS_BLOCK_X * RS_BLOCK_Y)

static thread_local char rs_err[256] = "";
static int rs_fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    snprintf(rs_err, sizeof(rs_err), "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
    return code;
}
#define RS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return rs_fail(GMS_E_CUDA, #call, e__); } while (0)
#define RS_LAUNCHED(name) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return rs_fail(GMS_E_CUDA, name, e__); } while (0)

static inline size_t rs_up(size_t x) { return (x + 127) / 128 * 128; }
template <typename T> static T* rs_take(char*& p, size_t n) { T* r = reinterpret_cast<T*>(p); p += rs_up(n * sizeof(T)); return r; }

struct RsGeom {     // stock GeometryState
    float* depths; uint8_t* clamped; float2* means2D; float* cov3D; float4* conic_opacity; float* rgb;
    uint32_t* point_offsets; uint32_t* tiles_touched; void* scan_space; size_t scan_bytes; size_t total;
};
static RsGeom rs_geom(void* base, int P) {
    RsGeom g; char* p = reinterpret_cast<char*>(base); const size_t n = P > 0 ? P : 1;
    g.depths = rs_take<float>(p, n); g.clamped = rs_take<uint8_t>(p, 3 * n); g.means2D = rs_take<float2>(p, n);
    g.cov3D = rs_take<float>(p, 6 * n); g.conic_opacity = rs_take<float4>(p, n); g.rgb = rs_take<float>(p, 3 * n);
    g.tiles_touched = rs_take<uint32_t>(p, n); g.point_offsets = rs_take<uint32_t>(p, n);
    g.scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, g.scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    g.scan_space = p; p += rs_up(g.scan_bytes);
    g.total = (size_t)(p - reinterpret_cast<char*>(base));
    return g;
}
struct RsBin {      // stock BinningState
    uint64_t* keys_unsorted; uint64_t* keys; uint32_t* vals_unsorted; uint32_t* vals; void* sort_space; size_t sort_bytes; size_t total;
};
static RsBin rs_bin(void* base, int64_t N) {
    RsBin b; char* p = reinterpret_cast<char*>(base); const size_t n = N > 0 ? (size_t)N : 1;
    b.keys_unsorted = rs_take<uint64_t>(p, n); b.keys = rs_take<uint64_t>(p, n);
    b.vals_unsorted = rs_take<uint32_t>(p, n); b.vals = rs_take<uint32_t>(p, n);
    b.sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b.sort_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    b.sort_space = p; p += rs_up(b.sort_bytes);
    b.total = (size_t)(p - reinterpret_cast<char*>(base));
    return b;
}
struct RsImg {      // stock ImageState
    uint2* ranges; uint32_t* n_contrib; float* accum_alpha; size_t total;
};
static RsImg rs_img(void* base, int W, int H) {
    RsImg m; char* p = reinterpret_cast<char*>(base); const size_t n = (size_t)W * H;
    m.accum_alpha = rs_take<float>(p, n); m.n_contrib = rs_take<uint32_t>(p, n); m.ranges = rs_take<uint2>(p, n);
    m.total = (size_t)(p - reinterpret_cast<char*>(base));
    return m;
}

struct RsCam { float view[16]; float proj[16]; float campos[3]; };

// ------------------------------------------------------------------------------------------------ forward kernels
__global__ void __launch_bounds__(256)
rs_preprocess(int P, int D, int M, const float* __restrict__ means, const float* __restrict__ scales, float mod,
              const float* __restrict__ rots, const float* __restrict__ opac, const float* __restrict__ shs,
              const float* __restrict__ colors_pre, const float* __restrict__ cov_pre, const float* __restrict__ view,
              const float* __restrict__ proj, const float* __restrict__ campos, int W, int H, float tanfovx, float tanfovy,
              float focal_x, float focal_y, int antialiasing, int gx, int gy, int* __restrict__ radii, RsGeom g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    radii[i] = 0;
    g.tiles_touched[i] = 0;
    const float mean[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    float sc[3] = {0, 0, 0}, rt[4] = {1, 0, 0, 0}, cv[6];
    const float* cvp = nullptr;
    if (cov_pre) { for (int k = 0; k < 6; k++) cv[k] = cov_pre[6 * (size_t)i + k]; cvp = cv; }
    else {
        sc[0] = scales[3 * i]; sc[1] = scales[3 * i + 1]; sc[2] = scales[3 * i + 2];
        rt[0] = rots[4 * i]; rt[1] = rots[4 * i + 1]; rt[2] = rots[4 * i + 2]; rt[3] = rots[4 * i + 3];
    }
    GmsPre o;
    if (!gms_preprocess_geom(mean, sc, rt, cvp, opac[i], view, proj, W, H, tanfovx, tanfovy, focal_x, focal_y, mod,
                             antialiasing, gx, gy, o)) return;
    float rgb[3];
    uint8_t cl[3] = {0, 0, 0};
    if (shs) {
        float sh[48];
        const float* row = shs + (size_t)i * M * 3;     // stock: glm::vec3 per coefficient, scalar loads, 192 B lane stride
        const int nf = 3 * (D + 1) * (D + 1);
        for (int k = 0; k < nf; k++) sh[k] = row[k];
        gms_sh_color(D, mean, campos, sh, rgb, cl);
    } else { rgb[0] = colors_pre[3 * i]; rgb[1] = colors_pre[3 * i + 1]; rgb[2] = colors_pre[3 * i + 2]; }
    g.rgb[3 * i] = rgb[0]; g.rgb[3 * i + 1] = rgb[1]; g.rgb[3 * i + 2] = rgb[2];
    g.clamped[3 * i] = cl[0]; g.clamped[3 * i + 1] = cl[1]; g.clamped[3 * i + 2] = cl[2];
    for (int k = 0; k < 6; k++) g.cov3D[6 * (size_t)i + k] = o.cov6[k];
    g.depths[i] = o.depth;
    radii[i] = o.radius;
    g.means2D[i] = make_float2(o.px, o.py);
    g.conic_opacity[i] = make_float4(o.conx, o.cony, o.conz, o.opac);
    g.tiles_touched[i] = o.tiles;
}

__global__ void __launch_bounds__(256)
rs_duplicate_with_keys(int P, int gx, int gy, const float2* __restrict__ means2D, const float* __restrict__ depths,
                       const uint32_t* __restrict__ offsets, const int* __restrict__ radii,
                       uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    uint32_t off = i == 0 ? 0u : offsets[i - 1];
    int x0, y0, x1, y1;
    gms_get_rect(means2D[i].x, means2D[i].y, radii[i], gx, gy, &x0, &y0, &x1, &y1);
    const uint64_t dbits = (uint64_t)__float_as_uint(depths[i]);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
            vals[off] = (uint32_t)i;
            off++;
        }
}

__global__ void __launch_bounds__(256)
rs_identify_tile_ranges(int64_t N, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const uint32_t t = (uint32_t)(keys[j] >> 32);
    if (j == 0) ranges[t].x = 0;
    else {
        const uint32_t tp = (uint32_t)(keys[j - 1] >> 32);
        if (tp != t) { ranges[tp].y = (uint32_t)j; ranges[t].x = (uint32_t)j; }
    }
    if (j == N - 1) ranges[t].y = (uint32_t)N;
}

__device__ __forceinline__ float rs_power(float cx, float cy, float cz, float dx, float dy) {
    // the oracle's canonical evaluation order of  -0.5 (cx dx^2 + cz dy^2) - cy dx dy  
    const float t = __fmaf_rn(__fmul_rn(cz, dy), dy, __fmul_rn(__fmul_rn(cx, dx), dx));
    return __fmaf_rn(-__fmul_rn(cy, dx), dy, __fmul_rn(-0.5f, t));
}

__global__ void __launch_bounds__(RS_BLOCK)
rs_render_forward(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                  const float2* __restrict__ means2D, const float* __restrict__ rgb, const float4* __restrict__ conic_opacity,
                  const float* __restrict__ depths, const float* __restrict__ bg, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_invdepth) {
    __shared__ int s_id[RS_BLOCK];
    __shared__ float2 s_xy[RS_BLOCK];
    __shared__ float4 s_co[RS_BLOCK];
    const int gx = (W + RS_BLOCK_X - 1) / RS_BLOCK_X;
    const int px = blockIdx.x * RS_BLOCK_X + threadIdx.x, py = blockIdx.y * RS_BLOCK_Y + threadIdx.y;
    const int tid = threadIdx.y * RS_BLOCK_X + threadIdx.x;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 rng = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = (int)((rng.y - rng.x + RS_BLOCK - 1) / RS_BLOCK);
    int todo = (int)(rng.y - rng.x);
    bool done = !inside;
    float T = 1.f, C[3] = {0.f, 0.f, 0.f}, Dacc = 0.f;
    uint32_t contributor = 0, last = 0;
    for (int r = 0; r < rounds; r++, todo -= RS_BLOCK) {
        if (__syncthreads_count(done) == RS_BLOCK) break;
        const int k = r * RS_BLOCK + tid;
        if (rng.x + k < rng.y) {
            const int id = (int)point_list[rng.x + k];
            s_id[tid] = id; s_xy[tid] = means2D[id]; s_co[tid] = conic_opacity[id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(RS_BLOCK, todo); j++) {
            contributor++;
            const float2 xy = s_xy[j]; const float4 co = s_co[j];
            const float power = rs_power(co.x, co.y, co.z, __fsub_rn(xy.x, pxf), __fsub_rn(xy.y, pyf));
            if (power > 0.f) continue;
            const float alpha = fminf(0.99f, __fmul_rn(co.w, expf(power)));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
            if (test_T < 0.0001f) { done = true; continue; }
            const int id = s_id[j];
            const float w = __fmul_rn(alpha, T);
            C[0] = __fmaf_rn(rgb[3 * id], w, C[0]); C[1] = __fmaf_rn(rgb[3 * id + 1], w, C[1]); C[2] = __fmaf_rn(rgb[3 * id + 2], w, C[2]);
            Dacc = __fmaf_rn(__fdiv_rn(1.f, depths[id]), w, Dacc);
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px, HW = (size_t)H * W;
        final_T[pix] = T; n_contrib[pix] = last;
        out_color[pix] = __fmaf_rn(T, bg[0], C[0]); out_color[HW + pix] = __fmaf_rn(T, bg[1], C[1]); out_color[2 * HW + pix] = __fmaf_rn(T, bg[2], C[2]);
        out_invdepth[pix] = Dacc;
    }
}

// ------------------------------------------------------------------------------------------------ backward kernels
__global__ void __launch_bounds__(RS_BLOCK)
rs_render_backward(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                   const float* __restrict__ bg, const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
                   const float* __restrict__ rgb, const float* __restrict__ depths, const float* __restrict__ final_T,
                   const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinvd,
                   float* __restrict__ dL_dmean2D /*[P,3]*/, float* __restrict__ dL_dconic /*[P,4]*/, float* __restrict__ dL_dopacity,
                   float* __restrict__ dL_dcolor /*[P,3]*/, float* __restrict__ dL_dinvdepth /*[P]*/) {
    __shared__ int s_id[RS_BLOCK];
    __shared__ float2 s_xy[RS_BLOCK];
    __shared__ float4 s_co[RS_BLOCK];
    __shared__ float s_rgb[3 * RS_BLOCK];
    __shared__ float s_inv[RS_BLOCK];
    const int gx = (W + RS_BLOCK_X - 1) / RS_BLOCK_X;
    const int px = blockIdx.x * RS_BLOCK_X + threadIdx.x, py = blockIdx.y * RS_BLOCK_Y + threadIdx.y;
    const int tid = threadIdx.y * RS_BLOCK_X + threadIdx.x;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t pix = (size_t)py * W + px, HW = (size_t)H * W;
    const uint2 rng = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = (int)((rng.y - rng.x + RS_BLOCK - 1) / RS_BLOCK);
    int todo = (int)(rng.y - rng.x);
    bool done = !inside;
    const float Tfin = inside ? final_T[pix] : 0.f;
    float T = Tfin;
    uint32_t contributor = (uint32_t)todo;
    const int last = inside ? (int)n_contrib[pix] : 0;
    float accum[3] = {0.f, 0.f, 0.f}, accum_inv = 0.f, dpix[3] = {0.f, 0.f, 0.f}, dinv = 0.f;
    if (inside) { dpix[0] = dL_dpix[pix]; dpix[1] = dL_dpix[HW + pix]; dpix[2] = dL_dpix[2 * HW + pix]; if (dL_dinvd) dinv = dL_dinvd[pix]; }
    float last_alpha = 0.f, last_col[3] = {0.f, 0.f, 0.f}, last_inv = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const float bg_dot = bg[0] * dpix[0] + bg[1] * dpix[1] + bg[2] * dpix[2];
    for (int r = 0; r < rounds; r++, todo -= RS_BLOCK) {
        __syncthreads();
        const int k = r * RS_BLOCK + tid;
        if (rng.x + k < rng.y) {
            const int id = (int)point_list[rng.y - k - 1];
            s_id[tid] = id; s_xy[tid] = means2D[id]; s_co[tid] = conic_opacity[id];
            s_rgb[3 * tid] = rgb[3 * id]; s_rgb[3 * tid + 1] = rgb[3 * id + 1]; s_rgb[3 * tid + 2] = rgb[3 * id + 2];
            s_inv[tid] = __fdiv_rn(1.f, depths[id]);
        }
        __syncthreads();
        for (int j = 0; !done && j < min(RS_BLOCK, todo); j++) {
            contributor--;
            if ((int)contributor >= last) continue;
            const float2 xy = s_xy[j]; const float4 co = s_co[j];
            const float dx = __fsub_rn(xy.x, pxf), dy = __fsub_rn(xy.y, pyf);
            const float power = rs_power(co.x, co.y, co.z, dx, dy);
            if (power > 0.f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, __fmul_rn(co.w, G));
            if (alpha < 1.0f / 255.0f) continue;
            T = T / (1.f - alpha);
            const float w = alpha * T;
            const int id = s_id[j];
            float dL_dalpha = 0.f;
            for (int c = 0; c < 3; c++) {
                const float col = s_rgb[3 * j + c];
                accum[c] = last_alpha * last_col[c] + (1.f - last_alpha) * accum[c];
                last_col[c] = col;
                dL_dalpha += (col - accum[c]) * dpix[c];
                atomicAdd(&dL_dcolor[3 * id + c], w * dpix[c]);
            }
            accum_inv = last_alpha * last_inv + (1.f - last_alpha) * accum_inv;
            last_inv = s_inv[j];
            dL_dalpha += (s_inv[j] - accum_inv) * dinv;
            atomicAdd(&dL_dinvdepth[id], w * dinv);
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-Tfin / (1.f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y, dG_ddely = -gdy * co.z - gdx * co.y;
            atomicAdd(&dL_dmean2D[3 * id], dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(&dL_dmean2D[3 * id + 1], dL_dG * dG_ddely * ddely_dy);
            atomicAdd(&dL_dconic[4 * id], -0.5f * gdx * dx * dL_dG);
            atomicAdd(&dL_dconic[4 * id + 1], -0.5f * gdx * dy * dL_dG);
            atomicAdd(&dL_dconic[4 * id + 3], -0.5f * gdy * dy * dL_dG);
            atomicAdd(&dL_dopacity[id], G * dL_dalpha);
        }
    }
}

// cov2D backward + the rest of preprocess backward as the stock's two kernels (same per-Gaussian maths as the product
// headers; stock splits it so that dL/dcov3D and dL/dmean3D make a round trip through global memory).
__global__ void __launch_bounds__(256)
rs_preprocess_backward(int P, int D, int M, const float* __restrict__ means, const int* __restrict__ radii,
                       const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
                       const float* __restrict__ rots, const float* __restrict__ cov_pre, float mod, const float* __restrict__ cov3D,
                       const float* __restrict__ opac, const float* __restrict__ view, const float* __restrict__ proj,
                       const float* __restrict__ campos, float tanfovx, float tanfovy, float focal_x, float focal_y, int antialiasing,
                       const float* __restrict__ dL_dmean2D, const float* __restrict__ dL_dconic, const float* __restrict__ dL_dinvdepth,
                       float* __restrict__ dL_dopacity, const float* __restrict__ dL_dcolor, float* __restrict__ dL_dmeans3D,
                       float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh, float* __restrict__ dL_dscale, float* __restrict__ dL_drot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    const float mean[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    float sc[3], rt[4];
    const float* scp = nullptr; const float* rtp = nullptr;
    if (!cov_pre) {
        sc[0] = scales[3 * i]; sc[1] = scales[3 * i + 1]; sc[2] = scales[3 * i + 2];
        rt[0] = rots[4 * i]; rt[1] = rots[4 * i + 1]; rt[2] = rots[4 * i + 2]; rt[3] = rots[4 * i + 3];
        scp = sc; rtp = rt;
    }
    float cov6[6];
    for (int k = 0; k < 6; k++) cov6[k] = cov3D[6 * (size_t)i + k];
    GmsPreGradIn gi;
    gi.dmean2D[0] = dL_dmean2D[3 * i]; gi.dmean2D[1] = dL_dmean2D[3 * i + 1];
    gi.dconic[0] = dL_dconic[4 * i]; gi.dconic[1] = dL_dconic[4 * i + 1]; gi.dconic[2] = dL_dconic[4 * i + 3];
    gi.dopac = dL_dopacity[i];
    gi.dcolor[0] = dL_dcolor[3 * i]; gi.dcolor[1] = dL_dcolor[3 * i + 1]; gi.dcolor[2] = dL_dcolor[3 * i + 2];
    gi.dinvdepth = dL_dinvdepth[i];
    GmsPreGradOut go;
    gms_preprocess_backward_geom(mean, scp, rtp, cov6, opac[i], view, proj, tanfovx, tanfovy, focal_x, focal_y, mod, antialiasing, gi, go);
    if (shs && dL_dsh) {
        float sh[48], dsh[48];
        const float* row = shs + (size_t)i * M * 3;
        const int nf = 3 * (D + 1) * (D + 1);
        for (int k = 0; k < nf; k++) sh[k] = row[k];
        const uint8_t cl[3] = {clamped[3 * i], clamped[3 * i + 1], clamped[3 * i + 2]};
        gms_sh_backward(D, M < 16 ? M : 16, mean, campos, sh, gi.dcolor, cl, dsh, go.dmean3D);
        float* out = dL_dsh + (size_t)i * M * 3;
        for (int k = 0; k < nf; k++) out[k] = dsh[k];       // the rest stays at the zero fill
    }
    dL_dmeans3D[3 * i] = go.dmean3D[0]; dL_dmeans3D[3 * i + 1] = go.dmean3D[1]; dL_dmeans3D[3 * i + 2] = go.dmean3D[2];
    dL_dopacity[i] = go.dopacity;
    if (dL_dcov3D) for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)i + k] = go.dcov6[k];
    if (dL_dscale) { dL_dscale[3 * i] = go.dscale[0]; dL_dscale[3 * i + 1] = go.dscale[1]; dL_dscale[3 * i + 2] = go.dscale[2]; }
    if (dL_drot) { dL_drot[4 * i] = go.drot[0]; dL_drot[4 * i + 1] = go.drot[1]; dL_drot[4 * i + 2] = go.drot[2]; dL_drot[4 * i + 3] = go.drot[3]; }
}

// ------------------------------------------------------------------------------------------------ C entry points
extern "C" {

const char* refstyle_last_error(void) { return rs_err; }

// per-Gaussian backward scratch (dL_dconic [P,4], dL_dcolor [P,3], dL_dinvdepth [P], dL_dcov3D [P,6]) -- torch::zeros in the stock wrapper
size_t refstyle_backward_scratch_bytes(int32_t P) { const size_t n = P > 0 ? P : 1; return rs_up(4 * n * 4) + rs_up(3 * n * 4) + rs_up(n * 4) + rs_up(6 * n * 4) + 256; }

int refstyle_rasterize_forward(const gms_raster_settings* s, const gms_raster_inputs* in, const gms_raster_outputs* out,
                               gms_alloc_fn alloc, void* user, gms_raster_saved* saved, void* cuda_stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (!s || !in || !out || !alloc || !saved) return rs_fail(GMS_E_ARG, "null argument");
    const int P = in->P, W = s->image_width, H = s->image_height;
    const int gx = (W + RS_BLOCK_X - 1) / RS_BLOCK_X, gy = (H + RS_BLOCK_Y - 1) / RS_BLOCK_Y, T = gx * gy;
    saved->geom = saved->binning = saved->image = nullptr; saved->num_rendered = 0; saved->num_visible = -1;
    if (P <= 0) return rs_fail(GMS_E_ARG, "refstyle: P must be > 0");
    const float focal_x = (float)W / (2.0f * s->tanfovx), focal_y = (float)H / (2.0f * s->tanfovy);
    void* graw = alloc(user, GMS_BUF_GEOM, rs_geom(nullptr, P).total + 256);
    void* iraw = alloc(user, GMS_BUF_IMAGE, rs_img(nullptr, W, H).total + 256);
    if (!graw || !iraw) return rs_fail(GMS_E_ALLOC, "scratch allocation failed");
    saved->geom = graw; saved->image = iraw;
    RsGeom G = rs_geom(reinterpret_cast<void*>(rs_up((size_t)graw)), P);
    RsImg I = rs_img(reinterpret_cast<void*>(rs_up((size_t)iraw)), W, H);
    rs_preprocess<<<(P + 255) / 256, 256, 0, st>>>(P, s->sh_degree, in->M, in->means3D, in->scales, s->scale_modifier, in->rotations,
                                                  in->opacities, in->shs, in->colors_precomp, in->cov3D_precomp, s->viewmatrix,
                                                  s->projmatrix, s->campos, W, H, s->tanfovx, s->tanfovy, focal_x, focal_y,
                                                  s->antialiasing, gx, gy, out->radii, G);
    RS_LAUNCHED("rs_preprocess");
    size_t sb = G.scan_bytes;
    RS_CUDA(cub::DeviceScan::InclusiveSum(G.scan_space, sb, G.tiles_touched, G.point_offsets, P, st));
    uint32_t n32 = 0;
    RS_CUDA(cudaMemcpyAsync(&n32, G.point_offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RS_CUDA(cudaStreamSynchronize(st));          // stock: blocking cudaMemcpy of num_rendered
    const int64_t N = n32;
    saved->num_rendered = N;
    void* braw = alloc(user, GMS_BUF_BINNING, rs_bin(nullptr, N).total + 256);
    if (!braw) return rs_fail(GMS_E_ALLOC, "binning allocation failed");
    saved->binning = braw;
    RsBin B = rs_bin(reinterpret_cast<void*>(rs_up((size_t)braw)), N);
    RS_CUDA(cudaMemsetAsync(I.ranges, 0, sizeof(uint2) * (size_t)T, st));
    if (N > 0) {
        rs_duplicate_with_keys<<<(P + 255) / 256, 256, 0, st>>>(P, gx, gy, G.means2D, G.depths, G.point_offsets, out->radii, B.keys_unsorted, B.vals_unsorted);
        RS_LAUNCHED("rs_duplicate_with_keys");
        size_t tb = B.sort_bytes;
        RS_CUDA(cub::DeviceRadixSort::SortPairs(B.sort_space, tb, B.keys_unsorted, B.keys, B.vals_unsorted, B.vals, (int)N, 0,
                                                32 + gms_tile_bits((uint32_t)T), st));
        rs_identify_tile_ranges<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, B.keys, I.ranges);
        RS_LAUNCHED("rs_identify_tile_ranges");
    }
    rs_render_forward<<<dim3(gx, gy), dim3(RS_BLOCK_X, RS_BLOCK_Y), 0, st>>>(I.ranges, B.vals, W, H, G.means2D, G.rgb, G.conic_opacity, G.depths,
                                                                              s->bg, I.accum_alpha, I.n_contrib, out->out_color, out->out_invdepth);
    RS_LAUNCHED("rs_render_forward");
    return GMS_OK;
}

int refstyle_rasterize_backward(const gms_raster_settings* s, const gms_raster_inputs* in, const int32_t* radii,
                                const gms_raster_saved* saved, const float* dL_dout_color, const float* dL_dout_invdepth,
                                const gms_raster_grads* gr, void* bwd_scratch, void* cuda_stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (!s || !in || !saved || !gr || !dL_dout_color || !bwd_scratch) return rs_fail(GMS_E_ARG, "null argument");
    const int P = in->P, W = s->image_width, H = s->image_height;
    const int gx = (W + RS_BLOCK_X - 1) / RS_BLOCK_X, gy = (H + RS_BLOCK_Y - 1) / RS_BLOCK_Y;
    const float focal_x = (float)W / (2.0f * s->tanfovx), focal_y = (float)H / (2.0f * s->tanfovy);
    RsGeom G = rs_geom(reinterpret_cast<void*>(rs_up((size_t)saved->geom)), P);
    RsImg I = rs_img(reinterpret_cast<void*>(rs_up((size_t)saved->image)), W, H);
    RsBin B = rs_bin(reinterpret_cast<void*>(rs_up((size_t)saved->binning)), saved->num_rendered);
    char* p = reinterpret_cast<char*>(rs_up((size_t)bwd_scratch));
    const size_t n = (size_t)P;
    float* dconic = rs_take<float>(p, 4 * n); float* dcolor = rs_take<float>(p, 3 * n); float* dinvd = rs_take<float>(p, n);
    float* dcov = rs_take<float>(p, 6 * n);
    // stock wrapper: every gradient tensor starts as torch::zeros
    RS_CUDA(cudaMemsetAsync(dconic, 0, 4 * n * 4, st)); RS_CUDA(cudaMemsetAsync(dcolor, 0, 3 * n * 4, st));
    RS_CUDA(cudaMemsetAsync(dinvd, 0, n * 4, st)); RS_CUDA(cudaMemsetAsync(dcov, 0, 6 * n * 4, st));
    RS_CUDA(cudaMemsetAsync(gr->dL_dmeans3D, 0, 3 * n * 4, st)); RS_CUDA(cudaMemsetAsync(gr->dL_dmeans2D, 0, 3 * n * 4, st));
    RS_CUDA(cudaMemsetAsync(gr->dL_dopacities, 0, n * 4, st));
    if (gr->dL_dshs) RS_CUDA(cudaMemsetAsync(gr->dL_dshs, 0, n * in->M * 3 * 4, st));
    if (gr->dL_dscales) RS_CUDA(cudaMemsetAsync(gr->dL_dscales, 0, 3 * n * 4, st));
    if (gr->dL_drotations) RS_CUDA(cudaMemsetAsync(gr->dL_drotations, 0, 4 * n * 4, st));
    float* dcol_out = in->colors_precomp ? gr->dL_dcolors_precomp : dcolor;
    if (in->colors_precomp) RS_CUDA(cudaMemsetAsync(dcol_out, 0, 3 * n * 4, st));
    float* dcov_out = in->cov3D_precomp ? gr->dL_dcov3D_precomp : dcov;
    if (in->cov3D_precomp) RS_CUDA(cudaMemsetAsync(dcov_out, 0, 6 * n * 4, st));
    rs_render_backward<<<dim3(gx, gy), dim3(RS_BLOCK_X, RS_BLOCK_Y), 0, st>>>(I.ranges, B.vals, W, H, s->bg, G.means2D, G.conic_opacity, G.rgb,
                                                                               G.depths, I.accum_alpha, I.n_contrib, dL_dout_color, dL_dout_invdepth,
                                                                               gr->dL_dmeans2D, dconic, gr->dL_dopacities, dcol_out, dinvd);
    RS_LAUNCHED("rs_render_backward");
    rs_preprocess_backward<<<(P + 255) / 256, 256, 0, st>>>(P, s->sh_degree, in->M, in->means3D, radii, in->shs, G.clamped, in->scales,
                                                           in->rotations, in->cov3D_precomp, s->scale_modifier, G.cov3D, in->opacities,
                                                           s->viewmatrix, s->projmatrix, s->campos, s->tanfovx, s->tanfovy, focal_x, focal_y,
                                                           s->antialiasing, gr->dL_dmeans2D, dconic, dinvd, gr->dL_dopacities, dcol_out,
                                                           gr->dL_dmeans3D, dcov_out, gr->dL_dshs, gr->dL_dscales, gr->dL_drotations);
    RS_LAUNCHED("rs_preprocess_backward");
    return GMS_OK;
}

}  // extern "C"
