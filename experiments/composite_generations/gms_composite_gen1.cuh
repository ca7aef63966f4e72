// gms_composite.cuh -- per-tile front-to-back alpha compositing, forward and backward, for sm_100a.
//
// Replaces [upstream forward.cu: renderCUDA] and [upstream backward.cu: renderCUDA] of
// graphdeco-inria/diff-gaussian-rasterization (reference call sites: renderer/gaussian_renderer/__init__.py:94-102,
// train.py:108).  Semantics: SURVEY.md Appendix A.2 / A.3, checked by the CPU oracle under oracle/.
//
// B200 design (differs from the stock one-thread-per-pixel / per-pixel-atomics kernels):
//  * a 16x16 tile is owned by a 128-thread CTA; each warp owns an 8x8 pixel QUAD, each lane two vertically
//    adjacent pixels (the x-dependent half of the quadratic form is shared between them, ILP = 2);
//  * the tile's depth-sorted splat list is streamed in batches of 128 through double-buffered shared memory;
//    global loads of batch b+1 (48-byte packed records, three 128-bit loads) are in flight while batch b is
//    composited; one __syncthreads per batch;
//  * while staging, every splat's conservative screen-space extent (where alpha can reach 1/255) is tested
//    against the four quads; warp ballots turn that into a per-quad bitmask, and a warp only visits the
//    splats whose bit is set (thin mesh Gaussians touch a tile's rect but few of its pixels);
//  * backward: per-lane partial sums of 10 moment accumulators per splat are folded across the warp with a
//    12-shuffle transpose-reduce, the four quads' partials meet in shared memory, and ONE thread per
//    (tile, splat) issues three vector reductions (red.global.add.v4.f32) -- ~1/256 of the stock atomics.
#pragma once
#include "gms_common.cuh"

#define GMS_CB 128                  // CTA size == splats per staged batch
#define GMS_LOG2E 1.4426950408889634f

__device__ __forceinline__ float gms_exp(float x) { return __expf(x); }

// canonical quadratic form (see oracle quad_power): shared x-part
__device__ __forceinline__ float gms_power(float m2, float m4, float conz, float dy) {
    const float m3 = __fmul_rn(conz, dy);
    const float t = __fmaf_rn(m3, dy, m2);
    const float h = __fmul_rn(-0.5f, t);
    return __fmaf_rn(-m4, dy, h);
}

struct GmsTileGeom {
    int tx0, ty0, px, py0;
    bool in0, in1;
};

__device__ __forceinline__ GmsTileGeom gms_tile_geom(int tile, int gx, int W, int H, int warp, int lane) {
    GmsTileGeom g;
    g.tx0 = (tile % gx) * GMS_TILE;
    g.ty0 = (tile / gx) * GMS_TILE;
    g.px = g.tx0 + (warp & 1) * 8 + (lane & 7);
    g.py0 = g.ty0 + (warp >> 1) * 8 + (lane >> 3) * 2;
    g.in0 = g.px < W && g.py0 < H;
    g.in1 = g.px < W && (g.py0 + 1) < H;
    return g;
}

// Can this splat reach alpha >= 1/255 anywhere in the pixel rectangle [rx0, rx0+7] x [ry0, ry0+7]?
// i.e. min over the rectangle of 0.5*(cx dx^2 + cz dy^2) + cy dx dy  <=  tau   (tau carries the safety margin).
__device__ __forceinline__ bool gms_reaches_quad(float x, float y, float cx, float cy, float cz, float tau, float rx0, float ry0) {
    if (!(tau > 0.f)) return false;                 // opacity < 1/255: never blends
    if (!(cx > 0.f) || !(cz > 0.f)) return true;    // degenerate conic: be conservative
    const float rx1 = rx0 + 7.0f, ry1 = ry0 + 7.0f;
    if (x >= rx0 && x <= rx1 && y >= ry0 && y <= ry1) return true;
    float best = 3.0e38f;
    const float icx = __fdividef(1.f, cx), icz = __fdividef(1.f, cz);
#pragma unroll
    for (int e = 0; e < 2; e++) {
        {   // horizontal edge
            const float dy = y - (e ? ry1 : ry0);
            const float t = fminf(fmaxf(x + cy * dy * icx, rx0), rx1);
            const float dx = x - t;
            best = fminf(best, 0.5f * (cx * dx * dx + cz * dy * dy) + cy * dx * dy);
        }
        {   // vertical edge
            const float dx = x - (e ? rx1 : rx0);
            const float t = fminf(fmaxf(y + cy * dx * icz, ry0), ry1);
            const float dy = y - t;
            best = fminf(best, 0.5f * (cx * dx * dx + cz * dy * dy) + cy * dx * dy);
        }
    }
    return best <= tau;
}

// which of the tile's four 8x8 quads can a splat reach?
__device__ __forceinline__ uint32_t gms_quad_mask(float x, float y, float cx, float cy, float cz, float tau, int tx0, int ty0) {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
        m |= (gms_reaches_quad(x, y, cx, cy, cz, tau, (float)(tx0 + (q & 1) * 8), (float)(ty0 + (q >> 1) * 8)) ? 1u : 0u) << q;
    return m;
}

// ------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(GMS_CB)
k_composite_fwd(const int2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                float* __restrict__ out_color, float* __restrict__ final_T, int* __restrict__ n_contrib,
                float* __restrict__ out_invdepth, int* __restrict__ tile_last, int use_masks) {
    __shared__ float4 s_a[2][GMS_CB];   // x, y, conx, cony
    __shared__ float4 s_b[2][GMS_CB];   // conz, opacity, r, g
    __shared__ float2 s_c[2][GMS_CB];   // b, 1/depth
    __shared__ uint32_t s_qm[2][4][4];  // [buffer][quad][staging warp]
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const GmsTileGeom g = gms_tile_geom(tile, gx, W, H, warp, lane);
    const int2 rng = ranges[tile];
    const int n = rng.y - rng.x;
    const int nb = (n + GMS_CB - 1) / GMS_CB;
    const float pxf = (float)g.px, pyf0 = (float)g.py0, pyf1 = (float)(g.py0 + 1);

    float T0 = 1.f, T1 = 1.f, D0 = 0.f, D1 = 0.f;
    float C0[3] = {0.f, 0.f, 0.f}, C1[3] = {0.f, 0.f, 0.f};
    int last0 = 0, last1 = 0;
    bool done0 = !g.in0, done1 = !g.in1;
    if (tid == 0) s_last = 0;

    int id_cur = (0 < n && tid < n) ? (int)point_list[rng.x + tid] : -1;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra;
    if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
    int id_nx = (GMS_CB + tid < n) ? (int)point_list[rng.x + GMS_CB + tid] : -1;

    for (int b = 0; b < nb; b++) {
        const int buf = b & 1;
        uint32_t qmask = 0;
        if (id_cur >= 0) {
            s_a[buf][tid] = ra;
            s_b[buf][tid] = rb;
            s_c[buf][tid] = make_float2(rc.x, rc.y);
            qmask = use_masks ? gms_quad_mask(ra.x, ra.y, ra.z, ra.w, rb.x, rc.z, g.tx0, g.ty0) : 0xFu;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (qmask >> q) & 1u);
            if (lane == 0) s_qm[buf][q][warp] = bal;
        }
        const int all_done = __syncthreads_and(done0 && done1);
        if (all_done) break;
        // software prefetch: records of batch b+1, ids of batch b+2
        id_cur = id_nx;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
        {
            const int k = (b + 2) * GMS_CB + tid;
            id_nx = (k < n) ? (int)point_list[rng.x + k] : -1;
        }
        if (__all_sync(0xffffffffu, done0 && done1)) continue;
#pragma unroll 1
        for (int sw = 0; sw < 4; sw++) {
            uint32_t m = s_qm[buf][warp][sw];
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int jj = sw * 32 + j;
                const float4 A = s_a[buf][jj];
                const float4 B = s_b[buf][jj];
                const float2 Cc = s_c[buf][jj];
                const int pos = b * GMS_CB + jj + 1;
                const float dx = __fsub_rn(A.x, pxf);
                const float m1 = __fmul_rn(A.z, dx);
                const float m2 = __fmul_rn(m1, dx);
                const float m4 = __fmul_rn(A.w, dx);
                {
                    const float dy = __fsub_rn(A.y, pyf0);
                    const float power = gms_power(m2, m4, B.x, dy);
                    const float alpha = fminf(GMS_ALPHA_MAX, __fmul_rn(B.y, gms_exp(power)));
                    bool ok = !done0 && power <= 0.0f && alpha >= GMS_ALPHA_MIN;
                    const float test_T = __fmul_rn(T0, __fsub_rn(1.f, alpha));
                    if (ok && test_T < GMS_T_STOP) { done0 = true; ok = false; }
                    if (ok) {
                        const float w = __fmul_rn(alpha, T0);
                        C0[0] = fmaf(B.z, w, C0[0]); C0[1] = fmaf(B.w, w, C0[1]); C0[2] = fmaf(Cc.x, w, C0[2]);
                        D0 = fmaf(Cc.y, w, D0);
                        T0 = test_T; last0 = pos;
                    }
                }
                {
                    const float dy = __fsub_rn(A.y, pyf1);
                    const float power = gms_power(m2, m4, B.x, dy);
                    const float alpha = fminf(GMS_ALPHA_MAX, __fmul_rn(B.y, gms_exp(power)));
                    bool ok = !done1 && power <= 0.0f && alpha >= GMS_ALPHA_MIN;
                    const float test_T = __fmul_rn(T1, __fsub_rn(1.f, alpha));
                    if (ok && test_T < GMS_T_STOP) { done1 = true; ok = false; }
                    if (ok) {
                        const float w = __fmul_rn(alpha, T1);
                        C1[0] = fmaf(B.z, w, C1[0]); C1[1] = fmaf(B.w, w, C1[1]); C1[2] = fmaf(Cc.x, w, C1[2]);
                        D1 = fmaf(Cc.y, w, D1);
                        T1 = test_T; last1 = pos;
                    }
                }
            }
        }
    }
    const size_t HW = (size_t)H * W;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    if (g.in0) {
        const size_t pix = (size_t)g.py0 * W + g.px;
        final_T[pix] = T0; n_contrib[pix] = last0;
        out_color[pix] = fmaf(T0, bg0, C0[0]); out_color[HW + pix] = fmaf(T0, bg1, C0[1]);
        out_color[2 * HW + pix] = fmaf(T0, bg2, C0[2]);
        out_invdepth[pix] = D0;
    }
    if (g.in1) {
        const size_t pix = (size_t)(g.py0 + 1) * W + g.px;
        final_T[pix] = T1; n_contrib[pix] = last1;
        out_color[pix] = fmaf(T1, bg0, C1[0]); out_color[HW + pix] = fmaf(T1, bg1, C1[1]);
        out_color[2 * HW + pix] = fmaf(T1, bg2, C1[2]);
        out_invdepth[pix] = D1;
    }
    const int wl = __reduce_max_sync(0xffffffffu, max(last0, last1));
    if (lane == 0 && wl > 0) atomicMax(&s_last, wl);
    __syncthreads();
    if (tid == 0) tile_last[tile] = s_last;
}

// ------------------------------------------------------------------------------------------ backward
// transpose-reduce of 10 per-lane values over the warp in 12 shuffles; lane (even, valid) ends up holding
// the warp-wide sum of value `idx`.
__device__ __forceinline__ void gms_fold10(const float (&v)[10], int lane, float& out, int& idx, bool& valid) {
    const unsigned F = 0xffffffffu;
    bool hi = (lane & 16) != 0;
    float w[6];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float send = hi ? v[i] : v[i + 5];
        const float keep = hi ? v[i + 5] : v[i];
        w[i] = keep + __shfl_xor_sync(F, send, 16);
    }
    w[5] = 0.f;
    hi = (lane & 8) != 0;
    float x[4];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float send = hi ? w[i] : w[i + 3];
        const float keep = hi ? w[i + 3] : w[i];
        x[i] = keep + __shfl_xor_sync(F, send, 8);
    }
    x[3] = 0.f;
    hi = (lane & 4) != 0;
    float y[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = hi ? x[i] : x[i + 2];
        const float keep = hi ? x[i + 2] : x[i];
        y[i] = keep + __shfl_xor_sync(F, send, 4);
    }
    hi = (lane & 2) != 0;
    float z;
    {
        const float send = hi ? y[0] : y[1];
        const float keep = hi ? y[1] : y[0];
        z = keep + __shfl_xor_sync(F, send, 2);
    }
    z += __shfl_xor_sync(F, z, 1);
    const int i1 = (lane & 8) ? 3 : 0, i2 = (lane & 4) ? 2 : 0, i3 = (lane & 2) ? 1 : 0;
    const int xi = i2 + i3;
    valid = (xi <= 2) && (i1 + xi <= 4) && ((lane & 1) == 0);
    idx = ((lane & 16) ? 5 : 0) + i1 + xi;
    out = z;
}

// Loop-hoisted form of gms_fold10 for the streaming kernels: the (lane-constant) destination slot is computed once
// before the splat loop -- gms_fold_slot(): idx of the sum this lane ends up with, or -1 -- and made opaque so that the
// compiler keeps it in a register instead of re-deriving it from %tid.x in every iteration (9 instructions + S2R).
__device__ __forceinline__ int gms_fold_slot(int lane) {
    const int i1 = (lane & 8) ? 3 : 0, i2 = (lane & 4) ? 2 : 0, i3 = (lane & 2) ? 1 : 0;
    const int xi = i2 + i3;
    const bool valid = (xi <= 2) && (i1 + xi <= 4) && ((lane & 1) == 0);
    int slot = valid ? ((lane & 16) ? 5 : 0) + i1 + xi : -1;
    asm volatile("" : "+r"(slot));
    return slot;
}
__device__ __forceinline__ float gms_fold10_sum(const float (&v)[10], int lane) {
    float out; int idx; bool valid;
    gms_fold10(v, lane, out, idx, valid);      // idx / valid are dead here and drop out
    return out;
}

struct GmsBwdPix {
    float T, Tfin, accum[3], lastc[3], accum_inv, last_inv, last_alpha, dpix[3], dinv, bg_dot;
    int last;
};

__device__ __forceinline__ void gms_bwd_pixel(GmsBwdPix& p, int pos, float dx, float m2, float m4, float dy,
                                              const float4& B, const float2& Cc, float (&v)[10], bool& any) {
    if (pos >= p.last) return;
    const float power = gms_power(m2, m4, B.x, dy);
    if (power > 0.0f) return;
    const float G = gms_exp(power);
    const float alpha = fminf(GMS_ALPHA_MAX, __fmul_rn(B.y, G));
    if (alpha < GMS_ALPHA_MIN) return;
    const float inv = __fdividef(1.f, 1.f - alpha);
    p.T = p.T * inv;
    const float w = alpha * p.T;
    float dLa = 0.f;
    const float la = p.last_alpha, ola = 1.f - la;
    const float col[3] = {B.z, B.w, Cc.x};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        p.accum[c] = la * p.lastc[c] + ola * p.accum[c];
        p.lastc[c] = col[c];
        dLa += (col[c] - p.accum[c]) * p.dpix[c];
    }
    p.accum_inv = la * p.last_inv + ola * p.accum_inv;
    p.last_inv = Cc.y;
    dLa += (Cc.y - p.accum_inv) * p.dinv;
    dLa *= p.T;
    p.last_alpha = alpha;
    dLa += (-p.Tfin * inv) * p.bg_dot;
    const float q = dLa * G;
    const float qx = q * dx, qy = q * dy;
    v[0] += qx; v[1] += qy; v[2] += qx * dx; v[3] += qx * dy; v[4] += qy * dy; v[5] += q;
    v[6] += w * p.dpix[0]; v[7] += w * p.dpix[1]; v[8] += w * p.dpix[2]; v[9] += w * p.dinv;
    any = true;
}

__global__ void __launch_bounds__(GMS_CB)
k_composite_bwd(const int2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                const float* __restrict__ final_T, const int* __restrict__ n_contrib,
                const int* __restrict__ tile_last, const float* __restrict__ dL_dpix,
                const float* __restrict__ dL_dinv, float4* __restrict__ dgeom, int use_masks) {
    __shared__ float4 s_a[GMS_CB];
    __shared__ float4 s_b[GMS_CB];
    __shared__ float2 s_c[GMS_CB];
    __shared__ int s_id[GMS_CB];
    __shared__ uint32_t s_qm[4][4];
    __shared__ uint32_t s_touch[4][4];
    __shared__ float s_part[4][GMS_CB][10];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int n_tile = tile_last[tile];
    if (n_tile <= 0) return;
    const GmsTileGeom g = gms_tile_geom(tile, gx, W, H, warp, lane);
    const int2 rng = ranges[tile];
    const float pxf = (float)g.px, pyf0 = (float)g.py0, pyf1 = (float)(g.py0 + 1);
    const size_t HW = (size_t)H * W;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const float halfW = 0.5f * (float)W, halfH = 0.5f * (float)H;

    GmsBwdPix p0, p1;
    {
        GmsBwdPix* pp[2] = {&p0, &p1};
        const bool in[2] = {g.in0, g.in1};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            GmsBwdPix& p = *pp[k];
            p.accum[0] = p.accum[1] = p.accum[2] = 0.f;
            p.lastc[0] = p.lastc[1] = p.lastc[2] = 0.f;
            p.accum_inv = p.last_inv = p.last_alpha = 0.f;
            if (in[k]) {
                const size_t pix = (size_t)(g.py0 + k) * W + g.px;
                p.Tfin = final_T[pix]; p.T = p.Tfin; p.last = n_contrib[pix];
                p.dpix[0] = dL_dpix[pix]; p.dpix[1] = dL_dpix[HW + pix]; p.dpix[2] = dL_dpix[2 * HW + pix];
                p.dinv = dL_dinv ? dL_dinv[pix] : 0.f;
            } else {
                p.Tfin = p.T = 1.f; p.last = 0; p.dpix[0] = p.dpix[1] = p.dpix[2] = 0.f; p.dinv = 0.f;
            }
            p.bg_dot = bg0 * p.dpix[0] + bg1 * p.dpix[1] + bg2 * p.dpix[2];
        }
    }
    const int wlast = __reduce_max_sync(0xffffffffu, max(p0.last, p1.last));

    const int nb = (n_tile + GMS_CB - 1) / GMS_CB;
    int id_cur;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra;
    {
        const int k = (nb - 1) * GMS_CB + tid;
        id_cur = (k < n_tile) ? (int)point_list[rng.x + k] : -1;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
    }
    int id_nx = (nb >= 2) ? (int)point_list[rng.x + (nb - 2) * GMS_CB + tid] : -1;   // full batch: always valid

    for (int b = nb - 1; b >= 0; b--) {
        // ---- stage batch b
        uint32_t qmask = 0;
        s_id[tid] = id_cur;
        if (id_cur >= 0) {
            s_a[tid] = ra; s_b[tid] = rb; s_c[tid] = make_float2(rc.x, rc.y);
            qmask = use_masks ? gms_quad_mask(ra.x, ra.y, ra.z, ra.w, rb.x, rc.z, g.tx0, g.ty0) : 0xFu;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (qmask >> q) & 1u);
            if (lane == 0) { s_qm[q][warp] = bal; }
        }
        if (tid < 16) s_touch[tid >> 2][tid & 3] = 0u;
        __syncthreads();
        // prefetch batch b-1 records and batch b-2 ids
        const float4 sa_keep = ra; (void)sa_keep;
        id_cur = id_nx;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
        id_nx = (b >= 2) ? (int)point_list[rng.x + (b - 2) * GMS_CB + tid] : -1;

        // ---- composite backward over the batch, back to front
        if (wlast > b * GMS_CB) {
#pragma unroll 1
            for (int sw = 3; sw >= 0; sw--) {
                uint32_t m = s_qm[warp][sw];
                uint32_t touched = 0;
                while (m) {
                    const int j = 31 - __clz(m);
                    m &= ~(1u << j);
                    const int jj = sw * 32 + j;
                    const int pos = b * GMS_CB + jj;
                    if (pos >= wlast) continue;
                    const float4 A = s_a[jj];
                    const float4 B = s_b[jj];
                    const float2 Cc = s_c[jj];
                    const float dx = __fsub_rn(A.x, pxf);
                    const float m1 = __fmul_rn(A.z, dx);
                    const float m2 = __fmul_rn(m1, dx);
                    const float m4 = __fmul_rn(A.w, dx);
                    float v[10];
#pragma unroll
                    for (int i = 0; i < 10; i++) v[i] = 0.f;
                    bool any = false;
                    gms_bwd_pixel(p0, pos, dx, m2, m4, __fsub_rn(A.y, pyf0), B, Cc, v, any);
                    gms_bwd_pixel(p1, pos, dx, m2, m4, __fsub_rn(A.y, pyf1), B, Cc, v, any);
                    if (!__any_sync(0xffffffffu, any)) continue;
                    float out; int idx; bool valid;
                    gms_fold10(v, lane, out, idx, valid);
                    if (valid) s_part[warp][jj][idx] = out;
                    touched |= 1u << j;
                }
                if (lane == 0) s_touch[warp][sw] = touched;
            }
        }
        __syncthreads();
        // ---- flush: one thread per (tile, splat)
        {
            const int id = s_id[tid];
            uint32_t tb = 0;
#pragma unroll
            for (int w = 0; w < 4; w++) tb |= ((s_touch[w][tid >> 5] >> (tid & 31)) & 1u) << w;
            if (id >= 0 && tb) {
                float s[10];
#pragma unroll
                for (int i = 0; i < 10; i++) s[i] = 0.f;
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    if ((tb >> w) & 1u) {
#pragma unroll
                        for (int i = 0; i < 10; i++) s[i] += s_part[w][tid][i];
                    }
                }
                const float4 A = s_a[tid];
                const float4 B = s_b[tid];
                const float conx = A.z, cony = A.w, conz = B.x, op = B.y;
                float4 g0, g1, g2;
                g0.x = (-conx * s[0] - cony * s[1]) * op * halfW;   // dL/dmean2D.x (NDC-scaled)
                g0.y = (-conz * s[1] - cony * s[0]) * op * halfH;   // dL/dmean2D.y
                g0.z = -0.5f * op * s[2];                          // dL/dconic.x
                g0.w = -0.5f * op * s[3];                          // dL/dconic.y (stock half convention)
                g1.x = -0.5f * op * s[4];                          // dL/dconic.z
                g1.y = s[5];                                       // dL/d(conic_opacity.w)
                g1.z = s[6]; g1.w = s[7];                          // dL/drgb
                g2.x = s[8]; g2.y = s[9]; g2.z = 0.f; g2.w = 0.f;  // dL/drgb.b, dL/dinvdepth
                atomicAdd(&dgeom[3 * id], g0);
                atomicAdd(&dgeom[3 * id + 1], g1);
                atomicAdd(reinterpret_cast<float2*>(&dgeom[3 * id + 2]), make_float2(g2.x, g2.y));
            }
        }
        // the next iteration's staging writes s_a/s_b/s_id: all flush reads must be done
        __syncthreads();
    }
}
