// gms_composite2.cuh -- composite forward/backward, generation 2: WARP-INDEPENDENT streaming.
//
// Same semantics and the same canonical fp32 sequences as gms_composite.cuh (generation 1, kept for A/B runs).
// What changed, and why (profiles/r1b_*: generation 1 spent ~1/3 of its issue slots stalled on the per-batch
// __syncthreads because the four 8x8 quads of a tile have very different amounts of work):
//  * each warp streams the tile's depth-sorted list ON ITS OWN, 32 splats per round (lane j stages splat j into a
//    warp-private shared-memory slab), decides with an exact ellipse-vs-rectangle test which of them can reach
//    ITS quad at alpha >= 1/255, and composites those; there is no block-level barrier anywhere in the loop and a
//    warp whose 64 pixels have saturated simply exits.  The four warps of a tile read the same records at about
//    the same time, so the replicated loads hit in L1.
//  * tiles are launched longest-list-first (a 32-bucket counting sort of the tile ranges) to cut the tail;
//  * exp() is a bare ex2.approx.ftz (the log2(e) multiply stays in fp32 -- bit-identical to __expf outside the
//    denormal range, which alpha >= 1/255 never reaches);
//  * backward: per (quad, splat) the 10 moment sums are folded across the warp (12 shuffles), parked in a warp-private
//    slab, and after each round lane j turns splat j's sums into the 10 gradients and issues three vector reductions.
#pragma once
#include "gms_composite.cuh"

#define GMS_WB 32   // splats staged per warp round

__device__ __forceinline__ float gms_exp_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__fmul_rn(x, GMS_LOG2E)));
    return y;
}

struct GmsWarpSlab {          // one per warp, two for the forward's double buffering
    float4 a[GMS_WB];         // x, y, conx, cony
    float4 b[GMS_WB];         // conz, opacity, r, g
    float2 c[GMS_WB];         // b, 1/depth
};

// ------------------------------------------------------------------------------------------- tile order
// Longest list first: counting sort of the T tiles into 32 buckets by bit-length of their list (one CTA).
__global__ void __launch_bounds__(1024) k_tile_order(int T, const int2* __restrict__ ranges, int* __restrict__ order) {
    __shared__ int s_cnt[33];
    __shared__ int s_off[33];
    if (threadIdx.x < 33) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int n = ranges[t].y - ranges[t].x;
        atomicAdd(&s_cnt[n > 0 ? 32 - __clz(n) : 0], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 32; b >= 0; b--) { s_off[b] = run; run += s_cnt[b]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int n = ranges[t].y - ranges[t].x;
        const int pos = atomicAdd(&s_off[n > 0 ? 32 - __clz(n) : 0], 1);
        order[pos] = t;
    }
}

// ------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(GMS_CB)
k_composite_fwd2(const int2* __restrict__ ranges, const int* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                 float* __restrict__ out_color, float* __restrict__ final_T, int* __restrict__ n_contrib,
                 float* __restrict__ out_invdepth) {
    __shared__ GmsWarpSlab s_slab[4][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;
    const GmsTileGeom g = gms_tile_geom(tile, gx, W, H, warp, lane);
    const int2 rng = ranges[tile];
    const int n = rng.y - rng.x;
    const float pxf = (float)g.px, pyf0 = (float)g.py0, pyf1 = (float)(g.py0 + 1);
    const float qx0 = (float)(g.tx0 + (warp & 1) * 8), qy0 = (float)(g.ty0 + (warp >> 1) * 8);

    float T0 = 1.f, T1 = 1.f, D0 = 0.f, D1 = 0.f;
    float C0[3] = {0.f, 0.f, 0.f}, C1[3] = {0.f, 0.f, 0.f};
    int last0 = 0, last1 = 0;
    bool live0 = g.in0, live1 = g.in1;

    int id_cur = (lane < n) ? (int)point_list[rng.x + lane] : -1;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra;
    if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
    int id_nx = (GMS_WB + lane < n) ? (int)point_list[rng.x + GMS_WB + lane] : -1;

    for (int base = 0; base < n; base += GMS_WB) {
        if (!__any_sync(0xffffffffu, live0 || live1)) break;
        GmsWarpSlab& S = s_slab[warp][(base >> 5) & 1];
        bool hit = false;
        if (id_cur >= 0) {
            S.a[lane] = ra; S.b[lane] = rb; S.c[lane] = make_float2(rc.x, rc.y);
            hit = gms_reaches_quad(ra.x, ra.y, ra.z, ra.w, rb.x, rc.z, qx0, qy0);
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        // prefetch the next round while this one is composited
        id_cur = id_nx;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
        {
            const int k = base + 2 * GMS_WB + lane;
            id_nx = (k < n) ? (int)point_list[rng.x + k] : -1;
        }
        __syncwarp();
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float4 A = S.a[j];
            const float4 B = S.b[j];
            const float2 Cc = S.c[j];
            const int pos = base + j + 1;
            const float dx = __fsub_rn(A.x, pxf);
            const float m1 = __fmul_rn(A.z, dx);
            const float m2 = __fmul_rn(m1, dx);
            const float m4 = __fmul_rn(A.w, dx);
            {
                const float dy = __fsub_rn(A.y, pyf0);
                const float power = gms_power(m2, m4, B.x, dy);
                const float alpha = fminf(GMS_ALPHA_MAX, __fmul_rn(B.y, gms_exp_fast(power)));
                bool ok = live0 && power <= 0.0f && alpha >= GMS_ALPHA_MIN;
                const float test_T = __fmul_rn(T0, __fsub_rn(1.f, alpha));
                if (ok && test_T < GMS_T_STOP) { live0 = false; ok = false; }
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
                const float alpha = fminf(GMS_ALPHA_MAX, __fmul_rn(B.y, gms_exp_fast(power)));
                bool ok = live1 && power <= 0.0f && alpha >= GMS_ALPHA_MIN;
                const float test_T = __fmul_rn(T1, __fsub_rn(1.f, alpha));
                if (ok && test_T < GMS_T_STOP) { live1 = false; ok = false; }
                if (ok) {
                    const float w = __fmul_rn(alpha, T1);
                    C1[0] = fmaf(B.z, w, C1[0]); C1[1] = fmaf(B.w, w, C1[1]); C1[2] = fmaf(Cc.x, w, C1[2]);
                    D1 = fmaf(Cc.y, w, D1);
                    T1 = test_T; last1 = pos;
                }
            }
        }
        // the slab written two rounds from now is this one: every lane must be done reading it
        __syncwarp();
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
}

// ------------------------------------------------------------------------------------------ backward
__device__ __forceinline__ void gms_bwd_pixel2(GmsBwdPix& p, int pos, float dx, float m2, float m4, float dy,
                                               const float4& B, const float2& Cc, float (&v)[10], bool& any) {
    if (pos >= p.last) return;
    const float power = gms_power(m2, m4, B.x, dy);
    if (power > 0.0f) return;
    const float G = gms_exp_fast(power);
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

struct GmsWarpSlabB {
    float4 a[GMS_WB];
    float4 b[GMS_WB];
    float2 c[GMS_WB];
    int id[GMS_WB];
    float part[GMS_WB][12];
};

__global__ void __launch_bounds__(GMS_CB)
k_composite_bwd2(const int2* __restrict__ ranges, const int* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                 const float* __restrict__ final_T, const int* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinv, float4* __restrict__ dgeom) {
    __shared__ GmsWarpSlabB s_slab[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;
    const GmsTileGeom g = gms_tile_geom(tile, gx, W, H, warp, lane);
    const int2 rng = ranges[tile];
    const float pxf = (float)g.px, pyf0 = (float)g.py0, pyf1 = (float)(g.py0 + 1);
    const float qx0 = (float)(g.tx0 + (warp & 1) * 8), qy0 = (float)(g.ty0 + (warp >> 1) * 8);
    const size_t HW = (size_t)H * W;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const float halfW = 0.5f * (float)W, halfH = 0.5f * (float)H;
    GmsWarpSlabB& S = s_slab[warp];

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
    if (wlast <= 0) return;
    const int nb = (wlast + GMS_WB - 1) / GMS_WB;

    int id_cur;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra;
    {
        const int k = (nb - 1) * GMS_WB + lane;
        id_cur = (k < wlast) ? (int)point_list[rng.x + k] : -1;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
    }
    int id_nx = (nb >= 2) ? (int)point_list[rng.x + (nb - 2) * GMS_WB + lane] : -1;

    for (int b = nb - 1; b >= 0; b--) {
        bool hit = false;
        S.id[lane] = id_cur;
        if (id_cur >= 0) {
            S.a[lane] = ra; S.b[lane] = rb; S.c[lane] = make_float2(rc.x, rc.y);
            hit = gms_reaches_quad(ra.x, ra.y, ra.z, ra.w, rb.x, rc.z, qx0, qy0);
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        id_cur = id_nx;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
        id_nx = (b >= 2) ? (int)point_list[rng.x + (b - 2) * GMS_WB + lane] : -1;
        __syncwarp();
        uint32_t touched = 0;
        while (m) {
            const int j = 31 - __clz(m);
            m &= ~(1u << j);
            const int pos = b * GMS_WB + j;
            const float4 A = S.a[j];
            const float4 B = S.b[j];
            const float2 Cc = S.c[j];
            const float dx = __fsub_rn(A.x, pxf);
            const float m1 = __fmul_rn(A.z, dx);
            const float m2 = __fmul_rn(m1, dx);
            const float m4 = __fmul_rn(A.w, dx);
            float v[10];
#pragma unroll
            for (int i = 0; i < 10; i++) v[i] = 0.f;
            bool any = false;
            gms_bwd_pixel2(p0, pos, dx, m2, m4, __fsub_rn(A.y, pyf0), B, Cc, v, any);
            gms_bwd_pixel2(p1, pos, dx, m2, m4, __fsub_rn(A.y, pyf1), B, Cc, v, any);
            if (!__any_sync(0xffffffffu, any)) continue;
            float out; int idx; bool valid;
            gms_fold10(v, lane, out, idx, valid);
            if (valid) S.part[j][idx] = out;
            touched |= 1u << j;
        }
        __syncwarp();
        if ((touched >> lane) & 1u) {
            const int id = S.id[lane];
            const float4 s0 = *reinterpret_cast<const float4*>(&S.part[lane][0]);
            const float4 s1 = *reinterpret_cast<const float4*>(&S.part[lane][4]);
            const float2 s2 = *reinterpret_cast<const float2*>(&S.part[lane][8]);
            const float4 A = S.a[lane];
            const float4 B = S.b[lane];
            const float conx = A.z, cony = A.w, conz = B.x, op = B.y;
            float4 g0, g1;
            g0.x = (-conx * s0.x - cony * s0.y) * op * halfW;   // dL/dmean2D.x (NDC-scaled)
            g0.y = (-conz * s0.y - cony * s0.x) * op * halfH;   // dL/dmean2D.y
            g0.z = -0.5f * op * s0.z;                           // dL/dconic.x
            g0.w = -0.5f * op * s0.w;                           // dL/dconic.y (stock half convention)
            g1.x = -0.5f * op * s1.x;                           // dL/dconic.z
            g1.y = s1.y;                                        // dL/d(conic_opacity.w)
            g1.z = s1.z; g1.w = s1.w;                           // dL/drgb.r, .g
            atomicAdd(&dgeom[3 * id], g0);
            atomicAdd(&dgeom[3 * id + 1], g1);
            atomicAdd(reinterpret_cast<float2*>(&dgeom[3 * id + 2]), s2);   // dL/drgb.b, dL/dinvdepth
        }
        __syncwarp();
    }
}
