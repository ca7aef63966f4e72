// gms_composite_fwd.cuh -- per-tile front-to-back compositing, forward (see gms_composite_common.cuh for the decomposition).
//
// k_composite_fwd2<EMIT>  default.  Scalar fp32, two pixels per lane.  EMIT: while compositing, every warp appends the list
//                         positions of the splats that actually BLENDED into at least one of its 64 pixels to a per-quad
//                         SURVIVOR LIST (a bit per visited splat in a lane-local mask, ONE warp OR-reduction per
//                         round, the set lanes store their positions compacted).  The backward pass then walks exactly those pairs (gms_composite_bwd.cuh):
//                         at 1M / 1080p 3.7 M of the 18.4 M (quad, splat) pairs, with no culling test and 1/5 of the rounds.
// k_composite_fwd3        A/B alternative: the same arithmetic as packed fp32x2 vectors (bit-identical result, slower:
//                         the forward is FMA/ALU-pipe bound and packing adds staging cost).
#pragma once
#include "gms_composite_common.cuh"

// ------------------------------------------------------------------------------------------- forward (default)
template <bool EMIT>
__global__ void __launch_bounds__(GMS_CB)
k_composite_fwd2(const int2* __restrict__ ranges, const int* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                 float* __restrict__ out_color, float* __restrict__ final_T, int* __restrict__ n_contrib,
                 float* __restrict__ out_invdepth, uint32_t* __restrict__ surv, uint32_t* __restrict__ nsurv) {
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
    // survivor list of this quad: region [4 * rng.x + warp * n, + n) of `surv`, positions relative to the tile's list
    uint32_t* const qlist = EMIT ? surv + 4 * (size_t)rng.x + (size_t)warp * n : nullptr;
    uint32_t scount = 0;

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
        uint32_t blended = 0;       // EMIT: bit j = splat j of this round blended into one of this lane's pixels
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
            bool b0, b1;
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
                b0 = ok;
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
                b1 = ok;
            }
            if (EMIT && (b0 || b1)) blended |= 1u << j;
        }
        if (EMIT) {     // one OR-reduction per round: the set lanes append their splat's list position, in order
            const uint32_t any = __reduce_or_sync(0xffffffffu, blended);
            if ((any >> lane) & 1u) qlist[scount + __popc(any & ((1u << lane) - 1u))] = (uint32_t)(base + lane);
            scount += __popc(any);
        }
        // the slab written two rounds from now is this one: every lane must be done reading it
        __syncwarp();
    }
    if (EMIT && lane == 0) nsurv[4 * tile + warp] = scount;
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


// ------------------------------------------------------------------------------------------- forward (packed f32x2, A/B)
__global__ void __launch_bounds__(GMS_CB)
k_composite_fwd3(const int2* __restrict__ ranges, const int* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                 float* __restrict__ out_color, float* __restrict__ final_T, int* __restrict__ n_contrib,
                 float* __restrict__ out_invdepth) {
    __shared__ GmsSlab3 s_slab[4][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;
    const GmsTileGeom g = gms_tile_geom(tile, gx, W, H, warp, lane);
    const int2 rng = ranges[tile];
    const int n = rng.y - rng.x;
    const f2 npx = make_float2(-(float)g.px, -(float)g.px);
    const f2 npy = make_float2(-(float)g.py0, -(float)(g.py0 + 1));
    const float qx0 = (float)(g.tx0 + (warp & 1) * 8), qy0 = (float)(g.ty0 + (warp >> 1) * 8);

    f2 T = make_float2(1.f, 1.f), D = make_float2(0.f, 0.f);
    f2 Cr = D, Cg = D, Cb = D;
    int last0 = 0, last1 = 0;
    bool live0 = g.in0, live1 = g.in1;

    int id_cur = (lane < n) ? (int)point_list[rng.x + lane] : -1;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, rc = ra;
    if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
    int id_nx = (GMS_WB + lane < n) ? (int)point_list[rng.x + GMS_WB + lane] : -1;

    for (int base = 0; base < n; base += GMS_WB) {
        if (!__any_sync(0xffffffffu, live0 || live1)) break;
        GmsSlab3& S = s_slab[warp][(base >> 5) & 1];
        bool hit = false;
        if (id_cur >= 0) {
            hit = gms_reaches_quad(ra.x, ra.y, ra.z, ra.w, rb.x, rc.z, qx0, qy0);
            if (hit) gms_slab3_store(S, lane, ra, rb, rc);
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
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
            const float4 Q0 = S.q0[j], Q1 = S.q1[j], Q2 = S.q2[j];
            const int pos = base + j + 1;
            f2 dx, dy;
            const f2 power = gms_power2(Q0, Q1, Q2, npx, npy, dx, dy);
            const f2 sc = f2mul(power, make_float2(GMS_LOG2E, GMS_LOG2E));
            const f2 G = make_float2(gms_ex2(sc.x), gms_ex2(sc.y));
            const f2 araw = f2mul(make_float2(Q2.z, Q2.w), G);
            const f2 alpha = make_float2(fminf(GMS_ALPHA_MAX, araw.x), fminf(GMS_ALPHA_MAX, araw.y));
            const f2 oma = f2fma(alpha, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));     // 1 - alpha, one rounding
            const f2 testT = f2mul(T, oma);
            bool ok0 = live0 && power.x <= 0.0f && alpha.x >= GMS_ALPHA_MIN;
            bool ok1 = live1 && power.y <= 0.0f && alpha.y >= GMS_ALPHA_MIN;
            if (ok0 && testT.x < GMS_T_STOP) { live0 = false; ok0 = false; }
            if (ok1 && testT.y < GMS_T_STOP) { live1 = false; ok1 = false; }
            if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;
            const float4 Q3 = S.q3[j], Q4 = S.q4[j];
            f2 w = f2mul(alpha, T);
            w.x = ok0 ? w.x : 0.f; w.y = ok1 ? w.y : 0.f;
            Cr = f2fma(make_float2(Q3.x, Q3.y), w, Cr);
            Cg = f2fma(make_float2(Q3.z, Q3.w), w, Cg);
            Cb = f2fma(make_float2(Q4.x, Q4.y), w, Cb);
            D = f2fma(make_float2(Q4.z, Q4.w), w, D);
            T.x = ok0 ? testT.x : T.x; T.y = ok1 ? testT.y : T.y;
            last0 = ok0 ? pos : last0; last1 = ok1 ? pos : last1;
        }
        __syncwarp();
    }
    const size_t HW = (size_t)H * W;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    if (g.in0) {
        const size_t pix = (size_t)g.py0 * W + g.px;
        final_T[pix] = T.x; n_contrib[pix] = last0;
        out_color[pix] = fmaf(T.x, bg0, Cr.x); out_color[HW + pix] = fmaf(T.x, bg1, Cg.x);
        out_color[2 * HW + pix] = fmaf(T.x, bg2, Cb.x);
        out_invdepth[pix] = D.x;
    }
    if (g.in1) {
        const size_t pix = (size_t)(g.py0 + 1) * W + g.px;
        final_T[pix] = T.y; n_contrib[pix] = last1;
        out_color[pix] = fmaf(T.y, bg0, Cr.y); out_color[HW + pix] = fmaf(T.y, bg1, Cg.y);
        out_color[2 * HW + pix] = fmaf(T.y, bg2, Cb.y);
        out_invdepth[pix] = D.y;
    }
}

