// gms_composite4.cuh -- composite forward/backward, generation 4: OCTET-INDEPENDENT lists.
//
// Generations 2/3 make a warp (an 8x8 pixel quad) visit every splat that reaches ANY of its 64 pixels; measured on the
// headline scene only ~22 of the 64 pixels blend a visited splat.  Here the warp's four lane-octets -- each an 8x2 pixel
// strip (lane = column, two rows per lane) -- advance through the 32-splat round INDEPENDENTLY: at staging every lane
// tests its splat against the four strips (exact ellipse-vs-rectangle test), four ballots give one bitmask per strip,
// and in each loop iteration every octet takes the next splat of ITS OWN mask.  Pixels of different strips are
// independent, and inside a strip the order is unchanged, so the result is bit-identical; the number of iterations per
// round drops from |union of the four masks| to max |mask| (0.71x on the headline scene, oracle counters), and the
// backward's shuffle reduction now folds up to FOUR splats at once (10 shuffles over 8 lanes instead of 12 over 32).
// Packed fp32x2 arithmetic as in generation 3 for the backward; scalar pair arithmetic (generation 2) for the forward.
#pragma once
#include "gms_composite3.cuh"

// min over the pixel rectangle [rx0, rx0+w] x [ry0, ry0+h] of 0.5*(cx dx^2 + cz dy^2) + cy dx dy  <=  tau ?
__device__ __forceinline__ bool gms_reaches_rect(float x, float y, float cx, float cy, float cz, float icx, float icz, float tau,
                                                 float rx0, float ry0, float w, float h) {
    const float rx1 = rx0 + w, ry1 = ry0 + h;
    if (x >= rx0 && x <= rx1 && y >= ry0 && y <= ry1) return true;
    float best = 3.0e38f;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        {
            const float dy = y - (e ? ry1 : ry0);
            const float t = fminf(fmaxf(x + cy * dy * icx, rx0), rx1);
            const float dx = x - t;
            best = fminf(best, 0.5f * (cx * dx * dx + cz * dy * dy) + cy * dx * dy);
        }
        {
            const float dx = x - (e ? rx1 : rx0);
            const float t = fminf(fmaxf(y + cy * dx * icz, ry0), ry1);
            const float dy = y - t;
            best = fminf(best, 0.5f * (cx * dx * dx + cz * dy * dy) + cy * dx * dy);
        }
    }
    return best <= tau;
}

// per-strip hit bits (bit s = 8x2 strip s of the quad at (qx0, qy0)) of one splat
__device__ __forceinline__ uint32_t gms_strip_hits(const float4& ra, const float4& rb, const float4& rc, float qx0, float qy0) {
    const float x = ra.x, y = ra.y, cx = ra.z, cy = ra.w, cz = rb.x, tau = rc.z;
    if (!(tau > 0.f)) return 0u;
    if (!(cx > 0.f) || !(cz > 0.f)) return 0xFu;
    const float icx = __fdividef(1.f, cx), icz = __fdividef(1.f, cz);
    if (!gms_reaches_rect(x, y, cx, cy, cz, icx, icz, tau, qx0, qy0, 7.f, 7.f)) return 0u;
    uint32_t h = 0;
#pragma unroll
    for (int s = 0; s < 4; s++)
        h |= (gms_reaches_rect(x, y, cx, cy, cz, icx, icz, tau, qx0, qy0 + 2.f * (float)s, 7.f, 1.f) ? 1u : 0u) << s;
    return h;
}

// ------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(GMS_CB)
k_composite_fwd4(const int2* __restrict__ ranges, const int* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
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
    const int octet = lane >> 3;

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
        uint32_t hits = 0;
        if (id_cur >= 0) {
            hits = gms_strip_hits(ra, rb, rc, qx0, qy0);
            // stored unconditionally: an idle octet (empty mask) reads row 0 as a harmless dummy, so every row of a valid
            // splat must hold finite data
            S.a[lane] = ra; S.b[lane] = rb; S.c[lane] = make_float2(rc.x, rc.y);
        }
        const uint32_t b0 = __ballot_sync(0xffffffffu, hits & 1u), b1 = __ballot_sync(0xffffffffu, hits & 2u);
        const uint32_t b2 = __ballot_sync(0xffffffffu, hits & 4u), b3 = __ballot_sync(0xffffffffu, hits & 8u);
        uint32_t m = octet == 0 ? b0 : (octet == 1 ? b1 : (octet == 2 ? b2 : b3));      // this octet's own list
        id_cur = id_nx;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
        {
            const int k = base + 2 * GMS_WB + lane;
            id_nx = (k < n) ? (int)point_list[rng.x + k] : -1;
        }
        __syncwarp();
        while (__any_sync(0xffffffffu, m != 0u)) {
            const bool has = m != 0u;
            const int j = has ? __ffs(m) - 1 : 0;
            m &= m - 1u;                                   // (0 & 0xffffffff) stays 0
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
                bool ok = has && live0 && power <= 0.0f && alpha >= GMS_ALPHA_MIN;
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
                bool ok = has && live1 && power <= 0.0f && alpha >= GMS_ALPHA_MIN;
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
// transpose-reduce of 10 per-lane values over the 8 lanes of an octet in 10 shuffles; every lane ends with two slots
// (o0, o1) holding the octet-wide sums of values i0 and i0+1 (each of the 10 values lands in exactly one valid slot).
__device__ __forceinline__ void gms_fold10_octet(const float (&v)[10], int lane, float& o0, float& o1, int& i0, bool& ok0, bool& ok1) {
    const unsigned F = 0xffffffffu;
    bool hi = (lane & 4) != 0;
    float w[6];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float send = hi ? v[i] : v[i + 5];
        const float keep = hi ? v[i + 5] : v[i];
        w[i] = keep + __shfl_xor_sync(F, send, 4);
    }
    w[5] = 0.f;
    hi = (lane & 2) != 0;
    float x[4];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float send = hi ? w[i] : w[i + 3];
        const float keep = hi ? w[i + 3] : w[i];
        x[i] = keep + __shfl_xor_sync(F, send, 2);
    }
    x[3] = 0.f;
    hi = (lane & 1) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = hi ? x[i] : x[i + 2];
        const float keep = hi ? x[i + 2] : x[i];
        const float r = keep + __shfl_xor_sync(F, send, 1);
        if (i == 0) o0 = r; else o1 = r;
    }
    const int b1 = (lane & 4) ? 5 : 0, b2 = (lane & 2) ? 3 : 0, b3 = (lane & 1) ? 2 : 0;
    // slot e (0/1): x-level index b3 + e must be a real entry (< 3), w-level index b2 + b3 + e must be < 5
    ok0 = (b3 + 0 <= 2) && (b2 + b3 + 0 <= 4);
    ok1 = (b3 + 1 <= 2) && (b2 + b3 + 1 <= 4);
    i0 = b1 + b2 + b3;
}

struct GmsSlab4B {
    float4 q0[GMS_WB], q1[GMS_WB], q2[GMS_WB], q3[GMS_WB], q4[GMS_WB];
    int id[GMS_WB];
    float part[4][GMS_WB][12];       // [octet][splat][value]
};

template <int MINB>
__global__ void __launch_bounds__(GMS_CB, MINB)
k_composite_bwd4(const int2* __restrict__ ranges, const int* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ recs, int W, int H, int gx, const float* __restrict__ bg,
                 const float* __restrict__ final_T, const int* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, const float* __restrict__ dL_dinv, float4* __restrict__ dgeom) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    GmsSlab4B* s_slab = reinterpret_cast<GmsSlab4B*>(s_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;
    const GmsTileGeom g = gms_tile_geom(tile, gx, W, H, warp, lane);
    const int2 rng = ranges[tile];
    const f2 npx = make_float2(-(float)g.px, -(float)g.px);
    const f2 npy = make_float2(-(float)g.py0, -(float)(g.py0 + 1));
    const float qx0 = (float)(g.tx0 + (warp & 1) * 8), qy0 = (float)(g.ty0 + (warp >> 1) * 8);
    const size_t HW = (size_t)H * W;
    const float halfW = 0.5f * (float)W, halfH = 0.5f * (float)H;
    const int octet = lane >> 3;
    GmsSlab4B& S = s_slab[warp];

    f2 T, nTfin, dpr, dpg, dpb, dpd, bgdot;
    int lastA = 0, lastB = 0;
    {
        float tf[2] = {1.f, 1.f}, r[2] = {0.f, 0.f}, gg[2] = {0.f, 0.f}, b[2] = {0.f, 0.f}, dd[2] = {0.f, 0.f};
        int la[2] = {0, 0};
        const bool in[2] = {g.in0, g.in1};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (in[k]) {
                const size_t pix = (size_t)(g.py0 + k) * W + g.px;
                tf[k] = final_T[pix]; la[k] = n_contrib[pix];
                r[k] = dL_dpix[pix]; gg[k] = dL_dpix[HW + pix]; b[k] = dL_dpix[2 * HW + pix];
                dd[k] = dL_dinv ? dL_dinv[pix] : 0.f;
            }
        }
        const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
        T = make_float2(tf[0], tf[1]); nTfin = make_float2(-tf[0], -tf[1]);
        dpr = make_float2(r[0], r[1]); dpg = make_float2(gg[0], gg[1]); dpb = make_float2(b[0], b[1]); dpd = make_float2(dd[0], dd[1]);
        bgdot = make_float2(bg0 * r[0] + bg1 * gg[0] + bg2 * b[0], bg0 * r[1] + bg1 * gg[1] + bg2 * b[1]);
        lastA = la[0]; lastB = la[1];
    }
    f2 Br = make_float2(0.f, 0.f), Bg = Br, Bb = Br, Bd = Br;
    const int wlast = __reduce_max_sync(0xffffffffu, max(lastA, lastB));
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
        uint32_t hits = 0;
        S.id[lane] = id_cur;
        if (id_cur >= 0) {
            hits = gms_strip_hits(ra, rb, rc, qx0, qy0);
            gms_slab3_store(reinterpret_cast<GmsSlab3&>(S), lane, ra, rb, rc);   // unconditional: row 0 doubles as the idle octets' dummy
        }
        const uint32_t b0 = __ballot_sync(0xffffffffu, hits & 1u), b1 = __ballot_sync(0xffffffffu, hits & 2u);
        const uint32_t b2 = __ballot_sync(0xffffffffu, hits & 4u), b3 = __ballot_sync(0xffffffffu, hits & 8u);
        uint32_t m = octet == 0 ? b0 : (octet == 1 ? b1 : (octet == 2 ? b2 : b3));
        id_cur = id_nx;
        if (id_cur >= 0) { ra = recs[3 * id_cur]; rb = recs[3 * id_cur + 1]; rc = recs[3 * id_cur + 2]; }
        id_nx = (b >= 2) ? (int)point_list[rng.x + (b - 2) * GMS_WB + lane] : -1;
        __syncwarp();
        uint32_t touched = 0;                    // bit j: this octet stored sums for splat j
        while (__any_sync(0xffffffffu, m != 0u)) {
            const bool has = m != 0u;
            const int j = has ? 31 - __clz(m) : 0;
            m &= ~(has ? (1u << j) : 0u);
            const int pos = b * GMS_WB + j;
            const float4 Q0 = S.q0[j], Q1 = S.q1[j], Q2 = S.q2[j];
            f2 dx, dy;
            const f2 power = gms_power2(Q0, Q1, Q2, npx, npy, dx, dy);
            const f2 sc = f2mul(power, make_float2(GMS_LOG2E, GMS_LOG2E));
            const f2 G = make_float2(gms_ex2(sc.x), gms_ex2(sc.y));
            const f2 araw = f2mul(make_float2(Q2.z, Q2.w), G);
            const float a0 = fminf(GMS_ALPHA_MAX, araw.x), a1 = fminf(GMS_ALPHA_MAX, araw.y);
            const bool v0 = has && pos < lastA && power.x <= 0.0f && a0 >= GMS_ALPHA_MIN;
            const bool v1 = has && pos < lastB && power.y <= 0.0f && a1 >= GMS_ALPHA_MIN;
            const uint32_t anyv = __ballot_sync(0xffffffffu, v0 || v1);
            if (anyv == 0u) continue;
            const float4 Q3 = S.q3[j], Q4 = S.q4[j];
            const f2 alpha = make_float2(v0 ? a0 : 0.f, v1 ? a1 : 0.f);
            const f2 oma = f2fma(alpha, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
            const f2 inv = make_float2(gms_rcp(oma.x), gms_rcp(oma.y));
            T = f2mul(T, inv);
            const f2 w = f2mul(alpha, T);
            const f2 cr = make_float2(Q3.x, Q3.y), cg = make_float2(Q3.z, Q3.w), cb = make_float2(Q4.x, Q4.y), cd = make_float2(Q4.z, Q4.w);
            const f2 neg1 = make_float2(-1.f, -1.f);
            f2 dLa = f2mul(f2fma(Br, neg1, cr), dpr);
            dLa = f2fma(f2fma(Bg, neg1, cg), dpg, dLa);
            dLa = f2fma(f2fma(Bb, neg1, cb), dpb, dLa);
            dLa = f2fma(f2fma(Bd, neg1, cd), dpd, dLa);
            Br = f2fma(alpha, cr, f2mul(oma, Br)); Bg = f2fma(alpha, cg, f2mul(oma, Bg));
            Bb = f2fma(alpha, cb, f2mul(oma, Bb)); Bd = f2fma(alpha, cd, f2mul(oma, Bd));
            dLa = f2mul(dLa, T);
            dLa = f2fma(f2mul(nTfin, inv), bgdot, dLa);
            f2 q = f2mul(dLa, G);
            q.x = v0 ? q.x : 0.f; q.y = v1 ? q.y : 0.f;
            const f2 qx = f2mul(q, dx), qy = f2mul(q, dy);
            const f2 pxx = f2mul(qx, dx), pxy = f2mul(qx, dy), pyy = f2mul(qy, dy);
            const f2 wr = f2mul(w, dpr), wg = f2mul(w, dpg), wb = f2mul(w, dpb), wd = f2mul(w, dpd);
            float v[10];
            v[0] = qx.x + qx.y; v[1] = qy.x + qy.y; v[2] = pxx.x + pxx.y; v[3] = pxy.x + pxy.y; v[4] = pyy.x + pyy.y;
            v[5] = q.x + q.y; v[6] = wr.x + wr.y; v[7] = wg.x + wg.y; v[8] = wb.x + wb.y; v[9] = wd.x + wd.y;
            float o0, o1; int i0; bool ok0, ok1;
            gms_fold10_octet(v, lane, o0, o1, i0, ok0, ok1);
            // did any pixel of THIS octet blend its splat?
            const bool oct_any = ((anyv >> (octet * 8)) & 0xFFu) != 0u;
            if (oct_any) {
                if (ok0) S.part[octet][j][i0] = o0;
                if (ok1) S.part[octet][j][i0 + 1] = o1;
                touched |= 1u << j;
            }
        }
        __syncwarp();
        // lanes of octet o hold the same `touched` word for o; gather the four words
        const uint32_t t0 = __shfl_sync(0xffffffffu, touched, 0), t1 = __shfl_sync(0xffffffffu, touched, 8);
        const uint32_t t2 = __shfl_sync(0xffffffffu, touched, 16), t3 = __shfl_sync(0xffffffffu, touched, 24);
        const uint32_t tl[4] = {t0, t1, t2, t3};
        if (((t0 | t1 | t2 | t3) >> lane) & 1u) {
            float s[10];
#pragma unroll
            for (int i = 0; i < 10; i++) s[i] = 0.f;
#pragma unroll
            for (int o = 0; o < 4; o++) {
                if ((tl[o] >> lane) & 1u) {
                    const float4 p0 = *reinterpret_cast<const float4*>(&S.part[o][lane][0]);
                    const float4 p1 = *reinterpret_cast<const float4*>(&S.part[o][lane][4]);
                    const float2 p2 = *reinterpret_cast<const float2*>(&S.part[o][lane][8]);
                    s[0] += p0.x; s[1] += p0.y; s[2] += p0.z; s[3] += p0.w; s[4] += p1.x; s[5] += p1.y; s[6] += p1.z; s[7] += p1.w;
                    s[8] += p2.x; s[9] += p2.y;
                }
            }
            const int id = S.id[lane];
            const float4 Q1 = S.q1[lane], Q2 = S.q2[lane];
            const float conx = Q1.x, ncony = Q1.z, conz = Q2.x, op = Q2.z;
            float4 g0, g1;
            g0.x = (-conx * s[0] + ncony * s[1]) * op * halfW;
            g0.y = (-conz * s[1] + ncony * s[0]) * op * halfH;
            g0.z = -0.5f * op * s[2];
            g0.w = -0.5f * op * s[3];
            g1.x = -0.5f * op * s[4];
            g1.y = s[5];
            g1.z = s[6]; g1.w = s[7];
            atomicAdd(&dgeom[3 * id], g0);
            atomicAdd(&dgeom[3 * id + 1], g1);
            atomicAdd(reinterpret_cast<float2*>(&dgeom[3 * id + 2]), make_float2(s[8], s[9]));
        }
        __syncwarp();
    }
}
