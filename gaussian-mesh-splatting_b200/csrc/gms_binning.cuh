// gms_binning.cuh -- tile binning WITHOUT a sort over the duplicates: one cooperative, persistent kernel, sm_100a.
//
// Replaces [upstream rasterizer_impl.cu: InclusiveSum + cudaMemcpy(num_rendered) + duplicateWithKeys +
// cub::DeviceRadixSort::SortPairs (64-bit keys, all N duplicates) + identifyTileRanges] behind the reference call
// renderer/gaussian_renderer/__init__.py:94-102.  What the stock pipeline computes is, for every tile, the list of
// Gaussians whose rectangle covers it, ordered by (depth bits, Gaussian index).  Once the P Gaussians are in that order
// (one stable 32-bit sort over P, gms_kernels.cu), the per-tile lists are a STABLE COUNTING SORT of the duplicates by
// tile id -- no keys have to be materialised and nothing has to be sorted again:
//
//   phase 1  every CTA owns a contiguous slice of the depth order and counts, per tile, how many of its Gaussians cover it
//            (one row of counters in shared memory, order-free: each lane walks its own rectangle)  -> M[cta][tile]  -> grid.sync
//   phase 2a per tile, exclusive scan of M[.][tile] over the CTAs (in place) and total[tile]                    -> grid.sync
//   phase 2b every CTA scans total[] (block scan in shared memory) into tile starts; base[tile] = start + M[cta][tile]
//            stays in shared memory; CTA 0 writes ranges[] and N.  N > capacity: overflow flag, empty ranges, no writes.
//   phase 3  the CTA walks its slice again, a batch of 32 Gaussians per warp, and writes
//            point_list[base[tile] + row[tile]++] = Gaussian id  with the batches COMMITTING IN DEPTH ORDER: a warp
//            prepares its batch on its own (loads two batches ahead, a warp scan numbers the batch's (Gaussian, tile) pairs
//            Gaussian-major, every step hands 32 consecutive pairs to the 32 lanes, lanes of a step that hit the same tile
//            are ranked with match.any), then waits for its turn (named barriers chain the warps: bar.arrive / bar.sync,
//            no spinning) and draws the slots with one shared-memory atomic per distinct tile and step.  Only those
//            atomics are serialised; 2 CTAs x 16 warps per SM keep everything else overlapped.
//
// The result is bit-identical to the stock (tile << 32 | depth) sort (tests/test_gpu_parity.py).  Device-side N, no host
// synchronisation, no scan over P, no key arrays.  Grid = 2 CTAs per SM (cooperative launch, grid.sync between the phases).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

struct GmsBinArgs {
    int P, T, gx;
    const uint32_t* order;      // [P] Gaussian ids in (depth bits, id) order; culled ones (key 0xFFFFFFFF) last
    const uint2* rect;          // [P] packed tile rectangle: .x = x0 | y0 << 16, .y = x1 | y1 << 16 (empty when culled)
    const uint32_t* nvis;       // device: number of visible Gaussians (they occupy order[0, nvis))
    uint32_t* M;                // [gridDim.x][T] scratch
    uint32_t* total;            // [T] scratch
    int2* ranges;               // [T] out
    uint32_t* point_list;       // [capacity] out
    uint32_t* tile_keys;        // [capacity] out, optional (debug views)
    uint32_t capacity;
    uint32_t* n_out;            // device [2]: N, overflow flag
    volatile uint32_t* n_host;  // mapped pinned host [2] or NULL: N, overflow flag (readable without a sync once the kernel ran)
};

constexpr int GMS_BIN_THREADS = 384;        // 12 warps per CTA (<= 15: one named barrier per warp hands the turn on)
constexpr int GMS_BIN_STEPS = 8;            // steps (of 32 pairs) prepared ahead of the ordered section

// dynamic shared memory: one row of T 32-bit counters + T 32-bit bases
static inline size_t gms_bin_smem_bytes(int T) { return 2 * (size_t)T * sizeof(uint32_t) + 64; }

// named barriers 1..W hand the "turn" from the warp that commits batch b to the one that commits batch b + 1 (barrier 1 + w
// is only ever waited on by warp w and arrived on by warp w - 1, so at most one hand-over is pending per barrier)
__device__ __forceinline__ void gms_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void gms_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void gms_bin_unpack(const uint2 r, int gx, int& t00, int& w, int& nt) {
    const int x0 = (int)(r.x & 0xffffu), y0 = (int)(r.x >> 16);
    w = (int)(r.y & 0xffffu) - x0;
    nt = w * ((int)(r.y >> 16) - y0);
    t00 = y0 * gx + x0;
    if (nt <= 0) { nt = 0; w = 1; }
}

__global__ void __launch_bounds__(GMS_BIN_THREADS, 2) k_bin_tiles(GmsBinArgs a) {
    extern __shared__ uint32_t bin_smem[];
    __shared__ uint32_t s_part[GMS_BIN_THREADS];
    __shared__ uint32_t s_grp[GMS_BIN_THREADS / 32][64];
    __shared__ uint32_t s_n;
    cg::grid_group grid = cg::this_grid();
    constexpr int W = GMS_BIN_THREADS / 32;
    const unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int T = a.T, G = (int)gridDim.x;
    uint32_t* row = bin_smem;               // [T] this CTA's running count per tile
    uint32_t* base = bin_smem + T;          // [T] first list position of this CTA's entries per tile
    for (int i = threadIdx.x; i < T; i += GMS_BIN_THREADS) row[i] = 0;
    // slice of the depth order owned by this CTA, in batches of 32 Gaussians dealt round-robin to the warps
    const uint32_t nvis = min(*a.nvis, (uint32_t)a.P);
    const uint32_t per = ((nvis + G - 1) / G + 31u) & ~31u;
    const uint32_t j_lo = min((uint32_t)blockIdx.x * per, nvis), j_hi = min(j_lo + per, nvis);
    const uint32_t nbatch = (j_hi - j_lo + 31u) / 32u;
    __syncthreads();

    // ---- phase 1: counting (order-free): lane = Gaussian, each lane walks its own rectangle
    {
        uint32_t b = warp;
        uint32_t g_nx = 0; uint2 r_cur = make_uint2(0u, 0u);
        if (b < nbatch) { const uint32_t j = j_lo + 32u * b + lane; if (j < j_hi) r_cur = a.rect[a.order[j]]; }
        if (b + W < nbatch) { const uint32_t j = j_lo + 32u * (b + W) + lane; if (j < j_hi) g_nx = a.order[j]; }
        for (; b < nbatch; b += W) {
            int t00, w, nt;
            gms_bin_unpack(r_cur, a.gx, t00, w, nt);
            {   // prefetch: rectangle of the next batch of this warp, id of the one after
                const uint32_t j1 = j_lo + 32u * (b + W) + lane, j2 = j_lo + 32u * (b + 2 * W) + lane;
                r_cur = (b + W < nbatch && j1 < j_hi) ? a.rect[g_nx] : make_uint2(0u, 0u);
                g_nx = (b + 2 * W < nbatch && j2 < j_hi) ? a.order[j2] : 0u;
            }
            const bool big = nt >= 64;
            if (!big) {
                int t = t00, x = 0;
                for (int k = 0; k < nt; k++) {
                    atomicAdd(&row[t], 1u);
                    if (++x == w) { x = 0; t += a.gx - w + 1; } else t++;
                }
            }
            uint32_t m = __ballot_sync(FULL, big);
            while (m) {     // large rectangles: the warp shares one
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int t0s = __shfl_sync(FULL, t00, src), ws = __shfl_sync(FULL, w, src), nts = __shfl_sync(FULL, nt, src);
                for (int k = lane; k < nts; k += 32) { const int yy = k / ws; atomicAdd(&row[t0s + yy * a.gx + (k - yy * ws)], 1u); }
            }
        }
    }
    __syncthreads();
    {   // CTA count per tile -> M; the row restarts from 0 and becomes the running count of phase 3
        uint32_t* Mrow = a.M + (size_t)blockIdx.x * T;
        for (int i = threadIdx.x; i < T; i += GMS_BIN_THREADS) { Mrow[i] = row[i]; row[i] = 0; }
    }
    grid.sync();

    // ---- phase 2a: per tile, exclusive scan over the CTAs.  64 adjacent tiles per CTA round (two per lane: coalesced rows
    // of M); the W warps split the CTA axis, combine their partial sums through shared memory, then write the offsets.
    {
        const int cper = (G + W - 1) / W;
        const int c_lo = min(warp * cper, G), c_hi = min(c_lo + cper, G);
        for (int chunk = blockIdx.x; chunk * 64 < T; chunk += G) {
            const int ta = chunk * 64 + lane, tb = ta + 32;
            const bool oka = ta < T, okb = tb < T;
            uint32_t suma = 0, sumb = 0;
            for (int c = c_lo; c < c_hi; c += 4) {
                uint32_t va[4], vb[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    va[u] = (oka && c + u < c_hi) ? a.M[(size_t)(c + u) * T + ta] : 0u;
                    vb[u] = (okb && c + u < c_hi) ? a.M[(size_t)(c + u) * T + tb] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) { suma += va[u]; sumb += vb[u]; }
            }
            s_grp[warp][lane] = suma; s_grp[warp][lane + 32] = sumb;
            __syncthreads();
            uint32_t runa = 0, tota = 0, runb = 0, totb = 0;
            for (int q = 0; q < W; q++) {
                const uint32_t xa = s_grp[q][lane], xb = s_grp[q][lane + 32];
                if (q < warp) { runa += xa; runb += xb; }
                tota += xa; totb += xb;
            }
            for (int c = c_lo; c < c_hi; c += 4) {
                uint32_t va[4], vb[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    va[u] = (oka && c + u < c_hi) ? a.M[(size_t)(c + u) * T + ta] : 0u;
                    vb[u] = (okb && c + u < c_hi) ? a.M[(size_t)(c + u) * T + tb] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (oka && c + u < c_hi) a.M[(size_t)(c + u) * T + ta] = runa;
                    if (okb && c + u < c_hi) a.M[(size_t)(c + u) * T + tb] = runb;
                    runa += va[u]; runb += vb[u];
                }
            }
            if (warp == 0) { if (oka) a.total[ta] = tota; if (okb) a.total[tb] = totb; }
            __syncthreads();
        }
    }
    grid.sync();

    // ---- phase 2b: tile starts (every CTA, redundantly, in shared memory); base[] = start + this CTA's offset
    {
        for (int i = threadIdx.x; i < T; i += GMS_BIN_THREADS) base[i] = a.total[i];
        __syncthreads();
        const int per_t = (T + GMS_BIN_THREADS - 1) / GMS_BIN_THREADS;
        const int t0 = min((int)threadIdx.x * per_t, T), t1 = min(t0 + per_t, T);
        uint32_t sum = 0;
        for (int t = t0; t < t1; t++) sum += base[t];
        s_part[threadIdx.x] = sum;
        __syncthreads();
        if (warp == 0) {        // exclusive scan of the partials by one warp
            uint32_t carry = 0;
            for (int c0 = 0; c0 < GMS_BIN_THREADS; c0 += 32) {
                const uint32_t v = s_part[c0 + lane];
                uint32_t x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
                s_part[c0 + lane] = carry + x - v;
                carry += __shfl_sync(FULL, x, 31);
            }
            if (lane == 0) s_n = carry;
        }
        __syncthreads();
        const uint32_t N = s_n;
        const bool overflow = N > a.capacity;
        uint32_t run = s_part[threadIdx.x];
        for (int t = t0; t < t1; t++) {
            const uint32_t tot = base[t];
            base[t] = run;
            if (blockIdx.x == 0) a.ranges[t] = (overflow || tot == 0) ? make_int2(0, 0) : make_int2((int)run, (int)(run + tot));   // untouched tiles stay (0, 0) like the stock's
            run += tot;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.n_out[0] = N; a.n_out[1] = overflow ? 1u : 0u;
            if (a.n_host) { a.n_host[0] = N; a.n_host[1] = overflow ? 1u : 0u; }
        }
        if (overflow) return;       // uniform over the grid: no further grid.sync follows
        __syncthreads();
        const uint32_t* Mrow = a.M + (size_t)blockIdx.x * T;
        for (int i = threadIdx.x; i < T; i += GMS_BIN_THREADS) base[i] += Mrow[i];
        __syncthreads();
    }

    // ---- phase 3: ordered placement.  Each warp prepares its batch (32 Gaussians -> their (Gaussian, tile) pairs numbered
    // Gaussian-major, 32 consecutive pairs per step; lanes of a step that hit the same tile are ranked with match.any), then
    // takes its TURN -- batches commit in depth order -- to draw the slots from the CTA's row with one shared-memory atomic
    // per distinct tile and step.  Only the atomics are serialised; loads, pair numbering and the stores overlap across warps.
    {
        uint32_t b = warp;
        uint32_t g_cur = 0, g_nx = 0; uint2 r_cur = make_uint2(0u, 0u);
        if (b < nbatch) { const uint32_t j = j_lo + 32u * b + lane; if (j < j_hi) { g_cur = a.order[j]; r_cur = a.rect[g_cur]; } }
        if (b + W < nbatch) { const uint32_t j = j_lo + 32u * (b + W) + lane; if (j < j_hi) g_nx = a.order[j]; }
        for (; b < nbatch; b += W) {
            const uint32_t g = g_cur;
            int t00, w, nt;
            gms_bin_unpack(r_cur, a.gx, t00, w, nt);
            {
                const uint32_t j1 = j_lo + 32u * (b + W) + lane, j2 = j_lo + 32u * (b + 2 * W) + lane;
                g_cur = g_nx;
                r_cur = (b + W < nbatch && j1 < j_hi) ? a.rect[g_cur] : make_uint2(0u, 0u);
                g_nx = (b + 2 * W < nbatch && j2 < j_hi) ? a.order[j2] : 0u;
            }
            int inc = nt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += y; }
            const int tot = __shfl_sync(FULL, inc, 31);
            const int exc = inc - nt;
            const float rw = __fdividef(1.0f, (float)w);
            bool mine = false;      // this warp holds the turn
            for (int p0 = 0; p0 < tot || !mine; p0 += 32 * GMS_BIN_STEPS) {
                int tt[GMS_BIN_STEPS]; uint32_t gg[GMS_BIN_STEPS], meta[GMS_BIN_STEPS];
#pragma unroll
                for (int s = 0; s < GMS_BIN_STEPS; s++) {
                    tt[s] = -1; gg[s] = 0; meta[s] = 0;
                    if (p0 + 32 * s >= tot) continue;       // (uniform) nothing beyond this step
                    const int p = p0 + 32 * s + lane;
                    const bool active = p < tot;
                    int lo = 0;         // owner = first lane whose inclusive prefix exceeds p
#pragma unroll
                    for (int st = 16; st >= 1; st >>= 1) { const int v = __shfl_sync(FULL, inc, lo + st - 1); if (v <= p) lo += st; }
                    const int owner = lo & 31;
                    const int k = p - __shfl_sync(FULL, exc, owner);
                    const int t0s = __shfl_sync(FULL, t00, owner), ws = __shfl_sync(FULL, w, owner);
                    const float rws = __shfl_sync(FULL, rw, owner);
                    gg[s] = __shfl_sync(FULL, g, owner);
                    const int yy = (int)(((float)k + 0.5f) * rws);          // floor(k / ws): exact for k < 2^20 (error of the approximate reciprocal << 0.5 / ws)
                    const int t = t0s + yy * a.gx + (k - yy * ws);
                    tt[s] = active ? t : -1;
                    const uint32_t peers = __match_any_sync(FULL, active ? (uint32_t)t : (0x40000000u | (uint32_t)lane));
                    meta[s] = (uint32_t)(__ffs(peers) - 1) | ((uint32_t)__popc(peers & lt) << 8) | ((uint32_t)__popc(peers) << 16);
                }
                if (!mine) {        // wait for batch b - 1 to have drawn its slots (named barrier: no spinning, ~tens of cycles)
                    if (b > 0) gms_bar_sync(1 + warp, 64);          // barrier (1 + w) belongs to the RECEIVING warp w: only warp w - 1 arrives on it
                    mine = true;
                }
                uint32_t old[GMS_BIN_STEPS];
#pragma unroll
                for (int s = 0; s < GMS_BIN_STEPS; s++) {       // the ordered section: nothing but the atomics, back to back
                    old[s] = 0;
                    if (tt[s] >= 0 && lane == (int)(meta[s] & 0xffu)) old[s] = atomicAdd(&row[tt[s]], meta[s] >> 16);
                }
#pragma unroll
                for (int s = 0; s < GMS_BIN_STEPS; s++) old[s] = __shfl_sync(FULL, old[s], (int)(meta[s] & 0xffu));   // (consumes the results: the atomics have been performed)
                const bool last_chunk = p0 + 32 * GMS_BIN_STEPS >= tot;
                if (last_chunk && b + 1 < nbatch) gms_bar_arrive(1 + (warp + 1) % W, 64);      // hand the turn to batch b + 1 (warp w + 1)
#pragma unroll
                for (int s = 0; s < GMS_BIN_STEPS; s++) {
                    if (tt[s] >= 0) {
                        const uint32_t pos = base[tt[s]] + old[s] + ((meta[s] >> 8) & 0xffu);
                        a.point_list[pos] = gg[s];
                        if (a.tile_keys) a.tile_keys[pos] = (uint32_t)tt[s];
                    }
                }
                if (last_chunk) break;
            }
        }
    }
}
