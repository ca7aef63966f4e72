// gms_binning.cuh -- tile binning WITHOUT a sort over the duplicates: one cooperative, persistent kernel, sm_100a.
//
// Replaces [upstream rasterizer_impl.cu: InclusiveSum + cudaMemcpy(num_rendered) + duplicateWithKeys +
// cub::DeviceRadixSort::SortPairs (64-bit keys, all N duplicates) + identifyTileRanges] behind the reference call
// renderer/gaussian_renderer/__init__.py:94-102.  What the stock pipeline computes is, for every tile, the list of
// Gaussians whose rectangle covers it, ordered by (depth bits, Gaussian index).  Once the P Gaussians are in that order
// (one stable 32-bit sort over P, gms_kernels.cu), the per-tile lists are a STABLE COUNTING SORT of the duplicates by
// tile id -- no keys have to be materialised and nothing has to be sorted again:
//
//   phase 1  every warp owns a contiguous slice of the depth order and counts, per tile, how many of its Gaussians cover
//            it (16-bit counters in the warp's own shared-memory row, two tiles per 32-bit word);  a CTA-level
//            exclusive scan over the warps' rows turns them into warp offsets and yields the CTA's per-tile count
//            M[cta][tile]                                                                                  -> grid.sync
//   phase 2a per tile, exclusive scan of M[.][tile] over the CTAs (in place) and total[tile]              -> grid.sync
//   phase 2b every CTA scans total[] (block scan in shared memory) into tile starts; base[tile] = start + M[cta][tile]
//            stays in shared memory; CTA 0 writes ranges[] and N.  N > capacity: overflow flag, empty ranges, no writes.
//   phase 3  every warp walks its slice again IN ORDER and writes
//            point_list[base[tile] + row[warp][tile]++] = Gaussian id.
// Both walks are DENSE over (Gaussian, tile) pairs: a batch of 32 Gaussians is loaded one per lane (ids and rectangles
// two / one batch ahead), a warp scan of the rectangle areas numbers the batch's pairs Gaussian-major, and every step hands
// 32 consecutive pairs to the 32 lanes (owner found by a 5-step shuffle search) -- full lanes whatever the rectangle sizes.
// In phase 3 the lanes of a step that hit the same tile (different Gaussians, lane order = depth order) are ranked with
// match.any and served by ONE shared-memory atomic; steps follow each other in program order, so slots come out in depth order.
//
// The result is bit-identical to the stock (tile << 32 | depth) sort (tests/test_gpu_parity.py).  Device-side N, no host
// synchronisation, no scan over P, no key arrays: at 1M Gaussians / 1080p it replaces 0.016 (scan) + 0.04 (emit) + 0.125
// (sort, cub) + 0.016 (ranges) ms.  The grid is one CTA per SM (cooperative launch, grid.sync between the phases).
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

// shared memory: W rows of T 16-bit counters (as ceil(T/2) words) + T 32-bit bases
static inline size_t gms_bin_smem_bytes(int T, int warps) {
    const size_t words = (size_t)(T + 1) / 2;
    return (words * warps + (size_t)T) * sizeof(uint32_t) + 64;
}
// largest warp count (power of two, <= 16) whose rows fit next to the bases in `budget` bytes; 0: does not fit at all
static inline int gms_bin_warps(int T, size_t budget) {
    for (int w = 16; w >= 2; w >>= 1)
        if (gms_bin_smem_bytes(T, w) <= budget) return w;
    return 0;
}

__device__ __forceinline__ void gms_bin_unpack(const uint2 r, int gx, int& t00, int& w, int& nt) {
    const int x0 = (int)(r.x & 0xffffu), y0 = (int)(r.x >> 16);
    w = (int)(r.y & 0xffffu) - x0;
    nt = w * ((int)(r.y >> 16) - y0);
    t00 = y0 * gx + x0;
    if (nt <= 0) { nt = 0; w = 1; }
}

// One pass over the warp's slice [j_lo, j_hi) of the depth order, 32 (Gaussian, tile) pairs per step.
// PLACE = false: count into `row`.  PLACE = true: ordered slots from `row`, ids to point_list.
template <bool PLACE>
__device__ __forceinline__ void gms_bin_walk(const GmsBinArgs& a, uint32_t j_lo, uint32_t j_hi, int lane, uint32_t* row, const uint32_t* base) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t lt = (1u << lane) - 1u;
    // software pipeline: ids two batches ahead, rectangles one batch ahead
    uint32_t g_cur = 0, g_nx = 0;
    uint2 r_cur = make_uint2(0u, 0u);
    if (j_lo + lane < j_hi) { g_cur = a.order[j_lo + lane]; r_cur = a.rect[g_cur]; }
    if (j_lo + 32 + lane < j_hi) g_nx = a.order[j_lo + 32 + lane];
    for (uint32_t j0 = j_lo; j0 < j_hi; j0 += 32) {
        const uint32_t g = g_cur;
        int t00, w, nt;
        gms_bin_unpack((j0 + lane < j_hi) ? r_cur : make_uint2(0u, 0u), a.gx, t00, w, nt);
        // prefetch
        g_cur = g_nx;
        r_cur = (j0 + 32 + lane < j_hi) ? a.rect[g_cur] : make_uint2(0u, 0u);
        g_nx = (j0 + 64 + lane < j_hi) ? a.order[j0 + 64 + lane] : 0u;
        // number the batch's pairs Gaussian-major
        int inc = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += y; }
        const int tot = __shfl_sync(FULL, inc, 31);
        const int exc = inc - nt;
        for (int p0 = 0; p0 < tot; p0 += 32) {
            const int p = p0 + lane;
            const bool active = p < tot;
            int lo = 0;         // owner = first lane whose inclusive prefix exceeds p
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) { const int v = __shfl_sync(FULL, inc, lo + st - 1); if (v <= p) lo += st; }
            const int owner = lo & 31;
            const int k = p - __shfl_sync(FULL, exc, owner);
            const int t0s = __shfl_sync(FULL, t00, owner), ws = __shfl_sync(FULL, w, owner);
            const int yy = (int)(((float)k + 0.5f) * __frcp_rn((float)ws));      // exact floor(k / ws) for k < 2^20
            const int t = t0s + yy * a.gx + (k - yy * ws);
            const int sh = (t & 1) * 16;
            if (!PLACE) {
                if (active) atomicAdd(&row[t >> 1], 1u << sh);
            } else {
                const uint32_t gs = __shfl_sync(FULL, g, owner);
                const uint32_t peers = __match_any_sync(FULL, active ? (uint32_t)t : (0x40000000u | (uint32_t)lane));
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (active && lane == leader) old = atomicAdd(&row[t >> 1], (uint32_t)__popc(peers) << sh);
                old = __shfl_sync(FULL, old, leader);
                if (active) {
                    const uint32_t pos = base[t] + ((old >> sh) & 0xffffu) + (uint32_t)__popc(peers & lt);
                    a.point_list[pos] = gs;
                    if (a.tile_keys) a.tile_keys[pos] = (uint32_t)t;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(512, 1) k_bin_tiles(GmsBinArgs a) {
    extern __shared__ uint32_t bin_smem[];
    __shared__ uint32_t s_part[512];
    __shared__ uint32_t s_grp[16][64];
    __shared__ uint32_t s_n;
    cg::grid_group grid = cg::this_grid();
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = a.T, words = (T + 1) >> 1, G = (int)gridDim.x;
    uint32_t* rows = bin_smem;                          // [W][words]
    uint32_t* base = bin_smem + (size_t)W * words;      // [T]
    uint32_t* row = rows + (size_t)warp * words;
    for (int i = threadIdx.x; i < W * words; i += blockDim.x) rows[i] = 0;
    // slice of the depth order owned by this warp
    const uint32_t nvis = min(*a.nvis, (uint32_t)a.P);
    const uint32_t nwarps = gridDim.x * W, gw = blockIdx.x * W + warp;
    const uint32_t per = (nvis + nwarps - 1) / nwarps;
    const uint32_t j_lo = min(gw * per, nvis), j_hi = min(j_lo + per, nvis);
    __syncthreads();

    // ---- phase 1: counting
    gms_bin_walk<false>(a, j_lo, j_hi, lane, row, nullptr);
    __syncthreads();
    // warps' rows -> exclusive offsets inside the CTA; CTA count per tile -> M
    {
        uint32_t* Mrow = a.M + (size_t)blockIdx.x * T;
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            uint32_t run = 0;   // two 16-bit lanes at once; a CTA's count of one tile stays below 65536 (host checks P / grid)
            for (int wq = 0; wq < W; wq++) { const uint32_t c = rows[(size_t)wq * words + i]; rows[(size_t)wq * words + i] = run; run += c; }
            if (2 * i + 1 < T) *reinterpret_cast<uint2*>(Mrow + 2 * i) = make_uint2(run & 0xffffu, run >> 16);
            else Mrow[2 * i] = run & 0xffffu;
        }
    }
    grid.sync();

    // ---- phase 2a: per tile, exclusive scan over the CTAs.  64 adjacent tiles per CTA (two per lane: coalesced rows of M);
    // the W warps split the CTA axis, combine their partial sums through shared memory, then write the offsets.
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
        for (int i = threadIdx.x; i < T; i += blockDim.x) base[i] = a.total[i];
        __syncthreads();
        const int per_t = (T + blockDim.x - 1) / blockDim.x;
        const int t0 = min((int)threadIdx.x * per_t, T), t1 = min(t0 + per_t, T);
        uint32_t s = 0;
        for (int t = t0; t < t1; t++) s += base[t];
        s_part[threadIdx.x] = s;
        __syncthreads();
        if (warp == 0) {        // exclusive scan of <= 512 partials by one warp
            uint32_t carry = 0;
            for (int c0 = 0; c0 < (int)blockDim.x; c0 += 32) {
                const uint32_t v = s_part[c0 + lane];
                uint32_t x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                s_part[c0 + lane] = carry + x - v;
                carry += __shfl_sync(0xffffffffu, x, 31);
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
            if (blockIdx.x == 0) a.ranges[t] = overflow ? make_int2(0, 0) : make_int2((int)run, (int)(run + tot));
            run += tot;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.n_out[0] = N; a.n_out[1] = overflow ? 1u : 0u;
            if (a.n_host) { a.n_host[0] = N; a.n_host[1] = overflow ? 1u : 0u; }
        }
        if (overflow) return;       // uniform over the grid: no further grid.sync follows
        __syncthreads();
        const uint32_t* Mrow = a.M + (size_t)blockIdx.x * T;
        for (int i = threadIdx.x; i < T; i += blockDim.x) base[i] += Mrow[i];
        __syncthreads();
    }

    // ---- phase 3: ordered placement
    gms_bin_walk<true>(a, j_lo, j_hi, lane, row, base);
}
