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
//            it (16-bit counters in the warp's own shared-memory row, two tiles per 32-bit word; order does not matter
//            here, so the lanes walk their own rectangles in parallel);  a CTA-level exclusive scan over the warps'
//            rows turns them into warp offsets and yields the CTA's per-tile count  M[cta][tile]          -> grid.sync
//   phase 2a one thread per tile scans M[.][tile] over the CTAs (exclusive, in place) and writes total[tile]  -> grid.sync
//   phase 2b every CTA scans total[] (T <= 16 K entries, block scan) into tile starts; base[tile] = start + M[cta][tile]
//            lives in shared memory; CTA 0 writes ranges[] and N.  N > capacity: overflow flag, empty ranges, no writes.
//   phase 3  every warp walks its slice again IN ORDER, one Gaussian per step with the lanes over its tiles, and writes
//            point_list[base[tile] + row[warp][tile]++] = Gaussian id.
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

__device__ __forceinline__ void gms_bin_unpack(const uint2 r, int& x0, int& y0, int& w, int& nt) {
    x0 = (int)(r.x & 0xffffu); y0 = (int)(r.x >> 16);
    w = (int)(r.y & 0xffffu) - x0;
    nt = w * ((int)(r.y >> 16) - y0);
}

__global__ void __launch_bounds__(512, 1) k_bin_tiles(GmsBinArgs a) {
    extern __shared__ uint32_t bin_smem[];
    cg::grid_group grid = cg::this_grid();
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = a.T, words = (T + 1) >> 1;
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

    // ---- phase 1: unordered counting
    for (uint32_t j0 = j_lo; j0 < j_hi; j0 += 32) {
        const uint32_t j = j0 + lane;
        int x0 = 0, y0 = 0, w = 0, nt = 0;
        if (j < j_hi) gms_bin_unpack(a.rect[a.order[j]], x0, y0, w, nt);
        const bool big = nt >= 64;
        if (!big) {
            int t = y0 * a.gx + x0, x = 0;
            for (int k = 0; k < nt; k++) {
                atomicAdd(&row[t >> 1], 1u << ((t & 1) * 16));
                if (++x == w) { x = 0; t += a.gx - w + 1; } else t++;
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, big);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int x0s = __shfl_sync(0xffffffffu, x0, src), y0s = __shfl_sync(0xffffffffu, y0, src);
            const int ws = __shfl_sync(0xffffffffu, w, src), nts = __shfl_sync(0xffffffffu, nt, src);
            for (int k = lane; k < nts; k += 32) {
                const int yy = k / ws, t = (y0s + yy) * a.gx + x0s + (k - yy * ws);
                atomicAdd(&row[t >> 1], 1u << ((t & 1) * 16));
            }
        }
    }
    __syncthreads();
    // warps' rows -> exclusive offsets inside the CTA; CTA count per tile -> M
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
        uint32_t run = 0;       // two 16-bit lanes at once; a CTA's count of one tile stays below 65536 (host checks per * W)
        for (int wq = 0; wq < W; wq++) { const uint32_t c = rows[(size_t)wq * words + i]; rows[(size_t)wq * words + i] = run; run += c; }
        uint32_t* Mrow = a.M + (size_t)blockIdx.x * T;
        Mrow[2 * i] = run & 0xffffu;
        if (2 * i + 1 < T) Mrow[2 * i + 1] = run >> 16;
    }
    grid.sync();

    // ---- phase 2a: per tile, exclusive scan over the CTAs.  A chunk of 32 adjacent tiles per CTA round (coalesced rows of
    // M); the W warps split the CTA axis, combine their partial sums through shared memory, then write the offsets.
    {
        __shared__ uint32_t s_grp[16][32];
        const int G = (int)gridDim.x, cper = (G + W - 1) / W;
        const int c_lo = min(warp * cper, G), c_hi = min(c_lo + cper, G);
        for (int chunk = blockIdx.x; chunk * 32 < T; chunk += G) {
            const int t = chunk * 32 + lane;
            const bool ok = t < T;
            uint32_t sum = 0;
            for (int c = c_lo; c < c_hi; c += 8) {
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = (ok && c + u < c_hi) ? a.M[(size_t)(c + u) * T + t] : 0u;
#pragma unroll
                for (int u = 0; u < 8; u++) sum += v[u];
            }
            s_grp[warp][lane] = sum;
            __syncthreads();
            uint32_t run = 0, tot = 0;
            for (int q = 0; q < W; q++) { const uint32_t x = s_grp[q][lane]; if (q < warp) run += x; tot += x; }
            for (int c = c_lo; c < c_hi; c += 8) {
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = (ok && c + u < c_hi) ? a.M[(size_t)(c + u) * T + t] : 0u;
#pragma unroll
                for (int u = 0; u < 8; u++) { if (ok && c + u < c_hi) a.M[(size_t)(c + u) * T + t] = run; run += v[u]; }
            }
            if (ok && warp == 0) a.total[t] = tot;
            __syncthreads();
        }
    }
    grid.sync();

    // ---- phase 2b: tile starts (every CTA, redundantly), base[] = start + this CTA's offset
    __shared__ uint32_t s_part[512];
    __shared__ uint32_t s_n;
    {
        const int per_t = (T + blockDim.x - 1) / blockDim.x;
        const int t0 = threadIdx.x * per_t, t1 = min(t0 + per_t, T);
        uint32_t s = 0;
        for (int t = t0; t < t1; t++) s += a.total[t];
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
        const uint32_t* Mrow = a.M + (size_t)blockIdx.x * T;
        for (int t = t0; t < t1; t++) {
            const uint32_t tot = a.total[t];
            base[t] = run + Mrow[t];
            if (blockIdx.x == 0) a.ranges[t] = overflow ? make_int2(0, 0) : make_int2((int)run, (int)(run + tot));
            run += tot;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.n_out[0] = N; a.n_out[1] = overflow ? 1u : 0u;
            if (a.n_host) { a.n_host[0] = N; a.n_host[1] = overflow ? 1u : 0u; }
        }
        __syncthreads();
        if (overflow) return;       // uniform over the grid: no further grid.sync follows
    }

    // ---- phase 3: ordered placement (one Gaussian per step, lanes over its tiles)
    for (uint32_t j0 = j_lo; j0 < j_hi; j0 += 32) {
        const uint32_t j = j0 + lane;
        int x0 = 0, y0 = 0, w = 0, nt = 0;
        uint32_t g = 0;
        if (j < j_hi) { g = a.order[j]; gms_bin_unpack(a.rect[g], x0, y0, w, nt); }
        const int t00 = y0 * a.gx + x0;
        uint32_t m = __ballot_sync(0xffffffffu, nt > 0);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t gs = __shfl_sync(0xffffffffu, g, src);
            const int t0s = __shfl_sync(0xffffffffu, t00, src);
            const int ws = __shfl_sync(0xffffffffu, w, src), nts = __shfl_sync(0xffffffffu, nt, src);
            const float inv = __frcp_rn((float)ws);
            for (int k = lane; k < nts; k += 32) {
                const int yy = (int)(((float)k + 0.5f) * inv);          // exact floor(k / ws) for k < 2^20
                const int t = t0s + yy * a.gx + (k - yy * ws);
                const int sh = (t & 1) * 16;
                const uint32_t old = atomicAdd(&row[t >> 1], 1u << sh);
                const uint32_t pos = base[t] + ((old >> sh) & 0xffffu);
                a.point_list[pos] = gs;
                if (a.tile_keys) a.tile_keys[pos] = (uint32_t)t;
            }
            __syncwarp();
        }
    }
}
