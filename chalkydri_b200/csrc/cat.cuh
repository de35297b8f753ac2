// cat.cuh -- rows P1, T1-T5 of SURVEY.md 8a: the in-house "CAT" detector stages (crates/chalkydri-apriltags).
//
//   T1 calc_otsu   (lib.rs:191-259): 5x5 clamped-window order statistics (statrs 0.18 R-8 quantiles) -> Color map
//   T2 thresh      (lib.rs:319-334): fixed ternary threshold
//   T3 detect_corners (lib.rs:291-309,345-400): FAST-like predicate, output in the reference's x-major scan order
//   T4 check_edges (lib.rs:409-499): all ordered corner pairs, second iterator reversed
//   T5 connected_components (lib.rs:501-549): shares the union-find kernels of ccl.cuh (MODE 1)
// Colors: 0 Black, 1 White, 2 Other (utils.rs:1-6).  Gray conversion is cat_gray() of threshold.cuh (utils.rs:43).
#pragma once
#include "common.cuh"
#include "threshold.cuh"

namespace cb {

__device__ __forceinline__ uint8_t f64_as_u8(double v)
{
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

__device__ __forceinline__ double statrs_quantile(const uint8_t *sorted, int n, double tau)
{
    const double h = ((double)n + 1.0 / 3.0) * tau + 1.0 / 3.0;
    const long long hf = (long long)h;
    if (hf <= 0 || tau == 0.0) return (double)sorted[0];
    if (hf >= n) return (double)sorted[n - 1];
    const double a = (double)sorted[hf - 1], b = (double)sorted[hf];
    return a + (h - (double)hf) * (b - a);
}

// one thread per pixel; gray plane precomputed (w*h u8)
__global__ void __launch_bounds__(128)
cat_otsu_kernel(const uint8_t *__restrict__ gray, uint8_t *__restrict__ color, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const int x_min = x >= 2 ? x - 2 : 0, x_max = min(x + 2, w - 1);
    const int y_min = y >= 2 ? y - 2 : 0, y_max = min(y + 2, h - 1);
    uint8_t px[25];
    int n = 0;
    for (int yy = y_min; yy <= y_max; yy++)
        for (int xx = x_min; xx <= x_max; xx++) {
            const uint8_t v = gray[(size_t)yy * w + xx];
            int j = n++;
            while (j > 0 && px[j - 1] > v) { px[j] = px[j - 1]; j--; }   // insertion sort
            px[j] = v;
        }
    const uint8_t p = gray[(size_t)y * w + x];
    uint8_t out;
    if ((y > 0 && x > 0) && ((double)px[n - 1] - (double)px[0]) < 5.0) {
        const int k = n / 2;
        const double med = (n % 2 != 0) ? (double)px[k] : ((double)px[k - 1] + (double)px[k]) / 2.0;
        out = med < 60.0 ? 0 : (med > 160.0 ? 1 : 2);
    } else {
        if (p >= f64_as_u8(statrs_quantile(px, n, 0.75))) out = 1;
        else if (p <= f64_as_u8(statrs_quantile(px, n, 0.25))) out = 0;
        else out = 2;
    }
    color[(size_t)y * w + x] = out;
}

__global__ void cat_thresh_kernel(const uint8_t *__restrict__ gray, uint8_t *__restrict__ color, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t g = gray[i];
    color[i] = g < 60 ? 0 : (g > 160 ? 1 : 2);
}

// ---- ordered compaction: flags[k] (0..2 outputs per item) -> stable output positions ----
// pass 1: per-block totals; pass 2 (single block): exclusive scan of block totals; pass 3: write.
constexpr int OC_BLOCK = 256;

__device__ __forceinline__ int cat_corner_flag(const uint8_t *__restrict__ c, int w, int h, int x, int y)
{
    // x in [3, w-3], y in [3, h-3] (inclusive, lib.rs:293-294).  px() is an unchecked linear index (utils.rs:27-29): at x = w-3
    // the (x+3, .) samples are column 0 of the next row, which `at` reproduces; pixels whose furthest sample lies beyond the
    // w*h buffer (undefined behaviour in the reference) are skipped
    if ((size_t)(y + 3) * w + (size_t)(x + 3) >= (size_t)w * h) return 0;
    auto at = [&](int xx, int yy) { return c[(size_t)yy * w + xx]; };
    if (at(x, y) != 0) return 0;
    const bool ul = at(x - 1, y - 1) == 0, ur = at(x + 1, y - 1) == 0, dl = at(x - 1, y + 1) == 0, dr = at(x + 1, y + 1) == 0;
    if (!(ul ^ ur ^ dl ^ dr)) return 0;
    const uint8_t p3 = at(x + 3, y - 3), p7 = at(x + 3, y + 3), p11 = at(x - 3, y + 3), p15 = at(x - 3, y - 3);
    if ((p3 != 2 && p7 != 2 && p11 != 2 && p15 != 2) && ((p3 == 0) ^ (p7 == 0) ^ (p11 == 0) ^ (p15 == 0))) return 1;
    return 0;
}

// item k (x-major): x = 3 + k / ny, y = 3 + k % ny with ny = h - 5
template <int PASS>
__global__ void __launch_bounds__(OC_BLOCK)
cat_corners_kernel(const uint8_t *__restrict__ color, int w, int h, long long nitems, uint32_t *__restrict__ block_tot,
                   int32_t *__restrict__ xy, long long cap)
{
    const long long k = (long long)blockIdx.x * OC_BLOCK + threadIdx.x;
    const int ny = h - 5;
    int flag = 0, x = 0, y = 0;
    if (k < nitems) { x = 3 + (int)(k / ny); y = 3 + (int)(k % ny); flag = cat_corner_flag(color, w, h, x, y); }
    __shared__ uint32_t wsum[OC_BLOCK / 32];
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsum[wid] = __popc(bal);
    __syncthreads();
    if (PASS == 0) {
        if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < OC_BLOCK / 32; i++) t += wsum[i]; block_tot[blockIdx.x] = t; }
    } else {
        long long pos = block_tot[blockIdx.x];   // exclusive prefix after the scan pass
        for (int i = 0; i < wid; i++) pos += wsum[i];
        pos += __popc(bal & ((1u << lane) - 1));
        if (flag && pos < cap) { xy[2 * pos] = x; xy[2 * pos + 1] = y; }
    }
}

// single block: in-place exclusive scan of block totals, total written to *total
__global__ void __launch_bounds__(1024)
oc_scan_kernel(uint32_t *__restrict__ tot, long long nblocks, unsigned long long *__restrict__ total)
{
    __shared__ unsigned long long carry;
    __shared__ uint32_t wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? tot[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        unsigned long long pre = carry;
        for (int k = 0; k < wid; k++) pre += wsum[k];
        if (i < nblocks) tot[i] = (uint32_t)(pre + incl - v);
        __syncthreads();
        if (threadIdx.x == 1023) carry = pre + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__device__ __forceinline__ uint8_t cat_at(const uint8_t *__restrict__ c, int w, int h, long long x, long long y)
{
    return (x < 0 || y < 0 || x >= w || y >= h) ? (uint8_t)2 : c[(size_t)y * w + x];
}

// pair item k: i = k / P, j = P - 1 - k % P (second iterator reversed); up to two lines (vertical test, horizontal test)
__device__ __forceinline__ int cat_edge_flags(const uint8_t *__restrict__ c, int w, int h, int x1, int y1, int x2, int y2)
{
    const int OFF = 5;
    const int mx = (x1 + x2) / 2, my = (y1 + y2) / 2;
    const int xdiff = max(x1, x2) - min(x1, x2), ydiff = max(y1, y2) - min(y1, y2);
    const bool vert = x1 == x2 || xdiff < ydiff, horiz = y1 == y2 || ydiff < xdiff;
    const int mw1x = (mx + x1) / 2, mw1y = (my + y1) / 2, mw2x = (mx + x2) / 2, mw2y = (my + y2) / 2;
    int f = 0;
    if (vert) {
        const uint8_t r1 = cat_at(c, w, h, mw1x + OFF, mw1y), r2 = cat_at(c, w, h, mw2x + OFF, mw2y);
        const uint8_t l1 = cat_at(c, w, h, mw1x - OFF, mw1y), l2 = cat_at(c, w, h, mw2x - OFF, mw2y);
        if (l1 != 2 && l2 != 2 && r1 != 2 && r2 != 2)
            if (((l1 == 0) ^ (r2 == 0)) && ((l2 == 0) ^ (r1 == 0)) && (l1 == l2)) f |= 1;
    }
    if (horiz) {
        const uint8_t t1 = cat_at(c, w, h, mw1x, mw1y - OFF), t2 = cat_at(c, w, h, mw2x, mw2y - OFF);
        const uint8_t b1 = cat_at(c, w, h, mw1x, mw1y + OFF), b2 = cat_at(c, w, h, mw2x, mw2y + OFF);
        if (t1 != 2 && t2 != 2 && b1 != 2 && b2 != 2)
            if (((t1 == 0) ^ (b2 == 0)) && ((t2 == 0) ^ (b1 == 0)) && (t1 == t2)) f |= 2;
    }
    return f;
}

template <int PASS>
__global__ void __launch_bounds__(OC_BLOCK)
cat_edges_kernel(const uint8_t *__restrict__ color, int w, int h, const int32_t *__restrict__ xy, long long P, uint32_t *__restrict__ block_tot,
                 int32_t *__restrict__ lines, long long cap)
{
    const long long k = (long long)blockIdx.x * OC_BLOCK + threadIdx.x;
    int f = 0, x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    if (k < P * P) {
        const long long i = k / P, j = P - 1 - (k % P);
        x1 = xy[2 * i]; y1 = xy[2 * i + 1]; x2 = xy[2 * j]; y2 = xy[2 * j + 1];
        f = cat_edge_flags(color, w, h, x1, y1, x2, y2);
    }
    const uint32_t cnt = (f & 1) + ((f >> 1) & 1);
    __shared__ uint32_t wsum[OC_BLOCK / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (PASS == 0) {
        if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < OC_BLOCK / 32; i++) t += wsum[i]; block_tot[blockIdx.x] = t; }
    } else {
        long long pos = block_tot[blockIdx.x];
        for (int i = 0; i < wid; i++) pos += wsum[i];
        pos += incl - cnt;
        for (uint32_t r = 0; r < cnt; r++, pos++)
            if (pos < cap) { lines[4 * pos] = x1; lines[4 * pos + 1] = y1; lines[4 * pos + 2] = x2; lines[4 * pos + 3] = y2; }
    }
}

// per-pixel size lookup for the CAT connected-components tap
__global__ void cat_sizes_kernel(const uint32_t *__restrict__ labels, const uint32_t *__restrict__ sizes, uint32_t *__restrict__ out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sizes[labels[i]];
}

// Colour map (0 Black, 1 White, 2 Other; pitch = width) -> upstream's ternary map (0 / 255 / 127; pitch tp) for the decode path
__global__ void cat_color_to_map_kernel(const uint8_t *__restrict__ color, uint8_t *__restrict__ map, int w, int h, int tp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t c = color[(size_t)y * w + x];
    map[(size_t)y * tp + x] = c == 0 ? 0 : (c == 1 ? 255 : 127);
}

}  // namespace cb
