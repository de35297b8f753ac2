// threshold.cuh -- rows A1+A2 of SURVEY.md 8a: image_u8_decimate(2) fused with threshold().
//
// Upstream (AprilTag-3 apriltag_quad_thresh.c threshold(), reached from crates/apriltags/src/lib.rs:301):
//   4x4 tiles over the decimated image, per-tile min/max, 3x3 tile dilation (max) / erosion (min),
//   (max-min) < min_white_black_diff -> 127 else v > min+(max-min)/2 ? 255 : 0; partial tiles on the
//   right / bottom reuse the last full tile without the low-contrast test.
//
// B200 mapping: HBM-bound byte work.  One CTA stages a 64x32-tile window (512x256 input pixels, even rows
// only) in registers with 128-bit coalesced loads, reduces tile min/max with the byte-SIMD video
// instructions, exchanges them through 4 KB of shared memory, dilates in shared memory and writes the
// ternary map with 64-bit coalesced stores.  Algorithmic traffic: W*H/2 read + W*H/4 written.
#pragma once
#include "common.cuh"

namespace cb {

constexpr int THR_TX = 32, THR_TY = 8, THR_RPT = 4;
constexpr int THR_CW = 64, THR_CH = THR_TY * THR_RPT;   // tiles covered by one CTA (with halo)
constexpr int THR_IW = 60, THR_IH = 30;                 // tiles written by one CTA

__device__ __forceinline__ uint32_t hmin4(uint32_t m) { m = __vminu4(m, m >> 16); m = __vminu4(m, m >> 8); return m & 0xffu; }
__device__ __forceinline__ uint32_t hmax4(uint32_t m) { m = __vmaxu4(m, m >> 16); m = __vmaxu4(m, m >> 8); return m & 0xffu; }

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Fast path: f == 2, input rows 16-byte aligned.  Writes the ternary map for all FULL tiles and the raw
// per-tile min/max (tiny) for the remainder kernel.
__global__ void __launch_bounds__(THR_TX * THR_TY)
threshold_f2_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint8_t *__restrict__ tmin,
                    uint8_t *__restrict__ tmax, Geom g, int min_diff)
{
    __shared__ uint8_t smin[THR_CH][THR_CW];
    __shared__ uint8_t smax[THR_CH][THR_CW];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.z;
    const int gtx0 = blockIdx.x * THR_IW - 2 + 2 * tx;             // first tile of this thread's pair
    const int gty0 = blockIdx.y * THR_IH - 1 + ty * THR_RPT;       // first of its tile rows
    const uint8_t *img = in + (size_t)b * g.frame_stride;

    uint32_t pa[THR_RPT][4], pb[THR_RPT][4];   // decimated pixels: tile A / tile B, 4 rows each
    const int x0 = gtx0 * 8;                    // full-resolution byte offset of the pair
    const bool xvec = (x0 >= 0) && (x0 + 16 <= g.stride);
#pragma unroll
    for (int r = 0; r < THR_RPT; r++) {
        const int gty = gty0 + r;
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            const int y = gty * 4 + dy;        // decimated row
            uint32_t a = 0, bb = 0;
            if (gty >= 0 && y < g.h) {
                const uint8_t *row = img + (size_t)(2 * y) * g.stride;
                if (xvec) {
                    uint4 v = ldg_stream(reinterpret_cast<const uint4 *>(row + x0));
                    a = __byte_perm(v.x, v.y, 0x6420);
                    bb = __byte_perm(v.z, v.w, 0x6420);
                } else if (x0 + 16 > 0 && x0 < g.W) {   // row edge: guarded scalar loads
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        int xa = x0 + 2 * k, xb = x0 + 8 + 2 * k;
                        if (xa >= 0 && xa < g.W) a |= (uint32_t)row[xa] << (8 * k);
                        if (xb >= 0 && xb < g.W) bb |= (uint32_t)row[xb] << (8 * k);
                    }
                }
            }
            pa[r][dy] = a;
            pb[r][dy] = bb;
        }
    }
#pragma unroll
    for (int r = 0; r < THR_RPT; r++) {
        const int gty = gty0 + r;
        const bool yok = gty >= 0 && gty < g.th;
        uint32_t mnA = 255, mxA = 0, mnB = 255, mxB = 0;
        if (yok && gtx0 >= 0 && gtx0 < g.tw) {
            mnA = hmin4(__vminu4(__vminu4(pa[r][0], pa[r][1]), __vminu4(pa[r][2], pa[r][3])));
            mxA = hmax4(__vmaxu4(__vmaxu4(pa[r][0], pa[r][1]), __vmaxu4(pa[r][2], pa[r][3])));
        }
        if (yok && gtx0 + 1 >= 0 && gtx0 + 1 < g.tw) {
            mnB = hmin4(__vminu4(__vminu4(pb[r][0], pb[r][1]), __vminu4(pb[r][2], pb[r][3])));
            mxB = hmax4(__vmaxu4(__vmaxu4(pb[r][0], pb[r][1]), __vmaxu4(pb[r][2], pb[r][3])));
        }
        const int ly = ty * THR_RPT + r;
        *reinterpret_cast<uchar2 *>(&smin[ly][2 * tx]) = make_uchar2((uint8_t)mnA, (uint8_t)mnB);
        *reinterpret_cast<uchar2 *>(&smax[ly][2 * tx]) = make_uchar2((uint8_t)mxA, (uint8_t)mxB);
    }
    __syncthreads();

    const bool xin = (2 * tx >= 2) && (2 * tx < 2 + THR_IW);       // pair is inside the written window
    if (!xin) return;
    uint8_t *o = out + (size_t)b * g.h * g.tp;
#pragma unroll
    for (int r = 0; r < THR_RPT; r++) {
        const int ly = ty * THR_RPT + r;
        const int gty = gty0 + r;
        if (ly < 1 || ly >= 1 + THR_IH || gty >= g.th) continue;
        uint32_t wout[2][4];
        bool valid[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int gtx = gtx0 + t, lx = 2 * tx + t;
            valid[t] = gtx < g.tw;
            uint32_t mn = 255, mx = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                int yy = ly + dy;
                if (yy < 0 || yy >= THR_CH) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    int xx = lx + dx;
                    if (xx < 0 || xx >= THR_CW) continue;
                    mn = min(mn, (uint32_t)smin[yy][xx]);
                    mx = max(mx, (uint32_t)smax[yy][xx]);
                }
            }
            if (valid[t]) {   // raw min/max of this tile for the remainder kernel
                tmin[((size_t)b * g.th + gty) * g.tw + gtx] = smin[ly][lx];
                tmax[((size_t)b * g.th + gty) * g.tw + gtx] = smax[ly][lx];
            }
            const uint32_t *px = t == 0 ? pa[r] : pb[r];
            if ((int)mx - (int)mn < min_diff) {
#pragma unroll
                for (int dy = 0; dy < 4; dy++) wout[t][dy] = 0x7f7f7f7fu;
            } else {
                uint32_t th = mn + (mx - mn) / 2;
                th = th * 0x01010101u;
#pragma unroll
                for (int dy = 0; dy < 4; dy++) wout[t][dy] = __vcmpgtu4(px[dy], th);
            }
        }
        if (!valid[0]) continue;
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            uint8_t *dst = o + (size_t)(gty * 4 + dy) * g.tp + gtx0 * 4;
            if (valid[1]) *reinterpret_cast<uint2 *>(dst) = make_uint2(wout[0][dy], wout[1][dy]);
            else *reinterpret_cast<uint32_t *>(dst) = wout[0][dy];
        }
    }
}

// ---- generic path (any integer decimation factor, any alignment): three simple kernels ----
__global__ void tile_minmax_generic_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ tmin, uint8_t *__restrict__ tmax, Geom g)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (t >= g.tw * g.th) return;
    int tx = t % g.tw, ty = t / g.tw;
    const uint8_t *img = in + (size_t)b * g.frame_stride;
    uint32_t mn = 255, mx = 0;
    for (int dy = 0; dy < 4; dy++)
        for (int dx = 0; dx < 4; dx++) {
            uint32_t v = img[(size_t)((ty * 4 + dy) * g.f) * g.stride + (tx * 4 + dx) * g.f];
            mn = min(mn, v); mx = max(mx, v);
        }
    tmin[(size_t)b * g.tw * g.th + t] = (uint8_t)mn;
    tmax[(size_t)b * g.tw * g.th + t] = (uint8_t)mx;
}

__device__ __forceinline__ void dilated_minmax(const uint8_t *tmin, const uint8_t *tmax, int tw, int th, int tx, int ty, int &mn, int &mx)
{
    mn = 255; mx = 0;
    for (int dy = -1; dy <= 1; dy++) {
        int yy = ty + dy;
        if (yy < 0 || yy >= th) continue;
        for (int dx = -1; dx <= 1; dx++) {
            int xx = tx + dx;
            if (xx < 0 || xx >= tw) continue;
            mn = min(mn, (int)tmin[yy * tw + xx]);
            mx = max(mx, (int)tmax[yy * tw + xx]);
        }
    }
}

// every pixel (full tiles and remainder); only_remainder restricts it to the right / bottom strips
__global__ void threshold_generic_kernel(const uint8_t *__restrict__ in, const uint8_t *__restrict__ tmin, const uint8_t *__restrict__ tmax,
                                         uint8_t *__restrict__ out, Geom g, int min_diff, int only_remainder)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int b = blockIdx.z;
    if (x >= g.w || y >= g.h) return;
    const bool full = x < g.tw * 4 && y < g.th * 4;
    if (only_remainder && full) return;
    uint8_t *o = out + (size_t)b * g.h * g.tp + (size_t)y * g.tp + x;
    if (g.tw == 0 || g.th == 0) { *o = 127; return; }
    int tx = min(x / 4, g.tw - 1), ty = min(y / 4, g.th - 1);
    int mn, mx;
    dilated_minmax(tmin + (size_t)b * g.tw * g.th, tmax + (size_t)b * g.tw * g.th, g.tw, g.th, tx, ty, mn, mx);
    if (full && mx - mn < min_diff) { *o = 127; return; }
    int thresh = mn + (mx - mn) / 2;
    int v = in[(size_t)b * g.frame_stride + (size_t)(y * g.f) * g.stride + x * g.f];
    *o = v > thresh ? 255 : 0;
}

// ---- pre-processing (row P1 and the "next" camera formats): produce full-resolution gray ----
// CAT grayscale (crates/chalkydri-apriltags/src/utils.rs:43): two fused multiply-adds in f32, truncating cast.
__device__ __forceinline__ uint8_t cat_gray(uint32_t r, uint32_t gch, uint32_t bch)
{
    float v = __fmaf_rn((float)r, 0.33f, __fmaf_rn((float)gch, 0.33f, __fmul_rn((float)bch, 0.33f)));
    return (uint8_t)min(255, max(0, (int)v));
}

// 16 pixels per thread: three 128-bit loads of packed RGB, one 128-bit store of gray
__global__ void rgb_to_gray_kernel(const uint8_t *__restrict__ rgb, uint8_t *__restrict__ gray, size_t npix)
{
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= npix) return;
    if (i + 16 <= npix && (((uintptr_t)rgb) & 15) == 0 && (((uintptr_t)gray) & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(rgb + 3 * i);
        uint4 a = ldg_stream(src), b2 = ldg_stream(src + 1), c = ldg_stream(src + 2);
        uint32_t wds[12] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w, c.x, c.y, c.z, c.w};
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t r = (wds[(3 * k) >> 2] >> (8 * ((3 * k) & 3))) & 0xff;
            uint32_t gg = (wds[(3 * k + 1) >> 2] >> (8 * ((3 * k + 1) & 3))) & 0xff;
            uint32_t bl = (wds[(3 * k + 2) >> 2] >> (8 * ((3 * k + 2) & 3))) & 0xff;
            o[k >> 2] |= (uint32_t)cat_gray(r, gg, bl) << (8 * (k & 3));
        }
        *reinterpret_cast<uint4 *>(gray + i) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        for (size_t k = i; k < npix && k < i + 16; k++) gray[k] = cat_gray(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]);
    }
}

// YUYV (Y0 U Y1 V): gray = Y.  32 bytes in, 16 out per thread.
__global__ void yuyv_to_gray_kernel(const uint8_t *__restrict__ yuyv, uint8_t *__restrict__ gray, size_t npix)
{
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= npix) return;
    if (i + 16 <= npix && (((uintptr_t)yuyv) & 15) == 0 && (((uintptr_t)gray) & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(yuyv + 2 * i);
        uint4 a = ldg_stream(src), b2 = ldg_stream(src + 1);
        *reinterpret_cast<uint4 *>(gray + i) = make_uint4(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420),
                                                          __byte_perm(b2.x, b2.y, 0x6420), __byte_perm(b2.z, b2.w, 0x6420));
    } else {
        for (size_t k = i; k < npix && k < i + 16; k++) gray[k] = yuyv[2 * k];
    }
}

}  // namespace cb
