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

constexpr int THR_ROLL_MAXIW = 124;      // inner tiles per strip (32 lanes x 4 tiles minus 2 + 2 halo tiles)
struct RollPlan { int strips, iw, ysegs, seg_rows, rowb, warp_bytes; };   // rowb: ring row pitch (bytes), warp_bytes: ring + barriers per warp

// ---- TMA-staged streaming variant (default) ----------------------------------------------------------------------
// The CTA-tiled kernel above serialises load / exchange / dilate / store phases behind block barriers, so with 3 CTAs per
// SM the memory system idles during the compute phases (ncu: 27 % of the HBM roofline, long-scoreboard bound).  Here one
// WARP owns a strip of up to 124 tiles x a segment of tile rows and streams down it:
//   * the 4 even input rows of each tile row are prefetched four tile rows ahead by the TMA engine (1-D bulk copies,
//     cp.async.bulk ... mbarrier::complete_tx) into a per-warp shared-memory ring: 4 stages x 4 KB per warp, 12 warps per
//     SM = 144 KB of loads in flight per SM, none of them holding registers.  One elected lane issues the copies; the
//     warp waits on the stage's mbarrier phase;
//   * lane l reads its 32 bytes per row (two LDS.128), reduces 4 tiles with the native 16-bit SIMD min/max
//     (VIMNMX.U16x2 on the even bytes, which are exactly the decimated pixels) -- the byte-SIMD intrinsics are emulated
//     on sm_100 and cost 5-7 instructions each, so they are kept off the per-pixel path except for the final compare;
//   * the 3x3 tile dilation needs no shared memory: vertical neighbours are the lane's own packed min/max words of the
//     last three tile rows, horizontal neighbours come from lane-1 / lane+1 through shuffles and a funnel shift;
//   * no block barrier; the only redundancy is one halo tile row at each end of a segment.
constexpr int THR_TMA_WARPS = 4, THR_TMA_STAGES = 4, THR_TMA_ROWB = 1024;

// per-warp shared memory: THR_TMA_STAGES x 4 rows x plan.rowb bytes, then the stages' mbarriers.  The row pitch follows the
// strip width (768 B for the c2 frame instead of the 1 KB maximum), which is what decides how many warps fit an SM.
struct ThrTmaWarp {
    uint8_t *buf;
    unsigned long long *bar;
    int rowb;
    __device__ __forceinline__ uint8_t *row(int stage, int dy) const { return buf + (size_t)(stage * 4 + dy) * rowb; }
};
__host__ __device__ inline int thr_tma_warp_bytes(int rowb) { return THR_TMA_STAGES * 4 * rowb + 128; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// lane 0: start the copies of tile row rr (4 even input rows, `nbytes` bytes starting at byte x0s) into `stage`
__device__ __forceinline__ void thr_tma_issue(const ThrTmaWarp &W, int stage, const uint8_t *__restrict__ img, const Geom &g, int rr, int x0s, int nbytes)
{
    const int xs = max(x0s, 0), xe = min(x0s + nbytes, g.stride);
    const int n = xe - xs;
    uint32_t total = 0;
    if (rr >= 0 && n > 0)
        for (int dy = 0; dy < 4; dy++) if (rr * 4 + dy < g.h) total += (uint32_t)n;
    if (total == 0) { mbar_arrive(&W.bar[stage]); return; }
    mbar_expect_tx(&W.bar[stage], total);
    for (int dy = 0; dy < 4; dy++) {
        const int y = rr * 4 + dy;
        if (y < g.h) tma_load_1d(W.row(stage, dy) + (xs - x0s), img + (size_t)(2 * y) * g.stride + xs, (uint32_t)n, &W.bar[stage]);
    }
}

__global__ void __launch_bounds__(THR_TMA_WARPS * 32)
threshold_f2_tma_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint8_t *__restrict__ tmin, uint8_t *__restrict__ tmax,
                        Geom g, int min_diff, RollPlan plan, int write_tiles)
{
    extern __shared__ __align__(128) unsigned char thr_smem[];
    ThrTmaWarp W;
    W.buf = thr_smem + (size_t)(threadIdx.x >> 5) * plan.warp_bytes;
    W.bar = reinterpret_cast<unsigned long long *>(W.buf + THR_TMA_STAGES * 4 * plan.rowb);
    W.rowb = plan.rowb;
    const int lane = threadIdx.x & 31;
    const long long widx = (long long)blockIdx.x * THR_TMA_WARPS + (threadIdx.x >> 5);
    const long long per_frame = (long long)plan.strips * plan.ysegs;
    if (widx >= per_frame * g.batch) return;
    const int b = (int)(widx / per_frame);
    const int rem = (int)(widx % per_frame);
    const int seg = rem / plan.strips, strip = rem % plan.strips;
    const int s0 = strip * plan.iw;
    const int s1 = min(s0 + plan.iw, g.tw);
    const int y0 = seg * plan.seg_rows, y1 = min(y0 + plan.seg_rows, g.th);
    if (y0 >= y1 || s0 >= s1) return;
    const int gtx0 = s0 - 2 + 4 * lane;
    const bool lane_on = gtx0 < s1 + 1;
    const int x0s = (s0 - 2) * 8;                                   // byte offset of lane 0's data in an input row
    const int nlanes = (s1 + 1 - (s0 - 2) + 3) / 4;                 // lanes with lane_on
    const int nbytes = nlanes * 32;
    const uint8_t *img = in + (size_t)b * g.frame_stride;
    uint8_t *o = out + (size_t)b * g.h * g.tp;
    const uint32_t full = 0xffffffffu;

    if (lane == 0) {
        for (int s = 0; s < THR_TMA_STAGES; s++) mbar_init(&W.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const int rfirst = y0 - 1;
    if (lane == 0)
        for (int s = 0; s < THR_TMA_STAGES; s++)
            if (rfirst + s <= y1) thr_tma_issue(W, s, img, g, rfirst + s, x0s, nbytes);

    uint32_t px_prev[4][4], px_cur[4][4];
    uint32_t mn_a = 0xffffffffu, mn_b = 0xffffffffu, mx_a = 0, mx_b = 0;
#pragma unroll
    for (int dy = 0; dy < 4; dy++)
#pragma unroll
        for (int k = 0; k < 4; k++) { px_prev[dy][k] = 0; px_cur[dy][k] = 0; }
    // lane constants: which of the 4 tiles exist, which tile pairs are written, output column
    uint32_t valid_mn = 0, valid_mx = 0;          // byte k = 0x00 (valid) / 0xff (force neutral min) resp. 0xff keep / 0x00 force neutral max
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool v = gtx0 + k >= 0 && gtx0 + k < g.tw;
        if (!v) valid_mn |= 0xffu << (8 * k); else valid_mx |= 0xffu << (8 * k);
    }
    bool pair_in[2], pair_two[2];
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        const int gtx = gtx0 + 2 * pr;
        pair_in[pr] = lane_on && gtx >= s0 && gtx < s1;
        pair_two[pr] = gtx + 1 < s1;
    }
    uint8_t *ocol = o + (ptrdiff_t)gtx0 * 4;
    for (int r = rfirst; r <= y1; r++) {
        const int it = r - rfirst, stage = it % THR_TMA_STAGES;
        mbar_wait(&W.bar[stage], (uint32_t)((it / THR_TMA_STAGES) & 1));
        // ---- read this lane's 4 x 32 bytes; tile min/max on 16-bit lanes (even bytes = decimated pixels) ----
        uint32_t mnp[4], mxp[4];                  // per tile: running min / max as a U16x2 pair
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
#pragma unroll
            for (int k = 0; k < 4; k++) px_prev[dy][k] = px_cur[dy][k];
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
            if (lane_on) {
                const uint4 *p = reinterpret_cast<const uint4 *>(W.row(stage, dy) + lane * 32);
                v0 = p[0]; v1 = p[1];
            }
            const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t e0 = w[2 * k] & 0x00ff00ffu, e1 = w[2 * k + 1] & 0x00ff00ffu;
                const uint32_t lo = __vminu2(e0, e1), hi = __vmaxu2(e0, e1);
                mnp[k] = dy == 0 ? lo : __vminu2(mnp[k], lo);
                mxp[k] = dy == 0 ? hi : __vmaxu2(mxp[k], hi);
                px_cur[dy][k] = __byte_perm(w[2 * k], w[2 * k + 1], 0x6420);
            }
        }
        __syncwarp();
        if (lane == 0 && r + THR_TMA_STAGES <= y1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            thr_tma_issue(W, stage, img, g, r + THR_TMA_STAGES, x0s, nbytes);
        }
        uint32_t mn_c = 0xffffffffu, mx_c = 0;
        if (r >= 0 && r < g.th) {
            uint32_t mn = 0, mx = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t tmn = __vminu2(mnp[k], mnp[k] >> 16) & 0xffu, tmx = __vmaxu2(mxp[k], mxp[k] >> 16) & 0xffu;
                mn |= tmn << (8 * k); mx |= tmx << (8 * k);
            }
            mn_c = mn | valid_mn; mx_c = mx & valid_mx;
        }
        const int ro = r - 1;
        const uint32_t vmn = __vminu4(mn_a, __vminu4(mn_b, mn_c)), vmx = __vmaxu4(mx_a, __vmaxu4(mx_b, mx_c));
        uint32_t lmn = __shfl_up_sync(full, vmn, 1), rmn = __shfl_down_sync(full, vmn, 1);
        uint32_t lmx = __shfl_up_sync(full, vmx, 1), rmx = __shfl_down_sync(full, vmx, 1);
        if (lane == 0) { lmn = 0xffffffffu; lmx = 0; }
        if (lane == 31) { rmn = 0xffffffffu; rmx = 0; }
        const uint32_t dmn = __vminu4(vmn, __vminu4(__funnelshift_r(lmn, vmn, 24), __funnelshift_r(vmn, rmn, 8)));
        const uint32_t dmx = __vmaxu4(vmx, __vmaxu4(__funnelshift_r(lmx, vmx, 24), __funnelshift_r(vmx, rmx, 8)));
        if (ro >= y0 && ro < y1 && (pair_in[0] || pair_in[1])) {
            uint32_t wout[4][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t mn = (dmn >> (8 * k)) & 0xff, mx = (dmx >> (8 * k)) & 0xff;
                const bool flat = (int)mx - (int)mn < min_diff;
                const uint32_t th = (mn + (mx - mn) / 2) * 0x01010101u;
#pragma unroll
                for (int dy = 0; dy < 4; dy++) wout[dy][k] = flat ? 0x7f7f7f7fu : __vcmpgtu4(px_prev[dy][k], th);
            }
            uint8_t *orow = ocol + (size_t)(ro * 4) * g.tp;
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                if (!pair_in[pr]) continue;
#pragma unroll
                for (int dy = 0; dy < 4; dy++) {
                    uint8_t *dst = orow + (size_t)dy * g.tp + 8 * pr;
                    if (pair_two[pr]) *reinterpret_cast<uint2 *>(dst) = make_uint2(wout[dy][2 * pr], wout[dy][2 * pr + 1]);
                    else *reinterpret_cast<uint32_t *>(dst) = wout[dy][2 * pr];
                }
                if (write_tiles) {
                    const size_t ti = ((size_t)b * g.th + ro) * g.tw + gtx0 + 2 * pr;
                    tmin[ti] = (uint8_t)(mn_b >> (8 * (2 * pr))); tmax[ti] = (uint8_t)(mx_b >> (8 * (2 * pr)));
                    if (pair_two[pr]) { tmin[ti + 1] = (uint8_t)(mn_b >> (8 * (2 * pr + 1))); tmax[ti + 1] = (uint8_t)(mx_b >> (8 * (2 * pr + 1))); }
                }
            }
        }
        mn_a = mn_b; mn_b = mn_c; mx_a = mx_b; mx_b = mx_c;
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

// One launch converts a whole batch: blockIdx.y = frame (frame f at rgb + f * in_stride, gray + f * out_stride).
// A warp owns RGB_CHUNKS consecutive chunks of 512 pixels = 1536 B of packed RGB each.  Lane 0 starts one TMA bulk copy per
// chunk into the warp's shared memory (cp.async.bulk + mbarrier::complete_tx: no per-thread load or shared-store
// instructions, every DRAM sector fetched once), all chunks in flight at once; then, chunk by chunk, each lane reads back its
// own 16 pixels (48 B: three uint4 at a 12-word stride -- the eight lanes of a quarter-warp land on banks
// {0,12,24,4,16,28,8,20} + 0..3, i.e. conflict-free) and stores 16 gray bytes.  The first version staged the chunk with
// per-lane coalesced loads + shared stores and was bound by the shared-memory instruction queue (ncu: mio_throttle).
constexpr int PRE_THREADS = 256;          // yuyv_to_gray_kernel
constexpr int PRE_PX_PER_WARP = 512;
constexpr int RGB_THREADS = 128;
constexpr int RGB_CHUNKS = 4;
constexpr int RGB_PX_PER_BLOCK = (RGB_THREADS / 32) * RGB_CHUNKS * 512;
__global__ void __launch_bounds__(RGB_THREADS) rgb_to_gray_kernel(const uint8_t *__restrict__ rgb, uint8_t *__restrict__ gray, size_t npix,
                                                                   size_t in_stride, size_t out_stride)
{
    __shared__ __align__(128) uint4 stage[RGB_THREADS / 32][RGB_CHUNKS][96];
    __shared__ __align__(8) unsigned long long bar[RGB_THREADS / 32][RGB_CHUNKS];
    rgb += (size_t)blockIdx.y * in_stride;
    gray += (size_t)blockIdx.y * out_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t w0 = ((size_t)blockIdx.x * (RGB_THREADS / 32) + warp) * (RGB_CHUNKS * 512);      // first pixel of the warp
    if (w0 >= npix) return;                       // barriers are per warp: no block-wide synchronisation in this kernel
    const bool aligned = (((uintptr_t)rgb) & 15) == 0 && (((uintptr_t)gray) & 15) == 0;
    const int nfull = aligned ? (int)((npix - w0) / 512 < (size_t)RGB_CHUNKS ? (npix - w0) / 512 : (size_t)RGB_CHUNKS) : 0;
    if (lane == 0 && nfull > 0) {
        for (int c = 0; c < nfull; c++) mbar_init(&bar[warp][c], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int c = 0; c < nfull; c++) {
            mbar_expect_tx(&bar[warp][c], 1536);
            tma_load_1d(&stage[warp][c][0], rgb + 3 * (w0 + (size_t)c * 512), 1536, &bar[warp][c]);
        }
    }
    __syncwarp();
    for (int c = 0; c < nfull; c++) {
        mbar_wait(&bar[warp][c], 0);
        const uint4 a = stage[warp][c][3 * lane], b2 = stage[warp][c][3 * lane + 1], cc = stage[warp][c][3 * lane + 2];
        const uint32_t wds[12] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w, cc.x, cc.y, cc.z, cc.w};
        uint32_t o[4] = {0, 0, 0, 0};
        // u8 -> f32 and f32 -> u8 without the quarter-rate conversion unit (the I2F / F2I form capped this kernel at 4.2 TB/s:
        // four conversions per pixel): PRMT drops the byte into the mantissa of 2^23 and one FADD removes the bias (exact);
        // the truncating cast is FADD.RZ with 2^23, whose low mantissa byte is floor(v) (0 <= v <= 252.45, so no clamp).
        uint32_t q[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float r = __fsub_rn(__uint_as_float(__byte_perm(wds[(3 * k) >> 2], 0x4B000000u, 0x7440 | ((3 * k) & 3))), 8388608.0f);
            const float gg = __fsub_rn(__uint_as_float(__byte_perm(wds[(3 * k + 1) >> 2], 0x4B000000u, 0x7440 | ((3 * k + 1) & 3))), 8388608.0f);
            const float bl = __fsub_rn(__uint_as_float(__byte_perm(wds[(3 * k + 2) >> 2], 0x4B000000u, 0x7440 | ((3 * k + 2) & 3))), 8388608.0f);
            const float v = __fmaf_rn(r, 0.33f, __fmaf_rn(gg, 0.33f, __fmul_rn(bl, 0.33f)));          // utils.rs:43
            q[k] = __float_as_uint(__fadd_rz(v, 8388608.0f));
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            o[k] = __byte_perm(__byte_perm(q[4 * k], q[4 * k + 1], 0x4040), __byte_perm(q[4 * k + 2], q[4 * k + 3], 0x4040), 0x5410);
        *reinterpret_cast<uint4 *>(gray + w0 + (size_t)c * 512 + 16 * lane) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // ragged tail of a frame, or a frame whose planes are not 16-byte aligned
    const size_t rest = w0 + (size_t)nfull * 512;
    const size_t end = w0 + RGB_CHUNKS * 512 < npix ? w0 + RGB_CHUNKS * 512 : npix;
    for (size_t k = rest + lane; k < end; k += 32) gray[k] = cat_gray(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]);
}

// YUYV (Y0 U Y1 V): gray = Y.  A warp owns 512 pixels = 1024 B; lane t loads uint4 t and t + 32 (coalesced) and stores the
// eight Y bytes of each as one 64-bit word (coalesced).
__global__ void __launch_bounds__(PRE_THREADS) yuyv_to_gray_kernel(const uint8_t *__restrict__ yuyv, uint8_t *__restrict__ gray, size_t npix,
                                                                    size_t in_stride, size_t out_stride)
{
    yuyv += (size_t)blockIdx.y * in_stride;
    gray += (size_t)blockIdx.y * out_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t p0 = ((size_t)blockIdx.x * (PRE_THREADS / 32) + warp) * PRE_PX_PER_WARP;
    if (p0 >= npix) return;
    if (p0 + PRE_PX_PER_WARP <= npix && (((uintptr_t)yuyv) & 15) == 0 && (((uintptr_t)gray) & 7) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(yuyv + 2 * p0);
        const uint4 a = ldg_stream(src + lane), b2 = ldg_stream(src + lane + 32);
        uint2 *dst = reinterpret_cast<uint2 *>(gray + p0);
        dst[lane] = make_uint2(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420));
        dst[lane + 32] = make_uint2(__byte_perm(b2.x, b2.y, 0x6420), __byte_perm(b2.z, b2.w, 0x6420));
    } else {
        const size_t end = p0 + PRE_PX_PER_WARP < npix ? p0 + PRE_PX_PER_WARP : npix;
        for (size_t k = p0 + lane; k < end; k += 32) gray[k] = yuyv[2 * k];
    }
}

}  // namespace cb
