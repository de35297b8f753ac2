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
#include <cuda.h>      // CUtensorMap (the type only: cuTensorMapEncodeTiled is looked up at run time, libcuda is not linked)

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

// ---- tensor-map variant (default since round 2) ---------------------------------------------------------------------
// Same warp-per-strip streaming structure as above, rebuilt around what ncu said about it (issue bound at 58 %, a third of
// the lanes idle at 1280 px, 100+ instructions per step spent on issuing four 1-D bulk copies from an elected lane):
//   * ONE tensor-map TMA per tile row.  The frames are described as a 3-D tensor of 8-byte elements (one element = the 8
//     input bytes of one tile row) with the row stride DOUBLED, so the map addresses the even input rows only; a box of
//     {32 T tiles, 4 rows, 1 frame} is one cp.async.bulk.tensor (SASS UTMALDG), out-of-range tiles / rows are zero-filled
//     by the engine (no edge cases in the issue code);
//   * T tiles per lane, T = 6 or 4 chosen per frame width so that one warp spans the frame (1280 px: 27 of 32 lanes busy with
//     T = 6 against 21 with two strips of T = 4);
//   * every reduction on native 16-bit SIMD (VIMNMX3.U16x2): tile min/max, the vertical and -- after a PRMT that shifts
//     the pairs by one tile -- the horizontal 3-tap of the dilation; the binarisation is an add that carries "px > thresh"
//     into bit 15 of each 16-bit lane and one PRMT in sign-replicate mode that turns four such bits into four 0x00 / 0xff
//     bytes (the byte-SIMD compare is emulated on sm_100);
//   * the pixel registers of the previous tile row ping-pong (loop unrolled twice) instead of being moved;
//   * neighbouring row segments run in opposite directions, so the halo rows two warps share are read at the same time
//     (second read hits L2);
//   * the STORES decide the rest (tools/cuda/thr_bench.cu: with the stores removed the kernel runs at the speed of its loads,
//     27 us per 256 x 1280x720 frames; with the first version's stores -- pair by pair, each pair's four rows -- 45 us, and
//     removing all the binarisation arithmetic changed nothing).  A lane owns 24 bytes of every output row, so a 32-byte
//     sector is completed by several 8-byte stores; issued pair-major those pieces reached the L2 far apart.  Two forms now:
//     BULK = 1 stages the four output rows of a tile row in shared memory (conflict-free 8-byte stores at a 24-byte lane
//     stride) and one lane writes each row with a TMA bulk store (cp.async.bulk.global.shared::cta, full lines; two staging
//     buffers, wait_group.read before reuse); BULK = 0 stores directly but row by row.  One warp per CTA.
struct TmPlan { int strips, iw, ysegs, seg_rows; };      // iw: interior tiles per strip (multiple of 4)
// T tiles per lane, S ring stages per warp, NW warps per CTA, MINB CTAs per SM the register allocation is held to
template <int T, int S = 3, int NW = 1, int MINB = 9, int BULK_STORE = 1> struct TmCfg {
    static constexpr int P = T / 2, ROWB = 32 * T * 8, STAGEB = 4 * ROWB, STAGES = S, WARPS = NW, MIN_CTAS = MINB;
    static constexpr int BULK = BULK_STORE;     // 1: output rows staged in shared memory and written by TMA bulk stores (full lines); 0: direct stores
    static constexpr int ROWOUT = 32 * T * 4, OUTB = BULK ? 2 * 4 * ROWOUT : 0;      // two staging buffers of 4 output rows per warp
    static constexpr int SMEM = NW * S * STAGEB + NW * OUTB + NW * S * 8;
    static constexpr int MAX_IW = 32 * T - 4;
};

__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));     // selector bit 3 of a nibble: replicate the byte's sign
    return d;
}

template <class C>
__global__ void __launch_bounds__(C::WARPS * 32, C::MIN_CTAS)
threshold_tm_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ out, uint8_t *__restrict__ tmin, uint8_t *__restrict__ tmax,
                    Geom g, int min_diff, TmPlan plan, int write_tiles)
{
    constexpr int P = C::P, T = 2 * C::P, TM_STAGES = C::STAGES, TM_WARPS = C::WARPS;
    extern __shared__ __align__(128) unsigned char tm_smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = tm_smem + (size_t)wid * (TM_STAGES * C::STAGEB);
    unsigned char *obuf = tm_smem + (size_t)TM_WARPS * TM_STAGES * C::STAGEB + (size_t)wid * C::OUTB;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(tm_smem + (size_t)TM_WARPS * (TM_STAGES * C::STAGEB + C::OUTB)) + wid * TM_STAGES;
    const long long widx = (long long)blockIdx.x * TM_WARPS + wid;
    const long long per_frame = (long long)plan.strips * plan.ysegs;
    if (widx >= per_frame * g.batch) return;                 // (no block-wide barrier in this kernel)
    const int b = (int)(widx / per_frame);
    const int rem = (int)(widx % per_frame);
    const int seg = rem / plan.strips, strip = rem % plan.strips;
    const int s0 = strip * plan.iw, s1 = min(s0 + plan.iw, g.tw);
    const int y0 = seg * plan.seg_rows, y1 = min(y0 + plan.seg_rows, g.th);
    if (y0 >= y1 || s0 >= s1) return;
    const int tbase = s0 - 2;                                 // first tile of lane 0 (even: outputs are stored as 8-byte tile pairs)
    const int t0 = tbase + T * lane;                          // first tile of this lane
    const int dir = (seg & 1) ? -1 : 1;                       // neighbouring segments meet at their shared halo rows at the same time
    const int nsteps = y1 - y0 + 2;
    const int rstart = dir > 0 ? y0 - 1 : y1;
    const uint32_t full = 0xffffffffu;
    uint8_t *o = out + (size_t)b * g.h * g.tp;

    if (lane == 0) {
        for (int s = 0; s < TM_STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < TM_STAGES && s < nsteps; s++) {
            mbar_expect_tx(&bars[s], (uint32_t)C::STAGEB);
            tma_load_3d(ring + (size_t)s * C::STAGEB, &tmap, tbase, 4 * (rstart + dir * s), b, &bars[s]);
        }
    }
    __syncwarp();

    // lane constants: validity of the lane's tile pairs (16-bit lane = one tile), which pairs are written
    uint32_t mn_or[P], mx_and[P];
    int pair_st[P];                                          // 0: nothing of the pair is written, 1: its first tile, 2: both
    bool any_in = false;
#pragma unroll
    for (int j = 0; j < P; j++) {
        const int ta = t0 + 2 * j, tb = ta + 1;
        const bool va = ta >= 0 && ta < g.tw, vb = tb >= 0 && tb < g.tw;
        mn_or[j] = (va ? 0u : 0x000000ffu) | (vb ? 0u : 0x00ff0000u);
        mx_and[j] = (va ? 0x0000ffffu : 0u) | (vb ? 0xffff0000u : 0u);
        pair_st[j] = (ta >= s0 && ta < s1) ? (tb < s1 ? 2 : 1) : 0;
        any_in |= pair_st[j] != 0;
    }
    const uint32_t fconst = (uint32_t)((0x8000 - min_diff) & 0xffff) * 0x00010001u;     // D + fconst has bit 15 set iff D >= min_diff
    uint32_t pxa[4][2 * T], pxb[4][2 * T];          // per tile and row: pixels (0, 1) and (2, 3) in 16-bit lanes, as the compare wants them
    uint32_t mnA[P], mnB[P], mxA[P], mxB[P];
#pragma unroll
    for (int j = 0; j < P; j++) { mnA[j] = mnB[j] = 0x00ff00ffu; mxA[j] = mxB[j] = 0u; }
#pragma unroll
    for (int dy = 0; dy < 4; dy++)
#pragma unroll
        for (int k = 0; k < 2 * T; k++) { pxa[dy][k] = 0; pxb[dy][k] = 0; }
    uint8_t *ocol = o + (ptrdiff_t)t0 * 4;
    const ptrdiff_t tp_ = g.tp;
    // bulk-store form: the leading nb16 bytes of each output row of the strip leave through shared memory (s0 is a multiple of 4
    // tiles and the map pitch a multiple of 16, so source and destination are 16-byte aligned); the last <= 3 tiles are stored directly
    const int nb16 = ((s1 - s0) * 4) & ~15;
    const int lane_off = (t0 - s0) * 4;                       // byte offset of the lane's first tile in a staged row (negative: halo)

    // one tile row: CUR receives its pixels, the row before it (pixels in PREV, min/max in *B) is written out
#define CB_TM_STEP(CUR, PREV, IT)                                                                                                     \
    {                                                                                                                                 \
        const int it_ = (IT), stage_ = it_ % TM_STAGES, r_ = rstart + dir * it_;                                                      \
        mbar_wait(&bars[stage_], (uint32_t)((it_ / TM_STAGES) & 1));                                                                  \
        const unsigned char *sp_ = ring + (size_t)stage_ * C::STAGEB + lane * (8 * T);                                                \
        uint32_t mnp_[T], mxp_[T];                                                                                                    \
        _Pragma("unroll") for (int dy = 0; dy < 4; dy++) {                                                                            \
            uint32_t w_[2 * T];                                                                                                       \
            _Pragma("unroll") for (int q = 0; q < T / 2; q++) {                                                                       \
                const uint4 v_ = *reinterpret_cast<const uint4 *>(sp_ + dy * C::ROWB + 16 * q);                                       \
                w_[4 * q] = v_.x; w_[4 * q + 1] = v_.y; w_[4 * q + 2] = v_.z; w_[4 * q + 3] = v_.w;                                   \
            }                                                                                                                         \
            _Pragma("unroll") for (int k = 0; k < T; k++) {                                                                           \
                const uint32_t e0_ = w_[2 * k] & 0x00ff00ffu, e1_ = w_[2 * k + 1] & 0x00ff00ffu;                                      \
                if (dy == 0) { mnp_[k] = __vminu2(e0_, e1_); mxp_[k] = __vmaxu2(e0_, e1_); }                                          \
                else { mnp_[k] = __vimin3_u16x2(mnp_[k], e0_, e1_); mxp_[k] = __vimax3_u16x2(mxp_[k], e0_, e1_); }                    \
                CUR[dy][2 * k] = e0_; CUR[dy][2 * k + 1] = e1_;                                                                       \
            }                                                                                                                         \
        }                                                                                                                             \
        __syncwarp();                                                                                                                 \
        if (lane == 0 && it_ + TM_STAGES < nsteps) {                                                                                  \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                                              \
            mbar_expect_tx(&bars[stage_], (uint32_t)C::STAGEB);                                                                       \
            tma_load_3d(ring + (size_t)stage_ * C::STAGEB, &tmap, tbase, 4 * (r_ + dir * TM_STAGES), b, &bars[stage_]);               \
        }                                                                                                                             \
        const bool rvalid_ = r_ >= 0 && r_ < g.th;                                                                                    \
        uint32_t mnC_[P], mxC_[P], vmn_[P + 2], vmx_[P + 2];                                                                          \
        _Pragma("unroll") for (int j = 0; j < P; j++) {                                                                               \
            const uint32_t a_ = __vminu2(prmt(mnp_[2 * j], mnp_[2 * j + 1], 0x5410u), prmt(mnp_[2 * j], mnp_[2 * j + 1], 0x7632u));   \
            const uint32_t c_ = __vmaxu2(prmt(mxp_[2 * j], mxp_[2 * j + 1], 0x5410u), prmt(mxp_[2 * j], mxp_[2 * j + 1], 0x7632u));   \
            mnC_[j] = rvalid_ ? (a_ | mn_or[j]) : 0x00ff00ffu;                                                                        \
            mxC_[j] = rvalid_ ? (c_ & mx_and[j]) : 0u;                                                                                \
            vmn_[j + 1] = __vimin3_u16x2(mnA[j], mnB[j], mnC_[j]);                                                                    \
            vmx_[j + 1] = __vimax3_u16x2(mxA[j], mxB[j], mxC_[j]);                                                                    \
        }                                                                                                                             \
        vmn_[0] = __shfl_up_sync(full, vmn_[P], 1); vmx_[0] = __shfl_up_sync(full, vmx_[P], 1);                                       \
        vmn_[P + 1] = __shfl_down_sync(full, vmn_[1], 1); vmx_[P + 1] = __shfl_down_sync(full, vmx_[1], 1);                           \
        if (lane == 0) { vmn_[0] = 0x00ff00ffu; vmx_[0] = 0u; }                                                                       \
        if (lane == 31) { vmn_[P + 1] = 0x00ff00ffu; vmx_[P + 1] = 0u; }                                                              \
        const int ro_ = r_ - dir;                                                                                                     \
        if (C::BULK && it_ >= 2) {      /* (warp-uniform) the bulk stores issued two steps ago have finished reading this staging buffer */ \
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");                                             \
            __syncwarp();                                                                                                             \
        }                                                                                                                             \
        if (it_ >= 2 && any_in) {                                                                                                     \
            uint8_t *const p0_ = ocol + (ptrdiff_t)(ro_ * 4) * tp_;                                                                   \
            unsigned char *const sb0_ = obuf + (it_ & 1) * (4 * C::ROWOUT) + lane_off;                                                \
            uint32_t sh_mn_ = prmt(vmn_[0], vmn_[1], 0x5432u), sh_mx_ = prmt(vmx_[0], vmx_[1], 0x5432u);                              \
            uint32_t oa_[P][4], ob_[P][4];                                                                                            \
            _Pragma("unroll") for (int j = 0; j < P; j++) {                                                                           \
                const uint32_t nx_mn_ = prmt(vmn_[j + 1], vmn_[j + 2], 0x5432u), nx_mx_ = prmt(vmx_[j + 1], vmx_[j + 2], 0x5432u);    \
                const uint32_t dmn_ = __vimin3_u16x2(vmn_[j + 1], sh_mn_, nx_mn_), dmx_ = __vimax3_u16x2(vmx_[j + 1], sh_mx_, nx_mx_); \
                sh_mn_ = nx_mn_; sh_mx_ = nx_mx_;                                                                                     \
                const uint32_t D_ = dmx_ - dmn_;                                                                                      \
                const uint32_t cc_ = 0x7fff7fffu - (dmn_ + ((D_ >> 1) & 0x7fff7fffu));                                                \
                const uint32_t F_ = D_ + fconst;                                                                                      \
                const uint32_t c0_ = prmt(cc_, 0u, 0x1010u), c1_ = prmt(cc_, 0u, 0x3232u);                                            \
                const uint32_t nf0_ = prmt(F_, 0u, 0x9999u), nf1_ = prmt(F_, 0u, 0xbbbbu);                                            \
                const uint32_t fl0_ = 0x7f7f7f7fu & ~nf0_, fl1_ = 0x7f7f7f7fu & ~nf1_;                                                \
                _Pragma("unroll") for (int dy = 0; dy < 4; dy++) {                                                                    \
                    oa_[j][dy] = (prmt(PREV[dy][4 * j] + c0_, PREV[dy][4 * j + 1] + c0_, 0xfdb9u) & nf0_) | fl0_;                     \
                    ob_[j][dy] = (prmt(PREV[dy][4 * j + 2] + c1_, PREV[dy][4 * j + 3] + c1_, 0xfdb9u) & nf1_) | fl1_;                 \
                }                                                                                                                     \
                if (C::BULK && pair_st[j] && lane_off + 8 * j < nb16) {        /* (a pair never straddles nb16: both multiples of 8) */ \
                    _Pragma("unroll") for (int dy = 0; dy < 4; dy++)                                                                  \
                        *reinterpret_cast<uint2 *>(sb0_ + dy * C::ROWOUT + 8 * j) = make_uint2(oa_[j][dy], ob_[j][dy]);               \
                }                                                                                                                     \
                if (write_tiles && pair_st[j]) {                                                                                      \
                    const size_t ti_ = ((size_t)b * g.th + ro_) * g.tw + t0 + 2 * j;                                                  \
                    tmin[ti_] = (uint8_t)mnB[j]; tmax[ti_] = (uint8_t)mxB[j];                                                         \
                    if (pair_st[j] == 2) { tmin[ti_ + 1] = (uint8_t)(mnB[j] >> 16); tmax[ti_ + 1] = (uint8_t)(mxB[j] >> 16); }        \
                }                                                                                                                     \
            }                                                                                                                         \
            /* direct stores, ROW BY ROW: the three 8-byte pieces a lane adds to a 32-byte sector reach the L2 back to back (the */   \
            /* pair-major order of the first version cost 10 % of the kernel); in the bulk form only the strip's last <= 3 tiles */   \
            _Pragma("unroll") for (int dy = 0; dy < 4; dy++)                                                                          \
                _Pragma("unroll") for (int j = 0; j < P; j++) {                                                                       \
                    if (C::BULK && lane_off + 8 * j < nb16) continue;                                                                 \
                    uint8_t *const q_ = p0_ + dy * tp_ + 8 * j;                                                                       \
                    if (pair_st[j] == 2) *reinterpret_cast<uint2 *>(q_) = make_uint2(oa_[j][dy], ob_[j][dy]);                         \
                    else if (pair_st[j] == 1) *reinterpret_cast<uint32_t *>(q_) = oa_[j][dy];                                         \
                }                                                                                                                     \
        }                                                                                                                             \
        if (C::BULK && it_ >= 2) {                                                                                                    \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                                              \
            __syncwarp();                                                                                                             \
            if (lane == 0 && nb16 > 0) {                                                                                              \
                const unsigned char *src_ = obuf + (it_ & 1) * (4 * C::ROWOUT);                                                       \
                uint8_t *dst_ = o + (ptrdiff_t)(ro_ * 4) * tp_ + s0 * 4;                                                              \
                _Pragma("unroll") for (int dy = 0; dy < 4; dy++)                                                                      \
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_ + dy * tp_),                \
                                 "r"(smem_u32(src_ + dy * C::ROWOUT)), "r"(nb16) : "memory");                                         \
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");                                                             \
            }                                                                                                                         \
        }                                                                                                                             \
        _Pragma("unroll") for (int j = 0; j < P; j++) { mnA[j] = mnB[j]; mnB[j] = mnC_[j]; mxA[j] = mxB[j]; mxB[j] = mxC_[j]; }       \
    }

    int it = 0;
    for (; it + 1 < nsteps; it += 2) {
        CB_TM_STEP(pxa, pxb, it)
        CB_TM_STEP(pxb, pxa, it + 1)
    }
    if (it < nsteps) CB_TM_STEP(pxa, pxb, it)
#undef CB_TM_STEP
    if (C::BULK && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the staging buffers live until they are read
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

// Persistent form of the conversion (default for aligned frames since round 2): the kernel above refills an SM block by block
// (a CTA loads its 16 chunks, converts them and exits; ncu: long_scoreboard, 45 % of the warps active, 0.69 of the HBM peak).  Here
// a warp keeps a ring of RGB_RING chunks in flight for as long as the launch lasts -- the structure of the threshold kernel: lane 0
// re-arms a stage with the chunk RGB_RING turns ahead as soon as the warp has read it, so the loads never drain; chunk g of the
// flattened (frame, chunk) space goes to warp g mod nwarps, i.e. at any moment the warps read one contiguous window of memory.
// Only whole 512-pixel chunks of 16-byte aligned frames; the tails (npix mod 512 pixels per frame) are converted by the first
// warps after their loop.
constexpr int RGB_RING = 4, RGB_RING_WARPS = 4;
__global__ void __launch_bounds__(RGB_RING_WARPS * 32) rgb_to_gray_ring_kernel(const uint8_t *__restrict__ rgb, uint8_t *__restrict__ gray, size_t npix,
                                                                               size_t in_stride, size_t out_stride, int nframes)
{
    __shared__ __align__(128) uint4 stage[RGB_RING_WARPS][RGB_RING][96];
    __shared__ __align__(8) unsigned long long bar[RGB_RING_WARPS][RGB_RING];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * RGB_RING_WARPS, w = (long long)blockIdx.x * RGB_RING_WARPS + warp;
    const long long cpf = (long long)(npix / 512), total = cpf * nframes;
    auto src_of = [&](long long gidx) { return rgb + (size_t)(gidx / cpf) * in_stride + (size_t)(gidx % cpf) * 1536; };
    if (lane == 0) {
        for (int s = 0; s < RGB_RING; s++) mbar_init(&bar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < RGB_RING; s++) {
            const long long gi = w + (long long)s * nwarps;
            if (gi < total) { mbar_expect_tx(&bar[warp][s], 1536); tma_load_1d(&stage[warp][s][0], src_of(gi), 1536, &bar[warp][s]); }
        }
    }
    __syncwarp();
    int it = 0;
    for (long long gi = w; gi < total; gi += nwarps, it++) {
        const int s = it % RGB_RING;
        mbar_wait(&bar[warp][s], (uint32_t)((it / RGB_RING) & 1));
        const uint4 a = stage[warp][s][3 * lane], b2 = stage[warp][s][3 * lane + 1], cc = stage[warp][s][3 * lane + 2];
        __syncwarp();
        if (lane == 0) {
            const long long gn = gi + (long long)RGB_RING * nwarps;
            if (gn < total) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bar[warp][s], 1536);
                tma_load_1d(&stage[warp][s][0], src_of(gn), 1536, &bar[warp][s]);
            }
        }
        const uint32_t wds[12] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w, cc.x, cc.y, cc.z, cc.w};
        uint32_t q[16], o[4];
#pragma unroll
        for (int k = 0; k < 16; k++) {      // (the conversion-free u8 <-> f32 arithmetic of rgb_to_gray_kernel)
            const float r = __fsub_rn(__uint_as_float(__byte_perm(wds[(3 * k) >> 2], 0x4B000000u, 0x7440 | ((3 * k) & 3))), 8388608.0f);
            const float gg = __fsub_rn(__uint_as_float(__byte_perm(wds[(3 * k + 1) >> 2], 0x4B000000u, 0x7440 | ((3 * k + 1) & 3))), 8388608.0f);
            const float bl = __fsub_rn(__uint_as_float(__byte_perm(wds[(3 * k + 2) >> 2], 0x4B000000u, 0x7440 | ((3 * k + 2) & 3))), 8388608.0f);
            const float v = __fmaf_rn(r, 0.33f, __fmaf_rn(gg, 0.33f, __fmul_rn(bl, 0.33f)));          // utils.rs:43
            q[k] = __float_as_uint(__fadd_rz(v, 8388608.0f));
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            o[k] = __byte_perm(__byte_perm(q[4 * k], q[4 * k + 1], 0x4040), __byte_perm(q[4 * k + 2], q[4 * k + 3], 0x4040), 0x5410);
        *reinterpret_cast<uint4 *>(gray + (size_t)(gi / cpf) * out_stride + (size_t)(gi % cpf) * 512 + 16 * lane) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // ragged tails: frame f by warp f
    const size_t rest = (size_t)cpf * 512;
    if (rest < npix)
        for (long long f = w; f < nframes; f += nwarps)
            for (size_t k = rest + lane; k < npix; k += 32)
                gray[(size_t)f * out_stride + k] = cat_gray(rgb[(size_t)f * in_stride + 3 * k], rgb[(size_t)f * in_stride + 3 * k + 1], rgb[(size_t)f * in_stride + 3 * k + 2]);
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
