// quads.cuh -- row A5 of SURVEY.md 8a: fit_quads() / fit_quad().
//
// Upstream (apriltag_quad_thresh.c fit_quad, ptsort, compute_lfps, fit_line, quad_segment_maxima).  The float /
// double arithmetic below follows upstream operation by operation (the library is built with --fmad=false), and
// every order-dependent accumulation keeps upstream's order:
//   * points are first put back into scan order (y, x, probe) -- the order upstream's hash map appends them in;
//   * ptsort()'s result is reproduced exactly: its leaf networks are emulated, the merges are replaced by a sort on a
//     composite key that has the same order (sort.cuh);
//   * the line-fit prefix moments are accumulated sequentially (six lanes, one per moment), 32 points at a time
//     through a shared-memory ring, so that the dependent chain is one DADD per point;
//   * the 4-corner search evaluates all <=210 subsets in parallel and keeps the first minimum in loop order.
//
// B200 mapping: the work is thousands of small, irregular, partly sequential jobs per frame (a c2 frame has ~700
// clusters of 24..7000 points), so the design maximises the number of clusters in flight instead of the threads per
// cluster.  Persistent kernels pull (frame, cluster) items from four device-side work lists (by cluster size):
//   sort #1, sort #2 (sort.cuh): five kernels each -- one warp per cluster up to 512 points, CTAs of 128 / 256 / 512
//           threads above -- on a register-blocked merge-path sort in shared memory;
//   lfps_kernel: prefix moments + window errors, one warp per cluster of any size, the prefix array stays in a
//           shared-memory ring;
//   fit_quads_kernel: maxima, corner search, quad tests, one warp per cluster of any size.
// A boundary point is fully described by its 32-bit scan key (pixel index, probe, gradient sign), so the only
// per-point input is 4 bytes; per point the stage moves 4 B (keys) + 4 B (sorted points) + 8 B (window error) + 6 B
// (checkpoints) through global memory.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "divby.cuh"

namespace cb {

struct LineFit { double Ex, Ey, nx, ny, err, mse; };

// One entry of upstream's prefix-moment array lfps[] (struct line_fit_pt).
struct M6 { double Mx, My, Mxx, Mxy, Myy, W; };

// fit_line() given the (at most) three entries of lfps[] it reads: a = lfps[i1], p = lfps[i0 - 1], l = lfps[sz - 1].
//   mode 0: i0 == 0 < i1 (a alone), mode 1: 0 < i0 < i1 (a - p), mode 2: i0 > i1, the range wraps ((l - p) + a).
// N = number of points in the range.
__device__ __forceinline__ void fit_line_m(const M6 &a, const M6 &p, const M6 &l, int mode, int N, bool want_params, LineFit &o)
{
    double Mx, My, Mxx, Myy, Mxy, W;
    if (mode != 2) {
        Mx = a.Mx; My = a.My; Mxx = a.Mxx; Mxy = a.Mxy; Myy = a.Myy; W = a.W;
        if (mode == 1) { Mx -= p.Mx; My -= p.My; Mxx -= p.Mxx; Mxy -= p.Mxy; Myy -= p.Myy; W -= p.W; }
    } else {
        Mx = l.Mx - p.Mx; My = l.My - p.My; Mxx = l.Mxx - p.Mxx; Mxy = l.Mxy - p.Mxy; Myy = l.Myy - p.Myy; W = l.W - p.W;
        Mx += a.Mx; My += a.My; Mxx += a.Mxx; Mxy += a.Mxy; Myy += a.Myy; W += a.W;
    }
    const DivBy by_w(W);
    const double Ex = by_w(Mx), Ey = by_w(My);
    const double Cxx = by_w(Mxx) - Ex * Ex, Cxy = by_w(Mxy) - Ex * Ey, Cyy = by_w(Myy) - Ey * Ey;
    const float disc = sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy));
    const double eig_small = 0.5 * (Cxx + Cyy - disc);
    if (want_params) {
        o.Ex = Ex; o.Ey = Ey;
        const double eig = 0.5 * (Cxx + Cyy + disc);
        const double nx1 = Cxx - eig, ny1 = Cxy, M1 = nx1 * nx1 + ny1 * ny1;
        const double nx2 = Cxy, ny2 = Cyy - eig, M2 = nx2 * nx2 + ny2 * ny2;
        double nx, ny, M;
        if (M1 > M2) { nx = nx1; ny = ny1; M = M1; } else { nx = nx2; ny = ny2; M = M2; }
        const double length = sqrtf((float)M);
        if (fabs(length) < 1e-12) { o.nx = 0; o.ny = 0; }
        else { const DivBy by_len(length); o.nx = by_len(nx); o.ny = by_len(ny); }
    }
    o.err = N * eig_small;
    o.mse = eig_small;
}

// The six terms point xy = px | py << 16 adds to the prefix moments (upstream compute_lfps): gradient-magnitude weight
// from the full-resolution image around the decimated pixel.
__device__ __forceinline__ void point_terms(const uint8_t *__restrict__ img, uint32_t xy, const Geom &g, double t[6])
{
    const int px = (int)(xy & 0xffff), py = (int)(xy >> 16);
    const double x = px * .5 + 0.5, y = py * .5 + 0.5;
    const int ix = (int)x, iy = (int)y;
    double W = 1;
    if (ix > 0 && ix + 1 < g.w && iy > 0 && iy + 1 < g.h) {
        const uint8_t *row = img + (size_t)(iy * g.f) * g.stride;
        const int grad_x = (int)row[(ix + 1) * g.f] - (int)row[(ix - 1) * g.f];
        const int grad_y = (int)row[(ptrdiff_t)ix * g.f + (ptrdiff_t)g.f * g.stride] - (int)row[(ptrdiff_t)ix * g.f - (ptrdiff_t)g.f * g.stride];
        W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
    }
    t[0] = W * x; t[1] = W * y; t[2] = W * x * x; t[3] = W * x * y; t[4] = W * y * y; t[5] = W;
}

// lfps[idx] rebuilt from the checkpoint before it (every LF_CP-th entry is kept) plus at most LF_CP - 1 replayed points;
// the additions happen in the same order as in the sequential pass, so the value is the same.
constexpr int LF_CP = 8;
__device__ __forceinline__ void replay_entry(const uint8_t *__restrict__ img, const uint32_t *__restrict__ XY, const double *__restrict__ cp,
                                             int idx, const Geom &g, double acc[6])
{
    const int blk = idx / LF_CP;
#pragma unroll
    for (int m = 0; m < 6; m++) acc[m] = blk > 0 ? cp[(size_t)(blk - 1) * 6 + m] : 0.0;
    for (int j = blk * LF_CP; j <= idx; j++) {
        double t[6];
        point_terms(img, XY[j], g, t);
#pragma unroll
        for (int m = 0; m < 6; m++) acc[m] += t[m];
    }
}

// 4-subsets of {0..9} packed (m0<<12|m1<<8|m2<<4|m3), in colex order so that the subsets of {0..k-1} are the first
// C(k,4) entries; filled by the host (api.cu)
__constant__ uint16_t c_combos[210];

// ---- maxima, corner search, corner / area / convexity tests: ONE WARP per cluster, every tier ------------------------------
// upstream quad_segment_maxima() from its second loop on, and the rest of fit_quad().  After the window errors exist the
// only per-point work left is a 7-tap smoothing and a local-maximum test, which stream through a 40-entry tile; all that
// follows works on <= 10 maxima.  So a cluster of any size needs one warp and ~7 KB of shared memory, no barriers:
//   * smoothed errors and local maxima, 32 points per step (wrap-around by modular loads of the global errs array);
//   * the best max_nmaxima + 1 maxima are kept in a sorted list, one per lane (insertion by shuffle) -- upstream's
//     "threshold = element [max_nmaxima] of the descending sort, keep err > threshold";
//   * the <= 21 prefix-moment entries the kept maxima refer to are rebuilt from checkpoints (replay_entry);
//   * pair table fit_line(kept[a], kept[c]), <= 210 four-subsets evaluated in parallel, first minimum in loop order;
//   * line intersections on four lanes, area and convexity on one.
constexpr int FQ_WARPS = 4;
struct FqWarp {
    double et[40];          // window errors of points j0 - 4 .. j0 + 35 (indices wrapped)
    double ys[34];          // smoothed errors of points j0 - 1 .. j0 + 32
    M6 ent[21];             // lfps[kept[m]], lfps[kept[m] - 1] (m < 10), lfps[n - 1]
    double p_err[10][10], p_mse[10][10], p_nx[10][10], p_ny[10][10], p_ex[10][10], p_ey[10][10];
    double lines[4][4];
    float qp[4][2];
    int kept[10];
};

// Smoothed errors, local maxima and the running list of the best max_nmaxima + 1 maxima (one per lane: te / ti, descending error;
// ties: lower index first) over the points [ja, jb) of a cluster of n points (ja a multiple of 32); nm counts the maxima seen.
// One warp; et / ys are its tiles in shared memory.
__device__ __forceinline__ void fq_scan(double *__restrict__ et, double *__restrict__ ys, const double *__restrict__ errs, int n, int ja, int jb,
                                        const DetParams &prm, int lane, double &te, int &ti, int &nm)
{
    const uint32_t full = 0xffffffffu;
    double t11 = __shfl_sync(full, te, 10);           // error of the 11th list entry (-inf while the list is short)
    // tile loads run one step ahead of their use (the errs array comes from L2 / HBM)
    auto load_err = [&](int idx) { if (idx < 0) idx += n; while (idx >= n) idx -= n; return errs[idx]; };
    double pre0 = load_err(ja - 4 + lane), pre1 = lane < 8 ? load_err(ja - 4 + 32 + lane) : 0.0;
    for (int j0 = ja; j0 < jb; j0 += 32) {
        et[lane] = pre0;
        if (lane < 8) et[32 + lane] = pre1;
        if (j0 + 32 < jb) {
            pre0 = load_err(j0 + 32 - 4 + lane);
            if (lane < 8) pre1 = load_err(j0 + 32 - 4 + 32 + lane);
        }
        __syncwarp();
        for (int u = lane; u < 34; u += 32) {
            double acc = 0;
#pragma unroll
            for (int i = 0; i < 7; i++) acc += et[u + i] * prm.smooth_f[i];
            ys[u] = acc;
        }
        __syncwarp();
        bool is_max = false;
        double e = 0;
        if (j0 + lane < jb) {
            e = ys[lane + 1];
            is_max = e > ys[lane + 2] && e > ys[lane];
        }
        nm += __popc(__ballot_sync(full, is_max));
        // only maxima above the current 11th best can enter the list (an equal error with a later index cannot)
        uint32_t bal = __ballot_sync(full, is_max && e > t11);
        while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const double ev = __shfl_sync(full, e, src);
            const int pos = __popc(__ballot_sync(full, lane < 11 && te >= ev));
            const double up_te = __shfl_up_sync(full, te, 1);
            const int up_ti = __shfl_up_sync(full, ti, 1);
            if (lane == pos) { te = ev; ti = j0 + src; }
            else if (lane > pos) { te = up_te; ti = up_ti; }
            t11 = __shfl_sync(full, te, 10);
        }
        __syncwarp();
    }
}

// The rest of fit_quad() once the maxima list of a cluster is known (one warp): threshold on the list, the <= 21 prefix-moment
// entries rebuilt from checkpoints, pair table, 4-corner search, corners, area / convexity tests, output.
__device__ __forceinline__ void fq_finish(FqWarp &S, double te, int ti, int nm, int n, int b, const ClusterRec &rec, size_t pbase,
                                          const uint8_t *__restrict__ in, const uint32_t *__restrict__ sorted_xy, const double *__restrict__ cp_all,
                                          QuadRec *__restrict__ quads, uint32_t *__restrict__ nquads, uint32_t *__restrict__ nquads_total,
                                          uint32_t *__restrict__ errflag, const Geom &g, const Caps &caps, const DetParams &prm, int lane)
{
    const int tid = lane;
    const uint32_t full = 0xffffffffu;
    const int reversed_border = 0;          // reversed clusters never get past the sort (tag36h11 has a normal border only)
    if (nm < 4) return;
    const int max_nmaxima = min(prm.max_nmaxima, 10);
    bool keep;
    if (nm > max_nmaxima) {
        const double thresh = __shfl_sync(full, te, max_nmaxima);
        keep = lane < max_nmaxima && te > thresh;
    } else {
        keep = lane < nm;
    }
    const uint32_t kb = __ballot_sync(full, keep);
    const int nk = __popc(kb);
    {
        int rank = 0;
        for (int q = 0; q < 11; q++) {
            const int oi = __shfl_sync(full, ti, q);
            if ((kb >> q) & 1) rank += oi < ti ? 1 : 0;
        }
        if (keep) S.kept[rank] = ti;
    }
    __syncwarp();
    if (nk < 4) return;   // (upstream's loops would simply find nothing)

    // ---- the prefix-moment entries the corner search reads: lfps[kept[m]], lfps[kept[m] - 1], lfps[n - 1] ----
    if (tid <= 2 * nk) {
        const int idx = tid == 2 * nk ? n - 1 : S.kept[tid >> 1] - (tid & 1);
        if (idx >= 0) {
            double acc[6];
            replay_entry(in + (size_t)b * g.frame_stride, sorted_xy + pbase, cp_all + (pbase / LF_CP) * 6, idx, g, acc);
            M6 &e = S.ent[tid];
            e.Mx = acc[0]; e.My = acc[1]; e.Mxx = acc[2]; e.Mxy = acc[3]; e.Myy = acc[4]; e.W = acc[5];
        }
    }
    __syncwarp();
    // pair table: fit_line(lfps, n, kept[a], kept[c]) for a != c
    for (int t = tid; t < nk * nk; t += 32) {
        const int a = t / nk, c = t % nk;
        if (a == c) continue;
        const int i0 = S.kept[a], i1 = S.kept[c];
        LineFit lf;
        if (i0 < i1) fit_line_m(S.ent[2 * c], S.ent[2 * a + 1], S.ent[2 * nk], i0 > 0 ? 1 : 0, i1 - i0 + 1, true, lf);
        else fit_line_m(S.ent[2 * c], S.ent[2 * a + 1], S.ent[2 * nk], 2, n - i0 + i1 + 1, true, lf);
        S.p_err[a][c] = lf.err; S.p_mse[a][c] = lf.mse; S.p_nx[a][c] = lf.nx; S.p_ny[a][c] = lf.ny;
        S.p_ex[a][c] = lf.Ex; S.p_ey[a][c] = lf.Ey;
    }
    __syncwarp();
    // 4-corner search; (m0,m1,m2,m3) packed big-endian orders like upstream's loop nest, so the minimum over
    // (err, packed) is upstream's "first minimum"
    double best_err = __longlong_as_double(0x7ff0000000000000ll);
    int best_combo = 1 << 30;
    {
        const double max_mse = (double)prm.max_line_fit_mse;
        const int ncomb = nk * (nk - 1) * (nk - 2) * (nk - 3) / 24;
        for (int ci = tid; ci < ncomb; ci += 32) {
            const int pk = c_combos[ci];
            const int m0 = pk >> 12, m1 = (pk >> 8) & 15, m2 = (pk >> 4) & 15, m3 = pk & 15;
            if (S.p_mse[m0][m1] > max_mse) continue;
            if (S.p_mse[m1][m2] > max_mse) continue;
            const double dt = S.p_nx[m0][m1] * S.p_nx[m1][m2] + S.p_ny[m0][m1] * S.p_ny[m1][m2];
            if (fabs(dt) > prm.cos_critical_rad) continue;
            if (S.p_mse[m2][m3] > max_mse) continue;
            if (S.p_mse[m3][m0] > max_mse) continue;
            const double err = S.p_err[m0][m1] + S.p_err[m1][m2] + S.p_err[m2][m3] + S.p_err[m3][m0];
            if (err < best_err || (err == best_err && pk < best_combo)) { best_err = err; best_combo = pk; }
        }
    }
    {
        double bmin = best_err;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double ob = __shfl_xor_sync(full, bmin, o); bmin = ob < bmin ? ob : bmin; }
        int cand = (best_err == bmin && best_combo != (1 << 30)) ? best_combo : (1 << 30);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(full, cand, o));
        best_err = bmin; best_combo = cand;
    }
    if (best_combo == (1 << 30)) return;
    if (!(best_err / n < (double)prm.max_line_fit_mse)) return;

    // ---- corners (four lanes), area and convexity tests (one lane) -----------------------------------------------
    const int mi[4] = {(best_combo >> 12) & 15, (best_combo >> 8) & 15, (best_combo >> 4) & 15, best_combo & 15};
    bool bad = false;
    if (tid < 4) {
        const int a = mi[tid], c = mi[(tid + 1) & 3];
        S.lines[tid][0] = S.p_ex[a][c]; S.lines[tid][1] = S.p_ey[a][c]; S.lines[tid][2] = S.p_nx[a][c]; S.lines[tid][3] = S.p_ny[a][c];
        bad = S.p_mse[a][c] > (double)prm.max_line_fit_mse;
    }
    if (__any_sync(full, bad)) return;
    if (tid < 4) {
        const int i = tid;
        const double A00 = S.lines[i][3], A01 = -S.lines[(i + 1) & 3][3];
        const double A10 = -S.lines[i][2], A11 = S.lines[(i + 1) & 3][2];
        const double B0 = -S.lines[i][0] + S.lines[(i + 1) & 3][0];
        const double B1 = -S.lines[i][1] + S.lines[(i + 1) & 3][1];
        const double det = A00 * A11 - A10 * A01;
        const double W00 = A11 / det, W01 = -A01 / det;
        if (fabs(det) < 0.001) bad = true;
        else {
            const double L0 = W00 * B0 + W01 * B1;
            S.qp[i][0] = (float)(S.lines[i][0] + L0 * A00);
            S.qp[i][1] = (float)(S.lines[i][1] + L0 * A10);
        }
    }
    if (__any_sync(full, bad)) return;
    if (tid == 0) {
        float qp[4][2];
        for (int i = 0; i < 4; i++) { qp[i][0] = S.qp[i][0]; qp[i][1] = S.qp[i][1]; }
        bool good = true;
        if (good) {
            double area = 0, length[3], p;
            for (int i = 0; i < 3; i++) {
                const int a = i, c = (i + 1) % 3;
                const double ddx = (double)qp[c][0] - (double)qp[a][0], ddy = (double)qp[c][1] - (double)qp[a][1];
                length[i] = sqrt(ddx * ddx + ddy * ddy);
            }
            p = (length[0] + length[1] + length[2]) / 2;
            area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
            const int idxs[4] = {2, 3, 0, 2};
            for (int i = 0; i < 3; i++) {
                const int a = idxs[i], c = idxs[i + 1];
                const double ddx = (double)qp[c][0] - (double)qp[a][0], ddy = (double)qp[c][1] - (double)qp[a][1];
                length[i] = sqrt(ddx * ddx + ddy * ddy);
            }
            p = (length[0] + length[1] + length[2]) / 2;
            area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
            if (area < 0.95 * prm.min_tag_width * prm.min_tag_width) good = false;
        }
        if (good) {
            for (int i = 0; i < 4; i++) {
                const int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
                const double dx1 = (double)qp[i1][0] - (double)qp[i0][0], dy1 = (double)qp[i1][1] - (double)qp[i0][1];
                const double dx2 = (double)qp[i2][0] - (double)qp[i1][0], dy2 = (double)qp[i2][1] - (double)qp[i1][1];
                const double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
                if ((cos_dtheta > prm.cos_critical_rad || cos_dtheta < -prm.cos_critical_rad) || dx1 * dy2 < dy1 * dx2) { good = false; break; }
            }
        }
        if (good) {
            const uint32_t qf = atomicAdd(&nquads[b], 1u);
            if (qf >= caps.quads_per_frame) flag_overflow(errflag, b, ERR_QUADS_FULL);
            else {
                const uint32_t qi = atomicAdd(nquads_total, 1u);   // < batch * quads_per_frame by construction
                QuadRec q;
                for (int i = 0; i < 4; i++) { q.p[i][0] = qp[i][0]; q.p[i][1] = qp[i][1]; }
                q.reversed_border = reversed_border; q.npoints = n; q.key = rec.key; q.frame = b; q.pad = 0;
                quads[qi] = q;
            }
        }
    }
}

// (eight CTAs per SM: the register allocation is held to 64 -- 12 bytes of spills -- because the kernel is latency bound: 24 -> 32
// warps per SM took 3 % off the quad stage; the decoder, with 450 bytes of spills at 64 registers, lost 3 % the same way)
__global__ void __launch_bounds__(FQ_WARPS * 32, 8)
fit_quads_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ sorted_xy, const ClusterRec *__restrict__ clusters,
                 const uint32_t *__restrict__ worklists, size_t wl_stride, const uint32_t *__restrict__ nwork /* stride 2, 4 tiers */,
                 uint32_t *__restrict__ work_counter, const double *__restrict__ errs_all, const double *__restrict__ cp_all,
                 QuadRec *__restrict__ quads, uint32_t *__restrict__ nquads, uint32_t *__restrict__ nquads_total,
                 uint32_t *__restrict__ errflag, Geom g, Caps caps, DetParams prm, int n_big)
{
    __shared__ FqWarp sh[FQ_WARPS];
    const int lane = threadIdx.x & 31;
    const uint32_t full = 0xffffffffu;
    FqWarp &S = sh[threadIdx.x >> 5];
    for (;;) {
        __syncwarp();
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(full, wi, 0);
        uint32_t item;
        if (!tier_item<0, 3>(wi, nwork, worklists, wl_stride, item)) return;
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        if (rec.cursor == 0xffffffffu || rec.count < 24 || (int)rec.count >= n_big) continue;      // (>= n_big: fit_quads_big_kernel)
        const int n = (int)rec.count;
        const int ksz = min(20, n / 12);
        if (ksz < 2) continue;
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;
        double te = __longlong_as_double(0xfff0000000000000ll);
        int ti = 1 << 30, nm = 0;
        fq_scan(S.et, S.ys, errs_all + pbase, n, 0, n, prm, lane, te, ti, nm);
        fq_finish(S, te, ti, nm, n, b, rec, pbase, in, sorted_xy, cp_all, quads, nquads, nquads_total, errflag, g, caps, prm, lane);
    }
}

// ONE CTA per LARGE cluster (small batches): the scan over the window errors is the part of fit_quads_kernel that grows with the
// cluster (0.65 us per 32 points, 100 us for 5 000 points).  Eight warps scan an eighth each, every warp keeps its own list of the best
// 11 maxima; warp 0 merges the lists by inserting the candidates in warp order -- within a warp's list equal errors are already in
// index order and lower warps hold lower indices, so the merged list is the one a single scan builds -- and finishes the cluster.
constexpr int FQB_WARPS = 8;
struct FqBig {
    FqWarp w0;
    double et[FQB_WARPS][40], ys[FQB_WARPS][34];
    double lte[FQB_WARPS][11];
    int lti[FQB_WARPS][11], lnm[FQB_WARPS];
    uint32_t work;
};
template <int T_LO>
__global__ void __launch_bounds__(FQB_WARPS * 32)
fit_quads_big_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ sorted_xy, const ClusterRec *__restrict__ clusters,
                     const uint32_t *__restrict__ worklists, size_t wl_stride, const uint32_t *__restrict__ nwork /* stride 2, 4 tiers */,
                     uint32_t *__restrict__ work_counter, const double *__restrict__ errs_all, const double *__restrict__ cp_all,
                     QuadRec *__restrict__ quads, uint32_t *__restrict__ nquads, uint32_t *__restrict__ nquads_total,
                     uint32_t *__restrict__ errflag, Geom g, Caps caps, DetParams prm, int n_big)
{
    extern __shared__ __align__(16) unsigned char fqb_smem[];
    FqBig &S = *reinterpret_cast<FqBig *>(fqb_smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) S.work = atomicAdd(work_counter, 1u);
        __syncthreads();
        uint32_t item;
        if (!tier_item<T_LO, 3>(S.work, nwork, worklists, wl_stride, item)) return;
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        if (rec.cursor == 0xffffffffu || (int)rec.count < n_big || rec.count < 24) continue;
        const int n = (int)rec.count;
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;
        const int seg = ((n + FQB_WARPS - 1) / FQB_WARPS + 31) / 32 * 32;
        const int ja = min(wid * seg, n), jb = min(ja + seg, n);
        double te = __longlong_as_double(0xfff0000000000000ll);
        int ti = 1 << 30, nm = 0;
        if (ja < jb) fq_scan(S.et[wid], S.ys[wid], errs_all + pbase, n, ja, jb, prm, lane, te, ti, nm);
        if (lane < 11) { S.lte[wid][lane] = te; S.lti[wid][lane] = ti; }
        if (lane == 0) S.lnm[wid] = nm;
        __syncthreads();
        if (wid != 0) continue;
        // merge: warp 0's own list is the start; the others' entries are inserted in warp order, best first
        for (int w = 1; w < FQB_WARPS; w++) {
            nm += S.lnm[w];
            for (int q = 0; q < 11; q++) {
                const int ci = S.lti[w][q];
                if (ci == (1 << 30)) break;                       // the rest of this list is empty
                const double ev = S.lte[w][q];
                const double t11 = __shfl_sync(full, te, 10);
                if (!(ev > t11)) break;                           // (the list is descending: nothing further down can enter either)
                const int pos = __popc(__ballot_sync(full, lane < 11 && te >= ev));
                const double up_te = __shfl_up_sync(full, te, 1);
                const int up_ti = __shfl_up_sync(full, ti, 1);
                if (lane == pos) { te = ev; ti = ci; }
                else if (lane > pos) { te = up_te; ti = up_ti; }
            }
        }
        if (min(20, n / 12) < 2) continue;
        fq_finish(S.w0, te, ti, nm, n, b, rec, pbase, in, sorted_xy, cp_all, quads, nquads, nquads_total, errflag, g, caps, prm, lane);
    }
}

// ---- prefix moments + window errors for every tier: ONE WARP per cluster ------------------------------------------------
// upstream compute_lfps() + the first loop of quad_segment_maxima().  The prefix is a dependent chain (one DADD per point
// and moment) that only six lanes can work on, so what matters is the number of clusters in flight, not the threads per
// cluster.  Writing the whole 48-byte-per-point prefix array and reading it back for the window errors was the largest
// HBM stream of the detector, so the array never leaves the SM:
//   * the prefix of the last 96 points lives in a shared-memory ring (moment-major, pitch 98: the six chain lanes and the
//     32 window lanes are both bank-conflict free); the window error errs[i] = fit_line(i - ksz, i + ksz) only needs
//     entries i + ksz and i - ksz - 1, i.e. at most 41 back, and is emitted as soon as entry i + ksz exists;
//   * the 2 ksz windows that wrap around the ends are done last, from the ring (tail) and a copy of the first 2 ksz
//     entries (head);
//   * every LF_CP-th entry is written to global memory as a checkpoint (6 bytes per point); the fit kernel rebuilds the
//     <= 21 entries the corner search needs by replaying <= LF_CP - 1 points from a checkpoint.
// The gradient taps of the next 32 points are issued before the current 32 are accumulated.
constexpr int LF_WARPS = 4;
constexpr int LF_RING = 96, LF_PITCH = 98, LF_HEAD = 40, LF_HPITCH = 41;   // pitch 98: rows 16-byte aligned, chain lanes 4 banks apart
struct __align__(16) LfWarp {
    double ring[6][LF_PITCH];
    double head[6][LF_HPITCH];
};
__global__ void __launch_bounds__(LF_WARPS * 32)
lfps_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ sorted_xy, const ClusterRec *__restrict__ clusters,
            const uint32_t *__restrict__ worklists, size_t wl_stride, const uint32_t *__restrict__ nwork /* stride 2, 4 tiers */,
            uint32_t *__restrict__ work_counter, double *__restrict__ errs_all, double *__restrict__ cp_all, Geom g, Caps caps, int n_big)
{
    __shared__ LfWarp sh[LF_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu;
    double (*ring)[LF_PITCH] = sh[wid].ring;
    double (*head)[LF_HPITCH] = sh[wid].head;
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(full, wi, 0);
        uint32_t item;
        if (!tier_item<0, 3>(wi, nwork, worklists, wl_stride, item)) return;
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        if (rec.cursor == 0xffffffffu || rec.count < 24 || (int)rec.count >= n_big) continue;      // (>= n_big: lfps_big_kernel)
        const int n = (int)rec.count;
        const int ksz = min(20, n / 12);
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;     // multiple of LF_CP
        const uint32_t *XY = sorted_xy + pbase;
        double *errs = errs_all + pbase;
        double *cp = cp_all + (pbase / LF_CP) * 6;
        const uint8_t *img = in + (size_t)b * g.frame_stride;
        // raw taps of one point: left/right/up/down neighbours of the decimated pixel, -1 = no gradient (W = 1)
        int gl = -1, gr = 0, gu = 0, gd = 0;
        // xy: the point whose taps are in flight; xy_ahead: the point one block further (its load is issued a whole block
        // before its taps need it, so the tap addresses never wait for it)
        uint32_t xy = 0, xy_ahead = lane < n ? XY[lane] : 0u;
        auto issue = [&](int j) {
            gl = -1;
            xy = xy_ahead;
            if (j + 32 < n) xy_ahead = XY[j + 32];
            if (j < n) {
                const int px = (int)(xy & 0xffff), py = (int)(xy >> 16);
                const double x = px * .5 + 0.5, y = py * .5 + 0.5;
                const int ix = (int)x, iy = (int)y;
                if (ix > 0 && ix + 1 < g.w && iy > 0 && iy + 1 < g.h) {
                    const uint8_t *row = img + (size_t)(iy * g.f) * g.stride;
                    gl = row[(ix - 1) * g.f]; gr = row[(ix + 1) * g.f];
                    gu = row[(ptrdiff_t)ix * g.f - (ptrdiff_t)g.f * g.stride]; gd = row[(ptrdiff_t)ix * g.f + (ptrdiff_t)g.f * g.stride];
                }
            }
        };
        auto ring_entry = [&](int idx, M6 &e) {
            const int s = idx % LF_RING;
            e.Mx = ring[0][s]; e.My = ring[1][s]; e.Mxx = ring[2][s]; e.Mxy = ring[3][s]; e.Myy = ring[4][s]; e.W = ring[5][s];
        };
        issue(lane);
        double acc = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane, s0 = j0 % LF_RING, s = s0 + lane;
            if (j < n) {
                double W = 1;
                if (gl >= 0) {
                    const int grad_x = gr - gl, grad_y = gd - gu;
                    W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
                }
                const int px = (int)(xy & 0xffff), py = (int)(xy >> 16);
                const double fx = px * .5 + 0.5, fy = py * .5 + 0.5;
                ring[0][s] = W * fx; ring[1][s] = W * fy; ring[2][s] = W * fx * fx; ring[3][s] = W * fx * fy; ring[4][s] = W * fy * fy; ring[5][s] = W;
            }
            __syncwarp();
            issue(j0 + 32 + lane);                   // gathers of the next block fly while this block is accumulated
            if (lane < 6) {
                const int cnt = min(32, n - j0);
                double *r = &ring[lane][s0];
                double *c = cp + (size_t)(j0 / LF_CP) * 6 + lane;
                if (cnt == 32) {
                    // full block: 128-bit shared-memory accesses, the additions stay one dependent chain
                    double2 *r2 = reinterpret_cast<double2 *>(r);
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        double2 v = r2[k];
                        acc += v.x; v.x = acc;
                        acc += v.y; v.y = acc;
                        r2[k] = v;
                        if ((k & (LF_CP / 2 - 1)) == LF_CP / 2 - 1) c[(size_t)(k / (LF_CP / 2)) * 6] = acc;
                    }
                } else {
#pragma unroll 8
                    for (int k = 0; k < cnt; k++) {
                        acc += r[k];
                        r[k] = acc;
                        if ((k & (LF_CP - 1)) == LF_CP - 1) c[(size_t)(k / LF_CP) * 6] = acc;
                    }
                }
            }
            __syncwarp();
            if (j < n) {
                if (j < 2 * ksz) {
#pragma unroll
                    for (int m = 0; m < 6; m++) head[m][j] = ring[m][s];
                } else {
                    // window centred on i = j - ksz: i0 = j - 2 ksz >= 0, i1 = j
                    M6 a, p;
                    ring_entry(j, a);
                    p = a;
                    const int i0 = j - 2 * ksz;
                    if (i0 > 0) ring_entry(i0 - 1, p);
                    LineFit lf;
                    fit_line_m(a, p, p, i0 > 0 ? 1 : 0, 2 * ksz + 1, false, lf);
                    errs[j - ksz] = lf.err;
                }
            }
            __syncwarp();                            // the next block overwrites the oldest third of the ring
        }
        // windows that wrap: i in [0, ksz) and [n - ksz, n)
        for (int t = lane; t < 2 * ksz; t += 32) {
            const int i = t < ksz ? t : n - 2 * ksz + t;
            int i0 = i - ksz; if (i0 < 0) i0 += n;
            int i1 = i + ksz; if (i1 >= n) i1 -= n;
            M6 a, p, l;
            a.Mx = head[0][i1]; a.My = head[1][i1]; a.Mxx = head[2][i1]; a.Mxy = head[3][i1]; a.Myy = head[4][i1]; a.W = head[5][i1];
            ring_entry(i0 - 1, p);
            ring_entry(n - 1, l);
            LineFit lf;
            fit_line_m(a, p, l, 2, n - i0 + i1 + 1, false, lf);
            errs[i] = lf.err;
        }
        __syncwarp();
    }
}

// ---- prefix moments + window errors, ONE CTA per LARGE cluster (small batches) -----------------------------------------------
// With one warp per cluster the largest cluster of a frame is a serial chain: per 32 points the warp computes the terms, runs
// the six-lane additions, then the window errors, one phase after the other -- 1 us per step, 168 us for the 5 000-point cluster
// of a 1280x720 frame, a quarter of that frame's latency.  Only the additions are sequential.  Here a CTA streams the cluster in
// chunks of LFB_C points through a four-chunk ring in shared memory as a three-stage pipeline, one __syncthreads per chunk:
//   workers (7 warps)  A(c): terms of chunk c (gradient taps, square root) -> ring
//   chain   (1 warp)   B(c - 1): the six prefix sums over chunk c - 1, in place, in point order (one DADD per point and moment)
//   workers            C(c - 2): window errors for the points whose window ends in chunk c - 2 (they read back at most 41 entries:
//                      chunks c - 2 and c - 3), checkpoints of that chunk, copy of the first 2 ksz entries
// so the cluster costs about its chain: 10 cycles per point.  Every value is computed by the same expressions in the same order as
// in lfps_kernel (same terms, same additions, same fit_line_m calls), so errs[] and the checkpoints are identical.  Used for the
// clusters of at least n_big points when the batch is small (api.cu); lfps_kernel skips those.
constexpr int LFB_THREADS = 256, LFB_C = 256, LFB_RING = 4 * LFB_C, LFB_PITCH = LFB_RING + 2;     // pitch: chain lanes 4 banks apart
constexpr int LFB_SMEM = (6 * LFB_PITCH + 6 * LF_HPITCH) * (int)sizeof(double) + 16;
template <int T_LO>
__global__ void __launch_bounds__(LFB_THREADS)
lfps_big_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ sorted_xy, const ClusterRec *__restrict__ clusters,
                const uint32_t *__restrict__ worklists, size_t wl_stride, const uint32_t *__restrict__ nwork /* stride 2, 4 tiers */,
                uint32_t *__restrict__ work_counter, double *__restrict__ errs_all, double *__restrict__ cp_all, Geom g, Caps caps, int n_big)
{
    extern __shared__ __align__(16) unsigned char lfb_smem[];
    double (*ring)[LFB_PITCH] = reinterpret_cast<double (*)[LFB_PITCH]>(lfb_smem);
    double (*head)[LF_HPITCH] = reinterpret_cast<double (*)[LF_HPITCH]>(lfb_smem + 6 * LFB_PITCH * sizeof(double));
    uint32_t *s_work = reinterpret_cast<uint32_t *>(lfb_smem + (6 * LFB_PITCH + 6 * LF_HPITCH) * sizeof(double));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NWORK = LFB_THREADS - 32;                 // worker threads (warps 1..7)
    for (;;) {
        __syncthreads();
        if (tid == 0) *s_work = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t wi = *s_work;
        uint32_t item;
        if (!tier_item<T_LO, 3>(wi, nwork, worklists, wl_stride, item)) return;
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        if (rec.cursor == 0xffffffffu || (int)rec.count < n_big || rec.count < 24) continue;
        const int n = (int)rec.count;
        const int ksz = min(20, n / 12);
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;     // multiple of LF_CP
        const uint32_t *XY = sorted_xy + pbase;
        double *errs = errs_all + pbase;
        double *cp = cp_all + (pbase / LF_CP) * 6;
        const uint8_t *img = in + (size_t)b * g.frame_stride;
        const int nchunks = (n + LFB_C - 1) / LFB_C;
        auto ring_entry = [&](int idx, M6 &e) {
            const int s = idx % LFB_RING;
            e.Mx = ring[0][s]; e.My = ring[1][s]; e.Mxx = ring[2][s]; e.Mxy = ring[3][s]; e.Myy = ring[4][s]; e.W = ring[5][s];
        };
        double acc = 0;                                     // chain lanes: running sum of moment `lane`
        for (int it = 0; it < nchunks + 2; it++) {
            if (wid == 0) {
                // ---- B(it - 1): the chain ----
                const int c = it - 1;
                if (c >= 0 && c < nchunks && lane < 6) {
                    const int cnt = min(LFB_C, n - c * LFB_C);
                    double *r = &ring[lane][(c * LFB_C) % LFB_RING];
                    int k = 0;
                    double2 *r2 = reinterpret_cast<double2 *>(r);
                    for (; k + 8 <= cnt; k += 8) {              // 128-bit shared-memory accesses, the additions stay one dependent chain
                        double2 v0 = r2[k / 2], v1 = r2[k / 2 + 1], v2 = r2[k / 2 + 2], v3 = r2[k / 2 + 3];
                        acc += v0.x; v0.x = acc; acc += v0.y; v0.y = acc;
                        acc += v1.x; v1.x = acc; acc += v1.y; v1.y = acc;
                        acc += v2.x; v2.x = acc; acc += v2.y; v2.y = acc;
                        acc += v3.x; v3.x = acc; acc += v3.y; v3.y = acc;
                        r2[k / 2] = v0; r2[k / 2 + 1] = v1; r2[k / 2 + 2] = v2; r2[k / 2 + 3] = v3;
                    }
                    for (; k < cnt; k++) { acc += r[k]; r[k] = acc; }
                }
            } else {
                const int wt = tid - 32;
                // ---- A(it): terms of chunk it ----
                if (it < nchunks) {
                    const int cnt = min(LFB_C, n - it * LFB_C);
                    for (int p = wt; p < cnt; p += NWORK) {
                        const int j = it * LFB_C + p, s = j % LFB_RING;
                        const uint32_t xy = XY[j];
                        const int px = (int)(xy & 0xffff), py = (int)(xy >> 16);
                        const double fx = px * .5 + 0.5, fy = py * .5 + 0.5;
                        const int ix = (int)fx, iy = (int)fy;
                        double W = 1;
                        if (ix > 0 && ix + 1 < g.w && iy > 0 && iy + 1 < g.h) {
                            const uint8_t *row = img + (size_t)(iy * g.f) * g.stride;
                            const int gl = row[(ix - 1) * g.f], gr = row[(ix + 1) * g.f];
                            const int gu = row[(ptrdiff_t)ix * g.f - (ptrdiff_t)g.f * g.stride], gd = row[(ptrdiff_t)ix * g.f + (ptrdiff_t)g.f * g.stride];
                            const int grad_x = gr - gl, grad_y = gd - gu;
                            W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
                        }
                        ring[0][s] = W * fx; ring[1][s] = W * fy; ring[2][s] = W * fx * fx; ring[3][s] = W * fx * fy; ring[4][s] = W * fy * fy; ring[5][s] = W;
                    }
                }
                // ---- C(it - 2): everything that reads the finished prefix of chunk it - 2 ----
                const int c = it - 2;
                if (c >= 0) {
                    const int cnt = min(LFB_C, n - c * LFB_C);
                    for (int q = wt; q < cnt; q += NWORK) {
                        const int j = c * LFB_C + q, s = j % LFB_RING;
                        if ((j & (LF_CP - 1)) == LF_CP - 1) {
#pragma unroll
                            for (int m = 0; m < 6; m++) cp[(size_t)(j / LF_CP) * 6 + m] = ring[m][s];
                        }
                        if (j < 2 * ksz) {
#pragma unroll
                            for (int m = 0; m < 6; m++) head[m][j] = ring[m][s];
                        } else {
                            // window centred on i = j - ksz: i0 = j - 2 ksz >= 0, i1 = j
                            M6 a, p;
                            ring_entry(j, a);
                            p = a;
                            const int i0 = j - 2 * ksz;
                            if (i0 > 0) ring_entry(i0 - 1, p);
                            LineFit lf;
                            fit_line_m(a, p, p, i0 > 0 ? 1 : 0, 2 * ksz + 1, false, lf);
                            errs[j - ksz] = lf.err;
                        }
                    }
                }
            }
            __syncthreads();
        }
        // windows that wrap: i in [0, ksz) and [n - ksz, n)
        for (int t = tid; t < 2 * ksz; t += LFB_THREADS) {
            const int i = t < ksz ? t : n - 2 * ksz + t;
            int i0 = i - ksz; if (i0 < 0) i0 += n;
            int i1 = i + ksz; if (i1 >= n) i1 -= n;
            M6 a, p, l;
            a.Mx = head[0][i1]; a.My = head[1][i1]; a.Mxx = head[2][i1]; a.Mxy = head[3][i1]; a.Myy = head[4][i1]; a.W = head[5][i1];
            ring_entry(i0 - 1, p);
            ring_entry(n - 1, l);
            LineFit lf;
            fit_line_m(a, p, l, 2, n - i0 + i1 + 1, false, lf);
            errs[i] = lf.err;
        }
    }
}

// tier limits of the four work lists, and the sort kernels' (threads, elements per thread, shared-memory points) per tier
constexpr int QT0 = 256, QT1 = 512, QT2 = 2048;
// the five kernels of each sort: <NT, E, MAXN, WHICH, T_LO, T_HI, NMIN, NMAX>
template <int W> using SortS8 = SortCfg<32, 8, QT0, W, 0, 0, 0, QT0>;
template <int W> using SortS16 = SortCfg<32, 16, QT1, W, 1, 1, 0, QT1>;
template <int W> using SortM = SortCfg<128, 16, QT2, W, 2, 2, 0, QT2>;
template <int W> using SortL1 = SortCfg<256, 16, 4096, W, 3, 3, 0, 4096>;
template <int W> using SortL2 = SortCfg<512, 16, 8192, W, 3, 3, 4097, (1 << 30)>;

}  // namespace cb
