// quads.cuh -- row A5 of SURVEY.md 8a: fit_quads() / fit_quad().
//
// Upstream (apriltag_quad_thresh.c fit_quad, ptsort, compute_lfps, fit_line, quad_segment_maxima).  The float /
// double arithmetic below follows upstream operation by operation (the library is built with --fmad=false), and
// every order-dependent accumulation keeps upstream's order:
//   * points are first put back into scan order (y, x, probe) -- the order upstream's hash map appends them in;
//   * ptsort()'s merge sort is emulated exactly: same recursive split (sz/2), same 2..5 element sorting networks
//     at the leaves, merges that take from the SECOND half on ties -- done as parallel rank merges;
//   * the line-fit prefix moments are accumulated sequentially (six lanes, one per moment), fed 32 points at a time
//     through a small shared-memory staging tile so that the dependent chain is one DADD per point;
//   * the 4-corner search evaluates all <=210 subsets in parallel and keeps the first minimum in loop order.
//
// B200 mapping: the work is thousands of small, irregular, partly sequential jobs per frame (a c2 frame has ~700
// clusters of 24..7000 points), so the design maximises the number of clusters in flight instead of the threads per
// cluster.  Two persistent kernels pull (frame, cluster) items from device-side work lists:
//   tier S: clusters of <= 256 points, ONE WARP per cluster (6 KB smem per warp, 32 warps per SM);
//   tier M: 257..2048 points, one 128-thread CTA per cluster (34 KB smem, 6 CTAs per SM);
//   tier L: larger clusters, one 256-thread CTA per cluster (98 KB smem, 2 CTAs per SM; clusters above 6144 points
//           run the same code out of a global scratch area).
// A boundary point is fully described by its 32-bit scan key (pixel index, probe, gradient sign), so the only
// per-point input is 4 bytes; sorting happens on packed (slope, scan key) 64-bit words in shared memory.
#pragma once
#include "common.cuh"
#include "sort.cuh"

namespace cb {

constexpr int QS_MAXN = 256;      // tier S: points per warp
constexpr int QS_WARPS = 8;
constexpr int QL_THREADS = 256;
constexpr int QL_MAXN = 6144;     // tier L: points per CTA in shared memory (tier M: 128 threads, 2048 points)

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
    const double Ex = Mx / W, Ey = My / W;
    const double Cxx = Mxx / W - Ex * Ex, Cxy = Mxy / W - Ex * Ey, Cyy = Myy / W - Ey * Ey;
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
        else { o.nx = nx / length; o.ny = ny / length; }
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

struct QfScratch {       // per group (warp or CTA)
    double red_d[8];
    int red_i[8];
    float red_f[8];
    uint32_t w_cnt[8];
    int nmax;
    int kept[16];
    int nkept;
    int taken[16];
    double thresh;
    M6 ent[21];               // prefix-moment entries of the kept maxima
};

// Processes one cluster.  A and B are 8-byte-per-point work arrays (shared or global), lfps the 48-byte-per-point
// prefix-moment array (global).  Every thread of the group must call it; control flow is group-uniform.
// The work on a cluster is cut into four phases, each its own kernel (per tier), because the one-warp-per-cluster tier is
// instruction-fetch bound when every warp of an SM sits in a different part of a 100 KB kernel (ncu: stall_no_instruction):
//   PHASE 1: bounding box, border polarity (rejected clusters get cursor = 0xffffffff), scan-order sort; the sorted scan
//            keys replace the unsorted ones in global memory;
//   PHASE 2: slopes, ptsort() emulation; the sorted points (px | py << 16) replace the keys;
//   PHASE 3: per-point weights and the sequential prefix moments (global lfps array);
//   PHASE 4: window errors, maxima, corner search, corner / area / convexity tests.
template <int NT, int PHASE>
__device__ void fit_quad_cluster(const uint8_t *__restrict__ img, uint32_t *__restrict__ K, int n, unsigned long long *A,
                                 unsigned long long *B, const double *__restrict__ errs_g, const double *__restrict__ cp, QfScratch &S, const ClusterRec &rec,
                                 ClusterRec *__restrict__ rec_global, int b,
                                 QuadRec *__restrict__ quads, uint32_t *__restrict__ nquads, uint32_t *__restrict__ nquads_total,
                                 uint32_t *__restrict__ errflag, const Geom &g, const Caps &caps, const DetParams &prm)
{
    typedef Grp<NT> G;
    const int tid = G::tid();
    const int lane = threadIdx.x & 31;

    const int reversed_border = 0;          // reversed clusters never get past sort #1 (tag36h11 has a normal border only)
    unsigned long long *src = A, *dst = B;
    // ---- PHASE 4: quad_segment_maxima ---------------------------------------------------------------------------
    const int ksz = min(20, n / 12);
    if (ksz < 2) return;
    double *errs = reinterpret_cast<double *>(src);
    double *ysm = reinterpret_cast<double *>(dst);
    for (int i = tid; i < n; i += NT) errs[i] = errs_g[i];      // window errors (lfps_kernel)
    G::sync();
    for (int iy = tid; iy < n; iy += NT) {
        double acc = 0;
#pragma unroll
        for (int i = 0; i < 7; i++) {
            int q = iy + i - 3;
            if (q < 0) q += n; else if (q >= n) q -= n;
            acc += errs[q] * prm.smooth_f[i];
        }
        ysm[iy] = acc;
    }
    G::sync();
    // local maxima, collected in index order (two maxima are never adjacent, so there are at most n/2)
    int *maxima = reinterpret_cast<int *>(errs);
    double *maxima_errs = reinterpret_cast<double *>(errs) + (n + 3) / 4;
    if (tid == 0) S.nmax = 0;
    G::sync();
    for (int i0 = 0; i0 < n; i0 += NT) {
        const int i = i0 + tid;
        bool is_max = false;
        double e = 0;
        if (i < n) {
            e = ysm[i];
            const int nx = i + 1 == n ? 0 : i + 1, pv = i == 0 ? n - 1 : i - 1;
            is_max = e > ysm[nx] && e > ysm[pv];
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, is_max);
        int base;
        if (NT == 32) {
            base = S.nmax;
        } else {
            const int wid = threadIdx.x >> 5;
            if (lane == 0) S.w_cnt[wid] = __popc(bal);
            __syncthreads();
            base = S.nmax;
            for (int k = 0; k < wid; k++) base += S.w_cnt[k];
        }
        if (is_max) {
            const int pos = base + __popc(bal & ((1u << lane) - 1));
            maxima[pos] = i;
            maxima_errs[pos] = e;
        }
        G::sync();
        if (tid == 0) {
            if (NT == 32) S.nmax += __popc(bal);
            else { int t = 0; for (int k = 0; k < NT / 32; k++) t += S.w_cnt[k]; S.nmax += t; }
        }
        G::sync();
    }
    const int nmaxima = S.nmax;
    if (nmaxima < 4) return;
    // keep only the best max_nmaxima: maxima_thresh = element [max_nmaxima] of the descending sort, found by
    // max_nmaxima + 1 rounds of "largest not yet taken" (ties -> lowest position, one element per round)
    const int max_nmaxima = min(prm.max_nmaxima, 10);
    if (nmaxima > max_nmaxima) {
        for (int r = 0; r <= max_nmaxima; r++) {
            double best = 0;
            int bpos = 1 << 30;       // 1<<30 = nothing found yet
            for (int m = tid; m < nmaxima; m += NT) {
                bool tk = false;
                for (int q = 0; q < r; q++) tk |= (S.taken[q] == m);
                if (tk) continue;
                const double e = maxima_errs[m];
                if (bpos == (1 << 30) || e > best) { best = e; bpos = m; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
                if (op != (1 << 30) && (bpos == (1 << 30) || ob > best || (ob == best && op < bpos))) { best = ob; bpos = op; }
            }
            if (NT != 32) {
                const int wid = threadIdx.x >> 5;
                __syncthreads();
                if (lane == 0) { S.red_d[wid] = best; S.red_i[wid] = bpos; }
                __syncthreads();
                best = S.red_d[0]; bpos = S.red_i[0];
                for (int k = 1; k < NT / 32; k++) {
                    const double ob = S.red_d[k]; const int op = S.red_i[k];
                    if (op != (1 << 30) && (bpos == (1 << 30) || ob > best || (ob == best && op < bpos))) { best = ob; bpos = op; }
                }
            }
            G::sync();
            if (tid == 0) { S.taken[r] = bpos; S.thresh = best; }
            G::sync();
        }
        if (tid == 0) {
            int out = 0;
            const double th = S.thresh;
            // the kept maxima are among the positions taken in the first max_nmaxima rounds; emit them in index order
            int pos[10];
            for (int q = 0; q < max_nmaxima; q++) pos[q] = S.taken[q];
            for (int a = 1; a < max_nmaxima; a++) { const int t = pos[a]; int c = a; while (c > 0 && pos[c - 1] > t) { pos[c] = pos[c - 1]; c--; } pos[c] = t; }
            for (int q = 0; q < max_nmaxima; q++) {
                if (maxima_errs[pos[q]] <= th) continue;
                S.kept[out++] = maxima[pos[q]];
            }
            S.nkept = out;
        }
    } else if (tid == 0) {
        for (int m = 0; m < nmaxima; m++) S.kept[m] = maxima[m];
        S.nkept = nmaxima;
    }
    G::sync();
    const int nk = S.nkept;
    if (nk < 4) return;   // (upstream's loops would simply find nothing)
    // pair table: fit_line(kept[a], kept[c]) for a != c (3200 B); lives in the now dead work arrays
    double (*p_err)[10] = reinterpret_cast<double (*)[10]>(A);    // A and B are adjacent (A first): >= 4 KB in every tier
    double (*p_mse)[10] = p_err + 10, (*p_nx)[10] = p_err + 20, (*p_ny)[10] = p_err + 30;
    // the prefix-moment entries the corner search reads: lfps[kept[m]], lfps[kept[m] - 1], lfps[n - 1]
    if (tid <= 2 * nk) {
        const int idx = tid == 2 * nk ? n - 1 : S.kept[tid >> 1] - (tid & 1);
        if (idx >= 0) {
            double acc[6];
            replay_entry(img, K, cp, idx, g, acc);
            M6 &e = S.ent[tid];
            e.Mx = acc[0]; e.My = acc[1]; e.Mxx = acc[2]; e.Mxy = acc[3]; e.Myy = acc[4]; e.W = acc[5];
        }
    }
    G::sync();
    // fit_line(lfps, n, kept[a], kept[c]) on the cached entries
    auto fit_kept = [&](int a, int c, LineFit &lf) {
        const int i0 = S.kept[a], i1 = S.kept[c];
        if (i0 < i1) fit_line_m(S.ent[2 * c], S.ent[2 * a + 1], S.ent[2 * nk], i0 > 0 ? 1 : 0, i1 - i0 + 1, true, lf);
        else fit_line_m(S.ent[2 * c], S.ent[2 * a + 1], S.ent[2 * nk], 2, n - i0 + i1 + 1, true, lf);
    };
    for (int t = tid; t < nk * nk; t += NT) {
        const int a = t / nk, c = t % nk;
        if (a == c) continue;
        LineFit lf;
        fit_kept(a, c, lf);
        p_err[a][c] = lf.err; p_mse[a][c] = lf.mse; p_nx[a][c] = lf.nx; p_ny[a][c] = lf.ny;
    }
    G::sync();
    // 4-corner search; (m0,m1,m2,m3) packed big-endian orders like upstream's loop nest, so the minimum over
    // (err, packed) is upstream's "first minimum"
    double best_err = __longlong_as_double(0x7ff0000000000000ll);
    int best_combo = 1 << 30;
    {
        const double max_mse = (double)prm.max_line_fit_mse;
        const int ncomb = nk * (nk - 1) * (nk - 2) * (nk - 3) / 24;
        for (int ci = tid; ci < ncomb; ci += NT) {
            const int pk = c_combos[ci];
            const int m0 = pk >> 12, m1 = (pk >> 8) & 15, m2 = (pk >> 4) & 15, m3 = pk & 15;
            if (p_mse[m0][m1] > max_mse) continue;
            if (p_mse[m1][m2] > max_mse) continue;
            const double dt = p_nx[m0][m1] * p_nx[m1][m2] + p_ny[m0][m1] * p_ny[m1][m2];
            if (fabs(dt) > prm.cos_critical_rad) continue;
            if (p_mse[m2][m3] > max_mse) continue;
            if (p_mse[m3][m0] > max_mse) continue;
            const double err = p_err[m0][m1] + p_err[m1][m2] + p_err[m2][m3] + p_err[m3][m0];
            if (err < best_err || (err == best_err && pk < best_combo)) { best_err = err; best_combo = pk; }
        }
    }
    {
        const double bmin = G::reduce(best_err, [](double a, double c) { return a < c ? a : c; }, S.red_d);
        int cand = (best_err == bmin && best_combo != (1 << 30)) ? best_combo : (1 << 30);
        cand = G::reduce(cand, [](int a, int c) { return min(a, c); }, S.red_i);
        best_err = bmin; best_combo = cand;
    }
    if (best_combo == (1 << 30)) return;
    if (!(best_err / n < (double)prm.max_line_fit_mse)) return;

    // ---- corners, area and convexity tests (one thread) -------------------------------------------------------------
    if (tid == 0) {
        const int mi[4] = {(best_combo >> 12) & 15, (best_combo >> 8) & 15, (best_combo >> 4) & 15, best_combo & 15};
        double lines[4][4];
        bool good = true;
        for (int i = 0; i < 4 && good; i++) {
            LineFit lf;
            fit_kept(mi[i], mi[(i + 1) & 3], lf);
            lines[i][0] = lf.Ex; lines[i][1] = lf.Ey; lines[i][2] = lf.nx; lines[i][3] = lf.ny;
            if (lf.mse > (double)prm.max_line_fit_mse) good = false;
        }
        float qp[4][2];
        for (int i = 0; i < 4 && good; i++) {
            const double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
            const double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
            const double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
            const double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
            const double det = A00 * A11 - A10 * A01;
            const double W00 = A11 / det, W01 = -A01 / det;
            if (fabs(det) < 0.001) { good = false; break; }
            const double L0 = W00 * B0 + W01 * B1;
            qp[i][0] = (float)(lines[i][0] + L0 * A00);
            qp[i][1] = (float)(lines[i][1] + L0 * A10);
        }
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
            if (qf >= caps.quads_per_frame) atomicOr(errflag, ERR_QUADS_FULL);
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

struct QsWarp {
    unsigned long long A[QS_MAXN];
    unsigned long long B[QS_MAXN];   // directly after A: the pair table of the corner search spans both
    QfScratch S;
};
struct QsShared { QsWarp w[QS_WARPS]; };

// tier S: persistent warps, one cluster (<= QS_MAXN points, work list 0) per warp at a time
template <int PHASE>
__global__ void __launch_bounds__(QS_WARPS * 32)
fit_quads_small_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ scankey, ClusterRec *__restrict__ clusters,
                       const uint32_t *__restrict__ worklist, const uint32_t *__restrict__ nwork, uint32_t *__restrict__ work_counter,
                       const double *__restrict__ errs_all, const double *__restrict__ cp_all, QuadRec *__restrict__ quads, uint32_t *__restrict__ nquads,
                       uint32_t *__restrict__ nquads_total, uint32_t *__restrict__ errflag, Geom g, Caps caps, DetParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QsShared &SH = *reinterpret_cast<QsShared *>(smem_raw);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t total = *nwork;
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= total) return;
        const uint32_t item = worklist[wi];
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        const int n = (int)rec.count;
        if (n < 24 || n > QS_MAXN) continue;
        if (rec.cursor == 0xffffffffu) continue;
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;
        fit_quad_cluster<32, PHASE>(in + (size_t)b * g.frame_stride, scankey + pbase, n, SH.w[wid].A, SH.w[wid].B, errs_all + pbase, cp_all + (pbase / LF_CP) * 6, SH.w[wid].S, rec,
                                    clusters + item, b,
                             quads, nquads, nquads_total, errflag, g, caps, prm);
        __syncwarp();
    }
}

template <int MAXN>
struct QlShared {
    unsigned long long A[MAXN];
    unsigned long long B[MAXN];    // directly after A
    QfScratch S;
    int work;
};

// tiers M / L: persistent CTAs of NT threads, one cluster from work lists T_LO..T_HI per CTA at a time; clusters above
// MAXN points run out of the global scratch area.  Tier M (NT = 128, MAXN = 2048, 34 KB smem) keeps 6 CTAs per SM
// resident, tier L (NT = 256, MAXN = 6144, 98 KB) two.
template <int NT, int MAXN, int PHASE, int T_LO, int T_HI>
__global__ void __launch_bounds__(NT)
fit_quads_cta_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ scankey, ClusterRec *__restrict__ clusters,
                     const uint32_t *__restrict__ worklists, size_t wl_stride, const uint32_t *__restrict__ nwork,
                     uint32_t *__restrict__ work_counter,
                     const double *__restrict__ errs_all, const double *__restrict__ cp_all, unsigned long long *__restrict__ scratch, QuadRec *__restrict__ quads,
                     uint32_t *__restrict__ nquads, uint32_t *__restrict__ nquads_total, uint32_t *__restrict__ errflag, Geom g, Caps caps,
                     DetParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QlShared<MAXN> &SH = *reinterpret_cast<QlShared<MAXN> *>(smem_raw);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) SH.work = (int)atomicAdd(work_counter, 1u);
        __syncthreads();
        uint32_t item;
        if (!tier_item<T_LO, T_HI>((uint32_t)SH.work, nwork, worklists, wl_stride, item)) return;
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        const int n = (int)rec.count;
        if (n < 24) continue;
        if (rec.cursor == 0xffffffffu) continue;
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;
        unsigned long long *A = SH.A, *B = SH.B;
        if (n > MAXN) { A = scratch + pbase * 2; B = A + n; }
        fit_quad_cluster<NT, PHASE>(in + (size_t)b * g.frame_stride, scankey + pbase, n, A, B, errs_all + pbase, cp_all + (pbase / LF_CP) * 6, SH.S, rec, clusters + item, b, quads, nquads,
                             nquads_total, errflag, g, caps, prm);
    }
}

// ---- prefix moments + window errors for every tier: ONE WARP per cluster ------------------------------------------------
// upstream compute_lfps() + the first loop of quad_segment_maxima().  The prefix is a dependent chain (one DADD per point
// and moment) that only six lanes can work on, so what matters is the number of clusters in flight, not the threads per
// cluster.  Writing the whole 48-byte-per-point prefix array and reading it back for the window errors was the largest
// HBM stream of the detector, so the array never leaves the SM:
//   * the prefix of the last 96 points lives in a shared-memory ring (moment-major, pitch 97: the six chain lanes and the
//     32 window lanes are both bank-conflict free); the window error errs[i] = fit_line(i - ksz, i + ksz) only needs
//     entries i + ksz and i - ksz - 1, i.e. at most 41 back, and is emitted as soon as entry i + ksz exists;
//   * the 2 ksz windows that wrap around the ends are done last, from the ring (tail) and a copy of the first 2 ksz
//     entries (head);
//   * every LF_CP-th entry is written to global memory as a checkpoint (6 bytes per point); the fit kernel rebuilds the
//     <= 21 entries the corner search needs by replaying <= LF_CP - 1 points from a checkpoint.
// The gradient taps of the next 32 points are issued before the current 32 are accumulated.
constexpr int LF_WARPS = 4;
constexpr int LF_RING = 96, LF_PITCH = 97, LF_HEAD = 40, LF_HPITCH = 41;
struct LfWarp {
    double ring[6][LF_PITCH];
    double head[6][LF_HPITCH];
};
__global__ void __launch_bounds__(LF_WARPS * 32)
lfps_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ sorted_xy, const ClusterRec *__restrict__ clusters,
            const uint32_t *__restrict__ worklists, size_t wl_stride, const uint32_t *__restrict__ nwork /* stride 2, 4 tiers */,
            uint32_t *__restrict__ work_counter, double *__restrict__ errs_all, double *__restrict__ cp_all, Geom g, Caps caps)
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
        if (rec.cursor == 0xffffffffu || rec.count < 24) continue;
        const int n = (int)rec.count;
        const int ksz = min(20, n / 12);
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;     // multiple of LF_CP
        const uint32_t *XY = sorted_xy + pbase;
        double *errs = errs_all + pbase;
        double *cp = cp_all + (pbase / LF_CP) * 6;
        const uint8_t *img = in + (size_t)b * g.frame_stride;
        // raw taps of one point: left/right/up/down neighbours of the decimated pixel, -1 = no gradient (W = 1)
        int gl = -1, gr = 0, gu = 0, gd = 0;
        uint32_t xy = 0;
        auto issue = [&](int j) {
            gl = -1;
            if (j < n) {
                xy = XY[j];
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
#pragma unroll 8
                for (int k = 0; k < cnt; k++) {
                    acc += r[k];
                    r[k] = acc;
                    if ((k & (LF_CP - 1)) == LF_CP - 1) c[(size_t)(k / LF_CP) * 6] = acc;
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

// tiers: S (one warp, <= QS_MAXN) | M1 (128 threads, <= 2048) | M2 (spare slot: same bound as M1, so it stays empty) | L (256 threads, the rest).
// Measured on the c2 workload (quad stage, ms per 256 frames): S512/L 24.6; S512/M2048/L 21.4; S256x8/M2048/L 18.7; S256/M1024(64 thr)/M3072/L 19.8.
constexpr int QM1_THREADS = 128, QM1_MAXN = 2048;
// tier limits of the four work lists, and the sort kernels' (threads, elements per thread, shared-memory points) per tier
constexpr int QT0 = 256, QT1 = 512, QT2 = 2048;
// the five kernels of each sort: <NT, E, MAXN, WHICH, T_LO, T_HI, NMIN, NMAX>
template <int W> using SortS8 = SortCfg<32, 8, QT0, W, 0, 0, 0, QT0>;
template <int W> using SortS16 = SortCfg<32, 16, QT1, W, 1, 1, 0, QT1>;
template <int W> using SortM = SortCfg<128, 16, QT2, W, 2, 2, 0, QT2>;
template <int W> using SortL1 = SortCfg<256, 16, 4096, W, 3, 3, 0, 4096>;
template <int W> using SortL2 = SortCfg<512, 16, 8192, W, 3, 3, 4097, (1 << 30)>;

}  // namespace cb
