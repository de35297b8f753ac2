// quads.cuh -- row A5 of SURVEY.md 8a: fit_quads() / fit_quad(), one CTA per gradient cluster.
//
// Upstream (apriltag_quad_thresh.c fit_quad, ptsort, compute_lfps, fit_line, quad_segment_maxima).  The float /
// double arithmetic below follows upstream operation by operation (the library is built with --fmad=false), and
// every order-dependent accumulation keeps upstream's order:
//   * points are first put back into scan order (y, x, probe) -- the order upstream's hash map appends them in;
//   * ptsort()'s merge sort is emulated exactly: same recursive split (sz/2), same 2..5 element sorting networks
//     at the leaves, merges that take from the SECOND half on ties -- done as parallel rank merges;
//   * the line-fit prefix moments are accumulated sequentially by six lanes (one lane per moment);
//   * the 4-corner search evaluates all <=210 subsets in parallel and keeps the first minimum in loop order.
// Work distribution: a persistent grid pulls (frame, cluster) items from a device-side work list, so cluster size
// imbalance is absorbed by the scheduler.  Clusters up to QF_NSM points are processed out of shared memory; larger
// ones use a global scratch area with the same code (generic pointers).
#pragma once
#include "common.cuh"

namespace cb {

constexpr int QF_THREADS = 256;
constexpr int QF_NSM = 2048;   // points handled in shared memory

struct LineFit { double Ex, Ey, nx, ny, err, mse; };

__device__ __forceinline__ uint32_t float_orderable(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// fit_line() on prefix moments lfps[j*6 + {Mx,My,Mxx,Mxy,Myy,W}]
__device__ __forceinline__ void fit_line(const double *__restrict__ lfps, int sz, int i0, int i1, bool want_params, LineFit &o)
{
    double Mx, My, Mxx, Myy, Mxy, W;
    int N;
    const double *a = lfps + (size_t)i1 * 6;
    if (i0 < i1) {
        N = i1 - i0 + 1;
        Mx = a[0]; My = a[1]; Mxx = a[2]; Mxy = a[3]; Myy = a[4]; W = a[5];
        if (i0 > 0) {
            const double *p = lfps + (size_t)(i0 - 1) * 6;
            Mx -= p[0]; My -= p[1]; Mxx -= p[2]; Mxy -= p[3]; Myy -= p[4]; W -= p[5];
        }
    } else {
        const double *l = lfps + (size_t)(sz - 1) * 6, *p = lfps + (size_t)(i0 - 1) * 6;
        Mx = l[0] - p[0]; My = l[1] - p[1]; Mxx = l[2] - p[2]; Mxy = l[3] - p[3]; Myy = l[4] - p[4]; W = l[5] - p[5];
        Mx += a[0]; My += a[1]; Mxx += a[2]; Mxy += a[3]; Myy += a[4]; W += a[5];
        N = sz - i0 + i1 + 1;
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

// node of ptsort()'s recursion tree that contains position i at depth d; returns false when the branch ended
// in a leaf (size <= 5) before reaching depth d.  leaf_here = node at depth d is itself a leaf.
__device__ __forceinline__ bool ptsort_node(int n, int i, int d, int &lo, int &hi)
{
    lo = 0; hi = n;
    for (int k = 0; k < d; k++) {
        if (hi - lo <= 5) return false;
        const int mid = lo + (hi - lo) / 2;
        if (i < mid) hi = mid; else lo = mid;
    }
    return true;
}

__device__ __forceinline__ uint32_t hi32(unsigned long long v) { return (uint32_t)(v >> 32); }

// block-wide helpers ---------------------------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T *scratch /* >= 8 entries */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T r = scratch[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) r = op(r, scratch[k]);
    return r;
}

struct QfShared {
    unsigned long long bufA[QF_NSM];
    unsigned long long bufB[QF_NSM];
    double bufD[QF_NSM];
    double red_d[8];
    int red_i[8];
    float red_f[8];
    unsigned long long red_u[8];
    int work;
    int nmax;
    int kept[16];
    int nkept;
    double thresh;
    int has_thresh;
    // pair tables for the 4-corner search
    double p_err[10][10], p_mse[10][10], p_nx[10][10], p_ny[10][10];
    uint32_t w_cnt[8];
    int ok;
};

__global__ void __launch_bounds__(QF_THREADS)
fit_quads_kernel(const uint8_t *__restrict__ in, const unsigned long long *__restrict__ pts, const uint32_t *__restrict__ scankey,
                 const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ worklist, const uint32_t *__restrict__ nwork,
                 uint32_t *__restrict__ work_counter, double *__restrict__ lfps_all, unsigned long long *__restrict__ scratch,
                 QuadRec *__restrict__ quads, uint32_t *__restrict__ nquads, uint32_t *__restrict__ nquads_total,
                 uint32_t *__restrict__ errflag, Geom g, Caps caps, DetParams prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QfShared &S = *reinterpret_cast<QfShared *>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t total = *nwork;

    for (;;) {
        __syncthreads();
        if (tid == 0) S.work = (int)atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t wi = (uint32_t)S.work;
        if (wi >= total) return;
        const uint32_t item = worklist[wi];
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        const int n = (int)rec.count;
        if (n < 24) continue;   // inert record (capacity overflow was flagged)
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;
        const unsigned long long *P = pts + pbase;
        const uint32_t *K = scankey + pbase;
        double *lfps = lfps_all + pbase * 6;
        unsigned long long *A, *B;
        double *D;
        if (n <= QF_NSM) { A = S.bufA; B = S.bufB; D = S.bufD; }
        else {
            A = scratch + pbase * 3; B = A + n; D = reinterpret_cast<double *>(B + n);
        }
        const uint8_t *img = in + (size_t)b * g.frame_stride;

        // ---- bounding box -----------------------------------------------------------------------------
        int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
        for (int i = tid; i < n; i += QF_THREADS) {
            const unsigned long long p = P[i];
            const int x = (int)(p & 0xffff), y = (int)((p >> 16) & 0xffff);
            xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
        }
        xmin = block_reduce(xmin, [](int a, int c) { return min(a, c); }, S.red_i);
        xmax = block_reduce(xmax, [](int a, int c) { return max(a, c); }, S.red_i);
        ymin = block_reduce(ymin, [](int a, int c) { return min(a, c); }, S.red_i);
        ymax = block_reduce(ymax, [](int a, int c) { return max(a, c); }, S.red_i);
        if ((xmax - xmin) * (ymax - ymin) < prm.min_tag_width) continue;
        const float cx = (float)((xmin + xmax) * 0.5 + 0.05118);
        const float cy = (float)((ymin + ymax) * 0.5 + -0.028581);

        // ---- restore scan order: merge sort on (scankey, index), keys are unique ------------------------
        for (int i = tid; i < n; i += QF_THREADS) A[i] = ((unsigned long long)K[i] << 32) | (uint32_t)i;
        __syncthreads();
        unsigned long long *src = A, *dst = B;
        for (int run = 1; run < n; run <<= 1) {
            for (int i = tid; i < n; i += QF_THREADS) {
                const int pair0 = (i / (2 * run)) * (2 * run);
                const int mid = min(pair0 + run, n), end = min(pair0 + 2 * run, n);
                const unsigned long long v = src[i];
                int lo, hi;
                if (i < mid) { lo = mid; hi = end; } else { lo = pair0; hi = mid; }
                const int sbase = lo;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (src[m] < v) lo = m + 1; else hi = m; }
                const int pos = (i < mid) ? (i + (lo - sbase)) : (pair0 + (i - mid) + (lo - sbase));
                dst[pos] = v;
            }
            __syncthreads();
            unsigned long long *t = src; src = dst; dst = t;
        }
        // ---- slopes in scan order (upstream fit_quad step 1) --------------------------------------------
        float dot = 0.f;
        for (int j = tid; j < n; j += QF_THREADS) {
            const uint32_t idx = (uint32_t)src[j];
            const unsigned long long p = P[idx];
            const int x = (int)(p & 0xffff), y = (int)((p >> 16) & 0xffff);
            const int gx = (int)(int16_t)((p >> 32) & 0xffff), gy = (int)(int16_t)((p >> 48) & 0xffff);
            float dx = (float)x - cx, dy = (float)y - cy;
            dot += dx * (float)gx + dy * (float)gy;
            float quadrant;
            if (dy > 0) quadrant = dx > 0 ? 65536.f : 131072.f; else quadrant = dx > 0 ? 0.f : -65536.f;
            if (dy < 0) { dy = -dy; dx = -dx; }
            if (dx < 0) { const float t = dx; dx = dy; dy = -t; }
            const float slope = quadrant + dy / dx;
            dst[j] = ((unsigned long long)float_orderable(slope) << 32) | idx;
        }
        dot = block_reduce(dot, [](float a, float c) { return a + c; }, S.red_f);
        const int reversed_border = dot < 0.f;
        if (reversed_border) continue;              // tag36h11 has a normal border only
        { unsigned long long *t = src; src = dst; dst = t; }
        __syncthreads();

        // ---- ptsort(): leaves (sorting networks), then merges bottom-up with "second half first on ties" --
        int maxd = 0;
        { int sz = n; while (sz > 5) { sz = sz - sz / 2; maxd++; } }
        for (int i = tid; i < n; i += QF_THREADS) {
            int lo = 0, hi = n;
            while (hi - lo > 5) { const int mid = lo + (hi - lo) / 2; if (i < mid) hi = mid; else lo = mid; }
            if (i != lo) continue;
            const int sz = hi - lo;
            unsigned long long *a = src + lo;
#define QF_SWAP(x, y) if (hi32(a[x]) > hi32(a[y])) { const unsigned long long t = a[x]; a[x] = a[y]; a[y] = t; }
            if (sz == 2) { QF_SWAP(0, 1); }
            else if (sz == 3) { QF_SWAP(0, 1); QF_SWAP(1, 2); QF_SWAP(0, 1); }
            else if (sz == 4) { QF_SWAP(0, 1); QF_SWAP(2, 3); QF_SWAP(0, 2); QF_SWAP(1, 3); QF_SWAP(1, 2); }
            else if (sz == 5) { QF_SWAP(0, 1); QF_SWAP(3, 4); QF_SWAP(2, 4); QF_SWAP(2, 3); QF_SWAP(0, 3); QF_SWAP(0, 2); QF_SWAP(1, 4); QF_SWAP(1, 3); QF_SWAP(1, 2); }
#undef QF_SWAP
        }
        __syncthreads();
        for (int d = maxd - 1; d >= 0; d--) {
            for (int i = tid; i < n; i += QF_THREADS) {
                int lo, hi;
                const unsigned long long v = src[i];
                if (!ptsort_node(n, i, d, lo, hi) || hi - lo <= 5) { dst[i] = v; continue; }
                const int mid = lo + (hi - lo) / 2;
                const uint32_t key = hi32(v);
                int pos;
                if (i < mid) {   // from the first half: all second-half keys <= key go before it
                    int l = mid, h = hi;
                    while (l < h) { const int m = (l + h) >> 1; if (hi32(src[m]) <= key) l = m + 1; else h = m; }
                    pos = i + (l - mid);
                } else {         // from the second half: only strictly smaller first-half keys go before it
                    int l = lo, h = mid;
                    while (l < h) { const int m = (l + h) >> 1; if (hi32(src[m]) < key) l = m + 1; else h = m; }
                    pos = lo + (i - mid) + (l - lo);
                }
                dst[pos] = v;
            }
            __syncthreads();
            unsigned long long *t = src; src = dst; dst = t;
        }
        // src: sorted (key, idx).  ---- compute_lfps: per-point weight, then sequential prefix moments ------
        uint32_t *XY = reinterpret_cast<uint32_t *>(dst);
        for (int j = tid; j < n; j += QF_THREADS) {
            const uint32_t idx = (uint32_t)src[j];
            const unsigned long long p = P[idx];
            const int px = (int)(p & 0xffff), py = (int)((p >> 16) & 0xffff);
            const double x = px * .5 + 0.5, y = py * .5 + 0.5;
            const int ix = (int)x, iy = (int)y;
            double W = 1;
            if (ix > 0 && ix + 1 < g.w && iy > 0 && iy + 1 < g.h) {
                const int grad_x = (int)img[(size_t)(iy * g.f) * g.stride + (ix + 1) * g.f] - (int)img[(size_t)(iy * g.f) * g.stride + (ix - 1) * g.f];
                const int grad_y = (int)img[(size_t)((iy + 1) * g.f) * g.stride + ix * g.f] - (int)img[(size_t)((iy - 1) * g.f) * g.stride + ix * g.f];
                W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
            }
            D[j] = W;
            XY[j] = (uint32_t)px | ((uint32_t)py << 16);
        }
        __syncthreads();
        if (tid < 6) {
            double acc = 0;
            for (int j = 0; j < n; j++) {
                const double W = D[j];
                const uint32_t xy = XY[j];
                const double fx = (double)(xy & 0xffff) * .5 + 0.5, fy = (double)(xy >> 16) * .5 + 0.5;
                double term;
                switch (tid) {
                    case 0: term = W * fx; break;
                    case 1: term = W * fy; break;
                    case 2: term = W * fx * fx; break;
                    case 3: term = W * fx * fy; break;
                    case 4: term = W * fy * fy; break;
                    default: term = W; break;
                }
                acc += term;
                lfps[(size_t)j * 6 + tid] = acc;
            }
        }
        __syncthreads();

        // ---- quad_segment_maxima ---------------------------------------------------------------------------
        const int ksz = min(20, n / 12);
        if (ksz < 2) continue;
        double *errs = reinterpret_cast<double *>(src);   // sorted keys are no longer needed
        double *ysm = reinterpret_cast<double *>(dst);
        for (int i = tid; i < n; i += QF_THREADS) {
            LineFit lf;
            fit_line(lfps, n, (i + n - ksz) % n, (i + ksz) % n, false, lf);
            errs[i] = lf.err;
        }
        __syncthreads();
        for (int iy = tid; iy < n; iy += QF_THREADS) {
            double acc = 0;
#pragma unroll
            for (int i = 0; i < 7; i++) acc += errs[(iy + i - 3 + n) % n] * prm.smooth_f[i];
            ysm[iy] = acc;
        }
        __syncthreads();
        // local maxima, collected in index order into (int) maxima[] / D-backed maxima_errs
        int *maxima = reinterpret_cast<int *>(errs);      // errs is dead after smoothing
        double *maxima_errs = D;
        if (tid == 0) S.nmax = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += QF_THREADS) {
            const int i = i0 + tid;
            bool is_max = false;
            double e = 0;
            if (i < n) { e = ysm[i]; is_max = e > ysm[(i + 1) % n] && e > ysm[(i + n - 1) % n]; }
            const uint32_t bal = __ballot_sync(0xffffffffu, is_max);
            const int lane = tid & 31, wid = tid >> 5;
            if (lane == 0) S.w_cnt[wid] = __popc(bal);
            __syncthreads();
            int base = S.nmax;
            for (int k = 0; k < wid; k++) base += S.w_cnt[k];
            if (is_max) {
                const int pos = base + __popc(bal & ((1u << lane) - 1));
                maxima[pos] = i;
                maxima_errs[pos] = e;
            }
            __syncthreads();
            if (tid == 0) { int t = 0; for (int k = 0; k < QF_THREADS / 32; k++) t += S.w_cnt[k]; S.nmax += t; }
            __syncthreads();
        }
        const int nmaxima = S.nmax;
        if (nmaxima < 4) continue;
        // keep only the best max_nmaxima
        if (tid == 0) { S.nkept = 0; S.has_thresh = 0; }
        __syncthreads();
        const int max_nmaxima = min(prm.max_nmaxima, 10);
        if (nmaxima > max_nmaxima) {
            // maxima_thresh = element [max_nmaxima] of the descending sort = value v with #(>v) <= max_nmaxima < #(>=v)
            for (int m = tid; m < nmaxima; m += QF_THREADS) {
                const double e = maxima_errs[m];
                int gt = 0, ge = 0;
                for (int k = 0; k < nmaxima; k++) { const double o = maxima_errs[k]; gt += o > e; ge += o >= e; }
                if (gt <= max_nmaxima && ge > max_nmaxima) { S.thresh = e; S.has_thresh = 1; }
            }
            __syncthreads();
            if (tid == 0) {
                int out = 0;
                const double th = S.thresh;
                for (int m = 0; m < nmaxima; m++) {
                    if (maxima_errs[m] <= th) continue;
                    if (out < 16) S.kept[out] = maxima[m];
                    out++;
                }
                S.nkept = min(out, 16);
            }
        } else if (tid == 0) {
            for (int m = 0; m < nmaxima; m++) S.kept[m] = maxima[m];
            S.nkept = nmaxima;
        }
        __syncthreads();
        const int nk = S.nkept;
        if (nk < 4) continue;   // (upstream's loops would simply find nothing)
        // pair table: fit_line(kept[a], kept[b]) for a != b
        for (int t = tid; t < nk * nk; t += QF_THREADS) {
            const int a = t / nk, c = t % nk;
            if (a == c) continue;
            LineFit lf;
            fit_line(lfps, n, S.kept[a], S.kept[c], true, lf);
            S.p_err[a][c] = lf.err; S.p_mse[a][c] = lf.mse; S.p_nx[a][c] = lf.nx; S.p_ny[a][c] = lf.ny;
        }
        __syncthreads();
        // 4-corner search: combination index in upstream's loop order; keep the first minimum
        double best_err = __longlong_as_double(0x7ff0000000000000ll);
        int best_combo = 1 << 30;
        {
            const double max_mse = (double)prm.max_line_fit_mse;
            int ci = 0;
            for (int m0 = 0; m0 < nk - 3; m0++)
                for (int m1 = m0 + 1; m1 < nk - 2; m1++)
                    for (int m2 = m1 + 1; m2 < nk - 1; m2++)
                        for (int m3 = m2 + 1; m3 < nk; m3++, ci++) {
                            if ((ci % QF_THREADS) != tid) continue;
                            if (S.p_mse[m0][m1] > max_mse) continue;
                            if (S.p_mse[m1][m2] > max_mse) continue;
                            const double dt = S.p_nx[m0][m1] * S.p_nx[m1][m2] + S.p_ny[m0][m1] * S.p_ny[m1][m2];
                            if (fabs(dt) > prm.cos_critical_rad) continue;
                            if (S.p_mse[m2][m3] > max_mse) continue;
                            if (S.p_mse[m3][m0] > max_mse) continue;
                            const double err = S.p_err[m0][m1] + S.p_err[m1][m2] + S.p_err[m2][m3] + S.p_err[m3][m0];
                            if (err < best_err) { best_err = err; best_combo = (m0 << 12) | (m1 << 8) | (m2 << 4) | m3; }
                        }
        }
        // (m0,m1,m2,m3) packed big-endian orders exactly like the loop nest, so min over (err, packed) = first minimum
        {
            // reduce on err first, then on combo among equal err
            const double bmin = block_reduce(best_err, [](double a, double c) { return a < c ? a : c; }, S.red_d);
            int cand = (best_err == bmin && best_combo != (1 << 30)) ? best_combo : (1 << 30);
            cand = block_reduce(cand, [](int a, int c) { return min(a, c); }, S.red_i);
            best_err = bmin; best_combo = cand;
        }
        if (best_combo == (1 << 30)) continue;
        if (!(best_err / n < (double)prm.max_line_fit_mse)) continue;

        // ---- corners, area and convexity tests (thread 0) --------------------------------------------------
        if (tid == 0) {
            S.ok = 0;
            const int mi[4] = {(best_combo >> 12) & 15, (best_combo >> 8) & 15, (best_combo >> 4) & 15, best_combo & 15};
            int indices[4];
            for (int i = 0; i < 4; i++) indices[i] = S.kept[mi[i]];
            double lines[4][4];
            bool good = true;
            for (int i = 0; i < 4 && good; i++) {
                LineFit lf;
                fit_line(lfps, n, indices[i], indices[(i + 1) & 3], true, lf);
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
}

}  // namespace cb
