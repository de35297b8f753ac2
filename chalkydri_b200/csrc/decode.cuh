// decode.cuh -- rows A6-A9 of SURVEY.md 8a: refine_edges, homography, tag36h11 decode, reconcile.
//
// Upstream (apriltag.c refine_edges / quad_update_homographies / quad_decode / quick_decode_codeword and the
// reconcile loop of apriltag_detector_detect; homography.c homography_compute2).  One warp per candidate quad:
// lanes sample edge / border / bit positions in parallel (sparse, L2-resident gathers from the full-resolution
// frame), while every accumulation whose result depends on summation order is replayed in upstream's order
// through warp shuffles, so the double arithmetic matches the CPU restatement operation by operation.
#pragma once
#include "common.cuh"

namespace cb {

// read with a different index per lane, so it lives in global memory (L1) instead of the constant bank
__device__ unsigned long long c_codes[kNumCodes];
__constant__ int c_bit_x[36];
__constant__ int c_bit_y[36];

struct DecodeConst {
    double rot_c[4], rot_s[4];    // host libm cos/sin of k*pi/2, like upstream's cos(theta)/sin(theta)
};

struct RawDet {
    cb_detection d;
    int32_t valid;
    int32_t pad;
};

constexpr int DEC_WARPS = 4;

__device__ __forceinline__ void homography_project(const double *H, double x, double y, double *ox, double *oy)
{
    const double xx = H[0] * x + H[1] * y + H[2];
    const double yy = H[3] * x + H[4] * y + H[5];
    const double zz = H[6] * x + H[7] * y + H[8];
    *ox = xx / zz;
    *oy = yy / zz;
}

// homography_compute2(): 8x9 Gaussian elimination with partial pivoting, executed by a full warp.  Lane j (0..8) keeps
// column j of the system in registers (no 72-double local array); the pivot search of column `col` runs on lane col, the
// multipliers are broadcast by shuffle, and every element sees exactly the operations of the sequential code
// (f = A[i][col] / A[col][col]; A[i][j] -= f * A[col][j]; back substitution summed in ascending i).
__device__ __forceinline__ bool homography_compute2(const double c[4][4], double *H)
{
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double a[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double cx = c[i][0], cy = c[i][1], cz = c[i][2], cw = c[i][3];
        double r0, r1;
        switch (lane) {
            case 0: r0 = cx; r1 = 0; break;
            case 1: r0 = cy; r1 = 0; break;
            case 2: r0 = 1; r1 = 0; break;
            case 3: r0 = 0; r1 = cx; break;
            case 4: r0 = 0; r1 = cy; break;
            case 5: r0 = 0; r1 = 1; break;
            case 6: r0 = -cx * cz; r1 = -cx * cw; break;
            case 7: r0 = -cy * cz; r1 = -cy * cw; break;
            case 8: r0 = cz; r1 = cw; break;
            default: r0 = 0; r1 = 0; break;
        }
        a[2 * i] = r0; a[2 * i + 1] = r1;
    }
    const double epsilon = 1e-10;
#pragma unroll
    for (int col = 0; col < 8; col++) {
        double max_val = 0;
        int max_val_idx = -1;
#pragma unroll
        for (int row = col; row < 8; row++) {
            const double val = fabs(a[row]);
            if (val > max_val) { max_val = val; max_val_idx = row; }
        }
        max_val = __shfl_sync(full, max_val, col);
        max_val_idx = __shfl_sync(full, max_val_idx, col);
        if (max_val_idx < 0) return false;
        if (max_val < epsilon) return false;
        if (max_val_idx != col) {
#pragma unroll
            for (int row = col + 1; row < 8; row++)
                if (row == max_val_idx) { const double t = a[col]; a[col] = a[row]; a[row] = t; }
        }
        const double pivot = __shfl_sync(full, a[col], col);
#pragma unroll
        for (int i = col + 1; i < 8; i++) {
            const double f = __shfl_sync(full, a[i], col) / pivot;
            if (lane == col) a[i] = 0;
            else if (lane > col) a[i] -= f * a[col];
        }
    }
    double x[8];
#pragma unroll
    for (int col = 7; col >= 0; col--) {
        double sum = 0;
#pragma unroll
        for (int i = col + 1; i < 8; i++) sum += __shfl_sync(full, a[col], i) * x[i];
        x[col] = (__shfl_sync(full, a[col], 8) - sum) / __shfl_sync(full, a[col], col);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) H[i] = x[i];
    H[8] = 1;
    return true;
}

// matd_plu() singular flag of the 3x3 H (matd_inverse inside quad_update_homographies)
__device__ bool mat33_plu_nonsingular(const double *H)
{
    double lu[9];
    for (int i = 0; i < 9; i++) lu[i] = H[i];
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < 3; i++) {
            const int kmax = i < j ? i : j;
            double acc = 0;
            for (int k = 0; k < kmax; k++) acc += lu[i * 3 + k] * lu[k * 3 + j];
            lu[i * 3 + j] -= acc;
        }
        int p = j;
        for (int i = j + 1; i < 3; i++)
            if (fabs(lu[i * 3 + j]) > fabs(lu[p * 3 + j])) p = i;
        if (p != j)
            for (int k = 0; k < 3; k++) { const double t = lu[p * 3 + k]; lu[p * 3 + k] = lu[j * 3 + k]; lu[j * 3 + k] = t; }
        const double LUjj = lu[j * 3 + j];
        if (fabs(LUjj) < 1e-8) return false;
        for (int i = j + 1; i < 3; i++) lu[i * 3 + j] /= LUjj;
    }
    return true;
}

struct GrayModel { double A00, A01, A02, A11, A12, A22, B0, B1, B2, C0, C1, C2; };

__device__ __forceinline__ void gm_solve(GrayModel &gm)
{
    // mat33_sym_solve: Cholesky, lower-triangular inverse, two triangular products
    double L0 = sqrt(gm.A00), L3 = gm.A01 / L0, L6 = gm.A02 / L0;
    double L4 = sqrt(gm.A11 - L3 * L3), L7 = (gm.A12 - L3 * L6) / L4;
    double L8 = sqrt(gm.A22 - L6 * L6 - L7 * L7);
    double M0 = 1 / L0, M3 = -L3 * M0 / L4, M4 = 1 / L4;
    double M6 = (-L6 * M0 - L7 * M3) / L8, M7 = -L7 * M4 / L8, M8 = 1 / L8;
    double t0 = M0 * gm.B0;
    double t1 = M3 * gm.B0 + M4 * gm.B1;
    double t2 = M6 * gm.B0 + M7 * gm.B1 + M8 * gm.B2;
    gm.C0 = M0 * t0 + M3 * t1 + M6 * t2;
    gm.C1 = M4 * t1 + M7 * t2;
    gm.C2 = M8 * t2;
}
__device__ __forceinline__ double gm_interp(const GrayModel &gm, double x, double y) { return gm.C0 * x + gm.C1 * y + gm.C2; }

__device__ __forceinline__ double value_for_pixel(const uint8_t *img, int W, int H, int stride, double px, double py)
{
    const int x1 = (int)floor(px - 0.5), x2 = (int)ceil(px - 0.5);
    const double x = px - 0.5 - x1;
    const int y1 = (int)floor(py - 0.5), y2 = (int)ceil(py - 0.5);
    const double y = py - 0.5 - y1;
    if (x1 < 0 || x2 >= W || y1 < 0 || y2 >= H) return -1;
    return img[(size_t)y1 * stride + x1] * (1 - x) * (1 - y) + img[(size_t)y1 * stride + x2] * x * (1 - y) +
           img[(size_t)y2 * stride + x1] * (1 - x) * y + img[(size_t)y2 * stride + x2] * x * y;
}

__device__ __forceinline__ unsigned long long rotate90_36(unsigned long long w)
{
    return ((w << 9) | (w >> 27)) & ((1ull << 36) - 1);
}

// one candidate quad, executed by a full warp
__device__ void decode_one_quad(const uint8_t *__restrict__ in, const QuadRec &q, RawDet *__restrict__ raw, uint32_t *__restrict__ nraw,
                                const Geom &g, const Caps &caps, const DetParams &prm, const DecodeConst &dc, double *values)
{
    const int lane = threadIdx.x & 31;
    const int b = q.frame;
    const uint32_t full = 0xffffffffu;
    const uint8_t *img = in + (size_t)b * g.frame_stride;
    const int W = g.W, H = g.H, stride = g.stride;

    float p[4][2];
    for (int j = 0; j < 4; j++) {
        if (prm.quad_decimate > 1) {
            p[j][0] = (float)(((double)q.p[j][0] - 0.5) * (double)prm.quad_decimate + 0.5);
            p[j][1] = (float)(((double)q.p[j][1] - 0.5) * (double)prm.quad_decimate + 0.5);
        } else { p[j][0] = q.p[j][0]; p[j][1] = q.p[j][1]; }
    }

    // ---- refine_edges ----------------------------------------------------------------------------------
    if (prm.refine_edges) {
        double lines[4][4];
        // Two edges at a time, one per half-warp (a tag edge below ~136 px has 16 samples, which would leave half of a warp
        // idle): lanes 0..15 take edge 2*pass, lanes 16..31 edge 2*pass + 1, sixteen samples per step each.
        const int half = lane >> 4, hl = lane & 15;
        for (int pass = 0; pass < 2; pass++) {
            const int edge = 2 * pass + half;
            const int a = edge, bb = (edge + 1) & 3;
            // (p[][] is indexed with a lane-dependent edge: select instead of dynamic indexing)
            const double pax = (double)(a == 0 ? p[0][0] : a == 1 ? p[1][0] : a == 2 ? p[2][0] : p[3][0]);
            const double pay = (double)(a == 0 ? p[0][1] : a == 1 ? p[1][1] : a == 2 ? p[2][1] : p[3][1]);
            const double pbx = (double)(bb == 0 ? p[0][0] : bb == 1 ? p[1][0] : bb == 2 ? p[2][0] : p[3][0]);
            const double pby = (double)(bb == 0 ? p[0][1] : bb == 1 ? p[1][1] : bb == 2 ? p[2][1] : p[3][1]);
            double nx = pby - pay;
            double ny = -pbx + pax;
            const double mag = sqrt(nx * nx + ny * ny);
            nx /= mag; ny /= mag;
            if (q.reversed_border) { nx = -nx; ny = -ny; }
            const int nsamples = max(16, (int)(mag / 8));
            const int nsamples_max = max(nsamples, __shfl_xor_sync(full, nsamples, 16));
            double Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0, N = 0;
            const double range = (double)prm.quad_decimate + 1;
            for (int s0 = 0; s0 < nsamples_max; s0 += 16) {
                const int s = s0 + hl;
                double bestx = 0, besty = 0;
                int have = 0;
                if (s < nsamples) {
                    const double alpha = (1.0 + s) / (nsamples + 1);
                    const double x0 = alpha * pax + (1 - alpha) * pbx;
                    const double y0 = alpha * pay + (1 - alpha) * pby;
                    double Mn = 0, Mcount = 0;
                    // upstream: for (double n = -range; n <= range; n += 0.25).  k*0.25 - range is exact in binary, so the steps are
                    // enumerated by index; five steps (ten gathers) are kept in flight at a time.  weight and weight*n are exact
                    // in double (integers times multiples of 1/4), so Mn / Mcount do not depend on the evaluation order.
                    const int nsteps = (int)(range * 8.0) + 1;
                    for (int i0 = 0; i0 < nsteps; i0 += 5) {
                        int g1v[5], g2v[5];
#pragma unroll
                        for (int u = 0; u < 5; u++) {
                            const double n = -range + 0.25 * (i0 + u);
                            const int x1 = (int)(x0 + (n + 1) * nx), y1 = (int)(y0 + (n + 1) * ny);
                            const int x2 = (int)(x0 + (n - 1) * nx), y2 = (int)(y0 + (n - 1) * ny);
                            const bool ok = i0 + u < nsteps && !(x1 < 0 || x1 >= W || y1 < 0 || y1 >= H) && !(x2 < 0 || x2 >= W || y2 < 0 || y2 >= H);
                            g1v[u] = -1; g2v[u] = 0;
                            if (ok) { g1v[u] = img[(size_t)y1 * stride + x1]; g2v[u] = img[(size_t)y2 * stride + x2]; }
                        }
#pragma unroll
                        for (int u = 0; u < 5; u++) {
                            if (g1v[u] < g2v[u]) continue;          // also skips the out-of-range steps (g1 = -1)
                            const double n = -range + 0.25 * (i0 + u);
                            const double weight = (double)((g2v[u] - g1v[u]) * (g2v[u] - g1v[u]));
                            Mn += weight * n;
                            Mcount += weight;
                        }
                    }
                    if (Mcount != 0) {
                        const double n0 = Mn / Mcount;
                        bestx = x0 + n0 * nx; besty = y0 + n0 * ny;
                        have = 1;
                    }
                }
                // replay the accumulation in sample order inside each half-warp (samples beyond an edge's count have have = 0)
                for (int k = 0; k < 16; k++) {
                    const int src = (lane & 16) | k;
                    const int hv = __shfl_sync(full, have, src);
                    const double bx = __shfl_sync(full, bestx, src), by = __shfl_sync(full, besty, src);
                    if (hv) { Mx += bx; My += by; Mxx += bx * bx; Mxy += bx * by; Myy += by * by; N++; }
                }
            }
            const double Ex = Mx / N, Ey = My / N;
            const double Cxx = Mxx / N - Ex * Ex, Cxy = Mxy / N - Ex * Ey, Cyy = Myy / N - Ey * Ey;
            const double normal_theta = .5 * atan2f((float)(-2 * Cxy), (float)(Cyy - Cxx));
            nx = cosf((float)normal_theta);
            ny = sinf((float)normal_theta);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                lines[2 * pass + h][0] = __shfl_sync(full, Ex, 16 * h); lines[2 * pass + h][1] = __shfl_sync(full, Ey, 16 * h);
                lines[2 * pass + h][2] = __shfl_sync(full, nx, 16 * h); lines[2 * pass + h][3] = __shfl_sync(full, ny, 16 * h);
            }
        }
        for (int i = 0; i < 4; i++) {
            const double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
            const double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
            const double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
            const double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
            const double det = A00 * A11 - A10 * A01;
            if (fabs(det) > 0.001) {
                const double W00 = A11 / det, W01 = -A01 / det;
                const double L0 = W00 * B0 + W01 * B1;
                p[i][0] = (float)(lines[i][0] + L0 * A00);
                p[i][1] = (float)(lines[i][1] + L0 * A10);
            }
        }
    }

    // ---- quad_update_homographies ------------------------------------------------------------------------
    double Hq[9];
    {
        double corr[4][4];
        for (int i = 0; i < 4; i++) {
            corr[i][0] = (i == 0 || i == 3) ? -1 : 1;
            corr[i][1] = (i == 0 || i == 1) ? -1 : 1;
            corr[i][2] = p[i][0];
            corr[i][3] = p[i][1];
        }
        if (!homography_compute2(corr, Hq)) return;
        if (!mat33_plu_nonsingular(Hq)) return;
    }

    // ---- quad_decode: gray models from the border samples ------------------------------------------------
    GrayModel wm = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, bm = wm;
    {
        const float wab = 8.f;
        // the 64 border samples in upstream's loop order: tag coordinates, pixel value, flag (0 skip, 1 white, 2 black)
        double *s_tagx = values, *s_tagy = values + 64;
        int *s_v = reinterpret_cast<int *>(values + 128), *s_flag = s_v + 64;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = lane + 32 * h, pi = k >> 3, i = k & 7;
            float p0, p1, p2, p3; int is_white;
            switch (pi) {
                case 0: p0 = -0.5f; p1 = 0.5f; p2 = 0; p3 = 1; is_white = 1; break;
                case 1: p0 = 0.5f; p1 = 0.5f; p2 = 0; p3 = 1; is_white = 0; break;
                case 2: p0 = wab + 0.5f; p1 = .5f; p2 = 0; p3 = 1; is_white = 1; break;
                case 3: p0 = wab - 0.5f; p1 = .5f; p2 = 0; p3 = 1; is_white = 0; break;
                case 4: p0 = 0.5f; p1 = -0.5f; p2 = 1; p3 = 0; is_white = 1; break;
                case 5: p0 = 0.5f; p1 = 0.5f; p2 = 1; p3 = 0; is_white = 0; break;
                case 6: p0 = 0.5f; p1 = wab + 0.5f; p2 = 1; p3 = 0; is_white = 1; break;
                default: p0 = 0.5f; p1 = wab - 0.5f; p2 = 1; p3 = 0; is_white = 0; break;
            }
            const double tagx01 = (double)(p0 + (float)i * p2) / 8.0;
            const double tagy01 = (double)(p1 + (float)i * p3) / 8.0;
            const double tagx = 2 * (tagx01 - 0.5), tagy = 2 * (tagy01 - 0.5);
            double px, py;
            homography_project(Hq, tagx, tagy, &px, &py);
            const int ix = (int)px, iy = (int)py;
            int sv = 0, sf = 0;
            if (!(ix < 0 || iy < 0 || ix >= W || iy >= H)) {
                sv = img[(size_t)iy * stride + ix];
                sf = is_white ? 1 : 2;
            }
            s_tagx[k] = tagx; s_tagy[k] = tagy; s_v[k] = sv; s_flag[k] = sf;
        }
        __syncwarp();
        // The nine sums of each model are independent chains whose terms are all products of two of {x, y, gray, 1}
        // (x * 1 and 1 * 1 are exact), so lane a (white) / 16 + a (black) accumulates sum a over the samples in order.
        const int a = lane & 15, my_flag = lane < 16 ? 1 : 2;
        const int su = (a == 0 || a == 1 || a == 2 || a == 6) ? 0 : ((a == 3 || a == 4 || a == 7) ? 1 : (a == 8 ? 2 : 3));
        const int sw = a == 0 ? 0 : ((a == 1 || a == 3) ? 1 : ((a == 6 || a == 7) ? 2 : 3));
        double acc = 0;
#pragma unroll 4
        for (int k = 0; k < 64; k++) {
            const double x = s_tagx[k], y = s_tagy[k], gray = (double)s_v[k];
            const double u = su == 0 ? x : (su == 1 ? y : (su == 2 ? gray : 1.0));
            const double w = sw == 0 ? x : (sw == 1 ? y : (sw == 2 ? gray : 1.0));
            if (s_flag[k] == my_flag) acc += u * w;
        }
        wm.A00 = __shfl_sync(full, acc, 0); wm.A01 = __shfl_sync(full, acc, 1); wm.A02 = __shfl_sync(full, acc, 2);
        wm.A11 = __shfl_sync(full, acc, 3); wm.A12 = __shfl_sync(full, acc, 4); wm.A22 = __shfl_sync(full, acc, 5);
        wm.B0 = __shfl_sync(full, acc, 6); wm.B1 = __shfl_sync(full, acc, 7); wm.B2 = __shfl_sync(full, acc, 8);
        bm.A00 = __shfl_sync(full, acc, 16); bm.A01 = __shfl_sync(full, acc, 17); bm.A02 = __shfl_sync(full, acc, 18);
        bm.A11 = __shfl_sync(full, acc, 19); bm.A12 = __shfl_sync(full, acc, 20); bm.A22 = __shfl_sync(full, acc, 21);
        bm.B0 = __shfl_sync(full, acc, 22); bm.B1 = __shfl_sync(full, acc, 23); bm.B2 = __shfl_sync(full, acc, 24);
        __syncwarp();
    }
    gm_solve(wm);
    gm_solve(bm);
    if ((gm_interp(wm, 0, 0) - gm_interp(bm, 0, 0) < 0) != false) return;

    // ---- bit samples, sharpening, code word ---------------------------------------------------------------
    for (int k = lane; k < 100; k += 32) values[k] = 0;
    __syncwarp();
    for (int i = lane; i < 36; i += 32) {
        const int bity = c_bit_y[i], bitx = c_bit_x[i];
        const double tagx01 = (bitx + 0.5) / 8.0, tagy01 = (bity + 0.5) / 8.0;
        const double tagx = 2 * (tagx01 - 0.5), tagy = 2 * (tagy01 - 0.5);
        double px, py;
        homography_project(Hq, tagx, tagy, &px, &py);
        const double v = value_for_pixel(img, W, H, stride, px, py);
        if (v == -1) continue;
        const double thresh = (gm_interp(bm, tagx, tagy) + gm_interp(wm, tagx, tagy)) / 2.0;
        values[10 * (bity + 1) + bitx + 1] = v - thresh;
    }
    __syncwarp();
    {
        double sh[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int k = lane + 32 * r;
            sh[r] = 0;
            if (k < 100) {
                const int y = k / 10, x = k % 10;
                double acc = 0;
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) {
                        if ((y + i - 1) < 0 || (y + i - 1) > 9 || (x + j - 1) < 0 || (x + j - 1) > 9) continue;
                        const double kern = (i == 1 && j == 1) ? 4.0 : ((i == 1 || j == 1) ? -1.0 : 0.0);
                        acc += values[(y + i - 1) * 10 + (x + j - 1)] * kern;
                    }
                sh[r] = acc;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int k = lane + 32 * r;
            if (k < 100) values[k] = values[k] + prm.decode_sharpening * sh[r];
        }
        __syncwarp();
    }
    unsigned long long rcode = 0;
    float black_score = 0, white_score = 0, black_score_count = 1, white_score_count = 1;
    for (int i = 0; i < 36; i++) {
        rcode <<= 1;
        const double v = values[(c_bit_y[i] + 1) * 10 + c_bit_x[i] + 1];
        if (v > 0) { white_score += (float)v; white_score_count++; rcode |= 1; }
        else { black_score -= (float)v; black_score_count++; }
    }
    // ---- quick_decode_codeword: first rotation with a code within bits_corrected ----------------------------
    int id = 65535, hamming = 255, rotation = 0;
    {
        unsigned long long rc = rcode;
        for (int ridx = 0; ridx < 4; ridx++) {
            int best = 1 << 30;   // (id << 8) | hamming, smallest id first
            for (int c = lane; c < kNumCodes; c += 32) {
                const int d = __popcll(rc ^ __ldg(&c_codes[c]));
                if (d <= prm.bits_corrected) { best = min(best, (c << 8) | d); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(full, best, o));
            if (best != (1 << 30)) { id = best >> 8; hamming = best & 255; rotation = ridx; break; }
            rc = rotate90_36(rc);
        }
    }
    const float decision_margin = fminf(white_score / white_score_count, black_score / black_score_count);
    if (!(decision_margin >= 0 && hamming < 255)) return;
    if (lane != 0) return;

    RawDet out;
    out.valid = 1; out.pad = 0;
    out.d.frame = b; out.d.id = id; out.d.hamming = hamming; out.d.decision_margin = decision_margin;
    {
        const double c = dc.rot_c[rotation], s = dc.rot_s[rotation];
        const double R[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double acc = 0;
                for (int k = 0; k < 3; k++) acc += Hq[i * 3 + k] * R[k * 3 + j];
                out.d.H[i * 3 + j] = acc;
            }
        homography_project(out.d.H, 0, 0, &out.d.c[0], &out.d.c[1]);
        for (int i = 0; i < 4; i++) {
            const int tcx = (i == 1 || i == 2) ? 1 : -1;
            const int tcy = (i < 2) ? 1 : -1;
            homography_project(out.d.H, tcx, tcy, &out.d.p[i][0], &out.d.p[i][1]);
        }
    }
    const uint32_t slot = atomicAdd(&nraw[b], 1u);
    if (slot < caps.quads_per_frame) raw[(size_t)b * caps.quads_per_frame + slot] = out;
}

// persistent warps pull candidate quads from the batch-wide list
__global__ void __launch_bounds__(DEC_WARPS * 32, 6)
decode_quads_kernel(const uint8_t *__restrict__ in, const QuadRec *__restrict__ quads, const uint32_t *__restrict__ nquads_total,
                    uint32_t *__restrict__ counter, RawDet *__restrict__ raw, uint32_t *__restrict__ nraw, Geom g, Caps caps,
                    DetParams prm, DecodeConst dc)
{
    __shared__ double s_values[DEC_WARPS][192];   // 10x10 bit grid; before that the 64 border samples (2 x 64 doubles + 2 x 64 ints)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t total = *nquads_total;
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(counter, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= total) return;
        const QuadRec q = quads[qi];
        decode_one_quad(in, q, raw, nraw, g, caps, prm, dc, s_values[wid]);
        __syncwarp();
    }
}

// ---- reconcile + sort: one thread per frame (a handful of detections) ------------------------------------
__device__ __forceinline__ double cross2(const double *a, const double *b, const double *c)
{
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}
__device__ bool on_seg(const double *a, const double *b, const double *c)
{
    return fmin(a[0], b[0]) <= c[0] && c[0] <= fmax(a[0], b[0]) && fmin(a[1], b[1]) <= c[1] && c[1] <= fmax(a[1], b[1]);
}
__device__ bool seg_intersect(const double *p0, const double *p1, const double *q0, const double *q1)
{
    const double d1 = cross2(q0, q1, p0), d2 = cross2(q0, q1, p1), d3 = cross2(p0, p1, q0), d4 = cross2(p0, p1, q1);
    if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0))) return true;
    if (d1 == 0 && on_seg(q0, q1, p0)) return true;
    if (d2 == 0 && on_seg(q0, q1, p1)) return true;
    if (d3 == 0 && on_seg(p0, p1, q0)) return true;
    if (d4 == 0 && on_seg(p0, p1, q1)) return true;
    return false;
}
__device__ bool poly_contains(const double (*poly)[2], const double *q)
{
    bool in = false;
    for (int i = 0, j = 3; i < 4; j = i++) {
        if (((poly[i][1] > q[1]) != (poly[j][1] > q[1])) &&
            (q[0] < (poly[j][0] - poly[i][0]) * (q[1] - poly[i][1]) / (poly[j][1] - poly[i][1]) + poly[i][0]))
            in = !in;
    }
    return in;
}
__device__ bool polygons_overlap(const double (*a)[2], const double (*b)[2])
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersect(a[i], a[(i + 1) & 3], b[j], b[(j + 1) & 3])) return true;
    double ca[2] = {0, 0}, cb2[2] = {0, 0};
    for (int i = 0; i < 4; i++) { ca[0] += a[i][0] / 4; ca[1] += a[i][1] / 4; cb2[0] += b[i][0] / 4; cb2[1] += b[i][1] / 4; }
    if (poly_contains(a, cb2)) return true;
    if (poly_contains(b, ca)) return true;
    return false;
}
__device__ __forceinline__ int prefer_smaller(int pref, double q0, double q1)
{
    if (pref) return pref;
    if (q0 < q1) return -1;
    if (q1 < q0) return 1;
    return 0;
}
__device__ __forceinline__ bool det_less(const cb_detection &a, const cb_detection &b)
{
    if (a.id != b.id) return a.id < b.id;
    if (a.hamming != b.hamming) return a.hamming < b.hamming;
    if (a.c[0] != b.c[0]) return a.c[0] < b.c[0];
    return a.c[1] < b.c[1];
}

// One warp per frame.  The bookkeeping (a handful of detections) runs on lane 0 over an index permutation in shared memory
// instead of moving 176-byte records around global memory; the surviving records are copied out by the whole warp.
// Frames with more raw detections than REC_MAX fall back to lane 0 working on the records in place.
constexpr int REC_WARPS = 4, REC_MAX = 128;
__global__ void __launch_bounds__(REC_WARPS * 32)
reconcile_kernel(RawDet *__restrict__ raw, const uint32_t *__restrict__ nraw, cb_detection *__restrict__ out,
                 int32_t *__restrict__ counts, Caps caps, int batch, const uint32_t *__restrict__ frame_err)
{
    __shared__ uint16_t s_idx[REC_WARPS][REC_MAX];
    __shared__ int s_n[REC_WARPS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * REC_WARPS + wid;
    if (b >= batch) return;
    RawDet *r = raw + (size_t)b * caps.quads_per_frame;
    int n = (int)min(nraw[b], caps.quads_per_frame);
    if (frame_err[b]) n = 0;          // a table of this frame overflowed: its list would be incomplete, it reports nothing (and its flag)
    const bool indexed = n <= REC_MAX;
    uint16_t *ix = s_idx[wid];
    if (lane == 0) {
        // The decode kernel appends in arbitrary order; upstream's outcome depends on list order only through the
        // swap-remove bookkeeping, not through which detection of an overlapping pair survives.  Put the list into a
        // canonical order first so the result is deterministic.
        if (indexed) {
            for (int i = 0; i < n; i++) ix[i] = (uint16_t)i;
            for (int i = 1; i < n; i++) {
                const uint16_t t = ix[i];
                int j = i;
                while (j > 0 && det_less(r[t].d, r[ix[j - 1]].d)) { ix[j] = ix[j - 1]; j--; }
                ix[j] = t;
            }
            for (int i0 = 0; i0 < n; i0++) {
                for (int i1 = i0 + 1; i1 < n; i1++) {
                    const cb_detection &d0 = r[ix[i0]].d, &d1 = r[ix[i1]].d;
                    if (d0.id != d1.id) continue;
                    if (!polygons_overlap(d0.p, d1.p)) continue;
                    int pref = 0;
                    pref = prefer_smaller(pref, d0.hamming, d1.hamming);
                    pref = prefer_smaller(pref, -d0.decision_margin, -d1.decision_margin);
                    for (int i = 0; i < 4; i++) {
                        pref = prefer_smaller(pref, d0.p[i][0], d1.p[i][0]);
                        pref = prefer_smaller(pref, d0.p[i][1], d1.p[i][1]);
                    }
                    if (pref < 0) { ix[i1] = ix[n - 1]; n--; i1--; }
                    else { ix[i0] = ix[n - 1]; n--; i0--; break; }
                }
            }
            for (int i = 1; i < n; i++) {
                const uint16_t t = ix[i];
                int j = i;
                while (j > 0 && det_less(r[t].d, r[ix[j - 1]].d)) { ix[j] = ix[j - 1]; j--; }
                ix[j] = t;
            }
        } else {
            for (int i = 1; i < n; i++) {
                RawDet t = r[i];
                int j = i;
                while (j > 0 && det_less(t.d, r[j - 1].d)) { r[j] = r[j - 1]; j--; }
                r[j] = t;
            }
            for (int i0 = 0; i0 < n; i0++) {
                for (int i1 = i0 + 1; i1 < n; i1++) {
                    const cb_detection &d0 = r[i0].d, &d1 = r[i1].d;
                    if (d0.id != d1.id) continue;
                    if (!polygons_overlap(d0.p, d1.p)) continue;
                    int pref = 0;
                    pref = prefer_smaller(pref, d0.hamming, d1.hamming);
                    pref = prefer_smaller(pref, -d0.decision_margin, -d1.decision_margin);
                    for (int i = 0; i < 4; i++) {
                        pref = prefer_smaller(pref, d0.p[i][0], d1.p[i][0]);
                        pref = prefer_smaller(pref, d0.p[i][1], d1.p[i][1]);
                    }
                    if (pref < 0) { r[i1] = r[n - 1]; n--; i1--; }
                    else { r[i0] = r[n - 1]; n--; i0--; break; }
                }
            }
            for (int i = 1; i < n; i++) {
                RawDet t = r[i];
                int j = i;
                while (j > 0 && det_less(t.d, r[j - 1].d)) { r[j] = r[j - 1]; j--; }
                r[j] = t;
            }
        }
        s_n[wid] = n;
    }
    __syncwarp();
    n = s_n[wid];
    const int m = min(n, (int)caps.dets_per_frame);
    static_assert(sizeof(cb_detection) % 4 == 0, "record copy works on 32-bit words");
    constexpr int WORDS = (int)(sizeof(cb_detection) / 4);
    uint32_t *o = reinterpret_cast<uint32_t *>(out + (size_t)b * caps.dets_per_frame);
    for (int i = 0; i < m; i++) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&r[indexed ? ix[i] : i].d);
        for (int w = lane; w < WORDS; w += 32) o[(size_t)i * WORDS + w] = src[w];
    }
    if (lane == 0) counts[b] = m;
}

}  // namespace cb
