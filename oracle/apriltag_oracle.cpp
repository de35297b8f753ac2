/*
 * oracle/apriltag_oracle.cpp -- CPU restatement of the UMich AprilTag-3 detector.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  The product never calls this file.
 *
 * The reference calls this detector through `apriltag::Detector::detect`
 * (/root/reference/crates/apriltags/src/lib.rs:258-261 builder, :301 detect, :306 id, :310-314 corners).
 * The C sources are an un-vendored, un-pinned git dependency (crates/apriltags/Cargo.toml:10-11),
 * so this file restates the published AprilTag 3.4.x algorithm stage by stage
 * with the library defaults the reference never changes: quad_decimate=2, quad_sigma=0, refine_edges=1,
 * decode_sharpening=0.25, nthreads=1, qtp {min_cluster_pixels=5, max_nmaxima=10, critical_rad=10deg,
 * max_line_fit_mse=10, min_white_black_diff=5, deglitch=0}.
 *
 * PARITY UNPINNED against the reference (it has no golden vectors).  Pins used instead:
 * tag36h11 known-answer codes, and id/corner agreement with cv2.aruco on synthetic frames (tests/).
 *
 * Frozen choices where upstream behaviour is order dependent (documented in DESIGN.md):
 *   - clusters are visited in ascending 64-bit cluster id (upstream: hash-bucket order);
 *   - points of a cluster enter the slope sort in scan order (y, x, neighbour order);
 *   - the final detection list is sorted by (id, hamming, centre x, centre y) (upstream: qsort by id).
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const uint64_t kCodes[587] = {
#include "tag36h11_codes.inc"
};

/* tag36h11.c: bit walk (four 9-bit quadrants) */
const int kBitX[36] = {1,2,3,4,5,2,3,4,3, 6,6,6,6,6,5,5,5,4, 6,5,4,3,2,5,4,3,4, 1,1,1,1,1,2,2,2,3};
const int kBitY[36] = {1,1,1,1,1,2,2,2,3, 1,2,3,4,5,2,3,4,3, 6,6,6,6,6,5,5,5,4, 6,5,4,3,2,5,4,3,4};
const int kNBits = 36, kNCodes = 587, kWidthAtBorder = 8, kTotalWidth = 10;

struct Image {
    int w, h, stride;
    const uint8_t *buf;
};

/* ---------- A1: image_u8_decimate (integer factor, point sampling) ---------- */
void decimate(const Image &im, int factor, std::vector<uint8_t> &out, int &sw, int &sh)
{
    sw = 1 + (im.w - 1) / factor;
    sh = 1 + (im.h - 1) / factor;
    out.assign((size_t)sw * sh, 0);
    int sy = 0;
    for (int y = 0; y < im.h; y += factor, sy++) {
        int sx = 0;
        for (int x = 0; x < im.w; x += factor, sx++)
            out[(size_t)sy * sw + sx] = im.buf[(size_t)y * im.stride + x];
    }
}

/* ---------- A2: threshold() ---------- */
void threshold(const Image &im, int min_white_black_diff, uint8_t *out /* stride = w */)
{
    const int w = im.w, h = im.h, s = im.stride;
    const int tilesz = 4;
    const int tw = w / tilesz, th = h / tilesz;
    if (tw == 0 || th == 0) { /* upstream would index out of range; define: everything 127 */
        memset(out, 127, (size_t)w * h);
        return;
    }
    std::vector<uint8_t> im_max((size_t)tw * th), im_min((size_t)tw * th);
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            uint8_t mx = 0, mn = 255;
            for (int dy = 0; dy < tilesz; dy++)
                for (int dx = 0; dx < tilesz; dx++) {
                    uint8_t v = im.buf[(size_t)(ty * tilesz + dy) * s + tx * tilesz + dx];
                    if (v < mn) mn = v;
                    if (v > mx) mx = v;
                }
            im_max[(size_t)ty * tw + tx] = mx;
            im_min[(size_t)ty * tw + tx] = mn;
        }
    /* 3x3 max/min over tiles */
    std::vector<uint8_t> mx2((size_t)tw * th), mn2((size_t)tw * th);
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            uint8_t mx = 0, mn = 255;
            for (int dy = -1; dy <= 1; dy++) {
                if (ty + dy < 0 || ty + dy >= th) continue;
                for (int dx = -1; dx <= 1; dx++) {
                    if (tx + dx < 0 || tx + dx >= tw) continue;
                    uint8_t m = im_max[(size_t)(ty + dy) * tw + tx + dx];
                    if (m > mx) mx = m;
                    m = im_min[(size_t)(ty + dy) * tw + tx + dx];
                    if (m < mn) mn = m;
                }
            }
            mx2[(size_t)ty * tw + tx] = mx;
            mn2[(size_t)ty * tw + tx] = mn;
        }
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            int mn = mn2[(size_t)ty * tw + tx], mx = mx2[(size_t)ty * tw + tx];
            if (mx - mn < min_white_black_diff) {
                for (int dy = 0; dy < tilesz; dy++)
                    for (int dx = 0; dx < tilesz; dx++)
                        out[(size_t)(ty * tilesz + dy) * w + tx * tilesz + dx] = 127;
                continue;
            }
            uint8_t thresh = (uint8_t)(mn + (mx - mn) / 2);
            for (int dy = 0; dy < tilesz; dy++)
                for (int dx = 0; dx < tilesz; dx++) {
                    int y = ty * tilesz + dy, x = tx * tilesz + dx;
                    out[(size_t)y * w + x] = im.buf[(size_t)y * s + x] > thresh ? 255 : 0;
                }
        }
    /* partial tiles on the right / bottom reuse the last full tile, WITHOUT the low-contrast test */
    for (int y = 0; y < h; y++) {
        int x0 = (y >= th * tilesz) ? 0 : tw * tilesz;
        int ty = y / tilesz;
        if (ty >= th) ty = th - 1;
        for (int x = x0; x < w; x++) {
            int tx = x / tilesz;
            if (tx >= tw) tx = tw - 1;
            int mx = mx2[(size_t)ty * tw + tx], mn = mn2[(size_t)ty * tw + tx];
            int thresh = mn + (mx - mn) / 2;
            out[(size_t)y * w + x] = im.buf[(size_t)y * s + x] > thresh ? 255 : 0;
        }
    }
}

/* ---------- A3: unionfind.h + connected_components() ---------- */
struct UnionFind {
    std::vector<uint32_t> parent, size; /* size holds (elements - 1) like upstream */
    explicit UnionFind(uint32_t n) : parent(n, 0xffffffffu), size(n, 0) {}
    uint32_t rep(uint32_t id)
    {
        uint32_t root = id;
        if (parent[root] == 0xffffffffu) return root; /* lazily initialised singleton */
        while (parent[root] != root) root = parent[root];
        while (parent[id] != root) { /* path compression */
            uint32_t t = parent[id];
            parent[id] = root;
            id = t;
        }
        return root;
    }
    uint32_t set_size(uint32_t id) { return size[rep(id)] + 1; }
    void connect(uint32_t a, uint32_t b)
    {
        uint32_t ar = rep(a), br = rep(b);
        if (parent[ar] == 0xffffffffu) parent[ar] = ar;
        if (parent[br] == 0xffffffffu) parent[br] = br;
        if (ar == br) return;
        uint32_t as = size[ar] + 1, bs = size[br] + 1;
        if (as > bs) { parent[br] = ar; size[ar] += bs; }
        else         { parent[ar] = br; size[br] += as; }
    }
};

void connected_components(const uint8_t *t, int w, int h, UnionFind &uf)
{
#define DO_UF(dx, dy) if (t[(size_t)(y + (dy)) * w + x + (dx)] == v) uf.connect((uint32_t)(y * w + x), (uint32_t)((y + (dy)) * w + x + (dx)))
    {   /* do_unionfind_first_line */
        int y = 0;
        for (int x = 1; x < w - 1; x++) {
            uint8_t v = t[x];
            if (v == 127) continue;
            DO_UF(-1, 0);
        }
    }
    for (int y = 1; y < h; y++) { /* do_unionfind_line2 */
        uint8_t v_m1_m1, v_0_m1 = t[(size_t)(y - 1) * w + 0], v_1_m1 = t[(size_t)(y - 1) * w + 1];
        uint8_t v_m1_0, v = t[(size_t)y * w + 0];
        for (int x = 1; x < w - 1; x++) {
            v_m1_m1 = v_0_m1;
            v_0_m1 = v_1_m1;
            v_1_m1 = t[(size_t)(y - 1) * w + x + 1];
            v_m1_0 = v;
            v = t[(size_t)y * w + x];
            if (v == 127) continue;
            DO_UF(-1, 0);
            if (x == 1 || !((v_m1_0 == v_m1_m1) && (v_m1_m1 == v_0_m1))) DO_UF(0, -1);
            if (v == 255) {
                if (x == 1 || !(v_m1_0 == v_m1_m1 || v_0_m1 == v_m1_m1)) DO_UF(-1, -1);
                if (!(v_0_m1 == v_1_m1)) DO_UF(1, -1);
            }
        }
    }
#undef DO_UF
}

/* ---------- A4: gradient_clusters() ---------- */
struct Pt {
    uint16_t x, y;
    int16_t gx, gy;
    float slope;
};
struct Cluster {
    uint64_t id;
    std::vector<Pt> pts;
};

void gradient_clusters(const uint8_t *t, int w, int h, UnionFind &uf, std::vector<Cluster> &clusters)
{
    /* chained hash map like upstream (bucket = u64hash_2(id) % nclustermap) */
    int nclustermap = (int)(0.2 * w * h);
    if (nclustermap < 1) nclustermap = 1;
    std::vector<int32_t> heads((size_t)nclustermap, -1);
    std::vector<int32_t> next;
    for (int y = 1; y < h - 1; y++) {
        bool connected_last = false;
        for (int x = 1; x < w - 1; x++) {
            uint8_t v0 = t[(size_t)y * w + x];
            if (v0 == 127) { connected_last = false; continue; }
            uint64_t rep0 = uf.rep((uint32_t)(y * w + x));
            if (uf.set_size((uint32_t)rep0) < 25) { connected_last = false; continue; }
            bool connected = false;
#define DO_CONN(dx, dy)                                                                            \
    do {                                                                                           \
        uint8_t v1 = t[(size_t)(y + (dy)) * w + x + (dx)];                                         \
        if (v0 + v1 == 255) {                                                                      \
            uint64_t rep1 = uf.rep((uint32_t)((y + (dy)) * w + x + (dx)));                         \
            if (uf.set_size((uint32_t)rep1) > 24) {                                                \
                uint64_t cid = rep0 < rep1 ? (rep1 << 32) + rep0 : (rep0 << 32) + rep1;            \
                uint32_t bucket = (uint32_t)(((cid >> 32) ^ cid) & 0xffffffffu) % nclustermap;     \
                int32_t e = heads[bucket];                                                         \
                while (e >= 0 && clusters[e].id != cid) e = next[e];                               \
                if (e < 0) {                                                                       \
                    e = (int32_t)clusters.size();                                                  \
                    clusters.push_back(Cluster{cid, {}});                                          \
                    next.push_back(heads[bucket]);                                                 \
                    heads[bucket] = e;                                                             \
                }                                                                                  \
                Pt p;                                                                              \
                p.x = (uint16_t)(2 * x + (dx)); p.y = (uint16_t)(2 * y + (dy));                    \
                p.gx = (int16_t)((dx) * ((int)v1 - v0)); p.gy = (int16_t)((dy) * ((int)v1 - v0));  \
                p.slope = 0;                                                                       \
                clusters[e].pts.push_back(p);                                                      \
                connected = true;                                                                  \
            }                                                                                      \
        }                                                                                          \
    } while (0)
            DO_CONN(1, 0);
            DO_CONN(0, 1);
            if (!connected_last) DO_CONN(-1, 1);
            connected = false;
            DO_CONN(1, 1);
            connected_last = connected;
#undef DO_CONN
        }
    }
}

/* ---------- A5: fit_quad() and helpers ---------- */
struct LineFitPt { double Mx, My, Mxx, Mxy, Myy, W; };

void ptsort(Pt *pts, int sz, Pt *tmp)
{
    /* upstream: sorting networks for sz<=5, then a merge sort that takes from the SECOND half on ties */
    if (sz <= 1) return;
#define MAYBE_SWAP(a, b) if (pts[a].slope - pts[b].slope > 0) std::swap(pts[a], pts[b])
    if (sz == 2) { MAYBE_SWAP(0, 1); return; }
    if (sz == 3) { MAYBE_SWAP(0, 1); MAYBE_SWAP(1, 2); MAYBE_SWAP(0, 1); return; }
    if (sz == 4) { MAYBE_SWAP(0, 1); MAYBE_SWAP(2, 3); MAYBE_SWAP(0, 2); MAYBE_SWAP(1, 3); MAYBE_SWAP(1, 2); return; }
    if (sz == 5) {
        MAYBE_SWAP(0, 1); MAYBE_SWAP(3, 4); MAYBE_SWAP(2, 4); MAYBE_SWAP(2, 3); MAYBE_SWAP(0, 3);
        MAYBE_SWAP(0, 2); MAYBE_SWAP(1, 4); MAYBE_SWAP(1, 3); MAYBE_SWAP(1, 2);
        return;
    }
#undef MAYBE_SWAP
    memcpy(tmp, pts, sizeof(Pt) * sz);
    int asz = sz / 2, bsz = sz - asz;
    Pt *as = tmp, *bs = tmp + asz;
    ptsort(as, asz, pts);
    ptsort(bs, bsz, pts + asz);
    int apos = 0, bpos = 0, out = 0;
    while (apos < asz && bpos < bsz) {
        if (as[apos].slope - bs[bpos].slope < 0) pts[out++] = as[apos++];
        else pts[out++] = bs[bpos++];
    }
    if (apos < asz) memcpy(&pts[out], &as[apos], (asz - apos) * sizeof(Pt));
    if (bpos < bsz) memcpy(&pts[out], &bs[bpos], (bsz - bpos) * sizeof(Pt));
}

void fit_line(const LineFitPt *lfps, int sz, int i0, int i1, double *lineparm, double *err, double *mse)
{
    double Mx, My, Mxx, Myy, Mxy, W;
    int N;
    if (i0 < i1) {
        N = i1 - i0 + 1;
        Mx = lfps[i1].Mx; My = lfps[i1].My; Mxx = lfps[i1].Mxx; Mxy = lfps[i1].Mxy; Myy = lfps[i1].Myy; W = lfps[i1].W;
        if (i0 > 0) {
            Mx -= lfps[i0 - 1].Mx; My -= lfps[i0 - 1].My; Mxx -= lfps[i0 - 1].Mxx;
            Mxy -= lfps[i0 - 1].Mxy; Myy -= lfps[i0 - 1].Myy; W -= lfps[i0 - 1].W;
        }
    } else {
        Mx = lfps[sz - 1].Mx - lfps[i0 - 1].Mx;   My = lfps[sz - 1].My - lfps[i0 - 1].My;
        Mxx = lfps[sz - 1].Mxx - lfps[i0 - 1].Mxx; Mxy = lfps[sz - 1].Mxy - lfps[i0 - 1].Mxy;
        Myy = lfps[sz - 1].Myy - lfps[i0 - 1].Myy; W = lfps[sz - 1].W - lfps[i0 - 1].W;
        Mx += lfps[i1].Mx; My += lfps[i1].My; Mxx += lfps[i1].Mxx; Mxy += lfps[i1].Mxy; Myy += lfps[i1].Myy; W += lfps[i1].W;
        N = sz - i0 + i1 + 1;
    }
    double Ex = Mx / W, Ey = My / W;
    double Cxx = Mxx / W - Ex * Ex, Cxy = Mxy / W - Ex * Ey, Cyy = Myy / W - Ey * Ey;
    double eig_small = 0.5 * (Cxx + Cyy - sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy)));
    if (lineparm) {
        lineparm[0] = Ex; lineparm[1] = Ey;
        double eig = 0.5 * (Cxx + Cyy + sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy)));
        double nx1 = Cxx - eig, ny1 = Cxy, M1 = nx1 * nx1 + ny1 * ny1;
        double nx2 = Cxy, ny2 = Cyy - eig, M2 = nx2 * nx2 + ny2 * ny2;
        double nx, ny, M;
        if (M1 > M2) { nx = nx1; ny = ny1; M = M1; } else { nx = nx2; ny = ny2; M = M2; }
        double length = sqrtf((float)M);
        if (fabs(length) < 1e-12) { lineparm[2] = lineparm[3] = 0; }
        else { lineparm[2] = nx / length; lineparm[3] = ny / length; }
    }
    if (err) *err = N * eig_small;
    if (mse) *mse = eig_small;
}

int quad_segment_maxima(const orc_params &prm, double cos_critical_rad, int sz, const LineFitPt *lfps, int indices[4])
{
    int ksz = std::min(20, sz / 12);
    if (ksz < 2) return 0;
    std::vector<double> errs(sz);
    for (int i = 0; i < sz; i++) fit_line(lfps, sz, (i + sz - ksz) % sz, (i + ksz) % sz, nullptr, &errs[i], nullptr);
    {   /* low-pass, sigma = 1, cutoff 0.05 -> fsz = 7 */
        std::vector<double> y(sz);
        double sigma = 1, cutoff = 0.05;
        int fsz = (int)(sqrt(-log(cutoff) * 2 * sigma * sigma) + 1);
        fsz = 2 * fsz + 1;
        std::vector<float> f(fsz);
        for (int i = 0; i < fsz; i++) {
            int j = i - fsz / 2;
            f[i] = (float)exp(-j * j / (2 * sigma * sigma));
        }
        for (int iy = 0; iy < sz; iy++) {
            double acc = 0;
            for (int i = 0; i < fsz; i++) acc += errs[(iy + i - fsz / 2 + sz) % sz] * f[i];
            y[iy] = acc;
        }
        errs = y;
    }
    std::vector<int> maxima; std::vector<double> maxima_errs;
    for (int i = 0; i < sz; i++)
        if (errs[i] > errs[(i + 1) % sz] && errs[i] > errs[(i + sz - 1) % sz]) { maxima.push_back(i); maxima_errs.push_back(errs[i]); }
    int nmaxima = (int)maxima.size();
    if (nmaxima < 4) return 0;
    int max_nmaxima = prm.max_nmaxima;
    if (nmaxima > max_nmaxima) {
        std::vector<double> copy = maxima_errs;
        std::sort(copy.begin(), copy.end(), [](double a, double b) { return a > b; });
        double maxima_thresh = copy[max_nmaxima];
        int out = 0;
        for (int in = 0; in < nmaxima; in++) {
            if (maxima_errs[in] <= maxima_thresh) continue;
            maxima[out++] = maxima[in];
        }
        nmaxima = out;
    }
    int best_indices[4] = {0, 0, 0, 0};
    double best_error = HUGE_VALF;
    double err01, err12, err23, err30, mse01, mse12, mse23, mse30;
    double params01[4], params12[4];
    double max_dot = cos_critical_rad;
    for (int m0 = 0; m0 < nmaxima - 3; m0++) {
        int i0 = maxima[m0];
        for (int m1 = m0 + 1; m1 < nmaxima - 2; m1++) {
            int i1 = maxima[m1];
            fit_line(lfps, sz, i0, i1, params01, &err01, &mse01);
            if (mse01 > prm.max_line_fit_mse) continue;
            for (int m2 = m1 + 1; m2 < nmaxima - 1; m2++) {
                int i2 = maxima[m2];
                fit_line(lfps, sz, i1, i2, params12, &err12, &mse12);
                if (mse12 > prm.max_line_fit_mse) continue;
                double dot = params01[2] * params12[2] + params01[3] * params12[3];
                if (fabs(dot) > max_dot) continue;
                for (int m3 = m2 + 1; m3 < nmaxima; m3++) {
                    int i3 = maxima[m3];
                    fit_line(lfps, sz, i2, i3, nullptr, &err23, &mse23);
                    if (mse23 > prm.max_line_fit_mse) continue;
                    fit_line(lfps, sz, i3, i0, nullptr, &err30, &mse30);
                    if (mse30 > prm.max_line_fit_mse) continue;
                    double err = err01 + err12 + err23 + err30;
                    if (err < best_error) {
                        best_error = err;
                        best_indices[0] = i0; best_indices[1] = i1; best_indices[2] = i2; best_indices[3] = i3;
                    }
                }
            }
        }
    }
    if (best_error == HUGE_VALF) return 0;
    for (int i = 0; i < 4; i++) indices[i] = best_indices[i];
    if (best_error / sz < prm.max_line_fit_mse) return 1;
    return 0;
}

struct Quad {
    float p[4][2];
    bool reversed_border;
    double H[9];
    int npoints;
    uint64_t cluster_id;
};

inline double sq(double v) { return v * v; }

int fit_quad(const orc_params &prm, double cos_critical_rad, const Image &im, std::vector<Pt> &cluster, Quad &quad,
             int tag_width, bool normal_border, bool reversed_border, std::vector<Pt> &tmp)
{
    int sz = (int)cluster.size();
    if (sz < 24) return 0;
    uint16_t xmax = cluster[0].x, xmin = cluster[0].x, ymax = cluster[0].y, ymin = cluster[0].y;
    for (int i = 1; i < sz; i++) {
        const Pt &p = cluster[i];
        if (p.x > xmax) xmax = p.x; else if (p.x < xmin) xmin = p.x;
        if (p.y > ymax) ymax = p.y; else if (p.y < ymin) ymin = p.y;
    }
    if ((xmax - xmin) * (ymax - ymin) < tag_width) return 0;
    float cx = (float)((xmin + xmax) * 0.5 + 0.05118);
    float cy = (float)((ymin + ymax) * 0.5 + -0.028581);
    float dot = 0;
    const float quadrants[2][2] = {{-1 * (2 << 15), 0}, {2 * (2 << 15), 2 << 15}};
    for (int i = 0; i < sz; i++) {
        Pt &p = cluster[i];
        float dx = p.x - cx, dy = p.y - cy;
        dot += dx * p.gx + dy * p.gy;
        float quadrant = quadrants[dy > 0][dx > 0];
        if (dy < 0) { dy = -dy; dx = -dx; }
        if (dx < 0) { float t = dx; dx = dy; dy = -t; }
        p.slope = quadrant + dy / dx;
    }
    quad.reversed_border = dot < 0;
    if (!reversed_border && quad.reversed_border) return 0;
    if (!normal_border && !quad.reversed_border) return 0;

    tmp.resize(sz);
    ptsort(cluster.data(), sz, tmp.data());

    /* compute_lfps */
    std::vector<LineFitPt> lfps(sz);
    for (int i = 0; i < sz; i++) {
        const Pt &p = cluster[i];
        if (i > 0) lfps[i] = lfps[i - 1]; else memset(&lfps[0], 0, sizeof(LineFitPt));
        double delta = 0.5;
        double x = p.x * .5 + delta, y = p.y * .5 + delta;
        int ix = (int)x, iy = (int)y;
        double W = 1;
        if (ix > 0 && ix + 1 < im.w && iy > 0 && iy + 1 < im.h) {
            int grad_x = im.buf[(size_t)iy * im.stride + ix + 1] - im.buf[(size_t)iy * im.stride + ix - 1];
            int grad_y = im.buf[(size_t)(iy + 1) * im.stride + ix] - im.buf[(size_t)(iy - 1) * im.stride + ix];
            W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
        }
        double fx = x, fy = y;
        lfps[i].Mx += W * fx; lfps[i].My += W * fy;
        lfps[i].Mxx += W * fx * fx; lfps[i].Mxy += W * fx * fy; lfps[i].Myy += W * fy * fy;
        lfps[i].W += W;
    }

    int indices[4];
    if (!quad_segment_maxima(prm, cos_critical_rad, sz, lfps.data(), indices)) return 0;

    double lines[4][4];
    for (int i = 0; i < 4; i++) {
        int i0 = indices[i], i1 = indices[(i + 1) & 3];
        double mse;
        fit_line(lfps.data(), sz, i0, i1, lines[i], nullptr, &mse);
        if (mse > prm.max_line_fit_mse) return 0;
    }
    for (int i = 0; i < 4; i++) {
        double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
        double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
        double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
        double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
        double det = A00 * A11 - A10 * A01;
        double W00 = A11 / det, W01 = -A01 / det;
        if (fabs(det) < 0.001) return 0;
        double L0 = W00 * B0 + W01 * B1;
        quad.p[i][0] = (float)(lines[i][0] + L0 * A00);
        quad.p[i][1] = (float)(lines[i][1] + L0 * A10);
    }
    {   /* area test */
        double area = 0, length[3], p;
        for (int i = 0; i < 3; i++) {
            int a = i, b = (i + 1) % 3;
            length[i] = sqrt(sq(quad.p[b][0] - quad.p[a][0]) + sq(quad.p[b][1] - quad.p[a][1]));
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        const int idxs[4] = {2, 3, 0, 2};
        for (int i = 0; i < 3; i++) {
            int a = idxs[i], b = idxs[i + 1];
            length[i] = sqrt(sq(quad.p[b][0] - quad.p[a][0]) + sq(quad.p[b][1] - quad.p[a][1]));
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        if (area < 0.95 * tag_width * tag_width) return 0;
    }
    for (int i = 0; i < 4; i++) { /* convexity / winding */
        int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
        double dx1 = quad.p[i1][0] - quad.p[i0][0], dy1 = quad.p[i1][1] - quad.p[i0][1];
        double dx2 = quad.p[i2][0] - quad.p[i1][0], dy2 = quad.p[i2][1] - quad.p[i1][1];
        double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
        if ((cos_dtheta > cos_critical_rad || cos_dtheta < -cos_critical_rad) || dx1 * dy2 < dy1 * dx2) return 0;
    }
    quad.npoints = sz;
    return 1;
}

/* ---------- A6: refine_edges() on the full-resolution image ---------- */
void refine_edges(const Image &im, double quad_decimate, Quad &quad)
{
    double lines[4][4];
    for (int edge = 0; edge < 4; edge++) {
        int a = edge, b = (edge + 1) & 3;
        double nx = quad.p[b][1] - quad.p[a][1];
        double ny = -quad.p[b][0] + quad.p[a][0];
        double mag = sqrt(nx * nx + ny * ny);
        nx /= mag; ny /= mag;
        if (quad.reversed_border) { nx = -nx; ny = -ny; }
        int nsamples = std::max(16, (int)(mag / 8));
        double Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0, N = 0;
        for (int s = 0; s < nsamples; s++) {
            double alpha = (1.0 + s) / (nsamples + 1);
            double x0 = alpha * quad.p[a][0] + (1 - alpha) * quad.p[b][0];
            double y0 = alpha * quad.p[a][1] + (1 - alpha) * quad.p[b][1];
            double Mn = 0, Mcount = 0;
            double range = quad_decimate + 1;
            for (double n = -range; n <= range; n += 0.25) {
                double grange = 1;
                int x1 = (int)(x0 + (n + grange) * nx), y1 = (int)(y0 + (n + grange) * ny);
                if (x1 < 0 || x1 >= im.w || y1 < 0 || y1 >= im.h) continue;
                int x2 = (int)(x0 + (n - grange) * nx), y2 = (int)(y0 + (n - grange) * ny);
                if (x2 < 0 || x2 >= im.w || y2 < 0 || y2 >= im.h) continue;
                int g1 = im.buf[(size_t)y1 * im.stride + x1], g2 = im.buf[(size_t)y2 * im.stride + x2];
                if (g1 < g2) continue;
                double weight = (double)((g2 - g1) * (g2 - g1));
                Mn += weight * n;
                Mcount += weight;
            }
            if (Mcount == 0) continue;
            double n0 = Mn / Mcount;
            double bestx = x0 + n0 * nx, besty = y0 + n0 * ny;
            Mx += bestx; My += besty; Mxx += bestx * bestx; Mxy += bestx * besty; Myy += besty * besty; N++;
        }
        double Ex = Mx / N, Ey = My / N;
        double Cxx = Mxx / N - Ex * Ex, Cxy = Mxy / N - Ex * Ey, Cyy = Myy / N - Ey * Ey;
        double normal_theta = .5 * atan2f((float)(-2 * Cxy), (float)(Cyy - Cxx));
        nx = cosf((float)normal_theta);
        ny = sinf((float)normal_theta);
        lines[edge][0] = Ex; lines[edge][1] = Ey; lines[edge][2] = nx; lines[edge][3] = ny;
    }
    for (int i = 0; i < 4; i++) {
        double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
        double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
        double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
        double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
        double det = A00 * A11 - A10 * A01;
        if (fabs(det) > 0.001) {
            double W00 = A11 / det, W01 = -A01 / det;
            double L0 = W00 * B0 + W01 * B1;
            quad.p[i][0] = (float)(lines[i][0] + L0 * A00);
            quad.p[i][1] = (float)(lines[i][1] + L0 * A10);
        }
    }
}

/* ---------- A7: homography_compute2 + invertibility ---------- */
bool homography_compute2(const double c[4][4], double H[9])
{
    double A[72];
    for (int i = 0; i < 4; i++) {
        double *r0 = &A[(2 * i) * 9], *r1 = &A[(2 * i + 1) * 9];
        r0[0] = c[i][0]; r0[1] = c[i][1]; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0;
        r0[6] = -c[i][0] * c[i][2]; r0[7] = -c[i][1] * c[i][2]; r0[8] = c[i][2];
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = c[i][0]; r1[4] = c[i][1]; r1[5] = 1;
        r1[6] = -c[i][0] * c[i][3]; r1[7] = -c[i][1] * c[i][3]; r1[8] = c[i][3];
    }
    const double epsilon = 1e-10;
    for (int col = 0; col < 8; col++) {
        double max_val = 0; int max_val_idx = -1;
        for (int row = col; row < 8; row++) {
            double val = fabs(A[row * 9 + col]);
            if (val > max_val) { max_val = val; max_val_idx = row; }
        }
        if (max_val_idx < 0) return false;
        if (max_val < epsilon) return false;
        if (max_val_idx != col)
            for (int i = col; i < 9; i++) std::swap(A[col * 9 + i], A[max_val_idx * 9 + i]);
        for (int i = col + 1; i < 8; i++) {
            double f = A[i * 9 + col] / A[col * 9 + col];
            A[i * 9 + col] = 0;
            for (int j = col + 1; j < 9; j++) A[i * 9 + j] -= f * A[col * 9 + j];
        }
    }
    for (int col = 7; col >= 0; col--) {
        double sum = 0;
        for (int i = col + 1; i < 8; i++) sum += A[col * 9 + i] * A[i * 9 + 8];
        A[col * 9 + 8] = (A[col * 9 + 8] - sum) / A[col * 9 + col];
    }
    for (int i = 0; i < 8; i++) H[i] = A[i * 9 + 8];
    H[8] = 1;
    return true;
}

/* matd_inverse() of a 3x3 goes through matd_plu(); only its singular flag matters here (MATD_EPS 1e-8) */
bool mat33_plu_nonsingular(const double H[9])
{
    double lu[9];
    memcpy(lu, H, sizeof(lu));
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < 3; i++) {
            int kmax = i < j ? i : j;
            double acc = 0;
            for (int k = 0; k < kmax; k++) acc += lu[i * 3 + k] * lu[k * 3 + j];
            lu[i * 3 + j] -= acc;
        }
        int p = j;
        for (int i = j + 1; i < 3; i++)
            if (fabs(lu[i * 3 + j]) > fabs(lu[p * 3 + j])) p = i;
        if (p != j)
            for (int k = 0; k < 3; k++) std::swap(lu[p * 3 + k], lu[j * 3 + k]);
        double LUjj = lu[j * 3 + j];
        if (fabs(LUjj) < 1e-8) return false;
        for (int i = j + 1; i < 3; i++) lu[i * 3 + j] /= LUjj;
    }
    return true;
}

bool quad_update_homographies(Quad &quad)
{
    double corr[4][4];
    for (int i = 0; i < 4; i++) {
        corr[i][0] = (i == 0 || i == 3) ? -1 : 1;
        corr[i][1] = (i == 0 || i == 1) ? -1 : 1;
        corr[i][2] = quad.p[i][0];
        corr[i][3] = quad.p[i][1];
    }
    if (!homography_compute2(corr, quad.H)) return false;
    return mat33_plu_nonsingular(quad.H);
}

inline void homography_project(const double H[9], double x, double y, double *ox, double *oy)
{
    double xx = H[0] * x + H[1] * y + H[2];
    double yy = H[3] * x + H[4] * y + H[5];
    double zz = H[6] * x + H[7] * y + H[8];
    *ox = xx / zz;
    *oy = yy / zz;
}

/* ---------- A8: quad_decode() ---------- */
struct GrayModel { double A[3][3], B[3], C[3]; };
void gm_add(GrayModel &gm, double x, double y, double gray)
{
    gm.A[0][0] += x * x; gm.A[0][1] += x * y; gm.A[0][2] += x;
    gm.A[1][1] += y * y; gm.A[1][2] += y; gm.A[2][2] += 1;
    gm.B[0] += x * gray; gm.B[1] += y * gray; gm.B[2] += gray;
}
void gm_solve(GrayModel &gm)
{
    const double *A = &gm.A[0][0];
    double L[9], M[9];
    /* mat33_chol */
    L[0] = sqrt(A[0]); L[3] = A[1] / L[0]; L[6] = A[2] / L[0];
    L[4] = sqrt(A[4] - L[3] * L[3]); L[7] = (A[5] - L[3] * L[6]) / L[4];
    L[8] = sqrt(A[8] - L[6] * L[6] - L[7] * L[7]);
    L[1] = L[2] = L[5] = 0;
    /* mat33_lower_tri_inv */
    M[0] = 1 / L[0]; M[3] = -L[3] * M[0] / L[4]; M[4] = 1 / L[4];
    M[6] = (-L[6] * M[0] - L[7] * M[3]) / L[8]; M[7] = -L[7] * M[4] / L[8]; M[8] = 1 / L[8];
    double tmp[3];
    tmp[0] = M[0] * gm.B[0];
    tmp[1] = M[3] * gm.B[0] + M[4] * gm.B[1];
    tmp[2] = M[6] * gm.B[0] + M[7] * gm.B[1] + M[8] * gm.B[2];
    gm.C[0] = M[0] * tmp[0] + M[3] * tmp[1] + M[6] * tmp[2];
    gm.C[1] = M[4] * tmp[1] + M[7] * tmp[2];
    gm.C[2] = M[8] * tmp[2];
}
inline double gm_interp(const GrayModel &gm, double x, double y) { return gm.C[0] * x + gm.C[1] * y + gm.C[2]; }

double value_for_pixel(const Image &im, double px, double py)
{
    int x1 = (int)floor(px - 0.5), x2 = (int)ceil(px - 0.5);
    double x = px - 0.5 - x1;
    int y1 = (int)floor(py - 0.5), y2 = (int)ceil(py - 0.5);
    double y = py - 0.5 - y1;
    if (x1 < 0 || x2 >= im.w || y1 < 0 || y2 >= im.h) return -1;
    return im.buf[(size_t)y1 * im.stride + x1] * (1 - x) * (1 - y) + im.buf[(size_t)y1 * im.stride + x2] * x * (1 - y) +
           im.buf[(size_t)y2 * im.stride + x1] * (1 - x) * y + im.buf[(size_t)y2 * im.stride + x2] * x * y;
}

void sharpen(double decode_sharpening, double *values, int size)
{
    std::vector<double> sharpened((size_t)size * size);
    const double kernel[9] = {0, -1, 0, -1, 4, -1, 0, -1, 0};
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            sharpened[y * size + x] = 0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    if ((y + i - 1) < 0 || (y + i - 1) > size - 1 || (x + j - 1) < 0 || (x + j - 1) > size - 1) continue;
                    sharpened[y * size + x] += values[(y + i - 1) * size + (x + j - 1)] * kernel[i * 3 + j];
                }
        }
    for (int i = 0; i < size * size; i++) values[i] = values[i] + decode_sharpening * sharpened[i];
}

inline uint64_t rotate90(uint64_t w, int numBits)
{
    int p = numBits; uint64_t l = 0;
    if (numBits % 4 == 1) { p = numBits - 1; l = 1; }
    w = ((w >> l) << (p / 4 + l)) | (w >> (3 * p / 4 + l) << l) | (w & l);
    w &= ((UINT64_C(1) << numBits) - 1);
    return w;
}

struct DecodeEntry { int id, hamming, rotation; };

/* quick_decode_codeword(): first rotation (0..3) for which a code within maxhamming exists.  tag36h11 has
 * minimum distance 11 over all rotations, so for maxhamming <= 3 the hash-table lookup upstream and this
 * exhaustive popcount search return the same (id, hamming, rotation). */
void quick_decode(uint64_t rcode, int maxhamming, DecodeEntry &e)
{
    for (int ridx = 0; ridx < 4; ridx++) {
        for (int id = 0; id < kNCodes; id++) {
            int d = __builtin_popcountll(rcode ^ kCodes[id]);
            if (d <= maxhamming) { e.id = id; e.hamming = d; e.rotation = ridx; return; }
        }
        rcode = rotate90(rcode, kNBits);
    }
    e.id = 65535; e.hamming = 255; e.rotation = 0;
}

float quad_decode(const orc_params &prm, const Image &im, const Quad &quad, DecodeEntry &entry)
{
    const float patterns[] = {
        -0.5f, 0.5f, 0, 1, 1,   0.5f, 0.5f, 0, 1, 0,
        kWidthAtBorder + 0.5f, .5f, 0, 1, 1,   kWidthAtBorder - 0.5f, .5f, 0, 1, 0,
        0.5f, -0.5f, 1, 0, 1,   0.5f, 0.5f, 1, 0, 0,
        0.5f, kWidthAtBorder + 0.5f, 1, 0, 1,   0.5f, kWidthAtBorder - 0.5f, 1, 0, 0};
    GrayModel white, black;
    memset(&white, 0, sizeof(white)); memset(&black, 0, sizeof(black));
    for (int pi = 0; pi < 8; pi++) {
        const float *pattern = &patterns[pi * 5];
        int is_white = (int)pattern[4];
        for (int i = 0; i < kWidthAtBorder; i++) {
            double tagx01 = (pattern[0] + i * pattern[2]) / (kWidthAtBorder);
            double tagy01 = (pattern[1] + i * pattern[3]) / (kWidthAtBorder);
            double tagx = 2 * (tagx01 - 0.5), tagy = 2 * (tagy01 - 0.5);
            double px, py;
            homography_project(quad.H, tagx, tagy, &px, &py);
            int ix = (int)px, iy = (int)py;
            if (ix < 0 || iy < 0 || ix >= im.w || iy >= im.h) continue;
            int v = im.buf[(size_t)iy * im.stride + ix];
            if (is_white) gm_add(white, tagx, tagy, v); else gm_add(black, tagx, tagy, v);
        }
    }
    gm_solve(white);
    gm_solve(black);
    /* reversed_border is false for tag36h11 */
    if ((gm_interp(white, 0, 0) - gm_interp(black, 0, 0) < 0) != false) return -1;

    float black_score = 0, white_score = 0, black_score_count = 1, white_score_count = 1;
    double values[kTotalWidth * kTotalWidth];
    memset(values, 0, sizeof(values));
    const int min_coord = (kWidthAtBorder - kTotalWidth) / 2;
    for (int i = 0; i < kNBits; i++) {
        int bity = kBitY[i], bitx = kBitX[i];
        double tagx01 = (bitx + 0.5) / (kWidthAtBorder), tagy01 = (bity + 0.5) / (kWidthAtBorder);
        double tagx = 2 * (tagx01 - 0.5), tagy = 2 * (tagy01 - 0.5);
        double px, py;
        homography_project(quad.H, tagx, tagy, &px, &py);
        double v = value_for_pixel(im, px, py);
        if (v == -1) continue;
        double thresh = (gm_interp(black, tagx, tagy) + gm_interp(white, tagx, tagy)) / 2.0;
        values[kTotalWidth * (bity - min_coord) + bitx - min_coord] = v - thresh;
    }
    sharpen(prm.decode_sharpening, values, kTotalWidth);
    uint64_t rcode = 0;
    for (int i = 0; i < kNBits; i++) {
        int bity = kBitY[i], bitx = kBitX[i];
        rcode = (rcode << 1);
        double v = values[(bity - min_coord) * kTotalWidth + bitx - min_coord];
        if (v > 0) { white_score += (float)v; white_score_count++; rcode |= 1; }
        else { black_score -= (float)v; black_score_count++; }
    }
    quick_decode(rcode, prm.bits_corrected, entry);
    return fminf(white_score / white_score_count, black_score / black_score_count);
}

/* ---------- A9: reconcile (g2d polygon overlap) ---------- */
inline double cross2(const double a[2], const double b[2], const double c[2])
{
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}
bool seg_intersect(const double p0[2], const double p1[2], const double q0[2], const double q1[2])
{
    /* proper or touching intersection of closed segments */
    double d1 = cross2(q0, q1, p0), d2 = cross2(q0, q1, p1), d3 = cross2(p0, p1, q0), d4 = cross2(p0, p1, q1);
    if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0))) return true;
    auto on = [](const double a[2], const double b[2], const double c[2]) {
        return std::min(a[0], b[0]) <= c[0] && c[0] <= std::max(a[0], b[0]) && std::min(a[1], b[1]) <= c[1] && c[1] <= std::max(a[1], b[1]);
    };
    if (d1 == 0 && on(q0, q1, p0)) return true;
    if (d2 == 0 && on(q0, q1, p1)) return true;
    if (d3 == 0 && on(p0, p1, q0)) return true;
    if (d4 == 0 && on(p0, p1, q1)) return true;
    return false;
}
bool poly_contains(const double poly[4][2], const double q[2])
{
    /* even-odd crossing test */
    bool in = false;
    for (int i = 0, j = 3; i < 4; j = i++) {
        if (((poly[i][1] > q[1]) != (poly[j][1] > q[1])) &&
            (q[0] < (poly[j][0] - poly[i][0]) * (q[1] - poly[i][1]) / (poly[j][1] - poly[i][1]) + poly[i][0]))
            in = !in;
    }
    return in;
}
bool polygons_overlap(const double a[4][2], const double b[4][2])
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersect(a[i], a[(i + 1) & 3], b[j], b[(j + 1) & 3])) return true;
    /* no edge crossing: one contains the other or they are disjoint; quads are convex so the vertex mean is interior */
    double ca[2] = {0, 0}, cb[2] = {0, 0};
    for (int i = 0; i < 4; i++) { ca[0] += a[i][0] / 4; ca[1] += a[i][1] / 4; cb[0] += b[i][0] / 4; cb[1] += b[i][1] / 4; }
    if (poly_contains(a, cb)) return true;
    if (poly_contains(b, ca)) return true;
    return false;
}
inline int prefer_smaller(int pref, double q0, double q1)
{
    if (pref) return pref;
    if (q0 < q1) return -1;
    if (q1 < q0) return 1;
    return 0;
}

/* ext_map: an externally supplied ternary map (0 / 127 / 255, pitch = decimated width) used in place of threshold()'s -- the
   CAT path (orc_detect_with_map): CAT's own colour map, then upstream's stages from connected_components() on */
int detect_impl(const uint8_t *buf, int W, int H, int stride, const orc_params &prm, orc_detection *out, int cap, orc_taps *taps,
                const uint8_t *ext_map = nullptr)
{
    if (!buf || W <= 0 || H <= 0 || stride < W) return -1;
    if (prm.bits_corrected < 0 || prm.bits_corrected > 3) return -2;
    Image im_orig{W, H, stride, buf};
    int factor = (int)prm.quad_decimate;
    if (factor < 1 || (float)factor != prm.quad_decimate) return -3; /* 1.5 not supported */
    std::vector<uint8_t> dec;
    Image quad_im = im_orig;
    if (prm.quad_decimate > 1) {
        int sw, sh;
        decimate(im_orig, factor, dec, sw, sh);
        quad_im = Image{sw, sh, sw, dec.data()};
    }
    const int w = quad_im.w, h = quad_im.h;
    std::vector<uint8_t> thr((size_t)w * h);
    if (ext_map) std::memcpy(thr.data(), ext_map, thr.size());
    else threshold(quad_im, prm.min_white_black_diff, thr.data());
    UnionFind uf((uint32_t)((size_t)w * h));
    connected_components(thr.data(), w, h, uf);
    std::vector<Cluster> clusters;
    gradient_clusters(thr.data(), w, h, uf, clusters);

    /* canonical labels (min pixel index per component) for taps and for deterministic cluster ordering */
    std::vector<uint32_t> minidx;
    auto build_minidx = [&]() {
        minidx.assign((size_t)w * h, 0xffffffffu);
        for (uint32_t i = 0; i < (uint32_t)((size_t)w * h); i++) {
            uint32_t r = uf.rep(i);
            if (minidx[r] == 0xffffffffu) minidx[r] = i; /* ascending scan: first hit is the minimum */
        }
    };
    build_minidx();
    for (auto &c : clusters) {
        uint64_t a = minidx[(uint32_t)(c.id & 0xffffffffu)], b = minidx[(uint32_t)(c.id >> 32)];
        c.id = a < b ? (b << 32) | a : (a << 32) | b;
    }
    std::sort(clusters.begin(), clusters.end(), [](const Cluster &a, const Cluster &b) { return a.id < b.id; });

    if (taps) {
        taps->w = w; taps->h = h;
        if (taps->thresh) memcpy(taps->thresh, thr.data(), (size_t)w * h);
        if (taps->labels || taps->comp_size)
            for (uint32_t i = 0; i < (uint32_t)((size_t)w * h); i++) {
                if (taps->labels) taps->labels[i] = minidx[uf.rep(i)];
                if (taps->comp_size) taps->comp_size[i] = uf.set_size(i);
            }
        int64_t np = 0; int nc = 0;
        for (auto &c : clusters) { np += (int64_t)c.pts.size(); if ((int)c.pts.size() >= prm.min_cluster_pixels) nc++; }
        taps->npoints = np; taps->nclusters = nc;
        if (taps->pts && taps->pts_cluster) {
            int64_t k = 0;
            for (auto &c : clusters) {
                std::vector<Pt> s = c.pts;
                std::sort(s.begin(), s.end(), [](const Pt &a, const Pt &b) {
                    if (a.y != b.y) return a.y < b.y;
                    if (a.x != b.x) return a.x < b.x;
                    if (a.gx != b.gx) return a.gx < b.gx;
                    return a.gy < b.gy;
                });
                for (auto &p : s) {
                    if (k >= taps->pts_cap) break;
                    taps->pts[k * 4 + 0] = (int16_t)p.x; taps->pts[k * 4 + 1] = (int16_t)p.y;
                    taps->pts[k * 4 + 2] = p.gx; taps->pts[k * 4 + 3] = p.gy;
                    taps->pts_cluster[k] = c.id;
                    k++;
                }
            }
        }
    }

    /* fit_quads */
    int min_tag_width = kWidthAtBorder;
    min_tag_width = (int)(min_tag_width / prm.quad_decimate);
    if (min_tag_width < 3) min_tag_width = 3;
    const double cos_critical_rad = cos(prm.critical_rad);
    std::vector<Quad> quads;
    std::vector<Pt> tmp;
    for (auto &c : clusters) {
        if ((int)c.pts.size() < prm.min_cluster_pixels) continue;
        if ((int)c.pts.size() > 3 * (2 * w + 2 * h)) continue;
        Quad q;
        memset(&q, 0, sizeof(q));
        q.cluster_id = c.id;
        if (fit_quad(prm, cos_critical_rad, quad_im, c.pts, q, min_tag_width, true, false, tmp)) quads.push_back(q);
    }
    if (taps) {
        taps->nquads = (int)quads.size();
        if (taps->quads)
            for (int i = 0; i < (int)quads.size() && i < taps->quads_cap; i++) {
                memcpy(taps->quads[i].p, quads[i].p, sizeof(quads[i].p));
                taps->quads[i].reversed_border = quads[i].reversed_border;
                taps->quads[i].npoints = quads[i].npoints;
                taps->quads[i].cluster_id = quads[i].cluster_id;
            }
    }
    /* back to full resolution */
    if (prm.quad_decimate > 1)
        for (auto &q : quads)
            for (int j = 0; j < 4; j++) {
                q.p[j][0] = (float)((q.p[j][0] - 0.5) * prm.quad_decimate + 0.5);
                q.p[j][1] = (float)((q.p[j][1] - 0.5) * prm.quad_decimate + 0.5);
            }
    /* decode */
    std::vector<orc_detection> dets;
    for (auto &q : quads) {
        if (prm.refine_edges) refine_edges(im_orig, prm.quad_decimate, q);
        if (!quad_update_homographies(q)) continue;
        DecodeEntry entry;
        float decision_margin = quad_decode(prm, im_orig, q, entry);
        if (decision_margin >= 0 && entry.hamming < 255) {
            orc_detection d;
            memset(&d, 0, sizeof(d));
            d.id = entry.id; d.hamming = entry.hamming; d.decision_margin = decision_margin;
            double theta = entry.rotation * M_PI / 2.0;
            double c = cos(theta), s = sin(theta);
            const double R[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double acc = 0;
                    for (int k = 0; k < 3; k++) acc += q.H[i * 3 + k] * R[k * 3 + j];
                    d.H[i * 3 + j] = acc;
                }
            homography_project(d.H, 0, 0, &d.c[0], &d.c[1]);
            for (int i = 0; i < 4; i++) {
                int tcx = (i == 1 || i == 2) ? 1 : -1;
                int tcy = (i < 2) ? 1 : -1;
                homography_project(d.H, tcx, tcy, &d.p[i][0], &d.p[i][1]);
            }
            dets.push_back(d);
        }
    }
    /* reconcile.  Upstream walks the detections in quad order, which follows its hash-bucket order; the list is put
       into a canonical order first (frozen choice) so that the outcome does not depend on that order. */
    auto det_less = [](const orc_detection &a, const orc_detection &b) {
        if (a.id != b.id) return a.id < b.id;
        if (a.hamming != b.hamming) return a.hamming < b.hamming;
        if (a.c[0] != b.c[0]) return a.c[0] < b.c[0];
        return a.c[1] < b.c[1];
    };
    std::sort(dets.begin(), dets.end(), det_less);
    for (int i0 = 0; i0 < (int)dets.size(); i0++) {
        bool removed0 = false;
        for (int i1 = i0 + 1; i1 < (int)dets.size(); i1++) {
            orc_detection &d0 = dets[i0], &d1 = dets[i1];
            if (d0.id != d1.id) continue;
            if (!polygons_overlap(d0.p, d1.p)) continue;
            int pref = 0;
            pref = prefer_smaller(pref, d0.hamming, d1.hamming);
            pref = prefer_smaller(pref, -d0.decision_margin, -d1.decision_margin);
            for (int i = 0; i < 4; i++) {
                pref = prefer_smaller(pref, d0.p[i][0], d1.p[i][0]);
                pref = prefer_smaller(pref, d0.p[i][1], d1.p[i][1]);
            }
            if (pref < 0) { /* keep d0 */
                dets[i1] = dets.back(); dets.pop_back(); i1--;
            } else {        /* keep d1 */
                dets[i0] = dets.back(); dets.pop_back(); i0--; removed0 = true;
                break;
            }
        }
        (void)removed0;
    }
    std::sort(dets.begin(), dets.end(), [](const orc_detection &a, const orc_detection &b) {
        if (a.id != b.id) return a.id < b.id;
        if (a.hamming != b.hamming) return a.hamming < b.hamming;
        if (a.c[0] != b.c[0]) return a.c[0] < b.c[0];
        return a.c[1] < b.c[1];
    });
    int n = std::min((int)dets.size(), cap);
    for (int i = 0; i < n; i++) out[i] = dets[i];
    return n;
}

}  // namespace

extern "C" {

void orc_default_params(orc_params *p)
{
    p->quad_decimate = 2.0f; p->refine_edges = 1; p->decode_sharpening = 0.25;
    p->min_cluster_pixels = 5; p->max_nmaxima = 10; p->critical_rad = (float)(10 * M_PI / 180);
    p->max_line_fit_mse = 10.0f; p->min_white_black_diff = 5; p->bits_corrected = 3;
}

const uint64_t *orc_tag36h11_codes(int *ncodes) { if (ncodes) *ncodes = kNCodes; return kCodes; }

void orc_decimated_size(int W, int H, float quad_decimate, int *w, int *h)
{
    int f = (int)quad_decimate;
    if (quad_decimate > 1) { *w = 1 + (W - 1) / f; *h = 1 + (H - 1) / f; } else { *w = W; *h = H; }
}

int orc_threshold(const uint8_t *im, int W, int H, int stride, const orc_params *prm, uint8_t *out)
{
    Image o{W, H, stride, im};
    int factor = (int)prm->quad_decimate;
    if (prm->quad_decimate > 1) {
        std::vector<uint8_t> dec; int sw, sh;
        decimate(o, factor, dec, sw, sh);
        Image q{sw, sh, sw, dec.data()};
        threshold(q, prm->min_white_black_diff, out);
    } else {
        threshold(o, prm->min_white_black_diff, out);
    }
    return 0;
}

int orc_detect(const uint8_t *im, int W, int H, int stride, const orc_params *prm, orc_detection *out, int cap, orc_taps *taps)
{
    return detect_impl(im, W, H, stride, *prm, out, cap, taps);
}

/* book/src/maintenance/apriltags.md:58-60 ("Decoding tags is done pretty much the same way the C library does it") with
   crates/chalkydri-apriltags/src/lib.rs:551-613 (the commented-out cluster() over connected_components()): the ternary map comes
   from CAT's own thresholding, everything after it is upstream's pipeline */
int orc_detect_with_map(const uint8_t *im, int W, int H, int stride, const uint8_t *map, const orc_params *prm, orc_detection *out, int cap)
{
    return detect_impl(im, W, H, stride, *prm, out, cap, nullptr, map);
}

int orc_detect_batch(const uint8_t *frames, int W, int H, int stride, int64_t frame_stride, int batch,
                     const orc_params *prm, orc_detection *out, int cap, int32_t *counts, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::atomic<int> err(0);
    auto work = [&]() {
        for (;;) {
            int b = next.fetch_add(1);
            if (b >= batch) break;
            int n = detect_impl(frames + (size_t)b * frame_stride, W, H, stride, *prm, out + (size_t)b * cap, cap, nullptr);
            if (n < 0) { err.store(n); counts[b] = 0; } else counts[b] = n;
        }
    };
    if (nthreads == 1) work();
    else {
        std::vector<std::thread> ts;
        for (int i = 0; i < nthreads; i++) ts.emplace_back(work);
        for (auto &t : ts) t.join();
    }
    return err.load();
}

}  // extern "C"
