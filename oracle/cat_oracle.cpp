/*
 * oracle/cat_oracle.cpp -- CPU restatement of the in-house "CAT" detector stages (TEST INFRASTRUCTURE ONLY).
 *
 * Follows /root/reference/crates/chalkydri-apriltags/src/lib.rs and src/utils.rs; citations per function.
 * Third-party arithmetic restated: statrs 0.18.0 `Data` order statistics (Cargo.toml:16), i.e. the R-8
 * quantile h = (n + 1/3) tau + 1/3 with linear interpolation, and median = middle / mean of two middles.
 * PARITY UNPINNED (the reference has no test or fixture for CAT; its bench input test.png is absent).
 *
 * Defined behaviour where the reference is undefined: `check_edge` (lib.rs:430-461) subtracts 5 from
 * usize coordinates without a bounds check; samples that fall outside the image are treated as Color::Other.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
enum : uint8_t { BLACK = 0, WHITE = 1, OTHER = 2 };   /* utils.rs:1-6 */

inline uint8_t grayscale(const uint8_t *p)
{
    /* utils.rs:43: (r as f32).mul_add(0.33, (g as f32).mul_add(0.33, (b as f32) * 0.33)) as u8 */
    float v = fmaf((float)p[0], 0.33f, fmaf((float)p[1], 0.33f, (float)p[2] * 0.33f));
    if (!(v > 0.0f)) return 0;            /* Rust `as u8` saturates */
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

/* statrs OrderStatistics::quantile on sorted data */
double quantile(const double *sorted, int n, double tau)
{
    double h = ((double)n + 1.0 / 3.0) * tau + 1.0 / 3.0;
    int64_t hf = (int64_t)h;
    if (hf <= 0 || tau == 0.0) return sorted[0];
    if (hf >= n) return sorted[n - 1];
    double a = sorted[hf - 1], b = sorted[hf];
    return a + (h - (double)hf) * (b - a);
}
inline uint8_t f64_as_u8(double v)
{
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

struct UF {
    std::vector<uint32_t> parent, size;
    explicit UF(size_t n) : parent(n), size(n, 1) { for (size_t i = 0; i < n; i++) parent[i] = (uint32_t)i; }
    uint32_t find(uint32_t id)
    {
        uint32_t root = id;
        while (parent[root] != root) root = parent[root];
        while (parent[id] != root) { uint32_t t = parent[id]; parent[id] = root; id = t; }   /* full compression, lib.rs:67-75 */
        return root;
    }
    void unite(uint32_t a, uint32_t b)
    {
        uint32_t r1 = find(a), r2 = find(b);
        if (r1 == r2) return;
        if (size[r1] < size[r2]) { parent[r1] = r2; size[r2] += size[r1]; }
        else { parent[r2] = r1; size[r1] += size[r2]; }
    }
};
}  // namespace

extern "C" {

uint8_t orc_cat_grayscale(uint8_t r, uint8_t g, uint8_t b)
{
    uint8_t p[3] = {r, g, b};
    return grayscale(p);
}

void orc_cat_calc_otsu(const uint8_t *rgb, int w, int h, uint8_t *color)
{
    /* lib.rs:191-259 */
    std::vector<uint8_t> gray((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) gray[i] = grayscale(rgb + 3 * i);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int x_min = x >= 2 ? x - 2 : 0, x_max = std::min(x + 2, w - 1);
            int y_min = y >= 2 ? y - 2 : 0, y_max = std::min(y + 2, h - 1);
            double px[25];
            int n = 0;
            for (int xx = x_min; xx <= x_max; xx++)
                for (int yy = y_min; yy <= y_max; yy++) px[n++] = (double)gray[(size_t)yy * w + xx];
            std::sort(px, px + n);
            size_t i = (size_t)y * w + x;
            uint8_t p = gray[i];
            uint8_t out;
            if ((y > 0 && x > 0) && (px[n - 1] - px[0]) < 5.0) {
                int k = n / 2;
                double med = (n % 2 != 0) ? px[k] : (px[k - 1] + px[k]) / 2.0;
                out = med < 60.0 ? BLACK : (med > 160.0 ? WHITE : OTHER);
            } else {
                if (p >= f64_as_u8(quantile(px, n, 0.75))) out = WHITE;
                else if (p <= f64_as_u8(quantile(px, n, 0.25))) out = BLACK;
                else out = OTHER;
            }
            color[i] = out;
        }
}

void orc_cat_thresh(const uint8_t *rgb, int w, int h, uint8_t *color)
{
    /* lib.rs:319-334 */
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t g = grayscale(rgb + 3 * i);
        color[i] = g < 60 ? BLACK : (g > 160 ? WHITE : OTHER);
    }
}

int64_t orc_cat_detect_corners(const uint8_t *c, int w, int h, int32_t *xy, int64_t cap)
{
    /* lib.rs:291-309 (x outer, y inner, inclusive upper bounds w-3 / h-3) and :345-400 */
    int64_t n = 0;
    auto at = [&](int x, int y) { return c[(size_t)y * w + x]; };
    for (int x = 3; x <= w - 3; x++)
        for (int y = 3; y <= h - 3; y++) {
            /* px() is an unchecked linear index (utils.rs:27-29), so at x = w-3 the reference's (x+3, y-3) / (x+3, y+3) samples
               are column 0 of the NEXT row -- defined behaviour, reproduced here by the same linear indexing.  Only pixels
               whose furthest sample (x+3, y+3) lies beyond the w*h buffer (row h-3, and (w-3, h-4)) read out of bounds in
               the reference (undefined behaviour): those cannot be evaluated and are skipped. */
            if ((size_t)(y + 3) * w + (size_t)(x + 3) >= (size_t)w * h) continue;
            if (at(x, y) != BLACK) continue;
            bool ul = at(x - 1, y - 1) == BLACK, ur = at(x + 1, y - 1) == BLACK;
            bool dl = at(x - 1, y + 1) == BLACK, dr = at(x + 1, y + 1) == BLACK;
            if (!(ul ^ ur ^ dl ^ dr)) continue;
            uint8_t p3 = at(x + 3, y - 3), p7 = at(x + 3, y + 3), p11 = at(x - 3, y + 3), p15 = at(x - 3, y - 3);
            if ((p3 != OTHER && p7 != OTHER && p11 != OTHER && p15 != OTHER) &&
                ((p3 == BLACK) ^ (p7 == BLACK) ^ (p11 == BLACK) ^ (p15 == BLACK))) {
                if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; }
                n++;
            }
        }
    return n;
}

int64_t orc_cat_check_edges(const uint8_t *c, int w, int h, const int32_t *xy, int64_t npts, int32_t *lines, int64_t cap)
{
    /* lib.rs:409-499 */
    const int OFF = 5;
    int64_t n = 0;
    auto at = [&](long x, long y) -> uint8_t { return (x < 0 || y < 0 || x >= w || y >= h) ? (uint8_t)OTHER : c[(size_t)y * w + x]; };
    auto push = [&](int x1, int y1, int x2, int y2) {
        if (n < cap) { lines[4 * n] = x1; lines[4 * n + 1] = y1; lines[4 * n + 2] = x2; lines[4 * n + 3] = y2; }
        n++;
    };
    for (int64_t i = 0; i < npts; i++)
        for (int64_t j = npts - 1; j >= 0; j--) {
            int x1 = xy[2 * i], y1 = xy[2 * i + 1], x2 = xy[2 * j], y2 = xy[2 * j + 1];
            int mx = (x1 + x2) / 2, my = (y1 + y2) / 2;
            int xdiff = std::max(x1, x2) - std::min(x1, x2), ydiff = std::max(y1, y2) - std::min(y1, y2);
            bool vert = x1 == x2 || xdiff < ydiff, horiz = y1 == y2 || ydiff < xdiff;
            int mw1x = (mx + x1) / 2, mw1y = (my + y1) / 2, mw2x = (mx + x2) / 2, mw2y = (my + y2) / 2;
            if (vert) {
                uint8_t r1 = at(mw1x + OFF, mw1y), r2 = at(mw2x + OFF, mw2y), l1 = at(mw1x - OFF, mw1y), l2 = at(mw2x - OFF, mw2y);
                if (l1 != OTHER && l2 != OTHER && r1 != OTHER && r2 != OTHER)
                    if (((l1 == BLACK) ^ (r2 == BLACK)) && ((l2 == BLACK) ^ (r1 == BLACK)) && (l1 == l2)) push(x1, y1, x2, y2);
            }
            if (horiz) {
                uint8_t t1 = at(mw1x, mw1y - OFF), t2 = at(mw2x, mw2y - OFF), b1 = at(mw1x, mw1y + OFF), b2 = at(mw2x, mw2y + OFF);
                if (t1 != OTHER && t2 != OTHER && b1 != OTHER && b2 != OTHER)
                    if (((t1 == BLACK) ^ (b2 == BLACK)) && ((t2 == BLACK) ^ (b1 == BLACK)) && (t1 == t2)) push(x1, y1, x2, y2);
            }
        }
    return n;
}

void orc_cat_connected_components(const uint8_t *c, int w, int h, uint32_t *labels, uint32_t *sizes)
{
    /* lib.rs:501-549 with UnionFind lib.rs:42-113 */
    UF uf((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 1; x < w - 1; x++) {
            uint32_t i = (uint32_t)((size_t)y * w + x);
            uint8_t p = c[i];
            if (p == OTHER) continue;
            if (c[i - 1] == p) uf.unite(i, i - 1);
            if (y > 0) {
                if (c[i - w] == p) uf.unite(i, i - w);
                if (p == WHITE) {
                    if (c[i - w - 1] == p) uf.unite(i, i - w - 1);
                    if (x < w - 1 && c[i - w + 1] == p) uf.unite(i, i - w + 1);
                }
            }
        }
    std::vector<uint32_t> minidx((size_t)w * h, 0xffffffffu);
    for (uint32_t i = 0; i < (uint32_t)((size_t)w * h); i++) {
        uint32_t r = uf.find(i);
        if (minidx[r] == 0xffffffffu) minidx[r] = i;
    }
    for (uint32_t i = 0; i < (uint32_t)((size_t)w * h); i++) {
        uint32_t r = uf.find(i);
        if (labels) labels[i] = minidx[r];
        if (sizes) sizes[i] = uf.size[r];
    }
}

}  // extern "C"
