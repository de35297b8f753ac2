// ccl.cuh -- row A3 of SURVEY.md 8a: connected_components() as a parallel union-find.
//
// Upstream (apriltag_quad_thresh.c do_unionfind_first_line / do_unionfind_line2): for every pixel with
// 1 <= x <= w-2 and v != 127, connect to the left, up, and for white pixels up-left / up-right neighbour of
// equal value, with redundancy guards.  The guards are evaluated here exactly as upstream writes them, so the
// edge set -- and therefore the partition -- is identical; only the tree shape differs (roots are the smallest
// pixel index of each component, which is also the canonical form the oracle reports).
//
// B200 mapping: latency / atomic bound.  Pass 1 resolves 64x16-pixel tiles in shared memory (atomicMin
// union-find on 32-bit smem words), pass 2 stitches tile borders with global atomicMin, pass 3 flattens every
// pixel to its root and histograms component sizes with warp-aggregated atomics (match.any on the label).
#pragma once
#include "common.cuh"

namespace cb {

constexpr int CCL_TW = 64, CCL_TH = 16, CCL_THREADS = 256;

enum : uint32_t { LINK_LEFT = 1, LINK_UP = 2, LINK_UPLEFT = 4, LINK_UPRIGHT = 8 };

// link mask of pixel (x,y) from the ternary map t (pitch tp).
// MODE 0: upstream AprilTag-3 (skip 127, white 255), exact transcription of its redundancy guards.
// MODE 1: CAT connected_components (crates/chalkydri-apriltags/src/lib.rs:501-549): Color map (skip Other = 2,
//         White = 1), same neighbourhood (left, up; white also up-left / up-right), no guards.
template <int MODE>
__device__ __forceinline__ uint32_t link_mask(const uint8_t *__restrict__ t, int tp, int w, int x, int y)
{
    if (x < 1 || x > w - 2) return 0;
    const uint8_t *row = t + (size_t)y * tp;
    const uint32_t v = row[x];
    const uint32_t SKIP = MODE == 0 ? 127u : 2u, WHITE = MODE == 0 ? 255u : 1u;
    if (v == SKIP) return 0;
    uint32_t m = 0;
    const uint32_t v_m1_0 = row[x - 1];
    if (v_m1_0 == v) m |= LINK_LEFT;
    if (y > 0) {
        const uint8_t *up = row - tp;
        const uint32_t v_m1_m1 = up[x - 1], v_0_m1 = up[x], v_1_m1 = up[x + 1];
        if (MODE == 0) {
            if (v_0_m1 == v && (x == 1 || !((v_m1_0 == v_m1_m1) && (v_m1_m1 == v_0_m1)))) m |= LINK_UP;
            if (v == WHITE) {
                if (v_m1_m1 == v && (x == 1 || !(v_m1_0 == v_m1_m1 || v_0_m1 == v_m1_m1))) m |= LINK_UPLEFT;
                if (v_1_m1 == v && !(v_0_m1 == v_1_m1)) m |= LINK_UPRIGHT;
            }
        } else {
            if (v_0_m1 == v) m |= LINK_UP;
            if (v == WHITE) {
                if (v_m1_m1 == v) m |= LINK_UPLEFT;
                if (v_1_m1 == v) m |= LINK_UPRIGHT;
            }
        }
    }
    return m;
}

template <typename T>
__device__ __forceinline__ uint32_t uf_find(const T *L, uint32_t i)
{
    const volatile T *V = L;
    uint32_t p = V[i];
    while (p != i) { i = p; p = V[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(uint32_t *L, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { uint32_t t = a; a = b; b = t; }   // a < b: hang b under a
        uint32_t old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;
    }
}

// pass 1: tile-local union-find in shared memory, then write global labels (index of the local root) and the local
// component sizes.  Horizontal runs are labelled without atomics (ballot of run starts inside each 32-pixel row
// segment); vertical links that are implied by the neighbouring column of the same two runs are skipped, so the
// number of smem atomics is about one per run overlap instead of one per pixel.
template <int MODE>
__global__ void __launch_bounds__(CCL_THREADS)
ccl_local_kernel(const uint8_t *__restrict__ thresh, uint32_t *__restrict__ labels, uint32_t *__restrict__ sizes, Geom g)
{
    __shared__ uint32_t L[CCL_TW * CCL_TH];
    __shared__ uint32_t Cnt[CCL_TW * CCL_TH];
    __shared__ uint8_t Ms[CCL_TW * CCL_TH];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * CCL_TW, y0 = blockIdx.y * CCL_TH;
    const uint8_t *t = thresh + (size_t)b * g.h * g.tp;
    const int lane = threadIdx.x & 31;
    constexpr int PER = CCL_TW * CCL_TH / CCL_THREADS;
    uint32_t masks[PER];
    uint8_t runlen[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * CCL_THREADS;
        const int lx = i % CCL_TW, ly = i / CCL_TW;
        const int x = x0 + lx, y = y0 + ly;
        uint32_t m = 0;
        if (x < g.w && y < g.h) m = link_mask<MODE>(t, g.tp, g.w, x, y);
        masks[k] = m;
        const uint32_t starts = __ballot_sync(0xffffffffu, !(m & LINK_LEFT) || lane == 0);
        const int start_lane = 31 - __clz(starts & ((2u << lane) - 1u));
        L[i] = (uint32_t)(i - lane + start_lane);
        // run starts remember the length of their run inside this 32-pixel segment (0 for every other pixel)
        const uint32_t higher = lane == 31 ? 0u : starts & ~((2u << lane) - 1u);
        runlen[k] = start_lane == lane ? (uint8_t)((higher ? __ffs(higher) - 1 : 32) - lane) : (uint8_t)0;
        Ms[i] = (uint8_t)m;
        Cnt[i] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * CCL_THREADS;
        const int lx = i % CCL_TW, ly = i / CCL_TW;
        const uint32_t m = masks[k];
        if (lane == 0 && (m & LINK_LEFT) && lx > 0) uf_union(L, i, i - 1);      // run continues across the 32-pixel segment
        if ((m & LINK_UP) && ly > 0) {
            const bool implied = (m & LINK_LEFT) && lx > 0 && (Ms[i - 1] & LINK_UP) && (Ms[i - CCL_TW] & LINK_LEFT);
            if (!implied) uf_union(L, i, i - CCL_TW);
        }
        if ((m & LINK_UPLEFT) && ly > 0 && lx > 0) uf_union(L, i, i - CCL_TW - 1);
        if ((m & LINK_UPRIGHT) && ly > 0 && lx < CCL_TW - 1) uf_union(L, i, i - CCL_TW + 1);
    }
    __syncthreads();
    // pointer jumping over the run starts: parallel hooking leaves chains as long as the component is tall (a start per
    // row), and a divergent walk costs the warp its longest chain; three uniform rounds cut the depth eightfold
#pragma unroll 1
    for (int round = 0; round < 3; round++) {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = threadIdx.x + k * CCL_THREADS;
            if (runlen[k]) {
                const uint32_t p = L[i];
                const uint32_t pp = L[p];
                if (pp != p) L[i] = pp;
            }
        }
        __syncthreads();
    }
    // Only run starts can be tree nodes (every other pixel still points at the start of its run), so only they walk to
    // the root; they compress their own link and credit the run's pixels to the root.  After that every pixel is two
    // loads away from its root.
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * CCL_THREADS;
        const int lx = i % CCL_TW, ly = i / CCL_TW;
        const bool in = (x0 + lx < g.w) && (y0 + ly < g.h);
        if (runlen[k] && in) {
            const uint32_t r = uf_find(L, (uint32_t)i);
            if (r != (uint32_t)i) L[i] = r;
            atomicAdd(&Cnt[r], (uint32_t)runlen[k]);
        }
    }
    __syncthreads();
    uint32_t roots[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * CCL_THREADS;
        roots[k] = L[L[i]];
    }
    const uint32_t base = (uint32_t)b * g.npix;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * CCL_THREADS;
        const int lx = i % CCL_TW, ly = i / CCL_TW;
        const int x = x0 + lx, y = y0 + ly;
        if (x < g.w && y < g.h) {
            const uint32_t r = roots[k];
            const int rx = x0 + (int)(r % CCL_TW), ry = y0 + (int)(r / CCL_TW);
            const size_t gi = (size_t)base + (size_t)y * g.w + x;
            labels[gi] = base + (uint32_t)(ry * g.w + rx);
            sizes[gi] = (r == (uint32_t)i) ? Cnt[i] : 0u;     // local component size at the local root, 0 elsewhere
        }
    }
}

// pass 2: links that cross a tile border.  One thread per pixel of a border row / column.
// mode 0: rows y = k*CCL_TH (k >= 1), all x: up / up-left / up-right links.
// mode 1: columns x = k*CCL_TW (k >= 1): left link always, up-left unless y is a tile-row border (done by mode 0);
//         columns x = k*CCL_TW - 1: up-right unless y is a tile-row border.
template <int MODE>
__global__ void ccl_merge_kernel(const uint8_t *__restrict__ thresh, uint32_t *__restrict__ labels, Geom g, int mode)
{
    const int b = blockIdx.z;
    const uint8_t *t = thresh + (size_t)b * g.h * g.tp;
    const uint32_t base = (uint32_t)b * g.npix;
    uint32_t *L = labels;   // global label space of the whole batch
    int x, y;
    if (mode == 0) {
        x = blockIdx.x * blockDim.x + threadIdx.x;
        y = (blockIdx.y + 1) * CCL_TH;
        if (x >= g.w || y >= g.h) return;
        const uint32_t m = link_mask<MODE>(t, g.tp, g.w, x, y);
        const uint32_t i = base + (uint32_t)(y * g.w + x);
        if (m & LINK_UP) uf_union(L, i, i - g.w);
        if (m & LINK_UPLEFT) uf_union(L, i, i - g.w - 1);
        if (m & LINK_UPRIGHT) uf_union(L, i, i - g.w + 1);
    } else {
        y = blockIdx.x * blockDim.x + threadIdx.x;
        const int k = blockIdx.y + 1;
        if (y >= g.h) return;
        const bool yborder = (y % CCL_TH) == 0;
        x = k * CCL_TW;
        if (x < g.w) {
            const uint32_t m = link_mask<MODE>(t, g.tp, g.w, x, y);
            const uint32_t i = base + (uint32_t)(y * g.w + x);
            if (m & LINK_LEFT) uf_union(L, i, i - 1);
            if ((m & LINK_UPLEFT) && !yborder) uf_union(L, i, i - g.w - 1);
        }
        x = k * CCL_TW - 1;
        if (x < g.w && !yborder) {
            const uint32_t m = link_mask<MODE>(t, g.tp, g.w, x, y);
            const uint32_t i = base + (uint32_t)(y * g.w + x);
            if (m & LINK_UPRIGHT) uf_union(L, i, i - g.w + 1);
        }
    }
}

// pass 3: flatten every pixel to its root; local roots that were merged into another root add their local
// component size to it (a few atomics per component instead of one per pixel).  sizes[root] ends up as the full size.
__global__ void ccl_flatten_kernel(uint32_t *__restrict__ labels, uint32_t *__restrict__ sizes, uint32_t total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t root = uf_find(labels, i);
    labels[i] = root;
    if (root != i) {
        const uint32_t c = sizes[i];
        if (c) atomicAdd(&sizes[root], c);
    }
}

// component-size gate of gradient_clusters(): pixels of components smaller than 25 become 127 ("ignore")
__global__ void ccl_mark_kernel(const uint8_t *__restrict__ thresh, const uint32_t *__restrict__ labels,
                                const uint32_t *__restrict__ sizes, uint8_t *__restrict__ mark, Geom g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= g.w) return;
    const size_t ti = (size_t)b * g.h * g.tp + (size_t)y * g.tp + x;
    uint8_t v = thresh[ti];
    if (v != 127) {
        const uint32_t root = labels[(size_t)b * g.npix + (size_t)y * g.w + x];
        if (sizes[root] < 25) v = 127;
    }
    mark[ti] = v;
}

}  // namespace cb
