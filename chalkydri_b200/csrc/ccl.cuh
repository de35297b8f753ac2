// ccl.cuh -- row A3 of SURVEY.md 8a: connected_components() as a parallel union-find.
//
// Upstream (apriltag_quad_thresh.c do_unionfind_first_line / do_unionfind_line2): for every pixel with
// 1 <= x <= w-2 and v != 127, connect to the left, up, and for white pixels up-left / up-right neighbour of
// equal value, with redundancy guards.  The guards are evaluated here exactly as upstream writes them, so the
// edge set -- and therefore the partition -- is identical; only the tree shape differs (roots are the smallest
// pixel index of each component, which is also the canonical form the oracle reports).
//
// B200 mapping: latency / atomic bound.  Pass 1 resolves 256x16-pixel tiles in shared memory (atomicMin
// union-find on 32-bit smem words), pass 2 stitches tile borders with global atomicMin, pass 3 flattens every
// pixel to its root and histograms component sizes with warp-aggregated atomics (match.any on the label).
#pragma once
#include "common.cuh"

namespace cb {

constexpr int CCL_TW = 256, CCL_TH = 16, CCL_THREADS = 256;   // 8 warps x (32-column strip), 16 rows

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

// the same from the six pixel values (v = (x,y); the others by offset); has_up = false leaves only the LEFT bit
template <int MODE>
__device__ __forceinline__ uint32_t link_mask_vals(int w, int x, bool has_up, uint32_t v, uint32_t v_m1_0, uint32_t v_m1_m1,
                                                   uint32_t v_0_m1, uint32_t v_1_m1)
{
    if (x < 1 || x > w - 2) return 0;
    const uint32_t SKIP = MODE == 0 ? 127u : 2u, WHITE = MODE == 0 ? 255u : 1u;
    if (v == SKIP) return 0;
    uint32_t m = 0;
    if (v_m1_0 == v) m |= LINK_LEFT;
    if (has_up) {
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
// component sizes.  A CTA resolves a 256x16-pixel tile.  Each warp first labels its own 32-column strip top-down, one
// row per step: the row's horizontal runs are labelled without atomics (ballot of run starts), runs are hooked to the
// runs of the previous row (vertical links implied by the neighbouring column of the same two runs are skipped), and
// the row is compressed at once -- every pixel points at its current root before the next row starts, so the walks of
// the union-find stay one or two hops long instead of growing with the height of the component.  The seven strip
// borders are stitched afterwards, then every pixel is flattened and the run starts credit their run to the root.
// root_list != nullptr (detector): instead of a dense sizes[] plane (local size at local roots, 0 elsewhere -- 4 bytes per pixel written
// here and read back by the flatten pass just to find the few non-zeros) the tile appends its local roots to a batch-wide list
// (one global atomic per CTA reserves the range) and writes sizes[] at the roots only; ccl_flatten_list_kernel walks the list.
template <int MODE>
__global__ void __launch_bounds__(CCL_THREADS)
ccl_local_kernel(const uint8_t *__restrict__ thresh, uint32_t *__restrict__ labels, uint32_t *__restrict__ sizes, Geom g,
                 uint32_t *__restrict__ root_list = nullptr, uint32_t *__restrict__ nroots = nullptr)
{
    __shared__ uint32_t L[CCL_TW * CCL_TH];
    __shared__ uint32_t Cnt[CCL_TW * CCL_TH];
    __shared__ uint32_t s_wsum[CCL_THREADS / 32 + 1];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * CCL_TW, y0 = blockIdx.y * CCL_TH;
    const uint8_t *t = thresh + (size_t)b * g.h * g.tp;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu;
    {
        const int lx = wid * 32 + lane, x = x0 + lx;
        uint32_t m_prev = 0, r_prev = 0;
        // pixel values of the previous row at x - 1, x, x + 1 stay in registers; a row costs one byte load per lane plus the
        // two halo columns (lanes 0 and 31).  Columns beyond the pitch never reach a link (x > w - 2 gives mask 0).
        uint32_t u_l = 0, u_c = 0, u_r = 0;
        const int xl = max(x - 1, 0), xr = min(x + 1, g.tp - 1), xc = min(x, g.tp - 1);
        for (int ly = 0; ly < CCL_TH; ly++) {
            const int y = y0 + ly, i = ly * CCL_TW + lx;
            uint32_t m = 0, v_c = 0, v_l = 0, v_r = 0;
            if (y < g.h) {
                const uint8_t *row = t + (size_t)y * g.tp;
                v_c = row[xc];
                v_l = __shfl_up_sync(full, v_c, 1);
                v_r = __shfl_down_sync(full, v_c, 1);
                if (lane == 0) v_l = row[xl];
                if (lane == 31) v_r = row[xr];
                if (x < g.w) m = link_mask_vals<MODE>(g.w, x, ly > 0, v_c, v_l, u_l, u_c, u_r);
            }
            u_l = v_l; u_c = v_c; u_r = v_r;
            const uint32_t starts = __ballot_sync(full, !(m & LINK_LEFT) || lane == 0);
            const int start_lane = 31 - __clz(starts & ((2u << lane) - 1u));
            const uint32_t higher = lane == 31 ? 0u : starts & ~((2u << lane) - 1u);
            const uint32_t my_start = (uint32_t)(i - lane + start_lane);
            L[i] = my_start;
            Cnt[i] = start_lane == lane ? (uint32_t)((higher ? __ffs(higher) - 1 : 32) - lane) : 0u;   // run length at run starts
            const uint32_t m_left = __shfl_up_sync(full, m, 1);
            // roots the three upper neighbours had when their row was compressed: unions start from them and from this run's
            // start instead of from the pixels, which saves a hop per walk
            const uint32_t r_ul = __shfl_up_sync(full, r_prev, 1), r_ur = __shfl_down_sync(full, r_prev, 1);
            __syncwarp();
            if (ly > 0) {
                const uint32_t NONE = 0xffffffffu;
                const bool up = (m & LINK_UP) && !((m & LINK_LEFT) && lane > 0 && (m_left & LINK_UP) && (m_prev & LINK_LEFT));
                const bool ul = (m & LINK_UPLEFT) && lane > 0, ur = (m & LINK_UPRIGHT) && lane < 31;
                // upstream's guards make UP exclude both diagonals, so two rounds cover a pixel's links (CAT: three)
                const uint32_t t1 = up ? r_prev : (ul ? r_ul : NONE);
                if (t1 != NONE) uf_union(L, my_start, t1);
                if (ur) uf_union(L, my_start, r_ur);
                if (MODE == 1 && up && ul) uf_union(L, my_start, r_ul);
                __syncwarp();
                uint32_t r = 0;
                if (start_lane == lane) r = uf_find(L, my_start);
                r = __shfl_sync(full, r, start_lane);
                L[i] = r;
                r_prev = r;
                __syncwarp();
            } else {
                r_prev = my_start;
            }
            m_prev = m;
        }
    }
    __syncthreads();
    // strip borders: columns 32k (left link, up-left link) and 32k - 1 (up-right link), k = 1..7
    if (threadIdx.x < 2 * 7 * CCL_TH) {
        const int side = threadIdx.x / (7 * CCL_TH), rem = threadIdx.x % (7 * CCL_TH);
        const int k = rem / CCL_TH + 1, ly = rem % CCL_TH;
        const int lx = side == 0 ? 32 * k : 32 * k - 1;
        const int x = x0 + lx, y = y0 + ly, i = ly * CCL_TW + lx;
        if (x < g.w && y < g.h) {
            const uint32_t m = link_mask<MODE>(t, g.tp, g.w, x, y);
            if (side == 0) {
                if (m & LINK_LEFT) uf_union(L, i, i - 1);
                if ((m & LINK_UPLEFT) && ly > 0) uf_union(L, i, i - CCL_TW - 1);
            } else {
                if ((m & LINK_UPRIGHT) && ly > 0) uf_union(L, i, i - CCL_TW + 1);
            }
        }
    }
    __syncthreads();
    constexpr int PER = CCL_TW * CCL_TH / CCL_THREADS;
    // final flatten in two steps: run starts (Cnt != 0; the only possible tree nodes) walk to the root, link straight to it
    // and hand it their run length; then every pixel is two loads from its root
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int i = threadIdx.x + k * CCL_THREADS;
        const uint32_t rl = Cnt[i];
        if (rl == 0) continue;
        const uint32_t r = uf_find(L, (uint32_t)i);
        if (r != (uint32_t)i) {
            // (a start's count is only ever added to by others when it IS the root, so rl is still its own run length)
            L[i] = r;
            const int lx = i % CCL_TW, ly = i / CCL_TW;
            if (x0 + lx < g.w && y0 + ly < g.h) atomicAdd(&Cnt[r], rl);
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
    uint32_t rootmask = 0;                                   // bit k: pixel k of this thread is a local root inside the image
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
            if (root_list == nullptr) sizes[gi] = (r == (uint32_t)i) ? Cnt[i] : 0u;     // local component size at the local root, 0 elsewhere
            else if (r == (uint32_t)i) { sizes[gi] = Cnt[i]; rootmask |= 1u << k; }
        }
    }
    if (root_list != nullptr) {
        // exclusive position of this thread's roots inside the CTA's range: warp scan, warp totals through shared memory, one atomic
        const uint32_t mine = __popc(rootmask);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(full, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < CCL_THREADS / 32; w++) { const uint32_t v = s_wsum[w]; s_wsum[w] = tot; tot += v; }
            s_wsum[CCL_THREADS / 32] = tot ? atomicAdd(nroots, tot) : 0u;
        }
        __syncthreads();
        uint32_t pos = s_wsum[CCL_THREADS / 32] + s_wsum[wid] + incl - mine;
#pragma unroll
        for (int k = 0; k < PER; k++)
            if ((rootmask >> k) & 1u) {
                const int i = threadIdx.x + k * CCL_THREADS;
                root_list[pos++] = base + (uint32_t)((y0 + i / CCL_TW) * g.w + x0 + i % CCL_TW);
            }
    }
}

// pass 3a of the detector, list form: every listed local root walks to its final root, links straight to it and adds its local size
__global__ void __launch_bounds__(256) ccl_flatten_list_kernel(uint32_t *__restrict__ labels, uint32_t *__restrict__ sizes, const uint32_t *__restrict__ root_list,
                                                               const uint32_t *__restrict__ nroots)
{
    const uint32_t n = *nroots;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = root_list[k];
        const uint32_t root = uf_find(labels, i);
        if (root != i) {
            const uint32_t c = sizes[i];
            labels[i] = root;
            atomicAdd(&sizes[root], c);
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
constexpr int FLAT_PER = 1;   // (four interleaved walks per thread measured slower: every thread then waits for its longest chain)
__global__ void __launch_bounds__(256) ccl_flatten_kernel(uint32_t *__restrict__ labels, uint32_t *__restrict__ sizes, uint32_t total)
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

// pass 3, detector variant, in two uniform steps instead of a pointer chase per pixel:
//   (a) only local roots (the pixels pass 1 left with a non-zero local size) walk to their final root, link straight to it
//       and add their local size to it;
//   (b) every pixel is then exactly two loads from its final root; the same kernel applies the component-size gate of
//       gradient_clusters() (pixels of components smaller than 25 become 127, "ignore") -- sizes[] is final because (a)
//       is a separate launch.
// Four pixels per thread: almost every thread only reads four zeros and leaves, so the launch is a quarter as many waves of
// the pointer chase's latency.
constexpr int ROOTS_PER = 4;
__global__ void __launch_bounds__(256) ccl_flatten_roots_kernel(uint32_t *__restrict__ labels, uint32_t *__restrict__ sizes, uint32_t total)
{
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * ROOTS_PER;
    if (i0 >= total) return;
    uint32_t c[ROOTS_PER];
    if (i0 + ROOTS_PER <= total && (reinterpret_cast<uintptr_t>(sizes) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sizes + i0);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < ROOTS_PER; k++) c[k] = i0 + k < total ? sizes[i0 + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < ROOTS_PER; k++) {
        if (c[k] == 0) continue;
        const uint32_t i = i0 + k;
        const uint32_t root = uf_find(labels, i);
        if (root != i) {
            labels[i] = root;
            atomicAdd(&sizes[root], c[k]);
        }
    }
}

constexpr int MARK_PER = 4;
__global__ void __launch_bounds__(256) ccl_finish_kernel(const uint8_t *__restrict__ thresh, uint32_t *__restrict__ labels,
                                                         const uint32_t *__restrict__ sizes, uint8_t *__restrict__ mark, Geom g)
{
    const int y = blockIdx.y, b = blockIdx.z;
    const size_t trow = (size_t)b * g.h * g.tp + (size_t)y * g.tp, lrow = (size_t)b * g.npix + (size_t)y * g.w;
    uint8_t v[MARK_PER];
    uint32_t root[MARK_PER];
#pragma unroll
    for (int k = 0; k < MARK_PER; k++) {
        const int x = (blockIdx.x * MARK_PER + k) * 256 + threadIdx.x;
        v[k] = 127; root[k] = 0xffffffffu;
        if (x < g.w) { v[k] = thresh[trow + x]; root[k] = labels[lrow + x]; }
    }
#pragma unroll
    for (int k = 0; k < MARK_PER; k++)
        if (root[k] != 0xffffffffu) root[k] = labels[root[k]];       // the local root's link is final after step (a)
#pragma unroll
    for (int k = 0; k < MARK_PER; k++) {
        const int x = (blockIdx.x * MARK_PER + k) * 256 + threadIdx.x;
        if (x >= g.w) continue;
        labels[lrow + x] = root[k];
        if (v[k] != 127 && sizes[root[k]] < 25) v[k] = 127;
        mark[trow + x] = v[k];
    }
}

}  // namespace cb
